/*
 * d2d_oracle.c - float64 CPU restatement of the GymD2D step path.  TEST INFRASTRUCTURE ONLY
 * (see d2d_oracle.h for the rules on who may load this and for the parity-pinning status).
 *
 * The code deliberately mirrors the reference's *evaluation order* (left-to-right Python float
 * arithmetic, math.pow / math.log10 / math.log2 from libm) so that it agrees with the running
 * reference to ~1e-14 relative; the only intentional difference is that the interference sum is
 * taken in link order instead of Python set order (simulator.py:95-101), which the reference itself
 * does not fix.
 */
#include "d2d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SPEED_OF_LIGHT 299792458.0 /* path_loss.py:9 */

/* conversion.py:4-13   pow(10, dB / 10) */
double d2d_oracle_dB_to_linear(double dB) { return pow(10.0, dB / 10.0); }

/* conversion.py:16-25  10 * log10(linear) */
double d2d_oracle_linear_to_dB(double lin) { return 10.0 * log10(lin); }

/* path_loss.py:28-39   10*ple*log10(f_GHz*1e9) + 10*ple*log10(4*pi/c) */
double d2d_oracle_pl_constant_dB(double f_GHz, double ple) {
    return 10.0 * ple * log10(f_GHz * 1e9) + 10.0 * ple * log10((4.0 * M_PI) / SPEED_OF_LIGHT);
}

/* path_loss.py:65-66    10*ple*log10(d) + pl_constant_dB */
double d2d_oracle_log_distance_pl(double dist_m, double f_GHz, double ple) {
    return 10.0 * ple * log10(dist_m) + d2d_oracle_pl_constant_dB(f_GHz, ple);
}

/* position.py:11-12     ((dx)**2 + (dy)**2) ** 0.5 */
double d2d_oracle_distance(double x1, double y1, double x2, double y2) {
    double dx = x1 - x2, dy = y1 - y2;
    return sqrt(dx * dx + dy * dy);
}

/* device.py:51-60 base; BaseStation :134-135 adds -cable +masthead; UserEquipment :158-159 adds -body */
double d2d_oracle_eirp_dBm(const d2d_oracle_device *d, double p) {
    double base = p + d->tx_antenna_gain_dBi - d->ix_margin_dB;
    if (d->is_bs) return base - d->cable_loss_dB + d->masthead_amplifier_gain_dB;
    return base - d->body_loss_dB;
}

/* device.py:62-72 base; BaseStation :137-140; UserEquipment :161-162 */
double d2d_oracle_rx_signal_level_dBm(const d2d_oracle_device *d, double eirp, double pl) {
    double base = eirp - pl + d->rx_antenna_gain_dBi;
    if (d->is_bs) return base - d->cable_loss_dB + d->masthead_amplifier_gain_dB;
    return base - d->body_loss_dB;
}

/* device.py:74-80   (noise_figure + thermal_noise) + sinr_dB */
double d2d_oracle_rx_sensitivity_dBm(const d2d_oracle_device *d) {
    return (d->noise_figure_dB + d->thermal_noise_dBm) + d->sinr_dB;
}

/* device.py:85-95   int(num_subcarriers) * int(subcarrier_spacing_kHz) */
double d2d_oracle_rb_bandwidth_kHz(const d2d_oracle_device *d) {
    return (double)((int64_t)d->num_subcarriers * (int64_t)d->subcarrier_spacing_kHz);
}

/* envs/d2d_env.py:93-101  Python floor division / modulo (n_pwr > 0), so a = -1 -> rb = -1, p = n-1 */
void d2d_oracle_decode_action(int64_t a, int64_t n_pwr, int64_t *rb, int64_t *pwr) {
    int64_t q = a / n_pwr, r = a % n_pwr;
    if (r != 0 && ((r < 0) != (n_pwr < 0))) { q -= 1; r += n_pwr; }
    *rb = q;
    *pwr = r;
}

/* path_loss.py:96-123 CostHataPathLoss.__call__ and _ms_h_correction (AreaType: 0 RURAL, 1 SUBURBAN, 2 URBAN) */
double d2d_oracle_cost_hata_pl(double dist_m, double f_GHz, int area_type, double h_tx, double h_rx) {
    const double f = f_GHz * 1000.0;            /* :100 MHz */
    const double d = dist_m / 1000.0;           /* :101 km */
    double a_hc;
    if (area_type == 2) {                       /* :116-120 */
        if (f >= 200.0) { const double l = log10(1.54 * h_rx); a_hc = 8.29 * l * l - 1.1; }
        else { const double l = log10(11.75 * h_rx); a_hc = 3.2 * l * l - 4.97; }
    } else {
        a_hc = (1.1 * log10(f) - 0.7) * h_rx - (1.56 * log10(f) - 0.8);     /* :122 */
    }
    const double c = area_type == 2 ? 3.0 : 0.0;                             /* :107 */
    return 46.3 + 33.9 * log10(f) - 13.82 * log10(h_tx) - a_hc + (44.9 - 6.55 * log10(h_tx)) * log10(d) + c;   /* :108 */
}

void d2d_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* The product's N(0,1) draw for one path-loss evaluation: Box-Muller on two 24-bit uniforms of one Philox block. */
double d2d_oracle_shadow_normal(uint64_t seed, uint64_t genv, uint32_t victim, uint32_t source, uint32_t kind, uint64_t step) {
    uint32_t ctr[4] = {(uint32_t)genv, (uint32_t)(genv >> 32) ^ (kind << 31) ^ ((uint32_t)(step >> 32) << 8), victim | (source << 16), (uint32_t)step};
    uint32_t key[2] = {(uint32_t)seed ^ 0x5bd1e995u, (uint32_t)(seed >> 32)}, o[4];
    d2d_oracle_philox4x32_10(ctr, key, o);
    const double u1 = ((double)(o[0] >> 8) + 0.5) * (1.0 / 16777216.0), u2 = ((double)(o[1] >> 8) + 0.5) * (1.0 / 16777216.0);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}

/* PathLoss.__call__(tx, rx) for the configured model (simulator.py:59,93,99).  (genv, victim, source, kind) identify the
 * evaluation for ShadowingPathLoss's per-call draw. */
static double path_loss_dB(const d2d_oracle_cfg *cfg, double K, const d2d_oracle_device *tx, const d2d_oracle_device *rx, double d,
                           uint64_t genv, uint32_t victim, uint32_t source, uint32_t kind) {
    if (cfg->path_loss_model == 2)
        return d2d_oracle_cost_hata_pl(d, cfg->carrier_freq_GHz, cfg->area_type, tx->antenna_height_m, rx->antenna_height_m);
    double pl = 10.0 * cfg->ple * log10(d) + K;      /* path_loss.py:65-66 */
    if (cfg->path_loss_model == 3 && d > cfg->shadow_d0_m)     /* path_loss.py:76-79: ldpl(d0) + 10 ple log10(d / d0) + gauss(0, chi) */
        pl += cfg->shadow_chi_dB * d2d_oracle_shadow_normal(cfg->rng_seed, genv, victim, source, kind, cfg->rng_step);
    return pl;
}

/* One environment.  Returns 0 / -1 (zero-distance link). */
static int step_one(const d2d_oracle_cfg *cfg, uint64_t genv, const d2d_oracle_device *dev, const double *pos,
                    const int32_t *act, const uint8_t *active,
                    int32_t *rb_o, int32_t *pwr_o, double *sinr_o, double *snr_o, double *rate_o,
                    double *cap_o, double *obs_o, double *reward_o,
                    int64_t *rb, int64_t *pw, int32_t *txd, int32_t *rxd, double *sinr, double *cap) {
    const int C = cfg->num_cues, D = cfg->num_due_pairs, N = C + D + (cfg->downlinks ? C : 0);
    const double K = d2d_oracle_pl_constant_dB(cfg->carrier_freq_GHz, cfg->ple);
    int status = 0;

    /* devices.py:20-25 device order; envs/d2d_env.py:55-60 link order; :80-91 link type by tx:
     * tx in due_pairs -> SIDELINK / 'due'; tx in cues -> UPLINK / 'cue'; else (the MBS) -> DOWNLINK / 'mbs' */
    for (int j = 0; j < N; ++j) {
        int64_t n_pwr;
        if (j < C) { txd[j] = 1 + j; rxd[j] = 0; n_pwr = cfg->n_pwr_cue; }
        else if (j < C + D) { txd[j] = 1 + C + 2 * (j - C); rxd[j] = txd[j] + 1; n_pwr = cfg->n_pwr_due; }
        else { txd[j] = 0; rxd[j] = 1 + (j - C - D); n_pwr = cfg->n_pwr_mbs; }
        d2d_oracle_decode_action((int64_t)act[j], n_pwr, &rb[j], &pw[j]);
    }

    int n_present = 0;
    double cap_sum = 0.0;
    for (int j = 0; j < N; ++j) {
        const int present = active ? active[j] != 0 : 1;
        double sinr_db = 0.0, snr_db = 0.0, rate = 0.0, capacity = 0.0;
        if (present) {
            const d2d_oracle_device *tx = &dev[txd[j]], *rx = &dev[rxd[j]];
            const double *pt = pos + 2 * txd[j], *pr = pos + 2 * rxd[j];
            /* simulator.py:93 */
            double d = d2d_oracle_distance(pt[0], pt[1], pr[0], pr[1]);
            if (!(d > 0.0)) status = -1;
            double pl = path_loss_dB(cfg, K, tx, rx, d, genv, (uint32_t)j, (uint32_t)j, 0);
            double rx_pwr = d2d_oracle_rx_signal_level_dBm(rx, d2d_oracle_eirp_dBm(tx, (double)pw[j]), pl);
            /* simulator.py:95-101: same-RB actions minus self; NO rx gain on interferers */
            double sum_ix = 0.0;
            for (int k = 0; k < N; ++k) {
                if (k == j || rb[k] != rb[j]) continue;
                if (active && !active[k]) continue;
                const double *pk = pos + 2 * txd[k];
                double dk = d2d_oracle_distance(pk[0], pk[1], pr[0], pr[1]);
                if (!(dk > 0.0)) status = -1;
                double ix_eirp = d2d_oracle_eirp_dBm(&dev[txd[k]], (double)pw[k]);
                double ix_pl = path_loss_dB(cfg, K, &dev[txd[k]], rx, dk, genv, (uint32_t)j, (uint32_t)k, 0);
                sum_ix += d2d_oracle_dB_to_linear(ix_eirp - ix_pl);
            }
            /* simulator.py:106-107 */
            sinr_db = rx_pwr - d2d_oracle_linear_to_dB(sum_ix + d2d_oracle_dB_to_linear(rx->thermal_noise_dBm));
            /* simulator.py:110-116: the SNR evaluates the path loss of the own link AGAIN (a second draw under shadowing) */
            if (cfg->path_loss_model == 3)
                snr_db = d2d_oracle_rx_signal_level_dBm(rx, d2d_oracle_eirp_dBm(tx, (double)pw[j]),
                                                        path_loss_dB(cfg, K, tx, rx, d, genv, (uint32_t)j, (uint32_t)j, 1)) - rx->thermal_noise_dBm;
            else
                snr_db = rx_pwr - rx->thermal_noise_dBm;
            /* simulator.py:118-127 and :144-154 (dB compared with dBm, as written) */
            if (sinr_db > d2d_oracle_rx_sensitivity_dBm(rx)) {
                rate = log2(1.0 + d2d_oracle_dB_to_linear(sinr_db));
                double b = d2d_oracle_rb_bandwidth_kHz(tx) * 1000.0;
                capacity = 1e-6 * b * log2(1.0 + d2d_oracle_dB_to_linear(sinr_db));
            }
            cap_sum += capacity;
            ++n_present;
        }
        sinr[j] = sinr_db;
        cap[j] = capacity;
        if (rb_o) rb_o[j] = present ? (int32_t)rb[j] : 0;
        if (pwr_o) pwr_o[j] = present ? (int32_t)pw[j] : 0;
        if (sinr_o) sinr_o[j] = sinr_db;
        if (snr_o) snr_o[j] = snr_db;
        if (rate_o) rate_o[j] = rate;
        if (cap_o) cap_o[j] = capacity;
        if (obs_o) { /* envs/obs_fn.py:55-61 */
            /* the reference has no row at all for an absent agent; the batched table keeps its positions and zeros */
            double *row = obs_o + 6 * j;
            row[0] = pos[2 * txd[j]]; row[1] = pos[2 * txd[j] + 1];
            row[2] = pos[2 * rxd[j]]; row[3] = pos[2 * rxd[j] + 1];
            row[4] = present ? sinr_db : 0.0; row[5] = present ? snr_db : 0.0;
        }
    }

    if (reward_o) { /* envs/reward_fn.py:27-44 */
        int bad = 0;
        for (int i = C; i < C + D && !bad; ++i) {        /* SIDELINK actions */
            if (active && !active[i]) continue;
            for (int k = 0; k < N; ++k) {                /* non-SIDELINK actions (uplink, downlink) on the same RB */
                if (k >= C && k < C + D) continue;
                if (active && !active[k]) continue;
                if (rb[k] == rb[i] && cap[k] <= cfg->min_capacity_mbps) { bad = 1; break; }
            }
        }
        /* len(actions) == 0 raises ZeroDivisionError in the reference; report NaN */
        *reward_o = bad ? -1.0 : (n_present ? cap_sum / (double)n_present : NAN);
    }
    return status;
}

int d2d_oracle_step_batch(const d2d_oracle_cfg *cfg, const d2d_oracle_device *devices,
                          int64_t E, const double *positions, const int32_t *actions,
                          const uint8_t *active, int32_t *rb, int32_t *pwr,
                          double *sinr_db, double *snr_db, double *rate, double *cap,
                          double *obs, double *reward, int nthreads) {
    const int C = cfg->num_cues, D = cfg->num_due_pairs, N = C + D + (cfg->downlinks ? C : 0), V = 1 + C + 2 * D;
    int status = 0;
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(min : status)
#endif
    {
        int64_t *s_rb = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(N ? N : 1));
        int64_t *s_pw = s_rb + N;
        int32_t *s_tx = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(N ? N : 1));
        int32_t *s_rx = s_tx + N;
        double *s_sinr = (double *)malloc(sizeof(double) * 2 * (size_t)(N ? N : 1));
        double *s_cap = s_sinr + N;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t e = 0; e < E; ++e) {
            int st = step_one(cfg, cfg->first_global_env + (uint64_t)e, devices, positions + (size_t)e * V * 2, actions + (size_t)e * N,
                              active ? active + (size_t)e * N : NULL,
                              rb ? rb + (size_t)e * N : NULL, pwr ? pwr + (size_t)e * N : NULL,
                              sinr_db ? sinr_db + (size_t)e * N : NULL, snr_db ? snr_db + (size_t)e * N : NULL,
                              rate ? rate + (size_t)e * N : NULL, cap ? cap + (size_t)e * N : NULL,
                              obs ? obs + (size_t)e * N * 6 : NULL, reward ? reward + e : NULL,
                              s_rb, s_pw, s_tx, s_rx, s_sinr, s_cap);
            if (st < status) status = st;
        }
        free(s_rb); free(s_tx); free(s_sinr);
    }
    return status;
}

/* envs/obs_fn.py:43-53 */
void d2d_oracle_per_agent_obs(const double *table, const int32_t *present, int32_t n, double *out) {
    for (int32_t i = 0; i < n; ++i) {
        double *dst = out + (size_t)i * 6 * n;
        memcpy(dst, table + 6 * present[i], 6 * sizeof(double));
        dst += 6;
        for (int32_t k = 0; k < n; ++k) {
            if (k == i) continue;
            memcpy(dst, table + 6 * present[k], 6 * sizeof(double));
            dst += 6;
        }
    }
}

/* ---- Philox4x32-10 (Salmon et al., SC'11): the product's counter-based reset sampler ---- */
static void philox4x32(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void d2d_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32(ctr, key, 10, out); }

/* uniform-in-disc draw from two 32-bit words: theta = 2*pi*u1, r = radius*sqrt(u2)  (position.py:24-28, 41-44); 24-bit
 * uniforms centred in their cell */
static void disc_from_words(uint32_t w0, uint32_t w1, double radius, double *x, double *y) {
    double u1 = ((double)(w0 >> 8) + 0.5) * (1.0 / 16777216.0), u2 = ((double)(w1 >> 8) + 0.5) * (1.0 / 16777216.0);
    double r = radius * sqrt(u2);
    *x = r * cos(2.0 * M_PI * u1);
    *y = r * sin(2.0 * M_PI * u1);
}
static void reset_block(uint64_t seed, uint64_t genv, uint32_t unit, uint32_t attempt, uint32_t o[4]) {
    uint32_t ctr[4] = {(uint32_t)genv, (uint32_t)(genv >> 32), unit, attempt};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    d2d_oracle_philox4x32_10(ctr, key, o);
}

/* The product's draw scheme (gym_d2d_b200/csrc/d2d_common.cuh): an env's devices are drawn in UNITS of one Philox block -
 * unit u < CU = ceil(C/2): CUE 2u from words (0,1), CUE 2u+1 from words (2,3); unit CU + d: DUE pair d, attempt 0 = transmitter
 * from words (0,1) and the first receiver offset from (2,3), attempt a >= 1 = offsets 2a-1 from (0,1) and 2a from (2,3). */
void d2d_oracle_reset_positions(const d2d_oracle_cfg *cfg, double cell_radius_m, double d2d_radius_m,
                                uint64_t seed, uint64_t first_global_env, int64_t E, double *positions) {
    const int C = cfg->num_cues, D = cfg->num_due_pairs, V = 1 + C + 2 * D, CU = (C + 1) / 2;
    for (int64_t e = 0; e < E; ++e) {
        double *p = positions + (size_t)e * V * 2;
        uint64_t g = first_global_env + (uint64_t)e;
        uint32_t o[4];
        p[0] = 0.0; p[1] = 0.0;                                    /* simulator.py:63-64 */
        for (int j = 0; j < C; ++j) {
            int v = 1 + j;
            reset_block(seed, g, (uint32_t)(j >> 1), 0, o);
            if (j & 1) disc_from_words(o[2], o[3], cell_radius_m, &p[2 * v], &p[2 * v + 1]);
            else disc_from_words(o[0], o[1], cell_radius_m, &p[2 * v], &p[2 * v + 1]);
        }
        for (int d = 0; d < D; ++d) {
            int t = 1 + C + 2 * d, r = t + 1;
            reset_block(seed, g, (uint32_t)(CU + d), 0, o);
            disc_from_words(o[0], o[1], cell_radius_m, &p[2 * t], &p[2 * t + 1]);
            /* position.py:38-44: redraw until inside the cell (bounded to 64 candidates here) */
            for (uint32_t k = 0; k < 64; ++k) {
                double ox, oy;
                if (k && (k & 1)) reset_block(seed, g, (uint32_t)(CU + d), (k + 1) >> 1, o);
                if (k & 1) disc_from_words(o[0], o[1], d2d_radius_m, &ox, &oy);
                else disc_from_words(o[2], o[3], d2d_radius_m, &ox, &oy);
                p[2 * r] = p[2 * t] + ox; p[2 * r + 1] = p[2 * t + 1] + oy;
                if (p[2 * r] * p[2 * r] + p[2 * r + 1] * p[2 * r + 1] <= cell_radius_m * cell_radius_m) break;
            }
        }
    }
}

/* The product's on-device Discrete(n).sample() (envs/d2d_env.py:54-60; uniform over 0 .. n-1): CUE l and DUE pair l share the
 * block Philox4x32-7 (seven rounds: the draws sit in the step kernel's hot loop) (counter = (global env, l, t >> 1), key = seed ^ 0xA511E9B3); word = 2 (t & 1) + (1 for the DUE link);
 * a = floor(word * n / 2^32).  Links beyond C + D (DOWNLINK) are absent (-1).  actions int32 [E][N]. */
void d2d_oracle_sample_actions(const d2d_oracle_cfg *cfg, int32_t num_links, uint64_t seed, uint64_t first_global_env, uint32_t t,
                               int64_t E, int32_t *actions) {
    const int C = cfg->num_cues, D = cfg->num_due_pairs, L = C > D ? C : D;
    const uint32_t n_cue = (uint32_t)(cfg->num_rbs * cfg->n_pwr_cue), n_due = (uint32_t)(cfg->num_rbs * cfg->n_pwr_due);
    for (int64_t e = 0; e < E; ++e) {
        uint64_t g = first_global_env + (uint64_t)e;
        int32_t *a = actions + (size_t)e * num_links;
        for (int j = C + D; j < num_links; ++j) a[j] = -1;
        for (int l = 0; l < L; ++l) {
            uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)l, t >> 1};
            uint32_t key[2] = {(uint32_t)seed ^ 0xA511E9B3u, (uint32_t)(seed >> 32)}, o[4];
            philox4x32(ctr, key, 7, o);
            if (l < C) a[l] = (int32_t)(((uint64_t)o[2 * (t & 1)] * n_cue) >> 32);
            if (l < D) a[C + l] = (int32_t)(((uint64_t)o[2 * (t & 1) + 1] * n_due) >> 32);
        }
    }
}
