// d2d_step_block.cuh - fused env.step, one thread block per environment, for N > 64 links
// (BASELINE config #3: 100 RBs / 100 CUEs / 500 DUE pairs -> N = 600, V = 1101).
//
// Same arithmetic contract as d2d_step_warp.cuh.  The same-RB grouping of Actions.get_actions_by_rb
// (actions.py:27-31) is a shared-memory counting sort by RB: count -> exclusive scan -> scatter; each
// victim link then walks its own bin (expected N/R peers) instead of all N links, so the work per
// env-step is O(N + N^2/R) pair evaluations, like the reference, not the dense N*Q product.
#pragma once

#include "d2d_common.cuh"

#define D2D_BLOCK_THREADS 256

// dynamic shared memory layout (sizes depend on N and nbins), see d2d_block_smem_bytes()
struct D2DBlockSmemView {
    float4 *rec;       // [N] (tx_x, tx_y, w, key)
    float4 *aux;       // [N] (rx_x, rx_y, p_lin, p as int bits)
    float4 *linkA;     // [N]
    float4 *linkB;     // [N]
    float *pwr_lin;    // [D2D_MAX_PWR_LEVELS]
    int32_t *bin_cnt;  // [nbins]
    int32_t *bin_off;  // [nbins]
    int32_t *slot;     // [N] position of the link inside its bin
    uint16_t *sorted;  // [N] link indices grouped by bin
    float *red;        // [32] reduction scratch
};

__host__ __device__ inline size_t d2d_block_smem_bytes(int N, int nbins) {
    size_t b = 0;
    b += (size_t)4 * N * sizeof(float4);
    b += D2D_MAX_PWR_LEVELS * sizeof(float);
    b += (size_t)2 * nbins * sizeof(int32_t);
    b += (size_t)N * sizeof(int32_t);
    b += 32 * sizeof(float);
    b += ((size_t)N * sizeof(uint16_t) + 15) & ~(size_t)15;
    return b;
}

__device__ __forceinline__ D2DBlockSmemView d2d_block_carve(unsigned char *raw, int N, int nbins) {
    D2DBlockSmemView v;
    v.rec = reinterpret_cast<float4 *>(raw);
    v.aux = v.rec + N;
    v.linkA = v.aux + N;
    v.linkB = v.linkA + N;
    v.pwr_lin = reinterpret_cast<float *>(v.linkB + N);
    v.bin_cnt = reinterpret_cast<int32_t *>(v.pwr_lin + D2D_MAX_PWR_LEVELS);
    v.bin_off = v.bin_cnt + nbins;
    v.slot = v.bin_off + nbins;
    v.red = reinterpret_cast<float *>(v.slot + N);
    v.sorted = reinterpret_cast<uint16_t *>(v.red + 32);
    return v;
}

template <bool PLE2>
__global__ void __launch_bounds__(D2D_BLOCK_THREADS) d2d_step_block_kernel(const D2DParams P) {
    extern __shared__ __align__(16) unsigned char d2d_smem_raw[];
    const int N = P.N, C = P.C, V = P.V, nbins = P.nbins;
    D2DBlockSmemView S = d2d_block_carve(d2d_smem_raw, N, nbins);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    d2d_pdl_launch_dependents();

    for (int i = tid; i < N; i += D2D_BLOCK_THREADS) {
        S.linkA[i] = reinterpret_cast<const float4 *>(P.linkA)[i];
        S.linkB[i] = reinterpret_cast<const float4 *>(P.linkB)[i];
    }
    for (int i = tid; i < D2D_MAX_PWR_LEVELS; i += D2D_BLOCK_THREADS) S.pwr_lin[i] = P.pwr_lin[i];
    const uint32_t magic_cue = d2d_div_magic(P.n_pwr_cue), magic_due = d2d_div_magic(P.n_pwr_due);

    double st_reward = 0.0, st_cap = 0.0, st_reward2 = 0.0, st_pen = 0.0, st_resc = 0.0, st_n = 0.0;
    d2d_pdl_wait();

    for (int64_t e = blockIdx.x; e < P.num_envs; e += gridDim.x) {
        const int32_t *act = P.actions + e * N;
        const float2 *pe = reinterpret_cast<const float2 *>(P.pos) + e * V;
        for (int i = tid; i < nbins; i += D2D_BLOCK_THREADS) S.bin_cnt[i] = 0;
        __syncthreads();

        // phase 1: decode (envs/d2d_env.py:93-101), stage link records, count links per RB bin
        for (int j = tid; j < N; j += D2D_BLOCK_THREADS) {
            const bool cue = j < C;
            const int txd = cue ? 1 + j : 1 + C + 2 * (j - C), rxd = cue ? 0 : txd + 1;
            const int a = __ldg(act + j);
            const float2 t = __ldg(pe + txd), r = __ldg(pe + rxd);
            const bool active = a >= 0;
            const int npw = cue ? P.n_pwr_cue : P.n_pwr_due;
            const int rb = d2d_div(a, cue ? magic_cue : magic_due), p = a - rb * npw;
            const uint32_t key = active ? (uint32_t)rb : (D2D_INACTIVE_KEY | (uint32_t)j);
            const float pl = active ? S.pwr_lin[p & (D2D_MAX_PWR_LEVELS - 1)] : 0.0f;
            S.rec[j] = make_float4(t.x, t.y, pl * S.linkA[j].x, __uint_as_float(key));
            S.aux[j] = make_float4(r.x, r.y, pl, __int_as_float(active ? p : -1));
            if (active) S.slot[j] = atomicAdd(&S.bin_cnt[key % (uint32_t)nbins], 1);
        }
        __syncthreads();

        // exclusive scan of the bin counts (first warp; nbins is small)
        if (warp == 0) {
            const int chunk = (nbins + 31) >> 5, lo = lane * chunk, hi = min(lo + chunk, nbins);
            int sum = 0;
            for (int i = lo; i < hi; ++i) sum += S.bin_cnt[i];
            int incl = sum;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= s) incl += n;
            }
            int run = incl - sum;
            for (int i = lo; i < hi; ++i) { S.bin_off[i] = run; run += S.bin_cnt[i]; }
        }
        __syncthreads();
        for (int j = tid; j < N; j += D2D_BLOCK_THREADS) {
            const uint32_t key = __float_as_uint(S.rec[j].w);
            if (!(key & D2D_INACTIVE_KEY)) S.sorted[S.bin_off[key % (uint32_t)nbins] + S.slot[j]] = (uint16_t)j;
        }
        __syncthreads();

        // phase 2: per-victim interference walk + epilogue + outputs
        float cap_part = 0.0f;
        int n_act = 0, bad = 0, resc = 0;
        for (int j = tid; j < N; j += D2D_BLOCK_THREADS) {
            const float4 rj = S.rec[j], xj = S.aux[j];
            const uint32_t key = __float_as_uint(rj.w);
            const int p = __float_as_int(xj.w);
            const bool active = p >= 0;
            D2DLinkOut o = {0.f, 0.f, 0.f, 0.f};
            if (active) {
                const int b = key % (uint32_t)nbins, beg = S.bin_off[b], end = beg + S.bin_cnt[b];
                float I = 0.0f, dmin2 = 3.0e38f;
                bool side = false;
                for (int q = beg; q < end; ++q) {
                    const int k = S.sorted[q];
                    const float4 rk = S.rec[k];
                    if (k != j && __float_as_uint(rk.w) == key) {
                        const float dx = rk.x - xj.x, dy = rk.y - xj.y;
                        const float d2 = fmaf(dx, dx, dy * dy);
                        I = fmaf(rk.z, d2d_gain<PLE2>(d2, P.neg_half_ple), I);
                        dmin2 = fminf(dmin2, d2);
                        side |= (k >= C);
                    }
                }
                const float4 Av = S.linkA[j], Bv = S.linkB[j];
                const float2 sb = make_float2(Bv.x, Bv.y);
                const float dx = rj.x - xj.x, dy = rj.y - xj.y;
                const float d2own = fmaf(dx, dx, dy * dy);
                const float lg = d2d_lg2(d2own);
                o = d2d_link_epilogue<PLE2>(p, xj.z, lg, PLE2 ? d2d_rcp(d2own) : d2d_ex2(P.neg_half_ple * lg), I, Av, sb, P);
                if (D2D_RESCUE_ENABLED && d2d_needs_rescue<true>(o, fminf(dmin2, d2own), P)) {   // rare: fp64 pass (d2d_common.cuh)
                    const double2 *pe64 = P.pos64 ? reinterpret_cast<const double2 *>(P.pos64) + e * V : nullptr;
                    const double2 rx = d2d_pos_f64(pe, pe64, d2d_rx_dev(j, C));
                    double I64 = 0.0;
                    for (int q = beg; q < end; ++q) {
                        const int k = S.sorted[q];
                        if (k != j && __float_as_uint(S.rec[k].w) == key) I64 += d2d_ix_term_f64<PLE2>(k, rx, pe, pe64, act, P);
                    }
                    o = d2d_link_f64<PLE2>(j, d2d_pos_f64(pe, pe64, d2d_tx_dev(j, C)), rx, I64, sb.x, act, P);
                    ++resc;
                }
                cap_part += o.cap;
                ++n_act;
                bad |= (j < C && side && o.cap <= P.min_cap);   // envs/reward_fn.py:30-39
            }
            const int64_t g = e * N + j;
            if (P.obs) {
                float2 *ob = reinterpret_cast<float2 *>(P.obs + g * 6);
                ob[0] = make_float2(rj.x, rj.y);      // an absent agent's row keeps its positions (sinr = snr = 0)
                ob[1] = make_float2(xj.x, xj.y);
                ob[2] = make_float2(o.sinr_dB, o.snr_dB);
            }
            if (P.cap) P.cap[g] = o.cap;
            if (P.rate) P.rate[g] = o.rate;
            if (P.rb_out) P.rb_out[g] = active ? (int16_t)key : (int16_t)0;
            if (P.pwr_out) P.pwr_out[g] = active ? (int16_t)p : (int16_t)0;
        }

        // block reduction for the reward (envs/reward_fn.py:27-44)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cap_part += __shfl_xor_sync(0xffffffffu, cap_part, s);
        const int n_act_w = __reduce_add_sync(0xffffffffu, n_act);
        const int resc_w = __reduce_add_sync(0xffffffffu, resc);
        if (lane == 0) { S.red[warp] = cap_part; S.red[8 + warp] = (float)n_act_w; S.red[16 + warp] = (float)resc_w; }
        const int any_bad = __syncthreads_or(bad);   // barrier: S.red is visible to thread 0 below
        if (tid == 0) {
            float cs = 0.f, na = 0.f, rs = 0.f;
            for (int w2 = 0; w2 < D2D_BLOCK_THREADS / 32; ++w2) { cs += S.red[w2]; na += S.red[8 + w2]; rs += S.red[16 + w2]; }
            const float reward = any_bad ? -1.0f : cs / na;
            int ns = P.step_count ? (int)P.step_count[e] + 1 : 1;
            if (ns > 255) ns = 255;
            if (P.step_count) P.step_count[e] = (uint8_t)ns;
            if (P.reward) P.reward[e] = reward;
            if (P.done) P.done[e] = ns >= P.episode_length ? 1 : 0;
            st_reward += reward; st_cap += cs; st_reward2 += (double)reward * reward;
            st_pen += any_bad ? 1.0 : 0.0; st_resc += rs; st_n += 1.0;
        }
        __syncthreads();
    }

    if (P.stats && tid == 0) {
        double *dst = P.stats + (blockIdx.x % 32) * 8;
        atomicAdd(dst + 0, st_reward); atomicAdd(dst + 1, st_cap); atomicAdd(dst + 2, st_reward2);
        atomicAdd(dst + 3, st_n); atomicAdd(dst + 4, st_pen); atomicAdd(dst + 5, st_resc);
    }
}
