// d2d_step_warp.cuh - fused env.step, one warp per environment, for C <= 32 CUEs and D <= 32 DUE pairs
// (the reference's default 25/25/25 configuration and everything around it).
//
// Replaces, for E environments at once, the reference call chain
//   D2DEnv.step (envs/d2d_env.py:62-71) -> _decode_action (:93-101) -> Simulator.step (simulator.py:77-154)
//   -> LinearObsFunction (envs/obs_fn.py:43-61) -> SystemCapacityRewardFunction (envs/reward_fn.py:27-44).
//
// Mapping: lane l owns CUE link l (slot A) and DUE pair l (slot B); peers are addressed by slot index
// (A: l, B: 32 + l).  The masked per-RB interference sum (Actions.get_actions_by_rb, actions.py:27-31 +
// simulator.py:95-101) is a warp-level segmented reduction:
//   * MATCH.ANY on the RB key gives every lane the mask of same-RB links inside its own slot;
//   * the two cross-slot masks go through a 64-entry shared-memory bin table per slot, tagged with the
//     warp's iteration number (all links of one RB store the same mask, so plain stores suffice);
//   * each lane walks the set bits of its peer masks.  All CUE links share one receiver (the MBS at the
//     origin), so an interferer's contribution there, u_k = w_k * g(|tx_k|), is a per-link scalar
//     computed once and a CUE victim's walk is a plain sum of u_k; a DUE victim's walk reads the peer's
//     (tx_x, tx_y, w_k) record and evaluates the gain to its own receiver.
// The sum always EXCLUDES the victim itself instead of subtracting it from a per-RB total: with SNRs of
// 70 dB the subtraction would cancel every significant bit of a weak interferer.
//
// HBM traffic per env-step is the compulsory 32N + 8V + 5 bytes: one coalesced pass over the
// env's actions and positions, one over its outputs; nothing is re-read.  There is no block-level
// prologue and no block barrier: a warp needs only its own 2 KB of shared memory.
#pragma once

#include "d2d_common.cuh"

#ifndef D2D_WARP_WARPS_PER_BLOCK
#define D2D_WARP_WARPS_PER_BLOCK 4
#endif
#ifndef D2D_WARP_MIN_BLOCKS
#define D2D_WARP_MIN_BLOCKS 8
#endif
#ifndef D2D_STATS_REPLICAS
#define D2D_STATS_REPLICAS 32
#endif

// Per warp: float4 rec[64]  [slot index] = (tx_x, tx_y, w, u):  w = 10^(p/10) tx_lin0,  u = w g(|tx|) (at the MBS)
//           uint2 bins[2][64] [slot][rb & 63] = (mask of that slot's lanes on this RB, iteration tag)

// Per-warp shared memory is addressed through one 32-bit base held in a register (explicit ld/st.shared), so the
// compiler never re-derives generic addresses from threadIdx.  Byte layout per warp:
#define D2D_W_REC 0u        // float4 rec[64]
#define D2D_W_BINS 1024u    // uint2 bins[2][64]
#define D2D_W_PWR 2048u     // float pwr_lin[128]: the warp's own copy of the integer-dBm -> mW table
#define D2D_W_BYTES 2560u
__device__ __forceinline__ void d2d_sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void d2d_sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ float4 d2d_lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint2 d2d_lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float d2d_lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}

// index of the highest set bit (FLO) and removal of that bit
__device__ __forceinline__ uint32_t d2d_pop_bit(uint32_t &mask) {
    uint32_t k;
    asm("bfind.u32 %0, %1;" : "=r"(k) : "r"(mask));
    mask ^= 1u << k;
    return k;
}

// CUE victim: every interferer is received at the MBS, so the walk sums the precomputed u_k (rec[k].w)
__device__ __forceinline__ float d2d_walk_mbs(uint32_t mask, uint32_t rec) {
    float I = 0.0f;
    while (mask) I += d2d_lds32(rec + 12u + (d2d_pop_bit(mask) << 4));
    return I;
}

// general victim: gain from each interferer's transmitter to this victim's receiver
template <bool PLE2, bool EXACT>
__device__ __forceinline__ float d2d_walk_rx(uint32_t mask, uint32_t rec, float rxx, float rxy, float nhp, float &dmin2) {
    float I = 0.0f;
    while (mask) {
        const float4 r = d2d_lds128(rec + (d2d_pop_bit(mask) << 4));
        const float dx = r.x - rxx, dy = r.y - rxy;
        const float d2 = fmaf(dx, dx, dy * dy);
        I = fmaf(r.z, d2d_gain<PLE2>(d2, nhp), I);
        if (EXACT) dmin2 = fminf(dmin2, d2);
    }
    return I;
}

// One env's inputs as a lane sees them: its CUE action + transmitter, its DUE action + (tx, rx) pair.
struct D2DLaneIn {
    int aA, aB, ns;   // ns: this env's step counter (lane 0 only)
    float2 tA;      // CUE transmitter (its receiver is the MBS at the origin)
    float4 pB;      // DUE (tx_x, tx_y, rx_x, rx_y)
};
__device__ __forceinline__ D2DLaneIn d2d_load_inputs(const D2DParams &P, uint32_t e, uint32_t lane, bool hasA, bool hasB) {
    D2DLaneIn in;
    in.aA = -1; in.aB = -1;
    in.ns = (lane == 0 && P.step_count) ? (int)P.step_count[e] : 0;
    in.tA = make_float2(0.f, 0.f);
    in.pB = make_float4(0.f, 0.f, 0.f, 0.f);
    const int32_t *act = P.actions + e * (uint32_t)P.N;
    const float2 *pe = reinterpret_cast<const float2 *>(P.pos) + e * (uint32_t)P.V;
    if (hasA) { in.aA = __ldg(act + lane); in.tA = __ldg(pe + (1u + lane)); }
    if (hasB) {
        in.aB = __ldg(act + ((uint32_t)P.C + lane));
        const float2 *q = pe + (uint32_t)(1 + P.C + 2 * lane);
        if (P.align4) {
            in.pB = __ldg(reinterpret_cast<const float4 *>(q));
        } else {
            const float2 t = __ldg(q), r = __ldg(q + 1);
            in.pB = make_float4(t.x, t.y, r.x, r.y);
        }
    }
    return in;
}

template <bool PLE2, bool EXACT>
__global__ void __launch_bounds__(D2D_WARP_WARPS_PER_BLOCK * 32, D2D_WARP_MIN_BLOCKS)
d2d_step_warp_kernel(const D2DParams P) {
    __shared__ __align__(16) unsigned char smem[D2D_WARP_WARPS_PER_BLOCK * D2D_W_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = P.N, C = P.C, V = P.V, D = N - C;
    uint32_t wb = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)warp * D2D_W_BYTES;
    asm volatile("mov.u32 %0, %0;" : "+r"(wb));      // opaque: keep the base in a register instead of re-deriving it
    const uint32_t recA = wb + D2D_W_REC, recB = recA + 512u, bins0 = wb + D2D_W_BINS, bins1 = bins0 + 512u;

    // bin tags start at 0 and `iter` at 1: shared memory left behind by an earlier block can never look current
    d2d_sts128(bins0 + (lane << 4), 0.f, 0.f, 0.f, 0.f);
    d2d_sts128(bins1 + (lane << 4), 0.f, 0.f, 0.f, 0.f);
    {   // 10^(p/10) table -> shared, so the lookup that depends on the action is an LDS, not a second global round trip
        const float4 t = __ldg(reinterpret_cast<const float4 *>(P.pwr_lin) + lane);
        d2d_sts128(wb + D2D_W_PWR + (lane << 4), t.x, t.y, t.z, t.w);
    }
    __syncwarp();

    const bool hasA = lane < C, hasB = lane < D;
    const uint32_t jA = lane, jB = C + lane;                  // canonical link indices (envs/d2d_env.py:55-60)
    // per-lane link constants stay in registers for every env this warp visits
    const float4 cA = hasA ? __ldg(reinterpret_cast<const float4 *>(P.linkA) + jA) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 cB = hasB ? __ldg(reinterpret_cast<const float4 *>(P.linkA) + jB) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float2 sA = hasA ? __ldg(reinterpret_cast<const float2 *>(P.linkB + jA)) : make_float2(0.f, 0.f);   // (sens, bw)
    const float2 sB = hasB ? __ldg(reinterpret_cast<const float2 *>(P.linkB + jB)) : make_float2(0.f, 0.f);
    const uint32_t magicA = P.magic_cue, magicB = P.magic_due;   // ceil(2^32 / n_pwr), folded on the host
    const uint32_t lane_bit = 1u << lane;

    // per-warp partial statistics (fp32 over the few envs one warp visits; flushed once as fp64 atomics)
    float st_reward = 0.f, st_cap = 0.f, st_reward2 = 0.f;
    int st_pen = 0, st_resc = 0;

    // 32-bit indexing: the host launches at most 2^31 / max(6N, 2V) envs per call (d2d_step chunks larger batches)
    uint32_t iter = 1;
    const uint32_t num_envs = (uint32_t)P.num_envs, stride = gridDim.x * D2D_WARP_WARPS_PER_BLOCK;
    uint32_t e = blockIdx.x * D2D_WARP_WARPS_PER_BLOCK + warp;
    D2DLaneIn nxt;
    if (e < num_envs) nxt = d2d_load_inputs(P, e, lane, hasA, hasB);
    for (; e < num_envs; e += stride, ++iter) {
        const uint32_t row0 = e * (uint32_t)N;
        const int32_t *act = P.actions + row0;
        const float2 *pe = reinterpret_cast<const float2 *>(P.pos) + e * (uint32_t)V;

        // ---- one coalesced pass over the env's inputs, software-pipelined: the NEXT env's loads are in flight
        // while this env computes, so a warp hides its own HBM latency ----------------------------------------
        const int aA = nxt.aA, aB = nxt.aB;
        const float2 tA = nxt.tA;
        const float4 pB = nxt.pB;
        const int ns_prev = nxt.ns;
        if (e + stride < num_envs) nxt = d2d_load_inputs(P, e + stride, lane, hasA, hasB);
        const bool actA = aA >= 0, actB = aB >= 0;

        // ---- envs/d2d_env.py:93-101: rb = a // n_pwr, p = a % n_pwr -----------------------------------------------
        const int rbA = d2d_div(aA, magicA), pA = aA - rbA * P.n_pwr_cue;
        const int rbB = d2d_div(aB, magicB), pB_ = aB - rbB * P.n_pwr_due;
        const uint32_t keyA = actA ? (uint32_t)rbA : (D2D_INACTIVE_KEY | (uint32_t)lane);
        const uint32_t keyB = actB ? (uint32_t)rbB : (D2D_INACTIVE_KEY | 32u | (uint32_t)lane);
        const float plA = actA ? d2d_lds32(wb + D2D_W_PWR + ((pA & (D2D_MAX_PWR_LEVELS - 1)) << 2)) : 0.0f;   // 10^(p/10)
        const float plB = actB ? d2d_lds32(wb + D2D_W_PWR + ((pB_ & (D2D_MAX_PWR_LEVELS - 1)) << 2)) : 0.0f;

        // ---- peer records: position, radiated weight w, and its value u at the MBS -------------------------------
        const float d2A = fmaf(tA.x, tA.x, tA.y * tA.y);                       // CUE -> MBS distance^2 (own link)
        const float lgA = d2d_lg2(d2A);
        const float gA = PLE2 ? d2d_rcp(d2A) : d2d_ex2(P.neg_half_ple * lgA);
        const float wA = plA * cA.x;
        const float d2Bm = fmaf(pB.x, pB.x, pB.y * pB.y);                      // DUE tx -> MBS distance^2 (as interferer)
        const float wB = plB * cB.x;
        if (hasA) d2d_sts128(recA + (lane << 4), tA.x, tA.y, wA, wA * gA);
        if (hasB) d2d_sts128(recB + (lane << 4), pB.x, pB.y, wB, wB * d2d_gain<PLE2>(d2Bm, P.neg_half_ple));

        // ---- same-RB masks (actions.py:27-31) ------------------------------------------------------------------
        const uint32_t mAA = __match_any_sync(0xffffffffu, keyA);
        const uint32_t mBB = __match_any_sync(0xffffffffu, keyB);
        if (actA) d2d_sts64(bins0 + ((keyA & 63u) << 3), mAA, iter);
        if (actB) d2d_sts64(bins1 + ((keyB & 63u) << 3), mBB, iter);
        __syncwarp();
        uint32_t mAB = 0, mBA = 0;                                             // DUE peers of my CUE / CUE peers of my DUE
        if (actA) { const uint2 b = d2d_lds64(bins1 + ((keyA & 63u) << 3)); mAB = b.y == iter ? b.x : 0u; }
        if (actB) { const uint2 b = d2d_lds64(bins0 + ((keyB & 63u) << 3)); mBA = b.y == iter ? b.x : 0u; }

        // ---- simulator.py:95-101: interference at each victim's receiver --------------------------------------------
        float IA = 0.0f, IB = 0.0f, dminA = 3.0e38f, dminB = 3.0e38f;
        if (EXACT) {
            if (actA) IA = d2d_walk_rx<PLE2, true>(mAA & ~lane_bit, recA, 0.f, 0.f, P.neg_half_ple, dminA) +
                           d2d_walk_rx<PLE2, true>(mAB, recB, 0.f, 0.f, P.neg_half_ple, dminA);
        } else {
            if (actA) IA = d2d_walk_mbs(mAA & ~lane_bit, recA) + d2d_walk_mbs(mAB, recB);
        }
        if (actB) IB = d2d_walk_rx<PLE2, EXACT>(mBA, recA, pB.z, pB.w, P.neg_half_ple, dminB) +
                       d2d_walk_rx<PLE2, EXACT>(mBB & ~lane_bit, recB, pB.z, pB.w, P.neg_half_ple, dminB);

        // ---- per-link epilogue (simulator.py:93,106-107,110-127,144-154) --------------------------------------------
        D2DLinkOut oA = {0.f, 0.f, 0.f, 0.f}, oB = oA;
        int need = 0;
        if (actA) {
            oA = d2d_link_epilogue<PLE2>(pA, plA, lgA, gA, IA, cA, sA, P);
            if (d2d_needs_rescue<EXACT>(oA, fminf(dminA, d2A), P)) need |= 1;
        }
        if (actB) {
            const float dx = pB.x - pB.z, dy = pB.y - pB.w;
            const float d2 = fmaf(dx, dx, dy * dy);
            const float lg = d2d_lg2(d2);
            oB = d2d_link_epilogue<PLE2>(pB_, plB, lg, PLE2 ? d2d_rcp(d2) : d2d_ex2(P.neg_half_ple * lg), IB, cB, sB, P);
            if (d2d_needs_rescue<EXACT>(oB, fminf(dminB, d2), P)) need |= 2;
        }

        // ---- envs/reward_fn.py:27-44 -----------------------------------------------------------------------------------
        const bool bad = __any_sync(0xffffffffu, actA && mAB != 0u && oA.cap <= P.min_cap);
        const int n_act = __popc(__ballot_sync(0xffffffffu, actA)) + __popc(__ballot_sync(0xffffffffu, actB));
        float cap_sum = oA.cap + oB.cap;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cap_sum += __shfl_xor_sync(0xffffffffu, cap_sum, s);
        const float reward = bad ? -1.0f : __fdividef(cap_sum, (float)n_act);

        // ---- outputs: compact observation table (envs/obs_fn.py:55-61), capacity, optional info -----------------------
        if (P.obs) {
            float2 *ob = reinterpret_cast<float2 *>(P.obs);
            if (hasA) {
                float2 *o = ob + (row0 + jA) * 3u;
                o[0] = actA ? tA : make_float2(0.f, 0.f);
                o[1] = make_float2(0.f, 0.f);
                o[2] = make_float2(oA.sinr_dB, oA.snr_dB);
            }
            if (hasB) {
                float2 *o = ob + (row0 + jB) * 3u;
                o[0] = actB ? make_float2(pB.x, pB.y) : make_float2(0.f, 0.f);
                o[1] = actB ? make_float2(pB.z, pB.w) : make_float2(0.f, 0.f);
                o[2] = make_float2(oB.sinr_dB, oB.snr_dB);
            }
        }
        if (P.cap) {
            float *c = P.cap;
            if (hasA) c[row0 + jA] = oA.cap;
            if (hasB) c[row0 + jB] = oB.cap;
        }
        if (P.rate) {
            float *c = P.rate;
            if (hasA) c[row0 + jA] = oA.rate;
            if (hasB) c[row0 + jB] = oB.rate;
        }
        if (P.rb_out) {
            int16_t *c = P.rb_out;
            if (hasA) c[row0 + jA] = actA ? (int16_t)rbA : (int16_t)0;
            if (hasB) c[row0 + jB] = actB ? (int16_t)rbB : (int16_t)0;
        }
        if (P.pwr_out) {
            int16_t *c = P.pwr_out;
            if (hasA) c[row0 + jA] = actA ? (int16_t)pA : (int16_t)0;
            if (hasB) c[row0 + jB] = actB ? (int16_t)pB_ : (int16_t)0;
        }
        if (lane == 0) {
            // envs/d2d_env.py:65,68: num_steps += 1; done = num_steps >= EPISODE_LENGTH
            const int ns = min(ns_prev + 1, 255);
            if (P.step_count) P.step_count[e] = (uint8_t)ns;
            if (P.reward) P.reward[e] = reward;
            if (P.done) P.done[e] = ns >= P.episode_length ? 1 : 0;
        }
        st_reward += reward; st_cap += cap_sum; st_reward2 = fmaf(reward, reward, st_reward2);
        st_pen += bad ? 1 : 0;

        // ---- rare: fp64 recomputation of flagged links (d2d_common.cuh).  Runs after the stores, when none of the
        // per-link state above is live; the whole warp cooperates on each flagged link. -------------------------------
        if (D2D_RESCUE_ENABLED && __any_sync(0xffffffffu, need != 0)) {
            const double2 *pe64 = P.pos64 ? reinterpret_cast<const double2 *>(P.pos64) + (int64_t)e * V : nullptr;
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                uint32_t todo = __ballot_sync(0xffffffffu, (need >> s) & 1);
                while (todo) {
                    const int L = (int)d2d_pop_bit(todo);
                    const int j = s ? C + L : L;
                    const uint32_t key = __shfl_sync(0xffffffffu, s ? keyB : keyA, L);
                    const double2 rx = d2d_pos_f64(pe, pe64, d2d_rx_dev(j, C));
                    double I = 0.0;
                    if (keyA == key && (int)jA != j) I += d2d_ix_term_f64<PLE2>((int)jA, rx, pe, pe64, act, P);
                    if (keyB == key && (int)jB != j) I += d2d_ix_term_f64<PLE2>((int)jB, rx, pe, pe64, act, P);
#pragma unroll
                    for (int sh = 16; sh > 0; sh >>= 1) I += __shfl_xor_sync(0xffffffffu, I, sh);
                    if (lane == 0) {
                        const D2DLinkOut o = d2d_link_f64<PLE2>(j, d2d_pos_f64(pe, pe64, d2d_tx_dev(j, C)), rx, I,
                                                                P.linkB[j].sens_dBm, act, P);
                        if (P.obs) *reinterpret_cast<float2 *>(P.obs + (int64_t)(row0 + j) * 6 + 4) = make_float2(o.sinr_dB, o.snr_dB);
                        if (P.cap) P.cap[row0 + j] = o.cap;
                        if (P.rate) P.rate[row0 + j] = o.rate;
                        ++st_resc;
                    }
                }
            }
        }
        __syncwarp();
    }

    if (P.stats && lane < 6) {
        // one fire-and-forget fp64 reduction per statistic and warp, spread over the replicas
        const int resc0 = __shfl_sync(0x3fu, st_resc, 0);
        const double v = lane == 0 ? (double)st_reward : lane == 1 ? (double)st_cap : lane == 2 ? (double)st_reward2
                       : lane == 3 ? (double)(iter - 1) : lane == 4 ? (double)st_pen : (double)resc0;
        const unsigned w_global = blockIdx.x * D2D_WARP_WARPS_PER_BLOCK + warp;
        if (v != 0.0) atomicAdd(P.stats + (w_global % D2D_STATS_REPLICAS) * 8 + lane, v);
    }
}
