// d2d_step_warp.cuh - fused env.step, one warp per environment, for N = C + D <= 64 links.
//
// Replaces, for E environments at once, the reference call chain
//   D2DEnv.step (envs/d2d_env.py:62-71) -> _decode_action (:93-101) -> Simulator.step (simulator.py:77-154)
//   -> LinearObsFunction (envs/obs_fn.py:43-61) -> SystemCapacityRewardFunction (envs/reward_fn.py:27-44).
//
// Mapping: lane l owns links l (slot 0) and l+32 (slot 1).  The masked per-RB interference sum
// (Actions.get_actions_by_rb, actions.py:27-31 + simulator.py:95-101) is a warp-level segmented
// reduction: MATCH.ANY on the RB key yields each lane's same-RB peer mask inside a slot; the
// cross-slot masks are exchanged through a 64-entry shared-memory bin table tagged with the
// iteration number (all writers of a bin store the same mask, so plain stores suffice); each lane then
// walks the set bits of its peer masks, reading the peer's (tx_x, tx_y, w, key) record from shared
// memory.  Every candidate is re-validated against the peer's real key, so a stale or colliding bin
// can only cost a wasted iteration, never a wrong sum.
//
// HBM traffic per env-step is the compulsory 32N + 8V + 5 bytes: one coalesced pass over
// the env's actions and positions, one over its outputs; nothing is re-read.
#pragma once

#include "d2d_common.cuh"

#define D2D_WARP_WARPS_PER_BLOCK 8
#ifndef D2D_WARP_MIN_BLOCKS
#define D2D_WARP_MIN_BLOCKS 4
#endif
#ifndef D2D_STATS_REPLICAS
#define D2D_STATS_REPLICAS 32
#endif

struct D2DWarpSmem {
    float4 linkA[D2D_WARP_MAX_LINKS];
    float4 linkB[D2D_WARP_MAX_LINKS];
    float pwr_lin[D2D_MAX_PWR_LEVELS];
    double stats[8];
    struct PerWarp {
        float4 rec[D2D_WARP_MAX_LINKS];   // (tx_x, tx_y, w, key)
        uint2 bins[2][64];                // [slot][key & 63] = (same-key lane mask within the slot, iteration tag)
    } w[D2D_WARP_WARPS_PER_BLOCK];
};

template <bool PLE2>
__device__ __forceinline__ float d2d_walk_peers(uint32_t mask, int base, uint32_t key, float rxx, float rxy,
                                                const float4 *rec, int C, float nhp, bool &sidelink_peer, float &dmin2) {
    float I = 0.0f;
    while (mask) {
        const int k = base + __ffs(mask) - 1;
        mask &= mask - 1;
        const float4 r = rec[k];
        if (__float_as_uint(r.w) == key) {
            const float dx = r.x - rxx, dy = r.y - rxy;
            const float d2 = fmaf(dx, dx, dy * dy);
            I = fmaf(r.z, d2d_gain<PLE2>(d2, nhp), I);
            dmin2 = fminf(dmin2, d2);
            sidelink_peer |= (k >= C);
        }
    }
    return I;
}

template <bool PLE2>
__global__ void __launch_bounds__(D2D_WARP_WARPS_PER_BLOCK * 32, D2D_WARP_MIN_BLOCKS)
d2d_step_warp_kernel(const D2DParams P) {
    extern __shared__ __align__(16) unsigned char d2d_smem_raw[];
    D2DWarpSmem &S = *reinterpret_cast<D2DWarpSmem *>(d2d_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = P.N, C = P.C, V = P.V;

    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        S.linkA[i] = reinterpret_cast<const float4 *>(P.linkA)[i];
        S.linkB[i] = reinterpret_cast<const float4 *>(P.linkB)[i];
    }
    for (int i = threadIdx.x; i < D2D_MAX_PWR_LEVELS; i += blockDim.x) S.pwr_lin[i] = P.pwr_lin[i];
    if (threadIdx.x < 8) S.stats[threadIdx.x] = 0.0;
    // bin tags start at 0 and `iter` at 1: shared memory left behind by an earlier block can never look current
    for (int i = threadIdx.x; i < D2D_WARP_WARPS_PER_BLOCK * 128; i += blockDim.x)
        S.w[i >> 7].bins[(i >> 6) & 1][i & 63] = make_uint2(0u, 0u);
    __syncthreads();

    D2DWarpSmem::PerWarp &W = S.w[warp];
    const int j0 = lane, j1 = lane + 32;
    const bool has0 = j0 < N, has1 = j1 < N;
    const bool cue0 = j0 < C, cue1 = j1 < C;
    const int npw0 = cue0 ? P.n_pwr_cue : P.n_pwr_due, npw1 = cue1 ? P.n_pwr_cue : P.n_pwr_due;
    const uint32_t magic0 = d2d_div_magic(npw0), magic1 = d2d_div_magic(npw1);
    const int tx0 = cue0 ? 1 + j0 : 1 + C + 2 * (j0 - C), rx0 = cue0 ? 0 : tx0 + 1;
    const int tx1 = cue1 ? 1 + j1 : 1 + C + 2 * (j1 - C), rx1 = cue1 ? 0 : tx1 + 1;
    const uint32_t lane_bit = 1u << lane;

    // per-warp partial statistics (fp32 over the few envs one warp visits; flushed to fp64 atomics)
    float st_reward = 0.f, st_cap = 0.f, st_reward2 = 0.f;
    int st_pen = 0, st_resc = 0;

    uint32_t iter = 1;
    const int64_t stride = (int64_t)gridDim.x * D2D_WARP_WARPS_PER_BLOCK;
    for (int64_t e = (int64_t)blockIdx.x * D2D_WARP_WARPS_PER_BLOCK + warp; e < P.num_envs; e += stride, ++iter) {
        const int32_t *act = P.actions + e * N;
        const float2 *pe = reinterpret_cast<const float2 *>(P.pos) + e * V;
        int a0 = -1, a1 = -1;
        float2 t0 = make_float2(0.f, 0.f), r0 = t0, t1 = t0, r1 = t0;
        if (has0) { a0 = __ldg(act + j0); t0 = __ldg(pe + tx0); r0 = __ldg(pe + rx0); }
        if (has1) { a1 = __ldg(act + j1); t1 = __ldg(pe + tx1); r1 = __ldg(pe + rx1); }
        const bool act0 = a0 >= 0, act1 = a1 >= 0;

        // envs/d2d_env.py:93-101: rb = a // n_pwr, p = a % n_pwr
        const int rb0 = d2d_div(a0, magic0), p0 = a0 - rb0 * npw0;
        const int rb1 = d2d_div(a1, magic1), p1 = a1 - rb1 * npw1;
        const uint32_t key0 = act0 ? (uint32_t)rb0 : (D2D_INACTIVE_KEY | (uint32_t)lane);
        const uint32_t key1 = act1 ? (uint32_t)rb1 : (D2D_INACTIVE_KEY | 32u | (uint32_t)lane);

        const float pl0 = act0 ? S.pwr_lin[p0 & (D2D_MAX_PWR_LEVELS - 1)] : 0.0f;
        const float pl1 = act1 ? S.pwr_lin[p1 & (D2D_MAX_PWR_LEVELS - 1)] : 0.0f;
        const float4 A0v = S.linkA[j0], A1v = S.linkA[j1];
        const D2DLinkA A0 = {A0v.x, A0v.y, A0v.z, A0v.w}, A1 = {A1v.x, A1v.y, A1v.z, A1v.w};

        // peer records + same-RB masks
        if (has0) W.rec[j0] = make_float4(t0.x, t0.y, pl0 * A0.tx_lin0, __uint_as_float(key0));
        if (has1) W.rec[j1] = make_float4(t1.x, t1.y, pl1 * A1.tx_lin0, __uint_as_float(key1));
        const uint32_t m00 = __match_any_sync(0xffffffffu, key0);
        const uint32_t m11 = __match_any_sync(0xffffffffu, key1);
        if (act0) W.bins[0][key0 & 63] = make_uint2(m00, iter);
        if (act1) W.bins[1][key1 & 63] = make_uint2(m11, iter);
        __syncwarp();
        uint32_t m01 = 0, m10 = 0;
        if (act0) { const uint2 b = W.bins[1][key0 & 63]; m01 = b.y == iter ? b.x : 0u; }
        if (act1) { const uint2 b = W.bins[0][key1 & 63]; m10 = b.y == iter ? b.x : 0u; }

        // simulator.py:95-101 interference at each victim's receiver
        bool side0 = false, side1 = false;
        float I0 = 0.0f, I1 = 0.0f, dmin0 = 3.0e38f, dmin1 = 3.0e38f;
        if (act0) {
            I0 = d2d_walk_peers<PLE2>(m00 & ~lane_bit, 0, key0, r0.x, r0.y, W.rec, C, P.neg_half_ple, side0, dmin0);
            I0 += d2d_walk_peers<PLE2>(m01, 32, key0, r0.x, r0.y, W.rec, C, P.neg_half_ple, side0, dmin0);
        }
        if (act1) {
            I1 = d2d_walk_peers<PLE2>(m10, 0, key1, r1.x, r1.y, W.rec, C, P.neg_half_ple, side1, dmin1);
            I1 += d2d_walk_peers<PLE2>(m11 & ~lane_bit, 32, key1, r1.x, r1.y, W.rec, C, P.neg_half_ple, side1, dmin1);
        }

        // per-link epilogue (simulator.py:93,106-107,110-127,144-154)
        D2DLinkOut o0 = {0.f, 0.f, 0.f, 0.f}, o1 = o0;
        int need = 0;
        if (act0) {
            const float4 Bv = S.linkB[j0];
            const D2DLinkB B0 = {Bv.x, Bv.y, 0, 0};
            const float dx = t0.x - r0.x, dy = t0.y - r0.y;
            const float d2 = fmaf(dx, dx, dy * dy);
            o0 = d2d_link_epilogue<PLE2>(p0, pl0, d2, I0, A0, B0, P);
            if (d2d_needs_rescue(o0, fminf(dmin0, d2), P)) need |= 1;
        }
        if (act1) {
            const float4 Bv = S.linkB[j1];
            const D2DLinkB B1 = {Bv.x, Bv.y, 0, 0};
            const float dx = t1.x - r1.x, dy = t1.y - r1.y;
            const float d2 = fmaf(dx, dx, dy * dy);
            o1 = d2d_link_epilogue<PLE2>(p1, pl1, d2, I1, A1, B1, P);
            if (d2d_needs_rescue(o1, fminf(dmin1, d2), P)) need |= 2;
        }

        // envs/reward_fn.py:27-44
        const bool bad0 = act0 && cue0 && side0 && o0.cap <= P.min_cap;
        const bool bad1 = act1 && cue1 && side1 && o1.cap <= P.min_cap;
        const bool bad = __any_sync(0xffffffffu, bad0 || bad1);
        const int n_act = __popc(__ballot_sync(0xffffffffu, act0)) + __popc(__ballot_sync(0xffffffffu, act1));
        float cap_sum = o0.cap + o1.cap;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cap_sum += __shfl_xor_sync(0xffffffffu, cap_sum, s);
        const float reward = bad ? -1.0f : cap_sum / (float)n_act;

        // outputs: compact observation table (envs/obs_fn.py:55-61) + capacity + optional info
        if (P.obs) {
            if (has0) {
                float2 *o = reinterpret_cast<float2 *>(P.obs + (e * N + j0) * 6);
                o[0] = act0 ? t0 : make_float2(0.f, 0.f);
                o[1] = act0 ? r0 : make_float2(0.f, 0.f);
                o[2] = make_float2(o0.sinr_dB, o0.snr_dB);
            }
            if (has1) {
                float2 *o = reinterpret_cast<float2 *>(P.obs + (e * N + j1) * 6);
                o[0] = act1 ? t1 : make_float2(0.f, 0.f);
                o[1] = act1 ? r1 : make_float2(0.f, 0.f);
                o[2] = make_float2(o1.sinr_dB, o1.snr_dB);
            }
        }
        if (P.cap) {
            if (has0) P.cap[e * N + j0] = o0.cap;
            if (has1) P.cap[e * N + j1] = o1.cap;
        }
        if (P.rate) {
            if (has0) P.rate[e * N + j0] = o0.rate;
            if (has1) P.rate[e * N + j1] = o1.rate;
        }
        if (P.rb_out) {
            if (has0) P.rb_out[e * N + j0] = act0 ? (int16_t)rb0 : (int16_t)0;
            if (has1) P.rb_out[e * N + j1] = act1 ? (int16_t)rb1 : (int16_t)0;
        }
        if (P.pwr_out) {
            if (has0) P.pwr_out[e * N + j0] = act0 ? (int16_t)p0 : (int16_t)0;
            if (has1) P.pwr_out[e * N + j1] = act1 ? (int16_t)p1 : (int16_t)0;
        }
        if (lane == 0) {
            // envs/d2d_env.py:65,68: num_steps += 1; done = num_steps >= EPISODE_LENGTH
            int ns = P.step_count ? (int)P.step_count[e] + 1 : 1;
            if (ns > 255) ns = 255;
            if (P.step_count) P.step_count[e] = (uint8_t)ns;
            if (P.reward) P.reward[e] = reward;
            if (P.done) P.done[e] = ns >= P.episode_length ? 1 : 0;
        }
        st_reward += reward; st_cap += cap_sum; st_reward2 = fmaf(reward, reward, st_reward2);
        st_pen += bad ? 1 : 0;

        // rare: fp64 recomputation of flagged links (d2d_common.cuh).  Runs after the stores so none of the
        // per-link state above is live; the whole warp cooperates on each flagged link.
        if (D2D_RESCUE_ENABLED && __any_sync(0xffffffffu, need != 0)) {
            const double2 *pe64 = P.pos64 ? reinterpret_cast<const double2 *>(P.pos64) + e * V : nullptr;
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                uint32_t todo = __ballot_sync(0xffffffffu, (need >> s) & 1);
                while (todo) {
                    const int j = 32 * s + __ffs(todo) - 1;
                    todo &= todo - 1;
                    const uint32_t key = __float_as_uint(W.rec[j].w);
                    const double2 rx = d2d_pos_f64(pe, pe64, d2d_rx_dev(j, C));
                    double I = 0.0;
#pragma unroll 1
                    for (int k = lane; k < N; k += 32)
                        if (k != j && __float_as_uint(W.rec[k].w) == key) I += d2d_ix_term_f64<PLE2>(k, rx, pe, pe64, act, P);
#pragma unroll
                    for (int sh = 16; sh > 0; sh >>= 1) I += __shfl_xor_sync(0xffffffffu, I, sh);
                    if (lane == 0) {
                        const D2DLinkOut o = d2d_link_f64<PLE2>(j, d2d_pos_f64(pe, pe64, d2d_tx_dev(j, C)), rx, I,
                                                                S.linkB[j].x, act, P);
                        const int64_t g = e * N + j;
                        if (P.obs) *reinterpret_cast<float2 *>(P.obs + g * 6 + 4) = make_float2(o.sinr_dB, o.snr_dB);
                        if (P.cap) P.cap[g] = o.cap;
                        if (P.rate) P.rate[g] = o.rate;
                        ++st_resc;
                    }
                }
            }
        }
        __syncwarp();
    }

    if (P.stats) {
        const int resc_w = st_resc;
        if (lane == 0) {
            atomicAdd(&S.stats[0], (double)st_reward); atomicAdd(&S.stats[1], (double)st_cap);
            atomicAdd(&S.stats[2], (double)st_reward2); atomicAdd(&S.stats[3], (double)(iter - 1));
            atomicAdd(&S.stats[4], (double)st_pen); atomicAdd(&S.stats[5], (double)resc_w);
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            const double v = S.stats[threadIdx.x];
            if (v != 0.0) atomicAdd(P.stats + (blockIdx.x % D2D_STATS_REPLICAS) * 8 + threadIdx.x, v);
        }
    }
}
