// d2d_step_warp.cuh - fused env.step, one warp per environment, for C <= 32 CUEs, D <= 32 DUE pairs and
// R <= 64 RBs (the reference's default 25/25/25 configuration and everything around it).
//
// Replaces, for E environments at once, the reference call chain
//   D2DEnv.step (envs/d2d_env.py:62-71) -> _decode_action (:93-101) -> Simulator.step (simulator.py:77-154)
//   -> LinearObsFunction (envs/obs_fn.py:43-61) -> SystemCapacityRewardFunction (envs/reward_fn.py:27-44).
//
// Mapping: lane l owns CUE link l (slot A) and DUE pair l (slot B).
//
// The masked per-RB interference sum (Actions.get_actions_by_rb, actions.py:27-31 + simulator.py:95-101) is a
// segmented reduction over links grouped by RB.  The grouping is a per-warp counting sort in shared memory:
//   rank   = atomicAdd(count[rb], 1)            one shared atomic per link (the DUE count rides in the high half)
//   offset = exclusive warp-shuffle scan of the 64 counts (two bins per lane, packed in one register)
//   sorted[offset[rb] + rank] = (tx_x, tx_y, w, u)   the link's peer record
// after which a victim's co-channel peers are ONE contiguous range of records.  Its first five entries are
// handled branch-free (plain indexed LDS, predicated terms, so the loads and gains of a lane's two victims
// overlap); an RB with more than five links (1.4 % of RBs at the default load) falls into a short serial loop.
// Measured alternatives this replaced: MATCH.ANY costs ~290 cycles per call on sm_100a, and a bit-serial walk over
// peer masks ~100 cycles per trip (FLO -> shift -> branch is one dependent chain).
//
// All CUE links share one receiver (the MBS at the origin), so there an interferer contributes a per-link scalar
// u_k = w_k g(|tx_k|) computed once, and a CUE victim's walk is a sum of u_k; a DUE victim evaluates
// w_k g(|tx_k - rx|).  The victim is excluded from its own range by index - never subtracted from a per-RB total,
// which would cancel a weak interferer against a 70 dB stronger self term.
//
// HBM traffic per env-step is the compulsory 32N + 8V + 5 bytes: one coalesced pass over the env's
// actions and positions (software-pipelined one env ahead), one over its outputs; nothing is re-read.  There is no
// block-level prologue and no block barrier: a warp needs only its own 2.4 KB of shared memory.
#pragma once

#include "d2d_common.cuh"

// Two launch shapes, picked by the host from the batch size (measured on B200, profiles/README.md):
//   WPB = 4, >= 7 blocks/SM (72 registers): finest block granularity - best when the batch is about one wave (E = 4096)
//   WPB = 8, >= 4 blocks/SM (64 registers): best sustained throughput for batches of many waves
#define D2D_WARP_MIN_BLOCKS(WPB) ((WPB) == 4 ? 7 : 4)
#ifndef D2D_STATS_REPLICAS
#define D2D_STATS_REPLICAS 32
#endif
#define D2D_WALK_INLINE 5      // range entries handled branch-free before the serial loop

// Per-warp shared memory, addressed through one 32-bit base held in a register (explicit ld/st.shared, so the
// compiler never re-derives generic addresses from threadIdx).  Byte layout:
#define D2D_W_SORTED 0u        // float4 sorted[64 + 8]: peer records grouped by RB (+ slack for the inline over-read)
#define D2D_W_CNT 1152u        // u32 count[2][64]: per-RB link count, double-buffered by iteration parity
#define D2D_W_OFF 1664u        // u32 offset[64]: exclusive scan of count
#define D2D_W_PWR 1920u        // float pwr_lin[128]: the warp's copy of the integer-dBm -> mW table
#define D2D_W_BYTES 2432u

__device__ __forceinline__ void d2d_sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void d2d_sts32(uint32_t a, uint32_t x) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}
__device__ __forceinline__ float4 d2d_lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float d2d_lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t d2d_lds32u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t d2d_atoms_add(uint32_t a, uint32_t x) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(x) : "memory");
    return old;
}
// index of the highest set bit (FLO) and removal of that bit - rescue path only
__device__ __forceinline__ uint32_t d2d_pop_bit(uint32_t &mask) {
    uint32_t k;
    asm("bfind.u32 %0, %1;" : "=r"(k) : "r"(mask));
    mask ^= 1u << k;
    return k;
}

// One interferer's term at a general receiver: w_k g(|tx_k - rx|)
template <bool PLE2, bool EXACT>
__device__ __forceinline__ float d2d_term_rx(uint32_t addr, bool valid, float rxx, float rxy, float nhp, float &dmin2) {
    const float4 r = d2d_lds128(addr);
    const float dx = r.x - rxx, dy = r.y - rxy;
    const float d2 = fmaf(dx, dx, dy * dy);
    if (EXACT) dmin2 = valid ? fminf(dmin2, d2) : dmin2;
    return valid ? r.z * d2d_gain<PLE2>(d2, nhp) : 0.0f;
}
// Interference at a general receiver from the records [beg, beg + n) minus the victim's own record `self`
template <bool PLE2, bool EXACT>
__device__ __forceinline__ float d2d_walk_rx(uint32_t sorted, uint32_t beg, uint32_t n, uint32_t self, float rxx, float rxy,
                                             float nhp, float &dmin2) {
    float I = 0.0f;
#pragma unroll
    for (uint32_t t = 0; t < D2D_WALK_INLINE; ++t)
        I += d2d_term_rx<PLE2, EXACT>(sorted + ((beg + t) << 4), t < n && beg + t != self, rxx, rxy, nhp, dmin2);
    for (uint32_t t = D2D_WALK_INLINE; t < n; ++t)       // rare: more than D2D_WALK_INLINE links on one RB
        I += d2d_term_rx<PLE2, EXACT>(sorted + ((beg + t) << 4), beg + t != self, rxx, rxy, nhp, dmin2);
    return I;
}
// Interference at the MBS: the records carry u_k = w_k g(|tx_k|) in .w
__device__ __forceinline__ float d2d_walk_mbs(uint32_t sorted, uint32_t beg, uint32_t n, uint32_t self) {
    float I = 0.0f;
#pragma unroll
    for (uint32_t t = 0; t < D2D_WALK_INLINE; ++t) {
        const float u = d2d_lds32(sorted + 12u + ((beg + t) << 4));
        I += (t < n && beg + t != self) ? u : 0.0f;
    }
    for (uint32_t t = D2D_WALK_INLINE; t < n; ++t) {
        const float u = d2d_lds32(sorted + 12u + ((beg + t) << 4));
        I += (beg + t != self) ? u : 0.0f;
    }
    return I;
}

// One env's inputs as a lane sees them: its CUE action + transmitter, its DUE action + (tx, rx) pair.
struct D2DLaneIn {
    int aA, aB, ns;   // ns: this env's step counter (lane 0 only)
    float2 tA;        // CUE transmitter (its receiver is the MBS at the origin)
    float4 pB;        // DUE (tx_x, tx_y, rx_x, rx_y)
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

#ifdef D2D_TIMELINE   // debug builds only: SM-clock timestamps of one warp's phases (profiles/timeline.py)
__device__ unsigned long long d2d_dbg[16];
#define D2D_TICK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) d2d_dbg[i] = clock64(); } while (0)
#else
#define D2D_TICK(i) do { } while (0)
#endif

template <bool PLE2, bool EXACT, int WPB>
__global__ void __launch_bounds__(WPB * 32, D2D_WARP_MIN_BLOCKS(WPB))
d2d_step_warp_kernel(const D2DParams P) {
    constexpr int D2D_WARP_WARPS_PER_BLOCK = WPB;
    __shared__ __align__(16) unsigned char smem[WPB * D2D_W_BYTES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = P.N, C = P.C, V = P.V, D = N - C;
    D2D_TICK(0);
    d2d_pdl_launch_dependents();
    uint32_t wb = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)warp * D2D_W_BYTES;
    asm volatile("mov.u32 %0, %0;" : "+r"(wb));      // opaque: keep the base in a register instead of re-deriving it
    const uint32_t sorted = wb + D2D_W_SORTED, offs = wb + D2D_W_OFF;

    // both parity buffers of the RB counters start at zero; each iteration re-zeroes the one it is not using
    d2d_sts128(wb + D2D_W_CNT + (lane << 4), 0.f, 0.f, 0.f, 0.f);
    // 10^(p/10) table -> shared, so the lookup that depends on the action is an LDS, not a second global round trip;
    // the load is issued here and parked in shared memory only after the first env's inputs are in flight
    const float4 lut = __ldg(reinterpret_cast<const float4 *>(P.pwr_lin) + lane);

    const bool hasA = lane < C, hasB = lane < D;
    const uint32_t jA = lane, jB = C + lane;                  // canonical link indices (envs/d2d_env.py:55-60)
    // per-lane link constants stay in registers for every env this warp visits
    const float4 cA = hasA ? __ldg(reinterpret_cast<const float4 *>(P.linkA) + jA) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 cB = hasB ? __ldg(reinterpret_cast<const float4 *>(P.linkA) + jB) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float2 sA = hasA ? __ldg(reinterpret_cast<const float2 *>(P.linkB + jA)) : make_float2(0.f, 0.f);   // (sens, bw)
    const float2 sB = hasB ? __ldg(reinterpret_cast<const float2 *>(P.linkB + jB)) : make_float2(0.f, 0.f);
    const uint32_t magicA = P.magic_cue, magicB = P.magic_due;   // ceil(2^32 / n_pwr), folded on the host

    // per-warp partial statistics (fp32 over the few envs one warp visits; flushed once as fp64 atomics)
    float st_reward = 0.f, st_cap = 0.f, st_reward2 = 0.f;
    int st_pen = 0, st_resc = 0;

    // 32-bit indexing: the host launches at most 2^31 / max(6N, 2V) envs per call (d2d_step chunks larger batches)
    uint32_t iter = 1;
    const uint32_t num_envs = (uint32_t)P.num_envs, stride = gridDim.x * D2D_WARP_WARPS_PER_BLOCK;
    uint32_t e = blockIdx.x * D2D_WARP_WARPS_PER_BLOCK + warp;
    D2DLaneIn nxt;
    D2D_TICK(1);
    d2d_pdl_wait();
    D2D_TICK(2);
    if (e < num_envs) nxt = d2d_load_inputs(P, e, lane, hasA, hasB);
    d2d_sts128(wb + D2D_W_PWR + (lane << 4), lut.x, lut.y, lut.z, lut.w);
    __syncwarp();
    for (; e < num_envs; e += stride, ++iter) {
        const uint32_t row0 = e * (uint32_t)N;
        const uint32_t cnt = wb + D2D_W_CNT + ((iter & 1u) << 8), cnt_other = wb + D2D_W_CNT + (((iter & 1u) ^ 1u) << 8);

        // ---- one coalesced pass over the env's inputs, software-pipelined: the NEXT env's loads are in flight
        // while this env computes, so a warp hides its own HBM latency ----------------------------------------
        const int aA = nxt.aA, aB = nxt.aB;
        const float2 tA = nxt.tA;
        const float4 pB = nxt.pB;
        const int ns_prev = nxt.ns;
        if (e + stride < num_envs) nxt = d2d_load_inputs(P, e + stride, lane, hasA, hasB);
        const bool actA = aA >= 0, actB = aB >= 0;
        if (actA || actB) D2D_TICK(3);     // inputs arrived

        // ---- envs/d2d_env.py:93-101: rb = a // n_pwr, p = a % n_pwr; rank inside the RB (actions.py:27-31) -------------
        const int rbA = d2d_div(aA, magicA), pA = aA - rbA * P.n_pwr_cue;
        const int rbB = d2d_div(aB, magicB), pB_ = aB - rbB * P.n_pwr_due;
        const uint32_t binA = (uint32_t)rbA & 63u, binB = (uint32_t)rbB & 63u;
        uint32_t rankA = 0, rankB = 0;
        if (actA) rankA = d2d_atoms_add(cnt + (binA << 2), 1u) & 0xffffu;
        if (actB) rankB = d2d_atoms_add(cnt + (binB << 2), 0x10001u) & 0xffffu;     // high half counts the SIDELINKs
        if (lane < 16) d2d_sts128(cnt_other + (lane << 4), 0.f, 0.f, 0.f, 0.f);       // next iteration's counters
        const float plA = actA ? d2d_lds32(wb + D2D_W_PWR + ((pA & (D2D_MAX_PWR_LEVELS - 1)) << 2)) : 0.0f;   // 10^(p/10)
        const float plB = actB ? d2d_lds32(wb + D2D_W_PWR + ((pB_ & (D2D_MAX_PWR_LEVELS - 1)) << 2)) : 0.0f;

        // ---- peer records: position, radiated weight w, and its value u at the MBS -------------------------------
        const float d2A = fmaf(tA.x, tA.x, tA.y * tA.y);                       // CUE -> MBS distance^2 (own link)
        const float lgA = d2d_lg2(d2A);
        const float gA = PLE2 ? d2d_rcp(d2A) : d2d_ex2(P.neg_half_ple * lgA);
        const float wA = plA * cA.x;
        const float d2Bm = fmaf(pB.x, pB.x, pB.y * pB.y);                      // DUE tx -> MBS distance^2 (as interferer)
        const float wB = plB * cB.x;
        const float uB = wB * d2d_gain<PLE2>(d2Bm, P.neg_half_ple);
        D2D_TICK(4);

        // ---- exclusive scan of the 64 RB counts: lane l scans bins l and l + 32, packed 16 + 16 bits ---------------
        __syncwarp();
        const uint32_t c0 = d2d_lds32u(cnt + (lane << 2)) & 0xffffu, c1 = d2d_lds32u(cnt + ((lane + 32) << 2)) & 0xffffu;
        uint32_t incl = c0 | (c1 << 16);
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += up;
        }
        const uint32_t total0 = __shfl_sync(0xffffffffu, incl, 31) & 0xffffu;   // links in bins 0..31
        d2d_sts32(offs + (lane << 2), (incl & 0xffffu) - c0);
        d2d_sts32(offs + ((lane + 32) << 2), total0 + (incl >> 16) - c1);
        __syncwarp();
        // ---- scatter the records into RB order --------------------------------------------------------------------
        uint32_t begA = 0, begB = 0, nA = 0, nB = 0, sideA = 0;
        if (actA) {
            begA = d2d_lds32u(offs + (binA << 2));
            const uint32_t c = d2d_lds32u(cnt + (binA << 2));
            nA = c & 0xffffu; sideA = c >> 16;
            d2d_sts128(sorted + ((begA + rankA) << 4), tA.x, tA.y, wA, wA * gA);
        }
        if (actB) {
            begB = d2d_lds32u(offs + (binB << 2));
            nB = d2d_lds32u(cnt + (binB << 2)) & 0xffffu;
            d2d_sts128(sorted + ((begB + rankB) << 4), pB.x, pB.y, wB, uB);
        }
        __syncwarp();
        D2D_TICK(5);

        // ---- simulator.py:95-101: interference at each victim's receiver (an absent victim has an empty range) ----------
        float dminA = 3.0e38f, dminB = 3.0e38f;
        const float IA = EXACT ? d2d_walk_rx<PLE2, true>(sorted, begA, nA, begA + rankA, 0.f, 0.f, P.neg_half_ple, dminA)
                               : d2d_walk_mbs(sorted, begA, nA, begA + rankA);
        const float IB = d2d_walk_rx<PLE2, EXACT>(sorted, begB, nB, begB + rankB, pB.z, pB.w, P.neg_half_ple, dminB);
        if (IA + IB >= 0.f) D2D_TICK(6);

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
        if (oA.cap + oB.cap >= 0.f) D2D_TICK(7);

        // ---- envs/reward_fn.py:27-44 -----------------------------------------------------------------------------------
        const bool bad = __any_sync(0xffffffffu, actA && sideA != 0u && oA.cap <= P.min_cap);
        const int n_act = __popc(__ballot_sync(0xffffffffu, actA)) + __popc(__ballot_sync(0xffffffffu, actB));
        float cap_sum = oA.cap + oB.cap;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cap_sum += __shfl_xor_sync(0xffffffffu, cap_sum, s);
        const float reward = bad ? -1.0f : __fdividef(cap_sum, (float)n_act);
        if (reward > -2.f) D2D_TICK(8);

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
        D2D_TICK(9);
        st_reward += reward; st_cap += cap_sum; st_reward2 = fmaf(reward, reward, st_reward2);
        st_pen += bad ? 1 : 0;

        // ---- rare: fp64 recomputation of flagged links (d2d_common.cuh).  Runs after the stores, when none of the
        // per-link state above is live; the whole warp cooperates on each flagged link. -------------------------------
        if (D2D_RESCUE_ENABLED && __any_sync(0xffffffffu, need != 0)) {
            const int32_t *act = P.actions + row0;
            const float2 *pe = reinterpret_cast<const float2 *>(P.pos) + e * (uint32_t)V;
            const double2 *pe64 = P.pos64 ? reinterpret_cast<const double2 *>(P.pos64) + (int64_t)e * V : nullptr;
            const uint32_t keyA = actA ? (uint32_t)rbA : (D2D_INACTIVE_KEY | (uint32_t)lane);
            const uint32_t keyB = actB ? (uint32_t)rbB : (D2D_INACTIVE_KEY | 32u | (uint32_t)lane);
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
    }

    D2D_TICK(10);
    if (P.stats && lane < 6) {
        // one fire-and-forget fp64 reduction per statistic and warp, spread over the replicas
        const int resc0 = __shfl_sync(0x3fu, st_resc, 0);
        const double v = lane == 0 ? (double)st_reward : lane == 1 ? (double)st_cap : lane == 2 ? (double)st_reward2
                       : lane == 3 ? (double)(iter - 1) : lane == 4 ? (double)st_pen : (double)resc0;
        const unsigned w_global = blockIdx.x * D2D_WARP_WARPS_PER_BLOCK + warp;
        if (v != 0.0) atomicAdd(P.stats + (w_global % D2D_STATS_REPLICAS) * 8 + lane, v);
    }
}
