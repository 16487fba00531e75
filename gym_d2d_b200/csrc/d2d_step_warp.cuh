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
#define D2D_W_SORTED 0u        // float4 sorted[64 + 8]: peer records grouped by RB; the 8 slack records take the dead
                               // lanes' records (dummy bin 64) and the inline walk's over-read
#define D2D_W_CNT 1152u        // u32 count[2][68]: per-RB link count (bin 64 = dummy), double-buffered by iteration parity
#define D2D_W_CNT_STRIDE 272u
#define D2D_W_OFF 1696u        // u32 offset[68]: exclusive scan of count; offset[64] = 64 (the slack records)
#define D2D_W_PWR 1968u        // float pwr_lin[128]: the warp's copy of the integer-dBm -> mW table
#define D2D_W_BYTES 2480u

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
// same, with a compile-time byte offset folded into the instruction (no address add per access)
template <int OFF>
__device__ __forceinline__ float4 d2d_lds128_at(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ float d2d_lds32_at(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t d2d_lds32u(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// predicated shared atomic add (returns 0 when the predicate is off): no branch around a one-instruction body
__device__ __forceinline__ uint32_t d2d_atoms_add_if(bool p, uint32_t a, uint32_t x) {
    uint32_t old = 0;
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %1, 0; @q atom.shared.add.u32 %0, [%2], %3; }"
                 : "+r"(old) : "r"((uint32_t)p), "r"(a), "r"(x) : "memory");
    return old;
}
// index of the highest set bit (FLO) and removal of that bit - rescue path only
__device__ __forceinline__ uint32_t d2d_pop_bit(uint32_t &mask) {
    uint32_t k;
    asm("bfind.u32 %0, %1;" : "=r"(k) : "r"(mask));
    mask ^= 1u << k;
    return k;
}

// A victim's co-channel records are the range [beg, beg + n) of `sorted`; `valid` has bit t set when entry t exists
// and is not the victim itself.  The first D2D_WALK_INLINE entries are straight-line code with immediate offsets.

// (an RB may hold up to 64 links - the one-RB corner case - so entries 32.. are walked by index in a second rare loop)
__device__ __forceinline__ uint32_t d2d_valid_mask(uint32_t n, uint32_t rank) {
    const uint32_t m = n >= 32u ? 0xffffffffu : (1u << n) - 1u;
    return rank < 32u ? m & ~(1u << rank) : m;
}

// One interferer's term at a general receiver: w_k g(|tx_k - rx|)
template <bool PLE2, bool EXACT>
__device__ __forceinline__ float d2d_term_rx(const float4 r, bool valid, float rxx, float rxy, float nhp, float &dmin2) {
    const float dx = r.x - rxx, dy = r.y - rxy;
    const float d2 = fmaf(dx, dx, dy * dy);
    if (EXACT) dmin2 = valid ? fminf(dmin2, d2) : dmin2;
    return valid ? r.z * d2d_gain<PLE2>(d2, nhp) : 0.0f;
}
template <bool PLE2, bool EXACT, int T>
__device__ __forceinline__ void d2d_inline_rx(float &I, uint32_t base, uint32_t valid, float rxx, float rxy, float nhp, float &dmin2) {
    if constexpr (T < D2D_WALK_INLINE) {
        I += d2d_term_rx<PLE2, EXACT>(d2d_lds128_at<16 * T>(base), (valid >> T) & 1u, rxx, rxy, nhp, dmin2);
        d2d_inline_rx<PLE2, EXACT, T + 1>(I, base, valid, rxx, rxy, nhp, dmin2);
    }
}
template <bool PLE2, bool EXACT>
__device__ __forceinline__ float d2d_walk_rx(uint32_t sorted, uint32_t beg, uint32_t valid, uint32_t n, uint32_t rank,
                                             float rxx, float rxy, float nhp, float &dmin2) {
    float I = 0.0f;
    const uint32_t base = sorted + (beg << 4);
    d2d_inline_rx<PLE2, EXACT, 0>(I, base, valid, rxx, rxy, nhp, dmin2);
    valid >>= D2D_WALK_INLINE;
    for (uint32_t a = base + 16u * D2D_WALK_INLINE; valid; valid >>= 1, a += 16u)   // rare: > D2D_WALK_INLINE links on one RB
        I += d2d_term_rx<PLE2, EXACT>(d2d_lds128(a), valid & 1u, rxx, rxy, nhp, dmin2);
    for (uint32_t t = 32u; t < n; ++t)                                              // corner case: > 32 links on one RB
        I += d2d_term_rx<PLE2, EXACT>(d2d_lds128(base + (t << 4)), t != rank, rxx, rxy, nhp, dmin2);
    return I;
}
// Interference at the MBS: the records carry u_k = w_k g(|tx_k|) in .w
template <int T>
__device__ __forceinline__ void d2d_inline_mbs(float &I, uint32_t base, uint32_t valid) {
    if constexpr (T < D2D_WALK_INLINE) {
        const float u = d2d_lds32_at<16 * T + 12>(base);
        I += ((valid >> T) & 1u) ? u : 0.0f;
        d2d_inline_mbs<T + 1>(I, base, valid);
    }
}
__device__ __forceinline__ float d2d_walk_mbs(uint32_t sorted, uint32_t beg, uint32_t valid, uint32_t n, uint32_t rank) {
    float I = 0.0f;
    const uint32_t base = sorted + (beg << 4);
    d2d_inline_mbs<0>(I, base, valid);
    valid >>= D2D_WALK_INLINE;
    for (uint32_t a = base + 16u * D2D_WALK_INLINE + 12u; valid; valid >>= 1, a += 16u) {
        const float u = d2d_lds32(a);
        I += (valid & 1u) ? u : 0.0f;
    }
    for (uint32_t t = 32u; t < n; ++t) {
        const float u = d2d_lds32(base + 12u + (t << 4));
        I += t != rank ? u : 0.0f;
    }
    return I;
}

// Predicated global stores (no branch, no reconvergence bookkeeping around a two-line body)
__device__ __forceinline__ void d2d_stg64_if(bool p, float2 *ptr, float x, float y) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.v2.f32 [%1], {%2, %3}; }" ::"r"((uint32_t)p), "l"(ptr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void d2d_stg32_if(bool p, float *ptr, float x) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.f32 [%1], %2; }" ::"r"((uint32_t)p), "l"(ptr), "f"(x) : "memory");
}
__device__ __forceinline__ void d2d_stg16_if(bool p, int16_t *ptr, int x) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.u16 [%1], %2; }" ::"r"((uint32_t)p), "l"(ptr), "h"((short)x) : "memory");
}

// One env's inputs as a lane sees them: its CUE action + transmitter, its DUE action + (tx, rx) pair.
// Lanes without a link (lane >= C or >= D) and agents absent this step (action < 0) are "dead": they keep running
// the same straight-line code on benign values, sort into a dummy RB bin and are masked out of every sum and store.
struct D2DLaneIn {
    int aA, aB, ns;   // ns: this env's step counter (lane 0 only)
    float2 tA;        // CUE transmitter (its receiver is the MBS at the origin)
    float4 pB;        // DUE (tx_x, tx_y, rx_x, rx_y)
};
__device__ __forceinline__ D2DLaneIn d2d_load_inputs(const D2DParams &P, uint32_t e, uint32_t lane, bool hasA, bool hasB) {
    D2DLaneIn in;
    in.aA = -1; in.aB = -1;
    in.tA = make_float2(1.f, 0.f);                    // benign: unit distance
    in.pB = make_float4(1.f, 0.f, 0.f, 0.f);
    const uint32_t row0 = e * (uint32_t)P.N, pos0 = e * (uint32_t)P.V;
    const float2 *pos = reinterpret_cast<const float2 *>(P.pos);
    if (hasA) { in.aA = __ldg(P.actions + (row0 + lane)); in.tA = __ldg(pos + (pos0 + 1u + lane)); }
    if (hasB) {
        in.aB = __ldg(P.actions + (row0 + (uint32_t)P.C + lane));
        const uint32_t q = pos0 + 1u + (uint32_t)P.C + 2u * lane;
        if (P.align4) {
            in.pB = __ldg(reinterpret_cast<const float4 *>(pos + q));
        } else {
            const float2 t = __ldg(pos + q), r = __ldg(pos + (q + 1u));
            in.pB = make_float4(t.x, t.y, r.x, r.y);
        }
    }
    in.ns = (lane == 0 && P.step_count) ? (int)P.step_count[e] : 0;
    return in;
}

#ifdef D2D_TIMELINE   // debug builds only: SM-clock timestamps of one warp's phases (profiles/timeline.py)
__device__ unsigned long long d2d_dbg[16];
#define D2D_TICK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) d2d_dbg[i] = clock64(); } while (0)
#else
#define D2D_TICK(i) do { } while (0)
#endif

// FULL: the caller passed exactly the core outputs (obs, capacity, reward, done) - the VecD2DEnv default - so no
// pointer is tested on the hot path; otherwise every output is optional and checked.
template <bool PLE2, bool EXACT, int WPB, bool FULL>
__global__ void __launch_bounds__(WPB * 32, D2D_WARP_MIN_BLOCKS(WPB))
d2d_step_warp_kernel(const D2DParams P) {
    __shared__ __align__(16) unsigned char smem[WPB * D2D_W_BYTES];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t N = (uint32_t)P.N, C = (uint32_t)P.C, V = (uint32_t)P.V, D = N - C;
    D2D_TICK(0);
    d2d_pdl_launch_dependents();
    uint32_t wb = (uint32_t)__cvta_generic_to_shared(smem) + warp * D2D_W_BYTES;
    asm volatile("mov.u32 %0, %0;" : "+r"(wb));      // opaque: keep the base in a register instead of re-deriving it
    const uint32_t sorted = wb + D2D_W_SORTED, offs = wb + D2D_W_OFF;

    // both parity buffers of the RB counters (64 bins + the dummy bin each) start at zero; each iteration re-zeroes the
    // one it is not using.  offset[64] = 64 forever: the dummy bin's records land in the slack behind the sorted array.
    for (uint32_t i = lane; i < 2u * D2D_W_CNT_STRIDE / 16u; i += 32u) d2d_sts128(wb + D2D_W_CNT + (i << 4), 0.f, 0.f, 0.f, 0.f);
    if (lane == 0) d2d_sts32(offs + 64u * 4u, 64u);
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
    const uint32_t dump = 64u + (lane & 7u);                  // where a dead lane parks its record

    // per-warp partial statistics (fp32 over the few envs one warp visits; flushed once as fp64 atomics)
    float st_reward = 0.f, st_cap = 0.f, st_reward2 = 0.f;
    int st_pen = 0, st_resc = 0;

    // 32-bit indexing: the host launches at most 2^31 / max(6N, 2V) envs per call (d2d_step chunks larger batches)
    uint32_t iter = 1;
    const uint32_t num_envs = (uint32_t)P.num_envs, stride = gridDim.x * WPB;
    uint32_t e = blockIdx.x * WPB + warp;
    D2DLaneIn nxt;
    D2D_TICK(1);
    d2d_pdl_wait();
    D2D_TICK(2);
    if (e < num_envs) nxt = d2d_load_inputs(P, e, lane, hasA, hasB);
    d2d_sts128(wb + D2D_W_PWR + (lane << 4), lut.x, lut.y, lut.z, lut.w);
    __syncwarp();
    for (; e < num_envs; e += stride, ++iter) {
        const uint32_t row0 = e * N;
        const uint32_t cnt = wb + D2D_W_CNT + (iter & 1u) * D2D_W_CNT_STRIDE;
        const uint32_t cnt_other = wb + D2D_W_CNT + ((iter & 1u) ^ 1u) * D2D_W_CNT_STRIDE;

        // ---- one coalesced pass over the env's inputs, software-pipelined: the NEXT env's loads are in flight
        // while this env computes, so a warp hides its own HBM latency ----------------------------------------
        const int aA = nxt.aA, aB = nxt.aB;
        const float2 tA = nxt.tA;
        const float4 pB = nxt.pB;
        const int ns_prev = nxt.ns;
        if (e + stride < num_envs) nxt = d2d_load_inputs(P, e + stride, lane, hasA, hasB);
        const bool liveA = aA >= 0, liveB = aB >= 0;       // has a link AND the agent acts this step
        if (liveA || liveB) D2D_TICK(3);     // inputs arrived

        // ---- envs/d2d_env.py:93-101: rb = a // n_pwr, p = a % n_pwr; rank inside the RB (actions.py:27-31) -------------
        const uint32_t rbA = __umulhi((uint32_t)aA, P.magic_cue) + ((uint32_t)aA & P.npw1_cue);
        const uint32_t rbB = __umulhi((uint32_t)aB, P.magic_due) + ((uint32_t)aB & P.npw1_due);
        const uint32_t pA = (uint32_t)aA - rbA * (uint32_t)P.n_pwr_cue, pB_ = (uint32_t)aB - rbB * (uint32_t)P.n_pwr_due;
        const uint32_t binA = liveA ? (rbA & 63u) : 64u, binB = liveB ? (rbB & 63u) : 64u;    // dead lanes: dummy bin (never counted)
        const uint32_t rankA = d2d_atoms_add_if(liveA, cnt + (binA << 2), 1u) & 0xffffu;
        const uint32_t rankB = d2d_atoms_add_if(liveB, cnt + (binB << 2), 0x10001u) & 0xffffu;   // high half counts the SIDELINKs
        if (lane < D2D_W_CNT_STRIDE / 16u) d2d_sts128(cnt_other + (lane << 4), 0.f, 0.f, 0.f, 0.f);   // next iteration's counters
        const float lutA = d2d_lds32(wb + D2D_W_PWR + ((pA & (D2D_MAX_PWR_LEVELS - 1)) << 2));   // 10^(p/10)
        const float lutB = d2d_lds32(wb + D2D_W_PWR + ((pB_ & (D2D_MAX_PWR_LEVELS - 1)) << 2));
        const float plA = liveA ? lutA : 0.0f, plB = liveB ? lutB : 0.0f;

        // ---- peer records: position, radiated weight w, and its value u at the MBS -------------------------------
        const float d2A = fmaf(tA.x, tA.x, tA.y * tA.y);                       // CUE -> MBS distance^2 (own link)
        const float lgA = d2d_lg2(d2A);
        const float gA = PLE2 ? d2d_rcp(d2A) : d2d_ex2(P.neg_half_ple * lgA);
        const float wA = plA * cA.x;
        const float d2Bm = fmaf(pB.x, pB.x, pB.y * pB.y);                      // DUE tx -> MBS distance^2 (as interferer)
        const float wB = plB * cB.x;
        const float uB = wB * d2d_gain<PLE2>(d2Bm, P.neg_half_ple);
        const float dxB = pB.x - pB.z, dyB = pB.y - pB.w;                      // DUE own link
        const float d2B = fmaf(dxB, dxB, dyB * dyB);
        const float lgB = d2d_lg2(d2B);
        const float gB = PLE2 ? d2d_rcp(d2B) : d2d_ex2(P.neg_half_ple * lgB);
        D2D_TICK(4);

        // ---- exclusive scan of the 64 RB counts: lane l scans bins l and l + 32, packed 16 + 16 bits ---------------
        __syncwarp();
        const uint32_t c0 = d2d_lds32u(cnt + (lane << 2)) & 0xffffu, c1 = d2d_lds32u(cnt + ((lane + 32u) << 2)) & 0xffffu;
        uint32_t incl = c0 | (c1 << 16);
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, s);
            incl += lane >= (uint32_t)s ? up : 0u;
        }
        const uint32_t total0 = __shfl_sync(0xffffffffu, incl, 31) & 0xffffu;   // links in bins 0..31
        d2d_sts32(offs + (lane << 2), (incl & 0xffffu) - c0);
        d2d_sts32(offs + ((lane + 32u) << 2), total0 + (incl >> 16) - c1);
        __syncwarp();
        // ---- scatter the records into RB order (dead lanes: offset[64] = 64 -> the slack records) ----------------------
        const uint32_t begA = d2d_lds32u(offs + (binA << 2)), begB = d2d_lds32u(offs + (binB << 2));
        const uint32_t ctA = d2d_lds32u(cnt + (binA << 2)), ctB = d2d_lds32u(cnt + (binB << 2));
        const uint32_t nA = ctA & 0xffffu, nB = ctB & 0xffffu;
        // validA/B: bit t set <=> entry t of the victim's RB range exists and is not the victim (dead lane: 0)
        const uint32_t validA = liveA ? d2d_valid_mask(nA, rankA) : 0u, validB = liveB ? d2d_valid_mask(nB, rankB) : 0u;
        d2d_sts128(sorted + ((liveA ? begA + rankA : dump) << 4), tA.x, tA.y, wA, wA * gA);
        d2d_sts128(sorted + ((liveB ? begB + rankB : dump) << 4), pB.x, pB.y, wB, uB);
        __syncwarp();
        D2D_TICK(5);

        // ---- simulator.py:95-101: interference at each victim's receiver (a dead lane has an empty range) ----------
        float dminA = 3.0e38f, dminB = 3.0e38f;
        const float IA = EXACT ? d2d_walk_rx<PLE2, true>(sorted, begA, validA, nA, rankA, 0.f, 0.f, P.neg_half_ple, dminA)
                               : d2d_walk_mbs(sorted, begA, validA, nA, rankA);
        const float IB = d2d_walk_rx<PLE2, EXACT>(sorted, begB, validB, nB, rankB, pB.z, pB.w, P.neg_half_ple, dminB);
        if (IA + IB >= 0.f) D2D_TICK(6);

        // ---- per-link epilogue (simulator.py:93,106-107,110-127,144-154); dead lanes are zeroed ------------------------
        D2DLinkOut oA = d2d_link_epilogue<PLE2>((int)pA, plA, lgA, gA, IA, cA, sA, P);
        D2DLinkOut oB = d2d_link_epilogue<PLE2>((int)pB_, plB, lgB, gB, IB, cB, sB, P);
        oA.sinr_dB = liveA ? oA.sinr_dB : 0.f; oA.snr_dB = liveA ? oA.snr_dB : 0.f; oA.cap = liveA ? oA.cap : 0.f;
        oB.sinr_dB = liveB ? oB.sinr_dB : 0.f; oB.snr_dB = liveB ? oB.snr_dB : 0.f; oB.cap = liveB ? oB.cap : 0.f;
        const bool needA = liveA && d2d_needs_rescue<EXACT>(oA, fminf(dminA, d2A), P);
        const bool needB = liveB && d2d_needs_rescue<EXACT>(oB, fminf(dminB, d2B), P);
        if (oA.cap + oB.cap >= 0.f) D2D_TICK(7);

        // ---- envs/reward_fn.py:27-44 -----------------------------------------------------------------------------------
        const bool bad = __any_sync(0xffffffffu, liveA && (ctA >> 16) != 0u && oA.cap <= P.min_cap);
        const int n_act = __popc(__ballot_sync(0xffffffffu, liveA)) + __popc(__ballot_sync(0xffffffffu, liveB));
        float cap_sum = oA.cap + oB.cap;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) cap_sum += __shfl_xor_sync(0xffffffffu, cap_sum, s);
        const float reward = bad ? -1.0f : __fdividef(cap_sum, (float)n_act);
        if (reward > -2.f) D2D_TICK(8);

        // ---- outputs: compact observation table (envs/obs_fn.py:55-61), capacity, optional info.  Rows of absent
        // agents carry their positions and zeros (the reference has no row for them). ---------------------------------
        const uint32_t iA = row0 + jA, iB = row0 + jB;
        if (FULL || P.obs) {
            float2 *oa = reinterpret_cast<float2 *>(P.obs) + iA * 3u, *ob = reinterpret_cast<float2 *>(P.obs) + iB * 3u;
            d2d_stg64_if(hasA, oa, tA.x, tA.y);
            d2d_stg64_if(hasA, oa + 1, 0.f, 0.f);
            d2d_stg64_if(hasA, oa + 2, oA.sinr_dB, oA.snr_dB);
            d2d_stg64_if(hasB, ob, pB.x, pB.y);
            d2d_stg64_if(hasB, ob + 1, pB.z, pB.w);
            d2d_stg64_if(hasB, ob + 2, oB.sinr_dB, oB.snr_dB);
        }
        if (FULL || P.cap) {
            d2d_stg32_if(hasA, P.cap + iA, oA.cap);
            d2d_stg32_if(hasB, P.cap + iB, oB.cap);
        }
        if (!FULL) {
            if (P.rate) {
                d2d_stg32_if(hasA, P.rate + iA, liveA ? oA.rate : 0.f);
                d2d_stg32_if(hasB, P.rate + iB, liveB ? oB.rate : 0.f);
            }
            if (P.rb_out) {
                d2d_stg16_if(hasA, P.rb_out + iA, liveA ? (int)rbA : 0);
                d2d_stg16_if(hasB, P.rb_out + iB, liveB ? (int)rbB : 0);
            }
            if (P.pwr_out) {
                d2d_stg16_if(hasA, P.pwr_out + iA, liveA ? (int)pA : 0);
                d2d_stg16_if(hasB, P.pwr_out + iB, liveB ? (int)pB_ : 0);
            }
        }
        if (lane == 0) {
            // envs/d2d_env.py:65,68: num_steps += 1; done = num_steps >= EPISODE_LENGTH
            const int ns = min(ns_prev + 1, 255);
            if (P.step_count) P.step_count[e] = (uint8_t)ns;
            if (FULL || P.reward) P.reward[e] = reward;
            if (FULL || P.done) P.done[e] = ns >= P.episode_length ? 1 : 0;
        }
        D2D_TICK(9);
        st_reward += reward; st_cap += cap_sum; st_reward2 = fmaf(reward, reward, st_reward2);
        st_pen += bad ? 1 : 0;

        // ---- rare: fp64 recomputation of flagged links (d2d_common.cuh).  Runs after the stores, when none of the
        // per-link state above is live; the whole warp cooperates on each flagged link. -------------------------------
        if (D2D_RESCUE_ENABLED && __any_sync(0xffffffffu, needA || needB)) {
            const int32_t *act = P.actions + row0;
            const float2 *pe = reinterpret_cast<const float2 *>(P.pos) + e * V;
            const double2 *pe64 = P.pos64 ? reinterpret_cast<const double2 *>(P.pos64) + (int64_t)e * V : nullptr;
            const uint32_t keyA = liveA ? rbA : (D2D_INACTIVE_KEY | lane), keyB = liveB ? rbB : (D2D_INACTIVE_KEY | 32u | lane);
#pragma unroll 1
            for (int s = 0; s < 2; ++s) {
                uint32_t todo = __ballot_sync(0xffffffffu, s ? needB : needA);
                while (todo) {
                    const int L = (int)d2d_pop_bit(todo);
                    const int j = s ? (int)C + L : L;
                    const uint32_t key = __shfl_sync(0xffffffffu, s ? keyB : keyA, L);
                    const double2 rx = d2d_pos_f64(pe, pe64, d2d_rx_dev(j, (int)C));
                    double I = 0.0;
                    if (keyA == key && (int)jA != j) I += d2d_ix_term_f64<PLE2>((int)jA, rx, pe, pe64, act, P);
                    if (keyB == key && (int)jB != j) I += d2d_ix_term_f64<PLE2>((int)jB, rx, pe, pe64, act, P);
#pragma unroll
                    for (int sh = 16; sh > 0; sh >>= 1) I += __shfl_xor_sync(0xffffffffu, I, sh);
                    if (lane == 0) {
                        const D2DLinkOut o = d2d_link_f64<PLE2>(j, d2d_pos_f64(pe, pe64, d2d_tx_dev(j, (int)C)), rx, I,
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
        const unsigned w_global = blockIdx.x * WPB + warp;
        if (v != 0.0) atomicAdd(P.stats + (w_global % D2D_STATS_REPLICAS) * 8 + lane, v);
    }
}
