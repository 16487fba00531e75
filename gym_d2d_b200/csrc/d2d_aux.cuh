// d2d_aux.cuh - the kernels either side of the step: position upload, device-side reset, and the
// optional per-agent observation materialisation.
#pragma once

#include "d2d_common.cuh"

// Device.set_position over a batch (device.py:82-83): float64 [count][V][2] -> float32 state.
// Device 0 (the MBS) is pinned to the origin like simulator.py:63-64.
__global__ void d2d_set_positions_kernel(const double *__restrict__ src, float *__restrict__ dst,
                                         double *__restrict__ dst64, int64_t count, int V) {
    const int64_t total = count * V;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % V);
        double2 s = reinterpret_cast<const double2 *>(src)[i];
        if (v == 0) s = make_double2(0.0, 0.0);
        reinterpret_cast<float2 *>(dst)[i] = make_float2((float)s.x, (float)s.y);
        if (dst64) reinterpret_cast<double2 *>(dst64)[i] = s;      // unrounded shadow for the fp64 rescue path
    }
}

// uniform-in-disc draw (position.py:24-28): theta = 2 pi u1, r = radius sqrt(u2)
__device__ __forceinline__ float2 d2d_disc_draw(uint64_t seed, uint64_t genv, uint32_t dev, uint32_t attempt, float radius) {
    const uint4 o = d2d_philox4x32_10(make_uint4((uint32_t)genv, (uint32_t)(genv >> 32), dev, attempt),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    // 24-bit uniforms centred in their cell, so u is never 0 (r = 0 would put a receiver on its transmitter)
    const float u1 = ((float)(o.x >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(o.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float s, c;
    sincospif(2.0f * u1, &s, &c);
    const float r = radius * sqrtf(u2);
    return make_float2(r * c, r * s);
}

// Simulator.reset (simulator.py:61-75): one thread per (env, link slot); a DUE thread draws its tx in
// the cell and re-draws its rx around the tx until it falls inside the cell (position.py:31-45).
__global__ void d2d_reset_kernel(float *__restrict__ pos, double *__restrict__ pos64, uint8_t *__restrict__ step_count,
                                 const uint8_t *__restrict__ env_mask, int64_t num_envs, int C, int D, float cell_radius,
                                 float d2d_radius, uint64_t seed, uint64_t first_global_env) {
    const int N = C + D, V = 1 + C + 2 * D;
    const int64_t total = num_envs * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / N;
        const int j = (int)(i - e * N);
        if (env_mask && !env_mask[e]) continue;
        float2 *pe = reinterpret_cast<float2 *>(pos) + e * V;
        const uint64_t g = first_global_env + (uint64_t)e;
        double2 *pe64 = pos64 ? reinterpret_cast<double2 *>(pos64) + e * V : nullptr;
        if (j == 0) {
            pe[0] = make_float2(0.f, 0.f);                       // simulator.py:63-64
            if (pe64) pe64[0] = make_double2(0.0, 0.0);
            if (step_count) step_count[e] = 0;                   // envs/d2d_env.py:46
        }
        if (j < C) {
            const float2 c = d2d_disc_draw(seed, g, (uint32_t)(1 + j), 0, cell_radius);
            pe[1 + j] = c;
            if (pe64) pe64[1 + j] = make_double2((double)c.x, (double)c.y);
        } else {
            const int t = 1 + C + 2 * (j - C);
            const float2 tx = d2d_disc_draw(seed, g, (uint32_t)t, 0, cell_radius);
            float2 rx = tx;
            for (uint32_t a = 0; a < 64; ++a) {
                const float2 o = d2d_disc_draw(seed, g, (uint32_t)(t + 1), a, d2d_radius);
                rx = make_float2(tx.x + o.x, tx.y + o.y);
                if (fmaf(rx.x, rx.x, rx.y * rx.y) <= cell_radius * cell_radius) break;
            }
            pe[t] = tx;
            pe[t + 1] = rx;
            if (pe64) { pe64[t] = make_double2((double)tx.x, (double)tx.y); pe64[t + 1] = make_double2((double)rx.x, (double)rx.y); }
        }
    }
}

// LinearObsFunction.get_state (envs/obs_fn.py:43-53): agent i's vector = its own row, then every other
// row in link order.  One block per (env, agent); float2 granularity (3 per row).
__global__ void d2d_per_agent_obs_kernel(const float *__restrict__ table, float *__restrict__ out, int N) {
    const int64_t e = blockIdx.x / N;
    const int i = (int)(blockIdx.x - e * N);
    const float2 *src = reinterpret_cast<const float2 *>(table) + e * N * 3;
    float2 *dst = reinterpret_cast<float2 *>(out) + (e * N + i) * (int64_t)N * 3;
    for (int m = threadIdx.x; m < 3 * N; m += blockDim.x) {
        const int row = m / 3, part = m - row * 3;
        const int link = row == 0 ? i : (row - 1 < i ? row - 1 : row);
        dst[m] = src[link * 3 + part];
    }
}


// Per-agent rewards from a step's results, one TEAM of threads per env (a warp for N <= 64, a block beyond):
//   mode 0  SystemCapacityRewardFunction's scalar broadcast to the acting agents (envs/reward_fn.py:44)
//   mode 1  ShannonRewardFunction (envs/reward_fn.py:47-57): log2(1 + 10^(sinr/10)) if sinr >= param else -1
//   mode 2  CueSinrShannonRewardFunction (envs/reward_fn.py:60-78): -1 if some OTHER action on the agent's RB is a
//           non-SIDELINK link with sinr < param, else log2(1 + 10^(sinr/10))
// Reads the decoded step results (sinr_dB from the observation table - after the fp64 pass, so threshold decisions see the
// rescued values) and re-derives the RB from the raw action (envs/d2d_env.py:93-101).  For modes 1 / 2 reward[e] becomes
// the mean agent reward and the reward statistics are accumulated here (the step kernel skips them).
template <int TEAM>
__global__ void d2d_agent_reward_kernel(const int32_t *__restrict__ actions, const float *__restrict__ obs,
                                        float *__restrict__ agent_reward, float *__restrict__ reward, double *__restrict__ stats,
                                        int64_t num_envs, int N, const int32_t *__restrict__ link_meta, int R, int mode, float param) {
    extern __shared__ uint32_t d2d_weak_smem[];
    const int teams = blockDim.x / TEAM, team = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
    uint32_t *weak = d2d_weak_smem + (size_t)team * R;
    auto team_sync = [&]() { if (TEAM == 32) __syncwarp(); else __syncthreads(); };
    float st_r = 0.f, st_r2 = 0.f;
    const int64_t rounds = (num_envs + (int64_t)gridDim.x * teams - 1) / ((int64_t)gridDim.x * teams);
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t e = (it * gridDim.x + blockIdx.x) * teams + team;
        const bool in = e < num_envs;                 // whole teams drop out together; barriers stay uniform per block
        const int32_t *act = actions + (in ? e : 0) * N;
        const float *ob = obs + (in ? e : 0) * N * 6;
        if (mode == 2) {
            for (int r = tl; r < R; r += TEAM) weak[r] = 0u;
            team_sync();
            for (int j = tl; in && j < N; j += TEAM) {
                const int meta = link_meta[j], npw = meta & 0xffff;          // power levels | SIDELINK << 16
                const uint32_t a = (uint32_t)act[j];
                if (!(meta >> 16) && a < (uint32_t)(R * npw) && ob[j * 6 + 4] < param) atomicAdd(&weak[a / (uint32_t)npw], 1u);
            }
            team_sync();
        }
        float sum = 0.f;
        int n_act = 0;
        for (int j = tl; in && j < N; j += TEAM) {
            const int meta = link_meta[j], npw = meta & 0xffff;
            const uint32_t a = (uint32_t)act[j];
            const bool live = a < (uint32_t)(R * npw);
            float rw = 0.f;
            if (live) {
                const float sinr = ob[j * 6 + 4];
                const float shannon = d2d_log2_1p(d2d_ex2(sinr * 0.33219280948873623f));     // log2(1 + 10^(sinr/10))
                if (mode == 0) rw = reward[e];
                else if (mode == 1) rw = sinr >= param ? shannon : -1.0f;
                else rw = weak[a / (uint32_t)npw] - ((!(meta >> 16) && sinr < param) ? 1u : 0u) > 0u ? -1.0f : shannon;
                sum += rw;
                ++n_act;
            }
            if (agent_reward) agent_reward[e * N + j] = rw;
        }
        if (mode != 0) {
            // mean over the acting agents -> reward[e]; team reduction through the warp, then shared memory for blocks
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, s); n_act += __shfl_xor_sync(0xffffffffu, n_act, s); }
            if (TEAM > 32) {
                __shared__ float red_s[8];
                __shared__ int red_n[8];
                team_sync();
                if ((tl & 31) == 0) { red_s[tl >> 5] = sum; red_n[tl >> 5] = n_act; }
                team_sync();
                sum = 0.f; n_act = 0;
                for (int w = 0; w < TEAM / 32; ++w) { sum += red_s[w]; n_act += red_n[w]; }
            }
            if (in && tl == 0) {
                const float r = n_act ? sum / (float)n_act : 0.f;
                reward[e] = r;
                st_r += r; st_r2 = fmaf(r, r, st_r2);
            }
        }
        team_sync();
    }
    if (mode != 0 && stats && tl == 0 && (st_r != 0.f || st_r2 != 0.f)) {
        double *dst = stats + ((blockIdx.x * teams + team) % 32) * 8;
        atomicAdd(dst + 0, (double)st_r);
        atomicAdd(dst + 2, (double)st_r2);
    }
}
