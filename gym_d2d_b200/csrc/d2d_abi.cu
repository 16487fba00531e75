// d2d_abi.cu - the C ABI of libd2d_b200.so (include/d2d_b200.h): handle management, host-side folding of
// the link-budget constants, and the kernel launches.  sm_100a only; there is no CPU path.
#include "../../include/d2d_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "d2d_aux.cuh"
#include "d2d_common.cuh"
#include "d2d_step_block.cuh"
#include "d2d_step_dense.cuh"
#include "d2d_step_warp.cuh"

// (links per thread, threads per block) instantiations of the dense kernel
#define D2D_DENSE_SHAPES(X) X(1, 256) X(2, 256) X(3, 256) X(4, 256) X(1, 320) X(2, 320) X(3, 320)
#define D2D_DENSE_PLAN_CASE(LPT_, BT_) if (h->lpt == LPT_ && h->dense_bt == BT_) rc = D2D_PLAN_DENSE(LPT_, BT_);
#define D2D_DENSE_LAUNCH_CASE(LPT_, BT_) if (h->lpt == LPT_ && h->dense_bt == BT_) err = D2D_LAUNCH_DENSE(LPT_, BT_);

static_assert(D2D_STATS_REPLICAS * 8 * sizeof(double) == 65536, "stats layout");

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

#define D2D_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t err__ = (call);                                                                 \
        if (err__ != cudaSuccess)                                                                   \
            return fail(D2D_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));        \
    } while (0)

constexpr double kSpeedOfLight = 299792458.0;   // path_loss.py:9

}  // namespace

struct d2d_handle {
    d2d_config_t cfg{};
    int N = 0, V = 0;
    int num_sms = 0;
    bool ple2 = true;
    bool use_warp = true;
    bool spec = false;         // warp kernel instantiated for the reference's default EnvConfig shape
    bool uniform = false;      // every CUE link has the same constants, and every DUE link (no per-device overrides)
    D2DLinkA u_cue{}, u_due{};
    D2DLinkD ud_cue{}, ud_due{};
    float us_cue[2] = {0, 0}, us_due[2] = {0, 0};
    int wpb = 4;               // warps per block of the warp kernel
    int dense_bt = 0;          // dense kernel (d2d_step_dense.cuh): threads per block, 0 = not used
    int bin_cap = 0;           // dense kernel: record slots per RB bin
    int lpt = 0;               // block / dense kernel: links per thread held in registers (0 = the generic shared-memory kernel)
    int64_t chunk_override = 0;  // D2D_B200_CHUNK: force small launch chunks (tests of the > 2^31-element path)
    bool pdl = true;           // programmatic dependent launch (D2D_B200_PDL=0 disables)
    int grid = 0, block = 0, smem = 0, envs_per_block = 0;
    double K_dB = 0.0, ple = 2.0;
    D2DLinkA *dA = nullptr;
    D2DLinkB *dB = nullptr;
    D2DLinkD *dD = nullptr;
    int32_t *dMeta = nullptr;  // [N] power levels | SIDELINK << 16 (general-topology kernel, fp64 helpers)
    float *dPwr = nullptr;
    double *dPwrD = nullptr;
    // bound state (caller-owned)
    float *pos = nullptr;
    double *pos64 = nullptr;
    uint8_t *step_count = nullptr;
    double *stats = nullptr;
    // staging for d2d_step_host / d2d_set_positions (handle-owned, allocated on first use)
    void *stage2[2][9] = {};
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {}, ev_kernel[2] = {}, ev_out[2] = {};
    bool slot_used[2] = {false, false};
    bool pipe_ready = false;
    double *stage_pos = nullptr;
    int64_t stage_pos_envs = 0;
    int64_t launches = 0;
    uint64_t rng_calls = 0;      // ShadowingPathLoss: number of step calls so far (every call draws fresh values)
};

namespace {

int ensure_device(const d2d_handle *h) {
    int cur = -1;
    D2D_CUDA(cudaGetDevice(&cur));
    if (cur != h->cfg.cuda_device) D2D_CUDA(cudaSetDevice(h->cfg.cuda_device));
    return D2D_OK;
}

D2DParams make_params(const d2d_handle *h, const d2d_step_io_t *io) {
    D2DParams P{};
    P.num_envs = h->cfg.num_envs;
    P.N = h->N; P.V = h->V; P.C = h->cfg.num_cues; P.R = h->cfg.num_rbs;
    P.n_pwr_cue = h->cfg.n_pwr_cue; P.n_pwr_due = h->cfg.n_pwr_due;
    P.episode_length = h->cfg.episode_length;
    P.nbins = h->cfg.num_rbs;
    P.bin_cap = h->bin_cap;
    P.magic_cue = d2d_div_magic(h->cfg.n_pwr_cue);
    P.magic_due = d2d_div_magic(h->cfg.n_pwr_due);
    P.npw1_cue = h->cfg.n_pwr_cue == 1 ? 0xffffffffu : 0u;
    P.npw1_due = h->cfg.n_pwr_due == 1 ? 0xffffffffu : 0u;
    P.align4 = (h->V % 2 == 0) && ((1 + h->cfg.num_cues) % 2 == 0) && ((uintptr_t)h->pos % 16 == 0);
    P.ple = (float)h->ple;
    P.neg_half_ple = (float)(-0.5 * h->ple);
    P.snr_slope = (float)(5.0 * h->ple * std::log10(2.0));
    P.min_cap = (float)h->cfg.min_capacity_mbps;
    // worst-case fp32 error of SINR_dB is ~6e-6 dB for ple = 2 (d2d_common.cuh); x 1e4 for a 1e-4 relative bound
    const bool shadowing = h->cfg.path_loss_model == D2D_PL_SHADOWING;
    P.rescue_band_dB = (h->ple2 && !shadowing) ? 0.0625f : 0.5f;
    P.shadow_chi_dB = shadowing ? (float)h->cfg.shadow_chi_dB : 0.0f;
    P.shadow_d0sq = (float)(h->cfg.shadow_d0_m * h->cfg.shadow_d0_m);
    P.shadow_chi_d = shadowing ? h->cfg.shadow_chi_dB : 0.0;
    P.shadow_d0sq_d = h->cfg.shadow_d0_m * h->cfg.shadow_d0_m;
    P.rng_seed = h->cfg.rng_seed; P.first_global_env = h->cfg.first_global_env; P.rng_step = h->rng_calls;
    if (h->pos64) {
        // positions were rounded to fp32: each coordinate is off by <= ulp(R)/2, a distance by <= ~sqrt(2) ulp(R),
        // i.e. 10 ple log10(e) * sqrt(2) ulp(R) / d dB per term; recompute whatever that could push past 1e-4 relative
        const double ulp = std::ldexp(1.0, std::ilogb(std::max(h->cfg.cell_radius_m, 1.0)) - 23);
        P.rescue_c = (float)(2.0 * 1e4 * 4.3429448190325 * h->ple * std::sqrt(2.0) * ulp);
        P.rescue_dmin2 = (float)std::pow(2.0 * 1e4 * 0.5 * h->ple * std::sqrt(2.0) * ulp, 2.0);   // capacity: d(ln r) = ple * dd/d
    }
    P.u_cue = make_float4(h->u_cue.tx_lin0, h->u_cue.a_lin, h->u_cue.inv_noise, h->u_cue.snr0_dB);
    P.u_due = make_float4(h->u_due.tx_lin0, h->u_due.a_lin, h->u_due.inv_noise, h->u_due.snr0_dB);
    P.us_cue = make_float2(h->us_cue[0], h->us_cue[1]);
    P.us_due = make_float2(h->us_due[0], h->us_due[1]);
    P.uniform = h->uniform ? 1 : 0;
    P.ud_cue = h->ud_cue; P.ud_due = h->ud_due;
    P.reward_fn = h->cfg.reward_fn;
    P.ple_d = h->ple;
    P.linkA = h->dA; P.linkB = h->dB; P.linkD = h->dD; P.link_meta = h->dMeta; P.pwr_lin = h->dPwr; P.pwr_lin_d = h->dPwrD;
    P.pos = h->pos; P.pos64 = h->pos64; P.step_count = h->step_count; P.stats = h->stats;
    P.actions = io->actions; P.obs = io->obs; P.cap = io->capacity_mbps; P.reward = io->reward;
    P.done = io->done; P.rate = io->rate_bps; P.rb_out = io->rb; P.pwr_out = io->tx_pwr_dBm;
    return P;
}

// Launch a step kernel with programmatic stream serialisation (PDL): it may begin launching while the previous
// kernel in the stream drains; the kernel itself orders its memory accesses with griddepcontrol.wait.
template <typename K>
cudaError_t launch_step(K kernel, int grid, int block, size_t smem, cudaStream_t st, const D2DParams &P, bool pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, P);
}

template <typename K>
int allow_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) D2D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return D2D_OK;
}

// every (EXACT, FULL) instantiation d2d_step may launch for this handle needs the dynamic shared-memory opt-in
template <bool PLE2, int WPB, bool SPEC>
int allow_smem_warp(size_t smem) {
    int rc = allow_smem(d2d_step_warp_kernel<PLE2, false, WPB, false, SPEC, false>, smem);
    if (!rc) rc = allow_smem(d2d_step_warp_kernel<PLE2, false, WPB, true, SPEC, false>, smem);
    if (!rc) rc = allow_smem(d2d_step_warp_kernel<PLE2, true, WPB, false, SPEC, false>, smem);
    if (!rc) rc = allow_smem(d2d_step_warp_kernel<PLE2, true, WPB, true, SPEC, false>, smem);
    if (!rc) rc = allow_smem(d2d_step_warp_kernel<PLE2, false, WPB, false, SPEC, true>, smem);
    if (!rc) rc = allow_smem(d2d_step_warp_kernel<PLE2, true, WPB, false, SPEC, true>, smem);
    return rc;
}

template <typename K>
int plan_geometry(d2d_handle *h, K kernel, int block, size_t smem, int envs_per_block) {
    int rc = allow_smem(kernel, smem);
    if (rc) return rc;
    int occ = 0;
    D2D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem));
    if (occ < 1) return fail(D2D_ERR_UNSUPPORTED, "step kernel does not fit on an SM for this configuration");
    const int64_t need = (h->cfg.num_envs + envs_per_block - 1) / envs_per_block;
    const int64_t resident = (int64_t)h->num_sms * occ;
    h->grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, resident));
    h->block = block;
    h->smem = (int)smem;
    h->envs_per_block = envs_per_block;
    return D2D_OK;
}

// every (FULL, EXACT) instantiation d2d_step may launch for this handle needs the dynamic shared-memory opt-in
template <bool PLE2, int LPT, int BT>
int plan_dense(d2d_handle *h, size_t smem) {
    int rc = allow_smem(d2d_step_dense_kernel<PLE2, LPT, BT, false, false>, smem);
    if (!rc) rc = allow_smem(d2d_step_dense_kernel<PLE2, LPT, BT, false, true>, smem);
    if (!rc) rc = allow_smem(d2d_step_dense_kernel<PLE2, LPT, BT, true, true>, smem);
    if (!rc) rc = plan_geometry(h, d2d_step_dense_kernel<PLE2, LPT, BT, true, false>, BT, smem, 1);
    return rc;
}

template <int WPB>
int plan_warp(d2d_handle *h, size_t smem) {
    int rc;
    if (h->spec) {          // the reference's default EnvConfig shape: counts and division magics are immediates
        rc = allow_smem_warp<true, WPB, true>(smem);
        if (!rc) rc = plan_geometry(h, d2d_step_warp_kernel<true, false, WPB, true, true, false>, WPB * 32, smem, WPB);
    } else if (h->ple2) {
        rc = allow_smem_warp<true, WPB, false>(smem);
        if (!rc) rc = plan_geometry(h, d2d_step_warp_kernel<true, false, WPB, true, false, false>, WPB * 32, smem, WPB);
    } else {
        rc = allow_smem_warp<false, WPB, false>(smem);
        if (!rc) rc = plan_geometry(h, d2d_step_warp_kernel<false, false, WPB, true, false, false>, WPB * 32, smem, WPB);
    }
    return rc;
}

}  // namespace

D2D_API int d2d_abi_version(void) { return D2D_ABI_VERSION; }

D2D_API const char *d2d_last_error(void) { return g_last_error.c_str(); }

D2D_API int d2d_create(const d2d_config_t *cfg, const d2d_link_t *links, d2d_handle_t **out) {
    if (!cfg || !links || !out) return fail(D2D_ERR_INVALID_ARG, "d2d_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != D2D_ABI_VERSION) return fail(D2D_ERR_INVALID_ARG, "d2d_create: abi_version mismatch");
    if (cfg->num_envs < 1) return fail(D2D_ERR_INVALID_ARG, "d2d_create: num_envs must be >= 1");
    if (cfg->num_rbs < 1 || cfg->num_cues < 0 || cfg->num_due_pairs < 0 || cfg->num_cues + cfg->num_due_pairs < 1)
        return fail(D2D_ERR_INVALID_ARG, "d2d_create: need num_rbs >= 1 and at least one link");
    if (cfg->num_downlinks != 0 && cfg->num_downlinks != cfg->num_cues)
        return fail(D2D_ERR_INVALID_ARG, "d2d_create: num_downlinks must be 0 or num_cues (one 'mbs:cueXX' link per CUE)");
    if (cfg->num_downlinks && (cfg->n_pwr_mbs < 1 || cfg->n_pwr_mbs > D2D_MAX_PWR_LEVELS))
        return fail(D2D_ERR_UNSUPPORTED, "d2d_create: n_pwr_mbs must be in [1, 128]");
    if (cfg->n_pwr_cue < 1 || cfg->n_pwr_due < 1 || cfg->n_pwr_cue > D2D_MAX_PWR_LEVELS || cfg->n_pwr_due > D2D_MAX_PWR_LEVELS)
        return fail(D2D_ERR_UNSUPPORTED, "d2d_create: power levels per link must be in [1, 128]");
    if (cfg->num_rbs > 32767) return fail(D2D_ERR_UNSUPPORTED, "d2d_create: num_rbs must be <= 32767");
    if (cfg->path_loss_model != D2D_PL_LOG_DISTANCE && cfg->path_loss_model != D2D_PL_FREE_SPACE &&
        cfg->path_loss_model != D2D_PL_COST_HATA && cfg->path_loss_model != D2D_PL_SHADOWING)
        return fail(D2D_ERR_UNSUPPORTED, "d2d_create: unsupported path_loss_model (LogDistance / FreeSpace / CostHata / Shadowing only)");
    if (cfg->path_loss_model == D2D_PL_SHADOWING && (!(cfg->shadow_chi_dB >= 0.0) || !(cfg->shadow_d0_m >= 0.0)))
        return fail(D2D_ERR_INVALID_ARG, "d2d_create: shadow_chi_dB and shadow_d0_m must be >= 0");
    if (cfg->obs_fn != D2D_OBS_LINEAR) return fail(D2D_ERR_UNSUPPORTED, "d2d_create: unsupported obs_fn (LinearObsFunction only)");
    if (cfg->reward_fn != D2D_REWARD_SYSTEM_CAPACITY && cfg->reward_fn != D2D_REWARD_SHANNON &&
        cfg->reward_fn != D2D_REWARD_CUE_SINR_SHANNON)
        return fail(D2D_ERR_UNSUPPORTED, "d2d_create: unsupported reward_fn (SystemCapacity / Shannon / CueSinrShannon only)");
    if (!(cfg->carrier_freq_GHz > 0.0)) return fail(D2D_ERR_INVALID_ARG, "d2d_create: carrier_freq_GHz must be > 0");
    const double ple = cfg->path_loss_model == D2D_PL_FREE_SPACE ? 2.0 : cfg->ple;
    if (!(ple > 0.0) || ple > 8.0) return fail(D2D_ERR_INVALID_ARG, "d2d_create: ple must be in (0, 8]");
    if (cfg->episode_length < 1 || cfg->episode_length > 255)
        return fail(D2D_ERR_UNSUPPORTED, "d2d_create: episode_length must be in [1, 255]");

    d2d_handle *h = new (std::nothrow) d2d_handle();
    if (!h) return fail(D2D_ERR_INVALID_ARG, "d2d_create: out of host memory");
    h->cfg = *cfg;
    h->N = cfg->num_cues + cfg->num_due_pairs + cfg->num_downlinks;
    h->V = 1 + cfg->num_cues + 2 * cfg->num_due_pairs;
    h->ple = ple;
    h->ple2 = ple == 2.0;
    if (h->N > 65535) { delete h; return fail(D2D_ERR_UNSUPPORTED, "d2d_create: at most 65535 links per env"); }
    // path_loss.py:28-39
    h->K_dB = 10.0 * ple * std::log10(cfg->carrier_freq_GHz * 1e9) + 10.0 * ple * std::log10((4.0 * M_PI) / kSpeedOfLight);

    auto bail = [&](int code) { d2d_destroy(h); return code; };
    cudaError_t e = cudaSetDevice(cfg->cuda_device);
    if (e != cudaSuccess) return bail(fail(D2D_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)));
    cudaDeviceProp prop{};
    e = cudaGetDeviceProperties(&prop, cfg->cuda_device);
    if (e != cudaSuccess) return bail(fail(D2D_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e)));
    if (prop.major != 10)
        return bail(fail(D2D_ERR_UNSUPPORTED, "libd2d_b200 is built for sm_100a (Blackwell B200) only; device is sm_" +
                                                  std::to_string(prop.major) + std::to_string(prop.minor)));
    h->num_sms = prop.multiProcessorCount;

    // fold the link-budget constants in fp64, store fp32 + fp64 copies
    std::vector<D2DLinkA> A(h->N);
    std::vector<D2DLinkB> B(h->N);
    std::vector<D2DLinkD> Dv(h->N);
    std::vector<int32_t> meta(h->N);
    for (int j = 0; j < h->N; ++j) {
        const d2d_link_t &L = links[j];
        const bool cue = j < cfg->num_cues, down = j >= cfg->num_cues + cfg->num_due_pairs;
        const int expect = cue ? D2D_LINK_UPLINK : down ? D2D_LINK_DOWNLINK : D2D_LINK_SIDELINK;
        if (L.link_type != expect)
            return bail(fail(D2D_ERR_UNSUPPORTED, "d2d_create: link " + std::to_string(j) +
                                                      " has an unsupported link_type (CUE uplinks, DUE sidelinks, then MBS downlinks)"));
        // the path-loss constant K belongs to the RECEIVER (CostHata: A(h_tx, h_rx) - 3 B; log-distance: one K for all), so it
        // is folded into the victim's constants and the radiated weights w carry none: I_true = 10^(-K_rx/10) sum w_k g_k
        const double Kj = cfg->path_loss_model == D2D_PL_COST_HATA ? L.path_loss_const_dB : h->K_dB;
        A[j].tx_lin0 = (float)std::pow(10.0, L.tx_eirp_offset_dB / 10.0);
        const double snr0 = L.tx_eirp_offset_dB + L.rx_offset_dB - Kj - L.rx_noise_dBm;
        A[j].a_lin = (float)std::pow(10.0, snr0 / 10.0);
        A[j].inv_noise = (float)std::pow(10.0, -(Kj + L.rx_noise_dBm) / 10.0);
        A[j].snr0_dB = (float)snr0;
        B[j].sens_dBm = (float)L.rx_sensitivity_dBm;
        B[j].bw_MHz = (float)(1e-6 * (L.tx_rb_bandwidth_kHz * 1000.0));
        // envs/d2d_env.py:80-91: uplink cue j -> mbs; sidelink due pair; downlink mbs -> cue (j - C - D)
        B[j].tx_dev = cue ? 1 + j : down ? 0 : 1 + cfg->num_cues + 2 * (j - cfg->num_cues);
        B[j].rx_dev = cue ? 0 : down ? 1 + (j - cfg->num_cues - cfg->num_due_pairs) : B[j].tx_dev + 1;
        meta[j] = (cue ? cfg->n_pwr_cue : down ? cfg->n_pwr_mbs : cfg->n_pwr_due) | ((!cue && !down) ? 1 << 16 : 0);
        Dv[j].a_lin = std::pow(10.0, snr0 / 10.0);
        Dv[j].t_lin = std::pow(10.0, L.tx_eirp_offset_dB / 10.0);
        Dv[j].inv_noise = std::pow(10.0, -(Kj + L.rx_noise_dBm) / 10.0);
        Dv[j].bw_MHz = 1e-6 * (L.tx_rb_bandwidth_kHz * 1000.0);
    }
    // one set of constants per link type (no per-device overrides)?  Then the default-shape kernel reads them from
    // the constant bank instead of a shared-memory table
    h->uniform = cfg->num_downlinks == 0;
    for (int j = 0; j < h->N && !cfg->num_downlinks; ++j) {
        const int j0 = j < cfg->num_cues ? 0 : cfg->num_cues;
        if (std::memcmp(&A[j], &A[j0], sizeof(D2DLinkA)) != 0 || std::memcmp(&Dv[j], &Dv[j0], sizeof(D2DLinkD)) != 0 ||
            B[j].sens_dBm != B[j0].sens_dBm || B[j].bw_MHz != B[j0].bw_MHz)
            h->uniform = false;
    }
    if (cfg->num_cues > 0) h->ud_cue = Dv[0];
    if (cfg->num_due_pairs > 0) h->ud_due = Dv[cfg->num_cues];
    if (cfg->num_cues > 0) { h->u_cue = A[0]; h->us_cue[0] = B[0].sens_dBm; h->us_cue[1] = B[0].bw_MHz; }
    if (cfg->num_due_pairs > 0) {
        h->u_due = A[cfg->num_cues]; h->us_due[0] = B[cfg->num_cues].sens_dBm; h->us_due[1] = B[cfg->num_cues].bw_MHz;
    }
    float pwr[D2D_MAX_PWR_LEVELS];
    double pwr_d[D2D_MAX_PWR_LEVELS];
    for (int p = 0; p < D2D_MAX_PWR_LEVELS; ++p) {   // conversion.py:4-13
        pwr_d[p] = std::pow(10.0, p / 10.0);
        pwr[p] = (float)pwr_d[p];
    }
    {   // the warp kernel's fp64 pass reads the fp64 table from the constant bank (per device: set at every create)
        cudaError_t err__ = cudaMemcpyToSymbol(d2d_pwr_lin_c, pwr_d, sizeof(pwr_d));
        if (err__ != cudaSuccess) return bail(fail(D2D_ERR_CUDA, std::string("cudaMemcpyToSymbol(d2d_pwr_lin_c): ") + cudaGetErrorString(err__)));
    }

#define D2D_CUDA_BAIL(call)                                                                              \
    do {                                                                                                 \
        cudaError_t err__ = (call);                                                                      \
        if (err__ != cudaSuccess)                                                                        \
            return bail(fail(D2D_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__)));       \
    } while (0)
    D2D_CUDA_BAIL(cudaMalloc(&h->dA, sizeof(D2DLinkA) * h->N));
    D2D_CUDA_BAIL(cudaMalloc(&h->dB, sizeof(D2DLinkB) * h->N));
    D2D_CUDA_BAIL(cudaMalloc(&h->dD, sizeof(D2DLinkD) * h->N));
    D2D_CUDA_BAIL(cudaMalloc(&h->dMeta, sizeof(int32_t) * h->N));
    D2D_CUDA_BAIL(cudaMemcpy(h->dMeta, meta.data(), sizeof(int32_t) * h->N, cudaMemcpyHostToDevice));
    D2D_CUDA_BAIL(cudaMalloc(&h->dPwr, sizeof(pwr)));
    D2D_CUDA_BAIL(cudaMalloc(&h->dPwrD, sizeof(pwr_d)));
    D2D_CUDA_BAIL(cudaMemcpy(h->dPwrD, pwr_d, sizeof(pwr_d), cudaMemcpyHostToDevice));
    D2D_CUDA_BAIL(cudaMemcpy(h->dA, A.data(), sizeof(D2DLinkA) * h->N, cudaMemcpyHostToDevice));
    D2D_CUDA_BAIL(cudaMemcpy(h->dB, B.data(), sizeof(D2DLinkB) * h->N, cudaMemcpyHostToDevice));
    D2D_CUDA_BAIL(cudaMemcpy(h->dD, Dv.data(), sizeof(D2DLinkD) * h->N, cudaMemcpyHostToDevice));
    D2D_CUDA_BAIL(cudaMemcpy(h->dPwr, pwr, sizeof(pwr), cudaMemcpyHostToDevice));
#undef D2D_CUDA_BAIL

    // warp kernel: one lane slot per CUE and per DUE pair, one shared-memory bin per RB
    // downlink links and ShadowingPathLoss run on the general-topology kernel
    const bool general = cfg->num_downlinks != 0 || cfg->path_loss_model == D2D_PL_SHADOWING;
    h->use_warp = cfg->num_cues <= 32 && cfg->num_due_pairs <= 32 && cfg->num_rbs <= 64 && !general;
    if (const char *c = std::getenv("D2D_B200_CHUNK")) h->chunk_override = std::atoll(c);
    const char *pdl = std::getenv("D2D_B200_PDL");
    h->pdl = !(pdl && std::strcmp(pdl, "0") == 0);
    int rc;
    if (h->use_warp) {
        // launch shape by batch size (d2d_step_warp.cuh): about one wave of envs -> 4-warp blocks, many waves -> 8-warp blocks
        h->wpb = cfg->num_envs >= 65536 ? 8 : cfg->num_envs <= D2D_LATENCY_ENVS ? 2 : D2D_WPB_MID;
        if (const char *w = std::getenv("D2D_B200_WPB")) { const int v = std::atoi(w); h->wpb = v == 8 ? 8 : v == 2 ? 2 : D2D_WPB_MID; }
        h->spec = h->ple2 && h->uniform && cfg->path_loss_model != D2D_PL_COST_HATA && cfg->num_rbs == 25 && cfg->num_cues == 25 && cfg->num_due_pairs == 25 && cfg->n_pwr_cue == 24 &&
                  cfg->n_pwr_due == 21;
        if (const char *sp = std::getenv("D2D_B200_SPEC")) h->spec = h->spec && std::atoi(sp) != 0;   // tests: force the generic shape
        const size_t smem = d2d_warp_smem_bytes(cfg->num_rbs, h->wpb);
        rc = h->wpb == 8 ? plan_warp<8>(h, smem) : h->wpb == 2 ? plan_warp<2>(h, smem) : plan_warp<D2D_WPB_MID>(h, smem);
    } else {
        // <= 1024 links: the binned one-barrier kernel (d2d_step_dense.cuh), LPT links per thread; when its double-buffered
        // bins do not fit (many RBs) the sorting block kernel; beyond 1024 links / other topologies: everything staged in shared memory
        h->lpt = (h->N <= D2D_BLOCK_THREADS * D2D_BLOCK_MAX_LPT && !general)
                     ? (h->N + D2D_BLOCK_THREADS - 1) / D2D_BLOCK_THREADS : 0;      // downlinks: the general-topology kernel
        h->bin_cap = d2d_dense_bin_cap(h->N, cfg->num_rbs);
        const char *dn = std::getenv("D2D_B200_DENSE");
        bool dense = h->lpt > 0 && d2d_dense_layout(h->N, cfg->num_rbs, h->bin_cap).total <= 72 * 1024 && !(dn && std::atoi(dn) == 0);
        if (dense) {
            // threads per block / links per thread: the shape with the fewest (warp, slot) bodies per env - every warp runs the
            // straight-line code of each of its slots whether or not all 32 lanes hold a link (N = 600: 10 warps x 2 slots,
            // all but one full, instead of 8 warps x 3 with the third slot three-quarters empty)
            int bt = ((h->N + 319) / 320) * 10 < ((h->N + 255) / 256) * 8 && (h->N + 319) / 320 <= 3 ? 320 : 256;
            if (dn && std::atoi(dn) >= 64) bt = std::atoi(dn);
            h->dense_bt = bt;
            h->lpt = (h->N + bt - 1) / bt;
        }
        const size_t smem = dense ? d2d_dense_layout(h->N, cfg->num_rbs, h->bin_cap).total
                          : h->lpt ? d2d_block2_smem_bytes(h->N, cfg->num_rbs) : d2d_block_smem_bytes(h->N, cfg->num_rbs);
        if (smem > 227 * 1024) return bail(fail(D2D_ERR_UNSUPPORTED, "d2d_create: too many links / RBs for one SM's shared memory"));
#define D2D_PLAN_BLOCK(LPT_) (h->ple2 ? plan_geometry(h, d2d_step_block_kernel<true, LPT_>, D2D_BLOCK_THREADS, smem, 1) \
                                      : plan_geometry(h, d2d_step_block_kernel<false, LPT_>, D2D_BLOCK_THREADS, smem, 1))
#define D2D_PLAN_DENSE(LPT_, BT_) (h->ple2 ? plan_dense<true, LPT_, BT_>(h, smem) \
                                           : plan_dense<false, LPT_, BT_>(h, smem))
        if (dense) {
            rc = fail(D2D_ERR_UNSUPPORTED, "d2d_create: no dense kernel instantiation for this shape");
            D2D_DENSE_SHAPES(D2D_DENSE_PLAN_CASE)
        } else switch (h->lpt) {
            case 1: rc = D2D_PLAN_BLOCK(1); break;
            case 2: rc = D2D_PLAN_BLOCK(2); break;
            case 3: rc = D2D_PLAN_BLOCK(3); break;
            case 4: rc = D2D_PLAN_BLOCK(4); break;
            default:
                rc = h->ple2 ? plan_geometry(h, d2d_step_block_generic_kernel<true>, D2D_BLOCK_THREADS, smem, 1)
                             : plan_geometry(h, d2d_step_block_generic_kernel<false>, D2D_BLOCK_THREADS, smem, 1);
        }
#undef D2D_PLAN_DENSE
#undef D2D_PLAN_BLOCK
    }
    if (rc != D2D_OK) return bail(rc);
    if (const char *gs = std::getenv("D2D_B200_GRID"))      // tests: few blocks, so every block steps many envs
        if (std::atoi(gs) > 0) h->grid = std::min(h->grid, std::atoi(gs));
    *out = h;
    return D2D_OK;
}

D2D_API int d2d_destroy(d2d_handle_t *h) {
    if (!h) return D2D_OK;
    cudaFree(h->dA); cudaFree(h->dB); cudaFree(h->dD); cudaFree(h->dMeta); cudaFree(h->dPwr); cudaFree(h->dPwrD); cudaFree(h->stage_pos);
    for (auto &slot : h->stage2)
        for (void *p : slot) cudaFree(p);
    if (h->pipe_ready) {
        cudaStreamDestroy(h->s_in); cudaStreamDestroy(h->s_out);
        for (int s = 0; s < 2; ++s) { cudaEventDestroy(h->ev_in[s]); cudaEventDestroy(h->ev_kernel[s]); cudaEventDestroy(h->ev_out[s]); }
    }
    delete h;
    return D2D_OK;
}

D2D_API int d2d_state_bytes(const d2d_handle_t *h, size_t *pos_bytes, size_t *step_bytes, size_t *stats_bytes) {
    if (!h) return fail(D2D_ERR_INVALID_ARG, "d2d_state_bytes: null handle");
    if (pos_bytes) *pos_bytes = (size_t)h->cfg.num_envs * h->V * 2 * sizeof(float);
    if (step_bytes) *step_bytes = (size_t)h->cfg.num_envs;
    if (stats_bytes) *stats_bytes = (size_t)D2D_STATS_REPLICAS * 8 * sizeof(double);
    return D2D_OK;
}

D2D_API int d2d_bind_state(d2d_handle_t *h, float *positions, uint8_t *step_count, double *stats) {
    if (!h || !positions) return fail(D2D_ERR_INVALID_ARG, "d2d_bind_state: positions buffer is required");
    if ((uintptr_t)positions % 16) return fail(D2D_ERR_INVALID_ARG, "d2d_bind_state: positions must be 16-byte aligned");
    h->pos = positions;
    h->step_count = step_count;
    h->stats = stats;
    return D2D_OK;
}

D2D_API int d2d_bind_positions_f64(d2d_handle_t *h, double *positions_f64) {
    if (!h) return fail(D2D_ERR_INVALID_ARG, "d2d_bind_positions_f64: null handle");
    if (positions_f64 && ((uintptr_t)positions_f64 % 16))
        return fail(D2D_ERR_INVALID_ARG, "d2d_bind_positions_f64: buffer must be 16-byte aligned");
    h->pos64 = positions_f64;
    return D2D_OK;
}

D2D_API int d2d_set_positions(d2d_handle_t *h, const double *src, int src_on_device, int64_t first_env, int64_t count,
                              void *stream) {
    if (!h || !src) return fail(D2D_ERR_INVALID_ARG, "d2d_set_positions: null argument");
    if (!h->pos) return fail(D2D_ERR_STATE, "d2d_set_positions: call d2d_bind_state first");
    if (first_env < 0 || count < 0 || first_env + count > h->cfg.num_envs)
        return fail(D2D_ERR_INVALID_ARG, "d2d_set_positions: env range out of bounds");
    if (count == 0) return D2D_OK;
    int rc = ensure_device(h);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const double *dsrc = src;
    if (!src_on_device) {
        if (h->stage_pos_envs < count) {
            cudaFree(h->stage_pos);
            h->stage_pos = nullptr; h->stage_pos_envs = 0;
            D2D_CUDA(cudaMalloc(&h->stage_pos, sizeof(double) * 2 * h->V * count));
            h->stage_pos_envs = count;
        }
        D2D_CUDA(cudaMemcpyAsync(h->stage_pos, src, sizeof(double) * 2 * h->V * count, cudaMemcpyHostToDevice, st));
        dsrc = h->stage_pos;
    }
    const int64_t total = count * h->V;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->num_sms * 8);
    d2d_set_positions_kernel<<<grid, 256, 0, st>>>(dsrc, h->pos + first_env * h->V * 2,
                                                    h->pos64 ? h->pos64 + first_env * h->V * 2 : nullptr, count, h->V);
    D2D_CUDA(cudaGetLastError());
    ++h->launches;
    if (!src_on_device) D2D_CUDA(cudaStreamSynchronize(st));
    return D2D_OK;
}

D2D_API int d2d_reset(d2d_handle_t *h, uint64_t seed, uint64_t first_global_env, const uint8_t *env_mask, void *stream) {
    if (!h) return fail(D2D_ERR_INVALID_ARG, "d2d_reset: null handle");
    if (!h->pos) return fail(D2D_ERR_STATE, "d2d_reset: call d2d_bind_state first");
    int rc = ensure_device(h);
    if (rc) return rc;
    const int64_t total = h->cfg.num_envs * h->N;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->num_sms * 16);
    d2d_reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h->pos, h->pos64, h->step_count, env_mask, h->cfg.num_envs, h->cfg.num_cues,
                                                             h->cfg.num_due_pairs, (float)h->cfg.cell_radius_m,
                                                             (float)h->cfg.d2d_radius_m, seed, first_global_env);
    D2D_CUDA(cudaGetLastError());
    ++h->launches;
    return D2D_OK;
}

namespace {

// One launch (per chunk of envs) that makes T consecutive steps: T == 1 is d2d_step; T > 1 needs the warp kernel.
int step_launch(d2d_handle *h, const d2d_step_io_t *io, int T, void *stream) {
    D2DParams P = make_params(h, io);
    P.T = T;
    P.t_stride = h->cfg.num_envs;
    cudaStream_t st = (cudaStream_t)stream;
    const bool many = T > 1;
    h->rng_calls += (uint64_t)T;
    // the kernels index with 32 bits: batches beyond 2^31 / (T max(6N, 2V)) envs (> 7 million default envs) go in chunks
    int64_t chunk = std::max<int64_t>(1, (int64_t)0x7fffffff / std::max(6 * h->N, 2 * h->V));
    if (many) chunk = h->cfg.num_envs;        // d2d_step_many checked that T slices fit 32-bit indices
    if (h->chunk_override > 0 && !many) chunk = std::min(chunk, h->chunk_override);
    for (int64_t e0 = 0; e0 < h->cfg.num_envs; e0 += chunk) {
        const int64_t n = std::min<int64_t>(chunk, h->cfg.num_envs - e0);
        if (e0 > 0) {
            const int64_t dl = chunk * h->N;
            P.first_global_env += (uint64_t)chunk;
            P.actions += dl; P.pos += chunk * h->V * 2;
            if (P.pos64) P.pos64 += chunk * h->V * 2;
            if (P.step_count) P.step_count += chunk;
            if (P.obs) P.obs += dl * 6;
            if (P.cap) P.cap += dl;
            if (P.reward) P.reward += chunk;
            if (P.done) P.done += chunk;
            if (P.rate) P.rate += dl;
            if (P.rb_out) P.rb_out += dl;
            if (P.pwr_out) P.pwr_out += dl;
        }
        P.num_envs = n;
#ifdef D2D_TIMELINE
        P.tl_slot = (int32_t)(h->launches % D2D_TL_SLOTS);
#endif
        const int grid = (int)std::min<int64_t>(h->grid, (n + h->envs_per_block - 1) / h->envs_per_block);
        if (h->use_warp) { const int64_t warps = (int64_t)grid * h->wpb; P.envs_per_warp = (uint32_t)((n + warps - 1) / warps); }
        const bool exact = h->pos64 != nullptr;
        cudaError_t err;
        // FULL: exactly the core outputs were passed, so the kernel tests no output pointer on its hot path
        const bool full = io->obs && io->capacity_mbps && io->reward && io->done && !io->rate_bps && !io->rb && !io->tx_pwr_dBm &&
                          h->step_count;
#define D2D_LAUNCH_WARP(PLE2_, EXACT_, WPB_, FULL_, SPEC_, MANY_) \
    launch_step(d2d_step_warp_kernel<PLE2_, EXACT_, WPB_, FULL_, SPEC_, MANY_>, grid, WPB_ * 32, h->smem, st, P, h->pdl)
#define D2D_PICK_FULL(PLE2_, EXACT_, WPB_, SPEC_)                                               \
    (many ? D2D_LAUNCH_WARP(PLE2_, EXACT_, WPB_, false, SPEC_, true)                             \
          : full ? D2D_LAUNCH_WARP(PLE2_, EXACT_, WPB_, true, SPEC_, false) : D2D_LAUNCH_WARP(PLE2_, EXACT_, WPB_, false, SPEC_, false))
#define D2D_PICK_EXACT(PLE2_, WPB_, SPEC_) (exact ? D2D_PICK_FULL(PLE2_, true, WPB_, SPEC_) : D2D_PICK_FULL(PLE2_, false, WPB_, SPEC_))
#define D2D_PICK_SHAPE(WPB_) \
    (h->spec ? D2D_PICK_EXACT(true, WPB_, true) : h->ple2 ? D2D_PICK_EXACT(true, WPB_, false) : D2D_PICK_EXACT(false, WPB_, false))
        if (h->use_warp && h->wpb == 8) {
            err = D2D_PICK_SHAPE(8);
        } else if (h->use_warp && h->wpb == 2) {
            err = D2D_PICK_SHAPE(2);
        } else if (h->use_warp) {
            err = D2D_PICK_SHAPE(D2D_WPB_MID);
        } else {
#define D2D_LAUNCH_BLOCK(LPT_) (h->ple2 ? launch_step(d2d_step_block_kernel<true, LPT_>, grid, h->block, h->smem, st, P, h->pdl) \
                                        : launch_step(d2d_step_block_kernel<false, LPT_>, grid, h->block, h->smem, st, P, h->pdl))
#define D2D_LAUNCH_DENSE4(PLE2_, LPT_, BT_, FULL_, EXACT_) \
    launch_step(d2d_step_dense_kernel<PLE2_, LPT_, BT_, FULL_, EXACT_>, grid, h->block, h->smem, st, P, h->pdl)
#define D2D_LAUNCH_DENSE3(PLE2_, LPT_, BT_)                                                                                    \
    (full && h->uniform ? (exact ? D2D_LAUNCH_DENSE4(PLE2_, LPT_, BT_, true, true) : D2D_LAUNCH_DENSE4(PLE2_, LPT_, BT_, true, false)) \
                        : (exact ? D2D_LAUNCH_DENSE4(PLE2_, LPT_, BT_, false, true) : D2D_LAUNCH_DENSE4(PLE2_, LPT_, BT_, false, false)))
#define D2D_LAUNCH_DENSE(LPT_, BT_) (h->ple2 ? D2D_LAUNCH_DENSE3(true, LPT_, BT_) : D2D_LAUNCH_DENSE3(false, LPT_, BT_))
            if (h->dense_bt) {
                err = cudaErrorInvalidValue;
                D2D_DENSE_SHAPES(D2D_DENSE_LAUNCH_CASE)
            } else
            switch (h->lpt) {
                case 1: err = D2D_LAUNCH_BLOCK(1); break;
                case 2: err = D2D_LAUNCH_BLOCK(2); break;
                case 3: err = D2D_LAUNCH_BLOCK(3); break;
                case 4: err = D2D_LAUNCH_BLOCK(4); break;
                default:
                    err = h->ple2 ? launch_step(d2d_step_block_generic_kernel<true>, grid, h->block, h->smem, st, P, h->pdl)
                                  : launch_step(d2d_step_block_generic_kernel<false>, grid, h->block, h->smem, st, P, h->pdl);
            }
#undef D2D_LAUNCH_BLOCK
#undef D2D_LAUNCH_DENSE
#undef D2D_LAUNCH_DENSE3
#undef D2D_LAUNCH_DENSE4
        }
        if (err != cudaSuccess) return fail(D2D_ERR_CUDA, std::string("step kernel launch: ") + cudaGetErrorString(err));
        ++h->launches;
    }
    D2D_CUDA(cudaGetLastError());
    // per-agent rewards (SHANNON / CUE_SINR_SHANNON always; SYSTEM_CAPACITY when the caller asked for the broadcast): one more
    // small kernel over the T x E env-steps just written
    if (h->cfg.reward_fn != D2D_REWARD_SYSTEM_CAPACITY || io->agent_reward) {
        if (!io->obs || !io->reward)
            return fail(D2D_ERR_INVALID_ARG, "per-agent rewards need the obs and reward outputs");
        const int64_t total = (int64_t)T * h->cfg.num_envs;
        const bool warp_team = h->N <= 64;
        const int teams = warp_team ? 8 : 1;
        const int grid = (int)std::min<int64_t>((total + teams - 1) / teams, (int64_t)h->num_sms * 8);
        const size_t smem = (size_t)teams * h->cfg.num_rbs * sizeof(uint32_t);
        double *stats = h->cfg.reward_fn != D2D_REWARD_SYSTEM_CAPACITY ? h->stats : nullptr;
        if (warp_team)
            d2d_agent_reward_kernel<32><<<grid, 256, smem, st>>>(io->actions, io->obs, io->agent_reward, io->reward, stats, total, h->N,
                                                                h->dMeta, h->cfg.num_rbs, h->cfg.reward_fn, (float)h->cfg.reward_param);
        else
            d2d_agent_reward_kernel<256><<<grid, 256, smem, st>>>(io->actions, io->obs, io->agent_reward, io->reward, stats, total, h->N,
                                                                 h->dMeta, h->cfg.num_rbs, h->cfg.reward_fn, (float)h->cfg.reward_param);
        D2D_CUDA(cudaGetLastError());
        ++h->launches;
    }
    return D2D_OK;
}

}  // namespace

D2D_API int d2d_step(d2d_handle_t *h, const d2d_step_io_t *io, void *stream) {
    if (!h || !io || !io->actions) return fail(D2D_ERR_INVALID_ARG, "d2d_step: handle, io and io->actions are required");
    if (!h->pos) return fail(D2D_ERR_STATE, "d2d_step: call d2d_bind_state first");
    if (io->obs && ((uintptr_t)io->obs % 8)) return fail(D2D_ERR_INVALID_ARG, "d2d_step: obs must be 8-byte aligned");
    return step_launch(h, io, 1, stream);
}

D2D_API int d2d_step_many(d2d_handle_t *h, const d2d_step_io_t *io, int32_t num_steps, void *stream) {
    if (!h || !io || !io->actions) return fail(D2D_ERR_INVALID_ARG, "d2d_step_many: handle, io and io->actions are required");
    if (!h->pos) return fail(D2D_ERR_STATE, "d2d_step_many: call d2d_bind_state first");
    if (num_steps < 1) return fail(D2D_ERR_INVALID_ARG, "d2d_step_many: num_steps must be >= 1");
    if (io->obs && ((uintptr_t)io->obs % 8)) return fail(D2D_ERR_INVALID_ARG, "d2d_step_many: obs must be 8-byte aligned");
    const int64_t E = h->cfg.num_envs, per_env = std::max(6 * h->N, 2 * h->V);
    // fused launches of as many steps as 32-bit indices allow (all of them unless T E N is astronomically large);
    // configurations served by the block kernel take one launch per step
    int64_t fuse = h->use_warp ? std::min<int64_t>(num_steps, (int64_t)0x7fffffff / (E * per_env)) : 1;
    if (fuse < 1) fuse = 1;
    d2d_step_io_t cur = *io;
    for (int64_t t0 = 0; t0 < num_steps; t0 += fuse) {
        const int T = (int)std::min<int64_t>(fuse, num_steps - t0);
        int rc = step_launch(h, &cur, T, stream);
        if (rc) return rc;
        const int64_t dl = (int64_t)T * E * h->N, de = (int64_t)T * E;
        cur.actions += dl;
        if (cur.obs) cur.obs += dl * 6;
        if (cur.capacity_mbps) cur.capacity_mbps += dl;
        if (cur.reward) cur.reward += de;
        if (cur.done) cur.done += de;
        if (cur.rate_bps) cur.rate_bps += dl;
        if (cur.rb) cur.rb += dl;
        if (cur.tx_pwr_dBm) cur.tx_pwr_dBm += dl;
        if (cur.agent_reward) cur.agent_reward += dl;
    }
    return D2D_OK;
}

// ---- host-buffer steps: a two-slot pipeline (copy-in stream -> caller's stream for the kernel -> copy-out stream) ----
namespace {

int host_pipeline_init(d2d_handle *h) {
    if (h->pipe_ready) return D2D_OK;
    D2D_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    D2D_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int s = 0; s < 2; ++s) {
        D2D_CUDA(cudaEventCreateWithFlags(&h->ev_in[s], cudaEventDisableTiming));
        D2D_CUDA(cudaEventCreateWithFlags(&h->ev_kernel[s], cudaEventDisableTiming));
        D2D_CUDA(cudaEventCreateWithFlags(&h->ev_out[s], cudaEventDisableTiming));
    }
    h->pipe_ready = true;
    return D2D_OK;
}

}  // namespace

D2D_API int d2d_step_host_async(d2d_handle_t *h, const d2d_step_io_t *hio, int slot, void *stream) {
    if (!h || !hio || !hio->actions) return fail(D2D_ERR_INVALID_ARG, "d2d_step_host_async: handle, io and io->actions are required");
    if (slot < 0 || slot > 1) return fail(D2D_ERR_INVALID_ARG, "d2d_step_host_async: slot must be 0 or 1");
    if (!h->pos) return fail(D2D_ERR_STATE, "d2d_step_host_async: call d2d_bind_state first");
    int rc = ensure_device(h);
    if (rc) return rc;
    rc = host_pipeline_init(h);
    if (rc) return rc;
    const size_t E = (size_t)h->cfg.num_envs, EN = E * h->N;
    const size_t bytes[9] = {EN * 4, EN * 24, EN * 4, E * 4, E, EN * 4, EN * 2, EN * 2, EN * 4};
    void *host[9] = {(void *)hio->actions, hio->obs, hio->capacity_mbps, hio->reward, hio->done, hio->rate_bps, hio->rb, hio->tx_pwr_dBm,
                     hio->agent_reward};
    void **stage = h->stage2[slot];
    for (int i = 0; i < 9; ++i)
        if (host[i] && !stage[i]) D2D_CUDA(cudaMalloc(&stage[i], bytes[i]));
    cudaStream_t st = (cudaStream_t)stream;
    // copy-in: this slot's action staging is free once the kernel that last read it has run
    if (h->slot_used[slot]) D2D_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_kernel[slot], 0));
    D2D_CUDA(cudaMemcpyAsync(stage[0], host[0], bytes[0], cudaMemcpyHostToDevice, h->s_in));
    D2D_CUDA(cudaEventRecord(h->ev_in[slot], h->s_in));
    // kernel on the caller's stream (steps stay ordered there): needs the actions in, and this slot's previous
    // outputs copied out
    D2D_CUDA(cudaStreamWaitEvent(st, h->ev_in[slot], 0));
    if (h->slot_used[slot]) D2D_CUDA(cudaStreamWaitEvent(st, h->ev_out[slot], 0));
    d2d_step_io_t dio{};
    dio.actions = (const int32_t *)stage[0];
    dio.obs = hio->obs ? (float *)stage[1] : nullptr;
    dio.capacity_mbps = hio->capacity_mbps ? (float *)stage[2] : nullptr;
    dio.reward = hio->reward ? (float *)stage[3] : nullptr;
    dio.done = hio->done ? (uint8_t *)stage[4] : nullptr;
    dio.rate_bps = hio->rate_bps ? (float *)stage[5] : nullptr;
    dio.rb = hio->rb ? (int16_t *)stage[6] : nullptr;
    dio.tx_pwr_dBm = hio->tx_pwr_dBm ? (int16_t *)stage[7] : nullptr;
    dio.agent_reward = hio->agent_reward ? (float *)stage[8] : nullptr;
    rc = d2d_step(h, &dio, stream);
    if (rc) return rc;
    D2D_CUDA(cudaEventRecord(h->ev_kernel[slot], st));
    // copy-out
    D2D_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_kernel[slot], 0));
    for (int i = 1; i < 9; ++i)
        if (host[i]) D2D_CUDA(cudaMemcpyAsync(host[i], stage[i], bytes[i], cudaMemcpyDeviceToHost, h->s_out));
    D2D_CUDA(cudaEventRecord(h->ev_out[slot], h->s_out));
    h->slot_used[slot] = true;
    return D2D_OK;
}

D2D_API int d2d_step_host_wait(d2d_handle_t *h, int slot) {
    if (!h || slot < 0 || slot > 1) return fail(D2D_ERR_INVALID_ARG, "d2d_step_host_wait: bad handle or slot");
    if (!h->slot_used[slot]) return D2D_OK;
    D2D_CUDA(cudaEventSynchronize(h->ev_out[slot]));
    return D2D_OK;
}

D2D_API int d2d_step_host(d2d_handle_t *h, const d2d_step_io_t *hio, void *stream) {
    int rc = d2d_step_host_async(h, hio, 0, stream);
    if (rc) return rc;
    return d2d_step_host_wait(h, 0);
}

D2D_API int d2d_per_agent_obs(d2d_handle_t *h, const float *table, float *out, int64_t num_envs, void *stream) {
    if (!h || !table || !out) return fail(D2D_ERR_INVALID_ARG, "d2d_per_agent_obs: null argument");
    if (num_envs < 0 || num_envs > h->cfg.num_envs) return fail(D2D_ERR_INVALID_ARG, "d2d_per_agent_obs: bad num_envs");
    if (num_envs == 0) return D2D_OK;
    if (num_envs * h->N > 0x7fffffffLL) return fail(D2D_ERR_UNSUPPORTED, "d2d_per_agent_obs: num_envs * N exceeds the grid limit");
    d2d_per_agent_obs_kernel<<<(unsigned)(num_envs * h->N), 128, 0, (cudaStream_t)stream>>>(table, out, h->N);
    D2D_CUDA(cudaGetLastError());
    ++h->launches;
    return D2D_OK;
}

D2D_API int d2d_stats_reset(d2d_handle_t *h, void *stream) {
    if (!h) return fail(D2D_ERR_INVALID_ARG, "d2d_stats_reset: null handle");
    if (!h->stats) return fail(D2D_ERR_STATE, "d2d_stats_reset: no stats buffer bound");
    D2D_CUDA(cudaMemsetAsync(h->stats, 0, (size_t)D2D_STATS_REPLICAS * 8 * sizeof(double), (cudaStream_t)stream));
    return D2D_OK;
}

#ifdef D2D_TIMELINE
// instrumented build only (profiles/timeline.py): copies the stamp table to the host
D2D_API int d2d_debug_timeline(void *host_out, size_t bytes) {
    if (bytes > sizeof(d2d_tl_buf)) bytes = sizeof(d2d_tl_buf);
    D2D_CUDA(cudaDeviceSynchronize());
    D2D_CUDA(cudaMemcpyFromSymbol(host_out, d2d_tl_buf, bytes));
    return D2D_OK;
}
#endif
D2D_API int64_t d2d_launch_count(const d2d_handle_t *h) { return h ? h->launches : -1; }

D2D_API int d2d_step_geometry(const d2d_handle_t *h, int32_t *grid, int32_t *block, int32_t *smem_bytes, int32_t *envs_per_block) {
    if (!h) return fail(D2D_ERR_INVALID_ARG, "d2d_step_geometry: null handle");
    if (grid) *grid = h->grid;
    if (block) *block = h->block;
    if (smem_bytes) *smem_bytes = h->smem;
    if (envs_per_block) *envs_per_block = h->envs_per_block;
    return D2D_OK;
}

