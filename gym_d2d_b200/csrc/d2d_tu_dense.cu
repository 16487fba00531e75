// d2d_tu_dense.cu - the instantiations of d2d_step_dense_kernel (d2d_step_dense.cuh): one block per env, 65 <= N <= 1024 links.
#include "d2d_internal.h"
#include "d2d_step_dense.cuh"

// (links per thread, threads per block) instantiations
#define D2D_DENSE_SHAPES(X) X(1, 256) X(2, 256) X(3, 256) X(4, 256) X(1, 320) X(2, 320) X(3, 320)

size_t d2d_dense_smem(int N, int R, int bin_cap, int bt, int V) { return d2d_dense_layout(N, R, bin_cap, bt, V).total; }
int d2d_dense_bin_cap_host(int N, int R) { return d2d_dense_bin_cap(N, R); }

namespace {
// every (FULL, EXACT) instantiation d2d_step may launch for this handle needs the dynamic shared-memory opt-in
template <bool PLE2, int LPT, int BT>
int plan(d2d_handle *h, size_t smem) {
    int rc = d2d_allow_smem(d2d_step_dense_kernel<PLE2, LPT, BT, false, false>, smem);
    if (!rc) rc = d2d_allow_smem(d2d_step_dense_kernel<PLE2, LPT, BT, false, true>, smem);
    if (!rc) rc = d2d_allow_smem(d2d_step_dense_kernel<PLE2, LPT, BT, true, true>, smem);
    if (!rc) rc = d2d_plan_geometry(h, d2d_step_dense_kernel<PLE2, LPT, BT, true, false>, BT, smem, 1);
    return rc;
}
// the instantiations for BASELINE config #3's shape (d2d_step_dense.cuh: SPEC)
int plan_spec(size_t smem) {
    int rc = d2d_allow_smem(d2d_step_dense_kernel<true, 2, 320, false, false, true>, smem);
    if (!rc) rc = d2d_allow_smem(d2d_step_dense_kernel<true, 2, 320, false, true, true>, smem);
    if (!rc) rc = d2d_allow_smem(d2d_step_dense_kernel<true, 2, 320, true, true, true>, smem);
    if (!rc) rc = d2d_allow_smem(d2d_step_dense_kernel<true, 2, 320, true, false, true>, smem);
    return rc;
}
template <bool PLE2, int LPT, int BT>
cudaError_t launch(const d2d_handle *h, const D2DParams &P, int grid, const D2DLaunchSel &sel, cudaStream_t st, bool pdl) {
#define D2D_GO(FULL_, EXACT_) d2d_launch_step(d2d_step_dense_kernel<PLE2, LPT, BT, FULL_, EXACT_>, grid, h->block, (size_t)h->smem, st, P, pdl)
    if (sel.full && h->uniform) return sel.exact ? D2D_GO(true, true) : D2D_GO(true, false);
    return sel.exact ? D2D_GO(false, true) : D2D_GO(false, false);
#undef D2D_GO
}
cudaError_t launch_spec(const d2d_handle *h, const D2DParams &P, int grid, const D2DLaunchSel &sel, cudaStream_t st, bool pdl) {
#define D2D_GO(FULL_, EXACT_) d2d_launch_step(d2d_step_dense_kernel<true, 2, 320, FULL_, EXACT_, true>, grid, h->block, (size_t)h->smem, st, P, pdl)
    if (sel.full && h->uniform) return sel.exact ? D2D_GO(true, true) : D2D_GO(true, false);
    return sel.exact ? D2D_GO(false, true) : D2D_GO(false, false);
#undef D2D_GO
}
}  // namespace

int d2d_dense_plan(d2d_handle *h, size_t smem) {
    int rc = d2d_fail(D2D_ERR_UNSUPPORTED, "d2d_create: no dense kernel instantiation for this shape");
#define D2D_CASE(LPT_, BT_) \
    if (h->lpt == LPT_ && h->dense_bt == BT_) rc = h->ple2 ? plan<true, LPT_, BT_>(h, smem) : plan<false, LPT_, BT_>(h, smem);
    D2D_DENSE_SHAPES(D2D_CASE)
#undef D2D_CASE
    if (!rc && h->spec) rc = plan_spec(smem);
    return rc;
}

cudaError_t d2d_dense_launch(const d2d_handle *h, const D2DParams &P, int grid, const D2DLaunchSel &sel, cudaStream_t st, bool pdl) {
    if (h->spec) return launch_spec(h, P, grid, sel, st, pdl);
    cudaError_t err = cudaErrorInvalidValue;
#define D2D_CASE(LPT_, BT_)                                                                                      \
    if (h->lpt == LPT_ && h->dense_bt == BT_)                                                                    \
        err = h->ple2 ? launch<true, LPT_, BT_>(h, P, grid, sel, st, pdl) : launch<false, LPT_, BT_>(h, P, grid, sel, st, pdl);
    D2D_DENSE_SHAPES(D2D_CASE)
#undef D2D_CASE
    return err;
}
