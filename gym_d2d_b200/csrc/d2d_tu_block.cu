// d2d_tu_block.cu - the sorting block kernel (shapes whose double-buffered bins do not fit the dense kernel) and the
// general-topology kernel (any N <= 65535, DOWNLINK links, ShadowingPathLoss): d2d_step_block.cuh.
#include "d2d_internal.h"
#include "d2d_step_block.cuh"

size_t d2d_block_smem(int N, int R, int lpt) { return lpt ? d2d_block2_smem_bytes(N, R) : d2d_block_smem_bytes(N, R); }

int d2d_block_plan(d2d_handle *h, size_t smem) {
#define D2D_PLAN(LPT_) (h->ple2 ? d2d_plan_geometry(h, d2d_step_block_kernel<true, LPT_>, D2D_BLOCK_THREADS, smem, 1) \
                                : d2d_plan_geometry(h, d2d_step_block_kernel<false, LPT_>, D2D_BLOCK_THREADS, smem, 1))
    switch (h->lpt) {
        case 1: return D2D_PLAN(1);
        case 2: return D2D_PLAN(2);
        case 3: return D2D_PLAN(3);
        case 4: return D2D_PLAN(4);
        default:
            return h->ple2 ? d2d_plan_geometry(h, d2d_step_block_generic_kernel<true>, D2D_BLOCK_THREADS, smem, 1)
                           : d2d_plan_geometry(h, d2d_step_block_generic_kernel<false>, D2D_BLOCK_THREADS, smem, 1);
    }
#undef D2D_PLAN
}

cudaError_t d2d_block_launch(const d2d_handle *h, const D2DParams &P, int grid, cudaStream_t st, bool pdl) {
#define D2D_GO(LPT_) (h->ple2 ? d2d_launch_step(d2d_step_block_kernel<true, LPT_>, grid, h->block, (size_t)h->smem, st, P, pdl) \
                              : d2d_launch_step(d2d_step_block_kernel<false, LPT_>, grid, h->block, (size_t)h->smem, st, P, pdl))
    switch (h->lpt) {
        case 1: return D2D_GO(1);
        case 2: return D2D_GO(2);
        case 3: return D2D_GO(3);
        case 4: return D2D_GO(4);
        default:
            return h->ple2 ? d2d_launch_step(d2d_step_block_generic_kernel<true>, grid, h->block, (size_t)h->smem, st, P, pdl)
                           : d2d_launch_step(d2d_step_block_generic_kernel<false>, grid, h->block, (size_t)h->smem, st, P, pdl);
    }
#undef D2D_GO
}
