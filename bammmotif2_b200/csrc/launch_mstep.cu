// launch_mstep.cu — instantiations and launcher of the packed M-step kernels (mstep.cuh), one per column count of a CTA.
#include "launch.h"
#include "mstep.cuh"

namespace bamm {

template <int NC, bool HG>
static int mstep_one(int mode, int grid, size_t smem, cudaStream_t stream, const PackedView* pv, const Plan* pl, const ActiveList* al,
                     uint32_t nregions, int nsplit, MTables mt, unsigned long long* d_part, const float* d_r, const float* d_scale,
                     const uint32_t* only_if) {
    if (mode == 0)
        return (cudaFuncSetAttribute(k_mstep_list_w<NC, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                cudaFuncSetAttribute(k_mstep_scan_w<NC, HG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) ? 0 : -1;
    if (mode == 1) k_mstep_list_w<NC, HG><<<grid, 1024, smem, stream>>>(*pv, *pl, *al, nregions, nsplit, mt, d_part);
    else k_mstep_scan_w<NC, HG><<<grid, 1024, smem, stream>>>(*pv, *pl, d_r, d_scale, only_if, nsplit, mt, d_part);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_mstep_packed(int nc, int mode, int grid, size_t smem, cudaStream_t stream, const PackedView* pv, const Plan* pl,
                        const ActiveList* al, uint32_t nregions, int nsplit, MTables mt, unsigned long long* d_part,
                        const float* d_r, const float* d_scale, const uint32_t* only_if) {
    if (mt.hi_global) {           // orders >= 5: at most 14 columns of low words fit (4^6 rows)
        switch (nc) {
#define X(w) case w: return mstep_one<w, true>(mode, grid, smem, stream, pv, pl, al, nregions, nsplit, mt, d_part, d_r, d_scale, only_if);
            X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)
#undef X
            default: return -1;
        }
    }
    switch (nc) {
#define X(w) case w: return mstep_one<w, false>(mode, grid, smem, stream, pv, pl, al, nregions, nsplit, mt, d_part, d_r, d_scale, only_if);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
        X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32)
#undef X
        default: return -1;
    }
}

}  // namespace bamm
