// stats.cuh — score statistics behind the scoring path: the sorts of FDR::calculatePR /
// ScoreSeqSet::calcPvalues and the per-window p-value of ScoreSeqSet::calcPvalues on the device.
// The sort is CUB's radix sort (library code, like cuBLAS for a GEMM: not a kernel this repository claims);
// the p-value kernel restates ScoreSeqSet.cpp:98-124 per window.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include "kernels.cuh"

namespace bamm {

// reference: src/seq_scoring/ScoreSeqSet.cpp:98-124. neg: ALL negative window scores sorted ascending.
//   FPl = number of negatives strictly above Sl (upper_bound, :103-104)
//   FPl == negN                      -> p = 1
//   FPl < 10 and |lambda| > eps      -> p = nTop / negN * expf( -(Sl - S_ntop) / lambda )
//   else                             -> p = ( FPl + (S_higher - Sl + eps) / (S_higher - S_lower + eps) ) / negN
// e-value = p * number of positive sequences (:122).
__global__ void __launch_bounds__(256)
k_mops_pvalues(const float* __restrict__ neg, unsigned long long negN, const float* __restrict__ pos, unsigned long long npos,
               float S_ntop, float lambda, float nTop, float posN, float* __restrict__ pval, float* __restrict__ eval) {
    const float eps = 1.0e-5;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < npos; i += (unsigned long long)gridDim.x * blockDim.x) {
        const float Sl = pos[i];
        unsigned long long lo = 0, hi = negN;                      // first index with neg[idx] > Sl
        while (lo < hi) {
            const unsigned long long mid = (lo + hi) >> 1;
            if (neg[mid] > Sl) hi = mid; else lo = mid + 1;
        }
        const unsigned long long FPl = negN - lo;
        float p;
        if (FPl == negN) p = 1.0f;
        else if (FPl < 10 && fabsf(lambda) > eps) p = nTop / (float)negN * expf(-(Sl - S_ntop) / lambda);
        else {
            // FPl == 0 (no negative above Sl) only gets here without a usable tail fit: the reference then reads one element
            // past its vector (ScoreSeqSet.cpp:112); Sl itself stands in for it, which gives p = 1 / negN
            const float SlHigher = neg[negN - FPl - 1], SlLower = FPl ? neg[negN - FPl] : Sl;
            p = ((float)FPl + (SlHigher - Sl + eps) / (SlHigher - SlLower + eps)) / (float)negN;
        }
        pval[i] = p;
        eval[i] = p * posN;
    }
}

}  // namespace bamm
