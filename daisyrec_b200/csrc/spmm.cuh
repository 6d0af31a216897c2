// spmm.cuh -- the segmented-CSR sparse x dense product of lightgcn.cu, shared with ngcf.cu.
#pragma once
#include "common.cuh"

namespace drb {

struct Adj {
    const int64_t *row_ptr;
    const int32_t *col;
    const float *val;
    const int32_t *seg_row;      // row of every <= 256-edge segment (drb_lgcn_segments)
    const int64_t *seg_ptr;
    long long nseg, n;
};

// Y = A X  ([n, F] fp32, row-major);  S != nullptr: S += A X as well (LightGCN's layer sum)
int launch_spmm(const Adj &a, const float *X, float *Y, float *S, int F, cudaStream_t st);

}  // namespace drb
