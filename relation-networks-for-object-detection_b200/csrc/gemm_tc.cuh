// gemm_tc.cuh -- internal interface of the tcgen05 GEMM (gemm_tc.cu)
#pragma once
#include <cuda_fp16.h>
#include <algorithm>
#include "common.cuh"

namespace rn {

// C[M,N] = act(A[M,K] . B[N,K]^T + bias); A,B fp16 with pitches lda/ldb (multiples of 8); outputs fp32 and/or fp16.
// bias_per_row: bias indexed by output row instead of column.  ws: optional split-K scratch (gemm_tc_workspace_bytes).
int gemm_tc(cudaStream_t st, const __half* A, long long lda, const __half* B, long long ldb, int M, int N, int K,
            const float* bias, int bias_per_row, int relu, float* C32, long long ldc32, __half* C16, long long ldc16,
            void* ws, size_t ws_bytes);
size_t gemm_tc_workspace_bytes(int M, int N, int K);
int cast_f32_f16(cudaStream_t st, const float* src, __half* dst, size_t n);
int cast_rows_f16(cudaStream_t st, const float* src, __half* dst, int rows, int cols, int ld);

}  // namespace rn
