// gemm_tc.cuh -- internal interface of the tcgen05 GEMM (gemm_tc.cu)
#pragma once
#include <cuda_fp16.h>
#include <algorithm>
#include "common.cuh"

namespace rn {

// optional output segmentation: columns [begin[i], begin[i] + cols[i]) go to the contiguous matrix ptr[i] [M, cols[i]]
struct GemmSegments { int n; int begin[4]; int cols[4]; float* ptr[4]; };

// C[M,N] = act(A[M,K] . B[N,K]^T + bias); A,B fp16 with pitches lda/ldb (multiples of 8); outputs fp32 and/or fp16.
// bias_per_row: bias indexed by output row instead of column.  ws: optional split-K scratch (gemm_tc_workspace_bytes).
int gemm_tc(cudaStream_t st, const __half* A, long long lda, const __half* B, long long ldb, int M, int N, int K,
            const float* bias, int bias_per_row, int relu, float* C32, long long ldc32, __half* C16, long long ldc16,
            void* ws, size_t ws_bytes, const GemmSegments* seg = nullptr, bool bf16 = false);   // bf16: A, B hold bf16 bits
size_t gemm_tc_workspace_bytes(int M, int N, int K);
int cast_f32_f16(cudaStream_t st, const float* src, __half* dst, size_t n);
int cast_rows_f16(cudaStream_t st, const float* src, __half* dst, int rows, int cols, int ld);
// several FC layers over one fp16 input as ONE GEMM (layer i's packed rows start at a multiple of 32)
size_t linear_multi_packed_bytes(const int32_t* outs, int nout, int in);
int linear_multi_pack(const float* const* W, const float* const* b, const int32_t* outs, int nout, int in, void* packed,
                      cudaStream_t st);
int linear_multi_packed_f16in(const void* x16, const void* packed, float* const* ys, const int32_t* outs, int nout, int rows,
                              int in, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace rn
