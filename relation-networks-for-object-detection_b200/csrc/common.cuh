// common.cuh -- error plumbing and small device helpers shared by every translation unit of librelnet_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "relnet_b200.h"

namespace rn {

void set_error(const char* fmt, ...);          // thread-local message returned by rn_last_error()

#define RN_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) { rn::set_error(__VA_ARGS__); return RN_ERR_INVALID; } \
  } while (0)

#define RN_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      rn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));         \
      return RN_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

#define RN_LAUNCH_CHECK()                                                                          \
  do {                                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                           \
    if (e_ != cudaSuccess) {                                                                       \
      rn::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));     \
      return RN_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// bump allocator over the caller's workspace (256-byte aligned slices)
struct Workspace {
  char* base; size_t size; size_t off;
  Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
  template <typename T> T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (off + bytes > size) return nullptr;
    T* r = (T*)(base + off); off += bytes; return r;
  }
};
static inline size_t ws_slice(size_t count, size_t elt) { return align_up(count * elt, 256); }

int sm_count();                // cached per process (device of the current context)
bool is_sm100();

// plain GEMM helper (cuBLAS): C[M,N] (row-major) = A[M,K] . B[N,K]^T, fp32, optional batch with strides
int sgemm_nt(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             int batch = 1, long long sA = 0, long long sB = 0, long long sC = 0);
// C[M,N] = A[M,K] . B[K,N] (row-major), fp32
int sgemm_nn(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             int batch = 1, long long sA = 0, long long sB = 0, long long sC = 0);

// general row-major C[M,N] = alpha * op(A) . op(B) + beta * C (cuBLAS fp32, pedantic)
int sgemm_rm(cudaStream_t st, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
             const float* B, int ldb, float beta, float* C, int ldc, int batch = 1, long long sA = 0, long long sB = 0,
             long long sC = 0);

// two-level batched variant: problem (o, i) at base + o*s_outer + i*s_inner; ptr_ws holds 3*outer*inner device pointers
int sgemm_rm_2level(cudaStream_t st, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
                    long long sAo, long long sAi, const float* B, int ldb, long long sBo, long long sBi, float beta, float* C,
                    int ldc, long long sCo, long long sCi, int outer, int inner, void* ptr_ws);

// tcgen05 tf32 GEMM on fp32 operands (gemm_tf32.cu): same contract as sgemm_rm_2level (no pointer workspace needed)
bool gemm_tf32_usable(const float* A, int lda, long long sAo, long long sAi, const float* B, int ldb, long long sBo, long long sBi);
int gemm_tf32(cudaStream_t st, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
              long long sAo, long long sAi, const float* B, int ldb, long long sBo, long long sBi, float beta, float* C,
              int ldc, long long sCo, long long sCi, int outer, int inner);
// backend of sgemm_nt / sgemm_nn / sgemm_rm / sgemm_rm_2level for the calling thread: 0 = cuBLAS fp32 (pedantic; the
// bit-conservative parity mode), 1 = the repo's tcgen05 tf32 GEMM wherever its alignment rules hold (training side)
int gemm_backend();
void set_gemm_backend(int b);
struct GemmBackendScope {
  int prev;
  explicit GemmBackendScope(int b) : prev(gemm_backend()) { set_gemm_backend(b); }
  ~GemmBackendScope() { set_gemm_backend(prev); }
};

}  // namespace rn
