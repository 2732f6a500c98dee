// api.cu -- error string, device info and the cuBLAS plumbing for the *plain* GEMMs of the path.
// (cuBLAS is used only for library-shaped dense layers in the fp32 parity mode; the fused relation kernels and the
//  fp16 GEMMs are hand-written tcgen05 code in gemm_tc.cu / relation_tc.cu.)
#include "common.cuh"
#include <cublas_v2.h>
#include <mutex>

namespace rn {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_sm_count = -1, g_cc_major = -1, g_cc_minor = -1, g_bound_dev = -1;
static std::once_flag g_dev_once;
static void query_dev() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { g_sm_count = 0; g_cc_major = g_cc_minor = 0; cudaGetLastError(); return; }
  g_bound_dev = dev;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { g_sm_count = 0; g_cc_major = g_cc_minor = 0; cudaGetLastError(); return; }
  g_sm_count = p.multiProcessorCount; g_cc_major = p.major; g_cc_minor = p.minor;
}
int sm_count() { std::call_once(g_dev_once, query_dev); return g_sm_count; }
// one device per process (header: PROCESS MODEL): every cached property / launch attribute belongs to g_bound_dev
static bool on_bound_device() {
  int dev = -1;
  if (g_bound_dev < 0 || cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return true; }
  if (dev == g_bound_dev) return true;
  set_error("librelnet_b200 is bound to device %d (first use) but device %d is current: one device per process", g_bound_dev, dev);
  return false;
}
bool is_sm100() { std::call_once(g_dev_once, query_dev); return g_cc_major == 10 && on_bound_device(); }
int check_bound_device() { std::call_once(g_dev_once, query_dev); return on_bound_device() ? RN_OK : RN_ERR_INVALID; }

struct CublasTls {
  cublasHandle_t h = nullptr;
  ~CublasTls() { if (h) cublasDestroy(h); }
};
static thread_local CublasTls g_cublas;

int check_bound_device();
static int get_cublas(cudaStream_t st, cublasHandle_t* out) {
  if (check_bound_device()) return RN_ERR_INVALID;
  if (!g_cublas.h) {
    cublasStatus_t s = cublasCreate(&g_cublas.h);
    if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasCreate failed: %d", (int)s); g_cublas.h = nullptr; return RN_ERR_CUDA; }
    cublasSetMathMode(g_cublas.h, CUBLAS_PEDANTIC_MATH);   // true fp32: no TF32, no reduced-precision reductions
  }
  cublasStatus_t s = cublasSetStream(g_cublas.h, st);
  if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetStream failed: %d", (int)s); return RN_ERR_CUDA; }
  *out = g_cublas.h;
  return RN_OK;
}

// row-major C[M,N] = A[M,K] . B[N,K]^T   <=> col-major C^T[N,M] = op_T(B as [K,N] ld ldb) . A^T ([K,M] ld lda)
int sgemm_nt(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             int batch, long long sA, long long sB, long long sC) {
  if (gemm_backend() == 1 && gemm_tf32_usable(A, lda, 0, sA, B, ldb, 0, sB))
    return gemm_tf32(st, false, true, M, N, K, 1.f, A, lda, 0, sA, B, ldb, 0, sB, 0.f, C, ldc, 0, sC, 1, batch);
  cublasHandle_t h;
  int r = get_cublas(st, &h);
  if (r) return r;
  const float one = 1.f, zero = 0.f;
  cublasStatus_t s;
  if (batch == 1)
    s = cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &one, B, ldb, A, lda, &zero, C, ldc);
  else
    s = cublasSgemmStridedBatched(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &one, B, ldb, sB, A, lda, sA, &zero, C, ldc, sC,
                                  batch);
  if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasSgemm(nt %dx%dx%d) failed: %d", M, N, K, (int)s); return RN_ERR_CUDA; }
  return RN_OK;
}

// row-major C[M,N] = A[M,K] . B[K,N]   <=> col-major C^T[N,M] = B^T ([N,K] ld ldb, op N) . A^T ([K,M] ld lda, op N)
int sgemm_nn(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
             int batch, long long sA, long long sB, long long sC) {
  if (gemm_backend() == 1 && gemm_tf32_usable(A, lda, 0, sA, B, ldb, 0, sB))
    return gemm_tf32(st, false, false, M, N, K, 1.f, A, lda, 0, sA, B, ldb, 0, sB, 0.f, C, ldc, 0, sC, 1, batch);
  cublasHandle_t h;
  int r = get_cublas(st, &h);
  if (r) return r;
  const float one = 1.f, zero = 0.f;
  cublasStatus_t s;
  if (batch == 1)
    s = cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, N, M, K, &one, B, ldb, A, lda, &zero, C, ldc);
  else
    s = cublasSgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, N, M, K, &one, B, ldb, sB, A, lda, sA, &zero, C, ldc, sC,
                                  batch);
  if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasSgemm(nn %dx%dx%d) failed: %d", M, N, K, (int)s); return RN_ERR_CUDA; }
  return RN_OK;
}

// general row-major C[M,N] = alpha * op(A) . op(B) + beta * C, optional strided batch
int sgemm_rm(cudaStream_t st, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
             const float* B, int ldb, float beta, float* C, int ldc, int batch, long long sA, long long sB, long long sC) {
  if (gemm_backend() == 1 && gemm_tf32_usable(A, lda, 0, sA, B, ldb, 0, sB))
    return gemm_tf32(st, transA, transB, M, N, K, alpha, A, lda, 0, sA, B, ldb, 0, sB, beta, C, ldc, 0, sC, 1, batch);
  cublasHandle_t h;
  int r = get_cublas(st, &h);
  if (r) return r;
  const cublasOperation_t opb = transB ? CUBLAS_OP_T : CUBLAS_OP_N, opa = transA ? CUBLAS_OP_T : CUBLAS_OP_N;
  cublasStatus_t s;
  if (batch == 1)
    s = cublasSgemm(h, opb, opa, N, M, K, &alpha, B, ldb, A, lda, &beta, C, ldc);
  else
    s = cublasSgemmStridedBatched(h, opb, opa, N, M, K, &alpha, B, ldb, sB, A, lda, sA, &beta, C, ldc, sC, batch);
  if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasSgemm(rm %dx%dx%d) failed: %d", M, N, K, (int)s); return RN_ERR_CUDA; }
  return RN_OK;
}

// ---- two-level batch: problem (o, i), o < outer, i < inner, operands at base + o*s_outer + i*s_inner.  One
// cublasSgemmBatched over outer*inner pointer triples built on the device (ptr_ws: 3*outer*inner pointers) -- the
// per-class x per-head GEMMs of the learn-NMS head (80 x 16 problems) become one launch instead of 80.
__global__ void fill_gemm_ptrs_kernel(const float* A, long long sAo, long long sAi, const float* B, long long sBo,
                                      long long sBi, float* C, long long sCo, long long sCi, int outer, int inner,
                                      const float** pa, const float** pb, float** pc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= outer * inner) return;
  const int o = t / inner, i = t % inner;
  pa[t] = A + o * sAo + i * sAi;
  pb[t] = B + o * sBo + i * sBi;
  pc[t] = C + o * sCo + i * sCi;
}

int sgemm_rm_2level(cudaStream_t st, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
                    long long sAo, long long sAi, const float* B, int ldb, long long sBo, long long sBi, float beta, float* C,
                    int ldc, long long sCo, long long sCi, int outer, int inner, void* ptr_ws) {
  if (gemm_backend() == 1 && gemm_tf32_usable(A, lda, sAo, sAi, B, ldb, sBo, sBi) && (long long)outer * inner <= 65535)
    return gemm_tf32(st, transA, transB, M, N, K, alpha, A, lda, sAo, sAi, B, ldb, sBo, sBi, beta, C, ldc, sCo, sCi, outer, inner);
  if (outer == 1) return sgemm_rm(st, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, inner, sAi, sBi, sCi);
  cublasHandle_t h;
  int r = get_cublas(st, &h);
  if (r) return r;
  const int cnt = outer * inner;
  const float** pa = (const float**)ptr_ws;
  const float** pb = pa + cnt;
  float** pc = (float**)(pb + cnt);
  fill_gemm_ptrs_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(A, sAo, sAi, B, sBo, sBi, C, sCo, sCi, outer, inner, pa, pb, pc);
  RN_LAUNCH_CHECK();
  const cublasOperation_t opb = transB ? CUBLAS_OP_T : CUBLAS_OP_N, opa = transA ? CUBLAS_OP_T : CUBLAS_OP_N;
  cublasStatus_t s = cublasSgemmBatched(h, opb, opa, N, M, K, &alpha, pb, ldb, pa, lda, &beta, pc, ldc, cnt);
  if (s != CUBLAS_STATUS_SUCCESS) { set_error("cublasSgemmBatched(%dx%dx%d x%d) failed: %d", M, N, K, cnt, (int)s); return RN_ERR_CUDA; }
  return RN_OK;
}

}  // namespace rn

extern "C" {
const char* rn_last_error(void) { return rn::g_err; }
int rn_version(void) { return 100; }
int rn_device_info(int* sm, int* major, int* minor) {
  int n = rn::sm_count();
  if (sm) *sm = n;
  if (major) *major = rn::g_cc_major;
  if (minor) *minor = rn::g_cc_minor;
  return rn::is_sm100() ? 1 : 0;
}
}
