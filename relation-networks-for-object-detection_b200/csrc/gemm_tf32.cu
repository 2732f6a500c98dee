// gemm_tf32.cu -- hand-written tcgen05 GEMM on fp32 operands (kind::tf32, fp32 accumulate in TMEM) with arbitrary
// transposes and a two-level strided batch:
//     C[o,i][M,N] = alpha * op(A[o,i]) . op(B[o,i]) + beta * C[o,i]         row-major, fp32 in HBM
// It is the contraction engine of the TRAINING side (rn_relation_bwd / rn_learn_nms_bwd: dV' = P^T dO, dP = dO V'^T,
// dQ = dS K, dK = dS^T Q, the projection-weight and input gradients, and the recomputed forward), replacing the cuBLAS
// fp32 calls of round 1.  The reference has no counterpart: MXNet differentiates SYM_REL:104-151 op by op on cuBLAS.
//
// Why tf32 and not fp16 operands: gradients span many decades (fp16 would need loss scaling) and the operands already
// sit in HBM as fp32 -- TMA loads them as they are, the tensor core reads the top 19 bits, no cast kernels, no packed
// copies.  Both operand majors come straight from the row-major matrices:
//   op(A) = A   (A stored [M,K]):  K-major tile, ONE TMA box [128 rows x 32 k] (SWIZZLE_128B rows of 32 floats)
//   op(A) = A^T (A stored [K,M]):  MN-major tile, FOUR boxes [32 k rows x 32 m].  MN-major tf32 has exactly one legal
//                                  shared-memory layout, SWIZZLE_128B with a 32-byte base (cute Layout_MN_SW128_32B_Atom,
//                                  Swizzle<2,5,2>: atoms of 32 m x 4 k): TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
//                                  descriptor layout type 1, LBO = 4096 (next 32 m), SBO = 512 (next 4 k), +1024 B per K step of 8
//   op(B): the same with N in place of M (B stored [N,K] = "transB" is the K-major case).
// Out-of-range rows / columns / K are zero-filled by TMA (tensor-map extents are the true per-problem extents), so no
// operand is ever padded in memory; the epilogue masks its stores.
//
// Small problems: short K loops run with a 2-stage ring (64 KB: three CTAs share an SM, which matters for the 1280 per-class x
// per-head problems of the learn-NMS head); long K loops of single problems with few output tiles are split over gridDim.z
// and combined with red.global.add.f32 (C zeroed first when beta = 0).
// Structure: one CTA = one 128 x BN output tile of one problem (and one K split); warp 4 = TMA producer, warp 5 = TMEM allocation + MMA
// issue (one lane each), warps 0-3 = epilogue (TMEM lane quarter each; alpha / beta applied on the way out).  4-stage
// ring of (A 16 KB + B <= 16 KB) stages, full/empty mbarriers, tcgen05.commit releases a stage.
// Roofline: tensor pipe for the weight-gradient GEMMs (K = rois is short: they are launch/latency sized), L2 bandwidth
// for the per-head attention gradients (operands are L2 resident: written by the previous kernel).
#include "common.cuh"
#include "umma.cuh"
#include <mutex>

namespace rn {
using namespace umma;

namespace {

constexpr int kTM = 128, kTK = 32, kTMaxStages = 4;
constexpr int kTA = kTM * kTK * 4;                   // 16 KB
constexpr int kTStage = 2 * kTA;                     // A + B (B up to 128 columns)
static inline int tf32_smem(int stages) { return stages * kTStage + 256 + 1024; }

struct Tf32Params {
  int M, N, K, BN;
  int a_mn, b_mn;                                    // 1: operand is MN-major in shared memory (stored transposed)
  int inner;                                         // problems per outer index (blockIdx.y = o * inner + i)
  int stages;                                        // ring depth (2 or 4); barriers live after the ring
  int splits, kb_per_split;                          // K splits (gridDim.z); > 1: the epilogue adds its part atomically
  int a_i, a_o, b_i, b_o;                            // 0: the operand is shared along that batch level (stride 0) -> coordinate 0
  float alpha, beta;
  float* C; long long ldc, sCo, sCi;
};

// shared-memory descriptor of an MN-major tf32 operand: layout type 1 = SWIZZLE_128B_BASE32B (cute UMMA::LayoutType)
__device__ __forceinline__ uint64_t make_smem_desc_sw128_base32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

// kind::tf32 instruction descriptor (cute InstrDescriptor): c_format F32, a/b_format 2 = TF32
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(192) gemm_tf32_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, const Tf32Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int kTStages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kTStages * kTStage);
  uint64_t* empty = full + kTMaxStages;
  uint64_t* tfull = empty + kTMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (p.N + p.BN - 1) / p.BN;
  const int m0 = (blockIdx.x / tiles_n) * kTM, n0 = (blockIdx.x % tiles_n) * p.BN;
  const int bo = blockIdx.y / p.inner, bi = blockIdx.y % p.inner;
  const int nkb_all = (p.K + kTK - 1) / kTK;
  const int kb0 = blockIdx.z * p.kb_per_split, nkb = max(0, min(nkb_all, kb0 + p.kb_per_split) - kb0);   // this split's K blocks

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < kTStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t bytesA = kTA;                                  // OOB parts of a box still count as transferred bytes
      const uint32_t bytesB = (uint32_t)p.BN * kTK * 4;
      for (int it = 0; it < nkb; ++it) {
        const int kb = kb0 + it;
        const int s = it % kTStages;
        mbar_wait(&empty[s], ((it / kTStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], bytesA + bytesB);
        uint8_t* sa = smem + s * kTStage;
        uint8_t* sb = sa + kTA;
        if (p.a_mn) {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(sa + j * 4096, &tmA, &full[s], m0 + 32 * j, kb * kTK, bi * p.a_i, bo * p.a_o);
        } else {
          tma_load_4d(sa, &tmA, &full[s], kb * kTK, m0, bi * p.a_i, bo * p.a_o);
        }
        if (p.b_mn) {
          for (int j = 0; j < p.BN / 32; ++j) tma_load_4d(sb + j * 4096, &tmB, &full[s], n0 + 32 * j, kb * kTK, bi * p.b_i, bo * p.b_o);
        } else {
          tma_load_4d(sb, &tmB, &full[s], kb * kTK, n0, bi * p.b_i, bo * p.b_o);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(kTM, p.BN, p.a_mn != 0, p.b_mn != 0);
      const uint32_t stepA = p.a_mn ? 1024u : 32u, stepB = p.b_mn ? 1024u : 32u;        // bytes per K step of 8
      for (int it = 0; it < nkb; ++it) {
        const int kb = kb0 + it;
        const int s = it % kTStages;
        mbar_wait(&full[s], (it / kTStages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * kTStage), sb = sa + kTA;
        const int ksteps = min(kTK, p.K - kb * kTK + 7) >> 3;          // whole K steps that hold at least one real k
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t da = p.a_mn ? make_smem_desc_sw128_base32(sa + k * stepA, 4096, 512) : make_smem_desc_sw128(sa + k * stepA, 16, 1024);
          const uint64_t db = p.b_mn ? make_smem_desc_sw128_base32(sb + k * stepB, 4096, 512) : make_smem_desc_sw128(sb + k * stepB, 16, 1024);
          mma_tf32_ss(tmem_base, da, db, idesc, (it > 0) || (k > 0));
        }
        mma_commit(&empty[s]);
      }
      mma_commit(tfull);
    }
  } else {
    // epilogue: thread (warp, lane) owns accumulator row 32 * warp + lane
    const int row = m0 + warp * 32 + lane;
    float* Cb = p.C + (long long)bo * p.sCo + (long long)bi * p.sCi;
    if (nkb > 0) { mbar_wait(tfull, 0); tc_fence_after(); }
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool vec = (p.ldc & 3) == 0 && (((size_t)Cb) & 15) == 0;
#pragma unroll 1
    for (int c = 0; c < p.BN; c += 16) {
      uint32_t v[16];
      if (nkb > 0) { tmem_ld_32x32b_x16(lane_base + c, v); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      const int col0 = n0 + c;
      if (row >= p.M || col0 >= p.N) continue;
      float* dst = Cb + (long long)row * p.ldc + col0;
      if (p.splits > 1) {                 // K split: every split adds its part (beta was applied by zeroing / keeping C)
        if (vec && col0 + 16 <= p.N) {    // 16-byte vector reductions: a quarter of the L2 atomic transactions
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(p.alpha * __uint_as_float(v[j])),
                         "f"(p.alpha * __uint_as_float(v[j + 1])), "f"(p.alpha * __uint_as_float(v[j + 2])),
                         "f"(p.alpha * __uint_as_float(v[j + 3]))
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col0 + j < p.N) atomicAdd(dst + j, p.alpha * __uint_as_float(v[j]));
        }
        continue;
      }
      if (vec && col0 + 16 <= p.N) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 o = make_float4(p.alpha * __uint_as_float(v[j]), p.alpha * __uint_as_float(v[j + 1]),
                                 p.alpha * __uint_as_float(v[j + 2]), p.alpha * __uint_as_float(v[j + 3]));
          if (p.beta != 0.f) {
            const float4 old = *reinterpret_cast<const float4*>(dst + j);
            o.x = fmaf(p.beta, old.x, o.x); o.y = fmaf(p.beta, old.y, o.y); o.z = fmaf(p.beta, old.z, o.z); o.w = fmaf(p.beta, old.w, o.w);
          }
          *reinterpret_cast<float4*>(dst + j) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.N) {
            float o = p.alpha * __uint_as_float(v[j]);
            if (p.beta != 0.f) o = fmaf(p.beta, dst[j], o);
            dst[j] = o;
          }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<128>(tmem_base);
}

// ---- host: fp32 tensor maps --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode32 = nullptr;
std::once_flag g_once32;
void load_encode32() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    g_encode32 = (EncodeTiledFn)fn;
  else
    cudaGetLastError();
}

// matrix stored row-major [rows, cols] (pitch ld floats) per problem, problems at base + o*sO + i*sI; box = 32 cols x box_rows
int encode_f32_4d(CUtensorMap* out, const float* base, int rows, int cols, long long ld, int inner, int outer, long long sI,
                  long long sO, int box_rows, bool mn_major) {
  std::call_once(g_once32, load_encode32);
  if (!g_encode32) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return RN_ERR_CUDA; }
  const long long fallback = ld * (long long)rows;            // any valid stride for extent-1 dimensions
  // a batch level the operand is shared along (stride 0, e.g. one weight matrix for every problem) becomes an extent-1
  // dimension; the kernel then passes coordinate 0 for it (Tf32Params::a_i ...)
  const bool use_i = inner > 1 && sI != 0, use_o = outer > 1 && sO != 0;
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(use_i ? inner : 1), (cuuint64_t)(use_o ? outer : 1)};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)(use_i ? sI : fallback) * 4, (cuuint64_t)(use_o ? sO : fallback) * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode32(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tf32: cuTensorMapEncodeTiled failed: %d (ptr=%p rows=%d cols=%d ld=%lld inner=%d outer=%d sI=%lld sO=%lld)", (int)r,
              (const void*)base, rows, cols, ld, inner, outer, sI, sO);
    return RN_ERR_CUDA;
  }
  return RN_OK;
}

thread_local int g_backend = 0;          // 0 = cuBLAS fp32 (pedantic), 1 = tcgen05 tf32

}  // namespace

bool gemm_tf32_usable(const float* A, int lda, long long sAo, long long sAi, const float* B, int ldb, long long sBo, long long sBi) {
  auto ok = [](const float* p, long long ld, long long so, long long si) {
    return (((size_t)p) & 15) == 0 && (ld & 3) == 0 && (so & 3) == 0 && (si & 3) == 0 && ld > 0;
  };
  return is_sm100() && ok(A, lda, sAo, sAi) && ok(B, ldb, sBo, sBi);
}

int gemm_tf32(cudaStream_t st, bool transA, bool transB, int M, int N, int K, float alpha, const float* A, int lda,
              long long sAo, long long sAi, const float* B, int ldb, long long sBo, long long sBi, float beta, float* C,
              int ldc, long long sCo, long long sCi, int outer, int inner) {
  if (M <= 0 || N <= 0 || outer <= 0 || inner <= 0) return RN_OK;
  RN_CHECK_ARG(gemm_tf32_usable(A, lda, sAo, sAi, B, ldb, sBo, sBi),
               "gemm_tf32: operands need 16-byte aligned bases and pitches / batch strides that are multiples of 4 floats");
  RN_CHECK_ARG((long long)outer * inner <= 65535, "gemm_tf32: %d x %d problems exceed the grid", outer, inner);
  Tf32Params p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = transA ? 1 : 0;                 // op(A) = A^T: A stored [K, M] -> M contiguous
  p.b_mn = transB ? 0 : 1;                 // op(B) = B:   B stored [K, N] -> N contiguous
  const int gran = p.b_mn ? 32 : 16;
  p.BN = std::min(128, (N + gran - 1) / gran * gran);
  p.inner = inner; p.alpha = alpha; p.beta = beta;
  p.a_i = (inner > 1 && sAi != 0) ? 1 : 0; p.a_o = (outer > 1 && sAo != 0) ? 1 : 0;
  p.b_i = (inner > 1 && sBi != 0) ? 1 : 0; p.b_o = (outer > 1 && sBo != 0) ? 1 : 0;
  p.C = C; p.ldc = ldc; p.sCo = sCo; p.sCi = sCi;
  CUtensorMap tmA, tmB;
  int r;
  if (p.a_mn) r = encode_f32_4d(&tmA, A, K, M, lda, inner, outer, sAi, sAo, kTK, true);
  else r = encode_f32_4d(&tmA, A, M, K, lda, inner, outer, sAi, sAo, kTM, false);
  if (r) return r;
  if (p.b_mn) r = encode_f32_4d(&tmB, B, K, N, ldb, inner, outer, sBi, sBo, kTK, true);
  else r = encode_f32_4d(&tmB, B, N, K, ldb, inner, outer, sBi, sBo, p.BN, false);
  if (r) return r;
  static thread_local bool configured = false;
  if (!configured) {
    RN_CUDA(cudaFuncSetAttribute(gemm_tf32_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tf32_smem(kTMaxStages)));
    configured = true;
  }
  const int tiles = cdiv(M, kTM) * cdiv(N, p.BN), nkb = cdiv(K, kTK);
  const int sms = sm_count() > 0 ? sm_count() : 148;
  p.splits = 1; p.kb_per_split = nkb > 0 ? nkb : 1;
  if (outer * inner == 1 && (beta == 0.f || beta == 1.f) && nkb >= 8 && tiles * 2 <= sms) {
    int want = std::min(std::min(sms / tiles, nkb / 4), 32);
    if (want > 1) {
      p.kb_per_split = cdiv(nkb, want);
      p.splits = cdiv(nkb, p.kb_per_split);
    }
  }
  if (p.splits > 1 && beta == 0.f)        // the splits accumulate into C
    RN_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
  p.stages = p.kb_per_split <= 2 ? 2 : kTMaxStages;
  gemm_tf32_tc_kernel<<<dim3((unsigned)tiles, (unsigned)(outer * inner), (unsigned)p.splits), 192, tf32_smem(p.stages), st>>>(tmA, tmB, p);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

// ---- backend switch of the general GEMM helpers (thread-local, set by the backward entry points) -----------------------
int gemm_backend() { return g_backend; }
void set_gemm_backend(int b) { g_backend = b; }

}  // namespace rn

// C ABI: exported so the contraction engine of the training side can be checked on its own (tests/test_gpu_backward.py)
extern "C" int rn_gemm_tf32(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K, float alpha, const float* A,
                            int32_t lda, int64_t sAo, int64_t sAi, const float* B, int32_t ldb, int64_t sBo, int64_t sBi, float beta,
                            float* C, int32_t ldc, int64_t sCo, int64_t sCi, int32_t outer, int32_t inner, rn_stream_t stream) {
  RN_CHECK_ARG(A && B && C && M >= 0 && N >= 0 && K >= 0, "rn_gemm_tf32: bad arguments");
  RN_CHECK_ARG(rn::is_sm100(), "rn_gemm_tf32: needs an sm_100 device");
  return rn::gemm_tf32((cudaStream_t)stream, transA != 0, transB != 0, M, N, K, alpha, A, lda, sAo, sAi, B, ldb, sBo, sBi, beta, C,
                       ldc, sCo, sCi, outer, inner);
}
