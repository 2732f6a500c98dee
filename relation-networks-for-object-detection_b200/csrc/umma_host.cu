// umma_host.cu -- host side of the TMA plumbing (tensor-map encode through the runtime's driver entry point, so the
// library has no link-time dependency on libcuda) and the tcgen05 self-test kernel exported as rn_umma_selftest.
#include "common.cuh"
#include "umma.cuh"
#include <mutex>

namespace rn {
namespace umma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_once;

static void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode = (EncodeTiledFn)fn;
  else
    cudaGetLastError();
}

static int encode(CUtensorMap* out, const void* gptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  std::call_once(g_once, load_encode);
  if (!g_encode) { set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old / no GPU)"); return RN_ERR_CUDA; }
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = box[0] * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : box[0] * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : box[0] * 2 == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(gptr), dims,
                        strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (ptr=%p dims=%llu,%llu pitch=%llu box=%u,%u)", (int)r, gptr,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0], box[0], box[1]);
    return RN_ERR_CUDA;
  }
  return RN_OK;
}

int encode_tmap_2d_f16(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                       uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(out, gptr, 2, dims, strides, box);
}

int encode_tmap_3d_f16(CUtensorMap* out, const void* gptr, uint64_t batch, uint64_t rows, uint64_t cols,
                       uint64_t pitch_elems, uint64_t batch_pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {pitch_elems * 2, batch_pitch_elems * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  return encode(out, gptr, 3, dims, strides, box);
}

// ---------------------------------------------------------------------------------------------------- self-test
// S[128,128] = A[128,64] . B[128,64]^T   (A,B K-major SWIZZLE_128B via TMA)
// O[128, 64] = P[128,128] . V[128,64]    (P K-major written by threads with the manual swizzle, V MN-major via TMA)
__global__ void __launch_bounds__(128) umma_selftest_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const __grid_constant__ CUtensorMap tmV,
                                                            const __half* __restrict__ P, float* __restrict__ outS,
                                                            float* __restrict__ outO) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sA = smem;                 // 16 KB
  uint8_t* sB = smem + 16384;         // 16 KB
  uint8_t* sP = smem + 32768;         // 32 KB (two [128 x 64] K sub-tiles)
  uint8_t* sV = smem + 65536;         // 16 KB
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + 81920);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc<256>(tmem_slot);
  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  // P: thread t owns row t; 16 chunks of 8 halfs
  {
    const uint4* src = reinterpret_cast<const uint4*>(P + (size_t)tid * 128);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      uint4 v = src[c];
      *reinterpret_cast<uint4*>(sP + (c >> 3) * 16384 + sw128_offset(tid, c & 7)) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_tma, 3 * 16384);
    tma_load_2d(sA, &tmA, bar_tma, 0, 0);
    tma_load_2d(sB, &tmB, bar_tma, 0, 0);
    tma_load_2d(sV, &tmV, bar_tma, 0, 0);
    mbar_wait(bar_tma, 0);
    tc_fence_after();
    const uint32_t idesc_s = make_idesc_f16(128, 128, false, false, false);
    const uint32_t idesc_o = make_idesc_f16(128, 64, false, false, true);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      mma_f16_ss(tmem_base, make_smem_desc_sw128(smem_u32(sA) + k * 32, 16, 1024),
                 make_smem_desc_sw128(smem_u32(sB) + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      mma_f16_ss(tmem_base + 128, make_smem_desc_sw128(smem_u32(sP) + (i >> 2) * 16384 + (i & 3) * 32, 16, 1024),
                 make_smem_desc_sw128(smem_u32(sV) + i * 2048, 1024, 1024), idesc_o, i > 0);
    mma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32];
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    tmem_ld_32x32b_x32(lane_base + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) outS[(size_t)tid * 128 + c * 32 + j] = __uint_as_float(v[j]);
  }
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    tmem_ld_32x32b_x32(lane_base + 128 + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) outO[(size_t)tid * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace umma
}  // namespace rn

extern "C" int rn_umma_selftest(const void* a, const void* b, const void* p, const void* v, float* out_s, float* out_o,
                                rn_stream_t stream) {
  using namespace rn::umma;
  RN_CHECK_ARG(rn::is_sm100(), "rn_umma_selftest: needs an sm_100 device");
  CUtensorMap tmA, tmB, tmV;
  int r;
  if ((r = encode_tmap_2d_f16(&tmA, a, 128, 64, 64, 128, 64))) return r;
  if ((r = encode_tmap_2d_f16(&tmB, b, 128, 64, 64, 128, 64))) return r;
  if ((r = encode_tmap_2d_f16(&tmV, v, 128, 64, 64, 128, 64))) return r;
  const int smem = 81920 + 64 + 1024;
  RN_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, tmV, (const __half*)p, out_s, out_o);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
