// geom_tc.cu -- geometry weight with the 64 -> H pair FC on tcgen05 (sm_100a), E = 64.  Two kernels, one arithmetic:
// geom_weight_tc1_kernel (one thread per pair; throughput form, used from ~600x600 pairs up) and geom_weight_tc_kernel
// (four threads per pair; latency form for the detection head's N = M = 300).  The latter:
//
// A CTA of 512 threads owns 128 consecutive (query, key) pairs (flattened n*M + m), 4 threads per pair (one per box coordinate): each evaluates its eps,
// the 8 sin/cos pairs of that coordinate, splits every value into fp16 hi + lo and writes its two 16-byte chunks of the
// pair's 64-wide row of the A operand (128 pairs x 64, SWIZZLE_128B K-major) straight into shared memory.  One elected thread then issues 12 UMMAs (M=128, N=16, K=16: A_hi.W_hi + A_lo.W_hi + A_hi.W_lo over 4 K-steps)
// into a 16-column TMEM accumulator, every thread reads 4 of its pair's 16 head values back with one tcgen05.ld, applies
// bias / max(.,1e-6) / log2 and stores them -- for a fixed head, consecutive threads are consecutive keys, so stores
// are fully coalesced.  Replaces the mma.sync form (geom.cu) on sm_100: legacy HMMA costs ~64 issue cycles per SMSP there
// (measured: 635 us of geometry at N = 3000, profiles/r01_relation_sweep_1.jsonl), while these UMMAs are ~free and the
// kernel becomes SFU/issue bound (~90 MUFU + ~700 instructions per pair).
#include "common.cuh"
#include "geom.cuh"
#include "umma.cuh"
#include <algorithm>

namespace rn {
using namespace umma;

constexpr int kGA = 16384, kGB = 2048;
constexpr int kGeomSmem = 2 * kGA + 2 * kGB + 64 + 1024;

__device__ __forceinline__ void sincos_2pi_tc(float x, float* s, float* c) {
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, x);
  r = fmaf(n, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

__device__ __forceinline__ void split2(float v0, float v1, uint32_t* hi, uint32_t* lo) {
  const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
  __half2 H = __halves2half2(h0, h1);
  __half2 L = __floats2half2_rn(v0 - __half2float(h0), v1 - __half2float(h1));
  *hi = *reinterpret_cast<uint32_t*>(&H);
  *lo = *reinterpret_cast<uint32_t*>(&L);
}

// Throughput variant: ONE thread per pair, 128-thread CTAs, ~5 CTAs per SM.  No redundant box loads / eps set-up across the
// four coordinate threads, so the instruction count per pair is ~25 % lower than the 4-threads-per-pair kernel below, which
// wins only when the problem is too small to fill the SMs (N = M = 300: 16 vs 20 us; N = M = 3000: 530 vs 720 us).
template <bool EXACT>
__global__ void __launch_bounds__(128) geom_weight_tc1_kernel(const float* __restrict__ boxes, const int* __restrict__ key_index,
                                                             int B, int N, int M, int H, GeomFreq fr,
                                                             const float* __restrict__ Wg, const float* __restrict__ bg,
                                                             float* __restrict__ out, int ldg, int log2_out,
                                                             int swap_roles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sAh = smem; uint8_t* sAl = smem + kGA; uint8_t* sBh = smem + 2 * kGA; uint8_t* sBl = sBh + kGB;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBl + kGB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  __shared__ float s_bias[16];
  const int tid = threadIdx.x, warp = tid >> 5;

  // B operand: Wg [16 heads (rows >= H zero)] x [64] as hi / lo fp16, K-major SWIZZLE_128B (16 rows x 128 B)
  {
    const int row = tid >> 3, chunk = tid & 7;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float w0 = row < H ? Wg[row * 64 + chunk * 8 + 2 * j] : 0.f, w1 = row < H ? Wg[row * 64 + chunk * 8 + 2 * j + 1] : 0.f;
      split2(w0, w1, &hi[j], &lo[j]);
    }
    *reinterpret_cast<uint4*>(sBh + sw128_offset(row, chunk)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(sBl + sw128_offset(row, chunk)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (tid < 16) s_bias[tid] = tid < H ? bg[tid] : 0.f;
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<32>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = make_idesc_f16(128, 16, false, false, false);

  float rdim[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) rdim[k] = EXACT ? fr.dim[k] : __frcp_rn(fr.dim[k]);

  const uint32_t tiles_m = (uint32_t)((M + 127) >> 7);
  const uint32_t per_b = (uint32_t)N * tiles_m, total = (uint32_t)B * per_b;       // launcher guarantees < 2^31
  uint32_t phase = 0;
  for (uint32_t item = blockIdx.x; item < total; item += gridDim.x) {
    const int b = (int)(item / per_b);
    const uint32_t it = item - (uint32_t)b * per_b;
    const int n = (int)(it / tiles_m), tm = (int)(it - (uint32_t)n * tiles_m);
    const int m = tm * 128 + tid;
    const int mc = min(m, M - 1);
    const float4 bfix = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + n];
    const float4 bvar = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + (key_index ? key_index[mc] : mc)];
    const float4 bq = swap_roles ? bvar : bfix, bk = swap_roles ? bfix : bvar;     // query box / key box
    float eps[4];
    if (EXACT) {
      pair_eps(bq, bk, eps);
    } else {
      const float wn = bq.z - bq.x + 1.f, hn = bq.w - bq.y + 1.f, wm = bk.z - bk.x + 1.f, hm = bk.w - bk.y + 1.f;
      const float dcx = 0.5f * (bq.x + bq.z) - 0.5f * (bk.x + bk.z), dcy = 0.5f * (bq.y + bq.w) - 0.5f * (bk.y + bk.w);
      eps[0] = __logf(fmaxf(fabsf(dcx * __frcp_rn(wn)), 1e-3f));
      eps[1] = __logf(fmaxf(fabsf(dcy * __frcp_rn(hn)), 1e-3f));
      eps[2] = __logf(wn * __frcp_rn(wm));
      eps[3] = __logf(hn * __frcp_rn(hm));
    }
    // row `tid` of A: [coord c][sin f0..f7 | cos f0..f7] -> chunk 2c = sins, chunk 2c+1 = coses (8 halfs each)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float a = 100.0f * eps[c];
      float sn[8], cs[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) sincos_2pi_tc(EXACT ? __fdiv_rn(a, rdim[k]) : a * rdim[k], &sn[k], &cs[k]);
      uint32_t sh[4], sl[4], ch[4], cl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split2(sn[2 * j], sn[2 * j + 1], &sh[j], &sl[j]);
        split2(cs[2 * j], cs[2 * j + 1], &ch[j], &cl[j]);
      }
      *reinterpret_cast<uint4*>(sAh + sw128_offset(tid, 2 * c)) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
      *reinterpret_cast<uint4*>(sAl + sw128_offset(tid, 2 * c)) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
      *reinterpret_cast<uint4*>(sAh + sw128_offset(tid, 2 * c + 1)) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
      *reinterpret_cast<uint4*>(sAl + sw128_offset(tid, 2 * c + 1)) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl), bh = smem_u32(sBh), bl = smem_u32(sBl);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_f16_ss(tmem_d, make_smem_desc_sw128(ah + k * 32, 16, 1024), make_smem_desc_sw128(bh + k * 32, 16, 1024), idesc, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_f16_ss(tmem_d, make_smem_desc_sw128(al + k * 32, 16, 1024), make_smem_desc_sw128(bh + k * 32, 16, 1024), idesc, 1);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_f16_ss(tmem_d, make_smem_desc_sw128(ah + k * 32, 16, 1024), make_smem_desc_sw128(bl + k * 32, 16, 1024), idesc, 1);
      mma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem_d + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    if (m < M) {
      float* o = out + (((size_t)b * H) * N + n) * ldg + m;
#pragma unroll
      for (int h = 0; h < 16; ++h)
        if (h < H) {
          const float gv = fmaxf(__uint_as_float(v[h]) + s_bias[h], 1e-6f);
          o[(size_t)h * N * ldg] = log2_out ? __log2f(gv) : gv;
        }
    }
    tc_fence_before();
    __syncthreads();          // TMEM accumulator and the A tiles are free again
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem_d);
}

template <bool EXACT, bool FLAT>
__global__ void __launch_bounds__(512, 2) geom_weight_tc_kernel(const float* __restrict__ boxes, const int* __restrict__ key_index,
                                                             int B, int N, int M, int H, GeomFreq fr,
                                                             const float* __restrict__ Wg, const float* __restrict__ bg,
                                                             float* __restrict__ out, int ldg, int log2_out,
                                                             int swap_roles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sAh = smem; uint8_t* sAl = smem + kGA; uint8_t* sBh = smem + 2 * kGA; uint8_t* sBl = sBh + kGB;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBl + kGB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  __shared__ float s_bias[16];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int pr = tid & 127, cc = tid >> 7;       // 4 threads per pair: thread (pr, cc) evaluates coordinate cc of pair pr

  // B operand: Wg [16 heads (rows >= H zero)] x [64] as hi / lo fp16, K-major SWIZZLE_128B (16 rows x 128 B)
  if (tid < 128) {
    const int row = tid >> 3, chunk = tid & 7;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float w0 = row < H ? Wg[row * 64 + chunk * 8 + 2 * j] : 0.f, w1 = row < H ? Wg[row * 64 + chunk * 8 + 2 * j + 1] : 0.f;
      split2(w0, w1, &hi[j], &lo[j]);
    }
    *reinterpret_cast<uint4*>(sBh + sw128_offset(row, chunk)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(sBl + sw128_offset(row, chunk)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (tid < 16) s_bias[tid] = tid < H ? bg[tid] : 0.f;
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<32>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = make_idesc_f16(128, 16, false, false, false);

  float rdim[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) rdim[k] = EXACT ? fr.dim[k] : __frcp_rn(fr.dim[k]);

  // work item = 128 pairs.  FLAT: 128 consecutive pairs of the flattened (n, m) index space of one problem -- no partially
  // filled key tiles (M = 300 would waste a third of its third tile); costs a per-thread divide.  !FLAT: (query n, 128-key
  // tile) -- n and the tile are CTA-uniform (uniform datapath, broadcast query-box load), used when the tile padding is
  // small: at N = M = 3000 the flat form measured 35 % slower (profiles/r01_relation_sweep_3.jsonl vs _2).
  // 32-bit index arithmetic: the launcher guarantees N*M < 2^31 and the item count < 2^31.
  const uint32_t pairs = (uint32_t)N * (uint32_t)M;
  const uint32_t tiles = FLAT ? (pairs + 127u) >> 7 : (uint32_t)N * (uint32_t)((M + 127) >> 7);
  const uint32_t total = (uint32_t)B * tiles;
  uint32_t phase = 0;
  for (uint32_t item = blockIdx.x; item < total; item += gridDim.x) {
    const int b = (int)(item / tiles);
    const uint32_t it = item - (uint32_t)b * tiles;
    int n, mc, m;
    if (FLAT) {
      const uint32_t p = it * 128u + (uint32_t)pr;
      const bool live = p < pairs;
      const uint32_t pc = live ? p : pairs - 1u;
      n = (int)(pc / (uint32_t)M); mc = (int)(pc - (uint32_t)n * (uint32_t)M);
      m = live ? mc : M;                               // m == M marks a dead lane (no store)
    } else {
      const uint32_t tiles_m = (uint32_t)((M + 127) >> 7);
      n = (int)(it / tiles_m);
      m = (int)(it - (uint32_t)n * tiles_m) * 128 + pr;
      mc = min(m, M - 1);
      m = min(m, M);
    }
    const float4 bfix = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + n];
    const float4 bvar = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + (key_index ? key_index[mc] : mc)];
    const float4 bq = swap_roles ? bvar : bfix, bk = swap_roles ? bfix : bvar;     // query box / key box
    float e;                                             // eps[cc] of SYM_REL:56-75
    {
      const float wn = bq.z - bq.x + 1.f, hn = bq.w - bq.y + 1.f, wm = bk.z - bk.x + 1.f, hm = bk.w - bk.y + 1.f;
      if (cc == 0) {
        const float dcx = 0.5f * (bq.x + bq.z) - 0.5f * (bk.x + bk.z);
        e = EXACT ? logf(fmaxf(fabsf(dcx / wn), 1e-3f)) : __logf(fmaxf(fabsf(dcx * __frcp_rn(wn)), 1e-3f));
      } else if (cc == 1) {
        const float dcy = 0.5f * (bq.y + bq.w) - 0.5f * (bk.y + bk.w);
        e = EXACT ? logf(fmaxf(fabsf(dcy / hn), 1e-3f)) : __logf(fmaxf(fabsf(dcy * __frcp_rn(hn)), 1e-3f));
      } else if (cc == 2) {
        e = EXACT ? logf(wn / wm) : __logf(wn * __frcp_rn(wm));
      } else {
        e = EXACT ? logf(hn / hm) : __logf(hn * __frcp_rn(hm));
      }
    }
    // row `pr` of A: [coord c][sin f0..f7 | cos f0..f7] -> chunk 2c = sins, chunk 2c+1 = coses (8 halfs each)
    {
      const float a = 100.0f * e;
      float sn[8], cs[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) sincos_2pi_tc(EXACT ? __fdiv_rn(a, rdim[k]) : a * rdim[k], &sn[k], &cs[k]);
      uint32_t sh[4], sl[4], ch[4], cl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split2(sn[2 * j], sn[2 * j + 1], &sh[j], &sl[j]);
        split2(cs[2 * j], cs[2 * j + 1], &ch[j], &cl[j]);
      }
      *reinterpret_cast<uint4*>(sAh + sw128_offset(pr, 2 * cc)) = make_uint4(sh[0], sh[1], sh[2], sh[3]);
      *reinterpret_cast<uint4*>(sAl + sw128_offset(pr, 2 * cc)) = make_uint4(sl[0], sl[1], sl[2], sl[3]);
      *reinterpret_cast<uint4*>(sAh + sw128_offset(pr, 2 * cc + 1)) = make_uint4(ch[0], ch[1], ch[2], ch[3]);
      *reinterpret_cast<uint4*>(sAl + sw128_offset(pr, 2 * cc + 1)) = make_uint4(cl[0], cl[1], cl[2], cl[3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl), bh = smem_u32(sBh), bl = smem_u32(sBl);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_f16_ss(tmem_d, make_smem_desc_sw128(ah + k * 32, 16, 1024), make_smem_desc_sw128(bh + k * 32, 16, 1024), idesc, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_f16_ss(tmem_d, make_smem_desc_sw128(al + k * 32, 16, 1024), make_smem_desc_sw128(bh + k * 32, 16, 1024), idesc, 1);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_f16_ss(tmem_d, make_smem_desc_sw128(ah + k * 32, 16, 1024), make_smem_desc_sw128(bl + k * 32, 16, 1024), idesc, 1);
      mma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      // warp w may read TMEM lanes 32*(w%4)..: thread (pr, cc) reads heads 4cc..4cc+3 of its own pair
      uint32_t v[4];
      tmem_ld_32x32b_x4(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + 4 * cc, v);
      tmem_ld_wait();
      if (m < M) {
        float* o = out + (((size_t)b * H) * N + n) * ldg + m;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int h = 4 * cc + j;
          if (h < H) {
            const float gv = fmaxf(__uint_as_float(v[j]) + s_bias[h], 1e-6f);
            o[(size_t)h * N * ldg] = log2_out ? __log2f(gv) : gv;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();          // TMEM accumulator and the A tiles are free again
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem_d);
}

int launch_geom_weight_tc(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H,
                          const GeomFreq& fr, const float* Wg, const float* bg, float* g, int ldg, int log2_out,
                          int swap_roles, bool exact) {
  const int sms = sm_count() > 0 ? sm_count() : 148;
  static thread_local bool configured1 = false;
  if (!configured1) {
    RN_CUDA(cudaFuncSetAttribute(geom_weight_tc1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGeomSmem));
    RN_CUDA(cudaFuncSetAttribute(geom_weight_tc1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGeomSmem));
    configured1 = true;
  }
  RN_CHECK_ARG((long long)B * N * cdiv(M, 128) < (1ll << 31), "geometry: too many pair tiles for the 32-bit item index");
  if ((long long)B * N * M >= 600ll * 600) {        // enough pairs to fill the machine: the one-thread-per-pair form
    const long long items1 = (long long)B * N * cdiv(M, 128);
    const int grid1 = (int)std::min<long long>(items1, (long long)sms * 5);
    if (exact)
      geom_weight_tc1_kernel<true><<<grid1, 128, kGeomSmem, st>>>(boxes, key_index, B, N, M, H, fr, Wg, bg, g, ldg, log2_out, swap_roles);
    else
      geom_weight_tc1_kernel<false><<<grid1, 128, kGeomSmem, st>>>(boxes, key_index, B, N, M, H, fr, Wg, bg, g, ldg, log2_out, swap_roles);
    RN_LAUNCH_CHECK();
    return RN_OK;
  }
  // flat pair tiles when padding M up to a multiple of 128 would waste more than ~8 % of the lanes
  const bool flat = (long long)cdiv(M, 128) * 128 * 100 > (long long)M * 108;
  const long long items = flat ? (long long)B * (((long long)N * M + 127) / 128) : (long long)B * N * cdiv(M, 128);
  // persistent: exactly as many CTAs as are co-resident (register-limited: 2 per SM), so the per-CTA prologue (Wg split,
  // TMEM allocation, barrier init) is paid once and no CTA waits for a second wave
  // __launch_bounds__(512, 2) keeps the kernel at <= 64 registers, so two CTAs are co-resident per SM: one computes while
  // the other waits for its UMMA / drains TMEM
  static thread_local bool configured = false;
  if (!configured) {
    RN_CUDA(cudaFuncSetAttribute(geom_weight_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGeomSmem));
    RN_CUDA(cudaFuncSetAttribute(geom_weight_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGeomSmem));
    RN_CUDA(cudaFuncSetAttribute(geom_weight_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGeomSmem));
    RN_CUDA(cudaFuncSetAttribute(geom_weight_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGeomSmem));
    configured = true;
  }
  RN_CHECK_ARG((long long)N * M < (1ll << 31) && items < (1ll << 31), "geometry: N*M = %lld exceeds the 32-bit pair index", (long long)N * M);
  const int grid = (int)std::min<long long>(items, (long long)sms * 2);
#define RN_GEOM_LAUNCH(E, F) geom_weight_tc_kernel<E, F><<<grid, 512, kGeomSmem, st>>>(boxes, key_index, B, N, M, H, fr, Wg, bg, g, ldg, log2_out, swap_roles)
  if (exact) { if (flat) RN_GEOM_LAUNCH(true, true); else RN_GEOM_LAUNCH(true, false); }
  else { if (flat) RN_GEOM_LAUNCH(false, true); else RN_GEOM_LAUNCH(false, false); }
#undef RN_GEOM_LAUNCH
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn
