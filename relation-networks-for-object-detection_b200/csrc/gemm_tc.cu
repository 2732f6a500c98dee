// gemm_tc.cu -- hand-written tcgen05 GEMM for the dense layers of the path (fp16 operands, fp32 accumulate in TMEM):
//     C[M,N] = act( A[M,K] . B[N,K]^T + bias )          A,B row-major with K contiguous ("K-major")
// Used for the Q/K/V' projections of the relation module, fc_new_1/2, cls/bbox heads, the learn-NMS projections and
// the deformable-conv GEMM.
//
// Structure (one CTA = one 128 x BN output tile, optionally one K-split): warp 4 = TMA producer (one elected lane),
// warp 5 = TMEM allocator + MMA issuer (one elected lane), warps 0-3 = epilogue (TMEM lane quarter w%4 each).
// 4-stage smem ring (A 128x64 + B BNx64 halfs per stage, SWIZZLE_128B, filled by cp.async.bulk.tensor), full/empty
// mbarriers, tcgen05.commit frees a stage and finally signals the epilogue.  Out-of-range rows/cols/K are zero-filled
// by TMA; the epilogue masks its stores.
// HBM/L2 traffic per CTA: (128 + BN) * K * 2 B in, 128*BN*4 B out; roofline: tensor pipe for large M, L2->SM
// bandwidth / latency for the M = 300 layers of this path.
#include "common.cuh"
#include "relation.cuh"
#include "umma.cuh"
#include "gemm_tc.cuh"

namespace rn {
using namespace umma;

constexpr int kBM = 128, kBK = 64;

// ring depth 4: measured (tools/gemm_bench2.py, graph replay, warm L2) the K loop of one CTA runs at 0.19 us per 24 KB
// K-block = the ~64 B/clk one SM can pull from L2, so a deeper ring (8 stages were tried) changes nothing, and 98 KB of
// shared memory leaves room for a second CTA of the concurrent stream on the same SM
template <int BN>
struct GemmSmem {
  static constexpr int kStages = 4;
  static constexpr int kA = kBM * kBK * 2;            // 16 KB
  static constexpr int kB = BN * kBK * 2;
  static constexpr int kStage = kA + kB;
  static constexpr int kBar = kStages * kStage;       // barriers live after the ring
  static constexpr int kTotal = kBar + 256 + 1024;    // + barriers + alignment slack
};

struct GemmParams {
  int M, N, K;
  int k_blocks_per_split;     // K blocks (of 64) per K-split
  int splits;                 // number of K-splits (work units = tiles * splits)
  const float* bias;          // [N] (bias_per_row == 0) or [M] (bias_per_row == 1) or nullptr
  int bias_per_row;
  int relu;
  int bf16;                           // operands are bf16 instead of fp16 (same bytes through TMA; only the instruction descriptor differs)
  float* C32; long long ldc32;        // may be nullptr
  __half* C16; long long ldc16;       // may be nullptr
  float* partial;                     // split-K partials [splits][M][N] (when gridDim.z > 1)
  // optional output segmentation (several FC layers sharing one input run as ONE GEMM): columns [seg_begin[i],
  // seg_begin[i] + seg_n[i]) go to the separate contiguous matrix seg_ptr[i] (leading dimension seg_n[i]).  Segment starts
  // are multiples of 32 (one epilogue chunk never straddles two outputs); columns between segments are padding.
  int nseg;
  int seg_begin[4], seg_n[4];
  float* seg_ptr[4];
};

// Persistent: CTA b processes work units b, b+grid, ... ; a unit = (K-split z, output tile (m, n)).  The smem ring and
// its phases run on across units, the accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of unit i
// overlaps the MMAs of unit i+1 (tfull / tempty barriers).
template <int BN>
__global__ void __launch_bounds__(192, 1) gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;      // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  __shared__ float sT[4][32][33];          // epilogue transpose buffers (segmented outputs only)
  constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = (p.K + kBK - 1) / kBK;
  const int tiles_n = (p.N + BN - 1) / BN, tiles = ((p.M + kBM - 1) / kBM) * tiles_n;
  const int units = tiles * p.splits;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int z = u / tiles, t = u % tiles;
        const int m0 = (t / tiles_n) * kBM, n0 = (t % tiles_n) * BN;
        const int kb_begin = z * p.k_blocks_per_split, kb_end = min(total_kb, kb_begin + p.k_blocks_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], S::kStage);
          uint8_t* sa = smem + s * S::kStage;
          tma_load_2d(sa, &tmA, &full[s], kb * kBK, m0);
          tma_load_2d(sa + S::kA, &tmB, &full[s], kb * kBK, n0);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBM, BN, p.bf16 != 0, false, false);
      int it = 0, li = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++li) {
        const int z = u / tiles;
        const int kb_begin = z * p.k_blocks_per_split, kb_end = min(total_kb, kb_begin + p.k_blocks_per_split);
        const int as = li & 1;
        mbar_wait(&tempty[as], ((li >> 1) & 1) ^ 1);           // epilogue has drained this accumulator stage
        tc_fence_after();
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&full[s], (it / kStages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * S::kStage), sb = sa + S::kA;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            mma_f16_ss(tmem_base + as * BN, make_smem_desc_sw128(sa + k * 32, 16, 1024),
                       make_smem_desc_sw128(sb + k * 32, 16, 1024), idesc, (kb > kb_begin) || k > 0);
          mma_commit(&empty[s]);
        }
        mma_commit(&tfull[as]);
      }
    }
  } else {
    // epilogue: thread t of warps 0..3 owns accumulator row (32*warp + lane)
    const bool split = p.splits > 1;
    int li = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++li) {
      const int z = u / tiles, t = u % tiles;
      const int m0 = (t / tiles_n) * kBM, n0 = (t % tiles_n) * BN;
      const int kb_begin = z * p.k_blocks_per_split, kb_end = min(total_kb, kb_begin + p.k_blocks_per_split);
      const int as = li & 1;
      mbar_wait(&tfull[as], (li >> 1) & 1);
      tc_fence_after();
      const int row = m0 + warp * 32 + lane;
      const uint32_t lane_base = tmem_base + as * BN + ((uint32_t)(warp * 32) << 16);
      const float rbias = (!split && p.bias && p.bias_per_row && row < p.M) ? p.bias[row] : 0.f;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        if (kb_end > kb_begin) { tmem_ld_32x32b_x32(lane_base + c * 32, v); tmem_ld_wait(); }
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        const int col0 = n0 + c * 32;
        if ((row < p.M || p.nseg) && col0 < p.N) {       // segmented stores are warp-cooperative: every lane takes part
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (split) {
            float* dst = p.partial + ((size_t)z * p.M + row) * p.N + col0;
            if (col0 + 32 <= p.N && (p.N & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < 32 && col0 + j < p.N; ++j) dst[j] = f[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float b = rbias;
              if (p.bias && !p.bias_per_row && col0 + j < p.N) b = __ldg(p.bias + col0 + j);
              f[j] += b;
              if (p.relu) f[j] = fmaxf(f[j], 0.f);
            }
            const bool full_chunk = col0 + 32 <= p.N;
            if (p.nseg) {
              // segmented outputs have arbitrary widths (81, 8, 128 ...): transpose the warp's 32x32 block through shared
              // memory so that every store instruction writes 32 consecutive columns of ONE row (the thread-per-row form
              // costs one 32-byte sector per element: measured 22 us for the 217-column head GEMM, 12 us this way)
              int sg = 0;
#pragma unroll
              for (int q = 1; q < 4; ++q) if (q < p.nseg && col0 >= p.seg_begin[q]) sg = q;
              const int off = col0 - p.seg_begin[sg], nv = p.seg_n[sg] - off;      // valid columns of this chunk
              float* tw = &sT[warp][0][0];
#pragma unroll
              for (int j = 0; j < 32; ++j) tw[lane * 33 + j] = f[j];
              __syncwarp();
              float* seg_base = p.seg_ptr[sg] + off;
              const int row0 = m0 + warp * 32;
#pragma unroll 4
              for (int rr = 0; rr < 32; ++rr)
                if (row0 + rr < p.M && lane < nv) seg_base[(size_t)(row0 + rr) * p.seg_n[sg] + lane] = tw[rr * 33 + lane];
              __syncwarp();
            } else if (p.C32) {
              float* dst = p.C32 + (size_t)row * p.ldc32 + col0;
              if (full_chunk && (p.ldc32 & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              } else {
                for (int j = 0; j < 32 && col0 + j < p.N; ++j) dst[j] = f[j];
              }
            }
            if (p.C16) {
              __half* dst = p.C16 + (size_t)row * p.ldc16 + col0;
              if (full_chunk && (p.ldc16 & 7) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  __half2 h0 = __floats2half2_rn(f[j], f[j + 1]), h1 = __floats2half2_rn(f[j + 2], f[j + 3]);
                  __half2 h2 = __floats2half2_rn(f[j + 4], f[j + 5]), h3 = __floats2half2_rn(f[j + 6], f[j + 7]);
                  uint4 uu;
                  uu.x = *reinterpret_cast<uint32_t*>(&h0); uu.y = *reinterpret_cast<uint32_t*>(&h1);
                  uu.z = *reinterpret_cast<uint32_t*>(&h2); uu.w = *reinterpret_cast<uint32_t*>(&h3);
                  *reinterpret_cast<uint4*>(dst + j) = uu;
                }
              } else {
                for (int j = 0; j < 32 && col0 + j < p.N; ++j) dst[j] = __float2half_rn(f[j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<kTmemCols>(tmem_base);
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, const float* bias,
                                     int bias_per_row, int relu, float* C32, long long ldc32, __half* C16, long long ldc16) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = i / N, c = i % N;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * total + i];
    if (bias) s += bias_per_row ? bias[r] : bias[c];
    if (relu) s = fmaxf(s, 0.f);
    if (C32) C32[(size_t)r * ldc32 + c] = s;
    if (C16) C16[(size_t)r * ldc16 + c] = __float2half_rn(s);
  }
}

__global__ void cast_f32_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(dst)[i] = u;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2half_rn(src[i]);
}

int cast_f32_f16(cudaStream_t st, const float* src, __half* dst, size_t n) {
  if (n == 0) return RN_OK;
  size_t blocks = (n / 4 + 255) / 256 + 1;
  const size_t cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  cast_f32_f16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(src, dst, n);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

size_t gemm_tc_workspace_bytes(int M, int N, int K) {
  // worst case split-K partials over both tile widths the launcher may pick
  const int kb = cdiv(K, kBK);
  const int sms = 148;
  int splits = 1;
  for (int bn = 64; bn <= 128; bn *= 2) {
    const int tiles = cdiv(M, kBM) * cdiv(N, bn);
    if (tiles < sms) splits = std::max(splits, std::min(std::max(sms / tiles, 1), std::max(kb / 4, 1)));
  }
  return splits > 1 ? ws_slice((size_t)splits * M * N, 4) : 0;
}

template <int BN>
static int launch(cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid) {
  static thread_local bool configured = false;
  if (!configured) {
    RN_CUDA(cudaFuncSetAttribute(gemm_f16_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::kTotal));
    configured = true;
  }
  gemm_f16_tc_kernel<BN><<<grid, 192, GemmSmem<BN>::kTotal, st>>>(tmA, tmB, p);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int gemm_tc(cudaStream_t st, const __half* A, long long lda, const __half* B, long long ldb, int M, int N, int K,
            const float* bias, int bias_per_row, int relu, float* C32, long long ldc32, __half* C16, long long ldc16,
            void* ws, size_t ws_bytes, const GemmSegments* seg, bool bf16) {
  RN_CHECK_ARG(is_sm100(), "tcgen05 GEMM needs an sm_100 device");
  RN_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_tc: bad sizes %dx%dx%d", M, N, K);
  RN_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0 && (((uintptr_t)A | (uintptr_t)B) & 15) == 0,
               "gemm_tc: operand pitch must be a multiple of 8 halfs and 16-byte aligned (lda=%lld ldb=%lld)", lda, ldb);
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const int tiles_m = cdiv(M, kBM);
  // BN = 64 when 128-wide tiles would leave most SMs idle -- unless K is long (fc_new_1: K = 12544): then split-K fills the
  // machine anyway and the 128-wide tile moves a third fewer operand bytes per output (per-SM L2 ingest is the bound)
  const int kb = cdiv(K, kBK);
  // (tried for the deformable-conv GEMM, 19 x 4 tiles of 128 on 76 of 148 SMs: 19 x 8 tiles of 64 = 152 units need a second
  // wave on 148 SMs and measured slower, 58 -> 67 us per layer)
  const bool bn64 = ((tiles_m * cdiv(N, 128) < sms / 2) && kb < 64) || N <= 64;
  const int BN = bn64 ? 64 : 128;
  const int tiles = tiles_m * cdiv(N, BN);
  // split K only when every split keeps >= 16 K-blocks (1024 of K): below that the separate reduce launch (~3-4 us)
  // costs more than the K loop it shortens (measured: fc_new_2 / cls_score / embedding GEMMs at M = 300, K = 1024)
  int splits = 1;
  if (tiles < sms) splits = std::min(std::max(sms / tiles, 1), std::max(kb / 16, 1));
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.k_blocks_per_split = cdiv(kb, splits);
  splits = cdiv(kb, p.k_blocks_per_split);
  p.bias = bias; p.bias_per_row = bias_per_row; p.relu = relu; p.bf16 = bf16 ? 1 : 0;
  p.C32 = C32; p.ldc32 = ldc32; p.C16 = C16; p.ldc16 = ldc16; p.partial = nullptr;
  p.nseg = 0;
  if (seg && seg->n > 0) {
    RN_CHECK_ARG(seg->n <= 4, "gemm_tc: at most 4 output segments");
    p.nseg = seg->n;
    for (int i = 0; i < seg->n; ++i) {
      RN_CHECK_ARG(seg->begin[i] % 32 == 0 && seg->ptr[i], "gemm_tc: segment %d must start at a multiple of 32 columns", i);
      p.seg_begin[i] = seg->begin[i]; p.seg_n[i] = seg->cols[i]; p.seg_ptr[i] = seg->ptr[i];
    }
    splits = 1; p.k_blocks_per_split = kb;       // the split-K reduce kernel does not know about segments
  }
  if (splits > 1) {
    const size_t need = ws_slice((size_t)splits * M * N, 4);
    if (!ws || ws_bytes < need) { splits = 1; p.k_blocks_per_split = kb; }     // no room: fall back to a single pass
    else p.partial = (float*)ws;
  }
  CUtensorMap tmA, tmB;
  int r;
  if ((r = encode_tmap_2d_f16(&tmA, A, M, K, lda, kBM, kBK))) return r;
  if ((r = encode_tmap_2d_f16(&tmB, B, N, K, ldb, BN, kBK))) return r;
  p.splits = splits;
  const int units = tiles * splits;
  dim3 grid(std::min(units, sms));                       // persistent: one CTA per SM, units round-robin
  r = bn64 ? launch<64>(st, tmA, tmB, p, grid) : launch<128>(st, tmA, tmB, p, grid);
  if (r) return r;
  if (splits > 1) {
    const size_t total = (size_t)M * N;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)sms * 8;
    splitk_reduce_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(p.partial, splits, M, N, bias, bias_per_row,
                                                                            relu, C32, ldc32, C16, ldc16);
    RN_LAUNCH_CHECK();
  }
  return RN_OK;
}

// ---------------------------------------------------------------------------------------------------- rn_linear (F16)
size_t linear_tc_workspace_bytes(int rows, int in, int out) {
  const int in8 = (int)align_up(in, 8);
  return ws_slice((size_t)rows * in8, 2) + ws_slice((size_t)out * in8, 2) + gemm_tc_workspace_bytes(rows, out, in) + 512;
}

__global__ void cast_pad_rows_kernel(const float* __restrict__ src, __half* __restrict__ dst, int rows, int cols, int ld) {
  const size_t total = (size_t)rows * ld;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = i / ld, c = i % ld;
    dst[i] = c < cols ? __float2half_rn(src[(size_t)r * cols + c]) : __float2half_rn(0.f);
  }
}

int cast_rows_f16(cudaStream_t st, const float* src, __half* dst, int rows, int cols, int ld) {
  if (ld == cols) return cast_f32_f16(st, src, dst, (size_t)rows * cols);
  const size_t total = (size_t)rows * ld;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  cast_pad_rows_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(src, dst, rows, cols, ld);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

size_t linear_tc_packed_bytes(int in, int out) { return ws_slice((size_t)out * align_up(in, 8), 2); }

int linear_tc_pack(const float* W, int in, int out, void* packed, cudaStream_t st) {
  return cast_rows_f16(st, W, (__half*)packed, out, in, (int)align_up(in, 8));
}

int linear_tc_packed(const float* x, const void* packed_W, const float* b, float* y, int rows, int in, int out, int relu,
                     void* wsp, size_t ws_bytes, cudaStream_t st) {
  const int in8 = (int)align_up(in, 8);
  Workspace ws(wsp, ws_bytes);
  __half* x16 = ws.take<__half>((size_t)rows * in8);
  if (!x16) { set_error("rn_linear_packed_fwd: workspace too small"); return RN_ERR_WORKSPACE; }
  int r;
  if ((r = cast_rows_f16(st, x, x16, rows, in, in8))) return r;
  return gemm_tc(st, x16, in8, (const __half*)packed_W, in8, rows, out, in8, b, 0, relu, y, out, nullptr, 0, ws.base + ws.off,
                 ws.size - ws.off);
}

// W [out, C*S] (MXNet FC weight over a flattened [C, S] = [channels, ph*pw] input) -> fp16 [out, S*C]: the K order of
// the channels-last pooled tensor written by rn_roi_pool_nhwc_f16_fwd
__global__ void pack_chw_to_hwc_kernel(const float* __restrict__ W, int out, int C, int S, __half* __restrict__ dst) {
  const size_t total = (size_t)out * C * S;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C, sidx = (i / C) % S;
    const size_t o = i / ((size_t)C * S);
    dst[i] = __float2half_rn(W[(o * C + c) * S + sidx]);
  }
}

int linear_tc_pack_chw_to_hwc(const float* W, int out, int C, int S, void* packed, cudaStream_t st) {
  RN_CHECK_ARG(((size_t)C * S) % 8 == 0, "rn_linear_pack_chw_to_hwc: C*S must be a multiple of 8");
  const size_t total = (size_t)out * C * S;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 16;
  pack_chw_to_hwc_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(W, out, C, S, (__half*)packed);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

// x already fp16 [rows, in] (in % 8 == 0): no cast, straight into the GEMM; optional fp16 copy of the output
int linear_tc_packed_f16in(const void* x16, const void* packed_W, const float* b, float* y, void* y16, int rows, int in,
                           int out, int relu, void* wsp, size_t ws_bytes, cudaStream_t st) {
  RN_CHECK_ARG(in % 8 == 0, "rn_linear_packed_f16in_fwd: in must be a multiple of 8");
  return gemm_tc(st, (const __half*)x16, in, (const __half*)packed_W, in, rows, out, in, b, 0, relu, y, out, (__half*)y16,
                 out, wsp, ws_bytes);
}

// ---- several FullyConnected layers over the same input as one GEMM (cls_score + bbox_pred + roi_feat_embedding, SURVEY 8f
// rank 3).  packed = fp16 [Npad, in] with layer i's rows starting at a multiple of 32 (zero rows between), then the
// fp32 bias [Npad] laid out the same way.
static int multi_layout(const int32_t* outs, int nout, int* begin) {
  int off = 0;
  for (int i = 0; i < nout; ++i) { begin[i] = off; off += (int)align_up(outs[i], 32); }
  return off;
}

size_t linear_multi_packed_bytes(const int32_t* outs, int nout, int in) {
  int begin[4];
  if (nout < 1 || nout > 4) return 0;
  const int npad = multi_layout(outs, nout, begin);
  return ws_slice((size_t)npad * in, 2) + ws_slice(npad, 4);
}

__global__ void fill_bias_kernel(const float* __restrict__ b, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = b ? b[i] : 0.f;
}

int linear_multi_pack(const float* const* W, const float* const* b, const int32_t* outs, int nout, int in, void* packed,
                      cudaStream_t st) {
  RN_CHECK_ARG(nout >= 1 && nout <= 4 && in % 8 == 0, "rn_linear_multi_pack: 1..4 layers, in %% 8 == 0");
  int begin[4];
  const int npad = multi_layout(outs, nout, begin);
  __half* w16 = (__half*)packed;
  float* bias = (float*)((char*)packed + ws_slice((size_t)npad * in, 2));
  RN_CUDA(cudaMemsetAsync(packed, 0, linear_multi_packed_bytes(outs, nout, in), st));
  for (int i = 0; i < nout; ++i) {
    RN_CHECK_ARG(W[i] && outs[i] > 0, "rn_linear_multi_pack: layer %d", i);
    int r = cast_rows_f16(st, W[i], w16 + (size_t)begin[i] * in, outs[i], in, in);
    if (r) return r;
    fill_bias_kernel<<<cdiv(outs[i], 128), 128, 0, st>>>(b ? b[i] : nullptr, outs[i], bias + begin[i]);
    RN_LAUNCH_CHECK();
  }
  return RN_OK;
}

int linear_multi_packed_f16in(const void* x16, const void* packed, float* const* ys, const int32_t* outs, int nout, int rows,
                              int in, void* wsp, size_t ws_bytes, cudaStream_t st) {
  RN_CHECK_ARG(nout >= 1 && nout <= 4 && in % 8 == 0, "rn_linear_multi_packed_f16in_fwd: 1..4 layers, in %% 8 == 0");
  GemmSegments seg;
  seg.n = nout;
  const int npad = multi_layout(outs, nout, seg.begin);
  for (int i = 0; i < nout; ++i) { seg.cols[i] = outs[i]; seg.ptr[i] = ys[i]; }
  const float* bias = (const float*)((const char*)packed + ws_slice((size_t)npad * in, 2));
  return gemm_tc(st, (const __half*)x16, in, (const __half*)packed, in, rows, npad, in, bias, 0, 0, nullptr, 0, nullptr, 0, wsp,
                 ws_bytes, &seg);
}

int linear_tc(const float* x, const float* W, const float* b, float* y, int rows, int in, int out, int relu, void* wsp,
              size_t ws_bytes, cudaStream_t st) {
  const int in8 = (int)align_up(in, 8);
  Workspace ws(wsp, ws_bytes);
  __half* x16 = ws.take<__half>((size_t)rows * in8);
  __half* w16 = ws.take<__half>((size_t)out * in8);
  if (!w16) { set_error("rn_linear_fwd(F16): workspace too small"); return RN_ERR_WORKSPACE; }
  int r;
  if ((r = cast_rows_f16(st, x, x16, rows, in, in8))) return r;
  if ((r = cast_rows_f16(st, W, w16, out, in, in8))) return r;
  return gemm_tc(st, x16, in8, w16, in8, rows, out, in8, b, 0, relu, y, out, nullptr, 0, ws.base + ws.off, ws.size - ws.off);
}

}  // namespace rn
