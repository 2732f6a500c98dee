// relation_tc.cu -- the fused object-relation kernel on tcgen05 tensor cores (RN_PREC_F16, sm_100a).
//
// Per module:   cast X -> fp16            (1 launch; skipped when the producer already emitted fp16)
//               QKV' = X . [Wq;Wk;Wout']^T + [bq;bk;bout]   tcgen05 GEMM (gemm_tc.cu), fp16 out, one launch
//               lg   = log2 max(Wg.phi(boxes)+bg, 1e-6)     geometry kernel (geom.cu), fp32 [B,H,N,ld]
//               out  = relu(X + softmax_m(lg + q.k/sqrt(dk)) . V')   relation_attn_tc_kernel, one launch
// Wout' is the grouped 1x1 conv weight padded to 64 output columns per head, so one kernel shape (dk = dv = 64)
// serves both the detection head (dv = 64) and the learn-NMS relation (dv = 8); bout is folded into V' (softmax rows
// sum to one).
//
// relation_attn_tc_kernel: one CTA per (128-query tile, head, problem).  Warp roles: warps 0-3 = softmax/epilogue (one
// query row per thread = one TMEM lane), warp 4 = MMA issuer + TMEM allocator, warp 5 = TMA producer.  Streams 128-key
// tiles (flash style, online softmax, any M): S = Q K^T (UMMA 128x128x64, fp32 in TMEM, double buffered) -> threads
// read their row from TMEM, add the geometry term, exp2, write P (fp16, SWIZZLE_128B K-major) to smem -> O_tile = P V'
// (UMMA 128x64x128, V' MN-major straight from the projection output) -> accumulated in registers with the usual
// rescale.  K/V' tiles arrive by TMA into a 3-stage ring; nothing N x M ever goes to HBM except the geometry term.
// Algorithmic work per CTA-tile: 2*128*128*64*2 FLOP on the tensor pipe, 128*128 exp2 on the SFU.
#include "common.cuh"
#include "geom.cuh"
#include "relation.cuh"
#include "gemm_tc.cuh"
#include "umma.cuh"

namespace rn {
using namespace umma;

constexpr int kKV = 3;                       // K/V' ring stages
constexpr int kMaxTileSplits = 24;           // <= this many key tiles (M <= 3072): one CTA per (query tile, key tile) + combine --
                                             // measured faster than the streaming kernel at every sweep point (r01_relation_sweep_5)
constexpr int kQ = 16384, kKt = 16384, kVt = 16384, kPt = 32768;
constexpr int kAttnBar = kQ + kKV * (kKt + kVt) + 2 * kPt;      // 180224
constexpr int kAttnSmem = kAttnBar + 256 + 1024;

struct AttnParams {
  int N, M, H, T;                 // T = number of 128-key tiles
  const float* lg; int ldg;       // [B,H,N,ldg] log2 geometry weight
  const float* X; int ldx;        // residual source [B,N,ldx] (nullptr: none)
  float* out; int ldo;            // [B,N,ldo]
  __half* out16; int ldo16;       // optional fp16 copy of out (the next GEMM's operand: saves its cast launch)
  int dv;                         // valid output columns per head (<= 64)
  int relu;
  float scale_log2;               // log2(e)/sqrt(dk)
  // optional gather mode (learn-NMS with class-agnostic boxes): lg is ONE table over the image's rois, stored TRANSPOSED
  // [H, key roi, query roi]; row i of problem b is roi gidx[i*gs_i + b*gs_b] -- the per-class geometry is a gather, not
  // 80 recomputations
  const int* gidx; int gs_i, gs_b; int R;
  // optional per-problem skip mask (learn-NMS: problem = fg class; pruned classes -- LNMS:298-303 -- never get their
  // conditional score used): a CTA of a problem with active[b] == 0 exits before touching anything
  const int* active;
};

__global__ void __launch_bounds__(192, 1) relation_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                  const __grid_constant__ CUtensorMap tmK,
                                                                  const __grid_constant__ CUtensorMap tmV, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kQ;                          // stage s: K at s*(32K), V' at +16K
  uint8_t* sP = smem + kQ + kKV * (kKt + kVt);       // 2 buffers of 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAttnBar);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;            // [kKV]
  uint64_t* kv_empty = kv_full + kKV;      // [kKV]
  uint64_t* s_full = kv_empty + kKV;       // [2]
  uint64_t* s_free = s_full + 2;           // [2]
  uint64_t* p_full = s_free + 2;           // [2]
  uint64_t* pv_full = p_full + 2;          // [2]
  uint64_t* pv_free = pv_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;

  if (warp == 5 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKV; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 128); mbar_init(&p_full[s], 128);
      mbar_init(&pv_full[s], 1); mbar_init(&pv_free[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tPV = tmem_base + 256;          // S[2] at +0,+128 ; PV[2] at +256,+320

  if (warp == 5) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQ);
      tma_load_3d(sQ, &tmQ, q_full, h * 64, q0, b);
      for (int j = 0; j < T; ++j) {
        const int s = j % kKV;
        mbar_wait(&kv_empty[s], ((j / kKV) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], kKt + kVt);
        tma_load_3d(sKV + s * (kKt + kVt), &tmK, &kv_full[s], h * 64, j * 128, b);
        tma_load_3d(sKV + s * (kKt + kVt) + kKt, &tmV, &kv_full[s], h * 64, j * 128, b);
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, 128, false, false, false);
      const uint32_t idesc_o = make_idesc_f16(128, 64, false, false, true);
      const uint32_t aQ = smem_u32(sQ);
      auto issue_pv = [&](int i) {
        const int bf = i & 1, s = i % kKV;
        mbar_wait(&p_full[bf], (i >> 1) & 1);
        if (i >= 2) mbar_wait(&pv_free[bf], ((i - 2) >> 1) & 1);
        tc_fence_after();
        const uint32_t aP = smem_u32(sP + bf * kPt), aV = smem_u32(sKV + s * (kKt + kVt) + kKt);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mma_f16_ss(tPV + bf * 64, make_smem_desc_sw128(aP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                     make_smem_desc_sw128(aV + k * 2048, 1024, 1024), idesc_o, k > 0);
        mma_commit(&pv_full[bf]);
        mma_commit(&kv_empty[s]);
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < T; ++j) {
        const int bf = j & 1, s = j % kKV;
        mbar_wait(&kv_full[s], (j / kKV) & 1);
        if (j >= 2) mbar_wait(&s_free[bf], ((j - 2) >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sKV + s * (kKt + kVt));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tS + bf * 128, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aK + k * 32, 16, 1024),
                     idesc_s, k > 0);
        mma_commit(&s_full[bf]);
        if (j >= 1) issue_pv(j - 1);
      }
      issue_pv(T - 1);
    }
  } else {
    // ------------------------------------------------------------------------------------------ softmax + epilogue
    const int r = warp * 32 + lane;                  // row inside the tile == TMEM lane
    const int n = q0 + r;
    const bool row_ok = n < p.N;
    const uint32_t lane_base = ((uint32_t)(warp * 32) << 16);
    const float* lg_row = p.lg + (((size_t)b * p.H + h) * p.N + (row_ok ? n : 0)) * p.ldg;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;

    auto fold_pv = [&](int i, float alpha) {       // o = o*alpha + PV_i
      const int bf = i & 1;
      mbar_wait(&pv_full[bf], (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tPV + bf * 64 + c * 32 + lane_base, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) o[c * 32 + q] = fmaf(o[c * 32 + q], alpha, __uint_as_float(v[q]));
      }
      tc_fence_before();
      mbar_arrive(&pv_free[bf]);
    };

    for (int j = 0; j < T; ++j) {
      const int bf = j & 1;
      const int m0 = j * 128;
      mbar_wait(&s_full[bf], (j >> 1) & 1);
      tc_fence_after();
      // pass 1: row maximum of t = lg + s*scale over the valid keys of this tile
      float mt = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int mc = m0 + c * 32;
        if (mc >= p.M) break;                                  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + bf * 128 + c * 32 + lane_base, v);
        float gl[32];
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (mc + q < p.ldg) t4 = __ldg(reinterpret_cast<const float4*>(lg_row + mc + q));
          gl[q] = t4.x; gl[q + 1] = t4.y; gl[q + 2] = t4.z; gl[q + 3] = t4.w;
        }
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const float t = fmaf(__uint_as_float(v[q]), p.scale_log2, gl[q]);
          mt = fmaxf(mt, (mc + q < p.M) ? t : -INFINITY);
        }
      }
      const float m_new = fmaxf(m_run, mt);
      const float alpha = exp2f(m_run - m_new);                // first tile: exp2(-inf) = 0
      m_run = m_new;
      // pass 2: p = exp2(t - m), row sum, fp16 P tile into shared memory (K-major, SWIZZLE_128B)
      float lsum = 0.f;
      uint8_t* sPb = sP + bf * kPt;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int mc = m0 + c * 32;
        uint32_t pk[16];
        if (mc < p.M) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + bf * 128 + c * 32 + lane_base, v);
          float gl[32];
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mc + q < p.ldg) t4 = __ldg(reinterpret_cast<const float4*>(lg_row + mc + q));
            gl[q] = t4.x; gl[q + 1] = t4.y; gl[q + 2] = t4.z; gl[q + 3] = t4.w;
          }
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            float p0 = exp2f(fmaf(__uint_as_float(v[q]), p.scale_log2, gl[q]) - m_new);
            float p1 = exp2f(fmaf(__uint_as_float(v[q + 1]), p.scale_log2, gl[q + 1]) - m_new);
            p0 = (mc + q < p.M) ? p0 : 0.f;
            p1 = (mc + q + 1 < p.M) ? p1 : 0.f;
            lsum += p0 + p1;
            __half2 hh = __floats2half2_rn(p0, p1);
            pk[q >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) pk[q] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {                          // 4 chunks of 8 halfs
          const int kc = c * 4 + i;                            // 16-byte chunk index along the 128 keys
          *reinterpret_cast<uint4*>(sPb + (kc >> 3) * 16384 + sw128_offset(r, kc & 7)) =
              make_uint4(pk[i * 4], pk[i * 4 + 1], pk[i * 4 + 2], pk[i * 4 + 3]);
        }
      }
      l_run = fmaf(l_run, alpha, lsum);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[bf]);
      mbar_arrive(&s_free[bf]);
      if (j >= 1) fold_pv(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    fold_pv(T - 1, alpha_prev);
    // epilogue: normalise, residual, relu, store the dv valid columns of this head
    if (row_ok) {
      const float inv = 1.f / l_run;
      float* dst = p.out + ((size_t)b * p.N + n) * p.ldo + (size_t)h * p.dv;
      const float* res = p.X ? p.X + ((size_t)b * p.N + n) * p.ldx + (size_t)h * p.dv : nullptr;
      if (p.dv == 64 && (p.ldo & 3) == 0 && (!res || (p.ldx & 3) == 0)) {
#pragma unroll
        for (int q = 0; q < 64; q += 4) {
          float4 y = make_float4(o[q] * inv, o[q + 1] * inv, o[q + 2] * inv, o[q + 3] * inv);
          if (res) {
            const float4 x4 = __ldg(reinterpret_cast<const float4*>(res + q));
            y.x += x4.x; y.y += x4.y; y.z += x4.z; y.w += x4.w;
          }
          if (p.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
          *reinterpret_cast<float4*>(dst + q) = y;
          if (p.out16) {
            __half2* d16 = reinterpret_cast<__half2*>(p.out16 + ((size_t)b * p.N + n) * p.ldo16 + (size_t)h * 64 + q);
            d16[0] = __floats2half2_rn(y.x, y.y); d16[1] = __floats2half2_rn(y.z, y.w);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 64; ++q)
          if (q < p.dv) {
            float y = o[q] * inv;
            if (res) y += res[q];
            if (p.relu) y = fmaxf(y, 0.f);
            dst[q] = y;
            if (p.out16) p.out16[((size_t)b * p.N + n) * p.ldo16 + (size_t)h * p.dv + q] = __float2half_rn(y);
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// Single-key-tile variant (M <= 512: the N = 300 detection head and the n = 100 learn-NMS batch).  One CTA per
// (128-query tile, head, problem, 128-key tile): 80 KB of shared memory and 256 TMEM columns, so two CTAs share an SM,
// and at N = M = 300 the grid is 3 x 16 x 3 = 144 CTAs = one wave of the 148 SMs instead of 48 CTAs streaming 3 tiles
// each.  Each thread pulls its whole 128-float geometry row into registers BEFORE the S tile lands (one L2 latency
// instead of eight dependent ones) and makes a single pass over TMEM.  With more than one key tile the CTAs write
// un-normalised partials (O, m, l) and relation_attn_combine_kernel merges them (flash-decoding style).
constexpr int kTileBar = 16384 * 3 + 32768;            // Q, K, V', P
constexpr int kTileSmem = kTileBar + 128 + 1024;

struct TileParams {
  AttnParams a;
  int splits;                    // key tiles per query tile (gridDim.x = qtiles * splits)
  float* part_o;                 // [splits][B][H][N][64] un-normalised partial outputs
  float* part_ml;                // [splits][B][H][N][2]  (row max in log2 units, row sum)
};

// 288 threads: warps 0..7 are the softmax / epilogue warps -- TWO threads per query row (warp w and w + 4 share the TMEM lane
// quarter w % 4; the first takes key columns 0..63 of the tile, the second 64..127), warp 8 issues TMA / UMMA and owns the
// TMEM allocation.  Twice the warps per SM of the one-thread-per-row form and half the serial work per thread: the kernel is
// bound by the latency of its dependent chain (gathered geometry loads, TMEM reads, exp2), not by a pipe.
__global__ void __launch_bounds__(288, 2) relation_attn_tile_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                    const __grid_constant__ CUtensorMap tmK,
                                                                    const __grid_constant__ CUtensorMap tmV, TileParams tp) {
  extern __shared__ uint8_t smem_raw[];
  const AttnParams& p = tp.a;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sQ = smem; uint8_t* sK = smem + 16384; uint8_t* sV = smem + 32768; uint8_t* sP = smem + 49152;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTileBar);
  uint64_t* ld_full = bars; uint64_t* s_full = bars + 1; uint64_t* p_full = bars + 2; uint64_t* pv_full = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  __shared__ int s_gidx[128], s_qidx[128];
  __shared__ float s_mx[2][128], s_sum[2][128];          // per-half row max / row sum, exchanged between the two threads of a row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x / tp.splits, kt = blockIdx.x % tp.splits;
  const int q0 = qt * 128, m0 = kt * 128, h = blockIdx.y, b = blockIdx.z;
  if (p.active && !p.active[b]) return;          // uniform for the whole CTA, before any barrier / TMEM allocation
  if (p.gidx && threadIdx.x < 128) {
    const int m = m0 + threadIdx.x, nq = q0 + threadIdx.x;
    s_gidx[threadIdx.x] = m < p.M ? p.gidx[(size_t)m * p.gs_i + (size_t)b * p.gs_b] : 0;
    s_qidx[threadIdx.x] = nq < p.N ? p.gidx[(size_t)nq * p.gs_i + (size_t)b * p.gs_b] : 0;
  }

  if (warp == 8) {
    if (lane == 0) {
      prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
      mbar_init(ld_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 256); mbar_init(pv_full, 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(ld_full, 3 * 16384);
      tma_load_3d(sQ, &tmQ, ld_full, h * 64, q0, b);
      tma_load_3d(sK, &tmK, ld_full, h * 64, m0, b);
      tma_load_3d(sV, &tmV, ld_full, h * 64, m0, b);
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tPV = tmem_base + 128;

  if (warp == 8) {
    if (lane == 0) {
      mbar_wait(ld_full, 0);
      tc_fence_after();
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aP = smem_u32(sP), aV = smem_u32(sV);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_f16_ss(tS, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aK + k * 32, 16, 1024),
                   make_idesc_f16(128, 128, false, false, false), k > 0);
      mma_commit(s_full);
      mbar_wait(p_full, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        mma_f16_ss(tPV, make_smem_desc_sw128(aP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                   make_smem_desc_sw128(aV + k * 2048, 1024, 1024), make_idesc_f16(128, 64, false, false, true), k > 0);
      mma_commit(pv_full);
    }
  } else {
    const int half = warp >> 2;                          // 0: key columns 0..63 of the tile, 1: 64..127
    const int r = (warp & 3) * 32 + lane, n = q0 + r;
    const int c0 = half * 64;
    const bool row_ok = n < p.N;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
    // this thread's half of the geometry row, issued before the S tile is ready
    float t[64];
    if (p.gidx) {
      // gather from the TRANSPOSED roi-level table lgT[h][key roi][query roi]: for a given key column every lane of the
      // warp reads the same 1.2 KB table row (at its own query offset) -> a few cache lines per load instruction, and all
      // 64 loads of the thread are in flight at once
      const int ri = s_qidx[r];
      const float* col0 = p.lg + (size_t)h * p.R * p.ldg + ri;
#pragma unroll
      for (int q = 0; q < 64; ++q) t[q] = (m0 + c0 + q < p.M) ? __ldg(col0 + (size_t)s_gidx[c0 + q] * p.ldg) : -INFINITY;
    } else {
      // query-major rows, float4 per lane.  (A key-major table with one coalesced scalar load per key was tried: 64 load
      // instructions per thread instead of 16 made this section slower, 2.8 k -> 4.9 k cycles; profiles/r01_tile_trace_n300.txt)
      const float* lg_row = p.lg + (((size_t)b * p.H + h) * p.N + (row_ok ? n : 0)) * p.ldg + m0 + c0;
#pragma unroll
      for (int q = 0; q < 64; q += 4) {
        float4 t4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (m0 + c0 + q < p.M) t4 = __ldg(reinterpret_cast<const float4*>(lg_row + q));      // ldg >= M rounded up to 4
        t[q] = t4.x; t[q + 1] = t4.y; t[q + 2] = t4.z; t[q + 3] = t4.w;
      }
    }
    mbar_wait(s_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tS + c0 + c * 32 + lane_base, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const float tt = (m0 + c0 + c * 32 + q < p.M) ? fmaf(__uint_as_float(v[q]), p.scale_log2, t[c * 32 + q]) : -INFINITY;
        t[c * 32 + q] = tt;
        mx = fmaxf(mx, tt);
      }
    }
    s_mx[half][r] = mx;
    asm volatile("bar.sync 1, 256;" ::: "memory");       // the 8 softmax warps only
    mx = fmaxf(mx, s_mx[half ^ 1][r]);
    float lsum = 0.f;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {                           // 16-byte chunks of 8 keys of this half
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p0 = exp2f(t[kc * 8 + 2 * i] - mx), p1 = exp2f(t[kc * 8 + 2 * i + 1] - mx);   // exp2(-inf) = 0
        lsum += p0 + p1;
        __half2 hh = __floats2half2_rn(p0, p1);
        pk[i] = *reinterpret_cast<uint32_t*>(&hh);
      }
      *reinterpret_cast<uint4*>(sP + half * 16384 + sw128_offset(r, kc)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    s_sum[half][r] = lsum;
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(p_full);
    mbar_wait(pv_full, 0);
    tc_fence_after();
    asm volatile("bar.sync 1, 256;" ::: "memory");       // both halves' row sums are in shared memory
    lsum += s_sum[half ^ 1][r];
    // this thread reads output columns 32*half .. 32*half + 31 of its row
    float o[32];
    {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tPV + half * 32 + lane_base, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) o[q] = __uint_as_float(v[q]);
    }
    if (tp.splits > 1) {
      // partial outputs: transpose the warp's 32 rows x 32 columns through shared memory (the Q / K / V' tiles are dead
      // once O is visible) so that each store instruction writes 128 contiguous bytes of ONE row instead of 16 bytes of
      // 32 different rows (2.5 k of the 9.7 k cycles of a CTA went into that, profiles/r01_tile_trace_n300.txt)
      float* tw = reinterpret_cast<float*>(smem) + warp * (32 * 33);
#pragma unroll
      for (int q = 0; q < 32; ++q) tw[lane * 33 + q] = o[q];
      __syncwarp();
      const int row_base = q0 + (warp & 3) * 32;
      const size_t prow0 = (((size_t)kt * gridDim.z + b) * p.H + h) * p.N;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr)
        if (row_base + rr < p.N) tp.part_o[(prow0 + row_base + rr) * 64 + half * 32 + lane] = tw[rr * 33 + lane];
      if (row_ok && half == 0) reinterpret_cast<float2*>(tp.part_ml)[prow0 + n] = make_float2(mx, lsum);
    } else if (row_ok) {
      {
        const float inv = 1.f / lsum;
        const int col0 = half * 32;
        float* dst = p.out + ((size_t)b * p.N + n) * p.ldo + (size_t)h * p.dv;
        const float* res = p.X ? p.X + ((size_t)b * p.N + n) * p.ldx + (size_t)h * p.dv : nullptr;
#pragma unroll
        for (int q = 0; q < 32; ++q)
          if (col0 + q < p.dv) {
            float y = o[q] * inv;
            if (res) y += res[col0 + q];
            if (p.relu) y = fmaxf(y, 0.f);
            dst[col0 + q] = y;
            if (p.out16) p.out16[((size_t)b * p.N + n) * p.ldo16 + (size_t)h * p.dv + col0 + q] = __float2half_rn(y);
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<256>(tmem_base);
}

// merge the per-key-tile partials: one thread per (b, n, h, 4 output columns)
__global__ void __launch_bounds__(256) relation_attn_combine_kernel(TileParams tp, int B) {
  const AttnParams& p = tp.a;
  const size_t total = (size_t)B * p.N * p.H * 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = i & 15, h = (i >> 4) % p.H;
    const int n = (i / (16 * p.H)) % p.N, b = i / ((size_t)16 * p.H * p.N);
    float mx = -INFINITY;
    for (int s = 0; s < tp.splits; ++s)
      mx = fmaxf(mx, tp.part_ml[((((size_t)s * B + b) * p.H + h) * p.N + n) * 2]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float l = 0.f;
    for (int s = 0; s < tp.splits; ++s) {
      const size_t row = (((size_t)s * B + b) * p.H + h) * p.N + n;
      const float2 ml = reinterpret_cast<const float2*>(tp.part_ml)[row];
      const float w = exp2f(ml.x - mx);
      const float4 o = reinterpret_cast<const float4*>(tp.part_o + row * 64)[c4];
      acc.x = fmaf(w, o.x, acc.x); acc.y = fmaf(w, o.y, acc.y); acc.z = fmaf(w, o.z, acc.z); acc.w = fmaf(w, o.w, acc.w);
      l = fmaf(w, ml.y, l);
    }
    const float inv = 1.f / l;
    float y[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
    float* dst = p.out + ((size_t)b * p.N + n) * p.ldo + (size_t)h * p.dv;
    const float* res = p.X ? p.X + ((size_t)b * p.N + n) * p.ldx + (size_t)h * p.dv : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = c4 * 4 + j;
      if (q < p.dv) {
        float v = y[j];
        if (res) v += res[q];
        if (p.relu) v = fmaxf(v, 0.f);
        dst[q] = v;
        if (p.out16) p.out16[((size_t)b * p.N + n) * p.ldo16 + (size_t)h * p.dv + q] = __float2half_rn(v);
      }
    }
  }
}

// [Wq; Wk; Wout'] -> fp16 [3*H*64, d8] and bias [3*H*64]: every head owns 64 columns of each part; Wq / Wk rows beyond dk
// and Wout' / bout rows beyond dv are zero (zero Q/K columns leave q.k unchanged, zero V' columns are never stored)
__global__ void pack_relation_weights_kernel(const float* __restrict__ Wq, const float* __restrict__ bq,
                                             const float* __restrict__ Wk, const float* __restrict__ bk,
                                             const float* __restrict__ Wout, const float* __restrict__ bout, int d, int d8,
                                             int H, int dk, int dv, __half* __restrict__ W16, float* __restrict__ bias) {
  const int rows = 3 * H * 64;
  const size_t total = (size_t)rows * d8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = i / d8, c = i % d8;
    const int part = r / (H * 64), rr = r % (H * 64);
    float v = 0.f, bv = 0.f;
    const int hh = rr / 64, jj = rr % 64;
    if (part == 0) { if (jj < dk) { if (c < d) v = Wq[(size_t)(hh * dk + jj) * d + c]; bv = bq[hh * dk + jj]; } }
    else if (part == 1) { if (jj < dk) { if (c < d) v = Wk[(size_t)(hh * dk + jj) * d + c]; bv = bk[hh * dk + jj]; } }
    else {
      if (jj < dv) { if (c < d) v = Wout[(size_t)(hh * dv + jj) * d + c]; bv = bout[hh * dv + jj]; }
    }
    W16[i] = __float2half_rn(v);
    if (c == 0) bias[r] = bv;
  }
}

// relation_fused.cu: geometry + attention in one cooperative launch
bool relation_fused_ok(const rn_relation_desc* d);
size_t relation_fused_ws_bytes(const rn_relation_desc* d);
int relation_fused_launch(const rn_relation_desc* d, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                          const float* boxes, const int* key_index, const float* Wg, const float* bg, const float* X,
                          float* out, void* out_f16, void* wsp, size_t ws_bytes, cudaStream_t st, bool phi_lo, int chunks);
// RN_RELATION_UNFUSED=1 (or rn_relation_fused_enable(0)) keeps the round-1 decomposition (geometry kernel -> [B,H,N,M]
// table -> tile attention + combine): the A/B arm of the measurements, not a fallback -- both are tcgen05 paths
// g_fused_on: 0 = unfused, 1 = fused with phi rounded to fp16 (default), 2 = fused with the fp16 residual of phi as well
static int g_fused_on = [] {
  const char* e = getenv("RN_RELATION_UNFUSED");
  if (e && e[0] == '1') return 0;
  const char* l = getenv("RN_FUSED_PHI_LO");
  return (l && l[0] == '1') ? 2 : 1;
}();
static bool use_fused(const rn_relation_desc* d) { return g_fused_on && relation_fused_ok(d); }

// Heads wider than the kernels' native 64 columns (d_k = d_v = 64 c, c > 1: BASELINE.json configs[4] d = 1024, H = 4 has
// d_k = 256) are run as H c VIRTUAL heads of 64: packing, the projection GEMM, the Q/K/V' layout and every tensor map are
// exactly those of the (d, H c) module; only the fused kernel knows that c consecutive virtual heads share one geometry
// weight and one softmax (it accumulates their c Q.K^T chunks into the same TMEM tile and each member keeps its own 64
// columns of P.V').  Returns c (1 = native shape), 0 = not covered.
static int head_chunks(const rn_relation_desc* d) {
  if (d->H < 1 || d->dq % d->H || d->dout % d->H) return 0;
  const int dk = d->dq / d->H, dv = d->dout / d->H;
  if (dk < 1 || dv < 1) return 0;
  if (dk <= 64 && dv <= 64) return 1;
  if (dk != dv || dk % 64) return 0;
  const int c = dk / 64, team = d->H * c;
  if (team > 16 || 128 % team || (128 / team) % 8) return 0;
  return c;
}
static bool tc_shape_ok(const rn_relation_desc* d) { return head_chunks(d) > 0; }
// the (d, H c) module the kernels see
static rn_relation_desc virtual_heads(const rn_relation_desc* d) {
  rn_relation_desc v = *d;
  const int c = head_chunks(d);
  if (c > 1) v.H = d->H * c;
  return v;
}

size_t relation_tc_workspace_bytes(const rn_relation_desc* d0) {
  if (!tc_shape_ok(d0)) return 0;
  const rn_relation_desc vd = virtual_heads(d0);
  const rn_relation_desc* d = &vd;
  const size_t B = d->batch, N = d->N, M = d->M, H = d->H;
  const size_t d8 = align_up(d->d, 8), W3 = 3 * H * 64, ldg = align_up(M, 4);
  size_t t = 0;
  t += ws_slice(B * N * d8, 2);          // X fp16
  t += ws_slice(B * M * d8, 2);          // gathered keys fp16
  t += relation_tc_packed_bytes(d);       // packed weights + bias (unpacked entry point)
  t += ws_slice(B * N * W3, 2);          // QKV' fp16
  t += ws_slice(B * M * 2 * H * 64, 2);  // KV' of gathered keys
  t += gemm_tc_workspace_bytes((int)(B * N), (int)W3, (int)d8) + 512;
  if (use_fused(d)) return t + relation_fused_ws_bytes(d) + 512;
  t += ws_slice(B * H * N * ldg, 4);     // log2 geometry weight
  const size_t T = (M + 127) / 128;
  if (T > 1 && T <= kMaxTileSplits) t += ws_slice(T * B * H * N * 64, 4) + ws_slice(T * B * H * N * 2, 4);
  return t;
}

__global__ void gather_rows_f16_kernel(const __half* __restrict__ X, const int* __restrict__ idx, int N, int M, int d8,
                                       __half* __restrict__ out) {
  const int m = blockIdx.x, b = blockIdx.y;
  const uint4* src = reinterpret_cast<const uint4*>(X + ((size_t)b * N + idx[m]) * d8);
  uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * M + m) * d8);
  for (int i = threadIdx.x; i < d8 / 8; i += blockDim.x) dst[i] = src[i];
}

int launch_geom_weight_log2(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H, int E,
                            float wave_length, const float* Wg, const float* bg, float* g, int ldg);

size_t relation_tc_packed_bytes(const rn_relation_desc* d0) {
  if (!tc_shape_ok(d0)) return 0;
  const rn_relation_desc vd = virtual_heads(d0);
  const rn_relation_desc* d = &vd;
  const size_t d8 = align_up(d->d, 8), W3 = 3 * (size_t)d->H * 64;
  return ws_slice(W3 * d8, 2) + ws_slice(W3, 4);
}

int relation_tc_pack(const rn_relation_desc* d0, const float* Wq, const float* bq, const float* Wk, const float* bk,
                     const float* Wout, const float* bout, void* packed, cudaStream_t st) {
  RN_CHECK_ARG(tc_shape_ok(d0), "rn_relation_pack: shape not covered by the tcgen05 kernel (dq=%d dout=%d H=%d)", d0->dq,
               d0->dout, d0->H);
  const rn_relation_desc vd = virtual_heads(d0);
  const rn_relation_desc* d = &vd;
  const int D = d->d, H = d->H, dv = d->dout / H, d8 = (int)align_up(D, 8), W3 = 3 * H * 64;
  __half* w16 = (__half*)packed;
  float* bias = (float*)((char*)packed + ws_slice((size_t)W3 * d8, 2));
  const size_t total = (size_t)W3 * d8;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  pack_relation_weights_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(Wq, bq, Wk, bk, Wout, bout, D, d8, H,
                                                                                 d->dq / H, dv, w16, bias);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

int relation_tc_packed(const rn_relation_desc* d0, const float* X, const float* boxes, const int* key_index,
                       const void* packed, const float* Wg, const float* bg, float* out, void* wsp, size_t ws_bytes,
                       cudaStream_t st, int stage_mask, const GeomGather* gg, const void* x_f16, void* out_f16) {
  const bool do_proj = stage_mask & 1, do_geom = stage_mask & 2, do_attn = stage_mask & 4;
  RN_CHECK_ARG(is_sm100(), "RN_PREC_F16 needs an sm_100 device (tcgen05); use RN_PREC_FP32");
  RN_CHECK_ARG(tc_shape_ok(d0), "RN_PREC_F16 relation kernel supports dq/H, dout/H <= 64, or dq/H = dout/H = 64 c with H c <= 16 "
               "(got dq=%d dout=%d H=%d); use RN_PREC_FP32 for this shape", d0->dq, d0->dout, d0->H);
  const int chunks = head_chunks(d0);
  const rn_relation_desc vd = virtual_heads(d0);
  const rn_relation_desc* d = &vd;
  const int B = d->batch, N = d->N, M = d->M, D = d->d, H = d->H, dv = d->dout / H;
  const int d8 = (int)align_up(D, 8), W3 = 3 * H * 64, ldg = (int)align_up(M, 4);
  const __half* w16 = (const __half*)packed;
  const float* bias = (const float*)((const char*)packed + ws_slice((size_t)W3 * d8, 2));
  Workspace ws(wsp, ws_bytes);
  __half* x16 = ws.take<__half>((size_t)B * N * d8);
  __half* xk16 = ws.take<__half>((size_t)B * M * d8);
  __half* qkv = ws.take<__half>((size_t)B * N * W3);
  __half* kv = ws.take<__half>((size_t)B * M * 2 * H * 64);
  const bool fused = !gg && use_fused(d);
  RN_CHECK_ARG(chunks == 1 || fused, "rn_relation (F16): heads wider than 64 columns run on the fused kernel only (enable it; no "
               "gathered-geometry form)");
  // the [B,H,N,M] table exists only in the unfused decomposition (gathered mode brings its own roi-level table)
  float* lg = (fused || gg) ? reinterpret_cast<float*>(kv) : ws.take<float>((size_t)B * H * N * ldg);
  const int T = cdiv(M, 128);
  float *part_o = nullptr, *part_ml = nullptr;
  if (!fused && T > 1 && T <= kMaxTileSplits) {
    part_o = ws.take<float>((size_t)T * B * H * N * 64);
    part_ml = ws.take<float>((size_t)T * B * H * N * 2);
  }
  if (!lg || !kv || (!fused && T > 1 && T <= kMaxTileSplits && !part_ml)) { set_error("rn_relation_fwd(F16): workspace too small (%zu < %zu)", ws_bytes, relation_tc_workspace_bytes(d)); return RN_ERR_WORKSPACE; }
  void* gws = ws.base + ws.off; const size_t gws_bytes = ws.size - ws.off;
  int r;
  const bool ext_qkv = gg && gg->qkv_ext;
  if (x_f16) {          // the producer already wrote an fp16 copy of X ([B*N, d], d % 8 == 0): no cast launch
    RN_CHECK_ARG(D == d8, "rn_relation: X_f16 needs d %% 8 == 0 (d = %d)", D);
    x16 = (__half*)x_f16;
  } else if (do_proj && !ext_qkv && (r = cast_rows_f16(st, X, x16, B * N, D, d8))) return r;
  const __half *Qp, *Kp, *Vp; long long ldq, ldk; long long bq_pitch, bk_pitch;
  if (ext_qkv) {
    const __half* e = (const __half*)gg->qkv_ext;
    Qp = e; Kp = e + H * 64; Vp = e + 2 * H * 64; ldq = ldk = W3; bq_pitch = bk_pitch = (long long)N * W3;
  } else if (key_index) {
    if (do_proj) {
      gather_rows_f16_kernel<<<dim3(M, B), 128, 0, st>>>(x16, key_index, N, M, d8, xk16);
      RN_LAUNCH_CHECK();
      if ((r = gemm_tc(st, x16, d8, w16, d8, B * N, H * 64, d8, bias, 0, 0, nullptr, 0, qkv, H * 64, gws, gws_bytes))) return r;
      if ((r = gemm_tc(st, xk16, d8, w16 + (size_t)H * 64 * d8, d8, B * M, 2 * H * 64, d8, bias + H * 64, 0, 0, nullptr, 0,
                       kv, 2 * H * 64, gws, gws_bytes))) return r;
    }
    Qp = qkv; ldq = H * 64; bq_pitch = (long long)N * ldq;
    Kp = kv; Vp = kv + H * 64; ldk = 2 * H * 64; bk_pitch = (long long)M * ldk;
  } else {
    if (do_proj && (r = gemm_tc(st, x16, d8, w16, d8, B * N, W3, d8, bias, 0, 0, nullptr, 0, qkv, W3, gws, gws_bytes))) return r;
    Qp = qkv; Kp = qkv + H * 64; Vp = qkv + 2 * H * 64; ldq = ldk = W3; bq_pitch = bk_pitch = (long long)N * W3;
  }
  if (gg) RN_CHECK_ARG(!key_index && cdiv(M, 128) <= kMaxTileSplits, "gathered geometry needs M <= %d and no key_index", 128 * kMaxTileSplits);
  if (!gg && !fused && do_geom && (r = launch_geom_weight_log2(st, boxes, key_index, B, N, M, H, d->E, d->wave_length, Wg, bg, lg, ldg))) return r;
  if (!do_attn) return RN_OK;

  CUtensorMap tmQ, tmK, tmV;
  if ((r = encode_tmap_3d_f16(&tmQ, Qp, B, N, H * 64, ldq, bq_pitch, 128, 64))) return r;
  if ((r = encode_tmap_3d_f16(&tmK, Kp, B, M, H * 64, ldk, bk_pitch, 128, 64))) return r;
  if ((r = encode_tmap_3d_f16(&tmV, Vp, B, M, H * 64, ldk, bk_pitch, 128, 64))) return r;
  if (fused)           // geometry + attention: one cooperative launch, nothing N x M in HBM
    return relation_fused_launch(d, tmQ, tmK, tmV, boxes, key_index, Wg, bg, X, out, out_f16, gws, gws_bytes, st,
                                 g_fused_on == 2 && chunks == 1, chunks);
  AttnParams p;
  p.N = N; p.M = M; p.H = H; p.T = T;
  p.lg = lg; p.ldg = ldg;
  p.gidx = nullptr; p.gs_i = p.gs_b = 0; p.R = 0; p.active = nullptr;
  if (gg) p.active = gg->active;
  if (gg) { p.lg = gg->lg_table; p.ldg = gg->ld; p.R = gg->R; p.gidx = gg->idx; p.gs_i = gg->stride_i; p.gs_b = gg->stride_b; }
  p.X = d->fuse_residual_relu ? X : nullptr; p.ldx = D;
  p.out = out; p.ldo = d->dout; p.dv = dv; p.relu = d->fuse_residual_relu;
  p.out16 = (__half*)out_f16; p.ldo16 = d->dout;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)(d->dq / H));
  static thread_local bool configured = false;
  if (!configured) {
    RN_CUDA(cudaFuncSetAttribute(relation_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    RN_CUDA(cudaFuncSetAttribute(relation_attn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmem));
    configured = true;
  }
  if (T <= kMaxTileSplits) {
    TileParams tp;
    tp.a = p; tp.splits = T; tp.part_o = part_o; tp.part_ml = part_ml;
    relation_attn_tile_kernel<<<dim3(cdiv(N, 128) * T, H, B), 288, kTileSmem, st>>>(tmQ, tmK, tmV, tp);
    RN_LAUNCH_CHECK();
    if (T > 1) {
      const size_t total = (size_t)B * N * H * 16;
      size_t blocks = (total + 255) / 256;
      const size_t cap = (size_t)sm_count() * 8;
      relation_attn_combine_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(tp, B);
      RN_LAUNCH_CHECK();
    }
    return RN_OK;
  }
  relation_attn_tc_kernel<<<dim3(cdiv(N, 128), H, B), 192, kAttnSmem, st>>>(tmQ, tmK, tmV, p);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

// unpacked entry (rn_relation_fwd with RN_PREC_F16): pack the weights into the workspace, then run
int relation_tc(const rn_relation_desc* d, const float* X, const float* boxes, const int* key_index, const float* Wq,
                const float* bq, const float* Wk, const float* bk, const float* Wg, const float* bg, const float* Wout,
                const float* bout, float* out, float* softmax_out, void* wsp, size_t ws_bytes, cudaStream_t st) {
  RN_CHECK_ARG(!softmax_out, "RN_PREC_F16 relation kernel never materialises the softmax; request it with RN_PREC_FP32");
  const size_t pk = relation_tc_packed_bytes(d);
  RN_CHECK_ARG(pk > 0, "RN_PREC_F16 relation kernel supports dq/H <= 64 and dout/H <= 64 (got dq=%d dout=%d H=%d); use "
               "RN_PREC_FP32 for this shape", d->dq, d->dout, d->H);
  if (ws_bytes < pk) { set_error("rn_relation_fwd(F16): workspace too small"); return RN_ERR_WORKSPACE; }
  int r = relation_tc_pack(d, Wq, bq, Wk, bk, Wout, bout, wsp, st);
  if (r) return r;
  return relation_tc_packed(d, X, boxes, key_index, wsp, Wg, bg, out, (char*)wsp + pk, ws_bytes - pk, st, 7, nullptr);
}

}  // namespace rn

namespace rn {
// learn-NMS entry: pack the weights into the workspace, then run with the geometry gathered from a roi-level table
int relation_tc_gathered(const rn_relation_desc* d, const float* X, const GeomGather* gg, const float* Wq, const float* bq,
                         const float* Wk, const float* bk, const float* Wout, const float* bout, float* out, void* wsp,
                         size_t ws_bytes, cudaStream_t st) {
  const size_t pk = relation_tc_packed_bytes(d);
  RN_CHECK_ARG(pk > 0 && gg, "relation_tc_gathered: unsupported shape");
  if (ws_bytes < pk) { set_error("relation_tc_gathered: workspace too small"); return RN_ERR_WORKSPACE; }
  int r = relation_tc_pack(d, Wq, bq, Wk, bk, Wout, bout, wsp, st);
  if (r) return r;
  return relation_tc_packed(d, X, nullptr, nullptr, wsp, nullptr, nullptr, out, (char*)wsp + pk, ws_bytes - pk, st, 7, gg);
}
bool relation_tc_shape_ok(const rn_relation_desc* d) { return tc_shape_ok(d); }

// qkv16[(b*n + i), :] = fp16(E[idx[i*gs_i + b*gs_b], :] + Rk[i, :]);  8 columns per thread
__global__ void __launch_bounds__(256) lnms_gather_add_qkv_kernel(const float* __restrict__ E, const float* __restrict__ Rk,
                                                                  const int* __restrict__ idx, int gs_i, int gs_b, int B,
                                                                  int n, int W3, const int* __restrict__ active,
                                                                  __half* __restrict__ out) {
  const int vec = W3 / 8;
  const size_t total = (size_t)B * n * vec;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int v = t % vec;
    const size_t row = t / vec;
    const int i = row % n, b = row / n;
    if (active && !active[b]) continue;
    const float4* e = reinterpret_cast<const float4*>(E + (size_t)idx[(size_t)i * gs_i + (size_t)b * gs_b] * W3) + 2 * v;
    const float4* rk = reinterpret_cast<const float4*>(Rk + (size_t)i * W3) + 2 * v;
    const float4 e0 = __ldg(e), e1 = __ldg(e + 1), r0 = __ldg(rk), r1 = __ldg(rk + 1);
    __half2 h0 = __floats2half2_rn(e0.x + r0.x, e0.y + r0.y), h1 = __floats2half2_rn(e0.z + r0.z, e0.w + r0.w);
    __half2 h2 = __floats2half2_rn(e1.x + r1.x, e1.y + r1.y), h3 = __floats2half2_rn(e1.z + r1.z, e1.w + r1.w);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    reinterpret_cast<uint4*>(out + row * W3)[v] = u;
  }
}

size_t relation_tc_lnms_extra_bytes(const rn_relation_desc* d, int R_emb) {
  const size_t W3 = 3 * (size_t)d->H * 64, d8 = align_up(d->d, 8);
  const size_t T = (d->M + 127) / 128;
  const size_t parts = T > 1 ? ws_slice(T * d->batch * d->H * d->N * 64, 4) + ws_slice(T * d->batch * d->H * d->N * 2, 4) : 0;
  return ws_slice((size_t)R_emb * d8, 2) + ws_slice((size_t)d->N * d8, 2) + ws_slice((size_t)R_emb * W3, 4) +
         ws_slice((size_t)d->N * W3, 4) + ws_slice((size_t)d->batch * d->N * W3, 2) + parts + 1024;
}

// The weight-only half of relation_tc_lnms: packed fp16 weights and RQKV = rank_feat . W^T + b  [n, 3*H*64] fp32
size_t relation_tc_lnms_prepared_bytes(const rn_relation_desc* d) {
  return relation_tc_packed_bytes(d) + ws_slice((size_t)d->N * 3 * d->H * 64, 4);
}

int relation_tc_lnms_prepare(const rn_relation_desc* d, const float* rank_feat, const float* Wq, const float* bq,
                             const float* Wk, const float* bk, const float* Wout, const float* bout, void* prepared,
                             void* wsp, size_t ws_bytes, cudaStream_t st) {
  const size_t pk = relation_tc_packed_bytes(d);
  RN_CHECK_ARG(pk > 0, "relation_tc_lnms_prepare: unsupported shape");
  const int n = d->N, D = d->d, H = d->H, d8 = (int)align_up(D, 8), W3 = 3 * H * 64;
  char* packed = (char*)prepared;
  float* Rk = (float*)(packed + pk);
  Workspace ws(wsp, ws_bytes);
  __half* rank16 = ws.take<__half>((size_t)n * d8);
  if (!rank16) { set_error("relation_tc_lnms_prepare: workspace too small"); return RN_ERR_WORKSPACE; }
  int r;
  if ((r = relation_tc_pack(d, Wq, bq, Wk, bk, Wout, bout, packed, st))) return r;
  const __half* w16 = (const __half*)packed;
  const float* bias = (const float*)(packed + ws_slice((size_t)W3 * d8, 2));
  if ((r = cast_rows_f16(st, rank_feat, rank16, n, D, d8))) return r;
  return gemm_tc(st, rank16, d8, w16, d8, n, W3, d8, bias, 0, 0, Rk, W3, nullptr, 0, ws.base + ws.off, ws.size - ws.off);
}

int relation_tc_lnms(const rn_relation_desc* d, const float* X, const float* emb, int R_emb, const float* rank_feat,
                     const GeomGather* gg, const float* Wq, const float* bq, const float* Wk, const float* bk,
                     const float* Wout, const float* bout, float* out, void* wsp, size_t ws_bytes, cudaStream_t st,
                     const void* prepared) {
  const size_t pk = relation_tc_packed_bytes(d);
  RN_CHECK_ARG(pk > 0 && gg && gg->idx, "relation_tc_lnms: unsupported shape");
  const int B = d->batch, n = d->N, D = d->d, H = d->H, d8 = (int)align_up(D, 8), W3 = 3 * H * 64;
  Workspace ws(wsp, ws_bytes);
  char* prep_local = prepared ? nullptr : ws.take<char>(relation_tc_lnms_prepared_bytes(d));
  __half* emb16 = ws.take<__half>((size_t)R_emb * d8);
  float* E = ws.take<float>((size_t)R_emb * W3);
  __half* qkv = ws.take<__half>((size_t)B * n * W3);
  if (!qkv) { set_error("relation_tc_lnms: workspace too small"); return RN_ERR_WORKSPACE; }
  int r;
  void* gws = ws.base + ws.off; const size_t gws_bytes = ws.size - ws.off;
  if (!prepared) {      // one-shot form: build the weight-only half now
    if ((r = relation_tc_lnms_prepare(d, rank_feat, Wq, bq, Wk, bk, Wout, bout, prep_local, gws, gws_bytes, st))) return r;
    prepared = prep_local;
  }
  const char* packed = (const char*)prepared;
  const float* Rk = (const float*)(packed + pk);
  const __half* w16 = (const __half*)packed;
  if ((r = cast_rows_f16(st, emb, emb16, R_emb, D, d8))) return r;
  if ((r = gemm_tc(st, emb16, d8, w16, d8, R_emb, W3, d8, nullptr, 0, 0, E, W3, nullptr, 0, gws, gws_bytes))) return r;
  {
    const size_t total = (size_t)B * n * (W3 / 8);
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 16;
    lnms_gather_add_qkv_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(E, Rk, gg->idx, gg->stride_i,
                                                                                  gg->stride_b, B, n, W3, gg->active, qkv);
    RN_LAUNCH_CHECK();
  }
  GeomGather g2 = *gg;
  g2.qkv_ext = qkv;
  return relation_tc_packed(d, X, nullptr, nullptr, packed, nullptr, nullptr, out, gws, gws_bytes, st, 7, &g2);
}
}  // namespace rn

extern "C" int rn_relation_fused_enable(int32_t on) {
  const int prev = rn::g_fused_on;
  rn::g_fused_on = on < 0 ? 0 : (on > 2 ? 2 : on);
  return prev;
}

