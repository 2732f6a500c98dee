// relation_fused.cu -- ONE launch for the whole N x M part of the object-relation module (SYM_REL:30-151 after the
// Q/K/V' projections): pair geometry eps -> sinusoid embedding phi -> 64->H pair FC -> max(.,1e-6) -> scores QK^T ->
// softmax -> P.V' -> (+residual, relu).  Nothing N x M ever reaches HBM; the only inter-SM traffic is a short-lived,
// L2-resident ring of fp16 geometry weights.
//
// Why a team of CTAs.  phi (4 log + 32 sincos per pair) is shared by all H heads, but the outputs of H heads do not
// fit one SM (H x 64 fp32 accumulator columns = 1024 > 512 TMEM columns).  So H CTAs (one per head, one CTA per SM)
// form a TEAM that walks the same (query tile, key tile) blocks in lock step.  For a 128 x 128 block, member h
//   1. PRODUCES: evaluates phi for the 128 queries x (128/H) keys of its slice, 4 threads per pair writing the fp16 hi/lo
//      rows of the A operand straight into SWIZZLE_128B shared memory; 12 tcgen05.mma (M=128 pairs, N=16 heads, K=64:
//      A_hi.W_hi + A_lo.W_hi + A_hi.W_lo, fp32-accurate) give all H heads of a key at once; g = max(x+b,1e-6) is
//      scaled by a power of two (from sum|Wg| so it sits in fp16's normal range), rounded to fp16 and written to the
//      team's ring slot as [producer][consumer head][query][key] -- coalesced 16-byte stores;
//   2. CONSUMES: waits for the team counter, reads its own head's 128 x 128 g tile (coalesced 16-byte L2 loads),
//      S = Q K^T by tcgen05 (TMA operands), p = g . exp2(s/sqrt(dk) - rowmax) -- softmax(log g + s) without a log per
//      element -- online softmax across key tiles, P (fp16, SWIZZLE_128B) -> tcgen05 P.V' (V' MN-major), O in
//      registers (16 columns per thread, 4 threads per query row).
// Team members only ever wait for teammates, the launch is cooperative (all CTAs co-resident), and every wait is for an
// event that does not depend on the waiter's own later work, so the schedule cannot deadlock.
// Key ranges: with few query tiles (N = 300) the key tiles are split over several teams; un-normalised partials
// (O, m, l) go to the workspace and the last team to finish a (query tile, head) merges them in the same launch.
#include "common.cuh"
#include "geom.cuh"
#include "relation.cuh"
#include "umma.cuh"
#include <algorithm>

namespace rn {
using namespace umma;

constexpr int kFThreads = 544;              // 16 compute warps + 1 TMA/UMMA warp
constexpr int kFSlots = 4;                  // ring depth of the g exchange (blocks in flight per team)
constexpr int kOffQ = 0, kOffK = 16384, kOffV = 49152, kOffP = 81920, kOffA = 114688, kOffB = 180224,
              kOffG = 184320, kOffBar = 217088;
static_assert(kFSlots >= 4, "a block is published two iterations before it is read: slot reuse needs a 4-deep ring");
constexpr int kFSmem = kOffBar + 256 + 1024;
#ifdef RN_AB_NO_ELECT
constexpr int kArriveA = 512;
#else
constexpr int kArriveA = 16;
#endif
constexpr int kStagePitch = 68;             // floats per row of the output staging tile (aliases the A buffers)
// MC (multi-chunk heads, d_k = d_v = 64 c): a 2-stage ring of (Q chunk | K chunk) pairs replaces the Q tile + 2 K tiles, the A
// stages shrink to one 16 KB tile each, and P / G sit next to each other so the staging tile can alias them
constexpr int kMOffQK = 0, kMOffV = 65536, kMOffP = 98304, kMOffG = 131072, kMOffA = 163840, kMOffB = 196608, kMOffBar = 200704;
static_assert(kMOffBar + 256 + 1024 <= kFSmem, "MC layout must fit the common allocation");

// Measurement build only (tools/fused_trace.py compiles a private copy of the library with -DRN_FUSED_TRACE; the product
// library has no trace code, no trace parameter and no trace symbol).
#ifdef RN_FUSED_TRACE
__device__ long long* g_fused_trace = nullptr;          // 16 stamps per CTA
#define RN_TRACE(i) do { if (threadIdx.x == 0 && g_fused_trace) g_fused_trace[blockIdx.x * 16 + (i)] = ((i) == 0 || (i) == 15) ? (long long)globaltimer_ns() : clock64(); } while (0)
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#else
#define RN_TRACE(i) do { } while (0)
#endif
// second measurement variant: per-stage stamps of the first phi round, for thread 0 (a centre-coordinate warp) and thread
// 256 (a size-coordinate warp): 32 slots per CTA
#ifdef RN_FUSED_TRACE2
__device__ long long* g_fused_trace2 = nullptr;
#define RN_T2(i) do { if ((threadIdx.x == 0 || threadIdx.x == 256) && g_fused_trace2) g_fused_trace2[blockIdx.x * 32 + (threadIdx.x == 256 ? 16 : 0) + (i)] = clock64(); } while (0)
#else
#define RN_T2(i) do { } while (0)
#endif

struct FusedParams {
  int B, N, M, H, dv;                       // H = team size = (virtual) heads of 64 columns
  int Hr, KC;                               // real heads (rows of Wg) and 64-column chunks per real head: H = Hr * KC
  int QT, T, R, Tr;                         // query tiles, key tiles, key ranges, key tiles per range
  int teams, ntasks;
  const float* boxes; const int* key_index;
  const float* Wg; const float* bg;
  float rdim[8];                            // 1 / wave_length^(k/8)              (constant bank, not registers)
  float crev[8];                            // 100 ln2 / (2 pi wave_length^(k/8)): log2(x) * crev[k] = angle in revolutions
  float crad[8];                            // 100 ln2 / wave_length^(k/8):        log2(x) * crad[k] = angle in radians
  float scale_log2;                         // log2(e) / sqrt(dk)
  const float* X; int ldx; float* out; int ldo; __half* out16; int ldo16; int relu;
  __half* gslots;                           // [teams][kFSlots][H producers][Hr heads][128 queries][128/H keys]
  unsigned* counters;                       // [teams][32]: published[16] PER MEMBER (+16 spare) ; then tickets [B*QT*H]
  float* part_o; float* part_ml;            // [R][B][H][N][64], [R][B][H][N][2]
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ void sincos_2pi_f(float x, float* s, float* c) {
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, x);
  r = fmaf(n, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}
// two fp32 values -> packed fp16 hi and the packed fp16 residual
__device__ __forceinline__ void split2_f(float v0, float v1, uint32_t* hi, uint32_t* lo) {
  const __half2 H2 = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(H2);
  const __half2 L2 = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  *hi = *reinterpret_cast<const uint32_t*>(&H2);
  *lo = *reinterpret_cast<const uint32_t*>(&L2);
}
// sin / cos of an angle given in REVOLUTIONS: exact removal of the integer part (magic-number rounding on the FMA pipe,
// no FRND on the XU pipe), then MUFU on [-pi, pi]
__device__ __forceinline__ void sincos_rev(float t, float* s, float* c) {
  const float n = (t + 12582912.0f) - 12582912.0f;            // rint(t) for |t| < 2^22
  const float y = (t - n) * 6.2831853071795865f;
  *s = __sinf(y);
  *c = __cosf(y);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void tmem_ld_32x32b_x4f(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

// LO: also feed the fp16 residual of phi to the pair FC (A_lo.W_hi), i.e. phi at ~fp32 accuracy; without it phi is
// rounded to fp16 (|err| <= 2.4e-4, the size of the reference's own float32 noise on the angles) and two key tiles share
// one A stage (half the handshakes).  W is split hi/lo in both forms.
template <bool LO, bool MC>
__global__ void __launch_bounds__(kFThreads, 1) relation_fused_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                      const __grid_constant__ CUtensorMap tmK,
                                                                      const __grid_constant__ CUtensorMap tmV,
                                                                      const FusedParams p) {
  static_assert(!(LO && MC), "the phi-residual form is not built for multi-chunk heads");
  constexpr int TPS = (LO || MC) ? 1 : 2;     // key tiles per A stage (a stage is 32 KB: hi + lo of one tile, or hi of two; MC: 16 KB)
  constexpr int NST = 8 / TPS;                // stages per round of 8 keys
  constexpr int kAStage = MC ? 16384 : 32768;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* sQ = smem + (MC ? kMOffQK : kOffQ);    // MC: ring of 2 x (Q chunk 16 KB | K chunk 16 KB)
  uint8_t* sK = smem + kOffK;               // 2 x 16 KB (unused under MC)
  uint8_t* sV = smem + (MC ? kMOffV : kOffV);     // 2 x 16 KB
  uint8_t* sP = smem + (MC ? kMOffP : kOffP);     // 32 KB: two [128 x 64-key] halves
  uint8_t* sA = smem + (MC ? kMOffA : kOffA);     // 2 stages x 32 KB (MC: 2 x 16 KB)
  uint8_t* sBh = smem + (MC ? kMOffB : kOffB); uint8_t* sBl = sBh + 2048;       // one 32-row tile: rows 0..15 = W_hi, rows 16..31 = W_lo
  uint8_t* sG = smem + (MC ? kMOffG : kOffG);     // this head's 128 x 128 fp16 g tile of the current block ([16-byte chunk][row])
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (MC ? kMOffBar : kOffBar));
  uint64_t* q_full = bars;                  // [1]
  uint64_t* k_full = bars + 1;              // [2]
  uint64_t* v_full = bars + 3;              // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* pv_full = bars + 7;
  uint64_t* a_full = bars + 8;              // [2]
  uint64_t* a_free = bars + 10;             // [2]
  uint64_t* g_full = bars + 12;
  uint64_t* g_free = bars + 13;
  uint64_t* qk_free = bars + 16;            // [2] (MC: a ring stage has been consumed by its S UMMAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  __shared__ float s_mx[4][128], s_sum[4][128];
  __shared__ float s_bias[16], s_rowabs[16];
  __shared__ __align__(16) float s_ktab[2][2][8][16];   // [round parity][size coord][key][sin f0..7 | cos f0..7] of the key boxes
  __shared__ float s_gscale;
  __shared__ int s_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, ks = 128 / H, rounds = ks >> 3;
  const int HR = p.Hr;                      // real heads: rows of Wg, consumer dimension of the g ring
  const int team = blockIdx.x / H, h = blockIdx.x % H;
  const int hc = MC ? h / p.KC : h;         // the real head whose geometry weight / softmax this member consumes
  RN_TRACE(0); RN_TRACE(1);
  // per-member progress counters (an aggregate count cannot express "EVERY member has ..."): member m has published
  // pub_ctr[m] blocks.  No "consumed" counter is needed: a member publishes block y only after it has read block y - 2, so
  // "every member has published block i" implies "every member has finished reading block i - 2", which is the block whose
  // ring slot block i + 2 overwrites (kFSlots = 4).
  unsigned* pub_ctr = p.counters + 32 * team;
  unsigned* tickets = p.counters + 32 * p.teams;
  const int ks_shift = 31 - __clz(ks);

  // ------------------------------------------------------------------------------------------------ prologue
  if (warp == 16) {
    if (lane == 0) {
      prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&a_full[i], kArriveA); mbar_init(&a_free[i], 1);
      }
      mbar_init(s_full, 1); mbar_init(p_full, 16); mbar_init(pv_full, 1); mbar_init(g_full, 1); mbar_init(g_free, 16);   // one arrival per compute WARP
      mbar_init(&qk_free[0], 1); mbar_init(&qk_free[1], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  } else {
    // per-head bound sum|Wg[h,:]| + |bg[h]| (fixed summation order: every CTA gets the same bits)
    float a = 0.f;
    if (warp < HR) a = fabsf(p.Wg[warp * 64 + lane]) + fabsf(p.Wg[warp * 64 + 32 + lane]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) s_rowabs[warp] = warp < HR ? a + fabsf(p.bg[warp]) : 0.f;
  }
  __syncthreads();
  if (tid == 0) {
    float gm = 1e-6f;
    for (int i = 0; i < 16; ++i) gm = fmaxf(gm, s_rowabs[i]);
    int e;
    frexpf(gm, &e);                          // gm = f * 2^e, f in [0.5, 1)  ->  g <= 2^e
    s_gscale = exp2f((float)(15 - e));       // g * scale <= 2^15 ; 1e-6 * scale stays a normal fp16 for gm < 512
  }
  __syncthreads();
  const float gscale = s_gscale;
  if (warp < 16) {
    // B operand of the pair FC: scale * Wg [16 head rows (>= H zero)] x [64] as fp16 hi / lo, K-major SWIZZLE_128B.  The power-of-two
    // scale that puts g into fp16's normal range is folded into the weights and the bias here (exact), not applied per element
    if (tid < 128) {
      const int row = tid >> 3, chunk = tid & 7;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w0 = row < HR ? p.Wg[row * 64 + chunk * 8 + 2 * j] * gscale : 0.f;
        const float w1 = row < HR ? p.Wg[row * 64 + chunk * 8 + 2 * j + 1] * gscale : 0.f;
        split2_f(w0, w1, &hi[j], &lo[j]);
      }
      *reinterpret_cast<uint4*>(sBh + sw128_offset(row, chunk)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(sBl + sw128_offset(row, chunk)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (tid < 16) s_bias[tid] = tid < HR ? p.bg[tid] * gscale : 0.f;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  RN_TRACE(2);                              // prologue done
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tG = tmem_base + 128, tPV = tmem_base + 384;       // S 128 | pair FC 8 keys x 32 | PV 64
  const float gfloor = 1e-6f * gscale;      // the reference's max(., 1e-6) in the scaled domain

  if (warp == 16) {
    // ============================================================================================ TMA + UMMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, 128, false, false, false);
      const uint32_t idesc_o = make_idesc_f16(128, 64, false, false, true);
      // [W_hi; W_lo] stacked as 32 B rows: ONE UMMA per K step (two N = 16 UMMAs into one accumulator were measured: the per-
      // instruction cost of an M = 128 UMMA is its A read, so doubling the count put ~1 k cycles per block on the g_full wait)
      const uint32_t idesc_g = make_idesc_f16(128, 32, false, false, false);
      const uint32_t idesc_gl = make_idesc_f16(128, 16, false, false, false);    // LO: A_lo . W_hi into the hi columns
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      const uint32_t bh = smem_u32(sBh);
      uint32_t x = 0, u = 0, rc = 0, tc = 0;                     // blocks, A stages, FC rounds, tasks so far
      auto geom_mma = [&]() {                                    // the UMMAs of one block's geometry (rounds x 8 pair tiles)
        for (int rd = 0; rd < rounds; ++rd) {
          if (rc >= 1) mbar_wait(g_free, (rc - 1) & 1);
          for (int st = 0; st < NST; ++st) {
            const uint32_t bf = u & 1;
            mbar_wait(&a_full[bf], (u >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < TPS; ++kk) {
              const uint32_t ah = smem_u32(sA + bf * kAStage) + (TPS == 2 ? kk * 16384 : 0);
              const uint32_t d = tG + (st * TPS + kk) * 32;      // per key: columns 0..15 = A.W_hi, 16..31 = A.W_lo
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_f16_ss(d, make_smem_desc_sw128(ah + k * 32, 16, 1024), make_smem_desc_sw128(bh + k * 32, 16, 1024), idesc_g, k > 0);
              if (LO) {
                const uint32_t al = ah + 16384;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma_f16_ss(d, make_smem_desc_sw128(al + k * 32, 16, 1024), make_smem_desc_sw128(bh + k * 32, 16, 1024), idesc_gl, 1);
              }
            }
            mma_commit(&a_free[bf]);
            ++u;
          }
          mma_commit(g_full);
          ++rc;
        }
      };
      auto issue_s = [&](uint32_t xb) {
        mbar_wait(&k_full[xb & 1], (xb >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + (xb & 1) * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_f16_ss(tS, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aK + k * 32, 16, 1024), idesc_s, k > 0);
        mma_commit(s_full);
      };
      if constexpr (MC) {
        // multi-chunk heads: S = sum_c Q_c K_c^T over the KC 64-column chunks of this member's real head, streamed through a
        // 2-stage ring of (Q chunk | K chunk) pairs (Q chunks are re-read from L2 per block: 16 KB each); V' / P.V' as before
        uint32_t gl = 0, gu = 0;                                   // ring chunks loaded / consumed so far (all tasks)
        const int KC = p.KC, col0 = hc * KC * 64;
        for (int task = team; task < p.ntasks; task += p.teams, ++tc) {
          const int rg = task % p.R, qt = (task / p.R) % p.QT, b = task / (p.R * p.QT);
          const int kt0 = rg * p.Tr, nT = min(p.T, kt0 + p.Tr) - kt0;
          const int q0 = qt * 128;
          const uint32_t total = (uint32_t)(nT * KC);
          uint32_t tl = 0, tu = 0;                                 // this task's chunks loaded / consumed
          if (x >= 1) mbar_wait(pv_full, (x - 1) & 1);             // the previous task has drained (V, P buffers free)
          mbar_arrive_expect_tx(&v_full[x & 1], 16384);
          tma_load_3d(sV + (x & 1) * 16384, &tmV, &v_full[x & 1], h * 64, kt0 * 128, b);
          auto load_next = [&]() {
            const uint32_t st = gl & 1;
            if (gl >= 2) mbar_wait(&qk_free[st], ((gl >> 1) - 1) & 1);
            const int blk = (int)tl / KC, c = (int)tl % KC;
            mbar_arrive_expect_tx(&k_full[st], 32768);
            tma_load_3d(sQ + st * 32768, &tmQ, &k_full[st], col0 + c * 64, q0, b);
            tma_load_3d(sQ + st * 32768 + 16384, &tmK, &k_full[st], col0 + c * 64, (kt0 + blk) * 128, b);
            ++gl; ++tl;
          };
          auto issue_s_mc = [&]() {
            for (int c = 0; c < KC; ++c) {
              while (tl < total && tl < tu + 2) load_next();
              const uint32_t st = gu & 1;
              mbar_wait(&k_full[st], (gu >> 1) & 1);
              tc_fence_after();
              const uint32_t aQc = smem_u32(sQ + st * 32768), aKc = aQc + 16384;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_f16_ss(tS, make_smem_desc_sw128(aQc + k * 32, 16, 1024), make_smem_desc_sw128(aKc + k * 32, 16, 1024), idesc_s,
                           (c > 0) || (k > 0));
              mma_commit(&qk_free[st]);
              ++gu; ++tu;
            }
            mma_commit(s_full);
          };
          issue_s_mc();
#pragma unroll 1
          for (int i = -2; i < nT; ++i) {
            if (i + 2 < nT) geom_mma();
            if (i >= 0) {
              mbar_wait(p_full, x & 1);
              tc_fence_after();
              if (i + 1 < nT) {
                mbar_arrive_expect_tx(&v_full[(x + 1) & 1], 16384);
                tma_load_3d(sV + ((x + 1) & 1) * 16384, &tmV, &v_full[(x + 1) & 1], h * 64, (kt0 + i + 1) * 128, b);
              }
              if (i + 1 < nT) issue_s_mc();
              mbar_wait(&v_full[x & 1], (x >> 1) & 1);
              tc_fence_after();
              const uint32_t aV = smem_u32(sV + (x & 1) * 16384);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                mma_f16_ss(tPV, make_smem_desc_sw128(aP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                           make_smem_desc_sw128(aV + k * 2048, 1024, 1024), idesc_o, k > 0);
              mma_commit(pv_full);
              ++x;
            }
          }
        }
      } else {
        for (int task = team; task < p.ntasks; task += p.teams, ++tc) {
          const int rg = task % p.R, qt = (task / p.R) % p.QT, b = task / (p.R * p.QT);
          const int kt0 = rg * p.Tr, nT = min(p.T, kt0 + p.Tr) - kt0;
          const int q0 = qt * 128;
          if (x >= 1) mbar_wait(pv_full, (x - 1) & 1);             // the previous task has drained (Q, K, V, P buffers free)
          mbar_arrive_expect_tx(q_full, 16384);
          tma_load_3d(sQ, &tmQ, q_full, h * 64, q0, b);
          mbar_arrive_expect_tx(&k_full[x & 1], 16384);
          tma_load_3d(sK + (x & 1) * 16384, &tmK, &k_full[x & 1], h * 64, kt0 * 128, b);
          mbar_arrive_expect_tx(&v_full[x & 1], 16384);
          tma_load_3d(sV + (x & 1) * 16384, &tmV, &v_full[x & 1], h * 64, kt0 * 128, b);
          if (nT > 1) {
            mbar_arrive_expect_tx(&k_full[(x + 1) & 1], 16384);
            tma_load_3d(sK + ((x + 1) & 1) * 16384, &tmK, &k_full[(x + 1) & 1], h * 64, (kt0 + 1) * 128, b);
          }
          mbar_wait(q_full, tc & 1);
          issue_s(x);
#pragma unroll 1
          for (int i = -2; i < nT; ++i) {
            if (i + 2 < nT) geom_mma();                            // the compute warps evaluate phi of block i + 2 first ...
            if (i >= 0) {
              mbar_wait(p_full, x & 1);                            // ... then the softmax of block x: S and the previous PV are consumed
              tc_fence_after();
              if (i + 1 < nT) {
                mbar_arrive_expect_tx(&v_full[(x + 1) & 1], 16384);
                tma_load_3d(sV + ((x + 1) & 1) * 16384, &tmV, &v_full[(x + 1) & 1], h * 64, (kt0 + i + 1) * 128, b);
              }
              if (i + 2 < nT) {
                mbar_arrive_expect_tx(&k_full[x & 1], 16384);
                tma_load_3d(sK + (x & 1) * 16384, &tmK, &k_full[x & 1], h * 64, (kt0 + i + 2) * 128, b);
              }
              if (i + 1 < nT) issue_s(x + 1);
              mbar_wait(&v_full[x & 1], (x >> 1) & 1);
              tc_fence_after();
              const uint32_t aV = smem_u32(sV + (x & 1) * 16384);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                mma_f16_ss(tPV, make_smem_desc_sw128(aP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                           make_smem_desc_sw128(aV + k * 2048, 1024, 1024), idesc_o, k > 0);
              mma_commit(pv_full);
              ++x;
            }
          }
        }
      }
    }
  } else {
    // ============================================================================================ compute warps
    const int j = warp >> 2;                                      // coordinate (geometry) / 32-key slice (softmax) / 16-col slice (O)
    const int r = (warp & 3) * 32 + lane;                         // pair row / query row == TMEM lane
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_off0 = sw128_offset(r, 2 * j), a_off1 = sw128_offset(r, 2 * j + 1);   // this thread's two chunks of an A row
    uint32_t x = 0, u = 0, rc = 0;                                // blocks consumed, A stages, FC rounds (CTA-local counts)
    uint32_t tb_base = 0;                                         // team-wide sequence number of the task's first block
    const int qbar = 3 + (warp & 3);                              // named barrier of the 4 warps that share a TMEM lane quarter

    for (int task = team; task < p.ntasks; task += p.teams) {
      const int rg = task % p.R, qt = (task / p.R) % p.QT, b = task / (p.R * p.QT);
      const int kt0 = rg * p.Tr, nT = min(p.T, kt0 + p.Tr) - kt0;
      const int q0 = qt * 128, n = q0 + r;
      const float4* boxes4 = reinterpret_cast<const float4*>(p.boxes) + (size_t)b * p.N;
      // query side of this thread's coordinate.  j = 0, 1 (centre distances, SYM_REL:56-69): eps = log max(|c_n - c_m| / s_n,
      // 1e-3) needs the pair -> MUFU per pair.  j = 2, 3 (size ratios, :70-75): eps = log s_n - log s_m is separable, so
      // sin / cos(alpha (a_n - a_m)) come from per-box tables by the angle-difference identities: no MUFU per pair.
      float qc0, qc1;                                             // j<2: centre, 1/size
      float qs[8], qcs[8];                                        // j>=2: sin / cos(100 log(size_n) / dim_k)
      {
        const float4 bq = __ldg(boxes4 + min(n, p.N - 1));
        const float wn = bq.z - bq.x + 1.f, hn = bq.w - bq.y + 1.f;
        qc0 = j == 0 ? 0.5f * (bq.x + bq.z) : 0.5f * (bq.y + bq.w);
        qc1 = __frcp_rn(j == 0 ? wn : hn);
        const float lq = log2f((j & 1) ? hn : wn);                // used by j >= 2 only (full-precision log: once per task)
#pragma unroll
        for (int k = 0; k < 8; ++k) sincos_rev(lq * p.crev[k], &qs[k], &qcs[k]);
      }

      // ---- producer, first half: phi of this member's key slice of block bi -> A stages (the UMMA warp turns them into
      // the pair FC).  Only for rounds == 1 (H = 16) can the read-back be deferred past the softmax; with more rounds the
      // TMEM accumulator of a round must be drained before the next round's UMMAs.
      // key-side value of this thread for the 8 keys of (block bi, round rd): j < 2: lane i < 8 holds the centre of key i
      // (broadcast by shuffle per tile); warps 8..11: thread (key i8, coord cc, freq k) holds the size of key i8.  Loaded
      // one round AHEAD (kpre) so that the L2 / DRAM latency of the box never sits at the head of a round.
      auto key_value = [&](int bi, int rd) -> float {
        const int mbase = (kt0 + bi) * 128 + h * ks + rd * 8;
        if (j < 2) {
          if (lane >= 8) return 0.f;
          const int m = min(mbase + lane, p.M - 1);
          const float4 bk = __ldg(boxes4 + (p.key_index ? p.key_index[m] : m));
          return j == 0 ? 0.5f * (bk.x + bk.z) : 0.5f * (bk.y + bk.w);
        }
        if (warp >= 12) return 0.f;
        const int tt = tid - 256, i8 = tt >> 4, cc = (tt >> 3) & 1;
        const int m = min(mbase + i8, p.M - 1);
        const float4 bk = __ldg(boxes4 + (p.key_index ? p.key_index[m] : m));
        return cc ? bk.w - bk.y + 1.f : bk.z - bk.x + 1.f;
      };
      float kpre = key_value(0, 0);
      auto phi_round = [&](int bi, int rd) {
        if (bi == 0) RN_T2(0);
#ifdef RN_AB_NO_PREFETCH
        const float kval = key_value(bi, rd);
#else
        const float kval = kpre;
        {
          const int nrd = rd + 1 < rounds ? rd + 1 : 0, nbi = rd + 1 < rounds ? bi : bi + 1;
          if (nbi < nT) kpre = key_value(nbi, nrd);
        }
#endif
        if (j >= 2) {
          if (warp < 12) {
            const int tt = tid - 256, i8 = tt >> 4, cc = (tt >> 3) & 1, k = tt & 7;
            const float lk = log2f(kval);
            float sk, ck;
            sincos_rev(lk * p.crev[k], &sk, &ck);
            s_ktab[rc & 1][cc][i8][k] = sk;
            s_ktab[rc & 1][cc][i8][8 + k] = ck;
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");        // warps 8..15
        }
        if (bi == 0) RN_T2(1);
#pragma unroll 1
        for (int st = 0; st < NST; ++st) {
          if (bi == 0 && st < 4) RN_T2(2 + 3 * st);
          uint32_t wh[TPS][8], wl[LO ? 8 : 1];
#pragma unroll
          for (int kk = 0; kk < TPS; ++kk) {
            const int i8 = st * TPS + kk;
            float sn[8], cs[8];
            if (j < 2) {
              const float km = __shfl_sync(0xffffffffu, kval, i8);
              const float lg = lg2_approx(fmaxf(fabsf((qc0 - km) * qc1), 1e-3f));
              if (LO) {
#pragma unroll
                for (int k = 0; k < 8; ++k) sincos_rev(lg * p.crev[k], &sn[k], &cs[k]);
              } else {
                // phi is rounded to fp16 (2.4e-4) in this form: the MUFU's own range reduction (|angle| <= 690 rad ->
                // <= 2.6e-4 rad) is at that level, and costs 2 FMA-pipe ops per frequency instead of 7
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float a = lg * p.crad[k]; sn[k] = __sinf(a); cs[k] = __cosf(a); }
              }
            } else {
              const float4* kt = reinterpret_cast<const float4*>(&s_ktab[rc & 1][j & 1][i8][0]);
              const float4 s0 = kt[0], s1 = kt[1], c0 = kt[2], c1 = kt[3];
              const float ksn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              const float kcs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
              for (int k = 0; k < 8; ++k) {                     // sin(x - y) = sx cy - cx sy ; cos(x - y) = cx cy + sx sy
                sn[k] = fmaf(qs[k], kcs[k], -qcs[k] * ksn[k]);
                cs[k] = fmaf(qcs[k], kcs[k], qs[k] * ksn[k]);
              }
            }
            if (LO) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                split2_f(sn[2 * q], sn[2 * q + 1], &wh[kk][q], &wl[q]);
                split2_f(cs[2 * q], cs[2 * q + 1], &wh[kk][4 + q], &wl[4 + q]);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                wh[kk][q] = pack_h2(sn[2 * q], sn[2 * q + 1]);
                wh[kk][4 + q] = pack_h2(cs[2 * q], cs[2 * q + 1]);
              }
            }
          }
          const uint32_t bf = u & 1;
          if (bi == 0 && st < 4) RN_T2(3 + 3 * st);
          if (u >= 2) mbar_wait(&a_free[bf], ((u >> 1) - 1) & 1);
          if (bi == 0 && st < 4) RN_T2(4 + 3 * st);
          uint8_t* As = sA + bf * kAStage;
          // row r of an A tile: [coord c][sin f0..7 | cos f0..7] -> chunk 2c = sins, chunk 2c+1 = coses
#pragma unroll
          for (int kk = 0; kk < TPS; ++kk) {
            uint8_t* Ah = As + (TPS == 2 ? kk * 16384 : 0);
            *reinterpret_cast<uint4*>(Ah + a_off0) = make_uint4(wh[kk][0], wh[kk][1], wh[kk][2], wh[kk][3]);
            *reinterpret_cast<uint4*>(Ah + a_off1) = make_uint4(wh[kk][4], wh[kk][5], wh[kk][6], wh[kk][7]);
          }
          if (LO) {
            *reinterpret_cast<uint4*>(As + 16384 + a_off0) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
            *reinterpret_cast<uint4*>(As + 16384 + a_off1) = make_uint4(wl[4], wl[5], wl[6], wl[7]);
          }
          if (bi == 0 && st < 4) RN_T2(14);                       // (overwritten per stage: last = after the final stage's stores)
          fence_proxy_async_smem();
#ifdef RN_AB_NO_ELECT
          mbar_arrive(&a_full[bf]);
#else
          __syncwarp();                                           // every lane's stores are fenced: one arrival per warp
          if (lane == 0) mbar_arrive(&a_full[bf]);
#endif
          ++u;
        }
      };
      // ---- producer, second half: the round's FC results (heads 4j..4j+3 of this thread's query row, 8 keys; hi and lo
      // partial sums) -> g = max(x + b, 1e-6) * scale -> fp16 -> the team's ring slot
      auto readback_round = [&](int bi, int rd) {
        const uint32_t tb = tb_base + bi;
        __half* slot = p.gslots + ((((size_t)team * kFSlots + (tb % kFSlots)) * H + h) * HR) * (size_t)(128 * ks);
        mbar_wait(g_full, rc & 1);
        tc_fence_after();
        float gsum[8][4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t vh[4][4], vl[4][4];
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            tmem_ld_32x32b_x4f(tG + lane_base + (half * 4 + i4) * 32 + 4 * j, vh[i4]);
            tmem_ld_32x32b_x4f(tG + lane_base + (half * 4 + i4) * 32 + 16 + 4 * j, vl[i4]);
          }
          tmem_ld_wait();
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
            for (int q = 0; q < 4; ++q) gsum[half * 4 + i4][q] = __uint_as_float(vh[i4][q]) + __uint_as_float(vl[i4][q]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(g_free);
        ++rc;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int hh = 4 * j + q;
          if (hh < HR) {
            const float bb = s_bias[hh];
            uint32_t pk[4];
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2)
              pk[i2] = pack_h2(fmaxf(gsum[2 * i2][q] + bb, gfloor), fmaxf(gsum[2 * i2 + 1][q] + bb, gfloor));
            *reinterpret_cast<uint4*>(slot + ((size_t)hh * 128 + r) * ks + rd * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      };
      auto publish = [&]() {
        bar_compute();                                            // every store of the slab is issued ...
#ifdef RN_AB_NO_RED
        if (tid == 0) { __threadfence(); atomicAdd(pub_ctr + h, 1u); }
#else
        if (tid == 0)                                             // ... and ordered before the team counter (release)
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(pub_ctr + h) : "memory");
#endif
      };

      float o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = 0.f;
      float m_run = -INFINITY, l_run = 0.f, alpha = 0.f;

      // iteration i:  wait for block i's publication -> async copy of its g tile -> phi of block i + 2 (UMMAs drain while
      // ...) -> softmax + P of block i -> read back / publish block i + 2 -> fold P.V' of block i.  Two blocks of lead hide
      // the exchange latency; i = -2, -1 only produce.  (One call site per phase keeps the code in the instruction cache.)
#pragma unroll 1
      for (int i = -2; i < nT; ++i) {
        const uint32_t tb = tb_base + i;
        if (i >= 0) {
          if (warp == 0) {                                        // EVERY teammate has published block tb
            if (lane < H) while (ld_acquire_gpu(pub_ctr + lane) < tb + 1) {}
            __syncwarp();
          }
          bar_compute();
          if (i == 0) RN_TRACE(6);                                // every teammate published block 0
          const __half* slot = p.gslots + (((size_t)team * kFSlots + (tb % kFSlots)) * H) * (size_t)HR * (128 * ks);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int kk = j * 32 + 8 * c, pp = kk >> ks_shift, within = kk & (ks - 1);
            const __half* src = slot + (((size_t)pp * HR + hc) * 128 + r) * ks + within;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sG + ((j * 4 + c) * 128 + r) * 16)), "l"(src) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        // ---------------------------------------------------------------------------------- phi of block i + 2
        const bool prod = i + 2 < nT;
        if (prod) {
          for (int rd = 0; rd < rounds; ++rd) {
            phi_round(i + 2, rd);
            if (rd + 1 < rounds) readback_round(i + 2, rd);       // H < 16: drain the accumulator between rounds
          }
          if (i == -2) RN_TRACE(3);                               // phi of the first block written
        }
        if (i >= 0) {
          // -------------------------------------------------------------------------------- softmax + P of block i
          const int m0 = (kt0 + i) * 128 + j * 32;                // first key of this thread's slice
          mbar_wait(s_full, x & 1);
          if (i == 0) RN_TRACE(7);                                // S visible
          tc_fence_after();
          uint32_t sv[32];
          tmem_ld_32x32b_x32(tS + lane_base + j * 32, sv);
          tmem_ld_wait();
          tc_fence_before();
          if (m0 + 32 > p.M) {                                    // tail tile (warp-uniform): keys >= M never win the max, get p = 0
#pragma unroll
            for (int q = 0; q < 32; ++q) if (m0 + q >= p.M) sv[q] = 0xff800000u;      // -inf
          }
          float mx = -INFINITY;                                   // row maximum of the RAW scores (scale > 0 commutes with max)
#pragma unroll
          for (int q = 0; q < 32; ++q) mx = fmaxf(mx, __uint_as_float(sv[q]));
          s_mx[j][r] = mx;
          asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");   // the 4 warps of this lane quarter
          mx = fmaxf(fmaxf(s_mx[0][r], s_mx[1][r]), fmaxf(s_mx[2][r], s_mx[3][r])) * p.scale_log2;
          const float m_new = fmaxf(m_run, mx);                   // finite: every key tile holds at least one valid key
          alpha = ex2_approx(m_run - m_new);                      // first block: 2^-inf = 0
          m_run = m_new;
          asm volatile("cp.async.wait_group 0;" ::: "memory");    // this thread's own 4 chunks of g
          if (i == 0) RN_TRACE(8);                                // g tile landed
          float lsum = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 gq = *reinterpret_cast<const uint4*>(sG + ((j * 4 + c) * 128 + r) * 16);
            const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w};
            uint32_t pk[4];
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
              const float2 g2 = __half22float2(*reinterpret_cast<const __half2*>(&gw[i2]));
              const float p0 = g2.x * ex2_approx(fmaf(__uint_as_float(sv[c * 8 + 2 * i2]), p.scale_log2, -m_new));
              const float p1 = g2.y * ex2_approx(fmaf(__uint_as_float(sv[c * 8 + 2 * i2 + 1]), p.scale_log2, -m_new));
              lsum += p0 + p1;
              pk[i2] = pack_h2(p0, p1);
            }
            *reinterpret_cast<uint4*>(sP + (j >> 1) * 16384 + sw128_offset(r, (j & 1) * 4 + c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          l_run = fmaf(l_run, alpha, lsum);
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
          if (i == 0) RN_TRACE(9);                                // P written
        }
        // ---------------------------------------------------------------------------------- publish block i + 2
        if (prod) {
          readback_round(i + 2, rounds - 1);
          if (i == -2) RN_TRACE(4);                               // FC read back, g stored
          publish();
          if (i == -2) RN_TRACE(5);                               // published
        }
        if (i < 0) continue;
        // ---------------------------------------------------------------------------------- fold O += P V'
        mbar_wait(pv_full, x & 1);
        if (i == 0) RN_TRACE(10);                                 // P.V' visible
        tc_fence_after();
        {
          uint32_t pv[16];
          tmem_ld_32x32b_x16(tPV + lane_base + j * 16, pv);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q) o[q] = fmaf(o[q], alpha, __uint_as_float(pv[q]));
        }
        tc_fence_before();
        ++x;
      }
      tb_base += nT;

      // ------------------------------------------------------------------------------------ epilogue
      s_sum[j][r] = l_run;
      bar_compute();
      const float l_tot = (s_sum[0][r] + s_sum[1][r]) + (s_sum[2][r] + s_sum[3][r]);
      // [128][kStagePitch]; the A buffers are idle here (MC: the adjacent P and G tiles, both owned by the compute warps)
      float* stage = reinterpret_cast<float*>(MC ? sP : sA);
      {
        const float sc = p.R == 1 ? 1.f / l_tot : 1.f;
#pragma unroll
        for (int q = 0; q < 16; q += 4)
          *reinterpret_cast<float4*>(stage + r * kStagePitch + j * 16 + q) = make_float4(o[q] * sc, o[q + 1] * sc, o[q + 2] * sc, o[q + 3] * sc);
      }
      if (p.R > 1 && j == 0 && n < p.N)
        reinterpret_cast<float2*>(p.part_ml)[(((size_t)rg * p.B + b) * H + h) * p.N + n] = make_float2(m_run, l_tot);
      bar_compute();
      RN_TRACE(11);                                               // tile staged
      bool finalize = p.R == 1;
      if (p.R > 1) {
        // un-normalised partial tile -> workspace (row-contiguous 256-byte segments), then the ticket
        for (int it = tid; it < 128 * 16; it += 512) {
          const int row = it >> 4, c4 = it & 15;
          if (q0 + row < p.N)
            reinterpret_cast<float4*>(p.part_o + ((((size_t)rg * p.B + b) * H + h) * p.N + q0 + row) * 64)[c4] =
                *reinterpret_cast<const float4*>(stage + row * kStagePitch + c4 * 4);
        }
        bar_compute();
        if (tid == 0) {
          __threadfence();
          const unsigned tk = atomicAdd(&tickets[((size_t)b * p.QT + qt) * H + h], 1u);
          s_last = (tk == (unsigned)(p.R - 1));
          if (s_last) __threadfence();
        }
        bar_compute();
        finalize = s_last != 0;
        RN_TRACE(12);                                             // partials stored, ticket taken
      }
      if (finalize) {
        // R > 1, last team of this (query tile, head): pull the R partial tiles through shared memory with cp.async (three
        // 32 KB tile buffers: the two A stages and the g tile; the (m, l) pairs go to the idle P buffer) so that the whole
        // merge costs ~one L2 round trip per group of three partials instead of five dependent round trips per item
        float4 acc[4]; float Mx[4], Ls[4];
        if (p.R > 1) {
          float2* sml = reinterpret_cast<float2*>(sP);            // [R][128]
          for (int it = tid; it < p.R * 128; it += 512) {
            const int sidx = it >> 7, row = it & 127;
            const int nn = min(q0 + row, p.N - 1);
            const float2* src = reinterpret_cast<const float2*>(p.part_ml) + (((size_t)sidx * p.B + b) * H + h) * p.N + nn;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(sml + it)), "l"(src) : "memory");
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) { acc[k] = make_float4(0.f, 0.f, 0.f, 0.f); Mx[k] = -INFINITY; Ls[k] = 0.f; }
          constexpr int NB = MC ? 2 : 3;                          // 32 KB tile buffers available for the merge
          for (int s0 = 0; s0 < p.R; s0 += NB) {
            const int ns = min(NB, p.R - s0);
            for (int g = 0; g < ns; ++g) {
              uint8_t* buf = MC ? (g == 0 ? sA : sG) : (g == 0 ? sA : (g == 1 ? sA + 32768 : sG));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int it = tid + k * 512, row = it >> 4, c4 = it & 15;
                const int nn = min(q0 + row, p.N - 1);
                const float* src = p.part_o + ((((size_t)(s0 + g) * p.B + b) * H + h) * p.N + nn) * 64 + c4 * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + it * 16)), "l"(src) : "memory");
              }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            bar_compute();
            for (int g = 0; g < ns; ++g) {
              const uint8_t* buf = MC ? (g == 0 ? sA : sG) : (g == 0 ? sA : (g == 1 ? sA + 32768 : sG));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int it = tid + k * 512, row = it >> 4;
                const float2 ml = sml[(s0 + g) * 128 + row];
                const float4 po = *reinterpret_cast<const float4*>(buf + it * 16);
                const float mn = fmaxf(Mx[k], ml.x);
                const float wo = ex2_approx(Mx[k] - mn), wn = ex2_approx(ml.x - mn);
                acc[k].x = fmaf(wn, po.x, acc[k].x * wo); acc[k].y = fmaf(wn, po.y, acc[k].y * wo);
                acc[k].z = fmaf(wn, po.z, acc[k].z * wo); acc[k].w = fmaf(wn, po.w, acc[k].w * wo);
                Ls[k] = fmaf(wn, ml.y, Ls[k] * wo);
                Mx[k] = mn;
              }
            }
            if (s0 + NB < p.R) bar_compute();                     // the tile buffers are refilled by the next group
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int it = tid + k * 512;
          const int row = it >> 4, c4 = it & 15;
          const int nn = q0 + row;
          if (nn >= p.N || c4 * 4 >= p.dv) continue;
          float4 y;
          if (p.R == 1) {
            y = *reinterpret_cast<const float4*>(stage + row * kStagePitch + c4 * 4);
          } else {
            const float inv = 1.f / Ls[k];
            y = make_float4(acc[k].x * inv, acc[k].y * inv, acc[k].z * inv, acc[k].w * inv);
          }
          float yy[4] = {y.x, y.y, y.z, y.w};
          float* dst = p.out + ((size_t)b * p.N + nn) * p.ldo + (size_t)h * p.dv + c4 * 4;
          const float* res = p.X ? p.X + ((size_t)b * p.N + nn) * p.ldx + (size_t)h * p.dv + c4 * 4 : nullptr;
          __half* d16 = p.out16 ? p.out16 + ((size_t)b * p.N + nn) * p.ldo16 + (size_t)h * p.dv + c4 * 4 : nullptr;
          if (c4 * 4 + 4 <= p.dv && (p.dv & 3) == 0 && (p.ldo & 3) == 0 && (!res || (p.ldx & 3) == 0) && (!d16 || (p.ldo16 & 3) == 0)) {
            if (res) {
              const float4 x4 = __ldg(reinterpret_cast<const float4*>(res));
              yy[0] += x4.x; yy[1] += x4.y; yy[2] += x4.z; yy[3] += x4.w;
            }
            if (p.relu) { yy[0] = fmaxf(yy[0], 0.f); yy[1] = fmaxf(yy[1], 0.f); yy[2] = fmaxf(yy[2], 0.f); yy[3] = fmaxf(yy[3], 0.f); }
            *reinterpret_cast<float4*>(dst) = make_float4(yy[0], yy[1], yy[2], yy[3]);
            if (d16) {
              const __half2 a2 = __floats2half2_rn(yy[0], yy[1]), b2 = __floats2half2_rn(yy[2], yy[3]);
              *reinterpret_cast<uint2*>(d16) = make_uint2(*reinterpret_cast<const uint32_t*>(&a2), *reinterpret_cast<const uint32_t*>(&b2));
            }
          } else {
            for (int q = 0; q < 4; ++q)
              if (c4 * 4 + q < p.dv) {
                float vv = yy[q];
                if (res) vv += res[q];
                if (p.relu) vv = fmaxf(vv, 0.f);
                dst[q] = vv;
                if (d16) d16[q] = __float2half_rn(vv);
              }
          }
        }
      }
      bar_compute();                                              // the staging tile aliases the next task's A buffers
      RN_TRACE(13);                                               // task done (merged / final rows stored)
    }
  }
  tc_fence_before();
  __syncthreads();
  RN_TRACE(14); RN_TRACE(15);
  if (warp == 16) tmem_dealloc<512>(tmem_base);
}

#ifdef RN_FUSED_TRACE
}  // namespace rn
extern "C" int rn_fused_trace_set(void* buf) {
  return (int)cudaMemcpyToSymbol(rn::g_fused_trace, &buf, sizeof(void*));
}
namespace rn {
#endif
#ifdef RN_FUSED_TRACE2
}  // namespace rn
extern "C" int rn_fused_trace2_set(void* buf) {
  return (int)cudaMemcpyToSymbol(rn::g_fused_trace2, &buf, sizeof(void*));
}
namespace rn {
#endif

// ------------------------------------------------------------------------------------------------------------ host
struct FusedPlan { int teams, QT, T, R, Tr, ntasks; };

// d: the module as the kernels see it (H = virtual heads of 64 columns, see relation_tc.cu:virtual_heads)
static bool fused_plan(const rn_relation_desc* d, FusedPlan* pl) {
  const int H = d->H;
  if (d->E != 64 || H < 1 || H > 16 || (128 % H) != 0 || (128 / H) % 8 != 0) return false;
  if (d->dq % H || d->dq / H > 64 || d->dout % H || d->dout / H > 64) return false;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  pl->teams = sms / H;
  if (pl->teams < 1) return false;
  pl->QT = cdiv(d->N, 128); pl->T = cdiv(d->M, 128);
  const long long qtasks = (long long)d->batch * pl->QT;
  // key ranges: split the key tiles over several teams only when that fills noticeably more of the machine (partials cost
  // a workspace round trip and the in-kernel merge)
  int bestR = 1; double best = -1.0;
  for (int R = 1; R <= std::min(pl->T, 8); ++R) {
    const int Tr = cdiv(pl->T, R);
    if ((R - 1) * Tr >= pl->T) continue;                          // would leave an empty range
    const long long tasks = qtasks * R;
    const long long waves = (tasks + pl->teams - 1) / pl->teams;
    const double eff = (double)qtasks * pl->T / ((double)waves * pl->teams * Tr);      // useful block slots / issued slots
    if (eff > best * 1.15 + 1e-9) { best = eff; bestR = R; }
  }
  pl->R = bestR; pl->Tr = cdiv(pl->T, bestR);
  pl->ntasks = (int)(qtasks * pl->R);
  pl->teams = (int)std::min<long long>(pl->teams, pl->ntasks);
  return true;
}

bool relation_fused_ok(const rn_relation_desc* d) { FusedPlan pl; return is_sm100() && fused_plan(d, &pl); }

size_t relation_fused_ws_bytes(const rn_relation_desc* d) {
  FusedPlan pl;
  if (!fused_plan(d, &pl)) return 0;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const size_t teams_max = sms / d->H;
  size_t t = ws_slice(teams_max * kFSlots * (size_t)d->H * 128 * 128, 2);                       // g ring
  t += ws_slice(32 * teams_max + (size_t)d->batch * pl.QT * d->H, 4);                           // counters + tickets
  if (pl.R > 1) t += ws_slice((size_t)pl.R * d->batch * d->H * d->N * 64, 4) + ws_slice((size_t)pl.R * d->batch * d->H * d->N * 2, 4);
  return t;
}

int relation_fused_launch(const rn_relation_desc* d, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                          const float* boxes, const int* key_index, const float* Wg, const float* bg, const float* X,
                          float* out, void* out_f16, void* wsp, size_t ws_bytes, cudaStream_t st, bool phi_lo, int chunks) {
  FusedPlan pl;
  RN_CHECK_ARG(chunks >= 1 && d->H % chunks == 0 && !(phi_lo && chunks > 1), "relation_fused: bad head chunking %d for H=%d", chunks, d->H);
  RN_CHECK_ARG(fused_plan(d, &pl), "relation_fused: shape not covered (H=%d E=%d dq=%d dout=%d)", d->H, d->E, d->dq, d->dout);
  const int H = d->H, sms = sm_count() > 0 ? sm_count() : 148;
  const size_t teams_max = sms / H;
  Workspace ws(wsp, ws_bytes);
  __half* gslots = ws.take<__half>(teams_max * kFSlots * (size_t)H * 128 * 128);
  const size_t nctr = 32 * teams_max + (size_t)d->batch * pl.QT * H;
  unsigned* counters = ws.take<unsigned>(nctr);
  float *part_o = nullptr, *part_ml = nullptr;
  if (pl.R > 1) {
    part_o = ws.take<float>((size_t)pl.R * d->batch * H * d->N * 64);
    part_ml = ws.take<float>((size_t)pl.R * d->batch * H * d->N * 2);
  }
  if (!counters || (pl.R > 1 && !part_ml)) { set_error("relation_fused: workspace too small (%zu < %zu)", ws_bytes, relation_fused_ws_bytes(d)); return RN_ERR_WORKSPACE; }
  RN_CUDA(cudaMemsetAsync(counters, 0, nctr * sizeof(unsigned), st));
  FusedParams p;
  p.B = d->batch; p.N = d->N; p.M = d->M; p.H = H; p.dv = d->dout / H;
  p.Hr = H / chunks; p.KC = chunks;
  p.QT = pl.QT; p.T = pl.T; p.R = pl.R; p.Tr = pl.Tr; p.teams = pl.teams; p.ntasks = pl.ntasks;
  p.boxes = boxes; p.key_index = key_index; p.Wg = Wg; p.bg = bg;
  GeomFreq fr;
  int r = make_freq(d->E, d->wave_length, &fr);
  if (r) return r;
  for (int k = 0; k < 8; ++k) {
    p.rdim[k] = 1.0f / fr.dim[k];
    p.crev[k] = (float)(100.0 * 0.6931471805599453 / (6.283185307179586 * (double)fr.dim[k]));
    p.crad[k] = (float)(100.0 * 0.6931471805599453 / (double)fr.dim[k]);
  }
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)(d->dq / p.Hr));           // the REAL head width d_k
  p.X = d->fuse_residual_relu ? X : nullptr; p.ldx = d->d;
  p.out = out; p.ldo = d->dout; p.out16 = (__half*)out_f16; p.ldo16 = d->dout; p.relu = d->fuse_residual_relu;
  p.gslots = gslots; p.counters = counters; p.part_o = part_o; p.part_ml = part_ml;
  static thread_local bool configured = false;
  if (!configured) {
    RN_CUDA(cudaFuncSetAttribute(relation_fused_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    RN_CUDA(cudaFuncSetAttribute(relation_fused_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    RN_CUDA(cudaFuncSetAttribute(relation_fused_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pl.teams * H));
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = kFSmem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;          // all CTAs co-resident: team members wait for each other
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (chunks > 1) RN_CUDA(cudaLaunchKernelEx(&cfg, relation_fused_kernel<false, true>, tmQ, tmK, tmV, p));
  else if (phi_lo) RN_CUDA(cudaLaunchKernelEx(&cfg, relation_fused_kernel<true, false>, tmQ, tmK, tmV, p));
  else RN_CUDA(cudaLaunchKernelEx(&cfg, relation_fused_kernel<false, false>, tmQ, tmK, tmV, p));
  return RN_OK;
}

}  // namespace rn
