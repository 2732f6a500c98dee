// geom.cu -- pairwise box geometry: eps (SYM_REL:47-83), sinusoid embedding (SYM_REL:30-44) and the geometry weight
// g = max(relu(Wg.phi + bg), 1e-6) (SYM_REL:107-116,139) WITHOUT ever materialising the [N,M,64] embedding.
//
// HBM layout:  boxes [B,N,4] fp32 in;  g [B,H,N,ldg] fp32 out (ldg >= M, rows padded so 16-byte row loads stay aligned).
// Roofline: transcendental (MUFU/FMA) bound -- per pair 4 logf + (E/2) sincosf + E*H FMA; algorithmic bytes are only
// the g write (4*B*H*N*M) + boxes.
#include "common.cuh"
#include "geom.cuh"
#include <algorithm>

namespace rn {

int launch_geom_weight_tc(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H,
                          const GeomFreq& fr, const float* Wg, const float* bg, float* g, int ldg, int log2_out,
                          int swap_roles, bool exact);

// one thread = one (query n, key m) pair; one block = 128 keys of one query row
template <int MAXH>
__global__ void __launch_bounds__(128) geom_weight_kernel(const float* __restrict__ boxes, const int* __restrict__ key_index,
                                                          int N, int M, int H, int E, GeomFreq fr,
                                                          const float* __restrict__ Wg, const float* __restrict__ bg,
                                                          float* __restrict__ g, int ldg, int log2_out) {
  extern __shared__ float wg_s[];   // [E][MAXH] transposed, then bias [MAXH]
  const int b = blockIdx.z, n = blockIdx.y;
  const int m = blockIdx.x * 128 + threadIdx.x;
  for (int i = threadIdx.x; i < E * MAXH; i += 128) {
    int e = i / MAXH, h = i % MAXH;
    wg_s[i] = h < H ? Wg[h * E + e] : 0.f;
  }
  if (threadIdx.x < MAXH) wg_s[E * MAXH + threadIdx.x] = threadIdx.x < H ? bg[threadIdx.x] : 0.f;
  __syncthreads();
  if (m >= M) return;
  const float4 bn = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + n];
  const int mi = key_index ? key_index[m] : m;
  const float4 bm = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + mi];
  float eps[4];
  pair_eps(bn, bm, eps);
  float acc[MAXH];
#pragma unroll
  for (int h = 0; h < MAXH; ++h) acc[h] = wg_s[E * MAXH + h];
  const int nf = E / 8;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float a = 100.0f * eps[c];
    for (int k = 0; k < nf; ++k) {
      float s, co;
      sincosf(a / fr.dim[k], &s, &co);
      const float* ws = wg_s + (c * 2 * nf + k) * MAXH;
      const float* wc = wg_s + (c * 2 * nf + nf + k) * MAXH;
#pragma unroll
      for (int h = 0; h < MAXH; h += 4) {
        float4 a4 = *reinterpret_cast<const float4*>(ws + h);
        float4 c4 = *reinterpret_cast<const float4*>(wc + h);
        acc[h + 0] = fmaf(a4.x, s, acc[h + 0]); acc[h + 1] = fmaf(a4.y, s, acc[h + 1]);
        acc[h + 2] = fmaf(a4.z, s, acc[h + 2]); acc[h + 3] = fmaf(a4.w, s, acc[h + 3]);
        acc[h + 0] = fmaf(c4.x, co, acc[h + 0]); acc[h + 1] = fmaf(c4.y, co, acc[h + 1]);
        acc[h + 2] = fmaf(c4.z, co, acc[h + 2]); acc[h + 3] = fmaf(c4.w, co, acc[h + 3]);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < MAXH; ++h)
    if (h < H) {
      const float gv = fmaxf(acc[h], 1e-6f);                                   // max(relu(x),1e-6) == max(x,1e-6)
      g[(((size_t)b * H + h) * N + n) * ldg + m] = log2_out ? log2f(gv) : gv;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// E = 64 fast path: the 64 -> H pair FC runs on mma.sync (m16n8k16, fp16 x fp16 -> fp32) with the A fragment produced
// IN REGISTERS: a warp owns 16 (query, key) pairs; for coordinate c the 16 embedding values [sin f0..f7, cos f0..f7] of
// a pair are exactly one K=16 step, and lane (g = lane/4, q = lane%4) needs sin/cos of frequencies 2q, 2q+1 for the
// pairs g and g+8 -- i.e. every sin/cos is computed once, by the lane whose fragment slot it fills.  phi and Wg are
// split into fp16 hi + lo parts (3 MMAs per step: hi*hi + lo*hi + hi*lo), so the result is fp32-accurate (~1e-6).
// sin/cos: Cody-Waite reduction by 2*pi in two FMAs, then MUFU.SIN/COS on [-pi, pi] (|err| < 5e-7); arguments reach
// +-690 rad.  The SIMT kernel above (libdevice sincosf + 1024 FMA per pair) measured 23 us at N = M = 300 and 145 us on
// the learn-NMS batch (profiles/r01_launches_hot_v1.csv); this form does ~1/50 of the instructions.
__device__ __forceinline__ void sincos_2pi(float x, float* s, float* c) {
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, x);
  r = fmaf(n, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

__device__ __forceinline__ void split_h2(float v0, float v1, uint32_t* hi, uint32_t* lo) {
  const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
  const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
  __half2 H = __halves2half2(h0, h1), L = __halves2half2(l0, l1);
  *hi = *reinterpret_cast<uint32_t*>(&H);
  *lo = *reinterpret_cast<uint32_t*>(&L);
}

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// persistent: grid ~ 8 CTAs/SM x 4 warps; work item = (problem b, query n, 16-key tile), consecutive warps take
// consecutive key tiles of the same query so their stores land next to each other.
// EXACT = true keeps the reference's IEEE divisions (|dcx|/w, w_n/w_m, a/dim: the fp32 parity mode); EXACT = false
// multiplies by IEEE reciprocals instead (<= 1 ulp on eps and on the sin/cos argument; changes the module output by
// ~2e-6 relative, measured on the oracle) and is what the tcgen05 path uses.
template <int NHALF, bool EXACT>     // NHALF: number of 8-head halves (1: H <= 8, 2: H <= 16)
__global__ void __launch_bounds__(128) geom_weight_mma_kernel(const float* __restrict__ boxes,
                                                              const int* __restrict__ key_index, int B, int N, int M,
                                                              int H, GeomFreq fr, const float* __restrict__ Wg,
                                                              const float* __restrict__ bg, float* __restrict__ out,
                                                              int ldg, int log2_out, int swap_roles) {
  // swap_roles: out[h][n][m] = weight(query = m, key = n), i.e. the transposed table, still written with m contiguous
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  // B fragments of Wg^T: element (k, col) = Wg[nh*8 + col][c*16 + k]; this lane holds col = g, k = 2q,2q+1 | 2q+8,2q+9
  uint32_t bh[4][NHALF][2], bl[4][NHALF][2];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int nh = 0; nh < NHALF; ++nh) {
      const int h = nh * 8 + g;
      const float* w = Wg + (size_t)(h < H ? h : 0) * 64 + c * 16;
      const float z = h < H ? 1.f : 0.f;
      split_h2(z * w[2 * q], z * w[2 * q + 1], &bh[c][nh][0], &bl[c][nh][0]);
      split_h2(z * w[2 * q + 8], z * w[2 * q + 9], &bh[c][nh][1], &bl[c][nh][1]);
    }
  float bias[NHALF][2];
#pragma unroll
  for (int nh = 0; nh < NHALF; ++nh) {
    const int h = nh * 8 + 2 * q;
    bias[nh][0] = h < H ? bg[h] : 0.f;
    bias[nh][1] = h + 1 < H ? bg[h + 1] : 0.f;
  }
  const float d0 = fr.dim[2 * q], d1 = fr.dim[2 * q + 1];
  const float r0 = __frcp_rn(d0), r1 = __frcp_rn(d1);
  const int tiles_m = (M + 15) >> 4;
  const long long total = (long long)B * N * tiles_m;
  for (long long item = (long long)blockIdx.x * 4 + warp; item < total; item += (long long)gridDim.x * 4) {
    const int tm = (int)(item % tiles_m);
    const int n = (int)((item / tiles_m) % N), b = (int)(item / ((long long)tiles_m * N));
    const int m0 = tm * 16;
    const float4 bn = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + n];
    float acc[NHALF][4];
#pragma unroll
    for (int nh = 0; nh < NHALF; ++nh) { acc[nh][0] = acc[nh][1] = acc[nh][2] = acc[nh][3] = 0.f; }
    float eps[2][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int m = min(m0 + g + 8 * rr, M - 1);
      const int mi = key_index ? key_index[m] : m;
      const float4 bm_ = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + mi];
      const float4 bq = swap_roles ? bm_ : bn, bm = swap_roles ? bn : bm_;       // bq = query box, bm = key box
      if (EXACT) {
        pair_eps(bq, bm, eps[rr]);
      } else {
        const float wn = bq.z - bq.x + 1.f, hn = bq.w - bq.y + 1.f;
        const float wm = bm.z - bm.x + 1.f, hm = bm.w - bm.y + 1.f;
        const float dcx = 0.5f * (bq.x + bq.z) - 0.5f * (bm.x + bm.z), dcy = 0.5f * (bq.y + bq.w) - 0.5f * (bm.y + bm.w);
        eps[rr][0] = __logf(fmaxf(fabsf(dcx * __frcp_rn(wn)), 1e-3f));
        eps[rr][1] = __logf(fmaxf(fabsf(dcy * __frcp_rn(hn)), 1e-3f));
        eps[rr][2] = __logf(wn * __frcp_rn(wm));
        eps[rr][3] = __logf(hn * __frcp_rn(hm));
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float sn[2][2], cs[2][2];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const float a = 100.0f * eps[rr][c];
        sincos_2pi(EXACT ? __fdiv_rn(a, d0) : a * r0, &sn[rr][0], &cs[rr][0]);
        sincos_2pi(EXACT ? __fdiv_rn(a, d1) : a * r1, &sn[rr][1], &cs[rr][1]);
      }
      uint32_t ah[4], al[4];
      split_h2(sn[0][0], sn[0][1], &ah[0], &al[0]);            // a0: row g,   cols 2q,2q+1   (sin)
      split_h2(sn[1][0], sn[1][1], &ah[1], &al[1]);            // a1: row g+8
      split_h2(cs[0][0], cs[0][1], &ah[2], &al[2]);            // a2: row g,   cols 2q+8,2q+9 (cos)
      split_h2(cs[1][0], cs[1][1], &ah[3], &al[3]);            // a3: row g+8
#pragma unroll
      for (int nh = 0; nh < NHALF; ++nh) {
        // hi/lo split of both operands -> fp32-accurate FC (plain fp16 operands cost ~1e-3 on the module output when
        // the geometry weights are O(0.1): measured, tests/test_gpu_parity.py relation_cfg0_fanin)
        mma16816(acc[nh], ah[0], ah[1], ah[2], ah[3], bh[c][nh][0], bh[c][nh][1]);
        mma16816(acc[nh], al[0], al[1], al[2], al[3], bh[c][nh][0], bh[c][nh][1]);
        mma16816(acc[nh], ah[0], ah[1], ah[2], ah[3], bl[c][nh][0], bl[c][nh][1]);
      }
    }
    // C fragment: c0 (row g, head 2q), c1 (row g, head 2q+1), c2 (row g+8, head 2q), c3 (row g+8, head 2q+1)
#pragma unroll
    for (int nh = 0; nh < NHALF; ++nh)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int h = nh * 8 + 2 * q + (i & 1), m = m0 + g + 8 * (i >> 1);
        if (h < H && m < M) {
          const float gv = fmaxf(acc[nh][i] + bias[nh][i & 1], 1e-6f);
          out[(((size_t)b * H + h) * N + n) * ldg + m] = log2_out ? __log2f(gv) : gv;
        }
      }
  }
}

__global__ void pos_embed_kernel(const float* __restrict__ boxes, const int* __restrict__ key_index, int N, int M, int E,
                                 GeomFreq fr, float* __restrict__ eps_out, float* __restrict__ emb_out) {
  const int n = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float4 bn = reinterpret_cast<const float4*>(boxes)[n];
  const float4 bm = reinterpret_cast<const float4*>(boxes)[key_index ? key_index[m] : m];
  float eps[4];
  pair_eps(bn, bm, eps);
  const size_t p = (size_t)n * M + m;
  if (eps_out) reinterpret_cast<float4*>(eps_out)[p] = make_float4(eps[0], eps[1], eps[2], eps[3]);
  if (emb_out) {
    const int nf = E / 8;
    for (int c = 0; c < 4; ++c) {
      const float a = 100.0f * eps[c];
      for (int k = 0; k < nf; ++k) {
        float s, co;
        sincosf(a / fr.dim[k], &s, &co);
        emb_out[p * E + c * 2 * nf + k] = s;
        emb_out[p * E + c * 2 * nf + nf + k] = co;
      }
    }
  }
}

int make_freq(int E, float wave_length, GeomFreq* fr) {
  if (E % 8 != 0 || E / 8 > GeomFreq::kMax || E <= 0) { set_error("embedding dim E=%d unsupported (multiple of 8, <= %d)", E, 8 * GeomFreq::kMax); return RN_ERR_INVALID; }
  // dim_mat = wave_length ** ((8/E) * k), float32 powf like MXNet's broadcast_power (SYM_REL:32-34)
  for (int k = 0; k < E / 8; ++k) fr->dim[k] = powf(wave_length, (8.0f / (float)E) * (float)k);
  for (int k = E / 8; k < GeomFreq::kMax; ++k) fr->dim[k] = 1.f;
  return RN_OK;
}

static int launch_geom_weight_impl(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H,
                                   int E, float wave_length, const float* Wg, const float* bg, float* g, int ldg,
                                   int log2_out, int swap_roles = 0);

int launch_geom_weight(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H, int E,
                       float wave_length, const float* Wg, const float* bg, float* g, int ldg) {
  return launch_geom_weight_impl(st, boxes, key_index, B, N, M, H, E, wave_length, Wg, bg, g, ldg, 0);
}

// log2 of the geometry weight, the form the fused tcgen05 kernel adds to the scaled logits (relation_tc.cu)
int launch_geom_weight_log2(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H, int E,
                            float wave_length, const float* Wg, const float* bg, float* g, int ldg) {
  return launch_geom_weight_impl(st, boxes, key_index, B, N, M, H, E, wave_length, Wg, bg, g, ldg, 1);
}

// transposed log2 table over one box set (N == M, no key_index): g[h][key][query]
int launch_geom_weight_log2_T(cudaStream_t st, const float* boxes, int N, int H, int E, float wave_length, const float* Wg,
                              const float* bg, float* g, int ldg) {
  RN_CHECK_ARG(E == 64, "transposed geometry table needs E == 64");
  return launch_geom_weight_impl(st, boxes, nullptr, 1, N, N, H, E, wave_length, Wg, bg, g, ldg, 1, 1);
}

static int launch_geom_weight_impl(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H,
                                   int E, float wave_length, const float* Wg, const float* bg, float* g, int ldg,
                                   int log2_out, int swap_roles) {
  GeomFreq fr;
  int r = make_freq(E, wave_length, &fr);
  if (r) return r;
  RN_CHECK_ARG(H >= 1 && H <= 16, "geometry heads H=%d unsupported (1..16)", H);
  if (E == 64 && is_sm100())      // sm_100a: pair FC on tcgen05 (geom_tc.cu)
    return launch_geom_weight_tc(st, boxes, key_index, B, N, M, H, fr, Wg, bg, g, ldg, log2_out, swap_roles, !log2_out);
  if (E == 64) {
    // mma.sync FC path: 16 pairs per warp-tile
    const long long items = (long long)B * N * cdiv(M, 16);
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int grid = (int)std::min<long long>((items + 3) / 4, (long long)sms * 8);
    const bool exact = !log2_out;               // fp32 parity mode asks for g, the tcgen05 path for log2 g
    if (H <= 8) {
      if (exact) geom_weight_mma_kernel<1, true><<<grid, 128, 0, st>>>(boxes, key_index, B, N, M, H, fr, Wg, bg, g, ldg, log2_out, swap_roles);
      else geom_weight_mma_kernel<1, false><<<grid, 128, 0, st>>>(boxes, key_index, B, N, M, H, fr, Wg, bg, g, ldg, log2_out, swap_roles);
    } else {
      if (exact) geom_weight_mma_kernel<2, true><<<grid, 128, 0, st>>>(boxes, key_index, B, N, M, H, fr, Wg, bg, g, ldg, log2_out, swap_roles);
      else geom_weight_mma_kernel<2, false><<<grid, 128, 0, st>>>(boxes, key_index, B, N, M, H, fr, Wg, bg, g, ldg, log2_out, swap_roles);
    }
    RN_LAUNCH_CHECK();
    return RN_OK;
  }
  dim3 grid(cdiv(M, 128), N, B);
  if (H <= 4) {
    size_t smem = (size_t)(E * 4 + 4) * sizeof(float);
    geom_weight_kernel<4><<<grid, 128, smem, st>>>(boxes, key_index, N, M, H, E, fr, Wg, bg, g, ldg, log2_out);
  } else {
    size_t smem = (size_t)(E * 16 + 16) * sizeof(float);
    geom_weight_kernel<16><<<grid, 128, smem, st>>>(boxes, key_index, N, M, H, E, fr, Wg, bg, g, ldg, log2_out);
  }
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn

extern "C" int rn_pos_embed_fwd(const float* boxes, const int32_t* key_index, int32_t N, int32_t M, int32_t E,
                                float wave_length, float* eps_out, float* emb_out, rn_stream_t stream) {
  RN_CHECK_ARG(boxes && N > 0 && M > 0, "rn_pos_embed_fwd: bad arguments");
  rn::GeomFreq fr;
  int r = rn::make_freq(E, wave_length, &fr);
  if (r) return r;
  dim3 grid(rn::cdiv(M, 128), N);
  rn::pos_embed_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(boxes, key_index, N, M, E, fr, eps_out, emb_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_geometry_weight_fwd(const float* boxes, const int32_t* key_index, int32_t batch, int32_t N, int32_t M,
                                      int32_t H, int32_t E, float wave_length, const float* Wg, const float* bg,
                                      float* g_out, rn_stream_t stream) {
  RN_CHECK_ARG(boxes && Wg && bg && g_out && batch > 0 && N > 0 && M > 0, "rn_geometry_weight_fwd: bad arguments");
  return rn::launch_geom_weight((cudaStream_t)stream, boxes, key_index, batch, N, M, H, E, wave_length, Wg, bg, g_out, M);
}
