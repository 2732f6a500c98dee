// geom.cu -- pairwise box geometry: eps (SYM_REL:47-83), sinusoid embedding (SYM_REL:30-44) and the geometry weight
// g = max(relu(Wg.phi + bg), 1e-6) (SYM_REL:107-116,139) WITHOUT ever materialising the [N,M,64] embedding.
//
// HBM layout:  boxes [B,N,4] fp32 in;  g [B,H,N,ldg] fp32 out (ldg >= M, rows padded so 16-byte row loads stay aligned).
// Roofline: transcendental (MUFU/FMA) bound -- per pair 4 logf + (E/2) sincosf + E*H FMA; algorithmic bytes are only
// the g write (4*B*H*N*M) + boxes.
#include "common.cuh"
#include "geom.cuh"

namespace rn {

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
                                   int log2_out);

int launch_geom_weight(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H, int E,
                       float wave_length, const float* Wg, const float* bg, float* g, int ldg) {
  return launch_geom_weight_impl(st, boxes, key_index, B, N, M, H, E, wave_length, Wg, bg, g, ldg, 0);
}

// log2 of the geometry weight, the form the fused tcgen05 kernel adds to the scaled logits (relation_tc.cu)
int launch_geom_weight_log2(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H, int E,
                            float wave_length, const float* Wg, const float* bg, float* g, int ldg) {
  return launch_geom_weight_impl(st, boxes, key_index, B, N, M, H, E, wave_length, Wg, bg, g, ldg, 1);
}

static int launch_geom_weight_impl(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H,
                                   int E, float wave_length, const float* Wg, const float* bg, float* g, int ldg,
                                   int log2_out) {
  GeomFreq fr;
  int r = make_freq(E, wave_length, &fr);
  if (r) return r;
  RN_CHECK_ARG(H >= 1 && H <= 16, "geometry heads H=%d unsupported (1..16)", H);
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
