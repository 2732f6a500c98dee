// geom.cuh -- shared device helpers for the pairwise box geometry
#pragma once
#include <cuda_runtime.h>

namespace rn {

struct GeomFreq {
  static constexpr int kMax = 16;   // E/8 frequencies (8 for E = 64)
  float dim[kMax];                  // wave_length ** ((8/E) k)
};

// eps[n,m,:] of SYM_REL:56-75, float32 op for op: the division is by the QUERY box's width/height
__device__ __forceinline__ void pair_eps(const float4 bn, const float4 bm, float* eps) {
  const float wn = bn.z - bn.x + 1.f, hn = bn.w - bn.y + 1.f;
  const float wm = bm.z - bm.x + 1.f, hm = bm.w - bm.y + 1.f;
  const float cxn = 0.5f * (bn.x + bn.z), cyn = 0.5f * (bn.y + bn.w);
  const float cxm = 0.5f * (bm.x + bm.z), cym = 0.5f * (bm.y + bm.w);
  eps[0] = logf(fmaxf(fabsf((cxn - cxm) / wn), 1e-3f));
  eps[1] = logf(fmaxf(fabsf((cyn - cym) / hn), 1e-3f));
  eps[2] = logf(wn / wm);
  eps[3] = logf(hn / hm);
}

int make_freq(int E, float wave_length, GeomFreq* fr);
int launch_geom_weight(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H, int E,
                       float wave_length, const float* Wg, const float* bg, float* g, int ldg);

}  // namespace rn
