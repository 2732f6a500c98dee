// proposal.cu -- device-resident `proposal` op (operator_py/proposal.py:51-168), NMS (lib/nms/nms_kernel.cu),
// fp64 IoU (lib/bbox/bbox.pyx) and `proposal_target` (operator_py/proposal_target.py:44-93, core/rcnn.py:288-325).
// Compiled with -fmad=false: decode/encode follow numpy's separate mul/add roundings so indices are bit-exact.
//
// Pipeline (no host round trip, no hidden allocation; the reference does 2 full D2H copies, a numpy argsort, a
// cudaMalloc'd NMS with a 4.5 MB mask D2H and a serial host sweep):
//   decode+clip+filter (fp64, 1 thread/anchor)  -> keys (order-preserving uint32 of the fg score, 0 = filtered out)
//   single-CTA bitonic sort of (key, index) in 192 KB of shared memory (n <= 32768), descending, ties -> larger index
//   gather top pre_nms boxes as float32 -> 64x64 IoU bitmask tiles (upper triangle) -> one-CTA block sweep with early
//   exit after post_nms kept -> emit rois (+ deterministic padding keep[i % kept]).
// HBM traffic is ~1 MB of maps + <= 4.5 MB of mask: latency bound, not bandwidth bound (DESIGN.md "proposal").
#include "common.cuh"
#include "umma.cuh"
#include <algorithm>

namespace rn {

constexpr int kMaxAnchors = 64;
struct AnchorSet { double a[kMaxAnchors][4]; int A; };

// lib/rpn/generate_anchor.py:22-86 in double (np.round = half-to-even; the products here never land on .5 for the
// configured ratios/scales, and nearbyint() in the default rounding mode implements half-to-even anyway)
static void generate_anchors_host(int base, const float* ratios, int nr, const float* scales, int ns, AnchorSet* out) {
  const double w = base, h = base, xc = 0.5 * (base - 1), yc = 0.5 * (base - 1);
  int k = 0;
  for (int i = 0; i < nr; ++i) {
    const double ws = nearbyint(sqrt(w * h / (double)ratios[i]));
    const double hs = nearbyint(ws * (double)ratios[i]);
    for (int j = 0; j < ns; ++j, ++k) {
      const double W = ws * (double)scales[j], H = hs * (double)scales[j];
      out->a[k][0] = xc - 0.5 * (W - 1); out->a[k][1] = yc - 0.5 * (H - 1);
      out->a[k][2] = xc + 0.5 * (W - 1); out->a[k][3] = yc + 0.5 * (H - 1);
    }
  }
  out->A = k;
}

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// proposal.py:75-137 + bbox_transform.py:103-140 (nonlinear_pred, float64) + :45-60 (clip_boxes)
__global__ void __launch_bounds__(256) proposal_decode_kernel(AnchorSet anc, const float* __restrict__ cls_prob,
                                                              const float* __restrict__ bbox_pred,
                                                              const float* __restrict__ im_info, int Hf, int Wf,
                                                              int feat_stride, float min_size, double* __restrict__ props,
                                                              uint32_t* __restrict__ keys, int* __restrict__ n_total_out) {
  const int A = anc.A;
  const float im_h = im_info[0], im_w = im_info[1], im_s = im_info[2];
  const int height = min((int)(im_h / (float)feat_stride), Hf), width = min((int)(im_w / (float)feat_stride), Wf);
  const int n = height * width * A;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_total_out = n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = i % A, x = (i / A) % width, y = i / (A * width);
  const size_t plane = (size_t)Hf * Wf, pix = (size_t)y * Wf + x;
  const float score = cls_prob[(size_t)(A + a) * plane + pix];
  const float dx = bbox_pred[(size_t)(4 * a + 0) * plane + pix], dy = bbox_pred[(size_t)(4 * a + 1) * plane + pix];
  const float dw = bbox_pred[(size_t)(4 * a + 2) * plane + pix], dh = bbox_pred[(size_t)(4 * a + 3) * plane + pix];
  const double sx = (double)(x * feat_stride), sy = (double)(y * feat_stride);
  const double x1 = anc.a[a][0] + sx, y1 = anc.a[a][1] + sy, x2 = anc.a[a][2] + sx, y2 = anc.a[a][3] + sy;
  const double w = x2 - x1 + 1.0, h = y2 - y1 + 1.0;
  const double cx = x1 + 0.5 * (w - 1.0), cy = y1 + 0.5 * (h - 1.0);
  const double pcx = (double)dx * w + cx, pcy = (double)dy * h + cy;
  // np.exp on the float32 delta: defined as the correctly rounded float32 exp (oracle/proposal_np.py:decode_boxes)
  const double pw = (double)(float)exp((double)dw) * w, ph = (double)(float)exp((double)dh) * h;
  double bx1 = pcx - 0.5 * (pw - 1.0), by1 = pcy - 0.5 * (ph - 1.0);
  double bx2 = pcx + 0.5 * (pw - 1.0), by2 = pcy + 0.5 * (ph - 1.0);
  const double lim_w = (double)(im_w - 1.0f), lim_h = (double)(im_h - 1.0f);
  bx1 = fmax(fmin(bx1, lim_w), 0.0); by1 = fmax(fmin(by1, lim_h), 0.0);
  bx2 = fmax(fmin(bx2, lim_w), 0.0); by2 = fmax(fmin(by2, lim_h), 0.0);
  const double ms = (double)(min_size * im_s);
  const bool ok = (bx2 - bx1 + 1.0 >= ms) && (by2 - by1 + 1.0 >= ms);
  double* p = props + (size_t)i * 4;
  p[0] = bx1; p[1] = by1; p[2] = bx2; p[3] = by2;
  keys[i] = ok ? float_to_ordered(score) : 0u;
}

// ---- sort: descending by (key, index), top pre_nms_top_n only.  Two launches, all SMs busy:
//  (1) proposal_sort_chunks_kernel: every CTA bitonic-sorts one 1024-element chunk in shared memory;
//  (2) proposal_rank_kernel: every CTA stages ALL sorted chunks (6 B/element, <= 192 KB) in shared memory; one warp per
//      element, one lane per chunk: three binary searches give the element's exact global rank and the bounds of its
//      run of equal scores, and the element is scattered straight to its final slot.  (A single-CTA bitonic sort of
//      32768 keys is shared-memory-bandwidth bound on ONE SM: 356 us measured, profiles/r01_launches_hot_v1.csv.)
constexpr int kChunk = 1024;
constexpr int kKS = 1025, kIS = 1026;      // padded shared-memory strides of a chunk (keys / indices)

__device__ __forceinline__ bool comp_before(uint32_t ka, uint32_t ia, uint32_t kb, uint32_t ib) {
  return (ka > kb) || (ka == kb && ia > ib);       // score descending, ties -> larger index first
}

__global__ void __launch_bounds__(512) proposal_sort_chunks_kernel(const uint32_t* __restrict__ keys,
                                                                   const int* __restrict__ n_ptr,
                                                                   uint32_t* __restrict__ skeys, uint16_t* __restrict__ sidx) {
  __shared__ uint32_t k[kChunk];
  __shared__ uint16_t ix[kChunk];
  const int n = *n_ptr, base = blockIdx.x * kChunk;
  for (int i = threadIdx.x; i < kChunk; i += 512) {
    const int g = base + i;
    k[i] = g < n ? keys[g] : 0u;
    ix[i] = (uint16_t)(g < n ? g : 0xFFFF);
  }
  __syncthreads();
  for (int size = 2; size <= kChunk; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int t = threadIdx.x;
      const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
      const bool desc = ((lo & size) == 0);
      const uint32_t ka = k[lo], kb = k[hi];
      const uint16_t ia = ix[lo], ib = ix[hi];
      if (comp_before(ka, ia, kb, ib) != desc) { k[lo] = kb; k[hi] = ka; ix[lo] = ib; ix[hi] = ia; }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < kChunk; i += 512) { skeys[base + i] = k[i]; sidx[base + i] = ix[i]; }
}

__global__ void __launch_bounds__(1024) proposal_rank_kernel(const uint32_t* __restrict__ skeys,
                                                             const uint16_t* __restrict__ sidx, int num_chunks,
                                                             int pre_nms_top_n, int* __restrict__ order,
                                                             int* __restrict__ n_pre_out) {
  extern __shared__ uint8_t sm[];
  const int P = num_chunks * kChunk;
  // chunk c lives at c*kKS keys / c*kIS indices: the odd strides put equal offsets of different chunks in different
  // banks (lane c probes chunk c at the same `mid` in the lock-step binary searches below)
  uint32_t* k = reinterpret_cast<uint32_t*>(sm);
  uint16_t* ix = reinterpret_cast<uint16_t*>(sm + (size_t)num_chunks * kKS * 4);
  __shared__ int s_n_pre;
  for (int i = threadIdx.x; i < P / 8; i += blockDim.x) {      // 8 elements per thread-iteration: 2 x uint4 keys, 1 x uint4 idx
    const int e0 = i * 8, c = e0 >> 10, o = e0 & (kChunk - 1);
    const uint4 k0 = reinterpret_cast<const uint4*>(skeys)[2 * i], k1 = reinterpret_cast<const uint4*>(skeys)[2 * i + 1];
    const uint4 i0 = reinterpret_cast<const uint4*>(sidx)[i];
    uint32_t* kd = k + c * kKS + o;
    kd[0] = k0.x; kd[1] = k0.y; kd[2] = k0.z; kd[3] = k0.w; kd[4] = k1.x; kd[5] = k1.y; kd[6] = k1.z; kd[7] = k1.w;
    uint32_t* id = reinterpret_cast<uint32_t*>(ix + c * kIS + o);    // kIS and o are even -> 4-byte aligned
    id[0] = i0.x; id[1] = i0.y; id[2] = i0.z; id[3] = i0.w;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_chunk = lane < num_chunks;
  const uint32_t* kc = k + lane * kKS;
  const uint16_t* ic = ix + lane * kIS;
  if (warp == 0) {
    // number of valid (key != 0) elements: zeros sit at the end of every sorted chunk
    int a = 0, b = kChunk;
    if (has_chunk) { while (a < b) { const int mid = (a + b) >> 1; if (kc[mid] != 0u) a = mid + 1; else b = mid; } } else a = 0;
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      int np = a;
      if (pre_nms_top_n > 0 && np > pre_nms_top_n) np = pre_nms_top_n;
      s_n_pre = np;
      if (blockIdx.x == 0) *n_pre_out = np;
    }
  }
  __syncthreads();
  const int n_pre = s_n_pre;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int e = blockIdx.x * (blockDim.x >> 5) + warp; e < P; e += warps_total) {
    const int li = e & (kChunk - 1), ce = e >> 10;
    const uint32_t ke = k[ce * kKS + li];
    const uint32_t ie = ix[ce * kIS + li];
    if (ke == 0u || li >= n_pre) continue;          // warp-uniform
    int gt_comp = 0, gt_key = 0, ge_key = 0;
    if (has_chunk) {
      int a = 0, b = kChunk;                         // first position not strictly before e in the composite order
      if (lane == ce) a = li;
      else while (a < b) { const int mid = (a + b) >> 1; if (comp_before(kc[mid], ic[mid], ke, ie)) a = mid + 1; else b = mid; }
      gt_comp = a;
      a = 0; b = kChunk;
      while (a < b) { const int mid = (a + b) >> 1; if (kc[mid] > ke) a = mid + 1; else b = mid; }
      gt_key = a;
      b = kChunk;                                    // a == gt_key is a valid lower bound
      while (a < b) { const int mid = (a + b) >> 1; if (kc[mid] >= ke) a = mid + 1; else b = mid; }
      ge_key = a;
    }
    for (int o = 16; o; o >>= 1) {
      gt_comp += __shfl_xor_sync(0xffffffffu, gt_comp, o);
      gt_key += __shfl_xor_sync(0xffffffffu, gt_key, o);
      ge_key += __shfl_xor_sync(0xffffffffu, ge_key, o);
    }
    if (lane == 0 && gt_comp < n_pre) {
      // gpu_nms re-sorts the selected boxes by score (lib/nms/gpu_nms.pyx:26, `argsort()[::-1]`): under the stable-sort
      // tie rule that reverses every run of equal scores once more -> mirror the rank inside its run [lo, hi]
      const int lo = gt_key, hi = min(ge_key, n_pre) - 1;
      order[lo + hi - gt_comp] = (int)ie;
    }
  }
}

// det = hstack(proposals, scores).astype(float32)  (proposal.py:149)
__global__ void proposal_gather_kernel(const double* __restrict__ props, const float* __restrict__ cls_prob,
                                       const int* __restrict__ order, const int* __restrict__ n_pre_ptr, int A, int Hf,
                                       int Wf, const float* __restrict__ im_info, int feat_stride,
                                       float* __restrict__ det) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_pre_ptr) return;
  const int src = order[i];
  const int width = min((int)(im_info[1] / (float)feat_stride), Wf);
  const int a = src % A, x = (src / A) % width, y = src / (A * width);
  const double* p = props + (size_t)src * 4;
  float* d = det + (size_t)i * 5;
  d[0] = (float)p[0]; d[1] = (float)p[1]; d[2] = (float)p[2]; d[3] = (float)p[3];
  d[4] = cls_prob[(size_t)(A + a) * Hf * Wf + (size_t)y * Wf + x];
}

// lib/nms/nms_kernel.cu:24-32, float32 op for op
__device__ __forceinline__ float dev_iou(const float* a, const float* b) {
  const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  const float width = fmaxf(right - left + 1, 0.f), height = fmaxf(bottom - top + 1, 0.f);
  const float interS = width * height;
  const float Sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
  const float Sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  return interS / (Sa + Sb - interS);
}

// 64x64 tile of the suppression matrix (nms_kernel.cu:34-78); only tiles on/above the diagonal are needed by the
// sweep, so the grid enumerates the upper triangle linearly (row-major over row blocks).
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ boxes, int box_dim,
                                                      const int* __restrict__ n_ptr, int cb, int col_blocks_ld,
                                                      float thresh, unsigned long long* __restrict__ mask) {
  const int n = *n_ptr;
  // t = row*cb - row*(row-1)/2 + (col - row)  ->  row = floor(((2cb+1) - sqrt((2cb+1)^2 - 8t)) / 2)
  const int t = blockIdx.x;
  int row_start = (int)(((2.0f * cb + 1.0f) - sqrtf((2.0f * cb + 1.0f) * (2.0f * cb + 1.0f) - 8.0f * (float)t)) * 0.5f);
  while (row_start > 0 && row_start * cb - row_start * (row_start - 1) / 2 > t) --row_start;
  while ((row_start + 1) * cb - (row_start + 1) * row_start / 2 <= t) ++row_start;
  const int col_start = row_start + (t - (row_start * cb - row_start * (row_start - 1) / 2));
  if (row_start * 64 >= n || col_start * 64 >= n) return;
  const int row_size = min(n - row_start * 64, 64), col_size = min(n - col_start * 64, 64);
  __shared__ float bb[64 * 4];
  if (threadIdx.x < col_size) {
    const float* s = boxes + (size_t)(64 * col_start + threadIdx.x) * box_dim;
    bb[threadIdx.x * 4 + 0] = s[0]; bb[threadIdx.x * 4 + 1] = s[1];
    bb[threadIdx.x * 4 + 2] = s[2]; bb[threadIdx.x * 4 + 3] = s[3];
  }
  __syncthreads();
  if (threadIdx.x < row_size) {
    const int cur = 64 * row_start + threadIdx.x;
    const float* cbx = boxes + (size_t)cur * box_dim;
    const float c4[4] = {cbx[0], cbx[1], cbx[2], cbx[3]};
    unsigned long long bits = 0;
    const int start = (row_start == col_start) ? threadIdx.x + 1 : 0;
    const bool shortcut = thresh >= 0.f;      // disjoint boxes have IoU == +0 exactly: skip the areas and the division
    for (int i = start; i < col_size; ++i) {
      const float* o = bb + i * 4;
      if (shortcut && (fminf(c4[2], o[2]) - fmaxf(c4[0], o[0]) + 1 <= 0.f || fminf(c4[3], o[3]) - fmaxf(c4[1], o[1]) + 1 <= 0.f))
        continue;
      if (dev_iou(c4, o) > thresh) bits |= 1ULL << i;
    }
    mask[((size_t)row_start * col_blocks_ld + col_start) * 64 + threadIdx.x] = bits;   // blocked layout [rb][w][r]
  }
}

// Greedy sweep (nms_kernel.cu:124-139) on the device, 64 boxes per step, early exit once max_keep are kept.
// The mask is stored blocked, [row block][word][row in block], so the 64 rows x (col_blocks - blk) words a step needs
// are ONE contiguous span: a single cp.async.bulk stages it in shared memory one step ahead (double buffer, mbarrier
// completion).  (64 per-row bulk copies cost ~50 cycles of TMA issue each and made the step slower: 164 us measured.)
// A step = wait barrier -> one thread resolves the block with a find-first-set loop over the still-alive bits -> all
// threads OR the kept rows into the running suppression words.
__global__ void __launch_bounds__(256) nms_sweep_kernel(const unsigned long long* __restrict__ mask,
                                                        const int* __restrict__ n_ptr, int col_blocks_ld, int max_keep,
                                                        int* __restrict__ keep_out, int* __restrict__ num_out) {
  extern __shared__ __align__(16) unsigned long long sm64[];       // stage[2][col_blocks_ld][64] | remv[col_blocks_ld]
  unsigned long long* stage = sm64;
  unsigned long long* remv = sm64 + (size_t)2 * 64 * col_blocks_ld;
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int kept_local[64];
  __shared__ int s_nk_local, s_total;
  const int n = *n_ptr;
  const int col_blocks = (n + 63) / 64;
  const int ld = col_blocks_ld;
  for (int i = threadIdx.x; i < ld; i += blockDim.x) remv[i] = 0ULL;
  if (threadIdx.x == 0) {
    s_total = 0;
    umma::mbar_init(&bar[0], 1); umma::mbar_init(&bar[1], 1);
    umma::fence_barrier_init();
  }
  __syncthreads();
  auto stage_block = [&](int blk) {                     // thread 0 only
    const uint32_t bytes = (uint32_t)(col_blocks - blk) * 64u * 8u;
    umma::fence_proxy_async_smem();
    umma::mbar_arrive_expect_tx(&bar[blk & 1], bytes);
    umma::bulk_copy_g2s(stage + (size_t)(blk & 1) * 64 * ld, mask + ((size_t)blk * ld + blk) * 64, bytes, &bar[blk & 1]);
  };
  if (threadIdx.x == 0 && col_blocks > 0) stage_block(0);
  for (int blk = 0; blk < col_blocks; ++blk) {
    const int base = blk * 64, cnt = min(64, n - base);
    if (threadIdx.x == 0 && blk + 1 < col_blocks) stage_block(blk + 1);
    umma::mbar_wait(&bar[blk & 1], (blk >> 1) & 1);
    const unsigned long long* rows = stage + (size_t)(blk & 1) * 64 * ld;      // rows[(w - blk)*64 + r]
    if (threadIdx.x == 0) {
      unsigned long long cur = remv[blk];
      const unsigned long long valid = cnt == 64 ? ~0ULL : ((1ULL << cnt) - 1ULL);
      unsigned long long avail = ~cur & valid;
      int nk = 0, total = s_total;
      while (avail && total < max_keep) {
        const int i = __ffsll((long long)avail) - 1;
        kept_local[nk++] = i;
        keep_out[total++] = base + i;
        cur |= rows[i];                                        // diagonal word: bits j > i of this block
        avail = ~cur & valid & ~((2ULL << i) - 1ULL);          // only bits above i remain candidates
      }
      s_nk_local = nk; s_total = total;
    }
    __syncthreads();
    if (s_total >= max_keep) break;
    const int nk = s_nk_local;
    if (nk > 0)
      for (int w = blk + 1 + threadIdx.x; w < col_blocks; w += blockDim.x) {
        unsigned long long acc = remv[w];
        for (int j = 0; j < nk; ++j) acc |= rows[(size_t)(w - blk) * 64 + kept_local[j]];
        remv[w] = acc;
      }
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_out = s_total;
}

// Greedy NMS without the n x n/64 mask (used when max_keep is small, e.g. post_nms_top_n = 300): one CTA walks the sorted
// boxes 64 at a time; the <= max_keep kept boxes live in shared memory.  Per block: (1) 64 candidates x kept boxes IoU
// in parallel, (2) the 64x64 upper-triangle IoU bits, (3) a 64-step serial resolve on one 64-bit word; stops as soon as
// max_keep boxes are kept.  Same float32 IoU and '>' as lib/nms/nms_kernel.cu, so the kept set is bit-identical to the
// mask + sweep form (which took 57 + 264 us at n = 6000: profiles/r01_launches_hot_v1.csv).
__global__ void __launch_bounds__(256) nms_greedy_kernel(const float* __restrict__ boxes, int box_dim,
                                                         const int* __restrict__ n_ptr, float thresh, int max_keep,
                                                         int* __restrict__ keep_out, int* __restrict__ num_out) {
  extern __shared__ float4 kept[];                 // [max_keep]
  __shared__ float4 cand[64];
  __shared__ unsigned int supp[2];                 // candidate suppressed by an earlier kept box (bit c)
  __shared__ unsigned short m16[64][4];            // intra-block suppression bits, 16 columns per slice
  __shared__ int s_nkept;
  const int n = *n_ptr;
  const int tid = threadIdx.x;
  if (tid == 0) s_nkept = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 64) {
    const int cnt = min(64, n - base);
    if (tid < 64) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < cnt) { const float* p = boxes + (size_t)(base + tid) * box_dim; b = make_float4(p[0], p[1], p[2], p[3]); }
      cand[tid] = b;
    }
    if (tid < 2) supp[tid] = 0u;
    __syncthreads();
    const int nkept = s_nkept;
    const int c = tid & 63, part = tid >> 6;
    {
      const float4 cb = cand[c];
      const float cf[4] = {cb.x, cb.y, cb.z, cb.w};
      bool hit = false;
      for (int kk = part; kk < nkept && !hit; kk += 4) {
        const float4 kb = kept[kk];
        const float kf[4] = {kb.x, kb.y, kb.z, kb.w};
        hit = dev_iou(kf, cf) > thresh;
      }
      if (hit && c < cnt) atomicOr(&supp[c >> 5], 1u << (c & 31));
      unsigned int bits = 0;
      for (int j = part * 16; j < part * 16 + 16; ++j) {
        if (j > c && j < cnt) {
          const float4 jb = cand[j];
          const float jf[4] = {jb.x, jb.y, jb.z, jb.w};
          if (dev_iou(cf, jf) > thresh) bits |= 1u << (j - part * 16);
        }
      }
      m16[c][part] = (unsigned short)bits;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long cur = (unsigned long long)supp[0] | ((unsigned long long)supp[1] << 32);
      int nk = nkept;
      for (int i = 0; i < cnt && nk < max_keep; ++i) {
        if (!((cur >> i) & 1ULL)) {
          kept[nk] = cand[i];
          keep_out[nk++] = base + i;
          cur |= (unsigned long long)m16[i][0] | ((unsigned long long)m16[i][1] << 16) |
                 ((unsigned long long)m16[i][2] << 32) | ((unsigned long long)m16[i][3] << 48);
        }
      }
      s_nkept = nk;
    }
    __syncthreads();
    if (s_nkept >= max_keep) break;
  }
  if (tid == 0) *num_out = s_nkept;
}

// proposal.py:151-168: take post_nms, pad, emit [0, x1, y1, x2, y2] float32 (+ scores)
__global__ void proposal_emit_kernel(const float* __restrict__ det, const int* __restrict__ keep,
                                     const int* __restrict__ num_kept_ptr, int post, float* __restrict__ rois,
                                     float* __restrict__ scores, int* __restrict__ num_kept_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nk = *num_kept_ptr;
  if (i == 0 && num_kept_out) *num_kept_out = nk;
  if (i >= post) return;
  float* r = rois + (size_t)i * 5;
  if (nk <= 0) { r[0] = r[1] = r[2] = r[3] = r[4] = 0.f; if (scores) scores[i] = 0.f; return; }
  const int src = keep[i < nk ? i : (i - nk) % nk];
  const float* d = det + (size_t)src * 5;
  r[0] = 0.f; r[1] = d[0]; r[2] = d[1]; r[3] = d[2]; r[4] = d[3];
  if (scores) scores[i] = d[4];
}

// lib/bbox/bbox.pyx:15-55
__device__ __forceinline__ double iou_f64(const double* b, const double* q) {
  const double iw = fmin(b[2], q[2]) - fmax(b[0], q[0]) + 1;
  if (iw <= 0) return 0.0;
  const double ih = fmin(b[3], q[3]) - fmax(b[1], q[1]) + 1;
  if (ih <= 0) return 0.0;
  const double qa = (q[2] - q[0] + 1) * (q[3] - q[1] + 1);
  const double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + qa - iw * ih;
  return iw * ih / ua;
}

__global__ void bbox_overlaps_kernel(const double* __restrict__ boxes, const double* __restrict__ query, int N, int K,
                                     double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * K) return;
  const int n = i / K, k = i % K;
  out[i] = iou_f64(boxes + (size_t)n * 4, query + (size_t)k * 4);
}

// proposal_target.py:44-93 (BATCH_ROIS = -1) -> rcnn.py:288-325 -> bbox_transform.py:74-100 -> bbox_regression.py:120-140
__global__ void proposal_target_kernel(rn_proposal_target_desc d, const float* __restrict__ rois,
                                       const float* __restrict__ gt, float* __restrict__ rois_out,
                                       float* __restrict__ label, float* __restrict__ bt, float* __restrict__ bw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int T = d.N + d.G;
  if (i >= T) return;
  float r[5];
  if (i < d.N) { for (int j = 0; j < 5; ++j) r[j] = rois[(size_t)i * 5 + j]; }
  else { const float* g = gt + (size_t)(i - d.N) * 5; r[0] = 0.f; r[1] = g[0]; r[2] = g[1]; r[3] = g[2]; r[4] = g[3]; }
  for (int j = 0; j < 5; ++j) rois_out[(size_t)i * 5 + j] = r[j];
  const int R = d.class_agnostic ? 2 : d.num_reg_classes;
  for (int j = 0; j < 4 * R; ++j) { bt[(size_t)i * 4 * R + j] = 0.f; bw[(size_t)i * 4 * R + j] = 0.f; }
  if (d.G == 0) { label[i] = 0.f; return; }
  const double b[4] = {(double)r[1], (double)r[2], (double)r[3], (double)r[4]};
  double best = -1.0; int arg = 0;
  for (int k = 0; k < d.G; ++k) {
    const float* g = gt + (size_t)k * 5;
    const double q[4] = {(double)g[0], (double)g[1], (double)g[2], (double)g[3]};
    const double ov = iou_f64(b, q);
    if (ov > best) { best = ov; arg = k; }            // first maximum, like ndarray.argmax
  }
  const float* g = gt + (size_t)arg * 5;
  float lab = g[4];
  if (best < (double)d.bg_thresh_hi) lab = 0.f;
  label[i] = lab;
  if (lab > 0.f) {
    // nonlinear_transform in float32 (inputs are float32 arrays)
    const float ew = r[3] - r[1] + 1.0f, eh = r[4] - r[2] + 1.0f;
    const float ecx = r[1] + 0.5f * (ew - 1.0f), ecy = r[2] + 0.5f * (eh - 1.0f);
    const float gw = g[2] - g[0] + 1.0f, gh = g[3] - g[1] + 1.0f;
    const float gcx = g[0] + 0.5f * (gw - 1.0f), gcy = g[1] + 0.5f * (gh - 1.0f);
    float t[4] = {(gcx - ecx) / (ew + 1e-14f), (gcy - ecy) / (eh + 1e-14f), logf(gw / ew), logf(gh / eh)};
    const int s = d.class_agnostic ? 4 : 4 * (int)lab;
    for (int j = 0; j < 4; ++j) {
      const float v = d.normalize ? (float)(((double)t[j] - d.means[j]) / d.stds[j]) : t[j];
      bt[(size_t)i * 4 * R + s + j] = v;
      bw[(size_t)i * 4 * R + s + j] = d.bbox_weights[j];
    }
  }
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

struct ProposalWs {
  double* props; uint32_t* keys; uint32_t* skeys; uint16_t* sidx; int* order; float* det; unsigned long long* mask; int* keep; int* counters;
  int n_max, pre, col_blocks;
};

static size_t carve(const rn_proposal_desc* d, void* base, size_t bytes, ProposalWs* w) {
  const int A = d->num_scales * d->num_ratios;
  const int n_max = d->Hf * d->Wf * A;
  const int pre = d->pre_nms_top_n > 0 ? (d->pre_nms_top_n < n_max ? d->pre_nms_top_n : n_max) : n_max;
  const int cb = ((pre + 63) / 64 + 1) & ~1;      // even: 16-byte aligned mask rows
  const int P = cdiv(n_max, kChunk) * kChunk;
  size_t need = ws_slice((size_t)n_max * 4, 8) + ws_slice(n_max, 4) + ws_slice(P, 4) + ws_slice(P, 2) + ws_slice(pre, 4) +
                ws_slice((size_t)pre * 5, 4) +
                ws_slice((size_t)(pre + 63) / 64 * 64 * cb, 8) + ws_slice(d->post_nms_top_n > 0 ? d->post_nms_top_n : pre, 4) + ws_slice(8, 4);
  if (!w) return need;
  Workspace ws(base, bytes);
  w->n_max = n_max; w->pre = pre; w->col_blocks = cb;
  w->props = ws.take<double>((size_t)n_max * 4);
  w->keys = ws.take<uint32_t>(n_max);
  w->skeys = ws.take<uint32_t>(P);
  w->sidx = ws.take<uint16_t>(P);
  w->order = ws.take<int>(pre);
  w->det = ws.take<float>((size_t)pre * 5);
  w->mask = ws.take<unsigned long long>((size_t)(pre + 63) / 64 * 64 * cb);
  w->keep = ws.take<int>(d->post_nms_top_n > 0 ? d->post_nms_top_n : pre);
  w->counters = ws.take<int>(8);
  return w->counters ? need : 0;
}

static int launch_nms(cudaStream_t st, const float* boxes, int box_dim, const int* n_ptr, int n_max, float thresh,
                      int max_keep, unsigned long long* mask, int* keep, int* num_out) {
  if (n_max <= 256) {
    const size_t smem = (size_t)max_keep * sizeof(float4);
    static thread_local size_t configured = 40 * 1024;
    if (smem > configured) {
      RN_CUDA(cudaFuncSetAttribute(nms_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    nms_greedy_kernel<<<1, 256, smem, st>>>(boxes, box_dim, n_ptr, thresh, max_keep, keep, num_out);
    RN_LAUNCH_CHECK();
    return RN_OK;
  }
  // (A keep-list form -- one CTA, all boxes in shared memory, candidates x kept boxes, no mask -- was built and measured in
  // round 2: bit-identical lists, but 250 us against 41 + 65 us for mask + sweep at 6000 -> 300: the ~0.9 M IoUs it needs are
  // one SM's work, while the 18 M of the triangular mask spread over 148.)
  const int cbv = (n_max + 63) / 64;              // valid 64-box blocks
  const int cb = (cbv + 1) & ~1;                  // even leading dimension of the mask rows
  nms_mask_kernel<<<cbv * (cbv + 1) / 2, 64, 0, st>>>(boxes, box_dim, n_ptr, cbv, cb, thresh, mask);
  RN_LAUNCH_CHECK();
  {
    const size_t smem = (size_t)cb * 8 * (1 + 2 * 64);
    RN_CHECK_ARG(smem <= 220 * 1024, "rn_nms: %d boxes exceed the staged sweep capacity (~13400)", n_max);
    static thread_local size_t configured = 40 * 1024;
    if (smem > configured) {
      RN_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    nms_sweep_kernel<<<1, 256, smem, st>>>(mask, n_ptr, cb, max_keep, keep, num_out);
    RN_LAUNCH_CHECK();
  }
  return RN_OK;
}

}  // namespace rn

extern "C" size_t rn_proposal_workspace_bytes(const rn_proposal_desc* d) {
  if (!d) return 0;
  return rn::carve(d, nullptr, 0, nullptr) + 256;
}

extern "C" int rn_proposal_fwd(const rn_proposal_desc* d, const float* scales_host, const float* ratios_host,
                               const float* cls_prob, const float* bbox_pred, const float* im_info, float* rois_out,
                               float* scores_out, int32_t* num_kept_out, void* wsp, size_t ws_bytes, rn_stream_t stream) {
  using namespace rn;
  RN_CHECK_ARG(d && scales_host && ratios_host && cls_prob && bbox_pred && im_info && rois_out && wsp,
               "rn_proposal_fwd: null argument");
  const int A = d->num_scales * d->num_ratios;
  RN_CHECK_ARG(A > 0 && A <= kMaxAnchors, "rn_proposal_fwd: %d anchors per cell unsupported (1..%d)", A, kMaxAnchors);
  RN_CHECK_ARG(d->post_nms_top_n > 0, "rn_proposal_fwd: post_nms_top_n must be > 0");
  const int n_max = d->Hf * d->Wf * A;
  RN_CHECK_ARG(n_max > 0 && n_max <= 32768, "rn_proposal_fwd: %d anchors exceed the single-CTA sort capacity 32768", n_max);
  ProposalWs w;
  if (!carve(d, wsp, ws_bytes, &w)) { set_error("rn_proposal_fwd: workspace too small (%zu < %zu)", ws_bytes, rn_proposal_workspace_bytes(d)); return RN_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  AnchorSet anc;
  generate_anchors_host(d->feat_stride, ratios_host, d->num_ratios, scales_host, d->num_scales, &anc);
  int* n_total = w.counters; int* n_pre = w.counters + 1; int* n_kept = w.counters + 2;
  proposal_decode_kernel<<<cdiv(n_max, 256), 256, 0, st>>>(anc, cls_prob, bbox_pred, im_info, d->Hf, d->Wf,
                                                           d->feat_stride, d->min_size, w.props, w.keys, n_total);
  RN_LAUNCH_CHECK();
  const int num_chunks = cdiv(n_max, kChunk);
  const size_t rank_smem = (size_t)num_chunks * (kKS * 4 + kIS * 2) + 16;
  static thread_local size_t configured = 0;
  if (rank_smem > configured) {
    RN_CUDA(cudaFuncSetAttribute(proposal_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rank_smem));
    configured = rank_smem;
  }
  proposal_sort_chunks_kernel<<<num_chunks, 512, 0, st>>>(w.keys, n_total, w.skeys, w.sidx);
  RN_LAUNCH_CHECK();
  {
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int g = std::min(sms, cdiv(num_chunks * kChunk, 256));
    proposal_rank_kernel<<<g, 1024, rank_smem, st>>>(w.skeys, w.sidx, num_chunks, d->pre_nms_top_n, w.order, n_pre);
    RN_LAUNCH_CHECK();
  }
  proposal_gather_kernel<<<cdiv(w.pre, 256), 256, 0, st>>>(w.props, cls_prob, w.order, n_pre, A, d->Hf, d->Wf, im_info,
                                                           d->feat_stride, w.det);
  RN_LAUNCH_CHECK();
  int r = launch_nms(st, w.det, 5, n_pre, w.pre, d->nms_thresh, d->post_nms_top_n, w.mask, w.keep, n_kept);
  if (r) return r;
  proposal_emit_kernel<<<cdiv(d->post_nms_top_n, 128), 128, 0, st>>>(w.det, w.keep, n_kept, d->post_nms_top_n, rois_out,
                                                                     scores_out, num_kept_out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" size_t rn_nms_workspace_bytes(int32_t n) {
  const size_t cb = (((size_t)n + 63) / 64 + 1) & ~(size_t)1;
  return rn::ws_slice(((size_t)n + 63) / 64 * 64 * cb, 8) + rn::ws_slice(4, 4) + 256;
}

extern "C" int rn_nms(const float* boxes_sorted, int32_t n, int32_t box_dim, float thresh, int32_t max_keep,
                      int32_t* keep_out, int32_t* num_out, void* wsp, size_t ws_bytes, rn_stream_t stream) {
  using namespace rn;
  RN_CHECK_ARG(keep_out && num_out && n >= 0 && box_dim >= 4 && max_keep > 0, "rn_nms: bad arguments");
  RN_CHECK_ARG(n == 0 || (boxes_sorted && wsp), "rn_nms: null boxes / workspace");
  cudaStream_t st = (cudaStream_t)stream;
  RN_CUDA(cudaMemsetAsync(num_out, 0, sizeof(int), st));
  if (n == 0) return RN_OK;
  Workspace ws(wsp, ws_bytes);
  const int cb = ((n + 63) / 64 + 1) & ~1;
  unsigned long long* mask = ws.take<unsigned long long>(((size_t)n + 63) / 64 * 64 * cb);
  int* n_dev = ws.take<int>(4);
  if (!n_dev) { set_error("rn_nms: workspace too small"); return RN_ERR_WORKSPACE; }

  // the mask/sweep kernels read n from device memory (in rn_proposal_fwd it is produced on the device); here it is a
  // host value, staged by a 1-thread kernel so the call stays asynchronous
  set_int_kernel<<<1, 1, 0, st>>>(n_dev, n);
  RN_LAUNCH_CHECK();
  return launch_nms(st, boxes_sorted, box_dim, n_dev, n, thresh, max_keep, mask, keep_out, num_out);
}

extern "C" int rn_bbox_overlaps(const double* boxes, const double* query, int32_t N, int32_t K, double* out,
                                rn_stream_t stream) {
  RN_CHECK_ARG(N >= 0 && K >= 0, "rn_bbox_overlaps: bad sizes");
  if (N == 0 || K == 0) return RN_OK;
  RN_CHECK_ARG(boxes && query && out, "rn_bbox_overlaps: null pointer");
  rn::bbox_overlaps_kernel<<<rn::cdiv(N * K, 256), 256, 0, (cudaStream_t)stream>>>(boxes, query, N, K, out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_proposal_target_fwd(const rn_proposal_target_desc* d, const float* rois, const float* gt_boxes,
                                      float* rois_out, float* label, float* bbox_target, float* bbox_weight,
                                      rn_stream_t stream) {
  RN_CHECK_ARG(d && rois_out && label && bbox_target && bbox_weight && d->N >= 0 && d->G >= 0, "rn_proposal_target_fwd: bad arguments");
  RN_CHECK_ARG((d->N == 0 || rois) && (d->G == 0 || gt_boxes), "rn_proposal_target_fwd: null input");
  if (d->N + d->G == 0) return RN_OK;
  rn::proposal_target_kernel<<<rn::cdiv(d->N + d->G, 128), 128, 0, (cudaStream_t)stream>>>(*d, rois, gt_boxes, rois_out,
                                                                                         label, bbox_target, bbox_weight);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
