// relation_bwd.cu -- rn_relation_bwd: gradients of the object-relation module.  Contractions: the repo's tcgen05 tf32 GEMM
// (gemm_tf32.cu) under RN_PREC_F16, cuBLAS fp32 under RN_PREC_FP32 (parity mode); everything else small fp32 kernels.
//
// The reference has no hand-written backward for this path: MXNet differentiates the symbol graph of
// attention_module_multi_head (resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16.py:104-151) op by op
// and keeps every intermediate ([N,M,64] embedding, [N,H,M] logits, softmax ...) alive for it.  Here nothing is kept:
// the forward intermediates are RECOMPUTED into the workspace (relation_fp32), then with dO = dOut * [out > 0]:
//     dbout = colsum dO                  dV'_h = P_h^T dO_h                 dP_h = dO_h V'_h^T
//     dS    = P o (dP - rowsum(P o dP))  (softmax over keys of  log g + s/sqrt(dk))
//     dQ_h  = dS_h K_h / sqrt(dk)        dK_h  = dS_h^T Q_h / sqrt(dk)
//     dx    = [g > 1e-6] dS / g          (x = Wg.phi + bg; the max(.,1e-6) clamp passes no gradient)
//     dWg   = sum_{n,m} dx phi^T, dbg = sum dx   -- phi re-evaluated from the boxes per 128-pair tile, never stored
//     dWq = dQ^T X, dWk = dK^T Xk, dWout = dV'^T Xk, biases = column sums
//     dX = dO (residual) + dQ Wq, and dX[key rows] += dK Wk + dV' Wout
// Boxes carry no gradient (they come from zero-gradient custom ops: proposal.py:170-173).
// HBM layout as in relation.cu; extra scratch: Y, dO [B*N,dout], dP [B,H,N,ld], dQ, dK, dV', dXk, per-CTA dWg partials.
#include "common.cuh"
#include "geom.cuh"
#include "relation.cuh"
#include <algorithm>

namespace rn {

__global__ void relu_mask_kernel(const float* __restrict__ Y, const float* __restrict__ dOut, size_t total, int masked,
                                 float* __restrict__ dO, float* __restrict__ dX) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float v = masked ? (Y[i] > 0.f ? dOut[i] : 0.f) : dOut[i];
    dO[i] = v;
    if (masked) dX[i] = v;            // residual branch: d(X + o)/dX = I
  }
}

// out[c] = sum_r A[r][c].  CTA = 32 columns x 32 row lanes: row lane y sums rows y, y + 32, ... (a warp reads 128 contiguous
// bytes per row), then a shared-memory tree over the row lanes.  Deterministic (no atomics); rows x 128 B per CTA.
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ A, int rows, int cols, float* __restrict__ out) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f;
  if (c < cols) {
    int r = threadIdx.y;
    for (; r + 32 < rows; r += 64) { s0 += A[(size_t)r * cols + c]; s1 += A[(size_t)(r + 32) * cols + c]; }
    if (r < rows) s0 += A[(size_t)r * cols + c];
  }
  red[threadIdx.y][threadIdx.x] = s0 + s1;
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) {
    if (threadIdx.y < o) red[threadIdx.y][threadIdx.x] += red[threadIdx.y + o][threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.y == 0 && c < cols) out[c] = red[0][threadIdx.x];
}

// warp per (b,h,n) row: dP <- scale * P o (dP - sum(P o dP));  g <- [g > 1e-6] * P o (dP - sum) / g
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ dP,
                                                               float* __restrict__ g, int rows, int M, int ld,
                                                               float scale) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* p = P + (size_t)warp * ld;
  float* dp = dP + (size_t)warp * ld;
  float* gr = g + (size_t)warp * ld;
  float sum = 0.f;
  for (int m = lane; m < M; m += 32) sum = fmaf(p[m], dp[m], sum);
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int m = lane; m < M; m += 32) {
    const float ds = p[m] * (dp[m] - sum);
    const float gv = gr[m];
    dp[m] = ds * scale;
    gr[m] = gv > 1e-6f ? ds / gv : 0.f;
  }
  for (int m = M + lane; m < ld; m += 32) { dp[m] = 0.f; gr[m] = 0.f; }
}

// dWg / dbg partials.  CTA = 128 threads = 128 (query n, key m) pairs of one query row; phi (E columns + a constant 1
// column for the bias) goes to shared memory, then every thread owns up to 17 of the H*(E+1) outputs and runs the
// 128-long dot products.  Per-CTA partial sums are written out and reduced by reduce_partials_kernel (deterministic).
constexpr int kGJ = 17;
__global__ void __launch_bounds__(128) geom_grad_kernel(const float* __restrict__ boxes, const int* __restrict__ key_index,
                                                        const float* __restrict__ dx, int B, int N, int M, int H, int E,
                                                        int ld, GeomFreq fr, float* __restrict__ partial) {
  extern __shared__ float sm[];
  const int EE = E + 1;
  float* phi_s = sm;                    // [128][EE]  (EE odd: conflict-free column writes)
  float* dx_s = sm + 128 * EE;          // [H][128]
  const int tid = threadIdx.x;
  const int nout = H * EE;
  float acc[kGJ];
#pragma unroll
  for (int j = 0; j < kGJ; ++j) acc[j] = 0.f;
  const int tiles_m = (M + 127) >> 7;
  const long long total = (long long)B * N * tiles_m;
  const int nf = E / 8;
  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int tm = (int)(item % tiles_m);
    const int n = (int)((item / tiles_m) % N), b = (int)(item / ((long long)tiles_m * N));
    const int m = tm * 128 + tid;
    __syncthreads();                    // previous tile's dot products are done
    if (m < M) {
      const float4 bn = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + n];
      const float4 bm = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + (key_index ? key_index[m] : m)];
      float eps[4];
      pair_eps(bn, bm, eps);
      float* row = phi_s + tid * EE;
      for (int c = 0; c < 4; ++c) {
        const float a = 100.0f * eps[c];
        for (int k = 0; k < nf; ++k) {
          float s, co;
          sincosf(a / fr.dim[k], &s, &co);
          row[c * 2 * nf + k] = s;
          row[c * 2 * nf + nf + k] = co;
        }
      }
      row[E] = 1.f;
      for (int h = 0; h < H; ++h) dx_s[h * 128 + tid] = dx[(((size_t)b * H + h) * N + n) * ld + m];
    } else {
      for (int h = 0; h < H; ++h) dx_s[h * 128 + tid] = 0.f;
      float* row = phi_s + tid * EE;
      for (int e = 0; e < EE; ++e) row[e] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kGJ; ++j) {
      const int o = tid + 128 * j;
      if (o < nout) {
        const int h = o / EE, e = o - h * EE;
        const float* dr = dx_s + h * 128;
        float a = acc[j];
#pragma unroll 8
        for (int p = 0; p < 128; ++p) a = fmaf(dr[p], phi_s[p * EE + e], a);
        acc[j] = a;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kGJ; ++j) {
    const int o = tid + 128 * j;
    if (o < nout) partial[(size_t)blockIdx.x * nout + o] = acc[j];
  }
}

// E = 64 form of the kernel above with a register tile: 160 threads build phi for 128 pairs (rows of 68 floats: 64 phi
// columns, the constant 1 of the bias, 3 zeros -> 16-byte aligned float4 reads), then thread (hq, eq), hq < 8, eq < 17, owns
// the 2 heads x 4 columns block (2 hq .. 2 hq + 1, 4 eq .. 4 eq + 3): per pair 2 scalar + 1 float4 shared loads feed 8 FMAs
// (the one-output-per-thread form above issues 2 loads per FMA and was LDS bound: 84 us at N = 300, 618 us for the 80-class
// learn-NMS head; profiles/r02_launches_bwd_*.csv).  phi: two-step Cody-Waite reduction + MUFU (|err| ~ 1e-6), as in the forward.
constexpr int kGP = 68;
__device__ __forceinline__ void sincos_cw(float x, float* s, float* c) {
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, x);
  r = fmaf(n, 1.7484555e-7f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}
__global__ void __launch_bounds__(160) geom_grad_tiled_kernel(const float* __restrict__ boxes, const int* __restrict__ key_index,
                                                              const float* __restrict__ dx, int B, int N, int M, int H, int ld,
                                                              GeomFreq fr, float* __restrict__ partial) {
  __shared__ __align__(16) float phi_s[128 * kGP];
  __shared__ float dx_s[16 * 128];
  const int tid = threadIdx.x;
  const int hq = tid / 17, eq = tid - hq * 17;               // compute role (tid < 136)
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  const int tiles_m = (M + 127) >> 7;
  const long long total = (long long)B * N * tiles_m;
  for (long long item = blockIdx.x; item < total; item += gridDim.x) {
    const int tm = (int)(item % tiles_m);
    const int n = (int)((item / tiles_m) % N), b = (int)(item / ((long long)tiles_m * N));
    __syncthreads();                                          // previous tile's dot products are done
    if (tid < 128) {
      const int m = tm * 128 + tid;
      float* row = phi_s + tid * kGP;
      if (m < M) {
        const float4 bn = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + n];
        const float4 bm = reinterpret_cast<const float4*>(boxes)[(size_t)b * N + (key_index ? key_index[m] : m)];
        float eps[4];
        pair_eps(bn, bm, eps);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float a = 100.0f * eps[c];
#pragma unroll
          for (int k = 0; k < 8; ++k) sincos_cw(a / fr.dim[k], &row[c * 16 + k], &row[c * 16 + 8 + k]);
        }
        row[64] = 1.f; row[65] = 0.f; row[66] = 0.f; row[67] = 0.f;
        for (int h = 0; h < 16; ++h) dx_s[h * 128 + tid] = h < H ? dx[(((size_t)b * H + h) * N + n) * ld + m] : 0.f;
      } else {
#pragma unroll 4
        for (int e = 0; e < kGP; ++e) row[e] = 0.f;
        for (int h = 0; h < 16; ++h) dx_s[h * 128 + tid] = 0.f;
      }
    }
    __syncthreads();
    if (tid < 136) {
      const float* d0 = dx_s + (2 * hq) * 128;
      const float* d1 = d0 + 128;
      const float4* ph = reinterpret_cast<const float4*>(phi_s) + eq;
#pragma unroll 8
      for (int p = 0; p < 128; ++p) {
        const float4 f = ph[p * (kGP / 4)];
        const float a0 = d0[p], a1 = d1[p];
        acc[0][0] = fmaf(a0, f.x, acc[0][0]); acc[0][1] = fmaf(a0, f.y, acc[0][1]);
        acc[0][2] = fmaf(a0, f.z, acc[0][2]); acc[0][3] = fmaf(a0, f.w, acc[0][3]);
        acc[1][0] = fmaf(a1, f.x, acc[1][0]); acc[1][1] = fmaf(a1, f.y, acc[1][1]);
        acc[1][2] = fmaf(a1, f.z, acc[1][2]); acc[1][3] = fmaf(a1, f.w, acc[1][3]);
      }
    }
  }
  if (tid < 136) {
    const int nout = H * 65;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int h = 2 * hq + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = 4 * eq + j;
        if (h < H && e < 65) partial[(size_t)blockIdx.x * nout + h * 65 + e] = acc[i][j];
      }
    }
  }
}

// one warp per output: lanes stride over the per-CTA partials, shuffle tree (fixed order: deterministic)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int nblocks, int H, int E,
                                                              float* __restrict__ dWg, float* __restrict__ dbg) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int EE = E + 1, nout = H * EE;
  if (o >= nout) return;
  float s = 0.f;
  for (int k = lane; k < nblocks; k += 32) s += partial[(size_t)k * nout + o];
#pragma unroll
  for (int w = 16; w > 0; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
  if (lane == 0) {
    const int h = o / EE, e = o - h * EE;
    if (e == E) dbg[h] = s; else dWg[h * E + e] = s;
  }
}

// dX[b, key(m)] += dXk[b, m]
__global__ void scatter_add_rows_kernel(const float* __restrict__ dXk, const int* __restrict__ idx, int N, int M, int D,
                                        float* __restrict__ dX) {
  const int m = blockIdx.x, b = blockIdx.y;
  const float* src = dXk + ((size_t)b * M + m) * D;
  float* dst = dX + ((size_t)b * N + (idx ? idx[m] : m)) * D;
  if (idx) {
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(dst + i, src[i]);   // duplicates in key_index stay correct
  } else {
    for (int i = threadIdx.x; i < D; i += blockDim.x) dst[i] += src[i];
  }
}

int launch_colsum(cudaStream_t st, const float* A, int rows, int cols, float* out) {
  colsum_kernel<<<cdiv(cols, 32), dim3(32, 32), 0, st>>>(A, rows, cols, out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

static int geom_grad_grid(const rn_relation_desc* d) {
  const long long items = (long long)d->batch * d->N * cdiv(d->M, 128);
  const int sms = sm_count() > 0 ? sm_count() : 148;
  // 5 CTAs per SM: the per-tile work is a long dependent chain (log / sincos of phi, then 128 accumulation steps), so it is
  // latency bound at 5 warps per CTA -- resident CTAs, not issue slots, set the rate (43 KB smem / 160 threads each)
  return (int)std::max<long long>(1, std::min<long long>(items, (long long)sms * 5));
}

size_t relation_bwd_ws_bytes(const rn_relation_desc* d) {
  const size_t B = d->batch, N = d->N, M = d->M, H = d->H, ld = align_up(M, 4);
  size_t t = relation_fp32_ws_bytes(d);
  t += 2 * ws_slice(B * N * d->dout, 4);          // Y, dO
  t += ws_slice(B * H * N * ld, 4);               // dP / dS
  t += ws_slice(B * N * d->dq, 4);                // dQ
  t += ws_slice(B * M * d->dq, 4);                // dK
  t += ws_slice(B * M * d->dout, 4);              // dV'
  t += ws_slice(B * M * d->d, 4);                 // dXk
  t += ws_slice((size_t)geom_grad_grid(d) * H * (d->E + 1), 4);   // dWg partials (grid <= 2 CTAs per SM)
  return t;
}

int relation_bwd(const rn_relation_desc* d, const float* X, const float* boxes, const int* key_index,
                        const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wg,
                        const float* bg, const float* Wout, const float* bout, const float* dOut, float* dX, float* dWq,
                        float* dbq, float* dWk, float* dbk, float* dWg, float* dbg, float* dWout, float* dbout,
                        void* wsp, size_t ws_bytes, cudaStream_t st, const Fp32State* forward_state, const float* forward_out,
                        const float* mask_out) {
  const int B = d->batch, N = d->N, M = d->M, D = d->d, dq = d->dq, dout = d->dout, H = d->H, E = d->E;
  const int dk = dq / H, dv = dout / H;
  // forward_state / forward_out: the intermediates (Q, K, V', g, P, gathered keys) and output of a relation_fp32 call the
  // caller has just made with the same arguments (the learn-NMS backward recomputes its whole forward anyway): no second
  // recomputation here.  g is consumed (overwritten with dx) either way.
  Fp32State fs;
  if (forward_state) {
    fs = *forward_state;
    fs.used = 0;
  } else if (!relation_fp32_carve(d, wsp, ws_bytes, &fs)) {
    set_error("rn_relation_bwd: workspace too small (%zu < %zu)", ws_bytes, relation_bwd_ws_bytes(d));
    return RN_ERR_WORKSPACE;
  }
  const int ld = fs.ld;
  const int ggrid = geom_grad_grid(d);
  Workspace ws((char*)wsp + fs.used, ws_bytes - fs.used);
  float* Y = ws.take<float>((size_t)B * N * dout);
  float* dO = ws.take<float>((size_t)B * N * dout);
  float* dP = ws.take<float>((size_t)B * H * N * ld);
  float* dQ = ws.take<float>((size_t)B * N * dq);
  float* dK = ws.take<float>((size_t)B * M * dq);
  float* dVp = ws.take<float>((size_t)B * M * dout);
  float* dXk = ws.take<float>((size_t)B * M * D);
  float* partial = ws.take<float>((size_t)ggrid * H * (E + 1));
  if (!partial) {
    set_error("rn_relation_bwd: workspace too small (%zu < %zu)", ws_bytes, relation_bwd_ws_bytes(d));
    return RN_ERR_WORKSPACE;
  }
  GeomFreq fr;
  int r;
  if ((r = make_freq(E, d->wave_length, &fr))) return r;
  // 1. recompute the forward: leaves Q, K, V', g, P (in fs.S), gathered keys (fs.Xk) in the workspace, out in Y
  if (forward_state) Y = const_cast<float*>(forward_out);
  else if ((r = relation_fp32(d, X, boxes, key_index, Wq, bq, Wk, bk, Wg, bg, Wout, bout, Y, nullptr, wsp, fs.used, st))) return r;
  const float* P = fs.S;
  // 2. dO, residual part of dX
  {
    const size_t total = (size_t)B * N * dout;
    if (!d->fuse_residual_relu) RN_CUDA(cudaMemsetAsync(dX, 0, (size_t)B * N * D * sizeof(float), st));
    int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    // relu mask: the executed forward's output when the caller has it, else the recomputed one
    relu_mask_kernel<<<blocks, 256, 0, st>>>(mask_out ? mask_out : Y, dOut, total, d->fuse_residual_relu, dO, dX);
    RN_LAUNCH_CHECK();
    colsum_kernel<<<cdiv(dout, 32), dim3(32, 32), 0, st>>>(dO, B * N, dout, dbout);
    RN_LAUNCH_CHECK();
  }
  // 3. dV'_h = P_h^T dO_h,  dP_h = dO_h V'_h^T   (heads batched)
  const long long sS = (long long)H * N * ld, sSh = (long long)N * ld;      // [B,H,N,ld] tensors: problem / head strides
  if ((r = sgemm_rm_2level(st, true, false, M, dv, N, 1.f, P, ld, sS, sSh, dO, dout, (long long)N * dout, dv, 0.f, dVp, dout,
                           (long long)M * dout, dv, B, H, fs.ptrs))) return r;
  if ((r = sgemm_rm_2level(st, false, true, N, M, dv, 1.f, dO, dout, (long long)N * dout, dv, fs.Vp, dout, (long long)M * dout,
                           dv, 0.f, dP, ld, sS, sSh, B, H, fs.ptrs))) return r;
  // 4. softmax backward; dP <- dS/sqrt(dk), g <- dx
  {
    const int rows = B * H * N;
    softmax_bwd_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(P, dP, fs.g, rows, M, ld, 1.0f / sqrtf((float)dk));
    RN_LAUNCH_CHECK();
  }
  // 5. dQ_h = dS_h K_h, dK_h = dS_h^T Q_h
  if ((r = sgemm_rm_2level(st, false, false, N, dk, M, 1.f, dP, ld, sS, sSh, fs.K, dq, (long long)M * dq, dk, 0.f, dQ, dq,
                           (long long)N * dq, dk, B, H, fs.ptrs))) return r;
  if ((r = sgemm_rm_2level(st, true, false, M, dk, N, 1.f, dP, ld, sS, sSh, fs.Q, dq, (long long)N * dq, dk, 0.f, dK, dq,
                           (long long)M * dq, dk, B, H, fs.ptrs))) return r;
  // 6. geometry FC gradients
  {
    const size_t smem = ((size_t)128 * (E + 1) + (size_t)H * 128) * sizeof(float);
    RN_CUDA(cudaFuncSetAttribute(geom_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (E == 64 && H <= 16) geom_grad_tiled_kernel<<<ggrid, 160, 0, st>>>(boxes, key_index, fs.g, B, N, M, H, ld, fr, partial);
    else geom_grad_kernel<<<ggrid, 128, smem, st>>>(boxes, key_index, fs.g, B, N, M, H, E, ld, fr, partial);
    RN_LAUNCH_CHECK();
    reduce_partials_kernel<<<cdiv(H * (E + 1) * 32, 256), 256, 0, st>>>(partial, ggrid, H, E, dWg, dbg);
    RN_LAUNCH_CHECK();
  }
  // 7. projection weights / biases
  if ((r = sgemm_rm(st, true, false, dq, D, B * N, 1.f, dQ, dq, X, D, 0.f, dWq, D))) return r;
  colsum_kernel<<<cdiv(dq, 32), dim3(32, 32), 0, st>>>(dQ, B * N, dq, dbq);
  RN_LAUNCH_CHECK();
  colsum_kernel<<<cdiv(dq, 32), dim3(32, 32), 0, st>>>(dK, B * M, dq, dbk);
  RN_LAUNCH_CHECK();
  if (key_index || M == N || B == 1) {
    // key rows are one contiguous [B*M, D] matrix (gathered copy, or X itself): a single GEMM each
    const float* keys = key_index ? fs.Xk : X;
    if ((r = sgemm_rm(st, true, false, dq, D, B * M, 1.f, dK, dq, keys, D, 0.f, dWk, D))) return r;
    if ((r = sgemm_rm(st, true, false, dout, D, B * M, 1.f, dVp, dout, keys, D, 0.f, dWout, D))) return r;
  } else {
    for (int b = 0; b < B; ++b) {
      const float* keys = X + (size_t)b * N * D;
      const float beta = b ? 1.f : 0.f;
      if ((r = sgemm_rm(st, true, false, dq, D, M, 1.f, dK + (size_t)b * M * dq, dq, keys, D, beta, dWk, D))) return r;
      if ((r = sgemm_rm(st, true, false, dout, D, M, 1.f, dVp + (size_t)b * M * dout, dout, keys, D, beta, dWout, D))) return r;
    }
  }
  // 8. input gradient
  if ((r = sgemm_rm(st, false, false, B * N, D, dq, 1.f, dQ, dq, Wq, D, 1.f, dX, D))) return r;
  if ((r = sgemm_rm(st, false, false, B * M, D, dq, 1.f, dK, dq, Wk, D, 0.f, dXk, D))) return r;
  if ((r = sgemm_rm(st, false, false, B * M, D, dout, 1.f, dVp, dout, Wout, D, 1.f, dXk, D))) return r;
  scatter_add_rows_kernel<<<dim3(M, B), 128, 0, st>>>(dXk, key_index, N, M, D, dX);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn

extern "C" size_t rn_relation_bwd_workspace_bytes(const rn_relation_desc* d) {
  if (!d || d->H < 1 || d->M < 1) return 0;
  return rn::relation_bwd_ws_bytes(d) + 256;
}

extern "C" int rn_relation_bwd_masked(const rn_relation_desc* d, const float* X, const float* boxes, const int32_t* key_index,
                               const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wg,
                               const float* bg, const float* Wout, const float* bout, const float* out_fwd, const float* dOut, float* dX,
                               float* dWq, float* dbq, float* dWk, float* dbk, float* dWg, float* dbg, float* dWout,
                               float* dbout, void* ws, size_t ws_bytes, rn_stream_t stream) {
  int r = rn::relation_check_desc(d);
  if (r) return r;
  if ((r = rn::relation_check_keys(d, key_index))) return r;
  RN_CHECK_ARG(X && boxes && Wq && bq && Wk && bk && Wg && bg && Wout && bout && dOut && ws,
               "rn_relation_bwd: null input pointer");
  RN_CHECK_ARG(dX && dWq && dbq && dWk && dbk && dWg && dbg && dWout && dbout, "rn_relation_bwd: null gradient pointer");
  RN_CHECK_ARG(d->E >= 8 && d->E % 8 == 0 && d->E <= 128, "rn_relation_bwd: E=%d unsupported (multiple of 8, <= 128)", d->E);
  // d->precision selects the contraction engine of the backward (recomputed forward included): RN_PREC_F16 = the repo's
  // tcgen05 tf32 GEMM (gemm_tf32.cu: fp32 operands read by TMA, kind::tf32, fp32 accumulate), RN_PREC_FP32 = cuBLAS fp32
  rn::GemmBackendScope backend(d->precision == RN_PREC_F16 && rn::is_sm100() ? 1 : 0);
  return rn::relation_bwd(d, X, boxes, key_index, Wq, bq, Wk, bk, Wg, bg, Wout, bout, dOut, dX, dWq, dbq, dWk, dbk, dWg,
                          dbg, dWout, dbout, ws, ws_bytes, (cudaStream_t)stream, nullptr, nullptr, out_fwd);
}

extern "C" int rn_relation_bwd(const rn_relation_desc* d, const float* X, const float* boxes, const int32_t* key_index,
                               const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wg,
                               const float* bg, const float* Wout, const float* bout, const float* dOut, float* dX,
                               float* dWq, float* dbq, float* dWk, float* dbk, float* dWg, float* dbg, float* dWout,
                               float* dbout, void* ws, size_t ws_bytes, rn_stream_t stream) {
  return rn_relation_bwd_masked(d, X, boxes, key_index, Wq, bq, Wk, bk, Wg, bg, Wout, bout, nullptr, dOut, dX, dWq, dbq, dWk, dbk,
                                dWg, dbg, dWout, dbout, ws, ws_bytes, stream);
}
