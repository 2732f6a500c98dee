// relation.cu -- rn_relation_fwd / rn_linear_fwd: argument checking, workspace carving and the fp32 parity path.
//
// RN_PREC_FP32 evaluates the module with the reference's own op decomposition (SYM_REL:104-151) on library GEMMs:
//   Q,K,V' projections (cuBLAS SGEMM, pedantic fp32) -> geometry weight kernel (geom.cu) -> S_h = Q_h K_h^T (batched
//   SGEMM over heads) -> row softmax of log g + s/sqrt(dk), evaluated as g*exp(s - max) (exp(log g) == g) -> O_h =
//   P_h V'_h (batched SGEMM) -> bias / residual / relu epilogue.   V' = X_keys Wout^T is the SURVEY section 3.3
//   identity (project once instead of P.V over the full d then a grouped 1x1 conv).
// RN_PREC_F16 dispatches to the fused tcgen05 kernels in relation_tc.cu (sm_100a).
#include "common.cuh"
#include "geom.cuh"
#include "relation.cuh"
#include "gemm_tc.cuh"

namespace rn {

__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, size_t rows, int cols, int relu) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = rows * (size_t)cols;
  for (; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float v = y[i] + (bias ? bias[i % cols] : 0.f);
    y[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

int launch_bias_act(cudaStream_t st, float* y, const float* bias, size_t rows, int cols, int relu) {
  if (!bias && !relu) return RN_OK;
  size_t total = rows * (size_t)cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bias_act_kernel<<<blocks, 256, 0, st>>>(y, bias, rows, cols, relu);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

__global__ void gather_rows_kernel(const float* __restrict__ X, const int* __restrict__ idx, int B, int N, int M, int d,
                                   float* __restrict__ out) {
  const int m = blockIdx.x, b = blockIdx.y;
  const float* src = X + ((size_t)b * N + idx[m]) * d;
  float* dst = out + ((size_t)b * M + m) * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) dst[i] = src[i];
}

// one warp per (b,h,n) row: p = g*exp2((s - max s)*c) / sum   (in place over s; optional copy to softmax_out [B,N,H,M])
__global__ void __launch_bounds__(256) geo_softmax_rows_kernel(float* __restrict__ S, const float* __restrict__ g,
                                                               int B, int H, int N, int M, int ld, float scale,
                                                               float* __restrict__ softmax_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int rows = B * H * N;
  if (warp >= rows) return;
  float* s = S + (size_t)warp * ld;
  const float* gr = g + (size_t)warp * ld;
  // the reference takes the max over log g + s; any stabiliser gives the same softmax -- we use the same one
  float mx = -INFINITY;
  for (int m = lane; m < M; m += 32) mx = fmaxf(mx, logf(gr[m]) + s[m] * scale);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int m = lane; m < M; m += 32) {
    float e = expf(logf(gr[m]) + s[m] * scale - mx);
    s[m] = e; sum += e;
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  const int n = warp % N, h = (warp / N) % H, b = warp / (N * H);
  for (int m = lane; m < M; m += 32) {
    float p = s[m] * inv;
    s[m] = p;
    if (softmax_out) softmax_out[(((size_t)b * N + n) * H + h) * M + m] = p;
  }
}

// out = (residual_relu ? relu(X + o + bout) : o + bout)
__global__ void relation_epilogue_kernel(const float* __restrict__ O, const float* __restrict__ bout,
                                         const float* __restrict__ X, size_t rows, int dout, int residual_relu,
                                         float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = rows * (size_t)dout;
  for (; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float v = O[i] + bout[i % dout];
    if (residual_relu) v = fmaxf(X[i] + v, 0.f);
    out[i] = v;
  }
}

int relation_check_desc(const rn_relation_desc* d) {
  RN_CHECK_ARG(d, "rn_relation: null descriptor");
  RN_CHECK_ARG(d->batch >= 1 && d->N >= 1 && d->M >= 1 && d->d >= 1, "rn_relation: bad sizes batch=%d N=%d M=%d d=%d",
               d->batch, d->N, d->M, d->d);
  RN_CHECK_ARG(d->H >= 1 && d->H <= 16, "rn_relation: H=%d unsupported (1..16)", d->H);
  RN_CHECK_ARG(d->dq % d->H == 0 && d->dout % d->H == 0, "rn_relation: dq=%d / dout=%d not divisible by H=%d", d->dq,
               d->dout, d->H);
  RN_CHECK_ARG(!d->fuse_residual_relu || d->dout == d->d, "rn_relation: residual needs dout == d (%d vs %d)", d->dout,
               d->d);
  RN_CHECK_ARG(d->precision == RN_PREC_FP32 || d->precision == RN_PREC_F16 || d->precision == RN_PREC_TF32,
               "rn_relation: unknown precision %d", d->precision);
  RN_CHECK_ARG(d->dq >= d->H && d->dout >= d->H, "rn_relation: dq=%d / dout=%d must be at least H=%d", d->dq, d->dout, d->H);
  // the geometry kernels evaluate E / 8 frequencies per coordinate (4 coordinates x sin, cos): a remainder would be dropped
  RN_CHECK_ARG(d->E >= 8 && d->E % 8 == 0 && d->E <= 128, "rn_relation: E=%d unsupported (a multiple of 8 in 8..128)", d->E);
  return RN_OK;
}

// without a key index the keys are the FIRST M rows of X / boxes (SYM_REL:106 slice_axis): M > N would read past them
int relation_check_keys(const rn_relation_desc* d, const int* key_index) {
  RN_CHECK_ARG(key_index || d->M <= d->N, "rn_relation: M=%d keys but only N=%d rows and no key_index", d->M, d->N);
  return RN_OK;
}

size_t relation_fp32_ws_bytes(const rn_relation_desc* d) {
  size_t B = d->batch, N = d->N, M = d->M;
  int ld = (int)align_up(M, 4);
  size_t t = 0;
  t += ws_slice(B * N * d->dq, 4);      // Q
  t += ws_slice(B * M * d->dq, 4);      // K
  t += ws_slice(B * M * d->dout, 4);    // V'
  t += ws_slice(B * d->H * N * ld, 4);  // g
  t += ws_slice(B * d->H * N * ld, 4);  // S / P
  t += ws_slice(B * N * d->dout, 4);    // O
  t += ws_slice(B * M * d->d, 4);       // gathered keys
  t += ws_slice(3 * B * d->H, 8);       // pointer triples of the (problem, head) batched GEMMs
  return t;
}

// carve the fp32 path's intermediates out of a workspace (the backward re-uses them after recomputing the forward)
bool relation_fp32_carve(const rn_relation_desc* d, void* wsp, size_t ws_bytes, Fp32State* fs) {
  const size_t B = d->batch, N = d->N, M = d->M, H = d->H;
  const size_t ld = align_up(M, 4);
  Workspace ws(wsp, ws_bytes);
  fs->ld = (int)ld;
  fs->Q = ws.take<float>(B * N * d->dq);
  fs->K = ws.take<float>(B * M * d->dq);
  fs->Vp = ws.take<float>(B * M * d->dout);
  fs->g = ws.take<float>(B * H * N * ld);
  fs->S = ws.take<float>(B * H * N * ld);
  fs->O = ws.take<float>(B * N * d->dout);
  fs->Xk = ws.take<float>(B * M * d->d);
  fs->ptrs = ws.take<void*>(3 * B * H);
  fs->used = ws.off;
  return fs->ptrs != nullptr;
}

int relation_fp32(const rn_relation_desc* d, const float* X, const float* boxes, const int* key_index,
                         const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wg,
                         const float* bg, const float* Wout, const float* bout, float* out, float* softmax_out,
                         void* wsp, size_t ws_bytes, cudaStream_t st) {
  const int B = d->batch, N = d->N, M = d->M, D = d->d, dq = d->dq, dout = d->dout, H = d->H;
  const int dk = dq / H, dv = dout / H, ld = (int)align_up(M, 4);
  Fp32State fs;
  const bool carved = relation_fp32_carve(d, wsp, ws_bytes, &fs);
  float *Q = fs.Q, *K = fs.K, *Vp = fs.Vp, *g = fs.g, *S = fs.S, *O = fs.O, *Xk = fs.Xk;
  if (!carved) { set_error("rn_relation_fwd: workspace too small (%zu < %zu)", ws_bytes, relation_fp32_ws_bytes(d)); return RN_ERR_WORKSPACE; }
  int r;
  // projections
  if ((r = sgemm_nt(st, B * N, dq, D, X, D, Wq, D, Q, dq))) return r;
  if ((r = launch_bias_act(st, Q, bq, (size_t)B * N, dq, 0))) return r;
  const float* keys = X; long long key_stride = (long long)N * D;
  if (key_index) {
    gather_rows_kernel<<<dim3(M, B), 128, 0, st>>>(X, key_index, B, N, M, D, Xk);
    RN_LAUNCH_CHECK();
    keys = Xk; key_stride = (long long)M * D;
  }
  if (key_index || B == 1 || M == N) {
    // keys are contiguous rows [B*M, D] (gathered, or the whole X when M == N, or a single prefix)
    const int rows = (key_index || M == N) ? B * M : M;
    if ((r = sgemm_nt(st, rows, dq, D, keys, D, Wk, D, K, dq))) return r;
    if ((r = sgemm_nt(st, rows, dout, D, keys, D, Wout, D, Vp, dout))) return r;
  } else {
    if ((r = sgemm_nt(st, M, dq, D, keys, D, Wk, D, K, dq, B, key_stride, 0, (long long)M * dq))) return r;
    if ((r = sgemm_nt(st, M, dout, D, keys, D, Wout, D, Vp, dout, B, key_stride, 0, (long long)M * dout))) return r;
  }
  if ((r = launch_bias_act(st, K, bk, (size_t)B * M, dq, 0))) return r;
  // geometry weights g [B,H,N,ld]
  if ((r = launch_geom_weight(st, boxes, key_index, B, N, M, H, d->E, d->wave_length, Wg, bg, g, ld))) return r;
  // scores, softmax, aggregate (per problem; heads batched)
  if ((r = sgemm_rm_2level(st, false, true, N, M, dk, 1.f, Q, dq, (long long)N * dq, dk, K, dq, (long long)M * dq, dk, 0.f, S,
                           ld, (long long)H * N * ld, (long long)N * ld, B, H, fs.ptrs))) return r;
  {
    const int rows = B * H * N;
    geo_softmax_rows_kernel<<<cdiv(rows, 8), 256, 0, st>>>(S, g, B, H, N, M, ld, 1.0f / sqrtf((float)dk), softmax_out);
    RN_LAUNCH_CHECK();
  }
  if ((r = sgemm_rm_2level(st, false, false, N, dv, M, 1.f, S, ld, (long long)H * N * ld, (long long)N * ld, Vp, dout,
                           (long long)M * dout, dv, 0.f, O, dout, (long long)N * dout, dv, B, H, fs.ptrs))) return r;
  {
    size_t total = (size_t)B * N * dout;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    relation_epilogue_kernel<<<blocks, 256, 0, st>>>(O, bout, X, (size_t)B * N, dout, d->fuse_residual_relu, out);
    RN_LAUNCH_CHECK();
  }
  return RN_OK;
}

}  // namespace rn

extern "C" size_t rn_relation_workspace_bytes(const rn_relation_desc* d) {
  if (!d) return 0;
  // the two precisions have disjoint layouts: size for the one the descriptor asks for (the fused tcgen05 path keeps
  // nothing N x M, so its workspace stays small at any N)
  if (d->precision == RN_PREC_F16) return rn::relation_tc_workspace_bytes(d) + 256;
  return rn::relation_fp32_ws_bytes(d) + 256;
}

extern "C" int rn_relation_fwd(const rn_relation_desc* d, const float* X, const float* boxes, const int32_t* key_index,
                               const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wg,
                               const float* bg, const float* Wout, const float* bout, float* out, float* softmax_out,
                               void* ws, size_t ws_bytes, rn_stream_t stream) {
  int r = rn::relation_check_desc(d);
  if (r) return r;
  if ((r = rn::relation_check_keys(d, key_index))) return r;
  RN_CHECK_ARG(X && boxes && Wq && bq && Wk && bk && Wg && bg && Wout && bout && out && ws,
               "rn_relation_fwd: null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->precision == RN_PREC_F16)
    return rn::relation_tc(d, X, boxes, key_index, Wq, bq, Wk, bk, Wg, bg, Wout, bout, out, softmax_out, ws, ws_bytes, st);
  RN_CHECK_ARG(d->precision != RN_PREC_TF32 || rn::is_sm100(), "RN_PREC_TF32 needs an sm_100 device (tcgen05)");
  rn::GemmBackendScope backend(d->precision == RN_PREC_TF32 ? 1 : rn::gemm_backend());   // inside a backward: the caller's engine
  return rn::relation_fp32(d, X, boxes, key_index, Wq, bq, Wk, bk, Wg, bg, Wout, bout, out, softmax_out, ws, ws_bytes, st);
}

extern "C" size_t rn_linear_workspace_bytes(int32_t rows, int32_t in, int32_t out, int32_t precision) {
  if (precision == RN_PREC_F16) return rn::linear_tc_workspace_bytes(rows, in, out) + 256;
  return 256;
}

extern "C" int rn_linear_fwd(const float* x, const float* W, const float* b, float* y, int32_t rows, int32_t in,
                             int32_t out, int32_t relu, int32_t precision, void* ws, size_t ws_bytes,
                             rn_stream_t stream) {
  RN_CHECK_ARG(x && W && y && rows > 0 && in > 0 && out > 0, "rn_linear_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RN_PREC_F16) return rn::linear_tc(x, W, b, y, rows, in, out, relu, ws, ws_bytes, st);
  int r = rn::sgemm_nt(st, rows, out, in, x, in, W, in, y, out);
  if (r) return r;
  return rn::launch_bias_act(st, y, b, (size_t)rows, out, relu);
}

// ---- pre-packed fp16 operands (RN_PREC_F16): pack once per weight update, reuse every forward
extern "C" size_t rn_relation_packed_bytes(const rn_relation_desc* d) { return d ? rn::relation_tc_packed_bytes(d) : 0; }

extern "C" int rn_relation_pack(const rn_relation_desc* d, const float* Wq, const float* bq, const float* Wk,
                                const float* bk, const float* Wout, const float* bout, void* packed, rn_stream_t stream) {
  int r = rn::relation_check_desc(d);
  if (r) return r;
  RN_CHECK_ARG(Wq && bq && Wk && bk && Wout && bout && packed, "rn_relation_pack: null pointer argument");
  return rn::relation_tc_pack(d, Wq, bq, Wk, bk, Wout, bout, packed, (cudaStream_t)stream);
}

extern "C" int rn_relation_packed_fwd(const rn_relation_desc* d, const float* X, const float* boxes,
                                      const int32_t* key_index, const void* packed, const float* Wg, const float* bg,
                                      float* out, void* ws, size_t ws_bytes, rn_stream_t stream) {
  int r = rn::relation_check_desc(d);
  if (r) return r;
  if ((r = rn::relation_check_keys(d, key_index))) return r;
  RN_CHECK_ARG(X && boxes && packed && Wg && bg && out && ws, "rn_relation_packed_fwd: null pointer argument");
  return rn::relation_tc_packed(d, X, boxes, key_index, packed, Wg, bg, out, ws, ws_bytes, (cudaStream_t)stream, 7, nullptr);
}

// measurement hook: run only the selected stages (1 = cast + projection GEMM, 2 = geometry, 4 = fused attention) on the
// intermediates a previous full call left in the same workspace
extern "C" int rn_relation_packed_stages(const rn_relation_desc* d, const float* X, const float* boxes,
                                         const int32_t* key_index, const void* packed, const float* Wg, const float* bg,
                                         float* out, void* ws, size_t ws_bytes, int32_t stage_mask, rn_stream_t stream) {
  int r = rn::relation_check_desc(d);
  if (r) return r;
  if ((r = rn::relation_check_keys(d, key_index))) return r;
  RN_CHECK_ARG(X && boxes && packed && Wg && bg && out && ws, "rn_relation_packed_stages: null pointer argument");
  return rn::relation_tc_packed(d, X, boxes, key_index, packed, Wg, bg, out, ws, ws_bytes, (cudaStream_t)stream,
                                stage_mask, nullptr);
}

// fp16 in / fp16 out variant: X_f16 (optional) is the producer's fp16 copy of X, out_f16 (optional) receives an fp16 copy
// of the output for the consumer GEMM -- the cast launches between the layers of the head disappear
extern "C" int rn_relation_packed_fwd_f16io(const rn_relation_desc* d, const float* X, const void* X_f16,
                                            const float* boxes, const int32_t* key_index, const void* packed,
                                            const float* Wg, const float* bg, float* out, void* out_f16, void* ws,
                                            size_t ws_bytes, int32_t stage_mask, rn_stream_t stream) {
  int r = rn::relation_check_desc(d);
  if (r) return r;
  if ((r = rn::relation_check_keys(d, key_index))) return r;
  RN_CHECK_ARG(X && boxes && packed && Wg && bg && out && ws, "rn_relation_packed_fwd_f16io: null pointer argument");
  return rn::relation_tc_packed(d, X, boxes, key_index, packed, Wg, bg, out, ws, ws_bytes, (cudaStream_t)stream,
                                stage_mask, nullptr, X_f16, out_f16);
}

extern "C" size_t rn_linear_packed_bytes(int32_t in, int32_t out) { return rn::linear_tc_packed_bytes(in, out); }

extern "C" int rn_linear_pack(const float* W, int32_t in, int32_t out, void* packed, rn_stream_t stream) {
  RN_CHECK_ARG(W && packed && in > 0 && out > 0, "rn_linear_pack: bad arguments");
  return rn::linear_tc_pack(W, in, out, packed, (cudaStream_t)stream);
}

extern "C" int rn_linear_packed_fwd(const float* x, const void* packed_W, const float* b, float* y, int32_t rows,
                                    int32_t in, int32_t out, int32_t relu, void* ws, size_t ws_bytes, rn_stream_t stream) {
  RN_CHECK_ARG(x && packed_W && y && rows > 0 && in > 0 && out > 0 && ws, "rn_linear_packed_fwd: bad arguments");
  return rn::linear_tc_packed(x, packed_W, b, y, rows, in, out, relu, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int rn_linear_pack_chw_to_hwc(const float* W, int32_t out, int32_t C, int32_t S, void* packed,
                                         rn_stream_t stream) {
  RN_CHECK_ARG(W && packed && out > 0 && C > 0 && S > 0, "rn_linear_pack_chw_to_hwc: bad arguments");
  return rn::linear_tc_pack_chw_to_hwc(W, out, C, S, packed, (cudaStream_t)stream);
}

extern "C" int rn_linear_packed_f16in_fwd(const void* x_f16, const void* packed_W, const float* b, float* y, void* y_f16,
                                          int32_t rows, int32_t in, int32_t out, int32_t relu, void* ws, size_t ws_bytes,
                                          rn_stream_t stream) {
  RN_CHECK_ARG(x_f16 && packed_W && (y || y_f16) && rows > 0 && in > 0 && out > 0, "rn_linear_packed_f16in_fwd: bad arguments");
  return rn::linear_tc_packed_f16in(x_f16, packed_W, b, y, y_f16, rows, in, out, relu, ws, ws_bytes, (cudaStream_t)stream);
}

// ---- several FullyConnected layers over the same fp16 input as one tcgen05 GEMM
extern "C" size_t rn_linear_multi_packed_bytes(const int32_t* outs, int32_t nout, int32_t in) {
  return outs ? rn::linear_multi_packed_bytes(outs, nout, in) : 0;
}

extern "C" int rn_linear_multi_pack(const float* const* W, const float* const* b, const int32_t* outs, int32_t nout,
                                    int32_t in, void* packed, rn_stream_t stream) {
  RN_CHECK_ARG(W && outs && packed, "rn_linear_multi_pack: null argument");
  return rn::linear_multi_pack(W, b, outs, nout, in, packed, (cudaStream_t)stream);
}

extern "C" int rn_linear_multi_packed_f16in_fwd(const void* x_f16, const void* packed, float* const* ys, const int32_t* outs,
                                                int32_t nout, int32_t rows, int32_t in, void* ws, size_t ws_bytes,
                                                rn_stream_t stream) {
  RN_CHECK_ARG(x_f16 && packed && ys && outs && rows > 0 && in > 0, "rn_linear_multi_packed_f16in_fwd: bad arguments");
  return rn::linear_multi_packed_f16in(x_f16, packed, ys, outs, nout, rows, in, ws, ws_bytes, (cudaStream_t)stream);
}
