// relation.cuh -- internal interfaces between relation.cu (dispatch + fp32 path) and relation_tc.cu / gemm_tc.cu
#pragma once
#include "common.cuh"

namespace rn {

int launch_bias_act(cudaStream_t st, float* y, const float* bias, size_t rows, int cols, int relu);

// fp32 path (relation.cu) -- also the recomputation step of rn_relation_bwd (relation_bwd.cu)
struct Fp32State { float *Q, *K, *Vp, *g, *S, *O, *Xk; void** ptrs; int ld; size_t used; };
int relation_check_desc(const rn_relation_desc* d);
int relation_check_keys(const rn_relation_desc* d, const int* key_index);
size_t relation_fp32_ws_bytes(const rn_relation_desc* d);
bool relation_fp32_carve(const rn_relation_desc* d, void* ws, size_t ws_bytes, Fp32State* fs);
int relation_fp32(const rn_relation_desc* d, const float* X, const float* boxes, const int* key_index, const float* Wq,
                  const float* bq, const float* Wk, const float* bk, const float* Wg, const float* bg, const float* Wout,
                  const float* bout, float* out, float* softmax_out, void* ws, size_t ws_bytes, cudaStream_t st);

// backward (relation_bwd.cu); every gradient buffer is overwritten
size_t relation_bwd_ws_bytes(const rn_relation_desc* d);
int relation_bwd(const rn_relation_desc* d, const float* X, const float* boxes, const int* key_index, const float* Wq,
                 const float* bq, const float* Wk, const float* bk, const float* Wg, const float* bg, const float* Wout,
                 const float* bout, const float* dOut, float* dX, float* dWq, float* dbq, float* dWk, float* dbk,
                 float* dWg, float* dbg, float* dWout, float* dbout, void* ws, size_t ws_bytes, cudaStream_t st,
                 const Fp32State* forward_state = nullptr, const float* forward_out = nullptr,
                 const float* mask_out = nullptr);
int launch_colsum(cudaStream_t st, const float* A, int rows, int cols, float* out);   // out[c] = sum_r A[r][c]

// fused tcgen05 path (relation_tc.cu)
size_t relation_tc_workspace_bytes(const rn_relation_desc* d);
int relation_tc(const rn_relation_desc* d, const float* X, const float* boxes, const int* key_index, const float* Wq,
                const float* bq, const float* Wk, const float* bk, const float* Wg, const float* bg, const float* Wout,
                const float* bout, float* out, float* softmax_out, void* ws, size_t ws_bytes, cudaStream_t st);

struct GeomGather;
size_t relation_tc_packed_bytes(const rn_relation_desc* d);
int relation_tc_pack(const rn_relation_desc* d, const float* Wq, const float* bq, const float* Wk, const float* bk,
                     const float* Wout, const float* bout, void* packed, cudaStream_t st);
int relation_tc_packed(const rn_relation_desc* d, const float* X, const float* boxes, const int* key_index,
                       const void* packed, const float* Wg, const float* bg, float* out, void* ws, size_t ws_bytes,
                       cudaStream_t st, int stage_mask, const struct GeomGather* gg = nullptr, const void* x_f16 = nullptr,
                       void* out_f16 = nullptr);
// geometry gathered from one roi-level table lg_table [H, R, ld]: row i of problem b is roi idx[i*stride_i + b*stride_b]
struct GeomGather { const float* lg_table; int ld; int R; const int* idx; int stride_i; int stride_b;
                    const void* qkv_ext; const int* active = nullptr; };   // active: optional per-problem skip mask   // qkv_ext: optional pre-computed fp16 [B*N, 3*H*64] projections (skips the GEMM)
int relation_tc_gathered(const rn_relation_desc* d, const float* X, const GeomGather* gg, const float* Wq, const float* bq,
                         const float* Wk, const float* bk, const float* Wout, const float* bout, float* out, void* ws,
                         size_t ws_bytes, cudaStream_t st);
// learn-NMS: Q/K/V' of row (c, i) = (emb . W^T)[idx[i,c]] + (rank_feat . W^T + b)[i]  -- two small GEMMs + a gather-add
int relation_tc_lnms(const rn_relation_desc* d, const float* X, const float* emb, int R_emb, const float* rank_feat,
                     const GeomGather* gg, const float* Wq, const float* bq, const float* Wk, const float* bk,
                     const float* Wout, const float* bout, float* out, void* ws, size_t ws_bytes, cudaStream_t st,
                     const void* prepared = nullptr);
// weight-only half (packed fp16 weights + RQKV), built once per weight update and passed as `prepared`
size_t relation_tc_lnms_prepared_bytes(const rn_relation_desc* d);
int relation_tc_lnms_prepare(const rn_relation_desc* d, const float* rank_feat, const float* Wq, const float* bq,
                             const float* Wk, const float* bk, const float* Wout, const float* bout, void* prepared,
                             void* ws, size_t ws_bytes, cudaStream_t st);
size_t relation_tc_lnms_extra_bytes(const rn_relation_desc* d, int R_emb);
bool relation_tc_shape_ok(const rn_relation_desc* d);
int launch_geom_weight_log2_T(cudaStream_t st, const float* boxes, int N, int H, int E, float wave_length, const float* Wg,
                              const float* bg, float* g, int ldg);
int launch_geom_weight_log2(cudaStream_t st, const float* boxes, const int* key_index, int B, int N, int M, int H, int E,
                            float wave_length, const float* Wg, const float* bg, float* g, int ldg);

// tcgen05 fp16 GEMM (gemm_tc.cu): y = act(x W^T + b)
size_t linear_tc_workspace_bytes(int rows, int in, int out);
int linear_tc(const float* x, const float* W, const float* b, float* y, int rows, int in, int out, int relu, void* ws,
              size_t ws_bytes, cudaStream_t st);
size_t linear_tc_packed_bytes(int in, int out);
int linear_tc_pack_chw_to_hwc(const float* W, int out, int C, int S, void* packed, cudaStream_t st);
int linear_tc_packed_f16in(const void* x16, const void* packed_W, const float* b, float* y, void* y16, int rows, int in,
                           int out, int relu, void* ws, size_t ws_bytes, cudaStream_t st);
int linear_tc_pack(const float* W, int in, int out, void* packed, cudaStream_t st);
int linear_tc_packed(const float* x, const void* packed_W, const float* b, float* y, int rows, int in, int out, int relu,
                     void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace rn
