/* relnet_b200.h -- C ABI of the B200-native Relation-Networks hot path (librelnet_b200.so).
 *
 * One shared library, extern "C", plain pointers and sizes: no torch / C++ types cross this boundary.
 * Every entry point
 *   - takes raw DEVICE pointers (unless a parameter says "host"), explicit shapes and a cudaStream_t (passed as
 *     void*; NULL = legacy default stream);
 *   - returns 0 on success, non-zero on error, with a thread-local message in rn_last_error();
 *   - allocates nothing on the device: the caller passes a workspace of at least *_workspace_bytes(...);
 *   - never synchronises the host with the device and never copies to the host;
 *   - is safe to call concurrently from different threads on different streams.
 * (One exception to "allocates nothing": the first plain-GEMM call on a thread creates that thread's cuBLAS handle.)
 *
 * All tensors are dense row-major float32 unless stated; MXNet layouts are kept (FC weight = [out, in]).
 * Reference = msracver/Relation-Networks-for-Object-Detection; aliases as in SURVEY.md:
 *   SYM_REL = relation_rcnn/symbols/resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16.py
 *   LNMS    = relation_rcnn/operator_py/learn_nms.py
 */
#ifndef RELNET_B200_H_
#define RELNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_OK 0
#define RN_ERR_INVALID 1   /* bad argument / unsupported shape */
#define RN_ERR_WORKSPACE 2 /* workspace too small */
#define RN_ERR_CUDA 3      /* CUDA / cuBLAS runtime error */

typedef void* rn_stream_t; /* cudaStream_t */

const char* rn_last_error(void);
int rn_version(void);
/* 1 when the running device is sm_100 (tcgen05/TMA paths usable); fills *sm_count if non-NULL.
 * PROCESS MODEL: one device per process (the reference's one executor per GPU; torchrun's one rank per GPU).  The device
 * properties, the per-kernel launch attributes (> 48 KB dynamic shared memory) and the cuBLAS handle of the parity mode
 * are cached for the device that is current at the first call; entry points called later with another device current
 * return RN_ERR_INVALID ("librelnet_b200 is bound to device N") instead of launching with stale attributes. */
int rn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------------------------
 * Object-relation module.  Replaces extract_position_matrix (SYM_REL:47-83, FPN: SYM_FPN_REL_NMS:860-905),
 * extract_position_embedding (SYM_REL:30-44), attention_module_multi_head (SYM_REL:85-151, FPN :907-977) and the
 * caller's `fc_all = fc_new + attention; relu` (SYM_REL:267-268).  Also the per-class relation inside learn_nms
 * (LNMS:45-127) through `batch`.
 *
 *   g[n,h,m] = max(relu(Wg[h,:].phi(eps(box_n, box_m)) + bg[h]), 1e-6)
 *   p[n,h,:] = softmax_m( log g[n,h,m] + <q_h[n], k_h[m]>/sqrt(dk) )          q = X Wq^T + bq, k = X[keys] Wk^T + bk
 *   o[n, h*dv:(h+1)*dv] = sum_m p[n,h,m] (X[keys] Wout^T)[m, h*dv:(h+1)*dv] + bout
 *   out = fuse_residual_relu ? relu(X + o) : o
 */
enum { RN_PREC_FP32 = 0,  /* SIMT fp32 attention + cuBLAS fp32 projections: the bit-conservative parity mode */
       RN_PREC_F16 = 1,   /* fp16 operands / fp32 accumulate on tcgen05 tensor cores (sm_100a only) */
       RN_PREC_TF32 = 2 }; /* rn_relation_fwd only: the general (materialising) kernels of RN_PREC_FP32 with every GEMM on the
                            * library's tcgen05 tf32 engine (rn_gemm_tf32) -- exactly the forward that rn_relation_bwd /
                            * rn_learn_nms_bwd recompute under RN_PREC_F16; also serves return_softmax on tensor cores */

typedef struct rn_relation_desc {
  int32_t batch;      /* independent problems sharing the weights (1 for the detection head, #classes for learn-NMS) */
  int32_t N;          /* queries per problem */
  int32_t M;          /* keys per problem: rows key_index[0..M) of X, or the first M rows when key_index == NULL */
  int32_t d;          /* feature dim of X */
  int32_t dq;         /* total query/key dim = H * dk          (1024) */
  int32_t dout;       /* total output dim    = H * dv          (1024; 128 in learn-NMS) */
  int32_t H;          /* heads ("group" / fc_dim in the reference, 16) */
  int32_t E;          /* position-embedding dim (64) */
  float wave_length;  /* 1000 */
  int32_t fuse_residual_relu; /* 1: out = relu(X + o) (needs dout == d) */
  int32_t precision;  /* RN_PREC_* */
} rn_relation_desc;

size_t rn_relation_workspace_bytes(const rn_relation_desc* desc);

/* X [batch, N, d]; boxes [batch, N, 4] (x1,y1,x2,y2); key_index int32 [M] or NULL (shared by the batch);
 * Wq,Wk [dq,d]; bq,bk [dq]; Wg [H,E]; bg [H]; Wout [dout,d] (grouped 1x1 conv weight [dout,d,1,1]); bout [dout];
 * out [batch, N, dout].  softmax_out (optional, may be NULL) [batch, N, H, M] = p, as the reference's 2nd output of
 * attention_module_nms_multi_head (SYM_REL_NMS:158-238). */
int rn_relation_fwd(const rn_relation_desc* desc, const float* X, const float* boxes, const int32_t* key_index,
                    const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wg,
                    const float* bg, const float* Wout, const float* bout, float* out, float* softmax_out,
                    void* workspace, size_t workspace_bytes, rn_stream_t stream);

/* Standalone a1+a2 for tests: eps_out [N,M,4] and/or emb_out [N,M,E] (either may be NULL). */
int rn_pos_embed_fwd(const float* boxes, const int32_t* key_index, int32_t N, int32_t M, int32_t E, float wave_length,
                     float* eps_out, float* emb_out, rn_stream_t stream);

/* Geometry weight only: g_out [batch, H, N, M] = max(relu(Wg.phi + bg), 1e-6)  (SYM_REL:107-116 + the clamp of :139) */
int rn_geometry_weight_fwd(const float* boxes, const int32_t* key_index, int32_t batch, int32_t N, int32_t M, int32_t H,
                           int32_t E, float wave_length, const float* Wg, const float* bg, float* g_out,
                           rn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Plain dense layer y = act(x W^T + b) used either side of the relation module (fc_new_1/2, cls_score, bbox_pred:
 * SYM_REL:261-280).  x [rows, in], W [out, in], b [out] or NULL, y [rows, out]. relu: 0/1.  precision as above
 * (RN_PREC_FP32 = cuBLAS SGEMM; RN_PREC_F16 = tcgen05 fp16 GEMM with fp32 accumulate). */
size_t rn_linear_workspace_bytes(int32_t rows, int32_t in, int32_t out, int32_t precision);
int rn_linear_fwd(const float* x, const float* W, const float* b, float* y, int32_t rows, int32_t in, int32_t out,
                  int32_t relu, int32_t precision, void* workspace, size_t workspace_bytes, rn_stream_t stream);

/* Pre-packed fp16 operands for RN_PREC_F16: pack once per weight update (one small kernel), reuse every forward.
 * The packed relation block is [Wq; Wk; Wout padded to 64 cols/head] as fp16 [3*H*64, ceil8(d)] followed by the fp32
 * bias [bq; bk; bout padded]; the packed linear weight is fp16 [out, ceil8(in)]. */
/* RN_PREC_F16 has two decompositions of the N x M part, both tcgen05: the FUSED one (default: pair geometry + pair FC +
 * QK^T + softmax + P.V' in one cooperative launch, relation_fused.cu) and the round-1 one (geometry table [B,H,N,M] in
 * HBM -> tile attention -> combine).  rn_relation_fused_enable(0) selects the latter for A/B measurements (process-wide;
 * also RN_RELATION_UNFUSED=1 in the environment); 1 = fused with the pair embedding phi rounded to fp16 before the pair FC
 * (default), 2 = fused with phi's fp16 residual fed to the FC as well (phi at ~fp32 accuracy; RN_FUSED_PHI_LO=1).
 * Returns the previous setting. */
int rn_relation_fused_enable(int32_t on);
size_t rn_relation_packed_bytes(const rn_relation_desc* desc);
int rn_relation_pack(const rn_relation_desc* desc, const float* Wq, const float* bq, const float* Wk, const float* bk,
                     const float* Wout, const float* bout, void* packed, rn_stream_t stream);
int rn_relation_packed_fwd(const rn_relation_desc* desc, const float* X, const float* boxes, const int32_t* key_index,
                           const void* packed, const float* Wg, const float* bg, float* out, void* workspace,
                           size_t workspace_bytes, rn_stream_t stream);
/* measurement hook: run only the stages in stage_mask (1 = cast + projection GEMM, 2 = geometry kernel, 4 = fused attention
 * kernel) on the intermediates a previous full call left in the same workspace -- used by bench.py to time one kernel
 * with CUDA events */
int rn_relation_packed_stages(const rn_relation_desc* desc, const float* X, const float* boxes, const int32_t* key_index,
                              const void* packed, const float* Wg, const float* bg, float* out, void* workspace,
                              size_t workspace_bytes, int32_t stage_mask, rn_stream_t stream);
/* Backward of the relation module (training: the reference differentiates the symbol graph of
 * attention_module_multi_head, resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16.py:104-151, through MXNet
 * autograd; there is no explicit backward source to cite).  fp32; the forward intermediates are recomputed, nothing has to
 * be saved from rn_relation_fwd.  dOut [batch*N, dout] is the gradient of the loss w.r.t. rn_relation_fwd's `out`; every
 * gradient buffer is OVERWRITTEN (not accumulated) and has the shape of its parameter: dX [batch*N, d], dWq/dWk [dq, d],
 * dbq/dbk [dq], dWg [H, E], dbg [H], dWout [dout, d], dbout [dout].  Boxes receive no gradient (proposal.py:170-173). */
/* desc->precision selects the contraction engine of the WHOLE backward (the recomputed forward included):
 * RN_PREC_F16 = the library's own tcgen05 GEMM on the fp32 operands as they lie in HBM (kind::tf32, fp32 accumulate in
 * TMEM: rn_gemm_tf32 below; sm_100 only); RN_PREC_FP32 = cuBLAS fp32, the bit-conservative parity mode. */
size_t rn_relation_bwd_workspace_bytes(const rn_relation_desc* desc);
int rn_relation_bwd(const rn_relation_desc* desc, const float* X, const float* boxes, const int32_t* key_index,
                    const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wg, const float* bg,
                    const float* Wout, const float* bout, const float* dOut, float* dX, float* dWq, float* dbq, float* dWk,
                    float* dbk, float* dWg, float* dbg, float* dWout, float* dbout, void* workspace,
                    size_t workspace_bytes, rn_stream_t stream);
/* Same with the output of the forward that was actually executed (out_fwd [batch*N, dout], any precision of rn_relation_*_fwd):
 * under fuse_residual_relu the relu mask is then [out_fwd > 0] -- what autograd over the executed graph does -- instead of
 * the mask of the recomputed forward.  The two differ only for units within rounding of the kink (|X + o| ~ 1e-3 max under
 * the tf32 engine, ~1e-6 under fp32), but each such unit switches its whole dOut entry on or off. */
int rn_relation_bwd_masked(const rn_relation_desc* desc, const float* X, const float* boxes, const int32_t* key_index,
                           const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wg, const float* bg,
                           const float* Wout, const float* bout, const float* out_fwd, const float* dOut, float* dX, float* dWq,
                           float* dbq, float* dWk, float* dbk, float* dWg, float* dbg, float* dWout, float* dbout,
                           void* workspace, size_t workspace_bytes, rn_stream_t stream);
/* The contraction engine of the training side, exported for checking on its own: row-major fp32
 *   C[o,i] (M x N, pitch ldc) = alpha * op(A[o,i]) . op(B[o,i]) + beta * C[o,i],   o < outer, i < inner,
 * problem (o, i) of an operand at base + o * s?o + i * s?i (floats).  transA: A is stored [K, M] (else [M, K]); transB: B is
 * stored [N, K] (else [K, N]).  Operands are read by TMA exactly as they lie (no cast, no packed copy), multiplied on
 * tcgen05 tensor cores as tf32 (10-bit mantissa operands, fp32 accumulate).  Requirements: sm_100; 16-byte aligned bases;
 * pitches and batch strides multiples of 4 floats; outer * inner <= 65535.  Replaces cuBLAS in rn_relation_bwd /
 * rn_learn_nms_bwd under RN_PREC_F16 (the reference's backward is MXNet autograd over cuBLAS: no source to cite). */
int rn_gemm_tf32(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K, float alpha, const float* A, int32_t lda,
                 int64_t sAo, int64_t sAi, const float* B, int32_t ldb, int64_t sBo, int64_t sBi, float beta, float* C,
                 int32_t ldc, int64_t sCo, int64_t sCi, int32_t outer, int32_t inner, rn_stream_t stream);
/* rn_relation_packed_stages with fp16 side channels: X_f16 (NULL or the producer's fp16 copy of X, [batch*N, d], d % 8 == 0)
 * replaces the module's own cast; out_f16 (NULL or [batch*N, dout] fp16) receives a copy of `out` for the consumer GEMM
 * (rn_linear_packed_f16in_fwd).  Same arithmetic as rn_relation_packed_fwd. */
int rn_relation_packed_fwd_f16io(const rn_relation_desc* desc, const float* X, const void* X_f16, const float* boxes,
                                 const int32_t* key_index, const void* packed, const float* Wg, const float* bg, float* out,
                                 void* out_f16, void* workspace, size_t workspace_bytes, int32_t stage_mask,
                                 rn_stream_t stream);
size_t rn_linear_packed_bytes(int32_t in, int32_t out);
int rn_linear_pack(const float* W, int32_t in, int32_t out, void* packed, rn_stream_t stream);
int rn_linear_packed_fwd(const float* x, const void* packed_W, const float* b, float* y, int32_t rows, int32_t in,
                         int32_t out, int32_t relu, void* workspace, size_t workspace_bytes, rn_stream_t stream);

/* Several FullyConnected layers over the SAME input as one GEMM (cls_score + bbox_pred + roi_feat_embedding all read
 * fc_all_2_relu: SYM_REL_NMS:360-365,469-472).  nout <= 4; W[i] is [outs[i], in], b[i] [outs[i]] or NULL; ys[i] receives
 * the contiguous [rows, outs[i]] output of layer i.  x_f16 is the producer's fp16 copy of the input ([rows, in], in % 8 == 0). */
size_t rn_linear_multi_packed_bytes(const int32_t* outs, int32_t nout, int32_t in);
int rn_linear_multi_pack(const float* const* W, const float* const* b, const int32_t* outs, int32_t nout, int32_t in,
                         void* packed, rn_stream_t stream);
int rn_linear_multi_packed_f16in_fwd(const void* x_f16, const void* packed, float* const* ys, const int32_t* outs, int32_t nout,
                                     int32_t rows, int32_t in, void* workspace, size_t workspace_bytes, rn_stream_t stream);

/* Fused-head fast path (same arithmetic, fewer bytes): ROI max-pool straight from the trunk's channels-last feature map
 * to an fp16 [R, PH*PW, C] tensor, consumed by fc_new_1 through a K-permuted packed weight (no fp32 pooled tensor, no
 * cast kernel).  data_nhwc [B,H,W,C] fp32; out_f16 [R, PH*PW*C] fp16. */
int rn_roi_pool_nhwc_f16_fwd(const float* data_nhwc, const float* rois, int32_t R, int32_t C, int32_t H, int32_t W,
                             int32_t PH, int32_t PW, float spatial_scale, void* out_f16, rn_stream_t stream);
/* W [out, C*S] (FC over a flattened [C, S] input, S = PH*PW) -> packed fp16 [out, S*C] */
/* same with a bf16 channels-last feature map (what the cuDNN trunk emits): no fp32 conversion pass; identical output values */
int rn_roi_pool_nhwc_bf16in_f16_fwd(const void* data_nhwc_bf16, const float* rois, int32_t R, int32_t C, int32_t H, int32_t W,
                                    int32_t PH, int32_t PW, float spatial_scale, void* out_f16, rn_stream_t stream);
int rn_linear_pack_chw_to_hwc(const float* W, int32_t out, int32_t C, int32_t S, void* packed, rn_stream_t stream);
/* y = act(x W^T + b) with x already fp16 [rows, in] (in % 8 == 0); y (fp32) and/or y_f16 may be NULL */
int rn_linear_packed_f16in_fwd(const void* x_f16, const void* packed_W, const float* b, float* y, void* y_f16,
                               int32_t rows, int32_t in, int32_t out, int32_t relu, void* workspace,
                               size_t workspace_bytes, rn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * learn_nms CustomOp forward (LNMS:238-401) + test-time merge (SYM_REL_NMS:553-560).
 * Inputs in the order of LearnNmsProp.list_arguments (LNMS:429-441). */
typedef struct rn_learn_nms_desc {
  int32_t R;              /* rows of cls_score / bbox_pred / rois / feat */
  int32_t num_classes;    /* incl. background (81) */
  int32_t num_reg_classes;/* 2 when class agnostic */
  int32_t feat_dim;       /* 1024 */
  int32_t first_n;        /* 100 */
  int32_t num_thresh;     /* 5 */
  double class_thresh;    /* TEST.LEARN_NMS_CLASS_SCORE_TH, 0.01 (python float in the reference) */
  int32_t class_agnostic; /* 1 */
  int32_t has_means_stds; /* 0 at test time (already folded into bbox_pred weights, SYM_REL_NMS:420-421) */
  float means[4], stds[4];
  int32_t nongt_dim;      /* >0: use the first nongt_dim rows; 0: use non_gt_index if given, else all rows */
  int32_t num_non_gt;     /* length of non_gt_index (when nongt_dim == 0 and non_gt_index != NULL) */
  int32_t merge_method;   /* -1 mean, -2 max, k>=0 pick threshold k   (config.TEST.MERGE_METHOD) */
  int32_t precision;      /* RN_PREC_* for the projections / relation inside */
} rn_learn_nms_desc;

typedef struct rn_learn_nms_weights {
  const float *nms_rank_weight, *nms_rank_bias;                 /* [128,1024],[128] */
  const float *roi_feat_embedding_weight, *roi_feat_embedding_bias; /* [128,feat_dim],[128] */
  const float *nms_pair_pos_fc1_1_weight, *nms_pair_pos_fc1_1_bias; /* [16,64],[16] */
  const float *nms_query_1_weight, *nms_query_1_bias;           /* [1024,128],[1024] */
  const float *nms_key_1_weight, *nms_key_1_bias;               /* [1024,128],[1024] */
  const float *nms_linear_out_1_weight, *nms_linear_out_1_bias; /* [128,128,1,1],[128] */
  const float *nms_logit_weight, *nms_logit_bias;               /* [T,128],[T] */
} rn_learn_nms_weights;

size_t rn_learn_nms_workspace_bytes(const rn_learn_nms_desc* desc);
/* cls_score [R,num_classes], bbox_pred [R,4*num_reg_classes], rois [R,5], im_info [3] (device), feat [R,feat_dim],
 * non_gt_index int32 [num_non_gt] or NULL.
 * Outputs: nms_multi_score [n,C,T], sorted_bbox [n,C,4], sorted_score [n,C], final_score [n,C] (may be NULL). */
int rn_learn_nms_fwd(const rn_learn_nms_desc* desc, const float* cls_score, const float* bbox_pred, const float* rois,
                     const float* im_info, const float* feat, const rn_learn_nms_weights* w,
                     const int32_t* non_gt_index, float* nms_multi_score, float* sorted_bbox, float* sorted_score,
                     float* final_score, void* workspace, size_t workspace_bytes, rn_stream_t stream);

/* Weight-only half of the head, done once per weight update (RN_PREC_F16 + class-agnostic boxes only; packed_bytes returns
 * 0 otherwise): rank embedding -> nms_rank FC (LNMS:328-331), fp16 copies of the roi_feat_embedding / query / key /
 * linear_out weights and the rank half of the factored projection.  rn_learn_nms_packed_fwd == rn_learn_nms_fwd minus that
 * work (`w` is still needed for the biases, pair_pos_fc1 and nms_logit); workspace as for rn_learn_nms_fwd (also for pack). */
size_t rn_learn_nms_packed_bytes(const rn_learn_nms_desc* desc);
int rn_learn_nms_pack(const rn_learn_nms_desc* desc, const rn_learn_nms_weights* w, void* packed, void* workspace,
                      size_t workspace_bytes, rn_stream_t stream);
int rn_learn_nms_packed_fwd(const rn_learn_nms_desc* desc, const float* cls_score, const float* bbox_pred, const float* rois,
                            const float* im_info, const float* feat, const void* feat_f16 /* fp16 copy of feat or NULL */,
                            const float* emb /* roi_feat_embedding(feat) [R,128] if the caller already has it, else NULL */,
                            const rn_learn_nms_weights* w, const void* packed,
                            const int32_t* non_gt_index, float* nms_multi_score, float* sorted_bbox, float* sorted_score,
                            float* final_score, void* workspace, size_t workspace_bytes, rn_stream_t stream);

/* ---- training side of the learn-NMS head ------------------------------------------------------------------------
 * Gradient buffers, one per entry of rn_learn_nms_weights (same shapes); all are OVERWRITTEN by rn_learn_nms_bwd. */
typedef struct rn_learn_nms_grads {
  float *nms_rank_weight, *nms_rank_bias;
  float *roi_feat_embedding_weight, *roi_feat_embedding_bias;
  float *nms_pair_pos_fc1_1_weight, *nms_pair_pos_fc1_1_bias;
  float *nms_query_1_weight, *nms_query_1_bias;
  float *nms_key_1_weight, *nms_key_1_bias;
  float *nms_linear_out_1_weight, *nms_linear_out_1_bias;
  float *nms_logit_weight, *nms_logit_bias;
} rn_learn_nms_grads;

/* Backward of the learn-NMS head (the reference differentiates the train graph, ..._multi_head_16_learn_nms.py:424-501,
 * with MXNet autograd; no backward source exists).  Inputs as rn_learn_nms_fwd (use class_thresh = 0 for the train graph,
 * which prunes no class); d_multi [n,C,T] = d loss / d nms_multi_score.  Outputs: the 14 weight gradients, d_cls_score
 * [R,num_classes] (through the sorted softmax scores) and d_feat [R,feat_dim] (through roi_feat_embedding); bbox_pred /
 * rois get none (BlockGrad, :428).  fp32; the forward is recomputed inside, nothing has to be saved. */
size_t rn_learn_nms_bwd_workspace_bytes(const rn_learn_nms_desc* desc);
int rn_learn_nms_bwd(const rn_learn_nms_desc* desc, const float* cls_score, const float* bbox_pred, const float* rois,
                     const float* im_info, const float* feat, const rn_learn_nms_weights* w, const int32_t* non_gt_index,
                     const float* d_multi, const rn_learn_nms_grads* grads, float* d_cls_score, float* d_feat,
                     void* workspace, size_t workspace_bytes, rn_stream_t stream);

/* learn-NMS loss (..._learn_nms.py:539-551): pos = -target*log(score+eps)*loss_scale/(first_n*num_thresh), neg likewise
 * with (1-target), (1-score); d_multi = pos_grad_scale * d pos + d neg  (MakeLoss sends grad_scale back; TRAIN.nms_pos_scale
 * = 4, nms_loss_scale = 1, eps = 1e-8: config.py:73-74).  Any of pos_loss / neg_loss / d_multi [n,C,T] may be NULL. */
int rn_nms_loss(const float* nms_multi_score, const float* nms_multi_target, int32_t first_n, int32_t C, int32_t num_thresh,
                float loss_scale, float pos_grad_scale, float eps, float* pos_loss, float* neg_loss, float* d_multi,
                rn_stream_t stream);

/* `BoxAnnotatorOHEM` CustomOp forward (relation_rcnn/operator_py/box_annotator_ohem.py:26-53): keep the roi_per_img rois
 * of largest (softmax cross-entropy + weighted smooth-L1) loss; the others get label -1 and zero box weights.
 * cls_score [R,num_classes], bbox_pred/bbox_targets/bbox_weights [R,4*num_reg_classes], labels [R] (float-valued).
 * per_roi_loss [R] is written too.  Ties: larger roi index ranks first. */
int rn_box_annotator_ohem(const float* cls_score, const float* bbox_pred, const float* labels, const float* bbox_targets,
                          const float* bbox_weights, int32_t R, int32_t num_classes, int32_t num_reg_classes,
                          int32_t roi_per_img, float* labels_ohem, float* bbox_weights_ohem, float* per_roi_loss,
                          rn_stream_t stream);

/* `nms_multi_target` CustomOp forward (relation_rcnn/operator_py/nms_multi_target.py:24-74): learn-NMS training labels.
 * bbox [n,C,4], gt_boxes [G,5] (x1,y1,x2,y2,cls), score [n,C], target_thresh HOST double[T] -> out [n,C,T] (0/1). */
int rn_nms_multi_target_fwd(const float* bbox, const float* gt_boxes, const float* score, int32_t n, int32_t C, int32_t G,
                            const double* target_thresh_host, int32_t T, float* out, rn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * `proposal` CustomOp forward (relation_rcnn/operator_py/proposal.py:51-168), device resident end to end.
 * cls_prob [1,2A,Hf,Wf], bbox_pred [1,4A,Hf,Wf], im_info [3] (device).  scales/ratios are HOST arrays.
 * rois_out [post,5], scores_out [post,1] or NULL.  num_kept_out (device int32[1], may be NULL) = boxes that survived
 * NMS before padding.  Tie/padding rules: DESIGN.md "proposal". */
typedef struct rn_proposal_desc {
  int32_t Hf, Wf;           /* feature-map size of the inputs */
  int32_t feat_stride;      /* 16 */
  int32_t num_scales, num_ratios;
  int32_t pre_nms_top_n;    /* 6000 */
  int32_t post_nms_top_n;   /* 300 */
  float nms_thresh;         /* 0.7 */
  float min_size;           /* 0 */
} rn_proposal_desc;
size_t rn_proposal_workspace_bytes(const rn_proposal_desc* desc);
int rn_proposal_fwd(const rn_proposal_desc* desc, const float* scales_host, const float* ratios_host,
                    const float* cls_prob, const float* bbox_pred, const float* im_info, float* rois_out,
                    float* scores_out, int32_t* num_kept_out, void* workspace, size_t workspace_bytes,
                    rn_stream_t stream);

/* Replaces `_nms` (lib/nms/gpu_nms.hpp:1, lib/nms/nms_kernel.cu:91-144) with device in/out and no host sweep.
 * boxes [n, box_dim>=4] sorted by score; keep_out int32 [max_keep]; num_out int32[1].  IoU in float32, '>' thresh. */
size_t rn_nms_workspace_bytes(int32_t n);
int rn_nms(const float* boxes_sorted, int32_t n, int32_t box_dim, float thresh, int32_t max_keep, int32_t* keep_out,
           int32_t* num_out, void* workspace, size_t workspace_bytes, rn_stream_t stream);

/* Replaces bbox_overlaps_cython (lib/bbox/bbox.pyx:15-55): float64, +1 areas.  boxes [N,4] f64, query [K,4] f64. */
int rn_bbox_overlaps(const double* boxes, const double* query, int32_t N, int32_t K, double* out, rn_stream_t stream);

/* `proposal_target` CustomOp forward, BATCH_ROIS == -1 path (proposal_target.py:44-93 -> core/rcnn.py:288-325).
 * rois [N,5], gt_boxes [G,5] -> rois_out [N+G,5], label [N+G], bbox_target [N+G,4R], bbox_weight [N+G,4R]. */
typedef struct rn_proposal_target_desc {
  int32_t N, G;
  int32_t num_reg_classes;  /* R = 2 when class agnostic */
  int32_t class_agnostic;
  float bg_thresh_hi;       /* TRAIN.BG_THRESH_HI 0.5 */
  int32_t normalize;        /* TRAIN.BBOX_NORMALIZATION_PRECOMPUTED */
  double means[4], stds[4];
  float bbox_weights[4];
} rn_proposal_target_desc;
int rn_proposal_target_fwd(const rn_proposal_target_desc* desc, const float* rois, const float* gt_boxes,
                           float* rois_out, float* label, float* bbox_target, float* bbox_weight, rn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * ROIPooling (max; MXNet built-in used at SYM_REL:252-253).  data [B,C,H,W], rois [R,5] -> out [R,C,PH,PW],
 * argmax int32 same shape (may be NULL).  rois[:,0] is the image index: a NEGATIVE index marks a padding roi (output 0,
 * argmax -1, as MXNet's kernel does); indices >= B are the caller's error -- the forward entry points carry no B and
 * cannot check them (rn_roi_pool_bwd, which knows B, skips such rows). */
int rn_roi_pool_fwd(const float* data, const float* rois, int32_t R, int32_t C, int32_t H, int32_t W, int32_t PH,
                    int32_t PW, float spatial_scale, float* out, int32_t* argmax, rn_stream_t stream);

/* ROIPooling backward (MXNet 1.1.0 ROIPoolBackwardAcc): ddata [B,C,H,W] is OVERWRITTEN with the gradient routed to the
 * argmax elements recorded by rn_roi_pool_fwd. */
int rn_roi_pool_bwd(const float* dout, const int32_t* argmax, const float* rois, int32_t R, int32_t B, int32_t C, int32_t H,
                    int32_t W, int32_t PH, int32_t PW, float* ddata, rn_stream_t stream);

/* DeformablePSROIPooling forward (operator_cxx/deformable_psroi_pooling.cu:52-138; = average ROIAlign when no_trans,
 * group_size 1).  trans [R, 2*num_classes, part, part] or NULL when no_trans.  top_count may be NULL. */
typedef struct rn_psroi_desc {
  int32_t R, channels, H, W;
  float spatial_scale;
  int32_t output_dim, group_size, pooled_size, part_size, sample_per_part;
  float trans_std;
  int32_t no_trans, num_classes;
} rn_psroi_desc;
int rn_deform_psroi_pool_fwd(const rn_psroi_desc* desc, const float* data, const float* rois, const float* trans,
                             float* out, float* top_count, rn_stream_t stream);
/* Same operator on a CHANNELS-LAST feature map [B, H, W, channels] (fp32, or bf16 when data_is_bf16: the layout the trunk's
 * conv_new_1 leaves it in) -- the fast form: every bilinear tap is a 16-byte channel-vector load.  Same arithmetic per
 * channel (a bf16 map is widened exactly); out / top_count keep the reference layout [R, output_dim, pooled, pooled]. */
int rn_deform_psroi_pool_nhwc_fwd(const rn_psroi_desc* desc, const void* data_nhwc, int32_t data_is_bf16, const float* rois,
                                  const float* trans, float* out, float* top_count, rn_stream_t stream);

/* DeformablePSROIPooling backward (operator_cxx/deformable_psroi_pooling.cu:177-289).  B = batch size of data; dout and
 * top_count as produced by the forward; ddata [B,channels,H,W] and dtrans (shape of trans; NULL when no_trans) are
 * OVERWRITTEN.  Accumulation uses atomics: equal to the reference up to float summation order. */
int rn_deform_psroi_pool_bwd(const rn_psroi_desc* desc, int32_t B, const float* dout, const float* top_count,
                             const float* data, const float* rois, const float* trans, float* ddata, float* dtrans,
                             rn_stream_t stream);

/* DeformableConvolution forward (operator_cxx/deformable_convolution-inl.h:91-144 + nn/deformable_im2col.cuh:216-309).
 * data [B,C,H,W], offset [B, dg*2*kh*kw, Ho, Wo], weight [Co, C/groups, kh, kw], bias [Co] or NULL -> out [B,Co,Ho,Wo] */
typedef struct rn_deform_conv_desc {
  int32_t B, C, H, W, Co;
  int32_t kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w;
  int32_t num_group, num_deformable_group;
  int32_t precision;
} rn_deform_conv_desc;
size_t rn_deform_conv_workspace_bytes(const rn_deform_conv_desc* desc);
int rn_deform_conv_fwd(const rn_deform_conv_desc* desc, const float* data, const float* offset, const float* weight,
                       const float* bias, float* out, void* workspace, size_t workspace_bytes, rn_stream_t stream);
/* Channels-last fast form of the same forward (one image, num_group 1): data NHWC fp32 or bf16 (the trunk's layout), the
 * weight packed once by rn_deform_conv_pack (fp16, K ordered tap-major to match the sampler), output NHWC as fp32 and / or
 * fp16 with bias and optional relu fused in the tcgen05 GEMM epilogue.  workspace: rn_deform_conv_workspace_bytes. */
size_t rn_deform_conv_packed_bytes(const rn_deform_conv_desc* desc);
int rn_deform_conv_pack(const rn_deform_conv_desc* desc, const float* weight, void* packed, rn_stream_t stream);
int rn_deform_conv_nhwc_fwd(const rn_deform_conv_desc* desc, const void* data_nhwc, int32_t data_is_bf16, const float* offset,
                            const void* packed_weight, const float* bias, int32_t relu, float* out_f32, void* out_f16,
                            void* workspace, size_t workspace_bytes, rn_stream_t stream);
/* DeformableConvolution backward (operator_cxx/deformable_convolution-inl.h:145-233; col2im / col2im_coord kernels
 * nn/deformable_im2col.cuh:315-458).  ddata, doffset, dweight, dbias (NULL when no bias) are OVERWRITTEN; workspace as for
 * the forward (rn_deform_conv_workspace_bytes).  weight_grad_deformed = 0 is the REFERENCE: its dWeight is computed from
 * the plain (un-deformed) im2col of data (:215); 1 uses the deformed sampling (the gradient later MXNet releases compute). */
int rn_deform_conv_bwd(const rn_deform_conv_desc* desc, const float* dout, const float* data, const float* offset,
                       const float* weight, int32_t weight_grad_deformed, float* ddata, float* doffset, float* dweight,
                       float* dbias, void* workspace, size_t workspace_bytes, rn_stream_t stream);
/* the im2col stage alone (tests): col [C*kh*kw, Ho, Wo] for image b */
int rn_deform_im2col(const rn_deform_conv_desc* desc, const float* data_b, const float* offset_b, float* col,
                     rn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Stem helpers around the (library) trunk: both are plain HBM-bound data movement.
 * rn_image_s2d_bf16: fp32 [3,H,W] image -> bf16 channels-last [(H+2pad)/2, (W+2pad)/2, 16]: space-to-depth(2) of the zero-
 *   padded image, channel = c*4 + (row parity)*2 + (col parity), channels 12..15 zero.  With pad = 3 the 7x7 stride-2 conv1
 *   (relation_rcnn/symbols/resnet_v1_101_rcnn_base.py conv1) is a 4x4 stride-1 convolution over this tensor.
 * rn_maxpool3x3s2_nhwc_bf16: pool1 (3x3, stride 2, pooling_convention='full' = ceil mode) on a channels-last bf16 map;
 *   out is [ceil((H-3)/2)+1, ceil((W-3)/2)+1, C], C % 8 == 0. */
/* rn_rpn_head_fwd: the RPN head after rpn_conv (resnet_v1_101_rcnn_base.py:685-693): rpn_cls_score (2A) and rpn_bbox_pred
 *   (4A) 1x1 convolutions over the channels-last bf16 map r [HW, Cin] (Cin % 64 == 0, A <= 16), the {bg, fg} softmax per anchor
 *   (channel c pairs with c +- A), outputs fp32 NCHW: prob [2A, HW], bbox [4A, HW] -- what rn_proposal_fwd reads.
 *   Weights / biases bf16 in the layout of the conv parameters ([out, Cin]). */
int rn_rpn_head_fwd(const void* r_nhwc_bf16, int32_t HW, int32_t Cin, int32_t A, const void* Wcls_bf16, const void* bcls_bf16,
                    const void* Wbbox_bf16, const void* bbbox_bf16, float* prob, float* bbox, rn_stream_t stream);
/* Tensor-core form of rn_rpn_head_fwd (sm_100): the two 1x1 heads are ONE [HW, Cin] x [Cin, 6A] GEMM on the library's tcgen05
 * kernel with the bf16 operands exactly as the trunk left them, followed by one small softmax / layout kernel.  Pack the
 * weights once per weight update (rn_rpn_head_pack: [Wcls; Wbbox] concatenated + fp32 biases); outputs as rn_rpn_head_fwd. */
size_t rn_rpn_head_packed_bytes(int32_t Cin, int32_t A);
int rn_rpn_head_pack(const void* Wcls_bf16, const void* bcls_bf16, const void* Wbbox_bf16, const void* bbbox_bf16, int32_t Cin,
                     int32_t A, void* packed, rn_stream_t stream);
size_t rn_rpn_head_workspace_bytes(int32_t HW, int32_t Cin, int32_t A);
int rn_rpn_head_packed_fwd(const void* r_nhwc_bf16, int32_t HW, int32_t Cin, int32_t A, const void* packed, float* prob,
                           float* bbox, void* workspace, size_t workspace_bytes, rn_stream_t stream);
int rn_image_s2d_bf16(const float* image_chw, int32_t H, int32_t W, int32_t pad, void* out_nhwc16_bf16, rn_stream_t stream);
int rn_maxpool3x3s2_nhwc_bf16(const void* in_nhwc, int32_t H, int32_t W, int32_t C, void* out_nhwc, rn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * tcgen05 self-test (sm_100a only): runs one 128x128x64 K-major and one 128x64x128 MN-major-B UMMA through TMA/TMEM and
 * writes the fp32 results to out_s [128,128], out_o [128,64] for checking against a host product. */
int rn_umma_selftest(const void* a_f16, const void* b_f16, const void* p_f16, const void* v_f16, float* out_s,
                     float* out_o, rn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RELNET_B200_H_ */
