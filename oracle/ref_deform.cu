// C entry points around the REFERENCE's own deformable CUDA code, compiled from where it lies under /root/reference
// (never copied): relation_rcnn/operator_cxx/nn/deformable_im2col.cuh (deformable_im2col / col2im / col2im_coord) and
// relation_rcnn/operator_cxx/deformable_psroi_pooling.cu (DeformablePSROIPoolForward / BackwardAcc).
// TEST INFRASTRUCTURE: built by oracle/Makefile into oracle/_ref/libref_deform.so, loaded only by tests/ to pin
// oracle/oracle_c.c and the product kernels against the reference's real kernels on the GPU box.
#include "ref_stub/stub_all.h"
// the operator-class header of the PS-ROI op needs the MXNet tree; the kernels and their launch wrappers do not
#define MXNET_OPERATOR_DEFORMABLE_PSROI_POOLING_INL_H_
#include "relation_rcnn/operator_cxx/nn/deformable_im2col.cuh"
#undef CUDA_KERNEL_LOOP
using mxnet::op::Tensor;
using mxnet::op::Stream;
using mshadow::cuda::kBaseThreadNum;
using mxnet::Operator;
using mxnet::op::DeformablePSROIPoolingParam;
#include "relation_rcnn/operator_cxx/deformable_psroi_pooling.cu"

using mxnet::TShape;

static TShape im_shape(int C, int H, int W) { return TShape{1u, (index_t)C, (index_t)H, (index_t)W}; }

extern "C" int ref_deformable_im2col(const float* data_im, const float* data_offset, int C, int H, int W, int kh, int kw,
                                     int pad_h, int pad_w, int stride_h, int stride_w, int dil_h, int dil_w,
                                     int deformable_group, float* data_col) {
  const int Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  mshadow::Stream<gpu> s;
  mxnet::op::deformable_im2col<float>(&s, data_im, data_offset, im_shape(C, H, W),
                                      TShape{(index_t)(C * kh * kw), (index_t)Ho, (index_t)Wo}, TShape{(index_t)kh, (index_t)kw},
                                      TShape{(index_t)pad_h, (index_t)pad_w}, TShape{(index_t)stride_h, (index_t)stride_w},
                                      TShape{(index_t)dil_h, (index_t)dil_w}, (uint32_t)deformable_group, data_col);
  return (int)cudaDeviceSynchronize();
}

extern "C" int ref_deformable_col2im(const float* data_col, const float* data_offset, int C, int H, int W, int kh, int kw,
                                     int pad_h, int pad_w, int stride_h, int stride_w, int dil_h, int dil_w,
                                     int deformable_group, float* grad_im) {
  const int Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  mshadow::Stream<gpu> s;
  mxnet::op::deformable_col2im<float>(&s, data_col, data_offset, im_shape(C, H, W),
                                      TShape{(index_t)(C * kh * kw), (index_t)Ho, (index_t)Wo}, TShape{(index_t)kh, (index_t)kw},
                                      TShape{(index_t)pad_h, (index_t)pad_w}, TShape{(index_t)stride_h, (index_t)stride_w},
                                      TShape{(index_t)dil_h, (index_t)dil_w}, (uint32_t)deformable_group, grad_im, mxnet::kWriteTo);
  return (int)cudaDeviceSynchronize();
}

extern "C" int ref_deformable_col2im_coord(const float* data_col, const float* data_im, const float* data_offset, int C,
                                           int H, int W, int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w,
                                           int dil_h, int dil_w, int deformable_group, float* grad_offset) {
  const int Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  mshadow::Stream<gpu> s;
  mxnet::op::deformable_col2im_coord<float>(&s, data_col, data_im, data_offset, im_shape(C, H, W),
                                            TShape{(index_t)(C * kh * kw), (index_t)Ho, (index_t)Wo},
                                            TShape{(index_t)kh, (index_t)kw}, TShape{(index_t)pad_h, (index_t)pad_w},
                                            TShape{(index_t)stride_h, (index_t)stride_w}, TShape{(index_t)dil_h, (index_t)dil_w},
                                            (uint32_t)deformable_group, grad_offset, mxnet::kWriteTo);
  return (int)cudaDeviceSynchronize();
}

template <int dim> static mshadow::Tensor<gpu, dim, float> tens(const float* p, std::initializer_list<index_t> shp) {
  mshadow::Tensor<gpu, dim, float> t;
  t.dptr_ = const_cast<float*>(p);
  int i = 0;
  for (index_t v : shp) t.shape_.shape_[i++] = v;
  return t;
}

// op-level semantics of DeformablePSROIPoolingOp::Forward (deformable_psroi_pooling-inl.h:64-95): out = -FLT_MAX and
// top_count = 0 are written before the kernel; the caller passes buffers, the fill is done here the same way
extern "C" int ref_deform_psroi_forward(const float* data, const float* rois, const float* trans, int R, int C, int H, int W,
                                        int no_trans, float spatial_scale, int output_dim, int group_size, int pooled_size,
                                        int part_size, int sample_per_part, float trans_std, int num_classes, float* out,
                                        float* top_count) {
  const index_t ps = pooled_size;
  auto t_out = tens<4>(out, {(index_t)R, (index_t)output_dim, ps, ps});
  auto t_cnt = tens<4>(top_count, {(index_t)R, (index_t)output_dim, ps, ps});
  auto t_data = tens<4>(data, {1u, (index_t)C, (index_t)H, (index_t)W});
  auto t_box = tens<2>(rois, {(index_t)R, 5u});
  const index_t part = part_size > 0 ? part_size : pooled_size;
  auto t_trans = tens<4>(trans, {(index_t)R, (index_t)(2 * num_classes), part, part});
  std::vector<float> fill(t_out.shape_.Size(), -FLT_MAX);
  cudaMemcpy(out, fill.data(), fill.size() * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemset(top_count, 0, fill.size() * sizeof(float));
  mshadow::DeformablePSROIPoolForward<float>(t_out, t_data, t_box, t_trans, t_cnt, no_trans != 0, spatial_scale, output_dim,
                                             group_size, pooled_size, (int)part, sample_per_part, trans_std);
  return (int)cudaDeviceSynchronize();
}

// DeformablePSROIPoolingOp::Backward (-inl.h:97-151): in_grad / trans_grad zeroed (kWriteTo), then the atomicAdd kernel
extern "C" int ref_deform_psroi_backward(const float* out_grad, const float* data, const float* rois, const float* trans,
                                         const float* top_count, int R, int C, int H, int W, int no_trans,
                                         float spatial_scale, int output_dim, int group_size, int pooled_size, int part_size,
                                         int sample_per_part, float trans_std, int num_classes, float* in_grad,
                                         float* trans_grad) {
  const index_t ps = pooled_size;
  const index_t part = part_size > 0 ? part_size : pooled_size;
  auto t_og = tens<4>(out_grad, {(index_t)R, (index_t)output_dim, ps, ps});
  auto t_cnt = tens<4>(top_count, {(index_t)R, (index_t)output_dim, ps, ps});
  auto t_data = tens<4>(data, {1u, (index_t)C, (index_t)H, (index_t)W});
  auto t_ig = tens<4>(in_grad, {1u, (index_t)C, (index_t)H, (index_t)W});
  auto t_box = tens<2>(rois, {(index_t)R, 5u});
  auto t_trans = tens<4>(trans, {(index_t)R, (index_t)(2 * num_classes), part, part});
  auto t_tg = tens<4>(trans_grad, {(index_t)R, (index_t)(2 * num_classes), part, part});
  cudaMemset(in_grad, 0, (size_t)C * H * W * sizeof(float));
  if (!no_trans) cudaMemset(trans_grad, 0, t_tg.shape_.Size() * sizeof(float));
  mshadow::DeformablePSROIPoolBackwardAcc<float>(t_ig, t_tg, t_og, t_data, t_box, t_trans, t_cnt, no_trans != 0,
                                                 spatial_scale, output_dim, group_size, pooled_size, (int)part,
                                                 sample_per_part, trans_std);
  return (int)cudaDeviceSynchronize();
}
