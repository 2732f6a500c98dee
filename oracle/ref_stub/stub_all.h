// Minimal stand-ins for the MXNet 1.1.0 / mshadow / dmlc declarations that the reference's two deformable CUDA files use
// (relation_rcnn/operator_cxx/nn/deformable_im2col.cuh, relation_rcnn/operator_cxx/deformable_psroi_pooling.cu), so that
// those files compile FROM WHERE THEY LIE under /root/reference without the MXNet source tree.  TEST INFRASTRUCTURE
// (oracle/): nothing here is part of the product and no reference source is copied -- only the names the reference's
// host wrappers mention are declared: TShape, index_t, OpReqType, Stream<gpu>, Tensor<gpu,dim,DType>, kBaseThreadNum,
// cuda_get_num_blocks, CHECK_* / LOG(FATAL), MSHADOW_CUDA_POST_KERNEL_CHECK, CUDA_KERNEL_LOOP.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cfloat>
#include <iostream>
#include <sstream>
#include <vector>
#include <algorithm>

struct RefStubFatal {
  std::ostringstream os;
  bool fatal;
  explicit RefStubFatal(bool f) : fatal(f) {}
  template <typename T> RefStubFatal& operator<<(const T& v) { os << v; return *this; }
  ~RefStubFatal() { if (fatal) { fprintf(stderr, "reference CHECK/LOG(FATAL): %s\n", os.str().c_str()); abort(); } }
};
#define LOG(sev) RefStubFatal(true)
#define CHECK_OP_(a, b, op) if (!((a) op (b))) RefStubFatal(true) << #a " " #op " " #b " failed "
#define CHECK_EQ(a, b) CHECK_OP_(a, b, ==)
#define CHECK_LT(a, b) CHECK_OP_(a, b, <)
#define CHECK_LE(a, b) CHECK_OP_(a, b, <=)
#define CHECK_NE(a, b) CHECK_OP_(a, b, !=)
#define CHECK(c) if (!(c)) RefStubFatal(true) << #c " failed "

struct gpu {};
struct cpu {};
typedef uint32_t index_t;

namespace mshadow {
typedef uint32_t index_t;
using ::gpu;
using ::cpu;
template <typename Device> struct Stream {
  cudaStream_t stream_ = 0;
  static cudaStream_t GetStream(Stream<Device>* s) { return s ? s->stream_ : 0; }
};
template <int dim> struct Shape {
  index_t shape_[dim];
  index_t operator[](int i) const { return shape_[i]; }
  size_t Size() const { size_t n = 1; for (int i = 0; i < dim; ++i) n *= shape_[i]; return n; }
};
template <typename Device, int dim, typename DType> struct Tensor {
  DType* dptr_ = nullptr;
  Shape<dim> shape_;
  Stream<Device>* stream_ = nullptr;
  index_t size(int i) const { return shape_[i]; }
};
namespace cuda {
const int kBaseThreadBits = 8;
const int kBaseThreadNum = 1 << kBaseThreadBits;        // mshadow/cuda/tensor_gpu-inl.cuh
}  // namespace cuda
}  // namespace mshadow
#define MSHADOW_CUDA_POST_KERNEL_CHECK(x)                                                    \
  do {                                                                                        \
    cudaError_t err = cudaPeekAtLastError();                                                  \
    CHECK_EQ(err, cudaSuccess) << "Name: " << #x << " ErrStr:" << cudaGetErrorString(err);    \
  } while (0)
#define MSHADOW_REAL_TYPE_SWITCH(type, DType, ...) { typedef float DType; { __VA_ARGS__ } }

namespace mxnet {
using mshadow::index_t;
using ::gpu;
using ::cpu;
struct TShape {                                          // nnvm::TShape: only ndim / [] / ProdShape are used
  std::vector<index_t> d;
  TShape() {}
  TShape(std::initializer_list<index_t> l) : d(l) {}
  index_t ndim() const { return (index_t)d.size(); }
  index_t operator[](int i) const { return d[i]; }
  index_t ProdShape(int b, int e) const { index_t n = 1; for (int i = b; i < e; ++i) n *= d[i]; return n; }
};
enum OpReqType { kNullOp, kWriteTo, kWriteInplace, kAddTo };   // include/mxnet/op_attr_types.h
class Operator { public: virtual ~Operator() {} };
namespace op {
using mshadow::Tensor;
using mshadow::Stream;
namespace mxnet_op {
using namespace mshadow::cuda;
inline int cuda_get_num_blocks(const int N) {            // src/operator/mxnet_op.h: min(kMaxGridNum, ceil(N / kBaseThreadNum))
  const int kMaxGridNum = 65535;
  return std::min(kMaxGridNum, (N + kBaseThreadNum - 1) / kBaseThreadNum);
}
}  // namespace mxnet_op
// the operator class lives in deformable_psroi_pooling-inl.h (skipped: needs the MXNet tree); the .cu only names these
struct DeformablePSROIPoolingParam {};
template <typename xpu, typename DType> class DeformablePSROIPoolingOp : public Operator {
 public:
  explicit DeformablePSROIPoolingOp(DeformablePSROIPoolingParam) {}
};
template <typename xpu> Operator* CreateOp(DeformablePSROIPoolingParam param, int dtype);
}  // namespace op
}  // namespace mxnet
// src/operator/mxnet_op.h
#ifndef CUDA_KERNEL_LOOP
#define CUDA_KERNEL_LOOP(i, n) \
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += blockDim.x * gridDim.x)
#endif
