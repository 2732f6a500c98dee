// relation_tc.cu -- fused tcgen05 relation path (placeholder until the kernels land in this file)
#include "common.cuh"
#include "relation.cuh"
namespace rn {
size_t relation_tc_workspace_bytes(const rn_relation_desc*) { return 0; }
int relation_tc(const rn_relation_desc*, const float*, const float*, const int*, const float*, const float*, const float*,
                const float*, const float*, const float*, const float*, const float*, float*, float*, void*, size_t,
                cudaStream_t) {
  set_error("RN_PREC_F16 relation path not built");
  return RN_ERR_INVALID;
}
size_t linear_tc_workspace_bytes(int, int, int) { return 0; }
int linear_tc(const float*, const float*, const float*, float*, int, int, int, int, void*, size_t, cudaStream_t) {
  set_error("RN_PREC_F16 linear path not built");
  return RN_ERR_INVALID;
}
}  // namespace rn
