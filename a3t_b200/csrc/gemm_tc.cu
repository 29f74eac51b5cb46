// tcgen05 tensor-core GEMM (placeholder until the kernel lands: reports "unsupported").
#include "common.cuh"
namespace a3t {
int gemm_tc_launch(const A3tGemmDesc*, const void*, const void*, void*, const float*, const float*, const void*,
                   const unsigned long long*, cudaStream_t, bool) {
  return A3T_ERR_UNSUPPORTED;
}
}  // namespace a3t
