// tcgen05 / TMEM / TMA implicit-GEMM kernels (placeholder until the kernels land).
#include "e2_common.cuh"
#include "e2_conv_internal.cuh"

bool e2_gather_gemm_tc_ok(const e2_handle*, const GatherGemm&) { return false; }
int e2_launch_gather_gemm_tc(e2_handle* h, const GatherGemm&, cudaStream_t) {
  return e2_fail(h, E2_ERR_UNSUPPORTED, "tcgen05 gather-GEMM not built");
}
bool e2_reduce_gemm_tc_ok(const e2_handle*, const ReduceGemm&) { return false; }
int e2_launch_reduce_gemm_tc(e2_handle* h, const ReduceGemm&, cudaStream_t) {
  return e2_fail(h, E2_ERR_UNSUPPORTED, "tcgen05 reduce-GEMM not built");
}
