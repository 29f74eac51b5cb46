// a3t_gemm: argument validation and dispatch between the tcgen05 tensor-core kernel (gemm_tc.cu)
// and the exact-fp32 CUDA-core kernel (gemm_simt.cu).  Contract: include/a3t_b200.h.
#include "common.cuh"

namespace a3t {
int gemm_simt_launch(const A3tGemmDesc* d, const void* A, const void* B, void* C, const float* bias,
                     const float* res, const void* mask, const unsigned long long* seed, cudaStream_t st);
// returns A3T_ERR_UNSUPPORTED (without setting an error) when the problem does not qualify
int gemm_tc_launch(const A3tGemmDesc* d, const void* A, const void* B, void* C, const float* bias,
                   const float* res, const void* mask, const unsigned long long* seed, cudaStream_t st,
                   bool probe_only);
}  // namespace a3t

using namespace a3t;

// bf16 problems that A3T_IMPL_AUTO handed to the CUDA-core kernel because they did not qualify for the
// tcgen05 one (a silent 5x slowdown when it happens on a hot shape): counted so callers can assert on it
static unsigned long long g_auto_fallbacks = 0;

extern "C" int a3t_gemm_fallback_count(int reset) {
  const unsigned long long n = __atomic_load_n(&g_auto_fallbacks, __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&g_auto_fallbacks, 0ull, __ATOMIC_RELAXED);
  return n > 0x7fffffffull ? 0x7fffffff : (int)n;
}

extern "C" int a3t_gemm(const A3tGemmDesc* d, const void* A, const void* B, void* C, const float* bias,
                        const float* res, const void* mask, const unsigned long long* seed, void* stream) {
  A3T_REQUIRE(d && A && B && C, "gemm: null pointer");
  A3T_REQUIRE(d->M >= 0 && d->N >= 0 && d->K >= 0 && d->batch1 >= 1 && d->batch2 >= 1, "gemm: bad sizes");
  A3T_REQUIRE(d->mode >= A3T_GEMM_PLAIN && d->mode <= A3T_GEMM_WGRAD, "gemm: bad mode %d", d->mode);
  if (d->mode == A3T_GEMM_CONV) {
    A3T_REQUIRE(d->taps >= 1 && d->cin >= 1 && d->K == d->taps * d->cin && d->seq >= 1 && d->M % d->seq == 0,
                "gemm(conv): need K == taps*cin and M %% seq == 0 (M=%d K=%d taps=%d cin=%d seq=%d)", d->M, d->K,
                d->taps, d->cin, d->seq);
  }
  if (d->mode == A3T_GEMM_WGRAD) {
    A3T_REQUIRE(d->taps >= 1 && d->cin >= 1 && d->N == d->taps * d->cin && d->seq >= 1 && d->K % d->seq == 0,
                "gemm(wgrad): need N == taps*cin and K %% seq == 0 (N=%d K=%d taps=%d cin=%d seq=%d)", d->N, d->K,
                d->taps, d->cin, d->seq);
  }
  A3T_REQUIRE(d->drop_p == 0.f || seed, "gemm: dropout needs a seed");
  A3T_REQUIRE(d->drop_p >= 0.f && d->drop_p < 1.f, "gemm: bad dropout p");
  if (d->M == 0 || d->N == 0) return A3T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->impl != A3T_IMPL_SIMT) {
    int rc = gemm_tc_launch(d, A, B, C, bias, res, mask, seed, st, false);
    if (rc != A3T_ERR_UNSUPPORTED) return rc;
    if (d->impl == A3T_IMPL_TC || d->impl == A3T_IMPL_TC_PAIR) {
      set_error("gemm: problem does not qualify for the tcgen05 kernel (M=%d N=%d K=%d mode=%d)", d->M, d->N, d->K,
                d->mode);
      return A3T_ERR_UNSUPPORTED;
    }
    if (d->dtype_a == A3T_BF16 && d->dtype_b == A3T_BF16) __atomic_fetch_add(&g_auto_fallbacks, 1ull, __ATOMIC_RELAXED);
  }
  return gemm_simt_launch(d, A, B, C, bias, res, mask, seed, st);
}

extern "C" int a3t_gemm_tc_supported(const A3tGemmDesc* d, const void* A, const void* B, void* C) {
  if (!d) return 0;
  return gemm_tc_launch(d, A, B, C, nullptr, nullptr, nullptr, nullptr, nullptr, true) == A3T_OK ? 1 : 0;
}
