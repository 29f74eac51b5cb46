// Shared device/host helpers for the a3t_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/a3t_b200.h"

namespace a3t {

// ---- error plumbing: never throw/exit across the C ABI ------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // returns A3T_OK or A3T_ERR_CUDA after cudaGetLastError

#define A3T_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      a3t::set_error(__VA_ARGS__);                \
      return A3T_ERR_ARG;                         \
    }                                             \
  } while (0)

// ---- dtype helpers ------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float load_as_f32(const void* p, int dtype, int64_t i) {
  return dtype == A3T_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void store_from_f32(void* p, int dtype, int64_t i, float v) {
  if (dtype == A3T_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

// ---- programmatic dependent launch --------------------------------------------------------
// Every kernel signals at its first instruction that a dependent grid may be scheduled: a kernel launched
// with cudaLaunchAttributeProgrammaticStreamSerialization (the persistent GEMM) then runs its prologue
// (barrier init, TMEM allocation, descriptor prefetch) while this grid drains, and blocks in
// A3T_PDL_WAIT() -- which returns only when this grid has completed and flushed -- before it touches memory.
// Without the attribute on the dependent's launch both instructions are no-ops.
#define A3T_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define A3T_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")

// ---- stateless dropout mask (bit-identical twin: oracle/a3t_oracle.py::keep_mask) -----------
// One 32-bit integer hash of (element-pair index, seed, site) decides TWO consecutive elements: the even
// element of the pair takes the low 16 bits, the odd one the high 16 bits; keep iff that 16-bit value
// >= p * 2^16.  About five integer instructions per element: the mask is regenerated in every
// epilogue/backward instead of stored.
struct Drop {
  uint32_t thr;      // keep iff r16 >= thr ; thr = (uint32)(p * 2^16)
  float inv_keep;    // 1/(1-p)
  uint32_t k0, k1;   // derived from seed and site
  bool on;
};
__device__ __forceinline__ Drop make_drop(float p, const unsigned long long* seed_ptr, uint32_t site) {
  Drop d;
  d.on = p > 0.f;
  d.thr = 0; d.inv_keep = 1.f; d.k0 = 0; d.k1 = 0;
  if (d.on) {
    unsigned long long seed = *seed_ptr;
    d.thr = (uint32_t)(p * 65536.0f);
    d.inv_keep = 1.0f / (1.0f - p);
    d.k0 = (uint32_t)(seed & 0xFFFFFFFFull) + site * 0x9E3779B9u;
    d.k1 = (uint32_t)(seed >> 32);
  }
  return d;
}
// fold a 64-bit element index to the 32-bit hash input (identity for tensors below 2^32 elements)
__device__ __host__ __forceinline__ uint32_t drop_fold(unsigned long long idx) {
  return (uint32_t)(idx & 0xFFFFFFFFull) ^ ((uint32_t)(idx >> 32) * 0x85EBCA6Bu);
}
// hash of one element pair (pair = folded index >> 1)
__device__ __forceinline__ uint32_t drop_hash(const Drop& d, uint32_t pair) {
  uint32_t x = (pair ^ d.k0) * 0x9E3779B1u + d.k1;
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  return x;
}
__device__ __forceinline__ bool drop_keep32(const Drop& d, uint32_t folded) {
  const uint32_t h = drop_hash(d, folded >> 1);
  return ((folded & 1u) ? (h >> 16) : (h & 0xFFFFu)) >= d.thr;
}
// keep flags of the 4 consecutive elements whose folded indices are f0 ^ 0..3 (element index % 4 == 0)
__device__ __forceinline__ void drop_keep4(const Drop& d, uint32_t f0, bool (&k)[4]) {
  uint32_t ha = drop_hash(d, f0 >> 1), hb = drop_hash(d, (f0 >> 1) ^ 1u);
  if (f0 & 1u) {  // only for tensors beyond 2^32 elements: the fold may flip the parity bit
    ha = __funnelshift_l(ha, ha, 16);
    hb = __funnelshift_l(hb, hb, 16);
  }
  k[0] = (ha & 0xFFFFu) >= d.thr; k[1] = (ha >> 16) >= d.thr;
  k[2] = (hb & 0xFFFFu) >= d.thr; k[3] = (hb >> 16) >= d.thr;
}
__device__ __forceinline__ bool drop_keep(const Drop& d, unsigned long long idx) {
  return drop_keep32(d, drop_fold(idx));
}
__device__ __forceinline__ float drop_apply(const Drop& d, unsigned long long idx, float v) {
  if (!d.on) return v;
  return drop_keep(d, idx) ? v * d.inv_keep : 0.f;
}

// ---- warp / block reductions ---------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Tuning / experiment switches (A3T_TC_*, A3T_SOFTMAX_SMEM ...) exist only in builds made with -DA3T_TUNING
// (`make TUNING=1`); the release library never reads the environment.
#ifdef A3T_TUNING
#include <stdlib.h>
static inline const char* tune_env(const char* name) { return getenv(name); }
#else
static inline const char* tune_env(const char*) { return nullptr; }
#endif

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace a3t
