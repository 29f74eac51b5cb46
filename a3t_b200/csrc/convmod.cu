// Conformer convolution-module core: GLU + depthwise Conv1d, forward and backward
// (conformer/convolution.py:70-75).  Channels-last tiles staged in shared memory with a halo.
#include "common.cuh"

namespace a3t {

constexpr int DW_TT = 64;    // time steps per tile
constexpr int DW_TC = 64;    // channels per tile
constexpr int DW_MAXK = 32;  // max depthwise kernel size
constexpr int DW_TPT = 16;   // time steps per thread (256 threads = 64 ch x 4 groups)

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

template <typename TU>
__global__ void __launch_bounds__(256) glu_dwconv_fwd_kernel(const TU* __restrict__ u, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ z,
                                                             int B, int S, int C, int K) {
  __shared__ float tile[DW_TT + DW_MAXK - 1][DW_TC];
  const int pad = (K - 1) / 2;
  const int ntt = (S + DW_TT - 1) / DW_TT;
  const int b = blockIdx.x / ntt, t0 = (blockIdx.x % ntt) * DW_TT;
  const int c0 = blockIdx.y * DW_TC;
  const int rows = DW_TT + K - 1;
  for (int idx = threadIdx.x; idx < rows * DW_TC; idx += 256) {
    int rr = idx / DW_TC, cc = idx % DW_TC;
    int t = t0 - pad + rr, c = c0 + cc;
    float v = 0.f;
    if (t >= 0 && t < S && c < C) {
      const TU* ur = u + ((int64_t)b * S + t) * 2 * C;
      v = to_f32<TU>(ur[c]) * sigmoidf_(to_f32<TU>(ur[C + c]));
    }
    tile[rr][cc] = v;
  }
  __syncthreads();
  const int cc = threadIdx.x % DW_TC, tg = threadIdx.x / DW_TC;
  const int c = c0 + cc;
  if (c >= C) return;
  float wk[DW_MAXK];
#pragma unroll
  for (int k = 0; k < DW_MAXK; k++) wk[k] = k < K ? w[(int64_t)c * K + k] : 0.f;
  const float bv = bias ? bias[c] : 0.f;
  for (int tt = 0; tt < DW_TPT; tt++) {
    int tl = tg * DW_TPT + tt;
    int t = t0 + tl;
    if (t >= S) break;
    float acc = bv;
#pragma unroll
    for (int k = 0; k < DW_MAXK; k++)
      if (k < K) acc = fmaf(wk[k], tile[tl + k][cc], acc);
    z[((int64_t)b * S + t) * C + c] = acc;
  }
}

// backward: du = GLU'(dwconv^T(dz)), partial dw/dbias per (b, t-tile)
template <typename TU, typename TDU>
__global__ void __launch_bounds__(256) glu_dwconv_bwd_kernel(const float* __restrict__ dz, const TU* __restrict__ u,
                                                             const float* __restrict__ w, TDU* __restrict__ du,
                                                             float* __restrict__ partial, int B, int S, int C, int K) {
  constexpr int ROWS = DW_TT + DW_MAXK - 1;
  __shared__ float smem_raw[2 * ROWS * DW_TC];
  float (*tdz)[DW_TC] = reinterpret_cast<float (*)[DW_TC]>(smem_raw);                 // dz, rows from t0-(K-1-pad)
  float (*tgl)[DW_TC] = reinterpret_cast<float (*)[DW_TC]>(smem_raw + ROWS * DW_TC);  // glu, rows from t0-pad
  float (*red)[DW_MAXK + 1][DW_TC] = reinterpret_cast<float (*)[DW_MAXK + 1][DW_TC]>(smem_raw);  // reused after sync
  static_assert(4 * (DW_MAXK + 1) * DW_TC <= 2 * ROWS * DW_TC, "reduction scratch must fit");
  const int pad = (K - 1) / 2;
  const int padr = K - 1 - pad;
  const int ntt = (S + DW_TT - 1) / DW_TT;
  const int b = blockIdx.x / ntt, t0 = (blockIdx.x % ntt) * DW_TT;
  const int c0 = blockIdx.y * DW_TC;
  const int rows = DW_TT + K - 1;
  for (int idx = threadIdx.x; idx < rows * DW_TC; idx += 256) {
    int rr = idx / DW_TC, cc = idx % DW_TC;
    int c = c0 + cc;
    int t = t0 - padr + rr;
    tdz[rr][cc] = (t >= 0 && t < S && c < C) ? dz[((int64_t)b * S + t) * C + c] : 0.f;
    t = t0 - pad + rr;
    float v = 0.f;
    if (t >= 0 && t < S && c < C) {
      const TU* ur = u + ((int64_t)b * S + t) * 2 * C;
      v = to_f32<TU>(ur[c]) * sigmoidf_(to_f32<TU>(ur[C + c]));
    }
    tgl[rr][cc] = v;
  }
  __syncthreads();
  const int cc = threadIdx.x % DW_TC, tg = threadIdx.x / DW_TC;
  const int c = c0 + cc;
  float wk[DW_MAXK], aw[DW_MAXK];
  float ab = 0.f;
#pragma unroll
  for (int k = 0; k < DW_MAXK; k++) {
    wk[k] = (k < K && c < C) ? w[(int64_t)c * K + k] : 0.f;
    aw[k] = 0.f;
  }
  if (c < C) {
    for (int tt = 0; tt < DW_TPT; tt++) {
      int tl = tg * DW_TPT + tt;
      int t = t0 + tl;
      if (t >= S) break;
      // dglu[t] = sum_k w[k] * dz[t - k + pad]  ->  tdz row index (t - k + pad) - (t0 - padr) = tl + padr + pad - k
      float dg = 0.f;
#pragma unroll
      for (int k = 0; k < DW_MAXK; k++)
        if (k < K) dg = fmaf(wk[k], tdz[tl + K - 1 - k][cc], dg);
      const TU* ur = u + ((int64_t)b * S + t) * 2 * C;
      float a = to_f32<TU>(ur[c]), g = to_f32<TU>(ur[C + c]);
      float sg = sigmoidf_(g);
      TDU* dur = du + ((int64_t)b * S + t) * 2 * C;
      dur[c] = from_f32<TDU>(dg * sg);
      dur[C + c] = from_f32<TDU>(dg * a * sg * (1.f - sg));
      // dw[k] += dz[t] * glu[t + k - pad] ; dz[t] at tdz row tl + padr ; glu at tgl row tl + k
      float dzt = tdz[tl + padr][cc];
      ab += dzt;
#pragma unroll
      for (int k = 0; k < DW_MAXK; k++)
        if (k < K) aw[k] = fmaf(dzt, tgl[tl + k][cc], aw[k]);
    }
  }
  __syncthreads();  // tiles are dead from here on; reuse them as reduction scratch
#pragma unroll
  for (int k = 0; k < DW_MAXK; k++)
    if (k < K) red[tg][k][cc] = aw[k];
  red[tg][DW_MAXK][cc] = ab;
  __syncthreads();
  // partial layout: [blockIdx.x][(K+1)][C]  (row K = dbias)
  for (int idx = threadIdx.x; idx < (K + 1) * DW_TC; idx += 256) {
    int k = idx / DW_TC, c2 = idx % DW_TC;
    if (c0 + c2 >= C) continue;
    int kk = k < K ? k : DW_MAXK;
    float t = red[0][kk][c2] + red[1][kk][c2] + red[2][kk][c2] + red[3][kk][c2];
    partial[((int64_t)blockIdx.x * (K + 1) + k) * C + c0 + c2] = t;
  }
}

__global__ void dwconv_bwd_final_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                        float* __restrict__ dbias, int nblk, int C, int K) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (K + 1) * C) return;
  int k = idx / C, c = idx % C;
  float t = 0.f;
  for (int b = 0; b < nblk; b++) t += partial[((int64_t)b * (K + 1) + k) * C + c];
  if (k < K) dw[(int64_t)c * K + k] = t;
  else if (dbias) dbias[c] = t;
}

}  // namespace a3t

using namespace a3t;

extern "C" int a3t_glu_dwconv_fwd(const void* u, int dtype_u, const float* w, const float* bias, float* z, int B,
                                  int S, int C, int k, void* stream) {
  A3T_REQUIRE(u && w && z, "glu_dwconv_fwd: null pointer");
  A3T_REQUIRE(k >= 1 && k <= DW_MAXK && (k & 1), "glu_dwconv_fwd: kernel size %d must be odd and <= %d", k, DW_MAXK);
  if (B == 0 || S == 0) return A3T_OK;
  dim3 grid(B * ((S + DW_TT - 1) / DW_TT), (C + DW_TC - 1) / DW_TC);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_u == A3T_BF16)
    glu_dwconv_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)u, w, bias, z, B, S, C, k);
  else
    glu_dwconv_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)u, w, bias, z, B, S, C, k);
  return check_launch("glu_dwconv_fwd");
}

extern "C" int a3t_dwconv_bwd_blocks(int B, int S) { return B * ((S + DW_TT - 1) / DW_TT); }

extern "C" int a3t_glu_dwconv_bwd(const float* dz, const void* u, int dtype_u, const float* w, void* du, int dtype_du,
                                  float* dw, float* dbias, float* partial, int B, int S, int C, int k, void* stream) {
  A3T_REQUIRE(dz && u && w && du && dw && partial, "glu_dwconv_bwd: null pointer");
  A3T_REQUIRE(k >= 1 && k <= DW_MAXK && (k & 1), "glu_dwconv_bwd: kernel size %d must be odd and <= %d", k, DW_MAXK);
  A3T_REQUIRE(dtype_u == dtype_du, "glu_dwconv_bwd: u and du must share a dtype");
  if (B == 0 || S == 0) return A3T_OK;
  int nblk = a3t_dwconv_bwd_blocks(B, S);
  dim3 grid(nblk, (C + DW_TC - 1) / DW_TC);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_u == A3T_BF16)
    glu_dwconv_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(
        dz, (const __nv_bfloat16*)u, w, (__nv_bfloat16*)du, partial, B, S, C, k);
  else
    glu_dwconv_bwd_kernel<float, float><<<grid, 256, 0, st>>>(dz, (const float*)u, w, (float*)du, partial, B, S, C, k);
  int rc = check_launch("glu_dwconv_bwd");
  if (rc) return rc;
  int n = (k + 1) * C;
  dwconv_bwd_final_kernel<<<(n + 255) / 256, 256, 0, st>>>(partial, dw, dbias, nblk, C, k);
  return check_launch("glu_dwconv_bwd_final");
}
