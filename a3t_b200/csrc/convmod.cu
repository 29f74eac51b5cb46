// Conformer convolution-module core: GLU + depthwise Conv1d, forward and backward
// (conformer/convolution.py:70-75).  Channels-last tiles staged in shared memory with a halo.
#include "common.cuh"

namespace a3t {

constexpr int DW_TT = 64;    // time steps per tile
constexpr int DW_TC = 64;    // channels per tile
constexpr int DW_MAXK = 32;  // max depthwise kernel size
constexpr int DW_TPT = 16;   // time steps per thread (256 threads = 64 ch x 4 groups)

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
// bf16 activations: ex2.approx + rcp.approx (a few ulp of fp32, far below one bf16 ulp); the fp32 parity mode keeps expf
template <typename TU> __device__ __forceinline__ float sigmoid_t(float v) { return sigmoidf_(v); }
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

// 4 consecutive channels of row `ur` (a at c, gate at C + c) -> GLU values
template <typename TU>
__device__ __forceinline__ void load4(const TU* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename TU>
__device__ __forceinline__ float4 glu4(const TU* ur, int C, int c) {
  float a[4], g[4];
  load4<TU>(ur + c, a);
  load4<TU>(ur + C + c, g);
  return make_float4(a[0] * sigmoid_t<TU>(g[0]), a[1] * sigmoid_t<TU>(g[1]), a[2] * sigmoid_t<TU>(g[2]),
                     a[3] * sigmoid_t<TU>(g[3]));
}

template <typename TU>
__global__ void __launch_bounds__(256) glu_dwconv_fwd_kernel(const TU* __restrict__ u, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ z,
                                                             int B, int S, int C, int K) {
  A3T_PDL_TRIGGER();
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
  A3T_PDL_TRIGGER();
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


// ---------------------------------------------------------------------------------------------
// Compile-time kernel size KT (7 / 15 / 31): each thread owns one channel and DW_TPT consecutive
// time steps; the DW_TPT + KT - 1 inputs it needs are read from shared memory ONCE into a
// register window, so the inner loops are pure FMAs (the generic kernels above issue one shared
// load per FMA and are shared-memory bound).
// ---------------------------------------------------------------------------------------------
template <typename TU, int KT>
__global__ void __launch_bounds__(256) glu_dwconv_fwd_kt_kernel(const TU* __restrict__ u, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ z,
                                                                int B, int S, int C) {
  A3T_PDL_TRIGGER();
  constexpr int ROWS = DW_TT + KT - 1, WIN = DW_TPT + KT - 1, pad = (KT - 1) / 2;
  __shared__ __align__(16) float tile[ROWS][DW_TC];
  const int ntt = (S + DW_TT - 1) / DW_TT;
  const int b = blockIdx.x / ntt, t0 = (blockIdx.x % ntt) * DW_TT;
  const int c0 = blockIdx.y * DW_TC;
  // GLU of the haloed tile: 4 channels per thread (8-byte bf16 / 16-byte fp32 loads; C % 4 == 0)
  for (int idx = threadIdx.x; idx < ROWS * (DW_TC / 4); idx += 256) {
    int rr = idx / (DW_TC / 4), cc = (idx % (DW_TC / 4)) * 4;
    int t = t0 - pad + rr, c = c0 + cc;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < S && c < C) v = glu4<TU>(u + ((int64_t)b * S + t) * 2 * C, C, c);
    *reinterpret_cast<float4*>(&tile[rr][cc]) = v;
  }
  __syncthreads();
  const int cc = threadIdx.x % DW_TC, tg = threadIdx.x / DW_TC;
  const int c = c0 + cc;
  if (c >= C) return;
  float wk[KT];
#pragma unroll
  for (int k = 0; k < KT; k++) wk[k] = w[(int64_t)c * KT + k];
  const float bv = bias ? bias[c] : 0.f;
  float win[WIN];
#pragma unroll
  for (int i = 0; i < WIN; i++) win[i] = tile[tg * DW_TPT + i][cc];
#pragma unroll
  for (int tt = 0; tt < DW_TPT; tt++) {
    int t = t0 + tg * DW_TPT + tt;
    float acc = bv;
#pragma unroll
    for (int k = 0; k < KT; k++) acc = fmaf(wk[k], win[tt + k], acc);
    if (t < S) z[((int64_t)b * S + t) * C + c] = acc;
  }
}

template <typename TU, typename TDU, int KT>
__global__ void __launch_bounds__(256) glu_dwconv_bwd_kt_kernel(const float* __restrict__ dz, const TU* __restrict__ u,
                                                                const float* __restrict__ w, TDU* __restrict__ du,
                                                                float* __restrict__ partial, int B, int S, int C) {
  A3T_PDL_TRIGGER();
  constexpr int ROWS = DW_TT + KT - 1, WIN = DW_TPT + KT - 1, pad = (KT - 1) / 2, padr = KT - 1 - pad;
  __shared__ __align__(16) float smem_raw[2 * ROWS * DW_TC > 4 * (KT + 1) * DW_TC ? 2 * ROWS * DW_TC : 4 * (KT + 1) * DW_TC];
  float (*tdz)[DW_TC] = reinterpret_cast<float (*)[DW_TC]>(smem_raw);                 // dz, rows from t0-padr
  float (*tgl)[DW_TC] = reinterpret_cast<float (*)[DW_TC]>(smem_raw + ROWS * DW_TC);  // glu, rows from t0-pad
  float (*red)[KT + 1][DW_TC] = reinterpret_cast<float (*)[KT + 1][DW_TC]>(smem_raw);  // reused after sync
  const int ntt = (S + DW_TT - 1) / DW_TT;
  const int b = blockIdx.x / ntt, t0 = (blockIdx.x % ntt) * DW_TT;
  const int c0 = blockIdx.y * DW_TC;
  for (int idx = threadIdx.x; idx < ROWS * (DW_TC / 4); idx += 256) {
    int rr = idx / (DW_TC / 4), cc = (idx % (DW_TC / 4)) * 4;
    int c = c0 + cc;
    int t = t0 - padr + rr;
    float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < S && c < C) dv = *reinterpret_cast<const float4*>(dz + ((int64_t)b * S + t) * C + c);
    *reinterpret_cast<float4*>(&tdz[rr][cc]) = dv;
    t = t0 - pad + rr;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < S && c < C) v = glu4<TU>(u + ((int64_t)b * S + t) * 2 * C, C, c);
    *reinterpret_cast<float4*>(&tgl[rr][cc]) = v;
  }
  __syncthreads();
  const int cc = threadIdx.x % DW_TC, tg = threadIdx.x / DW_TC;
  const int c = c0 + cc;
  float aw[KT];
  float ab = 0.f;
#pragma unroll
  for (int k = 0; k < KT; k++) aw[k] = 0.f;
  if (c < C) {
    float wz[WIN];
#pragma unroll
    for (int i = 0; i < WIN; i++) wz[i] = tdz[tg * DW_TPT + i][cc];
    {
      float wk[KT];
#pragma unroll
      for (int k = 0; k < KT; k++) wk[k] = w[(int64_t)c * KT + k];
      // dglu[t] = sum_k w[k] * dz[t - k + pad]  ->  window index tt + KT - 1 - k
#pragma unroll
      for (int tt = 0; tt < DW_TPT; tt++) {
        int t = t0 + tg * DW_TPT + tt;
        float dg = 0.f;
#pragma unroll
        for (int k = 0; k < KT; k++) dg = fmaf(wk[k], wz[tt + KT - 1 - k], dg);
        if (t < S) {
          const TU* ur = u + ((int64_t)b * S + t) * 2 * C;
          float a = to_f32<TU>(ur[c]), g = to_f32<TU>(ur[C + c]);
          float sg = sigmoid_t<TU>(g);
          TDU* dur = du + ((int64_t)b * S + t) * 2 * C;
          dur[c] = from_f32<TDU>(dg * sg);
          dur[C + c] = from_f32<TDU>(dg * a * sg * (1.f - sg));
        }
      }
    }
    // dw[k] += dz[t] * glu[t + k - pad]: dz[t] = wz[tt + padr] (zero outside the sequence), glu window from tgl
    float wg[WIN];
#pragma unroll
    for (int i = 0; i < WIN; i++) wg[i] = tgl[tg * DW_TPT + i][cc];
#pragma unroll
    for (int tt = 0; tt < DW_TPT; tt++) {
      const float dzt = (t0 + tg * DW_TPT + tt < S) ? wz[tt + padr] : 0.f;
      ab += dzt;
#pragma unroll
      for (int k = 0; k < KT; k++) aw[k] = fmaf(dzt, wg[tt + k], aw[k]);
    }
  }
  __syncthreads();  // tiles are dead from here on; reuse them as reduction scratch
#pragma unroll
  for (int k = 0; k < KT; k++) red[tg][k][cc] = aw[k];
  red[tg][KT][cc] = ab;
  __syncthreads();
  // partial layout: [blockIdx.x][(K+1)][C]  (row K = dbias)
  for (int idx = threadIdx.x; idx < (KT + 1) * DW_TC; idx += 256) {
    int k = idx / DW_TC, c2 = idx % DW_TC;
    if (c0 + c2 >= C) continue;
    float t = red[0][k][c2] + red[1][k][c2] + red[2][k][c2] + red[3][k][c2];
    partial[((int64_t)blockIdx.x * (KT + 1) + k) * C + c0 + c2] = t;
  }
}

__global__ void __launch_bounds__(1024) dwconv_bwd_final_kernel(const float* __restrict__ partial,
                                                                float* __restrict__ dw, float* __restrict__ dbias,
                                                                int nblk, int C, int K) {
  A3T_PDL_TRIGGER();
  __shared__ float red[8][128];
  const int col = threadIdx.x & 127, part = threadIdx.x >> 7;
  const int idx = blockIdx.x * 128 + col;
  const int n = (K + 1) * C;
  float t = 0.f;
  if (idx < n) {
    int b = part;
    for (; b + 24 < nblk; b += 32) {
      float t0 = partial[(int64_t)b * n + idx], t1 = partial[(int64_t)(b + 8) * n + idx];
      float t2 = partial[(int64_t)(b + 16) * n + idx], t3 = partial[(int64_t)(b + 24) * n + idx];
      t += (t0 + t1) + (t2 + t3);
    }
    for (; b < nblk; b += 8) t += partial[(int64_t)b * n + idx];
  }
  red[part][col] = t;
  __syncthreads();
  if (part == 0 && idx < n) {
#pragma unroll
    for (int p2 = 1; p2 < 8; p2++) t += red[p2][col];
    int k = idx / C, c = idx % C;
    if (k < K) dw[(int64_t)c * K + k] = t;
    else if (dbias) dbias[c] = t;
  }
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
#define A3T_DW_FWD(KT)                                                                                          \
  {                                                                                                             \
    if (dtype_u == A3T_BF16)                                                                                    \
      glu_dwconv_fwd_kt_kernel<__nv_bfloat16, KT><<<grid, 256, 0, st>>>((const __nv_bfloat16*)u, w, bias, z, B, S, C); \
    else                                                                                                        \
      glu_dwconv_fwd_kt_kernel<float, KT><<<grid, 256, 0, st>>>((const float*)u, w, bias, z, B, S, C);          \
  }
  if (k == 7 && C % 4 == 0) A3T_DW_FWD(7)
  else if (k == 15 && C % 4 == 0) A3T_DW_FWD(15)
  else if (k == 31 && C % 4 == 0) A3T_DW_FWD(31)
  else if (dtype_u == A3T_BF16)
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
#define A3T_DW_BWD(KT)                                                                                          \
  {                                                                                                             \
    if (dtype_u == A3T_BF16)                                                                                    \
      glu_dwconv_bwd_kt_kernel<__nv_bfloat16, __nv_bfloat16, KT><<<grid, 256, 0, st>>>(                         \
          dz, (const __nv_bfloat16*)u, w, (__nv_bfloat16*)du, partial, B, S, C);                                \
    else                                                                                                        \
      glu_dwconv_bwd_kt_kernel<float, float, KT><<<grid, 256, 0, st>>>(dz, (const float*)u, w, (float*)du,      \
                                                                      partial, B, S, C);                        \
  }
  if (k == 7 && C % 4 == 0) A3T_DW_BWD(7)
  else if (k == 15 && C % 4 == 0) A3T_DW_BWD(15)
  else if (k == 31 && C % 4 == 0) A3T_DW_BWD(31)
  else if (dtype_u == A3T_BF16)
    glu_dwconv_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(
        dz, (const __nv_bfloat16*)u, w, (__nv_bfloat16*)du, partial, B, S, C, k);
  else
    glu_dwconv_bwd_kernel<float, float><<<grid, 256, 0, st>>>(dz, (const float*)u, w, (float*)du, partial, B, S, C, k);
  int rc = check_launch("glu_dwconv_bwd");
  if (rc) return rc;
  int n = (k + 1) * C;
  dwconv_bwd_final_kernel<<<(n + 127) / 128, 1024, 0, st>>>(partial, dw, dbias, nblk, C, k);
  return check_launch("glu_dwconv_bwd_final");
}
