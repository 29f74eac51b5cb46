// LayerNorm / BatchNorm / column-sum / scale-dropout kernels (bandwidth-bound, fp32 math).
// Contracts: include/a3t_b200.h.  Reference arithmetic: transformer/layer_norm.py:23,
// conformer/encoder.py:404, conformer/convolution.py:76, tacotron2/decoder.py:203.
#include "common.cuh"

namespace a3t {

constexpr int LN_WARPS = 8;
constexpr int LN_MAXV = 4;  // float4 per lane -> C <= 512

// ---------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, row held in registers.
// ---------------------------------------------------------------------------------------------
template <typename TY>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    TY* __restrict__ y, float* __restrict__ mean_o, float* __restrict__ rstd_o, int64_t rows, int C,
    float eps, int relu, float out_scale, float drop_p, const unsigned long long* __restrict__ seed,
    uint32_t site) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = C >> 2;  // float4 per row
  Drop dr = make_drop(drop_p, seed, site);
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < rows; row += (int64_t)gridDim.x * LN_WARPS) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; i++) {
      int c4 = lane + i * 32;
      if (c4 < nv) {
        v[i] = xr[c4];
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
    }
    s = warp_sum(s);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; i++) {
      int c4 = lane + i * 32;
      if (c4 < nv) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (lane == 0) {
      if (mean_o) mean_o[row] = mean;
      if (rstd_o) rstd_o[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < LN_MAXV; i++) {
      int c4 = lane + i * 32;
      if (c4 < nv) {
        float4 g = reinterpret_cast<const float4*>(gamma)[c4];
        float4 b = reinterpret_cast<const float4*>(beta)[c4];
        float o[4] = {(v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                      (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float t = o[e];
          if (relu) t = fmaxf(t, 0.f);
          t *= out_scale;
          t = drop_apply(dr, (unsigned long long)row * C + c4 * 4 + e, t);
          y[row * C + c4 * 4 + e] = from_f32<TY>(t);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward: warp per row; per-lane dgamma/dbeta register accumulators, block-reduced
// into partial[blk][2][C]; a second kernel finishes the column sums.
// ---------------------------------------------------------------------------------------------
template <typename TDY>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(
    const TDY* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_i,
    const float* __restrict__ rstd_i, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ dres, float* __restrict__ dx, float* __restrict__ partial, int64_t rows,
    int C, int relu, float out_scale, float drop_p, const unsigned long long* __restrict__ seed,
    uint32_t site) {
  extern __shared__ float sm[];  // [LN_WARPS][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = C >> 2;
  Drop dr = make_drop(drop_p, seed, site);
  float4 ag[LN_MAXV], ab[LN_MAXV], gm[LN_MAXV], bt[LN_MAXV];
#pragma unroll
  for (int i = 0; i < LN_MAXV; i++) {
    ag[i] = make_float4(0, 0, 0, 0);
    ab[i] = make_float4(0, 0, 0, 0);
    int c4 = lane + i * 32;
    if (c4 < nv) {
      gm[i] = reinterpret_cast<const float4*>(gamma)[c4];
      bt[i] = reinterpret_cast<const float4*>(beta)[c4];
    }
  }
  const float scale_keep = out_scale * dr.inv_keep;
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < rows; row += (int64_t)gridDim.x * LN_WARPS) {
    const float mean = mean_i[row], rstd = rstd_i[row];
    float xh[LN_MAXV][4], dh[LN_MAXV][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; i++) {
      int c4 = lane + i * 32;
      if (c4 < nv) {
        float4 xv = reinterpret_cast<const float4*>(x + row * C)[c4];
        float xe[4] = {xv.x, xv.y, xv.z, xv.w};
        float ge[4] = {gm[i].x, gm[i].y, gm[i].z, gm[i].w};
        float be[4] = {bt[i].x, bt[i].y, bt[i].z, bt[i].w};
        float gacc[4], bacc[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          int64_t idx = row * C + c4 * 4 + e;
          float g = to_f32<TDY>(dy[idx]);
          if (dr.on) g = drop_keep(dr, (unsigned long long)idx) ? g : 0.f;
          g *= scale_keep;
          float h = (xe[e] - mean) * rstd;
          if (relu && (h * ge[e] + be[e]) <= 0.f) g = 0.f;
          xh[i][e] = h;
          gacc[e] = g * h;
          bacc[e] = g;
          float d = g * ge[e];
          dh[i][e] = d;
          s1 += d;
          s2 += d * h;
        }
        ag[i].x += gacc[0]; ag[i].y += gacc[1]; ag[i].z += gacc[2]; ag[i].w += gacc[3];
        ab[i].x += bacc[0]; ab[i].y += bacc[1]; ab[i].z += bacc[2]; ab[i].w += bacc[3];
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; i++) {
      int c4 = lane + i * 32;
      if (c4 < nv) {
        float4 o;
        o.x = rstd * (dh[i][0] - s1 - xh[i][0] * s2);
        o.y = rstd * (dh[i][1] - s1 - xh[i][1] * s2);
        o.z = rstd * (dh[i][2] - s1 - xh[i][2] * s2);
        o.w = rstd * (dh[i][3] - s1 - xh[i][3] * s2);
        if (dres) {
          float4 r = reinterpret_cast<const float4*>(dres + row * C)[c4];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        reinterpret_cast<float4*>(dx + row * C)[c4] = o;
      }
    }
  }
  // block reduce the parameter gradients
  float* sg = sm + (size_t)warp * 2 * C;
#pragma unroll
  for (int i = 0; i < LN_MAXV; i++) {
    int c4 = lane + i * 32;
    if (c4 < nv) {
      reinterpret_cast<float4*>(sg)[c4] = ag[i];
      reinterpret_cast<float4*>(sg + C)[c4] = ab[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) t += sm[(size_t)w * 2 * C + c];
    partial[(size_t)blockIdx.x * 2 * C + c] = t;
  }
}

// Final row reduction shared by every "partial[nblk][n] -> out[n]" step of this file: a block owns 128
// columns and splits the nblk rows over 8 thread groups (4 independent loads in flight per thread),
// then combines the 8 partial sums through shared memory.  (The first version walked the rows
// serially with one thread per column: 10-40 us of pure latency per call.)
template <typename T>
__device__ __forceinline__ T reduce_rows_128x8(const T* __restrict__ partial, int nblk, int64_t n, int64_t idx,
                                               T (*red)[128]) {
  const int col = threadIdx.x & 127, part = threadIdx.x >> 7;
  T t = 0;
  if (idx < n) {
    int b = part;
    for (; b + 24 < nblk; b += 32) {
      T t0 = partial[(int64_t)b * n + idx], t1 = partial[(int64_t)(b + 8) * n + idx];
      T t2 = partial[(int64_t)(b + 16) * n + idx], t3 = partial[(int64_t)(b + 24) * n + idx];
      t += (t0 + t1) + (t2 + t3);
    }
    for (; b < nblk; b += 8) t += partial[(int64_t)b * n + idx];
  }
  red[part][col] = t;
  __syncthreads();
  if (part == 0) {
#pragma unroll
    for (int p2 = 1; p2 < 8; p2++) t += red[p2][col];
  }
  return t;  // valid for part == 0
}

// out[c] = sum_b partial[b*2C + c]  (c < 2C): first C -> out0, rest -> out1
__global__ void __launch_bounds__(1024) reduce_partial_kernel(const float* __restrict__ partial,
                                                              float* __restrict__ out0, float* __restrict__ out1,
                                                              int nblk, int C) {
  __shared__ float red[8][128];
  const int c = blockIdx.x * 128 + (threadIdx.x & 127);
  float t = reduce_rows_128x8<float>(partial, nblk, 2 * C, c, red);
  if ((threadIdx.x >> 7) != 0 || c >= 2 * C) return;
  if (c < C) { if (out0) out0[c] = t; }
  else if (out1) out1[c - C] = t;
}

// ---------------------------------------------------------------------------------------------
// column sums  out[c] = sum_r x[r*ldx + c]
// grid (ceil(C/256), nblk); partial[nblk][C]
// ---------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 128;
static int colsum_blocks_host(int64_t rows) {
  int64_t n = (rows + CS_ROWS - 1) / CS_ROWS;
  if (n > 592) n = 592;
  if (n < 1) n = 1;
  return (int)n;
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ partial,
                                                     int64_t rows, int C, int64_t ldx) {
  int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  int nblk = gridDim.y;
  int64_t per = (rows + nblk - 1) / nblk;
  int64_t r0 = (int64_t)blockIdx.y * per, r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int64_t r = r0;
  for (; r + 3 < r1; r += 4) {  // 4 independent loads in flight
    float v0 = to_f32<T>(x[r * ldx + c]), v1 = to_f32<T>(x[(r + 1) * ldx + c]);
    float v2 = to_f32<T>(x[(r + 2) * ldx + c]), v3 = to_f32<T>(x[(r + 3) * ldx + c]);
    a0 += v0; a1 += v1; a2 += v2; a3 += v3;
  }
  for (; r < r1; r++) a0 += to_f32<T>(x[r * ldx + c]);
  partial[(size_t)blockIdx.y * C + c] = (a0 + a1) + (a2 + a3);
}
__global__ void __launch_bounds__(1024) colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                            int nblk, int C) {
  __shared__ float red[8][128];
  const int c = blockIdx.x * 128 + (threadIdx.x & 127);
  float t = reduce_rows_128x8<float>(partial, nblk, C, c, red);
  if ((threadIdx.x >> 7) == 0 && c < C) out[c] = t;
}

// masked column sums for NewMaskInputLayer backward (mlm_encoder.py:67-70)
__global__ void __launch_bounds__(256) masked_colsum_kernel(const float* __restrict__ x,
                                                            const uint8_t* __restrict__ masked,
                                                            float* __restrict__ partial, int64_t rows, int C) {
  int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  int nblk = gridDim.y;
  int64_t per = (rows + nblk - 1) / nblk;
  int64_t r0 = (int64_t)blockIdx.y * per, r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; r++)
    if (masked[r]) acc += x[r * C + c];
  partial[(size_t)blockIdx.y * C + c] = acc;
}

// ---------------------------------------------------------------------------------------------
// scale + dropout elementwise
// ---------------------------------------------------------------------------------------------
template <typename TY>
__global__ void __launch_bounds__(256) scale_dropout_kernel(const float* __restrict__ x, TY* __restrict__ y,
                                                            int64_t n, float scale, float drop_p,
                                                            const unsigned long long* __restrict__ seed,
                                                            uint32_t site) {
  Drop dr = make_drop(drop_p, seed, site);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = from_f32<TY>(drop_apply(dr, (unsigned long long)i, x[i] * scale));
}

// ---------------------------------------------------------------------------------------------
// BatchNorm1d statistics: double column sums of z and z^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_partial_kernel(const float* __restrict__ z, double* __restrict__ partial,
                                                         int64_t rows, int C) {
  int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  int nblk = gridDim.y;
  int64_t per = (rows + nblk - 1) / nblk;
  int64_t r0 = (int64_t)blockIdx.y * per, r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  double s = 0.0, q = 0.0;
  for (int64_t r = r0; r < r1; r++) {
    double v = (double)z[r * C + c];
    s += v;
    q += v * v;
  }
  partial[((size_t)blockIdx.y * 2) * C + c] = s;
  partial[((size_t)blockIdx.y * 2 + 1) * C + c] = q;
}
__global__ void __launch_bounds__(1024) bn_final_kernel(const double* __restrict__ partial, float* __restrict__ mean_o,
                                                        float* __restrict__ rstd_o, float* __restrict__ running_mean,
                                                        float* __restrict__ running_var, int64_t* __restrict__ nbt,
                                                        int nblk, int64_t rows, int C, float momentum, float eps,
                                                        int training) {
  __shared__ double red[8][128];
  const int c = blockIdx.x * 128 + (threadIdx.x & 127);
  const bool lead = (threadIdx.x >> 7) == 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && nbt) *nbt += 1;
  if (!training) {
    if (lead && c < C) {
      mean_o[c] = running_mean[c];
      rstd_o[c] = rsqrtf(running_var[c] + eps);
    }
    return;
  }
  // partial rows alternate [sum | sum of squares] with row length C: view as nblk rows of 2C
  double s = reduce_rows_128x8<double>(partial, nblk, 2 * (int64_t)C, c < C ? c : (int64_t)2 * C, red);
  __syncthreads();
  double q = reduce_rows_128x8<double>(partial, nblk, 2 * (int64_t)C, c < C ? (int64_t)C + c : (int64_t)2 * C, red);
  if (!lead || c >= C) return;
  double n = (double)rows;
  double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_o[c] = (float)mean;
  rstd_o[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    double unb = rows > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
  }
}

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == A3T_ACT_SWISH) return v / (1.f + expf(-v));
  if (act == A3T_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float act_grad(float v, int act) {
  if (act == A3T_ACT_SWISH) {
    float s = 1.f / (1.f + expf(-v));
    return s * (1.f + v * (1.f - s));
  }
  if (act == A3T_ACT_TANH) {
    float t = tanhf(v);
    return 1.f - t * t;
  }
  return 1.f;
}

template <typename TY>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta,
                                                         const float* __restrict__ res, TY* __restrict__ y,
                                                         int64_t rows, int C, int act, float drop_p,
                                                         const unsigned long long* __restrict__ seed, uint32_t site) {
  Drop dr = make_drop(drop_p, seed, site);
  int64_t n = rows * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c = (int)(i % C);
    float v = (z[i] - mean[c]) * rstd[c] * gamma[c] + beta[c];
    v = act_fwd(v, act);
    v = drop_apply(dr, (unsigned long long)i, v);
    if (res) v += res[i];
    y[i] = from_f32<TY>(v);
  }
}

// pass 1 of the backward: column sums of dy' and dy'*zhat (double)
__global__ void __launch_bounds__(256) bn_bwd_partial_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             double* __restrict__ partial, int64_t rows, int C,
                                                             int act, float drop_p,
                                                             const unsigned long long* __restrict__ seed,
                                                             uint32_t site) {
  int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  Drop dr = make_drop(drop_p, seed, site);
  int nblk = gridDim.y;
  int64_t per = (rows + nblk - 1) / nblk;
  int64_t r0 = (int64_t)blockIdx.y * per, r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  const float mu = mean[c], rs = rstd[c], g = gamma[c], b = beta[c];
  double sb = 0.0, sg = 0.0;
  for (int64_t r = r0; r < r1; r++) {
    int64_t i = r * C + c;
    float zh = (z[i] - mu) * rs;
    float d = dy[i];
    if (dr.on) d = drop_keep(dr, (unsigned long long)i) ? d * dr.inv_keep : 0.f;
    d *= act_grad(zh * g + b, act);
    sb += (double)d;
    sg += (double)d * (double)zh;
  }
  partial[((size_t)blockIdx.y * 2) * C + c] = sb;
  partial[((size_t)blockIdx.y * 2 + 1) * C + c] = sg;
}
__global__ void __launch_bounds__(1024) bn_bwd_final_kernel(const double* __restrict__ partial,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ coef, int nblk, int C) {
  __shared__ double red[8][128];
  const int c = blockIdx.x * 128 + (threadIdx.x & 127);
  double sb = reduce_rows_128x8<double>(partial, nblk, 2 * (int64_t)C, c < C ? c : (int64_t)2 * C, red);
  __syncthreads();
  double sg = reduce_rows_128x8<double>(partial, nblk, 2 * (int64_t)C, c < C ? (int64_t)C + c : (int64_t)2 * C, red);
  if ((threadIdx.x >> 7) != 0 || c >= C) return;
  dbeta[c] = (float)sb;
  dgamma[c] = (float)sg;
  coef[c] = (float)sb;
  coef[C + c] = (float)sg;
}
__global__ void __launch_bounds__(256) bn_bwd_dz_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ coef, float* __restrict__ dz,
                                                        int64_t rows, int C, int act, int training, float drop_p,
                                                        const unsigned long long* __restrict__ seed, uint32_t site) {
  Drop dr = make_drop(drop_p, seed, site);
  int64_t n = rows * C;
  const float inv_n = 1.f / (float)rows;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c = (int)(i % C);
    float zh = (z[i] - mean[c]) * rstd[c];
    float d = dy[i];
    if (dr.on) d = drop_keep(dr, (unsigned long long)i) ? d * dr.inv_keep : 0.f;
    d *= act_grad(zh * gamma[c] + beta[c], act);
    float o;
    if (training) o = gamma[c] * rstd[c] * (d - coef[c] * inv_n - zh * coef[C + c] * inv_n);
    else o = gamma[c] * rstd[c] * d;
    dz[i] = o;
  }
}

static int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace a3t

using namespace a3t;

extern "C" int a3t_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int dtype_y,
                                 float* mean, float* rstd, int64_t rows, int C, float eps, int relu,
                                 float out_scale, float drop_p, const unsigned long long* seed, uint32_t site,
                                 void* stream) {
  A3T_REQUIRE(x && gamma && beta && y, "layernorm_fwd: null pointer");
  A3T_REQUIRE(C % 4 == 0 && C <= 128 * LN_MAXV, "layernorm_fwd: C=%d must be a multiple of 4 and <= %d", C, 128 * LN_MAXV);
  A3T_REQUIRE(drop_p == 0.f || seed, "layernorm_fwd: dropout needs a seed");
  if (rows == 0) return A3T_OK;
  int blocks = (int)((rows + LN_WARPS - 1) / LN_WARPS);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_y == A3T_BF16)
    ln_fwd_kernel<__nv_bfloat16><<<blocks, LN_WARPS * 32, 0, st>>>(x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows,
                                                                  C, eps, relu, out_scale, drop_p, seed, site);
  else
    ln_fwd_kernel<float><<<blocks, LN_WARPS * 32, 0, st>>>(x, gamma, beta, (float*)y, mean, rstd, rows, C, eps, relu,
                                                          out_scale, drop_p, seed, site);
  return check_launch("layernorm_fwd");
}

extern "C" int a3t_layernorm_bwd_blocks(int64_t rows) {
  int64_t b = (rows + LN_WARPS * 4 - 1) / (LN_WARPS * 4);
  if (b > 148 * 2) b = 148 * 2;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int a3t_layernorm_bwd(const void* dy, int dtype_dy, const float* x, const float* mean, const float* rstd,
                                 const float* gamma, const float* beta, const float* dres, float* dx, float* dgamma,
                                 float* dbeta, float* partial, int64_t rows, int C, int relu, float out_scale,
                                 float drop_p, const unsigned long long* seed, uint32_t site, void* stream) {
  A3T_REQUIRE(dy && x && mean && rstd && gamma && beta && dx && partial, "layernorm_bwd: null pointer");
  A3T_REQUIRE(C % 4 == 0 && C <= 128 * LN_MAXV, "layernorm_bwd: C=%d must be a multiple of 4 and <= %d", C, 128 * LN_MAXV);
  A3T_REQUIRE(drop_p == 0.f || seed, "layernorm_bwd: dropout needs a seed");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = a3t_layernorm_bwd_blocks(rows);
  size_t smem = (size_t)LN_WARPS * 2 * C * sizeof(float);
  if (dtype_dy == A3T_BF16)
    ln_bwd_kernel<__nv_bfloat16><<<nblk, LN_WARPS * 32, smem, st>>>((const __nv_bfloat16*)dy, x, mean, rstd, gamma,
                                                                   beta, dres, dx, partial, rows, C, relu, out_scale,
                                                                   drop_p, seed, site);
  else
    ln_bwd_kernel<float><<<nblk, LN_WARPS * 32, smem, st>>>((const float*)dy, x, mean, rstd, gamma, beta, dres, dx,
                                                           partial, rows, C, relu, out_scale, drop_p, seed, site);
  int rc = check_launch("layernorm_bwd");
  if (rc) return rc;
  if (dgamma || dbeta) {
    reduce_partial_kernel<<<(2 * C + 127) / 128, 1024, 0, st>>>(partial, dgamma, dbeta, nblk, C);
    rc = check_launch("layernorm_bwd_reduce");
  }
  return rc;
}

extern "C" int a3t_colsum_blocks(int64_t rows) { return colsum_blocks_host(rows); }

extern "C" int a3t_colsum(const void* x, int dtype_x, float* out, float* partial, int64_t rows, int C, int64_t ldx,
                          void* stream) {
  A3T_REQUIRE(x && out && partial && C > 0, "colsum: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = colsum_blocks_host(rows);
  dim3 grid((C + 255) / 256, nblk);
  if (dtype_x == A3T_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, partial, rows, C, ldx);
  else
    colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)x, partial, rows, C, ldx);
  int rc = check_launch("colsum");
  if (rc) return rc;
  colsum_final_kernel<<<(C + 127) / 128, 1024, 0, st>>>(partial, out, nblk, C);
  return check_launch("colsum_final");
}

extern "C" int a3t_scale_dropout(const float* x, void* y, int dtype_y, int64_t n, float scale, float drop_p,
                                 const unsigned long long* seed, uint32_t site, void* stream) {
  A3T_REQUIRE(x && y, "scale_dropout: null pointer");
  A3T_REQUIRE(drop_p == 0.f || seed, "scale_dropout: dropout needs a seed");
  if (n == 0) return A3T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_y == A3T_BF16)
    scale_dropout_kernel<__nv_bfloat16><<<ew_blocks(n), 256, 0, st>>>(x, (__nv_bfloat16*)y, n, scale, drop_p, seed, site);
  else
    scale_dropout_kernel<float><<<ew_blocks(n), 256, 0, st>>>(x, (float*)y, n, scale, drop_p, seed, site);
  return check_launch("scale_dropout");
}

extern "C" int a3t_mask_input_bwd(const float* dx, const uint8_t* masked, float* dmask_feature, float* partial,
                                  int64_t rows, int C, void* stream) {
  A3T_REQUIRE(dx && masked && dmask_feature && partial, "mask_input_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = colsum_blocks_host(rows);
  dim3 grid((C + 255) / 256, nblk);
  masked_colsum_kernel<<<grid, 256, 0, st>>>(dx, masked, partial, rows, C);
  int rc = check_launch("mask_input_bwd");
  if (rc) return rc;
  colsum_final_kernel<<<(C + 127) / 128, 1024, 0, st>>>(partial, dmask_feature, nblk, C);
  return check_launch("mask_input_bwd_final");
}

extern "C" int a3t_bn_stats(const float* z, float* mean, float* rstd, float* running_mean, float* running_var,
                            int64_t* num_batches_tracked, double* partial, int64_t rows, int C, float momentum,
                            float eps, int training, void* stream) {
  A3T_REQUIRE(z && mean && rstd, "bn_stats: null pointer");
  A3T_REQUIRE(training ? partial != nullptr : (running_mean && running_var), "bn_stats: missing buffers");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = colsum_blocks_host(rows);
  if (training) {
    dim3 grid((C + 255) / 256, nblk);
    bn_partial_kernel<<<grid, 256, 0, st>>>(z, partial, rows, C);
    int rc = check_launch("bn_partial");
    if (rc) return rc;
  }
  bn_final_kernel<<<(C + 127) / 128, 1024, 0, st>>>(partial, mean, rstd, running_mean, running_var,
                                                   num_batches_tracked, nblk, rows, C, momentum, eps, training);
  return check_launch("bn_final");
}

extern "C" int a3t_bn_act_fwd(const float* z, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const float* res, void* y, int dtype_y, int64_t rows, int C, int act,
                              float drop_p, const unsigned long long* seed, uint32_t site, void* stream) {
  A3T_REQUIRE(z && mean && rstd && gamma && beta && y, "bn_act_fwd: null pointer");
  A3T_REQUIRE(drop_p == 0.f || seed, "bn_act_fwd: dropout needs a seed");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n = rows * C;
  if (n == 0) return A3T_OK;
  if (dtype_y == A3T_BF16)
    bn_act_fwd_kernel<__nv_bfloat16><<<ew_blocks(n), 256, 0, st>>>(z, mean, rstd, gamma, beta, res, (__nv_bfloat16*)y,
                                                                   rows, C, act, drop_p, seed, site);
  else
    bn_act_fwd_kernel<float><<<ew_blocks(n), 256, 0, st>>>(z, mean, rstd, gamma, beta, res, (float*)y, rows, C, act,
                                                           drop_p, seed, site);
  return check_launch("bn_act_fwd");
}

extern "C" int a3t_bn_act_bwd(const float* dy, const float* z, const float* mean, const float* rstd,
                              const float* gamma, const float* beta, float* dz, float* dgamma, float* dbeta,
                              double* partial, float* coef, int64_t rows, int C, int act, int training, float drop_p,
                              const unsigned long long* seed, uint32_t site, void* stream) {
  A3T_REQUIRE(dy && z && mean && rstd && gamma && beta && dz && dgamma && dbeta && partial && coef,
              "bn_act_bwd: null pointer");
  A3T_REQUIRE(drop_p == 0.f || seed, "bn_act_bwd: dropout needs a seed");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = colsum_blocks_host(rows);
  dim3 grid((C + 255) / 256, nblk);
  bn_bwd_partial_kernel<<<grid, 256, 0, st>>>(dy, z, mean, rstd, gamma, beta, partial, rows, C, act, drop_p, seed, site);
  int rc = check_launch("bn_bwd_partial");
  if (rc) return rc;
  bn_bwd_final_kernel<<<(C + 127) / 128, 1024, 0, st>>>(partial, dgamma, dbeta, coef, nblk, C);
  rc = check_launch("bn_bwd_final");
  if (rc) return rc;
  int64_t n = rows * C;
  bn_bwd_dz_kernel<<<ew_blocks(n), 256, 0, st>>>(dy, z, mean, rstd, gamma, beta, coef, dz, rows, C, act, training,
                                                 drop_p, seed, site);
  return check_launch("bn_bwd_dz");
}
