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
  A3T_PDL_TRIGGER();
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
template <typename TDY, int MAXV, typename TG>
__global__ void __launch_bounds__(LN_WARPS * 32, 2) ln_bwd_kernel(
    const TDY* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_i,
    const float* __restrict__ rstd_i, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ dres, float* __restrict__ dx, float* __restrict__ partial, int64_t rows,
    int C, int relu, float out_scale, float drop_p, const unsigned long long* __restrict__ seed,
    uint32_t site, TG* __restrict__ gnext, float gn_scale, float gn_p, uint32_t gn_site,
    float* __restrict__ acc_dgamma, float* __restrict__ acc_dbeta, float* __restrict__ acc_gsum) {
  A3T_PDL_TRIGGER();
  extern __shared__ float sm[];  // [LN_WARPS][3][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = C >> 2;
  Drop dr = make_drop(drop_p, seed, site);
  Drop dn = make_drop(gnext ? gn_p : 0.f, seed, gn_site);
  const float gn_mul = gn_scale * dn.inv_keep;
  float4 ag[MAXV], ab[MAXV], an[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    ag[i] = make_float4(0, 0, 0, 0);
    ab[i] = make_float4(0, 0, 0, 0);
    an[i] = make_float4(0, 0, 0, 0);
  }
  const float scale_keep = out_scale * dr.inv_keep;
  // Row prefetch: while a warp works on row r, the x / dy / dres pieces of ITS next row are already in flight
  // (cp.async into a per-thread staging slot; every lane later reads back only what it copied itself, so
  // cp.async.wait_group is the only synchronisation).  Layout [slot][array][i][thread] x 16 bytes: conflict-free.
  // The staging area is reused as the reduction scratch after the row loop.
  const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(sm);
  constexpr uint32_t ARR_BYTES = MAXV * LN_WARPS * 32 * 16;   // one array of one slot
  constexpr uint32_t SLOT_BYTES = 3 * ARR_BYTES;              // x, dres, dy
  const uint32_t my = stage0 + threadIdx.x * 16;
  const int64_t rstride = (int64_t)gridDim.x * LN_WARPS;
  auto issue = [&](int64_t row, int slot) {
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      const int c4 = lane + i * 32;
      if (c4 < nv) {
        const uint32_t d0 = my + slot * SLOT_BYTES + i * (LN_WARPS * 32 * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(x + row * C + c4 * 4) : "memory");
        if (dres)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + ARR_BYTES), "l"(dres + row * C + c4 * 4) : "memory");
        if constexpr (sizeof(TDY) == 2)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0 + 2 * ARR_BYTES), "l"(dy + row * C + c4 * 4) : "memory");
        else
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 2 * ARR_BYTES), "l"(dy + row * C + c4 * 4) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp;
  int slot = 0;
  float mean_n = 0.f, rstd_n = 0.f;
  if (row < rows) {
    issue(row, 0);
    mean_n = mean_i[row];
    rstd_n = rstd_i[row];
  }
  for (; row < rows; row += rstride, slot ^= 1) {
    const float mean = mean_n, rstd = rstd_n;
    if (row + rstride < rows) {
      issue(row + rstride, slot ^ 1);
      mean_n = mean_i[row + rstride];
      rstd_n = rstd_i[row + rstride];
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    float xh[MAXV][4], dh[MAXV][4];
    float4 rres[MAXV];
    float s1 = 0.f, s2 = 0.f;
    const uint32_t sb = my + slot * SLOT_BYTES;
    if (dres) {
#pragma unroll
      for (int i = 0; i < MAXV; i++) {
        int c4 = lane + i * 32;
        if (c4 < nv)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(rres[i].x), "=f"(rres[i].y), "=f"(rres[i].z), "=f"(rres[i].w)
                       : "r"(sb + ARR_BYTES + i * (LN_WARPS * 32 * 16)));
      }
    }
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      int c4 = lane + i * 32;
      if (c4 < nv) {
        float4 xv;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(xv.x), "=f"(xv.y), "=f"(xv.z), "=f"(xv.w)
                     : "r"(sb + i * (LN_WARPS * 32 * 16)));
        const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma) + c4);  // L1-resident: not kept in registers
        const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
        const float ge[4] = {gv.x, gv.y, gv.z, gv.w};
        float g[4];
        const int64_t idx0 = row * C + c4 * 4;
        if constexpr (sizeof(TDY) == 2) {
          uint2 t;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(t.x), "=r"(t.y) : "r"(sb + 2 * ARR_BYTES + i * (LN_WARPS * 32 * 16)));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
          float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
          g[0] = f0.x; g[1] = f0.y; g[2] = f1.x; g[3] = f1.y;
        } else {
          float4 t;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                       : "r"(sb + 2 * ARR_BYTES + i * (LN_WARPS * 32 * 16)));
          g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
        }
        if (dr.on) {
          bool kp[4];
          drop_keep4(dr, drop_fold((unsigned long long)idx0), kp);
#pragma unroll
          for (int e = 0; e < 4; e++) g[e] = kp[e] ? g[e] : 0.f;
        }
        float be[4] = {0.f, 0.f, 0.f, 0.f};
        if (relu) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(beta) + c4);
          be[0] = bv.x; be[1] = bv.y; be[2] = bv.z; be[3] = bv.w;
        }
        float gacc[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          float ge_ = g[e] * scale_keep;
          float h = (xe[e] - mean) * rstd;
          if (relu && (h * ge[e] + be[e]) <= 0.f) ge_ = 0.f;
          xh[i][e] = h;
          gacc[e] = ge_ * h;
          g[e] = ge_;
          float d = ge_ * ge[e];
          dh[i][e] = d;
          s1 += d;
          s2 += d * h;
        }
        ag[i].x += gacc[0]; ag[i].y += gacc[1]; ag[i].z += gacc[2]; ag[i].w += gacc[3];
        ab[i].x += g[0]; ab[i].y += g[1]; ab[i].z += g[2]; ab[i].w += g[3];
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < MAXV; i++) {
      int c4 = lane + i * 32;
      if (c4 < nv) {
        float4 o;
        o.x = rstd * (dh[i][0] - s1 - xh[i][0] * s2);
        o.y = rstd * (dh[i][1] - s1 - xh[i][1] * s2);
        o.z = rstd * (dh[i][2] - s1 - xh[i][2] * s2);
        o.w = rstd * (dh[i][3] - s1 - xh[i][3] * s2);
        if (dres) {
          const float4 r = rres[i];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        reinterpret_cast<float4*>(dx + row * C)[c4] = o;
        if (gnext) {
          // the next section's "grad prep" g = dropout'(dx * scale) in its GEMM dtype, and the column sums
          // of g (that section's bias gradient), produced while dx is still in registers
          const int64_t idx0 = row * C + c4 * 4;
          float v[4] = {o.x, o.y, o.z, o.w};
          if (dn.on) {
            bool kp[4];
            drop_keep4(dn, drop_fold((unsigned long long)idx0), kp);
#pragma unroll
            for (int e = 0; e < 4; e++) v[e] = kp[e] ? v[e] * gn_mul : 0.f;
          } else {
#pragma unroll
            for (int e = 0; e < 4; e++) v[e] *= gn_mul;
          }
          if constexpr (sizeof(TG) == 2) {
            __nv_bfloat162 h[2] = {__floats2bfloat162_rn(v[0], v[1]), __floats2bfloat162_rn(v[2], v[3])};
            *reinterpret_cast<uint2*>(gnext + idx0) = *reinterpret_cast<const uint2*>(h);
            const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
            an[i].x += f0.x; an[i].y += f0.y; an[i].z += f1.x; an[i].w += f1.y;
          } else {
            *reinterpret_cast<float4*>(gnext + idx0) = make_float4(v[0], v[1], v[2], v[3]);
            an[i].x += v[0]; an[i].y += v[1]; an[i].z += v[2]; an[i].w += v[3];
          }
        }
      }
    }
  }
  // block reduce the parameter gradients (and the column sums of gnext); the staging slots are dead
  __syncthreads();
  const int nacc = gnext ? 3 : 2;
  float* sg = sm + (size_t)warp * 3 * C;
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    int c4 = lane + i * 32;
    if (c4 < nv) {
      reinterpret_cast<float4*>(sg)[c4] = ag[i];
      reinterpret_cast<float4*>(sg + C)[c4] = ab[i];
      reinterpret_cast<float4*>(sg + 2 * C)[c4] = an[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < nacc * C; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) t += sm[(size_t)w * 3 * C + c];
    if (partial) {  // two-stage: a second kernel sums the per-block rows (dgamma, dbeta only)
      if (c < 2 * C) partial[(size_t)blockIdx.x * 2 * C + c] = t;
    } else {        // accumulate into zero-initialised outputs
      float* dst = c < C ? acc_dgamma : (c < 2 * C ? acc_dbeta : acc_gsum);
      if (dst) atomicAdd(dst + (c % C), t);
    }
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
  A3T_PDL_TRIGGER();
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
  A3T_PDL_TRIGGER();
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
  A3T_PDL_TRIGGER();
  __shared__ float red[8][128];
  const int c = blockIdx.x * 128 + (threadIdx.x & 127);
  float t = reduce_rows_128x8<float>(partial, nblk, C, c, red);
  if ((threadIdx.x >> 7) == 0 && c < C) out[c] = t;
}

// Vectorised variant: a thread owns 4 (fp32) or 8 (bf16) consecutive columns = one 16-byte load per
// row; a block is 32 column groups x 8 row lanes (a warp reads 512 contiguous bytes of a row), rows
// unrolled by 4 for memory-level parallelism; the 8 row lanes are combined through shared memory.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, float* __restrict__ partial,
                                                         int64_t rows, int C, int64_t ldx,
                                                         float* __restrict__ atomic_out = nullptr) {
  A3T_PDL_TRIGGER();
  constexpr int N = Vec16<T>::N;
  __shared__ float red[8][32 * N + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * N;
  const int nblk = gridDim.y;
  const int64_t per = (rows + nblk - 1) / nblk;
  const int64_t r0 = (int64_t)blockIdx.y * per;
  int64_t r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  float acc[N];
#pragma unroll
  for (int i = 0; i < N; i++) acc[i] = 0.f;
  if (c < C) {
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      float v0[N], v1[N], v2[N], v3[N];
      Vec16<T>::load(x + r * ldx + c, v0);
      Vec16<T>::load(x + (r + 8) * ldx + c, v1);
      Vec16<T>::load(x + (r + 16) * ldx + c, v2);
      Vec16<T>::load(x + (r + 24) * ldx + c, v3);
#pragma unroll
      for (int i = 0; i < N; i++) acc[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
    }
    for (; r < r1; r += 8) {
      float v0[N];
      Vec16<T>::load(x + r * ldx + c, v0);
#pragma unroll
      for (int i = 0; i < N; i++) acc[i] += v0[i];
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++) red[ty][tx * N + i] = acc[i];
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * N; idx += 256) {
    int cc = blockIdx.x * 32 * N + idx;
    if (cc < C) {
      float t = 0.f;
#pragma unroll
      for (int y = 0; y < 8; y++) t += red[y][idx];
      if (atomic_out) atomicAdd(atomic_out + cc, t);
      else partial[(size_t)blockIdx.y * C + cc] = t;
    }
  }
}

// masked column sums for NewMaskInputLayer backward (mlm_encoder.py:67-70)
__global__ void __launch_bounds__(256) masked_colsum_kernel(const float* __restrict__ x,
                                                            const uint8_t* __restrict__ masked,
                                                            float* __restrict__ partial, int64_t rows, int C) {
  A3T_PDL_TRIGGER();
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
__device__ __forceinline__ void store4(TY* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 h[2] = {__floats2bfloat162_rn(v[0], v[1]), __floats2bfloat162_rn(v[2], v[3])};
  *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(h);
}
// keep-mask of 4 consecutive elements starting at idx0 (idx0 % 4 == 0)
__device__ __forceinline__ void drop_apply4(const Drop& dr, unsigned long long idx0, float (&v)[4]) {
  if (!dr.on) return;
  bool kp[4];
  drop_keep4(dr, drop_fold(idx0), kp);
#pragma unroll
  for (int j = 0; j < 4; j++) v[j] = kp[j] ? v[j] * dr.inv_keep : 0.f;
}

template <typename TY>
__global__ void __launch_bounds__(256) scale_dropout_kernel(const float* __restrict__ x, TY* __restrict__ y,
                                                            int64_t n, float scale, float drop_p,
                                                            const unsigned long long* __restrict__ seed,
                                                            uint32_t site) {
  A3T_PDL_TRIGGER();
  Drop dr = make_drop(drop_p, seed, site);
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 t = reinterpret_cast<const float4*>(x)[i];
    float v[4] = {t.x * scale, t.y * scale, t.z * scale, t.w * scale};
    drop_apply4(dr, (unsigned long long)i * 4ull, v);
    store4<TY>(y + i * 4, v);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = from_f32<TY>(drop_apply(dr, (unsigned long long)i, x[i] * scale));
}

// ---------------------------------------------------------------------------------------------
// BatchNorm1d statistics: double column sums of z and z^2
// ---------------------------------------------------------------------------------------------
// block = 32 column groups (4 channels each) x 8 row lanes; fp32 partial sums over <= 16 rows are
// promoted to double before they are combined (the reference accumulates in fp32; double keeps the
// batch statistics independent of the blocking)
__global__ void __launch_bounds__(256) bn_partial_kernel(const float* __restrict__ z, double* __restrict__ partial,
                                                         int64_t rows, int C) {
  A3T_PDL_TRIGGER();
  __shared__ double red[8][2][128 + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 4;
  const int nblk = gridDim.y;
  const int64_t per = (rows + nblk - 1) / nblk;
  const int64_t r0 = (int64_t)blockIdx.y * per;
  int64_t r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (c < C) {
    for (int64_t rb = r0 + ty; rb < r1; rb += 8 * 16) {
      float fs[4] = {0, 0, 0, 0}, fq[4] = {0, 0, 0, 0};
#pragma unroll 4
      for (int k = 0; k < 16; k++) {
        int64_t r = rb + (int64_t)k * 8;
        if (r < r1) {
          float4 t = *reinterpret_cast<const float4*>(z + r * C + c);
          fs[0] += t.x; fs[1] += t.y; fs[2] += t.z; fs[3] += t.w;
          fq[0] = fmaf(t.x, t.x, fq[0]); fq[1] = fmaf(t.y, t.y, fq[1]);
          fq[2] = fmaf(t.z, t.z, fq[2]); fq[3] = fmaf(t.w, t.w, fq[3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; i++) { s[i] += (double)fs[i]; q[i] += (double)fq[i]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) { red[ty][0][tx * 4 + i] = s[i]; red[ty][1][tx * 4 + i] = q[i]; }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * 128; idx += 256) {
    int which = idx >> 7, cc = idx & 127;
    int col = blockIdx.x * 128 + cc;
    if (col < C) {
      double t = 0;
#pragma unroll
      for (int y = 0; y < 8; y++) t += red[y][which][cc];
      partial[((size_t)blockIdx.y * 2 + which) * C + col] = t;
    }
  }
}
__global__ void __launch_bounds__(1024) bn_final_kernel(const double* __restrict__ partial, float* __restrict__ mean_o,
                                                        float* __restrict__ rstd_o, float* __restrict__ running_mean,
                                                        float* __restrict__ running_var, int64_t* __restrict__ nbt,
                                                        int nblk, int64_t rows, int C, float momentum, float eps,
                                                        int training) {
  A3T_PDL_TRIGGER();
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

// swish through ex2.approx / rcp.approx: relative error ~1e-6, two orders below the parity gates (1e-4)
__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == A3T_ACT_SWISH) return __fdividef(v, 1.f + __expf(-v));
  if (act == A3T_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float act_grad(float v, int act) {
  if (act == A3T_ACT_SWISH) {
    float s = __fdividef(1.f, 1.f + __expf(-v));
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
  A3T_PDL_TRIGGER();
  Drop dr = make_drop(drop_p, seed, site);
  const int C4 = C >> 2;
  const int64_t n4 = rows * C4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = (int)(i % C4) * 4;
    const float4 zz = reinterpret_cast<const float4*>(z)[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float v[4] = {(zz.x - mu.x) * rs.x * g.x + b.x, (zz.y - mu.y) * rs.y * g.y + b.y,
                  (zz.z - mu.z) * rs.z * g.z + b.z, (zz.w - mu.w) * rs.w * g.w + b.w};
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = act_fwd(v[j], act);
    drop_apply4(dr, (unsigned long long)i * 4ull, v);
    if (res) {
      const float4 r = reinterpret_cast<const float4*>(res)[i];
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    store4<TY>(y + i * 4, v);
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
  A3T_PDL_TRIGGER();
  __shared__ double red[8][2][128 + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 4;
  Drop dr = make_drop(drop_p, seed, site);
  const int nblk = gridDim.y;
  const int64_t per = (rows + nblk - 1) / nblk;
  const int64_t r0 = (int64_t)blockIdx.y * per;
  int64_t r1 = r0 + per;
  if (r1 > rows) r1 = rows;
  double sb[4] = {0, 0, 0, 0}, sg[4] = {0, 0, 0, 0};
  if (c < C) {
    const float4 mu4 = *reinterpret_cast<const float4*>(mean + c), rs4 = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + c), b4 = *reinterpret_cast<const float4*>(beta + c);
    const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    for (int64_t rb = r0 + ty; rb < r1; rb += 8 * 16) {
      float fb[4] = {0, 0, 0, 0}, fg[4] = {0, 0, 0, 0};
#pragma unroll 2
      for (int k = 0; k < 16; k++) {
        int64_t r = rb + (int64_t)k * 8;
        if (r < r1) {
          const int64_t i = r * C + c;
          const float4 z4 = *reinterpret_cast<const float4*>(z + i);
          const float4 d4 = *reinterpret_cast<const float4*>(dy + i);
          const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
          float d[4] = {d4.x, d4.y, d4.z, d4.w};
          drop_apply4(dr, (unsigned long long)i, d);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            float zh = (zz[j] - mu[j]) * rs[j];
            float dd = d[j] * act_grad(zh * g[j] + b[j], act);
            fb[j] += dd;
            fg[j] = fmaf(dd, zh, fg[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++) { sb[j] += (double)fb[j]; sg[j] += (double)fg[j]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) { red[ty][0][tx * 4 + i] = sb[i]; red[ty][1][tx * 4 + i] = sg[i]; }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * 128; idx += 256) {
    int which = idx >> 7, cc = idx & 127;
    int col = blockIdx.x * 128 + cc;
    if (col < C) {
      double t = 0;
#pragma unroll
      for (int y = 0; y < 8; y++) t += red[y][which][cc];
      partial[((size_t)blockIdx.y * 2 + which) * C + col] = t;
    }
  }
}
__global__ void __launch_bounds__(1024) bn_bwd_final_kernel(const double* __restrict__ partial,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ coef, int nblk, int C) {
  A3T_PDL_TRIGGER();
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
  A3T_PDL_TRIGGER();
  Drop dr = make_drop(drop_p, seed, site);
  const int C4 = C >> 2;
  const int64_t n4 = rows * C4;
  const float inv_n = 1.f / (float)rows;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int c = (int)(i % C4) * 4;
    const float4 z4 = reinterpret_cast<const float4*>(z)[i], d4 = reinterpret_cast<const float4*>(dy)[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    const float4 k0 = *reinterpret_cast<const float4*>(coef + c), k1 = *reinterpret_cast<const float4*>(coef + C + c);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, mm[4] = {mu.x, mu.y, mu.z, mu.w}, rr[4] = {rs.x, rs.y, rs.z, rs.w};
    const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
    const float c0[4] = {k0.x, k0.y, k0.z, k0.w}, c1[4] = {k1.x, k1.y, k1.z, k1.w};
    float d[4] = {d4.x, d4.y, d4.z, d4.w}, o[4];
    drop_apply4(dr, (unsigned long long)i * 4ull, d);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float zh = (zz[j] - mm[j]) * rr[j];
      float dd = d[j] * act_grad(zh * gg[j] + bb[j], act);
      o[j] = training ? gg[j] * rr[j] * (dd - c0[j] * inv_n - zh * c1[j] * inv_n) : gg[j] * rr[j] * dd;
    }
    reinterpret_cast<float4*>(dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
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
  if (b > 148 * 2) b = 148 * 2;  // 2 resident blocks per SM (128 registers per thread)
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int a3t_layernorm_bwd(const void* dy, int dtype_dy, const float* x, const float* mean, const float* rstd,
                                 const float* gamma, const float* beta, const float* dres, float* dx, float* dgamma,
                                 float* dbeta, float* partial, int64_t rows, int C, int relu, float out_scale,
                                 float drop_p, const unsigned long long* seed, uint32_t site, void* gnext,
                                 int dtype_gnext, float gnext_scale, float gnext_drop_p, uint32_t gnext_site,
                                 float* gsum, void* stream) {
  A3T_REQUIRE(dy && x && mean && rstd && gamma && beta && dx, "layernorm_bwd: null pointer");
  A3T_REQUIRE(C % 4 == 0 && C <= 128 * LN_MAXV, "layernorm_bwd: C=%d must be a multiple of 4 and <= %d", C, 128 * LN_MAXV);
  A3T_REQUIRE(drop_p == 0.f || seed, "layernorm_bwd: dropout needs a seed");
  A3T_REQUIRE(!gnext || gnext_drop_p == 0.f || seed, "layernorm_bwd: dropout needs a seed");
  A3T_REQUIRE(!(gsum && partial), "layernorm_bwd: gsum needs the accumulate mode (partial == NULL, zero-initialised outputs)");
  A3T_REQUIRE(!gsum || gnext, "layernorm_bwd: gsum without gnext");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = a3t_layernorm_bwd_blocks(rows);
  // two prefetch slots of 3 arrays x MAXV x 256 threads x 16 B, reused as [LN_WARPS][3][C] reduction scratch
  const int mv = C <= 384 ? 3 : 4;
  size_t smem = (size_t)2 * 3 * mv * LN_WARPS * 32 * 16;
  if (smem < (size_t)LN_WARPS * 3 * C * sizeof(float)) smem = (size_t)LN_WARPS * 3 * C * sizeof(float);
#define A3T_LN_BWD2(T, MV, TG)                                                                                       \
  {                                                                                                                  \
    static int smem_set = 0; /* per instantiation: raise the dynamic shared-memory limit once */                     \
    if (smem_set < (int)smem) {                                                                                      \
      cudaFuncSetAttribute(ln_bwd_kernel<T, MV, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
      smem_set = (int)smem;                                                                                          \
    }                                                                                                                \
    ln_bwd_kernel<T, MV, TG><<<nblk, LN_WARPS * 32, smem, st>>>((const T*)dy, x, mean, rstd, gamma, beta, dres, dx,   \
                                                                partial, rows, C, relu, out_scale, drop_p, seed, site, \
                                                                (TG*)gnext, gnext_scale, gnext_drop_p, gnext_site,   \
                                                                dgamma, dbeta, gsum);                                 \
  }
#define A3T_LN_BWD(T, MV)                                        \
  do {                                                           \
    if (gnext && dtype_gnext == A3T_BF16) A3T_LN_BWD2(T, MV, __nv_bfloat16) \
    else A3T_LN_BWD2(T, MV, float)                               \
  } while (0)
  if (dtype_dy == A3T_BF16) {
    if (C <= 384) A3T_LN_BWD(__nv_bfloat16, 3);
    else A3T_LN_BWD(__nv_bfloat16, 4);
  } else {
    if (C <= 384) A3T_LN_BWD(float, 3);
    else A3T_LN_BWD(float, 4);
  }
  int rc = check_launch("layernorm_bwd");
  if (rc) return rc;
  if (partial && (dgamma || dbeta)) {
    reduce_partial_kernel<<<(2 * C + 127) / 128, 1024, 0, st>>>(partial, dgamma, dbeta, nblk, C);
    rc = check_launch("layernorm_bwd_reduce");
  }
  return rc;
}

extern "C" int a3t_colsum_blocks(int64_t rows) { return colsum_blocks_host(rows); }

extern "C" int a3t_colsum(const void* x, int dtype_x, float* out, float* partial, int64_t rows, int C, int64_t ldx,
                          void* stream) {
  A3T_REQUIRE(x && out && C > 0, "colsum: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = colsum_blocks_host(rows);
  dim3 grid((C + 255) / 256, nblk);
  const int vn = dtype_x == A3T_BF16 ? 8 : 4;
  const bool vec = C % vn == 0 && ldx % vn == 0 && ((uintptr_t)x & 15) == 0;
  A3T_REQUIRE(partial || vec, "colsum: the accumulate mode (partial == NULL) needs 16-byte aligned rows");
  if (vec) {
    // partial == NULL: one kernel, block sums accumulated into the zero-initialised `out` with atomics
    if (!partial && nblk > 296) nblk = 296;
    dim3 gv((C / vn + 31) / 32, nblk);
    float* acc = partial ? nullptr : out;
    if (dtype_x == A3T_BF16)
      colsum_vec_kernel<__nv_bfloat16><<<gv, 256, 0, st>>>((const __nv_bfloat16*)x, partial, rows, C, ldx, acc);
    else
      colsum_vec_kernel<float><<<gv, 256, 0, st>>>((const float*)x, partial, rows, C, ldx, acc);
    if (!partial) return check_launch("colsum");
  } else if (dtype_x == A3T_BF16)
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
  A3T_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "scale_dropout: buffers must be 16-byte aligned");
  A3T_REQUIRE(drop_p == 0.f || seed, "scale_dropout: dropout needs a seed");
  if (n == 0) return A3T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_y == A3T_BF16)
    scale_dropout_kernel<__nv_bfloat16><<<ew_blocks((n + 3) / 4), 256, 0, st>>>(x, (__nv_bfloat16*)y, n, scale, drop_p, seed, site);
  else
    scale_dropout_kernel<float><<<ew_blocks((n + 3) / 4), 256, 0, st>>>(x, (float*)y, n, scale, drop_p, seed, site);
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
  A3T_REQUIRE(C % 4 == 0, "bn_stats: C=%d must be a multiple of 4", C);
  A3T_REQUIRE(training ? partial != nullptr : (running_mean && running_var), "bn_stats: missing buffers");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = colsum_blocks_host(rows);
  if (training) {
    dim3 grid((C + 127) / 128, nblk);
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
  A3T_REQUIRE(C % 4 == 0, "bn_act_fwd: C=%d must be a multiple of 4", C);
  A3T_REQUIRE(drop_p == 0.f || seed, "bn_act_fwd: dropout needs a seed");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n = rows * C;
  if (n == 0) return A3T_OK;
  if (dtype_y == A3T_BF16)
    bn_act_fwd_kernel<__nv_bfloat16><<<ew_blocks(n / 4), 256, 0, st>>>(z, mean, rstd, gamma, beta, res, (__nv_bfloat16*)y,
                                                                   rows, C, act, drop_p, seed, site);
  else
    bn_act_fwd_kernel<float><<<ew_blocks(n / 4), 256, 0, st>>>(z, mean, rstd, gamma, beta, res, (float*)y, rows, C, act,
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
  A3T_REQUIRE(C % 4 == 0, "bn_act_bwd: C=%d must be a multiple of 4", C);
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = colsum_blocks_host(rows);
  dim3 grid((C + 127) / 128, nblk);
  bn_bwd_partial_kernel<<<grid, 256, 0, st>>>(dy, z, mean, rstd, gamma, beta, partial, rows, C, act, drop_p, seed, site);
  int rc = check_launch("bn_bwd_partial");
  if (rc) return rc;
  bn_bwd_final_kernel<<<(C + 127) / 128, 1024, 0, st>>>(partial, dgamma, dbeta, coef, nblk, C);
  rc = check_launch("bn_bwd_final");
  if (rc) return rc;
  int64_t n = rows * C;
  bn_bwd_dz_kernel<<<ew_blocks(n / 4), 256, 0, st>>>(dy, z, mean, rstd, gamma, beta, coef, dz, rows, C, act, training,
                                                 drop_p, seed, site);
  return check_launch("bn_bwd_dz");
}
