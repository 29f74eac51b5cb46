// Legacy relative-position masked softmax, forward and backward
// (transformer/attention.py:145-165 rel_shift, :205-207 scale, :79-88 mask/softmax/zero/dropout).
// One warp per (b,h,i) score row; the row lives in shared memory.
#include <float.h>
#include <stdlib.h>
#include "common.cuh"

namespace a3t {

constexpr int SM_WARPS = 4;

// rel_shift gather: padded view (S, S+1) with a zero first column, reinterpreted as (S+1, S),
// first row dropped.  shifted[i,j] = padded_flat[(i+1)*S + j].
// `ld` = row pitch of the (S, S) score matrices in elements (>= S; a multiple of 8 keeps every row 16-byte
// aligned for the TMA tensor maps of the contractions when S itself is not)
__device__ __forceinline__ float bd_shifted(const void* __restrict__ bd, int dt, int64_t mat, int S, int64_t ld, int i, int j) {
  int64_t f = (int64_t)(i + 1) * S + j;
  int r = (int)(f / (S + 1));
  int c = (int)(f - (int64_t)r * (S + 1));
  return c == 0 ? 0.f : load_as_f32(bd, dt, mat + (int64_t)r * ld + (c - 1));
}

template <typename TP>
__global__ void __launch_bounds__(SM_WARPS * 32) relpos_softmax_fwd_kernel(
    const void* __restrict__ ac, const void* __restrict__ bd_raw, int dt_in, const uint8_t* __restrict__ keymask,
    TP* __restrict__ P, TP* __restrict__ Pd, int B, int H, int S, int64_t ld, float scale, float drop_p,
    const unsigned long long* __restrict__ seed, uint32_t site) {
  A3T_PDL_TRIGGER();
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* row = sm + (size_t)warp * S;
  Drop dr = make_drop(drop_p, seed, site);
  const int64_t nrows = (int64_t)B * H * S;
  for (int64_t r = (int64_t)blockIdx.x * SM_WARPS + warp; r < nrows; r += (int64_t)gridDim.x * SM_WARPS) {
    const int i = (int)(r % S);
    const int64_t bh = r / S;
    const int b = (int)(bh / H);
    const uint8_t* km = keymask + (int64_t)b * S;
    float mx = -FLT_MAX;
    for (int j = lane; j < S; j += 32) {
      float s = (load_as_f32(ac, dt_in, r * ld + j) + bd_shifted(bd_raw, dt_in, bh * (int64_t)S * ld, S, ld, i, j)) * scale;
      if (!km[j]) s = -FLT_MAX;  // finfo(float32).min
      row[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      float e = expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < S; j += 32) {
      float p = km[j] ? row[j] * inv : 0.f;
      P[r * ld + j] = from_f32<TP>(p);
      if (dr.on) Pd[r * ld + j] = from_f32<TP>(drop_apply(dr, (unsigned long long)(r * S + j), p));
      else if (Pd != P) Pd[r * ld + j] = from_f32<TP>(p);
    }
    __syncwarp();
  }
}

// dS[i,j] = P * (dPu - sum_j dPu*P) * scale ; dPu = dPd*keep/(1-p)
template <typename TP, typename TO>
__global__ void __launch_bounds__(SM_WARPS * 32) relpos_softmax_bwd_kernel(
    const void* __restrict__ dPd, int dt_in, const TP* __restrict__ P, TO* __restrict__ dS, int64_t nrows, int S, int64_t ld,
    float scale, float drop_p, const unsigned long long* __restrict__ seed, uint32_t site) {
  A3T_PDL_TRIGGER();
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* row = sm + (size_t)warp * S;
  Drop dr = make_drop(drop_p, seed, site);
  for (int64_t r = (int64_t)blockIdx.x * SM_WARPS + warp; r < nrows; r += (int64_t)gridDim.x * SM_WARPS) {
    float dot = 0.f;
    for (int j = lane; j < S; j += 32) {
      float g = load_as_f32(dPd, dt_in, r * ld + j);
      if (dr.on) g = drop_keep(dr, (unsigned long long)(r * S + j)) ? g * dr.inv_keep : 0.f;
      row[j] = g;
      dot += g * to_f32<TP>(P[r * ld + j]);
    }
    dot = warp_sum(dot);
    for (int j = lane; j < S; j += 32) {
      float p = to_f32<TP>(P[r * ld + j]);
      dS[r * ld + j] = from_f32<TO>(p * (row[j] - dot) * scale);
    }
    __syncwarp();
  }
}

// dBD_raw[r, cc] = dS at the inverse rel_shift position (0 where BD_raw is never read)
template <typename TO>
__global__ void __launch_bounds__(256) relshift_bwd_kernel(const TO* __restrict__ dS, TO* __restrict__ dBD, int64_t nmat,
                                                           int S, int64_t ld) {
  A3T_PDL_TRIGGER();
  const int64_t per = (int64_t)S * S;
  const int64_t n = nmat * per;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    int64_t m = idx / per;
    int64_t e = idx - m * per;
    int r = (int)(e / S), cc = (int)(e - (int64_t)r * S);
    int64_t f = (int64_t)r * (S + 1) + cc + 1 - S;  // flat index into the dense (S,S) shifted matrix
    const int64_t fr = f / S, fc = f - fr * S;
    dBD[m * S * ld + (int64_t)r * ld + cc] = f >= 0 ? dS[m * S * ld + fr * ld + fc] : from_f32<TO>(0.f);
  }
}


// ---------------------------------------------------------------------------------------------
// Vectorised kernels for S % 4 == 0.  A lane owns 4 consecutive keys per step (float4 loads of AC /
// dPd, 8-byte bf16 stores).  rel_shift without integer division: for query row i
//   shifted[i, j] = BD_raw[i, S-1-i+j] (j <= i) | 0 (j == i+1) | BD_raw[i+1, j-i-2] (j >= i+2)
// which is transformer/attention.py:155-159 written out (pad one zero column, view (S+1, S), drop row 0).
// ---------------------------------------------------------------------------------------------
template <typename TP>
__device__ __forceinline__ void store_p4(TP* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store_p4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store_p4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 h[2] = {__floats2bfloat162_rn(a, b), __floats2bfloat162_rn(c, d)};
  *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(h);
}
template <typename TP>
__device__ __forceinline__ void load_p4(const TP* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load_p4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load_p4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename TP, typename TI>
__global__ void __launch_bounds__(SM_WARPS * 32) relpos_softmax_fwd_v4_kernel(
    const TI* __restrict__ ac, const TI* __restrict__ bd_raw, const uint8_t* __restrict__ keymask,
    TP* __restrict__ P, TP* __restrict__ Pd, int B, int H, int S, int64_t ld, float scale, float drop_p,
    const unsigned long long* __restrict__ seed, uint32_t site) {
  A3T_PDL_TRIGGER();
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* row = sm + (size_t)warp * S;
  Drop dr = make_drop(drop_p, seed, site);
  const int64_t nrows = (int64_t)B * H * S;
  for (int64_t r = (int64_t)blockIdx.x * SM_WARPS + warp; r < nrows; r += (int64_t)gridDim.x * SM_WARPS) {
    const int i = (int)(r % S);
    const int64_t bh = r / S;
    const int b = (int)(bh / H);
    const TI* acr = ac + r * ld;
    const TI* bd0 = bd_raw + (bh * S + i) * ld + (S - 1 - i);   // + j       for j <= i
    const TI* bd1 = bd_raw + (bh * S + i + 1) * ld - (i + 2);   // + j       for j >= i+2
    const uint8_t* km = keymask + (int64_t)b * S;
    float mx = -FLT_MAX;
    for (int j = lane * 4; j < S; j += 128) {
      float a[4];
      load_p4<TI>(acr + j, a);
      const uchar4 k4 = *reinterpret_cast<const uchar4*>(km + j);
      const unsigned char kk[4] = {k4.x, k4.y, k4.z, k4.w};
      float sv[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int jj = j + e;
        float bd = 0.f;
        if (jj <= i) bd = to_f32<TI>(bd0[jj]);
        else if (jj >= i + 2) bd = to_f32<TI>(bd1[jj]);
        float v = (a[e] + bd) * scale;
        if (!kk[e]) v = -FLT_MAX;  // finfo(float32).min
        sv[e] = v;
        mx = fmaxf(mx, v);
      }
      *reinterpret_cast<float4*>(row + j) = make_float4(sv[0], sv[1], sv[2], sv[3]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane * 4; j < S; j += 128) {
      float4 t = *reinterpret_cast<float4*>(row + j);
      t.x = __expf(t.x - mx); t.y = __expf(t.y - mx); t.z = __expf(t.z - mx); t.w = __expf(t.w - mx);
      sum += (t.x + t.y) + (t.z + t.w);
      *reinterpret_cast<float4*>(row + j) = t;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane * 4; j < S; j += 128) {
      const float4 t = *reinterpret_cast<float4*>(row + j);
      const uchar4 k4 = *reinterpret_cast<const uchar4*>(km + j);
      float pv[4] = {k4.x ? t.x * inv : 0.f, k4.y ? t.y * inv : 0.f, k4.z ? t.z * inv : 0.f, k4.w ? t.w * inv : 0.f};
      store_p4<TP>(P + r * ld + j, pv[0], pv[1], pv[2], pv[3]);
      if (dr.on) {
        const unsigned long long idx0 = (unsigned long long)(r * S + j);
        bool kp[4];
        drop_keep4(dr, drop_fold(idx0), kp);
#pragma unroll
        for (int e = 0; e < 4; e++) pv[e] = kp[e] ? pv[e] * dr.inv_keep : 0.f;
        store_p4<TP>(Pd + r * ld + j, pv[0], pv[1], pv[2], pv[3]);
      } else if (Pd != P) {
        store_p4<TP>(Pd + r * ld + j, pv[0], pv[1], pv[2], pv[3]);
      }
    }
    __syncwarp();
  }
}

// dS = P * (dPu - sum_j dPu*P) * scale, and dBD_raw = inverse rel_shift of dS written by the same warp:
//   row i of dS feeds dBD[i, S-1-i+j] (j <= i) and dBD[i+1, j-i-2] (j >= i+2); row 0 of dBD is zero
//   except its last element.  Every dBD element is written exactly once.
template <typename TP, typename TO, typename TI>
__global__ void __launch_bounds__(SM_WARPS * 32) relpos_softmax_bwd_v4_kernel(
    const TI* __restrict__ dPd, const TP* __restrict__ P, TO* __restrict__ dS, TO* __restrict__ dBD, int64_t nrows,
    int S, int64_t ld, float scale, float drop_p, const unsigned long long* __restrict__ seed, uint32_t site) {
  A3T_PDL_TRIGGER();
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* row = sm + (size_t)warp * S;
  Drop dr = make_drop(drop_p, seed, site);
  for (int64_t r = (int64_t)blockIdx.x * SM_WARPS + warp; r < nrows; r += (int64_t)gridDim.x * SM_WARPS) {
    const int i = (int)(r % S);
    float dot = 0.f;
    for (int j = lane * 4; j < S; j += 128) {
      float g[4];
      load_p4<TI>(dPd + r * ld + j, g);
      float pv[4];
      load_p4<TP>(P + r * ld + j, pv);
      if (dr.on) {
        const unsigned long long idx0 = (unsigned long long)(r * S + j);
        bool kp[4];
        drop_keep4(dr, drop_fold(idx0), kp);
#pragma unroll
        for (int e = 0; e < 4; e++) g[e] = kp[e] ? g[e] * dr.inv_keep : 0.f;
      }
      dot += (g[0] * pv[0] + g[1] * pv[1]) + (g[2] * pv[2] + g[3] * pv[3]);
      *reinterpret_cast<float4*>(row + j) = make_float4(g[0], g[1], g[2], g[3]);
    }
    dot = warp_sum(dot);
    TO* d0 = dBD + r * ld + (S - 1 - i);         // + j   for j <= i
    TO* d1 = dBD + (r + 1) * ld - (i + 2);       // + j   for j >= i+2 (row i+1 of the same matrix)
    for (int j = lane * 4; j < S; j += 128) {
      float pv[4];
      load_p4<TP>(P + r * ld + j, pv);
      const float4 g4 = *reinterpret_cast<float4*>(row + j);
      const float o[4] = {pv[0] * (g4.x - dot) * scale, pv[1] * (g4.y - dot) * scale, pv[2] * (g4.z - dot) * scale,
                          pv[3] * (g4.w - dot) * scale};
      store_p4<TO>(dS + r * ld + j, o[0], o[1], o[2], o[3]);
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int jj = j + e;
        if (jj <= i) d0[jj] = from_f32<TO>(o[e]);
        else if (jj >= i + 2 && i + 1 < S) d1[jj] = from_f32<TO>(o[e]);
      }
    }
    if (i == 0) {  // BD_raw[0, 0..S-2] is never read by the forward
      for (int j = lane; j < S - 1; j += 32) dBD[r * ld + j] = from_f32<TO>(0.f);
    }
    __syncwarp();
  }
}


// ---------------------------------------------------------------------------------------------
// bf16 register-resident kernels (the tensor-core mode's score tensors), S % 4 == 0, S <= 128 * NIT.
// A warp owns a row; every lane issues ALL of the row's loads before the first use (the v4 kernels above
// stage the row in shared memory and wait on each 8-byte load in turn: they are latency- and
// instruction-bound at ~40 % of HBM bandwidth).  rel_shift as a contiguous slice: with G = BD_raw[b,h]
// flattened, shifted[i, j] = G[o + j] (j <= i), 0 (j == i+1), G[o + j - 1] (j >= i+2), o = (i+1)(S-1)
// (transformer/attention.py:155-159).  A lane's 4 keys therefore sit in ONE 16-byte aligned-pair window of
// G: two aligned 8-byte loads and a funnel shift replace four 2-byte loads with per-element branches.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bf16x4_to_f32(uint32_t w0, uint32_t w1, float (&v)[4]) {
  v[0] = __uint_as_float(w0 << 16); v[1] = __uint_as_float(w0 & 0xFFFF0000u);
  v[2] = __uint_as_float(w1 << 16); v[3] = __uint_as_float(w1 & 0xFFFF0000u);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NIT, bool FULL>  // FULL: S == 128 * NIT, no tail guards
__global__ void __launch_bounds__(256, NIT <= 9 ? 2 : 1) relpos_softmax_fwd_reg_kernel(
    const __nv_bfloat16* __restrict__ ac, const __nv_bfloat16* __restrict__ bd_raw, const uint8_t* __restrict__ keymask,
    __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ Pd, int B, int H, int S, int ld, float scale, float drop_p,
    const unsigned long long* __restrict__ seed, uint32_t site) {
  A3T_PDL_TRIGGER();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Drop dr = make_drop(drop_p, seed, site);
  const int64_t nrows = (int64_t)B * H * S;
  const int SS = S * ld;  // elements per score matrix (S <= 2048, ld = row pitch, ld % 4 == 0)
  const float c2 = scale * 1.4426950408889634f;  // softmax in base 2: exp(s - m) = 2^((s - m) log2 e)
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < nrows; r += (int64_t)gridDim.x * 8) {
    const int i = (int)(r % S);
    const int64_t bh = r / S;
    const int b = (int)(bh / H);
    const __nv_bfloat16* acr = ac + r * ld;
    __nv_bfloat16* const Pr = P + r * ld;
    __nv_bfloat16* const Pdr = Pd + r * ld;
    const unsigned long long ebase = (unsigned long long)(r * S);
    const __nv_bfloat16* G = bd_raw + bh * SS;
    const uint8_t* km = keymask + (int64_t)b * S;
    // windows: key j left of the gap reads G[ol + j] (row i of BD_raw), right of it G[og + j] (row i+1);
    // with a dense matrix (ld == S) og == ol - 1: one contiguous slice with the zero of key i+1 squeezed out
    const int ol = i * ld + (S - 1 - i), og = (i + 1) * ld - (i + 2);
    const int basel = ol & ~3, baser = og & ~3;
    const int dl = ol & 3, dg = og & 3;
    uint2 a[NIT], w01[NIT], w23[NIT];
    uint32_t kw[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      // groups past the end of the row (S < 128 * NIT) re-load the last group: no branch in the load phase,
      // their scores are forced to -inf below and their stores are skipped
      const int j = FULL ? lane * 4 + it * 128 : min(lane * 4 + it * 128, S - 4);
      {
        a[it] = __ldcs(reinterpret_cast<const uint2*>(acr + j));
        const int x4 = ((j + 3 <= i) ? basel : baser) + j;
        w01[it] = __ldg(reinterpret_cast<const uint2*>(G + x4));
        w23[it] = (x4 + 8 <= SS) ? __ldg(reinterpret_cast<const uint2*>(G + x4 + 4)) : make_uint2(0u, 0u);
        kw[it] = *reinterpret_cast<const uint32_t*>(km + j);  // consumed only after every load is in flight
      }
    }
    uint32_t notall = 0;  // bit it: some key of the group is padded
#pragma unroll
    for (int it = 0; it < NIT; it++)
      if (kw[it] != 0x01010101u) notall |= 1u << it;
    float v[NIT][4];
    float mx = -FLT_MAX;
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const bool ok = FULL || lane * 4 + it * 128 < S;
      const int j = FULL ? lane * 4 + it * 128 : min(lane * 4 + it * 128, S - 4);
      {
        float av[4], bv[4];
        bf16x4_to_f32(a[it].x, a[it].y, av);
        const bool left = j + 3 <= i;
        const int d = left ? dl : dg;
        if (left || j >= i + 2) {  // 4 consecutive elements of G, d elements into the window
          const uint32_t sh = (d & 1) * 16;
          const bool hi = (d & 2) != 0;
          const uint32_t q0 = hi ? w01[it].y : w01[it].x, q1 = hi ? w23[it].x : w01[it].y, q2 = hi ? w23[it].y : w23[it].x;
          bf16x4_to_f32(__funnelshift_r(q0, q1, sh), __funnelshift_r(q1, q2, sh), bv);
        } else {                   // the one group per row that holds the zero at key i+1: element loads
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int jj = j + e;
            bv[e] = jj == i + 1 ? 0.f : __bfloat162float(G[(jj <= i ? ol : og) + jj]);
          }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) v[it][e] = (av[e] + bv[e]) * c2;
        if (notall & (1u << it)) {
          const uchar4 k4 = *reinterpret_cast<const uchar4*>(km + j);
          if (!k4.x) v[it][0] = -FLT_MAX;  // finfo(float32).min
          if (!k4.y) v[it][1] = -FLT_MAX;
          if (!k4.z) v[it][2] = -FLT_MAX;
          if (!k4.w) v[it][3] = -FLT_MAX;
        }
        if (!ok) v[it][0] = v[it][1] = v[it][2] = v[it][3] = -FLT_MAX;  // duplicate of the last group: exp -> 0
        mx = fmaxf(mx, fmaxf(fmaxf(v[it][0], v[it][1]), fmaxf(v[it][2], v[it][3])));
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      {
#pragma unroll
        for (int e = 0; e < 4; e++) v[it][e] = ex2_approx(v[it][e] - mx);
        sum += (v[it][0] + v[it][1]) + (v[it][2] + v[it][3]);
      }
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    const float inv_dk = inv * dr.inv_keep;
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int j = lane * 4 + it * 128;
      if (FULL || j < S) {
        if (notall & (1u << it)) {  // padded keys get probability 0 (attention.py:86)
          const uchar4 k4 = *reinterpret_cast<const uchar4*>(km + j);
          if (!k4.x) v[it][0] = 0.f;
          if (!k4.y) v[it][1] = 0.f;
          if (!k4.z) v[it][2] = 0.f;
          if (!k4.w) v[it][3] = 0.f;
        }
        store_p4<__nv_bfloat16>(Pr + j, v[it][0] * inv, v[it][1] * inv, v[it][2] * inv, v[it][3] * inv);
        if (dr.on) {
          bool kp[4];
          drop_keep4(dr, drop_fold(ebase + (unsigned long long)j), kp);
          store_p4<__nv_bfloat16>(Pdr + j, kp[0] ? v[it][0] * inv_dk : 0.f, kp[1] ? v[it][1] * inv_dk : 0.f,
                                  kp[2] ? v[it][2] * inv_dk : 0.f, kp[3] ? v[it][3] * inv_dk : 0.f);
        } else if (Pd != P) {
          store_p4<__nv_bfloat16>(Pdr + j, v[it][0] * inv, v[it][1] * inv, v[it][2] * inv, v[it][3] * inv);
        }
      }
    }
  }
}

template <int NIT, bool FULL>
__global__ void __launch_bounds__(256, NIT <= 9 ? 2 : 1) relpos_softmax_bwd_reg_kernel(
    const __nv_bfloat16* __restrict__ dPd, const __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ dS,
    __nv_bfloat16* __restrict__ dBD, int64_t nrows, int S, int ld, float scale, float drop_p,
    const unsigned long long* __restrict__ seed, uint32_t site) {
  A3T_PDL_TRIGGER();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Drop dr = make_drop(drop_p, seed, site);
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < nrows; r += (int64_t)gridDim.x * 8) {
    const int i = (int)(r % S);
    uint2 gw[NIT], pw[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      // groups past the end of the row re-load the last group (no branch in the load phase); they are zeroed below
      const int j = FULL ? lane * 4 + it * 128 : min(lane * 4 + it * 128, S - 4);
      gw[it] = __ldcs(reinterpret_cast<const uint2*>(dPd + r * ld + j));
      pw[it] = __ldcs(reinterpret_cast<const uint2*>(P + r * ld + j));
      if (!FULL && lane * 4 + it * 128 >= S) pw[it] = make_uint2(0u, 0u);
    }
    float g[NIT][4];
    float dot = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int j = FULL ? lane * 4 + it * 128 : min(lane * 4 + it * 128, S - 4);
      {
        float pv[4];
        bf16x4_to_f32(gw[it].x, gw[it].y, g[it]);
        bf16x4_to_f32(pw[it].x, pw[it].y, pv);
        if (dr.on) {
          bool kp[4];
          drop_keep4(dr, drop_fold((unsigned long long)(r * S + j)), kp);
#pragma unroll
          for (int e = 0; e < 4; e++) g[it][e] = kp[e] ? g[it][e] * dr.inv_keep : 0.f;
        }
        dot += (g[it][0] * pv[0] + g[it][1] * pv[1]) + (g[it][2] * pv[2] + g[it][3] * pv[3]);
      }
    }
    dot = warp_sum(dot);
    // inverse rel_shift: keys j <= i go to row i of dBD_raw at column S-1-i+j, keys j >= i+2 to row i+1 at
    // column j-i-2 (key i+1 is dropped); contiguous when the matrix is dense
    __nv_bfloat16* const runl = dBD + r * ld + (S - 1 - i);      // + j
    __nv_bfloat16* const rung = dBD + (r + 1) * ld - (i + 2);    // + j
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int j = lane * 4 + it * 128;
      if (FULL || j < S) {
        float pv[4];
        bf16x4_to_f32(pw[it].x, pw[it].y, pv);
        __nv_bfloat162 h[2] = {__floats2bfloat162_rn(pv[0] * (g[it][0] - dot) * scale, pv[1] * (g[it][1] - dot) * scale),
                               __floats2bfloat162_rn(pv[2] * (g[it][2] - dot) * scale, pv[3] * (g[it][3] - dot) * scale)};
        *reinterpret_cast<uint2*>(dS + r * ld + j) = *reinterpret_cast<const uint2*>(h);
        const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(h);
        if (j + 3 <= i) {
#pragma unroll
          for (int e = 0; e < 4; e++) runl[j + e] = hv[e];
        } else if (j >= i + 2) {
          if (i + 1 < S) {
#pragma unroll
            for (int e = 0; e < 4; e++) rung[j + e] = hv[e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int jj = j + e;
            if (jj <= i) runl[jj] = hv[e];
            else if (jj >= i + 2 && i + 1 < S) rung[jj] = hv[e];
          }
        }
      }
    }
    if (i == 0) {  // BD_raw[0, 0..S-2] is never read by the forward
      for (int j = lane; j < S - 1; j += 32) dBD[r * ld + j] = __float2bfloat16_rn(0.f);
    }
  }
}

}  // namespace a3t

using namespace a3t;

extern "C" int a3t_relpos_softmax_fwd(const void* ac, const void* bd_raw, int dtype_in, const uint8_t* keymask, void* P,
                                      void* Pd, int dtype_p, int B, int H, int S, int ld, float scale, float drop_p,
                                      const unsigned long long* seed, uint32_t site, void* stream) {
  A3T_REQUIRE(ac && bd_raw && keymask && P && Pd, "relpos_softmax_fwd: null pointer");
  if (ld == 0) ld = S;
  A3T_REQUIRE(ld >= S, "relpos_softmax_fwd: row pitch %d < S=%d", ld, S);
  A3T_REQUIRE(drop_p == 0.f || (seed && Pd != P), "relpos_softmax_fwd: dropout needs a seed and a separate Pd");
  A3T_REQUIRE(S > 0 && S <= 12000, "relpos_softmax_fwd: S=%d out of range", S);
  A3T_REQUIRE(dtype_in == A3T_F32 || dtype_in == A3T_BF16, "relpos_softmax_fwd: bad input dtype");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t nrows = (int64_t)B * H * S;
  int blocks = (int)((nrows + SM_WARPS - 1) / SM_WARPS);
  if (blocks > 148 * 16) blocks = 148 * 16;
  size_t smem = (size_t)SM_WARPS * S * sizeof(float);
  const bool v4 = (S % 4) == 0 && (ld % 4) == 0 &&
                  ((((uintptr_t)ac | (uintptr_t)P | (uintptr_t)Pd | (uintptr_t)keymask) & 15) == 0);
  if (v4 && dtype_p == A3T_BF16 && dtype_in == A3T_BF16 && S <= 2048 && (((uintptr_t)bd_raw & 15) == 0) &&
      !tune_env("A3T_SOFTMAX_SMEM")) {
    int rb = (int)((nrows + 7) / 8);
    if (rb > 148 * 2 * 8) rb = 148 * 2 * 8;
#define A3T_SM_FWD_REG(NIT, FULL)                                                                               \
  relpos_softmax_fwd_reg_kernel<NIT, FULL><<<rb, 256, 0, st>>>((const __nv_bfloat16*)ac, (const __nv_bfloat16*)bd_raw, keymask, \
                                                         (__nv_bfloat16*)P, (__nv_bfloat16*)Pd, B, H, S, ld, scale, drop_p,  \
                                                         seed, site)
    if (S == 128 * 9) A3T_SM_FWD_REG(9, true);
    else if (S <= 128 * 5) A3T_SM_FWD_REG(5, false);
    else if (S <= 128 * 9) A3T_SM_FWD_REG(9, false);
    else if (S <= 128 * 14) A3T_SM_FWD_REG(14, false);
    else A3T_SM_FWD_REG(16, false);
    return check_launch("relpos_softmax_fwd");
  }
  if (v4 && smem <= 48 * 1024) {
#define A3T_SM_FWD(TPT, TIT)                                                                              \
  relpos_softmax_fwd_v4_kernel<TPT, TIT><<<blocks, SM_WARPS * 32, smem, st>>>(                            \
      (const TIT*)ac, (const TIT*)bd_raw, keymask, (TPT*)P, (TPT*)Pd, B, H, S, ld, scale, drop_p, seed, site)
    if (dtype_p == A3T_BF16 && dtype_in == A3T_BF16) A3T_SM_FWD(__nv_bfloat16, __nv_bfloat16);
    else if (dtype_p == A3T_BF16) A3T_SM_FWD(__nv_bfloat16, float);
    else if (dtype_in == A3T_BF16) A3T_SM_FWD(float, __nv_bfloat16);
    else A3T_SM_FWD(float, float);
    return check_launch("relpos_softmax_fwd");
  }
  if (dtype_p == A3T_BF16) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(relpos_softmax_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    relpos_softmax_fwd_kernel<__nv_bfloat16><<<blocks, SM_WARPS * 32, smem, st>>>(
        ac, bd_raw, dtype_in, keymask, (__nv_bfloat16*)P, (__nv_bfloat16*)Pd, B, H, S, ld, scale, drop_p, seed, site);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(relpos_softmax_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    relpos_softmax_fwd_kernel<float><<<blocks, SM_WARPS * 32, smem, st>>>(ac, bd_raw, dtype_in, keymask, (float*)P,
                                                                         (float*)Pd, B, H, S, ld, scale, drop_p, seed, site);
  }
  return check_launch("relpos_softmax_fwd");
}

template <typename TP, typename TO>
static int softmax_bwd_launch(const void* dPd, int dtype_in, const void* P, void* dS, void* dBD, int B, int H, int S, int ld,
                              float scale, float drop_p, const unsigned long long* seed, uint32_t site, cudaStream_t st) {
  int64_t nrows = (int64_t)B * H * S;
  int blocks = (int)((nrows + SM_WARPS - 1) / SM_WARPS);
  if (blocks > 148 * 16) blocks = 148 * 16;
  size_t smem = (size_t)SM_WARPS * S * sizeof(float);
  if ((S % 4) == 0 && (ld % 4) == 0 && smem <= 48 * 1024 && ((((uintptr_t)dPd | (uintptr_t)P | (uintptr_t)dS) & 15) == 0)) {
    if (dtype_in == A3T_BF16)
      relpos_softmax_bwd_v4_kernel<TP, TO, __nv_bfloat16><<<blocks, SM_WARPS * 32, smem, st>>>(
          (const __nv_bfloat16*)dPd, (const TP*)P, (TO*)dS, (TO*)dBD, nrows, S, ld, scale, drop_p, seed, site);
    else
      relpos_softmax_bwd_v4_kernel<TP, TO, float><<<blocks, SM_WARPS * 32, smem, st>>>(
          (const float*)dPd, (const TP*)P, (TO*)dS, (TO*)dBD, nrows, S, ld, scale, drop_p, seed, site);
    return check_launch("relpos_softmax_bwd");
  }
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(relpos_softmax_bwd_kernel<TP, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  relpos_softmax_bwd_kernel<TP, TO><<<blocks, SM_WARPS * 32, smem, st>>>(dPd, dtype_in, (const TP*)P, (TO*)dS, nrows, S, ld,
                                                                        scale, drop_p, seed, site);
  int rc = check_launch("relpos_softmax_bwd");
  if (rc) return rc;
  int64_t n = nrows * S;
  int64_t b2 = (n + 255) / 256;
  if (b2 > 148 * 16) b2 = 148 * 16;
  relshift_bwd_kernel<TO><<<(int)b2, 256, 0, st>>>((const TO*)dS, (TO*)dBD, (int64_t)B * H, S, ld);
  return check_launch("relshift_bwd");
}

extern "C" int a3t_relpos_softmax_bwd(const void* dPd, int dtype_in, const void* P, int dtype_p, void* dS, void* dBD,
                                      int dtype_o, int B, int H, int S, int ld, float scale, float drop_p,
                                      const unsigned long long* seed, uint32_t site, void* stream) {
  A3T_REQUIRE(dPd && P && dS && dBD, "relpos_softmax_bwd: null pointer");
  if (ld == 0) ld = S;
  A3T_REQUIRE(ld >= S, "relpos_softmax_bwd: row pitch %d < S=%d", ld, S);
  A3T_REQUIRE(drop_p == 0.f || seed, "relpos_softmax_bwd: dropout needs a seed");
  A3T_REQUIRE(S > 0 && S <= 12000, "relpos_softmax_bwd: S=%d out of range", S);
  A3T_REQUIRE(dtype_in == A3T_F32 || dtype_in == A3T_BF16, "relpos_softmax_bwd: bad input dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype_in == A3T_BF16 && dtype_p == A3T_BF16 && dtype_o == A3T_BF16 && (S % 4) == 0 && (ld % 4) == 0 && S <= 2048 &&
      ((((uintptr_t)dPd | (uintptr_t)P | (uintptr_t)dS) & 15) == 0) && !tune_env("A3T_SOFTMAX_SMEM")) {
    const int64_t nrows = (int64_t)B * H * S;
    int rb = (int)((nrows + 7) / 8);
    if (rb > 148 * 2 * 8) rb = 148 * 2 * 8;
#define A3T_SM_BWD_REG(NIT, FULL)                                                                                      \
  relpos_softmax_bwd_reg_kernel<NIT, FULL><<<rb, 256, 0, st>>>((const __nv_bfloat16*)dPd, (const __nv_bfloat16*)P,        \
                                                         (__nv_bfloat16*)dS, (__nv_bfloat16*)dBD, nrows, S, ld, scale, \
                                                         drop_p, seed, site)
    if (S == 128 * 9) A3T_SM_BWD_REG(9, true);
    else if (S <= 128 * 5) A3T_SM_BWD_REG(5, false);
    else if (S <= 128 * 9) A3T_SM_BWD_REG(9, false);
    else if (S <= 128 * 14) A3T_SM_BWD_REG(14, false);
    else A3T_SM_BWD_REG(16, false);
    return check_launch("relpos_softmax_bwd");
  }
  if (dtype_p == A3T_BF16 && dtype_o == A3T_BF16)
    return softmax_bwd_launch<__nv_bfloat16, __nv_bfloat16>(dPd, dtype_in, P, dS, dBD, B, H, S, ld, scale, drop_p, seed, site, st);
  if (dtype_p == A3T_F32 && dtype_o == A3T_F32)
    return softmax_bwd_launch<float, float>(dPd, dtype_in, P, dS, dBD, B, H, S, ld, scale, drop_p, seed, site, st);
  if (dtype_p == A3T_BF16 && dtype_o == A3T_F32)
    return softmax_bwd_launch<__nv_bfloat16, float>(dPd, dtype_in, P, dS, dBD, B, H, S, ld, scale, drop_p, seed, site, st);
  set_error("relpos_softmax_bwd: unsupported dtype combination");
  return A3T_ERR_UNSUPPORTED;
}
