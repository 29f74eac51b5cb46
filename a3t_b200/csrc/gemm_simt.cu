// Generic strided/batched GEMM on CUDA cores with fp32 accumulation.
// This is the exact-fp32 parity path and the fallback for shapes the tcgen05 kernel (gemm_tc.cu)
// does not take (K=80 pre-net, N=80 head, tiny position GEMM).  Same epilogue as the tensor-core
// kernel.  See include/a3t_b200.h::a3t_gemm for the contract.
#include <stdarg.h>
#include "common.cuh"

namespace a3t {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return A3T_ERR_CUDA;
  }
  return A3T_OK;
}

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

// element fetchers --------------------------------------------------------------------------
__device__ __forceinline__ float fetch_a(const A3tGemmDesc& d, const void* A, int m, int k) {
  if (m >= d.M || k >= d.K) return 0.f;
  if (d.mode == A3T_GEMM_CONV) {
    int tap = k / d.cin, c = k - tap * d.cin;
    int t = m % d.seq;
    int ts = t + tap - d.pad;
    if (ts < 0 || ts >= d.seq) return 0.f;
    return load_as_f32(A, d.dtype_a, (int64_t)(m - t + ts) * d.sa_m + (int64_t)c * d.sa_k);
  }
  return load_as_f32(A, d.dtype_a, (int64_t)m * d.sa_m + (int64_t)k * d.sa_k);
}
__device__ __forceinline__ float fetch_b(const A3tGemmDesc& d, const void* B, int n, int k) {
  if (n >= d.N || k >= d.K) return 0.f;
  if (d.mode == A3T_GEMM_CONV) {
    int tap = k / d.cin, c = k - tap * d.cin;
    return load_as_f32(B, d.dtype_b, (int64_t)n * d.sb_n + (int64_t)tap * d.sb_tap + (int64_t)c * d.sb_k);
  }
  if (d.mode == A3T_GEMM_WGRAD) {
    int tap = n / d.cin, c = n - tap * d.cin;
    int t = k % d.seq;
    int ts = t + tap - d.pad;
    if (ts < 0 || ts >= d.seq) return 0.f;
    return load_as_f32(B, d.dtype_b, (int64_t)(k - t + ts) * d.sb_k + (int64_t)c * d.sb_n);
  }
  return load_as_f32(B, d.dtype_b, (int64_t)n * d.sb_n + (int64_t)k * d.sb_k);
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const A3tGemmDesc d, const void* __restrict__ A0,
                                                        const void* __restrict__ B0, void* __restrict__ C0,
                                                        const float* __restrict__ bias,
                                                        const float* __restrict__ res0,
                                                        const void* __restrict__ mask0,
                                                        const unsigned long long* __restrict__ seed) {
  A3T_PDL_TRIGGER();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int bz = blockIdx.z;
  const int b1 = bz / d.batch2, b2 = bz - b1 * d.batch2;
  const size_t esa = d.dtype_a == A3T_BF16 ? 2 : 4, esb = d.dtype_b == A3T_BF16 ? 2 : 4,
               esc = d.dtype_c == A3T_BF16 ? 2 : 4, esm = d.dtype_mask == A3T_BF16 ? 2 : 4;
  const void* A = (const char*)A0 + (b1 * d.sa_b1 + b2 * d.sa_b2) * (int64_t)esa;
  const void* B = (const char*)B0 + (b1 * d.sb_b1 + b2 * d.sb_b2) * (int64_t)esb;
  void* C = (char*)C0 + (b1 * d.sc_b1 + b2 * d.sc_b2) * (int64_t)esc;
  const void* mask = mask0 ? (const char*)mask0 + (b1 * d.sc_b1 + b2 * d.sc_b2) * (int64_t)esm : nullptr;
  const float* res = res0 ? res0 + (b1 * d.sr_b1 + b2 * d.sr_b2) : nullptr;

  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // loader mapping: if the reduction index is the contiguous one, let consecutive threads walk k.
  const bool a_kfast = (d.mode == A3T_GEMM_WGRAD) ? false : (d.sa_k == 1);
  const bool b_kfast = (d.mode == A3T_GEMM_WGRAD) ? false : (d.sb_k == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < d.K; k0 += BK) {
#pragma unroll
    for (int e = 0; e < (BM * BK) / 256; e++) {
      int idx = tid + e * 256;
      int mm, kk;
      if (a_kfast) { kk = idx & (BK - 1); mm = idx >> 4; } else { mm = idx & (BM - 1); kk = idx >> 6; }
      As[kk][mm] = fetch_a(d, A, m0 + mm, k0 + kk);
    }
#pragma unroll
    for (int e = 0; e < (BN * BK) / 256; e++) {
      int idx = tid + e * 256;
      int nn, kk;
      if (b_kfast) { kk = idx & (BK - 1); nn = idx >> 4; } else { nn = idx & (BN - 1); kk = idx >> 6; }
      Bs[kk][nn] = fetch_b(d, B, n0 + nn, k0 + kk);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; j++) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  Drop dr = make_drop(d.drop_p, seed, d.drop_site);
#pragma unroll
  for (int i = 0; i < TM; i++) {
    int m = m0 + ty * TM + i;
    if (m >= d.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      int n = n0 + tx * TN + j;
      if (n >= d.N) continue;
      float v = d.alpha * acc[i][j];
      if (bias) v += bias[n];
      if (d.relu) v = fmaxf(v, 0.f);
      int64_t coff;
      if (d.mode == A3T_GEMM_WGRAD) {
        int tap = n / d.cin, c = n - tap * d.cin;
        coff = (int64_t)m * d.sc_m + (int64_t)tap * d.sc_tap + (int64_t)c * d.sc_n;
      } else {
        coff = (int64_t)m * d.sc_m + (int64_t)n * d.sc_n;
      }
      if (mask) {
        float mv = load_as_f32(mask, d.dtype_mask, coff);
        v = (mv != 0.f) ? v * d.mask_scale : 0.f;
      }
      if (dr.on) v = drop_apply(dr, ((unsigned long long)bz * d.M + m) * (unsigned long long)d.N + n, v);
      v *= d.out_scale;
      if (res) v += res[(int64_t)m * d.sr_m + (int64_t)n * d.sr_n];
      store_from_f32(C, d.dtype_c, coff, v);
    }
  }
}

int gemm_simt_launch(const A3tGemmDesc* d, const void* A, const void* B, void* C, const float* bias,
                     const float* res, const void* mask, const unsigned long long* seed,
                     cudaStream_t st) {
  dim3 grid(ceil_div(d->N, BN), ceil_div(d->M, BM), d->batch1 * d->batch2);
  A3T_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm_simt: grid too large (M=%d batch=%d)", d->M,
              d->batch1 * d->batch2);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(*d, A, B, C, bias, res, mask, seed);
  return check_launch("gemm_simt");
}

// weight packing -------------------------------------------------------------------------------
// 32 (n) x 32 (c) x taps tile through shared memory: the fp32 read is contiguous in (c, tap), the
// forward pack is written c-fastest and the dgrad pack n-fastest, so all three streams are coalesced
// (the first version wrote the dgrad pack with stride N: 21 us per FFN weight instead of ~3).
constexpr int PK_T = 32;
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const float* __restrict__ w, int N, int C, int taps,
                                                               __nv_bfloat16* __restrict__ fwd,
                                                               __nv_bfloat16* __restrict__ dg) {
  A3T_PDL_TRIGGER();
  extern __shared__ float tile[];  // [PK_T][PK_T * taps + 1]
  const int rowlen = PK_T * taps, ld = rowlen + 1;
  const int n0 = blockIdx.x * PK_T, c0 = blockIdx.y * PK_T;
  for (int idx = threadIdx.x; idx < PK_T * rowlen; idx += 256) {
    int n = idx / rowlen, j = idx - n * rowlen;
    int c = c0 + j / taps;
    float v = 0.f;
    if (n0 + n < N && c < C) v = w[((int64_t)(n0 + n) * C + c0) * taps + j];
    tile[n * ld + j] = v;
  }
  __syncthreads();
  if (fwd) {
    for (int idx = threadIdx.x; idx < PK_T * rowlen; idx += 256) {  // (n, tap, cc) with cc fastest
      int cc = idx % PK_T, r = idx / PK_T;
      int tap = r % taps, n = r / taps;
      if (n0 + n < N && c0 + cc < C)
        fwd[(int64_t)(n0 + n) * taps * C + (int64_t)tap * C + c0 + cc] = __float2bfloat16_rn(tile[n * ld + cc * taps + tap]);
    }
  }
  if (dg) {
    for (int idx = threadIdx.x; idx < PK_T * rowlen; idx += 256) {  // (cc, tap, n) with n fastest
      int n = idx % PK_T, r = idx / PK_T;
      int tap = r % taps, cc = r / taps;
      if (n0 + n < N && c0 + cc < C)
        dg[(int64_t)(c0 + cc) * taps * N + (int64_t)(taps - 1 - tap) * N + n0 + n] =
            __float2bfloat16_rn(tile[n * ld + cc * taps + tap]);
    }
  }
}

// batched variant: block -> (item, n tile, c tile) through the items' running tile counts.
// 64 x 64 (x taps) tiles staged in shared memory as bf16; both packs are written as 4-byte bf16 pairs so a
// warp stores 128 contiguous bytes (64 channels of one (n, tap) row for `fwd`, 64 output rows of one
// (c, tap) row for `dgrad`).  Reads are 16-byte vectors when the rows allow it.
constexpr int PK_B = 64;
__global__ void __launch_bounds__(256) pack_conv_weights_kernel(const A3tPackItem* __restrict__ items, int n_items) {
  A3T_PDL_TRIGGER();
  extern __shared__ __align__(16) unsigned char pk_smem[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(pk_smem);  // [PK_B][PK_B * taps + 2]
  __shared__ int s_item;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {  // last item whose tile_start <= blockIdx.x
      int mid = (lo + hi + 1) >> 1;
      if (items[mid].tile_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_item = lo;
  }
  __syncthreads();
  const A3tPackItem it = items[s_item];
  const int local = blockIdx.x - it.tile_start;
  const int n0 = (local / it.tiles_c) * PK_B, c0 = (local % it.tiles_c) * PK_B;
  const int N = it.N, C = it.C, taps = it.taps;
  const int rowlen = PK_B * taps, ld = rowlen + 2;   // ld/2 odd: the column walks of the dgrad pass hit distinct banks
  // output row n0 + n comes from source (n0 + n) / seg_rows (stacked sources; a tile may straddle two of them)
  const int cw = min(PK_B, C - c0);                  // channels of this tile
  const int rl = cw * taps;                          // valid floats per row
  const bool vec = ((C * taps) % 4) == 0 && ((c0 * taps) % 4) == 0 && (rl % 4) == 0 &&
                   ((((uintptr_t)it.w[0] | (uintptr_t)it.w[1] | (uintptr_t)it.w[2] | (uintptr_t)it.w[3]) & 15) == 0);
  if (vec) {
    const int q = rl >> 2;
    for (int idx = threadIdx.x; idx < PK_B * q; idx += 256) {
      const int n = idx / q, j = (idx - n * q) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + n < N) {
        const int sg = (n0 + n) / it.seg_rows, rn = (n0 + n) - sg * it.seg_rows;
        v = __ldcs(reinterpret_cast<const float4*>(it.w[sg] + ((int64_t)rn * C + c0) * taps + j));
      }
      __nv_bfloat162* d = reinterpret_cast<__nv_bfloat162*>(tile + n * ld + j);
      d[0] = __floats2bfloat162_rn(v.x, v.y);
      d[1] = __floats2bfloat162_rn(v.z, v.w);
    }
  } else {
    for (int idx = threadIdx.x; idx < PK_B * rl; idx += 256) {
      const int n = idx / rl, j = idx - n * rl;
      float v = 0.f;
      if (n0 + n < N) {
        const int sg = (n0 + n) / it.seg_rows, rn = (n0 + n) - sg * it.seg_rows;
        v = it.w[sg][((int64_t)rn * C + c0) * taps + j];
      }
      tile[n * ld + j] = __float2bfloat16_rn(v);
    }
  }
  __syncthreads();
  __nv_bfloat16* fwd = (__nv_bfloat16*)it.fwd;
  __nv_bfloat16* dg = (__nv_bfloat16*)it.dgrad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nrows = min(PK_B, N - n0);  // valid output rows of this tile
  if (fwd) {  // fwd[n, tap*C + c]: a warp writes the 64 channels of one (n, tap)
    const bool pair = (C % 2) == 0;  // 4-byte aligned channel pairs
    for (int r = warp; r < nrows * taps; r += 8) {
      const int n = r / taps, tap = r - n * taps;
      __nv_bfloat16* dst = fwd + (int64_t)(n0 + n) * taps * C + (int64_t)tap * C + c0;
      const __nv_bfloat16* src = tile + n * ld + tap;
      const int cc = lane * 2;
      if (pair && cc + 1 < cw) {
        __nv_bfloat162 h;
        h.x = src[cc * taps];
        h.y = src[(cc + 1) * taps];
        *reinterpret_cast<__nv_bfloat162*>(dst + cc) = h;
      } else {
        if (cc < cw) dst[cc] = src[cc * taps];
        if (cc + 1 < cw) dst[cc + 1] = src[(cc + 1) * taps];
      }
    }
  }
  if (dg) {   // dgrad[c, (taps-1-tap)*N + n]: a warp writes the 64 rows n of one (c, tap)
    const bool pair = (N % 2) == 0;
    for (int r = warp; r < cw * taps; r += 8) {
      const int cc = r / taps, tap = r - cc * taps;
      __nv_bfloat16* dst = dg + (int64_t)(c0 + cc) * taps * N + (int64_t)(taps - 1 - tap) * N + n0;
      const __nv_bfloat16* src = tile + cc * taps + tap;
      const int n = lane * 2;
      if (pair && n + 1 < nrows) {
        __nv_bfloat162 h;
        h.x = src[n * ld];
        h.y = src[(n + 1) * ld];
        *reinterpret_cast<__nv_bfloat162*>(dst + n) = h;
      } else {
        if (n < nrows) dst[n] = src[n * ld];
        if (n + 1 < nrows) dst[n + 1] = src[(n + 1) * ld];
      }
    }
  }
}

__global__ void qkv4_bias_kernel(const float* __restrict__ bq, const float* __restrict__ bk, const float* __restrict__ bv,
                                 const float* __restrict__ u, const float* __restrict__ v, float* __restrict__ out,
                                 int D) {
  A3T_PDL_TRIGGER();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  out[i] = bq[i] + u[i];
  out[D + i] = bq[i] + v[i];
  out[2 * D + i] = bk[i];
  out[3 * D + i] = bv[i];
}

}  // namespace a3t

extern "C" const char* a3t_last_error(void) { return a3t::g_err; }
extern "C" int a3t_version(void) { return 100; }

extern "C" int a3t_pack_conv_weight(const float* w, int N, int C, int taps, void* fwd, void* dg, void* stream) {
  A3T_REQUIRE(w && N > 0 && C > 0 && taps > 0, "pack_conv_weight: bad args");
  A3T_REQUIRE(taps <= 11, "pack_conv_weight: taps=%d too large (max 11)", taps);
  dim3 grid((N + a3t::PK_T - 1) / a3t::PK_T, (C + a3t::PK_T - 1) / a3t::PK_T);
  size_t smem = (size_t)a3t::PK_T * (a3t::PK_T * taps + 1) * sizeof(float);
  a3t::pack_conv_weight_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(w, N, C, taps, (__nv_bfloat16*)fwd,
                                                                         (__nv_bfloat16*)dg);
  return a3t::check_launch("pack_conv_weight");
}

extern "C" int a3t_pack_conv_weights(const A3tPackItem* items, int n_items, int total_tiles, int max_taps, void* stream) {
  A3T_REQUIRE(items && n_items > 0 && total_tiles > 0, "pack_conv_weights: bad args");
  A3T_REQUIRE(max_taps >= 1 && max_taps <= 11, "pack_conv_weights: max_taps=%d out of range (1..11)", max_taps);
  size_t smem = (size_t)a3t::PK_B * (a3t::PK_B * max_taps + 2) * sizeof(__nv_bfloat16);
  if (smem > 48 * 1024) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(a3t::pack_conv_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr = true;
    }
  }
  a3t::pack_conv_weights_kernel<<<total_tiles, 256, smem, (cudaStream_t)stream>>>(items, n_items);
  return a3t::check_launch("pack_conv_weights");
}

extern "C" int a3t_qkv4_bias(const float* bq, const float* bk, const float* bv, const float* u, const float* v,
                             float* out, int D, void* stream) {
  A3T_REQUIRE(bq && bk && bv && u && v && out && D > 0, "qkv4_bias: bad args");
  a3t::qkv4_bias_kernel<<<(D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bq, bk, bv, u, v, out, D);
  return a3t::check_launch("qkv4_bias");
}
