// ParallelWaveGAN generator kernels (espnet2/gan_tts/parallel_wavegan/parallel_wavegan.py:136-229,
// upsample.py:160-189, espnet2/gan_tts/wavenet/residual_block.py:114-169).  Channels-first fp32,
// time contiguous, as in the reference.
#include "common.cuh"

namespace a3t {

// nearest-neighbour stretch x scale, then FIR of length 2*scale+1 (zero pad `scale`)
__global__ void __launch_bounds__(256) pwg_upsample_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                           float* __restrict__ out, int rows, int64_t T, int scale) {
  A3T_PDL_TRIGGER();
  const int64_t To = T * scale;
  const int64_t n = (int64_t)rows * To;
  const int taps = 2 * scale + 1;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    int64_t r = idx / To, t = idx - r * To;
    const float* ir = in + r * T;
    float acc = 0.f;
    for (int k = 0; k < taps; k++) {
      int64_t tt = t + k - scale;
      if (tt >= 0 && tt < To) acc = fmaf(w[k], ir[tt / scale], acc);
    }
    out[idx] = acc;
  }
}

// small dense conv1d; one thread per (b, o, t)
__global__ void __launch_bounds__(256) pwg_conv1d_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int B,
                                                         int Cin, int Cout, int64_t T, int K, int dil, int pad_mode,
                                                         int relu_in, float in_scale) {
  A3T_PDL_TRIGGER();
  const int64_t n = (int64_t)B * Cout * T;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int half = (K - 1) / 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    int64_t t = idx % T;
    int64_t r = idx / T;
    int o = (int)(r % Cout);
    int b = (int)(r / Cout);
    float acc = bias ? bias[o] : 0.f;
    for (int i = 0; i < Cin; i++) {
      const float* ir = in + ((int64_t)b * Cin + i) * T;
      const float* wr = w + ((int64_t)o * Cin + i) * K;
      for (int k = 0; k < K; k++) {
        int64_t tt = t + (int64_t)(k - half) * dil;
        float v;
        if (tt < 0 || tt >= T) {
          if (pad_mode == 0) continue;
          tt = tt < 0 ? 0 : T - 1;
        }
        v = ir[tt] * in_scale;
        if (relu_in) v = fmaxf(v, 0.f);
        acc = fmaf(wr[k], v, acc);
      }
    }
    out[idx] = acc;
  }
}

// ---- fused gated residual block -----------------------------------------------------------
// Tile: all channels x RB_TT samples.  Stage the stacked input [3R + A][RB_TT] in shared memory
// (x[t-d], x[t], x[t+d], c[t]), GEMM1 -> h[G][TT] (registers 8x8 per thread), gate through
// shared memory, GEMM2 -> o[R+S][TT], epilogue writes x_out and accumulates skip.
constexpr int RB_TT = 128;
constexpr int RB_THREADS = 256;

__global__ void __launch_bounds__(RB_THREADS) pwg_resblock_kernel(
    const float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ w_in_t,
    const float* __restrict__ b_in, const float* __restrict__ w_out_t, const float* __restrict__ b_out,
    float* __restrict__ x_out, float* __restrict__ skip, int64_t T, int dil, int first) {
  A3T_PDL_TRIGGER();
  constexpr int R = 64, G = 128, A = 80, KIN = 3 * R + A;  // 272
  extern __shared__ float smem[];
  float* sin = smem;               // [KIN][RB_TT]
  float* sh = smem;                // reused: h [G][RB_TT]
  float* sg = smem + G * RB_TT;    // g [R][RB_TT]
  const int b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * RB_TT;
  const int tid = threadIdx.x;
  const float* xb = x + (int64_t)b * R * T;
  const float* cb = c + (int64_t)b * A * T;
  for (int idx = tid; idx < KIN * RB_TT; idx += RB_THREADS) {
    int row = idx / RB_TT, tl = idx % RB_TT;
    int64_t t = t0 + tl;
    float v = 0.f;
    if (row < 3 * R) {
      int tap = row / R, i = row % R;
      int64_t tt = t + (int64_t)(tap - 1) * dil;
      if (tt >= 0 && tt < T) v = xb[(int64_t)i * T + tt];
    } else if (t < T) {
      v = cb[(int64_t)(row - 3 * R) * T + t];
    }
    sin[idx] = v;
  }
  __syncthreads();
  const int og = tid >> 4, tg = tid & 15;  // 16 x 16 thread grid, 8 outputs x 8 samples each
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float bv = b_in[og * 8 + i];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = bv;
  }
  for (int kk = 0; kk < KIN; kk++) {
    float4 wa = *reinterpret_cast<const float4*>(w_in_t + (int64_t)kk * G + og * 8);
    float4 wb = *reinterpret_cast<const float4*>(w_in_t + (int64_t)kk * G + og * 8 + 4);
    float4 xa = *reinterpret_cast<const float4*>(sin + kk * RB_TT + tg * 8);
    float4 xb4 = *reinterpret_cast<const float4*>(sin + kk * RB_TT + tg * 8 + 4);
    float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb4.x, xb4.y, xb4.z, xb4.w};
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
  }
  __syncthreads();  // everyone is done reading sin
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float4 a = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    float4 bq = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    *reinterpret_cast<float4*>(sh + (og * 8 + i) * RB_TT + tg * 8) = a;
    *reinterpret_cast<float4*>(sh + (og * 8 + i) * RB_TT + tg * 8 + 4) = bq;
  }
  __syncthreads();
  for (int idx = tid; idx < R * RB_TT; idx += RB_THREADS) {
    float ha = sh[idx], hb = sh[idx + R * RB_TT];
    sg[idx] = tanhf(ha) * (1.f / (1.f + expf(-hb)));
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float bv = b_out[og * 8 + i];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = bv;
  }
  for (int kk = 0; kk < R; kk++) {
    float4 wa = *reinterpret_cast<const float4*>(w_out_t + (int64_t)kk * G + og * 8);
    float4 wb = *reinterpret_cast<const float4*>(w_out_t + (int64_t)kk * G + og * 8 + 4);
    float4 xa = *reinterpret_cast<const float4*>(sg + kk * RB_TT + tg * 8);
    float4 xb4 = *reinterpret_cast<const float4*>(sg + kk * RB_TT + tg * 8 + 4);
    float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb4.x, xb4.y, xb4.z, xb4.w};
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
  }
  const float rs = 0.70710678118654752440f;  // math.sqrt(0.5)
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int o = og * 8 + i;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int64_t t = t0 + tg * 8 + j;
      if (t >= T) continue;
      if (o < R) {
        int64_t off = ((int64_t)b * R + o) * T + t;
        x_out[off] = (acc[i][j] + x[off]) * rs;
      } else {
        int64_t off = ((int64_t)b * R + (o - R)) * T + t;
        skip[off] = first ? acc[i][j] : skip[off] + acc[i][j];
      }
    }
  }
}

// ---- fused output stack: out = w2 . relu(W1 relu(skip * scale) + b1) + b2  (parallel_wavegan.py:119-126, 166-173)
// thread = sample (lanes = consecutive samples: the 64 channel-major skip loads are coalesced); the 64 x 64 weight
// lives in shared memory and is read as warp-wide broadcasts; 4 160 FMAs per sample instead of 64 strided passes
__global__ void __launch_bounds__(128) pwg_last_kernel(const float* __restrict__ skip, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ out, int B,
                                                       int64_t T, float scale) {
  A3T_PDL_TRIGGER();
  __shared__ float sw[64 * 64 + 64 + 64];
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) sw[i] = w1[i];
  if (threadIdx.x < 64) {
    sw[64 * 64 + threadIdx.x] = b1[threadIdx.x];
    sw[64 * 64 + 64 + threadIdx.x] = w2[threadIdx.x];
  }
  __syncthreads();
  const float bo = b2[0];
  const int64_t n = (int64_t)B * T;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / T, t = idx - b * T;
    const float* sp = skip + b * 64 * T + t;
    float x[64];
#pragma unroll
    for (int c = 0; c < 64; c++) x[c] = fmaxf(__ldcs(sp + (int64_t)c * T) * scale, 0.f);
    float acc = bo;
#pragma unroll 4
    for (int o = 0; o < 64; o++) {
      float h = sw[64 * 64 + o];
      const float4* wr = reinterpret_cast<const float4*>(sw + o * 64);
#pragma unroll
      for (int c4 = 0; c4 < 16; c4++) {
        const float4 w = wr[c4];
        h = fmaf(w.x, x[4 * c4], h);
        h = fmaf(w.y, x[4 * c4 + 1], h);
        h = fmaf(w.z, x[4 * c4 + 2], h);
        h = fmaf(w.w, x[4 * c4 + 3], h);
      }
      acc = fmaf(sw[64 * 64 + 64 + o], fmaxf(h, 0.f), acc);
    }
    out[idx] = acc;
  }
}

}  // namespace a3t

using namespace a3t;

extern "C" int a3t_pwg_last(const float* skip, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                            int B, int64_t T, float scale, void* stream) {
  A3T_REQUIRE(skip && w1 && b1 && w2 && b2 && out, "pwg_last: null pointer");
  if (B == 0 || T == 0) return A3T_OK;
  int64_t blocks = ((int64_t)B * T + 127) / 128;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pwg_last_kernel<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(skip, w1, b1, w2, b2, out, B, T, scale);
  return check_launch("pwg_last");
}

extern "C" int a3t_pwg_upsample(const float* in, const float* w, float* out, int rows, int64_t T, int scale,
                                void* stream) {
  A3T_REQUIRE(in && w && out && scale >= 1, "pwg_upsample: bad args");
  int64_t n = (int64_t)rows * T * scale;
  if (n == 0) return A3T_OK;
  int64_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  pwg_upsample_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(in, w, out, rows, T, scale);
  return check_launch("pwg_upsample");
}

extern "C" int a3t_pwg_conv1d(const float* in, const float* w, const float* bias, float* out, int B, int Cin, int Cout,
                              int64_t T, int K, int dil, int pad_mode, int relu_in, float in_scale, void* stream) {
  A3T_REQUIRE(in && w && out && (K & 1), "pwg_conv1d: bad args");
  int64_t n = (int64_t)B * Cout * T;
  if (n == 0) return A3T_OK;
  int64_t b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  pwg_conv1d_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(in, w, bias, out, B, Cin, Cout, T, K, dil, pad_mode,
                                                              relu_in, in_scale);
  return check_launch("pwg_conv1d");
}

extern "C" int a3t_pwg_resblock(const float* x, const float* c, const float* w_in_t, const float* b_in,
                                const float* w_out_t, const float* b_out, float* x_out, float* skip, int B, int64_t T,
                                int R, int G, int A, int S_, int dil, int first, void* stream) {
  A3T_REQUIRE(x && c && w_in_t && b_in && w_out_t && b_out && x_out && skip, "pwg_resblock: null pointer");
  A3T_REQUIRE(R == 64 && G == 128 && A == 80 && S_ == 64,
              "pwg_resblock: only the v1 generator shape (residual 64, gate 128, aux 80, skip 64) is built");
  A3T_REQUIRE(x != x_out, "pwg_resblock: x_out must not alias x (neighbouring tiles read the halo)");
  if (B == 0 || T == 0) return A3T_OK;
  size_t smem = (size_t)(3 * 64 + 80) * RB_TT * sizeof(float);
  cudaFuncSetAttribute(pwg_resblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)((T + RB_TT - 1) / RB_TT), B);
  pwg_resblock_kernel<<<grid, RB_THREADS, smem, (cudaStream_t)stream>>>(x, c, w_in_t, b_in, w_out_t, b_out, x_out,
                                                                       skip, T, dil, first);
  return check_launch("pwg_resblock");
}
