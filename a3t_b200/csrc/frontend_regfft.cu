// STFT -> log-mel, warp-per-frame register FFT (the default for n_fft = 1024 / 2048; frontend.cu keeps the
// shared-memory Stockham kernel for other sizes).  Same arithmetic as espnet2/layers/stft.py:56-124 (torch.stft:
// center / reflect pad, periodic Hann zero-padded centred to n_fft, onesided), espnet2/layers/log_mel.py:56-83 and
// log_mel_fbank.py:88-106.
//
// The n_fft real samples of a frame are packed as NC = n_fft/2 complex points z[n] = x[2n] + i x[2n+1] and transformed
// by a two-level Cooley-Tukey FFT, NC = 32 (lanes) x R (registers), R = NC/32:
//   1. lane n1 holds z[n1 + 32 n2], n2 < R, and runs an R-point radix-2 FFT over n2 entirely in registers
//      (compile-time twiddles from constant memory, trivial ones folded);
//   2. multiplies by the inter-level twiddles W_NC^(n1 k2) (powers of a per-lane base, by recurrence);
//   3. a transposition through a padded per-warp shared-memory tile gives lane k2 the 32 values of its column,
//      and a 32-point register FFT over n1 finishes Z[R k1 + k2];
//   4. the real-input unpack pairs Z[k] with Z[NC-k] (both bins of a pair from one twiddle), amplitude -> per-warp buffer;
//   5. each lane sums the mel filters it owns from a compact shared-memory copy of the triangular weights, writes log10.
// Two barriers-free shared-memory exchanges per frame instead of log4(NC) CTA barriers, no transcendental per butterfly.
// HBM traffic per frame is unchanged (hop new samples in, n_mels out); overlapping windows are served by L1/L2.
#include "common.cuh"

namespace a3t {
namespace fe2 {

constexpr int WARPS = 8;

__constant__ float2 c_tw32[16];   // exp(-2 pi i t / 32), t < 16
__constant__ float2 c_tw64[32];   // exp(-2 pi i t / 64), t < 32

__host__ __device__ constexpr int bitrev(int v, int bits) {
  int r = 0;
  for (int i = 0; i < bits; i++) r |= ((v >> i) & 1) << (bits - 1 - i);
  return r;
}
__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-register radix-2 decimation-in-frequency FFT of R points (R = 16 or 32): natural order in, X[k] at index
// bitrev(k) out.  Every index is a compile-time constant after unrolling, so `a` stays in registers and the twiddles
// are constant-bank operands.
template <int R>
__device__ __forceinline__ void fft_regs(float2 (&a)[R]) {
#pragma unroll
  for (int len = R; len >= 2; len >>= 1) {
    const int half = len >> 1;
#pragma unroll
    for (int b0 = 0; b0 < R; b0 += len) {
#pragma unroll
      for (int j = 0; j < half; j++) {
        const float2 u = a[b0 + j], v = a[b0 + j + half];
        a[b0 + j] = make_float2(u.x + v.x, u.y + v.y);
        const float2 d = make_float2(u.x - v.x, u.y - v.y);
        const int t = j * (32 / len);              // twiddle exp(-2 pi i j / len) = c_tw32[t]
        if (t == 0) a[b0 + j + half] = d;
        else if (t == 8) a[b0 + j + half] = make_float2(d.y, -d.x);   // times -i
        else a[b0 + j + half] = cmul(d, c_tw32[t]);
      }
    }
  }
}

// NCT = n_fft / 2 (512 or 1024)
constexpr int MAX_MELS = 128, MAX_CW = 2560;   // compact mel weights: sum of the filters' bin ranges (2 130 for 80 mels / 2048)

template <int NCT>
__global__ void __launch_bounds__(WARPS * 32, 2)
stft_logmel_regfft_kernel(const float* __restrict__ wav, const int64_t* __restrict__ ilens, const float* __restrict__ window,
                          const float* __restrict__ melmat, const int32_t* __restrict__ mel_range, float* __restrict__ mel,
                          int B, int64_t N, int T, int win_length, int hop, int n_mels) {
  A3T_PDL_TRIGGER();
  constexpr int R = NCT / 32;
  constexpr int RB = ilog2(R);
  constexpr int n_fft = 2 * NCT;
  extern __shared__ float2 smem2[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2* tile = smem2 + warp * (R * 33);                                   // per warp: [k2][n1] padded to 33
  float* ampb = reinterpret_cast<float*>(smem2 + WARPS * (R * 33)) + warp * (NCT + 8);
  float* cw = reinterpret_cast<float*>(smem2 + WARPS * (R * 33)) + WARPS * (NCT + 8);   // compact mel weights
  int* moff = reinterpret_cast<int*>(cw + MAX_CW);                          // [n_mels + 1] offsets into cw
  int* mlo = moff + MAX_MELS + 1;                                           // first bin of each filter
  // ---- compact copy of the (sparse, triangular) mel matrix: filter m = cw[moff[m] .. moff[m+1]) over bins mlo[m] ..
  for (int m = threadIdx.x; m < n_mels; m += WARPS * 32) {   // (ranges first, in parallel: the prefix sum then runs on shared memory)
    mlo[m] = mel_range[2 * m];
    moff[m + 1] = mel_range[2 * m + 1] - mel_range[2 * m];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int o = 0;
    for (int m = 0; m < n_mels; m++) {
      const int wdt = moff[m + 1];
      moff[m] = o;
      o += wdt;
    }
    moff[n_mels] = o;
  }
  __syncthreads();
  for (int m = warp; m < n_mels; m += WARPS)
    for (int j = lane; j < moff[m + 1] - moff[m]; j += 32) cw[moff[m] + j] = melmat[(int64_t)(mlo[m] + j) * n_mels + m];
  __syncthreads();
  float2 wl, wb;   // exp(-2 pi i lane / n_fft): base of this lane's unpack twiddles; exp(-2 pi i lane / NC): inter-level base
  {
    float sn, cs;
    sincospif(-2.0f * (float)lane / (float)n_fft, &sn, &cs);
    wl = make_float2(cs, sn);
    sincospif(-2.0f * (float)lane / (float)NCT, &sn, &cs);
    wb = make_float2(cs, sn);
  }
  const int woff = (n_fft - win_length) >> 1;
  const int64_t nframes = (int64_t)B * T;
  for (int64_t frame = (int64_t)blockIdx.x * WARPS + warp; frame < nframes; frame += (int64_t)gridDim.x * WARPS) {
    const int b = (int)(frame / T), t = (int)(frame - (int64_t)b * T);
    const int64_t ilen = ilens ? ilens[b] : N;
    const int64_t olen = (ilen + 2 * (win_length / 2) - win_length) / hop + 1;
    float* out = mel + frame * n_mels;
    if (t >= olen) {  // padded frame: log_mel.py:78 zero fill
      for (int m = lane; m < n_mels; m += 32) out[m] = 0.f;
      continue;
    }
    // ---- load: frame sample s sits at wav index t*hop - n_fft/2 + s (reflected at the ends), times the centred window
    const float* w = wav + (int64_t)b * N;
    const int64_t base = (int64_t)t * hop - NCT;
    const bool interior = base >= 0 && base + n_fft <= N && ((base & 1) == 0) && ((woff & 1) == 0) &&
                          ((reinterpret_cast<uintptr_t>(w) & 7) == 0);
    float2 a[R];
#pragma unroll
    for (int n2 = 0; n2 < R; n2++) {
      const int s = 2 * (lane + 32 * n2);
      const int wi = s - woff;
      float2 v = make_float2(0.f, 0.f);
      if (wi >= -1 && wi < win_length) {
        if (interior && wi >= 0 && wi + 1 < win_length) {
          const float2 x = __ldg(reinterpret_cast<const float2*>(w + base + s));
          const float2 h = __ldg(reinterpret_cast<const float2*>(window + wi));
          v = make_float2(x.x * h.x, x.y * h.y);
        } else {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int wie = wi + e;
            if (wie >= 0 && wie < win_length) {
              int64_t idx = base + s + e;
              if (idx < 0) idx = -idx;
              if (idx >= N) idx = 2 * (N - 1) - idx;
              const float val = w[idx] * window[wie];
              if (e == 0) v.x = val; else v.y = val;
            }
          }
        }
      }
      a[n2] = v;
    }
    // ---- level 1: R-point FFT over n2 (per lane); inter-level twiddle W_NC^(lane k2) by recurrence on wb; transposition
    fft_regs<R>(a);
    {
      float2 tw = wb;
      tile[lane] = a[0];
#pragma unroll
      for (int k2 = 1; k2 < R; k2++) {
        tile[k2 * 33 + lane] = cmul(a[bitrev(k2, RB)], tw);
        tw = cmul(tw, wb);
      }
    }
    __syncwarp();
    // ---- level 2: lane k2 < R runs the 32-point FFT over n1; Z[R k1 + k2] ends up at register bitrev(k1, 5)
    if (lane < R) {
      float2 z[32];
#pragma unroll
      for (int n1 = 0; n1 < 32; n1++) z[n1] = tile[lane * 33 + n1];
      fft_regs<32>(z);
      __syncwarp(R == 32 ? 0xffffffffu : ((1u << (R & 31)) - 1u));
      // spectrum back over all 32 lanes through the tile: bin k at [k >> 5][k & 31] (for R = 32 lane k2 keeps its column)
#pragma unroll
      for (int k1 = 0; k1 < 32; k1++) tile[(R * k1 + lane) % 32 + 33 * ((R * k1 + lane) / 32)] = z[bitrev(k1, 5)];
    }
    __syncwarp();
    // ---- real-input unpack, two bins per step: with E = (Z[k] + conj(Z[NC-k])) / 2 and Tw = W^k (Z[k] - conj(Z[NC-k])) / (2i),
    // X[k] = E + Tw and X[NC-k] = conj(E - Tw); amplitude = sqrt(max(|X|^2, 1e-10)).  Bins 0, NC/2 and NC fall out of the
    // k = 0 and k = NC/2 pairs.
#pragma unroll
    for (int sl = 0; sl <= R / 2; sl++) {
      const int k = 32 * sl + lane;
      if (sl == R / 2 && lane != 0) break;           // only bin NC/2 is left
      const int kc = (NCT - k) & (NCT - 1);
      const float2 zk = tile[(k & 31) + 33 * (k >> 5)];
      float2 zc = tile[(kc & 31) + 33 * (kc >> 5)];
      zc.y = -zc.y;
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
      const float2 o = make_float2(zk.x - zc.x, zk.y - zc.y);
      // exp(-2 pi i k / n_fft) = wl * exp(-2 pi i 32 sl / n_fft) = wl * c_tw64[sl * (64 * 32 / n_fft)]
      const float2 tw = sl == 0 ? wl : cmul(wl, c_tw64[(sl * (2048 / n_fft)) & 31]);
      const float2 r = cmul(tw, o);                  // Tw = r * (-i/2) = (r.y / 2, -r.x / 2)
      const float tx = 0.5f * r.y, ty = -0.5f * r.x;
      const float re = e.x + tx, im = e.y + ty, re2 = e.x - tx, im2 = e.y - ty;
      ampb[k] = sqrtf(fmaxf(re * re + im * im, 1.0e-10f));
      ampb[NCT - k] = sqrtf(fmaxf(re2 * re2 + im2 * im2, 1.0e-10f));   // k = 0: the Nyquist bin X[NC] = Re(Z0) - Im(Z0)
    }
    __syncwarp();
    // ---- mel projection: lane m owns filters m, m + 32, ...; then log10(max(., 1e-10))
    for (int m = lane; m < n_mels; m += 32) {
      const float* wv = cw + moff[m];
      const float* av = ampb + mlo[m];
      const int n = moff[m + 1] - moff[m];
      float acc = 0.f;
      for (int j = 0; j < n; j++) acc = fmaf(av[j], wv[j], acc);
      out[m] = log10f(fmaxf(acc, 1.0e-10f));
    }
    __syncwarp();
  }
}

static bool g_tables = false;
static int init_tables() {
  if (g_tables) return A3T_OK;
  float2 t32[16], t64[32];
  for (int t = 0; t < 16; t++) {
    const double a = -2.0 * 3.14159265358979323846 * t / 32.0;
    t32[t] = make_float2((float)cos(a), (float)sin(a));
  }
  for (int t = 0; t < 32; t++) {
    const double a = -2.0 * 3.14159265358979323846 * t / 64.0;
    t64[t] = make_float2((float)cos(a), (float)sin(a));
  }
  if (cudaMemcpyToSymbol(c_tw32, t32, sizeof(t32)) != cudaSuccess || cudaMemcpyToSymbol(c_tw64, t64, sizeof(t64)) != cudaSuccess) {
    set_error("stft_logmel: constant table upload failed");
    return A3T_ERR_CUDA;
  }
  g_tables = true;
  return A3T_OK;
}

template <int NCT>
static int launch(const float* wav, const int64_t* ilens, const float* window, const float* melmat, const int32_t* mel_range,
                  float* mel, int B, int64_t N, int T, int win_length, int hop, int n_mels, cudaStream_t st) {
  constexpr int R = NCT / 32;
  const size_t smem = (size_t)(WARPS * R * 33) * sizeof(float2) + (size_t)(WARPS * (NCT + 8) + MAX_CW + 2 * MAX_MELS + 8) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(stft_logmel_regfft_kernel<NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = true;
  }
  const int64_t nframes = (int64_t)B * T;
  int64_t blocks = (nframes + WARPS - 1) / WARPS;
  if (blocks > 148 * 2) blocks = 148 * 2;   // persistent: two CTAs per SM, each builds its tables once and strides over the frames
  stft_logmel_regfft_kernel<NCT><<<(int)blocks, WARPS * 32, smem, st>>>(wav, ilens, window, melmat, mel_range, mel, B, N, T,
                                                                         win_length, hop, n_mels);
  return check_launch("stft_logmel(regfft)");
}

}  // namespace fe2

// returns A3T_ERR_UNSUPPORTED when this n_fft is not built (the caller then takes the shared-memory kernel)
int stft_logmel_regfft(const float* wav, const int64_t* ilens, const float* window, const float* melmat,
                       const int32_t* mel_range, float* mel, int B, int64_t N, int T, int n_fft, int win_length, int hop,
                       int n_mels, cudaStream_t st) {
  if ((n_fft != 2048 && n_fft != 1024) || n_mels > fe2::MAX_MELS || !mel_range) return A3T_ERR_UNSUPPORTED;
  if (2 * (n_fft / 2 + 1) + n_mels > fe2::MAX_CW) return A3T_ERR_UNSUPPORTED;   // bound on the compact weights (<= 2 filters per bin)
  int rc = fe2::init_tables();
  if (rc) return rc;
  if (n_fft == 2048) return fe2::launch<1024>(wav, ilens, window, melmat, mel_range, mel, B, N, T, win_length, hop, n_mels, st);
  return fe2::launch<512>(wav, ilens, window, melmat, mel_range, mel, B, N, T, win_length, hop, n_mels, st);
}

}  // namespace a3t
