// STFT -> log-mel frontend and the collate integer math.
// Reference arithmetic: espnet2/layers/stft.py:56-124 (torch.stft, center/reflect, periodic Hann
// zero-padded centred to n_fft, onesided), espnet2/layers/log_mel.py:56-83, log_mel_fbank.py:88-106;
// espnet2/train/collate_fn.py:236-237 (align floor), :330-343 (segment pos), :346-385 (span expansion).
//
// One CTA per frame: the n_fft real samples are packed as n_fft/2 complex values in shared
// memory, transformed by a Stockham autosort FFT (radix-4 stages + one radix-2 stage when
// log2 is odd), unpacked to n_fft/2+1 bins, and reduced against the (sparse, triangular)
// mel matrix.  Only wav-in / mel-out touch HBM (1 520 B per frame at hop 300).  Twiddles come from the
// MUFU sin/cos (|angle| <= pi: absolute error ~5e-7, two orders below the 1e-4 log-mel gate).
#include "common.cuh"

namespace a3t {

constexpr int FE_THREADS = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// smem: two complex buffers of NC = n_fft/2 entries + amplitude buffer of NC+1 floats
__global__ void __launch_bounds__(FE_THREADS) stft_logmel_kernel(
    const float* __restrict__ wav, const int64_t* __restrict__ ilens, const float* __restrict__ window,
    const float* __restrict__ melmat, const int32_t* __restrict__ mel_range, float* __restrict__ mel,
    int B, int64_t N, int T, int n_fft, int win_length, int hop, int n_mels) {
  A3T_PDL_TRIGGER();
  extern __shared__ float2 smem[];
  const int NC = n_fft >> 1;
  float2* bufA = smem;
  float2* bufB = smem + NC;
  float* amp = reinterpret_cast<float*>(smem + 2 * NC);
  const int frame = blockIdx.x;
  const int b = frame / T, t = frame - b * T;
  const int tid = threadIdx.x;
  const int64_t ilen = ilens ? ilens[b] : N;
  const int64_t olen = (ilen + 2 * (win_length / 2) - win_length) / hop + 1;
  float* out = mel + ((int64_t)b * T + t) * n_mels;
  if (t >= olen) {  // padded frame: log_mel.py:78 zero fill
    for (int m = tid; m < n_mels; m += FE_THREADS) out[m] = 0.f;
    return;
  }
  // ---- load: frame sample n sits at padded index t*hop + n, i.e. wav index t*hop + n - n_fft/2
  const int woff = (n_fft - win_length) >> 1;
  const float* w = wav + (int64_t)b * N;
  const int64_t base = (int64_t)t * hop - NC;
  for (int n = tid; n < NC; n += FE_THREADS) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      int s = 2 * n + e;
      int wi = s - woff;
      float val = 0.f;
      if (wi >= 0 && wi < win_length) {
        int64_t idx = base + s;
        if (idx < 0) idx = -idx;
        if (idx >= N) idx = 2 * (N - 1) - idx;
        val = w[idx] * window[wi];
      }
      v[e] = val;
    }
    bufA[n] = make_float2(v[0], v[1]);
  }
  __syncthreads();
  // ---- Stockham FFT of NC complex points
  float2* x = bufA;
  float2* y = bufB;
  int Ns = 1;
  while (Ns * 4 <= NC) {
    const int Tq = NC >> 2;
    for (int j = tid; j < Tq; j += FE_THREADS) {
      int k = j & (Ns - 1);
      float sn, cs;
      __sincosf(-3.14159265358979f * (float)k / (float)(2 * Ns), &sn, &cs);  // exp(-2 pi i k / (4 Ns)); |angle| < pi/2
      float2 w1 = make_float2(cs, sn);
      float2 w2 = cmul(w1, w1);
      float2 w3 = cmul(w2, w1);
      float2 u0 = x[j], u1 = cmul(x[j + Tq], w1), u2 = cmul(x[j + 2 * Tq], w2), u3 = cmul(x[j + 3 * Tq], w3);
      float2 v0 = make_float2(u0.x + u2.x, u0.y + u2.y);
      float2 v1 = make_float2(u0.x - u2.x, u0.y - u2.y);
      float2 v2 = make_float2(u1.x + u3.x, u1.y + u3.y);
      float2 d = make_float2(u1.x - u3.x, u1.y - u3.y);
      float2 v3 = make_float2(d.y, -d.x);
      int j0 = ((j - k) << 2) + k;
      y[j0] = make_float2(v0.x + v2.x, v0.y + v2.y);
      y[j0 + Ns] = make_float2(v1.x + v3.x, v1.y + v3.y);
      y[j0 + 2 * Ns] = make_float2(v0.x - v2.x, v0.y - v2.y);
      y[j0 + 3 * Ns] = make_float2(v1.x - v3.x, v1.y - v3.y);
    }
    __syncthreads();
    float2* tmp = x; x = y; y = tmp;
    Ns <<= 2;
  }
  if (Ns < NC) {
    const int Th = NC >> 1;
    for (int j = tid; j < Th; j += FE_THREADS) {
      int k = j & (Ns - 1);
      float sn, cs;
      __sincosf(-3.14159265358979f * (float)k / (float)Ns, &sn, &cs);  // |angle| < pi
      float2 u0 = x[j], u1 = cmul(x[j + Th], make_float2(cs, sn));
      int j0 = ((j - k) << 1) + k;
      y[j0] = make_float2(u0.x + u1.x, u0.y + u1.y);
      y[j0 + Ns] = make_float2(u0.x - u1.x, u0.y - u1.y);
    }
    __syncthreads();
    float2* tmp = x; x = y; y = tmp;
  }
  // ---- unpack to the one-sided spectrum of the real signal; amplitude = sqrt(max(|X|^2, 1e-10))
  for (int k = tid; k <= NC; k += FE_THREADS) {
    float2 zk = x[k & (NC - 1)];
    float2 zc = x[(NC - k) & (NC - 1)];
    zc.y = -zc.y;
    float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    float2 o = make_float2(zk.x - zc.x, zk.y - zc.y);
    float sn, cs;
    __sincosf(-3.14159265358979f * (float)k / (float)NC, &sn, &cs);  // exp(-2 pi i k / n_fft); |angle| <= pi
    float2 r = cmul(make_float2(cs, sn), o);     // times -i/2: (re,im) -> (im/2, -re/2)
    float re = e.x + 0.5f * r.y, im = e.y - 0.5f * r.x;
    float p = re * re + im * im;
    amp[k] = sqrtf(fmaxf(p, 1.0e-10f));
  }
  __syncthreads();
  // ---- mel projection over each filter's non-zero bin range, then log10(max(., 1e-10))
  const int lane = tid & 31, warp = tid >> 5;
  const int nb = NC + 1;
  for (int m = warp; m < n_mels; m += FE_THREADS / 32) {
    int lo = mel_range ? mel_range[2 * m] : 0, hi = mel_range ? mel_range[2 * m + 1] : nb;
    float acc = 0.f;
    for (int k = lo + lane; k < hi; k += 32) acc = fmaf(amp[k], melmat[(int64_t)k * n_mels + m], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[m] = log10f(fmaxf(acc, 1.0e-10f));
  }
}

__global__ void olens_kernel(const int64_t* __restrict__ ilens, int64_t* __restrict__ olens, int B, int64_t N,
                             int win_length, int hop) {
  A3T_PDL_TRIGGER();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int64_t il = ilens ? ilens[b] : N;
  olens[b] = (il + 2 * (win_length / 2) - win_length) / hop + 1;
}

// ---- collate integer math -----------------------------------------------------------------
__global__ void align_to_frames_kernel(const float* __restrict__ t_sec, int32_t* __restrict__ frames, int64_t n,
                                       float fs, float hop) {
  A3T_PDL_TRIGGER();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // torch.floor(fs * t / hop).int(): two correctly-rounded fp32 ops, in this order
  float v = __fdiv_rn(__fmul_rn(fs, t_sec[i]), hop);
  frames[i] = (int32_t)floorf(v);
}

// one CTA per utterance; phones applied in order so later phones overwrite earlier ones
__global__ void __launch_bounds__(256) expand_phone_mask_kernel(
    const uint8_t* __restrict__ phone_mask, const int32_t* __restrict__ align_start,
    const int32_t* __restrict__ align_end, const int64_t* __restrict__ align_len,
    const uint8_t* __restrict__ speech_valid, uint8_t* __restrict__ masked_position, int Ts, int Tt) {
  A3T_PDL_TRIGGER();
  const int b = blockIdx.x;
  uint8_t* mp = masked_position + (int64_t)b * Ts;
  for (int t = threadIdx.x; t < Ts; t += blockDim.x) mp[t] = 0;
  __syncthreads();
  int L = (int)align_len[b];
  if (L > Tt) L = Tt;
  // python slice semantics mp[s:e] = 1 with clamping (negative indices are not produced by the
  // reference: aligns are floor(fs*t/hop) with t >= 0)
  for (int j = 0; j < L; j++) {
    if (!phone_mask[(int64_t)b * Tt + j]) continue;
    int s = align_start[(int64_t)b * Tt + j], e = align_end[(int64_t)b * Tt + j];
    if (s < 0) s = 0;
    if (e > Ts) e = Ts;
    for (int t = s + threadIdx.x; t < e; t += blockDim.x) mp[t] = 1;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < Ts; t += blockDim.x) mp[t] = mp[t] & (speech_valid ? speech_valid[(int64_t)b * Ts + t] : 1);
}

__global__ void __launch_bounds__(256) segment_pos_kernel(const int32_t* __restrict__ align_start,
                                                          const int32_t* __restrict__ align_end,
                                                          const int64_t* __restrict__ align_len,
                                                          int64_t* __restrict__ speech_seg,
                                                          int64_t* __restrict__ text_seg, int Ts, int Tt) {
  A3T_PDL_TRIGGER();
  const int b = blockIdx.x;
  int64_t* sp = speech_seg + (int64_t)b * Ts;
  int64_t* tp = text_seg + (int64_t)b * Tt;
  for (int t = threadIdx.x; t < Ts; t += blockDim.x) sp[t] = 0;
  int L = (int)align_len[b];
  if (L > Tt) L = Tt;
  for (int j = threadIdx.x; j < Tt; j += blockDim.x) tp[j] = j < L ? j + 1 : 0;
  __syncthreads();
  for (int j = 0; j < L; j++) {
    int s = align_start[(int64_t)b * Tt + j], e = align_end[(int64_t)b * Tt + j];
    if (s < 0) s = 0;
    if (e > Ts) e = Ts;
    for (int t = s + threadIdx.x; t < e; t += blockDim.x) sp[t] = j + 1;
    __syncthreads();  // keep the overwrite order of overlapping phones
  }
}

}  // namespace a3t

using namespace a3t;

namespace a3t {
int stft_logmel_regfft(const float* wav, const int64_t* ilens, const float* window, const float* melmat,
                       const int32_t* mel_range, float* mel, int B, int64_t N, int T, int n_fft, int win_length, int hop,
                       int n_mels, cudaStream_t st);
}

extern "C" int a3t_stft_logmel(const float* wav, const int64_t* ilens, const float* window, const float* melmat,
                               const int32_t* mel_range, float* mel, int64_t* olens, int B, int64_t N, int n_fft,
                               int win_length, int hop, int n_mels, void* stream) {
  A3T_REQUIRE(wav && window && melmat && mel, "stft_logmel: null pointer");
  A3T_REQUIRE(n_fft >= 256 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "stft_logmel: n_fft=%d must be a power of two in [256,4096]", n_fft);
  A3T_REQUIRE(win_length > 0 && win_length <= n_fft && hop > 0 && n_mels > 0, "stft_logmel: bad window/hop/mels");
  A3T_REQUIRE(N > n_fft / 2, "stft_logmel: input of %lld samples is too short for reflect padding %d", (long long)N, n_fft / 2);
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return A3T_OK;
  const int T = (int)(1 + N / hop);
  const int NC = n_fft / 2;
  // warp-per-frame register FFT for the recipe sizes (n_fft 2048 / 1024); the shared-memory Stockham kernel otherwise
  int rc = stft_logmel_regfft(wav, ilens, window, melmat, mel_range, mel, B, N, T, n_fft, win_length, hop, n_mels, st);
  if (rc == A3T_ERR_UNSUPPORTED) {
    size_t smem = (size_t)2 * NC * sizeof(float2) + (size_t)(NC + 1) * sizeof(float);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(stft_logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stft_logmel_kernel<<<B * T, FE_THREADS, smem, st>>>(wav, ilens, window, melmat, mel_range, mel, B, N, T, n_fft,
                                                        win_length, hop, n_mels);
    rc = check_launch("stft_logmel");
  }
  if (rc) return rc;
  if (olens) {
    olens_kernel<<<(B + 127) / 128, 128, 0, st>>>(ilens, olens, B, N, win_length, hop);
    rc = check_launch("stft_olens");
  }
  return rc;
}

extern "C" int a3t_align_to_frames(const float* t_sec, int32_t* frames, int64_t n, float fs, float hop, void* stream) {
  A3T_REQUIRE(t_sec && frames, "align_to_frames: null pointer");
  if (n == 0) return A3T_OK;
  align_to_frames_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(t_sec, frames, n, fs, hop);
  return check_launch("align_to_frames");
}

extern "C" int a3t_expand_phone_mask(const uint8_t* phone_mask, const int32_t* align_start, const int32_t* align_end,
                                     const int64_t* align_len, const uint8_t* speech_valid, uint8_t* masked_position,
                                     int B, int Ts, int Tt, void* stream) {
  A3T_REQUIRE(phone_mask && align_start && align_end && align_len && masked_position, "expand_phone_mask: null pointer");
  if (B == 0) return A3T_OK;
  expand_phone_mask_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(phone_mask, align_start, align_end, align_len,
                                                                speech_valid, masked_position, Ts, Tt);
  return check_launch("expand_phone_mask");
}

extern "C" int a3t_segment_pos(const int32_t* align_start, const int32_t* align_end, const int64_t* align_len,
                               int64_t* speech_seg, int64_t* text_seg, int B, int Ts, int Tt, void* stream) {
  A3T_REQUIRE(align_start && align_end && align_len && speech_seg && text_seg, "segment_pos: null pointer");
  if (B == 0) return A3T_OK;
  segment_pos_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(align_start, align_end, align_len, speech_seg, text_seg, Ts, Tt);
  return check_launch("segment_pos");
}
