// Mask-input layer and embedding assembly (espnet2/asr/encoder/mlm_encoder.py:67-70,
// conformer/encoder.py:539-553).  Contracts: include/a3t_b200.h.
#include "common.cuh"

namespace a3t {

template <typename TY>
__global__ void __launch_bounds__(256) mask_input_fwd_kernel(const float* __restrict__ speech,
                                                             const uint8_t* __restrict__ masked,
                                                             const float* __restrict__ mask_feature,
                                                             TY* __restrict__ y, int64_t rows, int C) {
  A3T_PDL_TRIGGER();
  int64_t n = rows * C;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t r = i / C;
    int c = (int)(i - r * C);
    // x.masked_fill(m,0) + mask_feature.masked_fill(~m,0): a select, bit-exact with the reference
    // except that -0.0 + 0.0 = +0.0 there; reproduce by adding the zero of the other branch.
    float v = masked[r] ? (0.f + mask_feature[c]) : (speech[i] + 0.f);
    y[i] = from_f32<TY>(v);
  }
}

// xs[b, t<Ts]  = dropout(speech_y[b,t]) + seg[sseg[b,t]]
// xs[b, Ts+j]  = dropout(emb[text[b,j]] * xscale) + seg[tseg[b,j]]
__global__ void __launch_bounds__(256) embed_assemble_fwd_kernel(
    const float* __restrict__ speech_y, const int64_t* __restrict__ text, const int64_t* __restrict__ sseg,
    const int64_t* __restrict__ tseg, const float* __restrict__ emb, const float* __restrict__ seg,
    float* __restrict__ xs, int B, int Ts, int Tt, int D, float xscale, float drop_p,
    const unsigned long long* __restrict__ seed, uint32_t site_speech, uint32_t site_text, int V, int nseg,
    int* __restrict__ err) {
  A3T_PDL_TRIGGER();
  Drop ds = make_drop(drop_p, seed, site_speech);
  Drop dt = make_drop(drop_p, seed, site_text);
  const int S = Ts + Tt;
  int64_t n = (int64_t)B * S * D;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int d = (int)(i % D);
    int64_t r = i / D;
    int s = (int)(r % S);
    int b = (int)(r / S);
    float v;
    if (s < Ts) {
      int64_t li = ((int64_t)b * Ts + s) * D + d;
      v = drop_apply(ds, (unsigned long long)li, speech_y[li]);
      if (sseg) {
        const int64_t id = sseg[(int64_t)b * Ts + s];
        if (id >= 0 && id < nseg) v += seg[id * D + d];
        else if (err && d == 0) atomicOr(err, 2);  // torch raises IndexError here: flag it, add nothing
      }
    } else {
      int j = s - Ts;
      int64_t li = ((int64_t)b * Tt + j) * D + d;
      int64_t tok = text[(int64_t)b * Tt + j];
      float e = 0.f;
      if (tok >= 0 && tok < V) e = emb[tok * D + d];
      else if (err && d == 0) atomicOr(err, 1);
      v = drop_apply(dt, (unsigned long long)li, e * xscale);
      if (tseg) {
        const int64_t id = tseg[(int64_t)b * Tt + j];
        if (id >= 0 && id < nseg) v += seg[id * D + d];
        else if (err && d == 0) atomicOr(err, 2);
      }
    }
    xs[i] = v;
  }
}

__global__ void __launch_bounds__(256) embed_assemble_bwd_kernel(
    const float* __restrict__ dxs, const int64_t* __restrict__ text, const int64_t* __restrict__ sseg,
    const int64_t* __restrict__ tseg, float* __restrict__ dspeech_y, float* __restrict__ demb,
    float* __restrict__ dseg, int B, int Ts, int Tt, int D, float xscale, int emb_pad, int seg_pad,
    float drop_p, const unsigned long long* __restrict__ seed, uint32_t site_speech, uint32_t site_text, int V,
    int nseg) {
  A3T_PDL_TRIGGER();
  Drop ds = make_drop(drop_p, seed, site_speech);
  Drop dt = make_drop(drop_p, seed, site_text);
  const int S = Ts + Tt;
  int64_t n = (int64_t)B * S * D;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int d = (int)(i % D);
    int64_t r = i / D;
    int s = (int)(r % S);
    int b = (int)(r / S);
    float g = dxs[i];
    if (s < Ts) {
      int64_t li = ((int64_t)b * Ts + s) * D + d;
      dspeech_y[li] = drop_apply(ds, (unsigned long long)li, g);
      if (sseg && dseg) {
        int64_t id = sseg[(int64_t)b * Ts + s];
        if (id != seg_pad && id >= 0 && id < nseg) atomicAdd(&dseg[id * D + d], g);
      }
    } else {
      int j = s - Ts;
      int64_t li = ((int64_t)b * Tt + j) * D + d;
      int64_t tok = text[(int64_t)b * Tt + j];
      if (demb && tok != emb_pad && tok >= 0 && tok < V) atomicAdd(&demb[tok * D + d], drop_apply(dt, (unsigned long long)li, g) * xscale);
      if (tseg && dseg) {
        int64_t id = tseg[(int64_t)b * Tt + j];
        if (id != seg_pad && id >= 0 && id < nseg) atomicAdd(&dseg[id * D + d], g);
      }
    }
  }
}

static int ew_blocks2(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace a3t

using namespace a3t;

extern "C" int a3t_mask_input_fwd(const float* speech, const uint8_t* masked, const float* mask_feature, void* y,
                                  int dtype_y, int64_t rows, int C, void* stream) {
  A3T_REQUIRE(speech && masked && mask_feature && y, "mask_input_fwd: null pointer");
  if (rows == 0) return A3T_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n = rows * C;
  if (dtype_y == A3T_BF16)
    mask_input_fwd_kernel<__nv_bfloat16><<<ew_blocks2(n), 256, 0, st>>>(speech, masked, mask_feature,
                                                                        (__nv_bfloat16*)y, rows, C);
  else
    mask_input_fwd_kernel<float><<<ew_blocks2(n), 256, 0, st>>>(speech, masked, mask_feature, (float*)y, rows, C);
  return check_launch("mask_input_fwd");
}

extern "C" int a3t_embed_assemble_fwd(const float* speech_y, const int64_t* text, const int64_t* sseg,
                                      const int64_t* tseg, const float* emb, const float* seg, float* xs, int B,
                                      int Ts, int Tt, int D, float xscale, float drop_p,
                                      const unsigned long long* seed, uint32_t site_speech, uint32_t site_text,
                                      int V, int nseg, int* err_flag, void* stream) {
  A3T_REQUIRE(speech_y && xs && (Tt == 0 || (text && emb)), "embed_assemble_fwd: null pointer");
  A3T_REQUIRE((sseg == nullptr && tseg == nullptr) || seg, "embed_assemble_fwd: segment ids without table");
  A3T_REQUIRE(drop_p == 0.f || seed, "embed_assemble_fwd: dropout needs a seed");
  int64_t n = (int64_t)B * (Ts + Tt) * D;
  if (n == 0) return A3T_OK;
  embed_assemble_fwd_kernel<<<ew_blocks2(n), 256, 0, (cudaStream_t)stream>>>(
      speech_y, text, sseg, tseg, emb, seg, xs, B, Ts, Tt, D, xscale, drop_p, seed, site_speech, site_text, V, nseg,
      err_flag);
  return check_launch("embed_assemble_fwd");
}

extern "C" int a3t_embed_assemble_bwd(const float* dxs, const int64_t* text, const int64_t* sseg, const int64_t* tseg,
                                      float* dspeech_y, float* demb, float* dseg, int B, int Ts, int Tt, int D,
                                      float xscale, int emb_pad, int seg_pad, float drop_p,
                                      const unsigned long long* seed, uint32_t site_speech, uint32_t site_text,
                                      int V, int nseg, void* stream) {
  A3T_REQUIRE(dxs && dspeech_y, "embed_assemble_bwd: null pointer");
  A3T_REQUIRE(drop_p == 0.f || seed, "embed_assemble_bwd: dropout needs a seed");
  int64_t n = (int64_t)B * (Ts + Tt) * D;
  if (n == 0) return A3T_OK;
  embed_assemble_bwd_kernel<<<ew_blocks2(n), 256, 0, (cudaStream_t)stream>>>(
      dxs, text, sseg, tseg, dspeech_y, demb, dseg, B, Ts, Tt, D, xscale, emb_pad, seg_pad, drop_p, seed, site_speech,
      site_text, V, nseg);
  return check_launch("embed_assemble_bwd");
}
