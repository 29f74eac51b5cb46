/*
 * a3t_b200 — C ABI of the Blackwell-native A3T (alignment-aware masked-mel pretraining) hot path.
 *
 * The reference (richardbaihe/a3t, an ESPnet fork) has NO FFI: every GPU op is a PyTorch library
 * call.  This header is therefore the boundary a maintainer would bind to replace
 * those calls; each entry point cites the reference code it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding and the ESPnet-side patch.
 *
 * Conventions
 *   - plain device pointers + sizes; no torch types.  The CALLER owns every buffer (inputs,
 *     outputs, workspaces) and keeps it alive until `stream` has passed the launch.
 *   - the library never allocates or frees device memory and holds no global state apart from
 *     per-process function attributes and a small cache of TMA descriptors keyed by arguments.
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     returns A3T_OK (0) or a negative error; `a3t_last_error()` gives a thread-local message.
 *   - activations are channels-last, row-major: (B, S, C) with C contiguous.
 *   - dtype codes: A3T_F32 = 0, A3T_BF16 = 1.
 *   - dropout: `p` = drop probability, `seed` = DEVICE pointer to one uint64 (graph-capturable),
 *     `site` = id of the dropout site.  keep(idx) is a stateless hash of (seed, site, linear index
 *     of the element in its contiguous tensor); kept elements are scaled by 1/(1-p).  p == 0
 *     disables it (seed may be NULL).
 */
#ifndef A3T_B200_H_
#define A3T_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A3T_OK 0
#define A3T_ERR_ARG (-1)
#define A3T_ERR_CUDA (-2)
#define A3T_ERR_UNSUPPORTED (-3)

#define A3T_F32 0
#define A3T_BF16 1

#define A3T_ACT_NONE 0
#define A3T_ACT_SWISH 1
#define A3T_ACT_TANH 2

#define A3T_GEMM_PLAIN 0   /* C[m,n] = sum_k A[m,k] B[n,k]                                          */
#define A3T_GEMM_CONV 1    /* implicit 1-D conv over rows of A: k=(tap,c), A row m -> m+tap-pad     */
#define A3T_GEMM_WGRAD 2   /* reduction over rows: C[n',(tap,c)] = sum_r A[r,n'] B[r+tap-pad,c]     */

#define A3T_IMPL_AUTO 0    /* tcgen05 tensor-core kernel when the shape/dtype qualifies, else SIMT  */
#define A3T_IMPL_SIMT 1    /* force the fp32-accumulate CUDA-core kernel (exact-fp32 parity mode)   */
#define A3T_IMPL_TC 2      /* force tcgen05 (error if the problem does not qualify)                 */
#define A3T_IMPL_TC_PAIR 3 /* as A3T_IMPL_TC, and force cta_group::2 tiles (256 x N on a CTA pair)      */

const char* a3t_last_error(void);
int a3t_version(void);

/* Generic strided (batched) GEMM with fused epilogue.  Replaces every dense contraction of the
 * path: nn.Linear (transformer/attention.py:55-57,96,186), Conv1d k=3 FFN
 * (transformer/multi_layer_conv.py:61-62), pointwise convs (conformer/convolution.py:28-54),
 * Postnet Conv1d k=5 (tacotron2/decoder.py:189-238), QK^T / PV bmm (attention.py:198-202,90), and
 * their autograd backward (dgrad = CONV mode with flipped taps, wgrad = WGRAD mode).
 *   v   = alpha * acc + bias[n]          (bias optional, fp32)
 *   v   = relu ? max(v,0) : v
 *   v   = mask ? v * (mask[m,n] != 0) * mask_scale : v      (mask has C's strides/dtype_mask)
 *   v   = dropout(v)                      (index = ((b1*batch2+b2)*M + m)*N + n)
 *   out = (res ? res[m,n] : 0) + out_scale * v              (res fp32, own strides)
 * All strides are in ELEMENTS.  In CONV mode: K = taps*cin, rows of A are grouped in sequences of
 * `seq` rows (M % seq == 0), A(m,(tap,c)) = A[(m+tap-pad)*sa_m + c*sa_k] if the shifted row stays
 * inside the sequence else 0, B(n,(tap,c)) = B[n*sb_n + tap*sb_tap + c*sb_k].
 * In WGRAD mode: reduction index r runs over K rows grouped in sequences of `seq`;
 * A(m,r) = A[r*sa_k + m*sa_m]; B((tap,c),r) = B[(r+tap-pad)*sb_k + c*sb_n] (0 outside the sequence);
 * C(m,(tap,c)) = C[m*sc_m + tap*sc_tap + c*sc_n], N = taps*cin. */
typedef struct A3tGemmDesc {
  int32_t M, N, K;
  int32_t mode;
  int32_t taps, pad, seq, cin;
  int32_t batch1, batch2;
  int32_t dtype_a, dtype_b, dtype_c, dtype_mask;
  int32_t relu;
  int32_t impl;
  float alpha, out_scale, mask_scale;
  float drop_p;
  uint32_t drop_site;
  int32_t c_zeroed; /* 1: C already holds zeros (or a value to accumulate into): a split-K WGRAD adds its partial
                     * sums without clearing C first; 0: the library clears C itself when it splits K */
  int64_t sa_m, sa_k, sa_b1, sa_b2;
  int64_t sb_n, sb_k, sb_b1, sb_b2, sb_tap;
  int64_t sc_m, sc_n, sc_b1, sc_b2, sc_tap;
  int64_t sr_m, sr_n, sr_b1, sr_b2;
} A3tGemmDesc;

int a3t_gemm(const A3tGemmDesc* d, const void* A, const void* B, void* C, const float* bias,
             const float* res, const void* mask, const unsigned long long* seed, void* stream);

/* 1 if a3t_gemm would run this problem on the tcgen05 tensor-core kernel (shape, dtype, stride and
 * alignment rules of gemm_tc.cu), 0 if it would take the CUDA-core kernel.  No launch. */
int a3t_gemm_tc_supported(const A3tGemmDesc* d, const void* A, const void* B, void* C);
/* Number of bf16 x bf16 problems A3T_IMPL_AUTO has handed to the CUDA-core kernel since the library was
 * loaded (or since the last call with reset != 0).  A hot path asserts this stays 0. */
int a3t_gemm_fallback_count(int reset);

/* Pack an fp32 conv/linear weight (N, C, taps) into bf16 K-major operands for the tensor-core
 * kernels: fwd[n, tap*C + c] = w[n,c,tap];  dgrad[c, tap'*N + n] = w[n,c,taps-1-tap'] (either may
 * be NULL).  Weight layout: multi_layer_conv.py:33-46 (torch Conv1d). */
int a3t_pack_conv_weight(const float* w, int N, int C, int taps, void* fwd_bf16, void* dgrad_bf16,
                         void* stream);

/* Batched form of a3t_pack_conv_weight: ONE launch repacks every GEMM weight of the model after an optimizer
 * step.  `items` is a DEVICE array; an item may stack up to 4 source weights along N (the fused
 * [q+u | q+v | k | v] projection packs [Wq, Wq, Wk, Wv] without materialising the concatenation; seg_rows rows
 * per source).  tile_start = running count of 64x64 tiles
 * (ceil(N/64) * ceil(C/64) per item); total_tiles = the sum; max_taps = largest taps of any item (<= 11). */
typedef struct A3tPackItem {
  const float* w[4];
  void* fwd;    /* bf16 (N, taps*C) or NULL */
  void* dgrad;  /* bf16 (C, taps*N) or NULL */
  int32_t N, C, taps, seg_rows;
  int32_t tile_start, tiles_c;
  int32_t _pad[2];
} A3tPackItem;
int a3t_pack_conv_weights(const A3tPackItem* items, int n_items, int total_tiles, int max_taps, void* stream);
/* bias of the fused projection: out = [bq + u | bq + v | bk | bv]  (each D floats; u, v = pos_bias_{u,v} flattened,
 * transformer/attention.py:190-202 adds them to q before the two score products) */
int a3t_qkv4_bias(const float* bq, const float* bk, const float* bv, const float* u, const float* v, float* out,
                  int D, void* stream);

/* LayerNorm over the last dim (transformer/layer_norm.py:23 eps 1e-12; conformer/encoder.py:404
 * eps 1e-5 followed by ReLU, embedding.py:168 x*sqrt(D), dropout).
 * y = dropout( relu?(LN(x)) * out_scale ); mean/rstd (rows) are saved for the backward. */
int a3t_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int dtype_y,
                      float* mean, float* rstd, int64_t rows, int C, float eps, int relu,
                      float out_scale, float drop_p, const unsigned long long* seed, uint32_t site,
                      void* stream);
/* dx = (dres ? dres : 0) + LN'(dy).  dy has dtype_dy.
 * dgamma / dbeta: with `partial` (fp32 workspace of 2*nblk*C floats; nblk = a3t_layernorm_bwd_blocks(rows))
 * they are WRITTEN by a second reduction kernel; with partial == NULL they are ACCUMULATED with atomics
 * into caller-zeroed buffers (single kernel).
 * Optional fused "grad prep" of the next backward section (the residual branch that consumes dx,
 * conformer/encoder_layer.py:120-165): gnext (dtype_gnext) = dropout'(dx * gnext_scale) with the keep
 * mask of (gnext_drop_p, gnext_site), and gsum[c] += sum_rows gnext[r,c] (accumulate mode only) --
 * the bias gradient of the layer whose output that branch scaled. */
int a3t_layernorm_bwd_blocks(int64_t rows);
int a3t_layernorm_bwd(const void* dy, int dtype_dy, const float* x, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, const float* dres,
                      float* dx, float* dgamma, float* dbeta, float* partial, int64_t rows, int C,
                      int relu, float out_scale, float drop_p, const unsigned long long* seed,
                      uint32_t site, void* gnext, int dtype_gnext, float gnext_scale,
                      float gnext_drop_p, uint32_t gnext_site, float* gsum, void* stream);

/* Column sums out[c] = sum_r x[r,c] (bias gradients); `partial` = nblk*C floats workspace with
 * nblk = a3t_colsum_blocks(rows).  partial == NULL: single kernel, block sums are ACCUMULATED into
 * the caller-zeroed `out` with atomics (needs 16-byte aligned rows: C, ldx multiples of 16 bytes). */
int a3t_colsum_blocks(int64_t rows);
int a3t_colsum(const void* x, int dtype_x, float* out, float* partial, int64_t rows, int C,
               int64_t ldx, void* stream);

/* y = dropout(x * scale) elementwise; fp32 in, dtype_y out (embedding.py:168-170 for the decoder
 * embed and pos_emb; also the backward "grad prep": g = dy*scale*keep/(1-p)). */
int a3t_scale_dropout(const float* x, void* y, int dtype_y, int64_t n, float scale, float drop_p,
                      const unsigned long long* seed, uint32_t site, void* stream);

/* NewMaskInputLayer (espnet2/asr/encoder/mlm_encoder.py:67-70): y = masked ? mask_feature : x. */
int a3t_mask_input_fwd(const float* speech, const uint8_t* masked, const float* mask_feature,
                       void* y, int dtype_y, int64_t rows, int C, void* stream);
/* d mask_feature[c] = sum over masked rows of dx[r,c] (dx fp32); partial = nblk*C floats with
 * nblk = a3t_colsum_blocks(rows). */
int a3t_mask_input_bwd(const float* dx, const uint8_t* masked, float* dmask_feature,
                       float* partial, int64_t rows, int C, void* stream);

/* Embedding assembly (conformer/encoder.py:539-553):
 *   xs[b, t<Ts]   = dropout(speech_y[b,t]) + seg[sseg[b,t]]
 *   xs[b, Ts+j]   = dropout(emb[text[b,j]] * xscale) + seg[tseg[b,j]]          xs: (B, Ts+Tt, D) fp32
 * bwd: dspeech_y = keep * dxs[:, :Ts]; demb/dseg += scatter (atomicAdd into zero-initialised
 * buffers; rows `emb_pad` / `seg_pad` (torch padding_idx) receive no gradient).
 * V / nseg: table sizes.  An id outside [0, V) / [0, nseg) (torch raises IndexError) reads as a zero
 * row, receives no gradient and sets bit 0 (token) / bit 1 (segment) of *err_flag (device int, may
 * be NULL) so the host can raise without a per-call synchronisation. */
int a3t_embed_assemble_fwd(const float* speech_y, const int64_t* text, const int64_t* sseg,
                           const int64_t* tseg, const float* emb, const float* seg, float* xs,
                           int B, int Ts, int Tt, int D, float xscale, float drop_p,
                           const unsigned long long* seed, uint32_t site_speech, uint32_t site_text,
                           int V, int nseg, int* err_flag, void* stream);
int a3t_embed_assemble_bwd(const float* dxs, const int64_t* text, const int64_t* sseg,
                           const int64_t* tseg, float* dspeech_y, float* demb, float* dseg, int B,
                           int Ts, int Tt, int D, float xscale, int emb_pad, int seg_pad,
                           float drop_p, const unsigned long long* seed, uint32_t site_speech,
                           uint32_t site_text, int V, int nseg, void* stream);

/* Fused legacy relative-position attention (tcgen05; csrc/attn_fused.cu) -- replaces the matmul / rel_shift /
 * softmax / dropout / matmul chain of transformer/attention.py:190-209 and its autograd backward without
 * materialising AC, P or dP.
 *   qkv4   (B,S,4D) bf16 = [q+u | q+v | k | v]            bd_raw (B,H,S,ld) bf16 = (q+v) p^T (from a3t_gemm)
 *   keymask (B,S) uint8, 1 = valid key                     scale = 1/sqrt(D/H)
 * fwd: ctx (B,S,D) bf16 = dropout(softmax((q+u)k^T + rel_shift(bd_raw)) * scale, masked) v;  lse (B,H,S) fp32 =
 *      log2-domain log-sum-exp of the scaled scores (1e30 for a row without a valid key).
 * bwd: given dctx (B,S,D) bf16: d(q+u) -> dqkv4[..., h*dk ..] (columns [0,D) of the (B,S,4D) bf16 buffer), and the
 *      operands of the remaining contractions, each written once, (B,H,S,ld) bf16: pd = dropped probabilities,
 *      ds = d scores (scaled), dbd = rel_shift^T(ds) (= d bd_raw).  ld % 8 == 0.  delta_ws: caller-owned (B,H,S) fp32
 *      workspace (row sums of dctx * ctx per head, written by a small pre-kernel of the same call).
 * Dropout element index = ((b*H + h)*S + i)*S + j (the contiguous (B,H,S,S) view), as the unfused kernels use.
 * a3t_attn_fused_supported: 1 if (B,H,S,D) qualifies (D/H in {64,128,192}, B*H*S*S < 2^32). */
int a3t_attn_fused_supported(int B, int H, int S, int D);
/* Tuning builds (-DA3T_TUNING) only: device buffer of 8 x 256 int64 that CTA 0 of the next fused-attention
 * launches fills with clock64 stamps per warp role (NULL switches it off; release builds ignore it). */
int a3t_attn_set_trace(void* device_buffer);
int a3t_relpos_attn_fwd(const void* qkv4, const void* bd_raw, int64_t ld, const uint8_t* keymask, void* ctx,
                        float* lse, int B, int H, int S, int D, float scale, float drop_p,
                        const unsigned long long* seed, uint32_t site, void* stream);
int a3t_relpos_attn_bwd(const void* qkv4, const void* bd_raw, int64_t ld, const uint8_t* keymask,
                        const void* ctx, const void* dctx, const float* lse, float* delta_ws, void* dqkv4,
                        void* pd, void* ds, void* dbd, int B, int H, int S, int D, float scale, float drop_p,
                        const unsigned long long* seed, uint32_t site, void* stream);

/* Legacy relative-position masked softmax (transformer/attention.py:145-165 rel_shift, :205-207
 * scale, :79-86 finfo.min fill / softmax / zero fill, :88 dropout).
 *   s[i,j] = (AC[i,j] + BDraw_shifted[i,j]) * scale ; key j invalid -> finfo(f32).min
 *   P = softmax_j(s), zeroed at invalid keys;  Pd = dropout(P)   (Pd may alias P when p == 0)
 * AC/BD (B,H,S,S) of dtype_in (fp32 in the parity mode, bf16 in the tensor-core mode: same rounding as P);
 * keymask (B,S) uint8, 1 = valid; P, Pd dtype_p.  ld = row pitch in elements of all four (B,H,S,ld) tensors
 * (0 = dense, ld == S); the host side pads it to a multiple of 8 so that every score row is 16-byte aligned for
 * the TMA tensor maps of the attention contractions whatever S is. */
int a3t_relpos_softmax_fwd(const void* ac, const void* bd_raw, int dtype_in, const uint8_t* keymask, void* P,
                           void* Pd, int dtype_p, int B, int H, int S, int ld, float scale, float drop_p,
                           const unsigned long long* seed, uint32_t site, void* stream);
/* dS = P * (dPu - sum_j dPu*P) * scale with dPu = dPd*keep/(1-p);  dBD_raw = inverse rel_shift of
 * dS.  dPd dtype_in; dS, dBD dtype_o (B,H,S,ld); ld as in the forward. */
int a3t_relpos_softmax_bwd(const void* dPd, int dtype_in, const void* P, int dtype_p, void* dS, void* dBD,
                           int dtype_o, int B, int H, int S, int ld, float scale, float drop_p,
                           const unsigned long long* seed, uint32_t site, void* stream);

/* Conformer conv module core (conformer/convolution.py:70-75): GLU over channel halves of
 * u (B,S,2C) then depthwise Conv1d (kernel k, zero 'same' padding, weight (C,k), bias (C)).
 * z fp32 (B,S,C). */
int a3t_glu_dwconv_fwd(const void* u, int dtype_u, const float* w, const float* bias, float* z,
                       int B, int S, int C, int k, void* stream);
/* du (dtype_du) = GLU'(dwconv^T(dz)); dw (C,k), dbias (C) written from `partial`
 * ((k+1)*C*nblk floats, nblk = a3t_dwconv_bwd_blocks(B,S)). */
int a3t_dwconv_bwd_blocks(int B, int S);
int a3t_glu_dwconv_bwd(const float* dz, const void* u, int dtype_u, const float* w, void* du,
                       int dtype_du, float* dw, float* dbias, float* partial, int B, int S, int C,
                       int k, void* stream);

/* BatchNorm1d statistics over all rows (conformer/convolution.py:76, tacotron2/decoder.py:203;
 * padded frames included, per rank).  training: batch mean / biased var -> mean, rstd; running
 * stats updated with `momentum` and the unbiased var, num_batches_tracked += 1.  eval: mean/rstd
 * from the running stats.  partial = 2*nblk*C doubles, nblk = a3t_colsum_blocks(rows). */
int a3t_bn_stats(const float* z, float* mean, float* rstd, float* running_mean, float* running_var,
                 int64_t* num_batches_tracked, double* partial, int64_t rows, int C, float momentum,
                 float eps, int training, void* stream);
/* y = (res ? res : 0) + dropout(act(gamma*(z-mean)*rstd + beta)) ; act in A3T_ACT_*. */
int a3t_bn_act_fwd(const float* z, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, const float* res, void* y, int dtype_y, int64_t rows, int C,
                   int act, float drop_p, const unsigned long long* seed, uint32_t site,
                   void* stream);
/* Backward of bn_act_fwd w.r.t. z, gamma, beta (dy fp32).  training=1 differentiates through the
 * batch statistics.  partial = 2*nblk*C doubles (nblk = a3t_colsum_blocks(rows)); coef = 2*C floats. */
int a3t_bn_act_bwd(const float* dy, const float* z, const float* mean, const float* rstd,
                   const float* gamma, const float* beta, float* dz, float* dgamma, float* dbeta,
                   double* partial, float* coef, int64_t rows, int C, int act, int training,
                   float drop_p, const unsigned long long* seed, uint32_t site, void* stream);

/* Masked L1 (espnet2/tts/sedit/sedit_model.py:320-340):
 * loss = sum_rows mask*(|before-y|_1 + |after-y|_1) / (sum mask + 1e-10).  out[0]=loss, out[1]=den.
 * partial = 2*nblk doubles, nblk = a3t_colsum_blocks(rows).  `after`/`dafter` may be NULL (no
 * postnet).  bwd: d before = gloss/den * mask * sign(before-y) (same for after); den = &out[1]. */
int a3t_masked_l1_fwd(const float* before, const float* after, const float* y, const uint8_t* mask,
                      float* out, double* partial, int64_t rows, int C, void* stream);
int a3t_masked_l1_bwd(const float* gloss, const float* before, const float* after, const float* y,
                      const uint8_t* mask, const float* den, float* dbefore, float* dafter,
                      int64_t rows, int C, void* stream);

/* Trainer glue (espnet2/train/trainer.py:631-675; schedulers/noam_lr.py:58-65): squared L2 norm of
 * a flat fp32 gradient buffer (sq[0], double; partial = nblk doubles with nblk = 1024), then
 * clip_grad_norm_(max_norm) + Adam + Noam LR in one pass.  `step` is a device int64 holding the
 * number of optimizer steps done so far (incremented by the kernel unless the norm is non-finite,
 * in which case the update is skipped, trainer.py:640-656).  The gradient is multiplied by
 * grad_scale and, when `denom` (device float, may be NULL) is given, divided by denom[0] — the
 * all-reduced sum of per-rank batch weights (trainer.py:583-595 + DDP mean).  warmup <= 0 = constant lr. */
int a3t_grad_sqnorm(const float* g, int64_t n, double* sq, double* partial, void* stream);
int a3t_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const double* sq,
                  int64_t* step, float base_lr, float model_size, float warmup, float beta1,
                  float beta2, float eps, float max_norm, float grad_scale, const float* denom,
                  void* stream);
/* advance a device-resident dropout seed: *seed = *seed * 6364136223846793005 + 1442695040888963407 */
int a3t_seed_advance(unsigned long long* seed, void* stream);

/* STFT -> log-mel frontend (espnet2/tts/feats_extract/log_mel_fbank.py:88-106, layers/stft.py:56-124,
 * layers/log_mel.py:56-83): reflect-pad n_fft/2, frame n_fft @ hop, periodic Hann(win_length)
 * centred in n_fft, |rFFT|, sqrt(max(.,1e-10)), @ melmat (n_fft/2+1, n_mels), max(.,1e-10), log10,
 * frames >= olens[b] zeroed.  wav (B,N) fp32, ilens (B) int64, window (win_length) fp32,
 * mel (B,T,n_mels) fp32 with T = 1 + N / hop, olens (B) int64 (may be NULL).  n_fft must be a
 * power of two in [256, 4096].  mel_range (optional, 2*n_mels int32): [lo,hi) bin range outside
 * which column m of melmat is exactly zero (the filters are triangles); NULL = dense.
 * ilens NULL = every utterance is N samples long. */
int a3t_stft_logmel(const float* wav, const int64_t* ilens, const float* window, const float* melmat,
                    const int32_t* mel_range, float* mel, int64_t* olens, int B, int64_t N, int n_fft,
                    int win_length, int hop, int n_mels, void* stream);

/* Collate integer math on device (espnet2/train/collate_fn.py:236-237, :330-343, :346-385). */
int a3t_align_to_frames(const float* t_sec, int32_t* frames, int64_t n, float fs, float hop,
                        void* stream);
/* masked_position[b, s_j:e_j] = 1 for phones j < len[b] with phone_mask[b,j], AND speech_valid. */
int a3t_expand_phone_mask(const uint8_t* phone_mask, const int32_t* align_start,
                          const int32_t* align_end, const int64_t* align_len,
                          const uint8_t* speech_valid, uint8_t* masked_position, int B, int Ts,
                          int Tt, void* stream);
/* speech_segment_pos[b, s_j:e_j] = j+1 (later phones overwrite earlier), text_segment_pos[b,j]=j+1 */
int a3t_segment_pos(const int32_t* align_start, const int32_t* align_end, const int64_t* align_len,
                    int64_t* speech_seg, int64_t* text_seg, int B, int Ts, int Tt, void* stream);

/* ParallelWaveGAN generator (espnet2/gan_tts/parallel_wavegan/parallel_wavegan.py:136-229,
 * upsample.py:160-189, espnet2/gan_tts/wavenet/residual_block.py:114-169), channels-first fp32. */
/* nearest-neighbour stretch x`scale` then 1-D FIR of length 2*scale+1 (zero pad `scale`):
 * in (rows, T) -> out (rows, T*scale); w (2*scale+1). */
int a3t_pwg_upsample(const float* in, const float* w, float* out, int rows, int64_t T, int scale,
                     void* stream);
/* generic small dense Conv1d, channels-first: out[b,o,t] = bias[o] + sum_{i,k} w[o,i,k] *
 * in[b,i,t + (k - (K-1)/2)*dil] ; pad_mode 0 = zero, 1 = replicate; relu_in applies ReLU to the
 * input first.  Used for conv_in (80->80,k5), first_conv (1->64) and the two last convs. */
int a3t_pwg_conv1d(const float* in, const float* w, const float* bias, float* out, int B, int Cin,
                   int Cout, int64_t T, int K, int dil, int pad_mode, int relu_in, float in_scale,
                   void* stream);
/* one gated residual block, fused: h = dilconv3(x) + bias + aux1x1(c); g = tanh(h[:G/2])*sigmoid(h[G/2:]);
 * o = out1x1(g)+bias; x_out = (o[:R]+x)*sqrt(.5); skip += o[R:]  (skip = o[R:] if first).
 * x,x_out (B,R,T), c (B,A,T), skip (B,S_,T); x_out must not alias x.
 * Weights are passed K-major (transposed once at load time by the host):
 *   w_in_t  (3*R + A, G): row tap*R+i = conv.weight[:, i, tap], row 3*R+a = conv1x1_aux.weight[:, a, 0]
 *   b_in (G) = conv.bias;   w_out_t (G/2, R+S_): row i = conv1x1_out.weight[:, i, 0];   b_out (R+S_). */
int a3t_pwg_resblock(const float* x, const float* c, const float* w_in_t, const float* b_in,
                     const float* w_out_t, const float* b_out, float* x_out, float* skip, int B,
                     int64_t T, int R, int G, int A, int S_, int dil, int first, void* stream);

/* Fused output stack of the generator (parallel_wavegan.py:119-126, 166-173):
 *   out[b,t] = w2 . relu(W1 relu(skip[b,:,t] * scale) + b1) + b2,  skip (B,64,T) fp32, W1 (64,64), w2 (64), out (B,T). */
int a3t_pwg_last(const float* skip, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                 int B, int64_t T, float scale, void* stream);

/* The same residual block on the tensor cores (csrc/pwg_tc.cu): tcgen05 split-fp16 GEMMs (every operand as hi + lo
 * fp16 planes, three MMAs per product, fp32 accumulation in TMEM) so that the waveform stays within atol 1e-4 of the
 * fp32 reference after 30 blocks.  Activations are CHANNELS-LAST planes: xh/xl (B,T,64) in, yh/yl (B,T,64) out (must
 * not alias the input), ch/cl (B,T,80) conditioning; skip (B,64,T) fp32 channel-major (written when first != 0,
 * accumulated otherwise); w1h/w1l (5,128,64) fp16 = K chunks [tap0 | tap1 | tap2 | aux 0..63 | aux 64..79 + zeros] of
 * the (128, 320) gate weight, w2h/w2l (128,64); b1/b2 (128) fp32.
 * flags: bit 0 = first block (skip is written, not accumulated); bit 1 = two-pass mode: the weights enter as single
 * fp16 ((x_hi + x_lo) * w_hi, the lo weight planes are not read): 2/3 of the MMAs and 3/4 of the operand traffic, weight
 * rounding 2^-12 instead of 2^-23.
 * a3t_pwg_split_planes: fp32 channel-major (B,C,T) -> channels-last hi / lo fp16 planes (B,T,C), C % 8 == 0. */
int a3t_pwg_split_planes(const float* src, void* hi, void* lo, int B, int C, int64_t T, void* stream);
int a3t_pwg_resblock_tc(const void* xh, const void* xl, const void* ch, const void* cl, const void* w1h,
                        const void* w1l, const void* w2h, const void* w2l, const float* b1, const float* b2,
                        void* yh, void* yl, float* skip, int B, int64_t T, int dil, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* A3T_B200_H_ */
