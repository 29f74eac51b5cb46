"""CPU oracle for the A3T masked-mel pretraining hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch/numpy on the CPU, the arithmetic of every
SURVEY.md section-8(a) row of the reference (richardbaihe/a3t, an ESPnet fork).  It exists
to CHECK the CUDA path; nothing in `a3t_b200/` imports it.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it.

Pinning: the reference holds no golden vector for any A3T file.  The oracle is
pinned instead against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by
`oracle/make_golden.py` (imports /root/reference with the stub modules in `oracle/shims/`) and
committed under `tests/golden/`; `tests/test_oracle_golden.py` replays them.  The four
known-answer values of SURVEY 8c (span sampler, positional table, rel_shift) are checked there too.

The ops below have the SAME names and argument meaning as the methods of
`a3t_b200.backend.CudaBackend` (the tensor-level wrapper of the C-ABI), so a test can compare
them call by call.  Every op cites the reference file:line it follows (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# dropout mask: stateless counter hash shared bit-for-bit with csrc/common.cuh::keep_mask
# (the reference uses torch's Philox stream, which cannot be matched; SURVEY 7 "Dropout")
# --------------------------------------------------------------------------------------


def keep_mask(numel: int, p: float, seed: int, site: int) -> torch.Tensor:
    """Bool keep-mask over a contiguous tensor of `numel` elements (True = kept)."""
    if p <= 0.0:
        return torch.ones(numel, dtype=torch.bool)
    with np.errstate(over="ignore"):
        idx = np.arange(numel, dtype=np.uint64)
        lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi = (idx >> np.uint64(32)).astype(np.uint32)
        f = lo ^ (hi * np.uint32(0x85EBCA6B))
        k0 = np.uint32(seed & 0xFFFFFFFF) + np.uint32(site) * np.uint32(0x9E3779B9)
        k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
        x = ((f >> np.uint32(1)) ^ k0) * np.uint32(0x9E3779B1) + k1   # one hash per element PAIR
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        r16 = np.where((f & np.uint32(1)) != 0, x >> np.uint32(16), x & np.uint32(0xFFFF))
    thr = np.uint32(np.float32(p) * np.float32(65536.0))
    return torch.from_numpy(r16 >= thr)


def _drop(x: torch.Tensor, drop) -> torch.Tensor:
    """drop = None | (p, seed, site); scales kept elements by fp32 1/(1-p)."""
    if drop is None or drop[0] <= 0.0:
        return x
    p, seed, site = drop
    keep = keep_mask(x.numel(), p, seed, site).view(x.shape)
    inv = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
    return x * keep.to(x.dtype) * inv


# --------------------------------------------------------------------------------------
# a5  NewMaskInputLayer  (espnet2/asr/encoder/mlm_encoder.py:67-70)
# --------------------------------------------------------------------------------------


def mask_input_fwd(speech, masked_position, mask_feature):
    m = masked_position.unsqueeze(-1)
    return speech.masked_fill(m, 0.0) + mask_feature.view(1, 1, -1).expand_as(speech).masked_fill(~m, 0.0)


# --------------------------------------------------------------------------------------
# a7.1 / a9  1-D convolutions and linears in channels-last form
#   MultiLayeredConv1d (espnet/nets/pytorch_backend/transformer/multi_layer_conv.py:52-62),
#   Linear layers (transformer/attention.py:55-57,96), Postnet convs (tacotron2/decoder.py:189-238)
# --------------------------------------------------------------------------------------


def conv_fwd(x, w, bias=None, *, relu=False, drop=None, residual=None, out_scale=1.0):
    """y[b,t,n] = sum_{tap,c} x[b,t+tap-pad,c] w[n,c,tap] (+bias) -> relu -> dropout;
    out = residual + out_scale*y  (or out_scale*y).  w is (N,C,taps) or (N,C)."""
    if w.dim() == 2:
        w = w.unsqueeze(-1)
    taps = w.shape[-1]
    y = F.conv1d(x.transpose(1, 2), w, bias, padding=(taps - 1) // 2).transpose(1, 2)
    if relu:
        y = torch.relu(y)
    y = _drop(y.contiguous(), drop)
    y = out_scale * y
    if residual is not None:
        y = residual + y
    return y.contiguous()


def colsum(x):
    return x.reshape(-1, x.shape[-1]).sum(0)


# --------------------------------------------------------------------------------------
# LayerNorm  (transformer/layer_norm.py:23 eps=1e-12; conformer/encoder.py:404 eps=1e-5 + ReLU)
# --------------------------------------------------------------------------------------


def ln_fwd(x, gamma, beta, eps, *, relu=False, out_scale=1.0, drop=None):
    y = F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)
    if relu:
        y = torch.relu(y)
    y = y * out_scale
    return _drop(y.contiguous(), drop)


# --------------------------------------------------------------------------------------
# positional table  (transformer/embedding.py:56-80,147-170): row t = position 4999-t
# --------------------------------------------------------------------------------------


def legacy_rel_pos_table(T: int, d_model: int, max_len: int = 5000) -> torch.Tensor:
    L = max(T, max_len)
    position = torch.arange(L - 1, -1, -1.0, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(L, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe[:T].contiguous()


# --------------------------------------------------------------------------------------
# a6  embedding assembly  (conformer/encoder.py:526-553)
# --------------------------------------------------------------------------------------


def embed_assemble_fwd(speech_y, text, sseg, tseg, emb, seg, xscale, *, drop_speech=None, drop_text=None):
    """speech_y: prenet output already LN->ReLU->*xscale (B,Ts,D); text (B,Tt) int64.
    xs[:, :Ts] = drop(speech_y) + seg[sseg];  xs[:, Ts:] = drop(emb[text]*xscale) + seg[tseg]."""
    sp = _drop(speech_y.contiguous(), drop_speech) + seg[sseg]
    tx = _drop((emb[text] * xscale).contiguous(), drop_text) + seg[tseg]
    return torch.cat([sp, tx], dim=1).contiguous()


def scale_dropout(x, scale, drop=None):
    return _drop((x * scale).contiguous(), drop)


# --------------------------------------------------------------------------------------
# a7.2  legacy relative-position attention  (transformer/attention.py:145-209, 64-96)
# --------------------------------------------------------------------------------------


def rel_shift(x):
    """attention.py:145-159 (zero_triu=False)."""
    zero_pad = torch.zeros((*x.size()[:3], 1), dtype=x.dtype)
    x_padded = torch.cat([zero_pad, x], dim=-1)
    x_padded = x_padded.view(*x.size()[:2], x.size(3) + 1, x.size(2))
    return x_padded[:, :, 1:].view_as(x)


def attn_scores_fwd(qkv4, p, H):
    """qkv4 (B,S,4D) = [q+u | q+v | k | v]; p (S,D).  Returns AC, BDraw (B,H,S,S)."""
    B, S, D4 = qkv4.shape
    D = D4 // 4
    dk = D // H
    qu = qkv4[..., 0:D].view(B, S, H, dk).transpose(1, 2)
    qv = qkv4[..., D : 2 * D].view(B, S, H, dk).transpose(1, 2)
    k = qkv4[..., 2 * D : 3 * D].view(B, S, H, dk).transpose(1, 2)
    pp = p.view(1, S, H, dk).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(-2, -1))
    bd = torch.matmul(qv, pp.transpose(-2, -1))
    return ac.contiguous(), bd.contiguous()


def relpos_softmax_fwd(ac, bd_raw, keymask, scale, *, drop=None):
    """scores=(AC+rel_shift(BDraw))*scale; key-pad fill finfo.min; softmax; zero-fill; dropout.
    keymask (B,S) bool, True = valid key.  Returns (P, P_dropped)."""
    scores = (ac + rel_shift(bd_raw)) * scale
    m = (~keymask).view(keymask.shape[0], 1, 1, -1)
    scores = scores.masked_fill(m, float(np.finfo(np.float32).min))
    attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
    return attn.contiguous(), _drop(attn.contiguous(), drop)


def attn_pv_fwd(pd, qkv4, H):
    B, S, D4 = qkv4.shape
    D = D4 // 4
    dk = D // H
    v = qkv4[..., 3 * D :].view(B, S, H, dk).transpose(1, 2)
    x = torch.matmul(pd, v)
    return x.transpose(1, 2).contiguous().view(B, S, D)


# --------------------------------------------------------------------------------------
# a7.3  conv module  (conformer/convolution.py:67-79) and BatchNorm1d
# --------------------------------------------------------------------------------------


def glu_dwconv_fwd(u, w, bias):
    """u (B,S,2C) -> glu over channel halves -> depthwise conv k (zero 'same' pad) + bias -> (B,S,C)."""
    C = u.shape[-1] // 2
    g = F.glu(u.transpose(1, 2), dim=1)
    k = w.shape[-1]
    z = F.conv1d(g, w.view(C, 1, k), bias, padding=(k - 1) // 2, groups=C)
    return z.transpose(1, 2).contiguous()


def bn_stats(z, running_mean, running_var, nbt, momentum, eps, training):
    """BatchNorm1d statistics over (B,S) per channel, padded frames included.
    Returns (mean, rstd); in training updates the running buffers in place."""
    zz = z.reshape(-1, z.shape[-1])
    if training:
        n = zz.shape[0]
        mean = zz.mean(0)
        var = zz.var(0, unbiased=False)
        with torch.no_grad():
            running_mean.mul_(1 - momentum).add_(momentum * mean)
            running_var.mul_(1 - momentum).add_(momentum * var * (n / max(n - 1, 1)))
            nbt.add_(1)
    else:
        mean, var = running_mean.clone(), running_var.clone()
    return mean, torch.rsqrt(var + eps)


ACT_NONE, ACT_SWISH, ACT_TANH = 0, 1, 2


def bn_act_fwd(z, mean, rstd, gamma, beta, act, *, drop=None, residual=None):
    y = (z - mean) * rstd * gamma + beta
    if act == ACT_SWISH:
        y = y * torch.sigmoid(y)
    elif act == ACT_TANH:
        y = torch.tanh(y)
    y = _drop(y.contiguous(), drop)
    if residual is not None:
        y = residual + y
    return y.contiguous()


# --------------------------------------------------------------------------------------
# a10  masked L1  (espnet2/tts/sedit/sedit_model.py:320-340)
# --------------------------------------------------------------------------------------


def masked_l1_fwd(before, after, y, mask):
    l = (before - y).abs().sum(-1) + (after - y).abs().sum(-1)
    m = mask.float()
    den = m.sum() + 1e-10
    return ((l * m).sum() / den).view(1), den.view(1)


# --------------------------------------------------------------------------------------
# a1  STFT -> log-mel  (espnet2/layers/stft.py:56-124, log_mel.py:56-83, log_mel_fbank.py:88-106)
# --------------------------------------------------------------------------------------


def slaney_mel_matrix(fs, n_fft, n_mels, fmin, fmax) -> torch.Tensor:
    """(n_fft/2+1, n_mels) float32; librosa.filters.mel(htk=False).T restated (log_mel.py:37-51)."""
    import sys, os

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "shims"))
    try:
        from librosa.filters import mel as _mel  # the restatement lives in the shim
    finally:
        sys.path.pop(0)
    fmin = 0 if fmin is None else fmin
    fmax = fs / 2 if fmax is None else fmax
    return torch.from_numpy(_mel(sr=fs, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax).T.copy()).float()


def stft_logmel(wav, ilens, *, fs, n_fft, win_length, hop, n_mels, fmin, fmax, double=False):
    """wav (B,N) f32, ilens (B,) -> (mel (B,T,n_mels) f32 log10, olens)."""
    dt = torch.float64 if double else torch.float32
    window = torch.hann_window(win_length, dtype=dt)
    spec = torch.stft(wav.to(dt), n_fft=n_fft, win_length=win_length, hop_length=hop, center=True,
                      window=window, normalized=False, onesided=True, return_complex=True)
    spec = spec.transpose(1, 2)  # (B,T,F)
    power = spec.real**2 + spec.imag**2
    pad = win_length // 2
    olens = (ilens + 2 * pad - win_length) // hop + 1
    T = power.shape[1]
    tmask = torch.arange(T).unsqueeze(0) >= olens.unsqueeze(1)
    power = power.masked_fill(tmask.unsqueeze(-1), 0.0)  # stft.py:120 zero-fill before amp
    amp = torch.sqrt(torch.clamp(power, min=1.0e-10))
    mel = torch.matmul(amp, slaney_mel_matrix(fs, n_fft, n_mels, fmin, fmax).to(dt))
    mel = torch.clamp(mel, min=1e-10).log10()
    mel = mel.masked_fill(tmask.unsqueeze(-1), 0.0)
    return mel.float(), olens


# --------------------------------------------------------------------------------------
# a2-a4  alignment floor, span mask, segment positions (espnet2/train/collate_fn.py:236-446)
# --------------------------------------------------------------------------------------


def align_to_frames(t_sec: torch.Tensor, fs: int, hop: int) -> torch.Tensor:
    """collate_fn.py:236-237: floor(fs*t/hop) in float32, then int32."""
    return torch.floor(fs * t_sec.float() / hop).int()


def random_spans_noise_mask(length, mlm_prob, mean_phn_span):
    """collate_fn.py:387-446 (T5 span sampler); consumes the GLOBAL numpy RandomState."""
    num_noise = int(np.round(length * mlm_prob))
    num_noise = min(max(num_noise, 1), length - 1)
    num_spans = max(int(np.round(num_noise / mean_phn_span)), 1)
    num_nonnoise = length - num_noise

    def seg(num_items, num_segments):
        first = np.arange(num_items - 1) < (num_segments - 1)
        np.random.shuffle(first)
        ids = np.cumsum(np.pad(first, [[1, 0]]))
        return np.unique(ids, return_counts=True)[1]

    noise = seg(num_noise, num_spans)
    nonnoise = seg(num_nonnoise, num_spans)
    inter = np.reshape(np.stack([nonnoise, noise], axis=1), [num_spans * 2])
    starts = np.cumsum(inter)[:-1]
    ind = np.zeros((length,), dtype=np.int8)
    ind[starts] = True
    return np.equal(np.cumsum(ind) % 2, 1)[:length]


def draw_phone_masks(align_lengths, mlm_prob, mean_phn_span, max_phones):
    """Host-side draw of the per-utterance phone mask (B,max_phones) uint8, in the reference's
    RNG order (collate_fn.py:368-376: one random_spans_noise_mask per utterance with >=2 phones)."""
    B = len(align_lengths)
    out = np.zeros((B, max_phones), dtype=np.uint8)
    for b in range(B):
        L = int(align_lengths[b])
        if L < 2:
            continue
        out[b, :L] = random_spans_noise_mask(L, mlm_prob, mean_phn_span)
    return out


def expand_phone_mask(phone_mask, align_start, align_end, align_lengths, speech_valid, span_boundary=None):
    """collate_fn.py:346-385: masked_position[b, s_j:e_j]=1 for masked phones j, AND non-pad.
    speech_valid (B,Ts) bool.  span_boundary: list per utterance of [s0,e0,s1,e1,...] (inference)."""
    B, Ts = speech_valid.shape
    mp = np.zeros((B, Ts), dtype=np.uint8)
    for b in range(B):
        if span_boundary is not None:
            sb = span_boundary[b]
            for s, e in zip(sb[::2], sb[1::2]):
                mp[b, int(s) : int(e)] = 1
        else:
            L = int(align_lengths[b])
            for j in range(L):
                if phone_mask[b, j]:
                    mp[b, int(align_start[b, j]) : int(align_end[b, j])] = 1
    return torch.from_numpy(mp).bool() & speech_valid


def segment_pos(align_start, align_end, align_lengths, Ts, Tt, sega_emb=True):
    """collate_fn.py:330-343."""
    B = align_start.shape[0]
    sp = torch.zeros(B, Ts, dtype=torch.int64)
    tp = torch.zeros(B, Tt, dtype=torch.int64)
    if not sega_emb:
        return sp, tp
    for b in range(B):
        for j in range(int(align_lengths[b])):
            s, e = int(align_start[b, j]), int(align_end[b, j])
            sp[b, s:e] = j + 1
            tp[b, j] = j + 1
    return sp, tp


# --------------------------------------------------------------------------------------
# a13  ParallelWaveGAN generator  (espnet2/gan_tts/parallel_wavegan/{parallel_wavegan.py:136-229,
#      upsample.py:160-189}, espnet2/gan_tts/wavenet/residual_block.py:114-169)
# --------------------------------------------------------------------------------------


def pwg_generate(c, z, P, *, upsample_scales, layers=30, stacks=3, aux_context=2):
    """c (B,80,T) mel, z (B,1,T*hop) noise, P = state_dict of the in-tree ParallelWaveGANGenerator
    (weight-norm removed).  Returns wav (B,1,T*hop)."""
    k = 2 * aux_context + 1
    c = F.conv1d(F.pad(c, (aux_context, aux_context), mode="replicate"), P["upsample_net.conv_in.weight"])
    c = c.unsqueeze(1)
    for i, s in enumerate(upsample_scales):
        c = F.interpolate(c, scale_factor=(1, s), mode="nearest")
        c = F.conv2d(c, P[f"upsample_net.upsample.up_layers.{2*i+1}.weight"], padding=(0, s))
    c = c.squeeze(1)
    x = F.conv1d(z, P["first_conv.weight"], P["first_conv.bias"])
    per = layers // stacks
    skips = 0
    for l in range(layers):
        d = 2 ** (l % per)
        pre = f"conv_layers.{l}."
        h = F.conv1d(x, P[pre + "conv.weight"], P[pre + "conv.bias"], padding=d, dilation=d)
        h = h + F.conv1d(c, P[pre + "conv1x1_aux.weight"])
        ha, hb = h.split(h.shape[1] // 2, dim=1)
        g = torch.tanh(ha) * torch.sigmoid(hb)
        o = F.conv1d(g, P[pre + "conv1x1_out.weight"], P[pre + "conv1x1_out.bias"])
        o_res, o_skip = o.split(o.shape[1] // 2, dim=1)
        x = (o_res + x) * math.sqrt(0.5)
        skips = skips + o_skip
    skips = skips * math.sqrt(1.0 / layers)
    y = F.conv1d(torch.relu(skips), P["last_conv_layers.1.weight"], P["last_conv_layers.1.bias"])
    y = F.conv1d(torch.relu(y), P["last_conv_layers.3.weight"], P["last_conv_layers.3.bias"])
    return y


# --------------------------------------------------------------------------------------
# a14  trainer glue: clip + Adam + Noam  (espnet2/train/trainer.py:631-675, schedulers/noam_lr.py:58-65)
# --------------------------------------------------------------------------------------


def noam_lr(base_lr, model_size, warmup, step):
    return base_lr * model_size ** (-0.5) * min(step ** (-0.5), step * warmup ** (-1.5))


def clip_adam_step(p, g, m, v, step, lr, *, max_norm=1.0, beta1=0.9, beta2=0.999, eps=1e-8):
    """clip_grad_norm_(max_norm) followed by torch.optim.Adam (no weight decay, no amsgrad) on flat
    fp32 buffers; `step` is the 1-based optimizer step.  Returns grad norm; updates p,m,v in place."""
    norm = g.double().pow(2).sum().sqrt().float()
    coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
    gg = g * coef
    m.mul_(beta1).add_(gg, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
    bc1 = 1 - beta1**step
    bc2 = 1 - beta2**step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
    return norm
