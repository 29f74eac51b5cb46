"""Device-side batch assembly: the reference's `MLMCollateFn` (espnet2/train/collate_fn.py:106-287)
with its hot spots moved to CUDA kernels.

Same constructor arguments and the same output dict (keys, dtypes, shapes) as the reference.
The host keeps what must stay bit-identical with the reference's RNG stream: the T5 span sampler
(`random_spans_noise_mask`, collate_fn.py:387-446) draws from numpy's GLOBAL RandomState in the
reference's order, one draw per utterance.  Everything after the draw runs on the GPU:
STFT->log-mel (`a3t_stft_logmel`), seconds->frames floor (`a3t_align_to_frames`), span expansion
(`a3t_expand_phone_mask`) and segment ids (`a3t_segment_pos`).
"""
from __future__ import annotations

import math
from typing import Collection, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def random_spans_noise_mask(length: int, mlm_prob: float, mean_phn_span: float) -> np.ndarray:
    """T5 random-span sampler, consuming np.random exactly as collate_fn.py:387-446 does
    (two in-place shuffles of a boolean vector per call)."""
    n_noise = int(np.round(length * mlm_prob))
    n_noise = min(max(n_noise, 1), length - 1)
    n_spans = max(int(np.round(n_noise / mean_phn_span)), 1)
    n_keep = length - n_noise

    def split(n_items: int, n_seg: int) -> np.ndarray:
        marks = np.arange(n_items - 1) < (n_seg - 1)
        np.random.shuffle(marks)
        seg_id = np.cumsum(np.pad(marks, [[1, 0]]))
        return np.unique(seg_id, return_counts=True)[1]

    noise_len = split(n_noise, n_spans)
    keep_len = split(n_keep, n_spans)
    starts = np.cumsum(np.stack([keep_len, noise_len], axis=1).reshape(2 * n_spans))[:-1]
    flag = np.zeros((length,), dtype=np.int8)
    flag[starts] = True
    return (np.cumsum(flag) % 2 == 1)[:length]


def draw_phone_masks(align_lengths, mlm_prob: float, mean_phn_span: float, max_phones: int) -> np.ndarray:
    """(B, max_phones) uint8: one sampler call per utterance with >= 2 phones, in batch order
    (collate_fn.py:368-376)."""
    out = np.zeros((len(align_lengths), max_phones), dtype=np.uint8)
    for b, L in enumerate(align_lengths):
        L = int(L)
        if L >= 2:
            out[b, :L] = random_spans_noise_mask(L, mlm_prob, mean_phn_span)
    return out


def align_to_frames(t_sec: torch.Tensor, fs: int, hop: int) -> torch.Tensor:
    """floor(fs * t / hop).int() in fp32 (collate_fn.py:236-237) on the GPU."""
    t = t_sec.contiguous().float()
    out = torch.empty(t.shape, dtype=torch.int32, device=t.device)
    _lib.call("a3t_align_to_frames", t.data_ptr(), out.data_ptr(), t.numel(), float(fs), float(hop), _st(t.device))
    return out


def phones_masking(xs_pad, src_mask, align_start, align_end, align_start_lengths, mlm_prob, mean_phn_span,
                   span_boundary=None):
    """collate_fn.py:346-385.  Returns (masked_position bool (B,Ts), None)."""
    B, Ts, _ = xs_pad.shape
    dev = xs_pad.device
    valid = src_mask.reshape(B, Ts).to(torch.uint8).contiguous()
    Tt = align_start.shape[1]
    lens = align_start_lengths.to(device=dev, dtype=torch.int64).contiguous()
    a_s = align_start.to(device=dev, dtype=torch.int32).contiguous()
    a_e = align_end.to(device=dev, dtype=torch.int32).contiguous()
    # branch order of the reference (collate_fn.py:353-376): mlm_prob == 1 -> mean_phn_span == 0 -> per-utterance
    # (span_boundary if given, else the sampler)
    if mlm_prob == 1.0:
        return valid.bool(), None
    if mean_phn_span == 0:
        # speech-only batches (collate_fn.py:357-361): one frame-level draw shared by the batch; a given
        # span_boundary is ignored on this branch, as in the reference
        span = min(Ts * mlm_prob // 3, 50)
        m = random_spans_noise_mask(Ts, mlm_prob, span)
        return (torch.from_numpy(m).to(dev).unsqueeze(0).expand(B, Ts) & valid.bool()).contiguous(), None
    if span_boundary is not None:
        # inference: the given frame ranges are the mask (collate_fn.py:364-366); encode them as
        # pseudo-phones so that the same kernel expands them.  zip(s[::2], s[1::2]) drops an odd trailing entry.
        sb = [list(map(int, (s.tolist() if hasattr(s, "tolist") else s))) for s in span_boundary]
        n = max(len(s) // 2 for s in sb)
        pm = np.zeros((B, max(n, 1)), dtype=np.uint8)
        st_ = np.zeros((B, max(n, 1)), dtype=np.int32)
        en_ = np.zeros((B, max(n, 1)), dtype=np.int32)
        ln = np.zeros((B,), dtype=np.int64)
        for b, s in enumerate(sb):
            k = len(s) // 2
            pm[b, :k], st_[b, :k], en_[b, :k], ln[b] = 1, s[0::2][:k], s[1::2][:k], k
        pm_d, a_s, a_e, lens = (torch.from_numpy(x).to(dev) for x in (pm, st_, en_, ln))
        Tt = pm.shape[1]
    else:
        pm = draw_phone_masks(align_start_lengths.tolist(), mlm_prob, mean_phn_span, Tt)
        pm_d = torch.from_numpy(pm).to(dev)
    out = torch.empty(B, Ts, dtype=torch.uint8, device=dev)
    _lib.call("a3t_expand_phone_mask", pm_d.data_ptr(), a_s.data_ptr(), a_e.data_ptr(), lens.data_ptr(),
              valid.data_ptr(), out.data_ptr(), B, Ts, Tt, _st(dev))
    return out.bool(), None


def get_segment_pos(speech_pad, text_pad, align_start, align_end, align_start_lengths, sega_emb):
    """collate_fn.py:330-343 -> (speech_segment_pos (B,Ts) int64, text_segment_pos (B,Tt) int64)."""
    B, Ts, _ = speech_pad.shape
    Tt = text_pad.shape[1]
    dev = speech_pad.device
    if not sega_emb:
        return (torch.zeros(B, Ts, dtype=torch.int64, device=dev), torch.zeros(B, Tt, dtype=torch.int64, device=dev))
    sp = torch.empty(B, Ts, dtype=torch.int64, device=dev)
    tp = torch.empty(B, Tt, dtype=torch.int64, device=dev)
    a_s = align_start.to(device=dev, dtype=torch.int32).contiguous()
    a_e = align_end.to(device=dev, dtype=torch.int32).contiguous()
    lens = align_start_lengths.to(device=dev, dtype=torch.int64).contiguous()
    assert a_s.shape[1] == Tt, "align_start must be padded to the text length"
    _lib.call("a3t_segment_pos", a_s.data_ptr(), a_e.data_ptr(), lens.data_ptr(), sp.data_ptr(), tp.data_ptr(), B, Ts,
              Tt, _st(dev))
    return sp, tp


def _pad_stack(arrays: List[np.ndarray], pad_value) -> torch.Tensor:
    n = max(a.shape[0] for a in arrays)
    out = np.full((len(arrays), n) + arrays[0].shape[1:], pad_value, dtype=arrays[0].dtype)
    for i, a in enumerate(arrays):
        out[i, : a.shape[0]] = a
    return torch.from_numpy(out)


class MLMCollateFn:
    """Functor with the reference's constructor (collate_fn.py:109-131).  `device` selects the GPU the
    batch is assembled on (the reference assembles on the CPU inside a DataLoader worker)."""

    def __init__(self, feats_extract, float_pad_value: Union[float, int] = 0.0, int_pad_value: int = -32768,
                 not_sequence: Collection[str] = (), mlm_prob: float = 0.8, mean_phn_span: int = 8,
                 attention_window: int = 0, pad_speech: bool = False, sega_emb: bool = False,
                 duration_collect: bool = False, device: Union[str, torch.device] = "cuda"):
        if attention_window > 0 or pad_speech or duration_collect:
            raise NotImplementedError("longformer padding / duration collection are outside the A3T hot path")
        self.feats_extract = feats_extract
        self.float_pad_value, self.int_pad_value = float_pad_value, int_pad_value
        self.not_sequence = set(not_sequence)
        self.mlm_prob, self.mean_phn_span, self.sega_emb = mlm_prob, mean_phn_span, sega_emb
        self.device = torch.device(device)

    def __call__(self, data) -> Tuple[List[str], Dict[str, torch.Tensor]]:
        uttids = [u for u, _ in data]
        data = [d for _, d in data]
        assert all(set(data[0]) == set(d) for d in data), "dict-keys mismatching"
        dev = self.device
        out = {}
        for key in data[0]:
            pad = self.int_pad_value if data[0][key].dtype.kind == "i" else self.float_pad_value
            out[key] = _pad_stack([d[key] for d in data], pad)
            if key not in self.not_sequence:
                out[key + "_lengths"] = torch.tensor([d[key].shape[0] for d in data], dtype=torch.long)
        feats, feats_lengths = self.feats_extract(out["speech"].to(dev), out["speech_lengths"].to(dev))
        mlm_prob, mean_phn_span, sega_emb = self.mlm_prob, self.mean_phn_span, self.sega_emb
        if "text" not in out:
            # speech-only batches (collate_fn.py:222-233): a one-token dummy text (-2), no alignment, one
            # frame-level span draw shared by the batch with mlm_prob 0.15, no segment ids
            text = torch.zeros_like(feats_lengths.unsqueeze(-1)) - 2
            text_lengths = (torch.zeros_like(feats_lengths) + 1).cpu()
            a_s = torch.zeros(text.shape, dtype=torch.int32, device=dev)
            a_e = torch.zeros(text.shape, dtype=torch.int32, device=dev)
            a_len = torch.zeros_like(feats_lengths).cpu()
            sega_emb, mean_phn_span, mlm_prob = False, 0, 0.15
        else:
            text, text_lengths = out["text"].to(dev), out["text_lengths"]
            fs, hop = self.feats_extract.fs, self.feats_extract.hop_length
            a_s = align_to_frames(out["align_start"].to(dev), fs, hop)
            a_e = align_to_frames(out["align_end"].to(dev), fs, hop)
            a_len = out["align_start_lengths"]
        max_slen = int(feats_lengths.max().item())
        speech_pad = feats[:, :max_slen].contiguous()
        ar_t = torch.arange(text.shape[1], device=dev)
        text_mask = (ar_t[None, :] < text_lengths.to(dev)[:, None]).unsqueeze(-2)
        speech_mask = (torch.arange(max_slen, device=dev)[None, :] < feats_lengths[:, None]).unsqueeze(-2)
        span_boundary = out.get("span_boundary")
        masked_position, _ = phones_masking(speech_pad, speech_mask, a_s, a_e, a_len, mlm_prob, mean_phn_span,
                                            span_boundary)
        sseg, tseg = get_segment_pos(speech_pad, text, a_s, a_e, a_len, sega_emb)
        return uttids, dict(speech=speech_pad, text=text, masked_position=masked_position, speech_mask=speech_mask,
                            text_mask=text_mask, speech_segment_pos=sseg, text_segment_pos=tseg,
                            speech_lengths=out["speech_lengths"], text_lengths=text_lengths)
