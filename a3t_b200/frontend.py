"""STFT -> log-mel frontend behind the reference's `AbsFeatsExtract` interface.

Mirror of `espnet2.tts.feats_extract.log_mel_fbank.LogMelFbank` (log_mel_fbank.py:21-106: same
constructor kwargs, `fs` / `hop_length` attributes, `output_size()`, `get_parameters()`,
`forward(wav, lens) -> (feats (B,T,n_mels), feats_lens)`), computed by ONE fused CUDA kernel
(`a3t_stft_logmel`: in-shared-memory FFT, |.|, mel filterbank, log10) instead of
torch.stft + matmul.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .espnet_plugin import espnet_base

# `AbsFeatsExtract` (espnet2/tts/feats_extract/abs_feats_extract.py) when the reference is importable:
# `feats_extractor_choices` type-checks against it (espnet2/tasks/mlm.py:58-67)
_FeatsBase = espnet_base("espnet2.tts.feats_extract.abs_feats_extract", "AbsFeatsExtract") or torch.nn.Module


def slaney_mel_filterbank(fs: float, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """(n_fft//2+1, n_mels) float32 triangular filterbank on the Slaney mel scale with Slaney
    area normalisation — what `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False)`
    returns, transposed (call site: espnet2/layers/log_mel.py:37-51)."""
    f_sp, brk = 200.0 / 3.0, 1000.0
    brk_mel, step = brk / f_sp, np.log(6.4) / 27.0

    def to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f < brk, f / f_sp, brk_mel + np.log(np.maximum(f, 1e-300) / brk) / step)

    def to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m < brk_mel, m * f_sp, brk * np.exp(step * (m - brk_mel)))

    edges = to_hz(np.linspace(to_mel(fmin), to_mel(fmax), n_mels + 2))           # (n_mels+2,)
    bins = np.linspace(0.0, fs / 2.0, n_fft // 2 + 1)                              # (F,)
    width = np.diff(edges)                                                         # (n_mels+1,)
    up = (bins[:, None] - edges[None, :-2]) / width[None, :-1]                     # rising slope
    down = (edges[None, 2:] - bins[:, None]) / width[None, 1:]                     # falling slope
    fb = np.maximum(0.0, np.minimum(up, down))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[None, :]
    return fb.astype(np.float32)


class LogMelFbank(_FeatsBase):
    def __init__(self, fs: Union[int, str] = 16000, n_fft: int = 1024, win_length: Optional[int] = None,
                 hop_length: int = 256, window: Optional[str] = "hann", center: bool = True,
                 normalized: bool = False, onesided: bool = True, n_mels: int = 80, fmin: Optional[int] = 80,
                 fmax: Optional[int] = 7600, htk: bool = False, log_base: Optional[float] = 10.0):
        super().__init__()
        if isinstance(fs, str):
            fs = int(float(fs.lower().replace("k", "e3"))) if not fs.isdigit() else int(fs)
        if not (window == "hann" and center and not normalized and onesided and not htk and log_base == 10.0):
            raise NotImplementedError("a3t_b200.LogMelFbank builds the A3T recipe frontend only: hann window, "
                                      "center=True, normalized=False, onesided=True, htk=False, log_base=10")
        self.fs, self.n_mels, self.n_fft, self.hop_length = fs, n_mels, n_fft, hop_length
        self.win_length = win_length if win_length is not None else n_fft
        self.window, self.fmin, self.fmax = window, fmin, fmax
        fmin_ = 0.0 if fmin is None else float(fmin)
        fmax_ = fs / 2.0 if fmax is None else float(fmax)
        fb = slaney_mel_filterbank(fs, n_fft, n_mels, fmin_, fmax_)
        nz = fb != 0
        lo = np.where(nz.any(0), nz.argmax(0), 0)
        hi = np.where(nz.any(0), fb.shape[0] - nz[::-1].argmax(0), 0)
        self.register_buffer("melmat", torch.from_numpy(fb), persistent=False)
        self.register_buffer("mel_range", torch.from_numpy(np.stack([lo, hi], 1).astype(np.int32)).contiguous(),
                             persistent=False)
        self.register_buffer("hann", torch.hann_window(self.win_length, dtype=torch.float32), persistent=False)

    def output_size(self) -> int:
        return self.n_mels

    def get_parameters(self) -> Dict[str, Any]:
        return dict(fs=self.fs, n_fft=self.n_fft, n_shift=self.hop_length, window=self.window, n_mels=self.n_mels,
                    win_length=self.win_length, fmin=self.fmin, fmax=self.fmax)

    def forward(self, input: torch.Tensor, input_lengths: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if input.device.type != "cuda":
            raise _lib.A3TError("a3t_b200.LogMelFbank runs on CUDA tensors only (no CPU fallback)")
        wav = input.contiguous().float()
        B, N = wav.shape
        dev = wav.device
        if self.melmat.device != dev:
            self.to(dev)
        T = 1 + N // self.hop_length
        mel = torch.empty(B, T, self.n_mels, dtype=torch.float32, device=dev)
        olens = torch.empty(B, dtype=torch.int64, device=dev)
        il = None if input_lengths is None else input_lengths.to(device=dev, dtype=torch.int64).contiguous()
        _lib.call("a3t_stft_logmel", wav.data_ptr(), None if il is None else il.data_ptr(), self.hann.data_ptr(),
                  self.melmat.data_ptr(), self.mel_range.data_ptr(), mel.data_ptr(), olens.data_ptr(), B, N,
                  self.n_fft, self.win_length, self.hop_length, self.n_mels,
                  torch.cuda.current_stream(dev).cuda_stream)
        return mel, olens
