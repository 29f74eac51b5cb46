"""ParallelWaveGAN generator on hand-written CUDA kernels.

Parameter layout (`state_dict` keys) and arithmetic of
`espnet2.gan_tts.parallel_wavegan.ParallelWaveGANGenerator` (parallel_wavegan.py:26-248) with
`ConvInUpsampleNetwork` (upsample.py:108-189) and `ResidualBlock` (wavenet/residual_block.py:
17-169), weight-norm already removed (parallel_wavegan_pretrained_vocoder.py:43-44), plus the
callable wrapper `sedit_inference.py` uses (`vocoder(feats[T,80]) -> wav[T*hop]`, `.fs`).
Batched: `generate(c (B,80,T), z (B,1,T*hop))`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


class ParallelWaveGANGenerator(nn.Module):
    def __init__(self, in_channels: int = 1, out_channels: int = 1, kernel_size: int = 3, layers: int = 30,
                 stacks: int = 3, residual_channels: int = 64, gate_channels: int = 128, skip_channels: int = 64,
                 aux_channels: int = 80, aux_context_window: int = 2, dropout_rate: float = 0.0, bias: bool = True,
                 use_weight_norm: bool = False, upsample_conditional_features: bool = True,
                 upsample_net: str = "ConvInUpsampleNetwork", upsample_params: Optional[dict] = None):
        super().__init__()
        upsample_params = dict(upsample_params or {"upsample_scales": [4, 4, 4, 4]})
        if not (in_channels == 1 and out_channels == 1 and kernel_size == 3 and residual_channels == 64
                and gate_channels == 128 and skip_channels == 64 and aux_channels == 80 and bias
                and upsample_conditional_features and upsample_net == "ConvInUpsampleNetwork"
                and dropout_rate == 0.0 and layers % stacks == 0):
            raise NotImplementedError("a3t_b200 builds the parallel_wavegan.v1 generator shape only "
                                      "(1->64/128/64 channels, aux 80, kernel 3, ConvInUpsampleNetwork)")
        self.layers, self.stacks, self.aux_context_window = layers, stacks, aux_context_window
        self.upsample_scales = list(upsample_params["upsample_scales"])
        self.upsample_factor = int(np.prod(self.upsample_scales))
        R, G, S, A = residual_channels, gate_channels, skip_channels, aux_channels
        k = 2 * aux_context_window + 1
        # parameters with the reference's names ------------------------------------------------
        self.first_conv = nn.Conv1d(1, R, 1)
        self.upsample_net = nn.Module()
        self.upsample_net.conv_in = nn.Conv1d(A, A, k, bias=False)
        self.upsample_net.upsample = nn.Module()
        ups = []
        for s in self.upsample_scales:
            ups += [nn.Identity(), nn.Conv2d(1, 1, (1, 2 * s + 1), padding=(0, s), bias=False)]
        self.upsample_net.upsample.up_layers = nn.ModuleList(ups)
        self.conv_layers = nn.ModuleList()
        for _ in range(layers):
            blk = nn.Module()
            blk.conv = nn.Conv1d(R, G, 3, padding=1)
            blk.conv1x1_aux = nn.Conv1d(A, G, 1, bias=False)
            blk.conv1x1_out = nn.Conv1d(G // 2, R + S, 1)
            self.conv_layers.append(blk)
        self.last_conv_layers = nn.ModuleList([nn.ReLU(), nn.Conv1d(S, S, 1), nn.ReLU(), nn.Conv1d(S, 1, 1)])
        self._packed = None
        self._packed_tc = None
        self.use_tensor_cores = True   # False: the fp32 CUDA-core residual block (csrc/pwg.cu)
        self.tc_passes = 3             # 3: full split-fp16 products; 2: weights as single fp16 (see include/a3t_b200.h)

    # K-major weight packs for the fused residual-block kernel (rebuilt when parameters change)
    def _packs(self):
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and self._packed[0] == sig:
            return self._packed[1]
        packs = []
        for blk in self.conv_layers:
            wc = blk.conv.weight.detach()                       # (G, R, 3)
            wa = blk.conv1x1_aux.weight.detach()[:, :, 0]       # (G, A)
            w_in_t = torch.cat([wc[:, :, 0].t(), wc[:, :, 1].t(), wc[:, :, 2].t(), wa.t()], 0).contiguous().float()
            w_out_t = blk.conv1x1_out.weight.detach()[:, :, 0].t().contiguous().float()   # (G/2, R+S)
            packs.append((w_in_t, blk.conv.bias.detach().float().contiguous(), w_out_t,
                          blk.conv1x1_out.bias.detach().float().contiguous()))
        self._packed = (sig, packs)
        return packs

    @staticmethod
    def _split16(w: torch.Tensor):
        """fp32 -> (hi, lo) fp16 with hi + lo = w to 22 mantissa bits."""
        hi = w.to(torch.float16)
        lo = (w - hi.float()).to(torch.float16)
        return hi.contiguous(), lo.contiguous()

    def _packs_tc(self):
        """Split-fp16 K-major weight packs of the tensor-core residual block (csrc/pwg_tc.cu):
        W1 (128, 320) = [tap0 | tap1 | tap2 | aux 80 | 48 zeros] stored as 5 contiguous (128, 64) K chunks, W2 (128, 64)."""
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed_tc is not None and self._packed_tc[0] == sig:
            return self._packed_tc[1]
        packs = []
        for blk in self.conv_layers:
            wc = blk.conv.weight.detach().float()                       # (128, 64, 3)
            wa = blk.conv1x1_aux.weight.detach().float()[:, :, 0]       # (128, 80)
            w1 = torch.zeros(128, 320, dtype=torch.float32, device=wc.device)
            w1[:, 0:64], w1[:, 64:128], w1[:, 128:192], w1[:, 192:272] = wc[:, :, 0], wc[:, :, 1], wc[:, :, 2], wa
            w2 = blk.conv1x1_out.weight.detach().float()[:, :, 0].contiguous()   # (128, 64): rows 0..63 residual, 64..127 skip
            w1 = w1.view(128, 5, 64).permute(1, 0, 2).contiguous()      # chunk-major (5, 128, 64): each K chunk contiguous
            packs.append(self._split16(w1) + self._split16(w2)
                         + (blk.conv.bias.detach().float().contiguous(), blk.conv1x1_out.bias.detach().float().contiguous()))
        self._packed_tc = (sig, packs)
        return packs

    @torch.no_grad()
    def generate(self, c: torch.Tensor, z: Optional[torch.Tensor] = None) -> torch.Tensor:
        """c (B, 80, T_feats) fp32, z (B, 1, T_wav) noise -> (B, 1, T_wav); parallel_wavegan.py:136-173."""
        if c.device.type != "cuda":
            raise _lib.A3TError("a3t_b200 ParallelWaveGAN runs on CUDA tensors only (no CPU fallback)")
        c = c.contiguous().float()
        B, A, T = c.shape
        dev = c.device
        st = torch.cuda.current_stream(dev).cuda_stream
        Tw = T * self.upsample_factor
        if z is None:
            z = torch.randn(B, 1, Tw, device=dev)
        z = z.contiguous().float()
        assert z.shape == (B, 1, Tw)
        # conv_in: ReplicationPad1d(ctx) + Conv1d(k=2ctx+1, no bias)  == 'replicate' padded same-conv
        k = 2 * self.aux_context_window + 1
        cc = torch.empty_like(c)
        _lib.call("a3t_pwg_conv1d", c.data_ptr(), self.upsample_net.conv_in.weight.data_ptr(), None, cc.data_ptr(), B,
                  A, A, T, k, 1, 1, 0, 1.0, st)
        Tc = T
        for i, s in enumerate(self.upsample_scales):
            w = self.upsample_net.upsample.up_layers[2 * i + 1].weight.reshape(-1).contiguous()
            nxt = torch.empty(B, A, Tc * s, device=dev)
            _lib.call("a3t_pwg_upsample", cc.data_ptr(), w.data_ptr(), nxt.data_ptr(), B * A, Tc, s, st)
            cc, Tc = nxt, Tc * s
        x = torch.empty(B, 64, Tw, device=dev)
        _lib.call("a3t_pwg_conv1d", z.data_ptr(), self.first_conv.weight.data_ptr(), self.first_conv.bias.data_ptr(),
                  x.data_ptr(), B, 1, 64, Tw, 1, 1, 0, 0, 1.0, st)
        skip = torch.empty(B, 64, Tw, device=dev)
        per = self.layers // self.stacks
        if self.use_tensor_cores:
            # tcgen05 residual blocks on split-fp16 planes (hi + lo = 22 mantissa bits), channels-last (B, T, C)
            planes = lambda ch_: (torch.empty(B, Tw, ch_, dtype=torch.float16, device=dev),
                                  torch.empty(B, Tw, ch_, dtype=torch.float16, device=dev))
            ch, cl = planes(80)
            _lib.call("a3t_pwg_split_planes", cc.data_ptr(), ch.data_ptr(), cl.data_ptr(), B, 80, Tw, st)
            del cc
            xh, xl = planes(64)
            _lib.call("a3t_pwg_split_planes", x.data_ptr(), xh.data_ptr(), xl.data_ptr(), B, 64, Tw, st)
            del x
            yh, yl = planes(64)   # ping-pong: a block reads the halo of neighbouring tiles
            for l, (w1h, w1l, w2h, w2l, b1, b2) in enumerate(self._packs_tc()):
                _lib.call("a3t_pwg_resblock_tc", xh.data_ptr(), xl.data_ptr(), ch.data_ptr(), cl.data_ptr(), w1h.data_ptr(),
                          w1l.data_ptr(), w2h.data_ptr(), w2l.data_ptr(), b1.data_ptr(), b2.data_ptr(), yh.data_ptr(),
                          yl.data_ptr(), skip.data_ptr(), B, Tw, 2 ** (l % per), int(l == 0) | (2 if self.tc_passes == 2 else 0), st)
                xh, xl, yh, yl = yh, yl, xh, xl
        else:
            x2 = torch.empty_like(x)
            for l, (w_in_t, b_in, w_out_t, b_out) in enumerate(self._packs()):
                _lib.call("a3t_pwg_resblock", x.data_ptr(), cc.data_ptr(), w_in_t.data_ptr(), b_in.data_ptr(),
                          w_out_t.data_ptr(), b_out.data_ptr(), x2.data_ptr(), skip.data_ptr(), B, Tw, 64, 128, 80, 64,
                          2 ** (l % per), int(l == 0), st)
                x, x2 = x2, x
        l1, l3 = self.last_conv_layers[1], self.last_conv_layers[3]
        out = torch.empty(B, 1, Tw, device=dev)
        _lib.call("a3t_pwg_last", skip.data_ptr(), l1.weight.data_ptr(), l1.bias.data_ptr(), l3.weight.data_ptr(),
                  l3.bias.data_ptr(), out.data_ptr(), B, Tw, math.sqrt(1.0 / self.layers), st)
        return out

    def forward(self, c: torch.Tensor, z: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.generate(c, z)

    def inference(self, c: torch.Tensor, z: Optional[torch.Tensor] = None) -> torch.Tensor:
        """c (T_feats, 80), z (T_wav, 1) -> (T_wav, 1); parallel_wavegan.py:214-229."""
        if z is not None:
            z = z.transpose(1, 0).unsqueeze(0)
        return self.generate(c.transpose(1, 0).unsqueeze(0), z).squeeze(0).transpose(1, 0)

    def load_reference_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Accepts a weight-norm-free state dict of the reference generator; legacy checkpoints with
        a separate `conv1x1_skip` are merged as the reference does (parallel_wavegan.py:231-247)."""
        sd = dict(sd)
        for l in range(self.layers):
            ks, ko = f"conv_layers.{l}.conv1x1_skip.weight", f"conv_layers.{l}.conv1x1_out.weight"
            if ks in sd:
                sd[ko] = torch.cat([sd[ko], sd.pop(ks)], 0)
                kb = f"conv_layers.{l}.conv1x1_skip.bias"
                sd[f"conv_layers.{l}.conv1x1_out.bias"] = torch.cat([sd[f"conv_layers.{l}.conv1x1_out.bias"], sd.pop(kb)], 0)
        return self.load_state_dict(sd, strict=True)


class ParallelWaveGANPretrainedVocoder(nn.Module):
    """Callable used by bin/sedit_inference.py (`vocoder(feats[T,80]) -> wav[T*hop]`, `.fs`);
    espnet2/tts/utils/parallel_wavegan_pretrained_vocoder.py:49-63.  `mean`/`scale` reproduce the
    package's `normalize_before` (feats - mean) / scale when the checkpoint carries stats."""

    def __init__(self, generator: ParallelWaveGANGenerator, fs: int = 24000, mean: Optional[torch.Tensor] = None,
                 scale: Optional[torch.Tensor] = None):
        super().__init__()
        self.vocoder = generator
        self.fs = fs
        self.normalize_before = mean is not None
        if mean is not None:
            self.register_buffer("mean", mean.float())
            self.register_buffer("scale", scale.float())

    @torch.no_grad()
    def forward(self, feats: torch.Tensor, z: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.normalize_before:
            feats = (feats - self.mean) / self.scale
        if feats.dim() == 2:
            return self.vocoder.inference(feats, z).view(-1)
        return self.vocoder.generate(feats.transpose(1, 2), z).squeeze(1)   # batched (B,T,80) -> (B,T*hop)
