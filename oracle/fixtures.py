"""Seeded parameter recipes shared by `oracle/make_golden.py` (which applies them to the REFERENCE modules) and the
tests (which apply them to the a3t_b200 modules).  TEST INFRASTRUCTURE ONLY.

Fixtures at the paper width (D=384) would need tens of MB of weights; instead the fixture stores the seed and both
sides regenerate identical weights with `fill_params` (torch's CPU generator is deterministic; parameters are
visited in sorted-name order so module registration order does not matter)."""
from __future__ import annotations

import math

import torch


def fill_params(module: torch.nn.Module, seed: int, scale: float = 1.0) -> None:
    """Deterministic, non-degenerate values for every parameter and BatchNorm buffer of `module`:
    matrices / conv kernels ~ N(0, 1/fan_in), LayerNorm / BatchNorm gains 1 + 0.2 N, biases 0.1 N,
    running_mean 0.1 N, running_var 1 + 0.2 U."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
            if p.dim() > 1:
                fan_in = p[0].numel() if p.dim() > 1 else p.numel()
                p.copy_(scale * torch.randn(p.shape, generator=g) / math.sqrt(max(fan_in, 1)))
            elif n.endswith("weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
        for n, b in sorted(module.named_buffers(), key=lambda kv: kv[0]):
            if n.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(1.0 + 0.2 * torch.rand(b.shape, generator=g))


def grad_probe(g: torch.Tensor, n: int = 96) -> torch.Tensor:
    """A fixed, reproducible sample of a gradient tensor: its first n/2 values and n/2 strided ones."""
    f = g.detach().reshape(-1).float().cpu()
    h = n // 2
    if f.numel() <= n:
        return f.clone()
    idx = torch.cat([torch.arange(h), torch.linspace(h, f.numel() - 1, h).long()])
    return f[idx].clone()
