#!/usr/bin/env python
"""CUDA-graph-replayed timing of single GEMM cases (no host launch overhead). usage: bench_gemm2.py case [case...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a3t_b200 import _lib
from a3t_b200.backend import CudaBackend

def g(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)

be = CudaBackend("cuda:0", torch.bfloat16, seed=1, impl=_lib.IMPL_TC)
B, S, D, FF, H = 16, 1152, 384, 1536, 2
M = B * S
x = g(B, S, D); u = g(B, S, FF); res = g(B, S, D, dtype=torch.float32)
w1 = be.pack_weight(g(FF, D, 3, dtype=torch.float32, scale=0.03)); w2 = be.pack_weight(g(D, FF, 3, dtype=torch.float32, scale=0.03))
wq = be.pack_weight(g(4 * D, D, dtype=torch.float32, scale=0.05)); wo = be.pack_weight(g(D, D, dtype=torch.float32, scale=0.05))
b1 = g(FF, dtype=torch.float32); b2 = g(D, dtype=torch.float32); bq = g(4 * D, dtype=torch.float32)
qkv4 = g(B, S, 4 * D, scale=0.5); pp = g(S, D, scale=0.5); Pd = g(B, H, S, S, scale=0.01)
F3 = 2 * M * FF * 3 * D
cases = {
    "w1": (F3, lambda: be.conv_fwd(x, w1, b1)),
    "w1drop": (F3, lambda: be.conv_fwd(x, w1, b1, relu=True, drop=(0.2, 1))),
    "w2": (F3, lambda: be.conv_fwd(u, w2, b2, drop=(0.2, 2), residual=res, out_scale=0.5)),
    "w2plain": (F3, lambda: be.conv_fwd(u, w2, b2)),
    "w1dgrad": (F3, lambda: be.conv_dgrad(u, w1)),
    "w2dgrad": (F3, lambda: be.conv_dgrad(x, w2, mask=u, mask_scale=1.25)),
    "w1wgrad": (F3, lambda: be.conv_wgrad(u, x, 3)),
    "w2wgrad": (F3, lambda: be.conv_wgrad(x, u, 3)),
    "qkv4": (2 * M * 4 * D * D, lambda: be.conv_fwd(x, wq, bq)),
    "out": (2 * M * D * D, lambda: be.conv_fwd(x, wo, b2, drop=(0.2, 3), residual=res)),
    "qkv4wgrad": (2 * M * 4 * D * D, lambda: be.conv_wgrad(qkv4, x, 1)),
    "scores": (4 * B * H * S * S * (D // H), lambda: be.attn_scores_fwd(qkv4, pp, H)),
    "pv": (2 * B * H * S * S * (D // H), lambda: be.attn_pv_fwd(Pd, qkv4, H)),
}
names = sys.argv[1:] or list(cases)
for name in names:
    flops, fn = cases[name]
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"{name:12s} {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s", flush=True)
