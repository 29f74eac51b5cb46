#!/usr/bin/env python
"""Times the tcgen05 GEMM on the contraction shapes of the cfg2 training step (B=16, S=1152) in isolation:
CUDA events on the launching stream, 3 warm-up + 20 timed launches, operands far larger than... no: the
operands (tens of MB) mostly fit the 126 MB L2, as they do inside the real step where the producer kernel
just wrote them.  Prints TFLOP/s per shape and the fraction of MEASURED_PEAKS.json bf16_tflops (burst)."""
import json, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a3t_b200 import _lib
from a3t_b200.backend import CudaBackend

def g(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    be = CudaBackend("cuda:0", torch.bfloat16, seed=1, impl=_lib.IMPL_TC)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    B, S, D, FF, H = 16, 1152, 384, 1536, 2
    x = g(B, S, D); u = g(B, S, FF); res = g(B, S, D, dtype=torch.float32)
    w1 = be.pack_weight(g(FF, D, 3, dtype=torch.float32, scale=0.03)); w2 = be.pack_weight(g(D, FF, 3, dtype=torch.float32, scale=0.03))
    wq = be.pack_weight(g(4 * D, D, dtype=torch.float32, scale=0.05)); wo = be.pack_weight(g(D, D, dtype=torch.float32, scale=0.05))
    b1 = g(FF, dtype=torch.float32); b2 = g(D, dtype=torch.float32); bq = g(4 * D, dtype=torch.float32)
    qkv4 = g(B, S, 4 * D, scale=0.5); pp = g(S, D, scale=0.5)
    Pd = g(B, H, S, S, scale=0.01)
    rows = []
    def add(name, flops, fn):
        ms = timeit(fn)
        rows.append((name, flops / ms / 1e9, ms))
    M = B * S
    add("ffn w1 fwd  (conv3 384->1536, relu+drop)", 2 * M * FF * 3 * D, lambda: be.conv_fwd(x, w1, b1, relu=True, drop=(0.2, 1)))
    add("ffn w1 fwd  (plain epilogue)", 2 * M * FF * 3 * D, lambda: be.conv_fwd(x, w1, b1))
    add("ffn w2 fwd  (conv3 1536->384, drop+res)", 2 * M * FF * 3 * D, lambda: be.conv_fwd(u, w2, b2, drop=(0.2, 2), residual=res, out_scale=0.5))
    add("ffn w2 dgrad (384->1536, relu mask)", 2 * M * FF * 3 * D, lambda: be.conv_dgrad(x, w2, mask=u, mask_scale=1.25))
    add("ffn w1 dgrad (1536->384)", 2 * M * FF * 3 * D, lambda: be.conv_dgrad(u, w1))
    add("ffn w1 wgrad", 2 * M * FF * 3 * D, lambda: be.conv_wgrad(u, x, 3))
    add("ffn w2 wgrad", 2 * M * FF * 3 * D, lambda: be.conv_wgrad(x, u, 3))
    add("qkv4 (384->1536, bias)", 2 * M * 4 * D * D, lambda: be.conv_fwd(x, wq, bq))
    add("linear_out (384->384, drop+res)", 2 * M * D * D, lambda: be.conv_fwd(x, wo, b2, drop=(0.2, 3), residual=res))
    add("qkv4 wgrad (split-K)", 2 * M * 4 * D * D, lambda: be.conv_wgrad(qkv4, x, 1))
    add("attn scores AC+BD (fp32 out)", 2 * 2 * B * H * S * S * (D // H), lambda: be.attn_scores_fwd(qkv4, pp, H))
    add("attn PV", 2 * B * H * S * S * (D // H), lambda: be.attn_pv_fwd(Pd, qkv4, H))
    print(f"{'shape':46s} {'TFLOP/s':>9s} {'ms':>8s}  frac of {peak:.0f} (measured burst)")
    for name, tf, ms in rows:
        print(f"{name:46s} {tf:9.1f} {ms:8.3f}  {tf / peak:.3f}")

if __name__ == "__main__":
    main()
