#!/usr/bin/env python
"""Times the bf16 relative-position softmax kernels alone (cfg2 shape). usage: bench_softmax.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a3t_b200.backend import CudaBackend
be = CudaBackend("cuda:0", torch.bfloat16, seed=1)
B, H, S = 16, 2, 1152
ac = torch.randn(B, H, S, S, device="cuda").to(torch.bfloat16)
bd = torch.randn(B, H, S, S, device="cuda").to(torch.bfloat16)
km = torch.ones(B, S, dtype=torch.bool, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for drop in (None, (0.2, 5)):
    t = timeit(lambda: be.relpos_softmax_fwd(ac, bd, km, 0.07, drop=drop))
    print(f"fwd drop={drop}: {t:.1f} us  ({4*2*B*H*S*S/t/1e6:.2f} TB/s at 8 B/score)")
P, Pd = be.relpos_softmax_fwd(ac, bd, km, 0.07, drop=(0.2, 5))
for drop in (None, (0.2, 5)):
    t = timeit(lambda: be.relpos_softmax_bwd(ac, P, 0.07, drop=drop))
    print(f"bwd drop={drop}: {t:.1f} us")
