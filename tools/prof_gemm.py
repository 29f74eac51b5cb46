#!/usr/bin/env python
"""Launch one GEMM case a few times (for `ncu --set full -k regex:gemm_tc`).  usage: prof_gemm.py <case>"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a3t_b200 import _lib
from a3t_b200.backend import CudaBackend

def g(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)

case = sys.argv[1] if len(sys.argv) > 1 else "w1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
be = CudaBackend("cuda:0", torch.bfloat16, seed=1, impl=_lib.IMPL_TC)
B, S, D, FF, H = 16, 1152, 384, 1536, 2
x = g(B, S, D); u = g(B, S, FF); res = g(B, S, D, dtype=torch.float32)
w1 = be.pack_weight(g(FF, D, 3, dtype=torch.float32, scale=0.03)); w2 = be.pack_weight(g(D, FF, 3, dtype=torch.float32, scale=0.03))
wq = be.pack_weight(g(4 * D, D, dtype=torch.float32, scale=0.05))
b1 = g(FF, dtype=torch.float32); b2 = g(D, dtype=torch.float32); bq = g(4 * D, dtype=torch.float32)
fns = {
    "w1": lambda: be.conv_fwd(x, w1, b1),
    "w1drop": lambda: be.conv_fwd(x, w1, b1, relu=True, drop=(0.2, 1)),
    "w2": lambda: be.conv_fwd(u, w2, b2, drop=(0.2, 2), residual=res, out_scale=0.5),
    "w1dgrad": lambda: be.conv_dgrad(u, w1),
    "w2dgrad": lambda: be.conv_dgrad(x, w2, mask=u, mask_scale=1.25),
    "w1wgrad": lambda: be.conv_wgrad(u, x, 3),
    "qkv4": lambda: be.conv_fwd(x, wq, bq),
}
for _ in range(n):
    fns[case]()
torch.cuda.synchronize()
