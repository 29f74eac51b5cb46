"""Times backend.attn_fwd_fused / attn_bwd_fused (fused kernel + the remaining batched GEMMs) at cfg2 shape."""
import math, sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from a3t_b200 import _lib
from a3t_b200.backend import CudaBackend
B, H, S, dk = 16, 2, 1152, 192
D = H * dk
impl = int(sys.argv[1]) if len(sys.argv) > 1 else _lib.IMPL_TC   # IMPL_TC_PAIR (3): CTA pairs for every GEMM (measured slower)
tc = CudaBackend("cuda:0", torch.bfloat16, seed=1, impl=impl)
g = torch.Generator().manual_seed(0)
qkv4 = torch.randn(B, S, 4 * D, generator=g).to(torch.bfloat16).cuda()
p = torch.randn(S, D, generator=g).to(torch.bfloat16).cuda()
km = torch.ones(B, S, dtype=torch.bool).cuda()
sc = 1 / math.sqrt(dk)
drop = (0.2, 3)
ctx, bd, lse = tc.attn_fwd_fused(qkv4, p, km, H, sc, drop=drop)
dctx = torch.randn_like(ctx); dq = torch.empty_like(qkv4)
def fwd(): tc.attn_fwd_fused(qkv4, p, km, H, sc, drop=drop)
def bwd(): tc.attn_bwd_fused(dctx, ctx, lse, bd, qkv4, p, km, H, sc, dq, drop=drop)
for name, f in (("fwd(all)", fwd), ("bwd(all)", bwd)):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us impl={impl}")
