"""Per-tile timeline of CTA 0 of the fused attention kernels (needs a library built with `make TUNING=1`)."""
import math, sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from a3t_b200 import _lib
from a3t_b200.backend import CudaBackend
B, H, S, dk = 16, 2, 1152, 192
D = H * dk
tc = CudaBackend("cuda:0", torch.bfloat16, seed=1, impl=_lib.IMPL_TC)
g = torch.Generator().manual_seed(0)
qkv4 = torch.randn(B, S, 4 * D, generator=g).to(torch.bfloat16).cuda()
p = torch.randn(S, D, generator=g).to(torch.bfloat16).cuda()
km = torch.ones(B, S, dtype=torch.bool).cuda()
sc = 1 / math.sqrt(dk)
drop = (0.2, 3)
ctx, bd, lse = tc.attn_fwd_fused(qkv4, p, km, H, sc, drop=drop)
dq = torch.empty_like(qkv4)
dctx = torch.randn_like(ctx)
tc.attn_bwd_fused(dctx, ctx, lse, bd, qkv4, p, km, H, sc, dq, drop=drop)
torch.cuda.synchronize()
for which in ("fwd", "bwd"):
    tr = torch.zeros(8 * 256, dtype=torch.int64, device="cuda")
    _lib.call("a3t_attn_set_trace", tr.data_ptr())
    if which == "fwd":
        tc.attn_fwd_fused(qkv4, p, km, H, sc, drop=drop)
    else:
        tc.attn_bwd_fused(dctx, ctx, lse, bd, qkv4, p, km, H, sc, dq, drop=drop)
    torch.cuda.synchronize()
    _lib.call("a3t_attn_set_trace", None)
    t = tr.cpu().view(8, 256)
    t0 = int(t[t > 0].min())
    print("====", which, "(clocks since first stamp; CTA 0 = query tile 0)")
    names = {0: "softmax warp0", 1: "K issue", 2: "V issue", 3: "bias issue", 4: "S mma issued", 5: "PV/dQ mma issued", 6: "store done"}
    nt = 9 if which == "fwd" else 18
    for r in range(1, 7):
        row = [int(x) - t0 for x in t[r][: nt + 2] if x > 0]
        print(f"{names[r]:18s}", row)
    print("preamble (warp 0)", [int(x) - t0 for x in t[7][:6] if x > 0])
    print("softmax warp0 per tile: start, bias ready, S ready, after max-exchange/compute, before wait prev MMA, after, arrived")
    for k in range(nt):
        row = [int(x) - t0 if x > 0 else -1 for x in t[0][8 * k: 8 * k + 7]]
        print(k, row)
