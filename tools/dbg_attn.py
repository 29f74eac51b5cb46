import math, sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from a3t_b200 import _lib
from a3t_b200.backend import CudaBackend
B, H, S, dk = [int(x) for x in sys.argv[1:5]] if len(sys.argv) > 4 else (1, 1, 128, 64)
D = H * dk
tc = CudaBackend("cuda:0", torch.bfloat16, seed=1, impl=_lib.IMPL_TC)
g = torch.Generator().manual_seed(0)
qkv4 = torch.randn(B, S, 4 * D, generator=g).to(torch.bfloat16).cuda()
p = torch.randn(S, D, generator=g).to(torch.bfloat16).cuda()
km = torch.ones(B, S, dtype=torch.bool).cuda()
ctx, bd, lse = tc.attn_fwd_fused(qkv4, p, km, H, 1 / math.sqrt(dk))
torch.cuda.synchronize()
print("fwd ok", float(ctx.float().abs().mean()))
dq = torch.empty_like(qkv4)
dp = tc.attn_bwd_fused(torch.randn_like(ctx), ctx, lse, bd, qkv4, p, km, H, 1 / math.sqrt(dk), dq)
torch.cuda.synchronize()
print("bwd ok", float(dq.float().abs().mean()), float(dp.abs().mean()))
