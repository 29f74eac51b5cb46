"""Times the fused attention kernels alone (CUDA events, warm) at a given shape: usage time_attn.py B H S dk [drop]"""
import math, sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from a3t_b200 import _lib
from a3t_b200.backend import CudaBackend
B, H, S, dk = [int(x) for x in sys.argv[1:5]] if len(sys.argv) > 4 else (16, 2, 1152, 192)
drop = (float(sys.argv[5]), 3) if len(sys.argv) > 5 and float(sys.argv[5]) > 0 else None
D = H * dk
tc = CudaBackend("cuda:0", torch.bfloat16, seed=1, impl=_lib.IMPL_TC)
g = torch.Generator().manual_seed(0)
qkv4 = torch.randn(B, S, 4 * D, generator=g).to(torch.bfloat16).cuda()
p = torch.randn(S, D, generator=g).to(torch.bfloat16).cuda()
km = torch.ones(B, S, dtype=torch.bool).cuda()
sc = 1 / math.sqrt(dk)
ctx, bd, lse = tc.attn_fwd_fused(qkv4, p, km, H, sc, drop=drop)
dctx = torch.randn_like(ctx)
dq = torch.empty_like(qkv4)
pd, ds, dbd = tc._like(bd), tc._like(bd), tc._like(bd)
delta = torch.empty(B, H, S, dtype=torch.float32, device='cuda')
pr, seed, site = tc._drop(drop)
st = torch.cuda.current_stream().cuda_stream
kmu = km.view(torch.uint8)
def fwd():
    _lib.call("a3t_relpos_attn_fwd", qkv4.data_ptr(), bd.data_ptr(), bd.stride(2), kmu.data_ptr(), ctx.data_ptr(), lse.data_ptr(), B, H, S, D, sc, pr, seed, site, st)
def bwd():
    _lib.call("a3t_relpos_attn_bwd", qkv4.data_ptr(), bd.data_ptr(), bd.stride(2), kmu.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), pd.data_ptr(), ds.data_ptr(), dbd.data_ptr(), B, H, S, D, sc, pr, seed, site, st)
for name, f in (("fwd", fwd), ("bwd", bwd)):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us  (B={B} H={H} S={S} dk={dk} drop={drop})")
