#!/usr/bin/env python
"""Sweep of tile configurations for the short-K (k=1) GEMMs of the step, tuning build only (make TUNING=1):
A3T_TC_CTA (1 = single CTA, 2 = CTA pair), A3T_TC_BN (tile width).  Each case is timed by CUDA-graph replay and its
result compared with the default configuration's.  (profiles/r02_gemm_k1_bstationary_sweep.txt is the output of this
sweep with an experimental B-stationary main loop, A3T_TC_BSTAT=1, that was measured and not kept.)"""
import itertools, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from a3t_b200 import _lib
from a3t_b200.backend import CudaBackend


def g(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).to(dtype)


be = CudaBackend("cuda:0", torch.bfloat16, seed=1, impl=_lib.IMPL_TC)
B, S, D, FF = int(os.environ.get("K1_B", 16)), int(os.environ.get("K1_S", 1152)), 384, 1536
M = B * S
x = g(B, S, D); x2 = g(B, S, 2 * D); x4 = g(B, S, 4 * D); res = g(B, S, D, dtype=torch.float32)
wq = be.pack_weight(g(4 * D, D, dtype=torch.float32, scale=0.05)); wo = be.pack_weight(g(D, D, dtype=torch.float32, scale=0.05))
wp1 = be.pack_weight(g(2 * D, D, dtype=torch.float32, scale=0.05))
b2 = g(D, dtype=torch.float32); bq = g(4 * D, dtype=torch.float32); bp1 = g(2 * D, dtype=torch.float32)
cases = {
    "qkv4   N1536 K384 bf16": (2 * M * 4 * D * D, lambda: be.conv_fwd(x, wq, bq)),
    "pw1    N768  K384 bf16": (2 * M * 2 * D * D, lambda: be.conv_fwd(x, wp1, bp1)),
    "out    N384  K384 f32 res drop": (2 * M * D * D, lambda: be.conv_fwd(x, wo, b2, drop=(0.2, 3), residual=res)),
    "dgrad  N384  K384 f32": (2 * M * D * D, lambda: be.conv_dgrad(x, wo, out_dtype=torch.float32)),
    "dgrad  N384  K384 bf16": (2 * M * D * D, lambda: be.conv_dgrad(x, wo)),
    "dgrad  N384  K768 bf16": (2 * M * 2 * D * D, lambda: be.conv_dgrad(x2, wp1)),
    "dgrad  N384  K1536 f32": (2 * M * 4 * D * D, lambda: be.conv_dgrad(x4, wq, out_dtype=torch.float32)),
}


def timed(fn):
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
    return e0.elapsed_time(e1) / 50


def setenv(**kw):
    for k in ("A3T_TC_BSTAT", "A3T_TC_CTA", "A3T_TC_BN"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)


configs = [dict()]
for cta, bn in itertools.product((1, 2), (64, 96, 128, 192, 256)):
    configs.append(dict(A3T_TC_CTA=cta, A3T_TC_BN=bn))
for name, (flops, fn) in cases.items():
    setenv()
    ref = fn().float()
    scale = ref.abs().max().item()
    print(f"== {name}  (M={M})", flush=True)
    for cfg in configs:
        setenv(**cfg)
        try:
            y = fn().float()
            err = (y - ref).abs().max().item() / scale
            ms = timed(fn)
            tag = " ".join(f"{k[7:]}={v}" for k, v in cfg.items()) or "default"
            print(f"   {tag:28s} {ms*1e3:7.1f} us {flops/ms/1e9:7.0f} TF/s  relerr {err:.1e}", flush=True)
        except Exception as e:  # noqa: BLE001  (an unsupported forced shape raises; the sweep goes on)
            print(f"   {cfg}: {str(e)[:80]}", flush=True)
setenv()
