#!/usr/bin/env python
"""One eager cfg2 training step with a CUDA-event pair around every a3t_gemm call: per-shape launch counts, mean
duration and TFLOP/s (sorted by total time).  usage: gemm_table.py [B]"""
import collections, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from a3t_b200 import _lib
import a3t_b200.backend as bk
from a3t_b200.model import build_model
from a3t_b200.trainer import DataParallelTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
enc, dec, mc = bench.paper_conf()
torch.manual_seed(0)
model = build_model(enc, dec, mc, act_dtype=torch.bfloat16).to(dev).train()
with torch.no_grad():
    for n, p in model.named_parameters():
        if p.dim() == 1 and n.endswith("weight"):
            p.fill_(1.0)
tr = DataParallelTrainer(model)
batch = bench.synthetic_batch(B, 1024, 128, device=dev)
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
orig = _lib.call
evs = []
def timing(name, *a):
    if name != "a3t_gemm":
        return orig(name, *a)
    d = a[0]
    key = (d.mode, d.M, d.N, d.K, d.batch1 * d.batch2, d.dtype_c, bool(a[5]), bool(a[6]), d.drop_p > 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = orig(name, *a); e1.record()
    evs.append((key, e0, e1, 2.0 * d.M * d.N * d.K * d.batch1 * d.batch2))
    return rc
_lib.call = bk.call = timing
torch.cuda._sleep(int(1.9e9 * 0.06))  # GPU stays behind the host: event pairs hold kernel time only
tr.step(batch)
torch.cuda.synchronize()
_lib.call = bk.call = orig
agg = collections.OrderedDict()
for key, e0, e1, fl in evs:
    t = e0.elapsed_time(e1)
    a = agg.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] += fl
tot_t = sum(a[1] for a in agg.values()); tot_f = sum(a[2] for a in agg.values())
print(f"total GEMM {tot_t:.2f} ms, {tot_f/1e12:.2f} TFLOP, {tot_f/tot_t/1e9:.0f} TFLOP/s")
print("mode M N K batch cdt res mask drop | n  us/launch  TFLOP/s  total_ms")
for key, (n, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(*key, "|", n, f"{1e3*t/n:8.1f} {f/t/1e9:8.0f} {t:7.3f}")
