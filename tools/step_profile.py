#!/usr/bin/env python
"""Warm-cache kernel shares of one eager cfg2 training step, from CUPTI (torch.profiler): unlike the ncu launch
list (cache flushed before every kernel) these are the durations the kernels have inside the running step.
usage: step_profile.py [B] [out.md]"""
import collections, os, re, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        k = re.sub(r"\(.*", "", ev.name); k = k.replace("void ", "").replace("a3t::", "")
        k = re.sub(r"at::native::.*", "at::native", k)
        agg[k[:90]][0] += 1; agg[k[:90]][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
lines = [f"one eager step: {sum(v[0] for v in agg.values())} device activities, {tot/1e3:.3f} ms of kernel time (warm, CUPTI)", "",
         "| kernel | launches | ms | share | us/launch |", "|---|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {k} | {c} | {t/1e3:.3f} | {100*t/tot:.1f}% | {t/c:.1f} |")
text = "\n".join(lines)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
