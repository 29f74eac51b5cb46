import collections, re, sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from a3t_b200.vocoder import ParallelWaveGANGenerator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
gen = ParallelWaveGANGenerator(upsample_params={"upsample_scales": [4, 5, 3, 5]}).cuda().eval()
c = torch.randn(B, 80, 1024, device="cuda"); z = torch.randn(B, 1, 1024 * 300, device="cuda")
for _ in range(2): gen.generate(c, z)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gen.generate(c, z); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        k = re.sub(r"\(.*", "", ev.name).replace("void ", "").replace("a3t::", "")
        agg[k[:70]][0] += 1; agg[k[:70]][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"B={B}: {tot/1e3:.2f} ms")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{k:70s} {n:4d} {t/1e3:9.3f} ms {t/n:9.1f} us")
