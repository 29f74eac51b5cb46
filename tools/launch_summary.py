#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: find one training step (between two
adam_kernel launches) and print per-kernel launches / time / share.  usage: launch_summary.py file.csv [out.md]"""
import collections, csv, io, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(io.StringIO(''.join(lines))); hdr = next(r)
ik, iv, ig = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
rows = [(x[ik], float(x[iv].replace(',', '')), x[ig]) for x in r]
adam = [i for i, x in enumerate(rows) if 'adam_kernel' in x[0]]
assert len(adam) >= 2, "need two optimizer steps in the capture"
step = rows[adam[0] + 1: adam[1] + 1]
tot = sum(t for _, t, _ in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t, g in step:
    k = re.sub(r'\(.*', '', n); k = re.sub(r'<.*', '', k).replace('void ', '')
    agg[k][0] += 1; agg[k][1] += t
out = [f"one step = {len(step)} launches, {tot/1e6:.3f} ms serialised (cold-cache, under ncu)", "",
       "| kernel | launches | ms | share |", "|---|---|---|---|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| {k} | {c} | {t/1e6:.3f} | {100*t/tot:.1f}% |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "a").write(text + "\n")
