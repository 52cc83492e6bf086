#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel launch count and total/mean duration, and the share of the
babe_b200 kernels (names starting with k_) in the whole run.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    short = name.split("(")[0].replace("void ", "").replace("babe::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
    tot[short][0] += 1
    tot[short][1] += v_us
total = sum(v[1] for v in tot.values())
ours = {k: v for k, v in tot.items() if k.startswith(("k_", "pfa::k_"))}
ours_t = sum(v[1] for v in ours.values())
print(f"# launch list summary: {path}\n")
print(f"total kernel time {total/1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches; "
      f"babe_b200 kernels {ours_t/1e3:.3f} ms ({100*ours_t/total:.2f} %) over {sum(v[0] for v in ours.values())} launches\n")
print("## babe_b200 kernels\n\n| kernel | launches | total us | mean us | share of all kernel time |\n|---|---:|---:|---:|---:|")
for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1]/v[0]:.2f} | {100*v[1]/total:.3f} % |")
print("\n## top 15 other kernels (PyTorch / cuDNN: the out-of-scope denoiser body)\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
other = sorted(((k, v) for k, v in tot.items() if not k.startswith(("k_", "pfa::k_"))), key=lambda kv: -kv[1][1])[:15]
for k, v in other:
    print(f"| `{k[:90]}` | {v[0]} | {v[1]/1e3:.2f} | {100*v[1]/total:.2f} % |")
