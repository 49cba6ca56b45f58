"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel (and grid size
for the fused kernel) launch count, total time and share.  Usage: python summarize_launches.py file.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    if "k_vignette_fused" in name:
        name = f"{name} block={r[7]}"
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for k, v in agg.items() if not k.startswith("k_synth"))
print(f"{'kernel':60s} {'n':>5s} {'total_us':>12s} {'avg_us':>10s} {'share(excl. k_synth)':>8s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {v[0]:5d} {v[1] / 1e3:12.1f} {v[1] / 1e3 / v[0]:10.1f} {v[1] / tot:8.3f}")
