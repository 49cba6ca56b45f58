"""profiles/ncu_traffic.json from an ncu launch list of `bench.py --steps 2 --warmup 3` (metrics gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum): DRAM bytes per pixel of the first k_band_front launch (= batch 0).
Usage: python tools/ncu_traffic.py profiles/launches_r2_bands.csv"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
first = None
vals = {}
for r in rows:
    if r[4].startswith("void k_band_front") or r[4].startswith("k_band_front"):
        if first is None:
            first = r[0]
        if r[0] == first:
            vals[r[12]] = float(r[14].replace(",", ""))
hs, ws = bench.job_sizes()
px = int((hs[:4096].astype(np.int64) * ws[:4096]).sum())
rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
rec = {"k_band_front": {"dram_bytes_per_px": (rd + wr) / px, "dram_bytes_read": rd, "dram_bytes_write": wr, "pixels": px,
                        "source": os.path.relpath(path, ROOT) + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                                  "dram__bytes_write.sum,... --clock-control none; first k_band_front launch = batch 0 of "
                                  "`bench.py --steps 2 --warmup 3`)"}}
json.dump(rec, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(rec))
