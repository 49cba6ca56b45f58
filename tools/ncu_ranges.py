"""Instruction / stall-sample share per source-line range of one kernel of an .ncu-rep (--import-source on).
Usage: python tools/ncu_ranges.py report.ncu-rep kernel_substring file:lo-hi[:name] ..."""
import csv, io, subprocess, sys
rep, flt, specs = sys.argv[1], sys.argv[2], sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and flt in (cur_fn or ""):
        try: s = int(r[hdr.index("# Samples")]); i = int(r[hdr.index("Instructions Executed")])
        except ValueError: continue
        a = agg.setdefault((cur_file, int(r[0])), [0, 0]); a[0] += s; a[1] += i
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
print(flt, "samples", ts, "warp-instructions", ti)
used = set()
for sp in specs:
    parts = sp.split(":")
    f, rng = parts[0], parts[1]
    name = parts[2] if len(parts) > 2 else sp
    lo, hi = (int(x) for x in rng.split("-"))
    s = i = 0
    for (cf, ln), v in agg.items():
        if cf == f and lo <= ln <= hi:
            s += v[0]; i += v[1]; used.add((cf, ln))
    print(f"  {i / ti * 100:5.1f}%inst {s / ts * 100:5.1f}%samples  {name}")
s = sum(v[0] for k, v in agg.items() if k not in used); i = sum(v[1] for k, v in agg.items() if k not in used)
print(f"  {i / ti * 100:5.1f}%inst {s / ts * 100:5.1f}%samples  (other)")
