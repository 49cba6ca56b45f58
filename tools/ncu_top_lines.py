"""Top source lines (by stall samples) of every kernel in an .ncu-rep captured with --import-source on.
Usage: python tools/ncu_top_lines.py report.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = hdr = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1][:44]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and r[0].isdigit():
        try:
            s = int(r[hdr.index("# Samples")]); i = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        a = agg.setdefault((cur_fn, cur_file, int(r[0]), r[1].strip()[:96]), [0, 0])
        a[0] += s; a[1] += i
for fn in sorted({k[0] for k in agg}):
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    ts = sum(v[0] for _, v in items) or 1; ti = sum(v[1] for _, v in items) or 1
    print("=====", fn, "samples", ts, "warp-instructions", ti)
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0] / ts * 100:5.1f}%s {v[1] / ti * 100:5.1f}%i {k[1][:16]:16s}:{k[2]:4d} {k[3]}")
