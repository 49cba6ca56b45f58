"""Instruction / stall-sample share per phase of k_vignette_fused (phases are found from the '// ---- N.' markers
and helper function definitions in maze_fused.cu).  Usage: python tools/ncu_phases.py report.ncu-rep [filter]"""
import csv, subprocess, io, re, sys, os
rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1][:44]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        try: s = int(r[hdr.index("# Samples")]); i = int(r[hdr.index("Instructions Executed")])
        except ValueError: continue
        a = agg.setdefault((cur_fn, cur_file, int(r[0]), r[1].strip()[:90]), [0, 0]); a[0] += s; a[1] += i
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "maze_image_processing_pipeline_b200/csrc/maze_fused.cu")).read().split("\n")
pat = re.compile(r"\s*// ---- [0-9]|.*if \(prm.high_order\) \{|.*// shared rows -> staging|^__device__|^template <int R, int T>")
marks = [(i + 1, l.strip()[:50]) for i, l in enumerate(src) if pat.match(l)]
def phase(line, file):
    if file != "maze_fused.cu": return "helpers(" + file + ")"
    name = "top"
    for ln, nm in marks:
        if line >= ln: name = nm
    return name
for fn in sorted({k[0] for k in agg}):
    if flt and flt not in fn: continue
    tot_s = sum(v[0] for k, v in agg.items() if k[0] == fn) or 1
    tot_i = sum(v[1] for k, v in agg.items() if k[0] == fn) or 1
    ph = {}
    for k, v in agg.items():
        if k[0] != fn: continue
        a = ph.setdefault(phase(k[2], k[1]), [0, 0]); a[0] += v[0]; a[1] += v[1]
    print("=====", fn, "warp-instructions", tot_i, "samples", tot_s)
    for p, v in sorted(ph.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {v[1] / tot_i * 100:5.1f}%inst {v[0] / tot_s * 100:5.1f}%samples  {p}")
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:14]:
        print(f"      {v[0] / tot_s * 100:5.1f}%s {v[1] / tot_i * 100:5.1f}%i {k[1][:14]}:{k[2]:4d} {k[3]}")
