"""Counts of the SASS mnemonics that matter (TMA / bulk copies, cluster barriers, DSMEM mapping, atomics, barriers,
vector stores ...) per kernel of libmaze_b200.so.  Usage: python tools/sass_summary.py > profiles/sass_r2.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "maze_image_processing_pipeline_b200", "csrc", "libmaze_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = ("UBLKCP", "UTMACMDFLUSH", "FENCE", "DEPBAR", "UCGABAR", "MAPA", "ATOM", "MEMBAR", "CCTL", "ATOMG", "ATOMS", "RED", "BAR", "REDUX",
        "CREDUX", "SHFL", "MATCH", "VOTE", "POPC", "UPOPC", "PRMT", "UPRMT", "SHF", "USHF", "VIMNMX", "LDG", "STG.E.128", "DADD", "DMUL",
        "DFMA", "MUFU")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for w in WANT:
        if op == w or op.startswith(w + ".") or op.startswith(w + "_"):
            counts[cur][w] += 1
            break
print("SASS mnemonics per kernel of libmaze_b200.so (cuobjdump -sass, sm_100a), round 2, tools/sass_summary.py.  UBLKCP = cp.async.bulk (TMA")
print("engine, bulk shared->global copies of the zero fill); FENCE = fence.proxy.async; DEPBAR on the bulk-group scoreboard =")
print("cp.async.bulk.wait_group; UCGABAR = barrier.cluster arrive / wait (thread-block cluster; its distributed shared memory is reached with generic LD / ST / ATOM).")
for fn in sorted(counts):
    c = counts[fn]
    if not c:
        continue
    print(f"{fn[:70]:70s} " + " ".join(f"{k}={c[k]}" for k in sorted(c)))
