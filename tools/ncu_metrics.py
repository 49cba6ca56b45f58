"""Key raw metrics of every kernel in an .ncu-rep.  Usage: python tools/ncu_metrics.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("---")
    for w in want:
        if w in idx:
            val = r[idx[w]]
            print(f"{w:62s} {val[:70]} {units[idx[w]]}")
