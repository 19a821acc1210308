#!/usr/bin/env python
"""Turn one `ncu --set full` capture into the files kept under profiles/:
   <prefix>_full_raw.csv (--page raw), <prefix>_summary.txt (selected metrics + stall shares), <prefix>_hot_sass.txt
   (--page source through tools/ncu_hot.py).   usage: ncu_summary.py capture.ncu-rep profiles/r1_score_kernel"""
import csv
import os
import subprocess
import sys

rep, prefix = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
open(prefix + "_full_raw.csv", "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
WANT = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
with open(prefix + "_summary.txt", "w") as fh:
    for k in WANT:
        if k in m:
            fh.write(f"{k:75s} {m[k][0]} {m[k][1]}\n")
    stalls = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: float(v) for h, (v, _) in m.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(stalls.values()) or 1.0
    fh.write("warp-state samples (share of all samples):\n")
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]):
        if v / tot >= 0.005:
            fh.write(f"    {k:28s} {100 * v / tot:5.1f} %\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
tmp = prefix + "_source.tmp.csv"
open(tmp, "w").write(src)
hot = subprocess.run([sys.executable, os.path.join(here, "ncu_hot.py"), tmp, "0.012"], stdout=subprocess.PIPE, text=True, check=True).stdout
open(prefix + "_hot_sass.txt", "w").write(hot)
os.remove(tmp)
print(open(prefix + "_summary.txt").read())
