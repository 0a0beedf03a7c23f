#!/usr/bin/env python
"""One line per profiled launch from an .ncu-rep (`ncu --set full`): duration, DRAM bytes, instruction / issue / pipe
counters.  usage: python tools/ncu_summary.py report.ncu-rep > profiles/summary.txt"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
hdr, units = rows[0], rows[1]
iK = hdr.index("Kernel Name")
cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
print("# " + " ".join(sys.argv[1:]))
print("kernel | " + " | ".join("%s [%s]" % (m, units[i]) for m, i in cols))
for r in rows[2:]:
    print(r[iK][:44] + " | " + " | ".join(r[i] for _, i in cols))
