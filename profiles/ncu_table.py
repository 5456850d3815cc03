#!/usr/bin/env python
"""Compact per-launch table from an .ncu-rep (read here with `ncu -i`):  ncu_table.py REPORT [metric ...]"""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
           "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
rep, metrics = sys.argv[1], sys.argv[2:] or DEFAULT
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
head, units, body = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(head)}
if metrics == ["list"]:
    for n in head:
        if "tensor" in n or "pipe" in n:
            print(n)
    sys.exit(0)
have = [m for m in metrics if m in col]
print("# " + rep)
for r in body:
    print(r[col["Kernel Name"]][:110])
    for m in have:
        print("    %-80s %16s %s" % (m, r[col[m]], units[col[m]]))
