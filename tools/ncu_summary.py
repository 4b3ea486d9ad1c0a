#!/usr/bin/env python3
"""Print the roofline-relevant counters of one kernel from `ncu -i X.ncu-rep --page raw --csv`.
usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py [extra-substring ...]"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "dram__throughput.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct",
        "sm__throughput.avg.pct", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "sm__pipe_fma_cycles_active", "sm__pipe_alu_cycles_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_fmalite",
        "sm__pipe_fmaheavy_cycles_active", "sm__pipe_fmalite_cycles_active", "smsp__cycles_active.avg", "sm__sass_thread_inst_executed_op_integer",
        "smsp__average_warp", "smsp__warps_issue_stalled", "sm__sass_inst_executed_op_local", "local_load", "local_store", "derived__smsp__sass_thread_inst_executed_op"]
WANT += sys.argv[1:]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")][:90], "grid", vals[hdr.index("Grid Size")], "block", vals[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, vals):
        if any(w in h for w in WANT):
            print(f"  {h:95s} {v:>18s} {u}")
