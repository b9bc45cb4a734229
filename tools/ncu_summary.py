#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the few lines DESIGN.md / profiles/ quote.  Usage: ncu_summary.py rep [title]"""
import csv, subprocess, sys
rep = sys.argv[1]; title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__sass_thread_inst_executed_op_integer_pred_on.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'sm__cycles_elapsed.avg']
print("#", title)
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    for k in keys:
        if k in d:
            print("%-78s %s %s" % (k, d[k][0], d[k][1]))
    st = [(float(v[0]), h) for h, v in d.items() if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    print("warp-issue stall reasons (per issue):")
    for v, h in sorted(st, reverse=True)[:6]:
        print("  %-40s %.3f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    print()
