#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full --import-source on) into the text summary kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"== kernel: {name}")
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            print(f"  {h:78s} {v} {u}")
    print("  -- warp stall reasons (warps per issue-active cycle)")
    st = [(h, float(v)) for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
          and "not_issued" not in h]
    for h, v in sorted(st, key=lambda kv: -kv[1])[:8]:
        print(f"     {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    h2 = srows[1]
    ix = {h: i for i, h in enumerate(h2)}
    c, cs = Counter(), Counter()
    for r in srows[2:]:
        try:
            parts = r[ix["Source"]].split()
            op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
            c[op] += float(r[ix["Instructions Executed"]] or 0)
            cs[op] += float(r[ix["# Samples"]] or 0)
        except (IndexError, ValueError, KeyError):
            continue
    tot, tots = sum(c.values()) or 1, sum(cs.values()) or 1
    print("  -- SASS opcode mix: share of executed warp instructions / share of stall samples")
    for op, v in c.most_common(14):
        print(f"     {op:10s} {100 * v / tot:5.1f}%  {100 * cs[op] / tots:5.1f}%")
