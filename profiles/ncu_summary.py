#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS stall samples) -- used to produce the profiles/*.txt files.
usage: python profiles/ncu_summary.py gpurun_out/foo.ncu-rep [n_top_sass]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "local"]
for h, u, v in zip(hdr, units, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.2) or "local_" in h and "sum" in h and float(v or 0) > 0:
        print(f"{h} [{u}] = {v}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h2 = src[1]; ci = {h: i for i, h in enumerate(h2)}
data = [r for r in src[2:] if len(r) > ci["# Samples"]]
tot = sum(float(r[ci["# Samples"]] or 0) for r in data) or 1
totex = sum(float(r[ci["Instructions Executed"]] or 0) for r in data) or 1
agg, aggex = collections.Counter(), collections.Counter()
for r in data:
    t = r[ci["Source"]].split()
    op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "?")).split(".")[0]
    agg[op] += float(r[ci["# Samples"]] or 0); aggex[op] += float(r[ci["Instructions Executed"]] or 0)
print(f"-- SASS: {len(data)} instructions, {tot:.0f} samples, {totex:.0f} warp-instructions executed")
for op, v in agg.most_common(ntop):
    print(f"   {op:10s} samples {v / tot * 100:5.1f}%  executed {aggex[op] / totex * 100:5.1f}%")
for name in ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_branch_resolving", "stall_selected", "stall_not_selected", "stall_dispatch", "stall_no_inst"]:
    if name in ci:
        print(f"   {name:24s} {sum(float(r[ci[name]] or 0) for r in data) / tot * 100:5.1f}%")
