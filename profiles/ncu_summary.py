#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key counters, stall mix, top stalled SASS.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index] > profiles/xxx.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2 + idx]
col = {h: i for i, h in enumerate(hdr)}
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]
print(f"# {rep} launch {idx}")
for w in want:
    if w in col:
        print(f"{w:70s} {data[col[w]]} {units[col[w]]}")
stalls = [(float(data[i]), h) for h, i in col.items()
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
          and data[i]]
print("\n# warp stall reasons (average warps stalled per issue-active cycle)")
for v, h in sorted(stalls, reverse=True)[:9]:
    print(f"{v:8.3f}  {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if starts:
    hi = starts[min(idx, len(starts) - 1)]
    h = rows[hi]
    end = starts[starts.index(hi) + 1] - 1 if starts.index(hi) + 1 < len(starts) else len(rows)
    body = [r for r in rows[hi + 1:end] if len(r) == len(h)]
    si, ei, ci = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[si] or 0) for r in body)
    ex = sum(int(r[ei] or 0) for r in body)
    print(f"\n# source page: {len(body)} SASS instructions, {ex} warp-instructions executed, {tot} samples")
    print("# top stalled instructions: samples  executed  sass  {top stall reasons}")
    for r in sorted(body, key=lambda r: -int(r[si] or 0))[:14]:
        st = sorted(((int(r[i] or 0), h[i]) for i in stall_cols), reverse=True)[:2]
        print(f"{r[si]:>6} {r[ei]:>9}  {r[ci][:64]:64s} {[(n, v) for v, n in st if v]}")
