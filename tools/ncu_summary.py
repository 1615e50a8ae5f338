#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline counters, stall reasons and a per-phase split of the SASS
profile at the barriers.  usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp32.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__thread_inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
for i, n in enumerate(h):
    if n in want:
        print("%-90s %-16s %s" % (n, u[i], v[i]))
print("--- stall reasons (warps stalled per issue) ---")
for i, n in enumerate(h):
    if "smsp__average_warps_issue_stalled" in n and n.endswith("per_issue_active.ratio") and float(v[i]) > 0.05:
        print("  %-40s %s" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v[i]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
si, ii, sm, ti, we = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed"), h.index("L1 Wavefronts Shared Excessive")
data = [(r[si].strip(), int(r[ii]), int(r[sm]), int(r[ti]), int(r[we])) for r in rows[2:] if len(r) > we and r[ii].isdigit()]
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
print("--- SASS profile split at barriers / branches: %d warp-instructions, %d samples, %d SASS lines ---" % (tot, tots, len(data)))
seg, acc, first = 0, [0, 0, 0, 0], 0
for n, d in enumerate(data):
    acc[0] += d[1]; acc[1] += d[2]; acc[2] += d[3]; acc[3] += d[4]
    if "BAR." in d[0] or "WARPSYNC" in d[0] or n == len(data) - 1:
        if acc[0] * 200 > tot or acc[1] * 200 > tots:
            print("seg %2d sass[%4d..%4d] inst %5.1f%% samples %5.1f%% excess_smem_wavefronts %10d  ends with %s" % (seg, first, n, 100 * acc[0] / tot, 100 * acc[1] / tots, acc[3], d[0][:40]))
        seg += 1; acc = [0, 0, 0, 0]; first = n + 1
top = sorted(data, key=lambda d: -d[2])[:12]
print("--- top stall-sample instructions ---")
for d in top:
    print("  %5.2f%%  %s" % (100 * d[2] / tots, d[0][:90]))
