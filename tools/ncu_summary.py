#!/usr/bin/env python3
"""Print the handful of ncu raw metrics this project tracks from an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
rep = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 476160.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in d:
        print(f"{k:86s} {d[k][0]:>16s} {d[k][1]}")
for k in sorted(d):
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        v = float(d[k][0].replace(",", ""))
        if v > 0.05:
            print(f"  stall {k.split('stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
def num(k):
    return float(d[k][0].replace(",", ""))
print(f"per frame: {num('smsp__inst_executed.sum')/frames:.0f} warp-inst, "
      f"{num('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')/frames:.0f} smem wavefronts, "
      f"{num('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')/frames:.0f} bank conflicts")
