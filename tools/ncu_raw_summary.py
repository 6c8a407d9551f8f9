#!/usr/bin/env python3
"""Per-kernel digest of an `ncu --page raw --csv` export (one row per profiled launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("==", r[idx["Kernel Name"]][:70], "grid", r[idx.get("Grid Size", 0)] if "Grid Size" in idx else "")
    for k in keys:
        if k in idx:
            print(f"  {k:84s} {r[idx[k]]:>18s} {units[idx[k]]}")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                v = float(r[idx[h]].replace(",", ""))
            except ValueError:
                continue
            if v > 0.15:
                print(f"    stall {h.split('stalled_')[1].split('_per_issue')[0]:26s} {v:.3f}")
