#!/usr/bin/env python3
"""Top SASS instructions by warp-stall samples from an .ncu-rep captured with --import-source on.
usage: ncu_hotspots.py <rep> [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {k: i for i, k in enumerate(hdr)}
key = "Warp Stall Sampling (All Samples)"
data = []
for idx, r in enumerate(rows[h + 1:]):
    try:
        data.append((float(r[ci[key]]), idx, r[ci["Source"]], float(r[ci["Instructions Executed"]] or 0)))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print(f"total samples {tot:.0f}, instructions {len(data)}")
for s, idx, src, n in sorted(data, reverse=True)[:top]:
    print(f"{100 * s / tot:5.1f}%  #{idx:5d}  exec {n:12.0f}  {src[:100]}")
