#!/usr/bin/env python3
"""Dynamic opcode mix / hot regions of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`.
usage: ncu_opmix.py src.csv [frames]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 476160.0
hdr = rows[1]
ia, isrc, iex, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iwf = hdr.index("L1 Wavefronts Shared")
ops = collections.Counter(); samples = collections.Counter(); wf = collections.Counter()
tot = 0; tots = 0
lines = []
for r in rows[2:]:
    if len(r) <= iwf: continue
    s = r[isrc].strip()
    parts = s.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
    base = op.split('.')[0]
    n = int(r[iex] or 0); sm = int(r[ism] or 0); w = int(r[iwf] or 0)
    ops[base] += n; samples[base] += sm; wf[base] += w; tot += n; tots += sm
    lines.append((n, sm, w, s))
print(f"total warp-inst {tot} = {tot/frames:.0f}/frame; samples {tots}")
for k, v in ops.most_common(32):
    print(f"  {k:10s} {v/frames:8.1f}/frame  {100*v/tot:5.1f}%   samples {100*samples[k]/max(tots,1):5.1f}%  smem wavefronts {wf[k]/frames:7.1f}/frame")
