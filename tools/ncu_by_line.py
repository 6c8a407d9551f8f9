#!/usr/bin/env python3
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass) with nvdisasm -gi -c line
info of the same cubin: warp-stall samples, executed instructions and shared-memory wavefronts per OUTERMOST source
line of the kernel's own .cu file.
usage: ncu_by_line.py src.csv lines.txt <function substring> <file.cu> [units] [source file] [deep]
(deep: attribute to the DEEPEST inlined frame that lies in <file.cu> instead of the outermost line)"""
import csv, re, sys, collections
src_csv, lines_txt, fn, cu = sys.argv[1:5]
units = float(sys.argv[5]) if len(sys.argv) > 5 else 476160.0
# ---- nvdisasm: per instruction offset the outermost line in `cu`
off2line = {}
deep = len(sys.argv) > 7 and sys.argv[7] == "deep"
cur = None; infn = False; last_cu = None; chain_first = None
for ln in open(lines_txt):
    if ln.startswith(".text."):
        infn = fn in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if m.group(1).endswith(cu) and "inlined at" not in ln: last_cu = int(m.group(2))
        elif m.group(1).endswith(cu): pass
        # outermost = the last "//## File" before the instruction without "inlined at"; track separately
        if deep:
            if chain_first is None and m.group(1).endswith(cu): chain_first = (m.group(1), int(m.group(2)))
            if "inlined at" not in ln:
                cur = chain_first if chain_first else (m.group(1), int(m.group(2)))
                chain_first = None
        elif "inlined at" not in ln: cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+\S", ln)
    if m and cur: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ci = {k: i for i, k in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[ci["Address"]], 16)
    if base is None: base = a
    key = off2line.get(a - base, ("?", 0))
    if not key[0].endswith(cu): key = (key[0].split("/")[-1], key[1])
    else: key = (cu, key[1])
    e = agg[key]
    e[0] += int(r[ci["# Samples"]] or 0); e[1] += int(r[ci["Instructions Executed"]] or 0); e[2] += int(r[ci["L1 Wavefronts Shared"]] or 0)
    for h in reasons:
        v = int(r[ci[h]] or 0)
        if v: e[3][h[6:]] += v
tot = sum(e[0] for e in agg.values())
print(f"samples {tot}; per-unit = per frame")
src = {}
try:
    for i, l in enumerate(open([p for p in [sys.argv[6]] if p][0]) if len(sys.argv) > 6 else [], 1): src[i] = l.rstrip()
except Exception: pass
for key in sorted(agg, key=lambda k: (k[0] != cu, k[1])):
    e = agg[key]
    if e[0] < tot * 0.002 and e[1] / units < 2: continue
    top = ", ".join(f"{k} {100*v/max(e[0],1):.0f}%" for k, v in e[3].most_common(3))
    print(f"{key[0]}:{key[1]:4d}  {100*e[0]/tot:5.1f}%  inst {e[1]/units:7.1f}  smem-wf {e[2]/units:6.1f}   [{top}]  {src.get(key[1], '')[:70].strip() if key[0]==cu else ''}")
