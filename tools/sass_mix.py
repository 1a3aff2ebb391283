#!/usr/bin/env python3
"""Dynamic SASS opcode mix of the first kernel in an ncu source-page CSV: python tools/sass_mix.py <src.csv> [frames_x_threads]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
isrc, iex, ismp = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
mix = collections.Counter(); smp = collections.Counter(); tot = 0; stot = 0
for r in rows[2:]:
    if len(r) <= iex or r[0] == "Kernel Name" or r[0] == "Address":
        if r and r[0] == "Kernel Name" and tot: break
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
    if not m: continue
    op = m.group(2)
    n = int(r[iex]); mix[op] += n; tot += n
    s = int(r[ismp] or 0); smp[op] += s; stot += s
norm = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"total warp inst {tot}")
for op, n in mix.most_common(40):
    extra = f"  per_thread_frame={n/norm:7.1f}" if norm else ""
    print(f"{op:10s} {n:12d} {n/tot*100:6.2f}%  samples {smp[op]/max(stot,1)*100:6.2f}%{extra}")
