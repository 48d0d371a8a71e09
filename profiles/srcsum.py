#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` by CUDA source line.
usage: srcsum.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""; hdr = None; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = {k: j for j, k in enumerate(r)}; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "":      # a CUDA line with aggregated metrics
        try:
            n = int(r[hdr["Instructions Executed"]]); s = int(r[hdr["# Samples"]] or 0)
            ti = int(r[hdr["Thread Instructions Executed"]])
        except ValueError:
            continue
        data.append((n, s, ti, cur_file, r[0], r[1].strip()[:100]))
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f"total warp-inst {tot}  samples {ts}")
for n, s, ti, f, ln, src in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/max(ts,1):5.1f}% smp  thr/inst {ti/max(n,1):4.1f}  {f}:{ln:>4s} {src}")
