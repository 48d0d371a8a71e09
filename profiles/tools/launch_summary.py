#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and share per kernel name.
usage: launch_summary.py launches.csv [skip-first-N-launches]"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = collections.Counter(); cnt = collections.Counter()
k = 0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k += 1
    if k <= skip:
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name] += us; cnt[name] += 1
T = sum(tot.values())
print(f"# {sum(cnt.values())} launches, {T/1e3:.3f} ms total (serialised, cold cache: compare SHARES)")
for name, us in tot.most_common():
    print(f"{name:60s} n={cnt[name]:4d}  {us:10.1f} us  {100*us/T:5.1f}%  avg {us/cnt[name]:9.1f} us")
