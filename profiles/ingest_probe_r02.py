#!/usr/bin/env python
"""File-to-result probe (run on the GPU box): bench.py's `ingest` leg alone -- config 5's chunk 0 (10 M records, 1 M-gene
header) as a level-1 BGZF BAM on tmpfs through `msamtools filter ... --besthit | msamtools profile ...` -- at the given
host thread counts, with both processes' phase timers.  Usage: ingest_probe_r02.py [records] [threads,threads,...] [ref]"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
thr = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (16,)
cfg = bench.CONFIGS[5]
plan = bench.chunk_plan(cfg, n, n, 1)
raw, off = bench.gen_chunk(cfg, 0, 0, plan)
tlen = bench.target_lengths(cfg)
out = bench.ingest_entry(cfg, plan, tlen, raw[:int(off[-1])], off, thread_counts=thr, with_reference=len(sys.argv) > 3)
print(json.dumps(out))
