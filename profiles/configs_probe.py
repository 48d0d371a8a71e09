#!/usr/bin/env python
"""Resident-throughput probe for the BASELINE.json configs other than the bench default (one B200):
   3: N name-sorted alignments, 10k references, 30 % multi-mappers: profile --multi=proportional (no filter stage)
   4: N alignments, 100 genomes: filter -l 80 -p 95 -z 80 fused with coverage --summary
   5: N alignments to a 1M-gene catalogue: filter --besthit | profile --multi=proportional
   1: N alignments, 100 genomes: filter -l 80 -p 95 -z 80 with record output (literal configs[1])
usage: configs_probe.py [records] [configs, e.g. 3,4,5] [warmup] [steps]   (defaults 20 M, 3,4,5, 30, 40)"""
import json, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import msamtools_b200 as m
from msamtools_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["3", "4", "5"]
W = int(sys.argv[3]) if len(sys.argv) > 3 else 30
K = int(sys.argv[4]) if len(sys.argv) > 4 else 40
CASES = [
    ("1_filter_records_100_genomes", "community", dict(l=80, p=95, z=80, records=True, kept=False)),
    ("3_profile_10k_refs", "catalog10k", dict(profile=True, multi="proportional", do_filter=False)),
    ("4_filter+coverage_100_genomes", "community", dict(l=80, p=95, z=80, coverage=True, coverage_summary=True, kept=False)),
    ("5_besthit+profile_1M_genes", "genes1m", dict(l=80, p=95, z=80, besthit=True, profile=True, multi="proportional", kept=False)),
]
for name, preset, opts in CASES:
    if name[0] not in which:
        continue
    p = synth.make_params(preset, n_records=n, seed=13579)
    t = time.time(); raw, off, st = synth.generate(p); tlen = synth.target_lengths(p)
    nrec = len(off) - 1
    kw = dict(n_targets=len(tlen))
    if opts.get("coverage"):
        kw["target_len"] = tlen
    with m.Context(**kw, **opts) as ctx:
        d_raw = ctx.device_alloc(raw.nbytes); d_off = ctx.device_alloc(off.nbytes)
        ctx.device_upload(d_raw, raw); ctx.device_upload(d_off, off)
        def step():
            ctx.reset()
            ctx.push_device(d_raw, raw.nbytes, d_off, nrec)
            if opts.get("profile"):
                return ctx.finish_profile()[1]
            if opts.get("coverage"):
                return ctx.finish_coverage()[0].sum()
            ctx.sync()
            return ctx.kept_count()
        for _ in range(W):           # also brings the SM clocks up: short loops right after the (CPU-side) data generation run at idle clocks
            res = step()
        ctx.timing(reset=True)
        ctx.sync(); ctx.mark(0)
        for _ in range(K):
            res = step()
        ctx.mark(1); ms = ctx.elapsed_ms(0, 1) / K
        tim = ctx.timing()
        out = dict(config=name, records=nrec, n_refs=len(tlen), ms_per_step=ms, M_aln_per_s=nrec / ms / 1e3,
                   decode_ms=tim["decode_ms"] / max(tim["decode_launches"], 1), launches_per_step=tim["kernel_launches"] / K,
                   alg_GBps=tim["alg_bytes"] / max(tim["decode_launches"], 1) / (tim["decode_ms"] / max(tim["decode_launches"], 1)) / 1e6,
                   raw_GB=raw.nbytes / 1e9, slow_records=tim["slow_records"] // K, gen_s=round(time.time() - t, 1),
                   result=(dict(inserts=res["mapped_inserts"], multi=res["multi"], iters=res["iterations"], lists=res["n_lists"]) if isinstance(res, dict) else int(res)))
        print(json.dumps(out), flush=True)
        ctx.device_free(d_raw); ctx.device_free(d_off)
    del raw, off
