#!/usr/bin/env python
"""Profile-only steps (BASELINE configs[2] shape): decode launch time with and without the finish (EM) kernels in the
loop, plus SM clocks sampled during the loop.  Usage: python profiles/profile_only_probe.py [records]"""
import os, sys
sys.path.insert(0, ".")
import msamtools_b200 as m
from msamtools_b200 import synth
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
p = synth.make_params("catalog10k", n_records=n, seed=13579)
raw, off, st = synth.generate(p); tlen = synth.target_lengths(p)
with m.Context(n_targets=len(tlen), profile=True, multi="proportional", do_filter=False) as ctx:
    d_raw = ctx.device_alloc(raw.nbytes); d_off = ctx.device_alloc(off.nbytes)
    ctx.device_upload(d_raw, raw); ctx.device_upload(d_off, off)
    for with_finish in (True, False, True):
        for _ in range(2):
            ctx.reset(); ctx.push_device(d_raw, raw.nbytes, d_off, len(off) - 1)
            if with_finish: ctx.finish_profile()
        ctx.timing(reset=True)
        s = bench.ClockSampler(0)
        if not os.environ.get("NOSAMPLER"): s.start()
        for _ in range(40):
            ctx.reset(); ctx.push_device(d_raw, raw.nbytes, d_off, len(off) - 1)
            if with_finish: ctx.finish_profile()
        ctx.sync()
        clocks = s.stop() if not os.environ.get("NOSAMPLER") else None
        t = ctx.timing()
        print("finish" if with_finish else "push only", "decode ms/launch %.4f" % (t["decode_ms"] / t["decode_launches"]),
              "total ms/chunk %.3f" % (t["total_ms"] / t["decode_launches"]), clocks, flush=True)
