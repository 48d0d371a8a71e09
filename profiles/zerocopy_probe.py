#!/usr/bin/env python
"""Probe: decode straight out of pinned HOST memory (zero-copy over PCIe: only the touched 64-byte granules
travel) versus the staged msg_push (H2D of the whole batch, then kernels).  Usage: python profiles/zerocopy_probe.py [records]"""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np
import bench
from msamtools_b200 import api

n_req = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
raw, off, tlen, _ = bench.make_batch(n_req, 0, pinned=True)
n = len(off) - 1
R = len(tlen)
ctx = api.Context(l=80, p=95, z=80, besthit=True, profile=True, multi="proportional", n_targets=R, n_features=R, target_len=tlen)
def staged():
    ctx.reset(); ctx.push(raw, off); return ctx.finish_profile()
d_off = ctx.device_alloc(off.nbytes); ctx.device_upload(d_off, off)
def zero_copy():
    ctx.reset(); ctx.push_device(C.c_void_p(raw.ctypes.data), raw.nbytes, d_off, n); return ctx.finish_profile()
for name, fn in (("staged msg_push", staged), ("zero-copy push_device(host pinned)", zero_copy), ("staged msg_push", staged)):
    r = fn(); fn()
    t0 = time.perf_counter()
    for _ in range(3): r2 = fn()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name:40s} {dt*1e3:8.2f} ms/step  {n/dt/1e6:8.1f} M aln/s  {raw.nbytes/dt/1e9:6.1f} GB/s of BAM", flush=True)
    assert np.allclose(r[0], r2[0])
