#!/usr/bin/env python
"""Host-ingest probe (run on the GPU box): BGZF-compressed BAM -> drop-in CLI `profile`, with the
CLI's MSAMTOOLS_TIMING line (host read+inflate GB/s vs push/GPU time), for 1 and N inflate threads."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import samutil
from msamtools_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
p = synth.make_params("community", n_records=n, seed=13579)
raw, off, _ = synth.generate(p); tlen = synth.target_lengths(p)
names = [f"g{i:03d}" for i in range(len(tlen))]
path = "/dev/shm/ingest_probe.bam"
t = time.time(); samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, raw, level=1)
print(f"wrote {os.path.getsize(path)/1e6:.0f} MB BGZF ({len(raw)/1e6:.0f} MB payload, {len(off)-1} records) in {time.time()-t:.1f}s", flush=True)
cli = os.path.join(ROOT, "msamtools_b200", "bin", "msamtools")
for thr in (1, 4, 16):
    for cmd in (["profile", "--label", "x", "--multi=prop", "-o", "/dev/shm/ip.gz", path],
                ["filter", "-b", "-u", "-l", "80", "-p", "95", "-z", "80", "--besthit", path]):
        env = dict(os.environ, MSAMTOOLS_TIMING="1", MSAMTOOLS_THREADS=str(thr))
        t = time.time()
        r = subprocess.run([cli] + cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        wall = time.time() - t
        line = [l for l in r.stderr.splitlines() if l.startswith("# timing")]
        print(f"threads={thr:2d} {cmd[0]:8s} wall {wall:.2f}s rc={r.returncode} {line[0] if line else r.stderr[-200:]}", flush=True)
os.remove(path)
