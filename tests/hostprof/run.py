#!/usr/bin/env python
"""Host-side profile of the drop-in CLI against a NULL DEVICE (tests/hostprof/nulldev.c): what bench.py's `ingest`
leg measures (level-1 BGZF BAM of config 5's chunk 0 -> `filter -b -u ... --besthit | profile --multi=proportional`),
with the GPU work replaced by nothing, so that reader / index / record output / header / table costs can be read on a
machine without a GPU.  The outputs are meaningless (see nulldev.c).

    python tests/hostprof/run.py [--records N] [--threads T] [--keep-file] [--perf]
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build():
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libmsamtools_b200.so")
    subprocess.check_call(["gcc", "-O2", "-g", "-std=gnu99", "-fPIC", "-shared", "-o", lib, os.path.join(HERE, "nulldev.c")])
    cs = os.path.join(ROOT, "msamtools_b200", "csrc")
    cli = os.path.join(BUILD, "msamtools")
    src = [os.path.join(cs, "cli", "msamtools_main.c")] + [os.path.join(cs, "host", f) for f in ("bamio.c", "finflate.c", "crc32x.c", "gzpar.c", "margs.c", "keyorder.c", "recwalk.c")]
    subprocess.check_call(["gcc", "-O2", "-g", "-std=gnu99", "-Wall", "-Wextra", "-o", cli] + src +
                          ["-L" + BUILD, "-lmsamtools_b200", "-Wl,-rpath," + BUILD, "-lz", "-lm", "-lpthread"])
    return cli


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=4_000_000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--file", default="/dev/shm/msb200_hostprof.bam")
    ap.add_argument("--only", choices=["filter", "profile", "pipe"], default="pipe")
    ap.add_argument("--perf", action="store_true")
    a = ap.parse_args()
    import bench
    cli = build()
    cfg = bench.CONFIGS[5]
    if not os.path.exists(a.file) or os.environ.get("REGEN"):
        plan = bench.chunk_plan(cfg, a.records, a.records, 1)
        tlen = bench.target_lengths(cfg)
        raw, off = bench.gen_chunk(cfg, 0, 0, plan, n_records=a.records)
        hdr = bench.bam_header_blob([f"g{i:07d}" for i in range(len(tlen))], tlen)
        t0 = time.perf_counter()
        n = bench.write_bgzf_threads(a.file, [hdr, raw[:int(off[-1])]], level=1, threads=a.threads)
        print(f"# wrote {a.file}: {len(off) - 1} records, payload {n / 1e9:.2f} GB, file {os.path.getsize(a.file) / 1e9:.2f} GB in {time.perf_counter() - t0:.1f} s")
    env = dict(os.environ, MSAMTOOLS_TIMING="1", MSAMTOOLS_THREADS=str(a.threads), LD_LIBRARY_PATH=BUILD)
    f_args = [cli] + cfg["ref_filter"] + [a.file]
    p_args = [cli] + cfg["ref_second"] + ["-o", "/dev/shm/msb200_hostprof.out.gz", "-"]
    for rep in range(2):
        t0 = time.perf_counter()
        if a.only == "filter":
            with open("/dev/null", "wb") as dn:
                p1 = subprocess.run(f_args, stdout=dn, stderr=subprocess.PIPE, env=env)
            err = p1.stderr.decode()
        elif a.only == "profile":
            # profile straight on the file (no pipe)
            p1 = subprocess.run(p_args[:-1] + [a.file], stderr=subprocess.PIPE, env=env)
            err = p1.stderr.decode()
        else:
            p1 = subprocess.Popen(f_args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
            p2 = subprocess.Popen(p_args, stdin=p1.stdout, stderr=subprocess.PIPE, env=env)
            p1.stdout.close()
            err = p1.stderr.read().decode() + p2.stderr.read().decode()
            p2.wait(); p1.wait()
        dt = time.perf_counter() - t0
        print(f"# run {rep}: {a.only} wall {dt:.2f} s")
        print("\n".join("    " + l for l in err.splitlines() if l.startswith("# timing") or l.startswith("# phase")))


if __name__ == "__main__":
    main()
