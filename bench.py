#!/usr/bin/env python
"""bench.py -- `M alignments/s filter+besthit+profile` on B200 (BASELINE.json metric).

A step = one pass of the hot path (decode + filter statistics, best-hit, proportional profile
incl. the EM loop and, at N>1, the NCCL allreduce) over one batch of synthetic name-sorted
PE150 alignments of BASELINE.json configs[1] shape (per-GPU batch = --records, default 10 M;
weak scaling: every rank owns its own QNAME-group shard).

  value      whole-job throughput, batch resident in HBM when the timed region starts
  e2e        same metric through msg_push with HOST (pinned) buffers: H2D of the batch and D2H of
             the abundance vector inside the timed region
  roofline   decode/filter kernel: algorithmic bytes per launch / CUDA-event launch time vs the
             measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (restated reference algorithm) on a bounded sample, host cores

`--impl reference` times the reference's CPU algorithm (oracle/_ref when built, else the oracle
port) on the box's host cores for the same config/metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FILTER_OPTS = dict(l=80, p=95, z=80, besthit=True)
MULTI = "proportional"
METRIC = "M alignments/s filter+besthit+profile"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def make_batch(records, rank, pinned=True):
    """configs[1]-shaped shard for this rank, generated straight into pinned host memory."""
    from msamtools_b200 import synth
    # disjoint insert numbers per rank; the stride keeps "sim%08llu" at 8 digits on every rank (up to 9 ranks at the default
    # size), so that per-GPU work really is the same -- a 1e9 stride gave ranks >= 1 two more QNAME bytes per record
    stride = 10_000_000 if records <= 20_000_000 else 1_000_000_000
    p = synth.make_params("community", n_records=records, seed=13579, qname_base=rank * stride)
    cap_b, cap_r = (records + 64) * 330, records + 66
    raw = off = None
    if pinned:
        try:
            import torch
            raw = torch.empty(cap_b, dtype=torch.uint8, pin_memory=True).numpy()
            off = torch.empty(cap_r, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        except Exception:
            raw = off = None
    if raw is None:
        raw, off = np.empty(cap_b, dtype=np.uint8), np.empty(cap_r, dtype=np.uint64)
    raw, off, st = synth.generate(p, raw, off)
    return raw, off, synth.target_lengths(p), st


def cpu_reference_run(raw, off, tlen, sample_records, threads):
    """The reference algorithm on host cores: `threads` independent QNAME-boundary shards of the first
    `sample_records` records, one orc_pipeline (filter -> besthit -> proportional profile) per shard.
    Returns (alignments processed, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    import msamtools_b200 as m
    from oracle import oracle as orc
    orc.load()
    n = min(sample_records, len(off) - 1)
    cuts = [0]
    for t in range(1, threads):
        k = m.split_point(raw, off, n * t // threads)
        cuts.append(max(k, cuts[-1]))
    cuts.append(m.split_point(raw, off, n) or n)
    cfg = orc.filter_cfg(**FILTER_OPTS)
    shards = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b > a:
            lo, hi = int(off[a]), int(off[b])
            shards.append((raw[lo:hi], (off[a:b + 1] - off[a]).copy()))

    def work(sh):
        return orc.pipeline(sh[0], sh[1], cfg, len(tlen), 3)[1]["n_kept"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:      # ctypes releases the GIL inside the C call
        list(ex.map(work, shards))
    dt = time.perf_counter() - t0
    return cuts[-1], dt, len(shards)


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "msamtools")


def have_ref_binary():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


class RefPipelines:
    """The reference's own object code (oracle/_ref/msamtools: msamtools v1.1.3 sources compiled against the
    I/O shim) run as its documented pipe `filter -b -u -l 80 -p 95 -z 80 --besthit in.bam | profile
    --multi=proportional -o out.gz -`, one pipe (two processes) per QNAME-boundary shard, shards on tmpfs as
    level-0 BGZF BAM so that inflate cost is negligible."""

    def __init__(self, raw, off, tlen, n_pipes, per_pipe):
        import tempfile
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import samutil
        import msamtools_b200 as m
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        self.dir = tempfile.mkdtemp(prefix="msb200_ref_", dir=base)
        names = [f"g{i:06d}" for i in range(len(tlen))]
        hdr = samutil.synth_header(names, tlen)
        n = len(off) - 1
        self.paths, self.n = [], 0
        a = 0
        for k in range(n_pipes):
            want = min(n, a + per_pipe)
            b = m.split_point(raw, off, want) if want < n else n
            if b <= a:
                break
            path = os.path.join(self.dir, f"shard{k}.bam")
            samutil.write_bam(path, hdr, names, tlen, raw[int(off[a]):int(off[b])], level=0)
            self.paths.append(path)
            self.n += b - a
            a = b

    def run(self):
        t0 = time.perf_counter()
        procs = []
        for i, path in enumerate(self.paths):
            f = subprocess.Popen([REF_BIN, "filter", "-b", "-u", "-l", "80", "-p", "95", "-z", "80", "--besthit", path],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            g = subprocess.Popen([REF_BIN, "profile", "--label", "S", "--multi=proportional", "-o", os.path.join(self.dir, f"out{i}.gz"), "-"],
                                 stdin=f.stdout, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            f.stdout.close()
            procs.append((f, g))
        for f, g in procs:
            if g.wait() != 0 or f.wait() != 0:
                raise RuntimeError("reference pipeline failed")
        return time.perf_counter() - t0

    def close(self):
        import shutil
        shutil.rmtree(self.dir, ignore_errors=True)


def cpu_baseline_entry(raw, off, tlen, args, n):
    """cpu_baseline object for the JSON line: the reference's own object code when oracle/_ref exists
    (kind "reference"), else the oracle port (kind "port")."""
    cores = os.cpu_count() or 1
    if have_ref_binary() and not args.cpu_port:
        pipes = args.cpu_threads // 2 if args.cpu_threads else max(1, min(cores // 2, 16))
        per = args.cpu_sample // pipes if args.cpu_sample else 500_000
        rp = RefPipelines(raw, off, tlen, pipes, min(per, n))
        try:
            rp.run()                                   # page-cache / exec warm-up
            dt = min(rp.run() for _ in range(2))
        finally:
            rp.close()
        return {"value": rp.n / dt / 1e6, "unit": "M alignments/s", "cores": 2 * len(rp.paths), "kind": "reference",
                "host_cores_available": cores,
                "sample": f"{rp.n} alignments of the same batch as {len(rp.paths)} QNAME-boundary shards (level-0 BGZF BAM on tmpfs), one "
                          f"`msamtools filter -b -u -l 80 -p 95 -z 80 --besthit | msamtools profile --multi=proportional` pipe per shard "
                          f"(2 processes each; the reference is single-threaded); reference arithmetic, shim I/O (not htslib 1.24); best of 2"}, rp.n / dt / 1e6
    threads = args.cpu_threads or min(cores, 32)
    sample = args.cpu_sample or 2_000_000 * threads
    nn, dt, nsh = cpu_reference_run(raw, off, tlen, min(sample, n), threads)
    return {"value": nn / dt / 1e6, "unit": "M alignments/s", "cores": threads, "kind": "port", "host_cores_available": cores,
            "sample": f"first {nn} alignments of the same batch in {nsh} QNAME-boundary shards, one in-memory oracle pipeline "
                      f"(filter+besthit+proportional profile, no file I/O) per thread; the reference is single-threaded"}, nn / dt / 1e6


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    use_ref = have_ref_binary() and not args.cpu_port
    if use_ref:
        pipes = args.cpu_threads // 2 if args.cpu_threads else max(1, min(cores // 2, 16))
        per = args.cpu_sample // pipes if args.cpu_sample else 500_000
        raw, off, tlen, _ = make_batch(min(args.records, pipes * per + 1000), 0, pinned=False)
        rp = RefPipelines(raw, off, tlen, pipes, per)
        try:
            for _ in range(max(1, min(args.warmup, 2))):
                rp.run()
            tot_t = sum(rp.run() for _ in range(args.steps))
        finally:
            rp.close()
        n, nsh, used = rp.n, len(rp.paths), 2 * len(rp.paths)
        tot_n = n * args.steps
        kind = "reference"
        sample = (f"{n} alignments per step as {nsh} QNAME-boundary shards (level-0 BGZF BAM on tmpfs), one `msamtools filter -b -u -l 80 -p 95 "
                  f"-z 80 --besthit | msamtools profile --multi=proportional` pipe per shard (2 processes each; the reference is "
                  f"single-threaded); reference arithmetic (oracle/_ref), shim I/O (not htslib 1.24)")
    else:
        threads = args.cpu_threads or min(cores, 32)
        sample_n = args.cpu_sample or 2_000_000 * threads
        raw, off, tlen, _ = make_batch(min(args.records, sample_n + 1000), 0, pinned=False)
        for _ in range(args.warmup):
            cpu_reference_run(raw, off, tlen, min(sample_n, 200_000), threads)
        tot_n, tot_t = 0, 0.0
        for _ in range(args.steps):
            n, dt, nsh = cpu_reference_run(raw, off, tlen, sample_n, threads)
            tot_n += n; tot_t += dt
        used, kind = threads, "port"
        sample = (f"{n} alignments per step in {nsh} QNAME-boundary shards, one in-memory oracle pipeline per thread "
                  f"(reference arithmetic restated in oracle/msam_oracle.c; the reference itself is single-threaded)")
    v = tot_n / tot_t / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "M alignments/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic community PE150, 100 genomes, filter -l 80 -p 95 -z 80 --besthit | profile --multi=proportional",
                       "records_per_step": n},
            "cpu_baseline": {"value": v, "unit": "M alignments/s", "cores": used, "kind": kind, "host_cores_available": cores, "sample": sample},
            "e2e": {"value": v, "unit": "M alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import msamtools_b200 as m
    rank, world, local = dist_env()
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    dist = None
    uid = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [m.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    raw, off, tlen, gen = make_batch(args.records, rank)
    n = len(off) - 1
    # kept=False: the profile is the filter stage's only consumer (the reference pipe's output is the profile, not the records)
    ctx = m.Context(profile=True, multi=MULTI, kept=False, n_targets=len(tlen), device=local, n_ranks=world, rank=rank,
                    nccl_unique_id=uid, **FILTER_OPTS)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident: batch uploaded once, steps run from HBM
    d_raw = ctx.device_alloc(raw.nbytes)
    d_off = ctx.device_alloc(off.nbytes)
    ctx.device_upload(d_raw, raw)
    ctx.device_upload(d_off, off)

    def step_resident():
        ctx.reset()
        ctx.push_device(d_raw, raw.nbytes, d_off, n)
        return ctx.finish_profile()

    for _ in range(args.warmup):
        ab, st = step_resident()
    ctx.timing(reset=True)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ctx.mark(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ab, st = step_resident()
    ctx.mark(1)
    dev_ms = ctx.elapsed_ms(0, 1)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    tim = ctx.timing(reset=True)
    step_ms = max_over_ranks(max(dev_ms, 0.0) / args.steps)
    wall_step_ms = max_over_ranks(wall_ms / args.steps)
    n_total = n * world if dist is None else int(max_over_ranks(0) or 0) or n * world
    if dist is not None:
        import torch
        t = torch.tensor([n], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_total = int(t.item())
    value = n_total / (step_ms * 1e-3) / 1e6

    # ---- end to end: host buffers through msg_push, H2D + D2H inside the timed region
    def step_e2e():
        ctx.reset()
        ctx.push(raw, off)
        return ctx.finish_profile()

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    ctx.timing(reset=True)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(e2e_steps):
        ab2, st2 = step_e2e()
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / e2e_steps)
    clocks = sampler.stop() if rank == 0 else None      # sampled across both timed regions (resident + e2e)
    tim2 = ctx.timing(reset=True)
    e2e_value = n_total / (e2e_ms * 1e-3) / 1e6

    ctx.device_free(d_raw); ctx.device_free(d_off)

    if rank == 0:
        peak, peak_src = peaks()
        dec_ms = tim["decode_ms"] / max(tim["decode_launches"], 1)
        alg_per_launch = tim["alg_bytes"] / max(tim["decode_launches"], 1)
        achieved = alg_per_launch / (dec_ms * 1e-3) / 1e9
        full_scan = (raw.nbytes + off.nbytes) / (dec_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "decode_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            if tj.get("records") == n:
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": "M alignments/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic community PE150, 100 genomes, filter -l 80 -p 95 -z 80 --besthit | profile --multi=proportional",
                       "records_per_gpu": n, "raw_bytes_per_gpu": int(raw.nbytes), "n_references": int(len(tlen)),
                       "l2": "inputs larger than L2 (%.1f GB batch vs 126 MB)" % (raw.nbytes / 1e9),
                       "kept_records": int(ctx.kept_count()), "inserts": int(st["mapped_inserts"]), "em_iterations": int(st["iterations"]),
                       "wall_ms_per_step": wall_step_ms},
            "roofline": {"bound": "hbm", "kernel": "decode_kernel (record decode + fused filter statistics)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": traffic, "alg_bytes_per_launch": alg_per_launch, "launch_ms": dec_ms,
                         "full_scan_gbs": full_scan, "kernel_share_of_step": dec_ms / step_ms},
            "e2e": {"value": e2e_value, "unit": "M alignments/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(tim2["h2d_bytes"] // e2e_steps), "d2h_bytes_per_step": int(tim2["d2h_bytes"] // e2e_steps),
                    "host_ingest_gbs": (tim2["h2d_bytes"] / e2e_steps) / (e2e_ms * 1e-3) / 1e9,
                    "bam_gbs": raw.nbytes / (e2e_ms * 1e-3) / 1e9,
                    "h2d_mode": ("zero-copy: decode windows pulled from the pinned host batch over PCIe (h2d bytes = window chunks requested"
                                 " + offset index DMA; SEQ/QUAL never leave the host)") if tim2.get("zero_copy_chunks", 0) else "staged: whole batch DMA"},
            "gpu_launches": int(tim["kernel_launches"]),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _ = cpu_baseline_entry(raw, off, tlen, args, n)
            one_n, one_dt, _ = cpu_reference_run(raw, off, tlen, min(1_000_000, n), 1)
            line["cpu_baseline"]["oracle_port_single_thread_value"] = one_n / one_dt / 1e6
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=10_000_000, help="alignments per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="time the in-memory oracle port even when oracle/_ref exists")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
