"""GPU, >= 2 devices: two ranks, each pushing its QNAME-boundary shard through its own context;
msg_finish_profile / msg_finish_coverage combine over NCCL and every rank must reproduce the
oracle's whole-stream result (integers exact, abundances within 1e-9 relative)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    try:
        import msamtools_b200 as m
        return m._lib.load().msg_device_count()
    except Exception:
        return 0


def _worker(rank, world, uid, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import msamtools_b200 as m
    from msamtools_b200 import synth, shard
    p = synth.make_params("mixed", n_records=200_000, seed=2024)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    cuts = shard.shard_bounds(raw, off, world)
    sraw, soff = shard.shard_view(raw, off, cuts, rank)
    with m.Context(l=80, p=95, z=80, besthit=True, profile=True, multi="proportional", coverage=True, n_targets=len(tlen),
                   target_len=tlen, device=rank, n_ranks=world, rank=rank, nccl_unique_id=uid) as ctx:
        ctx.push(sraw, soff)
        kept = ctx.kept_count()
        ab, st = ctx.finish_profile()
        cov, touched, total = ctx.finish_coverage()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), ab=ab, kept=kept, cov=cov, touched=touched, total=total,
             st=np.array([st["mapped_inserts"], st["uniq"], st["multi"], st["purged"], st["iterations"], st["n_lists"]]))


@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")
def test_two_gpu_profile_and_coverage(tmp_path, oracle):
    import msamtools_b200 as m
    from msamtools_b200 import synth
    uid = m.nccl_unique_id()
    mp.spawn(_worker, args=(2, uid, str(tmp_path)), nprocs=2, join=True)
    p = synth.make_params("mixed", n_records=200_000, seed=2024)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    idx = oracle.filter_stream(raw, off, oracle.filter_cfg(l=80, p=95, z=80, besthit=True))
    eab, est, _, _ = oracle.profile(raw, off, idx, len(tlen), 3)
    ecov = oracle.coverage(raw, off, idx, tlen)
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(2)]
    assert int(r[0]["kept"]) + int(r[1]["kept"]) == len(idx)
    for k in range(2):
        assert r[k]["st"].tolist() == [est["mapped_inserts"], est["uniq"], est["multi"], est["purged"], est["iterations"], est["n_lists"]]
        assert np.all(np.abs(r[k]["ab"] - eab) <= 1e-9 * np.maximum(np.abs(eab), np.abs(r[k]["ab"])))
        assert np.array_equal(r[k]["cov"], ecov[0]) and np.array_equal(r[k]["touched"], ecov[1]) and np.array_equal(r[k]["total"], ecov[2])
    assert np.array_equal(r[0]["ab"], r[1]["ab"])           # identical on every rank
