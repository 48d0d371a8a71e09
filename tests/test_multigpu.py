"""GPU, >= 2 devices: WORLD ranks (default 2; MSG_TEST_WORLD=4|8 on bigger boxes), each pushing its QNAME-boundary shard
through its own context; msg_finish_profile / msg_finish_coverage combine the ranks (peer-memory exchange inside the
PropSharing kernel, NCCL for the other modes) and every rank must reproduce the oracle's whole-stream result (integers
exact, abundances within 1e-9 relative, bit-identical vectors on all ranks)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = int(os.environ.get("MSG_TEST_WORLD", "2"))


def _ndev():
    try:
        import msamtools_b200 as m
        return m._lib.load().msg_device_count()
    except Exception:
        return 0


CASES = {
    # name: (preset, env, share mode, coverage?)
    "p2p_prop_cov": ("mixed", {}, "proportional", True),          # general pipeline (coverage keeps the kept stream)
    "p2p_prop_fused": ("mixed", {}, "proportional", False),       # fused filter->profile pass, exchange by CTA 0
    "p2p_prop_bigF": ("catalog10k", {}, "proportional", False),   # F + 8 > 8192: reduce-scatter / all-gather exchange (em_loop_rsag_kernel)
    "p2p_prop_genes1m": ("genes1m", {}, "proportional", False),   # 1 M-gene catalogue, same kernel, 123 blocks per slice
    "p2p_prop_mid": ("community", {"_n_refs": "3000"}, "proportional", False),   # 2048 < F <= 8184: flagged words by CTA 0, global-memory a[]
    "nccl_prop": ("mixed", {"MSG_NO_P2P": "1"}, "proportional", False),   # NCCL allreduce + per-iteration NCCL loop
    "equal": ("mixed", {}, "equal", False),                        # packed u32 allreduce + f64 allreduce
}


def _case_data(case):
    from msamtools_b200 import synth
    preset = CASES[case][0]
    over = {}
    if "_n_refs" in CASES[case][1]:
        over = dict(n_refs=int(CASES[case][1]["_n_refs"]), ref_len_min=2_000, ref_len_max=9_000, shared_fraction=0.3)
    p = synth.make_params(preset, n_records=200_000, seed=2024, **over)
    raw, off, _ = synth.generate(p)
    return raw, off, synth.target_lengths(p)


def _worker(rank, world, uid, out_dir, case):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    _, env, mode, cov_on = CASES[case]
    os.environ.update({k: v for k, v in env.items() if not k.startswith("_")})
    import msamtools_b200 as m
    from msamtools_b200 import shard
    raw, off, tlen = _case_data(case)
    cuts = shard.shard_bounds(raw, off, world)
    sraw, soff = shard.shard_view(raw, off, cuts, rank)
    with m.Context(l=80, p=95, z=80, besthit=True, profile=True, multi=mode, coverage=cov_on, kept=cov_on, n_targets=len(tlen),
                   target_len=tlen, device=rank, n_ranks=world, rank=rank, nccl_unique_id=uid) as ctx:
        ctx.push(sraw, soff)
        kept = ctx.kept_count()
        ab, st = ctx.finish_profile()
        ab2, st2 = ctx.finish_profile()          # a second finish on the same state: next epoch of the peer exchange
        # (not bit-identical: the order of the floating-point atomics inside one GPU differs from run to run)
        assert np.allclose(ab, ab2, rtol=1e-11, atol=0) and st2["iterations"] == st["iterations"] and st2["purged"] == st["purged"]
        cov = touched = total = np.zeros(0)
        if cov_on:
            cov, touched, total = ctx.finish_coverage()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), ab=ab, kept=kept, cov=cov, touched=touched, total=total,
             st=np.array([st["mapped_inserts"], st["uniq"], st["multi"], st["purged"], st["iterations"], st["n_lists"]]))


@pytest.mark.skipif(_ndev() < WORLD, reason=f"needs {WORLD} GPUs")
@pytest.mark.parametrize("case", list(CASES))
def test_two_gpu_profile_and_coverage(tmp_path, oracle, case):
    import msamtools_b200 as m
    uid = m.nccl_unique_id()
    mp.spawn(_worker, args=(WORLD, uid, str(tmp_path), case), nprocs=WORLD, join=True)
    raw, off, tlen = _case_data(case)
    _, _, mode, cov_on = CASES[case]
    idx = oracle.filter_stream(raw, off, oracle.filter_cfg(l=80, p=95, z=80, besthit=True))
    eab, est, _, _ = oracle.profile(raw, off, idx, len(tlen), {"all": 1, "equal": 2, "proportional": 3}[mode])
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(WORLD)]
    assert sum(int(x["kept"]) for x in r) == len(idx)
    for k in range(WORLD):
        assert r[k]["st"].tolist() == [est["mapped_inserts"], est["uniq"], est["multi"], est["purged"], est["iterations"], est["n_lists"]]
        assert np.all(np.abs(r[k]["ab"] - eab) <= 1e-9 * np.maximum(np.abs(eab), np.abs(r[k]["ab"])))
    if cov_on:
        ecov = oracle.coverage(raw, off, idx, tlen)
        for k in range(WORLD):
            assert np.array_equal(r[k]["cov"], ecov[0]) and np.array_equal(r[k]["touched"], ecov[1]) and np.array_equal(r[k]["total"], ecov[2])
    for k in range(1, WORLD):
        assert np.array_equal(r[0]["ab"], r[k]["ab"])       # identical on every rank


def _skew_worker(rank, world, uid, out_dir):
    """rank 1 holds NO records (zero lists, zero counts) and arrives 4 s late at msg_finish_profile: the ranks are lined up
    before the cooperative kernel starts, so nobody times out and rank 1 still gets the full result"""
    import time
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ["MSG_PEER_TIMEOUT_S"] = "2"          # far below the skew: only the pre-launch line-up can make this pass
    import msamtools_b200 as m
    raw, off, tlen = _case_data("p2p_prop_bigF")
    with m.Context(l=80, p=95, z=80, besthit=True, profile=True, multi="proportional", kept=False, n_targets=len(tlen),
                   device=rank, n_ranks=world, rank=rank, nccl_unique_id=uid) as ctx:
        if rank == 0:
            ctx.push(raw, off)
        elif rank == world - 1:
            time.sleep(4.0)
        ab, st = ctx.finish_profile()
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), ab=ab,
             st=np.array([st["mapped_inserts"], st["uniq"], st["multi"], st["purged"], st["iterations"], st["n_lists"]]))


@pytest.mark.skipif(_ndev() < WORLD, reason=f"needs {WORLD} GPUs")
def test_two_gpu_late_and_empty_rank(tmp_path, oracle):
    import msamtools_b200 as m
    uid = m.nccl_unique_id()
    mp.spawn(_skew_worker, args=(WORLD, uid, str(tmp_path)), nprocs=WORLD, join=True)
    raw, off, tlen = _case_data("p2p_prop_bigF")
    idx = oracle.filter_stream(raw, off, oracle.filter_cfg(l=80, p=95, z=80, besthit=True))
    eab, est, _, _ = oracle.profile(raw, off, idx, len(tlen), 3)
    r = [np.load(tmp_path / f"s{k}.npz") for k in range(WORLD)]
    for k in range(WORLD):
        assert r[k]["st"].tolist() == [est["mapped_inserts"], est["uniq"], est["multi"], est["purged"], est["iterations"], est["n_lists"]]
        assert np.all(np.abs(r[k]["ab"] - eab) <= 1e-9 * np.maximum(np.abs(eab), np.abs(r[k]["ab"])))
        assert np.array_equal(r[0]["ab"], r[k]["ab"])
