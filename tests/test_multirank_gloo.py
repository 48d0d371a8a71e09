"""CPU, world_size 2 over gloo: the multi-GPU host logic -- QNAME-boundary sharding plus the
additive combination that msg_finish_profile / msg_finish_coverage perform with NCCL -- gives the
same result as the unsharded stream.  Per-rank counting uses the CPU oracle (this is a test of the
sharding + collective algebra, not of the kernels)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, share_type, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msamtools_b200 import synth, shard
    from oracle import oracle as orc
    p = synth.make_params("mixed", n_records=30_000, seed=777)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    F = len(tlen)
    cuts = shard.shard_bounds(raw, off, world)
    sraw, soff = shard.shard_view(raw, off, cuts, rank)
    cfg = orc.filter_cfg(l=80, p=95, z=80, besthit=True)
    idx = orc.filter_stream(sraw, soff, cfg)
    # local counting (what every rank's GPU does with no collective)
    lib = orc.load()
    import ctypes as C
    h = lib.orc_profile_new(F, F, None, share_type)
    lib.orc_profile_push(h, sraw.ctypes.data_as(C.c_void_p), soff.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), len(idx))
    ui = np.zeros(F, dtype=np.uint32); d = np.zeros(F, dtype=np.float64)
    lib.orc_profile_counts(h, ui.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p))
    ab_l = np.zeros(F); po = orc.OrcProfileOut()
    lib.orc_profile_finish(h, ab_l.ctypes.data_as(C.c_void_p), C.byref(po))
    lib.orc_profile_free(h)
    # ONE allreduce of the count vectors + scalar counters (msg_finish_profile)
    t_ui = torch.from_numpy(ui.astype(np.int64)); t_d = torch.from_numpy(d.copy())
    t_cnt = torch.tensor([po.mapped_inserts, po.uniq, po.multi], dtype=torch.int64)
    for t in (t_ui, t_d, t_cnt):
        dist.all_reduce(t)
    U = t_ui.numpy() / 2.0 + (t_d.numpy() if share_type == 2 else 0)
    kept = torch.tensor([len(idx)]); dist.all_reduce(kept)
    cov = orc.coverage(sraw, soff, idx, tlen)
    t_touch_in = torch.from_numpy(cov[2].copy()); dist.all_reduce(t_touch_in)       # sums are additive across shards
    if rank == 0:
        np.savez(os.path.join(out_dir, "r.npz"), U=U, cnt=t_cnt.numpy(), kept=kept.numpy(), cuts=np.array(cuts), covsum=t_touch_in.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("share_type", [1, 2])
def test_two_rank_sharding_matches_whole(tmp_path, oracle, share_type):
    port = 29600 + share_type + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, share_type, str(tmp_path)), nprocs=2, join=True)
    from msamtools_b200 import synth
    r = np.load(tmp_path / "r.npz")
    p = synth.make_params("mixed", n_records=30_000, seed=777)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    idx = oracle.filter_stream(raw, off, oracle.filter_cfg(l=80, p=95, z=80, besthit=True))
    ab, st, ui, d = oracle.profile(raw, off, idx, len(tlen), share_type)
    assert int(r["kept"][0]) == len(idx)
    assert r["cnt"].tolist() == [st["mapped_inserts"], st["uniq"], st["multi"]]
    assert np.allclose(r["U"], ab, rtol=1e-12, atol=0)
    cov = oracle.coverage(raw, off, idx, tlen)
    assert np.array_equal(r["covsum"], cov[2])
    cuts = r["cuts"].tolist()
    assert cuts[0] == 0 and cuts[-1] == len(off) - 1 and 0 < cuts[1] < cuts[-1]
    # no QNAME group straddles the cut
    o0, o1 = int(off[cuts[1] - 1]), int(off[cuts[1]])
    assert bytes(raw[o0 + 36:o0 + 36 + int(raw[o0 + 12])]) != bytes(raw[o1 + 36:o1 + 36 + int(raw[o1 + 12])])
