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


# ---------------------------------------------------------------- PropSharing across ranks: the reduce-scatter / all-gather algebra
def _multimapper_lists(raw, off, idx):
    """(doubled unique counts, multi-mapper feature lists) of one shard's kept stream, restated in Python from
    msam_profile.c:204-243 (grouping on tid != -1, QNAME vs the last counted record) and :65-200 (distinct features in
    first-appearance order; one feature -> +2, several -> a list in proportional mode)"""
    ui, lists = {}, []
    prev, group = None, []

    def close():
        if not group:
            return
        d = list(dict.fromkeys(group))
        if len(d) == 1:
            ui[d[0]] = ui.get(d[0], 0) + 2
        else:
            lists.append(d)

    for i in idx:
        o = int(off[i])
        tid = int(np.frombuffer(raw[o + 4:o + 8].tobytes(), dtype="<i4")[0])
        if tid == -1:
            continue
        lq = int(raw[o + 12])
        name = bytes(raw[o + 36:o + 36 + lq])
        if name != prev:
            close()
            group = []
            prev = name
        group.append(tid)
    close()
    return ui, lists


def _rsag_worker(rank, world, port, out_dir):
    """what em_loop_rsag_kernel does, with gloo standing in for the peer-memory stores: every rank owns a slice of the feature
    axis; partial increments are added IN RANK ORDER by the owner, which applies a = U + inc, the 1e-20 flush and its share
    of sum (a_new - a_old)^2 (msam_profile.c:369-380); slices and delta shares are then gathered by everyone"""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msamtools_b200 import synth, shard
    from oracle import oracle as orc
    p = synth.make_params("mixed", n_records=30_000, seed=4321)
    raw, off, _ = synth.generate(p)
    F = len(synth.target_lengths(p))
    cuts = shard.shard_bounds(raw, off, world)
    sraw, soff = shard.shard_view(raw, off, cuts, rank)
    idx = orc.filter_stream(sraw, soff, orc.filter_cfg(l=80, p=95, z=80, besthit=True))
    ui, lists = _multimapper_lists(sraw, soff, idx)
    S = (F + world - 1) // world                                   # slice length (the kernel rounds it up to whole blocks)
    lo, hi = rank * S, min(F, (rank + 1) * S)

    def exchange(vec, first, U_slice, a_old_slice):
        parts = [torch.zeros(F, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(vec))             # "every rank stores its partial slice into the owner's region"
        tot = np.zeros(hi - lo)
        for r in range(world):                                     # rank order: the same sum on every owner, every run
            tot = tot + parts[r].numpy()[lo:hi]
        if first:
            new, share = tot / 2, 0.0                              # :286
        else:
            new = U_slice + tot
            new[new < 1e-20] = 0.0                                 # :372-376
            share = float(np.sum((new - a_old_slice) ** 2))        # :377-379
        pad = np.zeros(S); pad[:hi - lo] = new
        slices = [torch.zeros(S, dtype=torch.float64) for _ in range(world)]
        shares = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(slices, torch.from_numpy(pad))            # all-gather of the new abundances
        dist.all_gather(shares, torch.tensor([share], dtype=torch.float64))
        a = np.concatenate([s.numpy() for s in slices])[:F]
        delta = 0.0
        for r in range(world):                                     # same order everywhere -> same stop decision everywhere
            delta += float(shares[r][0])
        return a, new, delta / F

    counts = np.zeros(F)
    for f, v in ui.items():
        counts[f] = v
    a, U_slice, _ = exchange(counts, True, None, None)
    iters, conv = 0, 0
    for k in range(1, 20):                                         # :331
        inc = np.zeros(F)
        for d in lists:                                            # :341-365
            s = 0.0
            for f in d:
                s += a[f]
            if s > 0:
                for f in d:
                    inc[f] += a[f] / s
        a, _, delta = exchange(inc, False, U_slice, a[lo:hi].copy())
        iters = k
        if delta < 1e-10:                                          # :383
            conv = 1
            break
    purged = torch.tensor([sum(1 for d in lists if sum(a[f] for f in d) == 0)], dtype=torch.int64)
    nl = torch.tensor([len(lists)], dtype=torch.int64)
    dist.all_reduce(purged); dist.all_reduce(nl)
    np.savez(os.path.join(out_dir, f"e{rank}.npz"), a=a, st=np.array([iters, conv, int(purged[0]), int(nl[0])]))
    dist.destroy_process_group()


def test_two_rank_rsag_propsharing_matches_whole(tmp_path, oracle):
    port = 29650 + os.getpid() % 200
    mp.spawn(_rsag_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from msamtools_b200 import synth
    p = synth.make_params("mixed", n_records=30_000, seed=4321)
    raw, off, _ = synth.generate(p)
    F = len(synth.target_lengths(p))
    idx = oracle.filter_stream(raw, off, oracle.filter_cfg(l=80, p=95, z=80, besthit=True))
    ab, st, _, _ = oracle.profile(raw, off, idx, F, 3)
    r = [np.load(tmp_path / f"e{k}.npz") for k in range(2)]
    assert np.array_equal(r[0]["a"], r[1]["a"])                    # bit-identical on both ranks
    assert r[0]["st"].tolist() == r[1]["st"].tolist() == [st["iterations"], st["converged"], st["purged"], st["n_lists"]]
    assert st["iterations"] >= 3 and st["n_lists"] > 100
    assert np.all(np.abs(r[0]["a"] - ab) <= 1e-9 * np.maximum(np.abs(ab), np.abs(r[0]["a"])))
