"""GPU: msg_push_async / msg_wait (double-buffered chunk pipeline) give the same results as the synchronous
push and as the CPU oracle -- staged chunks (two device slots, copy stream), zero-copy chunks (three rotating pinned
buffers), device-resident chunks, a guard-declined chunk with another chunk already queued behind it, and deferred
error reporting."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OPTS = dict(l=80, p=95, z=80, besthit=True)


@pytest.fixture(scope="module")
def m():
    import msamtools_b200 as mod
    assert mod._lib.load().msg_device_count() > 0, "no CUDA device: the GPU tests must not silently pass"
    return mod


def close(a, b, rel=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= rel * np.maximum(np.abs(a), np.abs(b)))


@pytest.fixture(scope="module")
def data(oracle):
    from msamtools_b200 import synth
    import msamtools_b200 as mod
    p = synth.make_params("mixed", n_records=240_000, seed=4242)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    n = len(off) - 1
    cuts = [0] + [mod.split_point(raw, off, n * k // 7) for k in range(1, 7)] + [n]
    chunks = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        lo, hi = int(off[a]), int(off[b])
        chunks.append((np.ascontiguousarray(raw[lo:hi]), np.ascontiguousarray(off[a:b + 1] - off[a])))
    idx = oracle.filter_stream(raw, off, oracle.filter_cfg(**OPTS))
    exp = {mode: oracle.profile(raw, off, idx, len(tlen), st) for mode, st in (("proportional", 3), ("equal", 2))}
    return raw, off, tlen, chunks, idx, exp


def check(ab, st, ui, kept, exp, idx):
    eab, est, eui, ed = exp
    for k in ("mapped_inserts", "uniq", "multi", "purged", "iterations", "converged", "n_lists", "n_entries"):
        assert st[k] == est[k], k
    assert np.array_equal(ui, eui) and close(ab, eab)
    if kept is not None:
        assert kept == len(idx)


@pytest.mark.parametrize("mode", ["proportional", "equal"])
def test_async_staged_pageable(m, data, mode):
    raw, off, tlen, chunks, idx, exp = data
    with m.Context(profile=True, multi=mode, kept=False, n_targets=len(tlen), **OPTS) as ctx:
        for rep in range(2):                                    # second pass reuses both slots and the grown list storage
            ctx.reset(); ctx.timing(reset=True)
            for r, o in chunks:
                ctx.push_async(r, o)
            ctx.wait()
            ui, _ = ctx.pull_counts()
            ab, st = ctx.finish_profile()
            t = ctx.timing()
            assert (t["fused_chunks"], t["fused_fallbacks"], t["zero_copy_chunks"]) == (len(chunks), 0, 0)
            check(ab, st, ui, None, exp[mode], idx)


def test_async_zero_copy_three_buffers(m, data):
    raw, off, tlen, chunks, idx, exp = data
    cap = max(len(r) for r, _ in chunks)
    bufs = [m.PinnedBuffer(cap) for _ in range(3)]
    try:
        with m.Context(profile=True, multi="proportional", kept=False, n_targets=len(tlen), **OPTS) as ctx:
            kept = 0
            for k, (r, o) in enumerate(chunks):
                b = bufs[k % 3].array                              # the buffer of chunk k-3: completed when push k-1 returned
                b[:len(r)] = r
                ctx.push_async(b[:len(r)], o)
            ctx.wait()
            ui, _ = ctx.pull_counts()
            ab, st = ctx.finish_profile()
            t = ctx.timing()
        assert (t["zero_copy_chunks"], t["fused_chunks"], t["fused_fallbacks"]) == (len(chunks), len(chunks), 0)
        assert t["h2d_bytes"] < len(raw)
        check(ab, st, ui, None, exp["proportional"], idx)
    finally:
        for b in bufs:
            b.close()


def test_async_device_chunks_and_kept_counts(m, data):
    raw, off, tlen, chunks, idx, exp = data
    with m.Context(profile=True, multi="proportional", kept=False, n_targets=len(tlen), **OPTS) as ctx:
        dev = []
        for r, o in chunks:
            d_raw, d_off = ctx.device_alloc(r.nbytes), ctx.device_alloc(o.nbytes)
            ctx.device_upload(d_raw, r); ctx.device_upload(d_off, o)
            dev.append((d_raw, r.nbytes, d_off, len(o) - 1))
        for d in dev:
            ctx.push_device_async(*d)
        ui, _ = ctx.pull_counts()                                  # completes everything in flight
        ab, st = ctx.finish_profile()
        check(ab, st, ui, None, exp["proportional"], idx)
        # synchronous pushes on the same context: per-chunk kept counts add up
        ctx.reset()
        kept = 0
        for d in dev:
            ctx.push_device(*d)
            kept += ctx.kept_count()
        ab, st = ctx.finish_profile()
        ui, _ = ctx.pull_counts()
        check(ab, st, ui, kept, exp["proportional"], idx)
        for d in dev:
            ctx.device_free(d[0]); ctx.device_free(d[2])


def test_async_guard_fallback_with_chunk_behind(m, oracle):
    """chunk 1 makes the fused guard trip (x | dropped y | x again: one insert for the profile, two pools for the filter)
    while chunk 2 is already queued behind it: chunk 1 is rerun on the general pipeline, totals match the oracle."""
    import samutil
    hdr = "@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:A\tLN:1000\n@SQ\tSN:B\tLN:1000\n@SQ\tSN:C\tLN:500\n"
    a = "\t0\t%s\t10\t60\t%dM\t*\t0\t0\t%s\t%s\tNM:i:0\tAS:i:%d"
    rec = lambda n, ref, ln, sc: n + a % (ref, ln, "A" * ln, "I" * ln, sc)
    texts = [
        [rec("a1", "A", 20, 9), rec("a2", "B", 20, 9), rec("a2", "C", 20, 9)],
        [rec("u", "C", 20, 9), rec("x", "A", 20, 10), rec("y", "A", 5, 5), rec("x", "B", 20, 20), rec("z", "C", 20, 7)],
        [rec("q1", "B", 20, 9), rec("q2", "A", 20, 3), rec("q2", "B", 20, 3), rec("q3", "C", 20, 1)],
        [rec("r1", "B", 20, 9), rec("r1", "C", 20, 9)],
    ]
    parts = [samutil.parse_sam_text(hdr + "\n".join(t) + "\n") for t in texts]
    whole = samutil.parse_sam_text(hdr + "\n".join(sum(texts, [])) + "\n")
    opts = dict(l=8, besthit=True)
    idx = oracle.filter_stream(whole.raw, whole.off, oracle.filter_cfg(**opts))
    for mode, stype in (("proportional", 3), ("equal", 2), ("all", 1)):
        eab, est, eui, ed = oracle.profile(whole.raw, whole.off, idx, 3, stype)
        with m.Context(profile=True, multi=mode, kept=False, n_targets=3, **opts) as ctx:
            for s in parts:
                ctx.push_async(np.ascontiguousarray(s.raw), np.ascontiguousarray(s.off))
            ctx.wait()
            ui, d = ctx.pull_counts()
            ab, st = ctx.finish_profile()
            t = ctx.timing()
        assert (t["fused_chunks"], t["fused_fallbacks"]) == (4, 1)
        assert ui.tolist() == eui.tolist() and close(ab, eab) and close(d, ed)
        for k in ("mapped_inserts", "uniq", "multi", "purged", "n_lists", "n_entries"):
            assert st[k] == est[k], (mode, k)


def test_async_deferred_error_and_recovery(m):
    """a chunk without AS on a best-hit candidate: the error surfaces at the call that completes the chunk
    (msam_filter.c:219-221 text), the error word is cleared, and the context keeps working after msg_reset."""
    import samutil
    hdr = "@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:A\tLN:1000\n"
    good = samutil.parse_sam_text(hdr + "r1\t0\tA\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:5\n")
    bad = samutil.parse_sam_text(hdr + "r2\t0\tA\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\n")
    with m.Context(profile=True, multi="all", kept=False, besthit=True, n_targets=1) as ctx:
        ctx.push_async(good.raw, good.off)
        ctx.push_async(bad.raw, bad.off)                           # queued; nothing reported yet
        with pytest.raises(m.MsgError) as e:
            ctx.wait()
        assert e.value.code == m._lib.MSG_ENOAS and "Required field AS not found" in e.value.text
        ctx.reset()
        ctx.push_async(good.raw, good.off)
        ctx.wait()
        ab, st = ctx.finish_profile()
        assert st["mapped_inserts"] == 1 and ab.tolist() == [1.0]


def test_malformed_offset_index_is_reported(m):
    """rec_off[] that is not monotonic / runs past the chunk: MSG_EFORMAT, no out-of-bounds access (run under compute-sanitizer)"""
    import samutil
    hdr = "@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:A\tLN:1000\n"
    s = samutil.parse_sam_text(hdr + "".join(f"r{k}\t0\tA\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:5\n" for k in range(6)))
    for which in ("past_end", "backwards"):
        off = s.off.copy()
        if which == "past_end":
            off[-1] = len(s.raw) + 4096
        else:
            off[3] = off[1]
        for kw in (dict(l=5, records=True), dict(besthit=True, profile=True, multi="all", kept=False), dict(profile=True, multi="all")):
            with m.Context(n_targets=1, **kw) as ctx:
                with pytest.raises(m.MsgError) as e:
                    ctx.push(s.raw, off)
                assert e.value.code == m._lib.MSG_EFORMAT
                ctx.reset()
                ctx.push(s.raw, s.off)                             # the error word does not stick
