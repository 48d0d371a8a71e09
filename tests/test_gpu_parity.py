"""GPU parity tests proper: the CUDA path, called through the C ABI, against the committed golden
vectors and against the CPU oracle on the same seeded inputs.  Integer/byte/index results must be
bit-exact; float64 abundances within 1e-9 relative (BASELINE.json north_star).
"""
import numpy as np
import pytest

import goldenutil as G

pytestmark = pytest.mark.gpu

SHARE = {"all": 1, "equal": 2, "proportional": 3, "ignore": 4}
REL = 1e-9


@pytest.fixture(scope="module")
def m():
    import msamtools_b200 as mod
    assert mod._lib.load().msg_device_count() > 0, "no CUDA device: the GPU tests must not silently pass"
    return mod


def close(a, b, rel=REL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= rel * np.maximum(np.abs(a), np.abs(b)))


def gpu_filter(m, s, opts, force_slow=False, records=True):
    with m.Context(stats=True, records=records, n_targets=len(s.ref_names), force_slow=force_slow, **opts) as ctx:
        ctx.push(s.raw, s.off)
        kept = ctx.pull_kept()
        st = ctx.pull_stats()
        rec = ctx.pull_records()[0] if records else None
    return kept, st, rec


@pytest.mark.parametrize("force_slow", [False, True], ids=["staged", "slowpath"])
@pytest.mark.parametrize("c", G.cases("filter"), ids=G.case_id)
def test_golden_filter(m, c, force_slow):
    s = G.fixture(c["fixture"])
    kept, st, rec = gpu_filter(m, s, c["opts"], force_slow)
    assert s.name_flags(kept) == c["name_flags"]
    assert kept.tolist() == c["kept"]
    assert bytes(rec).hex() == c["records_hex"]
    need_stats = any(k in c["opts"] for k in ("l", "p", "ppt", "z", "rescore"))
    mapped = np.array([not (s.record_fields(i)[1] & 4) for i in range(s.n)])
    if need_stats:
        for k in ("alen", "qlen", "qclip", "edit"):
            assert st[k][mapped].tolist() == np.array(c["stats"][k])[mapped].tolist(), k
    if not c["opts"].get("rescore"):
        has_as = np.array(c["stats"]["has_as"], dtype=bool)
        assert st["score"][has_as].tolist() == np.array(c["stats"]["score"])[has_as].tolist()
        assert ((st["flags"] & 2) != 0).tolist() == has_as.tolist()


@pytest.mark.parametrize("force_slow,kept", [(False, True), (True, True), (False, False)], ids=["staged", "slowpath", "fused"])
@pytest.mark.parametrize("c", G.cases("profile"), ids=G.case_id)
def test_golden_profile(m, c, force_slow, kept):
    s = G.fixture(c["fixture"])
    pre = c["pre"] or {}
    with m.Context(profile=True, multi=c["mode"], kept=kept, n_targets=len(s.ref_names), force_slow=force_slow, **pre) as ctx:
        ctx.push(s.raw, s.off)
        ui, d = ctx.pull_counts()
        ab, st = ctx.finish_profile()
    exp = c["stats"]
    for k in ("mapped_inserts", "uniq", "multi", "purged", "iterations", "converged", "n_lists", "n_entries"):
        assert st[k] == exp[k], k
    assert ui.tolist() == G.dense(c, "ui", np.uint32).tolist()
    assert close(ab, G.dense(c, "abundance", np.float64))
    assert close(st["delta"], exp["delta"], 1e-6) or (np.array(exp["delta"]) < 1e-12).all()


@pytest.mark.parametrize("summary", [False, True], ids=["depth", "summary"])
@pytest.mark.parametrize("c", G.cases("coverage"), ids=G.case_id)
def test_golden_coverage(m, c, summary):
    s = G.fixture(c["fixture"])
    pre = c["pre"] or {}
    with m.Context(coverage=True, coverage_summary=summary, n_targets=len(s.ref_names), target_len=s.target_len, **pre) as ctx:
        ctx.push(s.raw, s.off)
        cov, touched, total = ctx.finish_coverage()
        assert np.nonzero(cov)[0].tolist() == c["nz"]
        assert touched.tolist() == G.dense(c, "touched", np.int64).tolist()
        assert total.tolist() == G.dense(c, "sum", np.int64).tolist()
        if summary:
            with pytest.raises(m.MsgError):
                ctx.pull_coverage(0)
        elif c["depth"] is not None:
            for t, dexp in enumerate(c["depth"]):
                assert ctx.pull_coverage(t).tolist() == dexp


# ---------------------------------------------------------------- seeded synthetic streams vs the oracle
FILTER_OPTS = [
    dict(l=80, p=95, z=80), dict(l=80, p=95, z=80, besthit=True), dict(l=80, p=95, z=80, uniqhit=True),
    dict(besthit=True), dict(uniqhit=True), dict(p=99, invert=True), dict(p=99, invert=True, keep_unmapped=True),
    dict(ppt=-990), dict(z=90), dict(l=140), dict(l=100, rescore=True, besthit=True), dict(rescore=True, uniqhit=True),
]


@pytest.fixture(scope="module", params=["mixed", "clippy"])
def mixed(request):
    """two seeded synthetic streams: the `mixed` preset, and a 12-reference one with heavy clipping / indels, unmapped
    pairs and up to 12 occurrences per insert (same shape as tests/test_ref_binary.py pins the oracle on)"""
    from msamtools_b200 import synth
    if request.param == "mixed":
        p = synth.make_params("mixed", n_records=60_000, seed=24680)
    else:
        p = synth.make_params("community", n_records=50_000, seed=7, n_refs=12, ref_len_min=3_000, ref_len_max=9_000, clip_fraction=0.35,
                              indel_fraction=0.25, unmapped_fraction=0.06, shared_fraction=0.45, single_fraction=0.15, max_occ=12)
    raw, off, st = synth.generate(p)
    return raw, off, synth.target_lengths(p), p


@pytest.mark.parametrize("opts", FILTER_OPTS, ids=lambda o: ",".join(f"{k}={v}" for k, v in o.items()))
def test_synth_filter(m, oracle, mixed, opts):
    raw, off, tlen, p = mixed
    cfg = oracle.filter_cfg(**opts)
    exp = oracle.filter_stream(raw, off, cfg)
    with m.Context(records=True, n_targets=len(tlen), **opts) as ctx:
        ctx.push(raw, off)
        kept = ctx.pull_kept()
        rec, nrec = ctx.pull_records()
    assert np.array_equal(kept, exp)
    assert 0 < len(exp) < len(off) - 1
    assert bytes(rec) == bytes(oracle.emit_records(raw, off, exp, cfg))


def test_synth_stats(m, oracle, mixed):
    raw, off, tlen, p = mixed
    exp = oracle.record_stats(raw, off)
    for slow in (False, True):
        with m.Context(l=1, stats=True, n_targets=len(tlen), force_slow=slow) as ctx:
            ctx.push(raw, off)
            st = ctx.pull_stats()
        tagged = exp["has_tag"] > 0
        for k in ("alen", "qlen", "qclip", "edit"):
            assert np.array_equal(st[k][tagged], exp[k][tagged]), k
        assert np.array_equal((st["flags"] & 8) != 0, np.full(len(off) - 1, slow))


@pytest.mark.parametrize("mode", ["all", "equal", "proportional", "ignore"])
@pytest.mark.parametrize("pre", [None, dict(l=80, p=95, z=80, besthit=True), dict(p=97)], ids=["plain", "besthit", "p97"])
def test_synth_profile(m, oracle, mixed, mode, pre):
    raw, off, tlen, p = mixed
    idx = None if pre is None else oracle.filter_stream(raw, off, oracle.filter_cfg(**pre))
    eab, est, eui, ed = oracle.profile(raw, off, idx, len(tlen), SHARE[mode])
    with m.Context(profile=True, multi=mode, n_targets=len(tlen), **(pre or {})) as ctx:
        ctx.push(raw, off)
        ui, d = ctx.pull_counts()
        ab, st = ctx.finish_profile()
    for k in ("mapped_inserts", "uniq", "multi", "purged", "iterations", "converged", "n_lists", "n_entries"):
        assert st[k] == est[k], k
    assert np.array_equal(ui, eui)
    assert close(d, ed) and close(ab, eab)
    assert est["multi"] > 0 and est["uniq"] > 0


FUSED_PRE = [dict(l=80, p=95, z=80, besthit=True), dict(besthit=True), dict(uniqhit=True), dict(p=97, uniqhit=True),
             dict(l=100, rescore=True, besthit=True)]


@pytest.mark.parametrize("mode", ["all", "equal", "proportional", "ignore"])
@pytest.mark.parametrize("pre", FUSED_PRE, ids=lambda o: ",".join(f"{k}={v}" for k, v in o.items()))
def test_synth_profile_fused(m, oracle, mixed, mode, pre):
    """kept=False: best-hit + profile run as the fused single pass (fused.cuh); same results as the oracle."""
    raw, off, tlen, p = mixed
    idx = oracle.filter_stream(raw, off, oracle.filter_cfg(**pre))
    eab, est, eui, ed = oracle.profile(raw, off, idx, len(tlen), SHARE[mode])
    with m.Context(profile=True, multi=mode, kept=False, n_targets=len(tlen), **pre) as ctx:
        ctx.push(raw, off)
        assert ctx.kept_count() == len(idx)
        ui, d = ctx.pull_counts()
        ab, st = ctx.finish_profile()
        t = ctx.timing()
    assert t["fused_chunks"] == 1 and t["fused_fallbacks"] == 0
    for k in ("mapped_inserts", "uniq", "multi", "purged", "iterations", "converged", "n_lists", "n_entries"):
        assert st[k] == est[k], k
    assert np.array_equal(ui, eui)
    assert close(d, ed) and close(ab, eab)


def test_fused_chunked_and_genome_map(m, oracle, mixed):
    raw, off, tlen, p = mixed
    n = len(off) - 1
    opts = dict(l=80, p=95, z=80, besthit=True)
    fmap = (np.arange(len(tlen)) % 7).astype(np.int32)
    eidx = oracle.filter_stream(raw, off, oracle.filter_cfg(**opts))
    eab, est, eui, _ = oracle.profile(raw, off, eidx, len(tlen), 3, fmap=fmap, n_features=7)
    cuts = [0, m.split_point(raw, off, n // 4), m.split_point(raw, off, n // 2), n]
    with m.Context(profile=True, multi="proportional", kept=False, n_targets=len(tlen), n_features=7, fmap=fmap, **opts) as ctx:
        kept = 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            ctx.push(raw[int(off[a]):int(off[b])], off[a:b + 1] - off[a])
            kept += ctx.kept_count()
        ab, st = ctx.finish_profile()
        ui, _ = ctx.pull_counts()
        assert ctx.timing()["fused_chunks"] == 3
    assert kept == len(eidx) and np.array_equal(ui, eui) and close(ab, eab)
    assert (st["mapped_inserts"], st["uniq"], st["multi"], st["purged"], st["n_lists"], st["n_entries"]) == \
           (est["mapped_inserts"], est["uniq"], est["multi"], est["purged"], est["n_lists"], est["n_entries"])


def test_synth_profile_genome_map(m, oracle, mixed):
    # --genome style feature map: several sequences per feature (msam_profile.c:764-843)
    raw, off, tlen, p = mixed
    fmap = (np.arange(len(tlen)) % 7).astype(np.int32)
    eab, est, eui, ed = oracle.profile(raw, off, None, len(tlen), 3, fmap=fmap, n_features=7)
    with m.Context(profile=True, multi="prop", n_targets=len(tlen), n_features=7, fmap=fmap) as ctx:
        ctx.push(raw, off)
        ui, _ = ctx.pull_counts()
        ab, st = ctx.finish_profile()
    assert np.array_equal(ui, eui) and close(ab, eab)
    assert (st["uniq"], st["multi"], st["purged"], st["iterations"]) == (est["uniq"], est["multi"], est["purged"], est["iterations"])


@pytest.mark.parametrize("summary", [False, True], ids=["depth", "summary"])
@pytest.mark.parametrize("pre", [None, dict(l=80, p=95, z=80), dict(l=80, p=95, z=80, besthit=True)], ids=["plain", "fused", "besthit"])
def test_synth_coverage(m, oracle, mixed, pre, summary):
    raw, off, tlen, p = mixed
    idx = None if pre is None else oracle.filter_stream(raw, off, oracle.filter_cfg(**pre))
    ecov, et, es, edepth = oracle.coverage(raw, off, idx, tlen, want_depth=True)
    with m.Context(coverage=True, coverage_summary=summary, n_targets=len(tlen), target_len=tlen, **(pre or {})) as ctx:
        for rep in range(2):                                     # the second pass checks that msg_reset clears the bitmap / sums
            ctx.reset()
            ctx.push(raw, off)
            cov, touched, total = ctx.finish_coverage()
            assert np.array_equal(cov, ecov) and np.array_equal(touched, et) and np.array_equal(total, es)
        if not summary:
            for t in (0, len(tlen) // 2, len(tlen) - 1):
                assert np.array_equal(ctx.pull_coverage(t), edepth[t])


def test_chunked_push_equals_single(m, oracle, mixed):
    # chunk boundaries from msg_split_point: no QNAME group straddles chunks
    raw, off, tlen, p = mixed
    n = len(off) - 1
    opts = dict(l=80, p=95, z=80, besthit=True)
    eidx = oracle.filter_stream(raw, off, oracle.filter_cfg(**opts))
    eab, est, eui, _ = oracle.profile(raw, off, eidx, len(tlen), 3)
    ecov = oracle.coverage(raw, off, eidx, tlen)
    cuts = [0]
    for want in (n // 3, 2 * n // 3):
        k = m.split_point(raw, off, want)
        assert 0 < k <= want
        cuts.append(k)
    cuts.append(n)
    kept_all = []
    with m.Context(profile=True, coverage=True, multi="proportional", n_targets=len(tlen), target_len=tlen, **opts) as ctx:
        for a, b in zip(cuts[:-1], cuts[1:]):
            lo, hi = int(off[a]), int(off[b])
            ctx.push(raw[lo:hi], off[a:b + 1] - off[a])
            kept_all.append(ctx.pull_kept().astype(np.int64) + a)
        ab, st = ctx.finish_profile()
        ui, _ = ctx.pull_counts()
        cov = ctx.finish_coverage()
    assert np.array_equal(np.concatenate(kept_all), eidx)
    assert np.array_equal(ui, eui) and close(ab, eab)
    assert (st["mapped_inserts"], st["uniq"], st["multi"], st["purged"]) == (est["mapped_inserts"], est["uniq"], est["multi"], est["purged"])
    for got, exp in zip(cov, ecov[:3]):
        assert np.array_equal(got, exp)


# ---------------------------------------------------------------- edge cases
def _sam(text):
    import samutil
    return samutil.parse_sam_text(text)


HDR = "@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:A\tLN:1000\n@SQ\tSN:B\tLN:1000\n@SQ\tSN:C\tLN:500\n"


def test_empty_and_unmapped(m):
    s = _sam(HDR)
    with m.Context(l=10, besthit=True, profile=True, coverage=True, n_targets=3, target_len=s.target_len) as ctx:
        ctx.push(s.raw, s.off)
        assert ctx.kept_count() == 0
        ab, st = ctx.finish_profile()
        assert not ab.any() and st["mapped_inserts"] == 0
        cov, t, sm = ctx.finish_coverage()
        assert not cov.any()
    s = _sam(HDR + "u1\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\nu2\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\tIIII\n")
    with m.Context(profile=True, multi="equal", n_targets=3) as ctx:
        ctx.push(s.raw, s.off)
        ab, st = ctx.finish_profile()
        assert not ab.any() and (st["mapped_inserts"], st["uniq"], st["multi"]) == (0, 0, 0)


def test_missing_tags_errors(m, oracle):
    s = _sam(HDR + "r1\t0\tA\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tAS:i:5\n")
    with m.Context(l=5, n_targets=3) as ctx:                      # msam_filter.c:150-152
        with pytest.raises(m.MsgError) as e:
            ctx.push(s.raw, s.off)
        assert e.value.code == m._lib.MSG_ENOTAG and "Either NM or MD must be present" in e.value.text
    with pytest.raises(oracle.OracleError) as oe:
        oracle.filter_stream(s.raw, s.off, oracle.filter_cfg(l=5))
    assert oe.value.code == oracle.ORC_ENOTAG
    with m.Context(besthit=True, n_targets=3) as ctx:             # plain --besthit never parses NM/MD (:104,145)
        ctx.push(s.raw, s.off)
        assert ctx.pull_kept().tolist() == [0]
    s = _sam(HDR + "r1\t0\tA\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\n")
    with m.Context(besthit=True, n_targets=3) as ctx:             # msam_filter.c:219-221
        with pytest.raises(m.MsgError) as e:
            ctx.push(s.raw, s.off)
        assert e.value.code == m._lib.MSG_ENOAS and "Required field AS not found" in e.value.text


def test_aux_types_and_long_fields(m, oracle):
    # every aux type before the tags we need, long Z/B fields (slow path), MD edge forms
    seq, q = "A" * 50, "I" * 50
    recs = [
        f"a\t0\tA\t10\t60\t50M\t*\t0\t0\t{seq}\t{q}\tXA:A:x\tXc:i:-3\tXs:i:-300\tXi:i:-70000\tXI:i:3000000000\tXf:f:1.5\tXB:B:s,1,2,3\tNM:i:2\tMD:Z:10A20^CG0T18\tAS:i:44",
        f"b\t0\tB\t10\t60\t10S40M\t*\t0\t0\t{seq}\t{q}\tXZ:Z:{'z' * 300}\tMD:Z:40\tNM:i:7\tAS:i:-5",
        f"c\t16\tC\t10\t60\t20M5I25M\t*\t0\t0\t{seq}\t{q}\tAS:i:300\tMD:Z:A19^T0C24\tNM:i:1",
        f"d\t0\tA\t10\t60\t25=1X24=\t*\t0\t0\t{seq}\t{q}\tNM:i:1\tAS:i:70000",
        f"e\t0\tA\t10\t60\t5H45M5H\t*\t0\t0\t{'A' * 45}\t{'I' * 45}\tXH:H:1AE301\tNM:i:0\tAS:i:45",
        f"{'n' * 200}\t0\tA\t10\t60\t50M\t*\t0\t0\t{seq}\t{q}\tNM:i:0\tAS:i:50",
        f"{'n' * 200}\t256\tB\t10\t60\t30M20S\t*\t0\t0\t*\t*\tNM:i:1\tAS:i:50",
    ]
    s = _sam(HDR + "\n".join(recs) + "\n")
    exp = oracle.record_stats(s.raw, s.off)
    for slow in (False, True):
        with m.Context(l=1, p=50, stats=True, besthit=True, n_targets=3, force_slow=slow) as ctx:
            ctx.push(s.raw, s.off)
            st = ctx.pull_stats()
            kept = ctx.pull_kept()
        for k in ("alen", "qlen", "qclip", "edit", "score"):
            assert st[k].tolist() == exp[k].tolist(), (k, slow)
        assert kept.tolist() == oracle.filter_stream(s.raw, s.off, oracle.filter_cfg(l=1, p=50, besthit=True)).tolist()


def test_reopened_qname_and_unmapped_between(m, oracle):
    # A, B, A is three pools for the filter (msam_filter.c:120-125); an unmapped record with a different
    # name between two same-named mapped records splits the pool too, but the profile stage merges the
    # two A groups when only tid != -1 records are compared (msam_profile.c:223-232).
    a = "\t0\tA\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:%d"
    text = HDR + "x" + a % 10 + "\ny" + a % 5 + "\nx\t256\tB\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:20\n"
    text += "z" + a % 7 + "\nw\t4\t*\t0\t0\t*\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\nz\t256\tC\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:9\n"
    s = _sam(text)
    for opts in (dict(besthit=True), dict(uniqhit=True), dict(l=5, besthit=True)):
        exp = oracle.filter_stream(s.raw, s.off, oracle.filter_cfg(**opts))
        with m.Context(n_targets=3, **opts) as ctx:
            ctx.push(s.raw, s.off)
            assert ctx.pull_kept().tolist() == exp.tolist()
    for mode in ("all", "equal", "proportional", "ignore"):
        eab, est, eui, ed = oracle.profile(s.raw, s.off, None, 3, SHARE[mode])
        with m.Context(profile=True, multi=mode, n_targets=3) as ctx:
            ctx.push(s.raw, s.off)
            ui, d = ctx.pull_counts()
            ab, st = ctx.finish_profile()
        assert ui.tolist() == eui.tolist() and close(ab, eab) and close(d, ed)
        assert (st["mapped_inserts"], st["uniq"], st["multi"]) == (est["mapped_inserts"], est["uniq"], est["multi"])
        # besthit | profile on the same non-grouped input: pools x(A) and x(B) are separate for the filter but the
        # profile merges their winners into one group -> the fused pass must notice and fall back
        idx = oracle.filter_stream(s.raw, s.off, oracle.filter_cfg(besthit=True))
        eab, est, eui, ed = oracle.profile(s.raw, s.off, idx, 3, SHARE[mode])
        with m.Context(profile=True, multi=mode, kept=False, besthit=True, n_targets=3) as ctx:
            ctx.push(s.raw, s.off)
            ui, d = ctx.pull_counts()
            ab, st = ctx.finish_profile()
            t = ctx.timing()
        assert ui.tolist() == eui.tolist() and close(ab, eab) and close(d, ed)
        assert (st["mapped_inserts"], st["uniq"], st["multi"]) == (est["mapped_inserts"], est["uniq"], est["multi"])
        assert t["fused_chunks"] == 1


def test_fused_guard_fallback(m, oracle):
    # x(A) | y (dropped by -l) | x(B): two pools for the filter, ONE insert for the profile (msam_profile.c:226 compares
    # with the previous surviving record).  The fused pass must detect the equal QNAME hashes and hand the chunk to the
    # general pipeline; counts must match the oracle either way.
    a = "\t0\t%s\t10\t60\t%dM\t*\t0\t0\t%s\t%s\tNM:i:0\tAS:i:%d"
    rec = lambda n, ref, ln, sc: n + a % (ref, ln, "A" * ln, "I" * ln, sc)
    text = HDR + "\n".join([rec("u", "C", 20, 9), rec("x", "A", 20, 10), rec("y", "A", 5, 5), rec("x", "B", 20, 20), rec("z", "C", 20, 7)]) + "\n"
    s = _sam(text)
    opts = dict(l=8, besthit=True)
    idx = oracle.filter_stream(s.raw, s.off, oracle.filter_cfg(**opts))
    assert s.name_flags(idx) == "u:0,x:0,x:0,z:0"
    for mode in ("all", "equal", "proportional", "ignore"):
        eab, est, eui, ed = oracle.profile(s.raw, s.off, idx, 3, SHARE[mode])
        assert (est["mapped_inserts"], est["multi"]) == (3, 1)
        with m.Context(profile=True, multi=mode, kept=False, n_targets=3, **opts) as ctx:
            ctx.push(s.raw, s.off)
            ui, d = ctx.pull_counts()
            ab, st = ctx.finish_profile()
            t = ctx.timing()
        assert (t["fused_chunks"], t["fused_fallbacks"]) == (1, 1)
        assert ui.tolist() == eui.tolist() and close(ab, eab) and close(d, ed)
        assert (st["mapped_inserts"], st["uniq"], st["multi"], st["purged"]) == (est["mapped_inserts"], est["uniq"], est["multi"], est["purged"])


@pytest.mark.parametrize("mode", ["proportional", "equal"])
def test_zero_copy_push(m, oracle, mixed, mode):
    """msg_push on a pinned host buffer: the decode kernel reads its windows straight from host memory (no bulk H2D);
    same results as the staged copy and the oracle.  Unaligned sub-chunks silently take the staged path."""
    raw, off, tlen, p = mixed
    n = len(off) - 1
    opts = dict(l=80, p=95, z=80, besthit=True)
    idx = oracle.filter_stream(raw, off, oracle.filter_cfg(**opts))
    eab, est, eui, ed = oracle.profile(raw, off, idx, len(tlen), SHARE[mode])
    with m.PinnedBuffer(raw.nbytes) as pb, m.Context(profile=True, multi=mode, kept=False, n_targets=len(tlen), **opts) as ctx:
        pb.array[:] = raw
        ctx.push(pb.array, off)
        assert ctx.kept_count() == len(idx)
        ui, d = ctx.pull_counts()
        ab, st = ctx.finish_profile()
        t = ctx.timing(reset=True)
        assert (t["zero_copy_chunks"], t["fused_chunks"], t["fused_fallbacks"]) == (1, 1, 0)
        assert t["h2d_bytes"] < raw.nbytes                     # windows + offsets only
        assert np.array_equal(ui, eui) and close(d, ed) and close(ab, eab)
        for k in ("mapped_inserts", "uniq", "multi", "purged", "iterations", "converged", "n_lists", "n_entries"):
            assert st[k] == est[k], k
        # chunked: sub-chunks start wherever a QNAME group starts, so most are not 16-byte aligned
        ctx.reset()
        cuts = [0, m.split_point(raw, off, n // 3), m.split_point(raw, off, 2 * n // 3), n]
        for a, b in zip(cuts[:-1], cuts[1:]):
            ctx.push(pb.array[int(off[a]):int(off[b])], off[a:b + 1] - off[a])
        ab2, st2 = ctx.finish_profile()
        ui2, _ = ctx.pull_counts()
        assert np.array_equal(ui2, eui) and close(ab2, eab) and st2["n_lists"] == est["n_lists"]
        # chunked, every chunk in its own pinned buffer (what a double-buffering host does): all of them zero-copy
        ctx.reset(); ctx.timing(reset=True)
        for a, b in zip(cuts[:-1], cuts[1:]):
            lo, hi = int(off[a]), int(off[b])
            with m.PinnedBuffer(hi - lo) as pc:
                pc.array[:] = raw[lo:hi]
                ctx.push(pc.array, off[a:b + 1] - off[a])
        ab3, st3 = ctx.finish_profile()
        ui3, _ = ctx.pull_counts()
        t3 = ctx.timing()
        assert (t3["zero_copy_chunks"], t3["fused_chunks"], t3["fused_fallbacks"]) == (3, 3, 0)
        assert np.array_equal(ui3, eui) and close(ab3, eab) and st3["n_lists"] == est["n_lists"] and st3["purged"] == est["purged"]


def test_zero_copy_guard_fallback(m, oracle):
    """zero-copy chunk that the fused pass declines (equal QNAME hashes around a dropped record): it is staged into
    device memory and rerun on the general pipeline."""
    a = "\t0\t%s\t10\t60\t%dM\t*\t0\t0\t%s\t%s\tNM:i:0\tAS:i:%d"
    rec = lambda n, ref, ln, sc: n + a % (ref, ln, "A" * ln, "I" * ln, sc)
    text = HDR + "\n".join([rec("u", "C", 20, 9), rec("x", "A", 20, 10), rec("y", "A", 5, 5), rec("x", "B", 20, 20), rec("z", "C", 20, 7)]) + "\n"
    s = _sam(text)
    opts = dict(l=8, besthit=True)
    idx = oracle.filter_stream(s.raw, s.off, oracle.filter_cfg(**opts))
    eab, est, eui, ed = oracle.profile(s.raw, s.off, idx, 3, SHARE["proportional"])
    with m.PinnedBuffer(len(s.raw)) as pb, m.Context(profile=True, multi="proportional", kept=False, n_targets=3, **opts) as ctx:
        pb.array[:] = s.raw
        ctx.push(pb.array, s.off)
        ui, d = ctx.pull_counts()
        ab, st = ctx.finish_profile()
        t = ctx.timing()
    assert (t["zero_copy_chunks"], t["fused_chunks"], t["fused_fallbacks"]) == (1, 1, 1)
    assert ui.tolist() == eui.tolist() and close(ab, eab)
    assert (st["mapped_inserts"], st["uniq"], st["multi"], st["purged"]) == (est["mapped_inserts"], est["uniq"], est["multi"], est["purged"])


def test_huge_group(m, oracle):
    # one QNAME with 5000 alignments over 40 references: exercises the oversized-group serial kernel
    rng = np.random.default_rng(7)
    hdr = "@HD\tVN:1.6\tSO:queryname\n" + "".join(f"@SQ\tSN:r{i}\tLN:1000\n" for i in range(40))
    lines = []
    for k in range(5000):
        lines.append(f"big\t{256 if k else 0}\tr{int(rng.integers(40))}\t{1 + int(rng.integers(900))}\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:{int(rng.integers(3))}")
    lines.append("small\t0\tr3\t5\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:1")
    s = _sam(hdr + "\n".join(lines) + "\n")
    exp = oracle.filter_stream(s.raw, s.off, oracle.filter_cfg(besthit=True))
    with m.Context(besthit=True, n_targets=40) as ctx:
        ctx.push(s.raw, s.off)
        assert ctx.pull_kept().tolist() == exp.tolist()
    for mode in ("all", "equal", "proportional", "ignore"):
        eab, est, eui, ed = oracle.profile(s.raw, s.off, None, 40, SHARE[mode])
        with m.Context(profile=True, multi=mode, n_targets=40) as ctx:
            ctx.push(s.raw, s.off)
            ui, d = ctx.pull_counts()
            ab, st = ctx.finish_profile()
        assert ui.tolist() == eui.tolist() and close(d, ed) and close(ab, eab), mode
        assert (st["mapped_inserts"], st["uniq"], st["multi"], st["purged"]) == (est["mapped_inserts"], est["uniq"], est["multi"], est["purged"])
        # a 5000-record pool is beyond the fused pass's walk limit -> general pipeline, same numbers
        eab, est, eui, ed = oracle.profile(s.raw, s.off, exp, 40, SHARE[mode])
        with m.Context(profile=True, multi=mode, kept=False, besthit=True, n_targets=40) as ctx:
            ctx.push(s.raw, s.off)
            ui, d = ctx.pull_counts()
            ab, st = ctx.finish_profile()
            assert ctx.timing()["fused_fallbacks"] == 1
        assert ui.tolist() == eui.tolist() and close(d, ed) and close(ab, eab), mode
        assert (st["mapped_inserts"], st["uniq"], st["multi"], st["purged"]) == (est["mapped_inserts"], est["uniq"], est["multi"], est["purged"])


# ---------------------------------------------------------------- BASELINE sizes: size-independent properties
def test_full_size_properties(m, oracle):
    """configs[1] size (10 M PE150 alignments): idempotence of the filter, conservation laws of the
    profile, and a bit-exact oracle comparison on a 1 M-record prefix."""
    from msamtools_b200 import synth
    p = synth.make_params("community", n_records=10_000_000, seed=13579)
    raw, off, st = synth.generate(p)
    tlen = synth.target_lengths(p)
    n = len(off) - 1
    opts = dict(l=80, p=95, z=80, besthit=True)
    with m.Context(profile=True, multi="proportional", records=True, n_targets=len(tlen), **opts) as ctx:
        ctx.push(raw, off)
        kept = ctx.pull_kept()
        rec, nrec = ctx.pull_records()
        ab, pst = ctx.finish_profile()
        ui, _ = ctx.pull_counts()
    assert 0 < len(kept) < n and nrec == len(kept)
    # kept indices are a permutation-free selection; within a QNAME group READ1 precedes READ2
    assert len(np.unique(kept)) == len(kept)
    # conservation: every insert is unique or multi; unique counts sum to 2*uniq (proportional mode)
    assert pst["uniq"] + pst["multi"] == pst["mapped_inserts"]
    assert int(ui.astype(np.int64).sum()) == 2 * pst["uniq"]
    assert abs(ab.sum() - (pst["mapped_inserts"] - pst["purged"])) <= 1e-6 * pst["mapped_inserts"]
    # idempotence: filtering the filtered stream again changes nothing (records are byte-identical)
    off2 = m.index_records(rec)
    with m.Context(records=True, n_targets=len(tlen), **opts) as ctx:
        ctx.push(rec, off2)
        assert ctx.kept_count() == nrec
        rec2, _ = ctx.pull_records()
    assert np.array_equal(rec, rec2)
    # bit-exact against the oracle on a 1 M prefix cut at a QNAME boundary
    k = m.split_point(raw, off, 1_000_000)
    sub_raw, sub_off = raw[:int(off[k])], off[:k + 1]
    exp = oracle.filter_stream(sub_raw, sub_off, oracle.filter_cfg(**opts))
    assert np.array_equal(kept[kept < k], exp)
    eab, est, eui, _ = oracle.profile(sub_raw, sub_off, exp, len(tlen), 3)
    with m.Context(profile=True, multi="proportional", n_targets=len(tlen), **opts) as ctx:
        ctx.push(sub_raw, sub_off)
        ab1, st1 = ctx.finish_profile()
        ui1, _ = ctx.pull_counts()
    assert np.array_equal(ui1, eui) and close(ab1, eab)
    assert (st1["iterations"], st1["purged"], st1["multi"]) == (est["iterations"], est["purged"], est["multi"])


def test_cg_tag_cigar_is_refused(m):
    """htslib's placeholder for CIGARs of > 65535 operations ("<l_seq>S<ref_len>N" + CG:B,I tag, SAM spec 4.2.2) is not
    expanded on the device: the record is refused (MSG_EFORMAT) instead of being silently mis-counted."""
    text = HDR + "r1\t0\tA\t10\t60\t10S100N\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:5\tCG:B:I,160,16\n"
    s = _sam(text)
    for kw in (dict(l=5), dict(coverage=True, target_len=s.target_len)):
        with m.Context(n_targets=3, **kw) as ctx:
            with pytest.raises(m.MsgError) as e:
                ctx.push(s.raw, s.off)
            assert e.value.code == m._lib.MSG_EFORMAT and "CG tag" in e.value.text
    with m.Context(besthit=True, n_targets=3) as ctx:            # --besthit alone never looks at the CIGAR (msam_filter.c:104,145)
        ctx.push(s.raw, s.off)
        assert ctx.pull_kept().tolist() == [0]
