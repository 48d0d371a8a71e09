"""oracle/_ref = the reference's OWN object code (its unmodified C sources compiled against the
shims in oracle/shim/).  These tests run only where /root/reference exists (build container):

1. the reference's ten `make check` suites pass against it  -> the shim is faithful;
2. the CPU oracle (our restatement) agrees with it on seeded synthetic streams -> the oracle is
   pinned against the reference itself, not only against its test fixtures.
"""
import gzip
import os
import subprocess

import numpy as np
import pytest

import samutil
from conftest import REF_TESTS, ROOT, needs_reference

pytestmark = needs_reference
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "msamtools")


@pytest.fixture(scope="module")
def ref_bin():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-f", "Makefile.ref"])
    return REF_BIN


@pytest.mark.parametrize("suite", ["smoke", "filter", "besthit", "profile", "coverage", "integration", "qname_order",
                                   "errors", "streaming", "summary"])
def test_reference_suite_passes_on_ref_binary(ref_bin, suite, tmp_path):
    env = dict(os.environ, MSAMTOOLS=ref_bin, TMPDIR=str(tmp_path))
    r = subprocess.run(["sh", os.path.join(REF_TESTS, f"test_{suite}.sh")], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


STREAMS = {"mixed": ("mixed", 40_000, 424242, {}),
           # second stream: few references, heavy clipping / indels, unmapped pairs, up to 12 occurrences per insert
           "clippy": ("community", 30_000, 7, dict(n_refs=12, ref_len_min=3_000, ref_len_max=9_000, clip_fraction=0.35, indel_fraction=0.25,
                                                  unmapped_fraction=0.06, shared_fraction=0.45, single_fraction=0.15, max_occ=12)),
           # third stream: a gene catalogue (10 000 short references, log-normal abundances, 30 % multi-mappers): the PropSharing loop
           # of the restatement against the reference's own over thousands of features and lists
           "catalog": ("catalog10k", 60_000, 2026, {})}


@pytest.fixture(scope="module", params=list(STREAMS))
def synth_bam(request, tmp_path_factory):
    from msamtools_b200 import synth
    preset, nrec, seed, over = STREAMS[request.param]
    p = synth.make_params(preset, n_records=nrec, seed=seed, **over)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    names = [f"ref{i:04d}" for i in range(len(tlen))]
    path = str(tmp_path_factory.mktemp("synth") / "in.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, raw)
    return path, raw, off, tlen, names


FILTERS = [(["-l", "80", "-p", "95", "-z", "80"], dict(l=80, p=95, z=80)),
           (["-l", "80", "-p", "95", "-z", "80", "--besthit"], dict(l=80, p=95, z=80, besthit=True)),
           (["--uniqhit"], dict(uniqhit=True)),
           (["--ppt", "-990"], dict(ppt=-990)),
           (["-p", "99", "-v", "-k"], dict(p=99, invert=True, keep_unmapped=True)),
           (["-l", "100", "--rescore", "--besthit"], dict(l=100, rescore=True, besthit=True))]


def _payload(bam_bytes):
    """record stream of an (uncompressed-BGZF) BAM"""
    import struct
    data = gzip.decompress(bam_bytes)
    l_text, = struct.unpack_from("<i", data, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, o); o += 4
    for _ in range(n_ref):
        ln, = struct.unpack_from("<i", data, o); o += 4 + ln + 4
    return data[o:]


@pytest.mark.parametrize("argv,opts", FILTERS, ids=lambda x: " ".join(x) if isinstance(x, list) else "")
def test_oracle_filter_equals_reference(ref_bin, oracle, synth_bam, argv, opts):
    path, raw, off, tlen, names = synth_bam
    out = subprocess.run([ref_bin, "filter", "-b", "-u"] + argv + [path], capture_output=True, check=True).stdout
    cfg = oracle.filter_cfg(**opts)
    idx = oracle.filter_stream(raw, off, cfg)
    assert _payload(out) == bytes(oracle.emit_records(raw, off, idx, cfg))


@pytest.mark.parametrize("mode", ["all", "equal", "proportional", "ignore"])
@pytest.mark.parametrize("pre", [None, 1], ids=["plain", "filter|profile"])
def test_oracle_profile_equals_reference(ref_bin, oracle, synth_bam, mode, pre, tmp_path):
    path, raw, off, tlen, names = synth_bam
    outp = str(tmp_path / "p.gz")
    prof = [ref_bin, "profile", "--label", "x", "--unit", "ab", "--nolen", "--multi", mode, "-o", outp]
    if pre:
        f = subprocess.Popen([ref_bin, "filter", "-b", "-u", "-l", "80", "-p", "95", "-z", "80", "--besthit", path], stdout=subprocess.PIPE)
        r = subprocess.run(prof + ["-"], stdin=f.stdout, capture_output=True, text=True)
        f.wait()
        idx = oracle.filter_stream(raw, off, oracle.filter_cfg(l=80, p=95, z=80, besthit=True))
    else:
        r = subprocess.run(prof + [path], capture_output=True, text=True)
        idx = None
    assert r.returncode == 0, r.stderr
    share = {"all": 1, "equal": 2, "proportional": 3, "ignore": 4}[mode]
    ab, st, _, _ = oracle.profile(raw, off, idx, len(tlen), share)
    comments, body = samutil.read_profile_gz(outp)
    text = "\n".join(comments)
    for label, val in (("Mapped inserts", st["mapped_inserts"]), ("- Multiple mapped ", st["multi"]), ("- Uniquely mapped ", st["uniq"])):
        line = next(l for l in comments if label in l)
        assert int(line.split(":")[1].split("(")[0]) == val, line
    got = {k: v for k, v in body[1:]}
    for i, n in enumerate(names):
        assert got[n] == "%.8g" % ab[i], (n, got[n], ab[i])
    if mode == "proportional":
        iters = [l for l in r.stderr.splitlines() if "PropSharing Iteration" in l]
        assert len(iters) == st["iterations"]
        assert f"Purged {st['purged']} inserts" in r.stderr


def test_oracle_coverage_equals_reference(ref_bin, oracle, synth_bam, tmp_path):
    path, raw, off, tlen, names = synth_bam
    outp = str(tmp_path / "c.gz")
    subprocess.run([ref_bin, "coverage", "--summary", "-o", outp, path], check=True)
    cov, touched, total, _ = oracle.coverage(raw, off, None, tlen)
    with gzip.open(outp, "rt") as fh:
        lines = fh.read().splitlines()
    exp = ["%s\t0\t0" % n if not cov[t] else "%s\t%.8f\t%.2f" % (n, touched[t] / int(tlen[t]), total[t] / int(tlen[t]))
           for t, n in enumerate(names)]
    assert lines == exp
