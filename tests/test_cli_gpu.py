"""GPU: the drop-in `msamtools` CLI (C host + libmsamtools_b200) end to end on files, against the
golden vectors.  Expected QNAME:FLAG strings and profile values are the ones the reference's shell
suites assert (restated in tests/test_oracle_golden.py); outputs are also compared with the oracle.
"""
import gzip
import os
import subprocess

import numpy as np
import pytest

import goldenutil as G
import samutil
import test_oracle_golden as T
from conftest import ROOT

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "msamtools_b200", "bin", "msamtools")


def write_fixture(tmp_path, name, fmt="bam"):
    s = G.fixture(name)
    hdr = "\n".join(s.header_lines + ["@SQ\tSN:%s\tLN:%d" % (n, int(l)) for n, l in zip(s.ref_names, s.target_len)]) + "\n"
    # keep @HD first
    lines = [l for l in hdr.splitlines() if l.startswith("@HD")] + [l for l in hdr.splitlines() if not l.startswith("@HD")]
    hdr = "\n".join(lines) + "\n"
    path = str(tmp_path / f"{name}.bam")
    samutil.write_bam(path, hdr, s.ref_names, s.target_len, s.raw)
    return path, s


def run(args, stdin=None, check=True):
    r = subprocess.run([CLI] + args, input=stdin, capture_output=True)
    if check:
        assert r.returncode == 0, r.stderr.decode()[-2000:]
    return r


def sam_name_flags(text):
    return ",".join("%s:%s" % tuple(l.split("\t")[:2]) for l in text.splitlines() if l and not l.startswith("@"))


def opts_to_argv(o):
    argv = []
    for k, v in o.items():
        if k in ("l", "p", "z"):
            argv += [f"-{k}", str(v)]
        elif k == "ppt":
            argv += ["--ppt", str(v)]
        elif k == "invert" and v:
            argv.append("-v")
        elif k == "keep_unmapped" and v:
            argv.append("-k")
        elif v:
            argv.append(f"--{k}")
    return argv


@pytest.mark.parametrize("opts,expected", T.FILTER_CASES, ids=lambda x: "" if isinstance(x, str) else ",".join(f"{k}={v}" for k, v in x.items()))
def test_cli_filter_sam_output(tmp_path, opts, expected):
    path, s = write_fixture(tmp_path, "filter")
    r = run(["filter", "-h"] + opts_to_argv(opts) + [path])
    out = r.stdout.decode()
    assert sam_name_flags(out) == expected
    assert "@PG\tID:msamtools" in out and "QNAME grouping check: not required for this operation" in out


@pytest.mark.parametrize("fixture,opts,expected", T.BESTHIT_CASES, ids=lambda x: "" if isinstance(x, str) and ":" in x else str(x))
def test_cli_besthit(tmp_path, fixture, opts, expected):
    path, s = write_fixture(tmp_path, fixture.replace(".sam", ""))
    out = run(["filter", "-S", "-h"] + opts_to_argv(opts) + [path]).stdout.decode()
    assert sam_name_flags(out) == expected
    assert "QNAME grouping check: confirmed by input header SO:queryname" in out
    if opts.get("rescore"):
        assert "AS:i:100" in [l for l in out.splitlines() if l.startswith("rescore\t256")][0]


def test_cli_filter_bam_roundtrip_and_pipe(tmp_path, oracle):
    # BASELINE configs[0]: filter -b -u -l 80 -p 95 -z 80 --besthit tiny_aln.bam | profile --multi=proportional
    path, s = write_fixture(tmp_path, "tiny_aln")
    f = run(["filter", "-b", "-u", "-l", "80", "-p", "95", "-z", "80", "--besthit", path])
    cfg = oracle.filter_cfg(l=80, p=95, z=80, besthit=True)
    idx = oracle.filter_stream(s.raw, s.off, cfg)
    import struct
    data = gzip.decompress(f.stdout)
    l_text, = struct.unpack_from("<i", data, 4)
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, o); o += 4
    for _ in range(n_ref):
        ln, = struct.unpack_from("<i", data, o); o += 4 + ln + 4
    assert data[o:] == bytes(oracle.emit_records(s.raw, s.off, idx, cfg))          # record stream byte-identical
    outp = str(tmp_path / "prof.gz")
    p = run(["profile", "--multi=proportional", "--label", "X", "-o", outp, "-"], stdin=f.stdout)
    comments, body = samutil.read_profile_gz(outp)
    text = "\n".join(comments)
    assert "Mapped inserts      :       7" in text and "- Multiple mapped :       3" in text and "- Uniquely mapped :       4" in text
    assert "PropSharing Iteration:  1; DELTA^2=0. CONVERGED!" in p.stderr.decode()
    assert "# Purged 3 inserts" in p.stderr.decode()
    assert body[0] == ["ID", "X"] and body[1][0] == "Unknown"
    nz = {k: v for k, v in body[2:] if float(v) != 0}
    assert set(nz) == {"MH0349_GL0038880", "MH0013_GL0018062", "479436.Vpar_1233", "MH0002_GL0008419"}


@pytest.mark.parametrize("mode,unknown,a,b", T.PROFILE_CASES)
def test_cli_profile_fixture(tmp_path, mode, unknown, a, b):
    # tests/test_profile.sh:16-67
    path, s = write_fixture(tmp_path, "profile")
    outp = str(tmp_path / f"{mode}.gz")
    run(["profile", "-S", "--label", "test", "--unit", "ab", "--nolen", "--total", "7", "--multi", mode, "--pandas", "-o", outp, path])
    comments, body = samutil.read_profile_gz(outp)
    text = "\n".join(comments)
    for needle in ("QNAME grouping check: confirmed by input header SO:queryname", "Total inserts       : 7", "Mapped inserts      : 7",
                   "- Multiple mapped : 1", "- Uniquely mapped : 6"):
        assert needle in text, needle
    vals = {k: float(v) for k, v in body[1:]}
    assert abs(vals["Unknown"] - unknown) <= 1e-9 and abs(vals["A"] - a) <= 1e-6 and abs(vals["B"] - b) <= 1e-6
    assert body[0] == ["ID", "test"]


def test_cli_profile_mincount_and_units(tmp_path):
    # tests/test_profile.sh:139-246
    path, s = write_fixture(tmp_path, "profile_fractional_mincount")
    outp = str(tmp_path / "m.gz")
    run(["profile", "-S", "--label", "f", "--unit", "ab", "--nolen", "--total", "3", "--multi", "equal", "--mincount", "1", "--pandas", "-o", outp, path])
    vals = {k: float(v) for k, v in samutil.read_profile_gz(outp)[1][1:]}
    assert abs(vals["Unknown"] - 1 / 3) < 1e-6 and abs(vals["A"] - 4 / 3) < 1e-6 and abs(vals["B"] - 4 / 3) < 1e-6 and vals["C"] == 0
    path, s = write_fixture(tmp_path, "profile")
    run(["profile", "-S", "--label", "legacy", "--unit", "ab", "--nolen", "--total", "7", "--multi", "equal", "--no-pandas", "-o", outp, path])
    assert samutil.read_profile_gz(outp)[1][0] == ["legacy"]
    # relative abundance (default unit) with length normalisation sums to one over Unknown + features
    run(["profile", "-S", "--label", "r", "--total", "10", "-o", outp, path])
    comments, body = samutil.read_profile_gz(outp)
    assert "# Estimated seq. length for 'Unknown': 1000bp" in comments
    assert abs(sum(float(v) for _, v in body[1:]) - 1.0) < 1e-7
    for fx in ("profile_empty", "profile_unmapped"):                       # tests/test_profile.sh:99-137
        path, s = write_fixture(tmp_path, fx)
        run(["profile", "-S", "--label", fx, "--unit", "ab", "--nolen", "--multi", "equal", "--pandas", "-o", outp, path])
        comments, body = samutil.read_profile_gz(outp)
        text = "\n".join(comments)
        assert "Mapped inserts      :       0" in text and "Effective inserts   :          0" in text
        assert all(float(v) == 0 for _, v in body[1:])


def test_cli_profile_genome_order(tmp_path, oracle):
    # --genome: rows follow the reference's hash key order (msam_profile.c:791-798); 40 genomes crosses two table growths
    from msamtools_b200 import synth
    p = synth.make_params("mixed", n_records=20_000, seed=5)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    names = [f"seq{i:03d}" for i in range(len(tlen))]
    path = str(tmp_path / "g.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, raw)
    gdef = str(tmp_path / "genomes.tsv")
    genome_of = [f"genome_{(i * 7) % 40:02d}" for i in range(len(tlen))]
    with open(gdef, "w") as fh:
        for g, n in zip(genome_of, names):
            fh.write(f"{g}\t{n}\n")
    outp = str(tmp_path / "g.gz")
    run(["profile", "--label", "g", "--unit", "ab", "--nolen", "--multi", "prop", "--genome", gdef, "-o", outp, path])
    comments, body = samutil.read_profile_gz(outp)
    order = [k for k, _ in body[2:]]
    assert sorted(order) == sorted(set(genome_of)) and len(order) == 40
    expected_order = os.path.join(ROOT, "tests", "golden", "genome_order_40.txt")
    with open(expected_order) as fh:                                         # produced by oracle/_ref (make_golden_cli.py)
        assert order == fh.read().split()
    fmap = np.array([order.index(g) for g in genome_of], dtype=np.int32)
    ab, st, _, _ = oracle.profile(raw, off, None, len(tlen), 3, fmap=fmap, n_features=40)
    for k, v in body[2:]:
        assert v == "%.8g" % ab[order.index(k)]


def test_cli_coverage(tmp_path):
    # tests/test_coverage.sh:28-81
    path, s = write_fixture(tmp_path, "coverage")
    outp = str(tmp_path / "c.gz")
    r = run(["coverage", "-S", "-o", outp, "-w", "4", path])
    assert r.stdout == b"" and r.stderr == b""
    with gzip.open(outp, "rt") as fh:
        assert fh.read() == ">A\n1 0 1 1\n2 2 1 0\n0 1\n>B\n0 0 0 0\n0\n>C\n0 0 0 0\n0 0 0 0\n0 1\n>D\n4 4 2 4\n3 0 0 0\n"
    run(["coverage", "-S", "-o", outp, "--summary", path])
    with gzip.open(outp, "rt") as fh:
        assert fh.read() == "A\t0.70000000\t0.90\nB\t0\t0\nC\t0.10000000\t0.10\nD\t0.62500000\t2.12\n"
    run(["coverage", "-S", "-o", outp, "--summary", "--skipuncovered", path])
    with gzip.open(outp, "rt") as fh:
        assert fh.read() == "A\t0.70000000\t0.90\nC\t0.10000000\t0.10\nD\t0.62500000\t2.12\n"


def test_cli_qname_order(tmp_path):
    # tests/test_qname_order.sh:13-63
    outp = str(tmp_path / "q.gz")
    path, s = write_fixture(tmp_path, "qname_coordinate")
    r = run(["profile", "-S", "--label", "t", "--unit", "ab", "--nolen", "--multi", "equal", "-o", outp, path], check=False)
    assert r.returncode == 1 and b"SO:coordinate" in r.stderr and b"samtools sort -n" in r.stderr and not os.path.exists(outp)
    path, s = write_fixture(tmp_path, "qname_reopened")
    r = run(["profile", "-S", "--label", "t", "--unit", "ab", "--nolen", "--multi", "equal", "-o", outp, path], check=False)
    assert r.returncode == 1 and b"not grouped by QNAME" in r.stderr and b"readA" in r.stderr and not os.path.exists(outp)


def test_cli_streaming_boundary(tmp_path):
    # tests/test_streaming.sh: a QNAME group straddling record 100000 must still give one best hit
    seq, q = "A" * 100, "I" * 100
    lines = ["@HD\tVN:1.6", "@SQ\tSN:A\tLN:1000", "@SQ\tSN:B\tLN:1000"]
    for i in range(1, 100000):
        lines.append(f"q{i:06d}\t0\tA\t{100 if i % 2 == 0 else 200}\t60\t100M\t*\t0\t0\t{seq}\t{q}\tAS:i:50\tNM:i:0")
    lines.append(f"boundary\t0\tA\t300\t60\t100M\t*\t0\t0\t{seq}\t{q}\tAS:i:10\tNM:i:0")
    lines.append(f"boundary\t256\tB\t400\t60\t100M\t*\t0\t0\t{seq}\t{q}\tAS:i:20\tNM:i:0")
    r = run(["filter", "-S", "-h", "--besthit", "-"], stdin=("\n".join(lines) + "\n").encode())
    out = r.stdout.decode()
    assert "QNAME grouping check: no QNAME grouping violation detected in first 10000 records" in out
    recs = [l for l in out.splitlines() if not l.startswith("@")]
    assert len(recs) == 100000
    assert [l.split("\t")[:4] for l in recs if l.startswith("boundary")] == [["boundary", "256", "B", "400"]]


def test_cli_data_errors(tmp_path):
    sam = "@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:A\tLN:1000\nr1\t0\tA\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tAS:i:5\n"
    r = run(["filter", "-S", "-l", "5", "-"], stdin=sam.encode(), check=False)
    assert r.returncode == 1 and b"Fatal Error: Either NM or MD must be present in SAM/BAM input for 'filter' command" in r.stderr
    sam = sam.replace("AS:i:5", "NM:i:0")
    r = run(["filter", "-S", "--besthit", "-"], stdin=sam.encode(), check=False)
    assert r.returncode == 1 and b"Fatal Error: Required field AS not found in SAM/BAM input" in r.stderr


def _synth_bam(tmp_path, n_records, preset="mixed", seed=99, level=1):
    from msamtools_b200 import synth
    p = synth.make_params(preset, n_records=n_records, seed=seed)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    names = [f"seq{i:03d}" for i in range(len(tlen))]
    path = str(tmp_path / "synth.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, raw, level=level)
    return path, raw, off, tlen, names


@pytest.mark.parametrize("pinned", ["0", "1"])
def test_cli_chunked_reader_thread(tmp_path, oracle, pinned):
    """many small chunks through the reader thread / buffer ring (bulk BGZF ingest, QNAME-boundary cuts, tails carried
    into the next buffer): filter output bytes, profile table and coverage summary equal the oracle's on the whole stream"""
    path, raw, off, tlen, names = _synth_bam(tmp_path, 120_000)
    env = dict(os.environ, MSAMTOOLS_CHUNK_RECORDS="7000", MSAMTOOLS_PINNED=pinned, MSAMTOOLS_THREADS="4", MSAMTOOLS_TIMING="1")
    cfg = oracle.filter_cfg(l=80, p=95, z=80, besthit=True)
    idx = oracle.filter_stream(raw, off, cfg)
    f = subprocess.run([CLI, "filter", "-b", "-u", "-l", "80", "-p", "95", "-z", "80", "--besthit", path], capture_output=True, env=env)
    assert f.returncode == 0, f.stderr.decode()[-2000:]
    assert b"# timing:" in f.stderr
    tmp = str(tmp_path / "f.bam")
    with open(tmp, "wb") as fh:
        fh.write(f.stdout)
    got = samutil.read_bam(tmp)
    assert bytes(got.raw) == bytes(oracle.emit_records(raw, off, idx, cfg))
    # same stream piped on stdin into `profile` (record-wise pre-flight, then bulk ingest from the pipe)
    outp = str(tmp_path / "p.gz")
    p = subprocess.run([CLI, "profile", "--label", "x", "--unit", "ab", "--nolen", "--multi", "prop", "-o", outp, "-"], input=f.stdout,
                       capture_output=True, env=env)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    ab, st, _, _ = oracle.profile(raw, off, idx, len(tlen), 3)
    comments, body = samutil.read_profile_gz(outp)
    assert [v for _, v in body[2:]] == ["%.8g" % x for x in ab]
    assert f"PropSharing Iteration: {st['iterations']:2d}" in p.stderr.decode()
    # coverage summary over the unfiltered file
    outc = str(tmp_path / "c.gz")
    c = subprocess.run([CLI, "coverage", "--summary", "-o", outc, path], capture_output=True, env=env)
    assert c.returncode == 0 and c.stdout == b"", c.stderr.decode()[-2000:]
    cov, touched, total, _ = oracle.coverage(raw, off, None, tlen)
    with gzip.open(outc, "rt") as fh:
        lines = fh.read().splitlines()
    exp = [f"{n}\t0\t0" if not cv else "%s\t%.8f\t%.2f" % (n, t / l, s / l) for n, cv, t, s, l in zip(names, cov, touched, total, tlen)]
    assert lines == exp


def test_cli_corrupt_input_is_fatal(tmp_path):
    """a flipped payload byte (CRC mismatch) or a truncated file must not end the stream silently with exit status 0"""
    path, raw, off, tlen, names = _synth_bam(tmp_path, 30_000, level=6)
    data = bytearray(open(path, "rb").read())
    env = dict(os.environ, MSAMTOOLS_CHUNK_RECORDS="5000")
    outp = str(tmp_path / "p.gz")
    for threads in ("1", "4"):
        bad = bytearray(data)
        bad[len(bad) // 2] ^= 0x5a
        p1 = str(tmp_path / f"bad{threads}.bam")
        open(p1, "wb").write(bad)
        r = subprocess.run([CLI, "profile", "--label", "x", "-o", outp, p1], capture_output=True, env=dict(env, MSAMTOOLS_THREADS=threads))
        assert r.returncode == 1 and (b"Fatal Error: Cannot read input" in r.stderr or b"Fatal Error: Cannot read header" in r.stderr), (threads, r.stderr[-500:])
        assert not os.path.exists(outp)
        p2 = str(tmp_path / f"trunc{threads}.bam")
        open(p2, "wb").write(data[:len(data) * 2 // 3])
        r = subprocess.run([CLI, "filter", "-b", "-l", "50", p2], capture_output=True, env=dict(env, MSAMTOOLS_THREADS=threads))
        assert r.returncode == 1 and (b"Fatal Error: Cannot read input" in r.stderr or b"Fatal Error: Cannot read header" in r.stderr), (threads, r.stderr[-500:])
    # malformed SAM text
    sam = "@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:A\tLN:1000\nr1\t0\tNOPE\t10\t60\t10M\t*\t0\t0\tAAAAAAAAAA\tIIIIIIIIII\tNM:i:0\tAS:i:5\n"
    r = run(["filter", "-S", "-l", "5", "-"], stdin=sam.encode(), check=False)
    assert r.returncode == 1 and b"Fatal Error: Cannot read input" in r.stderr
