"""CPU: the HOST plumbing of the drop-in CLI -- reader thread and buffer ring, read-ahead, multi-threaded record index, QNAME
cuts, the record-output writer thread, BGZF packing, the parallel gzip tables -- run against a NULL DEVICE
(tests/hostprof/nulldev.c: keeps the records whose POS is not a multiple of 5, returns fixed patterns for profile and
coverage).  Nothing here says anything about alignment arithmetic (that is the GPU parity suite); it checks that every byte
that goes in comes out where it should, for many chunk sizes and thread counts, and that the threads are race-free
(ThreadSanitizer build)."""
import gzip
import os
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest

import samutil
from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "hostprof"))
CS = os.path.join(ROOT, "msamtools_b200", "csrc")
HOSTSRC = [os.path.join(CS, "cli", "msamtools_main.c")] + [os.path.join(CS, "host", f) for f in
                                                          ("bamio.c", "finflate.c", "crc32x.c", "gzpar.c", "margs.c", "keyorder.c", "recwalk.c")]


def build(d, san):
    lib = os.path.join(d, "libmsamtools_b200.so")
    flags = ["-O1", "-g", "-std=gnu99"] + ([f"-fsanitize={san}", "-fno-sanitize-recover=all"] if san else [])
    subprocess.run(["gcc"] + flags + ["-fPIC", "-shared", "-o", lib, os.path.join(ROOT, "tests", "hostprof", "nulldev.c")], check=True)
    cli = os.path.join(d, "msamtools")
    subprocess.run(["gcc"] + flags + ["-o", cli] + HOSTSRC + ["-L" + d, "-lmsamtools_b200", "-Wl,-rpath," + d, "-lz", "-lm", "-lpthread"], check=True)
    return cli


@pytest.fixture(scope="module", params=["address,undefined", "thread"])
def cli(tmp_path_factory, request):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    d = str(tmp_path_factory.mktemp("nulldev_" + request.param.split(",")[0]))
    os.environ["ASAN_OPTIONS"] = "detect_leaks=0"            # a command-line tool: what it holds at exit is the OS's to free
    os.environ["TSAN_OPTIONS"] = "report_thread_leaks=0"     # ... including the threads still running when a fatal error exits
    return build(d, request.param)


@pytest.fixture(scope="module")
def bam(tmp_path_factory):
    from msamtools_b200 import synth
    d = tmp_path_factory.mktemp("nulldev_in")
    p = synth.make_params("mixed", n_records=115_000, seed=11)            # ~ 34 MB of records
    raw, off, _ = synth.generate(p)
    n = len(off) - 1
    raw = raw[:int(off[n])]
    tlen = synth.target_lengths(p)
    names = [f"ref{i:04d}" for i in range(len(tlen))]
    path = str(d / "in.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, raw, level=1)
    pos = np.array([struct.unpack_from("<I", raw, int(o) + 8)[0] for o in off[:n]])
    kept = b"".join(bytes(raw[int(off[i]):int(off[i + 1])]) for i in range(n) if pos[i] == 0xffffffff or pos[i] % 5 != 0)
    return path, bytes(raw), kept, names, tlen, n


ENVS = [dict(MSAMTOOLS_CHUNK_RECORDS="7000", MSAMTOOLS_THREADS="4"),          # many chunks: ring, tails, writer hand-over
        dict(MSAMTOOLS_CHUNK_RECORDS="9000", MSAMTOOLS_THREADS="3", MSAMTOOLS_PINNED="1"),   # buffers from msg_host_alloc (input ring and output)
        dict(MSAMTOOLS_CHUNK_MB="1", MSAMTOOLS_CHUNK_SLACK_KB="192", MSAMTOOLS_THREADS="4"),  # 1.2 MB buffers: ~30 bulk chunks, batches cut to fit,
                                                                                              # tails carried from buffer to buffer
        dict(MSAMTOOLS_CHUNK_RECORDS="50000", MSAMTOOLS_THREADS="1"),        # streaming inflate, record index without threads
        dict(MSAMTOOLS_THREADS="8")]                                          # default chunk size: one chunk


@pytest.mark.parametrize("env", ENVS, ids=lambda e: "-".join(e.values()))
def test_filter_record_output_and_pipe_into_profile(cli, bam, env, tmp_path):
    path, raw, kept, names, tlen, n = bam
    e = dict(os.environ, MSAMTOOLS_TIMING="1", **env)
    # (without --besthit `filter` reads no pre-flight sample, so the tiny-buffer case really runs on 1.2 MB buffers)
    hit = [] if "MSAMTOOLS_CHUNK_SLACK_KB" in env else ["--besthit"]
    for mode in ("-bu", "-b"):
        f = subprocess.run([cli, "filter", mode, "-l", "80"] + hit + [path], capture_output=True, env=e)
        assert f.returncode == 0, f.stderr.decode()[-3000:]
        out = str(tmp_path / "f.bam")
        open(out, "wb").write(f.stdout)
        got = samutil.read_bam(out)
        assert bytes(got.raw) == kept, mode
    # the -b stream piped into `profile` on stdin: every kept record arrives (the null device reports the count) and the table is the pattern
    outp = str(tmp_path / "p.gz")
    p = subprocess.run([cli, "profile", "--label", "S", "--unit", "ab", "--nolen", "-o", outp, "-"], input=f.stdout, capture_output=True, env=e)
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    comments, body = samutil.read_profile_gz(outp)
    n_kept = sum(1 for _ in _records(kept))
    assert any(c.startswith("# Mapped inserts") and f" {n_kept} (" in c for c in comments), comments
    want = [0.0 if i % 11 == 0 else ((i * 2654435761 & 0xffffffff) >> 7) / 1024.0 / (i % 13 + 1) for i in range(len(names))]
    assert list(body[0]) == ["ID", "S"] and body[1][0] == "Unknown"
    assert [k for k, _ in body[2:]] == names and [v for _, v in body[2:]] == ["%.8g" % x for x in want]


def _records(blob):
    o = 0
    while o < len(blob):
        bs = struct.unpack_from("<I", blob, o)[0]
        yield o
        o += 4 + bs


def test_coverage_writers(cli, bam, tmp_path):
    path, raw, kept, names, tlen, n = bam
    e = dict(os.environ, MSAMTOOLS_THREADS="3", MSAMTOOLS_CHUNK_RECORDS="20000")
    out = str(tmp_path / "c.gz")
    r = subprocess.run([cli, "coverage", "--summary", "-o", out, path], capture_output=True, env=e)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    lines = gzip.open(out, "rt").read().splitlines()
    want = [f"{nm}\t0\t0" if i % 3 == 1 else "%s\t%.8f\t%.2f" % (nm, (10 + i) / l, (20 + 3 * i) / l) for i, (nm, l) in enumerate(zip(names, tlen))]
    assert lines == want
    # per-position dump, -w 7, skipping uncovered sequences: the integer formatter, negative values, line breaks, the last value
    r = subprocess.run([cli, "coverage", "-x", "-w", "7", "-o", out, path], capture_output=True, env=e)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    text = gzip.open(out, "rt").read()
    exp = []
    for t, (nm, l) in enumerate(zip(names, tlen)):
        if t % 3 == 1:
            continue
        i = np.arange(int(l), dtype=np.uint64)
        d = (((i * 2654435761 + t) & 0xffffffff) >> 12).astype(np.int64) % 100003 - np.where(i % 97 == 0, 7, 0)
        exp.append(">" + nm)
        for a in range(0, int(l), 7):
            exp.append(" ".join(str(int(x)) for x in d[a:a + 7]))
    assert text == "\n".join(exp) + "\n"


def test_small_input_and_sam_output(cli, tmp_path):
    """inputs that fit the pre-flight sample take the thread-less path; SAM text in and out"""
    sam = "@HD\tVN:1.6\tSO:queryname\n@SQ\tSN:A\tLN:1000\n" + "".join(
        f"r{i // 2}\t{0 if i % 2 == 0 else 16}\tA\t{10 + i}\t60\t10M\t*\t0\t0\tACGTACGTAC\tIIIIIIIIII\tNM:i:0\tAS:i:5\n" for i in range(40))
    r = subprocess.run([cli, "filter", "-S", "-l", "5", "-"], input=sam.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    got = r.stdout.decode().splitlines()
    assert got == [l for l in sam.splitlines()[2:] if int(l.split("\t")[3]) % 5 != 1]      # POS is 0-based in the record: 1-based % 5 != 1


def test_sam_text_through_the_ring(cli):
    """more SAM records than the pre-flight sample: the record-wise reader feeds the ring (growable buffers), the writer thread formats SAM"""
    seq, q = "A" * 60, "I" * 60
    lines = ["@HD\tVN:1.6", "@SQ\tSN:A\tLN:100000000", "@SQ\tSN:B\tLN:1000"]
    for i in range(130_000):
        lines.append(f"q{i // 3:06d}\t0\tA\t{1 + i * 7 % 1000003}\t60\t60M\t*\t0\t0\t{seq}\t{q}\tAS:i:50\tNM:i:0")
    e = dict(os.environ, MSAMTOOLS_CHUNK_RECORDS="9000", MSAMTOOLS_THREADS="3")
    r = subprocess.run([cli, "filter", "-S", "-h", "--besthit", "-"], input=("\n".join(lines) + "\n").encode(), capture_output=True, env=e)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    out = [l for l in r.stdout.decode().splitlines() if not l.startswith("@")]
    assert out == [l for l in lines[3:] if (int(l.split("\t")[3]) - 1) % 5 != 0]


def test_error_paths_do_not_hang_or_leave_results(cli, bam, tmp_path):
    """mid-stream corruption and truncation are fatal from the reader thread while the GPU and writer threads are busy (exit 1, message,
    no profile file); a downstream process that closes the pipe early ends `filter` by SIGPIPE instead of leaving it blocked"""
    path, raw, kept, names, tlen, n = bam
    blob = open(path, "rb").read()
    e = dict(os.environ, MSAMTOOLS_CHUNK_RECORDS="5000", MSAMTOOLS_THREADS="3")
    bad = bytearray(blob)
    bad[len(bad) * 2 // 3] ^= 0x5a
    p1, p2, outp = str(tmp_path / "bad.bam"), str(tmp_path / "trunc.bam"), str(tmp_path / "p.gz")
    open(p1, "wb").write(bad)
    open(p2, "wb").write(blob[:len(blob) * 2 // 3])
    r = subprocess.run([cli, "filter", "-b", "-u", "-l", "80", "--besthit", p1], capture_output=True, env=e, timeout=120)
    assert r.returncode == 1 and b"Fatal Error: Cannot read input" in r.stderr, r.stderr[-500:]
    r = subprocess.run([cli, "profile", "--label", "x", "-o", outp, p2], capture_output=True, env=e, timeout=120)
    assert r.returncode == 1 and b"Fatal Error: Cannot read input" in r.stderr and not os.path.exists(outp), r.stderr[-500:]
    p = subprocess.Popen([cli, "filter", "-b", "-u", "-l", "80", "--besthit", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    p.stdout.read(3_000_000)
    p.stdout.close()
    assert p.wait(timeout=120) == -13                       # SIGPIPE
    p.stderr.close()


def _random_sam(n, seed):
    """well-formed SAM lines using every field form the spec allows: '*' fields, all CIGAR operators, every aux type"""
    import random
    rng = random.Random(seed)
    lines = []
    for i in range(n):
        lseq = rng.choice([0, 1, 2, 17, 50, 151])
        ops = []
        if rng.random() < 0.9 and lseq:
            left = lseq
            for _ in range(rng.randrange(1, 6)):
                if left <= 0:
                    break
                op = rng.choice("MIS=X")
                w = rng.randrange(1, left + 1)
                ops.append(f"{w}{op}")
                left -= w
                if rng.random() < 0.3:
                    ops.append(f"{rng.randrange(1, 2000)}{rng.choice('DNHP')}")
            if left > 0:
                ops.append(f"{left}M")
        cigar = "".join(ops) if ops else "*"
        seq = "".join(rng.choice("ACGTNRYKM=") for _ in range(lseq)) if lseq else "*"
        qual = "*" if (not lseq or rng.random() < 0.2) else "".join(chr(33 + rng.randrange(0, 60)) for _ in range(lseq))
        rname = rng.choice(["A", "B", "*"])
        aux = []
        tags = rng.sample(["NM", "AS", "XS", "MD", "RG", "XA", "XB", "XF", "XH", "ZZ", "X0", "Y1"], rng.randrange(0, 8))
        for t in tags:
            kind = rng.choice("iiiAfZHB") if t not in ("MD", "RG") else "Z"
            if kind == "i":
                v = rng.choice([0, 1, -1, 127, 128, 255, 256, -128, -129, 32767, 32768, 65535, 65536, -32768, -32769,
                                2147483647, -2147483648, 4294967295, rng.randrange(-1000, 1000)])
                aux.append(f"{t}:i:{v}")
            elif kind == "A":
                aux.append(f"{t}:A:{rng.choice('aZ!~5')}")
            elif kind == "f":
                aux.append(f"{t}:f:{rng.choice(['0', '1', '-2.5', '3.14159', '1e+10', '0.000123'])}")
            elif kind == "Z":
                aux.append(f"{t}:Z:" + "".join(rng.choice("ACGT^0123456789 _:;") for _ in range(rng.randrange(0, 40))))
            elif kind == "H":
                aux.append(f"{t}:H:" + "".join(rng.choice("0123456789ABCDEF") for _ in range(2 * rng.randrange(0, 9))))
            else:
                sub = rng.choice("cCsSiIf")
                lo, hi = {"c": (-128, 127), "C": (0, 255), "s": (-32768, 32767), "S": (0, 65535), "i": (-2 ** 31, 2 ** 31 - 1),
                          "I": (0, 2 ** 32 - 1), "f": (0, 0)}[sub]
                vals = [rng.choice(["1.5", "-0.25", "100"]) if sub == "f" else str(rng.choice([lo, hi, rng.randrange(lo, hi + 1)]))
                        for _ in range(rng.randrange(0, 6))]
                aux.append(f"{t}:B:{sub}" + "".join("," + v for v in vals))
        flag = rng.choice([0, 16, 4, 77, 141, 99, 147, 256, 2048, 1024 + 83])
        if rname == "*":                                      # an unplaced read: unmapped flag, no CIGAR (what aligners write)
            flag, cigar = flag | 4, "*"
        pos = 0 if rname == "*" else 5 * rng.randrange(0, 100_000) + 2            # (the null device keeps POS % 5 != 0, 0-based)
        fields = [f"r{i // 2:05d}", str(flag), rname, str(pos), str(rng.randrange(0, 255)), cigar,
                  rng.choice(["=", "*", "A", "B"]) if rname != "*" else "*", str(rng.randrange(0, 10 ** 6)), str(rng.randrange(-10 ** 5, 10 ** 5)), seq, qual] + aux
        lines.append("\t".join(fields))
    return lines


def test_sam_text_round_trip_is_a_fixed_point(cli):
    """SAM -> records -> SAM through the CLI's own reader and writer (null device keeps everything here): the output re-read
    gives itself again (one normalisation, then a fixed point), no record is lost, and the eleven mandatory fields, the tag
    order and every Z / A / H / integer value survive exactly as written"""
    hdr = "@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:A\tLN:1000000\n@SQ\tSN:B\tLN:2000000\n"
    lines = _random_sam(3000, 77)
    src = (hdr + "\n".join(lines) + "\n").encode()
    outs = []
    for _ in range(2):
        r = subprocess.run([cli, "filter", "-S", "-l", "1", "-"], input=src if not outs else (hdr.encode() + outs[-1]), capture_output=True)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1]
    got = outs[0].decode().splitlines()
    keep = lines                                              # every POS was chosen so that the null device keeps the record
    assert len(got) == len(keep)
    for a, b in zip(keep, got):
        fa, fb = a.split("\t"), b.split("\t")
        if fa[6] == fa[2] and fa[2] != "*":
            fa[6] = "="                                       # RNEXT equal to RNAME prints as "="
        assert fa[:11] == fb[:11], (a, b)
        assert [x[:2] for x in fa[11:]] == [x[:2] for x in fb[11:]]
        for x, y in zip(fa[11:], fb[11:]):
            t = x[3]
            if t in "ZAH":
                assert x == y
            elif t == "i":
                assert y[3] == "i" and int(x[5:]) == int(y[5:]), (x, y)
            elif t == "f":
                assert y[3] == "f" and abs(float(x[5:]) - float(y[5:])) <= 1e-6 * max(1.0, abs(float(x[5:]))), (x, y)
            else:
                assert x[:6] == y[:6] and [float(v) for v in x[7:].split(",") if v] == [float(v) for v in y[7:].split(",") if v], (x, y)


def test_sam_to_bam_encoding_matches_the_python_encoder(cli, tmp_path):
    """the C host's SAM -> BAM record encoding (bin, smallest integer aux type, 4-bit sequence, '*' quality, B arrays) against
    tests/samutil.py's independent encoder (the one the golden fixtures were built with), byte for byte, on the random lines"""
    hdr = "@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:A\tLN:1000000\n@SQ\tSN:B\tLN:2000000\n"
    lines = _random_sam(2000, 99)
    r = subprocess.run([cli, "filter", "-S", "-b", "-l", "1", "-"], input=(hdr + "\n".join(lines) + "\n").encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    out = str(tmp_path / "o.bam")
    open(out, "wb").write(r.stdout)
    got = bytes(samutil.read_bam(out).raw)
    want = [samutil.encode_record(l.split("\t"), {"A": 0, "B": 1}) for l in lines]
    o = 0
    for l, w in zip(lines, want):
        assert got[o:o + len(w)] == w, l
        o += len(w)
    assert o == len(got)


def test_long_records_through_small_buffers(cli, tmp_path):
    """records of 150 b to 60 kb (long reads): more than a chunk in the pre-flight sample, few records per buffer, records
    spanning many BGZF blocks -- the record stream comes out byte for byte through both filter paths and reaches `profile`"""
    import random
    rng = random.Random(5)
    names, tlen = ["A", "B"], [5_000_000, 5_000_000]
    recs = []
    for i in range(900):
        L = rng.choice([20000, 35000, 150, 60000])
        seq = ("".join(rng.choice("ACGT") for _ in range(64)) * (L // 64 + 1))[:L]
        f = [f"long{i // 2:05d}", "0", rng.choice(names), str(5 * rng.randrange(0, 900000) + 2), "60", f"{L}M", "*", "0", "0", seq, "*", "NM:i:3", "AS:i:100"]
        recs.append(samutil.encode_record(f, {"A": 0, "B": 1}))
    raw = b"".join(recs)
    path = str(tmp_path / "long.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, np.frombuffer(raw, dtype=np.uint8), level=1)
    for env in (dict(MSAMTOOLS_CHUNK_MB="1", MSAMTOOLS_CHUNK_SLACK_KB="192"), {}):
        e = dict(os.environ, MSAMTOOLS_THREADS="4", **env)
        for hit in ([], ["--besthit"]):
            r = subprocess.run([cli, "filter", "-bu", "-l", "80"] + hit + [path], capture_output=True, env=e, timeout=300)
            assert r.returncode == 0, r.stderr.decode()[-2000:]
            out = str(tmp_path / "o.bam")
            open(out, "wb").write(r.stdout)
            assert bytes(samutil.read_bam(out).raw) == raw, (env, hit)
        outp = str(tmp_path / "p.gz")
        p = subprocess.run([cli, "profile", "--label", "x", "-o", outp, path], capture_output=True, env=e, timeout=300)
        assert p.returncode == 0, p.stderr.decode()[-2000:]
        assert any(c.startswith("# Mapped inserts") and " 900 (" in c for c in samutil.read_profile_gz(outp)[0])
