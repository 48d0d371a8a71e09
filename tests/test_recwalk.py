"""CPU: the CLI's multi-threaded record index (csrc/host/recwalk.c) under ASan/UBSan equals the sequential walk of the
block_size chain (msg_index_records) -- on ordinary streams, on streams with a partial trailing record, on streams whose
payload CONTAINS runs of well-formed fake records (the guess lands off the chain and must be caught by the
verification), and reports a corrupt block_size on the true chain."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "msamtools_b200", "csrc", "host")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    d = tmp_path_factory.mktemp("recwalk")
    exe = str(d / "rw")
    subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", HOST,
                    os.path.join(ROOT, "tests", "c", "recwalk_harness.c"), os.path.join(HOST, "recwalk.c"), "-DRW_DEBUG", "-lpthread", "-o", exe], check=True)
    return exe


def walk(raw, start=0):
    offs, o = [start], start
    while o + 4 <= len(raw):
        bs = struct.unpack_from("<I", raw, o)[0]
        assert 32 <= bs
        if o + 4 + bs > len(raw):
            break
        o += 4 + bs
        offs.append(o)
    return offs


def run(exe, raw, threads, n_targets=0, start=0):
    r = subprocess.run([exe, str(threads), str(n_targets), str(start)], input=raw, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    lines = r.stdout.split()
    rc, n = int(lines[0]), int(lines[1])
    v = r.stderr.decode().split()                                   # "rw_index: V of T segments verified" (absent: sequential walk)
    run.verified = (int(v[1]), int(v[3])) if v else None
    return rc, [int(x) for x in lines[2:]] if rc == 0 else None, n


@pytest.fixture(scope="module")
def stream():
    from msamtools_b200 import synth
    p = synth.make_params("mixed", n_records=150_000, seed=5)          # ~ 45 MB: 8 threads get > 4 MB each
    raw, off, _ = synth.generate(p)
    n = len(off) - 1
    return bytes(raw[:int(off[n])]), [int(x) for x in off[:n + 1]], len(synth.target_lengths(p))


def test_equals_sequential_walk(harness, stream):
    raw, off, nt = stream
    for thr in (1, 2, 8, 11):
        rc, got, n = run(harness, raw, thr, nt)
        assert rc == 0 and got == off, thr
        assert run.verified is None if thr == 1 else run.verified[0] == run.verified[1] >= min(thr, 8)   # every guess was on the chain
    # not from the start of the buffer, and with a partial trailing record (what a bulk read leaves)
    k = 1234
    cut = raw[:len(raw) - 77]
    rc, got, n = run(harness, cut, 8, nt, start=off[k])
    assert rc == 0 and got == off[k:-1]
    # n_targets unknown
    rc, got, n = run(harness, raw, 8, 0)
    assert rc == 0 and got == off


def fake_record(tid, name, seqlen, aux=b""):
    body = struct.pack("<iiBBHHHiiii", tid, 100, len(name) + 1, 30, 4680, 1, 0, seqlen, -1, -1, 0) + name + b"\0" + struct.pack("<I", seqlen << 4)
    body += bytes((seqlen + 1) // 2) + bytes([30]) * seqlen + aux
    return struct.pack("<I", len(body)) + body


def test_fake_chains_inside_payload_are_caught(harness):
    """every record carries, inside a Z tag, 12 complete well-formed records at an odd offset: wherever a thread starts
    looking it finds a plausible run that is NOT on the chain; the result must still be the true chain"""
    inner = b"".join(fake_record(3, b"fake%04d" % i, 50) for i in range(12))
    assert b"\0" in inner                                           # (a real Z tag could not hold it; the index does not care)
    recs = [fake_record(i % 7, b"read%07d" % (i // 2), 100, aux=b"XYZ" + b"q" * ((i * i) % 13) + inner + b"\0") for i in range(24_011)]
    raw = b"".join(recs)
    assert len(raw) > 8 * (4 << 20)                                # enough for 8 segments
    want = walk(raw)
    for thr in (2, 8):
        rc, got, n = run(harness, raw, thr, 10)
        assert rc == 0 and got == want, thr
        if thr == 8:
            assert run.verified[0] < run.verified[1]                    # (deterministic input) a guess was off the chain and was caught


def test_random_bytes_and_corruption(harness, stream):
    raw, off, nt = stream
    bad = bytearray(raw)
    k = len(off) * 3 // 4
    bad[off[k]:off[k] + 4] = struct.pack("<I", 7)                   # block_size < 32 on the true chain
    for thr in (1, 8):
        rc, got, n = run(harness, bytes(bad), thr, nt)
        assert rc == -1, thr
    # a stream that is one huge record followed by small ones: segments without any record start
    big = fake_record(0, b"big", 30_000_000)
    tail = b"".join(fake_record(1, b"t%05d" % i, 80) for i in range(50_000))
    raw2 = big + tail
    for thr in (4, 8):
        rc, got, n = run(harness, raw2, thr, 5)
        assert rc == 0 and got == walk(raw2), thr
