"""CPU: the fast BGZF-payload inflater (msamtools_b200/csrc/host/finflate.c) against zlib.
Valid raw-DEFLATE streams of every block type must decode to zlib's output (or be declined -- only streams with an
unusual code set may be, never the common ones marked must_succeed); corrupted / truncated streams must never make it
read or write outside its buffers (the harness runs under AddressSanitizer + UBSan) and, if accepted, would be caught by
the CRC check in bamio.c, which is exercised separately below."""
import os
import random
import shutil
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "msamtools_b200", "csrc", "host")


def raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, memlevel, strategy)
    return c.compress(data) + c.flush()


def corpus():
    rng = random.Random(1234)
    nprng = np.random.default_rng(99)
    from msamtools_b200 import synth
    p = synth.make_params("mixed", n_records=3000, seed=5)
    bam = bytes(synth.generate(p)[0])
    texts = [b"", b"a", b"ab", b"abc" * 5, bytes(range(256)) * 4, b"\x00" * 70000, b"I" * 300 + b"ACGT" * 2000,
             bytes(nprng.integers(0, 256, 65536, dtype=np.uint8)), bytes(nprng.integers(0, 4, 65280, dtype=np.uint8)),
             bam[:65280], bam[65280:2 * 65280], bam[100000:100000 + 1500], bam[:17],
             b"".join(bytes([rng.randrange(256)]) * rng.randrange(1, 300) for _ in range(400))[:65280]]
    # periodic data: long matches whose distance is shorter than the match (overlapping copies), one text per period around the
    # 8- and 16-byte copy units of the decoder
    texts += [bytes(rng.randrange(256) for _ in range(per)) * (9000 // per) + b"x" for per in (2, 3, 5, 7, 8, 9, 10, 12, 15, 16, 17, 23, 31, 33)]
    cases = []
    for t in texts:
        for level in (0, 1, 6, 9):
            cases.append((raw_deflate(t, level), t, True))
        cases.append((raw_deflate(t, 6, zlib.Z_FIXED), t, True))           # fixed Huffman blocks
        cases.append((raw_deflate(t, 6, zlib.Z_HUFFMAN_ONLY), t, True))    # no matches: empty / one-code distance alphabet
        cases.append((raw_deflate(t, 6, zlib.Z_RLE), t, True))             # distance-1 matches only
        cases.append((raw_deflate(t, 9, zlib.Z_FILTERED, 1), t, True))     # memlevel 1: many small blocks
    # several deflate blocks in one stream (Z_FULL_FLUSH inserts empty stored blocks between them)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    multi = c.compress(bam[:20000]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(bam[20000:41000]) + c.flush(zlib.Z_SYNC_FLUSH) + c.compress(b"tail") + c.flush()
    cases.append((multi, bam[:41000] + b"tail", True))
    # corrupted and truncated variants: any answer is fine, crashes and wrong "ok" outputs are not
    valid = list(cases)
    for comp, t, _ in valid:
        if len(comp) < 4:
            continue
        for _ in range(6):
            b = bytearray(comp)
            k = rng.randrange(len(b))
            b[k] ^= 1 << rng.randrange(8)
            try:
                exp = zlib.decompressobj(-15).decompress(bytes(b))
            except zlib.error:
                exp = None
            if exp is not None and len(exp) == len(t):
                cases.append((bytes(b), exp, False))           # still a valid stream of the same size: output must match zlib's
            else:
                cases.append((bytes(b), b"\xee" * len(t), False))   # must be rejected; an (impossible) acceptance is flagged by the mismatch
        cut = rng.randrange(1, len(comp))
        cases.append((comp[:cut], b"\xee" * len(t), False))
        cases.append((comp, b"\xee" * (len(t) + 1), False))     # wrong expected size
        if len(t) > 0:
            cases.append((comp, b"\xee" * (len(t) - 1), False))
    return cases


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    d = tmp_path_factory.mktemp("finflate")
    exe = str(d / "harness")
    cmd = ["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", HOST,
           os.path.join(ROOT, "tests", "c", "finflate_harness.c"), os.path.join(HOST, "finflate.c"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe, d


def test_finflate_matches_zlib_and_stays_in_bounds(harness):
    exe, d = harness
    cases = corpus()
    path = str(d / "cases.bin")
    with open(path, "wb") as fh:
        for comp, exp, must in cases:
            fh.write(struct.pack("<IIB", len(comp), len(exp), 1 if must else 0))
            fh.write(comp)
            fh.write(exp)
    r = subprocess.run([exe, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n_must = sum(1 for c in cases if c[2])
    words = r.stdout.split()
    assert int(words[1]) == len(cases) and int(words[3]) >= n_must, r.stdout
