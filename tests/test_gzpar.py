"""CPU: the parallel gzip writer of the CLI's table outputs (csrc/host/gzpar.c) under ASan/UBSan -- every thread count and
write pattern gives ONE gzip member that zlib, Python's gzip module and `zcat` inflate to exactly the bytes written."""
import gzip
import os
import shutil
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "msamtools_b200", "csrc", "host")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    d = tmp_path_factory.mktemp("gzpar")
    exe = str(d / "gzp")
    subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", HOST,
                    os.path.join(ROOT, "tests", "c", "gzpar_harness.c"), os.path.join(HOST, "gzpar.c"), os.path.join(HOST, "crc32x.c"), "-lz", "-lpthread", "-o", exe], check=True)
    return exe, d


def table(rows, seed):
    rng = np.random.default_rng(seed)
    v = rng.random(rows) * (rng.random(rows) < 0.7)
    return ("ID\tS\n" + "".join("g%07d\t%.8g\n" % (i, x) for i, x in enumerate(v))).encode()


PAYLOADS = {
    "empty": b"",
    "one_byte": b"x",
    "small_table": table(50, 1),
    "just_under_a_block": table(1, 2) * 3 + b"a" * ((512 << 10) - 40),
    "table_9MB": table(400_000, 3),                                    # 18 deflate jobs, several batches at 1 and 2 threads
    "binary_3MB": np.random.default_rng(4).integers(0, 256, 3_000_000, dtype=np.uint8).tobytes(),   # incompressible: stored blocks
}


@pytest.mark.parametrize("name", list(PAYLOADS))
def test_one_member_and_same_bytes(harness, name):
    exe, d = harness
    data = PAYLOADS[name]
    for thr in (1, 2, 7):
        for mode in (("write", "pieces") if len(data) < 10_000_000 else ("write",)):
            out = str(d / f"{name}_{thr}_{mode}.gz")
            r = subprocess.run([exe, out, str(thr), mode], input=data, capture_output=True)
            assert r.returncode == 0, r.stderr.decode()
            blob = open(out, "rb").read()
            z = zlib.decompressobj(16 + 15)                            # exactly one gzip member: nothing may be left behind it
            got = z.decompress(blob) + z.flush()
            assert z.eof and z.unused_data == b"" and got == data, (thr, mode)
            assert gzip.decompress(blob) == data
            if shutil.which("zcat") and thr == 7:
                assert subprocess.run(["zcat", out], capture_output=True, check=True).stdout == data


def test_compresses_like_gzopen(harness):
    """same level as gzopen(..., "wb"): splitting into jobs may cost a little, never more than 3 %"""
    exe, d = harness
    data = PAYLOADS["table_9MB"]
    out = str(d / "ratio.gz")
    subprocess.run([exe, out, "4", "write"], input=data, check=True)
    assert os.path.getsize(out) <= 1.03 * len(gzip.compress(data, 6))
