"""CPU: the host BGZF reader (bamio.c) with parallel block inflation gives the same record stream as the single-threaded
streaming reader and as the stream the file was written from; a corrupted block is reported, not passed on."""
import os
import shutil
import subprocess

import pytest

import samutil
from conftest import ROOT

HOST = os.path.join(ROOT, "msamtools_b200", "csrc", "host")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    d = tmp_path_factory.mktemp("bamio")
    exe = str(d / "reader")
    srcs = [os.path.join(ROOT, "tests", "c", "bamio_read_harness.c"), os.path.join(HOST, "bamio.c")]
    if os.path.exists(os.path.join(HOST, "finflate.c")):
        srcs.append(os.path.join(HOST, "finflate.c"))
    subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", HOST] + srcs +
                   ["-lz", "-lpthread", "-o", exe], check=True)
    return exe, d


@pytest.fixture(scope="module")
def stream():
    from msamtools_b200 import synth
    p = synth.make_params("mixed", n_records=30_000, seed=77)
    raw, off, _ = synth.generate(p)
    tlen = synth.target_lengths(p)
    names = [f"ref{i:04d}" for i in range(len(tlen))]
    return bytes(raw), names, tlen


@pytest.mark.parametrize("level", [0, 1, 6])
def test_parallel_inflate_equals_streaming(harness, stream, level):
    exe, d = harness
    raw, names, tlen = stream
    path = str(d / f"in{level}.bam")
    import numpy as np
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, np.frombuffer(raw, dtype=np.uint8), level=level)
    outs = {}
    for thr in (1, 2, 5):
        r = subprocess.run([exe, path, str(thr)], capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        outs[thr] = r.stdout
        fast, zl = (int(x) for x in r.stderr.decode().split()[1::2])
        if thr == 1:
            assert (fast, zl) == (0, 0)                 # streaming zlib path
        else:
            assert fast > 0 and zl == 0                 # every block of a zlib-written file is taken by the fast decoder
    assert outs[1] == raw and outs[2] == raw and outs[5] == raw


def test_corrupt_block_is_reported(harness, stream):
    exe, d = harness
    raw, names, tlen = stream
    import numpy as np
    path = str(d / "bad.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, np.frombuffer(raw, dtype=np.uint8), level=6)
    blob = bytearray(open(path, "rb").read())
    blob[len(blob) // 2] ^= 0x10                       # somewhere inside a block's deflate payload
    open(path, "wb").write(bytes(blob))
    for thr in (1, 4):
        r = subprocess.run([exe, path, str(thr)], capture_output=True)
        assert r.returncode == 1 and (b"corrupt" in r.stderr or b"inflate" in r.stderr or b"CRC" in r.stderr or b"truncated" in r.stderr), r.stderr
