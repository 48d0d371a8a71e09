"""CPU: the host BGZF reader (bamio.c) with parallel block inflation gives the same record stream as the single-threaded
streaming reader and as the stream the file was written from; a corrupted block is reported, not passed on."""
import os
import shutil
import subprocess

import pytest

import samutil
from conftest import ROOT

HOST = os.path.join(ROOT, "msamtools_b200", "csrc", "host")


@pytest.fixture(scope="module", params=["slots of 24 MB", "slots of 70001 bytes"])
def harness(tmp_path_factory, request):
    """built twice: with the product's read-ahead slot size, and with slots of little more than one BGZF block, so that the
    hand-over between slots (tail of one carried in front of the next) happens thousands of times on a small file"""
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    d = tmp_path_factory.mktemp("bamio")
    exe = str(d / "reader")
    small = ["-DIO_BATCH=((size_t)70001)"] if "70001" in request.param else []
    srcs = [os.path.join(ROOT, "tests", "c", "bamio_read_harness.c"), os.path.join(HOST, "bamio.c"), os.path.join(HOST, "crc32x.c")]
    if os.path.exists(os.path.join(HOST, "finflate.c")):
        srcs.append(os.path.join(HOST, "finflate.c"))
    subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", HOST] + small + srcs +
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
    for thr in (0, 1, 2, 5):
        r = subprocess.run([exe, path, str(thr)], capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        outs[thr] = r.stdout
        fast, zl = (int(x) for x in r.stderr.decode().split()[1::2])
        if thr == 0:
            assert (fast, zl) == (0, 0)                 # bio_set_threads never called: streaming zlib path
        else:
            assert fast > 0 and zl == 0                 # block-wise (one thread included): every block of a zlib-written file is taken by the fast decoder
    assert outs[0] == raw and outs[1] == raw and outs[2] == raw and outs[5] == raw


def test_pipe_input_through_the_read_ahead_thread(harness, stream):
    """stdin instead of a file, fed in dribbles by a slow producer: the read-ahead thread's slots are handed over in order and the
    tail of one slot is carried into the next (a file much larger than one slot is covered by the bulk-reader test below)"""
    exe, d = harness
    raw, names, tlen = stream
    import numpy as np
    path = str(d / "pipe.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, np.frombuffer(raw, dtype=np.uint8), level=1)
    blob = open(path, "rb").read()
    for thr in (2, 4):
        p = subprocess.Popen([exe, "-", str(thr)], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        import threading
        out = []
        t = threading.Thread(target=lambda: out.append(p.stdout.read()))
        t.start()
        for o in range(0, len(blob), 300_001):
            p.stdin.write(blob[o:o + 300_001]); p.stdin.flush()
        p.stdin.close()
        t.join()
        assert p.wait() == 0, p.stderr.read().decode()
        assert out[0] == raw


def test_corrupt_block_is_reported(harness, stream):
    exe, d = harness
    raw, names, tlen = stream
    import numpy as np
    path = str(d / "bad.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, np.frombuffer(raw, dtype=np.uint8), level=6)
    blob = bytearray(open(path, "rb").read())
    blob[len(blob) // 2] ^= 0x10                       # somewhere inside a block's deflate payload
    open(path, "wb").write(bytes(blob))
    for thr in (0, 1, 4):
        r = subprocess.run([exe, path, str(thr)], capture_output=True)
        assert r.returncode == 1 and (b"corrupt" in r.stderr or b"inflate" in r.stderr or b"CRC" in r.stderr or b"truncated" in r.stderr), r.stderr


@pytest.fixture(scope="module", params=["slots of 24 MB", "slots of 70001 bytes"])
def bulk(tmp_path_factory, request):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    d = tmp_path_factory.mktemp("bamio_bulk")
    exe = str(d / "bulk")
    small = ["-DIO_BATCH=((size_t)70001)"] if "70001" in request.param else []
    srcs = [os.path.join(ROOT, "tests", "c", "bamio_bulk_harness.c"), os.path.join(HOST, "bamio.c"), os.path.join(HOST, "finflate.c"), os.path.join(HOST, "crc32x.c")]
    subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", HOST] + small + srcs +
                   ["-lz", "-lpthread", "-o", exe], check=True)
    return exe, d


@pytest.mark.parametrize("level", [0, 6])
def test_bulk_reader_into_fixed_buffer(bulk, stream, level):
    """bio_read_raw (what the CLI's reader thread uses): BGZF blocks inflated straight into a fixed-capacity buffer, for buffer
    sizes below / around / far above a batch of blocks, with and without worker threads: always the original record stream"""
    exe, d = bulk
    raw, names, tlen = stream
    import numpy as np
    path = str(d / f"r{level}.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, np.frombuffer(raw, dtype=np.uint8), level=level)
    for thr in (0, 1, 3):
        for cap in (70_000, 200_001, 64 << 20):
            r = subprocess.run([exe, "read", path, str(thr), str(cap)], capture_output=True)
            assert r.returncode == 0, r.stderr.decode()
            assert r.stdout == raw, (thr, cap)


@pytest.mark.parametrize("mode", ["wbu", "wb"])
def test_bulk_writer_equals_recordwise(bulk, stream, mode):
    """bio_write_raw (blocks packed on worker threads; level 0 as hand-made stored blocks) writes byte for byte what
    bio_write_record writes, and the file reads back as the original stream"""
    exe, d = bulk
    raw, names, tlen = stream
    import numpy as np
    path = str(d / "w_in.bam")
    samutil.write_bam(path, samutil.synth_header(names, tlen), names, tlen, np.frombuffer(raw, dtype=np.uint8), level=1)
    outs = {}
    for how, thr in (("copyr", 1), ("copy", 1), ("copy", 4)):
        out = str(d / f"w_{how}_{thr}_{mode}.bam")
        r = subprocess.run([exe, how, path, str(thr), mode, out], capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        outs[(how, thr)] = open(out, "rb").read()
        assert bytes(samutil.read_bam(out).raw) == raw
    assert outs[("copy", 1)] == outs[("copyr", 1)] and outs[("copy", 4)] == outs[("copyr", 1)]
