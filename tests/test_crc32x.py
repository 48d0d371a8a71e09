"""CPU: crc32x (csrc/host/crc32x.c: CRC-32 by carry-less multiplication, zlib's convention) returns zlib's crc32 for every
length 0..4100 at several alignments and seeds, on chained calls and on a large buffer -- under ASan/UBSan."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "msamtools_b200", "csrc", "host")


def test_crc32x_equals_zlib(tmp_path):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    exe = str(tmp_path / "crc")
    subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", HOST,
                    os.path.join(ROOT, "tests", "c", "crc32x_harness.c"), os.path.join(HOST, "crc32x.c"), "-lz", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
    if "pclmul" in open("/proc/cpuinfo").read():
        assert r.stdout.split()[1] == "1"                # the accelerated path is what was compared on this machine
