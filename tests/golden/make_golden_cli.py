#!/usr/bin/env python
"""Golden vector for --genome row order: runs the reference's own object code (oracle/_ref/msamtools,
build container only) on the same synthetic input tests/test_cli_gpu.py::test_cli_profile_genome_order
builds, and records the order of the feature rows (zoeHash key order, zoeTools.c:218-372)."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import samutil                       # noqa: E402
from msamtools_b200 import synth     # noqa: E402

p = synth.make_params("mixed", n_records=20_000, seed=5)
raw, off, _ = synth.generate(p)
tlen = synth.target_lengths(p)
names = [f"seq{i:03d}" for i in range(len(tlen))]
genome_of = [f"genome_{(i * 7) % 40:02d}" for i in range(len(tlen))]
with tempfile.TemporaryDirectory() as d:
    bam, gdef, out = os.path.join(d, "g.bam"), os.path.join(d, "g.tsv"), os.path.join(d, "g.gz")
    samutil.write_bam(bam, samutil.synth_header(names, tlen), names, tlen, raw)
    with open(gdef, "w") as fh:
        for g, n in zip(genome_of, names):
            fh.write(f"{g}\t{n}\n")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "msamtools"), "profile", "--label", "g", "--unit", "ab", "--nolen", "--multi", "prop",
                    "--genome", gdef, "-o", out, bam], check=True, stderr=subprocess.DEVNULL)
    _, body = samutil.read_profile_gz(out)
order = [k for k, _ in body[2:]]
with open(os.path.join(HERE, "genome_order_40.txt"), "w") as fh:
    fh.write("\n".join(order) + "\n")
print(len(order), order[:6])
