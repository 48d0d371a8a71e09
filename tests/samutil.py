"""Test helpers: SAM text / BGZF BAM -> the raw BAM record stream + offset index the C ABI takes.

Encoding follows the SAM spec (section 4.2) and htslib's sam_parse1 conventions that the
reference relies on: integer aux values get the smallest BAM type, `*` SEQ gives l_seq 0,
`bin` is reg2bin(pos, end).  Pure Python; only for fixture-sized inputs.
"""
import gzip
import re
import struct

import numpy as np

CIGAR_OPS = "MIDNSHP=XB"
_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _aux_int(tag, v):
    t = tag.encode()
    if v >= 0:
        if v <= 0xff: return t + b"C" + struct.pack("<B", v)
        if v <= 0xffff: return t + b"S" + struct.pack("<H", v)
        return t + b"I" + struct.pack("<I", v)
    if v >= -128: return t + b"c" + struct.pack("<b", v)
    if v >= -32768: return t + b"s" + struct.pack("<h", v)
    return t + b"i" + struct.pack("<i", v)


def encode_aux(field):
    tag, ty, val = field.split(":", 2)
    if ty == "i": return _aux_int(tag, int(val))
    if ty == "A": return tag.encode() + b"A" + val.encode()[:1]
    if ty == "f": return tag.encode() + b"f" + struct.pack("<f", float(val))
    if ty in "ZH": return tag.encode() + ty.encode() + val.encode() + b"\0"
    if ty == "B":
        parts = val.split(",")
        st = parts[0]
        fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[st]
        vals = [float(x) if st == "f" else int(x) for x in parts[1:]]
        return tag.encode() + b"B" + st.encode() + struct.pack("<i", len(vals)) + struct.pack("<%d%s" % (len(vals), fmt), *vals)
    raise ValueError(f"unsupported aux type {ty}")


def encode_record(fields, ref_index):
    qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = fields[:11]
    flag, pos, mapq, pnext, tlen = int(flag), int(pos) - 1, int(mapq), int(pnext) - 1, int(tlen)
    tid = -1 if rname == "*" else ref_index[rname]
    mtid = -1 if rnext == "*" else (tid if rnext == "=" else ref_index[rnext])
    ops = [] if cigar == "*" else [(int(n), CIGAR_OPS.index(o)) for n, o in re.findall(r"(\d+)([MIDNSHP=XB])", cigar)]
    rlen = sum(n for n, o in ops if o in (0, 2, 3, 7, 8))
    if (flag & 4) or rlen == 0:
        rlen = 1
    b = reg2bin(max(pos, 0), max(pos, 0) + rlen) if pos >= 0 else 4680
    l_seq = 0 if seq == "*" else len(seq)
    name = qname.encode() + b"\0"
    core = struct.pack("<iiBBHHHiiii", tid, pos, len(name), mapq, b, len(ops), flag, l_seq, mtid, pnext, tlen)
    cig = b"".join(struct.pack("<I", n << 4 | o) for n, o in ops)
    sq = bytearray((l_seq + 1) // 2)
    for i, ch in enumerate(seq if l_seq else ""):
        code = _SEQ_CODE.get(ch.upper(), 15)
        sq[i >> 1] |= code << (4 if i % 2 == 0 else 0)
    if l_seq:
        ql = bytes([0xff] * l_seq) if qual == "*" else bytes(ord(c) - 33 for c in qual)
    else:
        ql = b""
    aux = b"".join(encode_aux(f) for f in fields[11:])
    body = core + name + cig + bytes(sq) + ql + aux
    return struct.pack("<i", len(body)) + body


class Sam:
    """Parsed alignment file: header lines, references, raw record stream + offsets."""

    def __init__(self, header_lines, refs, raw, off):
        self.header_lines = header_lines
        self.ref_names = [r[0] for r in refs]
        self.target_len = np.array([r[1] for r in refs], dtype=np.uint32)
        self.raw = raw
        self.off = off

    @property
    def n(self):
        return len(self.off) - 1

    def sort_order(self):
        for h in self.header_lines:
            if h.startswith("@HD"):
                for f in h.split("\t")[1:]:
                    if f.startswith("SO:"):
                        return f[3:]
        return None

    def record_fields(self, i):
        """(qname, flag, tid) of record i -- enough for the reference tests' `QNAME:FLAG` assertions."""
        o = int(self.off[i])
        tid, = struct.unpack_from("<i", self.raw, o + 4)
        lq = int(self.raw[o + 12])
        flag, = struct.unpack_from("<H", self.raw, o + 18)
        return bytes(self.raw[o + 36:o + 36 + lq - 1]).decode(), int(flag), int(tid)

    def name_flags(self, idx):
        return ",".join("%s:%d" % self.record_fields(int(i))[:2] for i in idx)


def parse_sam_text(text):
    header, refs, recs = [], [], []
    for line in text.splitlines():
        if not line:
            continue
        if line.startswith("@"):
            header.append(line)
            if line.startswith("@SQ"):
                d = dict(f.split(":", 1) for f in line.split("\t")[1:])
                refs.append((d["SN"], int(d["LN"])))
            continue
        recs.append(line.split("\t"))
    ref_index = {n: i for i, (n, _) in enumerate(refs)}
    chunks, off, o = [], [0], 0
    for f in recs:
        b = encode_record(f, ref_index)
        chunks.append(b)
        o += len(b)
        off.append(o)
    raw = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy() if chunks else np.zeros(0, dtype=np.uint8)
    return Sam(header, refs, raw, np.array(off, dtype=np.uint64))


def read_sam(path):
    with open(path) as fh:
        return parse_sam_text(fh.read())


def read_bam(path):
    """BGZF = concatenated gzip members; python's gzip handles multi-member streams (through the file object: gzip.decompress
    re-slices the remaining input once per member, which is quadratic on files of thousands of blocks)."""
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    assert data[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", data, 4)
    text = data[8:8 + l_text].split(b"\0")[0].decode()
    o = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, o); o += 4
    refs = []
    for _ in range(n_ref):
        ln, = struct.unpack_from("<i", data, o); o += 4
        name = data[o:o + ln - 1].decode(); o += ln
        tl, = struct.unpack_from("<i", data, o); o += 4
        refs.append((name, tl))
    raw = np.frombuffer(data[o:], dtype=np.uint8).copy()
    off = [0]
    p = 0
    while p + 4 <= len(raw):
        bs, = struct.unpack_from("<i", raw, p)
        p += 4 + bs
        off.append(p)
    return Sam([l for l in text.splitlines() if l], refs, raw, np.array(off, dtype=np.uint64))


def split_records(raw):
    """raw record stream -> list of bytes objects"""
    out, p = [], 0
    raw = bytes(raw)
    while p + 4 <= len(raw):
        bs, = struct.unpack_from("<i", raw, p)
        out.append(raw[p:p + 4 + bs])
        p += 4 + bs
    return out


def write_bam(path, header_text, ref_names, target_len, raw, level=1):
    """Write a BGZF-compressed BAM file (concatenated gzip members with the BC extra field)."""
    import zlib
    text = header_text.encode()
    parts = [b"BAM\1", struct.pack("<i", len(text)), text, struct.pack("<i", len(ref_names))]
    for n, l in zip(ref_names, target_len):
        nb = n.encode() + b"\0"
        parts += [struct.pack("<i", len(nb)), nb, struct.pack("<i", int(l))]
    data = b"".join(parts) + bytes(raw)
    with open(path, "wb") as fh:
        for o in range(0, len(data), 0xff00):
            blk = data[o:o + 0xff00]
            co = zlib.compressobj(level, zlib.DEFLATED, -15)
            comp = co.compress(blk) + co.flush()
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25))
            fh.write(comp + struct.pack("<II", zlib.crc32(blk) & 0xffffffff, len(blk)))
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def synth_header(ref_names, target_len, so="queryname"):
    lines = ["@HD\tVN:1.6\tSO:%s" % so] if so else []
    lines += ["@SQ\tSN:%s\tLN:%d" % (n, int(l)) for n, l in zip(ref_names, target_len)]
    return "\n".join(lines) + "\n"


def read_profile_gz(path):
    """msamtools profile output -> (header comment lines, {feature: value string})"""
    with gzip.open(path, "rt") as fh:
        lines = fh.read().splitlines()
    comments = [l for l in lines if l.startswith("#")]
    body = [l.split("\t") for l in lines if not l.startswith("#")]
    return comments, body
