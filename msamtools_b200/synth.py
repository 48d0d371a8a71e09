"""ctypes view of libmsamsynth.so (csrc/synth.c): synthetic name-sorted BAM record streams.

Workload presets follow SURVEY.md 8(d) / BASELINE.json configs.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsamsynth.so")


class SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_records", C.c_uint64), ("qname_base", C.c_uint64),
        ("n_refs", C.c_int32), ("read_len", C.c_int32), ("insert_len", C.c_int32),
        ("ref_len_min", C.c_uint32), ("ref_len_max", C.c_uint32),
        ("abund_sigma", C.c_double), ("shared_fraction", C.c_double), ("single_fraction", C.c_double),
        ("clip_fraction", C.c_double), ("indel_fraction", C.c_double), ("unmapped_fraction", C.c_double),
        ("max_occ", C.c_int32), ("alt_noise", C.c_int32), ("minimal_aux", C.c_int32),
    ]


PRESETS = {
    # BASELINE.json configs[1]: synthetic community, PE150, 100 genomes
    "community": dict(n_refs=100, ref_len_min=2_000_000, ref_len_max=6_000_000, abund_sigma=1.0,
                      shared_fraction=0.20, single_fraction=0.05, clip_fraction=0.10, indel_fraction=0.05,
                      unmapped_fraction=0.0, max_occ=8, alt_noise=1, minimal_aux=0),
    # configs[2]: 10k references, 30 % multi-mappers, profile only
    "catalog10k": dict(n_refs=10_000, ref_len_min=1_000, ref_len_max=5_000, abund_sigma=2.0,
                       shared_fraction=0.30, single_fraction=0.05, clip_fraction=0.0, indel_fraction=0.0,
                       unmapped_fraction=0.0, max_occ=8, alt_noise=0, minimal_aux=0),
    # configs[4]: 1M-gene catalog
    "genes1m": dict(n_refs=1_000_000, ref_len_min=400, ref_len_max=1_400, abund_sigma=2.0,
                    shared_fraction=0.30, single_fraction=0.05, clip_fraction=0.10, indel_fraction=0.05,
                    unmapped_fraction=0.0, max_occ=8, alt_noise=1, minimal_aux=0),
    # small mixed case for parity tests (unmapped pairs included)
    "mixed": dict(n_refs=50, ref_len_min=2_000, ref_len_max=9_000, abund_sigma=1.5,
                  shared_fraction=0.35, single_fraction=0.10, clip_fraction=0.25, indel_fraction=0.15,
                  unmapped_fraction=0.08, max_occ=12, alt_noise=1, minimal_aux=0),
}

_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run `make`")
        lib = C.CDLL(LIB_PATH)
        lib.synth_generate.argtypes = [C.POINTER(SynthParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p]
        lib.synth_generate.restype = C.c_int
        lib.synth_target_lengths.argtypes = [C.POINTER(SynthParams), C.c_void_p]
        lib.synth_target_lengths.restype = None
        lib.synth_max_record_bytes.argtypes = [C.POINTER(SynthParams)]
        lib.synth_max_record_bytes.restype = C.c_size_t
        _lib = lib
    return _lib


def make_params(preset="community", n_records=10_000, seed=13579, qname_base=0, read_len=150, insert_len=350, **over):
    d = dict(PRESETS[preset])
    d.update(over)
    p = SynthParams()
    p.seed, p.n_records, p.qname_base = seed, n_records, qname_base
    p.read_len, p.insert_len = read_len, insert_len
    for k, v in d.items():
        setattr(p, k, v)
    return p


def target_lengths(p):
    lib = _load()
    tl = np.zeros(p.n_refs, dtype=np.uint32)
    lib.synth_target_lengths(C.byref(p), tl.ctypes.data_as(C.c_void_p))
    return tl


def generate(p, raw=None, off=None):
    """Returns (raw uint8[nbytes], rec_off uint64[n+1], stats dict).  raw/off may be preallocated
    (e.g. pinned) buffers; the returned arrays are views trimmed to the generated size."""
    lib = _load()
    maxrec = lib.synth_max_record_bytes(C.byref(p))
    slack = 2 * max(p.max_occ, 1)
    if raw is None:
        raw = np.empty((p.n_records + slack) * 330 if p.read_len == 150 else (p.n_records + slack) * maxrec, dtype=np.uint8)
    if off is None:
        off = np.empty(p.n_records + slack + 2, dtype=np.uint64)
    nb, nr = C.c_size_t(), C.c_size_t()
    st = np.zeros(4, dtype=np.uint64)
    rc = lib.synth_generate(C.byref(p), raw.ctypes.data_as(C.c_void_p), raw.nbytes, off.ctypes.data_as(C.c_void_p), len(off),
                            C.byref(nb), C.byref(nr), st.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"synth_generate failed ({rc}): buffers too small")
    stats = dict(inserts=int(st[0]), shared=int(st[1]), unmapped=int(st[2]), single=int(st[3]))
    return raw[:nb.value], off[:nr.value + 1], stats
