"""Loader for the committed golden bundle (tests/golden/fixtures.npz + cases.json)."""
import json
import os

import numpy as np

import samutil

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def bundle():
    if "npz" not in _cache:
        _cache["npz"] = np.load(os.path.join(HERE, "fixtures.npz"))
        with open(os.path.join(HERE, "cases.json")) as fh:
            _cache["cases"] = json.load(fh)
    return _cache["npz"], _cache["cases"]


def fixture(name):
    z, _ = bundle()
    names = [str(x) for x in z[name + ".names"]]
    refs = list(zip(names, [int(x) for x in z[name + ".tlen"]]))
    return samutil.Sam([str(x) for x in z[name + ".header"]], refs, z[name + ".raw"].copy(), z[name + ".off"].copy())


def cases(kind):
    return [c for c in bundle()[1] if c["kind"] == kind]


def case_id(c):
    opts = c.get("opts") or c.get("pre") or {}
    tag = ",".join(f"{k}={v}" for k, v in opts.items())
    return f"{c['fixture']}[{c.get('mode', '')}{tag}]"


def dense(c, key, dtype):
    out = np.zeros(c["n"], dtype=dtype)
    out[np.array(c["nz"], dtype=np.int64)] = np.array(c[key], dtype=dtype)
    return out
