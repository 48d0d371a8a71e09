"""QNAME-boundary sharding of a record stream across ranks: every rank gets a
contiguous run of whole QNAME groups, so no read straddles GPUs and the only cross-rank state is
additive (count vectors, EM increments, coverage diff arrays)."""
import numpy as np

from .api import split_point


def shard_bounds(raw, rec_off, world_size):
    """Record-index cut points [0, c1, ..., n]: cut k is the msg_split_point at or below n*k/world."""
    n = len(rec_off) - 1
    cuts = [0]
    for k in range(1, world_size):
        c = split_point(raw, rec_off, n * k // world_size)
        cuts.append(max(c, cuts[-1]))
    cuts.append(n)
    return cuts


def shard_view(raw, rec_off, cuts, rank):
    """(raw bytes, offsets rebased to 0) of one rank's shard; views, no copy of the record bytes."""
    a, b = cuts[rank], cuts[rank + 1]
    lo, hi = int(rec_off[a]), int(rec_off[b])
    return raw[lo:hi], (np.asarray(rec_off[a:b + 1]) - rec_off[a]).astype(np.uint64)
