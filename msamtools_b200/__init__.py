"""msamtools_b200 -- B200-native hot path of msamtools (filter / best-hit / profile / coverage).

The compute path is hand-written CUDA for sm_100a in libmsamtools_b200.so behind the C ABI
in include/msamtools_b200.h; importing this package does not touch the GPU, but creating a
Context does and fails loudly when the library or a device is missing (no CPU fallback).
"""
from . import _lib
from .api import Context, MsgError, PinnedBuffer, index_records, split_point, nccl_unique_id, parse_multi

__all__ = ["Context", "MsgError", "PinnedBuffer", "index_records", "split_point", "nccl_unique_id", "parse_multi", "_lib"]
