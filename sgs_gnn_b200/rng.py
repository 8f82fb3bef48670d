"""Host (numpy) mirror of the counter-based hash RNG in csrc/common.cuh, so parity tests can
materialise the exact dropout keep-masks the kernels use and inject them into the CPU oracle."""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def dropout_threshold(p):
    t = np.float32(p) * np.float32(65536.0) + np.float32(0.5)
    return 0 if t <= 0 else (65536 if t >= 65536 else int(t))


def _pair_hash(rowkey, colpair):
    with np.errstate(over="ignore"):
        mix = ((colpair >> np.uint32(4)) * np.uint32(0x9E3779B9)) ^ ((colpair & np.uint32(15)) * np.uint32(0x7FEB352D))
        x = rowkey ^ mix
        lo = (x * np.uint32(0x85EBCA6B)) >> np.uint32(16)
        hi = (x * np.uint32(0xC2B2AE35)) & np.uint32(0xFFFF0000)
    return lo | hi


def keep_mask(seed, rows, ncols, p):
    """bool [len(rows), ncols]: keep decisions for (row id, column) exactly as the kernels compute them
    (dropout_rowkey / dropout_pair in csrc/common.cuh)."""
    rows = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
    with np.errstate(over="ignore"):
        rowkey = (splitmix64(np.uint64(seed) + rows * np.uint64(0xD6E8FEB86659FD93)) >> np.uint64(32)).astype(np.uint32)
    cols = np.arange(ncols, dtype=np.uint32).reshape(1, -1)
    h = _pair_hash(rowkey, cols >> np.uint32(1))
    lane = np.where((cols & np.uint32(1)) == 0, h & np.uint32(0xFFFF), h >> np.uint32(16))
    return lane >= np.uint32(dropout_threshold(p))
