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


def keep_mask(seed, rows, ncols, p):
    """bool [len(rows), ncols]: keep decisions for (row id, column) exactly as dropout_keep()."""
    rows = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
    col4 = (np.arange(ncols, dtype=np.uint64) >> np.uint64(2)).reshape(1, -1)
    k = (np.arange(ncols, dtype=np.uint64) & np.uint64(3)).reshape(1, -1)
    with np.errstate(over="ignore"):
        rowkey = splitmix64(np.uint64(seed) + rows * np.uint64(0xD6E8FEB86659FD93))
        bits = splitmix64(rowkey ^ (col4 * np.uint64(0xA0761D6478BD642F)))
    lane = (bits >> (np.uint64(16) * k)) & np.uint64(0xFFFF)
    return lane >= np.uint64(dropout_threshold(p))
