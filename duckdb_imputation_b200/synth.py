"""Host twin of the device generators (cfb_gen_uniform_f32 / cfb_gen_int32, scatter_kernels.cuh).

Element i of a stream is a pure function of (seed, first + i): any slice of a synthetic
device column can be regenerated here bit-for-bit to feed the CPU oracle.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _counter(seed: int, first: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + np.uint64(first)
        return base + np.arange(n, dtype=np.uint64)


def uniform_f32(n: int, seed: int, first: int = 0) -> np.ndarray:
    h = _mix64(_counter(seed, first, n))
    return ((h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def int32(n: int, seed: int, first: int = 0, lo: int = 0, rng: int = 100) -> np.ndarray:
    h = _mix64(_counter(seed, first, n))
    return (np.int64(lo) + ((h >> np.uint64(33)) % np.uint64(rng)).astype(np.int64)).astype(np.int32)


def column_seed(table_seed: int, col: int) -> int:
    """Seed of column `col` of the synthetic table with seed `table_seed`."""
    return table_seed * 1000 + col
