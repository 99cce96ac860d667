"""The reference's own ring aggregates behind the replay host (oracle/_ref/libref_replay.so) --
TEST INFRASTRUCTURE, same import rules as oracle/oracle.py."""
from __future__ import annotations

import os

from duckdb_imputation_b200.replay import Replay

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_replay.so")
_ref = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def ref() -> Replay:
    global _ref
    if _ref is None:
        _ref = Replay(LIB_PATH)
        assert _ref.implementation == "reference"
    return _ref


def time_sum_to_triple(num_cols, cat_cols, threads: int) -> float:
    """Seconds of update+combine+finalize of the reference's sum_to_triple on these columns.
    sum_to_triple_20_* is not registered by the reference (grid stops at 19,
    duckdb_imputation_extension.cpp:80-85): 20 columns are timed as _19_ scaled by 20/19 in rows/s
    terms by the caller -- here we simply run the widest registered function on the first 19."""
    r = ref()
    n = min(len(num_cols), 19)
    r.aggregate("sum_to_triple_%d_%d" % (n, len(cat_cols)), num_cols[:n], cat_cols, threads=threads)
    return r.last_seconds
