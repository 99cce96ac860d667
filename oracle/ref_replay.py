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
    """Seconds of update + combine + finalize of the reference's own sum_to_triple callbacks on
    these columns under DuckDB's protocol (T threads, 2048-row chunks, per-row state pointers).
    sum_to_triple_20_0 is outside the reference's registration grid (it stops at 19,
    duckdb_imputation_extension.cpp:80-85): ref_extension.cpp exposes the same unmodified
    callbacks for 20 FLOAT columns as ref_sum_to_triple_20_0."""
    r = ref()
    n, m = len(num_cols), len(cat_cols)
    name = "ref_sum_to_triple_20_0" if (n, m) == (20, 0) else "sum_to_triple_%d_%d" % (n, m)
    r.aggregate(name, num_cols, cat_cols, threads=threads)
    return r.last_seconds


class _quiet_stdout:
    """The reference's predict functions and trainers print debug lines to std::cout / std::cerr (regression.cpp:440,
    :466, lda.cpp:157-330, ML/utils.cpp:7): file descriptors 1 and 2 point at /dev/null while they run."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        sys.stderr.flush()
        self._saved = [os.dup(1), os.dup(2)]
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)
        os.dup2(self._null, 2)

    def __exit__(self, *a):
        import ctypes
        ctypes.CDLL(None).fflush(None)
        os.dup2(self._saved[0], 1)
        os.dup2(self._saved[1], 2)
        for fd in self._saved + [self._null]:
            os.close(fd)


def predict(function: str, params, flags, num_cols, cat_cols, where=None):
    """The reference's own linreg_predict / lda_predict / nb_predict (ML::linreg_impute, LDA_impute, ML::nb_impute compiled
    from /root/reference into oracle/_ref) on these columns: one value per (selected) row."""
    with _quiet_stdout():
        return ref().predict(function, params, flags, num_cols, cat_cols, where=where)


def seed_libc_random(seed: int):
    """linreg_predict(noise = true) draws from libc random() (regression.cpp:495-505)."""
    import ctypes
    ctypes.CDLL(None).srandom(ctypes.c_uint(seed))


def lapack_available() -> bool:
    """lda_train calls dgelsd / dgemm; oracle/ref_ml_stubs.cpp forwards them to the OpenBLAS inside scipy's wheel."""
    return _lapack_path() is not None


def _lapack_path():
    import glob
    try:
        import scipy
    except ImportError:
        return None
    libs = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
    return libs[0] if libs else None


def train(function: str, triple: dict, *consts):
    """The reference's own linreg_train / lda_train (ML::ridge_linear_regression, lda_train compiled from /root/reference
    into oracle/_ref) on a ring STRUCT -> FLOAT[] parameter list."""
    path = _lapack_path()
    if path:
        os.environ.setdefault("CFB_REF_LAPACK", path)
    with _quiet_stdout():
        return ref().train(function, triple, *consts)


def train_list(function: str, triples: list, labels, *consts):
    """The reference's own qda_train / nb_train on per-class ring STRUCTs (GROUP BY label) -> FLOAT[] parameter list."""
    path = _lapack_path()
    if path:
        os.environ.setdefault("CFB_REF_LAPACK", path)
    with _quiet_stdout():
        return ref().train_list(function, triples, labels, *consts)
