"""ctypes front end of the hash-aggregate replay host (csrc/host/replay_host.h).

`Replay(path)` wraps one library built around the replay host:
  * duckdb_imputation_b200/lib/libduckdb_imputation_b200.so -- OUR extension (CUDA behind the
    DuckDB aggregate callbacks);
  * oracle/_ref/libref_replay.so -- the reference's own sources (test infrastructure; loaded
    only through oracle/ref_replay.py).
`query()` is the SQL-shaped entry the parity tests use:  SELECT fn(cols) FROM t [WHERE..] [GROUP BY gb].
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from ._native import ptr_array

_HERE = os.path.dirname(os.path.abspath(__file__))
GLUE_LIB_PATH = os.path.join(_HERE, "lib", "libduckdb_imputation_b200.so")


class ReplayError(RuntimeError):
    pass


class Replay:
    def __init__(self, path: str):
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing (build it: __graft_entry__.build())")
        # RTLD_LOCAL (+ -Bsymbolic at link time): our extension and the reference build define the same
        # Triple::* symbols and may be loaded side by side in one test or bench process
        l = C.CDLL(path, mode=C.RTLD_LOCAL)
        P = C.c_void_p
        l.replay_aggregate.restype = C.c_int
        l.replay_aggregate.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(P), C.POINTER(P), P, C.c_int, P,
                                       C.c_size_t, C.c_size_t, C.c_int, C.POINTER(P), C.POINTER(C.c_double)]
        l.replay_scalar.restype = C.c_int
        l.replay_scalar.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(P), C.POINTER(P), P, C.c_size_t, C.c_size_t,
                                    C.POINTER(P)]
        l.replay_scalar_structs.restype = C.c_int
        l.replay_scalar_structs.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(P)]
        l.replay_predict.restype = C.c_int
        l.replay_predict.argtypes = [C.c_char_p, P, C.c_size_t, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(P),
                                     C.POINTER(P), P, C.c_size_t, C.c_size_t, P]
        l.replay_set_option.restype = C.c_int
        l.replay_set_option.argtypes = [C.c_char_p, C.c_int]
        l.replay_load_via_entry_points.restype = C.c_int
        l.replay_load_via_entry_points.argtypes = [C.POINTER(C.c_char_p)]
        l.replay_train.restype = C.c_int
        l.replay_train.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.c_char_p, C.POINTER(P),
                                   C.POINTER(C.c_size_t)]
        l.replay_train_list.restype = C.c_int
        l.replay_train_list.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_size_t, C.c_int,
                                        C.POINTER(C.c_double), C.c_char_p, C.POINTER(P), C.POINTER(C.c_size_t)]
        l.replay_value_ring.restype = C.c_int
        l.replay_value_ring.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.POINTER(P)]
        l.replay_value_error.restype = C.c_char_p
        l.replay_free.argtypes = [P]
        l.replay_last_error.restype = C.c_char_p
        l.replay_list_functions.restype = P
        l.replay_implementation.restype = C.c_char_p
        self.lib = l
        self.last_seconds = 0.0

    def options(self, **kw):
        """Context manager: shapes of DuckDB's protocol a plain scan does not produce (replay_host.h,
        replay_set_option): no_simple, lift_shape, split_states, parallel_finalize."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            try:
                for k, v in kw.items():
                    if self.lib.replay_set_option(k.encode(), int(v)):
                        raise ReplayError(self.lib.replay_last_error().decode("utf-8", "replace"))
                yield self
            finally:
                self.lib.replay_set_option(b"reset", 0)
        return cm()

    def load_via_entry_points(self):
        """duckdb_imputation_init() on a fresh catalog -> (number of registered functions, version string)."""
        v = C.c_char_p()
        n = self.lib.replay_load_via_entry_points(C.byref(v))
        if n < 0:
            raise ReplayError(self.lib.replay_last_error().decode("utf-8", "replace"))
        return n, (v.value or b"").decode()

    @property
    def implementation(self) -> str:
        return self.lib.replay_implementation().decode()

    def functions(self):
        p = self.lib.replay_list_functions()
        try:
            return C.string_at(p).decode().split()
        finally:
            self.lib.replay_free(p)

    def aggregate(self, function: str, num_cols, cat_cols, group=None, n_groups=1, sel=None, threads=1, scalar=None):
        """Raw call: group = int32 slots; sel = ascending uint32 row ids; scalar = name of a lift to apply
        first (sum_triple(to_cofactor(..))).  -> list of STRUCT dicts."""
        kn = [np.ascontiguousarray(c, np.float32) for c in num_cols]
        kc = [np.ascontiguousarray(c, np.int32) for c in cat_cols]
        rows = len(kn[0]) if kn else (len(kc[0]) if kc else 0)
        g = None if group is None else np.ascontiguousarray(group, np.int32)
        s = None if sel is None else np.ascontiguousarray(sel, np.uint32)
        out = C.c_void_p()
        secs = C.c_double()
        rc = self.lib.replay_aggregate(function.encode(), (scalar or "").encode(), len(kn), len(kc), ptr_array([k.ctypes.data for k in kn]),
                                       ptr_array([k.ctypes.data for k in kc]), None if g is None else g.ctypes.data,
                                       n_groups, None if s is None else s.ctypes.data, 0 if s is None else len(s), rows,
                                       threads, C.byref(out), C.byref(secs))
        if rc:
            raise ReplayError(self.lib.replay_last_error().decode("utf-8", "replace"))
        try:
            self.last_seconds = secs.value
            return json.loads(C.string_at(out).decode())
        finally:
            self.lib.replay_free(out)

    def scalar(self, function: str, num_cols, cat_cols, where=None):
        """SELECT function(cols) FROM t [WHERE ..] -> one value per (selected) row."""
        kn = [np.ascontiguousarray(c, np.float32) for c in num_cols]
        kc = [np.ascontiguousarray(c, np.int32) for c in cat_cols]
        rows = len(kn[0]) if kn else (len(kc[0]) if kc else 0)
        s = None if where is None else np.nonzero(np.asarray(where))[0].astype(np.uint32)
        out = C.c_void_p()
        rc = self.lib.replay_scalar(function.encode(), len(kn), len(kc), ptr_array([k.ctypes.data for k in kn]),
                                    ptr_array([k.ctypes.data for k in kc]), None if s is None else s.ctypes.data,
                                    0 if s is None else len(s), rows, C.byref(out))
        if rc:
            raise ReplayError(self.lib.replay_last_error().decode("utf-8", "replace"))
        try:
            return json.loads(C.string_at(out).decode())
        finally:
            self.lib.replay_free(out)

    def predict(self, function: str, params, flags, num_cols, cat_cols, where=None):
        """SELECT function(params::FLOAT[], flags..., cols...) FROM t [WHERE ..] -> numpy array (float32 for
        linreg_predict, int32 for lda_predict / nb_predict / qda_predict), one value per (selected) row."""
        kn = [np.ascontiguousarray(c, np.float32) for c in num_cols]
        kc = [np.ascontiguousarray(c, np.int32) for c in cat_cols]
        rows = len(kn[0]) if kn else (len(kc[0]) if kc else 0)
        s = None if where is None else np.nonzero(np.asarray(where))[0].astype(np.uint32)
        p = np.ascontiguousarray(params, np.float32)
        out = np.zeros(rows if s is None else len(s), np.float32 if function.startswith("linreg") else np.int32)
        fl = (C.c_int * max(1, len(flags)))(*[int(bool(f)) for f in flags])
        rc = self.lib.replay_predict(function.encode(), p.ctypes.data, len(p), fl, len(flags), len(kn), len(kc),
                                     ptr_array([k.ctypes.data for k in kn]), ptr_array([k.ctypes.data for k in kc]),
                                     None if s is None else s.ctypes.data, 0 if s is None else len(s), rows, out.ctypes.data)
        if rc:
            raise ReplayError(self.lib.replay_last_error().decode("utf-8", "replace"))
        return out

    def scalar_structs(self, function: str, *columns):
        """SELECT function(A, B, ..) FROM j, every argument a column (list) of ring STRUCT dicts, e.g.
        multiply_triple over the rows of a join of aggregate results -> one STRUCT per row."""
        rows = len(columns[0])
        if rows == 0:
            return []
        nb = 0 if "quad_cat" in columns[0][0] else 1
        texts = (C.c_char_p * len(columns))(*[json.dumps(col).encode() for col in columns])
        out = C.c_void_p()
        rc = self.lib.replay_scalar_structs(function.encode(), nb, len(columns), texts, rows, C.byref(out))
        if rc:
            raise ReplayError(self.lib.replay_last_error().decode("utf-8", "replace"))
        try:
            return json.loads(C.string_at(out).decode())
        finally:
            self.lib.replay_free(out)

    def train(self, function: str, triple: dict, *consts):
        """SELECT function(triple, consts...): linreg_train(triple, label, step, lambda, max_iterations, variance,
        normalize) / lda_train(triple, label, shrinkage, normalize) -> the FLOAT[] parameter list (numpy float32).
        Python ints are handed over as INTEGER, floats as FLOAT, bools as BOOLEAN."""
        types = "".join("b" if isinstance(c, bool) else "i" if isinstance(c, (int, np.integer)) else "f" for c in consts)
        vals = (C.c_double * max(1, len(consts)))(*[float(c) for c in consts])
        out, n = C.c_void_p(), C.c_size_t()
        rc = self.lib.replay_train(function.encode(), json.dumps(triple).encode(), len(consts), vals, types.encode(),
                                   C.byref(out), C.byref(n))
        if rc:
            raise ReplayError(self.lib.replay_last_error().decode("utf-8", "replace"))
        try:
            return np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float)), shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
        finally:
            self.lib.replay_free(out)

    def train_list(self, function: str, triples: list, labels, *consts):
        """SELECT function(list(agg), list(label), consts...): qda_train(triples, labels, normalize) / nb_train(triples,
        labels) over per-class ring STRUCTs -> the FLOAT[] parameter list."""
        nb = 0 if (triples and "quad_cat" in triples[0]) else 1
        types = "".join("b" if isinstance(c, bool) else "i" if isinstance(c, (int, np.integer)) else "f" for c in consts)
        vals = (C.c_double * max(1, len(consts)))(*[float(c) for c in consts])
        lab = (C.c_int32 * max(1, len(labels)))(*[int(v) for v in labels])
        out, n = C.c_void_p(), C.c_size_t()
        rc = self.lib.replay_train_list(function.encode(), nb, json.dumps(triples).encode(), lab, len(labels), len(consts), vals,
                                        types.encode(), C.byref(out), C.byref(n))
        if rc:
            raise ReplayError(self.lib.replay_last_error().decode("utf-8", "replace"))
        try:
            return np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_float)), shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
        finally:
            self.lib.replay_free(out)

    VALUE_OPS = {"sum_triple": 0, "subtract_triple": 1, "sum_nb_triple": 2}

    def value_ring(self, op: str, a: dict, b: dict) -> dict:
        """Triple::sum_triple / subtract_triple / sum_nb_triple(a, b) on two ring STRUCT values (dicts; fields are
        read by position) -- the Value-level helpers of the MICE drivers (imputation/include/sum_sub.h:10-14)."""
        out = C.c_void_p()
        rc = self.lib.replay_value_ring(self.VALUE_OPS[op], json.dumps(a).encode(), json.dumps(b).encode(), C.byref(out))
        if rc:
            raise ReplayError(self.lib.replay_value_error().decode("utf-8", "replace"))
        try:
            return json.loads(C.string_at(out).decode())
        finally:
            self.lib.replay_free(out)

    def query(self, kind, num_cols, cat_cols, group_by=None, where=None, threads=1, lifted=False):
        """Same signature as oracle.aggregate / aggregates._aggregate (tests/sqlmini.py backend).
        lifted=True runs sum_triple(to_cofactor(..)) / sum_nb_agg(to_nb_agg(..)) instead."""
        fn = ("sum_to_triple_%d_%d" if kind == 0 else "sum_to_nb_agg_%d_%d") % (len(num_cols), len(cat_cols))
        scalar = None
        if lifted:
            fn, scalar = ("sum_triple", "to_cofactor") if kind == 0 else ("sum_nb_agg", "to_nb_agg")
        sel = None if where is None else np.nonzero(np.asarray(where))[0].astype(np.uint32)
        if group_by is None:
            return self.aggregate(fn, num_cols, cat_cols, sel=sel, threads=threads, scalar=scalar)[0]
        gb = np.asarray(group_by)
        labels = np.unique(gb)
        slots = np.searchsorted(labels, gb).astype(np.int32)
        return self.aggregate(fn, num_cols, cat_cols, group=slots, n_groups=max(1, len(labels)), sel=sel, threads=threads,
                              scalar=scalar)


_glue = None


def glue() -> Replay:
    """Our extension behind the replay host (needs a CUDA device to run queries)."""
    global _glue
    if _glue is None:
        from . import _native
        _native.lib()  # libcofactor_b200.so first: the glue links against it
        _glue = Replay(GLUE_LIB_PATH)
    return _glue
