"""ctypes wrapper of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from duckdb_imputation_b200._native import Result, ptr_array
from duckdb_imputation_b200.struct_result import arrays_to_struct, result_arrays

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
TRIPLE, NB = 0, 1
EXACT, FAITHFUL = 0, 1
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        P = C.c_void_p
        l.orc_aggregate.restype = C.c_int
        l.orc_aggregate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(P), C.POINTER(P), P, C.c_int, P,
                                    C.c_size_t, C.c_size_t, C.c_int, C.POINTER(Result)]
        l.orc_sum_of_lifted.restype = C.c_int
        l.orc_sum_of_lifted.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(P), C.POINTER(P), P, C.c_int,
                                        C.c_size_t, C.POINTER(Result)]
        l.orc_result_add.restype = C.c_int
        l.orc_result_add.argtypes = [C.POINTER(Result)] * 3
        l.orc_result_free.restype = None
        l.orc_result_free.argtypes = [C.POINTER(Result)]
        l.orc_last_seconds.restype = C.c_double
        _lib = l
    return _lib


def _cols(cols, dt):
    keep = [np.ascontiguousarray(c, dtype=dt) for c in cols]
    return keep, ptr_array([k.ctypes.data for k in keep])


def aggregate_arrays(kind, num_cols, cat_cols, group=None, n_groups=1, sel=None, mode=EXACT, threads=1):
    """-> list (one per group) of numpy-form results (see struct_result.result_arrays)."""
    kn, pn = _cols(num_cols, np.float32)
    kc, pc = _cols(cat_cols, np.int32)
    g = None if group is None else np.ascontiguousarray(group, np.int32)
    s = None if sel is None else np.ascontiguousarray(sel, np.uint32)
    table_rows = len(kn[0]) if kn else (len(kc[0]) if kc else 0)
    rows = len(s) if s is not None else table_rows
    out = (Result * n_groups)()
    rc = lib().orc_aggregate(kind, mode, len(kn), len(kc), pn, pc, None if g is None else g.ctypes.data, n_groups,
                             None if s is None else s.ctypes.data, rows, table_rows, threads, out)
    if rc:
        raise ValueError("orc_aggregate: bad arguments")
    try:
        return [result_arrays(out[i]) for i in range(n_groups)]
    finally:
        for i in range(n_groups):
            lib().orc_result_free(C.byref(out[i]))


def aggregate(kind, num_cols, cat_cols, group_by=None, where=None, mode=EXACT, threads=1, narrow=True):
    """SQL-shaped front end: returns a STRUCT dict, or a list of them in ascending group-key order."""
    sel = None if where is None else np.nonzero(np.asarray(where))[0].astype(np.uint32)
    if group_by is None:
        return arrays_to_struct(aggregate_arrays(kind, num_cols, cat_cols, sel=sel, mode=mode, threads=threads)[0], narrow)
    gb = np.asarray(group_by)
    used = gb[sel] if sel is not None else gb
    labels = np.unique(used)
    slots = np.searchsorted(labels, gb).clip(0, max(0, len(labels) - 1)).astype(np.int32)
    res = aggregate_arrays(kind, num_cols, cat_cols, group=slots, n_groups=max(1, len(labels)), sel=sel, mode=mode,
                           threads=threads)
    return [arrays_to_struct(r, narrow) for r in res]


def sum_of_lifted(kind, num_cols, cat_cols, group=None, n_groups=1):
    kn, pn = _cols(num_cols, np.float32)
    kc, pc = _cols(cat_cols, np.int32)
    g = None if group is None else np.ascontiguousarray(group, np.int32)
    rows = len(kn[0]) if kn else (len(kc[0]) if kc else 0)
    out = (Result * n_groups)()
    rc = lib().orc_sum_of_lifted(kind, len(kn), len(kc), pn, pc, None if g is None else g.ctypes.data, n_groups, rows, out)
    if rc:
        raise ValueError("orc_sum_of_lifted: bad arguments")
    try:
        return [arrays_to_struct(result_arrays(out[i])) for i in range(n_groups)]
    finally:
        for i in range(n_groups):
            lib().orc_result_free(C.byref(out[i]))


def last_seconds() -> float:
    return float(lib().orc_last_seconds())


def multiply(a: dict, b: dict) -> dict:
    """Ring product of two result STRUCTs -- restatement of Triple::MultiplyFunction
    (duckdb_extension/src/triple/mul.cpp:17-611) and Triple::multiply_nb (mul_nb.cpp) on the
    Python form of the STRUCT (pure Python: the operands are per-group results, tens of values).
    fp32 arithmetic like the reference (float * int, float * float)."""
    f = np.float32
    la, lb = a.get("lin_agg", a.get("lin_num")), b.get("lin_agg", b.get("lin_num"))
    qa, qb = a.get("quad_agg", a.get("quad_num")), b.get("quad_agg", b.get("quad_num"))
    na, nb_, Na, Nb = len(la), len(lb), a["N"], b["N"]
    is_nb = "quad_cat" not in a
    out = {"N": Na * Nb}                                                           # mul.cpp:42-47
    out["lin_num"] = [float(f(x) * f(Nb)) for x in la] + [float(f(x) * f(Na)) for x in lb]   # :96-107
    if is_nb:                                                                      # mul_nb.cpp: diagonals only
        out["quad_num"] = [float(f(x) * f(Nb)) for x in qa] + [float(f(x) * f(Na)) for x in qb]
    else:
        quad, p = [], 0
        for j in range(na):                                                        # mul.cpp:262-283
            for _ in range(j, na):
                quad.append(float(f(qa[p]) * f(Nb)))
                p += 1
            quad += [float(f(la[j]) * f(x)) for x in lb]
        quad += [float(f(x) * f(Na)) for x in qb]                                  # :285-288
        out["quad_num"] = quad
    scale = lambda lst, N: [{**e, "value": float(f(e["value"]) * f(N))} for e in lst]
    lca, lcb = a["lin_cat"], b["lin_cat"]
    ma, mb = len(lca), len(lcb)
    out["lin_cat"] = [scale(l, Nb) for l in lca] + [scale(l, Na) for l in lcb]      # :184-217
    if is_nb:
        return out
    nca, ncb = a["quad_num_cat"], b["quad_num_cat"]
    nc = []
    for j in range(na):                                                            # :377-411
        nc += [scale(nca[j * ma + k], Nb) for k in range(ma)]
        nc += [[{"key": e["key"], "value": float(f(la[j]) * f(e["value"]))} for e in lcb[k]] for k in range(mb)]
    for j in range(nb_):                                                           # :414-445
        nc += [[{"key": e["key"], "value": float(f(lb[j]) * f(e["value"]))} for e in lca[k]] for k in range(ma)]
        nc += [scale(ncb[j * mb + k], Na) for k in range(mb)]
    out["quad_num_cat"] = nc
    cca, ccb = a["quad_cat"], b["quad_cat"]
    cc, p = [], 0
    for c1 in range(ma):                                                           # :546-581
        for _ in range(c1, ma):
            cc.append(scale(cca[p], Nb))
            p += 1
        for c2 in range(mb):
            cc.append([{"key1": x["key"], "key2": y["key"], "value": float(f(x["value"]) * f(y["value"]))}
                       for x in lca[c1] for y in lcb[c2]])
    cc += [scale(l, Na) for l in ccb]                                              # :583-597
    out["quad_cat"] = cc
    return out
