"""Flat canonical result (cfb_result) -> the STRUCT the DuckDB aggregates return.

Mirrors what Triple::SumStateFinalize writes (sum_state.cpp:116-464) as the Python value the
duckdb client would hand back: a dict with N, lin_agg, quad_agg, lin_cat[, quad_num_cat,
quad_cat].  Sums and counts are narrowed to FLOAT (the STRUCT's declared type,
sum_no_lift.cpp:21-44) unless narrow=False.
"""
from __future__ import annotations

import numpy as np


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def result_arrays(res) -> dict:
    """Copy a cfb_result / orc_result into numpy arrays (exact: float64 / int64)."""
    n, m = res.n_num, res.n_cat
    tk = res.total_keys
    out = {
        "kind": res.kind, "n": n, "m": m, "N": int(res.N),
        "lin": _arr(res.lin, n, np.float64),
        "quad": _arr(res.quad, res.n_quad, np.float64),
        "cat_offsets": _arr(res.cat_offsets, m + 1, np.int64),
        "cat_keys": _arr(res.cat_keys, tk, np.int32),
        "cat_counts": _arr(res.cat_counts, tk, np.int64),
    }
    if res.kind == 0:
        out["numcat"] = _arr(res.numcat_sums, n * tk, np.float64).reshape(n, tk)
        npl = res.n_pair_lists
        out["pair_offsets"] = _arr(res.pair_offsets, npl + 1, np.int64)
        tp = int(out["pair_offsets"][-1]) if npl else 0
        out["pair_key1"] = _arr(res.pair_key1, tp, np.int32)
        out["pair_key2"] = _arr(res.pair_key2, tp, np.int32)
        out["pair_counts"] = _arr(res.pair_counts, tp, np.int64)
    return out


def arrays_to_struct(a: dict, narrow: bool = True) -> dict:
    """numpy form -> the STRUCT dict (field order and list order of sum_state.cpp:132-461)."""
    f = (lambda v: float(np.float32(v))) if narrow else float
    n, m = a["n"], a["m"]
    offs = a["cat_offsets"]
    s = {
        "N": a["N"],
        "lin_agg": [f(v) for v in a["lin"]],
        "quad_agg": [f(v) for v in a["quad"]],
        "lin_cat": [[{"key": int(a["cat_keys"][t]), "value": f(a["cat_counts"][t])}
                     for t in range(offs[c], offs[c + 1])] for c in range(m)],
    }
    if a["kind"] == 0:
        # sub-list index = num * m + cat (sum_state.cpp:383-404)
        s["quad_num_cat"] = [[{"key": int(a["cat_keys"][t]), "value": f(a["numcat"][i, t])}
                              for t in range(offs[c], offs[c + 1])] for i in range(n) for c in range(m)]
        po = a["pair_offsets"]
        s["quad_cat"] = [[{"key1": int(a["pair_key1"][t]), "key2": int(a["pair_key2"][t]),
                           "value": f(a["pair_counts"][t])} for t in range(po[p], po[p + 1])]
                         for p in range(len(po) - 1)]
    return s


def result_to_struct(res, narrow: bool = True) -> dict:
    return arrays_to_struct(result_arrays(res), narrow)
