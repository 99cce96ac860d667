"""The SQL surface of the hot path as Python calls (test harness; mirrors
duckdb_imputation_extension.cpp:80-113 / :146-179 registration names).

    sum_to_triple(num_cols, cat_cols)                  ~ SELECT sum_to_triple_n_m(...) FROM t
    sum_to_triple(num_cols, cat_cols, group_by=gb)     ~ ... GROUP BY gb   (results in ascending gb)
    sum_to_nb_agg(...)                                 ~ sum_to_nb_agg_n_m

Input is host data (numpy) fed through cfb_ctx_append in DuckDB-sized chunks (2048 rows), i.e.
the same call sequence the DuckDB glue issues; results are the STRUCT dicts the duckdb client
would return.
"""
from __future__ import annotations

import numpy as np

from ._native import CFB_NB, CFB_TRIPLE
from .context import CofactorContext

STANDARD_VECTOR_SIZE = 2048


def _aggregate(kind, num_cols, cat_cols, group_by=None, where=None, chunk=STANDARD_VECTOR_SIZE, device=0,
               narrow=True):
    num_cols = [np.ascontiguousarray(c, np.float32) for c in num_cols]
    cat_cols = [np.ascontiguousarray(c, np.int32) for c in cat_cols]
    rows = len(num_cols[0]) if num_cols else (len(cat_cols[0]) if cat_cols else 0)
    sel = None
    if where is not None:
        sel = np.nonzero(np.asarray(where))[0].astype(np.uint32)
        rows = len(sel)
    labels, slots = None, None
    n_groups = 1
    if group_by is not None:
        gb = np.asarray(group_by)
        used = gb[sel] if sel is not None else gb
        labels, inv = np.unique(used, return_inverse=True)
        n_groups = max(1, len(labels))
        slots = inv.astype(np.uint32)
    with CofactorContext(kind, len(num_cols), len(cat_cols), n_groups, device) as ctx:
        for lo in range(0, rows, chunk):
            hi = min(rows, lo + chunk)
            if sel is None:
                ctx.append([c[lo:hi] for c in num_cols], [c[lo:hi] for c in cat_cols],
                           group=None if slots is None else slots[lo:hi], count=hi - lo)
            else:
                s = sel[lo:hi]
                ctx.append(num_cols, cat_cols, group=None if slots is None else slots[lo:hi],
                           num_sel=[s] * len(num_cols), cat_sel=[s] * len(cat_cols), count=hi - lo)
        if group_by is None:
            return ctx.finalize(0, narrow)
        return [ctx.finalize(g, narrow) for g in range(len(labels))]


def sum_to_triple(num_cols, cat_cols=(), group_by=None, where=None, **kw):
    return _aggregate(CFB_TRIPLE, list(num_cols), list(cat_cols), group_by, where, **kw)


def sum_to_nb_agg(num_cols, cat_cols=(), group_by=None, where=None, **kw):
    return _aggregate(CFB_NB, list(num_cols), list(cat_cols), group_by, where, **kw)
