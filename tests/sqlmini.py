"""Just enough SQL to replay the reference's test statements against a backend.

    SELECT <fn>(<cols>) from test [where gb = K] [GROUP BY gb [HAVING gb = K]]

`backend(kind, num_cols, cat_cols, group_by=None, where=None)` returns one STRUCT dict or a
list of them in ascending group order -- the signature shared by oracle.aggregate and
duckdb_imputation_b200.aggregates.
"""
import re

import numpy as np


def table(fixture):
    cols = {}
    rows = fixture["rows"]
    for i, (name, ty) in enumerate(zip(fixture["columns"], fixture["types"])):
        dt = np.float32 if ty == "FLOAT" else np.int32
        cols[name] = np.array([r[i] for r in rows], dtype=dt)
    return cols, dict(zip(fixture["columns"], fixture["types"]))


def parse(sql):
    m = re.match(r"SELECT (\w+)\((.*?)\) from test\s*(.*)$", sql.strip().rstrip(";"), re.I)
    fn, args, tail = m.group(1), [a.strip() for a in m.group(2).split(",")], m.group(3)
    where = re.search(r"where gb = (\d+)", tail, re.I)
    having = re.search(r"HAVING gb = (\d+)", tail, re.I)
    return {"fn": fn, "args": args, "group": bool(re.search(r"GROUP BY gb", tail, re.I)),
            "where": int(where.group(1)) if where else None, "having": int(having.group(1)) if having else None}


def run_sum(sql, fixture, backend):
    """Execute a sum_to_triple_x_y / sum_to_nb_agg_x_y statement -> list of result STRUCTs."""
    q = parse(sql)
    cols, types = table(fixture)
    kind = 0 if q["fn"].startswith("sum_to_triple") else 1
    num = [cols[a] for a in q["args"] if types[a] == "FLOAT"]
    cat = [cols[a] for a in q["args"] if types[a] != "FLOAT"]
    where = None if q["where"] is None else (cols["gb"] == q["where"])
    if not q["group"]:
        return [backend(kind, num, cat, where=where)]
    res = backend(kind, num, cat, group_by=cols["gb"], where=where)
    labels = np.unique(cols["gb"] if where is None else cols["gb"][where])
    if q["having"] is not None:
        return [r for r, l in zip(res, labels) if l == q["having"]]
    return res


def run_lift(sql, fixture, scalar_backend):
    """Execute a to_cofactor / to_nb_agg statement -> list of per-row STRUCTs.
    scalar_backend(function_name, num_cols, cat_cols, where=None)."""
    q = parse(sql)
    cols, types = table(fixture)
    num, cat = [], []
    for a in q["args"]:
        if "+" in a:  # expression argument, e.g. a+b+c (test_lift.py:58-63): a FLOAT column
            num.append(sum(cols[t.strip()] for t in a.split("+")).astype(np.float32))
        elif types[a] == "FLOAT":
            num.append(cols[a])
        else:
            cat.append(cols[a])
    where = None if q["where"] is None else (cols["gb"] == q["where"])
    return scalar_backend(q["fn"], num, cat, where=where)


def run_mul(sql, fixture, backend, multiply, reference_layout=False):
    """Execute the reference's multiply statements (test_mul.py, test_nb_mul.py):
        SELECT multiply_x(A, B) FROM (SELECT [gb,] fn(..) AS A FROM test ..) INNER JOIN (SELECT [gb,] fn(..) AS B ..)
               ON TRUE | on a.gb = b.gb
    `backend` as in run_sum; multiply(list_of_A_rows, list_of_B_rows) -> list of product STRUCTs.
    Row order of the cross join as DuckDB 0.9.2 emits it (and the goldens index it): B-major, one chunk
    per B row holding all A rows next to a CONSTANT B vector.

    reference_layout=True reproduces what the reference computes on that chunk: mul.cpp reads B's N through
    the vector's selection (:42-47) but indexes B's list children by the row's position in the chunk
    (`k + i * num_attr_size_2`, :96-107, :262-288; `col + i * cat_attr_size_2`, :186-217, ...) instead of
    through the list offsets, so row i of a chunk gets N of the chunk's B row and the LISTS of B row i.
    Two of the four cross-join goldens (index 1 and 2) are such mixtures; the equi-join goldens and
    index 0 / 3 are true products."""
    cols, _ = table(fixture)
    subs = re.findall(r"\(SELECT (?:gb as gb, )?(\w+\(.*?\)) AS [AB] FROM test\s*(.*?)\)", sql, re.I)
    assert len(subs) == 2, sql
    sides = []
    for call, tail in subs:
        res = run_sum("SELECT %s from test %s" % (call, tail), fixture, backend)
        where = re.search(r"where gb = (\d+)", tail, re.I)
        gb = cols["gb"] if where is None else cols["gb"][cols["gb"] == int(where.group(1))]
        labels = list(np.unique(gb)) if re.search(r"GROUP BY", tail, re.I) else [None]
        sides.append(list(zip(labels, res)))
    if re.search(r"on a\.gb = b\.gb", sql, re.I):
        pairs = [(ra, rb) for la, ra in sides[0] for lb, rb in sides[1] if la == lb]
    else:
        B = [rb for _, rb in sides[1]]
        pairs = [(ra, {**B[i], "N": rb["N"]} if reference_layout else rb) for rb in B for i, (_, ra) in enumerate(sides[0])]
    return multiply([p[0] for p in pairs], [p[1] for p in pairs])
