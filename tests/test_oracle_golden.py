"""CPU suite: the oracle restatement against every golden STRUCT of the reference's sum tests
(test_sum.py:22-52, test_nb_sum.py), in both arithmetic modes, plus internal consistency."""
import numpy as np
import pytest

from oracle import oracle
from tests import sqlmini

SUM_FILES = ("test_sum.py", "test_nb_sum.py")


def _sum_cases(goldens):
    return [c for c in goldens["cases"] if c["file"] in SUM_FILES]


@pytest.mark.parametrize("mode", [oracle.EXACT, oracle.FAITHFUL])
def test_oracle_matches_reference_goldens(goldens, mode):
    cases = _sum_cases(goldens)
    assert len(cases) == 8
    for c in cases:
        got = sqlmini.run_sum(c["sql"], goldens["fixtures"][c["file"]],
                              lambda *a, **k: oracle.aggregate(*a, mode=mode, **k))
        assert got[c["index"]] == c["expected"], (c["file"], c["test"], c["index"])


def test_sum_equals_sum_of_lifted(goldens):
    """test_sum.py:40-52 / test_nb_sum.py: sum_to_triple == sum_triple(to_cofactor), grouped."""
    for fn, kind in (("test_sum.py", oracle.TRIPLE), ("test_nb_sum.py", oracle.NB)):
        cols, types = sqlmini.table(goldens["fixtures"][fn])
        num = [cols[c] for c in "abc"]
        cat = [cols[c] for c in "def"]
        labels, slots = np.unique(cols["gb"], return_inverse=True)
        a = oracle.aggregate(kind, num, cat, group_by=cols["gb"], mode=oracle.FAITHFUL)
        b = oracle.sum_of_lifted(kind, num, cat, group=slots.astype(np.int32), n_groups=len(labels))
        assert a == b


@pytest.mark.parametrize("kind", [oracle.TRIPLE, oracle.NB])
def test_faithful_threads_and_chunks_agree_with_exact(kind):
    rng = np.random.default_rng(7)
    rows = 10_000  # > 2048: several update calls per thread, and a combine
    num = [rng.integers(0, 8, rows).astype(np.float32) for _ in range(4)]  # small ints: fp32 exact
    cat = [rng.integers(-3, 5, rows).astype(np.int32) for _ in range(3)]
    gb = rng.integers(0, 3, rows).astype(np.int32)
    ex = oracle.aggregate(kind, num, cat, group_by=gb, mode=oracle.EXACT)
    for t in (1, 3):
        fa = oracle.aggregate(kind, num, cat, group_by=gb, mode=oracle.FAITHFUL, threads=t)
        assert fa == ex


def test_filtered_scan_equals_materialised_filter():
    rng = np.random.default_rng(8)
    rows = 5000
    num = [rng.random(rows).astype(np.float32) for _ in range(3)]
    cat = [rng.integers(0, 10, rows).astype(np.int32) for _ in range(2)]
    keep = rng.random(rows) < 0.8
    a = oracle.aggregate(oracle.TRIPLE, num, cat, where=keep)
    b = oracle.aggregate(oracle.TRIPLE, [c[keep] for c in num], [c[keep] for c in cat])
    assert a == b


def test_empty_and_shapes():
    e = oracle.aggregate(oracle.TRIPLE, [np.zeros(0, np.float32)] * 2, [np.zeros(0, np.int32)])
    assert e["N"] == 0 and e["lin_agg"] == [0.0, 0.0] and e["lin_cat"] == [[]]
    # numeric only: categorical lists are empty (test_lift.py:45 shape)
    r = oracle.aggregate(oracle.TRIPLE, [np.ones(5, np.float32)], [])
    assert r == {"N": 5, "lin_agg": [5.0], "quad_agg": [5.0], "lin_cat": [], "quad_num_cat": [], "quad_cat": []}
    # categorical only
    r = oracle.aggregate(oracle.TRIPLE, [], [np.array([2, 2, 7], np.int32)])
    assert r["lin_cat"] == [[{"key": 2, "value": 2.0}, {"key": 7, "value": 1.0}]]
    assert r["quad_cat"] == [[{"key1": 2, "key2": 2, "value": 2.0}, {"key1": 7, "key2": 7, "value": 1.0}]]


def test_result_add_is_combine():
    rng = np.random.default_rng(9)
    rows = 3000
    num = [rng.integers(0, 5, rows).astype(np.float32) for _ in range(2)]
    cat = [rng.integers(0, 6, rows).astype(np.int32) for _ in range(2)]
    whole = oracle.aggregate(oracle.TRIPLE, num, cat)
    h = rows // 3
    import ctypes as C
    from duckdb_imputation_b200._native import Result
    from duckdb_imputation_b200.struct_result import result_to_struct
    # add through the C entry point on raw results
    l = oracle.lib()

    def raw(nc, cc):
        from duckdb_imputation_b200._native import ptr_array
        out = Result()
        kn = [np.ascontiguousarray(c) for c in nc]
        kc = [np.ascontiguousarray(c) for c in cc]
        rc = l.orc_aggregate(0, 0, len(kn), len(kc), ptr_array([k.ctypes.data for k in kn]),
                             ptr_array([k.ctypes.data for k in kc]), None, 1, None, len(kn[0]), len(kn[0]), 1, C.byref(out))
        assert rc == 0
        return out

    a = raw([c[:h] for c in num], [c[:h] for c in cat])
    b = raw([c[h:] for c in num], [c[h:] for c in cat])
    s = Result()
    assert l.orc_result_add(C.byref(a), C.byref(b), C.byref(s)) == 0
    assert result_to_struct(s) == whole
    for r in (a, b, s):
        l.orc_result_free(C.byref(r))
