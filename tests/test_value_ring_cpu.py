"""CPU suite: the Value-level ring helpers of the MICE drivers -- Triple::sum_triple / subtract_triple / sum_nb_triple
(imputation/include/sum_sub.h:10-14) -- mirrored by host/value_glue.cpp over cfb_result_combine and compared, value
for value, with the REFERENCE's own imputation/triple/{sum,sub,sum_nb}.cpp compiled into oracle/_ref (same driver,
host/value_replay.cpp, linked with either).  Host functions on small results: no GPU involved on either side."""
import numpy as np
import pytest

from duckdb_imputation_b200 import replay
from duckdb_imputation_b200.struct_result import arrays_to_struct
from oracle import oracle, ref_replay

needs_ref = pytest.mark.skipif(not ref_replay.available(), reason="oracle/_ref not built")


def _value(kind, num, cat):
    """A ring value as the aggregates emit it (FLOAT-narrowed), in the Value-level field order."""
    s = arrays_to_struct(oracle.aggregate_arrays(kind, num, cat)[0])
    names = ["N", "lin_num", "quad_num", "lin_cat", "quad_num_cat", "quad_cat"]
    return {names[i]: v for i, v in enumerate(s.values())}


def _table(rng, rows, n, m, lo, hi):
    return ([(rng.random(rows) * 8 - 4).astype(np.float32) for _ in range(n)],
            [rng.integers(lo, hi, rows).astype(np.int32) for _ in range(m)])


def _same(got, want):
    """Bit-identical as FLOAT / INTEGER values (the JSON carries doubles)."""
    if isinstance(want, dict):
        assert list(got.keys()) == list(want.keys())
        for k in want:
            _same(got[k], want[k])
    elif isinstance(want, list):
        assert len(got) == len(want)
        for g, w in zip(got, want):
            _same(g, w)
    else:
        assert np.float32(got).tobytes() == np.float32(want).tobytes(), (got, want)


@needs_ref
@pytest.mark.parametrize("n,m", [(3, 2), (0, 2), (4, 0), (1, 1), (5, 3)])
def test_sum_triple_matches_the_reference(n, m):
    rng = np.random.default_rng(100 * n + m)
    a = _value(oracle.TRIPLE, *_table(rng, 500, n, m, 0, 6))
    b = _value(oracle.TRIPLE, *_table(rng, 300, n, m, 3, 11))  # keys the first operand lacks
    want = ref_replay.ref().value_ring("sum_triple", a, b)
    got = replay.glue().value_ring("sum_triple", a, b)
    _same(got, want)
    assert got["N"] == 800


@needs_ref
@pytest.mark.parametrize("n,m", [(3, 2), (0, 2), (4, 0), (2, 1), (5, 3)])
def test_subtract_triple_matches_the_reference(n, m):
    """full - delta, the delta's rows a subset of full's (imputation_low.cpp:85-110): keys whose count becomes 0 stay
    in the lists with value 0, as the reference's std::map merge leaves them (sub.cpp:14-38)."""
    rng = np.random.default_rng(7 * n + m)
    num, cat = _table(rng, 900, n, m, 0, 5)
    for c in cat:
        c[600:] += 5  # the last 300 rows hold keys of their own: full - delta empties them
    full = _value(oracle.TRIPLE, num, cat)
    delta = _value(oracle.TRIPLE, [c[600:] for c in num], [c[600:] for c in cat])
    want = ref_replay.ref().value_ring("subtract_triple", full, delta)
    got = replay.glue().value_ring("subtract_triple", full, delta)
    _same(got, want)
    if m:
        assert [e["key"] for e in got["lin_cat"][0]] == [e["key"] for e in full["lin_cat"][0]]
        assert any(e["value"] == 0 for e in got["lin_cat"][0])


@needs_ref
@pytest.mark.parametrize("n,m", [(3, 2), (0, 1), (4, 0)])
def test_sum_nb_triple_matches_the_reference(n, m):
    rng = np.random.default_rng(31 * n + m)
    a = _value(oracle.NB, *_table(rng, 400, n, m, 0, 6))
    b = _value(oracle.NB, *_table(rng, 250, n, m, 2, 9))
    _same(replay.glue().value_ring("sum_nb_triple", a, b), ref_replay.ref().value_ring("sum_nb_triple", a, b))


@needs_ref
def test_sum_with_the_empty_value_is_the_other_operand():
    """A value whose lists are all empty is the ring's zero (what a partition without rows contributes):
    sum.cpp:86-93 and the like take the other operand's lists."""
    rng = np.random.default_rng(9)
    a = _value(oracle.TRIPLE, *_table(rng, 200, 3, 2, 0, 4))
    zero = {"N": 0, "lin_num": [], "quad_num": [], "lin_cat": [], "quad_num_cat": [], "quad_cat": []}
    for x, y in ((a, zero), (zero, a)):
        _same(replay.glue().value_ring("sum_triple", x, y), ref_replay.ref().value_ring("sum_triple", x, y))


def test_round_trip_and_the_documented_differences():
    """(a + b) - b == a up to FLOAT rounding of the sums and exactly in the counts; a - zero == a; zero - b == -b
    (the reference returns +b there, sub.cpp:93-96: not a ring operation, not mirrored)."""
    rng = np.random.default_rng(4)
    g = replay.glue()
    a = _value(oracle.TRIPLE, *_table(rng, 300, 2, 2, 0, 4))
    b = _value(oracle.TRIPLE, *_table(rng, 100, 2, 2, 0, 4))
    back = g.value_ring("subtract_triple", g.value_ring("sum_triple", a, b), b)
    assert back["N"] == a["N"] and back["lin_cat"] == a["lin_cat"] and back["quad_cat"] == a["quad_cat"]
    np.testing.assert_allclose(back["quad_num"], a["quad_num"], rtol=1e-5, atol=1e-3)
    zero = {"N": 0, "lin_num": [], "quad_num": [], "lin_cat": [], "quad_num_cat": [], "quad_cat": []}
    _same(g.value_ring("subtract_triple", a, zero), a)
    neg = g.value_ring("subtract_triple", zero, b)
    assert neg["N"] == -b["N"] and neg["lin_num"] == [-v for v in b["lin_num"]]
    assert [e["value"] for e in neg["lin_cat"][1]] == [-e["value"] for e in b["lin_cat"][1]]


def test_shape_mismatch_is_an_invalid_input_error():
    rng = np.random.default_rng(5)
    a = _value(oracle.TRIPLE, *_table(rng, 50, 2, 1, 0, 3))
    b = _value(oracle.TRIPLE, *_table(rng, 50, 3, 1, 0, 3))
    with pytest.raises(replay.ReplayError, match="Invalid Input Error: sum_triple"):
        replay.glue().value_ring("sum_triple", a, b)
