"""CPU suite: the ring product multiply_triple / multiply_nb_agg.

* the oracle restatement (oracle.multiply) against every golden of the reference's test_mul.py /
  test_nb_mul.py;
* OUR scalar functions (host code behind the DuckDB callbacks -> cfb_result_multiply; no GPU
  involved) against the same goldens, fed with oracle-computed operands here (the `-m gpu` suite feeds
  them with GPU aggregates);
* the ring homomorphism: multiply(sum(A), sum(B)) == sum over the cross join A x B."""
import numpy as np
import pytest

from duckdb_imputation_b200 import replay
from oracle import oracle
from tests import sqlmini

MUL_FILES = ("test_mul.py", "test_nb_mul.py")


def _cases(goldens):
    cases = [c for c in goldens["cases"] if c["file"] in MUL_FILES]
    assert len(cases) == 12
    return cases


def test_oracle_multiply_matches_reference_goldens(goldens):
    mul = lambda A, B: [oracle.multiply(a, b) for a, b in zip(A, B)]
    for c in _cases(goldens):
        got = sqlmini.run_mul(c["sql"], goldens["fixtures"][c["file"]], oracle.aggregate, mul, reference_layout=True)
        assert got[c["index"]] == c["expected"], (c["file"], c["test"], c["index"])


def test_glue_multiply_matches_reference_goldens(goldens):
    g = replay.glue()
    for c in _cases(goldens):
        fn = "multiply_triple" if c["file"] == "test_mul.py" else "multiply_nb_agg"
        got = sqlmini.run_mul(c["sql"], goldens["fixtures"][c["file"]], oracle.aggregate,
                              lambda A, B: g.scalar_structs(fn, A, B), reference_layout=True)
        assert got[c["index"]] == c["expected"], (c["file"], c["test"], c["index"])
        # our function honours the list offsets: on the true join rows it is the true product
        true = sqlmini.run_mul(c["sql"], goldens["fixtures"][c["file"]], oracle.aggregate,
                               lambda A, B: g.scalar_structs(fn, A, B))
        want = sqlmini.run_mul(c["sql"], goldens["fixtures"][c["file"]], oracle.aggregate,
                               lambda A, B: [oracle.multiply(a, b) for a, b in zip(A, B)])
        assert true == want


def _cross(cols_a, cols_b):
    ra, rb = len(cols_a[0]), len(cols_b[0])
    return [np.repeat(c, rb) for c in cols_a], [np.tile(c, ra) for c in cols_b]


@pytest.mark.parametrize("kind", [oracle.TRIPLE, oracle.NB])
@pytest.mark.parametrize("shape", [((3, 2), (2, 1)), ((2, 0), (1, 3)), ((0, 2), (3, 0)), ((4, 1), (0, 2))])
def test_product_of_sums_is_sum_over_cross_join(kind, shape):
    (na, ma), (nb, mb) = shape
    rng = np.random.default_rng(na * 1000 + ma * 100 + nb * 10 + mb)
    ra, rb = 37, 23
    An = [rng.integers(0, 6, ra).astype(np.float32) for _ in range(na)]  # small ints: fp32 exact
    Ac = [rng.integers(-2, 4, ra).astype(np.int32) for _ in range(ma)]
    Bn = [rng.integers(0, 6, rb).astype(np.float32) for _ in range(nb)]
    Bc = [rng.integers(5, 9, rb).astype(np.int32) for _ in range(mb)]
    A = oracle.aggregate(kind, An, Ac)
    B = oracle.aggregate(kind, Bn, Bc)
    xa, xb = _cross(An + Ac, Bn + Bc)
    whole = oracle.aggregate(kind, xa[:na] + xb[:nb], xa[na:] + xb[nb:])
    fn = "multiply_triple" if kind == oracle.TRIPLE else "multiply_nb_agg"
    for got in (oracle.multiply(A, B), replay.glue().scalar_structs(fn, [A], [B])[0]):
        got = dict(got)
        got["lin_agg"], got["quad_agg"] = got.pop("lin_num"), got.pop("quad_num")
        assert got == whole


def test_multiply_many_rows_and_errors():
    rng = np.random.default_rng(3)
    rows = 2500  # > 2048: two chunks
    A = [oracle.aggregate(oracle.TRIPLE, [rng.integers(0, 5, 4).astype(np.float32)], [rng.integers(0, 3, 4).astype(np.int32)])
         for _ in range(rows)]
    B = A[::-1]
    got = replay.glue().scalar_structs("multiply_triple", A, B)
    assert len(got) == rows
    for r in (0, 2047, 2048, rows - 1):
        assert got[r] == oracle.multiply(A[r], B[r])
    with pytest.raises(replay.ReplayError):
        replay.glue().scalar_structs("multiply_triple", A[:1])  # one argument
    with pytest.raises(replay.ReplayError):
        replay.glue().scalar_structs("multiply_nope", A[:1], B[:1])
