"""CPU suite: ring sum / difference of two results (cfb_result_combine, a host function of the C ABI) -- the
arithmetic of the MICE drivers' Value-level sum_triple / subtract_triple (imputation/triple/sum.cpp, sub.cpp:71-219):
sum(A) + sum(B) == sum(A u B) and sum(A u B) - sum(B) == sum(A), keys merged, zero-count keys dropped."""
import ctypes as C

import numpy as np
import pytest

from duckdb_imputation_b200 import _native as nat
from duckdb_imputation_b200.struct_result import result_arrays
from oracle import oracle
from tests.parity import assert_parity, to_result


def _combine(a, b, sign, flags=0):
    ra, ka = to_result(a)
    rb, kb = to_result(b)
    out = nat.Result()
    nat.check(nat.lib().cfb_result_combine(C.byref(ra), C.byref(rb), sign, flags, C.byref(out)))
    try:
        return result_arrays(out)
    finally:
        nat.lib().cfb_result_free(C.byref(out))


@pytest.mark.parametrize("kind", [oracle.TRIPLE, oracle.NB])
@pytest.mark.parametrize("n,m", [(3, 2), (0, 3), (4, 0), (2, 1)])
def test_sum_and_difference_of_results(kind, n, m):
    rng = np.random.default_rng(10 * n + m + kind)
    ra, rb = 700, 400
    num = [rng.random(ra + rb).astype(np.float32) for _ in range(n)]
    cat = [np.concatenate([rng.integers(0, 6, ra), rng.integers(3, 11, rb)]).astype(np.int32) for _ in range(m)]  # B has keys A lacks
    A = oracle.aggregate_arrays(kind, [c[:ra] for c in num], [c[:ra] for c in cat])[0]
    B = oracle.aggregate_arrays(kind, [c[ra:] for c in num], [c[ra:] for c in cat])[0]
    W = oracle.aggregate_arrays(kind, num, cat)[0]
    assert_parity(_combine(A, B, +1), W, what="A + B")
    assert_parity(_combine(W, B, -1), A, what="(A u B) - B")  # the keys only B has disappear again
    assert_parity(_combine(W, A, -1), B, what="(A u B) - A")


def test_combine_rejects_mismatched_shapes_and_signs():
    rng = np.random.default_rng(1)
    A = oracle.aggregate_arrays(oracle.TRIPLE, [rng.random(10).astype(np.float32)], [])[0]
    B = oracle.aggregate_arrays(oracle.TRIPLE, [rng.random(10).astype(np.float32)] * 2, [])[0]
    with pytest.raises(nat.CofactorError, match="shapes"):
        _combine(A, B, 1)
    with pytest.raises(nat.CofactorError, match="sign"):
        _combine(A, A, 2)
    with pytest.raises(nat.CofactorError, match="flag"):
        _combine(A, A, 1, flags=6)


def test_keep_zero_keys_leaves_the_emptied_keys_in_place():
    """CFB_COMBINE_KEEP_ZERO_KEYS: what the reference's std::map merge does (sub.cpp:14-38): (A u B) - B still lists
    the keys only B had, with count 0."""
    rng = np.random.default_rng(3)
    cat = [np.concatenate([rng.integers(0, 4, 50), rng.integers(4, 8, 30)]).astype(np.int32)]
    num = [rng.random(80).astype(np.float32)]
    B = oracle.aggregate_arrays(oracle.TRIPLE, [num[0][50:]], [cat[0][50:]])[0]
    W = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    kept = _combine(W, B, -1, flags=1)
    dropped = _combine(W, B, -1)
    assert list(kept["cat_keys"]) == list(W["cat_keys"]) and len(dropped["cat_keys"]) < len(kept["cat_keys"])
    zero = [i for i, k in enumerate(kept["cat_keys"]) if k >= 4]
    assert zero and all(kept["cat_counts"][i] == 0 for i in zero)
    assert len(kept["pair_counts"]) == len(W["pair_counts"])


# ------------------------------------------------------------ the write-back as ring arithmetic (no second scan)
def _impute(a, model, target):
    from duckdb_imputation_b200.predict import LinearModel
    ra, ka = to_result(a)
    lm = LinearModel(model["bias"], model["w_num"], model["keys"], model["w_cat"], upload=False)
    out = nat.Result()
    nat.check(nat.lib().cfb_result_impute_linear(C.byref(ra), C.byref(lm.c), target, C.byref(out)))
    try:
        return result_arrays(out)
    finally:
        nat.lib().cfb_result_free(C.byref(out))


@pytest.mark.parametrize("n,m,target", [(4, 2, 1), (3, 0, 0), (2, 3, 1), (5, 1, 4)])
def test_impute_linear_equals_the_cofactor_of_the_overwritten_rows(n, m, target):
    """cfb_result_impute_linear(cofactor(R), model, j) == cofactor(R with x_j := model(R)): the delta triple of a
    linear-regression write-back in closed form (the reference rescans, imputation_low.cpp:85-110)."""
    rng = np.random.default_rng(100 * n + 10 * m + target)
    rows = 3000
    num = [rng.standard_normal(rows).astype(np.float32) for _ in range(n)]
    cat = [rng.integers(-1, 4, rows).astype(np.int32) for _ in range(m)]
    keys = [np.array([-1, 0, 1, 2, 3], np.int32) for _ in range(m)]
    model = {"bias": np.array([0.3]), "w_num": rng.standard_normal((1, n - 1)), "keys": keys,
             "w_cat": rng.standard_normal((1, 5 * m))}
    before = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    y = np.full(rows, 0.3)
    for f, i in enumerate(i for i in range(n) if i != target):
        y += model["w_num"][0, f] * num[i].astype(np.float64)
    for c in range(m):
        y += model["w_cat"][0, 5 * c + (cat[c] + 1)]
    after_cols = list(num)
    after_cols[target] = y.astype(np.float32)  # what the predict kernel stores
    want = oracle.aggregate_arrays(oracle.TRIPLE, after_cols, cat)[0]
    got = _impute(before, model, target)
    assert got["N"] == want["N"] and np.array_equal(got["cat_counts"], want["cat_counts"])
    assert np.array_equal(got["pair_counts"], want["pair_counts"])
    scale = max(1.0, float(np.abs(want["quad"]).max()))
    np.testing.assert_allclose(got["lin"], want["lin"], rtol=1e-6, atol=1e-6 * scale)
    np.testing.assert_allclose(got["quad"], want["quad"], rtol=1e-6, atol=1e-6 * scale)
    np.testing.assert_allclose(got["numcat"], want["numcat"], rtol=1e-6, atol=1e-6 * scale)
    # and the entries that do not involve the target are untouched bit for bit
    keep = [i for i in range(n) if i != target]
    assert np.array_equal(got["lin"][keep], before["lin"][keep])
    assert np.array_equal(got["numcat"][keep], before["numcat"][keep])


def test_impute_linear_checks_the_model_shape():
    rng = np.random.default_rng(2)
    A = oracle.aggregate_arrays(oracle.TRIPLE, [rng.random(20).astype(np.float32)] * 3, [rng.integers(0, 3, 20).astype(np.int32)])[0]
    bad = {"bias": np.array([0.0]), "w_num": np.zeros((1, 3)), "keys": [np.arange(3, dtype=np.int32)], "w_cat": np.zeros((1, 3))}
    with pytest.raises(nat.CofactorError, match="does not fit"):
        _impute(A, bad, 0)
    ok = dict(bad, w_num=np.zeros((1, 2)))
    with pytest.raises(nat.CofactorError, match="not a numeric column"):
        _impute(A, ok, 3)
    # an unknown key weighs 0, as in the predict kernels
    short = dict(ok, keys=[np.array([0, 1], np.int32)], w_cat=np.ones((1, 2)))
    got = _impute(A, short, 0)
    assert got["lin"][0] == pytest.approx(float(A["cat_counts"][:2].sum()))
