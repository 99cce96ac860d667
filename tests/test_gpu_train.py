"""GPU suite: SURVEY 8 f4 -- the trainers' sigma matrix assembled on the device and the solves run there
(cfb_sigma_*: ridge regression by the reference's gradient descent in one cooperative kernel, LDA by a blocked
Cholesky), through the C ABI and through the linreg_train / lda_train scalar functions, against the numpy restatement
(oracle.build_sigma / linreg_train / lda_train, pinned to the reference's own trainers in tests/test_train_ref_cpu.py),
the committed golden parameter lists, and the reference's trainers themselves where oracle/_ref can run them."""
import json
import os

import numpy as np
import pytest

from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext, replay
from duckdb_imputation_b200._native import CofactorError
from duckdb_imputation_b200.struct_result import arrays_to_struct
from duckdb_imputation_b200.train import Sigma
from oracle import oracle, ref_replay
from tests.train_cases import CASES, restated, same_model, table, triple

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "train_params.json")


def scanned(x, c, domains=None):
    ctx = CofactorContext(CFB_TRIPLE, len(x), len(c))
    if domains is not None:
        ctx.set_cat_domain([d[0] for d in domains], [d[1] for d in domains])
    ctx.append(x, c)
    return ctx


@pytest.mark.parametrize("label_cat,drop_first", [(-1, False), (-1, True), (0, False), (2, False), (1, True)])
def test_sigma_from_the_device_state_and_from_a_result(label_cat, drop_first):
    """Both assembly paths: identical to each other bit for bit (every entry of sigma is one value of the same device
    state), and equal to build_sigma_matrix's restatement over the oracle's cofactor within the aggregate's own
    tolerance (fp32 partial sums folded into fp64; counts exact)."""
    x, c = table(seed=21, rows=900, n=3, doms=(4, 3, 5))
    c[1] = (c[1] - 1).astype(np.int32)  # a negative key: ordered last in its column (uint64 order)
    want, cat_array, idxs = oracle.build_sigma(arrays_to_struct(oracle.aggregate_arrays(oracle.TRIPLE, x, c)[0], narrow=False),
                                               label_cat=label_cat, drop_first=drop_first)
    with scanned(x, c, domains=[(0, 5), (-1, 3), (0, 7)]) as ctx:
        with Sigma.from_context(ctx, label_cat=label_cat, drop_first=drop_first) as a:
            sa, sums_a = a.matrix()
            assert list(a.cat_array) == cat_array and list(a.cat_vars_idxs) == idxs
        res = ctx.finalize_result()
        with Sigma.from_result(res, label_cat=label_cat, drop_first=drop_first) as b:
            sb, sums_b = b.matrix()
            assert list(b.cat_array) == cat_array and list(b.cat_vars_idxs) == idxs
        res.close()
    np.testing.assert_array_equal(sa, sb)
    np.testing.assert_array_equal(sums_a, sums_b)
    np.testing.assert_allclose(sa, want, rtol=1e-5, atol=1e-3)
    counts = want == np.round(want)
    np.testing.assert_array_equal(sa[counts & (np.abs(want) > 3)], want[counts & (np.abs(want) > 3)])
    if label_cat >= 0 and not drop_first:
        # class sums: row c = the cofactor of the rows of class c, first row of ITS sigma
        classes = cat_array[idxs[label_cat]:idxs[label_cat + 1]]
        for ci, key in enumerate(classes):
            rows = c[label_cat] == key
            sub, _, _ = oracle.build_sigma(arrays_to_struct(oracle.aggregate_arrays(
                oracle.TRIPLE, [v[rows] for v in x], [v[rows] for v in c])[0], narrow=False))
            assert sums_a[ci, 0] == rows.sum()
            np.testing.assert_allclose(sums_a[ci, 1:4], sub[0, 1:4], rtol=1e-5, atol=1e-3)
            assert sums_a[ci].sum() > rows.sum()


def test_sigma_from_a_context_with_discovered_keys():
    """No declared domain: the context keeps key dictionaries / discovered ranges; the sigma path goes through the
    canonical result internally."""
    x, c = table(seed=22, rows=500, n=2, doms=(3, 4))
    c[0] = (c[0] * 1000 - 7).astype(np.int32)
    want, cat_array, idxs = oracle.build_sigma(arrays_to_struct(oracle.aggregate_arrays(oracle.TRIPLE, x, c)[0], narrow=False))
    with scanned(x, c) as ctx, Sigma.from_context(ctx) as s:
        np.testing.assert_allclose(s.matrix()[0], want, rtol=1e-5, atol=1e-3)
        assert list(s.cat_array) == cat_array


@pytest.mark.parametrize("case", sorted(CASES))
def test_train_functions_match_the_restatement_and_the_goldens(case):
    """linreg_train / lda_train through the DuckDB scalar-function glue (host/train_glue.cpp)."""
    targs, fn, consts = CASES[case]
    t, want = restated(case)
    got = replay.glue().train(fn, t, *consts)
    same_model(case, got, want)
    with open(GOLDEN) as f:
        same_model(case, got, json.load(f)[case])


@pytest.mark.skipif(not (ref_replay.available() and ref_replay.lapack_available()), reason="oracle/_ref cannot run the trainers here")
@pytest.mark.parametrize("case", sorted(CASES))
def test_train_functions_match_the_reference_trainers(case):
    targs, fn, consts = CASES[case]
    t = triple(*table(**targs))
    same_model(case, replay.glue().train(fn, t, *consts), ref_replay.train(fn, t, *consts))


def test_device_resident_training_step():
    """scan -> sigma -> train without a finalize: the parameters equal the restatement run on the un-narrowed cofactor
    and predict the label column well."""
    x, c = table(seed=31, rows=20_000, n=4, doms=(5, 3))
    t = arrays_to_struct(oracle.aggregate_arrays(oracle.TRIPLE, x, c)[0], narrow=False)
    want, iters = oracle.linreg_train(t, 1, 0.001, 0.0, 3000, True, False)
    with scanned(x, c, domains=[(0, 4), (0, 2)]) as ctx, Sigma.from_context(ctx) as s:
        fit = s.linreg_train(1, 0.001, 0.0, 3000)
        got = s.linreg_params(fit)
    print("iterations", fit["iterations"], iters)
    feats = [v for i, v in enumerate(x) if i != 1]
    a, b = oracle.linreg_predict(got, False, feats, c), oracle.linreg_predict(want, False, feats, c)
    # (the device state differs from the oracle's cofactor in the last bits, and the descent's path is sensitive to
    # them: both runs end near the same optimum, not on the same iterate)
    np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-3)
    assert np.mean((a - x[1]) ** 2) < 0.05 * np.var(x[1])
    # LDA on the last categorical column, from the same state
    with scanned(x, c, domains=[(0, 4), (0, 2)]) as ctx, Sigma.from_context(ctx, label_cat=1) as s:
        got = s.lda_params(s.lda_train(0.01))
    want = oracle.lda_train(t, 1, 0.01, False)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * float(np.abs(want).max()))


def test_wide_one_hot_expansion():
    """What the device path is for: hundreds of one-hot columns.  The gradient descent is the reference's, so the
    iterate after a fixed number of steps equals the restatement's (fp64, summation order aside)."""
    rng = np.random.default_rng(41)
    rows, n, doms = 60_000, 4, (90, 70, 60, 80, 50)
    x = [rng.standard_normal(rows).astype(np.float32) for _ in range(n)]
    c = [rng.integers(0, d, rows).astype(np.int32) for d in doms]
    w = [rng.standard_normal(d) for d in doms]
    x[0] = (x[1] - 0.5 * x[2] + sum(w[k][c[k]] for k in range(len(doms))) + 0.1 * rng.standard_normal(rows)).astype(np.float32)
    t = arrays_to_struct(oracle.aggregate_arrays(oracle.TRIPLE, x, c)[0], narrow=False)
    want, iters = oracle.linreg_train(t, 0, 0.001, 0.01, 400, True, False)
    with scanned(x, c, domains=[(0, d - 1) for d in doms]) as ctx, Sigma.from_context(ctx) as s:
        assert s.p == 1 + n + sum(doms)
        fit = s.linreg_train(0, 0.001, 0.01, 400)
        got = s.linreg_params(fit)
    print("iterations", fit["iterations"], iters)
    feats = x[1:]
    a, b = oracle.linreg_predict(got, False, feats, c), oracle.linreg_predict(want, False, feats, c)
    np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-3)
    # LDA with 349 features on the last column (50 classes)
    with scanned(x, c, domains=[(0, d - 1) for d in doms]) as ctx, Sigma.from_context(ctx, label_cat=4) as s:
        got = s.lda_params(s.lda_train(0.1))
    want = oracle.lda_train(t, 4, 0.1, False)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5 * float(np.abs(want).max()))


def test_training_errors_are_reported():
    x, c = table(seed=51, rows=400, n=2, doms=(3, 2))
    t = triple(x, c)
    g = replay.glue()
    with pytest.raises(replay.ReplayError, match="not a numeric column"):
        g.train("linreg_train", t, 5, 0.001, 0.0, 10, False, False)
    with pytest.raises(replay.ReplayError, match="not a categorical column"):
        g.train("lda_train", t, 2, 0.001, False)
    with pytest.raises(replay.ReplayError, match="not positive definite"):
        g.train("lda_train", t, 1, 0.0, False)  # the one-hot columns of column 0 are collinear: no shrinkage, no inverse
    with scanned(x, c, domains=[(0, 2), (0, 1)]) as ctx:
        with Sigma.from_context(ctx, label_cat=1) as s, pytest.raises(CofactorError, match="LDA"):
            s.linreg_train(0, 0.001, 0.0, 10)
        with Sigma.from_context(ctx) as s, pytest.raises(CofactorError, match="no categorical label"):
            s.lda_train(0.1)


def test_sigma_of_one_group_slot_numeric_only_and_empty():
    """cfb_sigma_from_ctx on a GROUP BY context (every slot has its own state), on a table without categorical columns,
    and on a context that has seen no rows."""
    rng = np.random.default_rng(61)
    rows = 5_000
    x = [rng.standard_normal(rows).astype(np.float32) for _ in range(3)]
    c = [rng.integers(0, 4, rows).astype(np.int32)]
    g = rng.integers(0, 3, rows).astype(np.int32)
    with CofactorContext(CFB_TRIPLE, 3, 1, n_groups=3) as ctx:
        ctx.set_cat_domain([0], [3])
        ctx.append(x, c, group=g.astype(np.uint32))
        for slot in range(3):
            sel = g == slot
            want, cat_array, _ = oracle.build_sigma(arrays_to_struct(
                oracle.aggregate_arrays(oracle.TRIPLE, [v[sel] for v in x], [v[sel] for v in c])[0], narrow=False))
            with Sigma.from_context(ctx, group=slot) as s:
                assert list(s.cat_array) == cat_array
                np.testing.assert_allclose(s.matrix()[0], want, rtol=1e-5, atol=1e-3)
    with CofactorContext(CFB_TRIPLE, 3, 0) as ctx:
        ctx.append(x, [])
        want, _, _ = oracle.build_sigma(arrays_to_struct(oracle.aggregate_arrays(oracle.TRIPLE, x, [])[0], narrow=False))
        with Sigma.from_context(ctx) as s:
            assert s.p == 4 and len(s.cat_array) == 0
            np.testing.assert_allclose(s.matrix()[0], want, rtol=1e-5, atol=1e-3)
            fit = s.linreg_train(2, 0.01, 0.0, 500)
            assert fit["coeff"][3] == -1.0 and np.all(np.isfinite(fit["coeff"]))
    with CofactorContext(CFB_TRIPLE, 2, 1) as ctx:
        with Sigma.from_context(ctx) as s:  # no rows: N = 0, no keys
            assert s.p == 3 and s.matrix()[0].sum() == 0.0
            fit = s.linreg_train(0, 0.01, 0.0, 10)
            assert list(fit["coeff"]) == [0.0, -1.0, 0.0]
