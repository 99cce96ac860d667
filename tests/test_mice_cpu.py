"""CPU suite: the host half of tools/mice_loop.py -- moment matrix from a cofactor, closed-form trainers, numpy
predictions -- against scikit-learn fitted on the rows themselves (the cofactor is a sufficient statistic)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import mice_loop  # noqa: E402
from tests import mice_host  # noqa: E402
from oracle import oracle  # noqa: E402

pytest.importorskip("sklearn")


def _table(rows=4000, seed=1):
    num, cat, _, _, _ = mice_loop.synthetic_table(rows, n=4, m=3, dom=5, null_num=(), null_cat=(), seed=seed)
    return num, cat


def test_linreg_from_cofactor_equals_least_squares_on_rows():
    from sklearn.linear_model import LinearRegression
    num, cat = _table()
    res = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    model = mice_loop.train_linreg(res, 1)
    X = np.hstack([np.stack([num[0], num[2], num[3]], 1)] + [np.eye(5)[c] for c in cat])
    ref = LinearRegression().fit(X, num[1]).predict(X)
    got = mice_loop.predict_np(model, [num[0], num[2], num[3]], cat)[:, 0]
    assert np.abs(got - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())


def test_lda_from_cofactor_equals_sklearn_lda():
    from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
    num, cat = _table(seed=2)
    res = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    model = mice_loop.train_lda(res, 0)
    X = np.hstack([np.stack(num, 1)] + [np.eye(5)[c] for c in cat[1:]])
    ref = LinearDiscriminantAnalysis(solver="lsqr").fit(X, cat[0]).predict(X)
    got = model["classes"][np.argmax(mice_loop.predict_np(model, num, cat[1:]), axis=1)]
    assert (got == ref).mean() > 0.98  # one-hot collinearity is resolved by pseudo-inverses on both sides


def test_cpu_mice_loop_imputes_better_than_the_mean():
    num, cat, mn, mc, truth = mice_loop.synthetic_table(6000, n=5, m=3, dom=5, null_num=(0,), null_cat=(1,), seed=3)
    before = np.abs(num[0][mn[0]] - truth[("n", 0)][mn[0]]).mean()
    num, cat = mice_host.mice_cpu(num, cat, mn, mc, 2)
    after = np.abs(num[0][mn[0]] - truth[("n", 0)][mn[0]]).mean()
    assert after < 0.6 * before
    assert (cat[1][mc[1]] == truth[("c", 1)][mc[1]]).mean() > 0.35  # 5 classes: chance is 0.2
