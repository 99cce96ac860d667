"""Shared cases of the trainer tests: small correlated tables, the trainer calls made on them, and
what "the same model" means for two parameter lists."""
import numpy as np

from duckdb_imputation_b200.struct_result import arrays_to_struct
from oracle import oracle


def table(seed, rows=600, n=3, doms=(3, 2), neg_keys=False):
    """Correlated columns: x_1 depends on the others and on the first categorical column."""
    rng = np.random.default_rng(seed)
    x = [rng.standard_normal(rows).astype(np.float32) for _ in range(n)]
    c = [rng.integers(0, d, rows).astype(np.int32) for d in doms]
    if c:
        x[min(1, n - 1)] = (2 * x[0] - x[n - 1] + 0.5 * c[0] + 0.1 * rng.standard_normal(rows)).astype(np.float32)
        c[-1] = np.clip((x[0] + 0.7 * rng.standard_normal(rows) > 0).astype(np.int32) + (x[0] > 1), 0, doms[-1] - 1).astype(np.int32)
    if neg_keys:
        c[0] = (c[0] - 1).astype(np.int32)
    return x, c


def triple(x, c):
    return arrays_to_struct(oracle.aggregate_arrays(oracle.TRIPLE, x, c)[0])


CASES = {
    # name: (table args, function, constants)
    "linreg_cat": (dict(seed=0), "linreg_train", (1, 0.001, 0.0, 10000, True, False)),
    "linreg_cat_norm": (dict(seed=1), "linreg_train", (1, 0.001, 0.0, 10000, True, True)),
    "linreg_ridge": (dict(seed=2, doms=(4,)), "linreg_train", (0, 0.001, 0.05, 5000, False, False)),
    "linreg_num_only": (dict(seed=3, n=4, doms=()), "linreg_train", (2, 0.001, 0.0, 10000, True, False)),
    "linreg_neg_keys": (dict(seed=4, neg_keys=True), "linreg_train", (1, 0.001, 0.0, 2000, False, False)),
    "lda_last_label": (dict(seed=5, doms=(3, 3)), "lda_train", (1, 0.001, False)),
    "lda_last_label_norm": (dict(seed=6, doms=(3, 3)), "lda_train", (1, 0.01, True)),
    "lda_only_label": (dict(seed=7, n=4, doms=(3,)), "lda_train", (0, 0.0, False)),
    "lda_shrink": (dict(seed=8, n=2, doms=(2, 4, 3)), "lda_train", (2, 0.4, False)),
}


def same_model(case, got, want, rtol=2e-5, atol=2e-6):
    """Two parameter lists describe the same model: identical header (one-hot layout), and either identical
    parameters or -- where one-hot columns make the solution non-unique (all categories of a column + the intercept
    are collinear; gradient descent then stops wherever rounding noise along the null space took it) -- identical
    predictions on the training rows."""
    targs, fn, consts = CASES[case]
    x, c = table(**targs)
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    assert len(got) == len(want)
    if fn == "linreg_train":
        label, normalize = consts[0], consts[5]
        m = len(c)
        head = 1 + ((m + 1) + int(want[1 + m]) if m else 0)
        assert np.array_equal(got[:head], want[:head])
        if not c or targs.get("neg_keys"):  # (a negative key is emitted as 2^64 + key: no predict function finds it again)
            np.testing.assert_allclose(got, want, rtol=1e-4 if c else rtol, atol=1e-4 if c else atol)
            return
        if consts[4]:
            np.testing.assert_allclose(got[-1], want[-1], rtol=1e-4)  # the residual's standard deviation
        feats = [col for i, col in enumerate(x) if i != label]
        a = oracle.linreg_predict(got, normalize, feats, c)
        b = oracle.linreg_predict(want, normalize, feats, c)
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)
    else:
        label, shrinkage = consts[0], consts[1]
        np.testing.assert_allclose(got, want, rtol=max(rtol, 1e-4), atol=max(atol, 1e-4 * float(np.abs(want).max())))


def restated(case):
    targs, fn, consts = CASES[case]
    t = triple(*table(**targs))
    if fn == "linreg_train":
        return t, oracle.linreg_train(t, *consts)[0]
    return t, oracle.lda_train(t, *consts)
