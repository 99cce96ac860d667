"""GPU suite: the reference's own end-to-end ML tests (duckdb_extension/test/python/test_regression.py:96-172,
test_LDA.py:94-207) replayed through this build -- iris, `sum_to_triple_x_y` on the GPU, `linreg_train` / `lda_train`
(sigma + solve on the device), `linreg_predict` / `lda_predict` (predict kernels) -- with the reference's own pass
criteria: R^2 / accuracy equal to scikit-learn's to 3 decimals (0.2 for the categorical regression case)."""
import numpy as np
import pytest

from duckdb_imputation_b200 import replay

pytestmark = pytest.mark.gpu

pd = pytest.importorskip("pandas")
sk = pytest.importorskip("sklearn")


def _iris(discretize):
    from sklearn.datasets import load_iris
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import KBinsDiscretizer
    data = load_iris(as_frame=True, return_X_y=True)
    df = data[0].rename(columns={"sepal length (cm)": "s_length", "sepal width (cm)": "s_width",
                                 "petal length (cm)": "p_length", "petal width (cm)": "p_width"})
    if discretize:
        est = KBinsDiscretizer(n_bins=4, encode="ordinal", strategy="uniform", subsample=None)
        df[discretize] = est.fit_transform(df[discretize])
    enc = pd.get_dummies(df, columns=discretize) if discretize else df.copy()
    enc["target"] = data[1]
    tr, te, ytr, yte = train_test_split(df, data[1], test_size=0.33, random_state=42)
    tr, te = tr.assign(target=ytr), te.assign(target=yte)
    etr, ete = train_test_split(enc, test_size=0.33, random_state=42)
    return tr, te, etr, ete


def _cols(df, num, cat):
    return [df[c].to_numpy(np.float32) for c in num], [df[c].to_numpy(np.int32) for c in cat]


def _triple(df, num, cat):
    g = replay.glue()
    return g.aggregate("sum_to_triple_%d_%d" % (len(num), len(cat)), *_cols(df, num, cat))[0]


@pytest.mark.parametrize("normalize", [False, True])
def test_linreg_like_test_regression_py(normalize):
    """test_linreg_no_norm / test_linreg_norm (test_regression.py:128-172): s_length from the other three + target."""
    from sklearn.linear_model import LinearRegression
    from sklearn.metrics import r2_score
    tr, te, _, _ = _iris([])
    g = replay.glue()
    t = _triple(tr, ["s_length", "s_width", "p_length", "p_width"], ["target"])
    params = g.train("linreg_train", t, 0, 0.001, 0.0, 10000, False, normalize)
    pred = g.predict("linreg_predict", params, [False, normalize], *_cols(te, ["s_width", "p_length", "p_width"], ["target"]))
    r2_ours = r2_score(te["s_length"], pred)
    trd, ted = pd.get_dummies(tr, columns=["target"]), pd.get_dummies(te, columns=["target"])
    reg = LinearRegression().fit(trd.drop(["s_length"], axis=1), trd["s_length"])
    r2_py = reg.score(ted.drop(["s_length"], axis=1), ted["s_length"])
    # the reference asserts round(.., 3) == round(.., 3); scikit-learn's 0.8606 sits next to a rounding boundary and the
    # descent stops a few 1e-5 away from the least-squares optimum, so the criterion here is the distance itself
    assert abs(r2_ours - r2_py) < 1e-3


def test_linreg_with_categorical_features_like_test_lr_no_norm_cat():
    """test_lr_no_norm_cat (test_regression.py:96-126): p_length from p_width + binned sepal columns + target."""
    from sklearn.linear_model import LinearRegression
    from sklearn.metrics import r2_score
    tr, te, etr, ete = _iris(["s_length", "s_width"])
    g = replay.glue()
    t = _triple(tr, ["p_width", "p_length"], ["s_length", "s_width", "target"])
    params = g.train("linreg_train", t, 1, 0.001, 0.0, 10000, False, False)
    pred = g.predict("linreg_predict", params, [False, False], *_cols(te, ["p_width"], ["s_length", "s_width", "target"]))
    r2_ours = r2_score(te["p_length"], pred)
    reg = LinearRegression().fit(etr.drop(["p_length"], axis=1), etr["p_length"])
    r2_py = reg.score(ete.drop(["p_length"], axis=1), ete["p_length"])
    assert round(r2_py, 2) >= round(r2_ours, 2) + 0.2 or round(r2_py, 2) + 0.2 >= round(r2_ours, 2)  # the reference's criterion
    assert abs(r2_py - r2_ours) < 0.05                                                               # and a real one


@pytest.mark.parametrize("normalize", [False, True])
def test_lda_like_test_lda_py(normalize):
    """test_lda_no_norm / test_lda_norm (test_LDA.py:164-207): target from the four measurements, shrinkage 0."""
    from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
    tr, te, _, _ = _iris([])
    g = replay.glue()
    num = ["s_length", "s_width", "p_length", "p_width"]
    t = _triple(tr, num, ["target"])
    params = g.train("lda_train", t, 0, 0.0, normalize)
    pred = g.predict("lda_predict", params, [normalize], *_cols(te, num, []))
    acc_ours = float(np.mean(pred == te["target"].to_numpy()))
    clf = LinearDiscriminantAnalysis(solver="lsqr", shrinkage=0).fit(tr[num], tr["target"])
    assert round(acc_ours, 3) == round(clf.score(te[num], te["target"]), 3)


@pytest.mark.parametrize("normalize", [False, True])
def test_lda_with_categorical_features_like_test_lda_cat(normalize):
    """test_lda_no_norm_cat / test_lda_norm_cat (test_LDA.py:94-162): target (the LAST categorical column) from p_width
    and three binned columns, shrinkage 0.01."""
    from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
    tr, te, etr, ete = _iris(["s_length", "s_width", "p_length"])
    g = replay.glue()
    t = _triple(tr, ["p_width"], ["s_length", "s_width", "p_length", "target"])
    params = g.train("lda_train", t, 3, 0.01, normalize)
    pred = g.predict("lda_predict", params, [normalize], *_cols(te, ["p_width"], ["s_length", "s_width", "p_length"]))
    acc_ours = float(np.mean(pred == te["target"].to_numpy()))
    clf = LinearDiscriminantAnalysis(solver="lsqr", shrinkage=0).fit(etr.drop(["target"], axis=1), etr["target"])
    acc_py = clf.score(ete.drop(["target"], axis=1), ete["target"])
    assert round(acc_ours, 3) == round(acc_py, 3)


def _per_class_params(kind_fn, tr, num, *consts):
    """list(agg), list(target) FROM (SELECT agg(...) GROUP BY target) -> the REFERENCE's own per-class trainer (this
    build has no qda_train / nb_train: out of scope) -> the FLOAT[] our predict functions must understand."""
    from oracle import ref_replay
    if not (ref_replay.available() and ref_replay.lapack_available()):
        pytest.skip("oracle/_ref cannot run the reference's trainers here")
    g = replay.glue()
    labels = sorted(int(v) for v in tr["target"].unique())
    agg = "sum_to_triple_%d_0" if kind_fn == "qda_train" else "sum_to_nb_agg_%d_0"
    triples = [g.aggregate(agg % len(num), *_cols(tr[tr.target == l], num, []))[0] for l in labels]  # GPU aggregates
    return ref_replay.train_list(kind_fn, triples, labels, *consts)


def test_qda_predict_like_test_qda_py():
    """test_qda_no_norm (test_QDA.py:46-68): per-class triples from the GPU aggregate, the reference's qda_train, OUR
    qda_predict, accuracy equal to scikit-learn's QDA.  (test_qda_norm is not mirrored: the reference's qda_train with
    normalize = true returns different parameters from call to call in one process -- uninitialised memory.)"""
    from sklearn.discriminant_analysis import QuadraticDiscriminantAnalysis
    tr, te, _, _ = _iris([])
    num = ["s_length", "s_width", "p_length", "p_width"]
    params = _per_class_params("qda_train", tr, num, False)
    pred = replay.glue().predict("qda_predict", params, [False], *_cols(te, num, []))
    clf = QuadraticDiscriminantAnalysis(store_covariance=True).fit(tr[num], tr["target"])
    assert round(float(np.mean(pred == te["target"].to_numpy())), 3) == round(clf.score(te[num], te["target"]), 3)


def test_nb_predict_like_test_nb_py():
    """test_nb_no_norm (test_NB.py:46-72): per-class NB aggregates from the GPU, the reference's nb_train, OUR nb_predict,
    accuracy equal to scikit-learn's GaussianNB."""
    from sklearn.naive_bayes import GaussianNB
    tr, te, _, _ = _iris([])
    num = ["s_length", "s_width", "p_length", "p_width"]
    params = _per_class_params("nb_train", tr, num)
    pred = replay.glue().predict("nb_predict", params, [False], *_cols(te, num, []))
    clf = GaussianNB().fit(tr[num], tr["target"])
    assert round(float(np.mean(pred == te["target"].to_numpy())), 3) == round(clf.score(te[num], te["target"]), 3)
