"""CPU suite: the reference's own ML tests (duckdb_extension/test/python/test_regression.py, test_LDA.py) replayed on
the checkers -- (a) the REFERENCE's trainers and predict functions compiled into oracle/_ref, (b) the numpy
restatements (oracle.linreg_train / lda_train / linreg_predict / lda_predict) -- with the reference's own criteria:
R^2 / accuracy equal to scikit-learn's to 3 decimals.  tests/test_gpu_iris.py runs the same cases through the GPU
build."""
import numpy as np
import pytest

pd = pytest.importorskip("pandas")
pytest.importorskip("sklearn")

from duckdb_imputation_b200.struct_result import arrays_to_struct
from oracle import oracle, ref_replay
from tests.test_gpu_iris import _cols, _iris

needs_ref = pytest.mark.skipif(not (ref_replay.available() and ref_replay.lapack_available()), reason="oracle/_ref cannot run the trainers here")
NUM = ["s_length", "s_width", "p_length", "p_width"]


def _struct(df, num, cat):
    return arrays_to_struct(oracle.aggregate_arrays(oracle.TRIPLE, *_cols(df, num, cat))[0])


def _sklearn_linreg(tr, te):
    from sklearn.linear_model import LinearRegression
    trd, ted = pd.get_dummies(tr, columns=["target"]), pd.get_dummies(te, columns=["target"])
    reg = LinearRegression().fit(trd.drop(["s_length"], axis=1), trd["s_length"])
    return reg.score(ted.drop(["s_length"], axis=1), ted["s_length"])


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("who", ["reference", "restatement"])
def test_linreg_iris(who, normalize):
    from sklearn.metrics import r2_score
    if who == "reference" and not (ref_replay.available()):
        pytest.skip("oracle/_ref not built")
    tr, te, _, _ = _iris([])
    t = _struct(tr, NUM, ["target"])
    feats = _cols(te, ["s_width", "p_length", "p_width"], ["target"])
    if who == "reference":
        p = ref_replay.train("linreg_train", t, 0, 0.001, 0.0, 10000, False, normalize)
        pred = ref_replay.predict("linreg_predict", p, [False, normalize], *feats)
    else:
        p, _ = oracle.linreg_train(t, 0, 0.001, 0.0, 10000, False, normalize)
        pred = oracle.linreg_predict(np.concatenate([p, [0.0]]).astype(np.float32), normalize, *feats)  # (+ the sigma slot)
    assert round(r2_score(te["s_length"], pred), 3) == round(_sklearn_linreg(tr, te), 3)


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("who", ["reference", "restatement"])
def test_lda_iris(who, normalize):
    from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
    if who == "reference" and not (ref_replay.available() and ref_replay.lapack_available()):
        pytest.skip("oracle/_ref cannot run lda_train here")
    tr, te, _, _ = _iris([])
    t = _struct(tr, NUM, ["target"])
    feats = _cols(te, NUM, [])
    if who == "reference":
        p = ref_replay.train("lda_train", t, 0, 0.0, normalize)
        pred = ref_replay.predict("lda_predict", p, [normalize], *feats)
    else:
        p = oracle.lda_train(t, 0, 0.0, normalize)
        idx, _ = oracle.lda_predict(p, normalize, *feats)
        pred = idx
    clf = LinearDiscriminantAnalysis(solver="lsqr", shrinkage=0).fit(tr[NUM], tr["target"])
    assert round(float(np.mean(np.asarray(pred) == te["target"].to_numpy())), 3) == round(clf.score(te[NUM], te["target"]), 3)


@needs_ref
@pytest.mark.parametrize("fn,pfn,kind", [("qda_train", "qda_predict", 0), ("nb_train", "nb_predict", 1)])
def test_per_class_models_iris(fn, pfn, kind):
    """test_QDA.py:46-68 / test_NB.py:46-72 on the reference build, and the restated predict functions on the same
    parameter lists (class by class identical)."""
    from sklearn.discriminant_analysis import QuadraticDiscriminantAnalysis
    from sklearn.naive_bayes import GaussianNB
    tr, te, _, _ = _iris([])
    labels = sorted(int(v) for v in tr["target"].unique())
    triples = [arrays_to_struct(oracle.aggregate_arrays(oracle.TRIPLE if kind == 0 else oracle.NB, *_cols(tr[tr.target == l], NUM, []))[0])
               for l in labels]
    params = ref_replay.train_list(fn, triples, labels, *((False,) if kind == 0 else ()))
    feats = _cols(te, NUM, [])
    ref = ref_replay.predict(pfn, params, [False], *feats)
    mine = oracle.qda_predict(params, False, *feats)[0] if kind == 0 else oracle.nb_predict(params, *feats)[0]
    assert np.array_equal(ref, mine)
    clf = (QuadraticDiscriminantAnalysis(store_covariance=True) if kind == 0 else GaussianNB()).fit(tr[NUM], tr["target"])
    assert round(float(np.mean(ref == te["target"].to_numpy())), 3) == round(clf.score(te[NUM], te["target"]), 3)
