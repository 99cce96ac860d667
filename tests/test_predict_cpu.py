"""CPU suite: the oracle restatement of the predict functions (ML::linreg_impute, LDA_impute), pinned the way
the reference's own tests pin them -- agreement with scikit-learn on iris (test_regression.py:100-160,
test_LDA.py:100-190; the reference compares R^2 / accuracy, here the predictions themselves) -- and the
parameter-list layout against a hand-computed case."""
import numpy as np
import pytest

from oracle import oracle

sklearn = pytest.importorskip("sklearn")


def _iris():
    from sklearn.datasets import load_iris
    X, y = load_iris(return_X_y=True)
    return X.astype(np.float32), y.astype(np.int32)


def test_linreg_predict_agrees_with_sklearn_on_iris():
    """test_regression.py:100-118: p_length ~ p_width, s_length, s_width + one-hot(target)."""
    from sklearn.linear_model import LinearRegression
    X, y = _iris()
    feat = np.stack([X[:, 3], X[:, 0], X[:, 1]], 1)
    design = np.hstack([feat, np.eye(3)[y]])
    reg = LinearRegression().fit(design, X[:, 2])
    params = oracle.linreg_params(reg.intercept_, reg.coef_[:3], [[0, 1, 2]], [reg.coef_[3:]])
    pred = oracle.linreg_predict(params, False, list(feat.T), [y])
    assert np.abs(pred - reg.predict(design)).max() < 1e-5
    # normalize = true: weights on centred features, means appended to the list -- same predictions
    means_num, means_cat = feat.mean(0), np.eye(3)[y].mean(0)
    icpt = reg.intercept_ + reg.coef_[:3] @ means_num + reg.coef_[3:] @ means_cat
    params = oracle.linreg_params(icpt, reg.coef_[:3], [[0, 1, 2]], [reg.coef_[3:]], means_num, [means_cat])
    pred_n = oracle.linreg_predict(params, True, list(feat.T), [y])
    assert np.abs(pred_n - reg.predict(design)).max() < 1e-5


def test_lda_predict_agrees_with_sklearn_on_iris():
    """test_LDA.py:155-170: the class index of the largest linear score."""
    from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
    X, y = _iris()
    lda = LinearDiscriminantAnalysis().fit(X, y)
    params = oracle.lda_params([0, 1, 2], lda.coef_, lda.intercept_, [])
    cls, _ = oracle.lda_predict(params, False, list(X.T), [])
    assert (cls == lda.predict(X)).all()
    mean = X.mean(0)
    params = oracle.lda_params([0, 1, 2], lda.coef_, lda.intercept_ + lda.coef_ @ mean, [], means=mean)
    cls_n, _ = oracle.lda_predict(params, True, list(X.T), [])
    assert (cls_n == cls).all()


def test_parameter_layout_by_hand():
    # 1 numeric + 1 categorical column with keys {4, 8}: y = 10 + 2 x + {4: 0.5, 8: -1}[c]
    p = oracle.linreg_params(10.0, [2.0], [[4, 8]], [[0.5, -1.0]], sigma=3.0)
    assert list(p) == [1, 0, 2, 4, 8, 10, 2, 0.5, -1, 3]  # regression.cpp:424-435, sigma last (:503)
    out = oracle.linreg_predict(p, False, [np.array([1.0, 2.0], np.float32)], [np.array([8, 4], np.int32)])
    assert list(out) == [11.0, 14.5]
    with pytest.raises(ValueError):
        oracle.linreg_predict(p, False, [np.array([1.0], np.float32)], [np.array([5], np.int32)])
    # LDA, 2 classes, 1 numeric + keys {3, 7}
    q = oracle.lda_params([20, 30], [[1.0, 0.0, 2.0], [-1.0, 5.0, 0.0]], [0.0, 0.5], [[3, 7]])
    assert list(q) == [2, 2, 0, 2, 3, 7, 20, 30, 1, 0, 2, -1, 5, 0, 0, 0.5]  # lda.cpp:450-500
    cls, scores = oracle.lda_predict(q, False, [np.array([1.0, 1.0], np.float32)], [np.array([3, 7], np.int32)])
    assert scores.tolist() == [[1.0, 4.5], [3.0, -0.5]] and list(cls) == [1, 0]
