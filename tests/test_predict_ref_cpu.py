"""CPU suite: the oracle's restatements of the predict functions are pinned to the REFERENCE ITSELF --
ML::linreg_impute (ML/regression.cpp:397-509), LDA_impute (ML/lda.cpp:421-590) and ML::nb_impute
(ML/naive_bayes.cpp:153-263) compiled unmodified from /root/reference into oracle/_ref (oracle/Makefile) and driven
through the replay host like any registered scalar function.  noise = false / deterministic paths: bit-identical."""
import numpy as np
import pytest

from oracle import oracle, ref_replay

pytestmark = pytest.mark.skipif(not ref_replay.available(), reason="oracle/_ref is not built (needs /root/reference)")


def _table(seed, rows=6000):
    rng = np.random.default_rng(seed)
    num = [rng.normal(size=rows).astype(np.float32) * (1 + k) for k in range(4)]
    cat = [rng.integers(0, 5, rows).astype(np.int32), rng.integers(-3, 1, rows).astype(np.int32), rng.integers(100, 103, rows).astype(np.int32)]
    keys = [[0, 1, 2, 3, 4], [-3, -2, -1, 0], [100, 101, 102]]
    return rng, num, cat, keys


@pytest.mark.parametrize("n,m", [(4, 3), (3, 0), (0, 2), (1, 1)])
@pytest.mark.parametrize("normalize", [False, True])
def test_linreg_predict_is_bit_identical_to_the_reference(n, m, normalize):
    rng, num, cat, keys = _table(10 * n + m)
    num, cat, keys = num[:n], cat[:m], keys[:m]
    total = sum(len(k) for k in keys)
    kw = dict(means_num=rng.normal(size=n), means_cat=[rng.random(len(k)) for k in keys]) if normalize else {}
    p = oracle.linreg_params(rng.normal(), rng.normal(size=n), keys, [rng.normal(size=len(k)) for k in keys], sigma=1.5, **kw)
    assert len(p) == 1 + m + (1 + total if m else 0) + 1 + n + total + (n + total if normalize else 0) + 1
    ref = ref_replay.predict("linreg_predict", p, [False, normalize], num, cat)
    assert np.array_equal(ref, oracle.linreg_predict(p, normalize, num, cat))
    keep = rng.random(len(ref)) < 0.3  # behind a filter: DICTIONARY vectors
    ref_f = ref_replay.predict("linreg_predict", p, [False, normalize], num, cat, where=keep)
    assert np.array_equal(ref_f, ref[keep])


@pytest.mark.parametrize("n,m,K", [(4, 3, 3), (4, 0, 5), (2, 2, 2)])
@pytest.mark.parametrize("normalize", [False, True])
def test_lda_predict_is_identical_to_the_reference(n, m, K, normalize):
    rng, num, cat, keys = _table(100 + 10 * n + m)
    num, cat, keys = num[:n], cat[:m], keys[:m]
    total = sum(len(k) for k in keys)
    means = rng.normal(size=n + total) if normalize else None
    p = oracle.lda_params(list(range(10, 10 + K)), rng.normal(size=(K, n + total)), rng.normal(size=K), keys, means=means)
    ref = ref_replay.predict("lda_predict", p, [normalize], num, cat)
    cls, _ = oracle.lda_predict(p, normalize, num, cat)
    assert np.array_equal(ref, cls)


@pytest.mark.parametrize("n,m,K", [(4, 3, 3), (3, 0, 4), (0, 2, 2)])
def test_nb_predict_is_identical_to_the_reference(n, m, K):
    rng, num, cat, keys = _table(200 + 10 * n + m)
    num, cat, keys = num[:n], cat[:m], keys[:m]
    total = sum(len(k) for k in keys)
    labels = [7, 3, 9, 1][:K]
    # column 1 of the model misses key 0: rows with that key get probability 0 in every class (naive_bayes.cpp:236-237)
    model_keys = [k if i != 1 else k[:-1] for i, k in enumerate(keys)]
    mtotal = sum(len(k) for k in model_keys)
    p = oracle.nb_params(labels, rng.random(K) + 0.1, rng.normal(size=(K, n)), rng.random((K, n)) * 3 + 0.05, model_keys,
                         rng.random((K, mtotal)) + 0.01)
    ref = ref_replay.predict("nb_predict", p, [False], num, cat)
    got, prob = oracle.nb_predict(p, num, cat)
    assert np.array_equal(ref, got)
    if m > 1:
        dead = cat[1] == 0
        assert dead.any() and (got[dead] == labels[0]).all() and (prob[dead] == 0).all()


def test_reference_noise_is_gaussian_with_the_models_sigma():
    """What noise = true adds in the reference (regression.cpp:495-505): sigma * N(0, 1) per row from libc random()."""
    rng, num, cat, keys = _table(300, rows=40_000)
    p = oracle.linreg_params(0.5, rng.normal(size=4), keys, [rng.normal(size=len(k)) for k in keys], sigma=2.5)
    clean = ref_replay.predict("linreg_predict", p, [False, False], num, cat)
    ref_replay.seed_libc_random(11)
    noisy = ref_replay.predict("linreg_predict", p, [True, False], num, cat)
    res = (noisy - clean).astype(np.float64)
    assert abs(res.mean()) < 0.05 and abs(res.std() - 2.5) < 0.05


@pytest.mark.parametrize("n,m,K", [(4, 3, 3), (3, 0, 4), (0, 2, 2), (2, 1, 2)])
@pytest.mark.parametrize("normalize", [False, True])
def test_qda_predict_is_identical_to_the_reference(n, m, K, normalize):
    """ML::qda_impute (ML/qda.cpp:338-498) -- compiled from the build-directory copy of qda.cpp that oracle/Makefile
    makes (one cast at qda.cpp:209, which g++ rejects as written)."""
    rng, num, cat, keys = _table(300 + 10 * n + m)
    num, cat, keys = num[:n], cat[:m], keys[:m]
    # the reference reads the keys of the parameter list as FLOAT -> int: keep them non-negative (a negative key is
    # emitted as 2^64 + key by the trainers and never found again)
    cat = [np.abs(c) for c in cat]
    keys = [sorted({abs(k) for k in ks}) for ks in keys]
    P = n + sum(len(k) for k in keys)
    labels = [7, 3, 9, 1][:K]
    quad = rng.normal(size=(K, P, P)) * 0.3
    p = oracle.qda_params(labels, quad, rng.normal(size=(K, P)), rng.normal(size=K), keys,
                          means=rng.normal(size=P) * 0.2 if normalize else None)
    ref = ref_replay.predict("qda_predict", p, [normalize], num, cat)
    got, scores = oracle.qda_predict(p, normalize, num, cat)
    # rows whose two best scores are closer than fp64 summation-order noise may legitimately differ
    top2 = np.sort(scores, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-9 * np.abs(top2).max()
    assert clear.mean() > 0.999
    assert np.array_equal(ref[clear], got[clear])
