"""GPU suite: linreg_predict / lda_predict (predict_kernel behind cfb_predict_host / cfb_predict_device) against
the oracle restatement of ML::linreg_impute / LDA_impute.

Bar: regression scores <= 1e-5 relative (fp64 accumulation on both sides, FLOAT result); LDA class indices
identical wherever the two best scores are not within 1e-9 of each other."""
import numpy as np
import pytest

from duckdb_imputation_b200 import predict, replay
from oracle import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _random_model(rng, n, doms, K):
    keys = [np.sort(rng.choice(np.arange(-5, 4 * d), d, replace=False)).astype(np.int32) for d in doms]
    w_num = rng.standard_normal((K, n))
    w_cat = [rng.standard_normal((K, d)) for d in doms]
    bias = rng.standard_normal(K)
    return keys, w_num, w_cat, bias


def _rows(rng, rows, n, keys):
    num = [rng.standard_normal(rows).astype(np.float32) for _ in range(n)]
    cat = [k[rng.integers(0, len(k), rows)].astype(np.int32) for k in keys]
    return num, cat


@pytest.mark.parametrize("n,doms,normalize", [(3, [3], False), (20, [100] * 10, False), (20, [100] * 10, True), (0, [7, 9], True),
                                              (5, [], False), (32, [4], True)])
def test_linreg_predict_through_the_callbacks(n, doms, normalize):
    rng = np.random.default_rng(31 * n + len(doms) + normalize)
    keys, w_num, w_cat, bias = _random_model(rng, n, doms, 1)
    rows = 10_007  # several 2048-row chunks and a ragged tail
    num, cat = _rows(rng, rows, n, keys)
    means_num = rng.standard_normal(n) if normalize else None
    means_cat = [rng.random(len(k)) for k in keys] if normalize else None
    params = oracle.linreg_params(bias[0], w_num[0], keys, [w[0] for w in w_cat], means_num, means_cat, sigma=0.7)
    ref = oracle.linreg_predict(params, normalize, num, cat)
    got = replay.glue().predict("linreg_predict", params, [False, normalize], num, cat)
    assert got.dtype == np.float32 and np.allclose(got, ref, rtol=1e-5, atol=1e-5)
    keep = rng.random(rows) < 0.2  # the NULL cells of a MICE step
    got_w = replay.glue().predict("linreg_predict", params, [False, normalize], num, cat, where=keep)
    assert np.allclose(got_w, ref[keep], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,doms,K,normalize", [(4, [], 3, False), (12, [10, 20], 10, False), (12, [10, 20], 10, True), (0, [5], 2, False)])
def test_lda_predict_through_the_callbacks(n, doms, K, normalize):
    rng = np.random.default_rng(7 * n + K + normalize)
    keys, w_num, w_cat, bias = _random_model(rng, n, doms, K)
    rows = 6_000
    num, cat = _rows(rng, rows, n, keys)
    coef = np.hstack([w_num] + w_cat)
    means = rng.standard_normal(coef.shape[1]) if normalize else None
    params = oracle.lda_params(np.arange(K) + 10, coef, bias, keys, means)
    ref, scores = oracle.lda_predict(params, normalize, num, cat)
    got = replay.glue().predict("lda_predict", params, [normalize], num, cat)
    top2 = np.sort(scores, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-6
    assert got.dtype == np.int32 and clear.mean() > 0.99 and (got[clear] == ref[clear]).all()


def test_device_resident_predict_overwrites_only_masked_cells():
    """cfb_predict_device with the output aliasing the imputed column: NULL cells (mask != 0) get the score,
    observed cells keep their value."""
    rng = np.random.default_rng(5)
    rows, n = 1_000_003, 6
    keys, w_num, w_cat, bias = _random_model(rng, n, [50, 60], 1)
    num, cat = _rows(rng, rows, n, keys)
    target = rng.standard_normal(rows).astype(np.float32)
    mask = (rng.random(rows) < 0.2).astype(np.int32)
    model = predict.LinearModel(bias, w_num, keys, np.hstack(w_cat))
    d_num = [torch.from_numpy(c).cuda() for c in num]
    d_cat = [torch.from_numpy(c).cuda() for c in cat]
    d_col = torch.from_numpy(target.copy()).cuda()
    predict.predict_device(model, d_num, d_cat, rows, predict.SCORE, d_col, d_mask=torch.from_numpy(mask).cuda())
    torch.cuda.synchronize()
    got = d_col.cpu().numpy()
    params = oracle.linreg_params(bias[0], w_num[0], keys, [w[0] for w in w_cat])
    ref = oracle.linreg_predict(params, False, num, cat)
    assert np.array_equal(got[mask == 0], target[mask == 0])
    assert np.allclose(got[mask != 0], ref[mask != 0], rtol=1e-5, atol=1e-5)
    # unknown key: contributes 0 (documented difference; the reference reads past the weights)
    cat2 = [c.copy() for c in cat]
    cat2[0][:] = 1_000_000
    out = predict.predict_host(model, num, cat2, predict.SCORE)
    cat_only = np.hstack(w_cat)[0][np.searchsorted(keys[0], cat[0])]
    assert np.allclose(out, ref - cat_only, rtol=1e-4, atol=1e-4)


def test_bad_parameter_lists_are_query_errors():
    p = oracle.linreg_params(1.0, [2.0], [], [])
    x = [np.ones(10, np.float32)]
    with pytest.raises(replay.ReplayError, match="too short"):
        replay.glue().predict("linreg_predict", p[:2], [False, False], x, [])
    with pytest.raises(replay.ReplayError, match="categorical"):
        replay.glue().predict("linreg_predict", p, [False, False], x, [np.ones(10, np.int32)])


def test_less_common_predict_paths():
    """Unaligned device columns (scalar path of the score kernel), score_0 of a multi-output model, argmax of a
    single-output model (always index 0)."""
    rng = np.random.default_rng(77)
    rows, n = 50_003, 5
    keys, w_num, w_cat, bias = _random_model(rng, n, [9], 3)
    num, cat = _rows(rng, rows + 1, n, keys)
    d_num = [torch.from_numpy(c).cuda()[1:] for c in num]  # 4-byte aligned only
    d_cat = [torch.from_numpy(c).cuda()[1:] for c in cat]
    one = predict.LinearModel(bias[:1], w_num[:1], keys, w_cat[0][:1])
    out = torch.zeros(rows, dtype=torch.float32, device="cuda")
    predict.predict_device(one, d_num, d_cat, rows, predict.SCORE, out)
    params = oracle.linreg_params(bias[0], w_num[0], keys, [w_cat[0][0]])
    ref = oracle.linreg_predict(params, False, [c[1:] for c in num], [c[1:] for c in cat])
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    idx = torch.full((rows,), 7, dtype=torch.int32, device="cuda")
    predict.predict_device(one, d_num, d_cat, rows, predict.ARGMAX, idx)
    assert int(idx.abs().sum().item()) == 0
    multi = predict.LinearModel(bias, w_num, keys, w_cat[0])
    predict.predict_device(multi, d_num, d_cat, rows, predict.SCORE, out)
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)  # score of output 0
    torch.cuda.synchronize()


# ------------------------------------------------------------------ round 2: pinned to the reference itself; noise; nb / qda
def _ref():
    from oracle import ref_replay
    if not ref_replay.available():
        pytest.skip("oracle/_ref is not built")
    return ref_replay


@pytest.mark.parametrize("normalize", [False, True])
def test_gpu_linreg_and_lda_equal_the_references_own_functions(normalize):
    """The GPU predictions against ML::linreg_impute / LDA_impute THEMSELVES (compiled from /root/reference into
    oracle/_ref), same parameter list, same chunks: regression scores equal to one float32 ulp (fp64 sums in a
    different order, rounded to FLOAT), class indices identical away from ties."""
    ref_replay = _ref()
    rng = np.random.default_rng(5 + normalize)
    keys, w_num, w_cat, bias = _random_model(rng, 6, [5, 9], 4)
    num, cat = _rows(rng, 9_001, 6, keys)
    mn = rng.standard_normal(6) if normalize else None
    mc = [rng.random(len(k)) for k in keys] if normalize else None
    p = oracle.linreg_params(bias[0], w_num[0], keys, [w[0] for w in w_cat], mn, mc, sigma=0.3)
    want = ref_replay.predict("linreg_predict", p, [False, normalize], num, cat)
    got = replay.glue().predict("linreg_predict", p, [False, normalize], num, cat)
    ulp = np.spacing(np.abs(want).astype(np.float32))
    # normalize: the reference forms each centred categorical term as a FLOAT product before widening it
    # (regression.cpp:478-491); the device keeps fp64 throughout, so a few float32 ulps of the partial sums remain
    assert (np.abs(got - want) <= (8 if normalize else 1) * np.maximum(ulp, np.float32(1.2e-7))).all()
    assert (got == want).mean() > (0.5 if normalize else 0.98)
    coef = np.hstack([w_num] + w_cat)
    means = rng.standard_normal(coef.shape[1]) if normalize else None
    q = oracle.lda_params(np.arange(4), coef, bias, keys, means)
    want_c = ref_replay.predict("lda_predict", q, [normalize], num, cat)
    got_c = replay.glue().predict("lda_predict", q, [normalize], num, cat)
    _, scores = oracle.lda_predict(q, normalize, num, cat)
    top2 = np.sort(scores, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-9
    assert clear.mean() > 0.999 and (got_c[clear] == want_c[clear]).all()


def test_stochastic_regression_noise():
    """linreg_predict(params, noise = true, ...): what the MICE driver always calls (imputation_base.cpp:133).  The
    residual against the noise-free prediction is N(0, sigma^2) with sigma = the list's last entry (regression.cpp:503);
    the draw of a row is a function of (seed, position in the stream): reproducible, independent of the chunking."""
    import ctypes as C
    import os
    from duckdb_imputation_b200 import _native as nat
    rng = np.random.default_rng(9)
    keys, w_num, w_cat, bias = _random_model(rng, 4, [6], 1)
    rows = 200_000
    num, cat = _rows(rng, rows, 4, keys)
    sigma = 1.75
    p = oracle.linreg_params(bias[0], w_num[0], keys, [w[0] for w in w_cat], sigma=sigma)
    clean = replay.glue().predict("linreg_predict", p, [False, False], num, cat)
    os.environ["CFB_NOISE_SEED"] = "1234"
    noisy = replay.glue().predict("linreg_predict", p, [True, False], num, cat)
    res = (noisy - clean).astype(np.float64)
    assert abs(res.mean()) < 4 * sigma / np.sqrt(rows) and abs(res.std() - sigma) < 0.01 * sigma
    # a normal, not just the right two moments: skewness ~ 0, kurtosis ~ 3, no row-to-row correlation
    z = res / res.std()
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3.0) < 0.08 and abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 0.01
    # the reference's own noise (libc random, oracle/_ref) has the same distribution: compare a few quantiles
    ref_replay = _ref()
    ref_replay.seed_libc_random(3)
    ref_res = (ref_replay.predict("linreg_predict", p, [True, False], num, cat) - clean).astype(np.float64)
    qs = [0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99]
    assert np.abs(np.quantile(res, qs) - np.quantile(ref_res, qs)).max() < 0.03 * sigma
    # C ABI: the stream position makes chunked scoring identical to scoring in one call
    M = predict.LinearModel(bias, w_num, keys, np.hstack(w_cat))
    lib = nat.lib()
    dn = [torch.from_numpy(c).cuda() for c in num]
    dc = [torch.from_numpy(c).cuda() for c in cat]
    whole = torch.empty(rows, dtype=torch.float32, device="cuda")
    parts = torch.empty(rows, dtype=torch.float32, device="cuda")
    nat.check(lib.cfb_model_set_noise(M._h, sigma, 77, 0))
    predict.predict_device(M, dn, dc, rows, predict.SCORE, whole)
    for lo in range(0, rows, 40_000):
        hi = min(rows, lo + 40_000)
        nat.check(lib.cfb_model_set_noise(M._h, sigma, 77, lo))
        predict.predict_device(M, [t[lo:hi] for t in dn], [t[lo:hi] for t in dc], hi - lo, predict.SCORE, parts[lo:hi])
    torch.cuda.synchronize()
    assert torch.equal(whole, parts)
    nat.check(lib.cfb_model_set_noise(M._h, sigma, 78, 0))  # another seed: another draw
    predict.predict_device(M, dn, dc, rows, predict.SCORE, parts)
    torch.cuda.synchronize()
    assert not torch.equal(whole, parts)
    M.close()


@pytest.mark.parametrize("n,doms,K", [(4, [5, 9], 3), (12, [], 10), (0, [4, 6], 2)])
def test_nb_predict_equals_the_references_own_function(n, doms, K):
    """nb_predict against ML::nb_impute itself (oracle/_ref) and the oracle restatement: the class LABEL of the largest
    product; identical wherever the two largest products differ by more than rounding."""
    ref_replay = _ref()
    rng = np.random.default_rng(3 * n + K)
    keys = [np.sort(rng.choice(np.arange(0, 3 * d), d, replace=False)).astype(np.int32) for d in doms]
    total = int(sum(doms))
    rows = 8_003
    num = [(rng.standard_normal(rows) * 2).astype(np.float32) for _ in range(n)]
    cat = [k[rng.integers(0, len(k), rows)].astype(np.int32) for k in keys]
    if doms:
        cat[0][::97] = 10_000  # a key the model does not hold: every class gets probability 0 -> labels[0]
    labels = (np.arange(K) * 3 + 2).tolist()
    p = oracle.nb_params(labels, rng.random(K) + 0.1, rng.standard_normal((K, n)), rng.random((K, n)) * 3 + 0.1, [k.tolist() for k in keys],
                         rng.random((K, total)) + 0.01)
    want = ref_replay.predict("nb_predict", p, [False], num, cat)
    mine, prob = oracle.nb_predict(p, num, cat)
    assert np.array_equal(want, mine)
    got = replay.glue().predict("nb_predict", p, [False], num, cat)
    top2 = np.sort(prob, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-12 * top2[:, 1]
    dead = top2[:, 1] == 0
    assert (clear | dead).mean() > 0.999 and (got[clear | dead] == want[clear | dead]).all()
    if doms:
        assert dead.any() and (got[dead] == labels[0]).all()


@pytest.mark.parametrize("n,doms,K,normalize", [(4, [3, 5], 3, False), (4, [3, 5], 3, True), (6, [], 4, True), (0, [4, 4], 2, False)])
def test_qda_predict_matches_the_restatement(n, doms, K, normalize):
    """qda_predict against the numpy restatement of ML::qda_impute (pinned to the reference in
    tests/test_predict_ref_cpu.py) and, where oracle/_ref is present and the keys are non-negative (the reference reads
    the parameter list's keys as FLOAT -> int), against the reference's own function."""
    rng = np.random.default_rng(11 * n + K + normalize)
    keys = [np.sort(rng.choice(np.arange(-4, 3 * d), d, replace=False)).astype(np.int32) for d in doms]
    P = n + int(sum(doms))
    rows = 5_001
    num = [rng.standard_normal(rows).astype(np.float32) for _ in range(n)]
    cat = [k[rng.integers(0, len(k), rows)].astype(np.int32) for k in keys]
    labels = (np.arange(K) + 20).tolist()
    p = oracle.qda_params(labels, rng.standard_normal((K, P, P)), rng.standard_normal((K, P)), rng.standard_normal(K),
                          [k.tolist() for k in keys], means=rng.standard_normal(P) if normalize else None)
    want, scores = oracle.qda_predict(p, normalize, num, cat)
    got = replay.glue().predict("qda_predict", p, [normalize], num, cat)
    top2 = np.sort(scores, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-9 * np.abs(top2).max(axis=1)
    assert clear.mean() > 0.999 and (got[clear] == want[clear]).all()
    from oracle import ref_replay
    if ref_replay.available() and all((k >= 0).all() for k in keys):
        ref = ref_replay.predict("qda_predict", p, [normalize], num, cat)
        assert (got[clear] == ref[clear]).all()
