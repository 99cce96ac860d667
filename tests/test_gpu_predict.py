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


def test_noise_and_bad_parameter_lists_are_query_errors():
    p = oracle.linreg_params(1.0, [2.0], [], [])
    x = [np.ones(10, np.float32)]
    with pytest.raises(replay.ReplayError, match="noise"):
        replay.glue().predict("linreg_predict", p, [True, False], x, [])
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
