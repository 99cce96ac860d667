"""Small end-to-end exercise of every kernel, meant to run under compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from duckdb_imputation_b200 import CFB_NB, CFB_TRIPLE, CofactorContext, replay
from oracle import oracle
from tests.parity import assert_parity, assert_struct_parity

rng = np.random.default_rng(0)
rows = 20_003
num = [rng.random(rows).astype(np.float32) for _ in range(20)]
cat = [rng.integers(-2, 9, rows).astype(np.int32) for _ in range(3)]
grp = rng.integers(0, 4, rows).astype(np.int32)
dn = [torch.from_numpy(c).cuda() for c in num]
dc = [torch.from_numpy(c).cuda() for c in cat]
dg = torch.from_numpy(grp).cuda()
for kind, n, m in ((CFB_TRIPLE, 20, 0), (CFB_TRIPLE, 5, 3), (CFB_NB, 12, 2), (CFB_TRIPLE, 17, 1)):
    with CofactorContext(kind, n, m) as ctx:                      # gram + slab
        ctx.scan_device(dn[:n], dc[:m], rows)
        assert_parity(ctx.finalize_arrays(), oracle.aggregate_arrays(kind, num[:n], cat[:m])[0], what=f"{kind},{n},{m}")
    with CofactorContext(kind, n, m, n_groups=4) as ctx:           # group kernel + slab
        ctx.scan_device(dn[:n], dc[:m], rows, d_group=dg)
        ref = oracle.aggregate_arrays(kind, num[:n], cat[:m], group=grp, n_groups=4)
        for g in range(4):
            assert_parity(ctx.finalize_arrays(g), ref[g], what=f"group {g}")
with CofactorContext(CFB_TRIPLE, 3, 3) as a, CofactorContext(CFB_TRIPLE, 3, 3) as b:   # host feed, domain growth, combine
    a.append([c[:9000] for c in num[:3]], [c[:9000] for c in cat])
    b.append([c[9000:] for c in num[:3]], [c[9000:] + 40 for c in cat])
    a.combine(b)
    assert_parity(a.finalize_arrays(), oracle.aggregate_arrays(CFB_TRIPLE, num[:3], [np.concatenate([c[:9000], c[9000:] + 40]) for c in cat])[0])
big = [rng.integers(0, 50_000, rows).astype(np.int32) for _ in range(2)]                 # hashed pairs
with CofactorContext(CFB_TRIPLE, 2, 2) as ctx:
    ctx.append(num[:2], big)
    assert_parity(ctx.finalize_arrays(), oracle.aggregate_arrays(CFB_TRIPLE, num[:2], big)[0], what="hashed")
g = replay.glue()                                                                        # callbacks incl. lifted sums
assert_struct_parity(g.query(0, num[:4], cat, group_by=grp, threads=2, lifted=True)[1],
                     oracle.aggregate(0, num[:4], cat, group_by=grp)[1], what="lifted")
from duckdb_imputation_b200 import predict                                               # predict kernels
model = predict.LinearModel([0.5], rng.standard_normal((1, 5)), [np.arange(-2, 9)] * 3, rng.standard_normal((1, 33)))
out = torch.zeros(rows, dtype=torch.float32, device="cuda")
predict.predict_device(model, dn[:5], dc, rows, predict.SCORE, out, d_mask=(dg > 1).to(torch.int32))
lda = predict.LinearModel(rng.standard_normal(3), rng.standard_normal((3, 5)), [np.arange(-2, 9)] * 3, rng.standard_normal((3, 33)))
cls = predict.predict_host(lda, num[:5], cat, predict.ARGMAX)
torch.cuda.synchronize()
assert cls.min() >= 0 and cls.max() <= 2
# trainers: sigma from the state / a result, gradient descent (one CTA and several), Cholesky with panels
from duckdb_imputation_b200.struct_result import arrays_to_struct
from duckdb_imputation_b200.train import Sigma
wide = [rng.integers(0, 70, rows).astype(np.int32) for _ in range(3)]
with CofactorContext(CFB_TRIPLE, 4, 3) as ctx:
    ctx.set_cat_domain([0] * 3, [69] * 3)
    ctx.append(num[:4], wide)
    t = arrays_to_struct(oracle.aggregate_arrays(CFB_TRIPLE, num[:4], wide)[0], narrow=False)
    with Sigma.from_context(ctx) as s:                              # p = 215: several CTAs + grid barrier
        fit = s.linreg_train(1, 0.001, 0.01, 40)
        np.testing.assert_allclose(s.matrix()[0], oracle.build_sigma(t)[0], rtol=1e-5, atol=1e-3)
    with Sigma.from_context(ctx, label_cat=2) as s:                 # q = 144: three Cholesky panels
        got = s.lda_params(s.lda_train(0.05))
        want = oracle.lda_train(t, 2, 0.05, False)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * float(np.abs(want).max()))
    res = ctx.finalize_result()
    with Sigma.from_result(res, label_cat=-1, drop_first=True) as s:
        s.linreg_train(0, 0.001, 0.0, 5, normalize=True)
    res.close()
with CofactorContext(CFB_TRIPLE, 3, 1) as ctx:                      # p = 15: the single-CTA descent
    ctx.append(num[:3], cat[:1])
    with Sigma.from_context(ctx) as s:
        s.linreg_train(2, 0.001, 0.0, 30)
print("sanitize smoke ok")
