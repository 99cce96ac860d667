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
print("sanitize smoke ok")
