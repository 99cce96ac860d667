"""Time the device trainers on one-hot expansions of growing width, next to the reference's own
trainers (oracle/_ref, host, one core) on the same triple:

    python tools/train_bench.py [rows] [max_iterations] [shape index]

Per shape: p (sigma is p x p fp64), ms to assemble sigma from the device state, ms and iterations of the ridge
gradient descent (cfb_sigma_linreg_train, one cooperative kernel), ms of the LDA solve (blocked Cholesky), and the
reference's linreg_train / lda_train wall time on the finalized STRUCT.  One JSON line per shape."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext
    from duckdb_imputation_b200.struct_result import arrays_to_struct
    from duckdb_imputation_b200.train import Sigma
    from oracle import ref_replay

    rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
    max_it = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    shapes = [(19, (10,) * 10), (10, (30,) * 10), (10, (100,) * 10), (10, (200,) * 16)]
    if len(sys.argv) > 3:  # one shape only (profiling)
        shapes = [shapes[int(sys.argv[3])]]
    from oracle import oracle
    rng = np.random.default_rng(3)
    for n, doms in shapes:
        m = len(doms)
        num = [torch.from_numpy(rng.standard_normal(rows).astype(np.float32)).cuda() for _ in range(n)]
        cat = [torch.from_numpy(rng.integers(0, d, rows).astype(np.int32)).cuda() for d in doms]
        w = [torch.from_numpy(rng.standard_normal(d).astype(np.float32)).cuda() for d in doms]
        num[0] = num[1] - 0.5 * num[2] + sum(w[k][cat[k].long()] for k in range(m)) + 0.1 * num[0]
        with CofactorContext(CFB_TRIPLE, n, m) as ctx:
            ctx.set_cat_domain([0] * m, [d - 1 for d in doms])
            ctx.scan_device(num, cat, rows)
            ctx.sync()
            out = {"n_num": n, "cat_domains": f"{m} x {doms[0]}", "rows": rows}
            t0 = time.perf_counter()
            s = Sigma.from_context(ctx)
            out["p"] = s.p
            out["sigma_from_state_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
            s.linreg_train(0, 0.001, 0.01, 2)  # warm-up: module load, shared-memory attribute
            t0 = time.perf_counter()
            fit = s.linreg_train(0, 0.001, 0.01, max_it)
            dt = time.perf_counter() - t0
            out["linreg_train_ms"] = round(dt * 1e3, 2)
            out["linreg_iterations"] = fit["iterations"]
            out["linreg_products"] = fit["products"]
            out["us_per_product"] = round(dt * 1e6 / max(1, fit["products"]), 2)
            t0 = time.perf_counter()
            one = s.linreg_train(0, 0.001, 0.01, 1)  # one gradient step: the fixed cost of a call
            fixed = time.perf_counter() - t0
            out["call_overhead_ms"] = round(fixed * 1e3, 3)
            out["us_per_product_steady"] = round((dt - fixed) * 1e6 / max(1, fit["products"] - one["products"]), 2)
            s.close()
            t0 = time.perf_counter()
            s = Sigma.from_context(ctx, label_cat=m - 1)
            out["sigma_with_class_sums_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
            s.lda_train(0.01)
            t0 = time.perf_counter()
            s.lda_train(0.01)
            out["lda_train_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            s.close()
            t0 = time.perf_counter()
            res = ctx.finalize_arrays()
            out["finalize_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        if ref_replay.available() and s.p <= 1100:
            t = arrays_to_struct(res)
            t0 = time.perf_counter()
            ref_replay.train("linreg_train", t, 0, 0.001, 0.01, max_it, True, False)
            dt = time.perf_counter() - t0
            out["reference_linreg_train_ms"] = round(dt * 1e3, 1)
            if s.p <= 400:  # the restatement walks the reference's path (same inputs, same arithmetic): its product count
                oracle.linreg_train(t, 0, 0.001, 0.01, max_it, True, False)
                out["reference_products"] = oracle.linreg_train.last_products
                out["reference_us_per_product"] = round(dt * 1e6 / oracle.linreg_train.last_products, 2)
            if ref_replay.lapack_available() and s.p <= 400:  # (p = 1011: 270 s, mostly its p^2 x classes debug lines)
                t0 = time.perf_counter()
                ref_replay.train("lda_train", t, m - 1, 0.01, False)
                out["reference_lda_train_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        print(json.dumps(out), flush=True)
        del num, cat
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
