"""A GPU-resident MICE loop over a synthetic table (BASELINE config 5) -- what run_MICE_baseline
(imputation/algorithms/imputation_base.cpp:4-145) does with SQL, done here with the C ABI:

    for every iteration, for every column c with NULLs:
        cofactor = sum_to_triple(all columns) over the rows where c is observed       (cfb_triple_device, row filter)
        model    = train(cofactor, label = c)                                         (host, small matrices)
        c[NULL rows] = predict(model, the other columns)                              (cfb_predict_device, in place)

The trainers are out of scope of this repository (the small solves stay on the host): the tool closes
the loop with closed-form numpy solvers on the cofactor -- least squares for FLOAT columns, LDA (pooled covariance)
for INTEGER columns -- NOT the reference's gradient-descent linreg_train / LAPACK lda_train, and without the
stochastic noise term.  tests/mice_host.py holds the same loop on the host (oracle cofactors + numpy predictions):
the checker.

    python tools/mice_loop.py [rows] [iterations]        one JSON line per iteration + a summary line
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


# ------------------------------------------------------------------ host side: moments -> models
def moment_matrix(res):
    """Cofactor (struct_result.result_arrays form) -> the moment matrix over z = [1 | x_0..x_{n-1} | one-hot keys]
    (what ML/utils.cpp:176-310 builds as `sigma`), plus the feature layout."""
    n, m = res["n"], res["m"]
    offs, keys = res["cat_offsets"], res["cat_keys"]
    tk = len(keys)
    d = 1 + n + tk
    M = np.zeros((d, d))
    M[0, 0] = res["N"]
    M[0, 1:1 + n] = M[1:1 + n, 0] = res["lin"]
    iu = np.triu_indices(n)
    Q = np.zeros((n, n))
    Q[iu] = res["quad"]
    Q = Q + Q.T - np.diag(np.diag(Q))
    M[1:1 + n, 1:1 + n] = Q
    M[0, 1 + n:] = M[1 + n:, 0] = res["cat_counts"]
    if n:
        M[1:1 + n, 1 + n:] = res["numcat"]
        M[1 + n:, 1:1 + n] = res["numcat"].T
    p = 0
    for k in range(m):
        for l in range(k, m):
            lo, hi = res["pair_offsets"][p], res["pair_offsets"][p + 1]
            k1, k2, cnt = res["pair_key1"][lo:hi], res["pair_key2"][lo:hi], res["pair_counts"][lo:hi]
            r = 1 + n + offs[k] + np.searchsorted(keys[offs[k]:offs[k + 1]], k1)
            c = 1 + n + offs[l] + np.searchsorted(keys[offs[l]:offs[l + 1]], k2)
            M[r, c] = cnt
            M[c, r] = cnt
            p += 1
    return M


def train_linreg(res, label):
    """Least squares of numeric column `label` on [1 | other numerics | one-hot of every categorical column]."""
    n = res["n"]
    M = moment_matrix(res)
    feat = [i for i in range(M.shape[0]) if i != 1 + label]
    w = np.linalg.pinv(M[np.ix_(feat, feat)], rcond=1e-10, hermitian=True) @ M[feat, 1 + label]
    offs, keys = res["cat_offsets"], res["cat_keys"]
    return {"bias": np.array([w[0]]), "w_num": w[1:n][None, :], "keys": [keys[offs[c]:offs[c + 1]] for c in range(res["m"])],
            "w_cat": w[n:][None, :]}


def train_lda(res, label):
    """LDA (pooled covariance, class priors) of categorical column `label` on [numerics | one-hot of the other
    categorical columns]; scores are linear: w_k = Sigma^-1 mu_k, b_k = -mu_k.w_k / 2 + log prior."""
    n, m = res["n"], res["m"]
    M = moment_matrix(res)
    offs, keys = res["cat_offsets"], res["cat_keys"]
    lab = np.arange(1 + n + offs[label], 1 + n + offs[label + 1])
    feat = np.array([i for i in range(1, M.shape[0]) if i not in set(lab)])
    nk = M[0, lab]
    mu = M[np.ix_(lab, feat)] / np.maximum(nk, 1)[:, None]           # class means of the features
    Sw = M[np.ix_(feat, feat)] - (mu * nk[:, None]).T @ mu           # within-class scatter
    W = (np.linalg.pinv(Sw / max(res["N"] - len(lab), 1), rcond=1e-10, hermitian=True) @ mu.T).T
    b = -0.5 * np.einsum("kf,kf->k", mu, W) + np.log(np.maximum(nk, 1e-300) / res["N"])
    other = [c for c in range(m) if c != label]
    cat_w = W[:, n:]                                                  # columns follow `feat`: numerics, then the other cat columns in order
    return {"bias": b, "w_num": W[:, :n], "keys": [keys[offs[c]:offs[c + 1]] for c in other], "w_cat": cat_w,
            "classes": keys[offs[label]:offs[label + 1]]}


def predict_np(model, num_cols, cat_cols):
    """numpy scores of rows (fp64) -- the checker for cfb_predict_*."""
    rows = len(num_cols[0]) if num_cols else len(cat_cols[0])
    s = np.tile(model["bias"][None, :], (rows, 1))
    for i, x in enumerate(num_cols):
        s += np.asarray(x, np.float64)[:, None] * model["w_num"][:, i][None, :]
    off = 0
    for c, col in enumerate(cat_cols):
        k = model["keys"][c]
        pos = np.searchsorted(k, col)
        hit = (pos < len(k)) & (k[np.minimum(pos, len(k) - 1)] == col)
        s[hit] += model["w_cat"][:, off + pos[hit]].T
        off += len(k)
    return s


# ------------------------------------------------------------------ the device loop
def mice_gpu(d_num, d_cat, d_null_num, d_null_cat, iters, rows, domains=None, log=None):
    """Device loop.  d_num / d_cat: lists of torch CUDA tensors (modified in place); d_null_*: {column: int32 mask
    tensor, 1 = NULL}.  Returns per-iteration timings (ms): scan / train / predict."""
    import torch
    from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext, predict
    n, m = len(d_num), len(d_cat)
    slots = {("c", c): torch.where(msk != 0, -1, 0).to(torch.int32) for c, msk in d_null_cat.items()}
    slots.update({("n", c): torch.where(msk != 0, -1, 0).to(torch.int32) for c, msk in d_null_num.items()})
    out = []
    for it in range(iters):
        t = {"scan": 0.0, "train": 0.0, "predict": 0.0}
        for kind, cols in (("c", d_null_cat), ("n", d_null_num)):
            for c, msk in cols.items():
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                with CofactorContext(CFB_TRIPLE, n, m) as ctx:
                    if domains is not None:
                        ctx.set_cat_domain([d[0] for d in domains], [d[1] for d in domains])
                    ctx.scan_device(d_num, d_cat, rows, d_group=slots[(kind, c)])
                    res = ctx.finalize_arrays()
                t1 = time.perf_counter()
                if kind == "c":
                    mdl = train_lda(res, c)
                    lm = predict.LinearModel(mdl["bias"], mdl["w_num"], mdl["keys"], mdl["w_cat"])
                    t2 = time.perf_counter()
                    predict.predict_device(lm, d_num, [x for k, x in enumerate(d_cat) if k != c], rows, predict.ARGMAX, d_cat[c], d_mask=msk)
                else:
                    mdl = train_linreg(res, c)
                    lm = predict.LinearModel(mdl["bias"], mdl["w_num"], mdl["keys"], mdl["w_cat"])
                    t2 = time.perf_counter()
                    predict.predict_device(lm, [x for k, x in enumerate(d_num) if k != c], d_cat, rows, predict.SCORE, d_num[c], d_mask=msk)
                torch.cuda.synchronize()
                t3 = time.perf_counter()
                lm.close()
                t["scan"] += (t1 - t0) * 1e3
                t["train"] += (t2 - t1) * 1e3
                t["predict"] += (t3 - t2) * 1e3
        out.append(t)
        if log:
            log(it, t)
    return out


# ------------------------------------------------------------------ the device loop with delta cofactors
def partition_by_null_pattern(d_num, d_cat, d_null_num, d_null_cat):
    """Reorder the table (in place, once) so that the rows with the same NULL pattern over the imputed columns are
    contiguous -- what the reference's partition step does to the base table (imputation/partition.cpp).  Returns
    {(kind, column): [(lo, hi), ...]}: the row ranges where that column is NULL (each range 4-row aligned start is NOT
    guaranteed: ranges are cut to multiples of 4 rows, the few boundary rows are returned as `ragged` ranges too)."""
    import torch
    cols = [("n", c) for c in d_null_num] + [("c", c) for c in d_null_cat]
    masks = [d_null_num[c] if k == "n" else d_null_cat[c] for k, c in cols]
    code = torch.zeros_like(masks[0], dtype=torch.int64)
    for b, msk in enumerate(masks):
        code += (msk != 0).to(torch.int64) << b
    # every pattern's block is padded in the ORDER only: rows keep their identity, blocks start where the previous ends
    order = torch.argsort(code, stable=True)
    for t in list(d_num) + list(d_cat) + masks:
        t.copy_(t[order])
    counts = torch.bincount(code, minlength=1 << len(cols)).cpu().tolist()
    starts = np.concatenate([[0], np.cumsum(counts)])
    ranges = {}
    for b, key in enumerate(cols):
        rs = []
        for pat in range(1 << len(cols)):
            if (pat >> b) & 1 and counts[pat]:
                lo, hi = int(starts[pat]), int(starts[pat + 1])
                if rs and rs[-1][1] == lo:
                    rs[-1] = (rs[-1][0], hi)
                else:
                    rs.append((lo, hi))
        ranges[key] = rs
    ranges["patterns"] = {pat: (int(starts[pat]), int(starts[pat + 1])) for pat in range(1 << len(cols)) if counts[pat]}
    ranges["bit"] = {key: b for b, key in enumerate(cols)}
    return ranges, order


def _scan_ranges(ctx, d_num, d_cat, ranges):
    """Cofactor of the rows in `ranges`: device pointers must be 16-byte aligned, so a range [lo, hi) is scanned as the
    aligned middle plus (at most 3 + 3) boundary rows through a slot filter over the enclosing aligned window."""
    import torch
    for lo, hi in ranges:
        a_lo, a_hi = (lo + 3) // 4 * 4, hi
        if a_lo >= a_hi:
            a_lo = a_hi = lo
        if a_hi > a_lo:
            ctx.scan_device([t[a_lo:a_hi] for t in d_num], [t[a_lo:a_hi] for t in d_cat], a_hi - a_lo)
        if lo < a_lo:  # the ragged head: rows [lo, a_lo) inside the aligned window [a_lo - 4, a_lo)
            w = a_lo - 4
            slot = torch.full((4,), -1, dtype=torch.int32, device=d_num[0].device if d_num else d_cat[0].device)
            slot[lo - w:] = 0
            ctx.scan_device([t[w:a_lo] for t in d_num], [t[w:a_lo] for t in d_cat], 4, d_group=slot)


def mice_gpu_delta(d_num, d_cat, d_null_num, d_null_cat, iters, rows, domains=None, log=None, per_pattern=True, closed_form=False):
    """The loop with DELTA COFACTORS (the idea of imputation_low.cpp:85-110 and the Value-level subtract_triple /
    sum_triple helpers, imputation/triple/sub.cpp:71-219): the table is partitioned by NULL pattern once; the cofactor
    of the whole table is computed once and then maintained; per imputed column only its NULL rows (20 %) are scanned.

    per_pattern=False, the reference's scheme (two scans of the NULL rows per column):
        nulls    = cofactor(rows where c is NULL)                      one scan of the NULL ranges
        observed = total - nulls                                       cfb_result_combine(total, nulls, -1)
        model    = train(observed);  c[NULL rows] = predict(model)     contiguous ranges, no mask
        total    = observed + cofactor(rows where c is NULL)           second scan of the NULL ranges

    per_pattern=True (default) keeps one cofactor PER NULL PATTERN P (the partitions), total = sum over P: the first of
    the two scans disappears because nulls = sum of the kept cofactors of the patterns that contain c; after the
    write-back each of those partitions is rescanned once and its cofactor replaced (ONE scan of the NULL rows per
    column, the rest is ring arithmetic on small results).

    closed_form=True (with per_pattern): after a LINEAR-REGRESSION write-back not even that scan is needed -- every entry
    of a partition's new cofactor that involves the imputed column is theta . (a row of the partition's old sigma
    matrix), cfb_result_impute_linear; only classifier write-backs (LDA: the new keys are not linear in anything)
    rescan their NULL rows.

    Returns (per-iteration timings, the permutation applied to the rows, setup ms)."""
    import torch
    from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext, predict
    n, m = len(d_num), len(d_cat)
    ranges, order = partition_by_null_pattern(d_num, d_cat, d_null_num, d_null_cat)

    def cofactor(rs):
        with CofactorContext(CFB_TRIPLE, n, m) as ctx:
            if domains is not None:
                ctx.set_cat_domain([d[0] for d in domains], [d[1] for d in domains])
            _scan_ranges(ctx, d_num, d_cat, rs)
            return ctx.finalize_result()

    def ring_sum(handles):
        """(sum of the results, whether the caller owns it): one handle is returned as it is."""
        acc, owned = handles[0], False
        for h in handles[1:]:
            nxt = acc.combine(h, +1)
            if owned:
                acc.close()
            acc, owned = nxt, True
        return acc, owned

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    by_pattern = {}
    if per_pattern:
        by_pattern = {pat: cofactor([r]) for pat, r in ranges["patterns"].items()}
        total, owned = ring_sum(list(by_pattern.values()))
        if not owned:  # a single pattern: keep `total` a handle of its own
            only = total
            total = only.combine(only, +1).combine(only, -1)
    else:
        total = cofactor([(0, rows)])
    setup_ms = (time.perf_counter() - t0) * 1e3
    out = []
    for it in range(iters):
        t = {"scan": 0.0, "train": 0.0, "predict": 0.0}
        for kind, cols in (("c", d_null_cat), ("n", d_null_num)):
            for c in cols:
                rs = ranges[(kind, c)]
                pats = [p for p in by_pattern if (p >> ranges["bit"][(kind, c)]) & 1]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if per_pattern:
                    nulls, nulls_owned = ring_sum([by_pattern[p] for p in pats])
                else:
                    nulls, nulls_owned = cofactor(rs), True
                observed = total.combine(nulls, -1)
                res = observed.arrays()
                t1 = time.perf_counter()
                mdl = train_lda(res, c) if kind == "c" else train_linreg(res, c)
                lm = predict.LinearModel(mdl["bias"], mdl["w_num"], mdl["keys"], mdl["w_cat"])
                t2 = time.perf_counter()
                for lo, hi in rs:
                    if kind == "c":
                        predict.predict_device(lm, [x[lo:hi] for x in d_num], [x[lo:hi] for k, x in enumerate(d_cat) if k != c], hi - lo,
                                               predict.ARGMAX, d_cat[c][lo:hi])
                    else:
                        predict.predict_device(lm, [x[lo:hi] for k, x in enumerate(d_num) if k != c], [x[lo:hi] for x in d_cat], hi - lo,
                                               predict.SCORE, d_num[c][lo:hi])
                torch.cuda.synchronize()
                t3 = time.perf_counter()
                if per_pattern:
                    for p in pats:
                        old = by_pattern[p]
                        by_pattern[p] = old.impute_linear(lm, c) if closed_form and kind == "n" else cofactor([ranges["patterns"][p]])
                        old.close()
                    new_nulls, new_owned = ring_sum([by_pattern[p] for p in pats])
                else:
                    new_nulls, new_owned = cofactor(rs), True
                total2 = observed.combine(new_nulls, +1)
                for h, own in ((total, True), (nulls, nulls_owned), (observed, True), (new_nulls, new_owned)):
                    if own:
                        h.close()
                total = total2
                t4 = time.perf_counter()
                lm.close()
                t["scan"] += (t1 - t0 + t4 - t3) * 1e3
                t["train"] += (t2 - t1) * 1e3
                t["predict"] += (t3 - t2) * 1e3
        out.append(t)
        if log:
            log(it, t)
    total.close()
    for h in by_pattern.values():
        h.close()
    return out, order, setup_ms


def synthetic_table(rows, n=20, m=10, dom=10, null_num=(0, 1), null_cat=(0,), null_frac=0.2, seed=5):
    """Correlated columns (so that there is something to impute), NULL cells pre-filled with the column mean / mode
    as init_baseline does (partition.cpp:700-712)."""
    rng = np.random.default_rng(seed)
    latent = rng.standard_normal((rows, 3)).astype(np.float32)
    num = [(latent @ rng.standard_normal(3).astype(np.float32) + 0.3 * rng.standard_normal(rows).astype(np.float32)).astype(np.float32)
           for _ in range(n)]
    cat = []
    for _ in range(m):
        z = latent @ rng.standard_normal(3).astype(np.float32) + 0.5 * rng.standard_normal(rows).astype(np.float32)
        cat.append(np.clip(((z - z.min()) / (z.max() - z.min() + 1e-9) * dom).astype(np.int32), 0, dom - 1))
    masks_num = {c: rng.random(rows) < null_frac for c in null_num}
    masks_cat = {c: rng.random(rows) < null_frac for c in null_cat}
    truth = {("n", c): num[c].copy() for c in null_num}
    truth.update({("c", c): cat[c].copy() for c in null_cat})
    for c, msk in masks_num.items():
        num[c][msk] = num[c][~msk].mean()
    for c, msk in masks_cat.items():
        cat[c][msk] = np.bincount(cat[c][~msk]).argmax()
    return num, cat, masks_num, masks_cat, truth


def main():
    import torch
    rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rows -= rows % 4
    n, m, dom = 20, 10, 10
    # the table is generated in slices on the host (numpy) and moved over; 100 M rows x 30 columns = 12 GB
    d_num = [torch.empty(rows, dtype=torch.float32, device="cuda") for _ in range(n)]
    d_cat = [torch.empty(rows, dtype=torch.int32, device="cuda") for _ in range(m)]
    null_num, null_cat = (0, 1), (0,)
    d_nn = {c: torch.empty(rows, dtype=torch.int32, device="cuda") for c in null_num}
    d_nc = {c: torch.empty(rows, dtype=torch.int32, device="cuda") for c in null_cat}
    step = 5_000_000
    for lo in range(0, rows, step):
        k = min(step, rows - lo)
        num, cat, mn, mc, _ = synthetic_table(k, n, m, dom, null_num, null_cat, seed=5 + lo // step)
        for i in range(n):
            d_num[i][lo:lo + k] = torch.from_numpy(num[i])
        for i in range(m):
            d_cat[i][lo:lo + k] = torch.from_numpy(cat[i])
        for c in null_num:
            d_nn[c][lo:lo + k] = torch.from_numpy(mn[c].astype(np.int32))
        for c in null_cat:
            d_nc[c][lo:lo + k] = torch.from_numpy(mc[c].astype(np.int32))
    torch.cuda.synchronize()
    mice_gpu(d_num, d_cat, d_nn, d_nc, 1, rows, domains=[(0, dom - 1)] * m)  # warm-up iteration (also a real one)

    def log(it, t):
        print(json.dumps({"mice_iteration": it, "rows": rows, "columns": f"{n} FLOAT + {m} INT (domain {dom})",
                          "null_columns": len(null_num) + len(null_cat), "ms": {k: round(v, 2) for k, v in t.items()},
                          "ms_total": round(sum(t.values()), 2)}), flush=True)

    ts = mice_gpu(d_num, d_cat, d_nn, d_nc, iters, rows, domains=[(0, dom - 1)] * m, log=log)
    tot = sum(sum(t.values()) for t in ts)
    scans = iters * (len(null_num) + len(null_cat))
    # the same loop with delta cofactors: one partition + one full scan up front, then only the NULL rows per step
    for per_pattern, closed_form in ((False, False), (True, False), (True, True)):
        td, _, setup_ms = mice_gpu_delta(d_num, d_cat, d_nn, d_nc, iters, rows, domains=[(0, dom - 1)] * m, per_pattern=per_pattern,
                                         closed_form=closed_form)
        totd = sum(sum(t.values()) for t in td)
        what = ("one kept cofactor per NULL pattern, linear write-backs in closed form: NO scan for the numeric columns" if closed_form
                else "one kept cofactor per NULL pattern, ONE scan of the NULL rows per column" if per_pattern
                else "total - nulls, two scans of the NULL rows per column")
        print(json.dumps({"summary": f"MICE loop with delta cofactors ({what})", "rows": rows,
                          "iterations": iters, "ms_per_iteration": round(totd / iters, 2), "setup_full_scan_ms": round(setup_ms, 2),
                          "ms_per_column_scans": round(sum(t["scan"] for t in td) / scans, 2),
                          "ms_per_predict": round(sum(t["predict"] for t in td) / scans, 2),
                          "ms_per_train_host": round(sum(t["train"] for t in td) / scans, 2)}), flush=True)
    print(json.dumps({"summary": "MICE loop on one B200, table resident in HBM", "rows": rows, "iterations": iters,
                      "ms_per_iteration": round(tot / iters, 2), "cofactor_scans": scans,
                      "ms_per_scan": round(sum(t["scan"] for t in ts) / scans, 2),
                      "scan_rows_per_s": round(rows * scans / (sum(t["scan"] for t in ts) * 1e-3)),
                      "ms_per_predict": round(sum(t["predict"] for t in ts) / scans, 2),
                      "ms_per_train_host": round(sum(t["train"] for t in ts) / scans, 2)}), flush=True)


if __name__ == "__main__":
    main()
