"""ctypes wrapper of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from duckdb_imputation_b200._native import Result, ptr_array
from duckdb_imputation_b200.struct_result import arrays_to_struct, result_arrays

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
TRIPLE, NB = 0, 1
EXACT, FAITHFUL = 0, 1
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        P = C.c_void_p
        l.orc_aggregate.restype = C.c_int
        l.orc_aggregate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(P), C.POINTER(P), P, C.c_int, P,
                                    C.c_size_t, C.c_size_t, C.c_int, C.POINTER(Result)]
        l.orc_sum_of_lifted.restype = C.c_int
        l.orc_sum_of_lifted.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(P), C.POINTER(P), P, C.c_int,
                                        C.c_size_t, C.POINTER(Result)]
        l.orc_result_add.restype = C.c_int
        l.orc_result_add.argtypes = [C.POINTER(Result)] * 3
        l.orc_result_free.restype = None
        l.orc_result_free.argtypes = [C.POINTER(Result)]
        l.orc_last_seconds.restype = C.c_double
        _lib = l
    return _lib


def _cols(cols, dt):
    keep = [np.ascontiguousarray(c, dtype=dt) for c in cols]
    return keep, ptr_array([k.ctypes.data for k in keep])


def aggregate_arrays(kind, num_cols, cat_cols, group=None, n_groups=1, sel=None, mode=EXACT, threads=1):
    """-> list (one per group) of numpy-form results (see struct_result.result_arrays)."""
    kn, pn = _cols(num_cols, np.float32)
    kc, pc = _cols(cat_cols, np.int32)
    g = None if group is None else np.ascontiguousarray(group, np.int32)
    s = None if sel is None else np.ascontiguousarray(sel, np.uint32)
    table_rows = len(kn[0]) if kn else (len(kc[0]) if kc else 0)
    rows = len(s) if s is not None else table_rows
    out = (Result * n_groups)()
    rc = lib().orc_aggregate(kind, mode, len(kn), len(kc), pn, pc, None if g is None else g.ctypes.data, n_groups,
                             None if s is None else s.ctypes.data, rows, table_rows, threads, out)
    if rc:
        raise ValueError("orc_aggregate: bad arguments")
    try:
        return [result_arrays(out[i]) for i in range(n_groups)]
    finally:
        for i in range(n_groups):
            lib().orc_result_free(C.byref(out[i]))


def aggregate(kind, num_cols, cat_cols, group_by=None, where=None, mode=EXACT, threads=1, narrow=True):
    """SQL-shaped front end: returns a STRUCT dict, or a list of them in ascending group-key order."""
    sel = None if where is None else np.nonzero(np.asarray(where))[0].astype(np.uint32)
    if group_by is None:
        return arrays_to_struct(aggregate_arrays(kind, num_cols, cat_cols, sel=sel, mode=mode, threads=threads)[0], narrow)
    gb = np.asarray(group_by)
    used = gb[sel] if sel is not None else gb
    labels = np.unique(used)
    slots = np.searchsorted(labels, gb).clip(0, max(0, len(labels) - 1)).astype(np.int32)
    res = aggregate_arrays(kind, num_cols, cat_cols, group=slots, n_groups=max(1, len(labels)), sel=sel, mode=mode,
                           threads=threads)
    return [arrays_to_struct(r, narrow) for r in res]


def sum_of_lifted(kind, num_cols, cat_cols, group=None, n_groups=1):
    kn, pn = _cols(num_cols, np.float32)
    kc, pc = _cols(cat_cols, np.int32)
    g = None if group is None else np.ascontiguousarray(group, np.int32)
    rows = len(kn[0]) if kn else (len(kc[0]) if kc else 0)
    out = (Result * n_groups)()
    rc = lib().orc_sum_of_lifted(kind, len(kn), len(kc), pn, pc, None if g is None else g.ctypes.data, n_groups, rows, out)
    if rc:
        raise ValueError("orc_sum_of_lifted: bad arguments")
    try:
        return [arrays_to_struct(result_arrays(out[i])) for i in range(n_groups)]
    finally:
        for i in range(n_groups):
            lib().orc_result_free(C.byref(out[i]))


def last_seconds() -> float:
    return float(lib().orc_last_seconds())


def multiply(a: dict, b: dict) -> dict:
    """Ring product of two result STRUCTs -- restatement of Triple::MultiplyFunction
    (duckdb_extension/src/triple/mul.cpp:17-611) and Triple::multiply_nb (mul_nb.cpp) on the
    Python form of the STRUCT (pure Python: the operands are per-group results, tens of values).
    fp32 arithmetic like the reference (float * int, float * float)."""
    f = np.float32
    la, lb = a.get("lin_agg", a.get("lin_num")), b.get("lin_agg", b.get("lin_num"))
    qa, qb = a.get("quad_agg", a.get("quad_num")), b.get("quad_agg", b.get("quad_num"))
    na, nb_, Na, Nb = len(la), len(lb), a["N"], b["N"]
    is_nb = "quad_cat" not in a
    out = {"N": Na * Nb}                                                           # mul.cpp:42-47
    out["lin_num"] = [float(f(x) * f(Nb)) for x in la] + [float(f(x) * f(Na)) for x in lb]   # :96-107
    if is_nb:                                                                      # mul_nb.cpp: diagonals only
        out["quad_num"] = [float(f(x) * f(Nb)) for x in qa] + [float(f(x) * f(Na)) for x in qb]
    else:
        quad, p = [], 0
        for j in range(na):                                                        # mul.cpp:262-283
            for _ in range(j, na):
                quad.append(float(f(qa[p]) * f(Nb)))
                p += 1
            quad += [float(f(la[j]) * f(x)) for x in lb]
        quad += [float(f(x) * f(Na)) for x in qb]                                  # :285-288
        out["quad_num"] = quad
    scale = lambda lst, N: [{**e, "value": float(f(e["value"]) * f(N))} for e in lst]
    lca, lcb = a["lin_cat"], b["lin_cat"]
    ma, mb = len(lca), len(lcb)
    out["lin_cat"] = [scale(l, Nb) for l in lca] + [scale(l, Na) for l in lcb]      # :184-217
    if is_nb:
        return out
    nca, ncb = a["quad_num_cat"], b["quad_num_cat"]
    nc = []
    for j in range(na):                                                            # :377-411
        nc += [scale(nca[j * ma + k], Nb) for k in range(ma)]
        nc += [[{"key": e["key"], "value": float(f(la[j]) * f(e["value"]))} for e in lcb[k]] for k in range(mb)]
    for j in range(nb_):                                                           # :414-445
        nc += [[{"key": e["key"], "value": float(f(lb[j]) * f(e["value"]))} for e in lca[k]] for k in range(ma)]
        nc += [scale(ncb[j * mb + k], Na) for k in range(mb)]
    out["quad_num_cat"] = nc
    cca, ccb = a["quad_cat"], b["quad_cat"]
    cc, p = [], 0
    for c1 in range(ma):                                                           # :546-581
        for _ in range(c1, ma):
            cc.append(scale(cca[p], Nb))
            p += 1
        for c2 in range(mb):
            cc.append([{"key1": x["key"], "key2": y["key"], "value": float(f(x["value"]) * f(y["value"]))}
                       for x in lca[c1] for y in lcb[c2]])
    cc += [scale(l, Na) for l in ccb]                                              # :583-597
    out["quad_cat"] = cc
    return out


# ------------------------------------------------------------------ predict (MICE write-back step)
def linreg_params(intercept, w_num, cat_keys, w_cat, means_num=None, means_cat=None, sigma=0.0):
    """The FLOAT[] that linreg_train emits, in the layout linreg_impute reads (ML/regression.cpp:424-435):
    [n_cat | idx_0..idx_{n_cat} | unique keys | intercept | w_num | w_cat | (means_num | means_cat) | sigma].
    cat_keys / w_cat / means_cat: one list per categorical column."""
    p = [float(len(cat_keys))]
    if cat_keys:
        idx = np.concatenate([[0], np.cumsum([len(k) for k in cat_keys])])
        p += [float(i) for i in idx] + [float(k) for col in cat_keys for k in col]
    p += [float(intercept)] + [float(w) for w in w_num] + [float(w) for col in w_cat for w in col]
    if means_num is not None:
        p += [float(v) for v in means_num] + [float(v) for col in (means_cat or []) for v in col]
    p.append(float(sigma))
    return np.asarray(p, np.float32)


def linreg_predict(params, normalize, num_cols, cat_cols):
    """Restatement of ML::linreg_impute (ML/regression.cpp:397-509), noise = false, row by row in fp64 like
    the reference.  A key the parameter list does not hold raises (the reference reads past the weights)."""
    p = np.asarray(params, np.float32)
    n, m = len(num_cols), len(cat_cols)
    rows = len(num_cols[0]) if n else len(cat_cols[0])
    n_cat = int(p[0])
    start, max_idx = 1 + n_cat, 0                                             # :428-435
    if n_cat > 0:
        max_idx = int(p[start])
        start += max_idx + 1
    out = np.full(rows, float(p[start]), np.float64)                           # :439 intercept
    for i in range(n):                                                        # :442-453
        x = np.asarray(num_cols[i], np.float32).astype(np.float64)
        if normalize:
            x = x - float(p[1 + n + max_idx + start + i])
        out += float(p[i + start + 1]) * x
    for i in range(m):                                                        # :455-493
        b, e = int(p[1 + i]), int(p[2 + i])
        keys = p[b + 2 + n_cat:e + 2 + n_cat].astype(np.int64)
        col = np.asarray(cat_cols[i], np.int64)
        pos = np.full(rows, -1, np.int64)
        for j, k in enumerate(keys):
            pos[(col == k) & (pos < 0)] = b + j
        if (pos < 0).any():
            raise ValueError("key not in the parameter list")
        w = p[start + n + 1:]
        if normalize:
            # (double)(w_j * (onehot_j - mean_j)): the product is a FLOAT product (regression.cpp:478-491), added in
            # key order
            mean = p[1 + 2 * n + max_idx + start:]
            for j in range(b, e):
                out += (w[j] * ((pos == j).astype(np.float32) - mean[j])).astype(np.float64)
        else:
            out += w[pos].astype(np.float64)
    return out.astype(np.float32)                                             # FLOAT result (:366-375)


def lda_params(labels, coef, intercept, cat_keys, means=None):
    """The FLOAT[] that lda_train emits, in the layout LDA_impute reads (ML/lda.cpp:450-500):
    [K | S | idx_0..idx_{S-1} | unique keys | labels | coef[K][n + total] | intercept[K] | (means[n + total])]."""
    K = len(labels)
    p = [float(K)]
    if cat_keys:
        idx = np.concatenate([[0], np.cumsum([len(k) for k in cat_keys])])
        p += [float(len(idx))] + [float(i) for i in idx] + [float(k) for col in cat_keys for k in col]
    else:
        p += [0.0]
    p += [float(l) for l in labels] + [float(v) for v in np.asarray(coef, np.float64).reshape(-1)] + [float(v) for v in intercept]
    if means is not None:
        p += [float(v) for v in means]
    return np.asarray(p, np.float32)


def lda_predict(params, normalize, num_cols, cat_cols):
    """Restatement of LDA_impute (ML/lda.cpp:421-590): returns the INDEX of the class with the largest score
    (:566-575), first index on ties; also returns the scores (for tie-aware comparisons)."""
    p = np.asarray(params, np.float32)
    n, m = len(num_cols), len(cat_cols)
    rows = len(num_cols[0]) if n else len(cat_cols[0])
    K, S = int(p[0]), int(p[1])
    off = 2
    idx = [int(v) for v in p[off:off + S]]
    off += S
    total = idx[-1] if S else 0
    keys = p[off:off + total].astype(np.int64)
    off += total + K                                                           # labels are not used (:475-480)
    num_params = n + total
    coef = p[off:off + K * num_params].astype(np.float64).reshape(K, num_params)  # :487-491
    off += K * num_params
    intercept = p[off:off + K].astype(np.float64)                              # float intercepts (:495-500)
    off += K
    feats = np.zeros((rows, num_params), np.float64)                           # :506-531
    for j in range(n):
        feats[:, j] = np.asarray(num_cols[j], np.float32).astype(np.float64)
    for j in range(m):
        col = np.asarray(cat_cols[j], np.int64)
        hit = np.zeros(rows, bool)
        for t in range(idx[j], idx[j + 1]):
            sel = (col == keys[t]) & ~hit
            feats[sel, n + t] = 1.0
            hit |= sel
        if not hit.all():
            raise ValueError("key not in the parameter list")
    if normalize:                                                             # :533-549
        feats -= p[off:off + num_params].astype(np.float64)[None, :]
    scores = feats @ coef.T + intercept[None, :]                               # dgemv (:560-562) + intercept
    return np.argmax(scores, axis=1).astype(np.int32), scores


def nb_params(labels, priors, means, variances, cat_keys, cat_probs):
    """The FLOAT[] that nb_train emits, in the layout ML::nb_impute reads (ML/naive_bayes.cpp:186-210):
    [K | S | idx_0..idx_{S-1} | unique keys | labels[K] | priors[K] | per class: (mean, variance) per numeric column,
    then one probability per (column, key)].  means / variances: [K][n]; cat_probs: [K][total keys]."""
    K = len(labels)
    p = [float(K)]
    if cat_keys:
        idx = np.concatenate([[0], np.cumsum([len(k) for k in cat_keys])])
        p += [float(len(idx))] + [float(i) for i in idx] + [float(k) for col in cat_keys for k in col]
    else:
        p += [0.0]
    p += [float(l) for l in labels] + [float(v) for v in priors]
    for k in range(K):
        for j in range(len(means[k]) if len(means) else 0):
            p += [float(means[k][j]), float(variances[k][j])]
        if cat_keys:
            p += [float(v) for v in cat_probs[k]]
    return np.asarray(p, np.float32)


def nb_predict(params, num_cols, cat_cols):
    """Restatement of ML::nb_impute (ML/naive_bayes.cpp:153-263): per row and class the product of the prior, one
    Gaussian density per numeric column (variance + 1e-9) and one probability per categorical column (0 for a key
    the model does not hold); the LABEL of the first class with the largest product (class 0 when every product is
    0).  fp64 like the reference; also returns the products."""
    p = np.asarray(params, np.float32)
    n, m = len(num_cols), len(cat_cols)
    rows = len(num_cols[0]) if n else len(cat_cols[0])
    K, S = int(p[0]), int(p[1])
    idx, keys, total = [], np.zeros(0, np.int64), 0
    if S > 0:
        idx = [int(v) for v in p[2:2 + S]]
        total = idx[-1]
        keys = p[2 + S:2 + S + total].astype(np.uint64).astype(np.int64)       # uint64_t cat_vars (:196-202)
    label_off = 2 + (S + total if S > 0 else 0)
    prior_off = label_off + K
    k0 = prior_off + K
    per_class = 2 * n + (idx[m] if S > 0 else 0)
    prob = np.zeros((rows, K), np.float64)
    for c in range(K):
        base = k0 + c * per_class
        tot = np.full(rows, np.float64(p[prior_off + c]))
        for j in range(n):
            var = np.float64(p[base + 2 * j + 1]) + 0.000000001
            mean = np.float64(p[base + 2 * j])
            x = np.asarray(num_cols[j], np.float32)
            # (in_cont_data - mean): float - double -> double
            tot = tot * ((1.0 / np.sqrt(2 * np.pi * var)) * np.exp(-np.power(x.astype(np.float64) - mean, 2) / (2.0 * var)))
        if S > 0:
            cb = base + 2 * n
            for j in range(m):
                col = np.asarray(cat_cols[j], np.int64)
                f = np.zeros(rows, np.float64)
                seen = np.zeros(rows, bool)
                for t in range(idx[j], idx[j + 1]):
                    sel = (col == keys[t]) & ~seen
                    f[sel] = np.float64(p[cb + t])
                    seen |= sel
                tot = tot * f
        prob[:, c] = tot
    best = np.zeros(rows, np.int64)
    mx = np.zeros(rows, np.float64)
    for c in range(K):                                                        # strict '>' from 0 (:247-250)
        better = prob[:, c] > mx
        best[better] = c
        mx[better] = prob[better, c]
    return p[label_off + best].astype(np.int32), prob


def qda_params(labels, quad, lin, intercept, cat_keys, means=None):
    """The FLOAT[] that qda_train emits, in the layout ML::qda_impute reads (ML/qda.cpp:367-400, :437-448):
    [K | S | idx | unique keys | labels[K] | per class: Q[p x p] column-major, l[p], intercept | (means[p])],
    p = n + total keys."""
    K = len(labels)
    out = [float(K)]
    if cat_keys:
        idx = np.concatenate([[0], np.cumsum([len(k) for k in cat_keys])])
        out += [float(len(idx))] + [float(i) for i in idx] + [float(k) for col in cat_keys for k in col]
    else:
        out += [0.0]
    out += [float(l) for l in labels]
    for k in range(K):
        out += [float(v) for v in np.asarray(quad[k], np.float64).reshape(-1, order="F")]
        out += [float(v) for v in lin[k]] + [float(intercept[k])]
    if means is not None:
        out += [float(v) for v in means]
    return np.asarray(out, np.float32)


def qda_predict(params, normalize, num_cols, cat_cols):
    """Restatement of ML::qda_impute (ML/qda.cpp:338-498): score_k = intercept_k + f^T Q_k f + l_k . f over the
    features f = [numeric | one-hot] (centred when normalize; a key the model does not hold leaves its one-hot
    block zero); the LABEL of the first class with the largest score.  Pinned to the reference's own ML::qda_impute
    (tests/test_predict_ref_cpu.py; oracle/Makefile compiles ML/qda.cpp through a one-cast sed stream because g++
    rejects qda.cpp:209 as written)."""
    p = np.asarray(params, np.float32)
    n, m = len(num_cols), len(cat_cols)
    rows = len(num_cols[0]) if n else len(cat_cols[0])
    K, S = int(p[0]), int(p[1])
    idx, keys, total, start = [], np.zeros(0, np.int64), 0, 2
    if S > 0:
        idx = [int(v) for v in p[2:2 + S]]
        total = idx[-1]
        keys = p[2 + S:2 + S + total].astype(np.uint64).astype(np.int64)
        start = 2 + total + S
    label_off = start
    start += K
    P = n + total
    feats = np.zeros((rows, P), np.float64)
    for j in range(n):
        feats[:, j] = np.asarray(num_cols[j], np.float32).astype(np.float64)
    for j in range(m):
        col = np.asarray(cat_cols[j], np.int64)
        seen = np.zeros(rows, bool)
        for t in range(idx[j], idx[j + 1]):
            sel = (col == keys[t]) & ~seen
            feats[sel, n + t] = 1.0
            seen |= sel
    if normalize:
        off = start + (P * P + P + 1) * K
        feats = feats - p[off:off + P].astype(np.float64)[None, :]
    scores = np.zeros((rows, K), np.float64)
    for k in range(K):
        base = start + k * (P * P + P + 1)
        Q = p[base:base + P * P].astype(np.float64).reshape(P, P, order="F")
        l = p[base + P * P:base + P * P + P].astype(np.float64)
        scores[:, k] = np.float64(p[base + P * P + P]) + np.einsum("ri,ij,rj->r", feats, Q, feats) + feats @ l
    return p[label_off + np.argmax(scores, axis=1)].astype(np.int32), scores


# ------------------------------------------------------------------ trainers (sigma matrix + solves)
def _ref_key_order(keys):
    """n_cols_1hot_expansion sorts a column's keys as uint64 (ML/utils.cpp:541-556): negative keys after the others."""
    return sorted(keys, key=lambda k: int(k) & 0xFFFFFFFFFFFFFFFF)


def _emit_key(k):
    """cat_array is uint64 (ML/utils.cpp:534): a negative key reaches the FLOAT[] as 2^64 + key."""
    return float(int(k) & 0xFFFFFFFFFFFFFFFF)


def one_hot_layout(s: dict, drop_first=False):
    """cat_array, cat_vars_idxs of n_cols_1hot_expansion (ML/utils.cpp:522-576) for a ring STRUCT dict."""
    cat_array, idxs = [], [0]
    for col in s["lin_cat"]:
        ks = _ref_key_order([e["key"] for e in col])
        if drop_first:
            ks = ks[1:]
        cat_array += ks
        idxs.append(len(cat_array))
    return cat_array, idxs


def build_sigma(s: dict, label_cat=-1, drop_first=False):
    """Restatement of build_sigma_matrix (ML/utils.cpp:176-310) over a ring STRUCT dict (values as the STRUCT holds
    them: FLOAT).  Returns (sigma [p, p] float64, cat_array, cat_vars_idxs).  With a categorical label the column is
    left out and the later columns move up (`skipped_var_categories`)."""
    n, m = len(s["lin_agg"]), len(s["lin_cat"])
    cat_array, idxs = one_hot_layout(s, drop_first)
    label_keys = idxs[label_cat + 1] - idxs[label_cat] if label_cat >= 0 else 0
    p = 1 + n + idxs[m] - label_keys
    sig = np.zeros((p, p))
    sig[0, 0] = s["N"]
    lin = np.asarray(s["lin_agg"], np.float64)
    sig[0, 1:n + 1] = lin
    sig[1:n + 1, 0] = lin
    q = np.asarray(s["quad_agg"], np.float64)
    for i in range(n):
        for j in range(i, n):
            sig[1 + i, 1 + j] = sig[1 + j, 1 + i] = q[i * n - i * (i + 1) // 2 + j]

    def index(k, key):
        if k == label_cat:
            return -1
        col = cat_array[idxs[k]:idxs[k + 1]]
        if key not in col:
            return -1                                            # dropped first key
        return 1 + n + idxs[k] + col.index(key) - (label_keys if label_cat >= 0 and k > label_cat else 0)

    for k in range(m):
        for e in s["lin_cat"][k]:
            i = index(k, e["key"])
            if i >= 0:
                sig[0, i] = sig[i, 0] = sig[i, i] = e["value"]
    for num in range(n):
        for k in range(m):
            for e in s["quad_num_cat"][num * m + k]:
                i = index(k, e["key"])
                if i >= 0:
                    sig[i, 1 + num] = sig[1 + num, i] = e["value"]
    t = 0
    for k in range(m):
        for l in range(k, m):
            for e in s["quad_cat"][t]:
                a, b = index(k, e["key1"]), index(l, e["key2"])
                if a >= 0 and b >= 0:
                    sig[a, b] = sig[b, a] = e["value"]
            t += 1
    return sig, cat_array, idxs


def standardize_sigma(sig):
    """standardize_sigma (ML/utils.cpp:580-599); returns (sigma', means, stds)."""
    N = sig[0, 0]
    means = sig[0] / N
    with np.errstate(invalid="ignore", divide="ignore"):
        stds = np.sqrt(np.diag(sig) / N - (sig[0] / N) ** 2)
        out = sig.copy()
        out[1:, 1:] = (sig[1:, 1:] - np.outer(means[1:], sig[0, 1:]) - np.outer(sig[0, 1:], means[1:])
                       + N * np.outer(means[1:], means[1:])) / np.outer(stds[1:], stds[1:])
    out[0, 1:] = 0
    out[1:, 0] = 0
    return out, means, stds


def linreg_train(s: dict, label, step_size, lam, max_iterations, compute_variance, normalize):
    """Restatement of ML::ridge_linear_regression (ML/regression.cpp:113-356): batch gradient descent with
    Barzilai-Borwein steps and backtracking line search on the sigma matrix; step_size and lambda are FLOATs.
    Returns the FLOAT[] parameter list linreg_predict reads."""
    f32 = np.float32
    n, m = len(s["lin_agg"]), len(s["lin_cat"])
    sig, cat_array, idxs = build_sigma(s)
    p = sig.shape[0]
    means = stds = None
    if normalize:
        sig, means, stds = standardize_sigma(sig)
    step, lam = f32(step_size), f32(lam)
    N = sig[0, 0]
    theta = np.zeros(p)
    lab = label + 1
    theta[lab] = -1
    prev_theta = theta.copy()

    def gradient(th):
        g = sig @ th / N if N != 0 else np.zeros(p)
        g[lab] = 0
        return g

    def error(th):
        if N == 0:
            return 0.0
        return (th @ (sig @ th) / N + float(lam) * (np.sum(th[1:] ** 2) - 1)) / 2

    products = 1                                                 # Sigma * theta products spent (bench bookkeeping)
    grad = gradient(theta)
    upd = grad + float(lam) * theta
    upd[0] = grad[0]
    first_norm = np.sqrt(np.sum(upd ** 2) - float(lam) * float(lam))
    prev_error = error(theta)
    it = 1
    while True:
        update = grad + float(lam) * theta
        update[0] = grad[0]
        sq = np.sum(update ** 2)
        prev_theta, prev_grad = theta.copy(), grad.copy()
        theta = theta - float(step) * update
        theta[lab] = -1
        gnorm = sq - float(lam) * float(lam)
        dparam = float(step) * np.sqrt(sq)
        err = error(theta)
        products += 1
        bt = 0
        while err > prev_error - float(step / f32(2)) * gnorm and bt < 500:
            step = step / f32(2)
            newp = prev_theta - float(step) * update
            dparam = np.sqrt(np.sum((theta - newp) ** 2))
            theta = newp
            theta[lab] = -1
            err = error(theta)
            products += 1
            bt += 1
        if dparam < 1e-20 or np.sqrt(gnorm) / (first_norm + 0.001) < 1e-8:
            break
        grad = gradient(theta)
        pd, gd = theta - prev_theta, grad - prev_grad
        dss, gss, dgs = np.sum(pd * pd), np.sum(gd * gd), np.sum(pd * gd)
        if dgs != 0 and gss != 0:
            ts, tm = dss / dgs, dgs / gss
            if not (tm < 0 or ts < 0):
                step = f32(tm if tm / ts > 0.5 else ts - 0.5 * tm)
        prev_error = err
        it += 1
        if not it < max_iterations:
            break
    variance = theta @ (sig @ theta) / N if compute_variance else None
    if normalize:
        theta[1:] = theta[1:] / stds[1:] * stds[lab]
        theta[0] = theta[0] * stds[lab] + means[lab]
    out = [float(m)]
    if m > 0:
        out += [float(i) for i in idxs] + [_emit_key(k) for k in cat_array]
    out += [theta[i] for i in range(p) if i != lab]
    if normalize:
        out += [means[i] for i in range(1, p) if i != lab]
    if compute_variance:
        out.append(np.sqrt(variance))
    linreg_train.last_products = products
    return np.asarray(out, np.float32), it


def lda_train(s: dict, label, shrinkage, normalize, unshifted_sums=False):
    """Restatement of lda_train (ML/lda.cpp:154-410).  The class sums of the OTHER categorical columns are placed at
    their sigma positions.  unshifted_sums=True reproduces the reference where it differs: build_sum_vector adds the
    UN-shifted cat_array index (lda.cpp:131) although the sigma matrix moved the columns behind the label up by the
    label's key count, so for those columns the sums land `label keys` slots too far right (into the next class's
    row of the flat array; past its end for the last class -- dropped here).  dgelsd -> numpy lstsq."""
    n, m = len(s["lin_agg"]), len(s["lin_cat"])
    sig, cat_array, idxs = build_sigma(s, label_cat=label)
    p = sig.shape[0]
    classes = cat_array[idxs[label]:idxs[label + 1]]
    C = len(classes)
    label_keys = C
    sums = np.zeros((C, p))
    for e in s["lin_cat"][label]:
        sums[classes.index(e["key"]), 0] = e["value"]
    for num in range(n):
        for e in s["quad_num_cat"][num * m + label]:
            sums[classes.index(e["key"]), 1 + num] = e["value"]

    def index(k, key):
        col = cat_array[idxs[k]:idxs[k + 1]]
        return 1 + n + idxs[k] + col.index(key) - (label_keys if k > label and not unshifted_sums else 0)

    flat = sums.reshape(-1)
    t = 0
    for k in range(m):
        for l in range(k, m):
            if k != l and (k == label or l == label):
                for e in s["quad_cat"][t]:
                    at = (classes.index(e["key1"]) * p + index(l, e["key2"]) if k == label
                          else classes.index(e["key2"]) * p + index(k, e["key1"]))
                    if at < flat.size:
                        flat[at] = e["value"]
            t += 1
    means = stds = None
    if normalize:
        sig, means, stds = standardize_sigma(sig)
        sums[:, 1:] = (sums[:, 1:] - means[None, 1:] * sums[:, :1]) / stds[None, 1:]
    S = sig[1:, 1:].copy()
    q = p - 1
    for c in range(C):
        S -= np.outer(sums[c, 1:], sums[c, 1:]) / sums[c, 0]
    mean_vec = sums[:, 1:] / sums[:, :1]                        # [C, q]
    sh = np.float32(shrinkage)
    mu = np.trace(S) / np.float32(q)
    S = S * float(np.float32(1) - sh)
    S[np.diag_indices(q)] += float(sh) * mu
    S /= float(s["N"])
    coef = np.linalg.lstsq(S, mean_vec.T, rcond=None)[0].T       # [C, q]
    intercept = -0.5 * np.sum(mean_vec * coef, axis=1) + np.log(sums[:, 0] / float(s["N"]))
    if normalize:
        coef = coef / stds[None, 1:]
    out = [float(C), float(0 if m == 1 else m)]
    if q - n > 0:
        remove = 0
        for i in range(m + 1):
            if i == label:
                remove = label_keys
                continue
            out.append(float(idxs[i] - remove))
        out += [_emit_key(k) for k in cat_array[:idxs[label]]] + [_emit_key(k) for k in cat_array[idxs[label + 1]:]]
    out += [_emit_key(k) for k in classes] + [v for v in coef.reshape(-1)] + [v for v in intercept]
    if normalize:
        out += [means[i + 1] for i in range(q)]
    return np.asarray(out, np.float32)
