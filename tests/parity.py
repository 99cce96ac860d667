"""Comparison of a GPU result with the oracle's, in the numpy form of struct_result.result_arrays.

The bar: N, keys and every categorical count bit-exact; floating-point
sums within RTOL = 1e-5 relative of the fp64 oracle (with an absolute floor scaled by the sum of
magnitudes, for sums that cancel to ~0)."""
import numpy as np

RTOL = 1e-5


def assert_parity(got: dict, ref: dict, rtol: float = RTOL, abs_scale=None, what=""):
    assert got["kind"] == ref["kind"] and got["n"] == ref["n"] and got["m"] == ref["m"], what
    assert got["N"] == ref["N"], f"{what}: N {got['N']} != {ref['N']}"
    for k in ("cat_offsets", "cat_keys", "cat_counts"):
        assert np.array_equal(got[k], ref[k]), f"{what}: {k} differs"
    if ref["kind"] == 0:
        for k in ("pair_offsets", "pair_key1", "pair_key2", "pair_counts"):
            assert np.array_equal(got[k], ref[k]), f"{what}: {k} differs"
    fields = ["lin", "quad"] + (["numcat"] if ref["kind"] == 0 else [])
    for k in fields:
        g, r = np.asarray(got[k], np.float64), np.asarray(ref[k], np.float64)
        assert g.shape == r.shape, f"{what}: {k} shape {g.shape} != {r.shape}"
        if g.size == 0:
            continue
        floor = 0.0 if abs_scale is None else rtol * abs_scale
        err = np.abs(g - r)
        tol = rtol * np.abs(r) + floor
        bad = err > tol
        assert not bad.any(), (f"{what}: {k} off at {np.argwhere(bad)[:4].tolist()}: got {g[bad][:4]} ref {r[bad][:4]} "
                               f"max rel {np.max(err / np.maximum(np.abs(r), 1e-300)):.3e}")


def max_rel_err(got: dict, ref: dict) -> float:
    worst = 0.0
    for k in ("lin", "quad", "numcat"):
        if k in ref and np.asarray(ref[k]).size:
            g, r = np.asarray(got[k], np.float64), np.asarray(ref[k], np.float64)
            nz = np.abs(r) > 0
            if nz.any():
                worst = max(worst, float(np.max(np.abs(g - r)[nz] / np.abs(r)[nz])))
    return worst


def assert_struct_parity(got: dict, ref: dict, rtol: float = RTOL, abs_floor: float = 0.0, what=""):
    """Same bar on two result STRUCT dicts (the Python form of the DuckDB value): N, every key and
    every count exact; lin_agg / quad_agg / quad_num_cat values within rtol."""
    assert list(got.keys()) == list(ref.keys()), f"{what}: fields {list(got.keys())} != {list(ref.keys())}"
    assert got["N"] == ref["N"], what
    assert got["lin_cat"] == ref["lin_cat"], f"{what}: lin_cat differs"
    if "quad_cat" in ref:
        assert got["quad_cat"] == ref["quad_cat"], f"{what}: quad_cat differs"

    def close(a, b, where):
        assert abs(a - b) <= rtol * abs(b) + abs_floor, f"{what}: {where}: {a} vs {b}"

    for f in ("lin_agg", "quad_agg"):
        assert len(got[f]) == len(ref[f]), f"{what}: {f} length"
        for i, (a, b) in enumerate(zip(got[f], ref[f])):
            close(a, b, f"{f}[{i}]")
    if "quad_num_cat" in ref:
        assert len(got["quad_num_cat"]) == len(ref["quad_num_cat"])
        for li, (la, lb) in enumerate(zip(got["quad_num_cat"], ref["quad_num_cat"])):
            assert [e["key"] for e in la] == [e["key"] for e in lb], f"{what}: quad_num_cat[{li}] keys"
            for a, b in zip(la, lb):
                close(a["value"], b["value"], f"quad_num_cat[{li}][{a['key']}]")


def to_result(a: dict):
    """numpy form (struct_result.result_arrays) -> a ctypes cfb_result that borrows the arrays; returns (result, keepalive)."""
    import ctypes as C

    from duckdb_imputation_b200._native import Result
    keep = []

    def ptr(x, dt, ct):
        arr = np.ascontiguousarray(x, dtype=dt)
        keep.append(arr)
        return arr.ctypes.data_as(C.POINTER(ct))

    r = Result()
    r.kind, r.n_num, r.n_cat, r.N = a["kind"], a["n"], a["m"], a["N"]
    r.n_quad = len(a["quad"])
    r.lin, r.quad = ptr(a["lin"], np.float64, C.c_double), ptr(a["quad"], np.float64, C.c_double)
    r.total_keys = len(a["cat_keys"])
    r.cat_offsets = ptr(a["cat_offsets"], np.int64, C.c_int64)
    r.cat_keys = ptr(a["cat_keys"], np.int32, C.c_int32)
    r.cat_counts = ptr(a["cat_counts"], np.int64, C.c_int64)
    if a["kind"] == 0:
        r.numcat_sums = ptr(a["numcat"].reshape(-1), np.float64, C.c_double)
        r.n_pair_lists = len(a["pair_offsets"]) - 1
        r.pair_offsets = ptr(a["pair_offsets"], np.int64, C.c_int64)
        r.pair_key1, r.pair_key2 = ptr(a["pair_key1"], np.int32, C.c_int32), ptr(a["pair_key2"], np.int32, C.c_int32)
        r.pair_counts = ptr(a["pair_counts"], np.int64, C.c_int64)
    else:
        r.pair_offsets = ptr(np.zeros(1), np.int64, C.c_int64)
    return r, keep
