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
