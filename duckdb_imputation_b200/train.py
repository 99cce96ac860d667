"""The trainers' moment matrix and solves on the device over the C ABI: cfb_sigma_*.

    sigma = Sigma.from_context(ctx)            # straight from the dense device state, no finalize
    sigma = Sigma.from_result(handle)          # from a finalized result (ResultHandle)
    fit = sigma.linreg_train(label=1, step_size=0.001, lam=0.0, max_iterations=10000)
    fit = Sigma.from_context(ctx, label_cat=0).lda_train(shrinkage=0.001)

`linreg_params` / `lda_params` lay the fit out as the FLOAT[] the reference's trainers emit (ML/regression.cpp:276-354,
ML/lda.cpp:335-385), i.e. what linreg_predict / lda_predict read."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


class Sigma:
    def __init__(self, handle, n_num: int, n_cat: int, label_cat: int):
        self.h = handle
        self.n_num, self.n_cat, self.label_cat = n_num, n_cat, label_cat
        p, c, v = C.c_int32(), C.c_int32(), C.c_int64()
        nat.check(nat.lib().cfb_sigma_shape(self.h, C.byref(p), C.byref(c), C.byref(v)))
        self.p, self.n_classes = p.value, c.value
        self.cat_array = np.zeros(v.value, np.int64)
        self.cat_vars_idxs = np.zeros(n_cat + 1, np.int32)
        nat.check(nat.lib().cfb_sigma_layout(self.h, self.cat_array.ctypes.data, self.cat_vars_idxs.ctypes.data))

    @classmethod
    def from_context(cls, ctx, group: int = 0, label_cat: int = -1, drop_first: bool = False) -> "Sigma":
        h = C.c_void_p()
        nat.check(nat.lib().cfb_sigma_from_ctx(ctx._h, group, label_cat, int(drop_first), C.byref(h)))
        return cls(h, ctx.n, ctx.m, label_cat)

    @classmethod
    def from_result(cls, result, device: int = 0, label_cat: int = -1, drop_first: bool = False) -> "Sigma":
        res = result.res if hasattr(result, "res") else result
        h = C.c_void_p()
        nat.check(nat.lib().cfb_sigma_from_result(device, C.byref(res), label_cat, int(drop_first), C.byref(h)))
        return cls(h, res.n_num, res.n_cat, label_cat)

    def matrix(self):
        """(sigma [p, p], class sums [n_classes, p]) as float64."""
        sig = np.zeros((self.p, self.p))
        sums = np.zeros((self.n_classes, self.p))
        nat.check(nat.lib().cfb_sigma_download(self.h, sig.ctypes.data, sums.ctypes.data if self.n_classes else None))
        return sig, sums

    def linreg_train(self, label: int, step_size: float, lam: float, max_iterations: int, normalize: bool = False) -> dict:
        coeff, means = np.zeros(self.p), np.zeros(self.p)
        var, it, prod = C.c_double(), C.c_int32(), C.c_int32()
        nat.check(nat.lib().cfb_sigma_linreg_train(self.h, label, step_size, lam, max_iterations, int(normalize),
                                                   coeff.ctypes.data, means.ctypes.data, C.byref(var), C.byref(it), C.byref(prod)))
        return {"label": label, "coeff": coeff, "means": means if normalize else None, "variance": var.value,
                "iterations": it.value, "products": prod.value}

    def lda_train(self, shrinkage: float, normalize: bool = False) -> dict:
        q = self.p - 1
        coef, icpt, means = np.zeros((self.n_classes, q)), np.zeros(self.n_classes), np.zeros(self.p)
        nat.check(nat.lib().cfb_sigma_lda_train(self.h, shrinkage, int(normalize), coef.ctypes.data, icpt.ctypes.data,
                                                means.ctypes.data))
        return {"coef": coef, "intercept": icpt, "means": means if normalize else None}

    def linreg_params(self, fit: dict, compute_variance: bool = True) -> np.ndarray:
        """The FLOAT[] of linreg_train (ML/regression.cpp:276-354)."""
        m, lab = self.n_cat, fit["label"] + 1
        out = [float(m)]
        if m > 0:
            out += [float(i) for i in self.cat_vars_idxs] + [float(np.uint64(k)) for k in self.cat_array]
        out += [fit["coeff"][i] for i in range(self.p) if i != lab]
        if fit["means"] is not None:
            out += [fit["means"][i] for i in range(1, self.p) if i != lab]
        if compute_variance:
            out.append(np.sqrt(fit["variance"]))
        return np.asarray(out, np.float32)

    def lda_params(self, fit: dict) -> np.ndarray:
        """The FLOAT[] of lda_train (ML/lda.cpp:335-385)."""
        m, lab, idx = self.n_cat, self.label_cat, self.cat_vars_idxs
        out = [float(self.n_classes), float(0 if m == 1 else m)]
        if self.p - 1 - self.n_num > 0:
            remove = 0
            for i in range(m + 1):
                if i == lab:
                    remove = int(idx[lab + 1] - idx[lab])
                    continue
                out.append(float(idx[i] - remove))
            out += [float(np.uint64(k)) for k in self.cat_array[:idx[lab]]] + [float(np.uint64(k)) for k in self.cat_array[idx[lab + 1]:]]
        out += [float(np.uint64(k)) for k in self.cat_array[idx[lab]:idx[lab + 1]]]
        out += [v for v in fit["coef"].reshape(-1)] + [v for v in fit["intercept"]]
        if fit["means"] is not None:
            out += [fit["means"][i + 1] for i in range(self.p - 1)]
        return np.asarray(out, np.float32)

    def close(self):
        if self.h:
            nat.lib().cfb_sigma_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
