"""ctypes front end of the model-score entry points (cfb_predict_device / cfb_predict_host) -- test / bench
plumbing over the C ABI; nothing here computes."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat

SCORE, ARGMAX = 0, 1


class _Model(C.Structure):
    _fields_ = [("n_num", C.c_int32), ("n_cat", C.c_int32), ("n_out", C.c_int32), ("bias", C.c_void_p), ("w_num", C.c_void_p),
                ("cat_offsets", C.c_void_p), ("cat_keys", C.c_void_p), ("w_cat", C.c_void_p)]


class LinearModel:
    """score_o = bias[o] + w_num[o] . x + sum_c w_cat[o][position of key_c]  (include/cofactor_b200.h: cfb_linear_model)."""

    def __init__(self, bias, w_num, cat_keys=(), w_cat=None, device: int = 0, upload: bool = True):
        self.bias = np.ascontiguousarray(np.atleast_1d(bias), np.float64)
        K = len(self.bias)
        self.w_num = np.ascontiguousarray(np.asarray(w_num, np.float64).reshape(K, -1))
        self.n = self.w_num.shape[1]
        self.keys = [np.asarray(k, np.int32) for k in cat_keys]
        self.offs = np.ascontiguousarray(np.concatenate([[0], np.cumsum([len(k) for k in self.keys])]), np.int64)
        total = int(self.offs[-1])
        self.flat_keys = np.ascontiguousarray(np.concatenate(self.keys) if self.keys else np.zeros(0), np.int32)
        self.w_cat = np.ascontiguousarray(np.zeros((K, total)) if w_cat is None else np.asarray(w_cat, np.float64).reshape(K, total))
        self.c = _Model(self.n, len(self.keys), K, self.bias.ctypes.data, self.w_num.ctypes.data, self.offs.ctypes.data,
                        self.flat_keys.ctypes.data, self.w_cat.ctypes.data)
        self.device = device
        self._h = C.c_void_p()
        if upload:  # (upload=False: only the host description, e.g. for cfb_result_impute_linear)
            nat.check(nat.lib().cfb_model_create(device, C.byref(self.c), C.byref(self._h)))  # upload once

    def close(self):
        if self._h:
            nat.lib().cfb_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def predict_device(model: LinearModel, d_num, d_cat, rows: int, mode: int, d_out, d_mask=None, stream: int = 0):
    """Device-resident torch tensors (or raw addresses) in, d_out written in place (asynchronous on `stream`)."""
    addr = lambda t: None if t is None else (t if isinstance(t, int) else int(t.data_ptr()))
    nat.check(nat.lib().cfb_predict_device(model._h, nat.ptr_array([addr(t) for t in d_num]),
                                           nat.ptr_array([addr(t) for t in d_cat]), addr(d_mask), rows, mode, addr(d_out),
                                           stream or None))


def predict_host(model: LinearModel, num_cols, cat_cols, mode: int):
    kn = [np.ascontiguousarray(c, np.float32) for c in num_cols]
    kc = [np.ascontiguousarray(c, np.int32) for c in cat_cols]
    rows = len(kn[0]) if kn else len(kc[0])
    out = np.zeros(rows, np.float32 if mode == SCORE else np.int32)
    nat.check(nat.lib().cfb_predict_host(model._h, nat.ptr_array([k.ctypes.data for k in kn]), None,
                                         nat.ptr_array([k.ctypes.data for k in kc]), None, rows, mode, out.ctypes.data))
    return out
