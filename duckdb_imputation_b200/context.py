"""Python handle on one aggregate context (cfb_ctx) -- test / bench plumbing over the C ABI.

Nothing here computes: every method is one call into libcofactor_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native as nat
from .struct_result import result_arrays, arrays_to_struct


def _addr(x) -> Optional[int]:
    """Address of a numpy array / torch tensor / raw int (None -> NULL)."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return int(x.data_ptr())  # torch.Tensor


class CofactorContext:
    """One aggregate state: create -> append / scan_device ... -> finalize -> close."""

    def __init__(self, kind: int, n_num: int, n_cat: int, n_groups: int = 1, device: int = 0):
        self._h = C.c_void_p()
        nat.check(nat.lib().cfb_ctx_create(device, kind, n_num, n_cat, n_groups, C.byref(self._h)))
        self.kind, self.n, self.m, self.n_groups, self.device = kind, n_num, n_cat, n_groups, device

    # -- lifetime
    def close(self):
        if self._h:
            nat.lib().cfb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- input
    def set_cat_domain(self, lo: Sequence[int], hi: Sequence[int]):
        a = (C.c_int32 * max(1, self.m))(*lo)
        b = (C.c_int32 * max(1, self.m))(*hi)
        nat.check(nat.lib().cfb_ctx_set_cat_domain(self._h, a, b))

    def append(self, num_cols, cat_cols, group=None, num_sel=None, cat_sel=None, count=None):
        """Host columns (numpy float32 / int32), optional per-column uint32 selection vectors."""
        keep = []

        def prep(cols, dt):
            out = []
            for c in cols:
                c = np.ascontiguousarray(c, dtype=dt)
                keep.append(c)
                out.append(c.ctypes.data)
            return out

        np_ = prep(num_cols, np.float32)
        cp_ = prep(cat_cols, np.int32)
        if count is None:
            first = (num_sel[0] if num_sel and num_sel[0] is not None else None)
            if first is None and cat_sel and cat_sel[0] is not None:
                first = cat_sel[0]
            if first is not None:
                count = len(first)
            else:
                count = len(keep[0]) if keep else 0
        ns = cs = None
        if num_sel:
            ns = nat.ptr_array([None if k is None else k.ctypes.data for k in _keep_sel(num_sel, keep)])
        if cat_sel:
            cs = nat.ptr_array([None if k is None else k.ctypes.data for k in _keep_sel(cat_sel, keep)])
        g = None
        if group is not None:
            g = np.ascontiguousarray(group, dtype=np.uint32)
            keep.append(g)
        nat.check(nat.lib().cfb_ctx_append(self._h, nat.ptr_array(np_), ns, nat.ptr_array(cp_), cs,
                                           _addr(g), count))

    def scan_device(self, d_num_cols, d_cat_cols, n_rows: int, d_group=None, stream: int = 0):
        """Device-resident SoA columns (torch CUDA tensors or raw device addresses)."""
        nat.check(nat.lib().cfb_triple_device(
            self._h, nat.ptr_array([_addr(c) for c in d_num_cols]), nat.ptr_array([_addr(c) for c in d_cat_cols]),
            _addr(d_group), n_rows, stream or None))

    # -- output
    def sync(self):
        nat.check(nat.lib().cfb_ctx_sync(self._h))

    def combine(self, src: "CofactorContext"):
        nat.check(nat.lib().cfb_ctx_combine(self._h, src._h))

    def finalize_arrays(self, group: int = 0) -> dict:
        res = nat.Result()
        nat.check(nat.lib().cfb_ctx_finalize(self._h, group, C.byref(res)))
        try:
            return result_arrays(res)
        finally:
            nat.lib().cfb_result_free(C.byref(res))

    def finalize_result(self, group: int = 0) -> "ResultHandle":
        """The canonical result as an owned cfb_result (for cfb_result_combine / cfb_result_multiply)."""
        h = ResultHandle()
        nat.check(nat.lib().cfb_ctx_finalize(self._h, group, C.byref(h.res)))
        return h

    def finalize(self, group: int = 0, narrow: bool = True) -> dict:
        return arrays_to_struct(self.finalize_arrays(group), narrow)

    def last_scan_ms(self) -> float:
        return float(nat.lib().cfb_last_scan_ms(self._h))

    # -- multi-GPU partial exchange
    def allreduce(self, nccl_comm: int, stream: int = 0):
        """SUM-all-reduce the dense state in place over an ncclComm_t (cfb_ctx_allreduce): the exchange step of the
        path, inside the library.  Stream-ordered on `stream` (0 = the context's stream, synchronous)."""
        nat.check(nat.lib().cfb_ctx_allreduce(self._h, nccl_comm, stream or None))

    def partial_sizes(self):
        a, b = C.c_size_t(), C.c_size_t()
        nat.check(nat.lib().cfb_ctx_partial_sizes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def export_partial(self, d_f64, d_u64, stream: int = 0):
        nat.check(nat.lib().cfb_ctx_export_partial(self._h, _addr(d_f64), _addr(d_u64), stream or None))

    def import_partial(self, d_f64, d_u64, stream: int = 0):
        nat.check(nat.lib().cfb_ctx_import_partial(self._h, _addr(d_f64), _addr(d_u64), stream or None))


class ResultHandle:
    """An owned cfb_result: ring sum / difference of results through the C ABI (cfb_result_combine) -- the delta
    cofactors of a MICE loop (imputation/triple/sub.cpp:71-219): observed = total - nulls."""

    def __init__(self):
        self.res = nat.Result()

    def combine(self, other: "ResultHandle", sign: int, keep_zero_keys: bool = False) -> "ResultHandle":
        out = ResultHandle()
        nat.check(nat.lib().cfb_result_combine(C.byref(self.res), C.byref(other.res), sign, 1 if keep_zero_keys else 0, C.byref(out.res)))
        return out

    def impute_linear(self, model, target: int) -> "ResultHandle":
        """The cofactor of the same rows after numeric column `target` is overwritten by the linear model's
        predictions (predict.LinearModel, n_out = 1) -- closed form, no scan (cfb_result_impute_linear)."""
        out = ResultHandle()
        nat.check(nat.lib().cfb_result_impute_linear(C.byref(self.res), C.byref(model.c), target, C.byref(out.res)))
        return out

    def arrays(self) -> dict:
        return result_arrays(self.res)

    def close(self):
        nat.lib().cfb_result_free(C.byref(self.res))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _keep_sel(sels, keep):
    out = []
    for s in sels:
        if s is None:
            out.append(None)
        else:
            a = np.ascontiguousarray(s, np.uint32)
            keep.append(a)
            out.append(a)
    return out
