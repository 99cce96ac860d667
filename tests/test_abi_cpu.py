"""CPU suite for the boundary: the C-ABI library loads, exports every symbol that
include/cofactor_b200.h declares, and fails loudly (no fallback) without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from duckdb_imputation_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cofactor_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    l = nat.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/cofactor_b200.h but not exported"
    # and the binding table covers exactly the header
    assert sorted(nat.SIGNATURES) == names


def test_abi_version_and_result_layout():
    assert nat.lib().cfb_abi_version() == 2
    # cfb_result: 3 x int32 (+pad) then 14 eight-byte fields
    assert C.sizeof(nat.Result) == 16 + 14 * 8


def test_no_device_means_error_not_fallback():
    l = nat.lib()
    if l.cfb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    h = C.c_void_p()
    rc = l.cfb_ctx_create(0, nat.CFB_TRIPLE, 3, 0, 1, C.byref(h))
    assert rc == nat.CFB_ERR_NO_DEVICE and not h
    assert b"no CPU fallback" in l.cfb_last_error()
    with pytest.raises(nat.CofactorError):
        from duckdb_imputation_b200 import sum_to_triple
        sum_to_triple([np.ones(4, np.float32)], [])


def test_predict_entry_points_without_device():
    """The write-back entry points fail loudly too: no model upload, no scoring on the host."""
    l = nat.lib()
    if l.cfb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from duckdb_imputation_b200 import predict
    with pytest.raises(nat.CofactorError, match="no CPU fallback"):
        predict.LinearModel([0.0], [[1.0, 2.0]])
    assert l.cfb_predict_device(None, None, None, None, 4, 0, None, None) == nat.CFB_ERR_INVALID
    assert l.cfb_predict_host(None, None, None, None, None, 4, 0, None) == nat.CFB_ERR_INVALID
    l.cfb_model_destroy(None)


def test_argument_validation_without_device():
    l = nat.lib()
    h = C.c_void_p()
    assert l.cfb_ctx_create(0, 7, 3, 0, 1, C.byref(h)) == nat.CFB_ERR_INVALID
    assert l.cfb_ctx_create(0, 0, 33, 0, 1, C.byref(h)) == nat.CFB_ERR_INVALID
    assert l.cfb_ctx_create(0, 0, 3, 0, 0, C.byref(h)) == nat.CFB_ERR_INVALID
    assert l.cfb_ctx_destroy(None) == 0


def test_host_synth_is_deterministic_and_in_range():
    from duckdb_imputation_b200 import synth
    a = synth.uniform_f32(1000, 5)
    b = synth.uniform_f32(500, 5, first=500)
    assert np.array_equal(a[500:], b) and a.dtype == np.float32
    assert 0.0 <= a.min() and a.max() < 1.0 and abs(a.mean() - 0.5) < 0.05
    k = synth.int32(1000, 6, lo=-3, rng=7)
    assert k.min() >= -3 and k.max() <= 3 and len(np.unique(k)) == 7
