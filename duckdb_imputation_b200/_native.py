"""ctypes binding of the C ABI in include/cofactor_b200.h (libcofactor_b200.so).

The library is the product: there is no Python or CPU fallback.  Importing this module
without the built library raises; calling a compute entry point without a CUDA device
returns CFB_ERR_NO_DEVICE, surfaced as CofactorError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CFB_LIB_PATH") or os.path.join(_HERE, "lib", "libcofactor_b200.so")  # override: A/B builds

CFB_TRIPLE, CFB_NB = 0, 1
CFB_OK, CFB_ERR_INVALID, CFB_ERR_NO_DEVICE, CFB_ERR_CUDA, CFB_ERR_OOM, CFB_ERR_DOMAIN, CFB_ERR_STATE = (
    0, -1, -2, -3, -4, -5, -6)
CFB_MAX_NUM = 32
CFB_MAX_CAT = 32


class CofactorError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"cofactor_b200 error {code}: {message}")
        self.code = code


class Result(C.Structure):
    """cfb_result / orc_result (identical layout)."""
    _fields_ = [
        ("kind", C.c_int32), ("n_num", C.c_int32), ("n_cat", C.c_int32),
        ("N", C.c_int64), ("n_quad", C.c_int64),
        ("lin", C.POINTER(C.c_double)), ("quad", C.POINTER(C.c_double)),
        ("total_keys", C.c_int64),
        ("cat_offsets", C.POINTER(C.c_int64)), ("cat_keys", C.POINTER(C.c_int32)),
        ("cat_counts", C.POINTER(C.c_int64)), ("numcat_sums", C.POINTER(C.c_double)),
        ("n_pair_lists", C.c_int64),
        ("pair_offsets", C.POINTER(C.c_int64)), ("pair_key1", C.POINTER(C.c_int32)),
        ("pair_key2", C.POINTER(C.c_int32)), ("pair_counts", C.POINTER(C.c_int64)),
    ]


# every symbol include/cofactor_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "cfb_abi_version": (C.c_int, []),
    "cfb_device_count": (C.c_int, []),
    "cfb_last_error": (C.c_char_p, []),
    "cfb_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "cfb_ctx_destroy": (C.c_int, [_P]),
    "cfb_ctx_set_cat_domain": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "cfb_ctx_append": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), _P, C.c_size_t]),
    "cfb_ctx_append_triples": (C.c_int, [_P, C.c_size_t] + [_P] * 13),
    "cfb_ctx_append_triples_slot": (C.c_int, [_P, C.c_int, C.c_size_t] + [_P] * 13),
    "cfb_triple_device": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), _P, C.c_size_t, _P]),
    "cfb_ctx_sync": (C.c_int, [_P]),
    "cfb_ctx_combine": (C.c_int, [_P, _P]),
    "cfb_ctx_combine_slots": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "cfb_ctx_finalize": (C.c_int, [_P, C.c_int, C.POINTER(Result)]),
    "cfb_result_free": (None, [C.POINTER(Result)]),
    "cfb_result_combine": (C.c_int, [C.POINTER(Result), C.POINTER(Result), C.c_int, C.c_int, C.POINTER(Result)]),
    "cfb_result_impute_linear": (C.c_int, [C.POINTER(Result), _P, C.c_int, C.POINTER(Result)]),
    "cfb_result_multiply": (C.c_int, [C.POINTER(Result), C.POINTER(Result), C.POINTER(Result)]),
    "cfb_model_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "cfb_model_destroy": (None, [_P]),
    "cfb_model_set_noise": (C.c_int, [_P, C.c_double, C.c_uint64, C.c_uint64]),
    "cfb_model_create_nb": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "cfb_model_create_qda": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "cfb_predict_device": (C.c_int, [_P, _P, _P, _P, C.c_size_t, C.c_int, _P, _P]),
    "cfb_predict_host": (C.c_int, [_P, _P, _P, _P, _P, C.c_size_t, C.c_int, _P]),
    "cfb_ctx_partial_sizes": (C.c_int, [_P, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "cfb_ctx_export_partial": (C.c_int, [_P, _P, _P, _P]),
    "cfb_ctx_import_partial": (C.c_int, [_P, _P, _P, _P]),
    "cfb_ctx_allreduce": (C.c_int, [_P, _P, _P]),
    "cfb_nccl_unique_id": (C.c_int, [_P]),
    "cfb_nccl_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "cfb_nccl_comm_destroy": (C.c_int, [_P]),
    "cfb_nccl_agree_domain": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, _P]),
    "cfb_cat_minmax_device": (C.c_int, [C.c_int, C.POINTER(_P), C.c_int, C.c_size_t,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "cfb_sigma_from_result": (C.c_int, [C.c_int, C.POINTER(Result), C.c_int, C.c_int, C.POINTER(_P)]),
    "cfb_sigma_from_ctx": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "cfb_sigma_destroy": (None, [_P]),
    "cfb_sigma_shape": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "cfb_sigma_layout": (C.c_int, [_P, _P, _P]),
    "cfb_sigma_download": (C.c_int, [_P, _P, _P]),
    "cfb_sigma_linreg_train": (C.c_int, [_P, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, _P, _P,
                                         C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "cfb_sigma_lda_train": (C.c_int, [_P, C.c_float, C.c_int, _P, _P, _P]),
    "cfb_gen_uniform_f32": (C.c_int, [C.c_int, _P, C.c_size_t, C.c_uint64, C.c_uint64, _P]),
    "cfb_gen_int32": (C.c_int, [C.c_int, _P, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int32, C.c_uint32, _P]),
    "cfb_kernel_launches": (C.c_uint64, []),
    "cfb_set_timing": (C.c_int, [C.c_int]),
    "cfb_last_scan_ms": (C.c_double, [_P]),
    "cfb_stage_isa": (C.c_char_p, []),
    "cfb_host_copy_ceiling": (C.c_double, [C.c_size_t, C.c_int, C.c_int, C.c_int]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libcofactor_b200.so (built by `make -C duckdb_imputation_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no fallback implementation)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != CFB_OK:
        raise CofactorError(rc, lib().cfb_last_error().decode("utf-8", "replace"))


def ptr_array(ptrs):
    """A C array of void* from a list of integer addresses (NULL for None)."""
    arr = (_P * max(1, len(ptrs)))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
