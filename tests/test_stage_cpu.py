"""CPU suite: the host half of the feed (stage_copy.cpp) -- non-temporal staging copies, selection gathers and the key
range scan -- in every vector ISA the CPU offers (CFB_STAGE_ISA picks one at load time, so each runs in its own
process).  Pure host code: no device involved."""
import subprocess
import sys

import pytest

SCRIPT = r"""
import ctypes as C, numpy as np, sys
from duckdb_imputation_b200 import _native as nat
l = C.CDLL(nat.LIB_PATH)
l.cfb_stage_isa.restype = C.c_char_p
l.cfb_stage_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
l.cfb_stage_gather32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
l.cfb_stage_minmax32.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
rng = np.random.default_rng(3)
for n in (0, 1, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 2048, 2049, 10007):
    a = rng.integers(-2**31, 2**31 - 1, n, dtype=np.int64).astype(np.int32)
    lo, hi = C.c_int32(2**31 - 1), C.c_int32(-2**31)
    l.cfb_stage_minmax32(a.ctypes.data, n, C.byref(lo), C.byref(hi))
    assert (lo.value, hi.value) == ((int(a.min()), int(a.max())) if n else (2**31 - 1, -2**31)), n
    lo, hi = C.c_int32(-5), C.c_int32(7)      # running range: only ever widened
    b = (a % 4).astype(np.int32)
    l.cfb_stage_minmax32(b.ctypes.data, n, C.byref(lo), C.byref(hi))
    assert (lo.value, hi.value) == (-5, 7)
    for shift in (0, 4, 12, 60):              # unaligned destinations
        dst = np.zeros(n + 32, np.int32)
        l.cfb_stage_copy(dst.ctypes.data + shift, a.ctypes.data, 4 * n)
        assert np.array_equal(dst.view(np.uint8)[shift:shift + 4 * n].view(np.int32), a), (n, shift)
    if n:
        sel = rng.integers(0, n, 3 * n + 1).astype(np.uint32)
        out = np.zeros(len(sel), np.int32)
        l.cfb_stage_gather32(out.ctypes.data, a.ctypes.data, sel.ctypes.data, len(sel))
        assert np.array_equal(out, a[sel])
print(l.cfb_stage_isa().decode())
"""


@pytest.mark.parametrize("isa", ["sse2", "avx2", "avx512"])
def test_staging_primitives_in_every_isa(isa):
    import os
    env = dict(os.environ, CFB_STAGE_ISA=isa)
    out = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert out.returncode == 0, out.stderr[-2000:]
    got = out.stdout.strip().splitlines()[-1]
    assert got in ("sse2", "avx2", "avx512")  # the CPU may not offer the one asked for: it then runs the next one down
    if got != isa:
        pytest.skip(f"this CPU runs {got} when asked for {isa}")
