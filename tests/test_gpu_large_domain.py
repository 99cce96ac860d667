"""GPU suite: the sparse (hashed) pair-count fallback for large categorical domains
(pair_hash.cuh) -- the GPU twin of the reference's std::map<std::pair<int,int>, float>."""
import os
import subprocess
import sys

import numpy as np
import pytest

from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext, CofactorError
from duckdb_imputation_b200 import _native as nat
from oracle import oracle
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_domain_1e5_uses_the_hash_fallback_and_matches_oracle():
    rng = np.random.default_rng(77)
    rows = 300_000
    num = [rng.random(rows).astype(np.float32) for _ in range(3)]
    # 3 columns x 10^5 keys: dense pair tables would need 3 x 10^10 counters
    cat = [rng.integers(0, 100_000, rows).astype(np.int32) for _ in range(3)]
    cat[2] = (cat[2] % 1000) - 500  # one small column with negative keys
    with CofactorContext(CFB_TRIPLE, 3, 3) as ctx:
        for lo in range(0, rows, 50_000):  # several appends: the table grows between tiles
            ctx.append([c[lo:lo + 50_000] for c in num], [c[lo:lo + 50_000] for c in cat])
        got = ctx.finalize_arrays()
        with pytest.raises(CofactorError):  # no dense partial for NCCL in this mode
            ctx.partial_sizes() and ctx.export_partial(0, 0)
    assert_parity(got, oracle.aggregate_arrays(CFB_TRIPLE, num, cat)[0], what="domain 1e5")
    assert len(got["pair_key1"]) > 2 * rows  # almost every row is a new (key1,key2) pair


def test_large_domain_device_scan_and_combine():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(78)
    rows = 400_000
    num = [rng.random(rows).astype(np.float32) for _ in range(2)]
    cat = [rng.integers(0, 200_000, rows).astype(np.int32) for _ in range(2)]
    h = rows // 2
    dn = [torch.from_numpy(c).cuda() for c in num]
    dc = [torch.from_numpy(c).cuda() for c in cat]
    with CofactorContext(CFB_TRIPLE, 2, 2) as a, CofactorContext(CFB_TRIPLE, 2, 2) as b:
        a.scan_device([t[:h] for t in dn], [t[:h] for t in dc], h)
        b.scan_device([t[h:] for t in dn], [t[h:] for t in dc], rows - h)
        a.combine(b)
        got = a.finalize_arrays()
    assert_parity(got, oracle.aggregate_arrays(CFB_TRIPLE, num, cat)[0], what="scan + combine, hashed pairs")


def test_whole_suite_with_hashed_pairs_forced():
    """Re-run the parity suites with CFB_DENSE_PAIR_BYTES=1: every pair count of every test goes
    through the hash table instead of the dense tables (domain growth, GROUP BY partitions,
    combine, sum_triple scatter, the DuckDB callbacks)."""
    env = dict(os.environ, CFB_DENSE_PAIR_BYTES="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_glue.py", "-m", "gpu", "-x",
                        "-q", "-k", "not partial_export"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def _wide_table(rng, rows):
    num = [rng.random(rows).astype(np.float32) for _ in range(3)]
    # -1 is in the pool on purpose: (pending, key = -1) must not look like a free dictionary slot (key_dict.cuh)
    pool = np.array([-2_000_000_000, -7, -1, 0, 3, 65_536, 1_999_999_999, 2_147_483_647, -2_147_483_648], np.int64)
    cat = [pool[rng.integers(0, len(pool), rows)].astype(np.int32),                     # ids all over int32
           rng.integers(0, 6, rows).astype(np.int32),                                    # ordinary small column
           (rng.integers(0, 300, rows).astype(np.int64) * 7_000_003 - 1_000_000_000).astype(np.int32)]
    return num, cat


def test_keys_spread_over_the_whole_int32_range():
    """Any int32 is a valid key for the reference's std::map; wide-range columns go through the key
    dictionary (key_dict.cuh).  Host feed in several tiles, GROUP BY, NB ring, filtered scan."""
    rng = np.random.default_rng(5)
    rows = 60_000
    num, cat = _wide_table(rng, rows)
    with CofactorContext(CFB_TRIPLE, 3, 3) as ctx:
        for lo in range(0, rows, 7_000):
            ctx.append([c[lo:lo + 7_000] for c in num], [c[lo:lo + 7_000] for c in cat])
            ctx.sync()  # force a tile per append: dictionaries and domains grow step by step
        got = ctx.finalize_arrays()
    assert_parity(got, oracle.aggregate_arrays(CFB_TRIPLE, num, cat)[0], what="wide keys, host feed")
    assert got["cat_keys"][0] == -2_147_483_648 and 2_147_483_647 in got["cat_keys"]
    grp = rng.integers(0, 3, rows).astype(np.int32)
    for kind in (CFB_TRIPLE, 1):
        with CofactorContext(kind, 3, 3, n_groups=3) as ctx:
            ctx.append(num, cat, group=grp.astype(np.uint32))
            res = [ctx.finalize_arrays(g) for g in range(3)]
        ref = oracle.aggregate_arrays(kind, num, cat, group=grp, n_groups=3)
        for g in range(3):
            assert_parity(res[g], ref[g], what=f"wide keys kind {kind} group {g}")


def test_wide_keys_device_scan_combine_and_dense_to_dictionary_switch():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(6)
    rows = 50_000
    num, cat = _wide_table(rng, rows)
    h = rows // 2
    narrow = [np.where(np.abs(c.astype(np.int64)) < 100_000, c, 5).astype(np.int32) for c in cat]
    dn = [torch.from_numpy(c).cuda() for c in num]
    dcat = [torch.from_numpy(np.concatenate([narrow[k][:h], cat[k][h:]])).cuda() for k in range(3)]
    with CofactorContext(CFB_TRIPLE, 3, 3) as a, CofactorContext(CFB_TRIPLE, 3, 3) as b:
        a.scan_device([t[:h] for t in dn], [t[:h] for t in dcat], h)      # narrow keys: dense slots
        a.scan_device([t[h:] for t in dn], [t[h:] for t in dcat], rows - h)  # wide keys arrive: switch to codes
        whole = a.finalize_arrays()
        b.append([c[h:] for c in num], [c[h:] for c in cat])              # dictionary state
        a2 = CofactorContext(CFB_TRIPLE, 3, 3)
        a2.append([c[:h] for c in num], [c[:h] for c in narrow])          # dense state
        a2.combine(b)                                                     # dense += dictionary -> merged by key
        merged = a2.finalize_arrays()
        a2.close()
    ref = oracle.aggregate_arrays(CFB_TRIPLE, num, [np.concatenate([narrow[k][:h], cat[k][h:]]) for k in range(3)])[0]
    assert_parity(whole, ref, what="dense -> dictionary switch")
    assert_parity(merged, ref, what="dense += dictionary combine")


def test_wide_keys_through_the_callbacks_and_lifted_sum():
    from duckdb_imputation_b200 import replay
    from tests.parity import assert_struct_parity
    rng = np.random.default_rng(7)
    rows = 30_000
    num, cat = _wide_table(rng, rows)
    gb = rng.integers(0, 3, rows)
    ref = oracle.aggregate(0, num, cat, group_by=gb)
    for lifted in (False, True):
        got = replay.glue().query(0, num, cat, group_by=gb, threads=3, lifted=lifted)
        for a, b in zip(got, ref):
            assert_struct_parity(a, b, what=f"wide keys lifted={lifted}")


def test_whole_suite_with_dictionaries_forced():
    """Re-run the parity suites with CFB_DICT_RANGE=1: every categorical column without a declared
    domain is keyed through a dictionary."""
    env = dict(os.environ, CFB_DICT_RANGE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_glue.py", "-m", "gpu", "-x",
                        "-q", "-k", "not partial_export and not spread_over"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
