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
