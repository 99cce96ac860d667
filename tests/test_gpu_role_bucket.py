"""GPU suite: the shared-memory categorical path -- role_scan_kernel (pair counts in shared-memory tables
cut into roles) and bucket_sum_kernel (per-key payload sums by tile bucketing) -- against the fp64 oracle
and against the L2-reduction path it replaces (slab_scan_kernel, CFB_NO_ROLE / CFB_NO_BUCKET).

Bar: N, keys, key counts and pair counts bit-exact; sums <= 1e-5 relative (tests/parity.py)."""
import os

import numpy as np
import pytest

from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext, CofactorError
from oracle import oracle
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(cols, dt):
    return [torch.from_numpy(np.ascontiguousarray(c, dt)).cuda() for c in cols]


def _scan(num, cat, group=None, n_groups=1, domain=None, env=None, offset=0):
    """Device-resident scan through cfb_triple_device; `offset` shifts the column pointers by that many rows
    (unaligned input: the 128-bit loads of the role kernel do not apply)."""
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        rows = len(cat[0]) if cat else len(num[0])
        dn, dc = _dev(num, np.float32), _dev(cat, np.int32)
        dg = None if group is None else torch.from_numpy(np.ascontiguousarray(group, np.int32)).cuda()
        with CofactorContext(CFB_TRIPLE, len(num), len(cat), n_groups) as ctx:
            if domain is not None:
                ctx.set_cat_domain([domain[0]] * len(cat), [domain[1]] * len(cat))
            ctx.scan_device([t[offset:] for t in dn], [t[offset:] for t in dc], rows - offset,
                            d_group=None if dg is None else dg[offset:])
            return [ctx.finalize_arrays(g) for g in range(n_groups)]
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


LEGACY = {"CFB_NO_ROLE": "1", "CFB_NO_BUCKET": "1"}


def _table(seed, rows, n, doms, lo=0):
    rng = np.random.default_rng(seed)
    num = [rng.random(rows).astype(np.float32) for _ in range(n)]
    cat = [rng.integers(lo, lo + d, rows).astype(np.int32) for d in doms]
    return num, cat


@pytest.mark.parametrize("n,doms,rows", [
    (10, [100] * 10, 300_003),          # C3: 45 tables of 10^4 cells -> 9 roles of 32-bit cells
    (20, [100] * 10, 120_001),          # MICE shape: 24-float payload rows
    (3, [7, 300, 2, 41], 77_777),       # ragged domains, one wide table
    (0, [50, 60], 50_000),              # no numeric columns: payload = the count
    (5, [1000], 65_536),                # one categorical column: bucket kernel alone, no pairs
    (2, [3] * 12, 40_000),              # 66 tiny tables in one role
    (32, [10, 10], 33_001),             # widest payload
])
def test_matches_oracle_and_legacy_path(n, doms, rows):
    num, cat = _table(7 * n + len(doms), rows, n, doms, lo=-3)
    got = _scan(num, cat)[0]
    ref = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    assert_parity(got, ref, what=f"n={n} doms={doms}")
    old = _scan(num, cat, env=LEGACY)[0]
    assert_parity(old, ref, what=f"legacy n={n} doms={doms}")
    for k in ("N", "cat_keys", "cat_counts", "pair_key1", "pair_key2", "pair_counts"):
        assert np.array_equal(got[k], old[k]), k


@pytest.mark.parametrize("env", [{"CFB_ROLE_BITS": "16"}, {"CFB_ROLE_BITS": "32"}, {"CFB_NO_BUCKET": "1"},
                                 {"CFB_NO_BUCKET": "1", "CFB_ROLE_SUBSLABS": "1"}, {"CFB_NO_ROLE": "1"}])
def test_every_variant_is_exact(env):
    """16-bit packed cells (folded every 65 K rows), 32-bit cells, payloads as L2 reductions, bucket sums only."""
    num, cat = _table(11, 200_000, 6, [60, 70, 80, 90])
    ref = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    assert_parity(_scan(num, cat, env=env)[0], ref, what=str(env))


def test_sixteen_bit_cells_cannot_overflow():
    """Every row hits the same cell: 16-bit cells are folded before 65 536 increments."""
    rows = 400_000
    num = [np.ones(rows, np.float32)]
    cat = [np.full(rows, 5, np.int32), np.full(rows, 9, np.int32), np.zeros(rows, np.int32)]
    for bits in ("16", "32"):
        got = _scan(num, cat, domain=(0, 15), env={"CFB_ROLE_BITS": bits})[0]
        assert got["N"] == rows and list(got["cat_counts"]) == [rows] * 3 and list(got["pair_counts"]) == [rows] * 6
        assert abs(got["numcat"].sum() - 3 * rows) < 1e-6 * rows


def test_row_filter_and_group_slots():
    """WHERE (slot < 0 rows are skipped) and GROUP BY slots (MICE: slot 0 = observed rows, slot 1 = NULL rows)."""
    num, cat = _table(3, 150_000, 8, [30, 40, 50])
    rng = np.random.default_rng(5)
    keep = rng.random(150_000) < 0.8
    got = _scan(num, cat, group=np.where(keep, 0, -1))[0]
    assert_parity(got, oracle.aggregate_arrays(oracle.TRIPLE, num, cat, sel=np.nonzero(keep)[0].astype(np.uint32))[0], what="filter")
    for G in (2, 5):
        slots = rng.integers(-1, G, 150_000).astype(np.int32)
        got = _scan(num, cat, group=slots, n_groups=G)
        for g in range(G):
            ref = oracle.aggregate_arrays(oracle.TRIPLE, num, cat, sel=np.nonzero(slots == g)[0].astype(np.uint32))[0]
            assert_parity(got[g], ref, what=f"G={G} slot {g}")


def test_unaligned_columns_and_tails():
    num, cat = _table(9, 100_003, 4, [25, 35])
    for off in (0, 1, 3):
        ref = oracle.aggregate_arrays(oracle.TRIPLE, [c[off:] for c in num], [c[off:] for c in cat])[0]
        if off == 0:
            assert_parity(_scan(num, cat, offset=off)[0], ref, what=f"offset {off}")
        else:
            # numeric columns must be 16-byte aligned for the Gram kernel (CFB_ERR_INVALID); categorical-only
            # input may sit anywhere: the role kernel steps aside for the scalar path
            got = _scan([], cat, offset=off)[0]
            assert_parity(got, oracle.aggregate_arrays(oracle.TRIPLE, [], [c[off:] for c in cat])[0], what=f"offset {off}")


def test_key_outside_the_declared_domain_is_an_error():
    num, cat = _table(13, 50_000, 2, [20, 20])
    cat[1][31_234] = 99
    with pytest.raises(CofactorError):
        _scan(num, cat, domain=(0, 19))


def test_shapes_the_tables_do_not_fit_stay_on_the_l2_path():
    num, cat = _table(17, 60_000, 2, [2000, 700])  # 1.4 M cells: no shared-memory table
    got = _scan(num, cat)[0]
    assert_parity(got, oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0], what="wide pair table")


def test_skewed_keys_and_nb_key_counts():
    """Zipf-like keys (one hot bucket / hot cells: the worst case for bucket owners and for same-address atomics)
    and the Naive-Bayes ring's shared-memory key histogram, with GROUP BY slots."""
    from duckdb_imputation_b200 import CFB_NB
    rng = np.random.default_rng(23)
    rows = 200_000
    num = [rng.random(rows).astype(np.float32) for _ in range(5)]
    cat = [np.minimum(rng.zipf(1.2, rows) - 1, 99).astype(np.int32) for _ in range(4)]
    ref = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    assert ref["cat_counts"].max() > rows // 5  # really skewed
    assert_parity(_scan(num, cat)[0], ref, what="zipf keys")
    slots = rng.integers(-1, 3, rows).astype(np.int32)
    dn, dc = _dev(num, np.float32), _dev(cat, np.int32)
    with CofactorContext(CFB_NB, 5, 4, 3) as ctx:
        ctx.scan_device(dn, dc, rows, d_group=torch.from_numpy(slots).cuda())
        got = [ctx.finalize_arrays(g) for g in range(3)]
    for g in range(3):
        ref = oracle.aggregate_arrays(oracle.NB, num, cat, sel=np.nonzero(slots == g)[0].astype(np.uint32))[0]
        assert_parity(got[g], ref, what=f"nb slot {g}")
