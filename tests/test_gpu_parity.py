"""GPU suite: the CUDA path, called through the C ABI, against the CPU oracle.

Bar: N / keys / counts bit-exact; sums <= 1e-5 relative to the fp64 oracle (tests/parity.py);
exact equality with the reference's golden STRUCTs."""
import ctypes as C

import numpy as np
import pytest

from duckdb_imputation_b200 import CFB_NB, CFB_TRIPLE, CofactorContext, CofactorError, sum_to_nb_agg, sum_to_triple
from duckdb_imputation_b200 import _native as nat
from duckdb_imputation_b200 import synth
from oracle import oracle
from tests import sqlmini
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu


def _backend(kind, num, cat, group_by=None, where=None):
    f = sum_to_triple if kind == CFB_TRIPLE else sum_to_nb_agg
    return f(num, cat, group_by=group_by, where=where)


def test_reference_goldens_through_the_c_abi(goldens):
    cases = [c for c in goldens["cases"] if c["file"] in ("test_sum.py", "test_nb_sum.py")]
    assert len(cases) == 8
    for c in cases:
        got = sqlmini.run_sum(c["sql"], goldens["fixtures"][c["file"]], _backend)
        assert got[c["index"]] == c["expected"], (c["file"], c["test"], c["index"])


def _table(rng, rows, n, m, dom=100, lo=0, dist="uniform"):
    if dist == "normal":
        num = [rng.standard_normal(rows).astype(np.float32) for _ in range(n)]
    else:
        num = [rng.random(rows).astype(np.float32) for _ in range(n)]
    cat = [rng.integers(lo, lo + dom, rows).astype(np.int32) for _ in range(m)]
    return num, cat


def _gpu_host(kind, num, cat, group=None, n_groups=1, chunk=2048, sel=None):
    """Feed host columns through cfb_ctx_append the way the DuckDB glue does."""
    rows = len(sel) if sel is not None else (len(num[0]) if num else len(cat[0]))
    with CofactorContext(kind, len(num), len(cat), n_groups) as ctx:
        for lo in range(0, rows, chunk):
            hi = min(rows, lo + chunk)
            g = None if group is None else group[lo:hi]
            if sel is None:
                ctx.append([c[lo:hi] for c in num], [c[lo:hi] for c in cat], group=g, count=hi - lo)
            else:
                s = sel[lo:hi]
                ctx.append(num, cat, group=g, num_sel=[s] * len(num), cat_sel=[s] * len(cat), count=hi - lo)
        return [ctx.finalize_arrays(g) for g in range(n_groups)]


SHAPES = [
    (CFB_TRIPLE, 5, 0), (CFB_TRIPLE, 20, 0), (CFB_TRIPLE, 10, 10), (CFB_TRIPLE, 3, 3), (CFB_TRIPLE, 0, 2),
    (CFB_TRIPLE, 1, 0), (CFB_TRIPLE, 11, 1), (CFB_TRIPLE, 17, 0), (CFB_TRIPLE, 21, 2), (CFB_TRIPLE, 32, 0),
    (CFB_NB, 12, 4), (CFB_NB, 4, 0), (CFB_NB, 32, 1), (CFB_NB, 0, 3),
]


@pytest.mark.parametrize("kind,n,m", SHAPES)
def test_host_feed_matches_oracle(kind, n, m):
    rng = np.random.default_rng(100 + 7 * n + m)
    rows = 150_001
    num, cat = _table(rng, rows, n, m, dom=100 if m <= 4 else 20, lo=-5)
    got = _gpu_host(kind, num, cat, chunk=2048)[0]
    ref = oracle.aggregate_arrays(kind, num, cat)[0]
    assert_parity(got, ref, what=f"kind={kind} n={n} m={m}")


@pytest.mark.parametrize("kind,n,m,G", [(CFB_NB, 12, 4, 10), (CFB_TRIPLE, 12, 0, 10), (CFB_TRIPLE, 4, 2, 3)])
def test_group_by_matches_oracle(kind, n, m, G):
    rng = np.random.default_rng(5)
    rows = 60_000
    num, cat = _table(rng, rows, n, m, dom=30)
    group = rng.integers(0, G, rows).astype(np.uint32)
    got = _gpu_host(kind, num, cat, group=group, n_groups=G)
    ref = oracle.aggregate_arrays(kind, num, cat, group=group.astype(np.int32), n_groups=G)
    for g in range(G):
        assert_parity(got[g], ref[g], what=f"group {g}")


def test_filtered_scan_through_selection_vectors():
    rng = np.random.default_rng(6)
    rows = 40_000
    num, cat = _table(rng, rows, 6, 2, dom=12)
    sel = np.nonzero(rng.random(rows) < 0.8)[0].astype(np.uint32)
    got = _gpu_host(CFB_TRIPLE, num, cat, sel=sel)[0]
    ref = oracle.aggregate_arrays(CFB_TRIPLE, num, cat, sel=sel)[0]
    assert_parity(got, ref)


def test_cancelling_sums_normal_data():
    rng = np.random.default_rng(11)
    rows = 300_000
    num, cat = _table(rng, rows, 8, 0, dist="normal")
    got = _gpu_host(CFB_TRIPLE, num, cat, chunk=65536)[0]
    ref = oracle.aggregate_arrays(CFB_TRIPLE, num, cat)[0]
    # off-diagonal sums of N(0,1) products cancel to ~sqrt(rows): floor on sum of |terms| ~ rows
    assert_parity(got, ref, abs_scale=float(rows))


def _device_cols(torch, rows, n, m, seed, dom=100, lo=0):
    l = nat.lib()
    dn = [torch.empty(rows, dtype=torch.float32, device="cuda") for _ in range(n)]
    dc = [torch.empty(rows, dtype=torch.int32, device="cuda") for _ in range(m)]
    for k, t in enumerate(dn):
        nat.check(l.cfb_gen_uniform_f32(0, t.data_ptr(), rows, synth.column_seed(seed, k), 0, None))
    for k, t in enumerate(dc):
        nat.check(l.cfb_gen_int32(0, t.data_ptr(), rows, synth.column_seed(seed, 100 + k), 0, lo, dom, None))
    torch.cuda.synchronize()
    hn = [synth.uniform_f32(rows, synth.column_seed(seed, k)) for k in range(n)]
    hc = [synth.int32(rows, synth.column_seed(seed, 100 + k), lo=lo, rng=dom) for k in range(m)]
    return dn, dc, hn, hc


@pytest.mark.parametrize("kind,n,m,rows", [
    (CFB_TRIPLE, 5, 0, 1_000_000), (CFB_TRIPLE, 20, 0, 2_000_003), (CFB_TRIPLE, 10, 10, 500_001),
    (CFB_NB, 12, 4, 700_002), (CFB_TRIPLE, 20, 0, 3), (CFB_TRIPLE, 7, 0, 513), (CFB_TRIPLE, 32, 0, 300_000),
    (CFB_TRIPLE, 24, 0, 100_001), (CFB_TRIPLE, 13, 0, 65_536),
])
def test_device_resident_scan_matches_oracle(kind, n, m, rows):
    torch = pytest.importorskip("torch")
    dn, dc, hn, hc = _device_cols(torch, rows, n, m, seed=42)
    # the device generator and its host twin must agree bit for bit
    assert np.array_equal(dn[0].cpu().numpy(), hn[0])
    if m:
        assert np.array_equal(dc[0].cpu().numpy(), hc[0])
    with CofactorContext(kind, n, m) as ctx:
        ctx.scan_device(dn, dc, rows)
        got = ctx.finalize_arrays()
    ref = oracle.aggregate_arrays(kind, hn, hc)[0]
    assert_parity(got, ref, what=f"device n={n} m={m} rows={rows}")


def test_device_scan_is_deterministic_and_accumulates():
    torch = pytest.importorskip("torch")
    rows = 1_000_000
    dn, dc, hn, hc = _device_cols(torch, rows, 20, 0, seed=3)
    outs = []
    for _ in range(2):
        with CofactorContext(CFB_TRIPLE, 20, 0) as ctx:
            ctx.scan_device(dn, dc, rows)
            outs.append(ctx.finalize_arrays())
    assert np.array_equal(outs[0]["quad"], outs[1]["quad"]) and np.array_equal(outs[0]["lin"], outs[1]["lin"])
    with CofactorContext(CFB_TRIPLE, 20, 0) as ctx:  # two scans into one state == 2x
        ctx.scan_device(dn, dc, rows)
        ctx.scan_device(dn, dc, rows)
        twice = ctx.finalize_arrays()
    assert twice["N"] == 2 * rows
    np.testing.assert_allclose(twice["quad"], 2 * outs[0]["quad"], rtol=1e-12)


def test_device_group_by():
    torch = pytest.importorskip("torch")
    rows, G = 400_000, 10
    dn, dc, hn, hc = _device_cols(torch, rows, 12, 4, seed=9)
    dg = torch.empty(rows, dtype=torch.int32, device="cuda")
    nat.check(nat.lib().cfb_gen_int32(0, dg.data_ptr(), rows, 777, 0, 0, G, None))
    torch.cuda.synchronize()
    hg = synth.int32(rows, 777, lo=0, rng=G)
    for kind in (CFB_NB, CFB_TRIPLE):
        with CofactorContext(kind, 12, 4, n_groups=G) as ctx:
            ctx.scan_device(dn, dc, rows, d_group=dg)
            got = [ctx.finalize_arrays(g) for g in range(G)]
        ref = oracle.aggregate_arrays(kind, hn, hc, group=hg, n_groups=G)
        for g in range(G):
            assert_parity(got[g], ref[g], what=f"kind {kind} group {g}")


def test_combine_equals_single_scan():
    rng = np.random.default_rng(21)
    rows = 50_000
    num, cat = _table(rng, rows, 6, 3, dom=15, lo=-4)
    h = 20_000
    with CofactorContext(CFB_TRIPLE, 6, 3) as a, CofactorContext(CFB_TRIPLE, 6, 3) as b, \
            CofactorContext(CFB_TRIPLE, 6, 3) as empty:
        a.append([c[:h] for c in num], [c[:h] % 7 for c in cat])  # a sees a narrower key range than b
        b.append([c[h:] for c in num], [c[h:] for c in cat])
        a.combine(b)
        a.combine(empty)
        got = a.finalize_arrays()
        b_alone = b.finalize_arrays()  # src stays valid (sum_state.h:48-52 destroys it later)
    cat2 = [np.concatenate([c[:h] % 7, c[h:]]) for c in cat]
    assert_parity(got, oracle.aggregate_arrays(CFB_TRIPLE, num, cat2)[0])
    assert_parity(b_alone, oracle.aggregate_arrays(CFB_TRIPLE, [c[h:] for c in num], [c[h:] for c in cat])[0])


def test_domain_grows_across_appends_and_negative_keys():
    num = [np.array([1, 2, 3, 4], np.float32)]
    with CofactorContext(CFB_TRIPLE, 1, 1) as ctx:
        ctx.append([num[0][:2]], [np.array([5, 6], np.int32)])
        ctx.sync()
        ctx.append([num[0][2:]], [np.array([-1000, 70000], np.int32)])
        got = ctx.finalize()
    assert got["lin_cat"] == [[{"key": -1000, "value": 1.0}, {"key": 5, "value": 1.0}, {"key": 6, "value": 1.0},
                               {"key": 70000, "value": 1.0}]]
    assert [e["value"] for e in got["quad_num_cat"][0]] == [3.0, 1.0, 2.0, 4.0]


def test_declared_domain_violation_is_an_error():
    with CofactorContext(CFB_TRIPLE, 0, 1) as ctx:
        ctx.set_cat_domain([0], [9])
        ctx.append([], [np.array([1, 2, 10], np.int32)])
        with pytest.raises(CofactorError) as e:
            ctx.sync()
        assert e.value.code == nat.CFB_ERR_DOMAIN


def test_empty_input_and_tiny_inputs():
    with CofactorContext(CFB_TRIPLE, 2, 1) as ctx:
        got = ctx.finalize()
    assert got == {"N": 0, "lin_agg": [0.0, 0.0], "quad_agg": [0.0, 0.0, 0.0], "lin_cat": [[]],
                   "quad_num_cat": [[], []], "quad_cat": [[]]}
    one = sum_to_triple([np.array([3.0], np.float32)], [np.array([4], np.int32)])
    assert one == {"N": 1, "lin_agg": [3.0], "quad_agg": [9.0], "lin_cat": [[{"key": 4, "value": 1.0}]],
                   "quad_num_cat": [[{"key": 4, "value": 3.0}]], "quad_cat": [[{"key1": 4, "key2": 4, "value": 1.0}]]}


def test_kernels_actually_launch():
    before = nat.lib().cfb_kernel_launches()
    sum_to_triple([np.ones(10, np.float32)] * 3, [])
    assert nat.lib().cfb_kernel_launches() > before


def test_partial_export_matches_host_layout_and_roundtrips():
    """The dense device state (state_layout.h) == multi_gpu.pack_dense of the finalized result; and
    import(export(x)) is the identity.  This is the buffer the NCCL all-reduce sums."""
    torch = pytest.importorskip("torch")
    from duckdb_imputation_b200 import multi_gpu
    rng = np.random.default_rng(2)
    rows = 30_000
    num, cat = _table(rng, rows, 3, 3, dom=9, lo=-2)
    lo, hi = [-4, -2, -2], [8, 6, 10]
    with CofactorContext(CFB_TRIPLE, 3, 3) as ctx, CofactorContext(CFB_TRIPLE, 3, 3) as other:
        ctx.set_cat_domain(lo, hi)
        other.set_cat_domain(lo, hi)
        ctx.append(num, cat)
        nf, nu = ctx.partial_sizes()
        assert (nf, nu) == multi_gpu.dense_sizes(0, 3, 3, lo, hi)
        f = torch.zeros(nf, dtype=torch.float64, device="cuda")
        u = torch.zeros(nu, dtype=torch.int64, device="cuda")
        ctx.export_partial(f, u)
        arrays = ctx.finalize_arrays()
        pf, pu = multi_gpu.pack_dense(arrays, lo, hi)
        assert np.array_equal(u.cpu().numpy(), pu)
        np.testing.assert_allclose(f.cpu().numpy(), pf, rtol=0, atol=0)
        other.import_partial(f * 2, u * 2)  # what a 2-rank all-reduce of identical slices would give
        twice = other.finalize_arrays()
    assert twice["N"] == 2 * rows and np.array_equal(twice["cat_counts"], 2 * arrays["cat_counts"])
    np.testing.assert_allclose(twice["quad"], 2 * arrays["quad"], rtol=1e-15)


def test_mice_style_filtered_device_scan_via_group_slots():
    """MICE issues sum_to_triple(...) WHERE col_IS_NULL IS FALSE over ~80% of the rows
    (imputation_base.cpp:21-34).  On device-resident data the filter is a 2-slot GROUP BY on the
    null flag: slot 0 = rows that pass.  Must equal the oracle's filtered scan."""
    torch = pytest.importorskip("torch")
    rows = 300_000
    dn, dc, hn, hc = _device_cols(torch, rows, 6, 3, seed=5, dom=12)
    flag = synth.int32(rows, 4242, lo=0, rng=5)          # 0..4
    is_null = (flag == 0).astype(np.int32)               # ~20% NULL in the imputed column
    dflag = torch.from_numpy(is_null).cuda()
    with CofactorContext(CFB_TRIPLE, 6, 3, n_groups=2) as ctx:
        ctx.scan_device(dn, dc, rows, d_group=dflag)
        got = ctx.finalize_arrays(0)
        nulls = ctx.finalize_arrays(1)
    sel = np.nonzero(is_null == 0)[0].astype(np.uint32)
    assert_parity(got, oracle.aggregate_arrays(CFB_TRIPLE, hn, hc, sel=sel)[0], what="rows that pass the filter")
    assert got["N"] + nulls["N"] == rows


@pytest.mark.parametrize("kind,n,m", [(CFB_TRIPLE, 20, 0), (CFB_TRIPLE, 6, 3), (CFB_NB, 12, 4), (CFB_TRIPLE, 32, 0), (CFB_TRIPLE, 0, 2)])
def test_negative_slot_drops_the_row(kind, n, m):
    """A filtered device scan is n_groups = 1 with slot 0 (keep) / -1 (drop)."""
    torch = pytest.importorskip("torch")
    rows = 200_003
    dn, dc, hn, hc = _device_cols(torch, rows, n, m, seed=6, dom=9)
    keep = synth.int32(rows, 99, lo=0, rng=5) != 0
    dslot = torch.from_numpy(np.where(keep, 0, -1).astype(np.int32)).cuda()
    with CofactorContext(kind, n, m) as ctx:
        ctx.scan_device(dn, dc, rows, d_group=dslot)
        got = ctx.finalize_arrays()
    sel = np.nonzero(keep)[0].astype(np.uint32)
    assert_parity(got, oracle.aggregate_arrays(kind, hn, hc, sel=sel)[0], what=f"filtered n={n} m={m}")


def test_many_groups_fall_back_when_tables_do_not_fit():
    torch = pytest.importorskip("torch")
    rows, G = 100_000, 300  # 300 slots x 231 entries do not fit the warp-private tables
    dn, dc, hn, hc = _device_cols(torch, rows, 20, 0, seed=8)
    hg = synth.int32(rows, 555, lo=0, rng=G)
    dg = torch.from_numpy(hg).cuda()
    with CofactorContext(CFB_TRIPLE, 20, 0, n_groups=G) as ctx:
        ctx.scan_device(dn, dc, rows, d_group=dg)
        got = [ctx.finalize_arrays(g) for g in (0, 17, G - 1)]
    ref = oracle.aggregate_arrays(CFB_TRIPLE, hn, hc, group=hg, n_groups=G)
    for g, a in zip((0, 17, G - 1), got):
        assert_parity(a, ref[g], what=f"group {g}")


def test_cross_device_combine():
    """cfb_ctx_combine across GPUs (peer copy of the source state), dense and hashed pair counts."""
    if nat.lib().cfb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(33)
    rows = 40_000
    for dom in (15, 60_000):  # dense tables / hash fallback
        num, cat = _table(rng, rows, 4, 2, dom=dom)
        h = rows // 2
        with CofactorContext(CFB_TRIPLE, 4, 2, device=0) as a, CofactorContext(CFB_TRIPLE, 4, 2, device=1) as b:
            a.append([c[:h] for c in num], [c[:h] for c in cat])
            b.append([c[h:] for c in num], [c[h:] for c in cat])
            a.combine(b)
            got = a.finalize_arrays()
        assert_parity(got, oracle.aggregate_arrays(CFB_TRIPLE, num, cat)[0], what=f"cross-device dom={dom}")


@pytest.mark.parametrize("n,m,G", [(10, 10, 1), (20, 10, 2), (3, 4, 1)])
def test_skewed_keys_switch_the_list_plan(n, m, G, monkeypatch):
    """Zipf-distributed keys over a domain of 100: a hot key's bucket is far longer than the static sub-lists expect,
    so chain_sum_kernel hands that bucket more sub-lists for the CTA's NEXT tile (and works tiles off in two pieces
    while that plan is in force).  Small tiles (CFB_CHAIN_TILE) give every CTA several tiles at test size, so the
    switch to the plan, its use, the short rest-of-tile pieces and the way back all run.  Counts exact, sums <= 1e-5."""
    import torch
    monkeypatch.setenv("CFB_CHAIN_TILE", "512")
    rng = np.random.default_rng(77 + n)
    rows = 420_013
    num = [rng.random(rows).astype(np.float32) for _ in range(n)]
    cat = [np.minimum(rng.zipf(1.2, rows) - 1, 99).astype(np.int32) for _ in range(m)]
    # the second half of the table is uniform again in two columns: the plan must fall back to the static one
    for c in cat[:2]:
        c[rows // 2:] = rng.integers(0, 100, rows - rows // 2)
    group = (rng.random(rows) < 0.2).astype(np.int32) if G > 1 else None
    dn = [torch.from_numpy(c).cuda() for c in num]
    dc = [torch.from_numpy(c).cuda() for c in cat]
    dg = torch.from_numpy(group).cuda() if G > 1 else None
    with CofactorContext(CFB_TRIPLE, n, m, G) as ctx:
        ctx.set_cat_domain([0] * m, [99] * m)
        ctx.scan_device(dn, dc, rows, d_group=dg)
        got = [ctx.finalize_arrays(g) for g in range(G)]
    want = oracle.aggregate_arrays(oracle.TRIPLE, num, cat, group=group, n_groups=G)
    for g in range(G):
        assert_parity(got[g], want[g], what=f"skewed keys, slot {g}")
