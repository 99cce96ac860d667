"""CPU suite: the reference's OWN ring-aggregate sources (compiled unmodified from /root/reference
over the DuckDB-vector shim into oracle/_ref, driven by the hash-aggregate replay host).

This pins (a) the shim + replay protocol -- the reference must reproduce its own golden STRUCTs --
and (b) the oracle restatement -- its ref-faithful mode must agree with the reference bit for bit
on random data, including multi-threaded combine and filtered (dictionary-vector) scans, which no
reference test exercises."""
import numpy as np
import pytest

from oracle import oracle, ref_replay
from tests import sqlmini

pytestmark = pytest.mark.skipif(not ref_replay.available(), reason="oracle/_ref not built (needs /root/reference)")


def _f32(x):
    if isinstance(x, dict):
        return {k: _f32(v) for k, v in x.items()}
    if isinstance(x, list):
        return [_f32(v) for v in x]
    return float(np.float32(x)) if isinstance(x, float) else x


def test_reference_reproduces_its_own_goldens(goldens):
    r = ref_replay.ref()
    assert r.implementation == "reference"
    cases = [c for c in goldens["cases"] if c["file"] in ("test_sum.py", "test_nb_sum.py")]
    assert len(cases) == 8
    for c in cases:
        got = sqlmini.run_sum(c["sql"], goldens["fixtures"][c["file"]], r.query)
        assert got[c["index"]] == c["expected"], (c["file"], c["test"], c["index"])


def test_reference_registration_grid_stops_at_19():
    fns = set(ref_replay.ref().functions())
    assert "sum_to_triple_19_19" in fns and "sum_to_nb_agg_19_0" in fns
    assert "sum_to_triple_20_0" not in fns and "sum_to_triple_0_0" not in fns  # ext.cpp:80-85
    assert len(fns) == 2 * 399 + 1 and "ref_sum_to_triple_20_0" in fns  # + the bench-only extra


@pytest.mark.parametrize("kind", [oracle.TRIPLE, oracle.NB])
@pytest.mark.parametrize("threads", [1, 4])
def test_faithful_oracle_is_bit_identical_to_the_reference(kind, threads):
    rng = np.random.default_rng(31 + kind)
    rows = 25_000
    num = [rng.random(rows).astype(np.float32) for _ in range(6)]
    cat = [rng.integers(-4, 40, rows).astype(np.int32) for _ in range(3)]
    gb = rng.integers(0, 5, rows)
    r = ref_replay.ref()
    assert _f32(r.query(kind, num, cat, group_by=gb, threads=threads)) == \
        _f32(oracle.aggregate(kind, num, cat, group_by=gb, mode=oracle.FAITHFUL, threads=threads))
    keep = rng.random(rows) < 0.8  # MICE-style filtered scan (imputation_base.cpp:29)
    assert _f32(r.query(kind, num, cat, where=keep, threads=threads)) == \
        _f32(oracle.aggregate(kind, num, cat, where=keep, mode=oracle.FAITHFUL, threads=threads))


def test_exact_oracle_agrees_with_reference_within_fp32_error():
    from tests.parity import assert_struct_parity
    rng = np.random.default_rng(5)
    rows = 50_000
    num = [rng.random(rows).astype(np.float32) for _ in range(8)]
    cat = [rng.integers(0, 12, rows).astype(np.int32) for _ in range(2)]
    ref = ref_replay.ref().query(oracle.TRIPLE, num, cat, threads=2)
    exact = oracle.aggregate(oracle.TRIPLE, num, cat)
    assert_struct_parity(ref, exact, rtol=2e-4, what="reference fp32 vs exact")  # sequential fp32 drift at 50k rows
