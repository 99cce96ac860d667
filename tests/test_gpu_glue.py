"""GPU suite: OUR extension behind the DuckDB aggregate callbacks (triple_glue.cpp), driven by the
hash-aggregate replay host with DuckDB's protocol, against the reference goldens and the oracle."""
import numpy as np
import pytest

from duckdb_imputation_b200 import replay
from oracle import oracle
from tests import sqlmini
from tests.parity import assert_struct_parity

pytestmark = pytest.mark.gpu


def test_registration_superset_of_the_reference():
    fns = set(replay.glue().functions())
    assert replay.glue().implementation == "b200"
    for name in ("sum_to_triple_3_3", "sum_to_triple_19_19", "sum_to_triple_20_0", "sum_to_nb_agg_12_4", "sum_to_triple_20_10"):
        assert name in fns
    assert "sum_to_triple_0_0" not in fns and len(fns) == 2 * (21 * 21 - 1) + 2  # + sum_triple, sum_nb_agg
    assert {"sum_triple", "sum_nb_agg"} <= fns


def test_goldens_through_the_duckdb_callbacks(goldens):
    g = replay.glue()
    cases = [c for c in goldens["cases"] if c["file"] in ("test_sum.py", "test_nb_sum.py")]
    for threads in (1, 3):
        for c in cases:
            got = sqlmini.run_sum(c["sql"], goldens["fixtures"][c["file"]],
                                  lambda *a, **k: g.query(*a, threads=threads, **k))
            assert got[c["index"]] == c["expected"], (c["file"], c["test"], c["index"], threads)


@pytest.mark.parametrize("kind,n,m", [(0, 20, 0), (0, 10, 10), (1, 12, 4), (0, 5, 0), (0, 3, 3), (0, 0, 2)])
@pytest.mark.parametrize("threads", [1, 4])
def test_callbacks_match_oracle(kind, n, m, threads):
    rng = np.random.default_rng(17 * n + m + threads)
    rows = 100_003
    num = [rng.random(rows).astype(np.float32) for _ in range(n)]
    cat = [rng.integers(-2, 25, rows).astype(np.int32) for _ in range(m)]
    got = replay.glue().query(kind, num, cat, threads=threads)
    assert_struct_parity(got, oracle.aggregate(kind, num, cat), what=f"{kind} {n} {m} T={threads}")


def test_group_by_and_where_through_callbacks():
    rng = np.random.default_rng(3)
    rows = 80_000
    num = [rng.random(rows).astype(np.float32) for _ in range(12)]
    cat = [rng.integers(0, 30, rows).astype(np.int32) for _ in range(4)]
    gb = rng.integers(0, 10, rows)
    keep = rng.random(rows) < 0.8
    for kind in (0, 1):
        got = replay.glue().query(kind, num, cat, group_by=gb, where=keep, threads=4)
        ref = oracle.aggregate(kind, num, cat, group_by=gb, where=keep)
        assert len(got) == len(ref) == 10
        for g, (a, b) in enumerate(zip(got, ref)):
            assert_struct_parity(a, b, what=f"kind {kind} group {g}")


def test_unknown_function_is_a_query_error():
    with pytest.raises(replay.ReplayError, match="does not exist"):
        replay.glue().aggregate("sum_to_triple_21_0", [np.zeros(4, np.float32)] * 21, [])


def test_sum_of_lifted_equals_sum_goldens(goldens):
    """test_sum.py:40-52 / test_nb_sum.py: sum_to_triple(..) == sum_triple(to_cofactor(..)), grouped and
    with HAVING -- here additionally pinned to the golden STRUCTs themselves."""
    g = replay.glue()
    cases = [c for c in goldens["cases"] if c["file"] in ("test_sum.py", "test_nb_sum.py")]
    for c in cases:
        fx = goldens["fixtures"][c["file"]]
        direct = sqlmini.run_sum(c["sql"], fx, g.query)
        lifted = sqlmini.run_sum(c["sql"], fx, lambda *a, **k: g.query(*a, lifted=True, **k))
        assert direct == lifted
        assert lifted[c["index"]] == c["expected"], (c["file"], c["test"], c["index"])


@pytest.mark.parametrize("kind", [0, 1])
def test_sum_of_lifted_matches_oracle(kind):
    rng = np.random.default_rng(8)
    rows = 20_000
    num = [rng.random(rows).astype(np.float32) for _ in range(5)]
    cat = [rng.integers(-3, 12, rows).astype(np.int32) for _ in range(3)]
    gb = rng.integers(0, 4, rows)
    got = replay.glue().query(kind, num, cat, group_by=gb, threads=3, lifted=True)
    ref = oracle.aggregate(kind, num, cat, group_by=gb)
    for a, b in zip(got, ref):
        assert_struct_parity(a, b, what=f"lifted kind {kind}")
    # and the oracle's own restatement of sum(lift(..)) (fp32, like the reference) agrees
    labels, slots = np.unique(gb, return_inverse=True)
    lifted_ref = oracle.sum_of_lifted(kind, num, cat, group=slots.astype(np.int32), n_groups=len(labels))
    for a, b in zip(got, lifted_ref):
        assert_struct_parity(a, b, rtol=2e-4, what="vs fp32 restatement")


def test_states_spread_over_devices_when_asked():
    """CFB_DEVICES=all: the callbacks place aggregate states round-robin on the visible GPUs and
    SumStateCombine merges them across devices.  (Runs in a subprocess: the policy is read once.)"""
    import os, subprocess, sys
    from duckdb_imputation_b200 import _native as nat
    if nat.lib().cfb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    code = (
        "import numpy as np\n"
        "from duckdb_imputation_b200 import replay\n"
        "from oracle import oracle\n"
        "from tests.parity import assert_struct_parity\n"
        "rng = np.random.default_rng(1); rows = 120_000\n"
        "num = [rng.random(rows).astype(np.float32) for _ in range(6)]\n"
        "cat = [rng.integers(0, 20, rows).astype(np.int32) for _ in range(2)]\n"
        "got = replay.glue().query(0, num, cat, threads=4)\n"
        "assert_struct_parity(got, oracle.aggregate(0, num, cat))\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, CFB_DEVICES="all"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("kind,n,m,G", [(0, 6, 0, 700), (1, 4, 2, 150), (0, 3, 2, 40)])
def test_many_groups_span_several_arenas(kind, n, m, G):
    """More groups than one arena has slots: a worker thread chains arenas (32, 64, 128, ... slots;
    8 per arena when pair tables make a slot large) and ships a chunk once per arena present in it."""
    rng = np.random.default_rng(G)
    rows = 120_000
    num = [rng.random(rows).astype(np.float32) for _ in range(n)]
    cat = [rng.integers(0, 7, rows).astype(np.int32) for _ in range(m)]
    gb = rng.integers(0, G, rows)
    got = replay.glue().query(kind, num, cat, group_by=gb, threads=3)
    ref = oracle.aggregate(kind, num, cat, group_by=gb)
    assert len(got) == len(ref) == len(np.unique(gb))
    for g in (0, 1, len(ref) // 2, len(ref) - 1):
        assert_struct_parity(got[g], ref[g], what=f"group {g} of {G}")
    assert all(a["N"] == b["N"] for a, b in zip(got, ref))


def test_multiply_over_gpu_aggregates(goldens):
    """test_mul.py / test_nb_mul.py end to end: the operands come from OUR aggregates on the GPU, the product
    from OUR multiply_triple / multiply_nb_agg; plus the ring homomorphism on random data:
    multiply(sum A, sum B) == sum over the cross join."""
    g = replay.glue()
    cases = [c for c in goldens["cases"] if c["file"] in ("test_mul.py", "test_nb_mul.py")]
    assert len(cases) == 12
    for c in cases:
        fn = "multiply_triple" if c["file"] == "test_mul.py" else "multiply_nb_agg"
        got = sqlmini.run_mul(c["sql"], goldens["fixtures"][c["file"]], g.query, lambda A, B: g.scalar_structs(fn, A, B),
                              reference_layout=True)
        assert got[c["index"]] == c["expected"], (c["file"], c["test"], c["index"])
    rng = np.random.default_rng(11)
    ra, rb = 700, 300
    An = [rng.integers(0, 6, ra).astype(np.float32) for _ in range(3)]
    Ac = [rng.integers(-2, 4, ra).astype(np.int32) for _ in range(2)]
    Bn = [rng.integers(0, 6, rb).astype(np.float32) for _ in range(2)]
    Bc = [rng.integers(5, 9, rb).astype(np.int32) for _ in range(1)]
    for kind, fn in ((0, "multiply_triple"), (1, "multiply_nb_agg")):
        prod = g.scalar_structs(fn, [g.query(kind, An, Ac)], [g.query(kind, Bn, Bc)])[0]
        whole = g.query(kind, [np.repeat(c, rb) for c in An] + [np.tile(c, ra) for c in Bn],
                        [np.repeat(c, rb) for c in Ac] + [np.tile(c, ra) for c in Bc])
        prod["lin_agg"], prod["quad_agg"] = prod.pop("lin_num"), prod.pop("quad_num")
        assert prod == whole


# ---------------------------------------------------------------- protocol shapes a plain scan does not produce
def _small_table(seed, rows, n=4, m=2):
    rng = np.random.default_rng(seed)
    num = [rng.integers(0, 9, rows).astype(np.float32) for _ in range(n)]  # small ints: fp sums exact in any order
    cat = [rng.integers(-3, 9, rows).astype(np.int32) for _ in range(m)]
    return rng, num, cat


@pytest.mark.parametrize("kind", [0, 1])
def test_simple_update_equals_the_hash_aggregate_protocol(kind):
    """Ungrouped queries go through simple_update (PhysicalUngroupedAggregate: one state per thread, no per-row state
    pointers), which the reference leaves null (ext.cpp:53,106); the result equals the update() protocol's."""
    _, num, cat = _small_table(31, 70_001)
    g = replay.glue()
    for lifted in (False, True):
        fast = g.query(kind, num, cat, threads=3, lifted=lifted)
        with g.options(no_simple=1):
            slow = g.query(kind, num, cat, threads=3, lifted=lifted)
        assert fast == slow
        assert_struct_parity(fast, oracle.aggregate(kind, num, cat), what=f"simple_update kind {kind} lifted {lifted}")


@pytest.mark.parametrize("shape", [1, 2, 3])
@pytest.mark.parametrize("kind", [0, 1])
def test_sum_triple_over_non_flat_struct_vectors(kind, shape):
    """the reference flattens the argument of sum_triple / sum_nb_agg (sum.cpp:72 ->
    utils.cpp:3-18) because a join or filter below the aggregate hands over DICTIONARY / CONSTANT vectors.
    shape 1: DICTIONARY STRUCT; 2: flat STRUCT with DICTIONARY children and leaves; 3: CONSTANT STRUCT (the
    chunk's first lifted row for every row of the chunk -- a CROSS JOIN side)."""
    rng, num, cat = _small_table(32 + shape, 9_000)
    gb = rng.integers(0, 3, len(num[0]))
    g = replay.glue()
    with g.options(lift_shape=shape, no_simple=1):
        got = g.query(kind, num, cat, group_by=gb, threads=2, lifted=True)
        got1 = g.query(kind, num, cat, threads=2, lifted=True)
    if shape == 3:
        first = (np.arange(len(num[0])) // 2048) * 2048  # every row of a chunk carries the chunk's first row
        num, cat = [c[first] for c in num], [c[first] for c in cat]
    ref = oracle.aggregate(kind, num, cat, group_by=gb)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert_struct_parity(a, b, what=f"kind {kind} shape {shape}")
    assert_struct_parity(got1, oracle.aggregate(kind, num, cat), what=f"ungrouped kind {kind} shape {shape}")


@pytest.mark.parametrize("shape", [1, 2, 3])
def test_multiply_over_non_flat_struct_vectors(shape):
    """mul.cpp:24-28 flattens both arguments; here they are read in place whatever their shape."""
    g = replay.glue()
    rng = np.random.default_rng(40 + shape)
    rows = 5
    A, B = [], []
    for r in range(rows):
        ra, rb = 50 + r, 20 + 3 * r
        A.append(g.query(0, [rng.integers(0, 6, ra).astype(np.float32) for _ in range(2)], [rng.integers(0, 3 + r, ra).astype(np.int32)]))
        B.append(g.query(0, [rng.integers(0, 6, rb).astype(np.float32)], [rng.integers(5, 7 + r, rb).astype(np.int32) for _ in range(2)]))
    flat = g.scalar_structs("multiply_triple", A, B)
    with g.options(lift_shape=shape):
        got = g.scalar_structs("multiply_triple", A, B)
    if shape == 3:  # the last argument is a CONSTANT vector: its first row for every row
        flat = g.scalar_structs("multiply_triple", A, [B[0]] * rows)
    assert got == flat


@pytest.mark.parametrize("kind,n,m", [(0, 4, 2), (1, 3, 1), (0, 5, 0)])
def test_two_states_of_one_group_in_one_thread_are_merged(kind, n, m):
    """DuckDB's radix-partitioned hash aggregate can emit a group twice from one thread (hash table reset when
    full) and combine the two states: both live in the SAME arena (context).  split_states = 3: every worker keeps
    three state sets per group and merges them itself."""
    rng, num, cat = _small_table(50 + n, 60_000, n, m)
    gb = rng.integers(0, 5, len(num[0]))
    g = replay.glue()
    with g.options(split_states=3):
        got = g.query(kind, num, cat, group_by=gb, threads=2)
        got1 = g.query(kind, num, cat, threads=2)
    for a, b in zip(got, oracle.aggregate(kind, num, cat, group_by=gb)):
        assert_struct_parity(a, b, what=f"split states kind {kind}")
    assert_struct_parity(got1, oracle.aggregate(kind, num, cat), what="split states, ungrouped")


@pytest.mark.parametrize("kind,n,m", [(0, 4, 2), (1, 6, 2)])
def test_parallel_combine_and_finalize_of_one_arena(kind, n, m):
    """DuckDB finalizes the radix partitions of a hash aggregate in parallel: SumStateCombine / SumStateFinalize run
    concurrently for DIFFERENT groups whose states share the workers' arenas (one cfb_ctx each).  Four threads
    combine + finalize disjoint groups at once; the arena lock serialises the contexts."""
    rng, num, cat = _small_table(60 + n, 200_000, n, m)
    gb = rng.integers(0, 24, len(num[0]))
    g = replay.glue()
    ref = oracle.aggregate(kind, num, cat, group_by=gb)
    for _ in range(3):
        with g.options(parallel_finalize=4):
            got = g.query(kind, num, cat, group_by=gb, threads=4)
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert_struct_parity(a, b, what=f"parallel finalize kind {kind}")


def test_sum_triple_group_by_shares_one_arena_per_thread():
    """sum_triple ... GROUP BY with hundreds of groups: the states of a worker thread are slots of shared contexts
    (round 1 gave every group a private context + stream)."""
    rng, num, cat = _small_table(70, 30_000, 3, 1)
    gb = rng.integers(0, 300, len(num[0]))
    got = replay.glue().query(0, num, cat, group_by=gb, threads=2, lifted=True)
    ref = oracle.aggregate(0, num, cat, group_by=gb)
    assert len(got) == len(ref) == 300
    for a, b in zip(got, ref):
        assert_struct_parity(a, b, what="sum_triple GROUP BY 300")
