"""CPU suite for the N>1 path: world_size-2 gloo processes run the host side of the multi-GPU
aggregate (row sharding, domain agreement, dense partial layout, all-reduce) on oracle partials and
must reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from duckdb_imputation_b200 import multi_gpu
from oracle import oracle
from tests.parity import assert_parity


def _table(rows=20_001):
    rng = np.random.default_rng(12)
    num = [rng.integers(0, 9, rows).astype(np.float32) for _ in range(4)]  # small ints: sums exact in any order
    cat = [rng.integers(-3, 11, rows).astype(np.int32), (np.arange(rows) // 7000 * 5).astype(np.int32),
           rng.integers(100, 104, rows).astype(np.int32)]
    return num, cat


def _worker(rank, world, port, kind, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        num, cat = _table()
        rows = len(num[0])
        lo, hi = multi_gpu.shard_rows(rows, rank, world)
        assert lo % 4 == 0
        part = oracle.aggregate_arrays(kind, [c[lo:hi] for c in num], [c[lo:hi] for c in cat])[0]
        # ranks see different key ranges (column 1 is sorted): agree on the union before packing
        my_lo = [int(c[lo:hi].min()) for c in cat]
        my_hi = [int(c[lo:hi].max()) for c in cat]
        g_lo, g_hi = multi_gpu.agree_domain(my_lo, my_hi)
        f, u = multi_gpu.pack_dense(part, g_lo, g_hi)
        tf, tu = torch.from_numpy(f), torch.from_numpy(u)
        multi_gpu.allreduce_dense(tf, tu)
        total = multi_gpu.unpack_dense(kind, len(num), len(cat), g_lo, g_hi, tf.numpy(), tu.numpy())
        q.put((rank, (lo, hi), g_lo, g_hi, total))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind", [oracle.TRIPLE, oracle.NB])
def test_two_rank_gloo_reduce_equals_single_process(kind):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    num, cat = _table()
    whole = oracle.aggregate_arrays(kind, num, cat)[0]
    ranges = sorted(r[1] for r in results)
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == len(num[0])
    for rank, _, g_lo, g_hi, total in results:
        assert g_lo == [int(c.min()) for c in cat] and g_hi == [int(c.max()) for c in cat]
        assert_parity(total, whole, rtol=0.0, what=f"rank {rank}")  # every rank holds the global triple


def test_pack_unpack_roundtrip_and_sizes():
    num, cat = _table(3000)
    a = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    lo = [int(c.min()) - 2 for c in cat]
    hi = [int(c.max()) + 3 for c in cat]
    f, u = multi_gpu.pack_dense(a, lo, hi)
    assert (len(f), len(u)) == multi_gpu.dense_sizes(0, 4, 3, lo, hi)
    assert_parity(multi_gpu.unpack_dense(0, 4, 3, lo, hi, f, u), a, rtol=0.0)


def test_shard_rows_covers_everything():
    for rows in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = multi_gpu.shard_rows(rows, r, world)
                assert lo == prev and lo % 4 == 0 or lo == rows
                prev = hi
            assert prev == rows
