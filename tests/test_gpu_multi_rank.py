"""GPU suite for the N>1 path: the exchange step inside the C ABI (cfb_ctx_allreduce over NCCL).

One process per GPU: every rank scans its row range of a `_10_10` table, the categorical domains are agreed with
a MIN/MAX all-reduce (cfb_nccl_agree_domain), the dense states are summed in place by ONE fused NCCL group, and every
rank must hold the oracle's result for the whole table (counts exact, sums <= 1e-5).  States without a dense partial
(undeclared domains; wide key ranges with key dictionaries and hashed pair counts) take the fallback: canonical results
all-gathered and merged by key.  Needs >= 2 GPUs for the 2-rank case; the world-size-1 case runs the same calls on one
GPU."""
import os
import socket

import numpy as np
import pytest

from oracle import oracle
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu


def _table(rows=300_004):
    rng = np.random.default_rng(21)
    num = [rng.random(rows).astype(np.float32) for _ in range(10)]
    # ranks see different key ranges: column 1 is sorted, column 2 has a rare far key
    cat = [rng.integers(0, 100, rows).astype(np.int32) for _ in range(10)]
    cat[1] = (np.arange(rows) * 100 // rows).astype(np.int32)
    cat[2][rows - 5] = 140
    return num, cat


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext, multi_gpu
    from duckdb_imputation_b200 import _native as nat
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # only carries the 128-byte NCCL id
    try:
        torch.cuda.set_device(rank)
        comm = multi_gpu.Communicator(rank)
        num, cat = _table()
        rows = len(num[0])
        lo, hi = multi_gpu.shard_rows(rows, rank, world)
        dn = [torch.from_numpy(c[lo:hi]).cuda() for c in num]
        dc = [torch.from_numpy(c[lo:hi]).cuda() for c in cat]
        m = len(cat)
        a, b = (C.c_int32 * m)(), (C.c_int32 * m)()
        nat.check(nat.lib().cfb_cat_minmax_device(rank, nat.ptr_array([t.data_ptr() for t in dc]), m, hi - lo, a, b, None))
        g_lo, g_hi = comm.agree_domain(list(a), list(b))
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with CofactorContext(CFB_TRIPLE, 10, m, 1, rank) as ctx:
                ctx.set_cat_domain(g_lo, g_hi)
                ctx.scan_device(dn, dc, hi - lo, stream=s.cuda_stream)
                multi_gpu.allreduce_context(ctx, comm)  # default stream argument: torch's current stream
                total = ctx.finalize_arrays()
            # a state whose domain was never declared has no dense partial that means the same on every rank: the
            # ranks exchange canonical results instead (SURVEY 8e fallback); afterwards the context is sealed
            with CofactorContext(CFB_TRIPLE, 10, m, 1, rank) as ctx:
                ctx.scan_device(dn, dc, hi - lo, stream=s.cuda_stream)
                ctx.allreduce(comm.handle, stream=s.cuda_stream)
                undeclared = ctx.finalize_arrays()
                try:
                    ctx.scan_device(dn, dc, hi - lo, stream=s.cuda_stream)
                    refused = False
                except nat.CofactorError as e:
                    refused = e.code == nat.CFB_ERR_STATE
            # wide key ranges: key dictionaries + hashed pair counts, two GROUP BY slots
            wide = [(c.astype(np.int64) * 40_000_003 % 2_000_000_011 - 1_000_000_000).astype(np.int32) for c in cat[:2]]
            dw = [torch.from_numpy(c[lo:hi]).cuda() for c in wide]
            slot = torch.from_numpy((np.arange(lo, hi) % 2).astype(np.int32)).cuda()
            with CofactorContext(CFB_TRIPLE, 3, 2, 2, rank) as ctx:
                ctx.scan_device(dn[:3], dw, hi - lo, d_group=slot, stream=s.cuda_stream)
                ctx.allreduce(comm.handle, stream=s.cuda_stream)
                sparse = [ctx.finalize_arrays(g) for g in range(2)]
        comm.close()
        q.put((rank, g_lo, g_hi, total, refused, undeclared, sparse))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [1, 2])
def test_allreduce_inside_the_library_equals_the_oracle(world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    num, cat = _table()
    whole = oracle.aggregate_arrays(oracle.TRIPLE, num, cat)[0]
    wide = [(c.astype(np.int64) * 40_000_003 % 2_000_000_011 - 1_000_000_000).astype(np.int32) for c in cat[:2]]
    rows = len(num[0])
    # the slot of a row is (its index inside the rank's shard + the shard's first row) % 2 = its global index % 2
    by_slot = oracle.aggregate_arrays(oracle.TRIPLE, num[:3], wide, group=(np.arange(rows) % 2).astype(np.int32), n_groups=2)
    for rank, g_lo, g_hi, total, refused, undeclared, sparse in results:
        assert g_lo == [int(c.min()) for c in cat] and g_hi == [int(c.max()) for c in cat]
        assert_parity(total, whole, what=f"rank {rank} of {world}")  # every rank holds the global triple
        assert_parity(undeclared, whole, what=f"rank {rank} of {world}, undeclared domain (results exchanged)")
        assert refused
        for g in range(2):
            assert_parity(sparse[g], by_slot[g], what=f"rank {rank} of {world}, wide keys, slot {g}")
