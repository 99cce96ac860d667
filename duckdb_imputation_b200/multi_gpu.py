"""Multi-GPU plumbing of the aggregate: one process per GPU.

Rows are range-partitioned; every rank aggregates its slice into a dense state with the SAME
categorical domain; one all-reduce (SUM) of the state (fp64 sums, u64 counts) yields the global triple on
every rank -- the GPU twin of Triple::SumStateCombine across DuckDB threads (sum_state.cpp:10-114).

On GPUs the exchange is INSIDE the C ABI: `Communicator` creates an ncclComm_t through cfb_nccl_comm_create
(the 128-byte id travels over torch.distributed, the only thing Python contributes) and
`allreduce_context` is one call to cfb_ctx_allreduce -- export, collective and import happen in the library,
in place, stream-ordered.  The torch.distributed functions below (`agree_domain`, `allreduce_dense`) are the
same host logic on CPU tensors for the gloo tests, which drive it with oracle-produced partials.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_rows(rows: int, rank: int, world: int, align: int = 4):
    """Contiguous row range [lo, hi) of `rank`; boundaries aligned so every slice stays 16-byte aligned."""
    per = -(-rows // world)
    per += (-per) % align
    lo = min(rows, rank * per)
    return lo, min(rows, lo + per)


def agree_domain(lo, hi, device="cpu"):
    """Element-wise global [min lo, max hi] of the per-rank categorical key ranges."""
    if len(lo) == 0:
        return [], []
    t_lo = torch.tensor(list(lo), dtype=torch.int64, device=device)
    t_hi = torch.tensor(list(hi), dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t_lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(t_hi, op=dist.ReduceOp.MAX)
    return t_lo.tolist(), t_hi.tolist()


def allreduce_dense(f64: torch.Tensor, i64: torch.Tensor):
    """SUM-all-reduce the dense partial state in place (fp64 sums, int64 counts: exact, order-independent)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if f64.numel():
            dist.all_reduce(f64, op=dist.ReduceOp.SUM)
        dist.all_reduce(i64, op=dist.ReduceOp.SUM)
    return f64, i64


class Communicator:
    """An ncclComm_t owned by the C library (one per process / GPU).  Rank 0 draws the NCCL unique id, the id is
    broadcast over the already-initialised torch.distributed group, every rank joins."""

    def __init__(self, device: int):
        import ctypes as C
        from . import _native as nat
        self._nat = nat
        self.device = device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        buf = (C.c_char * 128)()
        if self.rank == 0:
            nat.check(nat.lib().cfb_nccl_unique_id(buf))
        box = [bytes(buf)]
        if self.world > 1:
            dist.broadcast_object_list(box, src=0)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        nat.check(nat.lib().cfb_nccl_comm_create(device, self.world, self.rank, ident, C.byref(self._h)))

    @property
    def handle(self) -> int:
        return self._h.value

    def agree_domain(self, lo, hi, stream: int = 0):
        """Element-wise global [min lo, max hi] over the ranks (cfb_nccl_agree_domain)."""
        import ctypes as C
        m = len(lo)
        a = (C.c_int32 * max(1, m))(*lo)
        b = (C.c_int32 * max(1, m))(*hi)
        self._nat.check(self._nat.lib().cfb_nccl_agree_domain(self._h, self.device, a, b, m, stream or None))
        return list(a)[:m], list(b)[:m]

    def close(self):
        if self._h:
            self._nat.lib().cfb_nccl_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allreduce_context(ctx, comm: "Communicator", stream: int = None):
    """Replace the context's state by the sum over ranks: ONE call into the library (cfb_ctx_allreduce), stream-ordered
    on `stream` -- default: torch's current stream, the one the caller's scans and events are on."""
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    ctx.allreduce(comm.handle, stream=stream)


# ---- dense partial layout on the host (mirror of csrc/state_layout.h) -----------------------
def dense_sizes(kind: int, n: int, m: int, lo, hi):
    dom = [int(h) - int(l) + 1 for l, h in zip(lo, hi)]
    total = sum(dom)
    nq = n if kind == 1 else n * (n + 1) // 2
    pairs = 0 if kind == 1 else sum(dom[k] * dom[l] for k in range(m) for l in range(k + 1, m))
    F = n + nq + (0 if kind == 1 else n * total)
    U = 1 + total + pairs
    return F, U


def pack_dense(a: dict, lo, hi):
    """numpy-form result (struct_result.result_arrays) -> (f64[F], i64[U]) in the device layout."""
    kind, n, m = a["kind"], a["n"], a["m"]
    dom = [int(h) - int(l) + 1 for l, h in zip(lo, hi)]
    cat_off = np.concatenate([[0], np.cumsum(dom)]).astype(np.int64)
    total = int(cat_off[-1])
    F, U = dense_sizes(kind, n, m, lo, hi)
    f = np.zeros(F, np.float64)
    u = np.zeros(U, np.int64)
    nq = len(a["quad"])
    f[:n] = a["lin"]
    f[n:n + nq] = a["quad"]
    u[0] = a["N"]
    offs = a["cat_offsets"]
    dense_t = np.zeros(len(a["cat_keys"]), np.int64)
    for c in range(m):
        for t in range(offs[c], offs[c + 1]):
            dense_t[t] = cat_off[c] + int(a["cat_keys"][t]) - int(lo[c])
            u[1 + dense_t[t]] = a["cat_counts"][t]
    if kind == 0:
        for i in range(n):
            f[n + nq + i * total + dense_t] = a["numcat"][i]
        po = a["pair_offsets"]
        base = 1 + total
        p = 0
        pair_off = 0
        for k in range(m):
            for l in range(k, m):
                if k != l:
                    for t in range(po[p], po[p + 1]):
                        sk = int(a["pair_key1"][t]) - int(lo[k])
                        sl = int(a["pair_key2"][t]) - int(lo[l])
                        u[base + pair_off + sk * dom[l] + sl] = a["pair_counts"][t]
                    pair_off += dom[k] * dom[l]
                p += 1
    return f, u


def unpack_dense(kind: int, n: int, m: int, lo, hi, f: np.ndarray, u: np.ndarray) -> dict:
    """Inverse of pack_dense: dense arrays -> numpy-form result (what cfb_ctx_finalize emits)."""
    dom = [int(h) - int(l) + 1 for l, h in zip(lo, hi)]
    cat_off = np.concatenate([[0], np.cumsum(dom)]).astype(np.int64)
    total = int(cat_off[-1])
    nq = n if kind == 1 else n * (n + 1) // 2
    out = {"kind": kind, "n": n, "m": m, "N": int(u[0]), "lin": f[:n].copy(), "quad": f[n:n + nq].copy()}
    keys, counts, offs, dense_t = [], [], [0], []
    for c in range(m):
        cnt = u[1 + cat_off[c]:1 + cat_off[c + 1]]
        nzs = np.nonzero(cnt)[0]
        keys += [int(lo[c]) + int(s) for s in nzs]
        counts += [int(cnt[s]) for s in nzs]
        dense_t += [int(cat_off[c]) + int(s) for s in nzs]
        offs.append(len(keys))
    out["cat_offsets"] = np.array(offs, np.int64)
    out["cat_keys"] = np.array(keys, np.int32)
    out["cat_counts"] = np.array(counts, np.int64)
    if kind == 0:
        dt = np.array(dense_t, np.int64)
        out["numcat"] = np.stack([f[n + nq + i * total + dt] for i in range(n)]) if n else np.zeros((0, len(keys)))
        k1, k2, pc, po = [], [], [], [0]
        base = 1 + total
        pair_off = 0
        for k in range(m):
            for l in range(k, m):
                if k == l:
                    for t in range(offs[k], offs[k + 1]):
                        k1.append(keys[t]); k2.append(keys[t]); pc.append(counts[t])
                else:
                    tab = u[base + pair_off:base + pair_off + dom[k] * dom[l]].reshape(dom[k], dom[l])
                    for sk, sl in zip(*np.nonzero(tab)):
                        k1.append(int(lo[k]) + int(sk)); k2.append(int(lo[l]) + int(sl)); pc.append(int(tab[sk, sl]))
                    pair_off += dom[k] * dom[l]
                po.append(len(k1))
        out["pair_offsets"] = np.array(po, np.int64)
        out["pair_key1"] = np.array(k1, np.int32)
        out["pair_key2"] = np.array(k2, np.int32)
        out["pair_counts"] = np.array(pc, np.int64)
    return out
