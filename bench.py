#!/usr/bin/env python
"""bench.py -- the headline benchmark of the cofactor hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): sum_to_triple_20_0 over 1 B rows of 20 FLOAT columns,
synthetic U[0,1) data generated on the device (counter-based, regenerable on the host).  One
"step" = one aggregate query over the rank's resident table: a fresh aggregate state, one scan,
and for N>1 the NCCL reduce of the per-GPU partial triples inside the C ABI (cfb_ctx_allreduce,
SURVEY 8e).  Weak scaling: every rank holds its own `rows` rows; the `strong` sub-record (N>1) times
the same 1 B rows range-partitioned over the ranks, `c3_multi_gpu` one categorical step (domain
agreement + 3.7 MB reduce).  The table (80 GB at full size) is far larger than L2, so no flush is needed
between iterations.  After the timed region every rank's rows are reduced again by an independent fp64
checker (torch) and compared with the library's result at full size (`check`).  At N=1 the line also
carries `configs`: the other BASELINE.json configs (C1, C3, C4a, C4b, C5) device-resident, end to end
through the DuckDB callbacks, and next to the reference's CPU callbacks.

One JSON line on stdout; everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
from ctypes import c_int32 as C_int32
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NUM = 20
BYTES_PER_ROW = 4 * N_NUM  # algorithmic bytes per row: every input value read once
FULL_ROWS = 1_000_000_000
WORKLOAD = "sum_to_triple_20_0"
SEED = 2


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Only the final JSON line may reach stdout: libraries (NCCL prints its version banner) write to fd 1,
# so fd 1 is pointed at stderr for the whole run and the JSON goes to a private duplicate of the
# original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thr = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU baselines
def cpu_reference_run(rows: int, threads: int, seed: int):
    """Time the reference's CPU algorithm (the ref-faithful oracle port, or oracle/_ref when it was
    built) on `rows` rows of the workload with `threads` worker threads.  -> (seconds, kind)"""
    import numpy as np
    from oracle import oracle
    rng = np.random.default_rng(seed)
    cols = [rng.random(rows, dtype=np.float32) for _ in range(N_NUM)]
    try:
        from oracle import ref_replay
        if ref_replay.available():
            secs = ref_replay.time_sum_to_triple(cols, [], threads)
            return secs, "reference"
    except ImportError:
        pass
    oracle.aggregate_arrays(oracle.TRIPLE, cols, [], mode=oracle.FAITHFUL, threads=threads)
    return oracle.last_seconds(), "port"


def cpu_baseline(budget_s: float = 12.0):
    cores = os.cpu_count() or 1
    probe_rows = 200_000 * max(1, min(cores, 8))
    secs, kind = cpu_reference_run(probe_rows, cores, 1)
    rate = probe_rows / max(secs, 1e-9)
    rows = int(min(64_000_000, max(1_000_000, rate * budget_s)))
    secs, kind = cpu_reference_run(rows, cores, 2)
    return {"value": rows / secs, "unit": "rows/s", "cores": cores, "kind": kind,
            "sample": f"{WORKLOAD} over {rows} rows x {N_NUM} FLOAT U[0,1), 2048-row chunks, {cores} threads + combine"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    probe_rows = 200_000 * max(1, min(cores, 8))
    secs, kind = cpu_reference_run(probe_rows, cores, 1)
    rate = probe_rows / max(secs, 1e-9)
    total_steps = args.steps + args.warmup
    rows = int(min(32_000_000, max(500_000, rate * min(20.0, 150.0 / max(1, total_steps)))))
    for _ in range(args.warmup):
        cpu_reference_run(rows, cores, 3)
    t = 0.0
    for i in range(args.steps):
        s, kind = cpu_reference_run(rows, cores, 4 + i)
        t += s
    val = rows * args.steps / t
    line = {
        "impl": "reference", "metric": "sum_to_triple rows/s", "value": val, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_step": rows, "n_float": N_NUM, "n_int": 0,
                   "note": "reference CPU algorithm on a bounded sample of the workload; rows/s is size-independent"},
        "cpu_baseline": {"value": val, "unit": "rows/s", "cores": cores, "kind": kind,
                         "sample": f"{rows} rows per step, {cores} threads"},
        "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------ the other BASELINE.json configs
# (tag, workload, kind, n, m, domain, GROUP BY slots, full rows, callback name, reference callback name)
# Algorithmic bytes per row: 4 (n + m) (+4 when a group / filter column is read).
CONFIGS = [
    ("C1", "sum_to_triple_5_0, 1 M rows", 0, 5, 0, 0, 1, 1_000_000, "sum_to_triple_5_0", "sum_to_triple_5_0"),
    ("C3", "sum_to_triple_10_10, domain 100, 500 M rows", 0, 10, 10, 100, 1, 500_000_000, "sum_to_triple_10_10", "sum_to_triple_10_10"),
    ("C4a", "sum_to_nb_agg_12_4 GROUP BY label(10), 500 M rows", 1, 12, 4, 100, 10, 500_000_000, "sum_to_nb_agg_12_4", "sum_to_nb_agg_12_4"),
    ("C4b", "sum_to_triple_12_0 GROUP BY label(10) (QDA per-class triples), 500 M rows", 0, 12, 0, 0, 10, 500_000_000,
     "sum_to_triple_12_0", "sum_to_triple_12_0"),
    ("C5", "MICE scan: sum_to_triple_20_10 over the rows where the imputed column is observed (20 % NULL), 100 M rows", 0, 20, 10,
     100, 2, 100_000_000, "sum_to_triple_20_10", "sum_to_triple_19_10"),
]


def train_leg(ctx, n, m):
    """linreg_train / lda_train on the device from a scanned context (cfb_sigma_*): wall ms of the second call of each
    (the first loads the kernels), the gradient descent also per Sigma * theta product."""
    from duckdb_imputation_b200.train import Sigma
    out = {}
    t0 = time.perf_counter()
    with Sigma.from_context(ctx) as s:
        out["p"] = s.p
        out["sigma_from_state_ms"] = (time.perf_counter() - t0) * 1e3
        s.linreg_train(0, 0.001, 0.01, 2)
        t0 = time.perf_counter()
        fit = s.linreg_train(0, 0.001, 0.01, 300)
        dt = time.perf_counter() - t0
        out.update({"linreg_train_ms": dt * 1e3, "linreg_iterations": fit["iterations"], "linreg_products": fit["products"],
                    "linreg_us_per_product": dt * 1e6 / max(1, fit["products"])})
    with Sigma.from_context(ctx, label_cat=m - 1) as s:
        s.lda_train(0.01)
        t0 = time.perf_counter()
        s.lda_train(0.01)
        out["lda_train_ms"] = (time.perf_counter() - t0) * 1e3
        out["lda_classes"] = s.n_classes
    return out


def run_configs(lib, local, stream, peak, scale, e2e_rows):
    """Device-resident rows/s + roofline fraction, full-size integer / fp64 checks against torch (checker only), the
    same aggregate end to end through the DuckDB callbacks from host columns, and the reference's CPU callbacks
    (oracle/_ref) on a bounded sample of the same columns.  One dict per config."""
    import numpy as np
    import torch

    from duckdb_imputation_b200 import CofactorContext, replay, synth
    from duckdb_imputation_b200 import _native as nat
    try:
        from oracle import ref_replay
        ref = ref_replay.ref() if ref_replay.available() else None
    except ImportError:
        ref = None
    cores = os.cpu_count() or 1
    g = replay.glue()
    out = []
    for tag, name, kind, n, m, dom, G, full, fn, ref_fn in CONFIGS:
        t_cfg = time.perf_counter()
        rows = full if full <= 1_000_000 else int(full * scale)
        rows -= rows % 4
        dn = [torch.empty(rows, dtype=torch.float32, device="cuda") for _ in range(n)]
        dc = [torch.empty(rows, dtype=torch.int32, device="cuda") for _ in range(m)]
        for k, t in enumerate(dn):
            nat.check(lib.cfb_gen_uniform_f32(local, t.data_ptr(), rows, synth.column_seed(3, k), 0, None))
        for k, t in enumerate(dc):
            nat.check(lib.cfb_gen_int32(local, t.data_ptr(), rows, synth.column_seed(3, 100 + k), 0, 0, dom, None))
        dg = None
        if G > 1:
            dg = torch.empty(rows, dtype=torch.int32, device="cuda")
            if tag == "C5":  # slot 0: the imputed column is observed (80 %), slot 1: it is NULL
                nat.check(lib.cfb_gen_int32(local, dg.data_ptr(), rows, 4242, 0, 0, 5, None))
                dg = (dg == 0).to(torch.int32)
            else:
                nat.check(lib.cfb_gen_int32(local, dg.data_ptr(), rows, 777, 0, 0, G, None))
        torch.cuda.synchronize()
        times, res = [], None
        reps = 6 if rows <= 1_000_000 else 5
        for rep in range(reps):
            with CofactorContext(kind, n, m, G, local) as ctx:
                if m:
                    ctx.set_cat_domain([0] * m, [dom - 1] * m)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.scan_device(dn, dc, rows, d_group=dg, stream=stream.cuda_stream)
                e1.record(stream)
                ctx.sync()
                torch.cuda.synchronize()
                if rep >= 2:
                    times.append(e0.elapsed_time(e1))
                if rep == reps - 1:
                    res = [ctx.finalize_arrays(gg) for gg in range(G)]
                    if tag == "C3":  # SURVEY 8 f4 on the state just scanned: sigma on the device, both trainers
                        train_rec = train_leg(ctx, n, m)
        ms = sum(times) / len(times)
        bpr = 4 * (n + m + (1 if G > 1 else 0))
        rec = {"config": tag, "workload": name, "rows": rows, "ms_per_scan": ms, "rows_per_s": rows / ms * 1e3,
               **({"train": train_rec} if tag == "C3" else {}),
               "bytes_per_row": bpr, "roofline": {"bound": "hbm", "achieved": rows * bpr / ms / 1e6, "peak": peak, "unit": "GB/s",
                                                    "frac": rows * bpr / ms / 1e6 / peak}}
        # ---- full-size checks of the scan just timed, against torch reductions (exact integers; fp64 sums)
        chk = {}
        slot = dg if dg is not None else None
        Ns = torch.bincount(slot, minlength=G).cpu().numpy() if slot is not None else np.array([rows])
        chk["N_ok"] = bool(all(int(res[gg]["N"]) == int(Ns[gg]) for gg in range(G)))
        worst = 0.0
        for gg in range(min(G, 2)):
            mask = None if slot is None else (slot == gg)
            if n:
                x0 = dn[0] if mask is None else dn[0][mask]
                xl = dn[n - 1] if mask is None else dn[n - 1][mask]
                lin0 = float(x0.double().sum())
                worst = max(worst, abs(res[gg]["lin"][0] - lin0) / abs(lin0))
                q = float((x0.double() * xl.double()).sum())
                got_q = res[gg]["quad"][n - 1] if kind == 0 else None  # packed triangle: (0, n-1) is entry n-1
                if got_q is not None:
                    worst = max(worst, abs(got_q - q) / abs(q))
            if m:
                k0 = dc[0] if mask is None else dc[0][mask]
                cnt = torch.bincount(k0, minlength=dom).cpu().numpy()
                off = res[gg]["cat_offsets"]
                chk["key_counts_ok"] = chk.get("key_counts_ok", True) and bool(
                    np.array_equal(res[gg]["cat_counts"][off[0]:off[1]], cnt[cnt > 0]))
                if kind == 0 and n:
                    x0 = dn[0] if mask is None else dn[0][mask]
                    sums = torch.zeros(dom, dtype=torch.float64, device="cuda").index_add_(0, k0.long(), x0.double()).cpu().numpy()
                    got = res[gg]["numcat"][0][off[0]:off[1]]
                    worst = max(worst, float(np.max(np.abs(got - sums[cnt > 0]) / np.abs(sums[cnt > 0]))))
                if kind == 0 and m >= 2:
                    k1 = dc[m - 1] if mask is None else dc[m - 1][mask]
                    pc = torch.bincount(k0.long() * dom + k1.long(), minlength=dom * dom).cpu().numpy()
                    po = res[gg]["pair_offsets"]
                    p = m - 1  # pair lists are in (k<=l) row-major order: (0,0), (0,1), ..., (0,m-1)
                    chk["pair_counts_ok"] = chk.get("pair_counts_ok", True) and bool(
                        np.array_equal(res[gg]["pair_counts"][po[p]:po[p + 1]], pc[pc > 0]))
        chk["sums_max_rel_err"] = worst
        chk["checker"] = "torch bincount / fp64 sums over all %d rows (first and last column, slots 0-1)" % rows
        assert chk["N_ok"] and chk.get("key_counts_ok", True) and chk.get("pair_counts_ok", True) and worst < 1e-5, (tag, chk)
        rec["check"] = chk
        del dn, dc, dg, slot
        torch.cuda.empty_cache()
        # ---- end to end through the callbacks (host columns), and the reference's CPU callbacks beside it
        er = min(e2e_rows, rows)
        rng = np.random.default_rng(11)
        hn = [rng.random(er, dtype=np.float32) for _ in range(n)]
        hc = [rng.integers(0, max(dom, 1), er).astype(np.int32) for _ in range(m)]
        hg = None
        sel = None
        kw = dict(threads=cores)
        if tag == "C5":
            sel = np.nonzero(rng.integers(0, 5, er) != 0)[0].astype(np.uint32)  # WHERE col IS NOT NULL: 80 % of the rows
            kw["sel"] = sel
        elif G > 1:
            hg = rng.integers(0, G, er).astype(np.int32)
            kw.update(group=hg, n_groups=G)
        eff_rows = er if sel is None else len(sel)
        g.aggregate(fn, hn, hc, **kw)  # cold: contexts, staging tiles, kernels load here
        cold_s = g.last_seconds
        best = 1e30
        for _ in range(2):
            r_ = g.aggregate(fn, hn, hc, **kw)
            best = min(best, g.last_seconds)
        assert sum(x["N"] for x in r_) == eff_rows
        rec["e2e"] = {"value": eff_rows / best, "unit": "rows/s", "rows": eff_rows, "host_threads": cores, "seconds": best,
                      "first_call_seconds": cold_s, "h2d_bytes": eff_rows * bpr,
                      "path": "host columns -> %s callbacks (replay host, %d threads) -> staging -> kernels -> STRUCT" % (fn, cores)}
        if ref is not None:
            nn, mm = (19, 10) if tag == "C5" else (n, m)
            rr = min(er, 1_000_000 if mm >= 2 else 8_000_000)
            kw2 = dict(threads=cores)
            if tag == "C5":
                kw2["sel"] = sel[sel < rr]
            elif G > 1:
                kw2.update(group=hg[:rr], n_groups=G)
            ref.aggregate(ref_fn, [c[:rr] for c in hn[:nn]], [c[:rr] for c in hc[:mm]], **kw2)
            eff = rr if "sel" not in kw2 else len(kw2["sel"])
            rec["cpu_baseline"] = {"value": eff / ref.last_seconds, "unit": "rows/s", "cores": cores, "kind": "reference",
                                   "seconds": ref.last_seconds,
                                   "sample": "%s over %d rows, %d threads (oracle/_ref: the reference's own callbacks)" % (ref_fn, eff, cores)}
        rec["wall_s"] = time.perf_counter() - t_cfg
        log(f"[config {tag}] {rec['rows_per_s'] / 1e9:.2f} G rows/s device-resident (frac {rec['roofline']['frac']:.3f}), "
            f"e2e {rec['e2e']['value'] / 1e6:.0f} M rows/s, cpu {rec.get('cpu_baseline', {}).get('value', 0) / 1e6:.2f} M rows/s")
        out.append(rec)
        del hn, hc
    return out


def fp64_checker(cols, rows, chunk=4_000_000):
    """Independent fp64 reduction of a resident table: lin = sum x, gram = X^T X, by torch in chunks (cuBLAS as the
    CHECKER, never the thing measured).  -> (lin[n], gram[n, n]) device tensors."""
    import torch
    n = len(cols)
    lin = torch.zeros(n, dtype=torch.float64, device=cols[0].device)
    gram = torch.zeros(n, n, dtype=torch.float64, device=cols[0].device)
    for lo in range(0, rows, chunk):
        X = torch.stack([c[lo:min(rows, lo + chunk)] for c in cols], dim=1).double()
        lin += X.sum(dim=0)
        gram += X.T @ X
        del X
    return lin, gram


# ------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=FULL_ROWS, help="rows per GPU (default: the 1 B of BASELINE.json)")
    ap.add_argument("--e2e-rows", type=int, default=64_000_000)
    ap.add_argument("--config-scale", type=float, default=1.0, help="row-count scale of the extra configs (C3..C5)")
    ap.add_argument("--config-e2e-rows", type=int, default=24_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3) if args.warmup >= 0 else 3

    import numpy as np
    import torch
    import torch.distributed as dist

    from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext, multi_gpu, synth
    from duckdb_imputation_b200 import _native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = multi_gpu.Communicator(local)  # the library's own ncclComm_t: the exchange runs inside the C ABI
    lib = nat.lib()
    dev = torch.device("cuda", local)
    peak, peak_src = measured_peak_gbs()

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident table
    free, _total = torch.cuda.mem_get_info(dev)
    rows = min(args.rows, int(free * 0.88) // BYTES_PER_ROW)
    rows -= rows % 4
    if world > 1:
        t = torch.tensor([rows], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        rows = int(t.item())
    log(f"[rank {rank}] rows per GPU = {rows} ({rows * BYTES_PER_ROW / 1e9:.1f} GB), steps={args.steps} warmup={args.warmup}")
    cols = [torch.empty(rows, dtype=torch.float32, device=dev) for _ in range(N_NUM)]
    first = rank * rows  # rank r holds rows [r*rows, (r+1)*rows) of the global synthetic table
    for k, c in enumerate(cols):
        nat.check(lib.cfb_gen_uniform_f32(local, c.data_ptr(), rows, synth.column_seed(SEED, k), first, None))
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(dev)  # explicit (non-default) stream: kernels, NCCL and the timing events share it
    torch.cuda.set_stream(stream)
    total = args.steps + args.warmup
    ctxs = [CofactorContext(CFB_TRIPLE, N_NUM, 0, 1, local) for _ in range(total)]

    def step(i, ev=None):
        ctx = ctxs[i]
        if ev:
            ev[0].record(stream)
        ctx.scan_device(cols, [], rows, stream=stream.cuda_stream)
        if ev:
            ev[1].record(stream)
        if world > 1:
            # the one exchange step of the path: sum the per-GPU partial triples (fp64 sums, u64 counts), in place,
            # one fused NCCL group issued by the library (cfb_ctx_allreduce)
            ctx.allreduce(comm.handle, stream=stream.cuda_stream)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.cfb_kernel_launches()
    torch.cuda.synchronize()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i, kev[i])
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lib.cfb_kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = allmax(e0.elapsed_time(e1))
    per_step = [a.elapsed_time(b) for a, b in kev]
    log(f"[rank {rank}] scan kernel ms per step: " + " ".join(f"{t:.2f}" for t in per_step))
    kernel_ms = sum(per_step) / args.steps
    ms_per_step = ms_total / args.steps
    value = world * rows / (ms_per_step * 1e-3)

    # ---- parity of what was just timed, at FULL size, against an independent fp64 reduction (torch / cuBLAS as the
    # checker): every rank reduces its own rows, the checker's sums are all-reduced by torch.distributed, and the
    # library's (all-reduced) lin / quad must agree to <= 1e-5 relative.
    check = {}
    last = ctxs[-1].finalize_arrays()
    check["N_ok"] = bool(last["N"] == rows * world)
    iu = np.triu_indices(N_NUM)
    c_lin, c_gram = fp64_checker(cols, rows)
    own = None
    if world > 1:
        # this rank's own partial through the library (no exchange), for the all-reduce check below
        with CofactorContext(CFB_TRIPLE, N_NUM, 0, 1, local) as c:
            c.scan_device(cols, [], rows, stream=stream.cuda_stream)
            own = c.finalize_arrays()
        part = torch.from_numpy(np.concatenate([own["lin"], own["quad"]])).to(dev)
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        summed = part.cpu().numpy()
        lib_all = np.concatenate([last["lin"], last["quad"]])
        # cfb_ctx_allreduce (NCCL inside the library) against torch.distributed's sum of the same per-rank partials
        check["allreduce_max_rel_err"] = float(np.max(np.abs(lib_all - summed) / np.abs(summed)))
        dist.all_reduce(c_lin, op=dist.ReduceOp.SUM)
        dist.all_reduce(c_gram, op=dist.ReduceOp.SUM)
    ref_lin, ref_quad = c_lin.cpu().numpy(), c_gram.cpu().numpy()[iu]
    check["full_rows"] = rows * world
    check["full_max_rel_err"] = max(float(np.max(np.abs(last["quad"] - ref_quad) / np.abs(ref_quad))),
                                    float(np.max(np.abs(last["lin"] - ref_lin) / np.abs(ref_lin))))
    check["checker"] = "chunked X.double().T @ X.double() over every resident row (torch), summed over ranks"
    assert check["N_ok"] and check["full_max_rel_err"] < 1e-5, f"full-size parity failed: {check}"
    assert check.get("allreduce_max_rel_err", 0.0) < 1e-12, f"all-reduce parity failed: {check}"
    del c_lin, c_gram
    for c in ctxs:
        c.close()

    # ---- strong scaling (N > 1): the 1 B-row table range-partitioned N ways -- every rank scans rows/N rows, then the
    # exchange; fixed costs (launch, last-CTA fold, NCCL) are no longer hidden behind a 13 ms scan.
    strong = None
    if world > 1:
        srows = (min(args.rows, rows * world) // world) // 4 * 4
        scols = [c[:srows] for c in cols]
        sctx = [CofactorContext(CFB_TRIPLE, N_NUM, 0, 1, local) for _ in range(args.steps + 3)]
        sev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        for i in range(3):
            sctx[i].scan_device(scols, [], srows, stream=stream.cuda_stream)
            sctx[i].allreduce(comm.handle, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for i in range(args.steps):
            sev[i][0].record(stream)
            sctx[3 + i].scan_device(scols, [], srows, stream=stream.cuda_stream)
            sev[i][1].record(stream)
            sctx[3 + i].allreduce(comm.handle, stream=stream.cuda_stream)
            sev[i][2].record(stream)
        s1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        s_ms = allmax(s0.elapsed_time(s1)) / args.steps
        scan_us = allmax(sum(e[0].elapsed_time(e[1]) for e in sev) / args.steps * 1e3)
        coll_us = allmax(sum(e[1].elapsed_time(e[2]) for e in sev) / args.steps * 1e3)
        got = sctx[-1].finalize_arrays()
        sl, sg = fp64_checker(scols, srows)
        dist.all_reduce(sl, op=dist.ReduceOp.SUM)
        dist.all_reduce(sg, op=dist.ReduceOp.SUM)
        s_err = max(float(np.max(np.abs(got["quad"] - sg.cpu().numpy()[iu]) / np.abs(sg.cpu().numpy()[iu]))),
                    float(np.max(np.abs(got["lin"] - sl.cpu().numpy()) / np.abs(sl.cpu().numpy()))))
        assert got["N"] == srows * world and s_err < 1e-5, f"strong-scaling parity failed: {s_err}"
        strong = {"rows_total": srows * world, "rows_per_gpu": srows, "ms_per_step": s_ms, "rows_per_s": srows * world / (s_ms * 1e-3),
                  "scan_us": scan_us, "collective_us": coll_us,
                  "efficiency": (kernel_ms * srows / rows) / s_ms,
                  "efficiency_definition": "(this run's 1-GPU scan time for rows/N rows, scaled from the weak step) / (strong step time)",
                  "full_max_rel_err": s_err}
        for c in sctx:
            c.close()
        del scols
    del cols
    torch.cuda.empty_cache()

    # ---- C3 across GPUs (N > 1): domain agreement (MIN/MAX all-reduce) + scan + 3.7 MB all-reduce of the dense state
    c3_multi = None
    if world > 1:
        m3, dom3 = 10, 100
        r3 = (min(500_000_000, int(500_000_000 * args.config_scale)) // world) // 4 * 4
        dn = [torch.empty(r3, dtype=torch.float32, device=dev) for _ in range(10)]
        dc = [torch.empty(r3, dtype=torch.int32, device=dev) for _ in range(m3)]
        for k, t in enumerate(dn):
            nat.check(lib.cfb_gen_uniform_f32(local, t.data_ptr(), r3, synth.column_seed(3, k), rank * r3, None))
        for k, t in enumerate(dc):
            nat.check(lib.cfb_gen_int32(local, t.data_ptr(), r3, synth.column_seed(3, 100 + k), rank * r3, 0, dom3, None))
        torch.cuda.synchronize()
        times, coll = [], []
        got = None
        for rep in range(5):
            with CofactorContext(CFB_TRIPLE, 10, m3, 1, local) as ctx:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                dist.barrier()
                t_host = time.perf_counter()
                lo = (C_int32 * m3)()
                hi = (C_int32 * m3)()
                nat.check(lib.cfb_cat_minmax_device(local, nat.ptr_array([t.data_ptr() for t in dc]), m3, r3, lo, hi, stream.cuda_stream))
                glo, ghi = comm.agree_domain(list(lo), list(hi), stream=stream.cuda_stream)
                ctx.set_cat_domain(glo, ghi)
                agree_ms = (time.perf_counter() - t_host) * 1e3
                ev[0].record(stream)
                ctx.scan_device(dn, dc, r3, stream=stream.cuda_stream)
                ev[1].record(stream)
                ctx.allreduce(comm.handle, stream=stream.cuda_stream)
                ev[2].record(stream)
                ctx.sync()
                torch.cuda.synchronize()
                if rep >= 2:
                    times.append(agree_ms + ev[0].elapsed_time(ev[2]))
                    coll.append(ev[1].elapsed_time(ev[2]))
                if rep == 4:
                    got = ctx.finalize_arrays()
        cnt = torch.bincount(dc[0], minlength=dom3)
        pc = torch.bincount(dc[0].long() * dom3 + dc[m3 - 1].long(), minlength=dom3 * dom3)
        sums = torch.zeros(dom3, dtype=torch.float64, device=dev).index_add_(0, dc[0].long(), dn[0].double())
        for t in (cnt, pc, sums):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        off, po = got["cat_offsets"], got["pair_offsets"]
        cn, pn, sn = cnt.cpu().numpy(), pc.cpu().numpy(), sums.cpu().numpy()
        ok = (got["N"] == r3 * world and np.array_equal(got["cat_counts"][off[0]:off[1]], cn[cn > 0])
              and np.array_equal(got["pair_counts"][po[m3 - 1]:po[m3]], pn[pn > 0]))
        err = float(np.max(np.abs(got["numcat"][0][off[0]:off[1]] - sn[cn > 0]) / np.abs(sn[cn > 0])))
        assert ok and err < 1e-5, f"C3 multi-GPU parity failed (counts ok={ok}, sums err={err})"
        ms3 = allmax(sum(times) / len(times))
        c3_multi = {"workload": "sum_to_triple_10_10 domain 100, %d rows over %d GPUs" % (r3 * world, world), "ms_per_step": ms3,
                    "rows_per_s": r3 * world / (ms3 * 1e-3), "collective_us": allmax(sum(coll) / len(coll) * 1e3),
                    "counts_exact": bool(ok), "sums_max_rel_err": err,
                    "step": "min/max pre-pass + MIN/MAX all-reduce (domain agreement) + scan + all-reduce of the dense state"}
        del dn, dc
        torch.cuda.empty_cache()

    # ---- end to end: host (DuckDB-side) buffers -> the extension's aggregate callbacks under DuckDB's
    # protocol (T threads x 2048-row chunks, update/combine/finalize) -> pinned staging ->
    # cudaMemcpyAsync -> kernels -> result STRUCT back on the host.
    e2e = None
    if not args.no_e2e:
        from duckdb_imputation_b200 import replay
        er = args.e2e_rows
        threads = max(1, (os.cpu_count() or 1) // world)
        rng = np.random.default_rng(SEED + rank)
        host = [rng.random(er, dtype=np.float32) for _ in range(N_NUM)]
        os.environ["CFB_DEVICE"] = str(local)
        g = replay.glue()

        def e2e_step():
            res = g.aggregate(f"sum_to_triple_{N_NUM}_0", host, [], threads=threads)
            return res[0]["N"]

        for _ in range(2):
            e2e_step()
        if world > 1:
            dist.barrier()
        esteps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(esteps):
            assert e2e_step() == er
        dt = allmax(time.perf_counter() - t0)
        # the host-memory ceiling of the staging copy on this box with the same thread count on every rank at once
        ceil_nt = lib.cfb_host_copy_ceiling(256 << 20, threads, 1, 2)
        if world > 1:
            t = torch.tensor([ceil_nt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ceil_nt = float(t.item())
        e2e = {"value": world * er * esteps / dt, "unit": "rows/s", "h2d_bytes_per_step": er * BYTES_PER_ROW,
               "d2h_bytes_per_step": 8 * (1 + N_NUM + N_NUM * (N_NUM + 1) // 2) * threads,
               "rows_per_step": er, "steps": esteps, "host_threads": threads,
               "gb_per_s": world * er * esteps * BYTES_PER_ROW / dt / 1e9,
               "host_copy_ceiling_gb_per_s": ceil_nt, "stage_isa": lib.cfb_stage_isa().decode(),
               # bytes the host memory system moves: the feed reads the source, writes the staging tile and the DMA engine
               # reads it again (3 per payload byte); the copy-only ceiling moves 2 per payload byte
               "host_dram_traffic_gb_per_s": 3 * world * er * esteps * BYTES_PER_ROW / dt / 1e9,
               "host_copy_ceiling_traffic_gb_per_s": 2 * ceil_nt,
               "host_copy_ceiling_note": "aggregate GB/s of the staging copy alone (non-temporal stores, 8 KB pieces, private "
                                         "buffers, %d threads x %d ranks at the same time): the host-memory bound of the feed" % (threads, world),
               "path": "pageable host columns -> DuckDB aggregate callbacks (replay host: %d threads, 2048-row chunks, "
                       "simple_update) -> cfb_ctx_append (pinned double-buffered staging, cudaMemcpyAsync) -> combine -> finalize" % threads}
        del host

    if rank != 0:
        if world > 1:
            comm.close()
            dist.destroy_process_group()
        return 0
    achieved = rows * BYTES_PER_ROW / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic per launch: ncu --set full on this kernel (profiles/r02_gram_ncu_summary.txt; the same in round 1,
    # profiles/r01_gram_final_ncu_summary.txt) measured dram__bytes_read + dram__bytes_write = 20.0004 GB + 4.4 MB for
    # 20.0000 GB of algorithmic bytes (250 M rows); the kernel has no size-dependent re-reads, so the ratio is applied
    # to this launch's algorithmic bytes.
    NCU_TRAFFIC_RATIO = (20.000386e9 + 4.380928e6) / 20.0e9
    line = {
        "metric": "sum_to_triple rows/s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_gpu": rows, "rows_total": rows * world, "n_float": N_NUM, "n_int": 0,
                   "bytes_per_row": BYTES_PER_ROW, "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush" % (rows * BYTES_PER_ROW / 1e9),
                   "exchange": "none" if world == 1 else "cfb_ctx_allreduce: one fused NCCL group (fp64 sums + u64 counts), in place, per step",
                   "accumulate": "fp32x2 FMA over bounded runs, folded into fp64"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": rows * BYTES_PER_ROW * NCU_TRAFFIC_RATIO, "algorithmic_bytes": rows * BYTES_PER_ROW,
                     "traffic_source": "ncu --set full, profiles/r02_gram_ncu_summary.txt (bytes per launch, scaled by rows)", "kernel": "cfb::gram_scan_kernel<20,false,768>", "kernel_ms": kernel_ms,
                     "peak_source": peak_src, "hbm_gbs_whole_step": world * rows * BYTES_PER_ROW / (ms_per_step * 1e-3) / 1e9},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "check": check,
    }
    if strong:
        line["strong"] = strong
    if c3_multi:
        line["c3_multi_gpu"] = c3_multi
    if world == 1 and not args.no_configs:
        line["configs"] = run_configs(lib, local, stream, peak, args.config_scale, args.config_e2e_rows)
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    if world > 1:
        comm.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
