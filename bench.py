#!/usr/bin/env python
"""bench.py -- the headline benchmark of the cofactor hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): sum_to_triple_20_0 over 1 B rows of 20 FLOAT columns,
synthetic U[0,1) data generated on the device (counter-based, regenerable on the host).  One
"step" = one aggregate query over the rank's resident table: a fresh aggregate state, one scan,
and for N>1 the NCCL reduce of the per-GPU partial triples.  Weak scaling: every rank
holds its own `rows` rows.  The table (80 GB at full size) is far larger than L2, so no flush is
needed between iterations.

One JSON line on stdout; everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
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


# ------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=FULL_ROWS, help="rows per GPU (default: the 1 B of BASELINE.json)")
    ap.add_argument("--e2e-rows", type=int, default=64_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = nat.lib()
    dev = torch.device("cuda", local)

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
    nf, nu = ctxs[0].partial_sizes()
    pf = torch.zeros(nf, dtype=torch.float64, device=dev)
    pu = torch.zeros(nu, dtype=torch.int64, device=dev)

    def step(i, ev=None):
        ctx = ctxs[i]
        if ev:
            ev[0].record(stream)
        ctx.scan_device(cols, [], rows, stream=stream.cuda_stream)
        if ev:
            ev[1].record(stream)
        if world > 1:
            # the one exchange step of the path: sum the per-GPU partial triples (fp64 sums, int64 counts)
            multi_gpu.allreduce_context(ctx, pf, pu, stream=stream.cuda_stream)

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
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    per_step = [a.elapsed_time(b) for a, b in kev]
    log(f"[rank {rank}] scan kernel ms per step: " + " ".join(f"{t:.2f}" for t in per_step))
    kernel_ms = sum(per_step) / args.steps
    ms_per_step = ms_total / args.steps
    value = world * rows / (ms_per_step * 1e-3)

    # ---- parity of what was just timed (rank 0)
    check = {}
    last = ctxs[-1].finalize_arrays()
    check["N_ok"] = bool(last["N"] == rows * world)
    if rank == 0:
        # size-independent checks of what was just timed (the CPU oracle is only used by the cpu_baseline
        # leg and by tests/; here the checker is an fp64 reduction of a prefix by torch on the device)
        pre = min(rows, 1_000_000)
        with CofactorContext(CFB_TRIPLE, N_NUM, 0, 1, local) as c:
            c.scan_device([t[:pre] for t in cols], [], pre, stream=stream.cuda_stream)
            got = c.finalize_arrays()
        X = torch.stack([t[:pre] for t in cols], dim=1).double()
        gram = (X.T @ X).cpu().numpy()
        lin = X.sum(dim=0).cpu().numpy()
        iu = np.triu_indices(N_NUM)
        rel = max(float(np.max(np.abs(got["quad"] - gram[iu]) / np.abs(gram[iu]))), float(np.max(np.abs(got["lin"] - lin) / np.abs(lin))))
        assert got["N"] == pre and rel < 1e-5, f"prefix parity failed: rel err {rel}"
        check["prefix_rows"] = pre
        check["prefix_max_rel_err"] = rel
        del X
        if world == 1:
            h = (rows // 2) - (rows // 2) % 4
            with CofactorContext(CFB_TRIPLE, N_NUM, 0, 1, local) as c:
                c.scan_device([t[:h] for t in cols], [], h, stream=stream.cuda_stream)
                c.scan_device([t[h:] for t in cols], [], rows - h, stream=stream.cuda_stream)
                halves = c.finalize_arrays()
            check["halves_vs_whole_rel"] = float(np.max(np.abs(halves["quad"] - last["quad"]) / np.abs(last["quad"])))
            assert check["halves_vs_whole_rel"] < 1e-6 and halves["N"] == last["N"]
            # E[x]=1/2 for U[0,1): a distribution-level sanity check at full size
            check["mean_lin"] = float(np.mean(last["lin"]) / last["N"])
    for c in ctxs:
        c.close()

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
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * er * esteps / dt, "unit": "rows/s", "h2d_bytes_per_step": er * BYTES_PER_ROW,
               "d2h_bytes_per_step": 8 * (1 + N_NUM + N_NUM * (N_NUM + 1) // 2) * threads,
               "rows_per_step": er, "steps": esteps, "host_threads": threads,
               "path": "pageable host columns -> DuckDB aggregate callbacks (replay host: %d threads, 2048-row chunks) -> "
                       "cfb_ctx_append (pinned double-buffered staging, cudaMemcpyAsync) -> combine -> finalize" % threads}
        del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak, peak_src = measured_peak_gbs()
    achieved = rows * BYTES_PER_ROW / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic per launch: ncu --set full on this kernel (profiles/r01_gram_final_ncu_summary.txt) measured
    # dram__bytes_read + dram__bytes_write = 20.0052 GB for 20.0000 GB of algorithmic bytes (250 M rows);
    # the kernel has no size-dependent re-reads, so the ratio is applied to this launch's algorithmic bytes.
    NCU_TRAFFIC_RATIO = (20.000506e9 + 4.665088e6) / 20.0e9
    line = {
        "metric": "sum_to_triple rows/s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_gpu": rows, "rows_total": rows * world, "n_float": N_NUM, "n_int": 0,
                   "bytes_per_row": BYTES_PER_ROW, "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush" % (rows * BYTES_PER_ROW / 1e9),
                   "exchange": "none" if world == 1 else "NCCL all_reduce of the partial triple (fp64 sums + int64 counts) per step",
                   "accumulate": "fp32x2 FMA over bounded runs, folded into fp64"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": rows * BYTES_PER_ROW * NCU_TRAFFIC_RATIO, "algorithmic_bytes": rows * BYTES_PER_ROW,
                     "traffic_source": "ncu --set full, profiles/r01_gram_final_ncu_summary.txt (bytes per launch, scaled by rows)", "kernel": "cfb::gram_scan_kernel<20,false,768>", "kernel_ms": kernel_ms,
                     "peak_source": peak_src, "hbm_gbs_whole_step": world * rows * BYTES_PER_ROW / (ms_per_step * 1e-3) / 1e9},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "check": check,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
