"""One JSON line per BASELINE.json config (device-resident), for profiles/: rows/s, GB/s, fraction of
the measured HBM peak.  Full sizes by default (needs ~80 GB of HBM for C2)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from duckdb_imputation_b200 import CFB_NB, CFB_TRIPLE, CofactorContext, synth
from duckdb_imputation_b200 import _native as nat

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None  # e.g. C3,C5
lib = nat.lib()
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6551.0
CONFIGS = [
    ("C1", "sum_to_triple_5_0", CFB_TRIPLE, 5, 0, 0, 1, 1_000_000),
    ("C2", "sum_to_triple_20_0", CFB_TRIPLE, 20, 0, 0, 1, 1_000_000_000),
    ("C3", "sum_to_triple_10_10 (domain 100)", CFB_TRIPLE, 10, 10, 100, 1, 500_000_000),
    ("C3z", "sum_to_triple_10_10, Zipf(1.2) keys in [0,100) (hot keys / hot cells)", CFB_TRIPLE, 10, 10, 100, 1, 50_000_000),
    ("C3w", "sum_to_triple_10_2, domain 100 000 (pair counts in the hashed table)", CFB_TRIPLE, 10, 2, 100_000, 1, 50_000_000),
    ("C4a", "sum_to_nb_agg_12_4 GROUP BY label(10)", CFB_NB, 12, 4, 100, 10, 500_000_000),
    ("C4b", "sum_to_triple_12_0 GROUP BY label(10) (QDA per-class triples)", CFB_TRIPLE, 12, 0, 0, 10, 500_000_000),
    ("C5", "sum_to_triple_20_10 WHERE not null (MICE scan, 2-slot filter)", CFB_TRIPLE, 20, 10, 100, 2, 100_000_000),
]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for tag, name, kind, n, m, dom, G, full in CONFIGS:
    if only and tag not in only:
        continue
    rows = full if full <= 1_000_000 else int(full * scale)
    rows -= rows % 4
    dn = [torch.empty(rows, dtype=torch.float32, device="cuda") for _ in range(n)]
    dc = [torch.empty(rows, dtype=torch.int32, device="cuda") for _ in range(m)]
    for k, t in enumerate(dn):
        nat.check(lib.cfb_gen_uniform_f32(0, t.data_ptr(), rows, synth.column_seed(3, k), 0, None))
    for k, t in enumerate(dc):
        if tag == "C3z":  # host-generated skew (numpy), copied over
            import numpy as np
            t.copy_(torch.from_numpy(np.minimum(np.random.default_rng(k).zipf(1.2, rows) - 1, dom - 1).astype(np.int32)))
        else:
            nat.check(lib.cfb_gen_int32(0, t.data_ptr(), rows, synth.column_seed(3, 100 + k), 0, 0, dom, None))
    dg = None
    if G > 1:
        dg = torch.empty(rows, dtype=torch.int32, device="cuda")
        if tag == "C5":  # 20% of the rows are NULL in the imputed column -> slot 1
            nat.check(lib.cfb_gen_int32(0, dg.data_ptr(), rows, 4242, 0, 0, 5, None))
            dg = (dg == 0).to(torch.int32)
        else:
            nat.check(lib.cfb_gen_int32(0, dg.data_ptr(), rows, 777, 0, 0, G, None))
    torch.cuda.synchronize()
    times = []
    for rep in range(5):
        with CofactorContext(kind, n, m, G) as ctx:
            if m:
                ctx.set_cat_domain([0] * m, [dom - 1] * m)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.scan_device(dn, dc, rows, d_group=dg, stream=stream.cuda_stream)
            e1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            N = sum(ctx.finalize_arrays(g)["N"] for g in range(G))
            assert N == rows or os.environ.get('CFB_ROLE_DEBUG')
            if rep >= 2:
                times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    bpr = 4 * (n + m + (1 if G > 1 else 0))
    print(json.dumps({"config": tag, "workload": name, "rows": rows, "ms": ms, "rows_per_s": rows / ms * 1e3,
                      "best_rows_per_s": rows / min(times) * 1e3,
                      "bytes_per_row": bpr, "gb_per_s": rows * bpr / ms / 1e6, "frac_of_measured_hbm_peak": rows * bpr / ms / 1e6 / PEAK,
                      }), flush=True)
    del dn, dc, dg
    torch.cuda.empty_cache()
