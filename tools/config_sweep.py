"""Device-resident timing of the BASELINE.json configs other than the headline (C1, C3, C4, C5-shape).

    python tools/config_sweep.py [scale]      scale = fraction of the full row counts (default 0.1)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from duckdb_imputation_b200 import CFB_NB, CFB_TRIPLE, CofactorContext, synth
from duckdb_imputation_b200 import _native as nat

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
lib = nat.lib()
CONFIGS = [
    # name, kind, n, m, domain, groups, full rows
    ("C1 sum_to_triple_5_0", CFB_TRIPLE, 5, 0, 0, 1, 1_000_000),
    ("C3 sum_to_triple_10_10 dom100", CFB_TRIPLE, 10, 10, 100, 1, 500_000_000),
    ("C4a sum_to_nb_agg_12_4 GROUP BY 10", CFB_NB, 12, 4, 100, 10, 500_000_000),
    ("C4b sum_to_triple_12_0 GROUP BY 10", CFB_TRIPLE, 12, 0, 0, 10, 500_000_000),
    ("C4c sum_to_nb_agg_12_4 ungrouped", CFB_NB, 12, 4, 100, 1, 500_000_000),
    ("C5 sum_to_triple_20_10 dom100", CFB_TRIPLE, 20, 10, 100, 1, 100_000_000),
]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for name, kind, n, m, dom, G, full in CONFIGS:
    if only and not any(name.startswith(o) for o in only):
        continue
    rows = max(1000, int(full * (1.0 if full <= 1_000_000 else scale)))
    dn = [torch.empty(rows, dtype=torch.float32, device="cuda") for _ in range(n)]
    dc = [torch.empty(rows, dtype=torch.int32, device="cuda") for _ in range(m)]
    for k, t in enumerate(dn):
        nat.check(lib.cfb_gen_uniform_f32(0, t.data_ptr(), rows, synth.column_seed(3, k), 0, None))
    for k, t in enumerate(dc):
        nat.check(lib.cfb_gen_int32(0, t.data_ptr(), rows, synth.column_seed(3, 100 + k), 0, 0, dom, None))
    dg = None
    if G > 1:
        dg = torch.empty(rows, dtype=torch.int32, device="cuda")
        nat.check(lib.cfb_gen_int32(0, dg.data_ptr(), rows, 777, 0, 0, G, None))
    torch.cuda.synchronize()
    best = 1e30
    for rep in range(4):
        with CofactorContext(kind, n, m, G) as ctx:
            if m:
                ctx.set_cat_domain([0] * m, [dom - 1] * m)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.scan_device(dn, dc, rows, d_group=dg, stream=stream.cuda_stream)
            e1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
    bpr = 4 * (n + m + (1 if G > 1 else 0))
    print(f"{name:40s} rows={rows:>11,d}  {best:9.3f} ms  {rows/best/1e6:9.2f} G rows/s  {rows*bpr/best/1e6:8.1f} GB/s "
          f"({100*rows*bpr/best/1e6/6551:5.1f}% of 6551)", flush=True)
    del dn, dc, dg
    torch.cuda.empty_cache()
