"""Host-fed (DuckDB callbacks) throughput of the categorical / GROUP BY configs, next to the
reference's own callbacks on the same columns (oracle/_ref) -- both through the replay host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from duckdb_imputation_b200 import replay
from oracle import ref_replay

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
T = os.cpu_count() or 1
rng = np.random.default_rng(0)
CONFIGS = [("C3 sum_to_triple_10_10", "sum_to_triple_10_10", 10, 10, None),
           ("C4a sum_to_nb_agg_12_4 GROUP BY 10", "sum_to_nb_agg_12_4", 12, 4, 10),
           ("C4b sum_to_triple_12_0 GROUP BY 10", "sum_to_triple_12_0", 12, 0, 10),
           ("C2 sum_to_triple_19_0 (reference grid)", "sum_to_triple_19_0", 19, 0, None),
           ("C5 sum_to_triple_19_10 (MICE scan shape, reference grid)", "sum_to_triple_19_10", 19, 10, None)]
g, r = replay.glue(), (ref_replay.ref() if ref_replay.available() else None)
for name, fn, n, m, G in CONFIGS:
    num = [rng.random(rows, dtype=np.float32) for _ in range(n)]
    cat = [rng.integers(0, 100, rows).astype(np.int32) for _ in range(m)]
    grp = None if G is None else rng.integers(0, G, rows).astype(np.int32)
    kw = dict(group=grp, n_groups=G or 1, threads=T)
    for _ in range(2):
        g.aggregate(fn, num, cat, **kw)
    g.aggregate(fn, num, cat, **kw); dt = g.last_seconds  # update + combine + finalize (not the JSON rendering of the result)
    line = f"{name:42s} rows={rows:>10,d} T={T}: b200 {rows/dt/1e6:8.1f} M rows/s"
    if r is not None:
        rr = min(rows, 2_000_000)
        kw2 = dict(group=None if grp is None else grp[:rr], n_groups=G or 1, threads=T)
        r.aggregate(fn, [c[:rr] for c in num], [c[:rr] for c in cat], **kw2)
        line += f" | reference CPU {rr/r.last_seconds/1e6:7.2f} M rows/s  -> x{(rows/dt)/(rr/r.last_seconds):6.1f}"
    print(line, flush=True)
