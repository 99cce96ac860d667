import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from duckdb_imputation_b200 import replay
g = replay.glue()
T = os.cpu_count()
rng = np.random.default_rng(0)
for rows in (8_000_000, 32_000_000, 64_000_000):
    num = [rng.random(rows, dtype=np.float32) for _ in range(10)]
    cat = [rng.integers(0, 100, rows).astype(np.int32) for _ in range(10)]
    for _ in range(2):
        g.aggregate("sum_to_triple_10_10", num, cat, threads=T)
    os.environ["CFB_REPLAY_TRACE"] = "1"
    g.aggregate("sum_to_triple_10_10", num, cat, threads=T); dt = g.last_seconds
    os.environ.pop("CFB_REPLAY_TRACE")
    print(f"rows={rows:,} T={T}: {dt*1e3:.1f} ms -> {rows/dt/1e6:.1f} M rows/s", flush=True)
    for th in (4, 8):
        g.aggregate("sum_to_triple_10_10", num, cat, threads=th); dt = g.last_seconds
        print(f"   T={th}: {dt*1e3:.1f} ms -> {rows/dt/1e6:.1f} M rows/s", flush=True)
