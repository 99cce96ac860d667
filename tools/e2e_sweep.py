"""Sweep the host-fed path (DuckDB-protocol replay of our extension) over thread counts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from duckdb_imputation_b200 import replay

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
n = 20
rng = np.random.default_rng(0)
host = [rng.random(rows, dtype=np.float32) for _ in range(n)]
g = replay.glue()
for T in (1, 2, 4, 8, 16):
    for rep in range(3):
        t0 = time.perf_counter()
        r = g.aggregate(f"sum_to_triple_{n}_0", host, [], threads=T)
        dt = time.perf_counter() - t0
    print(f"T={T:2d}: {rows/dt/1e6:8.1f} M rows/s  ({rows*80/dt/1e9:5.1f} GB/s)  wall {dt*1e3:7.1f} ms  internal {g.last_seconds*1e3:7.1f} ms", flush=True)
