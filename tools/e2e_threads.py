"""Where does the host-fed path spend its time?  sum_to_triple_20_0 through the callbacks with T = 1, 2, 4, ... worker
threads, next to the staging copy alone (cfb_host_copy_ceiling) with the same T -- per-thread and aggregate GB/s."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from duckdb_imputation_b200 import replay
from duckdb_imputation_b200 import _native as nat

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
lib = nat.lib()
cores = os.cpu_count() or 1
rng = np.random.default_rng(0)
host = [rng.random(rows, dtype=np.float32) for _ in range(20)]
g = replay.glue()
print(json.dumps({"cores": cores, "stage_isa": lib.cfb_stage_isa().decode(), "rows": rows}), flush=True)
T = 1
while True:
    T = min(T, cores)
    for _ in range(2):
        g.aggregate("sum_to_triple_20_0", host, [], threads=T)
    best = 1e30
    for _ in range(3):
        g.aggregate("sum_to_triple_20_0", host, [], threads=T)
        best = min(best, g.last_seconds)
    with g.options(no_simple=1):
        g.aggregate("sum_to_triple_20_0", host, [], threads=T)
        g.aggregate("sum_to_triple_20_0", host, [], threads=T)
        hashed = g.last_seconds
    nt = lib.cfb_host_copy_ceiling(128 << 20, T, 1, 2)
    mc = lib.cfb_host_copy_ceiling(128 << 20, T, 0, 2)
    print(json.dumps({"threads": T, "e2e_M_rows_s": rows / best / 1e6, "e2e_gb_s": rows * 80 / best / 1e9,
                      "e2e_gb_s_per_thread": rows * 80 / best / 1e9 / T, "update_protocol_M_rows_s": rows / hashed / 1e6,
                      "copy_nt_gb_s": nt, "copy_memcpy_gb_s": mc}), flush=True)
    if T >= cores:
        break
    T *= 2
