import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext
rng = np.random.default_rng(0)
rows = 200_000
num = [rng.random(rows, dtype=np.float32) for _ in range(10)]
cat = [rng.integers(0, 100, rows).astype(np.int32) for _ in range(10)]
for rep in range(3):
    t0 = time.perf_counter()
    ctxs = [CofactorContext(CFB_TRIPLE, 10, 10) for _ in range(16)]
    t1 = time.perf_counter()
    for c in ctxs:
        c.append(num, cat)
    t2 = time.perf_counter()
    for c in ctxs:
        c.sync()
    t3 = time.perf_counter()
    for c in ctxs[1:]:
        ctxs[0].combine(c)
    ctxs[0].sync()
    t4 = time.perf_counter()
    r = ctxs[0].finalize_arrays()
    t5 = time.perf_counter()
    for c in ctxs:
        c.close()
    t6 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.1f}  append {1e3*(t2-t1):.1f}  sync {1e3*(t3-t2):.1f}  combine {1e3*(t4-t3):.1f}  finalize {1e3*(t5-t4):.1f}  close {1e3*(t6-t5):.1f} ms; N={r['N']}")
