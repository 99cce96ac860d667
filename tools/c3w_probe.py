"""C3w (sum_to_triple_10_2, domain 100 000: pair counts in the hashed table) device-resident, fresh context per
repetition, WITHOUT the finalize (50 M distinct pairs take seconds to extract on the host): scan time with the table
reserved once per call (default) and grown slice by slice (CFB_HASH_RESERVE_MB=0).

    python tools/c3w_probe.py [rows] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from duckdb_imputation_b200 import CFB_TRIPLE, CofactorContext, synth
from duckdb_imputation_b200 import _native as nat

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
lib = nat.lib()
n, m, dom = 10, 2, 100_000
dn = [torch.empty(rows, dtype=torch.float32, device="cuda") for _ in range(n)]
dc = [torch.empty(rows, dtype=torch.int32, device="cuda") for _ in range(m)]
for k, t in enumerate(dn):
    nat.check(lib.cfb_gen_uniform_f32(0, t.data_ptr(), rows, synth.column_seed(3, k), 0, None))
for k, t in enumerate(dc):
    nat.check(lib.cfb_gen_int32(0, t.data_ptr(), rows, synth.column_seed(3, 100 + k), 0, 0, dom, None))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
torch.cuda.synchronize()
for mode, mb in (("reserved once per call", None), ("grown slice by slice", "0"), ("reserved once per call", None)):
    if mb is None:
        os.environ.pop("CFB_HASH_RESERVE_MB", None)
    else:
        os.environ["CFB_HASH_RESERVE_MB"] = mb
    times = []
    for rep in range(reps):
        with CofactorContext(CFB_TRIPLE, n, m, 1) as ctx:
            ctx.set_cat_domain([0] * m, [dom - 1] * m)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.scan_device(dn, dc, rows, stream=stream.cuda_stream)
            e1.record(stream)
            ctx.sync()
            torch.cuda.synchronize()
            times.append(round(e0.elapsed_time(e1), 2))
    print(json.dumps({"config": "C3w", "rows": rows, "table": mode, "ms": times, "best_rows_per_s": rows / min(times) * 1e3,
                      "median_rows_per_s": rows / sorted(times)[len(times) // 2] * 1e3}), flush=True)
