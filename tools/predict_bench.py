"""Device-resident throughput of the MICE write-back kernel (predict_kernel) on the MICE table shape
(20 FLOAT + 10 INTEGER columns of domain 100): linear regression (1 output) and LDA (10 classes), with a
20 % NULL mask and the output aliasing a column.  One JSON line per case, for profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from duckdb_imputation_b200 import predict, synth
from duckdb_imputation_b200 import _native as nat

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
rows -= rows % 4
N, M, DOM = 20, 10, 100
lib = nat.lib()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6551.0
dn = [torch.empty(rows, dtype=torch.float32, device="cuda") for _ in range(N)]
dc = [torch.empty(rows, dtype=torch.int32, device="cuda") for _ in range(M)]
for k, t in enumerate(dn):
    nat.check(lib.cfb_gen_uniform_f32(0, t.data_ptr(), rows, synth.column_seed(5, k), 0, None))
for k, t in enumerate(dc):
    nat.check(lib.cfb_gen_int32(0, t.data_ptr(), rows, synth.column_seed(5, 100 + k), 0, 0, DOM, None))
mask = torch.empty(rows, dtype=torch.int32, device="cuda")
nat.check(lib.cfb_gen_int32(0, mask.data_ptr(), rows, 4242, 0, 0, 5, None))
mask = (mask == 0).to(torch.int32)  # 20 % of the cells are NULL
rng = np.random.default_rng(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for name, K, mode, dt in (("linreg_predict", 1, predict.SCORE, torch.float32), ("lda_predict (10 classes)", 10, predict.ARGMAX, torch.int32)):
    model = predict.LinearModel(rng.standard_normal(K), rng.standard_normal((K, N)), [np.arange(DOM)] * M, rng.standard_normal((K, M * DOM)))
    for masked in (False, True):
        out = torch.zeros(rows, dtype=dt, device="cuda")
        times = []
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            predict.predict_device(model, dn, dc, rows, mode, out, d_mask=mask if masked else None, stream=stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            if rep >= 2:
                times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        frac_rows = 0.2 if masked else 1.0
        # algorithmic bytes: every input value of a scored row once + the 4-byte result (+ the mask of every row)
        bpr = (4 * (N + M) + 4) * frac_rows + (4 if masked else 0)
        print(json.dumps({"kernel": "predict_kernel", "case": name + (" over the 20 % NULL cells" if masked else " over all rows"),
                          "rows": rows, "ms": ms, "rows_per_s": rows / ms * 1e3, "algorithmic_bytes_per_row": bpr,
                          "gb_per_s": rows * bpr / ms / 1e6, "frac_of_measured_hbm_peak": rows * bpr / ms / 1e6 / PEAK}), flush=True)
