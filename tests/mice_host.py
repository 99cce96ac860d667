"""The MICE loop of tools/mice_loop.py on the host -- oracle cofactors + numpy predictions: TEST INFRASTRUCTURE,
the checker of the device loop (tests/test_gpu_mice.py) and of the closed-form trainers (tests/test_mice_cpu.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from mice_loop import predict_np, train_lda, train_linreg  # noqa: E402


def mice_cpu(num, cat, null_num, null_cat, iters):
    """Host loop (oracle cofactors + numpy predictions).  num / cat: lists of numpy columns (modified in place);
    null_num / null_cat: {column index: boolean NULL mask}."""
    from oracle import oracle
    for _ in range(iters):
        for c, mask in null_cat.items():
            res = oracle.aggregate_arrays(oracle.TRIPLE, num, cat, sel=np.nonzero(~mask)[0].astype(np.uint32))[0]
            model = train_lda(res, c)
            s = predict_np(model, [x[mask] for x in num], [x[mask] for k, x in enumerate(cat) if k != c])
            cat[c][mask] = np.argmax(s, axis=1).astype(np.int32)      # the class INDEX, as LDA_impute (lda.cpp:575)
        for c, mask in null_num.items():
            res = oracle.aggregate_arrays(oracle.TRIPLE, num, cat, sel=np.nonzero(~mask)[0].astype(np.uint32))[0]
            model = train_linreg(res, c)
            s = predict_np(model, [x[mask] for k, x in enumerate(num) if k != c], [x[mask] for x in cat])
            num[c][mask] = s[:, 0].astype(np.float32)
    return num, cat
