"""GPU suite: the device-resident MICE loop of tools/mice_loop.py (filtered cofactor scans + in-place predict
write-back through the C ABI) against the same loop on the host (oracle cofactors + numpy predictions)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import mice_loop  # noqa: E402
from tests import mice_host  # noqa: E402

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_device_loop_matches_host_loop():
    rows = 40_000
    num, cat, mn, mc, truth = mice_loop.synthetic_table(rows, n=6, m=4, dom=6, null_num=(0, 2), null_cat=(1,), seed=9)
    d_num = [torch.from_numpy(c.copy()).cuda() for c in num]
    d_cat = [torch.from_numpy(c.copy()).cuda() for c in cat]
    d_nn = {c: torch.from_numpy(m.astype(np.int32)).cuda() for c, m in mn.items()}
    d_nc = {c: torch.from_numpy(m.astype(np.int32)).cuda() for c, m in mc.items()}
    mice_loop.mice_gpu(d_num, d_cat, d_nn, d_nc, 2, rows)
    h_num, h_cat = mice_host.mice_cpu([c.copy() for c in num], [c.copy() for c in cat], mn, mc, 2)
    for c, msk in mn.items():
        got = d_num[c].cpu().numpy()
        assert np.array_equal(got[~msk], num[c][~msk])  # observed cells untouched
        assert np.abs(got[msk] - h_num[c][msk]).max() < 2e-3 * max(1.0, np.abs(h_num[c][msk]).max())
    for c, msk in mc.items():
        got = d_cat[c].cpu().numpy()
        assert np.array_equal(got[~msk], cat[c][~msk])
        assert (got[msk] == h_cat[c][msk]).mean() > 0.995  # near-ties between two classes may fall either way
    # and the imputations are informative
    assert np.abs(d_num[0].cpu().numpy()[mn[0]] - truth[("n", 0)][mn[0]]).mean() < 0.6 * np.abs(num[0][mn[0]] - truth[("n", 0)][mn[0]]).mean()
