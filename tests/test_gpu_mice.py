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


@pytest.mark.parametrize("per_pattern,closed_form", [(False, False), (True, False), (True, True)])
def test_delta_cofactor_loop_matches_the_filtered_scan_loop(per_pattern, closed_form):
    """with the table partitioned by NULL pattern, the cofactor over the rows where a column is
    observed is total - nulls (cfb_result_combine, the arithmetic of subtract_triple, imputation/triple/sub.cpp:71-219)
    and only the NULL rows (20 %) are scanned per step.  Same models, same imputations as the loop that rescans the
    whole table behind a row filter.  per_pattern: one kept cofactor per NULL pattern, one scan per column; closed_form: the cofactor after a linear-regression
    write-back from cfb_result_impute_linear, no scan."""
    rows = 60_001
    num, cat, mn, mc, _ = mice_loop.synthetic_table(rows, n=5, m=3, dom=5, null_num=(0, 3), null_cat=(2,), seed=21)

    def dev():
        return ([torch.from_numpy(c.copy()).cuda() for c in num], [torch.from_numpy(c.copy()).cuda() for c in cat],
                {c: torch.from_numpy(m.astype(np.int32)).cuda() for c, m in mn.items()},
                {c: torch.from_numpy(m.astype(np.int32)).cuda() for c, m in mc.items()})

    a_num, a_cat, a_nn, a_nc = dev()
    mice_loop.mice_gpu(a_num, a_cat, a_nn, a_nc, 2, rows)
    b_num, b_cat, b_nn, b_nc = dev()
    _, order, _ = mice_loop.mice_gpu_delta(b_num, b_cat, b_nn, b_nc, 2, rows, per_pattern=per_pattern, closed_form=closed_form)
    order = order.cpu().numpy()
    for c in mn:
        want, got = a_num[c].cpu().numpy()[order], b_num[c].cpu().numpy()
        assert np.abs(got - want).max() < 1e-3 * max(1.0, np.abs(want).max())
    for c in mc:
        want, got = a_cat[c].cpu().numpy()[order], b_cat[c].cpu().numpy()
        assert (got == want).mean() > 0.999
