"""CPU suite: the numpy restatements of the trainers (oracle.build_sigma / linreg_train / lda_train -- the checkers of
SURVEY 8 f4) against the REFERENCE's own ML::ridge_linear_regression and lda_train compiled into oracle/_ref
(ML/regression.cpp, ML/lda.cpp, ML/utils.cpp; dgelsd / dgemm forwarded to the OpenBLAS inside scipy's wheel).  Also
writes / checks tests/golden/train_params.json, the fixtures the GPU tests read when /root/reference is absent."""
import json
import os

import numpy as np
import pytest

from duckdb_imputation_b200.struct_result import arrays_to_struct
from oracle import oracle, ref_replay
from tests.train_cases import CASES, restated, same_model, table, triple

needs_ref = pytest.mark.skipif(not ref_replay.available(), reason="oracle/_ref not built")
needs_lapack = pytest.mark.skipif(not (ref_replay.available() and ref_replay.lapack_available()), reason="no LAPACK for the reference's lda_train")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "train_params.json")


@needs_lapack
@pytest.mark.parametrize("case", sorted(CASES))
def test_restatement_matches_the_reference_trainer(case):
    targs, fn, consts = CASES[case]
    t, mine = restated(case)
    same_model(case, mine, ref_replay.train(fn, t, *consts))


@needs_ref
def test_sigma_matrix_matches_what_the_reference_trains_on():
    """build_sigma is pinned through the trainer: with max_iterations = 1 the reference takes exactly one gradient step
    from theta = -e_label, theta_1 = -step * (Sigma theta_0 / N) -- every column of Sigma shows up in the parameters."""
    x, c = table(seed=11, doms=(3, 2, 4))
    t = triple(x, c)
    sig, cat_array, idxs = oracle.build_sigma(t)
    assert sig.shape[0] == 1 + 3 + len(cat_array) and len(cat_array) >= 7 and np.array_equal(sig, sig.T)
    for label in range(3):
        step = 2.0 ** -12
        params = ref_replay.train("linreg_train", t, label, step, 0.0, 1, False, False)
        theta = np.zeros(sig.shape[0])
        theta[label + 1] = -1
        expect = -np.float32(step) * (sig @ theta) / sig[0, 0]
        got = params[1 + len(idxs) + len(cat_array):]
        np.testing.assert_allclose(got, np.delete(expect, label + 1), rtol=1e-6, atol=1e-9)


@needs_lapack
def test_the_reference_misplaces_class_sums_behind_the_label():
    """lda.cpp:131 adds the un-shifted cat_array index: with a categorical feature column BEHIND the label the class
    sums land `label keys` slots too far right (and past the end of the array for the last class -- the reference
    corrupts its heap there, so only the restatement of that behaviour is exercised, never the reference itself)."""
    t = triple(*table(seed=0))
    quirk = oracle.lda_train(t, 0, 0.001, False, unshifted_sums=True)
    right = oracle.lda_train(t, 0, 0.001, False)
    assert len(quirk) == len(right) and np.abs(quirk - right).max() > 1.0


def test_golden_fixture_is_current():
    """tests/golden/train_params.json = the restatement's parameter lists for CASES (regenerate: CFB_WRITE_GOLDEN=1)."""
    out = {case: [float(v) for v in restated(case)[1]] for case in sorted(CASES)}
    if os.environ.get("CFB_WRITE_GOLDEN"):
        with open(GOLDEN, "w") as f:
            json.dump(out, f)
    with open(GOLDEN) as f:
        want = json.load(f)
    assert sorted(want) == sorted(out)
    for case in out:
        np.testing.assert_allclose(out[case], want[case], rtol=1e-5, atol=1e-6)
