"""CPU suite: the scalar lifts to_cofactor / to_nb_agg of OUR extension (host functions, no GPU
involved) against every per-row golden of the reference's test_lift.py:22-63 and test_nb_lift.py."""
from duckdb_imputation_b200 import replay
from tests import sqlmini


def test_lift_goldens(goldens):
    g = replay.glue()
    assert {"to_cofactor", "to_nb_agg"} <= set(_scalars(g))
    cases = [c for c in goldens["cases"] if c["file"] in ("test_lift.py", "test_nb_lift.py")]
    assert len(cases) == 40
    for c in cases:
        got = sqlmini.run_lift(c["sql"], goldens["fixtures"][c["file"]], g.scalar)
        assert got[c["index"]] == c["expected"], (c["file"], c["test"], c["index"])


def _scalars(g):
    # the catalog listing only names aggregates; probe the two lifts directly
    import numpy as np
    out = []
    for name in ("to_cofactor", "to_nb_agg"):
        g.scalar(name, [np.ones(1, np.float32)], [])
        out.append(name)
    return out


def test_lift_output_over_many_chunks():
    import numpy as np
    rng = np.random.default_rng(4)
    rows = 5000  # > 2048: several chunks
    a = rng.integers(0, 9, rows).astype(np.float32)
    d = rng.integers(0, 5, rows).astype(np.int32)
    out = replay.glue().scalar("to_cofactor", [a], [d])
    assert len(out) == rows
    r = 4321
    assert out[r] == {"N": 1, "lin_num": [float(a[r])], "quad_num": [float(a[r] * a[r])],
                      "lin_cat": [[{"key": int(d[r]), "value": 1.0}]],
                      "quad_num_cat": [[{"key": int(d[r]), "value": float(a[r])}]],
                      "quad_cat": [[{"key1": int(d[r]), "key2": int(d[r]), "value": 1.0}]]}
