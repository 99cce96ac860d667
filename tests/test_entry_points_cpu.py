"""CPU suite: the loadable-extension entry points the reference exports (duckdb_imputation_extension.cpp:269-279)
are exported by our glue library and register the whole catalog (no device needed: registration is host code)."""
import ctypes as C

from duckdb_imputation_b200 import replay


def test_duckdb_imputation_init_and_version_are_exported_and_register_the_catalog():
    g = replay.glue()
    for sym in ("duckdb_imputation_init", "duckdb_imputation_version"):
        assert hasattr(g.lib, sym), sym
    n, version = g.load_via_entry_points()
    # 2 x (21 x 21 - 1) grid aggregates + sum_triple + sum_nb_agg, and the 10 scalar functions (to_cofactor, to_nb_agg,
    # multiply_triple, multiply_nb_agg, 4 x *_predict, linreg_train, lda_train)
    assert n == 2 * (21 * 21 - 1) + 2 + 10
    assert version.startswith("v0.9.2")  # the DuckDB build the reference pins (README.md:35-42)
    g.lib.duckdb_imputation_version.restype = C.c_char_p
    assert g.lib.duckdb_imputation_version().decode() == version


def test_simple_update_is_registered_for_the_ring_aggregates():
    # ungrouped queries may be planned as PhysicalUngroupedAggregate (allowed improvement)
    assert {"sum_to_triple_20_0", "sum_to_nb_agg_12_4", "sum_triple", "sum_nb_agg"} <= set(replay.glue().functions())
