"""duckdb_imputation_b200 -- B200-native cofactor / triple aggregate behind the
eddbase/duckdb-imputation aggregate API (sum_to_triple_x_y, sum_to_nb_agg_x_y, sum_triple,
sum_nb_agg, to_cofactor).

The product is csrc/ (CUDA sm_100a kernels + the C ABI of include/cofactor_b200.h + the C++
DuckDB-callback glue).  The Python modules are plumbing for tests and bench.py.
"""
from ._native import CFB_NB, CFB_TRIPLE, CofactorError, lib  # noqa: F401
from .context import CofactorContext  # noqa: F401
from .aggregates import sum_to_nb_agg, sum_to_triple  # noqa: F401
