/*
 * cofactor_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the cofactor / triple ("ring") sum aggregates of
 * eddbase/duckdb-imputation.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (libcofactor_b200.so) never links, loads or calls it.
 *
 * Pinning: the restatement is checked against every golden STRUCT of the
 * reference's own tests (test_sum.py:22-52, test_nb_sum.py, test_lift.py:22-63;
 * see tests/test_oracle_golden.py) and against the reference's own sources
 * compiled here over a DuckDB-vector shim (oracle/_ref, see oracle/Makefile).
 */
#ifndef COFACTOR_ORACLE_H
#define COFACTOR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field layout as cfb_result (include/cofactor_b200.h) so one ctypes
 * Structure reads both. */
typedef struct orc_result {
  int32_t kind, n_num, n_cat;
  int64_t N;
  int64_t n_quad;
  double *lin;
  double *quad;
  int64_t total_keys;
  int64_t *cat_offsets;
  int32_t *cat_keys;
  int64_t *cat_counts;
  double *numcat_sums;
  int64_t n_pair_lists;
  int64_t *pair_offsets;
  int32_t *pair_key1;
  int32_t *pair_key2;
  int64_t *pair_counts;
} orc_result;

enum { ORC_TRIPLE = 0, ORC_NB = 1 };
enum {
  ORC_EXACT = 0,   /* int64 counts, fp64 sums: the parity target at scale          */
  ORC_FAITHFUL = 1 /* fp32 sums AND fp32 counts, row order, 2048-row chunks,
                      per-row state pointers, T threads + combine: reproduces the
                      reference's arithmetic; doubles as the CPU baseline "port"   */
};

/* Aggregate `rows` rows of columnar input into n_groups results (out[g]).
 * group == NULL -> everything goes to group 0.  sel == NULL -> identity; else
 * the aggregate runs over rows sel[0..rows) (ascending) of the columns, a filtered scan of a
 * table of table_rows physical rows (only used to cut ORC_FAITHFUL chunks and morsels).
 * Follows sum_no_lift.cpp:83-214 (ORC_TRIPLE) / sum_to_nb_agg.cpp:61-145
 * (ORC_NB), sum_state.cpp:23-112 (combine) and :132-461 (output order).
 * Returns 0, or -1 on bad arguments. */
int orc_aggregate(int kind, int mode, int n_num, int n_cat, const float *const *num_cols,
                  const int32_t *const *cat_cols, const int32_t *group, int n_groups,
                  const uint32_t *sel, size_t rows, size_t table_rows, int threads, orc_result *out);

/* sum_triple(to_cofactor(..)) / sum_nb_agg(to_nb_agg(..)): lift every row to a
 * singleton triple (lift.cpp:85-241, lift_to_nb_agg.cpp:13-136) and add the
 * lifted triples with the Sum update (sum.cpp:86-260, sum_nb_agg.cpp:69-174).
 * Arithmetic is fp32 like the reference.  */
int orc_sum_of_lifted(int kind, int n_num, int n_cat, const float *const *num_cols,
                      const int32_t *const *cat_cols, const int32_t *group, int n_groups,
                      size_t rows, orc_result *out);

/* out = a + b on two results (Value-level sum_triple, sum.cpp:319-460). */
int orc_result_add(const orc_result *a, const orc_result *b, orc_result *out);

void orc_result_free(orc_result *r);

/* Wall-clock seconds of the most recent orc_aggregate call on this thread
 * (update + combine + finalize only; inputs already in host RAM). */
double orc_last_seconds(void);

#ifdef __cplusplus
}
#endif
#endif
