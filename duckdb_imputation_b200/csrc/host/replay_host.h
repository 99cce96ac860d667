/* replay_host.h -- C API of the hash-aggregate replay ("mini DuckDB operator").
 *
 * DuckDB itself is not available here, so this is the caller side of the drop-in boundary:
 * it drives a registered duckdb::AggregateFunction with the exact protocol DuckDB 0.9.2's
 * PhysicalHashAggregate uses for the ring aggregates: bind once; T worker
 * threads, each with thread-local states (state_size bytes, initialize on first use), scanning
 * contiguous morsels in chunks of <= 2048 rows and calling
 *   update(inputs[], aggr_input_data, input_count, state_vector /-* one pointer per row *-/, count)
 * then combine(source, target) of the thread-local states, destructor on the sources,
 * finalize(states, result, count, 0) and destructor on the targets.
 * The same file is linked (a) with our extension (ring_extension.cpp + triple_glue.cpp) and
 * (b) with the reference's own sources (oracle/ref_extension.cpp) -- one driver, two
 * implementations behind the same callback API.
 */
#ifndef CFB_REPLAY_HOST_H
#define CFB_REPLAY_HOST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Run  SELECT <function>(num..., cat...) FROM t [WHERE row in sel] [GROUP BY group]
 *  or   SELECT <function>(<scalar>(num..., cat...)) FROM t ...   when `scalar` is not NULL/empty
 *       (e.g. sum_triple(to_cofactor(...)): the scalar function is evaluated chunk by chunk and its
 *       result vector is the aggregate's single input)
 *   function   registered name, e.g. "sum_to_triple_3_3"
 *   group      per-row group slot in [0, n_groups) or NULL; results come back in slot order,
 *              slots without rows are omitted
 *   sel        ascending row ids that pass the filter, or NULL (n_sel ignored); filtered chunks
 *              reach update() as DICTIONARY vectors (selection on top of the scan vectors)
 *   json_out   malloc'd JSON array with one STRUCT per group (free with replay_free)
 *   seconds    wall time of update + combine + finalize
 * Returns 0, or -1 with the message in replay_last_error().                             */
int replay_aggregate(const char *function, const char *scalar, int n_num, int n_cat, const float *const *num,
                     const int32_t *const *cat, const int32_t *group, int n_groups, const uint32_t *sel,
                     size_t n_sel, size_t rows, int threads, char **json_out, double *seconds);
/* Run  SELECT <scalar>(num..., cat...) FROM t [WHERE row in sel]  -> JSON array, one value per row. */
int replay_scalar(const char *scalar, int n_num, int n_cat, const float *const *num, const int32_t *const *cat,
                  const uint32_t *sel, size_t n_sel, size_t rows, char **json_out);
/* Run  SELECT <scalar>(A, B, ...) FROM j  where every argument is a column of ring STRUCTs (the output of
 * a join over aggregate results): json_args[k] is a JSON array with one STRUCT per row, in the rendering
 * of json_out (field names are ignored; nb != 0: the four-field Naive-Bayes ring).  e.g. multiply_triple. */
int replay_scalar_structs(const char *scalar, int nb, int n_args, const char *const *json_args, size_t rows, char **json_out);
/* Run  SELECT <scalar>(params::FLOAT[], flag..., num..., cat...) FROM t [WHERE row in sel]  -- the predict functions
 * (linreg_predict: flags = noise, normalize; lda_predict: flags = normalize).  `out` receives one 4-byte value per
 * (selected) row: float or int32, the function's return type.                                            */
int replay_predict(const char *scalar, const float *params, size_t n_params, const int *flags, int n_flags, int n_num,
                   int n_cat, const float *const *num, const int32_t *const *cat, const uint32_t *sel, size_t n_sel,
                   size_t rows, void *out);
/* Run  SELECT <scalar>(<ring STRUCT>, c1, c2, ...)  -- the trainers (linreg_train: label INTEGER, step FLOAT, lambda
 * FLOAT, max_iterations INTEGER, variance BOOLEAN, normalize BOOLEAN; lda_train: label INTEGER, shrinkage FLOAT,
 * normalize BOOLEAN).  const_types[i] in "ifb" says how consts[i] is handed over.  *params_out: the FLOAT[] the
 * function returns (malloc'd, free with replay_free).                                                       */
int replay_train(const char *scalar, const char *json_triple, int n_consts, const double *consts, const char *const_types,
                 float **params_out, size_t *n_params_out);
/* Run  SELECT <scalar>(list(agg), list(label), c1, ...)  -- the per-class trainers (qda_train: normalize BOOLEAN;
 * nb_train: no constants) over a LIST of ring STRUCTs (nb != 0: the four-field ring) and an INTEGER[] of labels.  */
int replay_train_list(const char *scalar, int nb, const char *json_triples, const int32_t *labels, size_t n_labels,
                      int n_consts, const double *consts, const char *const_types, float **params_out, size_t *n_params_out);
/* Shapes of DuckDB's protocol that a plain table scan does not produce (process-wide; "reset" restores all):
 *   no_simple = 1          never plan an ungrouped aggregate: use update() with per-row state pointers
 *   lift_shape = 1|2|3     the lifted STRUCT reaches the aggregate as a DICTIONARY vector (1), as a flat vector with
 *                          DICTIONARY children and leaves (2), or as a CONSTANT vector -- the chunk's first row
 *                          repeated (3): what joins and filters below sum_triple / multiply_triple hand over
 *   split_states = k       every worker keeps k state sets per group and merges them itself (hash table reset)
 *   parallel_finalize = P  P threads combine + finalize disjoint groups at the same time (radix partitions)  */
int replay_set_option(const char *name, int value);
/* Load a fresh catalog through duckdb_imputation_init(); returns the number of registered functions, -1 if the build
 * does not export the entry points.  *version_out = duckdb_imputation_version().                               */
int replay_load_via_entry_points(const char **version_out);
void replay_free(char *p);
const char *replay_last_error(void);
/* Names of all registered aggregate functions, '\n'-separated (malloc'd). */
char *replay_list_functions(void);
/* Which implementation is behind this library: "b200" or "reference". */
const char *replay_implementation(void);

#ifdef __cplusplus
}
#endif
#endif
