/*
 * cofactor_b200.h -- C ABI of the B200-native cofactor / triple aggregate.
 *
 * This is the drop-in boundary for the ONE hot path of eddbase/duckdb-imputation:
 * the ring ("triple") sum aggregates sum_to_triple_x_y, sum_to_nb_agg_x_y,
 * sum_triple, sum_nb_agg (and the to_cofactor lift that feeds sum_triple).
 * The reference has no FFI of its own -- the only callers are the DuckDB
 * aggregate callbacks -- so every entry point below cites the reference
 * callback (file:line under /root/reference/duckdb_extension/src) whose work it
 * replaces.  The DuckDB-side glue that binds them lives in
 * duckdb_imputation_b200/csrc/host/{triple,lift,mul,predict}_glue.cpp and is described in INTEGRATION.md.
 *
 * Conventions
 *   - plain C, POD arguments, no torch / DuckDB / CUDA types in any signature
 *     (streams and device pointers travel as void*);
 *   - every function returns CFB_OK (0) or a negative cfb_status; the message
 *     for the calling thread's last failure is cfb_last_error();
 *   - there is NO CPU fallback: without a usable CUDA device every compute
 *     entry point fails with CFB_ERR_NO_DEVICE;
 *   - a cfb_ctx is the aggregate state ("SumState" shrinks to {cfb_ctx*}); it is
 *     re-entrant per context: distinct threads may drive distinct contexts
 *     concurrently, one thread at a time per context.
 */
#ifndef COFACTOR_B200_H
#define COFACTOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFB_ABI_VERSION 2

typedef enum cfb_status {
  CFB_OK = 0,
  CFB_ERR_INVALID = -1,   /* bad argument / shape mismatch                    */
  CFB_ERR_NO_DEVICE = -2, /* no CUDA device or driver; never falls back to CPU */
  CFB_ERR_CUDA = -3,      /* a CUDA runtime call or kernel failed              */
  CFB_ERR_OOM = -4,       /* host or device allocation failed                  */
  CFB_ERR_DOMAIN = -5,    /* key outside a declared domain / state too large / no dense partial */
  CFB_ERR_STATE = -6      /* call sequence violation (e.g. append after free)  */
} cfb_status;

/* Which ring the context accumulates.
 * CFB_TRIPLE  : N, lin_agg, quad_agg (packed upper triangle), lin_cat,
 *               quad_num_cat, quad_cat      -- Triple::SumNoLift, sum_no_lift.cpp:53-216
 * CFB_NB      : N, lin_agg, quad_agg (diagonal only), lin_cat
 *                                            -- Triple::sum_to_nb_agg, sum_to_nb_agg.cpp:39-146 */
typedef enum cfb_kind { CFB_TRIPLE = 0, CFB_NB = 1 } cfb_kind;

#define CFB_MAX_NUM 32 /* FLOAT columns per aggregate (reference grid: 0..19, README: 20) */
#define CFB_MAX_CAT 32 /* INTEGER columns per aggregate                                   */

typedef struct cfb_ctx cfb_ctx;

/* ------------------------------------------------------------------ runtime */

/* ABI version of the loaded library (== CFB_ABI_VERSION of the header it was built from). */
int cfb_abi_version(void);

/* Number of visible CUDA devices (0 when there is no driver/GPU). Never fails. */
int cfb_device_count(void);

/* Message of the calling thread's most recent failing call ("" if none). */
const char *cfb_last_error(void);

/* -------------------------------------------------------- aggregate context */

/* Replaces StateFunction::Initialize (sum_state.h:33-45) plus the lazy shape
 * allocation inside update (sum_no_lift.cpp:96-116 / sum_to_nb_agg.cpp:75-95).
 * n_groups >= 1: the context holds n_groups independent states (GROUP BY slots);
 * rows are routed by a per-row slot id in cfb_ctx_append / cfb_triple_device.  */
int cfb_ctx_create(int device, int kind, int n_num, int n_cat, int n_groups, cfb_ctx **out);

/* Replaces StateFunction::Destroy (sum_state.h:48-52).  NULL is accepted. */
int cfb_ctx_destroy(cfb_ctx *ctx);

/* Optional: declare the value range [lo[k], hi[k]] of categorical column k
 * (e.g. from DuckDB column statistics) so the dense key->slot tables can be laid
 * out without a min/max pre-pass.  Keys outside a declared range are an error
 * (CFB_ERR_DOMAIN) reported by the next synchronising call.                    */
int cfb_ctx_set_cat_domain(cfb_ctx *ctx, const int32_t *lo, const int32_t *hi);

/* Replaces the per-chunk body of Triple::SumNoLift (sum_no_lift.cpp:83-214) and
 * Triple::sum_to_nb_agg (sum_to_nb_agg.cpp:61-145) for HOST (DuckDB) vectors.
 *   num_cols[i] / cat_cols[k] : base pointers of the unified-format column data
 *   num_sel[i]  / cat_sel[k]  : that column's selection vector (NULL = identity);
 *                               row r of the chunk reads data[sel ? sel[r] : r]
 *                               (UnifiedVectorFormat, sum_no_lift.cpp:120)
 *   group_slot                : per-row state slot in [0, n_groups) or NULL (slot 0)
 *                               -- the states[sdata.sel->get_index(j)] indirection
 *   count                     : rows in this chunk (any size; DuckDB sends <= 2048)
 * Rows are gathered into pinned columnar staging; a full staging tile is shipped
 * with cudaMemcpyAsync and reduced on the device while the next tile fills.
 * Validity masks are not consulted (the reference never reads them).           */
int cfb_ctx_append(cfb_ctx *ctx, const float *const *num_cols, const uint32_t *const *num_sel,
                   const int32_t *const *cat_cols, const uint32_t *const *cat_sel,
                   const uint32_t *group_slot, size_t count);

/* Same reduction for DEVICE-resident columnar (SoA) input: the device-resident
 * measurement path and the multi-GPU range partition.  Column pointers are
 * device pointers (16-byte aligned); d_group_slot is a device int32 array of
 * per-row slots or NULL; a NEGATIVE slot drops the row (a filtered scan such as
 * MICE's WHERE col_IS_NULL IS FALSE is n_groups = 1 with slots 0 / -1).  `stream` is a cudaStream_t (NULL = the context's own
 * stream).  Asynchronous: results are visible after cfb_ctx_sync/finalize.     */
int cfb_triple_device(cfb_ctx *ctx, const float *const *d_num_cols, const int32_t *const *d_cat_cols,
                      const int32_t *d_group_slot, size_t n_rows, void *stream);

/* One (offset,length) range into a flattened list child -- DuckDB's list_entry_t. */
typedef struct cfb_list_entry {
  uint64_t offset;
  uint64_t length;
} cfb_list_entry;

/* Replaces the per-chunk body of Triple::Sum (sum.cpp:57-261) and Triple::sum_nb_agg
 * (sum_nb_agg.cpp:45-175): add `count` ALREADY-LIFTED triples (rows of to_cofactor /
 * multiply_triple output) to the state.  Arguments are the flattened children of the input
 * STRUCT vector, exactly as DuckDB lays them out:
 *   N[count]                                  child 0
 *   lin[count*n], quad[count*nq]              list children 1, 2 (nq = n(n+1)/2, or n for CFB_NB)
 *   lin_cat lists   [count*m]      -> (lc_key, lc_val)      value = count of key
 *   num_cat lists   [count*n*m]    -> (nc_key, nc_val)      list (r*n + l)*m + c, value = sum x_l | key
 *   cat_cat lists   [count*m(m+1)/2] -> (cc_key1, cc_key2, cc_val)   (k<=l) order, diagonal included
 * The numeric children are summed on the device; the sparse (key,value) entries are
 * scatter-added into the dense device tables.  num_cat / cat_cat are NULL for CFB_NB.   */
int cfb_ctx_append_triples(cfb_ctx *ctx, size_t count, const int32_t *N, const float *lin, const float *quad,
                           const cfb_list_entry *lin_cat_lists, const int32_t *lc_key, const float *lc_val,
                           const cfb_list_entry *num_cat_lists, const int32_t *nc_key, const float *nc_val,
                           const cfb_list_entry *cat_cat_lists, const int32_t *cc_key1, const int32_t *cc_key2,
                           const float *cc_val);

/* Same, into GROUP BY slot `slot` of the context (sum_triple ... GROUP BY: the states of a worker thread share
 * one context, like the states of the sum_to_triple aggregates).                                       */
int cfb_ctx_append_triples_slot(cfb_ctx *ctx, int slot, size_t count, const int32_t *N, const float *lin,
                                const float *quad, const cfb_list_entry *lin_cat_lists, const int32_t *lc_key,
                                const float *lc_val, const cfb_list_entry *num_cat_lists, const int32_t *nc_key,
                                const float *nc_val, const cfb_list_entry *cat_cat_lists, const int32_t *cc_key1,
                                const int32_t *cc_key2, const float *cc_val);

/* Drain the context's streams; surfaces asynchronous kernel errors. */
int cfb_ctx_sync(cfb_ctx *ctx);

/* Replaces Triple::SumStateCombine (sum_state.cpp:10-114): dst += src, group by
 * group.  Shapes must match (an empty dst adopts src's categorical domains).
 * src stays valid and destroyable.  Works across devices (peer or staged copy). */
int cfb_ctx_combine(cfb_ctx *dst, const cfb_ctx *src);

/* Same for individual GROUP BY slots: dst[dst_slots[i]] += src[src_slots[i]], i < n_pairs; all other
 * slots of both contexts are untouched (n_groups may differ).  This is what the DuckDB glue calls when
 * the thread-local hash tables are merged: one context holds all groups of a worker thread.
 * dst == src is allowed (two states of one thread's context: a group that the radix-partitioned hash
 * aggregate emitted twice); then no slot may be source and target in the same call.            */
int cfb_ctx_combine_slots(cfb_ctx *dst, const cfb_ctx *src, size_t n_pairs, const int32_t *dst_slots,
                          const int32_t *src_slots);

/* ------------------------------------------------------------------- result */

/* Flat canonical result of ONE state (group).  Mirrors the STRUCT written by
 * Triple::SumStateFinalize (sum_state.cpp:116-464):
 *   N            <- state->count                      (:132-135)
 *   lin[i]       <- lin_agg                            (:162-171)
 *   quad[p]      <- quadratic_agg, p = packed upper-triangle index
 *                   i*n - i*(i+1)/2 + j (i<=j) for CFB_TRIPLE, p = i for CFB_NB (:177-199)
 *   cat_*        <- per categorical column k, keys ascending (std::map order):
 *                   cat_keys / cat_counts are the concatenation over k of the
 *                   lin_cat lists; cat_offsets[k]..cat_offsets[k+1] is column k (:372-395)
 *   numcat_sums  <- quad_num_cat: entry [i * total_keys + t] is SUM x_i over rows whose
 *                   column-k key is cat_keys[t]  (sub-list index num*m+cat, :383-404)
 *   pair_*       <- quad_cat: m(m+1)/2 lists in (k<=l) row-major order, diagonal
 *                   included, entries ascending by (key1,key2)        (:440-461)
 * Counts are exact integers; sums are fp64 (the glue narrows both to FLOAT, the
 * STRUCT's declared type).  All arrays are owned by the result; free with
 * cfb_result_free.                                                            */
typedef struct cfb_result {
  int32_t kind, n_num, n_cat;
  int64_t N;
  int64_t n_quad;        /* n(n+1)/2 (triple) or n (nb)                         */
  double *lin;           /* [n_num]                                             */
  double *quad;          /* [n_quad]                                            */
  int64_t total_keys;    /* sum over k of distinct keys                         */
  int64_t *cat_offsets;  /* [n_cat + 1]                                         */
  int32_t *cat_keys;     /* [total_keys]                                        */
  int64_t *cat_counts;   /* [total_keys]                                        */
  double *numcat_sums;   /* [n_num * total_keys]   (NULL for CFB_NB)            */
  int64_t n_pair_lists;  /* n_cat(n_cat+1)/2       (0 for CFB_NB)               */
  int64_t *pair_offsets; /* [n_pair_lists + 1]                                  */
  int32_t *pair_key1;    /* [pair_offsets[n_pair_lists]]                        */
  int32_t *pair_key2;
  int64_t *pair_counts;
} cfb_result;

/* Replaces Triple::SumStateFinalize for one group slot: synchronises, reads the
 * device partials back and emits the canonical result.                        */
int cfb_ctx_finalize(cfb_ctx *ctx, int group, cfb_result *out);
void cfb_result_free(cfb_result *res);

/* Ring product of two results -- replaces the arithmetic of Triple::MultiplyFunction
 * (mul.cpp:19-611) and Triple::multiply_nb (mul_nb.cpp) for factorised joins:
 *   N = Na*Nb;  lin = [Nb*lin_a | Na*lin_b];
 *   quad = packed upper triangle over the concatenated columns: [Nb*quad_a | lin_a (x) lin_b | Na*quad_b]
 *          (CFB_NB: [Nb*quad_a | Na*quad_b], the diagonal only);
 *   categorical columns of a, then of b: key counts scaled by the other side's N; per-key numeric
 *   sums scaled likewise or, across the two sides, lin_x[i] * count_y[key]; pair counts scaled
 *   likewise or, across the two sides, the outer product count_a[key1] * count_b[key2].
 * A host-side function on two small per-group results (the reference runs it per joined row);
 * `out` is owned by the caller afterwards (cfb_result_free).                               */
int cfb_result_multiply(const cfb_result *a, const cfb_result *b, cfb_result *out);

/* Ring sum / difference of two results: out = a + sign * b, sign = +1 or -1 -- the arithmetic of the Value-level
 * helpers Triple::sum_triple / subtract_triple / sum_nb_triple of the MICE drivers (imputation/triple/sum.cpp:69-209,
 * sub.cpp:71-216, sum_nb.cpp:38-82), used to maintain delta cofactors (all rows minus the rows where a column is
 * NULL).  Shapes must match (kind, n_num, n_cat).  Categorical parts are merged by key (a missing key counts 0).
 * flags = 0: keys (and key pairs) whose count becomes 0 are dropped, as a finalize of the corresponding state would;
 * flags = CFB_COMBINE_KEEP_ZERO_KEYS: they stay with count 0, as the reference's std::map merge leaves them
 * (sub.cpp:14-38) -- host/value_glue.cpp, the Value-level mirror, uses this.  A host-side function on two small
 * results; `out` is owned by the caller afterwards (cfb_result_free).                                          */
#define CFB_COMBINE_KEEP_ZERO_KEYS 1
int cfb_result_combine(const cfb_result *a, const cfb_result *b, int sign, int flags, cfb_result *out);

/* The write-back step as ring arithmetic ("predict -> delta-triple"): `rows` is the cofactor of some rows
 * R; numeric column `target` of those rows is about to be overwritten by the linear model's prediction
 *     y = bias + SUM_i w_num[i] * x_i + SUM_c w_cat[position of key_c]      (n_out = 1; the features are the other
 * numeric columns in order -- model->n_num == n_num - 1 -- and every categorical column).  Every entry of the NEW
 * cofactor that involves the target is a linear or quadratic function of entries that do not change:
 *     SUM y          = theta . (first row of the sigma matrix of R)
 *     SUM y x_i      = theta . (row i), SUM y [key_d = k] = theta . (row of that one-hot column)
 *     SUM y^2        = theta . (the new cross terms)
 * so `out` = the cofactor of R after the write-back, WITHOUT scanning the rows again (the reference recomputes the
 * delta with a second scan, imputation_low.cpp:85-110).  Exact up to the FLOAT rounding of the stored predictions;
 * not applicable to stochastic regression (noise = true) or to classifiers.  A host-side function on a small
 * result; `out` is owned by the caller afterwards (cfb_result_free).                                           */
struct cfb_linear_model;
int cfb_result_impute_linear(const cfb_result *rows, const struct cfb_linear_model *model, int target, cfb_result *out);

/* ------------------------------------------------- multi-GPU partial exchange */

/* Dense partial layout for the NCCL reduce of SURVEY 8(e): the caller
 * (one process per GPU) all-reduces / reduces two device buffers with SUM:
 *   f64 part: [group][lin | quad | numcat dense]      (cfb_ctx_partial_sizes)
 *   u64 part: [group][N | cat counts dense | pair counts dense]
 * export copies the context's device state into caller-provided device buffers,
 * import REPLACES the context's state with the buffers' contents.  Categorical
 * domains must have been agreed with cfb_ctx_set_cat_domain on every rank.
 * With a non-NULL `stream` (the stream the scans and the collective run on) both
 * calls are stream-ordered and do not synchronise the host; with NULL they run on
 * the context's stream and return when done.  Contexts whose pair counts are
 * hashed (large domains) have no dense partial: CFB_ERR_DOMAIN.               */
int cfb_ctx_partial_sizes(cfb_ctx *ctx, size_t *n_f64, size_t *n_u64);
int cfb_ctx_export_partial(cfb_ctx *ctx, void *d_f64, void *d_u64, void *stream);
int cfb_ctx_import_partial(cfb_ctx *ctx, const void *d_f64, const void *d_u64, void *stream);

/* The exchange step itself, inside the library ("one NCCL reduce", mirroring SumStateCombine,
 * sum_state.cpp:23-112, across GPUs): SUM-all-reduce of the context's dense state IN PLACE over `nccl_comm`
 * (an ncclComm_t) -- the fp64 sums and the u64 counts as one fused NCCL group, stream-ordered on `stream`
 * (cudaStream_t; NULL = the context's stream, then the call returns when done).  Every rank must have
 * declared the same categorical domain (cfb_nccl_agree_domain + cfb_ctx_set_cat_domain) and the same
 * n_groups.  NCCL is resolved from libnccl.so.2 at first use (CFB_NCCL_LIB overrides the name); the library
 * itself does not link it.  CFB_ERR_STATE when NCCL cannot be loaded.
 * States WITHOUT a dense partial that means the same on every rank -- hashed pair counts (large domains), key
 * dictionaries, or a domain every rank discovered for itself -- take SURVEY 8e's fallback instead: every rank
 * finalizes, the canonical results are all-gathered and merged by key in rank order on every rank (host-mediated,
 * synchronous); the context then HOLDS the global result: cfb_ctx_finalize returns it, further input is refused
 * (CFB_ERR_STATE).  All ranks must be in the same case: declare the domain on every rank or on none.      */
int cfb_ctx_allreduce(cfb_ctx *ctx, void *nccl_comm, void *stream);

/* Communicator plumbing for hosts that do not link NCCL themselves (one process per GPU): rank 0 asks for an
 * id, ships the 128 bytes to the other ranks by any channel, every rank creates its communicator.         */
#define CFB_NCCL_UNIQUE_ID_BYTES 128
int cfb_nccl_unique_id(void *id128);
int cfb_nccl_comm_create(int device, int world, int rank, const void *id128, void **nccl_comm_out);
int cfb_nccl_comm_destroy(void *nccl_comm);
/* Element-wise global [min lo, max hi] of per-rank key ranges (host arrays, in place): the domain agreement
 * that precedes a categorical multi-GPU scan.                                                              */
int cfb_nccl_agree_domain(void *nccl_comm, int device, int32_t *lo, int32_t *hi, int n_cat, void *stream);

/* Observed [min,max] of each categorical column over device-resident input
 * (device pre-pass; used to agree domains across ranks before the scan).      */
int cfb_cat_minmax_device(int device, const int32_t *const *d_cat_cols, int n_cat, size_t n_rows,
                          int32_t *lo_out, int32_t *hi_out, void *stream);

/* ---------------------------------------------- MICE write-back: model scores of rows */

/* A linear model over n_num FLOAT and n_cat INTEGER (one-hot) columns with n_out outputs:
 *     score_o(row) = bias[o] + SUM_i w_num[o][i] * x_i + SUM_c w_cat[o][ position of key_c in column c ]
 * This is the arithmetic of ML::linreg_impute (ML/regression.cpp:397-509, n_out = 1) and of
 * LDA_impute (ML/lda.cpp:421-590, n_out = #classes, result = index of the largest score); the
 * glue (host/predict_glue.cpp) parses the reference's parameter lists into this form and folds
 * the `normalize` centering into bias.  A key that the model does not know contributes 0.      */
typedef struct cfb_linear_model {
  int32_t n_num, n_cat, n_out;
  const double *bias;         /* [n_out]                                              */
  const double *w_num;        /* [n_out][n_num]                                       */
  const int64_t *cat_offsets; /* [n_cat + 1] into cat_keys / the columns of w_cat     */
  const int32_t *cat_keys;    /* [cat_offsets[n_cat]], ascending within a column      */
  const double *w_cat;        /* [n_out][cat_offsets[n_cat]]                          */
} cfb_linear_model;

/* The model on a device (weights and dense key maps uploaded once; a MICE step scores many chunks with it). */
typedef struct cfb_model cfb_model;
int cfb_model_create(int device, const cfb_linear_model *model, cfb_model **out);
void cfb_model_destroy(cfb_model *model);

/* Stochastic regression: linreg_predict(params, noise = true, ...) adds sigma * N(0, 1) to every prediction
 * (ML/regression.cpp:495-505; the MICE driver always asks for it, imputation_base.cpp:133).  The reference draws from
 * libc random() seeded from /dev/urandom; here the normal of row r is a pure function of (seed, first_row + r) --
 * Philox-4x32-10 + Box-Muller on the device -- so a query is reproducible whatever its chunking.  sigma = 0 switches
 * the noise off.  Applies to CFB_PREDICT_SCORE of a single-output linear model; first_row is where the next
 * cfb_predict_* call's rows start in the stream.                                                              */
int cfb_model_set_noise(cfb_model *model, double sigma, uint64_t seed, uint64_t first_row);

/* Gaussian naive Bayes (ML::nb_impute, ML/naive_bayes.cpp:153-263): per class k
 *     p_k(row) = prior[k] * PROD_j N(x_j; mean[k][j], var[k][j] + 1e-9) * PROD_c prob[k][position of key_c]
 * (0 for a key the model does not hold); the result is labels[first k with the largest p_k] (labels[0] when every
 * p_k is 0), computed in fp64 with the reference's expression.                                                  */
typedef struct cfb_nb_model {
  int32_t n_num, n_cat, n_classes;
  const int32_t *labels;      /* [n_classes]                                          */
  const double *prior;        /* [n_classes]                                          */
  const double *mean, *var;   /* [n_classes][n_num]                                   */
  const int64_t *cat_offsets; /* [n_cat + 1]                                          */
  const int32_t *cat_keys;    /* [cat_offsets[n_cat]], ascending within a column      */
  const double *cat_prob;     /* [n_classes][cat_offsets[n_cat]]                      */
} cfb_nb_model;
int cfb_model_create_nb(int device, const cfb_nb_model *model, cfb_model **out);

/* Quadratic discriminant analysis (ML::qda_impute, ML/qda.cpp:338-498): over the features f = [numeric | one-hot]
 * minus `center` (NULL = not normalized),  score_k = intercept[k] + f^T Q_k f + lin[k] . f ;  the result is
 * labels[first k with the largest score].  quad is [n_classes][p][p] with p = n_num + number of keys, each Q_k
 * column-major as the reference hands it to dgemv.                                                             */
typedef struct cfb_qda_model {
  int32_t n_num, n_cat, n_classes;
  const int32_t *labels;      /* [n_classes]                                          */
  const double *quad;         /* [n_classes][p][p]                                    */
  const double *lin;          /* [n_classes][p]                                       */
  const double *intercept;    /* [n_classes]                                          */
  const double *center;       /* [p] or NULL                                          */
  const int64_t *cat_offsets; /* [n_cat + 1]                                          */
  const int32_t *cat_keys;    /* [cat_offsets[n_cat]], ascending within a column      */
} cfb_qda_model;
int cfb_model_create_qda(int device, const cfb_qda_model *model, cfb_model **out);

#define CFB_PREDICT_SCORE 0  /* out: float  [rows], score_0                              */
#define CFB_PREDICT_ARGMAX 1 /* out: int32  [rows], first index of the largest score     */
#define CFB_PREDICT_LABEL 2  /* out: int32  [rows], the class label (naive Bayes / QDA models) */

/* Device-resident columns -> d_out (device).  d_row_mask (nullable, int32 per row): only rows with a
 * non-zero mask are written, the others keep their value -- with d_out aliasing the imputed column
 * this is the in-place overwrite of its NULL cells (imputation_base.cpp:75-83, :133-139).
 * Asynchronous on `stream` (cudaStream_t, NULL = the legacy default stream).                        */
int cfb_predict_device(cfb_model *model, const float *const *d_num_cols, const int32_t *const *d_cat_cols,
                       const int32_t *d_row_mask, size_t n_rows, int mode, void *d_out, void *stream);
/* Host columns (optional per-column selection vectors as in cfb_ctx_append) -> out (host); synchronous.
 * What the DuckDB scalar functions linreg_predict / lda_predict call per chunk.                       */
int cfb_predict_host(cfb_model *model, const float *const *num_cols, const uint32_t *const *num_sel,
                     const int32_t *const *cat_cols, const uint32_t *const *cat_sel, size_t count, int mode, void *out);

/* --------------------------------------------- the trainers' moment matrix and solves */

/* The p x p "sigma" matrix of the one-hot expanded design [1 | numeric columns | one column per (categorical column,
 * key)] -- what build_sigma_matrix assembles on the host for every trainer (ML/utils.cpp:176-310) -- assembled on the
 * device in fp64, with the reference's one-hot layout (n_cols_1hot_expansion, ML/utils.cpp:522-576: the keys that
 * occur, per column, ascending as uint64; drop_first removes each column's first key).
 *   label_cat >= 0   that categorical column is the class label of LDA: it is left out of the matrix, and the handle
 *                    also holds the per-class sums (build_sum_vector, ML/lda.cpp:58-144); -1 keeps every column.
 * cfb_sigma_from_ctx reads the dense device state of a context directly (no finalize; the only read-back is which
 * keys occur); contexts with hashed pair counts or key dictionaries go through cfb_ctx_finalize internally.      */
typedef struct cfb_sigma cfb_sigma;
int cfb_sigma_from_result(int device, const cfb_result *res, int label_cat, int drop_first, cfb_sigma **out);
int cfb_sigma_from_ctx(cfb_ctx *ctx, int group, int label_cat, int drop_first, cfb_sigma **out);
void cfb_sigma_destroy(cfb_sigma *sigma);
/* p; the number of classes (keys of label_cat; 0 without one); the length of cat_array (all columns' keys). */
int cfb_sigma_shape(const cfb_sigma *sigma, int32_t *p, int32_t *n_classes, int64_t *n_cat_values);
/* cat_array [n_cat_values] and cat_vars_idxs [n_cat + 1] as the trainers store them in their parameter lists. */
int cfb_sigma_layout(const cfb_sigma *sigma, int64_t *cat_array, int32_t *cat_vars_idxs);
/* The matrix (row-major, [p][p]) and the class sums ([n_classes][p]); either pointer may be NULL. */
int cfb_sigma_download(const cfb_sigma *sigma, double *sigma_out, double *class_sums_out);

/* ML::ridge_linear_regression (ML/regression.cpp:113-356) on the device: batch gradient descent with
 * Barzilai-Borwein steps and backtracking line search on sigma, one persistent cooperative kernel (the matrix stays
 * in L1 / L2; one grid barrier per matrix-vector product).  `label` is the numeric column to predict (0-based);
 * step_size and lambda are FLOAT as in the reference.  coeff [p]: intercept, then one coefficient per matrix column
 * (coeff[label + 1] = -1), already rescaled when normalize != 0; means [p] (normalize only); *variance =
 * theta^T Sigma theta / N of the final parameters (the reference emits its square root); *iterations = gradient steps
 * taken, *products = matrix-vector products spent (steps + backtracking trials + 1).  The last four may be NULL.   */
int cfb_sigma_linreg_train(cfb_sigma *sigma, int label, float step_size, float lambda, int max_iterations, int normalize,
                           double *coeff, double *means, double *variance, int32_t *iterations, int32_t *products);
/* lda_train (ML/lda.cpp:154-330) on the device: within-class covariance with shrinkage, divided by N, solved against
 * the class means by a blocked Cholesky factorisation (the reference calls dgelsd; for the positive definite matrix
 * shrinkage > 0 produces the two agree; a matrix that is not positive definite is CFB_ERR_STATE here).
 * coef [n_classes][p - 1], intercept [n_classes], means [p] (normalize only).                                     */
int cfb_sigma_lda_train(cfb_sigma *sigma, float shrinkage, int normalize, double *coef, double *intercept, double *means);

/* ------------------------------------------------------ synthetic inputs (tests, bench) */

/* Counter-based generators for device-resident synthetic columns: element i of the
 * stream is a pure function of (seed, first + i), so the host can regenerate any slice
 * bit-for-bit (duckdb_imputation_b200/synth.py) to feed the CPU oracle.
 *   uniform: float in [0,1) with 24 random bits;  int32: lo + (hash % range).        */
int cfb_gen_uniform_f32(int device, float *d_out, size_t n, uint64_t seed, uint64_t first, void *stream);
int cfb_gen_int32(int device, int32_t *d_out, size_t n, uint64_t seed, uint64_t first, int32_t lo,
                  uint32_t range, void *stream);

/* ------------------------------------------------------------- introspection */

/* Number of CUDA kernels this library has launched since load (bench.py's
 * gpu_launches claim), and time of the most recent cfb_triple_device scan kernel
 * in milliseconds measured with CUDA events on its launching stream (bench only:
 * enabled by cfb_set_timing(1)).                                               */
uint64_t cfb_kernel_launches(void);
int cfb_set_timing(int enabled);
double cfb_last_scan_ms(cfb_ctx *ctx);

/* The host half of the feed (cfb_ctx_append stages DuckDB vectors with non-temporal stores): which vector ISA
 * the staging copy picked on this CPU ("avx512" / "avx2" / "sse2"), and the host-memory ceiling of that copy on
 * this box -- `threads` threads copying private `bytes_per_thread` buffers in 8 KB pieces, aggregate GB/s of
 * payload (nt = 1: the staging copy, nt = 0: memcpy).  bench.py quotes it beside the end-to-end number.   */
const char *cfb_stage_isa(void);
double cfb_host_copy_ceiling(size_t bytes_per_thread, int threads, int nt, int reps);

#ifdef __cplusplus
}
#endif
#endif /* COFACTOR_B200_H */
