// triple_glue.h -- the DuckDB aggregate callbacks of the ring aggregates, B200 build.
//
// Same names, signatures and registration shape as the reference's
// duckdb_extension/src/include/triple/sum/{sum_state.h,sum_no_lift.h,sum_to_nb_agg.h}; the
// bodies call the C ABI of include/cofactor_b200.h instead of looping on the CPU.
#pragma once
#include <atomic>
#include <mutex>

#include <duckdb.hpp>

#include "../../../include/cofactor_b200.h"

namespace Triple {

// An arena is one device context (cfb_ctx) that holds the aggregate states of MANY GROUP BY groups of
// one worker thread as slots: DuckDB hands update() a state pointer per row, the glue turns it into a
// slot id per row and ships the chunk once -- the GPU routes rows to slots.  Arenas are reference
// counted by the states that live in them (and by the thread that is still filling them).
struct Arena {
  cfb_ctx *ctx = nullptr;
  int kind = 0, n = 0, m = 0;
  int capacity = 1;          // GROUP BY slots of ctx
  int next_slot = 0;         // only the owning thread hands out slots
  std::atomic<int> refs{0};  // live states + 1 while a thread keeps it open
  // A cfb_ctx is driven by one thread at a time, but the states of ONE arena can reach combine / finalize on
  // different threads at once (DuckDB finalizes the radix partitions of a hash aggregate in parallel; the groups
  // of a partition come from every worker's arena).  Every entry into the context takes this lock; combine
  // takes both arenas' locks in address order.
  std::mutex mu;
  void Release() {
    if (refs.fetch_sub(1) == 1) {
      cfb_ctx_destroy(ctx);
      delete this;
    }
  }
};

// sum_state.h:14-28 shrunk to a handle.  DuckDB relocates states with memcpy (hash-table row
// storage), so the state is plain data: which arena, which slot.
struct SumState {
  Arena *arena;
  int32_t slot;
};

struct StateFunction {
  template <class STATE>
  static void Initialize(STATE &state) {  // sum_state.h:33-45
    state.arena = nullptr;
    state.slot = 0;
  }
  template <class STATE>
  static void Destroy(STATE &state, duckdb::AggregateInputData &aggr_input_data);  // sum_state.h:48-52
  static bool IgnoreNull() { return false; }  // NULL rows are delivered (sum_state.h:54-56)
};

// The context of a state that is fed alone (ungrouped aggregates, sum_triple / sum_nb_agg): a private
// arena with one slot, created on first use.
cfb_ctx *PrivateContext(SumState &state, int kind, int n_num, int n_cat);
// GROUP BY: a state that has no arena yet gets a slot in the calling thread's open arena of this shape.
void AssignSlot(SumState &state, int kind, int n_num, int n_cat);

duckdb::unique_ptr<duckdb::FunctionData> SumNoLiftBind(duckdb::ClientContext &context, duckdb::AggregateFunction &function,
                                                       duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void SumNoLift(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, idx_t input_count,
               duckdb::Vector &state_vector, idx_t count);
// simple_update (ungrouped aggregates: one state, no per-row state pointers).  The reference leaves the slot
// null (duckdb_imputation_extension.cpp:53,106), which forces the hash aggregate even without GROUP BY.
void SumNoLiftSimple(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, idx_t input_count,
                     duckdb::data_ptr_t state, idx_t count);
void sum_to_nb_agg_simple(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, idx_t input_count,
                          duckdb::data_ptr_t state, idx_t count);

duckdb::unique_ptr<duckdb::FunctionData> sum_to_nb_agg_bind(duckdb::ClientContext &context, duckdb::AggregateFunction &function,
                                                            duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void sum_to_nb_agg(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, duckdb::idx_t cols,
                   duckdb::Vector &state_vector, duckdb::idx_t count);

// sum_triple / sum_nb_agg over lifted triples (sum.h, sum_nb_agg.h) and the lifts (lift.h, lift_to_nb_agg.h)
duckdb::unique_ptr<duckdb::FunctionData> SumBind(duckdb::ClientContext &context, duckdb::AggregateFunction &function,
                                                 duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void Sum(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, idx_t input_count,
         duckdb::Vector &state_vector, idx_t count);
duckdb::unique_ptr<duckdb::FunctionData> sum_nb_agg_bind(duckdb::ClientContext &context, duckdb::AggregateFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void sum_nb_agg(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, idx_t input_count,
                duckdb::Vector &state_vector, idx_t count);
void SumSimple(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, idx_t input_count,
               duckdb::data_ptr_t state, idx_t count);
void sum_nb_agg_simple(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, idx_t input_count,
                       duckdb::data_ptr_t state, idx_t count);
void CustomLift(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> CustomLiftBind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                        duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void to_nb_lift(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> to_nb_lift_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);

// multiply_triple / multiply_nb_agg: the ring product as a scalar function (mul.h, mul_nb.h)
void MultiplyFunction(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> MultiplyBind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                      duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void multiply_nb(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> multiply_nb_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                          duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);

void SumStateCombine(duckdb::Vector &state, duckdb::Vector &combined, duckdb::AggregateInputData &aggr_input_data, idx_t count);
void SumStateFinalize(duckdb::Vector &state_vector, duckdb::AggregateInputData &aggr_input_data, duckdb::Vector &result,
                      idx_t count, idx_t offset);

// Writes canonical results (one per output row) into the nested STRUCT vector (sum_state.cpp:132-461).
void WriteResults(const std::vector<cfb_result> &res, duckdb::Vector &result, bool nb, int n_num, int n_cat);

}  // namespace Triple

// linreg_predict / lda_predict / nb_predict / qda_predict: the write-back step of MICE (ML/regression.h, ML/lda.h, ...)
namespace ML {
void linreg_impute(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> linreg_impute_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                            duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void LDA_impute(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> LDA_impute_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
// nb_predict / qda_predict (ML/naive_bayes.h, ML/qda.h)
void nb_impute(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> nb_impute_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                        duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void qda_impute(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> qda_impute_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
// linreg_train (ML/regression.h): sigma assembly + gradient descent on the device (train_glue.cpp)
void ridge_linear_regression(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> ridge_linear_regression_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                                      duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
}  // namespace ML
// lda_train (ML/lda.h:8, a global function in the reference too)
void lda_train(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> lda_train_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                        duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
