// triple_glue.h -- the DuckDB aggregate callbacks of the ring aggregates, B200 build.
//
// Same names, signatures and registration shape as the reference's
// duckdb_extension/src/include/triple/sum/{sum_state.h,sum_no_lift.h,sum_to_nb_agg.h}; the
// bodies call the C ABI of include/cofactor_b200.h instead of looping on the CPU.
#pragma once
#include <duckdb.hpp>

struct cfb_ctx;

namespace Triple {

// sum_state.h:14-28 shrunk to a handle.  DuckDB relocates states with memcpy (hash-table row
// storage), so the state is a plain pointer to extension-owned memory and nothing else.
struct SumState {
  cfb_ctx *ctx;
};

struct StateFunction {
  template <class STATE>
  static void Initialize(STATE &state) {  // sum_state.h:33-45
    state.ctx = nullptr;
  }
  template <class STATE>
  static void Destroy(STATE &state, duckdb::AggregateInputData &aggr_input_data);  // sum_state.h:48-52
  static bool IgnoreNull() { return false; }  // NULL rows are delivered (sum_state.h:54-56)
};

duckdb::unique_ptr<duckdb::FunctionData> SumNoLiftBind(duckdb::ClientContext &context, duckdb::AggregateFunction &function,
                                                       duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void SumNoLift(duckdb::Vector inputs[], duckdb::AggregateInputData &aggr_input_data, idx_t input_count,
               duckdb::Vector &state_vector, idx_t count);

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
void CustomLift(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> CustomLiftBind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                        duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);
void to_nb_lift(duckdb::DataChunk &args, duckdb::ExpressionState &state, duckdb::Vector &result);
duckdb::unique_ptr<duckdb::FunctionData> to_nb_lift_bind(duckdb::ClientContext &context, duckdb::ScalarFunction &function,
                                                         duckdb::vector<duckdb::unique_ptr<duckdb::Expression>> &arguments);

void SumStateCombine(duckdb::Vector &state, duckdb::Vector &combined, duckdb::AggregateInputData &aggr_input_data, idx_t count);
void SumStateFinalize(duckdb::Vector &state_vector, duckdb::AggregateInputData &aggr_input_data, duckdb::Vector &result,
                      idx_t count, idx_t offset);

}  // namespace Triple
