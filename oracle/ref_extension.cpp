// ref_extension.cpp -- TEST INFRASTRUCTURE.  Registers the REFERENCE's own ring aggregates
// (compiled unmodified from /root/reference/duckdb_extension/src/triple/sum/*.cpp over the
// DuckDB-vector shim) under the names and with the constructor arguments the reference uses
// in load_ring / load_nb_ring (duckdb_imputation_extension.cpp:80-113, :146-179), so that the
// replay host can drive them exactly like ours.  Grid [0,19] as in the reference.
#include <duckdb.hpp>

#include <triple/sum/sum_no_lift.h>
#include <triple/sum/sum_state.h>
#include <triple/sum/sum_to_nb_agg.h>

namespace duckdb_ring {

const char *Implementation() { return "reference"; }

void Load(duckdb::DatabaseInstance &db) {
  using namespace duckdb;
  for (int i = 0; i < 20; i++)
    for (int j = 0; j < 20; j++) {
      if (i == 0 && j == 0) continue;
      vector<LogicalType> args;
      for (int k = 0; k < i; k++) args.push_back(LogicalType::FLOAT);
      for (int k = 0; k < j; k++) args.push_back(LogicalType::INTEGER);
      const std::string suffix = std::to_string(i) + "_" + std::to_string(j);
      AggregateFunction triple("sum_to_triple_" + suffix, args, LogicalTypeId::STRUCT,
                               AggregateFunction::StateSize<Triple::SumState>,
                               AggregateFunction::StateInitialize<Triple::SumState, Triple::StateFunction>,
                               Triple::SumNoLift, Triple::SumStateCombine, Triple::SumStateFinalize, nullptr,
                               Triple::SumNoLiftBind,
                               AggregateFunction::StateDestroy<Triple::SumState, Triple::StateFunction>, nullptr, nullptr);
      triple.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
      ExtensionUtil::RegisterFunction(db, triple);
      AggregateFunction nb("sum_to_nb_agg_" + suffix, args, LogicalTypeId::STRUCT,
                           AggregateFunction::StateSize<Triple::SumState>,
                           AggregateFunction::StateInitialize<Triple::SumState, Triple::StateFunction>,
                           Triple::sum_to_nb_agg, Triple::SumStateCombine, Triple::SumStateFinalize, nullptr,
                           Triple::sum_to_nb_agg_bind,
                           AggregateFunction::StateDestroy<Triple::SumState, Triple::StateFunction>, nullptr, nullptr);
      nb.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
      ExtensionUtil::RegisterFunction(db, nb);
    }
}

}  // namespace duckdb_ring
