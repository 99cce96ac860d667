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
  // Outside the reference's grid: the headline config sum_to_triple_20_0 (BASELINE.json) is
  // not registered by the reference (its loops stop at 19 although README.md:136 says 20).
  // The reference's callbacks are generic in the column count, so the CPU baseline times the
  // unmodified Triple::SumNoLift on 20 FLOAT columns under a name that cannot be mistaken
  // for a registered function.
  {
    vector<LogicalType> args(20, LogicalType::FLOAT);
    AggregateFunction extra("ref_sum_to_triple_20_0", args, LogicalTypeId::STRUCT,
                            AggregateFunction::StateSize<Triple::SumState>,
                            AggregateFunction::StateInitialize<Triple::SumState, Triple::StateFunction>,
                            Triple::SumNoLift, Triple::SumStateCombine, Triple::SumStateFinalize, nullptr,
                            Triple::SumNoLiftBind,
                            AggregateFunction::StateDestroy<Triple::SumState, Triple::StateFunction>, nullptr, nullptr);
    ExtensionUtil::RegisterFunction(db, extra);
  }
}

}  // namespace duckdb_ring
