// ref_extension.cpp -- TEST INFRASTRUCTURE.  Registers the REFERENCE's own ring aggregates
// (compiled unmodified from /root/reference/duckdb_extension/src/triple/sum/*.cpp over the
// DuckDB-vector shim) under the names and with the constructor arguments the reference uses
// in load_ring / load_nb_ring (duckdb_imputation_extension.cpp:80-113, :146-179), so that the
// replay host can drive them exactly like ours.  Grid [0,19] as in the reference.
#include <duckdb.hpp>

#include <triple/sum/sum_no_lift.h>
#include <triple/sum/sum_state.h>
#include <triple/sum/sum_to_nb_agg.h>

#include <ML/lda.h>
#include <ML/naive_bayes.h>
#include <ML/qda.h>
#include <ML/regression.h>

namespace {
// The local state ML::linreg_impute expects (regression.cpp:399).  The reference's own init function
// (regression.cpp:377-394) casts the BIND data to RegressionState and seeds libc random() from /dev/urandom; here the
// state is simply constructed -- the generator is seeded by the test that asks for noise.
duckdb::unique_ptr<duckdb::FunctionLocalState> FreshRegressionState(duckdb::ExpressionState &, const duckdb::BoundFunctionExpression &,
                                                                  duckdb::FunctionData *) {
  return duckdb::make_uniq<RegressionState>();
}
}  // namespace

namespace duckdb_ring {

const char *Implementation() { return "reference"; }

void Load(duckdb::DatabaseInstance &db) {
  using namespace duckdb;
  // the predict side of the write-back step, as load_ml registers it (duckdb_imputation_extension.cpp:193-249); the
  {
    ScalarFunction lda_predict("lda_predict", {LogicalType::ANY}, LogicalTypeId::INTEGER, LDA_impute, LDA_impute_bind, nullptr, LDA_impute_stats);
    lda_predict.varargs = LogicalType::ANY;
    lda_predict.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
    ExtensionUtil::RegisterFunction(db, lda_predict);
    ScalarFunction linreg_predict("linreg_predict", {LogicalType::ANY}, LogicalTypeId::INTEGER, ML::linreg_impute, ML::linreg_impute_bind, nullptr,
                                  nullptr, FreshRegressionState);
    linreg_predict.varargs = LogicalType::ANY;
    linreg_predict.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
    ExtensionUtil::RegisterFunction(db, linreg_predict);
    // the trainers of load_ml (duckdb_imputation_extension.cpp:184-189, :202-207): pins for SURVEY 8 f4
    ScalarFunction lda_train_func("lda_train", {LogicalType::ANY}, LogicalTypeId::LIST, lda_train, lda_train_bind, nullptr);
    lda_train_func.varargs = LogicalType::ANY;
    lda_train_func.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
    ExtensionUtil::RegisterFunction(db, lda_train_func);
    ScalarFunction linreg_train_func("linreg_train", {LogicalType::ANY}, LogicalTypeId::LIST, ML::ridge_linear_regression,
                                     ML::ridge_linear_regression_bind, nullptr);
    linreg_train_func.varargs = LogicalType::ANY;
    linreg_train_func.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
    ExtensionUtil::RegisterFunction(db, linreg_train_func);
    // qda_predict (duckdb_imputation_extension.cpp:234-240); ML/qda.cpp through the build copy of oracle/Makefile
    ScalarFunction qda_predict("qda_predict", {LogicalType::ANY}, LogicalTypeId::INTEGER, ML::qda_impute, ML::qda_impute_bind, nullptr);
    qda_predict.varargs = LogicalType::ANY;
    qda_predict.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
    ExtensionUtil::RegisterFunction(db, qda_predict);
    // the per-class trainers (duckdb_imputation_extension.cpp:219-224, :236-241): they make the parameter lists the
    // reference's own QDA / naive-Bayes tests feed to qda_predict / nb_predict
    ScalarFunction qda_train_func("qda_train", {LogicalType::ANY}, LogicalTypeId::LIST, ML::qda_train, ML::qda_train_bind, nullptr);
    qda_train_func.varargs = LogicalType::ANY;
    qda_train_func.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
    ExtensionUtil::RegisterFunction(db, qda_train_func);
    ScalarFunction nb_train_func("nb_train", {LogicalType::ANY}, LogicalTypeId::LIST, ML::nb_train, ML::nb_train_bind, nullptr);
    nb_train_func.varargs = LogicalType::ANY;
    nb_train_func.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
    ExtensionUtil::RegisterFunction(db, nb_train_func);
    ScalarFunction nb_predict("nb_predict", {LogicalType::ANY}, LogicalTypeId::INTEGER, ML::nb_impute, ML::nb_impute_bind, nullptr);
    nb_predict.varargs = LogicalType::ANY;
    nb_predict.null_handling = FunctionNullHandling::SPECIAL_HANDLING;
    ExtensionUtil::RegisterFunction(db, nb_predict);
  }
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
