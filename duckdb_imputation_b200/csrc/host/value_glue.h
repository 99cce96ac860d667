/* value_glue.h -- the Value-level ring helpers of the MICE drivers, same names and signatures as the reference's
 * imputation/include/sum_sub.h:10-14 (Triple::subtract_triple, Triple::sum_triple, Triple::sum_nb_triple): a driver
 * that maintains delta cofactors (imputation_low.cpp:85-110: full - delta, train, predict, full + new delta) links
 * these instead of imputation/triple/{sum,sub,sum_nb}.cpp.  The arithmetic runs in cfb_result_combine (C ABI). */
#ifndef CFB_VALUE_GLUE_H
#define CFB_VALUE_GLUE_H
#include <duckdb.hpp>

namespace Triple {
duckdb::Value subtract_triple(duckdb::Value &triple_1, duckdb::Value &triple_2);
duckdb::Value sum_triple(const duckdb::Value &triple_1, const duckdb::Value &triple_2);
duckdb::Value sum_nb_triple(const duckdb::Value &triple_1, const duckdb::Value &triple_2);
}  // namespace Triple
#endif
