// forwarding header: the reference includes <duckdb/function/scalar/nested_functions.hpp>
// for VariableReturnBindData, which the shim's duckdb.hpp already declares.
#pragma once
#include "../../../duckdb.hpp"
