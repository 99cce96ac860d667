// forwarding header: the include path the reference uses (lda.cpp:12)
#pragma once
#include "../../../duckdb.hpp"
