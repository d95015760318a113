// "auto": host-side selector (row-length statistics -> merge / vector / light) in libspmvb200.
// New: the reference has no cross-kind selector (BASELINE.json north_star adds it).
#pragma once
#include "abi_dispatch.hpp"
SPMV_DEFINE_KIND_TEMPLATE(SpMV_auto_select, auto)
