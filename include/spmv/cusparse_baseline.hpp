// "cusparse": cusparseSpMV(CUSPARSE_SPMV_ALG_DEFAULT) with handle, descriptors and buffer kept
// across calls.  Takes the place of SpMV_cusparse (reference/include/spmv/cusparse.cuh:37-88),
// which times handle creation and cudaMalloc on every call.  Comparison baseline only.
#pragma once
#include "abi_dispatch.hpp"
SPMV_DEFINE_KIND_TEMPLATE(SpMV_cusparse, cusparse)
