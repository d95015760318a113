// "vector": CSR-vector SpMV (sub-warp per row, 128-bit loads, shuffle reduce) in libspmvb200.
// Takes the place of SpMV_cusp_origin / SpMV_cusp_warp_reduce / SpMV_cusp_warp_read_reduce
// (reference/include/spmv/cusp/cusp.cuh:227, cusp_warp_reduce.cuh:138,
//  cusp_warp_read_reduce.cuh:144).
#pragma once
#include "abi_dispatch.hpp"
SPMV_DEFINE_KIND_TEMPLATE(SpMV_csr_vector, vector)
