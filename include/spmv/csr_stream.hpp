// "stream": CSR-stream SpMV (persistent CTAs, row tiles staged in shared memory by TMA bulk
// copies, one thread per row) in libspmvb200 -- the kernel for short regular rows, the case the
// reference gives to SpMV_cusp_* with 2- and 4-lane vectors
// (reference/include/spmv/cusp/cusp.cuh:189-203).
#pragma once
#include "abi_dispatch.hpp"
SPMV_DEFINE_KIND_TEMPLATE(SpMV_csr_stream, stream)
