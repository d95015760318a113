// "merge": merge-path SpMV (partition kernel + tile kernel + carry fixup) in libspmvb200.
// Takes the place of SpMV_merge_based / SpMV_merge_based_generalized / SpMV_cub_merge_based
// (reference/include/spmv/merge_based/merge_based.cuh:22, merge_genl/merge_genl.cuh:41,
//  cub_merge.cuh:20).
#pragma once
#include "abi_dispatch.hpp"
SPMV_DEFINE_KIND_TEMPLATE(SpMV_merge_path, merge)
