// "light": LightSpMV-style dynamic row hand-out in libspmvb200.
// Takes the place of SpMV_light_vector / SpMV_light_warp
// (reference/include/spmv/LightSpMV.cuh:379, :399).
#pragma once
#include "abi_dispatch.hpp"
SPMV_DEFINE_KIND_TEMPLATE(SpMV_dynamic_rows, light)
