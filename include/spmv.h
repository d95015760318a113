// spmv.h -- the SpMV registry: same entry point, same X-macro shape as
// reference/include/spmv.h:18-48; the kinds behind it are the sm_100a kernels of
// libspmvb200 reached through the C ABI (include/spmv_b200.h).
//
// Adding a kind is what it is in the reference (README.md:28-46): one header with a template
// of the 8-argument per-kind signature, one X(...) line here.
#pragma once

#include <iostream>
#include <string>

#include "common.cuh"
#include "spmv/auto_select.hpp"
#include "spmv/csr_stream.hpp"
#include "spmv/csr_vector.hpp"
#include "spmv/cusparse_baseline.hpp"
#include "spmv/dynamic_rows.hpp"
#include "spmv/host_check.hpp"
#include "spmv/merge_generalized.hpp"
#include "spmv/merge_path.hpp"

/// SPMV kind strings and its function.  The second block keeps every label of the reference's
/// table (spmv.h:18-27) working on a reference-style command line: each names the kernel here
/// that replaces the reference kernel of that label.
#define SPMV_KINDS                                                             \
    X("merge", SpMV_merge_path)                                                \
    X("merge_genl", SpMV_merge_generalized)                                    \
    X("vector", SpMV_csr_vector)                                               \
    X("light", SpMV_dynamic_rows)                                              \
    X("stream", SpMV_csr_stream)                                               \
    X("auto", SpMV_auto_select)                                                \
    X("cusparse", SpMV_cusparse)                                               \
    X("cusp", SpMV_csr_vector)                                                 \
    X("cusp1", SpMV_csr_vector)                                                \
    X("cusp2", SpMV_csr_vector)                                                \
    X("light_vec", SpMV_dynamic_rows)                                          \
    X("light_warp", SpMV_dynamic_rows)                                         \
    X("cub_merge", SpMV_merge_path)

template <typename index_t, typename offset_t, typename mat_value_t,
          typename vec_x_value_t, typename vec_y_value_t>
void SpMV(const std::string& kind_str,
    index_t n_rows,  index_t n_cols, offset_t nnz,
    const offset_t *Ap, const index_t *Aj, const mat_value_t *Ax,
    const vec_x_value_t *x, vec_y_value_t *y) {

    // the X-macro expands into a label -> function table for this instantiation
    using kind_fn = void (*)(index_t, index_t, offset_t, const offset_t *, const index_t *,
                             const mat_value_t *, const vec_x_value_t *, vec_y_value_t *);
    struct kind_entry { const char *label; kind_fn fn; };
    static const kind_entry kinds[] = {
    #define X(a, b) {a, &b<index_t, offset_t, mat_value_t, vec_x_value_t, vec_y_value_t>},
        SPMV_KINDS
    #undef X
    };
    for (const kind_entry &k : kinds) {
        if (kind_str == k.label) {
            Timer::total_start();
            k.fn(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y);
            Timer::total_stop();
            return;
        }
    }
    // the reference's message and exit code (spmv.h:46-47)
    std::cerr << "SpMV kind \"" << kind_str << "\" is NOT SUPPROT\n";
    exit(EXIT_FAILURE);
}
