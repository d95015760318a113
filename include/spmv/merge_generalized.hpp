// "merge_genl": the generalised (semiring) merge-path SpMV of libspmvb200.
// Takes the place of SpMV_merge_based_generalized (reference/include/spmv/merge_genl/
// merge_genl.cuh:41-84).  The reference parameterises it with a functor_t{initialize, combine,
// reduce} (merge_genl.cuh:19-38); a template functor cannot cross a C ABI, so the semiring is
// one of a fixed menu (SPMVB200_SEMIRING_*).  The registered kind uses plus-times, like the
// reference's registry entry (spmv.h:27); SpMV_merge_semiring<S> reaches the others.
#pragma once
#include <cstring>

#include "abi_dispatch.hpp"

template <int SEMIRING, typename index_t, typename offset_t, typename mat_value_t,
          typename vec_x_value_t, typename vec_y_value_t>
void SpMV_merge_semiring(index_t n_rows, index_t n_cols, offset_t nnz, const offset_t *Ap,
                         const index_t *Aj, const mat_value_t *Ax, const vec_x_value_t *x,
                         vec_y_value_t *y) {
    static_assert(spmv_abi::check_types<index_t, offset_t, mat_value_t, vec_x_value_t,
                                        vec_y_value_t>::ok, "");
    spmvb200_args_t a;
    std::memset(&a, 0, sizeof(a));
    a.kind = SPMVB200_KIND_MERGE;
    a.offset_bits = (int32_t)sizeof(offset_t) * 8;
    a.value_bits = (int32_t)sizeof(mat_value_t) * 8;
    a.n_rows = n_rows;
    a.n_cols = n_cols;
    a.nnz = (int64_t)nnz;
    a.Ap = Ap;
    a.Aj = Aj;
    a.Ax = Ax;
    a.x = x;
    a.y = y;
    a.stream = (void *)SpmvStream::get();
    a.semiring = SEMIRING;
    Timer::kernel_start();
    const int status = spmvb200_spmv(&a);
    Timer::kernel_stop();
    checkSpmvStatus(status);
}

template <typename index_t, typename offset_t, typename mat_value_t, typename vec_x_value_t,
          typename vec_y_value_t>
void SpMV_merge_generalized(index_t n_rows, index_t n_cols, offset_t nnz, const offset_t *Ap,
                            const index_t *Aj, const mat_value_t *Ax, const vec_x_value_t *x,
                            vec_y_value_t *y) {
    SpMV_merge_semiring<SPMVB200_SEMIRING_PLUS_TIMES>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y);
}
