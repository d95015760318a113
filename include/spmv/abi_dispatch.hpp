// abi_dispatch.hpp -- maps the reference's template parameters
// <index_t, offset_t, mat_value_t, vec_x_value_t, vec_y_value_t>
// (reference/include/spmv.h:29-34) onto the typed C-ABI symbols of libspmvb200
// (include/spmv_b200.h).  Unsupported combinations are compile-time errors, the way an
// unsupported type is a missing cuSPARSE mapping in reference/include/spmv/cusparse.cuh:23-33.
#pragma once

#include <cstdint>
#include <type_traits>

#include "../common.cuh"
#include "../spmv_b200.h"

namespace spmv_abi {

template <typename index_t, typename offset_t, typename mat_value_t, typename vec_x_value_t,
          typename vec_y_value_t>
struct check_types {
    static_assert(std::is_same<index_t, int32_t>::value, "index_t must be a 32-bit int");
    static_assert(std::is_same<offset_t, int32_t>::value || std::is_same<offset_t, int64_t>::value ||
                      (std::is_same<offset_t, long long>::value && sizeof(long long) == 8),
                  "offset_t must be int32 or int64");
    static_assert(std::is_same<mat_value_t, float>::value || std::is_same<mat_value_t, double>::value,
                  "mat_value_t must be float or double");
    static_assert(std::is_same<mat_value_t, vec_x_value_t>::value &&
                      std::is_same<mat_value_t, vec_y_value_t>::value,
                  "matrix, x and y must share one value type");
    static constexpr bool ok = true;
};

#define SPMV_ABI_KIND(KIND)                                                                        \
    inline int call_##KIND(int32_t r, int32_t c, int32_t nnz, const int32_t *Ap, const int32_t *Aj, \
                           const float *Ax, const float *x, float *y, void *s) {                   \
        return spmvb200_##KIND##_i32_o32_f32(r, c, nnz, Ap, Aj, Ax, x, y, s);                      \
    }                                                                                              \
    inline int call_##KIND(int32_t r, int32_t c, int32_t nnz, const int32_t *Ap, const int32_t *Aj, \
                           const double *Ax, const double *x, double *y, void *s) {                \
        return spmvb200_##KIND##_i32_o32_f64(r, c, nnz, Ap, Aj, Ax, x, y, s);                      \
    }                                                                                              \
    inline int call_##KIND(int32_t r, int32_t c, int64_t nnz, const int64_t *Ap, const int32_t *Aj, \
                           const float *Ax, const float *x, float *y, void *s) {                   \
        return spmvb200_##KIND##_i32_o64_f32(r, c, nnz, Ap, Aj, Ax, x, y, s);                      \
    }                                                                                              \
    inline int call_##KIND(int32_t r, int32_t c, int64_t nnz, const int64_t *Ap, const int32_t *Aj, \
                           const double *Ax, const double *x, double *y, void *s) {                \
        return spmvb200_##KIND##_i32_o64_f64(r, c, nnz, Ap, Aj, Ax, x, y, s);                      \
    }
SPMV_ABI_KIND(merge)
SPMV_ABI_KIND(vector)
SPMV_ABI_KIND(light)
SPMV_ABI_KIND(stream)
SPMV_ABI_KIND(auto)
SPMV_ABI_KIND(cusparse)
#undef SPMV_ABI_KIND

// `long long` offsets (8 bytes, distinct from int64_t = long on LP64) forward as int64_t
template <typename offset_t>
struct abi_offset {
    using type = typename std::conditional<sizeof(offset_t) == 8, int64_t, int32_t>::type;
};

}  // namespace spmv_abi

// Defines  template <...> void NAME(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y)  with the
// reference's per-kind signature (reference/include/spmv.h:39), forwarding to KIND.
#define SPMV_DEFINE_KIND_TEMPLATE(NAME, KIND)                                                      \
    template <typename index_t, typename offset_t, typename mat_value_t, typename vec_x_value_t,   \
              typename vec_y_value_t>                                                              \
    void NAME(index_t n_rows, index_t n_cols, offset_t nnz, const offset_t *Ap, const index_t *Aj, \
              const mat_value_t *Ax, const vec_x_value_t *x, vec_y_value_t *y) {                   \
        static_assert(spmv_abi::check_types<index_t, offset_t, mat_value_t, vec_x_value_t,         \
                                            vec_y_value_t>::ok, "");                               \
        using abi_off = typename spmv_abi::abi_offset<offset_t>::type;                             \
        Timer::kernel_start();                                                                     \
        const int status = spmv_abi::call_##KIND(                                                  \
            (int32_t)n_rows, (int32_t)n_cols, (abi_off)nnz, reinterpret_cast<const abi_off *>(Ap), \
            Aj, Ax, x, y, (void *)SpmvStream::get());                                              \
        Timer::kernel_stop();                                                                      \
        checkSpmvStatus(status);                                                                   \
    }
