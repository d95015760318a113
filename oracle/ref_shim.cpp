// ref_shim.cpp -- C-ABI doorway onto the UNMODIFIED reference CPU code.
//
// TEST INFRASTRUCTURE ONLY (see oracle/spmv_oracle.c for the rules).  This file
// contains no algorithm of its own: it #includes the reference headers from where
// they lie under /root/reference/include (never copied into this repo) and
// forwards extern "C" calls to the reference templates.  oracle/Makefile compiles it
// into oracle/_ref/libspmv_ref.so, which pins oracle/spmv_oracle.c and may serve as
// bench.py's cpu_baseline with kind "reference".
//
//   reference/include/spmv/cpu_navie.hpp:3-17    SpMV_cpu_navie
//   reference/include/spmv/cpu_navie.hpp:20-35   SpMV_genl_cpu_navie
//   reference/include/load.hpp:268-408           LoadCoo
//   reference/include/load.hpp:420-474           ToCsr
//   reference/include/spmv/merge_based/thread_search.cuh:16-49  SearchMergePath
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>   // load.hpp uses std::cerr without including it (load.hpp:279)
#include <limits>     // load.hpp uses std::numeric_limits without including it (load.hpp:302)
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "load.hpp"
#include "spmv/cpu_navie.hpp"

// thread_search.cuh is a CUDA header, but its one function is plain C++ once the
// execution-space qualifiers and CUB's min/max macros have host definitions.
#define __host__
#define __device__
#define __forceinline__ inline
#define CUB_MAX(a, b) (((b) > (a)) ? (b) : (a))
#define CUB_MIN(a, b) (((b) < (a)) ? (b) : (a))
#include "spmv/merge_based/thread_search.cuh"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

// The "b" list of the merge: the natural numbers (cub::CountingInputIterator in
// the reference, merge_based/dispatch_spmv_orig.cuh:131).
template <typename T>
struct Counting {
    using iterator_category = std::random_access_iterator_tag;
    using value_type = T;
    using difference_type = T;
    using pointer = const T *;
    using reference = T;
    T base;
    T operator[](T i) const { return base + i; }
};

template <typename T>
struct Coord {
    T x, y;
};

struct AbsFunctor {
    static double initialize() { return 0.0; }
    static double combine(double a, double x) { return std::fabs(a * x); }
    static double reduce(double a, double b) { return a + b; }
};

// the fixed semiring menu of include/spmv_b200.h as functor_t's for SpMV_genl_cpu_navie
template <typename T> struct MinPlusFunctor {
    static T initialize() { return std::numeric_limits<T>::infinity(); }
    static T combine(T a, T x) { return a + x; }
    static T reduce(T u, T v) { return u < v ? u : v; }
};
template <typename T> struct MaxPlusFunctor {
    static T initialize() { return -std::numeric_limits<T>::infinity(); }
    static T combine(T a, T x) { return a + x; }
    static T reduce(T u, T v) { return u > v ? u : v; }
};
template <typename T> struct OrAndFunctor {
    static T initialize() { return T(0); }
    static T combine(T a, T x) { return (a != T(0) && x != T(0)) ? T(1) : T(0); }
    static T reduce(T u, T v) { return u > v ? u : v; }
};

}  // namespace

// ---- SpMV_genl_cpu_navie with the semiring menu (1 min-plus, 2 max-plus, 3 or-and) ----
REF_API int ref_genl_o32_f32(int semiring, int32_t n_rows, int32_t n_cols, int32_t nnz,
                             const int32_t *Ap, const int32_t *Aj, const float *Ax, const float *x,
                             float *y) {
    switch (semiring) {
        case 1: SpMV_genl_cpu_navie<MinPlusFunctor<float>, int, int, float, float, float>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y); return 0;
        case 2: SpMV_genl_cpu_navie<MaxPlusFunctor<float>, int, int, float, float, float>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y); return 0;
        case 3: SpMV_genl_cpu_navie<OrAndFunctor<float>, int, int, float, float, float>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y); return 0;
        default: return 1;
    }
}
REF_API int ref_genl_o32_f64(int semiring, int32_t n_rows, int32_t n_cols, int32_t nnz,
                             const int32_t *Ap, const int32_t *Aj, const double *Ax, const double *x,
                             double *y) {
    switch (semiring) {
        case 1: SpMV_genl_cpu_navie<MinPlusFunctor<double>, int, int, double, double, double>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y); return 0;
        case 2: SpMV_genl_cpu_navie<MaxPlusFunctor<double>, int, int, double, double, double>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y); return 0;
        case 3: SpMV_genl_cpu_navie<OrAndFunctor<double>, int, int, double, double, double>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y); return 0;
        default: return 1;
    }
}

// ---- SpMV_cpu_navie, the instantiations the configs need (32-bit offsets only:
// the reference's inner counter is index_t, cpu_navie.hpp:12) ----
REF_API void ref_spmv_o32_f32(int32_t n_rows, int32_t n_cols, int32_t nnz, const int32_t *Ap,
                              const int32_t *Aj, const float *Ax, const float *x, float *y) {
    SpMV_cpu_navie<int, int, float, float, float>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y);
}
REF_API void ref_spmv_o32_f64(int32_t n_rows, int32_t n_cols, int32_t nnz, const int32_t *Ap,
                              const int32_t *Aj, const double *Ax, const double *x, double *y) {
    SpMV_cpu_navie<int, int, double, double, double>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y);
}
REF_API void ref_spmv_o32_f32_acc64(int32_t n_rows, int32_t n_cols, int32_t nnz,
                                    const int32_t *Ap, const int32_t *Aj, const float *Ax,
                                    const double *x, double *y) {
    SpMV_cpu_navie<int, int, float, double, double>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y);
}
// int64 offsets with the reference's own loop: valid only while every offset < 2^31.
REF_API void ref_spmv_o64_f32(int32_t n_rows, int32_t n_cols, int64_t nnz, const int64_t *Ap,
                              const int32_t *Aj, const float *Ax, const float *x, float *y) {
    SpMV_cpu_navie<int, int64_t, float, float, float>(n_rows, n_cols, nnz, Ap, Aj, Ax, x, y);
}

// ---- SpMV_genl_cpu_navie with the |a*x| functor: the per-row tolerance scale ----
REF_API void ref_abs_o32_f32(int32_t n_rows, int32_t n_cols, int32_t nnz, const int32_t *Ap,
                             const int32_t *Aj, const float *Ax, const double *x, double *s) {
    SpMV_genl_cpu_navie<AbsFunctor, int, int, float, double, double>(n_rows, n_cols, nnz, Ap, Aj,
                                                                     Ax, x, s);
}
REF_API void ref_abs_o32_f64(int32_t n_rows, int32_t n_cols, int32_t nnz, const int32_t *Ap,
                             const int32_t *Aj, const double *Ax, const double *x, double *s) {
    SpMV_genl_cpu_navie<AbsFunctor, int, int, double, double, double>(n_rows, n_cols, nnz, Ap,
                                                                      Aj, Ax, x, s);
}

// ---- the reference loop on all host threads: each thread calls the unmodified
// SpMV_cpu_navie on a contiguous row block (Ap + r0, y + r0).  Returns threads used. ----
REF_API int ref_spmv_mt_o32_f32(int32_t n_rows, int32_t n_cols, int32_t nnz, const int32_t *Ap,
                                const int32_t *Aj, const float *Ax, const float *x, float *y,
                                int n_threads) {
    int used = 1;
    if (n_threads < 1) n_threads = 1;
    int64_t n_blocks = std::max<int64_t>(1, std::min<int64_t>((int64_t)n_threads * 16, n_rows));
#pragma omp parallel num_threads(n_threads)
    {
#ifdef _OPENMP
#pragma omp single
        used = omp_get_num_threads();
#endif
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < n_blocks; ++b) {
            int32_t r0 = (int32_t)((int64_t)n_rows * b / n_blocks);
            int32_t r1 = (int32_t)((int64_t)n_rows * (b + 1) / n_blocks);
            SpMV_cpu_navie<int, int, float, float, float>(r1 - r0, n_cols, nnz, Ap + r0, Aj, Ax,
                                                          x, y + r0);
        }
    }
    return used;
}
REF_API int ref_spmv_mt_o64_f32(int32_t n_rows, int32_t n_cols, int64_t nnz, const int64_t *Ap,
                                const int32_t *Aj, const float *Ax, const float *x, float *y,
                                int n_threads) {
    int used = 1;
    if (n_threads < 1) n_threads = 1;
    int64_t n_blocks = std::max<int64_t>(1, std::min<int64_t>((int64_t)n_threads * 16, n_rows));
#pragma omp parallel num_threads(n_threads)
    {
#ifdef _OPENMP
#pragma omp single
        used = omp_get_num_threads();
#endif
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < n_blocks; ++b) {
            int32_t r0 = (int32_t)((int64_t)n_rows * b / n_blocks);
            int32_t r1 = (int32_t)((int64_t)n_rows * (b + 1) / n_blocks);
            SpMV_cpu_navie<int, int64_t, float, float, float>(r1 - r0, n_cols, nnz, Ap + r0, Aj,
                                                              Ax, x, y + r0);
        }
    }
    return used;
}
REF_API int ref_spmv_mt_o32_f64(int32_t n_rows, int32_t n_cols, int32_t nnz, const int32_t *Ap,
                                const int32_t *Aj, const double *Ax, const double *x, double *y,
                                int n_threads) {
    int used = 1;
    if (n_threads < 1) n_threads = 1;
    int64_t n_blocks = std::max<int64_t>(1, std::min<int64_t>((int64_t)n_threads * 16, n_rows));
#pragma omp parallel num_threads(n_threads)
    {
#ifdef _OPENMP
#pragma omp single
        used = omp_get_num_threads();
#endif
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < n_blocks; ++b) {
            int32_t r0 = (int32_t)((int64_t)n_rows * b / n_blocks);
            int32_t r1 = (int32_t)((int64_t)n_rows * (b + 1) / n_blocks);
            SpMV_cpu_navie<int, int, double, double, double>(r1 - r0, n_cols, nnz, Ap + r0, Aj,
                                                             Ax, x, y + r0);
        }
    }
    return used;
}

// ---- SearchMergePath exactly as DeviceSpmvSearchKernel calls it ----
REF_API void ref_merge_path_search_o32(int32_t diagonal, int32_t n_rows, int32_t nnz,
                                       const int32_t *Ap, int32_t *out_x, int32_t *out_y) {
    Coord<int32_t> c{0, 0};
    Counting<int32_t> nz{0};
    merge_spmv::SearchMergePath(diagonal, Ap + 1, nz, n_rows, nnz, c);
    *out_x = c.x;
    *out_y = c.y;
}
REF_API void ref_merge_path_search_o64(int64_t diagonal, int64_t n_rows, int64_t nnz,
                                       const int64_t *Ap, int64_t *out_x, int64_t *out_y) {
    Coord<int64_t> c{0, 0};
    Counting<int64_t> nz{0};
    merge_spmv::SearchMergePath(diagonal, Ap + 1, nz, n_rows, nnz, c);
    *out_x = c.x;
    *out_y = c.y;
}

// ---- Matrix Market -> COO -> CSR through the reference loader ----
// Two-call protocol: ref_load_mtx opens + converts and returns a handle; the sizes
// are read back, the caller allocates, ref_load_mtx_copy fills and frees.
struct RefCsrF32 {
    csr_t<int, int, float> csr;
};
REF_API void *ref_load_mtx_f32(const char *filename, int64_t *n_rows, int64_t *n_cols,
                               int64_t *nnz) {
    try {
        auto *h = new RefCsrF32{ToCsr(LoadCoo<int, int, float>(filename))};
        *n_rows = h->csr.number_of_rows;
        *n_cols = h->csr.number_of_columns;
        *nnz = h->csr.number_of_nonzeros;
        return h;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_load_mtx_f32: %s\n", e.what());
        return nullptr;
    }
}
REF_API void ref_load_mtx_f32_copy(void *handle, int32_t *Ap, int32_t *Aj, float *Ax) {
    auto *h = static_cast<RefCsrF32 *>(handle);
    std::copy(h->csr.row_offsets.begin(), h->csr.row_offsets.end(), Ap);
    std::copy(h->csr.column_indices.begin(), h->csr.column_indices.end(), Aj);
    std::copy(h->csr.nonzero_values.begin(), h->csr.nonzero_values.end(), Ax);
    delete h;
}

// ---- ToCsr on caller-provided COO arrays ----
REF_API void ref_coo_to_csr_o32_f32(int32_t n_rows, int32_t n_cols, int32_t nnz,
                                    const int32_t *rows, const int32_t *cols, const float *vals,
                                    int32_t *Ap, int32_t *Aj, float *Ax) {
    coo_t<int, int, float> coo(n_rows, n_cols, nnz);
    std::copy(rows, rows + nnz, coo.row_indices.begin());
    std::copy(cols, cols + nnz, coo.column_indices.begin());
    std::copy(vals, vals + nnz, coo.nonzero_values.begin());
    auto csr = ToCsr(coo);
    std::copy(csr.row_offsets.begin(), csr.row_offsets.end(), Ap);
    std::copy(csr.column_indices.begin(), csr.column_indices.end(), Aj);
    std::copy(csr.nonzero_values.begin(), csr.nonzero_values.end(), Ax);
}

REF_API int ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
