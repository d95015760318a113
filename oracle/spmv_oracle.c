/*
 * spmv_oracle.c -- CPU restatement of the reference's CSR SpMV hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (spmv_samples_b200/, include/,
 * main.cu) links, imports or calls this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs,
 * and only as the checker or the reported CPU baseline.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle.py)
 *   (a) against the reference's one known-answer vector, the 3x3 lattice of
 *       /root/reference/include/spmv/merge_based/device_spmv.cuh:95-128, and
 *   (b) against the reference's own code compiled from where it lies
 *       (oracle/_ref/libspmv_ref.so, built by oracle/Makefile from
 *       /root/reference/include/spmv/cpu_navie.hpp, load.hpp and
 *       merge_based/thread_search.cuh), on seeded random inputs, bit for bit;
 *       the outputs of (b) are also committed under tests/golden/ so the check
 *       still runs where /root/reference does not exist.
 *
 * Each function cites the reference lines it restates.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- *
 * y = A*x, sequential row loop, k ascending, accumulator in the y type.
 * Reference: include/spmv/cpu_navie.hpp:3-17 (SpMV_cpu_navie).
 *
 * Deviation, flagged: the reference's inner counter is index_t (cpu_navie.hpp:12)
 * and so cannot pass 2^31-1; here it is the offset type, which is what the int64
 * offset configuration (R-MAT scale 27, nnz = 2^31) needs.  For nnz < 2^31 the
 * two loops execute the same operations in the same order.
 * ------------------------------------------------------------------------- */
#define DEF_SPMV(NAME, OFF_T, MAT_T, X_T, Y_T)                                         \
    ORACLE_API void NAME(int64_t n_rows, const OFF_T *Ap, const int32_t *Aj,           \
                         const MAT_T *Ax, const X_T *x, Y_T *y) {                      \
        for (int64_t row = 0; row < n_rows; ++row) {                                   \
            Y_T sum = (Y_T)0;                                                          \
            for (OFF_T k = Ap[row]; k < Ap[row + 1]; ++k) {                            \
                sum += Ax[k] * x[Aj[k]];                                               \
            }                                                                          \
            y[row] = sum;                                                              \
        }                                                                              \
    }

/* native-precision instantiations (what main.cu:79-81 runs for fp32) */
DEF_SPMV(oracle_spmv_o32_f32, int32_t, float, float, float)
DEF_SPMV(oracle_spmv_o64_f32, int64_t, float, float, float)
DEF_SPMV(oracle_spmv_o32_f64, int32_t, double, double, double)
DEF_SPMV(oracle_spmv_o64_f64, int64_t, double, double, double)
/* fp64 host reference for fp32 matrices: SpMV_cpu_navie<int, off, float, double, double>;
 * the product Ax[k]*x[j] is float*double -> double, the sum is double. */
DEF_SPMV(oracle_spmv_o32_f32_acc64, int32_t, float, double, double)
DEF_SPMV(oracle_spmv_o64_f32_acc64, int64_t, float, double, double)

/* ------------------------------------------------------------------------- *
 * Per-row tolerance scale  s[r] = sum_k |Ax[k] * x[Aj[k]]|  in fp64.
 * Reference: include/spmv/cpu_navie.hpp:20-35 (SpMV_genl_cpu_navie) with
 * functor {initialize = 0, combine = |a*x|, reduce = +}.
 * ------------------------------------------------------------------------- */
#define DEF_ABS(NAME, OFF_T, MAT_T)                                                    \
    ORACLE_API void NAME(int64_t n_rows, const OFF_T *Ap, const int32_t *Aj,           \
                         const MAT_T *Ax, const double *x, double *s) {                \
        for (int64_t row = 0; row < n_rows; ++row) {                                   \
            double sum = 0.0;                                                          \
            for (OFF_T k = Ap[row]; k < Ap[row + 1]; ++k) {                            \
                sum = sum + fabs((double)Ax[k] * x[Aj[k]]);                            \
            }                                                                          \
            s[row] = sum;                                                              \
        }                                                                              \
    }
DEF_ABS(oracle_abs_o32_f32, int32_t, float)
DEF_ABS(oracle_abs_o64_f32, int64_t, float)
DEF_ABS(oracle_abs_o32_f64, int32_t, double)
DEF_ABS(oracle_abs_o64_f64, int64_t, double)

/* ------------------------------------------------------------------------- *
 * Generalised (semiring) SpMV: y[r] = REDUCE_k COMBINE(Ax[k], x[Aj[k]]) from IDENTITY,
 * sequential, k ascending.  Reference: include/spmv/cpu_navie.hpp:20-35
 * (SpMV_genl_cpu_navie<functor_t>) with the functors of the fixed menu that
 * include/spmv_b200.h offers in place of a template functor:
 *   1 min-plus (+inf, a+x, min)   2 max-plus (-inf, a+x, max)   3 or-and (0, a!=0&&x!=0, max)
 * ------------------------------------------------------------------------- */
#define DEF_GENL(NAME, OFF_T, VAL_T)                                                   \
    ORACLE_API int NAME(int semiring, int64_t n_rows, const OFF_T *Ap, const int32_t *Aj, \
                        const VAL_T *Ax, const VAL_T *x, VAL_T *y) {                   \
        if (semiring < 1 || semiring > 3) return 1;                                    \
        for (int64_t row = 0; row < n_rows; ++row) {                                   \
            VAL_T sum = semiring == 1 ? (VAL_T)INFINITY                                \
                        : semiring == 2 ? (VAL_T)-INFINITY : (VAL_T)0;                 \
            for (OFF_T k = Ap[row]; k < Ap[row + 1]; ++k) {                            \
                const VAL_T a = Ax[k], xv = x[Aj[k]];                                  \
                if (semiring == 1) { const VAL_T c = a + xv; sum = sum < c ? sum : c; } \
                else if (semiring == 2) { const VAL_T c = a + xv; sum = sum > c ? sum : c; } \
                else { const VAL_T c = (a != 0 && xv != 0) ? (VAL_T)1 : (VAL_T)0; sum = sum > c ? sum : c; } \
            }                                                                          \
            y[row] = sum;                                                              \
        }                                                                              \
        return 0;                                                                      \
    }
DEF_GENL(oracle_genl_o32_f32, int32_t, float)
DEF_GENL(oracle_genl_o64_f32, int64_t, float)
DEF_GENL(oracle_genl_o32_f64, int32_t, double)
DEF_GENL(oracle_genl_o64_f64, int64_t, double)

/* ------------------------------------------------------------------------- *
 * Row-parallel driver of the same loop for the CPU baseline timing: threads take
 * contiguous row blocks, each block runs the sequential loop above unchanged.
 * Not reference code (the reference is single threaded); reported with its
 * thread count.  Returns the number of threads used.
 * ------------------------------------------------------------------------- */
#define DEF_SPMV_MT(NAME, SEQ, OFF_T, MAT_T, X_T, Y_T)                                 \
    ORACLE_API int NAME(int64_t n_rows, const OFF_T *Ap, const int32_t *Aj,            \
                        const MAT_T *Ax, const X_T *x, Y_T *y, int n_threads) {        \
        int used = 1;                                                                  \
        if (n_threads < 1) n_threads = 1;                                              \
        int64_t n_blocks = (int64_t)n_threads * 16;                                    \
        if (n_blocks > n_rows) n_blocks = n_rows > 0 ? n_rows : 1;                     \
        _Pragma("omp parallel num_threads(n_threads)")                                 \
        {                                                                              \
            _Pragma("omp single")                                                      \
            { used = omp_get_num_threads(); }                                          \
            _Pragma("omp for schedule(dynamic, 1)")                                    \
            for (int64_t b = 0; b < n_blocks; ++b) {                                   \
                int64_t r0 = n_rows * b / n_blocks;                                    \
                int64_t r1 = n_rows * (b + 1) / n_blocks;                              \
                SEQ(r1 - r0, Ap + r0, Aj, Ax, x, y + r0);                              \
            }                                                                          \
        }                                                                              \
        return used;                                                                   \
    }
#ifdef _OPENMP
DEF_SPMV_MT(oracle_spmv_mt_o32_f32, oracle_spmv_o32_f32, int32_t, float, float, float)
DEF_SPMV_MT(oracle_spmv_mt_o64_f32, oracle_spmv_o64_f32, int64_t, float, float, float)
DEF_SPMV_MT(oracle_spmv_mt_o32_f64, oracle_spmv_o32_f64, int32_t, double, double, double)
DEF_SPMV_MT(oracle_spmv_mt_o64_f64, oracle_spmv_o64_f64, int64_t, double, double, double)
#endif

/* ------------------------------------------------------------------------- *
 * Merge-path diagonal search.
 * Reference: include/spmv/merge_based/thread_search.cuh:16-49 (SearchMergePath)
 * as called from merge_based/dispatch_spmv_orig.cuh:129-146 with
 *   a = row end offsets = Ap + 1 (length n_rows), b = 0,1,2,... (length nnz).
 * Writes (x = rows consumed, y = nonzeros consumed) for the given diagonal.
 * ------------------------------------------------------------------------- */
#define DEF_SEARCH(NAME, OFF_T)                                                        \
    ORACLE_API void NAME(int64_t diagonal, int64_t n_rows, int64_t nnz,                \
                         const OFF_T *Ap, int64_t *out_x, int64_t *out_y) {            \
        const OFF_T *row_end = Ap + 1;                                                 \
        int64_t lo = diagonal - nnz > 0 ? diagonal - nnz : 0;                          \
        int64_t hi = diagonal < n_rows ? diagonal : n_rows;                            \
        while (lo < hi) {                                                              \
            int64_t pivot = (lo + hi) >> 1;                                            \
            if ((int64_t)row_end[pivot] <= diagonal - pivot - 1) {                     \
                lo = pivot + 1;                                                        \
            } else {                                                                   \
                hi = pivot;                                                            \
            }                                                                          \
        }                                                                              \
        *out_x = lo < n_rows ? lo : n_rows;                                            \
        *out_y = diagonal - lo;                                                        \
    }
DEF_SEARCH(oracle_merge_path_search_o32, int32_t)
DEF_SEARCH(oracle_merge_path_search_o64, int64_t)

/* Tile start coordinates for all tiles+1 diagonals t*tile_items, clamped to the
 * path length n_rows+nnz.  Reference: merge_based/dispatch_spmv_orig.cuh:109-148
 * (DeviceSpmvSearchKernel), :613-623 (num_merge_items, num_merge_tiles). */
#define DEF_TILES(NAME, SEARCH, OFF_T)                                                 \
    ORACLE_API void NAME(int64_t n_rows, int64_t nnz, const OFF_T *Ap,                 \
                         int64_t tile_items, int64_t n_coords, int64_t *cx,            \
                         int64_t *cy) {                                                \
        int64_t total = n_rows + nnz;                                                  \
        for (int64_t t = 0; t < n_coords; ++t) {                                       \
            int64_t d = t * tile_items;                                                \
            if (d > total) d = total;                                                  \
            SEARCH(d, n_rows, nnz, Ap, cx + t, cy + t);                                \
        }                                                                              \
    }
DEF_TILES(oracle_merge_tile_coords_o32, oracle_merge_path_search_o32, int32_t)
DEF_TILES(oracle_merge_tile_coords_o64, oracle_merge_path_search_o64, int64_t)

/* nnz-balanced row split for P shards: diagonal
 * d_g = floor(g*(n_rows+nnz)/P), boundary row = x coordinate of the search. */
#define DEF_SPLIT(NAME, SEARCH, OFF_T)                                                 \
    ORACLE_API void NAME(int64_t n_rows, int64_t nnz, const OFF_T *Ap, int64_t parts,  \
                         int64_t *row_bounds) {                                        \
        int64_t total = n_rows + nnz;                                                  \
        for (int64_t g = 0; g <= parts; ++g) {                                         \
            int64_t d = (int64_t)(((__int128)g * (__int128)total) / parts);            \
            int64_t cx, cy;                                                            \
            SEARCH(d, n_rows, nnz, Ap, &cx, &cy);                                      \
            row_bounds[g] = cx;                                                        \
        }                                                                              \
        row_bounds[parts] = n_rows;                                                    \
    }
DEF_SPLIT(oracle_row_split_o32, oracle_merge_path_search_o32, int32_t)
DEF_SPLIT(oracle_row_split_o64, oracle_merge_path_search_o64, int64_t)

/* ------------------------------------------------------------------------- *
 * COO -> CSR by counting sort, stable in input order, duplicates kept, columns
 * left unsorted within a row.
 * Reference: include/load.hpp:420-474 (ToCsr).  Deviation, flagged: the
 * reference keeps running sums / scatter cursors in index_t (load.hpp:448-452,
 * :458-459, :467-471); here they are 64-bit so nnz >= 2^31 works.  Below 2^31 the
 * result is identical.
 * ------------------------------------------------------------------------- */
#define DEF_TOCSR(NAME, OFF_T, VAL_T)                                                  \
    ORACLE_API int NAME(int64_t n_rows, int64_t nnz, const int32_t *rows,              \
                        const int32_t *cols, const VAL_T *vals, OFF_T *Ap,             \
                        int32_t *Aj, VAL_T *Ax) {                                      \
        int64_t *cursor = (int64_t *)calloc((size_t)n_rows + 1, sizeof(int64_t));      \
        if (!cursor) return 1;                                                         \
        for (int64_t n = 0; n < nnz; ++n) cursor[rows[n]]++;                           \
        int64_t sum = 0;                                                               \
        for (int64_t i = 0; i < n_rows; ++i) {                                         \
            int64_t c = cursor[i];                                                     \
            cursor[i] = sum;                                                           \
            Ap[i] = (OFF_T)sum;                                                        \
            sum += c;                                                                  \
        }                                                                              \
        Ap[n_rows] = (OFF_T)nnz;                                                       \
        for (int64_t n = 0; n < nnz; ++n) {                                            \
            int64_t dest = cursor[rows[n]]++;                                          \
            Aj[dest] = cols[n];                                                        \
            Ax[dest] = vals[n];                                                        \
        }                                                                              \
        free(cursor);                                                                  \
        return 0;                                                                      \
    }
DEF_TOCSR(oracle_coo_to_csr_o32_f32, int32_t, float)
DEF_TOCSR(oracle_coo_to_csr_o64_f32, int64_t, float)
DEF_TOCSR(oracle_coo_to_csr_o32_f64, int32_t, double)
DEF_TOCSR(oracle_coo_to_csr_o64_f64, int64_t, double)

ORACLE_API int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
