// Test doorway onto include/load.hpp (the product's host loader): plain C entry points so
// tests/test_loader.py can drive LoadCoo/ToCsr through ctypes.
#include <cstdint>
#include <cstring>
#include <string>

#include "load.hpp"

namespace {
template <typename O, typename V>
struct Held {
    csr_t<int, O, V> csr;
};
}  // namespace

#define API extern "C" __attribute__((visibility("default")))

// returns MM_* code, or 100 + n for exception_t (message copied to err)
template <typename O, typename V>
static int load_impl(const char *file, void **handle, int64_t *n_rows, int64_t *n_cols, int64_t *nnz,
                     int *scheme, char *err, int errlen) {
    try {
        coo_t<int, O, V> coo(0, 0, 0);
        mm_header_t h;
        const int rc = TryLoadCoo(std::string(file), coo, &h);
        if (rc != MM_OK) return rc;
        auto *held = new Held<O, V>{ToCsr(coo)};
        *handle = held;
        *n_rows = held->csr.number_of_rows;
        *n_cols = held->csr.number_of_columns;
        *nnz = (int64_t)held->csr.number_of_nonzeros;
        *scheme = (int)h.scheme;
        return 0;
    } catch (const exception_t &e) {
        std::strncpy(err, e.what(), (size_t)errlen - 1);
        err[errlen - 1] = 0;
        return 100;
    }
}
API int shim_load_o32_f32(const char *file, void **handle, int64_t *n_rows, int64_t *n_cols, int64_t *nnz,
                          int *scheme, char *err, int errlen) {
    return load_impl<int, float>(file, handle, n_rows, n_cols, nnz, scheme, err, errlen);
}
API int shim_load_o64_f64(const char *file, void **handle, int64_t *n_rows, int64_t *n_cols, int64_t *nnz,
                          int *scheme, char *err, int errlen) {
    return load_impl<int64_t, double>(file, handle, n_rows, n_cols, nnz, scheme, err, errlen);
}
API void shim_copy_o32_f32(void *handle, int32_t *Ap, int32_t *Aj, float *Ax) {
    auto *h = static_cast<Held<int, float> *>(handle);
    std::copy(h->csr.row_offsets.begin(), h->csr.row_offsets.end(), Ap);
    std::copy(h->csr.column_indices.begin(), h->csr.column_indices.end(), Aj);
    std::copy(h->csr.nonzero_values.begin(), h->csr.nonzero_values.end(), Ax);
    delete h;
}
API void shim_copy_o64_f64(void *handle, int64_t *Ap, int32_t *Aj, double *Ax) {
    auto *h = static_cast<Held<int64_t, double> *>(handle);
    std::copy(h->csr.row_offsets.begin(), h->csr.row_offsets.end(), Ap);
    std::copy(h->csr.column_indices.begin(), h->csr.column_indices.end(), Aj);
    std::copy(h->csr.nonzero_values.begin(), h->csr.nonzero_values.end(), Ax);
    delete h;
}
