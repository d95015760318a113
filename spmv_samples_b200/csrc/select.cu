// select.cu -- row-length statistics, the host-side kernel selector ("auto"), and the
// cuSPARSE comparison baseline.
//
// The reference picks sub-warp widths from nnz/n_rows alone (reference/include/spmv/cusp/
// cusp.cuh:187-221, LightSpMV.cuh:345-370) and has no cross-kind selector; BASELINE.json's
// north_star adds one.  Statistics come from one device pass over Ap and are cached per
// matrix (keyed on the Ap pointer, n_rows, nnz): stale statistics can only cost speed, never
// correctness, because every kernel is correct for every CSR matrix.
//
// The cuSPARSE wrapper replaces reference/include/spmv/cusparse.cuh:37-88 with the handle,
// descriptors, work buffer and cusparseSpMV_preprocess hoisted into a one-entry plan cache,
// so the timed call is cusparseSpMV alone (hard part 7).
#include <cusparse.h>

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace spmvb200 {

namespace {

struct StatsDev {
    unsigned long long max_len;
    unsigned long long empty;
    double sum_sq;
    double pad;
};

template <typename OffT>
__global__ void __launch_bounds__(256)
row_stats_kernel(int64_t n_rows, const OffT *__restrict__ Ap, StatsDev *out) {
    unsigned long long mx = 0, empty = 0;
    double sq = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long len = (unsigned long long)(__ldg(Ap + r + 1) - __ldg(Ap + r));
        mx = len > mx ? len : mx;
        empty += (len == 0);
        sq += (double)len * (double)len;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long omx = __shfl_xor_sync(0xffffffffu, mx, s);
        mx = omx > mx ? omx : mx;
        empty += __shfl_xor_sync(0xffffffffu, empty, s);
        sq += __shfl_xor_sync(0xffffffffu, sq, s);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&out->max_len, mx);
        atomicAdd(&out->empty, empty);
        atomicAdd(&out->sum_sq, sq);
    }
}

using StatsKey = std::tuple<int, const void *, int64_t, int64_t>;
std::mutex g_stats_mu;
std::map<StatsKey, spmvb200_row_stats_t> g_stats;

void choose(spmvb200_row_stats_t &st) {
    const int64_t forced = option_get("auto_kind", -1);
    st.chosen_width = pick_width_from_mean(st.mean_row_len);
    if (forced >= 0) {
        st.chosen_kind = (int32_t)forced;
        return;
    }
    // Regular matrices (every row close to the mean) take the CSR-vector kernel, whose sub-warp
    // width is picked from the mean; anything else takes merge-path, whose cost does not depend
    // on how the nonzeros are spread over rows.  Thresholds from tools/selector_sweep.py
    // (profiles/r2_selector_sweep.txt, 13 matrices between regular and power-law, 4 Mi rows): a
    // sub-warp sized for the mean already loses to merge-path at std = 0.5 mean (rows of 1..31
    // nonzeros: 358 us against 310; log-normal sigma 0.5: 368 against 310; half the rows empty:
    // 344 against 312), so the line is drawn at 0.4 mean.  The dynamic-row kernel ("light") is
    // never the fastest of the four on any of the 13 and is not selected.
    const double mean = st.mean_row_len > 1.0 ? st.mean_row_len : 1.0;
    const bool heavy_tail = (double)st.max_row_len > 16.0 * mean + 64.0 ||
                            st.std_row_len > 0.4 * mean;
    const bool mostly_empty = st.n_rows > 0 && st.empty_rows * 2 > st.n_rows;
    st.chosen_kind = (heavy_tail || mostly_empty) ? SPMVB200_KIND_MERGE : SPMVB200_KIND_VECTOR;
    // Short regular rows: a sub-warp per row wastes most of its 128-bit load slots (a 5-nonzero
    // row fills 5 of 8) and every row is a chain of dependent round trips; the CSR-stream kernel
    // moves the matrix with TMA bulk copies and gives a row to a thread.  Only worth its pipeline
    // on a matrix large enough to fill the persistent grid a few times over.
    // Up to 6 nonzeros per row: at 8 the thread-per-row reads of the staged tile are 8-way bank
    // conflicts and the 2-lane CSR-vector kernel wins (band matrix, 8 per row: 86 us against 70);
    // at 3 and 5 per row it is 70 against 82 and 20.5 against 22.5 us.
    if (st.chosen_kind == SPMVB200_KIND_VECTOR && st.mean_row_len <= 6.0 && st.max_row_len <= 64 &&
        st.n_rows >= (int64_t)1 << 16)
        st.chosen_kind = SPMVB200_KIND_STREAM;
}

}  // namespace

void stats_cache_clear() {
    std::lock_guard<std::mutex> lk(g_stats_mu);
    g_stats.clear();
}

template <typename OffT>
int row_stats(int64_t n_rows, int64_t nnz, const OffT *Ap, spmvb200_row_stats_t *out,
              cudaStream_t stream, bool use_cache) {
    int dev = -1;
    SPMV_CUDA_TRY(cudaGetDevice(&dev));
    const StatsKey key{dev, (const void *)Ap, n_rows, nnz};
    if (use_cache) {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        auto it = g_stats.find(key);
        if (it != g_stats.end()) {
            *out = it->second;
            choose(*out);  // options may have changed since
            return SPMVB200_OK;
        }
    }
    spmvb200_row_stats_t st{};
    st.n_rows = n_rows;
    st.nnz = nnz;
    st.mean_row_len = n_rows > 0 ? (double)nnz / (double)n_rows : 0.0;
    if (n_rows > 0) {
        const DeviceInfo *di = nullptr;
        SPMV_TRY(current_device_info(&di));
        void *dbuf = nullptr;
        SPMV_TRY(scratch_get(stream, SCRATCH_STATS, sizeof(StatsDev), &dbuf));
        SPMV_CUDA_TRY(cudaMemsetAsync(dbuf, 0, sizeof(StatsDev), stream));
        int64_t blocks = (n_rows + 255) / 256;
        const int64_t cap = (int64_t)di->sm_count * 8;
        if (blocks > cap) blocks = cap;
        row_stats_kernel<OffT><<<(unsigned)blocks, 256, 0, stream>>>(n_rows, Ap,
                                                                     static_cast<StatsDev *>(dbuf));
        SPMV_LAUNCH_CHECK();
        StatsDev h{};
        SPMV_CUDA_TRY(cudaMemcpyAsync(&h, dbuf, sizeof(h), cudaMemcpyDeviceToHost, stream));
        SPMV_CUDA_TRY(cudaStreamSynchronize(stream));
        st.max_row_len = (int64_t)h.max_len;
        st.empty_rows = (int64_t)h.empty;
        const double var = h.sum_sq / (double)n_rows - st.mean_row_len * st.mean_row_len;
        st.std_row_len = var > 0.0 ? std::sqrt(var) : 0.0;
    }
    choose(st);
    {
        std::lock_guard<std::mutex> lk(g_stats_mu);
        if (g_stats.size() > 64) g_stats.clear();
        g_stats[key] = st;
    }
    *out = st;
    return SPMVB200_OK;
}
template int row_stats<int32_t>(int64_t, int64_t, const int32_t *, spmvb200_row_stats_t *, cudaStream_t, bool);
template int row_stats<int64_t>(int64_t, int64_t, const int64_t *, spmvb200_row_stats_t *, cudaStream_t, bool);

template <typename OffT, typename ValT>
int launch_auto(const SpmvProblem<OffT, ValT> &p) {
    if (p.n_rows <= 0) return SPMVB200_OK;
    spmvb200_row_stats_t st;
    SPMV_TRY(row_stats<OffT>(p.n_rows, (int64_t)p.nnz, p.Ap, &st, p.stream, true));
    switch (st.chosen_kind) {
        case SPMVB200_KIND_VECTOR: return launch_vector<OffT, ValT>(p, st.chosen_width);
        case SPMVB200_KIND_LIGHT: return launch_light<OffT, ValT>(p, st.chosen_width);
        case SPMVB200_KIND_STREAM: return launch_stream<OffT, ValT>(p);
        case SPMVB200_KIND_CUSPARSE: return launch_cusparse<OffT, ValT>(p);
        case SPMVB200_KIND_MERGE:
        default: return launch_merge<OffT, ValT>(p);
    }
}
template int launch_auto<int32_t, float>(const SpmvProblem<int32_t, float> &);
template int launch_auto<int32_t, double>(const SpmvProblem<int32_t, double> &);
template int launch_auto<int64_t, float>(const SpmvProblem<int64_t, float> &);
template int launch_auto<int64_t, double>(const SpmvProblem<int64_t, double> &);

// ------------------------------------------------------------------- cuSPARSE baseline
namespace {

struct CusparsePlan {
    int dev = -1;
    const void *Ap = nullptr, *Aj = nullptr, *Ax = nullptr;
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    int off_bits = 0, val_bits = 0;
    cusparseHandle_t handle = nullptr;
    cusparseSpMatDescr_t mat = nullptr;
    cusparseDnVecDescr_t vx = nullptr, vy = nullptr;
    void *buffer = nullptr;
    size_t buffer_bytes = 0;
    bool valid = false;
    bool preprocessed = false;
    int alg = 0;
    cudaStream_t stream = nullptr;  // the plan's work buffer belongs to one stream at a time
};
std::mutex g_cs_mu;
CusparsePlan g_plan;

void plan_destroy(CusparsePlan &pl) {
    if (pl.mat) cusparseDestroySpMat(pl.mat);
    if (pl.vx) cusparseDestroyDnVec(pl.vx);
    if (pl.vy) cusparseDestroyDnVec(pl.vy);
    if (pl.buffer) cudaFree(pl.buffer);
    if (pl.handle) cusparseDestroy(pl.handle);
    pl = CusparsePlan{};
}

#define SPMV_CUSPARSE_TRY(expr)                                  \
    do {                                                         \
        cusparseStatus_t _s = (expr);                            \
        if (_s != CUSPARSE_STATUS_SUCCESS) {                     \
            record_cuda_error(cudaErrorUnknown, #expr, cusparseGetErrorString(_s), __LINE__); \
            plan_destroy(g_plan);                                \
            return SPMVB200_ERR_CUSPARSE;                        \
        }                                                        \
    } while (0)

}  // namespace

void cusparse_plan_clear() {
    std::lock_guard<std::mutex> lk(g_cs_mu);
    plan_destroy(g_plan);
}

template <typename OffT, typename ValT>
int launch_cusparse(const SpmvProblem<OffT, ValT> &p) {
    if (p.n_rows <= 0 || p.n_cols <= 0) return SPMVB200_OK;
    // the baseline is plain y = A*x: no device alpha, no peer fan-out
    if (p.peers.n != 0 || p.alpha_dev) return SPMVB200_ERR_UNSUPPORTED;
    // cusparseCreateCsr rejects 64-bit offsets with 32-bit column indices: say so before any
    // handle or descriptor exists (the reference's wrapper maps `int` only, cusparse.cuh:23-26)
    if (sizeof(OffT) == 8) return SPMVB200_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lk(g_cs_mu);
    int dev = -1;
    SPMV_CUDA_TRY(cudaGetDevice(&dev));
    const cudaDataType vt = sizeof(ValT) == 4 ? CUDA_R_32F : CUDA_R_64F;
    const cusparseIndexType_t ot = sizeof(OffT) == 4 ? CUSPARSE_INDEX_32I : CUSPARSE_INDEX_64I;
    const ValT one = (ValT)1, zero = (ValT)0;
    CusparsePlan &pl = g_plan;
    // a preprocessed plan depends on the matrix CONTENTS, which a pointer key cannot see
    // (allocators hand the same addresses out again); it is only reused while the option
    // that asked for it is still on
    const bool want_pre = option_get("cusparse_preprocess", 0) > 0;
    // "cusparse_alg": 0 = CUSPARSE_SPMV_ALG_DEFAULT (what the reference calls, cusparse.cuh:76-78),
    // 1 = CUSPARSE_SPMV_CSR_ALG1, 2 = CUSPARSE_SPMV_CSR_ALG2
    const int want_alg = (int)option_get("cusparse_alg", 0);
    const cusparseSpMVAlg_t alg = want_alg == 1 ? CUSPARSE_SPMV_CSR_ALG1
                                : want_alg == 2 ? CUSPARSE_SPMV_CSR_ALG2 : CUSPARSE_SPMV_ALG_DEFAULT;
    // one plan = one handle + one work buffer: a call on another stream gets a fresh plan
    // (destroying the old one frees its buffer, which waits for the work that uses it)
    const bool hit = pl.valid && pl.preprocessed == want_pre && pl.alg == want_alg && pl.stream == p.stream && pl.dev == dev && pl.Ap == p.Ap && pl.Aj == p.Aj && pl.Ax == p.Ax &&
                     pl.n_rows == p.n_rows && pl.n_cols == p.n_cols && pl.nnz == (int64_t)p.nnz &&
                     pl.off_bits == (int)sizeof(OffT) * 8 && pl.val_bits == (int)sizeof(ValT) * 8;
    if (!hit) {
        plan_destroy(pl);
        pl.dev = dev;
        pl.Ap = p.Ap; pl.Aj = p.Aj; pl.Ax = p.Ax;
        pl.n_rows = p.n_rows; pl.n_cols = p.n_cols; pl.nnz = (int64_t)p.nnz;
        pl.off_bits = (int)sizeof(OffT) * 8; pl.val_bits = (int)sizeof(ValT) * 8;
        pl.alg = want_alg;
        pl.stream = p.stream;
        SPMV_CUSPARSE_TRY(cusparseCreate(&pl.handle));
        SPMV_CUSPARSE_TRY(cusparseCreateCsr(&pl.mat, p.n_rows, p.n_cols, (int64_t)p.nnz,
                                            const_cast<OffT *>(p.Ap), const_cast<int32_t *>(p.Aj),
                                            const_cast<ValT *>(p.Ax), ot, CUSPARSE_INDEX_32I,
                                            CUSPARSE_INDEX_BASE_ZERO, vt));
        SPMV_CUSPARSE_TRY(cusparseCreateDnVec(&pl.vx, p.n_cols, const_cast<ValT *>(p.x), vt));
        SPMV_CUSPARSE_TRY(cusparseCreateDnVec(&pl.vy, p.n_rows, p.y, vt));
        SPMV_CUSPARSE_TRY(cusparseSetStream(pl.handle, p.stream));
        SPMV_CUSPARSE_TRY(cusparseSpMV_bufferSize(pl.handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &one,
                                                  pl.mat, pl.vx, &zero, pl.vy, vt,
                                                  alg, &pl.buffer_bytes));
        SPMV_CUDA_TRY(cudaMalloc(&pl.buffer, pl.buffer_bytes ? pl.buffer_bytes : 16));
        pl.preprocessed = option_get("cusparse_preprocess", 0) > 0;
        if (pl.preprocessed)
            SPMV_CUSPARSE_TRY(cusparseSpMV_preprocess(pl.handle, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                                      &one, pl.mat, pl.vx, &zero, pl.vy, vt,
                                                      alg, pl.buffer));
        pl.valid = true;
    }
    SPMV_CUSPARSE_TRY(cusparseSetStream(pl.handle, p.stream));
    SPMV_CUSPARSE_TRY(cusparseDnVecSetValues(pl.vx, const_cast<ValT *>(p.x)));
    SPMV_CUSPARSE_TRY(cusparseDnVecSetValues(pl.vy, p.y));
    SPMV_CUSPARSE_TRY(cusparseSpMV(pl.handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, pl.mat, pl.vx,
                                   &zero, pl.vy, vt, alg, pl.buffer));
    return SPMVB200_OK;
}
template int launch_cusparse<int32_t, float>(const SpmvProblem<int32_t, float> &);
template int launch_cusparse<int32_t, double>(const SpmvProblem<int32_t, double> &);
template int launch_cusparse<int64_t, float>(const SpmvProblem<int64_t, float> &);
template int launch_cusparse<int64_t, double>(const SpmvProblem<int64_t, double> &);

}  // namespace spmvb200
