// hotx_fill.cu -- the per-call half of the hot-x plan (hotx.cu builds it): x_hot[r] = x[hot_cols[r]].
//
// 32 us per SpMV on R-MAT scale 27 (8.4 M hot columns).  Kept apart from hotx.cu, which pulls in CUB
// for the plan build.
#include "common.cuh"

namespace spmvb200 {

namespace {

// x_hot by compaction (option hot_x_fill = 2, kept for the comparison): a warp reads 32
// consecutive values of x wherever at least one of them is hot and writes the hot ones to
// consecutive slots.  143 us on R-MAT scale 27 against 32 us for the gather: the hot columns are
// 1 in 16, so the sweep reads 16 times what it keeps.
template <typename ValT>
__global__ void __launch_bounds__(256)
hot_compact_kernel(const ValT *__restrict__ x, const uint32_t *__restrict__ bitmap,
                   const uint32_t *__restrict__ rank32, int64_t words, int64_t n_cols, ValT *__restrict__ x_hot) {
    const int lane = threadIdx.x & 31;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < words;
         w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const unsigned mask = __ldg(bitmap + w);
        if (mask == 0u) continue;
        const int64_t c = w * 32 + lane;
        if ((mask >> lane) & 1u)
            x_hot[__ldg(rank32 + w) + __popc(mask & ((1u << lane) - 1u))] = __ldg(x + c);
    }
}

template <typename ValT>
__global__ void __launch_bounds__(256)
hot_gather_kernel(const ValT *__restrict__ x, const int32_t *__restrict__ hot_cols, int64_t K,
                  ValT *__restrict__ x_hot) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < K; r += (int64_t)gridDim.x * blockDim.x)
        x_hot[r] = __ldg(x + __ldg(hot_cols + r));
}

}  // namespace

template <typename ValT>
int hot_gather(const HotPlan &plan, const ValT *x, cudaStream_t stream, const ValT **x_hot) {
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    void *buf = nullptr;
    SPMV_TRY(scratch_get(stream, SCRATCH_XHOT, (size_t)plan.K * sizeof(ValT), &buf));
    *x_hot = static_cast<const ValT *>(buf);
    const int64_t cap = (int64_t)di->sm_count * 16;
    // "hot_x_fill": 0 / 1 = gather x[hot_cols[r]] (0 goes through side_fork / side_join, which is
    // the caller's stream unless option "side_stream" is on), 2 = sweep over x with the bitmap,
    // 3 = no refill at all (timing experiments only: x_hot goes stale)
    const int64_t fill = option_get("hot_x_fill", 0);
    if (fill == 3) return SPMVB200_OK;
    if (fill == 2) {
        const int64_t words = ((int64_t)plan.n_cols + 31) / 32;
        int64_t blocks = (words * 32 + 255) / 256;
        if (blocks > cap) blocks = cap;
        hot_compact_kernel<ValT><<<(unsigned)blocks, 256, 0, stream>>>(x, plan.bitmap, plan.rank32, words,
                                                                       plan.n_cols, static_cast<ValT *>(buf));
        SPMV_LAUNCH_CHECK();
        return SPMVB200_OK;
    }
    int64_t blocks = (plan.K + 255) / 256;
    if (blocks > cap) blocks = cap;
    if (fill == 1) {
        hot_gather_kernel<ValT><<<(unsigned)blocks, 256, 0, stream>>>(x, plan.hot_cols, plan.K,
                                                                      static_cast<ValT *>(buf));
        SPMV_LAUNCH_CHECK();
        return SPMVB200_OK;
    }
    cudaStream_t side = nullptr;
    SPMV_TRY(side_fork(stream, &side));
    hot_gather_kernel<ValT><<<(unsigned)blocks, 256, 0, side>>>(x, plan.hot_cols, plan.K, static_cast<ValT *>(buf));
    SPMV_LAUNCH_CHECK();
    return side_join(stream);
}
template int hot_gather<float>(const HotPlan &, const float *, cudaStream_t, const float **);
template int hot_gather<double>(const HotPlan &, const double *, cudaStream_t, const double **);

}  // namespace spmvb200
