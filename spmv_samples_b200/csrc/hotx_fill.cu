// hotx_fill.cu -- the per-call half of the hot-x plan (hotx.cu builds it): x_hot[r] = x[hot_cols[r]].
//
// 32 us per SpMV on R-MAT scale 27 (8.4 M hot columns).  Kept apart from hotx.cu, which pulls in CUB
// for the plan build.
#include "common.cuh"

namespace spmvb200 {

namespace {

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
    // the caller's stream unless option "side_stream" is on), 3 = no refill at all (timing
    // experiments only: x_hot goes stale).  A sweep over x with a bitmap of the hot columns was
    // tried as well: 143 us on R-MAT scale 27 against 32 us for the gather (the hot columns are
    // 1 in 16, so the sweep reads 16 times what it keeps).
    const int64_t fill = option_get("hot_x_fill", 0);
    if (fill == 3) return SPMVB200_OK;
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
