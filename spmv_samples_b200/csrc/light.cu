// light.cu -- LightSpMV-style dynamic row distribution for sm_100a.
//
// Covers reference/include/spmv/LightSpMV.cuh:114-263 (csrDynamicVector / csrDynamicWarp):
// a persistent grid whose warps claim rows from one global counter with atomicAdd, so that
// warps that drew long rows simply claim fewer.  Differences by design:
//   * a claim is a block of rows worth ~4K nonzeros, not 1 row (vector mode) or 32/T rows
//     (warp mode): B200's L2 serialises same-address atomics at roughly one per clock
//     (B300_MICROARCH.md, atomics table), so one atomic per row would cost more than the
//     SpMV itself at 16.7M rows;
//   * the grid is SMs x resident CTAs from the occupancy API -- the reference's launch has
//     grid and block transposed;
//   * a row far longer than its sub-warp is wide is reduced by the whole warp instead
//     (row_dot.cuh warp_long_rows), the warp-level half of the load balancing the reference
//     leaves entirely to the row counter;
//   * no texture object, no __constant__ row count; x goes through L2 with an evict-last
//     policy; the inner loop is the 128-bit one of row_dot.cuh.
#include "common.cuh"
#include "row_dot.cuh"

namespace spmvb200 {

namespace {

constexpr int kLightBlock = 256;

template <int T, typename OffT, typename ValT>
__global__ void __launch_bounds__(kLightBlock)
light_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap,
             const int32_t *__restrict__ Aj, const ValT *__restrict__ Ax,
             const ValT *__restrict__ x, ValT *__restrict__ y,
             const ValT *__restrict__ alpha_dev, PeerOut peers,
             unsigned long long *__restrict__ row_counter, int rows_per_claim) {
    constexpr int ROWS_PER_STEP = 32 / T;
    const int wlane = threadIdx.x & 31;
    const int lane = wlane & (T - 1);
    const int sub = wlane / T;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_x = policy_evict_last();
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;

    for (;;) {
        unsigned long long base = 0;
        if (wlane == 0) base = atomicAdd(row_counter, (unsigned long long)rows_per_claim);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned long long)n_rows) break;
        const int64_t limit = min((int64_t)base + rows_per_claim, (int64_t)n_rows);
        for (int64_t r0 = (int64_t)base; r0 < limit; r0 += ROWS_PER_STEP) {
            const int64_t row = r0 + sub;
            const bool active = row < limit;
            ValT sum = (ValT)0;
            OffT s = 0, e = 0;
            if (active) {
                s = __ldg(Ap + row);
                e = __ldg(Ap + row + 1);
            }
            const bool is_long = row_is_long<T, OffT>(e - s);
            if (active && !is_long)
                sum = row_partial<T, OffT, ValT>(s, e, nnz, lane, Aj, Ax, x, pol_stream, pol_x);
            sum = subwarp_sum<T>(sum);
            if (active && !is_long && lane == 0) store_y_nonempty(y, peers, row, alpha * sum, e > s);
            warp_long_rows<T, OffT, ValT>(is_long, s, e, row, nnz, Aj, Ax, x, y, peers, alpha,
                                          pol_stream, pol_x);
        }
    }
}

template <int T, typename OffT, typename ValT>
int launch_T(const SpmvProblem<OffT, ValT> &p, int rows_per_claim) {
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    static int blocks_per_sm = 0;  // per instantiation
    if (blocks_per_sm == 0) {
        SPMV_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &blocks_per_sm, light_kernel<T, OffT, ValT>, kLightBlock, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    void *counter = nullptr;
    SPMV_TRY(scratch_get(p.stream, SCRATCH_COUNTER, sizeof(unsigned long long), &counter));
    SPMV_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), p.stream));

    // never launch more warps than there are claims to make
    const int64_t claims = ((int64_t)p.n_rows + rows_per_claim - 1) / rows_per_claim;
    int64_t blocks = (int64_t)di->sm_count * blocks_per_sm;
    const int64_t need = (claims + (kLightBlock / 32) - 1) / (kLightBlock / 32);
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    LaunchCfg lc;
    make_launch_cfg(lc, dim3((unsigned)blocks), dim3(kLightBlock), 0, p.stream, p.x,
                    (size_t)p.n_cols * sizeof(ValT));
    {
        KernelTimerScope timed(p.stream);
        SPMV_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, light_kernel<T, OffT, ValT>, p.n_rows, p.nnz, p.Ap,
                                         p.Aj, p.Ax, p.x, p.y, p.alpha_dev, p.peers,
                                         static_cast<unsigned long long *>(counter), rows_per_claim));
    }
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

}  // namespace

template <typename OffT, typename ValT>
int launch_light(const SpmvProblem<OffT, ValT> &p, int width) {
    if (p.n_rows <= 0) return SPMVB200_OK;
    const double mean = (double)p.nnz / (double)p.n_rows;
    if (width <= 0) {
        width = (int)option_get("light_width", 0);
        if (width <= 0) width = pick_width_from_mean(mean);
    }
    int rpc = (int)option_get("light_rows_per_claim", 0);
    if (rpc <= 0) {
        // ~4096 nonzeros per claim, a whole number of sub-warp steps, at most 4096 rows
        const int step = 32 / width;
        int64_t r = (int64_t)(4096.0 / (mean > 1.0 ? mean : 1.0));
        r = (r + step - 1) / step * step;
        if (r < step) r = step;
        if (r > 4096) r = 4096;
        // a small matrix must still give every resident warp a few claims: 1M rows in claims of
        // 816 is 1285 claims for 9472 warps (the 1024^2 Laplacian ran at 78 us against 20 us for
        // the static CSR-vector kernel)
        const DeviceInfo *di = nullptr;
        SPMV_TRY(current_device_info(&di));
        const int64_t warps = (int64_t)di->sm_count * (di->max_threads_per_sm / 32);
        int64_t cap = (int64_t)p.n_rows / (4 * warps);
        cap = (cap + step - 1) / step * step;
        if (cap < step) cap = step;
        if (r > cap) r = cap;
        rpc = (int)r;
    }
    switch (width) {
        case 1: return launch_T<1>(p, rpc);
        case 2: return launch_T<2>(p, rpc);
        case 4: return launch_T<4>(p, rpc);
        case 8: return launch_T<8>(p, rpc);
        case 16: return launch_T<16>(p, rpc);
        case 32: return launch_T<32>(p, rpc);
        default: return SPMVB200_ERR_INVALID;
    }
}

template int launch_light<int32_t, float>(const SpmvProblem<int32_t, float> &, int);
template int launch_light<int32_t, double>(const SpmvProblem<int32_t, double> &, int);
template int launch_light<int64_t, float>(const SpmvProblem<int64_t, float> &, int);
template int launch_light<int64_t, double>(const SpmvProblem<int64_t, double> &, int);

}  // namespace spmvb200
