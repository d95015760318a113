// vector.cu -- CSR-vector SpMV for sm_100a: T lanes per row, 128-bit loads, shuffle reduce.
//
// Covers what the three CUSP kinds of the reference cover
// (reference/include/spmv/cusp/cusp.cuh:23-236, cusp_warp_reduce.cuh:15-147,
//  cusp_warp_read_reduce.cuh:15-153): one sub-warp per row, width from nnz/n_rows.
// Differences by design: every width reduces with shuffles (the reference falls back to a
// shared-memory tree for T < 32); each lane covers 4 nonzeros per step, so the width table
// is T = ceil_pow2(mean/4) in {1..32} instead of {2..32} on mean; 64-bit safe thread ids
// (SURVEY.md A.2).
#include "common.cuh"
#include "row_dot.cuh"

namespace spmvb200 {

namespace {

constexpr int kVecBlock = 256;

template <int T, typename OffT, typename ValT>
__global__ void __launch_bounds__(kVecBlock)
vector_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap,
              const int32_t *__restrict__ Aj, const ValT *__restrict__ Ax,
              const ValT *__restrict__ x, ValT *__restrict__ y,
              const ValT *__restrict__ alpha_dev, PeerOut peers) {
    const int64_t gtid = (int64_t)blockIdx.x * kVecBlock + threadIdx.x;
    const int64_t row = gtid / T;
    const int lane = threadIdx.x & (T - 1);
    const bool active = row < n_rows;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_x = policy_evict_last();
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;
    ValT sum = (ValT)0;
    OffT s = 0, e = 0;
    if (active) {
        s = __ldg(Ap + row);
        e = __ldg(Ap + row + 1);
    }
    // rows far longer than the sub-warp is wide go to the whole warp afterwards
    const bool is_long = row_is_long<T, OffT>(e - s);
    if (active && !is_long)
        sum = row_partial<T, OffT, ValT>(s, e, nnz, lane, Aj, Ax, x, pol_stream, pol_x);
    sum = subwarp_sum<T>(sum);
    if (active && !is_long && lane == 0) store_y_nonempty(y, peers, row, alpha * sum, e > s);
    warp_long_rows<T, OffT, ValT>(is_long, s, e, row, nnz, Aj, Ax, x, y, peers, alpha, pol_stream,
                                  pol_x);
}

template <int T, typename OffT, typename ValT>
int launch_T(const SpmvProblem<OffT, ValT> &p) {
    const int64_t threads = (int64_t)p.n_rows * T;
    const int64_t blocks = (threads + kVecBlock - 1) / kVecBlock;
    if (blocks > 0x7fffffffLL) return SPMVB200_ERR_UNSUPPORTED;
    LaunchCfg lc;
    make_launch_cfg(lc, dim3((unsigned)blocks), dim3(kVecBlock), 0, p.stream, p.x,
                    (size_t)p.n_cols * sizeof(ValT));
    {
        KernelTimerScope timed(p.stream);
        SPMV_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, vector_kernel<T, OffT, ValT>, p.n_rows, p.nnz,
                                         p.Ap, p.Aj, p.Ax, p.x, p.y, p.alpha_dev, p.peers));
    }
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

}  // namespace

// lanes per row from the mean row length: each lane covers 4 nonzeros per step
int pick_width_from_mean(double mean_row_len) {
    int t = 1;
    while (t < 32 && 4.0 * t < mean_row_len) t <<= 1;
    return t;
}

template <typename OffT, typename ValT>
int launch_vector(const SpmvProblem<OffT, ValT> &p, int width) {
    if (p.n_rows <= 0) return SPMVB200_OK;
    if (width <= 0) {
        width = (int)option_get("vector_width", 0);
        if (width <= 0) width = pick_width_from_mean((double)p.nnz / (double)p.n_rows);
    }
    switch (width) {
        case 1: return launch_T<1>(p);
        case 2: return launch_T<2>(p);
        case 4: return launch_T<4>(p);
        case 8: return launch_T<8>(p);
        case 16: return launch_T<16>(p);
        case 32: return launch_T<32>(p);
        default: return SPMVB200_ERR_INVALID;
    }
}

template int launch_vector<int32_t, float>(const SpmvProblem<int32_t, float> &, int);
template int launch_vector<int32_t, double>(const SpmvProblem<int32_t, double> &, int);
template int launch_vector<int64_t, float>(const SpmvProblem<int64_t, float> &, int);
template int launch_vector<int64_t, double>(const SpmvProblem<int64_t, double> &, int);

}  // namespace spmvb200
