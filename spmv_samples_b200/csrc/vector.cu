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

// Ablation variant (option "vector_rows_per_subwarp" = 4): RPS rows per sub-warp, interleaved
// so that all their loads are in flight at once.  Sub-warp s of the CTA takes rows
// base + k * (256/T) + s, k = 0..RPS-1, so consecutive sub-warps still read consecutive rows.
// The hypothesis was that one row per sub-warp is latency bound on very short rows (a chain of
// dependent round trips Ap -> Aj/Ax -> x -> y with 32 bytes per lane in flight); the
// measurement did not bear it out, see launch_vector.
template <int T, int RPS, typename OffT, typename ValT>
__global__ void __launch_bounds__(kVecBlock)
vector_multi_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap,
                    const int32_t *__restrict__ Aj, const ValT *__restrict__ Ax,
                    const ValT *__restrict__ x, ValT *__restrict__ y,
                    const ValT *__restrict__ alpha_dev, PeerOut peers) {
    constexpr int SW = kVecBlock / T;  // sub-warps per CTA
    const int lane = threadIdx.x & (T - 1);
    const int sw = threadIdx.x / T;
    const int64_t row0 = (int64_t)blockIdx.x * (SW * RPS) + sw;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_x = policy_evict_last();
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;

    OffT s[RPS], e[RPS];
    bool active[RPS], is_long[RPS];
#pragma unroll
    for (int k = 0; k < RPS; ++k) {
        const int64_t row = row0 + (int64_t)k * SW;
        active[k] = row < n_rows;
        s[k] = active[k] ? __ldg(Ap + row) : (OffT)0;
        e[k] = active[k] ? __ldg(Ap + row + 1) : (OffT)0;
    }
    // first chunk of every row: all Aj/Ax loads, then all gathers, then the arithmetic
    Chunk<ValT> ch[RPS];
#pragma unroll
    for (int k = 0; k < RPS; ++k) {
        is_long[k] = row_is_long<T, OffT>(e[k] - s[k]);
        const OffT a = s[k] & ~(OffT)3;
        ch[k] = fetch_chunk<OffT, ValT>(a + (OffT)(4 * lane), s[k], (active[k] && !is_long[k]) ? e[k] : s[k],
                                        nnz, Aj, Ax, pol_stream);
    }
    ValT xv[RPS][4];
#pragma unroll
    for (int k = 0; k < RPS; ++k) {
        xv[k][0] = (ch[k].mask & 1u) ? ldg_hint(x + ch[k].c.x, pol_x) : (ValT)0;
        xv[k][1] = (ch[k].mask & 2u) ? ldg_hint(x + ch[k].c.y, pol_x) : (ValT)0;
        xv[k][2] = (ch[k].mask & 4u) ? ldg_hint(x + ch[k].c.z, pol_x) : (ValT)0;
        xv[k][3] = (ch[k].mask & 8u) ? ldg_hint(x + ch[k].c.w, pol_x) : (ValT)0;
    }
    ValT sum[RPS];
#pragma unroll
    for (int k = 0; k < RPS; ++k) {
        ValT t = (ValT)0;
        if (ch[k].mask & 1u) t += ch[k].v.x * xv[k][0];
        if (ch[k].mask & 2u) t += ch[k].v.y * xv[k][1];
        if (ch[k].mask & 4u) t += ch[k].v.z * xv[k][2];
        if (ch[k].mask & 8u) t += ch[k].v.w * xv[k][3];
        sum[k] = t;
    }
    // whatever is left of rows longer than 4*T (uncommon when T was picked from the mean)
#pragma unroll
    for (int k = 0; k < RPS; ++k) {
        if (active[k] && !is_long[k]) {
            const OffT a = s[k] & ~(OffT)3;
            for (OffT p = a + (OffT)(4 * (lane + T)); p < e[k]; p += (OffT)(4 * T)) {
                const Chunk<ValT> c0 = fetch_chunk<OffT, ValT>(p, s[k], e[k], nnz, Aj, Ax, pol_stream);
                sum[k] = consume_chunk<ValT>(c0, x, pol_x, sum[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < RPS; ++k) {
        const int64_t row = row0 + (int64_t)k * SW;
        const ValT tot = subwarp_sum<T>(sum[k]);
        if (active[k] && !is_long[k] && lane == 0)
            store_y_nonempty(y, peers, row, alpha * tot, e[k] > s[k]);
        warp_long_rows<T, OffT, ValT>(active[k] && is_long[k], s[k], e[k], row, nnz, Aj, Ax, x, y, peers,
                                      alpha, pol_stream, pol_x);
    }
}

template <int T, int RPS, typename OffT, typename ValT>
int launch_multi(const SpmvProblem<OffT, ValT> &p) {
    constexpr int rows_per_block = (kVecBlock / T) * RPS;
    const int64_t blocks = ((int64_t)p.n_rows + rows_per_block - 1) / rows_per_block;
    if (blocks > 0x7fffffffLL) return SPMVB200_ERR_UNSUPPORTED;
    LaunchCfg lc;
    make_launch_cfg(lc, dim3((unsigned)blocks), dim3(kVecBlock), 0, p.stream, p.x,
                    (size_t)p.n_cols * sizeof(ValT));
    {
        KernelTimerScope timed(p.stream);
        SPMV_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, vector_multi_kernel<T, RPS, OffT, ValT>, p.n_rows,
                                         p.nnz, p.Ap, p.Aj, p.Ax, p.x, p.y, p.alpha_dev, p.peers));
    }
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
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
    // rows per sub-warp: "vector_rows_per_subwarp" = 4 selects the interleaved kernel.  Measured:
    // no gain on the Laplacian (22.5 us either way -- a 54 MB problem is bounded by launch and
    // ramp-up, a plain 27 MB device copy takes about as long) and a loss on uniform 16/row
    // (276 -> 315 us, registers), so one row per sub-warp stays the default.
    int64_t rps = option_get("vector_rows_per_subwarp", 0);
    if (rps <= 0) rps = 1;
    if (rps >= 4) {
        switch (width) {
            case 1: return launch_multi<1, 4>(p);
            case 2: return launch_multi<2, 4>(p);
            case 4: return launch_multi<4, 4>(p);
            case 8: return launch_multi<8, 4>(p);
            default: break;
        }
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
