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
//   * a block of rows that holds far more than a claim's worth of nonzeros (R-MAT's first 256
//     rows are 25M nonzeros: 21 ms for the one warp that drew them) is not worked off by the
//     claiming warp: a classify pass lists such blocks first, the main kernel hands their rows
//     out one ticket at a time to whole warps once the ordinary blocks are gone, and rows too
//     long even for a warp go to a third pass, a whole CTA per row.  Three counters instead of
//     one; still nothing but atomics on global counters, and every row is summed by one
//     sub-warp, one warp or one CTA in a fixed order whoever draws it;
//   * no texture object, no __constant__ row count; x goes through L2 with an evict-last
//     policy; the inner loop is the 128-bit one of row_dot.cuh.
#include "common.cuh"
#include "row_dot.cuh"

namespace spmvb200 {

namespace {

constexpr int kLightBlock = 256;
constexpr int kLightMegaBlock = 512;
constexpr long long kLightHeavyNnz = 16384;  // a claim with more nonzeros than this is "heavy"
constexpr long long kLightMegaRow = 8192;    // a row of a heavy block longer than this goes to a CTA
constexpr int kLightRowsPerTicket = 4;

struct LightCtl {                    // zeroed before every call
    unsigned long long row_counter;  // tier 1: next unclaimed row
    unsigned long long heavy_ticket; // tier 2: next (heavy block, row) pair
    unsigned int heavy_count;        // heavy blocks listed by the classify pass
    unsigned int mega_count;         // rows listed by tier 2 for tier 3
    unsigned int mega_ticket;        // tier 3: next listed row
    unsigned int pad;
};

template <typename OffT>
__device__ __forceinline__ long long block_nnz(const OffT *__restrict__ Ap, int64_t base, int rows_per_claim,
                                               int32_t n_rows) {
    const int64_t lim = base + rows_per_claim < (int64_t)n_rows ? base + rows_per_claim : (int64_t)n_rows;
    return (long long)__ldg(Ap + lim) - (long long)__ldg(Ap + base);
}

// one thread per block of rows_per_claim rows: list the heavy ones (any order)
template <typename OffT>
__global__ void __launch_bounds__(256)
light_classify_kernel(int32_t n_rows, const OffT *__restrict__ Ap, int rows_per_claim, int64_t n_blocks,
                      LightCtl *__restrict__ ctl, int32_t *__restrict__ heavy_base) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const int64_t base = b * rows_per_claim;
    if (block_nnz<OffT>(Ap, base, rows_per_claim, n_rows) > kLightHeavyNnz)
        heavy_base[atomicAdd(&ctl->heavy_count, 1u)] = (int32_t)base;
}

template <int T, typename OffT, typename ValT>
__global__ void __launch_bounds__(kLightBlock)
light_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap,
             const int32_t *__restrict__ Aj, const ValT *__restrict__ Ax,
             const ValT *__restrict__ x, ValT *__restrict__ y,
             const ValT *__restrict__ alpha_dev, PeerOut peers, LightCtl *__restrict__ ctl,
             const int32_t *__restrict__ heavy_base, int32_t *__restrict__ mega_row, int rows_per_claim) {
    constexpr int ROWS_PER_STEP = 32 / T;
    const int wlane = threadIdx.x & 31;
    const int lane = wlane & (T - 1);
    const int sub = wlane / T;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_x = policy_evict_last();
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;

    // ---- tier 1: blocks of rows, a sub-warp per row; heavy blocks are left for tier 2
    for (;;) {
        unsigned long long base = 0;
        int heavy = 0;
        if (wlane == 0) {
            base = atomicAdd(&ctl->row_counter, (unsigned long long)rows_per_claim);
            if (base < (unsigned long long)n_rows)
                heavy = block_nnz<OffT>(Ap, (int64_t)base, rows_per_claim, n_rows) > kLightHeavyNnz;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        heavy = __shfl_sync(0xffffffffu, heavy, 0);
        if (base >= (unsigned long long)n_rows) break;
        if (heavy) continue;
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

    // ---- tier 2: the rows of the heavy blocks, kLightRowsPerTicket per ticket, a warp per row
    const unsigned long long total = (unsigned long long)ctl->heavy_count * (unsigned long long)rows_per_claim;
    if (total == 0) return;
    for (;;) {
        unsigned long long t = 0;
        if (wlane == 0) t = atomicAdd(&ctl->heavy_ticket, (unsigned long long)kLightRowsPerTicket);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= total) break;
        for (int j = 0; j < kLightRowsPerTicket; ++j) {
            const unsigned long long idx = t + j;
            if (idx >= total) break;
            const int64_t row = (int64_t)heavy_base[idx / rows_per_claim] + (int64_t)(idx % rows_per_claim);
            if (row >= (int64_t)n_rows) continue;  // the last block of the matrix may be short
            const OffT s = __ldg(Ap + row);
            const OffT e = __ldg(Ap + row + 1);
            if ((long long)(e - s) > kLightMegaRow) {
                if (wlane == 0) mega_row[atomicAdd(&ctl->mega_count, 1u)] = (int32_t)row;
                continue;
            }
            ValT ps = row_partial<32, OffT, ValT>(s, e, nnz, wlane, Aj, Ax, x, pol_stream, pol_x);
            ps = subwarp_sum<32>(ps);
            if (wlane == 0) store_y_nonempty(y, peers, row, alpha * ps, e > s);
        }
    }
}

// ---- tier 3: the rows tier 2 listed, a whole CTA per row
template <typename OffT, typename ValT>
__global__ void __launch_bounds__(kLightMegaBlock)
light_mega_kernel(OffT nnz, const OffT *__restrict__ Ap, const int32_t *__restrict__ Aj,
                  const ValT *__restrict__ Ax, const ValT *__restrict__ x, ValT *__restrict__ y,
                  const ValT *__restrict__ alpha_dev, PeerOut peers, LightCtl *__restrict__ ctl,
                  const int32_t *__restrict__ mega_row) {
    const unsigned int count = ctl->mega_count;
    if (count == 0) return;
    __shared__ long long s_row;
    __shared__ ValT s_red[kLightMegaBlock / 32];
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_x = policy_evict_last();
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int m = atomicAdd(&ctl->mega_ticket, 1u);
            s_row = m < count ? (long long)mega_row[m] : -1;
        }
        __syncthreads();
        const long long row = s_row;
        if (row < 0) break;
        const OffT s = __ldg(Ap + row);
        const OffT e = __ldg(Ap + row + 1);
        ValT ps = row_partial<kLightMegaBlock, OffT, ValT>(s, e, nnz, (int)threadIdx.x, Aj, Ax, x, pol_stream, pol_x);
        ps = subwarp_sum<32>(ps);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ps;
        __syncthreads();
        if (threadIdx.x == 0) {
            ValT tot = (ValT)0;
#pragma unroll
            for (int w = 0; w < kLightMegaBlock / 32; ++w) tot += s_red[w];
            store_y_nonempty(y, peers, (int64_t)row, alpha * tot, true);
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
    // control block (zeroed per call) + the two lists (written before they are read).  Heavy
    // blocks are disjoint with more than kLightHeavyNnz nonzeros each, listed rows have more than
    // kLightMegaRow: both lists are bounded by nnz.
    const size_t heavy_cap = (size_t)((long long)p.nnz / kLightHeavyNnz) + 1;
    const size_t mega_cap = (size_t)((long long)p.nnz / kLightMegaRow) + 1;
    const size_t off_heavy = 64, off_mega = off_heavy + heavy_cap * sizeof(int32_t);
    void *buf = nullptr;
    SPMV_TRY(scratch_get(p.stream, SCRATCH_COUNTER, off_mega + mega_cap * sizeof(int32_t), &buf));
    SPMV_CUDA_TRY(cudaMemsetAsync(buf, 0, 64, p.stream));
    char *base = static_cast<char *>(buf);
    LightCtl *ctl = reinterpret_cast<LightCtl *>(base);
    int32_t *heavy_base = reinterpret_cast<int32_t *>(base + off_heavy);
    int32_t *mega_row = reinterpret_cast<int32_t *>(base + off_mega);

    const int64_t claims = ((int64_t)p.n_rows + rows_per_claim - 1) / rows_per_claim;
    light_classify_kernel<OffT><<<(unsigned)((claims + 255) / 256), 256, 0, p.stream>>>(
        p.n_rows, p.Ap, rows_per_claim, claims, ctl, heavy_base);
    SPMV_LAUNCH_CHECK();

    // never launch more warps than there are claims to make
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
                                         p.Aj, p.Ax, p.x, p.y, p.alpha_dev, p.peers, ctl,
                                         (const int32_t *)heavy_base, mega_row, rows_per_claim));
    }
    SPMV_LAUNCH_CHECK();
    light_mega_kernel<OffT, ValT><<<(unsigned)(di->sm_count * 2), kLightMegaBlock, 0, p.stream>>>(
        p.nnz, p.Ap, p.Aj, p.Ax, p.x, p.y, p.alpha_dev, p.peers, ctl, (const int32_t *)mega_row);
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
