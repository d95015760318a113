// merge.cu -- merge-path CSR SpMV for sm_100a: partition kernel, tile kernel, carry fixup.
//
// What it computes is what the reference's vendored CUB 1.15 merge-based SpMV computes
// (reference/include/spmv/merge_based/dispatch_spmv_orig.cuh:109-229,
//  agent_spmv_orig.cuh:454-760, agent_segment_fixup.cuh:229-358, thread_search.cuh:16-49):
// the (row-end, nonzero) merge path of length n_rows + nnz is cut into equal tiles, each
// tile reduces its nonzeros into the rows that end inside it and hands the unfinished
// tail to a fixup pass.  How it does it is new:
//   * a tile is 2048 path items (not 896/320) so that one CTA keeps 16-24 KB of Aj/Ax in
//     flight and 8 CTAs/SM cover the HBM latency-bandwidth product of a B200 SM;
//   * the tile's row offsets and its Aj / Ax segments are staged into shared memory with
//     TMA 1-D bulk copies (cp.async.bulk -> UBLKCP) completing on one mbarrier, with an L2
//     evict-first policy; x is gathered with an evict-last policy;
//   * only the row coordinate of each tile boundary is stored (int32); the nonzero
//     coordinate is diagonal - row, which also makes the scratch 64-bit safe for free;
//   * finished rows are stored straight from the merge loop; the cross-thread carry is a
//     warp-shuffle segmented scan; the tile carry-out goes to a fixup kernel that is
//     deterministic (run-head threads sum their run in tile order; no atomics, unlike
//     agent_segment_fixup.cuh:257,269).
#include <climits>

#include "common.cuh"

namespace spmvb200 {

namespace {

constexpr int kMergeBlock = 256;
constexpr int kMergeIPT = 8;
constexpr int kMergeTile = kMergeBlock * kMergeIPT;
constexpr int kPad = 8;  // slack for the 16-byte align-down shift and the sentinel

template <typename OffT, typename ValT>
constexpr size_t merge_smem_bytes() {
    return 16 + (size_t)(kMergeTile + kPad) * (sizeof(OffT) + sizeof(int32_t) + sizeof(ValT));
}

// ---------------------------------------------------------------- partition (search) kernel
// thread t: row coordinate of the merge path on diagonal min(t*tile_items, n_rows+nnz).
template <typename OffT>
__global__ void __launch_bounds__(256)
merge_partition_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap, int64_t tile_items,
                       int64_t n_coords, int32_t *__restrict__ coords_x) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_coords) return;
    const int64_t total = (int64_t)n_rows + (int64_t)nnz;
    int64_t diag = t * tile_items;
    if (diag > total) diag = total;
    const OffT *__restrict__ row_end = Ap + 1;
    int64_t lo = diag - (int64_t)nnz > 0 ? diag - (int64_t)nnz : 0;
    int64_t hi = diag < n_rows ? diag : n_rows;
    while (lo < hi) {
        const int64_t pivot = (lo + hi) >> 1;
        if ((int64_t)__ldg(row_end + pivot) <= diag - pivot - 1) lo = pivot + 1;
        else hi = pivot;
    }
    coords_x[t] = (int32_t)lo;
}

// stage count elements of g[gbeg ...) into s[(gbeg - a0) ...), a0 = gbeg aligned down to 16 B:
// the 16-byte aligned interior by one bulk copy (thread 0), the ragged tail by plain loads.
template <typename T>
struct StagePlan {
    int64_t a0;          // first element of the bulk copy (aligned)
    uint32_t bulk_bytes; // 0 = no bulk copy
    int64_t tail_beg;    // first element loaded by threads
    int tail_cnt;
    int shift;           // gbeg - a0
};
template <typename T>
__device__ __forceinline__ StagePlan<T> plan_stage(int64_t gbeg, int count) {
    constexpr int V = 16 / sizeof(T);
    StagePlan<T> p;
    p.a0 = gbeg & ~(int64_t)(V - 1);
    p.shift = (int)(gbeg - p.a0);
    const int64_t gend = gbeg + count;
    const int64_t be = gend & ~(int64_t)(V - 1);
    if (be > p.a0 && count > 0) {
        p.bulk_bytes = (uint32_t)((be - p.a0) * sizeof(T));
        p.tail_beg = be;
    } else {
        p.bulk_bytes = 0;
        p.tail_beg = gbeg;
    }
    p.tail_cnt = (int)(gend - p.tail_beg);
    return p;
}

// ---------------------------------------------------------------------------- tile kernel
template <typename OffT, typename ValT>
__global__ void __launch_bounds__(kMergeBlock)
merge_tile_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap,
                  const int32_t *__restrict__ Aj, const ValT *__restrict__ Ax,
                  const ValT *__restrict__ x, ValT *__restrict__ y,
                  const ValT *__restrict__ alpha_dev, PeerOut peers,
                  const int32_t *__restrict__ coords_x, int32_t *__restrict__ carry_row,
                  ValT *__restrict__ carry_val) {
    constexpr int TILE = kMergeTile;
    constexpr int IPT = kMergeIPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    OffT *s_rend = reinterpret_cast<OffT *>(smem_raw + 16);
    int32_t *s_col = reinterpret_cast<int32_t *>(s_rend + TILE + kPad);
    ValT *s_val = reinterpret_cast<ValT *>(s_col + TILE + kPad);
    __shared__ ValT s_wval[kMergeBlock / 32];
    __shared__ int s_wflag[kMergeBlock / 32];

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t total = (int64_t)n_rows + (int64_t)nnz;
    const int64_t d0 = tile * TILE;
    const int64_t d1 = d0 + TILE < total ? d0 + TILE : total;
    const int32_t sx = __ldg(coords_x + tile);
    const int32_t ex = __ldg(coords_x + tile + 1);
    const int64_t sy = d0 - sx;
    const int R = ex - sx;                     // rows that end inside this tile
    const int Z = (int)((d1 - ex) - sy);       // nonzeros inside this tile
    const int nr = R + (ex < n_rows ? 1 : 0);  // row ends staged (one past, for the open row)

    const StagePlan<OffT> pr = plan_stage<OffT>((int64_t)sx + 1, nr);
    const StagePlan<int32_t> pc = plan_stage<int32_t>(sy, Z);
    const StagePlan<ValT> pv = plan_stage<ValT>(sy, Z);

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint64_t pol = policy_evict_first();
        mbar_arrive_expect_tx(bar, pr.bulk_bytes + pc.bulk_bytes + pv.bulk_bytes);
        if (pr.bulk_bytes) bulk_g2s(s_rend, Ap + pr.a0, pr.bulk_bytes, bar, pol);
        if (pc.bulk_bytes) bulk_g2s(s_col, Aj + pc.a0, pc.bulk_bytes, bar, pol);
        if (pv.bulk_bytes) bulk_g2s(s_val, Ax + pv.a0, pv.bulk_bytes, bar, pol);
        // sentinel behind the staged row ends: "no further row ends here"
        s_rend[pr.shift + nr] = (OffT)(sizeof(OffT) == 8 ? LLONG_MAX : INT_MAX);
    }
    // ragged tails (fewer than one 16-byte vector each, or a whole tiny segment)
    if (tid < pr.tail_cnt) s_rend[pr.tail_beg - pr.a0 + tid] = __ldg(Ap + pr.tail_beg + tid);
    if (tid < pc.tail_cnt) s_col[pc.tail_beg - pc.a0 + tid] = __ldg(Aj + pc.tail_beg + tid);
    if (tid < pv.tail_cnt) s_val[pv.tail_beg - pv.a0 + tid] = __ldg(Ax + pv.tail_beg + tid);

    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;
    const uint64_t pol_x = policy_evict_last();

    mbar_wait(bar, 0);
    __syncthreads();

    // ---- products: s_val[i] *= x[s_col[i]], block-strided (bank-conflict free), all
    // gathers of a thread issued before the first use
    {
        const int32_t *cc = s_col + pc.shift;
        ValT *vv = s_val + pv.shift;
        ValT xv[IPT];
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int i = tid + k * kMergeBlock;
            xv[k] = (i < Z) ? ldg_hint(x + cc[i], pol_x) : (ValT)0;
        }
#pragma unroll
        for (int k = 0; k < IPT; ++k) {
            const int i = tid + k * kMergeBlock;
            if (i < Z) vv[i] *= xv[k];
        }
    }
    __syncthreads();

    // ---- per-thread merge-path search inside the tile (diagonal tid*IPT)
    const OffT *rend = s_rend + pr.shift;
    const ValT *prod = s_val + pv.shift;
    const int items = R + Z;
    const int beg = min(tid * IPT, items);
    const int n_my = min(IPT, items - beg);
    const OffT syo = (OffT)sy;
    int lo = max(beg - Z, 0), hi = min(beg, R);
    while (lo < hi) {
        const int p = (lo + hi) >> 1;
        if (rend[p] <= syo + (OffT)(beg - p - 1)) lo = p + 1;
        else hi = p;
    }
    int tx = lo, ty = beg - lo;

    // ---- serial merge of this thread's items
    ValT run = (ValT)0, first_part = (ValT)0;
    int first_row = -1;
    OffT next_end = rend[tx];
#pragma unroll
    for (int k = 0; k < IPT; ++k) {
        if (k < n_my) {
            if (syo + (OffT)ty < next_end) {
                run += prod[ty];
                ++ty;
            } else {
                if (first_row < 0) {
                    first_row = tx;
                    first_part = run;
                } else {
                    store_y(y, peers, (int64_t)sx + tx, alpha * run);
                }
                run = (ValT)0;
                ++tx;
                next_end = rend[tx];
            }
        }
    }

    // ---- segmented scan of (thread saw a row end, tail sum) across the block
    const int lane = tid & 31, warp = tid >> 5;
    int flag = first_row >= 0;
    ValT val = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const ValT pvv = __shfl_up_sync(0xffffffffu, val, d);
        const int pf = __shfl_up_sync(0xffffffffu, flag, d);
        if (lane >= d) {
            if (!flag) val += pvv;
            flag |= pf;
        }
    }
    if (lane == 31) {
        s_wval[warp] = val;
        s_wflag[warp] = flag;
    }
    ValT ev = __shfl_up_sync(0xffffffffu, val, 1);
    int ef = __shfl_up_sync(0xffffffffu, flag, 1);
    if (lane == 0) {
        ev = (ValT)0;
        ef = 0;
    }
    __syncthreads();
    ValT wv = (ValT)0;
    int wf = 0;
#pragma unroll
    for (int w = 0; w < kMergeBlock / 32; ++w) {
        if (w < warp) {
            const int f = s_wflag[w];
            const ValT v = s_wval[w];
            wv = f ? v : wv + v;
            wf |= f;
        }
    }
    const ValT carry_in = ef ? ev : wv + ev;
    if (first_row >= 0) store_y(y, peers, (int64_t)sx + first_row, alpha * (first_part + carry_in));

    if (tid == kMergeBlock - 1) {
        // inclusive over the whole block = tail after the last row end of the tile
        const ValT tot = flag ? val : wv + val;
        carry_row[tile] = ex;
        carry_val[tile] = tot;
    }
}

// ------------------------------------------------------------------------- carry fixup
// One thread per tile.  Consecutive tiles whose carry lands in the same row form a run; the
// head of the run adds the run's carries, in tile order, to y[row].
template <typename ValT>
__global__ void __launch_bounds__(256)
merge_fixup_kernel(int32_t n_rows, int64_t num_tiles, const int32_t *__restrict__ carry_row,
                   const ValT *__restrict__ carry_val, ValT *__restrict__ y,
                   const ValT *__restrict__ alpha_dev, PeerOut peers) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    const int32_t row = carry_row[t];
    if (row >= n_rows) return;
    if (t > 0 && carry_row[t - 1] == row) return;
    ValT sum = carry_val[t];
    for (int64_t u = t + 1; u < num_tiles && carry_row[u] == row; ++u) sum += carry_val[u];
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;
    store_y(y, peers, (int64_t)row, y[row] + alpha * sum);
}

}  // namespace

template <typename OffT>
int launch_partition(int32_t n_rows, OffT nnz, const OffT *Ap, int64_t tile_items, int64_t n_coords,
                     int32_t *coords_x, cudaStream_t stream) {
    if (n_coords <= 0) return SPMVB200_OK;
    const int64_t blocks = (n_coords + 255) / 256;
    merge_partition_kernel<OffT><<<(unsigned)blocks, 256, 0, stream>>>(n_rows, nnz, Ap, tile_items,
                                                                       n_coords, coords_x);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}
template int launch_partition<int32_t>(int32_t, int32_t, const int32_t *, int64_t, int64_t,
                                       int32_t *, cudaStream_t);
template int launch_partition<int64_t>(int32_t, int64_t, const int64_t *, int64_t, int64_t,
                                       int32_t *, cudaStream_t);

int64_t merge_tile_items() { return kMergeTile; }

template <typename OffT, typename ValT>
int launch_merge(const SpmvProblem<OffT, ValT> &p) {
    const int64_t total = (int64_t)p.n_rows + (int64_t)p.nnz;
    const int64_t num_tiles = (total + kMergeTile - 1) / kMergeTile;
    if (num_tiles <= 0) return SPMVB200_OK;
    if (num_tiles > 0x7fffffffLL) return SPMVB200_ERR_UNSUPPORTED;

    void *coords = nullptr, *crow = nullptr, *cval = nullptr;
    SPMV_TRY(scratch_get(p.stream, SCRATCH_COORDS, (size_t)(num_tiles + 1) * sizeof(int32_t), &coords));
    SPMV_TRY(scratch_get(p.stream, SCRATCH_CARRY_ROW, (size_t)num_tiles * sizeof(int32_t), &crow));
    SPMV_TRY(scratch_get(p.stream, SCRATCH_CARRY_VAL, (size_t)num_tiles * sizeof(ValT), &cval));

    SPMV_TRY(launch_partition<OffT>(p.n_rows, p.nnz, p.Ap, kMergeTile, num_tiles + 1,
                                    static_cast<int32_t *>(coords), p.stream));

    static bool attr_set = false;  // per instantiation
    constexpr size_t smem = merge_smem_bytes<OffT, ValT>();
    if (!attr_set) {
        SPMV_CUDA_TRY(cudaFuncSetAttribute(merge_tile_kernel<OffT, ValT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    LaunchCfg lc;
    make_launch_cfg(lc, dim3((unsigned)num_tiles), dim3(kMergeBlock), smem, p.stream, p.x,
                    (size_t)p.n_cols * sizeof(ValT));
    SPMV_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, merge_tile_kernel<OffT, ValT>, p.n_rows, p.nnz, p.Ap,
                                     p.Aj, p.Ax, p.x, p.y, p.alpha_dev, p.peers,
                                     (const int32_t *)coords, static_cast<int32_t *>(crow),
                                     static_cast<ValT *>(cval)));
    SPMV_LAUNCH_CHECK();

    if (num_tiles > 1) {
        const int64_t blocks = (num_tiles + 255) / 256;
        merge_fixup_kernel<ValT><<<(unsigned)blocks, 256, 0, p.stream>>>(
            p.n_rows, num_tiles, (const int32_t *)crow, (const ValT *)cval, p.y, p.alpha_dev, p.peers);
        SPMV_LAUNCH_CHECK();
    }
    return SPMVB200_OK;
}

template int launch_merge<int32_t, float>(const SpmvProblem<int32_t, float> &);
template int launch_merge<int32_t, double>(const SpmvProblem<int32_t, double> &);
template int launch_merge<int64_t, float>(const SpmvProblem<int64_t, float> &);
template int launch_merge<int64_t, double>(const SpmvProblem<int64_t, double> &);

}  // namespace spmvb200
