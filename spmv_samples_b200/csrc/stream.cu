// stream.cu -- CSR-stream SpMV for sm_100a: persistent CTAs, tiles of rows staged in shared
// memory by TMA bulk copies, one thread per row.
//
// The kernel for matrices of short, regular rows (the 5-point Laplacian of BASELINE.json's first
// configuration; the selector routes mean <= 6 and max <= 64 nonzeros per row here).  It covers
// the case the reference gives to its CSR-vector kernels with 2- and 4-lane vectors
// (reference/include/spmv/cusp/cusp.cuh:23-142 with THREADS_PER_VECTOR = 2/4, :189-203): there a
// 5-nonzero row occupies 2 lanes x 4 load slots (3 of 8 wasted), every lane starts from an
// unaligned row offset, and each row is a chain of dependent round trips Ap -> Aj/Ax -> x -> y.
// Here nothing of the matrix is loaded by a load instruction:
//   * a CTA owns a contiguous range of 256-row tiles; one producer warp walks it and, per tile,
//     issues three cp.async.bulk copies (the tile's row offsets, column indices and values:
//     contiguous in CSR whatever the row lengths) into a ring of shared-memory stages, completion
//     on an mbarrier per stage -- the stream runs `kStages` tiles ahead of the arithmetic;
//   * eight consumer warps take a stage when its barrier flips: thread t owns row t of the tile,
//     reads its offsets, then its nonzeros from shared memory in order, gathers x through L1
//     (rows of a banded matrix touch neighbouring x entries, so a warp's gathers coalesce) and
//     stores y coalesced; the stage goes back to the producer through an "empty" barrier;
//   * a tile holding more nonzeros than a stage (rows far longer than the mean) is streamed in
//     several chunks, the row sums living in registers in between, so any CSR matrix is handled
//     correctly -- just not quickly if its rows are long (that is what the other kinds are for).
// The sum of a row is formed in the order of its nonzeros, like the reference's CPU loop
// (reference/include/spmv/cpu_navie.hpp:9-16).
#include "common.cuh"

namespace spmvb200 {

namespace {

#ifndef SPMV_STREAM_ROWS
#define SPMV_STREAM_ROWS 256
#endif
#ifndef SPMV_STREAM_CAP
#define SPMV_STREAM_CAP 2048
#endif
#ifndef SPMV_STREAM_STAGES
#define SPMV_STREAM_STAGES 3
#endif
// ablation builds only (tools/build_variants.sh): 1 = no x gathers, 2 = consumers only take and
// release stages (the TMA stream alone), 3 = as 2 and no row offsets staged
#ifndef SPMV_STREAM_ABLATE
#define SPMV_STREAM_ABLATE 0
#endif
constexpr int kRows = SPMV_STREAM_ROWS;      // rows per tile = consumer threads
constexpr int kConsumerWarps = kRows / 32;
constexpr int kStreamBlock = kRows + 32;     // + one producer warp
constexpr int kCap = SPMV_STREAM_CAP;        // nonzeros per stage
constexpr int kStages = SPMV_STREAM_STAGES;
#ifndef SPMV_STREAM_ALIGN
#define SPMV_STREAM_ALIGN 32
#endif
constexpr int kAlign = SPMV_STREAM_ALIGN;    // a stage's Aj / Ax copies start at a multiple of this many
                                             // nonzeros: 32 = 128 bytes of Aj (16-byte starts ran the
                                             // TMA stream at 2.5 TB/s on the Laplacian)
static_assert(kAlign % 4 == 0 && kCap % kAlign == 0, "stage geometry");

struct alignas(16) StageMeta {
    long long abase;      // nonzero index held by slot 0 of the stage's Aj / Ax buffers (kAlign-aligned)
    long long lo, hi;     // the stage holds the tile's nonzeros [lo, hi)
    int row0;             // first row of the tile
    int rows;             // rows in the tile (0 = end of this CTA's work)
    int first, last;      // first / last chunk of its tile
};

template <typename OffT, typename ValT>
struct StreamSmem {
    // every staged array starts on a 128-byte boundary of shared memory (and of global memory,
    // see kAlign): the bulk-copy engine moves whole 128-byte lines that way
    static constexpr size_t off_bytes = ((size_t)(kRows + 1) * sizeof(OffT) + 127) / 128 * 128;
    static constexpr size_t col_bytes = (size_t)kCap * sizeof(int32_t);
    static constexpr size_t val_bytes = (size_t)kCap * sizeof(ValT);
    static constexpr size_t stage_bytes = off_bytes + col_bytes + val_bytes;
    static constexpr size_t header = (128 + sizeof(StageMeta) * kStages + 127) / 128 * 128;
    static constexpr size_t total = header + stage_bytes * kStages;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool HAS_PEERS, typename OffT, typename ValT>
__global__ void __launch_bounds__(kStreamBlock)
stream_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap, const int32_t *__restrict__ Aj,
              const ValT *__restrict__ Ax, const ValT *__restrict__ x, ValT *__restrict__ y,
              const ValT *__restrict__ alpha_dev, PeerOut peers, int64_t num_tiles) {
    using L = StreamSmem<OffT, ValT>;
    constexpr int VO = 16 / sizeof(OffT);  // offsets per 16 bytes
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);  // [kStages]
    uint64_t *empty = full + kStages;                         // [kStages]
    StageMeta *meta = reinterpret_cast<StageMeta *>(smem_raw + 128);
    unsigned char *stages = smem_raw + L::header;
    auto s_off = [&](int s) { return reinterpret_cast<OffT *>(stages + (size_t)s * L::stage_bytes); };
    auto s_col = [&](int s) { return reinterpret_cast<int32_t *>(stages + (size_t)s * L::stage_bytes + L::off_bytes); };
    auto s_val = [&](int s) {
        return reinterpret_cast<ValT *>(stages + (size_t)s * L::stage_bytes + L::off_bytes + L::col_bytes);
    };

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // this CTA's tiles: an equal share of the tile range, contiguous
    const int64_t t_begin = num_tiles * (int64_t)blockIdx.x / gridDim.x;
    const int64_t t_end = num_tiles * ((int64_t)blockIdx.x + 1) / gridDim.x;

    if (tid >= kRows) {
        // ------------------------------------------------------------------ producer warp
        const int lane = tid - kRows;
        const uint64_t pol = policy_evict_first();
        int64_t it = 0;  // stage items issued
        // tile boundaries, 32 at a time: lane i holds Ap[first row of tile tb + i]; the batch after
        // this one is loaded a batch early so that its latency is never waited for
        auto load_bounds = [&](int64_t tb) -> long long {
            const int64_t t = tb + lane;
            if (t > t_end) return 0;
            const int64_t r = t * kRows < (int64_t)n_rows ? t * kRows : (int64_t)n_rows;
            return (long long)__ldg(Ap + r);
        };
        long long bounds_next = load_bounds(t_begin);
        for (int64_t tb = t_begin; tb < t_end; tb += 31) {
            const long long bounds = bounds_next;
            bounds_next = load_bounds(tb + 31);
            const int in_batch = (int)((t_end - tb) < 31 ? (t_end - tb) : 31);
            for (int j = 0; j < in_batch; ++j) {
                const int64_t tile = tb + j;
                const long long B = __shfl_sync(0xffffffffu, bounds, j);
                const long long E = __shfl_sync(0xffffffffu, bounds, j + 1);
                const int64_t row0 = tile * kRows;
                const int rows = (int)(((int64_t)n_rows - row0) < kRows ? ((int64_t)n_rows - row0) : kRows);
                const long long abase0 = B & ~(long long)(kAlign - 1);
                const long long span = E - abase0;
                const int chunks = span > 0 ? (int)((span + kCap - 1) / kCap) : 1;
                for (int ch = 0; ch < chunks; ++ch, ++it) {
                    const int s = (int)(it % kStages);
                    const uint32_t phase = (uint32_t)((it / kStages) & 1);
                    if (lane == 0) mbar_wait(&empty[s], phase ^ 1u);
                    __syncwarp();
                    const long long abase = abase0 + (long long)ch * kCap;
                    const long long lo = ch == 0 ? B : abase;
                    const long long hi = abase + kCap < E ? abase + kCap : E;
                    // bulk part of Aj / Ax: [abase, up4(hi)), kept inside the arrays
                    long long bend = (hi + 3) & ~3ll;
                    if (bend > (long long)nnz) bend = (long long)nnz & ~3ll;
                    const uint32_t n_bulk = bend > abase ? (uint32_t)(bend - abase) : 0u;
                    // the last few elements of the matrix, when nnz is not a multiple of 4
                    const long long tail_beg = abase + n_bulk > lo ? abase + n_bulk : lo;
                    for (long long k = tail_beg + lane; k < hi; k += 32) {
                        s_col(s)[k - abase] = __ldg(Aj + k);
                        s_val(s)[k - abase] = __ldg(Ax + k);
                    }
                    // row offsets travel with the first chunk: whole 16-byte groups by bulk copy,
                    // the remainder (the last tile of the matrix only, or just Ap[row0 + rows] = E)
                    // by plain stores
                    uint32_t n_off = 0;
                    if (ch == 0) {
                        n_off = (uint32_t)((rows + 1) / VO * VO);
                        for (int k = (int)n_off + lane; k <= rows; k += 32)
                            s_off(s)[k] = k == rows ? (OffT)E : __ldg(Ap + row0 + k);
                    }
                    if (lane == 0) {
                        StageMeta m;
                        m.abase = abase; m.lo = lo; m.hi = hi;
                        m.row0 = (int)row0; m.rows = rows; m.first = ch == 0; m.last = ch == chunks - 1;
                        meta[s] = m;
                    }
                    __syncwarp();  // the plain stores above are ordered before the arrive below
                    if (lane == 0) {
                        const uint32_t bytes = n_off * (uint32_t)sizeof(OffT) +
                                               n_bulk * (uint32_t)(sizeof(int32_t) + sizeof(ValT));
                        mbar_arrive_expect_tx(&full[s], bytes);
                        if (n_off) bulk_g2s(s_off(s), Ap + row0, n_off * (uint32_t)sizeof(OffT), &full[s], pol);
                        if (n_bulk) {
                            bulk_g2s(s_col(s), Aj + abase, n_bulk * (uint32_t)sizeof(int32_t), &full[s], pol);
                            bulk_g2s(s_val(s), Ax + abase, n_bulk * (uint32_t)sizeof(ValT), &full[s], pol);
                        }
                    }
                }
            }
        }
        // end marker
        const int s = (int)(it % kStages);
        const uint32_t phase = (uint32_t)((it / kStages) & 1);
        if (lane == 0) {
            mbar_wait(&empty[s], phase ^ 1u);
            StageMeta m;
            m.abase = 0; m.lo = 0; m.hi = 0; m.row0 = 0; m.rows = 0; m.first = 0; m.last = 0;
            meta[s] = m;
            mbar_arrive_expect_tx(&full[s], 0);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const uint64_t pol_x = policy_evict_last();
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;
    long long b = 0, e = 0;
    ValT sum = (ValT)0;
    for (int64_t it = 0;; ++it) {
        const int s = (int)(it % kStages);
        const uint32_t phase = (uint32_t)((it / kStages) & 1);
        mbar_wait(&full[s], phase);
        const StageMeta m = meta[s];
        if (m.rows == 0) break;
        const bool mine = tid < m.rows;
        if (m.first) {
            b = mine ? (long long)s_off(s)[tid] : 0;
            e = mine ? (long long)s_off(s)[tid + 1] : 0;
            sum = (ValT)0;
        }
        // the row's part of this chunk, as slots of the stage
        const int lo = (int)((b > m.lo ? b : m.lo) - m.abase);
        const int hi = (int)((e < m.hi ? e : m.hi) - m.abase);
        const int32_t *cs = s_col(s);
        const ValT *vs = s_val(s);
        // four nonzeros per trip: their gathers are in flight together, the additions stay in
        // the order of the row
#if SPMV_STREAM_ABLATE >= 2
        if (lo < 0)
#endif
        for (int i = lo; i < hi; i += 4) {
            const int n = hi - i;
            const int c0 = cs[i];
            const int c1 = n > 1 ? cs[i + 1] : c0;
            const int c2 = n > 2 ? cs[i + 2] : c0;
            const int c3 = n > 3 ? cs[i + 3] : c0;
#if SPMV_STREAM_ABLATE == 1
            const ValT x0 = (ValT)c0, x1 = (ValT)c1, x2 = (ValT)c2, x3 = (ValT)c3;
#else
            const ValT x0 = ldg_hint(x + c0, pol_x);
            const ValT x1 = ldg_hint(x + c1, pol_x);
            const ValT x2 = ldg_hint(x + c2, pol_x);
            const ValT x3 = ldg_hint(x + c3, pol_x);
#endif
            sum += vs[i] * x0;
            if (n > 1) sum += vs[i + 1] * x1;
            if (n > 2) sum += vs[i + 2] * x2;
            if (n > 3) sum += vs[i + 3] * x3;
        }
        if (m.last && mine) {
            const int64_t row = (int64_t)m.row0 + tid;
            if (HAS_PEERS) store_y_nonempty(y, peers, row, alpha * sum, e > b);
            else y[row] = alpha * sum;
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
}

}  // namespace

template <typename OffT, typename ValT>
int launch_stream(const SpmvProblem<OffT, ValT> &p) {
    if (p.n_rows <= 0) return SPMVB200_OK;
    using L = StreamSmem<OffT, ValT>;
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    const int64_t num_tiles = ((int64_t)p.n_rows + kRows - 1) / kRows;
    const bool has_peers = p.peers.n != 0;
    auto kernel = has_peers ? stream_kernel<true, OffT, ValT> : stream_kernel<false, OffT, ValT>;
    static bool configured[2] = {false, false};  // per instantiation
    if (!configured[has_peers]) {
        SPMV_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
        configured[has_peers] = true;
    }
    // persistent grid: "stream_ctas_per_sm" CTAs per SM, never more CTAs than tiles.  Measured on
    // the 1024^2 Laplacian (L2 flushed): 1 -> 30.7 us, 2 -> 21.5, 3 -> 20.5, 4 -> 24.6 (shared memory
    // squeezes the L1 the gathers need); the stage geometry (256/512-row tiles, 3-6 stages) moves
    // nothing, and with the consumers switched off the TMA stream alone takes 18.4 us: what is left
    // is launch + two dependent DRAM round trips before the first tile arrives.
    int64_t per_sm = option_get("stream_ctas_per_sm", 3);
    if (per_sm < 1) per_sm = 1;
    int occ = 1;
    SPMV_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kStreamBlock, L::total));
    if (occ < 1) occ = 1;
    if (per_sm > occ) per_sm = occ;
    int64_t grid = (int64_t)di->sm_count * per_sm;
    if (grid > num_tiles) grid = num_tiles;
    LaunchCfg lc;
    make_launch_cfg(lc, dim3((unsigned)grid), dim3(kStreamBlock), L::total, p.stream, p.x,
                    (size_t)p.n_cols * sizeof(ValT));
    {
        KernelTimerScope timed(p.stream);
        SPMV_CUDA_TRY(cudaLaunchKernelEx(&lc.cfg, kernel, p.n_rows, p.nnz, p.Ap, p.Aj, p.Ax, p.x, p.y,
                                         p.alpha_dev, p.peers, num_tiles));
    }
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

template int launch_stream<int32_t, float>(const SpmvProblem<int32_t, float> &);
template int launch_stream<int32_t, double>(const SpmvProblem<int32_t, double> &);
template int launch_stream<int64_t, float>(const SpmvProblem<int64_t, float> &);
template int launch_stream<int64_t, double>(const SpmvProblem<int64_t, double> &);

}  // namespace spmvb200
