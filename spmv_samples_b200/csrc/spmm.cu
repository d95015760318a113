// spmm.cu -- CSR sparse matrix times K dense vectors at once (K in {2, 4, 8}), Y = alpha*A*X with
// X [n_cols x K] and Y [n_rows x K] row-major.
//
// SURVEY.md 8(f) rank 4 ("multi-vector: the only way past the SpMV byte roofline"); the
// reference has no such kind.  Why it matters on B200: a CSR SpMV whose columns are scattered
// is bound by the L1TEX gather rate, one wavefront per nonzero for 4 useful bytes (DESIGN.md
// section 3).  With K right-hand sides interleaved row-major the same wavefront returns K values,
// so the gather cost per nonzero stays and the useful work grows K-fold.
//
// Kernel: the CSR-vector scheme of vector.cu (sub-warp per row, 128-bit loads of Aj / Ax from a
// 16-byte aligned position, masking, shuffle reduction) with K accumulators per lane and the
// same three tiers for long rows (sub-warp, whole warp, whole CTA).
#include "common.cuh"
#include "row_dot.cuh"

namespace spmvb200 {

namespace {

constexpr int kSpmmBlock = 256;

// K consecutive values with the widest aligned loads (rows of X are 16-byte aligned when
// K * sizeof(ValT) >= 16, 8-byte aligned for K = 2 floats)
template <int K>
__device__ __forceinline__ void load_row(const float *p, float (&v)[K]) {
    if constexpr (K == 2) {
        const float2 a = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = a.x; v[1] = a.y;
    } else {
#pragma unroll
        for (int i = 0; i < K / 4; ++i) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(p) + i);
            v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
        }
    }
}
template <int K>
__device__ __forceinline__ void load_row(const double *p, double (&v)[K]) {
#pragma unroll
    for (int i = 0; i < K / 2; ++i) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(p) + i);
        v[2 * i] = a.x; v[2 * i + 1] = a.y;
    }
}
template <int K>
__device__ __forceinline__ void store_row(float *p, const float (&v)[K]) {
    if constexpr (K == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < K / 4; ++i)
            reinterpret_cast<float4 *>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
}
template <int K>
__device__ __forceinline__ void store_row(double *p, const double (&v)[K]) {
#pragma unroll
    for (int i = 0; i < K / 2; ++i) reinterpret_cast<double2 *>(p)[i] = make_double2(v[2 * i], v[2 * i + 1]);
}

// one masked nonzero: acc += a * X[col, :]
template <int K, typename ValT>
__device__ __forceinline__ void fma_row(bool valid, ValT a, int col, const ValT *__restrict__ X, int64_t ldx,
                                        ValT (&acc)[K]) {
    if (valid) {
        ValT xr[K];
        load_row<K>(X + (int64_t)col * ldx, xr);
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += a * xr[k];
    }
}

// partial sums of row [s, e) seen by lane `lane` of a T-lane group
template <int T, int K, typename OffT, typename ValT>
__device__ __forceinline__ void row_partial_k(OffT s, OffT e, OffT nnz, int lane,
                                              const int32_t *__restrict__ Aj, const ValT *__restrict__ Ax,
                                              const ValT *__restrict__ X, int64_t ldx, uint64_t pol_stream,
                                              ValT (&acc)[K]) {
    const OffT a = s & ~(OffT)3;
    for (OffT p = a + (OffT)(4 * lane); p < e; p += (OffT)(4 * T)) {
        const Chunk<ValT> ch = fetch_chunk<OffT, ValT>(p, s, e, nnz, Aj, Ax, pol_stream);
        fma_row<K>(ch.mask & 1u, ch.v.x, ch.c.x, X, ldx, acc);
        fma_row<K>(ch.mask & 2u, ch.v.y, ch.c.y, X, ldx, acc);
        fma_row<K>(ch.mask & 4u, ch.v.z, ch.c.z, X, ldx, acc);
        fma_row<K>(ch.mask & 8u, ch.v.w, ch.c.w, X, ldx, acc);
    }
}

template <int T, int K, typename ValT>
__device__ __forceinline__ void subwarp_sum_k(ValT (&acc)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = subwarp_sum<T>(acc[k]);
}

template <int T, int K, typename OffT, typename ValT>
__global__ void __launch_bounds__(kSpmmBlock)
spmm_vector_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap, const int32_t *__restrict__ Aj,
                   const ValT *__restrict__ Ax, const ValT *__restrict__ X, int64_t ldx,
                   ValT *__restrict__ Y, int64_t ldy, const ValT *__restrict__ alpha_dev) {
    const int64_t gtid = (int64_t)blockIdx.x * kSpmmBlock + threadIdx.x;
    const int64_t row = gtid / T;
    const int lane = threadIdx.x & (T - 1);
    const int wlane = threadIdx.x & 31;
    const bool active = row < n_rows;
    const uint64_t pol_stream = policy_evict_first();
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;
    __shared__ HugeList hl;
    __shared__ ValT s_red[kSpmmBlock / 32][K];
    if (threadIdx.x == 0) hl.count = 0;
    __syncthreads();

    OffT s = 0, e = 0;
    if (active) {
        s = __ldg(Ap + row);
        e = __ldg(Ap + row + 1);
    }
    int queued = (active && lane == 0 && e - s > (OffT)kHugeRow) ? (int)push_huge(hl, s, e, row) : 0;
    queued = __shfl_sync(0xffffffffu, queued, wlane & ~(T - 1));
    const bool is_long = !queued && row_is_long<T, OffT>(e - s);

    // tier 1: the sub-warp
    ValT acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = (ValT)0;
    if (active && !is_long && !queued) row_partial_k<T, K, OffT, ValT>(s, e, nnz, lane, Aj, Ax, X, ldx, pol_stream, acc);
    subwarp_sum_k<T, K>(acc);
    if (active && !is_long && !queued && lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] *= alpha;
        store_row<K>(Y + row * ldy, acc);
    }
    // tier 2: the whole warp, one long row at a time
    if (T < 32) {
        unsigned todo = __ballot_sync(0xffffffffu, is_long && lane == 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            todo &= todo - 1;
            const OffT ls = __shfl_sync(0xffffffffu, s, leader);
            const OffT le = __shfl_sync(0xffffffffu, e, leader);
            const int64_t lrow = __shfl_sync(0xffffffffu, row, leader);
            ValT a2[K];
#pragma unroll
            for (int k = 0; k < K; ++k) a2[k] = (ValT)0;
            row_partial_k<32, K, OffT, ValT>(ls, le, nnz, wlane, Aj, Ax, X, ldx, pol_stream, a2);
            subwarp_sum_k<32, K>(a2);
            if (wlane == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) a2[k] *= alpha;
                store_row<K>(Y + lrow * ldy, a2);
            }
        }
    }
    // tier 3: the whole CTA, one hub row at a time
    __syncthreads();
    const int n_huge = hl.count < kMaxHugePerCta ? hl.count : kMaxHugePerCta;
    for (int h = 0; h < n_huge; ++h) {
        ValT a3[K];
#pragma unroll
        for (int k = 0; k < K; ++k) a3[k] = (ValT)0;
        row_partial_k<kSpmmBlock, K, OffT, ValT>((OffT)hl.s[h], (OffT)hl.e[h], nnz, (int)threadIdx.x, Aj, Ax, X,
                                                 ldx, pol_stream, a3);
        subwarp_sum_k<32, K>(a3);
        if (wlane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) s_red[threadIdx.x >> 5][k] = a3[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            ValT tot[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                ValT t = (ValT)0;
#pragma unroll
                for (int w = 0; w < kSpmmBlock / 32; ++w) t += s_red[w][k];
                tot[k] = alpha * t;
            }
            store_row<K>(Y + hl.row[h] * ldy, tot);
        }
        __syncthreads();
    }
}

template <int T, int K, typename OffT, typename ValT>
int launch_TK(int32_t n_rows, OffT nnz, const OffT *Ap, const int32_t *Aj, const ValT *Ax, const ValT *X,
              int64_t ldx, ValT *Y, int64_t ldy, const ValT *alpha_dev, cudaStream_t stream) {
    const int64_t blocks = ((int64_t)n_rows * T + kSpmmBlock - 1) / kSpmmBlock;
    if (blocks > 0x7fffffffLL) return SPMVB200_ERR_UNSUPPORTED;
    KernelTimerScope timed(stream);
    spmm_vector_kernel<T, K, OffT, ValT><<<(unsigned)blocks, kSpmmBlock, 0, stream>>>(n_rows, nnz, Ap, Aj, Ax, X,
                                                                                    ldx, Y, ldy, alpha_dev);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

template <int K, typename OffT, typename ValT>
int launch_K(int width, int32_t n_rows, OffT nnz, const OffT *Ap, const int32_t *Aj, const ValT *Ax,
             const ValT *X, int64_t ldx, ValT *Y, int64_t ldy, const ValT *alpha_dev, cudaStream_t stream) {
#define GO(T) return launch_TK<T, K, OffT, ValT>(n_rows, nnz, Ap, Aj, Ax, X, ldx, Y, ldy, alpha_dev, stream)
    switch (width) {
        case 1: GO(1);
        case 2: GO(2);
        case 4: GO(4);
        case 8: GO(8);
        case 16: GO(16);
        default: GO(32);
    }
#undef GO
}

// ------------------------------------------------------------ merge-path tile kernel, K-wide
// The tile algorithm of merge.cu (register-staged nonzeros, byte flags at row starts, segmented
// scan, one thread per row end) with a K-vector per nonzero: one gather fetches the K values of
// X's row, the scan runs K times over values that share one set of flags.  Balances hub rows
// exactly like the SpMV kernel and keeps the K-fold amortisation of the gather.
constexpr int kMsBlock = 128;
// items per thread: 8 for two floats per nonzero, else 4.  The scan's shuffles and barriers are
// paid per thread and the gathers per item, so more items per thread pay until the K-vectors
// cost occupancy: R-MAT scale 24, fp32, 4 -> 8 items: K = 2 1590 -> 1422 us (40 registers),
// K = 4 2111 -> 2281 us (40 -> 64 registers).
template <int K, typename ValT> constexpr int ms_ipt() { return K * (int)sizeof(ValT) <= 8 ? 8 : 4; }
template <int K, typename ValT> constexpr int ms_tile() { return kMsBlock * ms_ipt<K, ValT>() - 4; }

template <int K, typename OffT, typename ValT>
__global__ void __launch_bounds__(kMsBlock)
merge_spmm_tile_kernel(int32_t n_rows, OffT nnz, const OffT *__restrict__ Ap, const int32_t *__restrict__ Aj,
                       const ValT *__restrict__ Ax, const ValT *__restrict__ X, int64_t ldx,
                       ValT *__restrict__ Y, int64_t ldy, const ValT *__restrict__ alpha_dev,
                       const int32_t *__restrict__ coords_x, int32_t *__restrict__ carry_row,
                       ValT *__restrict__ carry_val) {
    constexpr int IPT = ms_ipt<K, ValT>();
    constexpr int kMsSlots = kMsBlock * IPT;
    constexpr int kMsTile = kMsSlots - 4;  // path items per tile (slot 0 sits on a 16-byte boundary)
    __shared__ __align__(16) ValT s_scan[kMsSlots * K];
    __shared__ __align__(16) unsigned char s_flag[kMsSlots];
    __shared__ ValT s_wval[kMsBlock / 32][K];
    __shared__ int s_wflag[kMsBlock / 32];

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int64_t total = (int64_t)n_rows + (int64_t)nnz;
    const int64_t d0 = tile * kMsTile;
    const int64_t d1 = d0 + kMsTile < total ? d0 + kMsTile : total;
    const int32_t sx = __ldg(coords_x + tile);
    const int32_t ex = __ldg(coords_x + tile + 1);
    const int64_t sy = d0 - sx;
    const int R = ex - sx;                // rows that end inside this tile
    const int Z = (int)((d1 - ex) - sy);  // nonzeros inside this tile
    const int shift = (int)(sy & 3);      // slot s holds tile-local nonzero s - shift
    const int64_t a0 = sy - shift;

#pragma unroll
    for (int i = 0; i < IPT; i += 4) *reinterpret_cast<uint32_t *>(s_flag + tid * IPT + i) = 0u;
    const int slot0 = tid * IPT;
    const uint64_t pol_stream = policy_evict_first();
    int c[IPT];
    ValT a[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        c[i] = 0;
        a[i] = (ValT)0;
    }
    if (slot0 < shift + Z) {
        const int64_t g = a0 + slot0;
        if (g + IPT <= (int64_t)nnz) {
#pragma unroll
            for (int i = 0; i < IPT; i += 4) {
                const int4 cv = ldg_stream_int4(Aj + g + i, pol_stream);
                const typename Val4<ValT>::type av = ldg_stream_val4(Ax + g + i, pol_stream);
                c[i] = cv.x; c[i + 1] = cv.y; c[i + 2] = cv.z; c[i + 3] = cv.w;
                a[i] = av.x; a[i + 1] = av.y; a[i + 2] = av.z; a[i + 3] = av.w;
            }
        } else {  // the last vector of the matrix: element-wise, inside the arrays
#pragma unroll
            for (int i = 0; i < IPT; ++i) {
                if (g + i < (int64_t)nnz) {
                    c[i] = __ldg(Aj + g + i);
                    a[i] = __ldg(Ax + g + i);
                }
            }
        }
    }
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;
    __syncthreads();  // flags are clear

    for (int j = tid; j < R; j += kMsBlock) {
        const int64_t q = (int64_t)__ldg(Ap + sx + 1 + j) - sy;
        if (q < Z) s_flag[(int)q + shift] = 1;
    }
    // ---- the K-wide gathers, all issued before the first use
    ValT p[IPT][K];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int t = slot0 + i - shift;
        if (t >= 0 && t < Z) {
            load_row<K>(X + (int64_t)c[i] * ldx, p[i]);
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) p[i][k] = (ValT)0;
            a[i] = (ValT)0;
        }
    }
#pragma unroll
    for (int i = 0; i < IPT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k) p[i][k] *= a[i];
    __syncthreads();  // flags are set

    // ---- segmented scan: one set of flags, K values
    unsigned long long fbits = 0ull;
#pragma unroll
    for (int i = 0; i < IPT; i += 4)
        fbits |= (unsigned long long)*reinterpret_cast<const uint32_t *>(s_flag + slot0 + i) << (8 * i);
    int flag = fbits != 0ull;
    ValT val[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        ValT v = (ValT)0;
#pragma unroll
        for (int i = 0; i < IPT; ++i) v = ((fbits >> (8 * i)) & 1ull) ? p[i][k] : v + p[i][k];
        val[k] = v;
    }
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int pf = __shfl_up_sync(0xffffffffu, flag, d);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const ValT pv = __shfl_up_sync(0xffffffffu, val[k], d);
            if (lane >= d && !flag) val[k] += pv;
        }
        if (lane >= d) flag |= pf;
    }
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < K; ++k) s_wval[warp][k] = val[k];
        s_wflag[warp] = flag;
    }
    int ef = __shfl_up_sync(0xffffffffu, flag, 1);
    ValT run[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        run[k] = __shfl_up_sync(0xffffffffu, val[k], 1);
        if (lane == 0) run[k] = (ValT)0;
    }
    if (lane == 0) ef = 0;
    __syncthreads();
    if (!ef) {  // nothing in this warp before the thread starts a row: the earlier warps carry in
#pragma unroll
        for (int k = 0; k < K; ++k) {
            ValT wv = (ValT)0;
#pragma unroll
            for (int w = 0; w < kMsBlock / 32; ++w) {
                if (w < warp) {
                    const ValT v = s_wval[w][k];
                    wv = s_wflag[w] ? v : wv + v;
                }
            }
            run[k] += wv;
        }
    }
    // K planes of kMsSlots values: a thread's four consecutive slots of one plane are one
    // 16-byte (fp32) store at a 16-byte lane stride -- conflict-free, where a [slot][K] layout
    // put the lanes 64 bytes apart (4-way conflicts on the pipe the gathers also need)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        ValT r = run[k];
        ValT out4[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            r = ((fbits >> (8 * i)) & 1ull) ? p[i][k] : r + p[i][k];
            out4[i] = r;
        }
#pragma unroll
        for (int i = 0; i < IPT; i += 4)
            store_row<4>(s_scan + k * kMsSlots + slot0 + i, reinterpret_cast<const ValT(&)[4]>(out4[i]));
    }
    __syncthreads();

    // ---- one thread per row end; row sx+j covers tile-local nonzeros [max(Ap[sx+j]-sy,0), Ap[sx+j+1]-sy)
    for (int j = tid; j < R; j += kMsBlock) {
        const int64_t b64 = (int64_t)__ldg(Ap + sx + j) - sy;
        const int q = (int)((int64_t)__ldg(Ap + sx + 1 + j) - sy);
        const int b = b64 > 0 ? (int)b64 : 0;
        ValT out[K];
        if (q > b) {
#pragma unroll
            for (int k = 0; k < K; ++k) out[k] = alpha * s_scan[k * kMsSlots + q - 1 + shift];
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) out[k] = (ValT)0;
        }
        store_row<K>(Y + ((int64_t)sx + j) * ldy, out);
    }
    if (tid == 0) {
        const int64_t lq = R > 0 ? (int64_t)__ldg(Ap + ex) - sy : 0;
        const int lastq = lq > 0 ? (int)lq : 0;
        carry_row[tile] = ex;
#pragma unroll
        for (int k = 0; k < K; ++k)
            carry_val[tile * K + k] = Z > lastq ? s_scan[k * kMsSlots + Z - 1 + shift] : (ValT)0;
    }
}

// One thread per tile: the head of a run of tiles whose carry lands in one row adds the run's
// carries to that row of Y, in tile order (deterministic).
template <int K, typename ValT>
__global__ void __launch_bounds__(256)
merge_spmm_fixup_kernel(int32_t n_rows, int64_t num_tiles, const int32_t *__restrict__ carry_row,
                        const ValT *__restrict__ carry_val, ValT *__restrict__ Y, int64_t ldy,
                        const ValT *__restrict__ alpha_dev) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    const int32_t row = carry_row[t];
    if (row >= n_rows) return;
    if (t > 0 && carry_row[t - 1] == row) return;
    ValT sum[K];
#pragma unroll
    for (int k = 0; k < K; ++k) sum[k] = carry_val[t * K + k];
    for (int64_t u = t + 1; u < num_tiles && carry_row[u] == row; ++u)
#pragma unroll
        for (int k = 0; k < K; ++k) sum[k] += carry_val[u * K + k];
    const ValT alpha = alpha_dev ? __ldg(alpha_dev) : (ValT)1;
    ValT *y = Y + (int64_t)row * ldy;
#pragma unroll
    for (int k = 0; k < K; ++k) y[k] += alpha * sum[k];
}

template <int K, typename OffT, typename ValT>
int launch_spmm_merge_K(int32_t n_rows, OffT nnz, const OffT *Ap, const int32_t *Aj, const ValT *Ax,
                        const ValT *X, int64_t ldx, ValT *Y, int64_t ldy, const ValT *alpha_dev,
                        cudaStream_t stream) {
    const int64_t total = (int64_t)n_rows + (int64_t)nnz;
    constexpr int kMsTile = ms_tile<K, ValT>();
    const int64_t num_tiles = (total + kMsTile - 1) / kMsTile;
    if (num_tiles > 0x7fffffffLL) return SPMVB200_ERR_UNSUPPORTED;
    void *coords = nullptr, *crow = nullptr, *cval = nullptr;
    SPMV_TRY(scratch_get(stream, SCRATCH_COORDS, (size_t)(num_tiles + 1) * sizeof(int32_t), &coords));
    SPMV_TRY(scratch_get(stream, SCRATCH_CARRY_ROW, (size_t)num_tiles * sizeof(int32_t), &crow));
    SPMV_TRY(scratch_get(stream, SCRATCH_CARRY_VAL, (size_t)num_tiles * K * sizeof(ValT), &cval));
    SPMV_TRY(launch_partition<OffT>(n_rows, nnz, Ap, kMsTile, num_tiles + 1, static_cast<int32_t *>(coords),
                                    stream));
    SPMV_TRY(apply_carveout(reinterpret_cast<const void *>(&merge_spmm_tile_kernel<K, OffT, ValT>),
                            option_get("spmm_carveout", -1)));
    {
        KernelTimerScope timed(stream);
        merge_spmm_tile_kernel<K, OffT, ValT><<<(unsigned)num_tiles, kMsBlock, 0, stream>>>(
            n_rows, nnz, Ap, Aj, Ax, X, ldx, Y, ldy, alpha_dev, static_cast<const int32_t *>(coords),
            static_cast<int32_t *>(crow), static_cast<ValT *>(cval));
    }
    SPMV_LAUNCH_CHECK();
    merge_spmm_fixup_kernel<K, ValT><<<(unsigned)((num_tiles + 255) / 256), 256, 0, stream>>>(
        n_rows, num_tiles, static_cast<const int32_t *>(crow), static_cast<const ValT *>(cval), Y, ldy,
        alpha_dev);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

// row-major [n x k] (leading dimension ld) <-> k contiguous vectors of length n
template <typename ValT>
__global__ void __launch_bounds__(256)
split_columns_kernel(int64_t n, int k, const ValT *__restrict__ X, int64_t ld, ValT *__restrict__ xt) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * k;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / k;
        const int j = (int)(t % k);
        xt[(int64_t)j * n + i] = X[i * ld + j];
    }
}
template <typename ValT>
__global__ void __launch_bounds__(256)
join_columns_kernel(int64_t n, int k, const ValT *__restrict__ yt, ValT *__restrict__ Y, int64_t ld) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * k;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / k;
        const int j = (int)(t % k);
        Y[i * ld + j] = yt[(int64_t)j * n + i];
    }
}

}  // namespace

// Power-law matrices: the row-per-sub-warp scheme cannot balance hub rows of 10^5..10^6
// nonzeros (measured on R-MAT scale 24: 6.2 ms for k = 4 against 4 x 1.46 ms of merge-path
// SpMV), so for matrices the selector sends to merge-path the k right-hand sides are split into
// contiguous vectors, multiplied one at a time by the merge-path kernel, and joined again.  A
// merge-path tile kernel carrying k-vectors through its scan is the "next" item that would make
// this case profit from the multi-vector gather too.
template <typename OffT, typename ValT>
static int launch_spmm_by_columns(int k, int32_t n_rows, int32_t n_cols, OffT nnz, const OffT *Ap,
                                  const int32_t *Aj, const ValT *Ax, const ValT *X, int64_t ldx, ValT *Y,
                                  int64_t ldy, const ValT *alpha_dev, cudaStream_t stream) {
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    void *xt = nullptr, *yt = nullptr;
    SPMV_TRY(scratch_get(stream, SCRATCH_SPMM_X, (size_t)k * n_cols * sizeof(ValT), &xt));
    SPMV_TRY(scratch_get(stream, SCRATCH_SPMM_Y, (size_t)k * n_rows * sizeof(ValT), &yt));
    const unsigned grid = (unsigned)(di->sm_count * 8);
    split_columns_kernel<ValT><<<grid, 256, 0, stream>>>(n_cols, k, X, ldx, static_cast<ValT *>(xt));
    SPMV_LAUNCH_CHECK();
    for (int j = 0; j < k; ++j) {
        SpmvProblem<OffT, ValT> p;
        p.n_rows = n_rows;
        p.n_cols = n_cols;
        p.nnz = nnz;
        p.Ap = Ap;
        p.Aj = Aj;
        p.Ax = Ax;
        p.x = static_cast<const ValT *>(xt) + (size_t)j * n_cols;
        p.y = static_cast<ValT *>(yt) + (size_t)j * n_rows;
        p.alpha_dev = alpha_dev;
        p.peers.n = 0;
        for (auto &q : p.peers.ptr) q = nullptr;
        p.stream = stream;
        SPMV_TRY((launch_merge<OffT, ValT>(p)));
    }
    join_columns_kernel<ValT><<<grid, 256, 0, stream>>>(n_rows, k, static_cast<const ValT *>(yt), Y, ldy);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

template <typename OffT, typename ValT>
int launch_spmm(int k, int32_t n_rows, int32_t n_cols, OffT nnz, const OffT *Ap, const int32_t *Aj,
                const ValT *Ax, const ValT *X, int64_t ldx, ValT *Y, int64_t ldy, const ValT *alpha_dev,
                cudaStream_t stream) {
    if (n_rows <= 0 || n_cols <= 0) return SPMVB200_OK;
    if (k != 2 && k != 4 && k != 8) return SPMVB200_ERR_UNSUPPORTED;
    spmvb200_row_stats_t st;
    SPMV_TRY(row_stats<OffT>(n_rows, (int64_t)nnz, Ap, &st, stream, true));
    const bool force_merge = option_get("spmm_force_merge", 0) > 0;
    if (force_merge || (st.chosen_kind == SPMVB200_KIND_MERGE && option_get("spmm_force_vector", 0) == 0)) {
        if (!force_merge && option_get("spmm_by_columns", 0) > 0)  // ablation: K merge-path SpMVs
            return launch_spmm_by_columns<OffT, ValT>(k, n_rows, n_cols, nnz, Ap, Aj, Ax, X, ldx, Y, ldy,
                                                      alpha_dev, stream);
        switch (k) {
            case 2: return launch_spmm_merge_K<2, OffT, ValT>(n_rows, nnz, Ap, Aj, Ax, X, ldx, Y, ldy, alpha_dev, stream);
            case 4: return launch_spmm_merge_K<4, OffT, ValT>(n_rows, nnz, Ap, Aj, Ax, X, ldx, Y, ldy, alpha_dev, stream);
            default: return launch_spmm_merge_K<8, OffT, ValT>(n_rows, nnz, Ap, Aj, Ax, X, ldx, Y, ldy, alpha_dev, stream);
        }
    }
    int width = (int)option_get("vector_width", 0);
    if (width <= 0) width = pick_width_from_mean((double)nnz / (double)n_rows);
    switch (k) {
        case 2: return launch_K<2, OffT, ValT>(width, n_rows, nnz, Ap, Aj, Ax, X, ldx, Y, ldy, alpha_dev, stream);
        case 4: return launch_K<4, OffT, ValT>(width, n_rows, nnz, Ap, Aj, Ax, X, ldx, Y, ldy, alpha_dev, stream);
        case 8: return launch_K<8, OffT, ValT>(width, n_rows, nnz, Ap, Aj, Ax, X, ldx, Y, ldy, alpha_dev, stream);
        default: return SPMVB200_ERR_UNSUPPORTED;
    }
}
#define INST(OffT, ValT)                                                                                  \
    template int launch_spmm<OffT, ValT>(int, int32_t, int32_t, OffT, const OffT *, const int32_t *,      \
                                         const ValT *, const ValT *, int64_t, ValT *, int64_t,           \
                                         const ValT *, cudaStream_t);
INST(int32_t, float)
INST(int32_t, double)
INST(int64_t, float)
INST(int64_t, double)
#undef INST

}  // namespace spmvb200
