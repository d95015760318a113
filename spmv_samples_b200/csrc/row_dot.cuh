// row_dot.cuh -- the sub-warp row dot product shared by the CSR-vector and the dynamic-row
// kernels.
//
// Computes what the inner loops of reference/include/spmv/cusp/cusp_warp_reduce.cuh:33-50
// and LightSpMV.cuh:147-170 compute (T lanes stride over one row, then a shuffle reduce),
// but each lane moves four nonzeros per step with one 128-bit load of Aj and of Ax from a
// 16-byte aligned position (the row start rounded down), masking the elements that fall
// outside [row_start, row_end).  The
// reference only aligns for T == 32 and loads 4 bytes per lane.
//
// Rows much longer than the sub-warp can chew (row_is_long) are handed to the whole warp by
// the callers, so a power-law matrix does not serialise a 4-lane sub-warp on a 300K row.
#pragma once

#include "common.cuh"

namespace spmvb200 {

template <typename ValT>
struct Chunk {
    int4 c;
    typename Val4<ValT>::type v;
    unsigned mask;  // bit k: element k lies inside the row
};

// Four consecutive nonzeros starting at the 4-aligned position p, masked to [s, e).
// nnz bounds the arrays: the last vector of the matrix is read element-wise if it would run
// past them.
template <typename OffT, typename ValT>
__device__ __forceinline__ Chunk<ValT> fetch_chunk(OffT p, OffT s, OffT e, OffT nnz,
                                                   const int32_t *__restrict__ Aj,
                                                   const ValT *__restrict__ Ax,
                                                   uint64_t pol_stream) {
    Chunk<ValT> ch;
    ch.mask = 0;
    ch.c = make_int4(0, 0, 0, 0);
    ch.v.x = ch.v.y = ch.v.z = ch.v.w = (ValT)0;
    if (p < e) {
        ch.mask = ((p >= s) ? 1u : 0u) | ((p + 1 >= s && p + 1 < e) ? 2u : 0u) |
                  ((p + 2 >= s && p + 2 < e) ? 4u : 0u) | ((p + 3 >= s && p + 3 < e) ? 8u : 0u);
        if (p + 4 <= nnz) {
            ch.c = ldg_stream_int4(Aj + p, pol_stream);
            ch.v = ldg_stream_val4(Ax + p, pol_stream);
        } else {
            if (ch.mask & 1u) { ch.c.x = __ldg(Aj + p); ch.v.x = __ldg(Ax + p); }
            if (ch.mask & 2u) { ch.c.y = __ldg(Aj + p + 1); ch.v.y = __ldg(Ax + p + 1); }
            if (ch.mask & 4u) { ch.c.z = __ldg(Aj + p + 2); ch.v.z = __ldg(Ax + p + 2); }
            if (ch.mask & 8u) { ch.c.w = __ldg(Aj + p + 3); ch.v.w = __ldg(Ax + p + 3); }
        }
    }
    return ch;
}

// gather x for the valid elements (all four gathers issued before the first use) and
// accumulate; the product is masked too, because a neighbouring row's value may be Inf/NaN
template <typename ValT>
__device__ __forceinline__ ValT consume_chunk(const Chunk<ValT> &ch, const ValT *__restrict__ x,
                                              uint64_t pol_x, ValT sum) {
    const ValT x0 = (ch.mask & 1u) ? ldg_hint(x + ch.c.x, pol_x) : (ValT)0;
    const ValT x1 = (ch.mask & 2u) ? ldg_hint(x + ch.c.y, pol_x) : (ValT)0;
    const ValT x2 = (ch.mask & 4u) ? ldg_hint(x + ch.c.z, pol_x) : (ValT)0;
    const ValT x3 = (ch.mask & 8u) ? ldg_hint(x + ch.c.w, pol_x) : (ValT)0;
    if (ch.mask & 1u) sum += ch.v.x * x0;
    if (ch.mask & 2u) sum += ch.v.y * x1;
    if (ch.mask & 4u) sum += ch.v.z * x2;
    if (ch.mask & 8u) sum += ch.v.w * x3;
    return sum;
}

// Partial sum of row [s, e) seen by lane `lane` of a T-lane group.
template <int T, typename OffT, typename ValT>
__device__ __forceinline__ ValT row_partial(OffT s, OffT e, OffT nnz, int lane,
                                            const int32_t *__restrict__ Aj,
                                            const ValT *__restrict__ Ax,
                                            const ValT *__restrict__ x, uint64_t pol_stream,
                                            uint64_t pol_x) {
    ValT sum = (ValT)0;
    const OffT a = s & ~(OffT)3;
    // one chunk per trip: two in flight per trip was measured slower on every configuration
    // (registers 32 -> 48-60, occupancy down; c2 273 -> 291 us, c4 298 -> 319 us)
    for (OffT p = a + (OffT)(4 * lane); p < e; p += (OffT)(4 * T)) {
        const Chunk<ValT> c0 = fetch_chunk<OffT, ValT>(p, s, e, nnz, Aj, Ax, pol_stream);
        sum = consume_chunk<ValT>(c0, x, pol_x, sum);
    }
    return sum;
}

// A row is "long" for a T-lane sub-warp when it would take more than 16 steps.
template <int T, typename OffT>
__device__ __forceinline__ bool row_is_long(OffT len) {
    return T < 32 && len > (OffT)(64 * T);
}

// The long rows of one warp step, each reduced by all 32 lanes.  `is_long` is per sub-warp;
// s, e, row are the sub-warp's values.  Returns nothing: stores directly.
template <int T, typename OffT, typename ValT>
__device__ __forceinline__ void warp_long_rows(bool is_long, OffT s, OffT e, int64_t row, OffT nnz,
                                               const int32_t *__restrict__ Aj,
                                               const ValT *__restrict__ Ax,
                                               const ValT *__restrict__ x, ValT *__restrict__ y,
                                               const PeerOut &peers, ValT alpha, uint64_t pol_stream,
                                               uint64_t pol_x) {
    if (T == 32) return;
    const int wlane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, is_long && (wlane & (T - 1)) == 0);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        todo &= todo - 1;
        const OffT ls = __shfl_sync(0xffffffffu, s, leader);
        const OffT le = __shfl_sync(0xffffffffu, e, leader);
        const int64_t lrow = __shfl_sync(0xffffffffu, row, leader);
        ValT ps = row_partial<32, OffT, ValT>(ls, le, nnz, wlane, Aj, Ax, x, pol_stream, pol_x);
        ps = subwarp_sum<32>(ps);
        if (wlane == 0) store_y(y, peers, lrow, alpha * ps);
    }
}

// Third tier: rows longer than kHugeRow are queued in shared memory and reduced by the whole
// CTA (every thread strides the row with 128-bit loads, block reduction in a fixed order).  A
// power-law matrix has a few rows of 10^5..10^6 nonzeros; one warp on such a row is a
// millisecond-long tail (the CSR-vector kernel took 6.2 ms on R-MAT scale 24 before this).
constexpr int kHugeRow = 16384;
constexpr int kMaxHugePerCta = 8;
struct HugeList {
    int count;
    long long s[kMaxHugePerCta], e[kMaxHugePerCta], row[kMaxHugePerCta];
};

// leader lane of a sub-warp queues its row; false when the list is full (the caller then keeps
// the row on the warp-level path)
__device__ __forceinline__ bool push_huge(HugeList &hl, long long s, long long e, long long row) {
    const int slot = atomicAdd(&hl.count, 1);
    if (slot >= kMaxHugePerCta) return false;
    hl.s[slot] = s;
    hl.e[slot] = e;
    hl.row[slot] = row;
    return true;
}

// all threads of the CTA; BLOCK = blockDim.x.  s_red: BLOCK/32 values of shared memory.
template <int BLOCK, typename OffT, typename ValT>
__device__ __forceinline__ void cta_huge_rows(HugeList &hl, ValT *s_red, OffT nnz,
                                              const int32_t *__restrict__ Aj,
                                              const ValT *__restrict__ Ax,
                                              const ValT *__restrict__ x, ValT *__restrict__ y,
                                              const PeerOut &peers, ValT alpha, uint64_t pol_stream,
                                              uint64_t pol_x) {
    __syncthreads();  // the list is complete
    const int n = hl.count < kMaxHugePerCta ? hl.count : kMaxHugePerCta;
    for (int h = 0; h < n; ++h) {
        ValT ps = row_partial<BLOCK, OffT, ValT>((OffT)hl.s[h], (OffT)hl.e[h], nnz, (int)threadIdx.x, Aj,
                                                 Ax, x, pol_stream, pol_x);
        ps = subwarp_sum<32>(ps);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ps;
        __syncthreads();
        if (threadIdx.x == 0) {
            ValT tot = (ValT)0;
#pragma unroll
            for (int w = 0; w < BLOCK / 32; ++w) tot += s_red[w];
            store_y(y, peers, hl.row[h], alpha * tot);
        }
        __syncthreads();
    }
}

}  // namespace spmvb200
