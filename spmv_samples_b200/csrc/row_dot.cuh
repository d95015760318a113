// row_dot.cuh -- the sub-warp row dot product shared by the CSR-vector and the dynamic-row
// kernels.
//
// Computes what the inner loops of reference/include/spmv/cusp/cusp_warp_reduce.cuh:33-50
// and LightSpMV.cuh:147-170 compute (T lanes stride over one row, then a shuffle reduce),
// but each lane moves four nonzeros per step with one 128-bit load of Aj and of Ax from a
// 16-byte aligned position (the row start rounded down), masking the elements that fall
// outside [row_start, row_end).  The reference only aligns for T == 32 and loads 4 bytes
// per lane.
#pragma once

#include "common.cuh"

namespace spmvb200 {

// Partial sum of row [s, e) seen by lane `lane` of a T-lane sub-warp.  nnz bounds the
// arrays: the last vector of the matrix is read element-wise if it would run past them.
template <int T, typename OffT, typename ValT>
__device__ __forceinline__ ValT row_partial(OffT s, OffT e, OffT nnz, int lane,
                                            const int32_t *__restrict__ Aj,
                                            const ValT *__restrict__ Ax,
                                            const ValT *__restrict__ x, uint64_t pol_stream,
                                            uint64_t pol_x) {
    ValT sum = (ValT)0;
    const OffT a = s & ~(OffT)3;
    for (OffT p = a + (OffT)(4 * lane); p < e; p += (OffT)(4 * T)) {
        if (p + 4 <= nnz) {
            const int4 c = ldg_stream_int4(Aj + p, pol_stream);
            const typename Val4<ValT>::type v = ldg_stream_val4(Ax + p, pol_stream);
            const bool m0 = p >= s;  // p < e holds
            const bool m1 = (p + 1 >= s) && (p + 1 < e);
            const bool m2 = (p + 2 >= s) && (p + 2 < e);
            const bool m3 = (p + 3 >= s) && (p + 3 < e);
            const ValT x0 = m0 ? ldg_hint(x + c.x, pol_x) : (ValT)0;
            const ValT x1 = m1 ? ldg_hint(x + c.y, pol_x) : (ValT)0;
            const ValT x2 = m2 ? ldg_hint(x + c.z, pol_x) : (ValT)0;
            const ValT x3 = m3 ? ldg_hint(x + c.w, pol_x) : (ValT)0;
            // mask the product, not just the gather: a neighbouring row's value may be Inf/NaN
            if (m0) sum += v.x * x0;
            if (m1) sum += v.y * x1;
            if (m2) sum += v.z * x2;
            if (m3) sum += v.w * x3;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const OffT q = p + k;
                if (q >= s && q < e) sum += __ldg(Ax + q) * __ldg(x + __ldg(Aj + q));
            }
        }
    }
    return sum;
}

}  // namespace spmvb200
