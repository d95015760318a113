// host_check.hpp -- the driver's self-check, not an SpMV kind.
//
// reference/main.cu:78-97 computes y on the host and prints sum|y_ref - y| for each kind.
// This header gives main.cu the same check with two changes the north_star asks for: the
// host sums are fp64 whatever the matrix type, and the per-row scale sum|a x| is returned so
// the driver can say pass/fail (|y - y_ref| <= tol * scale).  It is deliberately NOT in
// SPMV_KINDS: SpMV(kind_str, ...) can never dispatch to host code.
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

template <typename index_t, typename offset_t, typename mat_value_t, typename vec_x_value_t>
void SpMV_host_check(index_t n_rows, const offset_t *Ap, const index_t *Aj, const mat_value_t *Ax,
                     const vec_x_value_t *x, std::vector<double> &y_ref, std::vector<double> &scale) {
    y_ref.assign((size_t)n_rows, 0.0);
    scale.assign((size_t)n_rows, 0.0);
    for (int64_t r = 0; r < (int64_t)n_rows; ++r) {
        double acc = 0.0, mag = 0.0;
        for (offset_t k = Ap[r]; k < Ap[r + 1]; ++k) {
            const double t = (double)Ax[k] * (double)x[Aj[k]];
            acc += t;
            mag += std::fabs(t);
        }
        y_ref[(size_t)r] = acc;
        scale[(size_t)r] = mag;
    }
}
