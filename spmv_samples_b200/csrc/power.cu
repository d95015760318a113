// power.cu -- the two scalar helpers of the power iteration (BASELINE.json configs[4]:
// "100 power-iteration SpMVs"): a deterministic sum of squares and 1/sqrt on the device, so the
// normalisation x <- A x / ||A x|| never synchronises the host.  The scale is applied by the
// next SpMV through its device alpha (spmvb200_args_t.alpha_dev).
// The reference has no iteration driver at all; reference/main.cu:102-113 just repeats the call.
#include "common.cuh"

namespace spmvb200 {
namespace {

constexpr int kRedBlock = 256;

template <typename ValT>
__global__ void __launch_bounds__(kRedBlock)
sumsq_partial_kernel(int64_t n, const ValT *__restrict__ v, double *__restrict__ partial) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kRedBlock + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kRedBlock) {
        const double t = (double)v[i];
        acc += t * t;
    }
    __shared__ double s[kRedBlock / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kRedBlock / 32; ++w) t += s[w];
        partial[blockIdx.x] = t;
    }
}

// one block, fixed order: the result does not depend on scheduling
__global__ void __launch_bounds__(kRedBlock)
sumsq_final_kernel(int n_partials, const double *__restrict__ partial, double *__restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += kRedBlock) acc += partial[i];
    __shared__ double s[kRedBlock / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kRedBlock / 32; ++w) t += s[w];
        *out = t;
    }
}

template <typename ValT>
__global__ void inv_sqrt_kernel(const double *__restrict__ sumsq, ValT *__restrict__ alpha) {
    const double s = *sumsq;
    *alpha = (ValT)(s > 0.0 ? 1.0 / sqrt(s) : 1.0);
}

}  // namespace

template <typename ValT>
int sum_squares(int64_t n, const ValT *v, double *sumsq_dev, cudaStream_t stream) {
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    int64_t blocks = (n + kRedBlock - 1) / kRedBlock;
    const int64_t cap = (int64_t)di->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    void *partial = nullptr;
    SPMV_TRY(scratch_get(stream, SCRATCH_MISC, (size_t)blocks * sizeof(double), &partial));
    sumsq_partial_kernel<ValT><<<(unsigned)blocks, kRedBlock, 0, stream>>>(n, v, static_cast<double *>(partial));
    SPMV_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, kRedBlock, 0, stream>>>((int)blocks, static_cast<const double *>(partial), sumsq_dev);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}
template int sum_squares<float>(int64_t, const float *, double *, cudaStream_t);
template int sum_squares<double>(int64_t, const double *, double *, cudaStream_t);

template <typename ValT>
int inv_sqrt(const double *sumsq_dev, ValT *alpha_dev, cudaStream_t stream) {
    inv_sqrt_kernel<ValT><<<1, 1, 0, stream>>>(sumsq_dev, alpha_dev);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}
template int inv_sqrt<float>(const double *, float *, cudaStream_t);
template int inv_sqrt<double>(const double *, double *, cudaStream_t);

}  // namespace spmvb200
