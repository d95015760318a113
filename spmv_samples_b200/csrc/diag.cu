// diag.cu -- the gather-rate yardstick the SpMV numbers are read against.
//
// Not part of any SpMV: a kernel that does nothing but what bounds the gather-heavy
// configurations -- every thread streams 8 consecutive column indices (two 128-bit loads) and
// gathers x[col] with the same load the kernels use -- over uniformly random columns in a
// footprint of the caller's choice.  bench.py reports it beside c2/c3 (x L2-resident: the L1TEX
// ceiling of one gathered line per cycle per SM) and beside c5 (x = 512 MB).  tools/
// l2_gather_probe.cu is the long form with the column distributions and load flavours.
#include "common.cuh"

namespace spmvb200 {

namespace {

__device__ __forceinline__ uint64_t diag_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
diag_fill_idx_kernel(int64_t n_x, int64_t count, int32_t *__restrict__ idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = diag_mix64(0xD1A6ull + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull);
        idx[i] = (int32_t)(((h >> 32) * (uint64_t)n_x) >> 32);
    }
}

__global__ void __launch_bounds__(256)
diag_gather_kernel(const float *__restrict__ x, const int32_t *__restrict__ idx, int64_t count, float *out) {
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_x = policy_evict_last();
    const int64_t g = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
    if (g + 8 > count) return;
    const int4 a = ldg_stream_int4(idx + g, pol_stream), b = ldg_stream_int4(idx + g + 4, pol_stream);
    const int c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ldg_hint(x + c[k], pol_x);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
    if (acc == 123.456f) out[0] = acc;
}

}  // namespace

// best of `reps` launches (after one warm-up) of `count` uniformly random gathers over n_x floats
int gather_yardstick(int64_t n_x, int64_t count, int reps, cudaStream_t stream, double *best_ms) {
    if (n_x <= 0 || n_x > 0x7fffffffLL || count < 8 || reps < 1 || !best_ms) return SPMVB200_ERR_INVALID;
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    float *x = nullptr, *out = nullptr;
    int32_t *idx = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int status = SPMVB200_OK;
    auto ok = [&](cudaError_t e, const char *what) {
        if (e == cudaSuccess) return true;
        record_cuda_error(e, what, __FILE__, __LINE__);
        status = SPMVB200_ERR_CUDA;
        return false;
    };
    do {
        if (!ok(cudaMalloc(&x, (size_t)n_x * 4), "cudaMalloc x")) break;
        if (!ok(cudaMalloc(&idx, (size_t)count * 4), "cudaMalloc idx")) break;
        if (!ok(cudaMalloc(&out, 64), "cudaMalloc out")) break;
        if (!ok(cudaMemsetAsync(x, 0, (size_t)n_x * 4, stream), "memset")) break;
        if (!ok(cudaEventCreate(&e0), "event") || !ok(cudaEventCreate(&e1), "event")) break;
        diag_fill_idx_kernel<<<(unsigned)di->sm_count * 8, 256, 0, stream>>>(n_x, count, idx);
        const unsigned grid = (unsigned)((count / 8 + 255) / 256);
        double best = 1e30;
        for (int r = 0; r <= reps; ++r) {
            if (!ok(cudaEventRecord(e0, stream), "record")) break;
            diag_gather_kernel<<<grid, 256, 0, stream>>>(x, idx, count, out);
            if (!ok(cudaEventRecord(e1, stream), "record")) break;
            if (!ok(cudaStreamSynchronize(stream), "sync")) break;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        count_launch(reps + 2);
        if (!ok(cudaGetLastError(), "launch")) break;
        *best_ms = best;
    } while (false);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (x) cudaFree(x);
    if (idx) cudaFree(idx);
    if (out) cudaFree(out);
    return status;
}

}  // namespace spmvb200
