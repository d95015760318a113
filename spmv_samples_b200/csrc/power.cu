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

// ------------------------------------------------------------------ norm exchange
// One kernel for everything between two SpMVs of the row-sharded power iteration: the sum of
// squares of this rank's slice, its exchange with the other ranks, and alpha = 1 / ||A x||.
//   * every block reduces its part of y to a partial; the last block to finish (a counter, the
//     threadfence-reduction pattern) adds the partials in block order -- deterministic;
//   * that block publishes the rank's sum into slot [step % 3][rank] of every rank's mailbox: a
//     peer-mapped store per rank, or one multimem.st through the NVLink multicast mapping;
//   * it then waits until the `world` entries of its own slot are there and adds them in rank
//     order, so every rank computes the same total bit for bit.
// A sum of squares is never negative and an empty slot holds -1, so the value is its own flag
// (one 8-byte store, no ordering between a value and a flag to get wrong).  The slot of the step
// after this one is cleared before publishing: a peer writes it only after it has seen this
// step's value.  Because a rank publishes only after its own SpMV kernels of the step are done
// (stream order + a system fence), leaving the wait means every peer's stores into this rank's
// replica of x have landed: the kernel is also the step barrier, which is what the NCCL
// all-reduce of the norm was used for before (three launches and a collective ago).
// The wait is bounded (~4 s of globaltimer): a rank that never arrives raises *error instead of
// hanging the others.
constexpr int kMailboxRanks = 8;
constexpr int kMailboxSlots = 3;

__device__ __forceinline__ void st_release_sys(double *p, double v) {
    asm volatile("st.release.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_acquire_sys(const double *p) {
    double v;
    asm volatile("ld.acquire.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct MailboxOut {
    double *ptr[kMailboxRanks];  // mailbox base of rank q as seen from this rank (q = rank: local)
    double *multicast;           // or one multicast address reaching every rank's mailbox
};

template <typename ValT>
__global__ void __launch_bounds__(kRedBlock)
norm_exchange_kernel(int64_t n, const ValT *__restrict__ v, double *__restrict__ partial,
                     unsigned int *__restrict__ done_counter, int rank, int world,
                     unsigned long long step, double *mailbox_local, MailboxOut out,
                     double *__restrict__ sumsq_out, ValT *__restrict__ alpha_out, int *error) {
    __shared__ double s[kRedBlock / 32];
    __shared__ bool is_last;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kRedBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kRedBlock) {
        const double t = (double)v[i];
        acc += t * t;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kRedBlock / 32; ++w) t += s[w];
        partial[blockIdx.x] = t;
        __threadfence();
        is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // ---- the last block: partials in block order
    acc = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kRedBlock) acc += partial[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    const int slot = (int)(step % kMailboxSlots), next = (int)((step + 1) % kMailboxSlots);
    if (threadIdx.x == 0) {
        double mine = 0.0;
#pragma unroll
        for (int w = 0; w < kRedBlock / 32; ++w) mine += s[w];
        *done_counter = 0;  // ready for the next launch
        for (int q = 0; q < world; ++q) mailbox_local[next * kMailboxRanks + q] = -1.0;
        __threadfence_system();  // this rank's SpMV stores and the cleared slot before the value
        if (out.multicast) {
            asm volatile("multimem.st.release.sys.global.f64 [%0], %1;" ::"l"(out.multicast + slot * kMailboxRanks + rank),
                         "d"(mine)
                         : "memory");
        } else {
            for (int q = 0; q < world; ++q) st_release_sys(out.ptr[q] + slot * kMailboxRanks + rank, mine);
        }
    }
    __syncthreads();
    if (threadIdx.x < world) {
        const double *src = mailbox_local + slot * kMailboxRanks + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(src) < 0.0) {
            if (global_timer_ns() - t0 > 4000000000ull) {
                atomicExch(error, 1 + (int)threadIdx.x);
                break;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int q = 0; q < world; ++q) total += ld_acquire_sys(mailbox_local + slot * kMailboxRanks + q);
        *sumsq_out = total;
        *alpha_out = (ValT)(total > 0.0 ? 1.0 / sqrt(total) : 1.0);
    }
}

}  // namespace

template <typename ValT>
int norm_exchange(int64_t n, const ValT *v, int rank, int world, uint64_t step, double *mailbox_local,
                  void *const *mailbox_of_rank, void *mailbox_multicast, double *sumsq_dev, ValT *alpha_dev,
                  int *error_dev, cudaStream_t stream) {
    if (world < 1 || world > kMailboxRanks || rank < 0 || rank >= world || !mailbox_local || !sumsq_dev ||
        !alpha_dev || !error_dev || (!mailbox_multicast && !mailbox_of_rank))
        return SPMVB200_ERR_INVALID;
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    int64_t blocks = (n + kRedBlock - 1) / kRedBlock;
    const int64_t cap = (int64_t)di->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    // partials, then the block counter (zero on first use: scratch_get clears what it allocates,
    // and the kernel's last block resets it)
    void *scratch = nullptr;
    SPMV_TRY(scratch_get(stream, SCRATCH_NORM, (size_t)(cap + 2) * sizeof(double), &scratch));
    double *partial = static_cast<double *>(scratch);
    unsigned int *counter = reinterpret_cast<unsigned int *>(partial + cap);
    MailboxOut out{};
    out.multicast = static_cast<double *>(mailbox_multicast);
    for (int q = 0; q < world; ++q) out.ptr[q] = mailbox_of_rank ? static_cast<double *>(mailbox_of_rank[q]) : nullptr;
    if (!out.multicast)
        for (int q = 0; q < world; ++q)
            if (!out.ptr[q]) return SPMVB200_ERR_INVALID;
    norm_exchange_kernel<ValT><<<(unsigned)blocks, kRedBlock, 0, stream>>>(
        n, v, partial, counter, rank, world, (unsigned long long)step, mailbox_local, out, sumsq_dev, alpha_dev,
        error_dev);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}
template int norm_exchange<float>(int64_t, const float *, int, int, uint64_t, double *, void *const *, void *,
                                  double *, float *, int *, cudaStream_t);
template int norm_exchange<double>(int64_t, const double *, int, int, uint64_t, double *, void *const *, void *,
                                   double *, double *, int *, cudaStream_t);

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
