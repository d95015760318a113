// hotx.cu -- the hot part of x, compacted: a per-matrix "gather plan" for very long x.
//
// Why (tools/l2_gather_probe.cu, profiles/r2_l2_probe.md): on R-MAT scale 27 the gathers of x
// alone -- no row structure, no values -- take as long as the whole SpMV (1.75 ms per 2^28
// gathers, 14 ms per SpMV).  x is 512 MB: twice the reach of the TLB (2 MB pages) and eight
// times what one die's L2 holds, and the L2 keeps 128-byte lines of which a random gather uses
// one 32-byte sector.  The column distribution of a power-law matrix is as skewed as its row
// distribution, though: on this matrix the 8 M most frequent columns (6 % of them) receive 91 %
// of the gathers.  Copied into one dense 32 MB array they occupy 16 pages and a quarter of one
// die's L2, and the same gathers run at the L1TEX rate again (1.00 ms per 2^28).
//
// The plan is built once per matrix (column histogram, threshold, ranks, a remapped copy of Aj in
// which a hot column c is stored as 0x80000000 | rank(c)); every SpMV then starts with a small
// kernel x_hot[r] = x[hot_cols[r]] and the tile kernel picks its gather base by the sign of the
// index.  Products and their order are unchanged, so y is bit-identical to the plain kernel's.
// The CSR arrays the caller passed are not modified; the plan costs nnz * 4 bytes of HBM.
// Nothing like it in the reference: its merge kernel gathers x[Aj[k]] as is
// (merge_based/agent_spmv_orig.cuh:474-506).
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <map>
#include <mutex>

#include "common.cuh"

namespace spmvb200 {

namespace {

constexpr int kCountBuckets = 1026;  // bucket b < 1025: columns seen exactly b times; 1025: more

__global__ void __launch_bounds__(256)
hot_count_kernel(const int32_t *__restrict__ Aj, int64_t nnz, uint32_t *__restrict__ counts) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; k < nnz; k += stride) {
        if (k + 4 <= nnz) {
            const int4 c = *reinterpret_cast<const int4 *>(Aj + k);
            atomicAdd(counts + c.x, 1u);
            atomicAdd(counts + c.y, 1u);
            atomicAdd(counts + c.z, 1u);
            atomicAdd(counts + c.w, 1u);
        } else {
            for (int64_t j = k; j < nnz; ++j) atomicAdd(counts + Aj[j], 1u);
        }
    }
}

// columns per count bucket, and the gathers they receive
__global__ void __launch_bounds__(256)
hot_hist_kernel(const uint32_t *__restrict__ counts, int64_t n_cols, unsigned long long *__restrict__ hist_cols,
                unsigned long long *__restrict__ hist_mass) {
    __shared__ unsigned int s_cols[kCountBuckets];
    __shared__ unsigned long long s_mass[kCountBuckets];
    for (int i = threadIdx.x; i < kCountBuckets; i += blockDim.x) {
        s_cols[i] = 0;
        s_mass[i] = 0;
    }
    __syncthreads();
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t n = counts[c];
        const int b = n < (uint32_t)(kCountBuckets - 1) ? (int)n : kCountBuckets - 1;
        atomicAdd(&s_cols[b], 1u);
        atomicAdd(&s_mass[b], (unsigned long long)n);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kCountBuckets; i += blockDim.x) {
        if (s_cols[i]) atomicAdd(&hist_cols[i], (unsigned long long)s_cols[i]);
        if (s_mass[i]) atomicAdd(&hist_mass[i], s_mass[i]);
    }
}

struct IsHot {
    uint32_t threshold;
    __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t &n) const { return n >= threshold ? 1u : 0u; }
};

// rank[] holds the exclusive scan of the hot flags on entry and the remap table on exit
__global__ void __launch_bounds__(256)
hot_remap_table_kernel(const uint32_t *__restrict__ counts, int64_t n_cols, uint32_t threshold,
                       uint32_t *__restrict__ rank, int32_t *__restrict__ hot_cols) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += (int64_t)gridDim.x * blockDim.x) {
        if (counts[c] >= threshold) {
            const uint32_t r = rank[c];
            hot_cols[r] = (int32_t)c;
            rank[c] = 0x80000000u | r;
        } else {
            rank[c] = (uint32_t)c;
        }
    }
}

__global__ void __launch_bounds__(256)
hot_remap_kernel(const int32_t *__restrict__ Aj, int64_t nnz, const uint32_t *__restrict__ table,
                 int32_t *__restrict__ Aj2) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; k < nnz; k += stride) {
        if (k + 4 <= nnz) {
            const int4 c = *reinterpret_cast<const int4 *>(Aj + k);
            int4 o;
            o.x = (int32_t)__ldg(table + c.x);
            o.y = (int32_t)__ldg(table + c.y);
            o.z = (int32_t)__ldg(table + c.z);
            o.w = (int32_t)__ldg(table + c.w);
            *reinterpret_cast<int4 *>(Aj2 + k) = o;
        } else {
            for (int64_t j = k; j < nnz; ++j) Aj2[j] = (int32_t)__ldg(table + Aj[j]);
        }
    }
}

// One warp per 32 columns: which of them are hot (a word of the bitmap) and the rank of the first
// (the exclusive scan at the word's first column).  Runs before the remap-table kernel turns the
// scan into the table.
__global__ void __launch_bounds__(256)
hot_bitmap_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ rank, int64_t n_cols,
                  uint32_t threshold, uint32_t *__restrict__ bitmap, uint32_t *__restrict__ rank32) {
    const int lane = threadIdx.x & 31;
    const int64_t words = (n_cols + 31) / 32;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < words;
         w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t c = w * 32 + lane;
        const unsigned mask = __ballot_sync(0xffffffffu, c < n_cols && counts[c] >= threshold);
        if (lane == 0) {
            bitmap[w] = mask;
            rank32[w] = rank[w * 32];
        }
    }
}

struct PlanEntry {
    HotPlan plan;
    int64_t nnz = 0;
    int32_t n_cols = 0;
    bool none = false;  // built and found not worth it: do not try again
};
using PlanKey = std::pair<int, const void *>;
std::mutex g_plan_mu;
std::map<PlanKey, PlanEntry> g_plans;

void entry_free(PlanEntry &e) {
    if (e.plan.Aj2) cudaFree(const_cast<int32_t *>(e.plan.Aj2));
    if (e.plan.hot_cols) cudaFree(const_cast<int32_t *>(e.plan.hot_cols));
    if (e.plan.bitmap) cudaFree(const_cast<uint32_t *>(e.plan.bitmap));
    if (e.plan.rank32) cudaFree(const_cast<uint32_t *>(e.plan.rank32));
    e = PlanEntry{};
}

int build(const int32_t *Aj, int64_t nnz, int32_t n_cols, size_t val_bytes, cudaStream_t stream, PlanEntry &e) {
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    const unsigned grid = (unsigned)di->sm_count * 8;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    SPMV_CUDA_TRY(cudaEventCreate(&t0));
    SPMV_CUDA_TRY(cudaEventCreate(&t1));
    SPMV_CUDA_TRY(cudaEventRecord(t0, stream));

    uint32_t *counts = nullptr, *rank = nullptr;
    unsigned long long *hist = nullptr;
    void *scan_tmp = nullptr;
    auto cleanup = [&]() {
        if (counts) cudaFree(counts);
        if (rank) cudaFree(rank);
        if (hist) cudaFree(hist);
        if (scan_tmp) cudaFree(scan_tmp);
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
    };
#define HOT_TRY(expr)                                                   \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) {                                        \
            record_cuda_error(_e, #expr, __FILE__, __LINE__);           \
            cleanup();                                                  \
            entry_free(e);                                              \
            return SPMVB200_ERR_CUDA;                                   \
        }                                                               \
    } while (0)
    HOT_TRY(cudaMalloc(&counts, (size_t)n_cols * 4));
    HOT_TRY(cudaMalloc(&hist, sizeof(unsigned long long) * 2 * kCountBuckets));
    HOT_TRY(cudaMemsetAsync(counts, 0, (size_t)n_cols * 4, stream));
    HOT_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * 2 * kCountBuckets, stream));
    hot_count_kernel<<<grid, 256, 0, stream>>>(Aj, nnz, counts);
    hot_hist_kernel<<<grid, 256, 0, stream>>>(counts, n_cols, hist, hist + kCountBuckets);
    count_launch(2);
    HOT_TRY(cudaGetLastError());
    unsigned long long h[2 * kCountBuckets];
    HOT_TRY(cudaMemcpyAsync(h, hist, sizeof(h), cudaMemcpyDeviceToHost, stream));
    HOT_TRY(cudaStreamSynchronize(stream));

    // threshold: as many of the most frequent columns as fit "hot_x_max_bytes" of x_hot, a column
    // seen once gaining nothing
    const int64_t k_max = option_get("hot_x_max_bytes", 32 << 20) / (int64_t)val_bytes;
    uint32_t threshold = 0;
    int64_t K = 0;
    unsigned long long mass = 0;
    {
        int64_t cols = 0;
        unsigned long long m = 0;
        for (int b = kCountBuckets - 1; b >= 2; --b) {
            if (cols + (int64_t)h[b] > k_max) break;
            cols += (int64_t)h[b];
            m += h[kCountBuckets + b];
            threshold = (uint32_t)b;
            K = cols;
            mass = m;
        }
    }
    e.nnz = nnz;
    e.n_cols = n_cols;
    // not worth a second copy of Aj unless the hot columns take a good share of the gathers
    if (K == 0 || (double)mass < 0.25 * (double)nnz) {
        cleanup();
        e.none = true;
        return SPMVB200_OK;
    }

    HOT_TRY(cudaMalloc(&rank, (size_t)n_cols * 4));
    cub::TransformInputIterator<uint32_t, IsHot, const uint32_t *> flags(counts, IsHot{threshold});
    size_t tmp_bytes = 0;
    HOT_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags, rank, (int)n_cols, stream));
    HOT_TRY(cudaMalloc(&scan_tmp, tmp_bytes ? tmp_bytes : 16));
    HOT_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, tmp_bytes, flags, rank, (int)n_cols, stream));
    int32_t *hot_cols = nullptr, *Aj2 = nullptr;
    HOT_TRY(cudaMalloc(&hot_cols, (size_t)K * 4));
    e.plan.hot_cols = hot_cols;
    HOT_TRY(cudaMalloc(&Aj2, (size_t)(nnz > 0 ? nnz : 1) * 4));
    e.plan.Aj2 = Aj2;
    {
        const int64_t words = ((int64_t)n_cols + 31) / 32;
        uint32_t *bitmap = nullptr, *rank32 = nullptr;
        HOT_TRY(cudaMalloc(&bitmap, (size_t)words * 4));
        e.plan.bitmap = bitmap;
        HOT_TRY(cudaMalloc(&rank32, (size_t)words * 4));
        e.plan.rank32 = rank32;
        hot_bitmap_kernel<<<grid, 256, 0, stream>>>(counts, rank, n_cols, threshold, bitmap, rank32);
    }
    hot_remap_table_kernel<<<grid, 256, 0, stream>>>(counts, n_cols, threshold, rank, hot_cols);
    hot_remap_kernel<<<grid, 256, 0, stream>>>(Aj, nnz, rank, Aj2);
    count_launch(4);
    HOT_TRY(cudaGetLastError());
    HOT_TRY(cudaEventRecord(t1, stream));
    HOT_TRY(cudaStreamSynchronize(stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    e.plan.K = K;
    e.plan.n_cols = n_cols;
    e.plan.threshold = threshold;
    e.plan.hot_share = (double)mass / (double)(nnz > 0 ? nnz : 1);
    e.plan.build_ms = ms;
    cleanup();
#undef HOT_TRY
    return SPMVB200_OK;
}

}  // namespace

void hot_plan_clear() {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    for (auto &kv : g_plans) entry_free(kv.second);
    g_plans.clear();
}

// The plan for (Aj, nnz, n_cols) on the current device: found, or built when `may_build`.
// *out = nullptr when there is none (not built yet, or the matrix has no hot set worth having).
int hot_plan_get(const int32_t *Aj, int64_t nnz, int32_t n_cols, size_t val_bytes, cudaStream_t stream,
                 bool may_build, const HotPlan **out) {
    *out = nullptr;
    int dev = -1;
    SPMV_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_plan_mu);
    const PlanKey key{dev, (const void *)Aj};
    auto it = g_plans.find(key);
    if (it != g_plans.end() && (it->second.nnz != nnz || it->second.n_cols != n_cols)) {
        entry_free(it->second);  // the address now holds another matrix
        g_plans.erase(it);
        it = g_plans.end();
    }
    if (it == g_plans.end()) {
        if (!may_build) return SPMVB200_OK;
        PlanEntry e;
        SPMV_TRY(build(Aj, nnz, n_cols, val_bytes, stream, e));
        it = g_plans.emplace(key, e).first;
    }
    if (!it->second.none) *out = &it->second.plan;
    return SPMVB200_OK;
}

const HotPlan *hot_plan_peek(const int32_t *Aj) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(PlanKey{dev, (const void *)Aj});
    return (it == g_plans.end() || it->second.none) ? nullptr : &it->second.plan;
}

void hot_plan_drop(const int32_t *Aj) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(PlanKey{dev, (const void *)Aj});
    if (it == g_plans.end()) return;
    entry_free(it->second);
    g_plans.erase(it);
}


}  // namespace spmvb200
