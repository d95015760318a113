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
// Ranks 0 .. K_table-1 are the plan's most frequent columns (whole count buckets from the top,
// then part of the next one, to the capacity asked for): the persistent tile kernel
// (merge.cu, merge_tile_table_kernel) keeps their x values in shared memory.  For an x that the L2
// holds anyway the plan is only that: every hot column is a table column.
// Nothing like it in the reference: its merge kernel gathers x[Aj[k]] as is
// (merge_based/agent_spmv_orig.cuh:474-506).
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <map>
#include <mutex>

#include "common.cuh"

namespace spmvb200 {

namespace {

// Occurrence counts are binned on a logarithmic scale, eight bins per octave (exact up to 15):
// bin(n) = 8 * msb(n) + the three bits below the leading one.  "count >= lower edge of bin b" is
// then the same set as "bin >= b", which is what a threshold needs.
constexpr int kCountBuckets = 256;
__host__ __device__ __forceinline__ int count_bucket(uint32_t n) {   // n >= 1
#ifdef __CUDA_ARCH__
    const int msb = 31 - __clz(n);
#else
    int msb = 0;
    while ((n >> msb) > 1u) ++msb;
#endif
    const uint32_t sub = msb >= 3 ? (n >> (msb - 3)) & 7u : (n << (3 - msb)) & 7u;
    return msb * 8 + (int)sub;
}
inline uint32_t bucket_lower_edge(int b) {
    const int msb = b / 8;
    const uint64_t m = 8u + (uint32_t)(b % 8);
    const uint64_t e = msb >= 3 ? m << (msb - 3) : (m << msb) >> 3;
    return e > 0xffffffffull ? 0xffffffffu : (uint32_t)e;
}

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
        if (n == 0u) continue;
        const int b = count_bucket(n);
        atomicAdd(&s_cols[b], 1u);
        atomicAdd(&s_mass[b], (unsigned long long)n);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kCountBuckets; i += blockDim.x) {
        if (s_cols[i]) atomicAdd(&hist_cols[i], (unsigned long long)s_cols[i]);
        if (s_mass[i]) atomicAdd(&hist_mass[i], s_mass[i]);
    }
}

// Classes of columns: 1 = table (the most frequent ones, which the persistent tile kernel keeps in
// shared memory), 2 = the other hot columns, 0 = not hot.  The table class is every column seen
// at least `table_threshold` times plus, to fill the table to its capacity, the first `part_take`
// columns (in column order) of the next lower count bucket [part_lo, table_threshold); the rest of
// that bucket is class `part_rest`.
struct InBand {
    uint32_t lo, hi;   // lo <= n < hi
    __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t &n) const {
        return (n >= lo && n < hi) ? 1u : 0u;
    }
};
struct IsClass {
    unsigned char which;
    __host__ __device__ __forceinline__ uint32_t operator()(const unsigned char &c) const { return c == which ? 1u : 0u; }
};

// part_rank: exclusive scan of InBand{part_lo, table_threshold} over the counts
__global__ void __launch_bounds__(256)
hot_class_kernel(const uint32_t *__restrict__ counts, int64_t n_cols, uint32_t threshold, uint32_t table_threshold,
                 uint32_t part_lo, uint32_t part_take, unsigned char part_rest, const uint32_t *__restrict__ part_rank,
                 unsigned char *__restrict__ cls, unsigned long long *__restrict__ table_mass) {
    unsigned long long mass = 0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t n = counts[c];
        unsigned char k = 0;
        if (n >= table_threshold) k = 1;
        else if (part_take > 0u && n >= part_lo) k = part_rank[c] < part_take ? 1 : part_rest;
        else if (n >= threshold) k = 2;
        cls[c] = k;
        if (k == 1) mass += n;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mass += __shfl_down_sync(0xffffffffu, mass, d);
    if ((threadIdx.x & 31) == 0 && mass) atomicAdd(table_mass, mass);
}

// rank_t / rank_w hold the exclusive scans of the two class flags on entry; rank_t is the remap
// table on exit (hot column -> 0x80000000 | rank, any other -> itself).  Ranks: the table class
// first (0 .. K_table-1), the other hot columns after it, both in column order.
__global__ void __launch_bounds__(256)
hot_remap_table_kernel(const unsigned char *__restrict__ cls, int64_t n_cols, uint32_t K_table,
                       uint32_t *__restrict__ rank_t, const uint32_t *__restrict__ rank_w,
                       int32_t *__restrict__ hot_cols) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += (int64_t)gridDim.x * blockDim.x) {
        const unsigned char k = cls[c];
        if (k) {
            const uint32_t r = k == 1 ? rank_t[c] : K_table + rank_w[c];
            hot_cols[r] = (int32_t)c;
            rank_t[c] = 0x80000000u | r;
        } else {
            rank_t[c] = (uint32_t)c;
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

struct PlanEntry {
    HotPlan plan;
    int64_t nnz = 0;
    int32_t n_cols = 0;
    int64_t k_max = 0, k_table = 0;  // what it was built for
    bool none = false;  // built and found not worth it: do not try again
};
using PlanKey = std::pair<int, const void *>;
std::mutex g_plan_mu;
std::map<PlanKey, PlanEntry> g_plans;

void entry_free(PlanEntry &e) {
    if (e.plan.Aj2) cudaFree(const_cast<int32_t *>(e.plan.Aj2));
    if (e.plan.hot_cols) cudaFree(const_cast<int32_t *>(e.plan.hot_cols));
    e = PlanEntry{};
}

// k_max: how many columns x_hot may hold; k_table: how many of them the shared-memory table of
// the persistent tile kernel may hold (0: no table class)
int build(const int32_t *Aj, int64_t nnz, int32_t n_cols, int64_t k_max, int64_t k_table, cudaStream_t stream,
          PlanEntry &e) {
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    const unsigned grid = (unsigned)di->sm_count * 8;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    SPMV_CUDA_TRY(cudaEventCreate(&t0));
    SPMV_CUDA_TRY(cudaEventCreate(&t1));
    SPMV_CUDA_TRY(cudaEventRecord(t0, stream));

    uint32_t *counts = nullptr, *rank_t = nullptr, *rank_w = nullptr;
    unsigned char *cls = nullptr;
    unsigned long long *hist = nullptr;
    void *scan_tmp = nullptr;
    auto cleanup = [&]() {
        if (counts) cudaFree(counts);
        if (rank_t) cudaFree(rank_t);
        if (rank_w) cudaFree(rank_w);
        if (cls) cudaFree(cls);
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
            if (_e == cudaErrorMemoryAllocation) {                      \
                /* no room for the plan (nnz * 4 bytes and its scratch): the SpMV does without */ \
                (void)cudaGetLastError();                               \
                e.nnz = nnz;                                            \
                e.n_cols = n_cols;                                      \
                e.k_max = k_max;                                        \
                e.k_table = k_table;                                    \
                e.none = true;                                          \
                return SPMVB200_OK;                                     \
            }                                                           \
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

    // thresholds: as many of the most frequent columns as fit x_hot (a column seen once gains
    // nothing), and among them as many as fit the shared-memory table: whole count buckets, then
    // part of the next one
    uint32_t threshold = 0, table_threshold = 0xffffffffu, part_lo = 0;
    int64_t K = 0, K_table = 0, part_take = 0;
    unsigned long long mass = 0;
    bool part_rest_hot = true;   // the columns of the split bucket that do not fit the table: still hot?
    {
        int64_t cols = 0;
        unsigned long long m = 0;
        bool table_open = k_table > 0;
        for (int b = kCountBuckets - 1; b >= count_bucket(2u); --b) {
            if (h[b] == 0) continue;
            if (cols + (int64_t)h[b] > k_max) {
                // a plan that is all table (k_max == k_table) still fills the table from this bucket;
                // the rest of the bucket stays cold
                if (table_open && k_table > K_table && k_max == k_table) {
                    part_lo = bucket_lower_edge(b);
                    part_take = k_table - K_table;
                    part_rest_hot = false;
                    threshold = part_lo;
                    K = cols + part_take;
                    mass = m + (unsigned long long)part_take * part_lo;   // a lower bound; the exact sum comes from the device
                }
                break;
            }
            cols += (int64_t)h[b];
            m += h[kCountBuckets + b];
            threshold = bucket_lower_edge(b);
            K = cols;
            mass = m;
            if (table_open) {
                if (cols <= k_table) {
                    table_threshold = threshold;
                    K_table = cols;
                } else {
                    part_lo = threshold;
                    part_take = k_table - K_table;
                    table_open = false;
                }
            }
        }
    }
    e.nnz = nnz;
    e.n_cols = n_cols;
    e.k_max = k_max;
    e.k_table = k_table;
    // not worth a second copy of Aj unless the hot columns take a good share of the gathers: a
    // quarter for the dense x_hot, a tenth when all of them sit in the shared-memory table
    if (K == 0 || (double)mass < (K == K_table + part_take ? 0.10 : 0.25) * (double)nnz) {
        cleanup();
        e.none = true;
        return SPMVB200_OK;
    }

    HOT_TRY(cudaMalloc(&rank_t, (size_t)n_cols * 4));
    HOT_TRY(cudaMalloc(&rank_w, (size_t)n_cols * 4));
    HOT_TRY(cudaMalloc(&cls, (size_t)n_cols));
    {
        size_t tmp_bytes = 0, tb = 0;
        cub::TransformInputIterator<uint32_t, InBand, const uint32_t *> in_part(counts, InBand{part_lo, table_threshold});
        cub::TransformInputIterator<uint32_t, IsClass, const unsigned char *> in_table(cls, IsClass{1});
        cub::TransformInputIterator<uint32_t, IsClass, const unsigned char *> in_rest(cls, IsClass{2});
        HOT_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in_part, rank_w, (int)n_cols, stream));
        HOT_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, in_table, rank_t, (int)n_cols, stream));
        if (tb > tmp_bytes) tmp_bytes = tb;
        HOT_TRY(cudaMalloc(&scan_tmp, tmp_bytes ? tmp_bytes : 16));
        if (part_take > 0) HOT_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, tmp_bytes, in_part, rank_w, (int)n_cols, stream));
        HOT_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long), stream));
        hot_class_kernel<<<grid, 256, 0, stream>>>(counts, n_cols, threshold, table_threshold, part_lo, (uint32_t)part_take,
                                                   part_rest_hot ? 2 : 0, rank_w, cls, hist);
        HOT_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, tmp_bytes, in_table, rank_t, (int)n_cols, stream));
        HOT_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, tmp_bytes, in_rest, rank_w, (int)n_cols, stream));
    }
    K_table += part_take;
    unsigned long long table_mass = 0;
    HOT_TRY(cudaMemcpyAsync(&table_mass, hist, sizeof(table_mass), cudaMemcpyDeviceToHost, stream));
    int32_t *hot_cols = nullptr, *Aj2 = nullptr;
    HOT_TRY(cudaMalloc(&hot_cols, (size_t)K * 4));
    e.plan.hot_cols = hot_cols;
    HOT_TRY(cudaMalloc(&Aj2, (size_t)(nnz > 0 ? nnz : 1) * 4));
    e.plan.Aj2 = Aj2;
    hot_remap_table_kernel<<<grid, 256, 0, stream>>>(cls, n_cols, (uint32_t)K_table, rank_t, rank_w, hot_cols);
    hot_remap_kernel<<<grid, 256, 0, stream>>>(Aj, nnz, rank_t, Aj2);
    count_launch(6);
    HOT_TRY(cudaGetLastError());
    HOT_TRY(cudaEventRecord(t1, stream));
    HOT_TRY(cudaStreamSynchronize(stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    e.plan.K = K;
    e.plan.K_table = K_table;
    e.plan.n_cols = n_cols;
    e.plan.threshold = threshold;
    if (!part_rest_hot) mass = table_mass;   // an all-table plan: the device's exact sum replaces the lower bound
    e.plan.hot_share = (double)mass / (double)(nnz > 0 ? nnz : 1);
    e.plan.table_share = (double)table_mass / (double)(nnz > 0 ? nnz : 1);
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
int hot_plan_get(const int32_t *Aj, int64_t nnz, int32_t n_cols, int64_t k_max, int64_t k_table, cudaStream_t stream,
                 bool may_build, const HotPlan **out) {
    *out = nullptr;
    int dev = -1;
    SPMV_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_plan_mu);
    const PlanKey key{dev, (const void *)Aj};
    auto it = g_plans.find(key);
    if (it != g_plans.end() && (it->second.nnz != nnz || it->second.n_cols != n_cols ||
                                (may_build && (it->second.k_max != k_max || it->second.k_table != k_table)))) {
        entry_free(it->second);  // the address now holds another matrix, or the sizes asked for changed
        g_plans.erase(it);
        it = g_plans.end();
    }
    if (it == g_plans.end()) {
        if (!may_build) return SPMVB200_OK;
        PlanEntry e;
        SPMV_TRY(build(Aj, nnz, n_cols, k_max, k_table, stream, e));
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
