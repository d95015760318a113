// gen.cu -- device-side data layer: synthetic matrix generators and a 64-bit-safe stable
// COO -> CSR conversion.
//
// The reference has no generator (its inputs are .mtx files, reference/README.md:24-26) and
// builds CSR on one host thread with index_t counters (reference/include/load.hpp:420-474,
// which breaks at nnz >= 2^31).  BASELINE.json's configurations are synthetic
// shapes up to 2^31 nonzeros, so they are generated where they are used: on the device,
// from a counter-based RNG (splitmix64) that oracle/generators.py restates on the host bit
// for bit.
//
//   key  = mix64(seed ^ (stream * PHI));   draw(ctr) = mix64(key + (ctr + 1) * PHI)
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace spmvb200 {

namespace {

constexpr uint64_t PHI = 0x9E3779B97F4A7C15ull;
enum { STREAM_VAL = 1, STREAM_X = 2, STREAM_COL = 3, STREAM_RMAT = 4 };

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t stream_key(uint64_t seed, uint32_t stream) { return mix64(seed ^ ((uint64_t)stream * PHI)); }
__device__ __forceinline__ uint64_t draw(uint64_t key, uint64_t ctr) { return mix64(key + (ctr + 1) * PHI); }

template <typename ValT> __device__ __forceinline__ ValT pm1(uint64_t h);
template <> __device__ __forceinline__ float pm1<float>(uint64_t h) {
    return (float)((int32_t)(h >> 40) - (1 << 23)) * 1.1920928955078125e-07f;  // 2^-23
}
template <> __device__ __forceinline__ double pm1<double>(uint64_t h) {
    return (double)((int64_t)(h >> 11) - (1ll << 52)) * 2.220446049250313e-16;  // 2^-52
}

template <typename ValT>
__global__ void __launch_bounds__(256)
uniform_pm1_kernel(uint64_t key, uint64_t first, int64_t count, ValT *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = pm1<ValT>(draw(key, first + (uint64_t)i));
}

// 5-point Laplacian on an n x n grid, row-major, columns ascending, values 4 / -1.
// nnz before row r in closed form: 5r minus the neighbours missing so far.
__host__ __device__ __forceinline__ int64_t lap2d_offset(int64_t r, int64_t n) {
    const int64_t up = r < n ? r : n;
    const int64_t down = r - n * (n - 1) > 0 ? r - n * (n - 1) : 0;
    return 5 * r - up - down - (r + n - 1) / n - r / n;
}
template <typename OffT, typename ValT>
__global__ void __launch_bounds__(256)
lap2d_kernel(int32_t n, OffT *__restrict__ Ap, int32_t *__restrict__ Aj, ValT *__restrict__ Ax) {
    const int64_t N = (int64_t)n * n;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= N;
         r += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = lap2d_offset(r, n);
        Ap[r] = (OffT)k;
        if (r == N) continue;
        const int64_t i = r / n, j = r % n;
        if (i > 0) { Aj[k] = (int32_t)(r - n); Ax[k] = (ValT)-1; ++k; }
        if (j > 0) { Aj[k] = (int32_t)(r - 1); Ax[k] = (ValT)-1; ++k; }
        Aj[k] = (int32_t)r; Ax[k] = (ValT)4; ++k;
        if (j < n - 1) { Aj[k] = (int32_t)(r + 1); Ax[k] = (ValT)-1; ++k; }
        if (i < n - 1) { Aj[k] = (int32_t)(r + n); Ax[k] = (ValT)-1; ++k; }
    }
}

// K columns per row, one per stratum of width n_cols / K: distinct and ascending.
template <typename OffT, typename ValT>
__global__ void __launch_bounds__(256)
uniform_rows_kernel(int32_t n_rows, int32_t K, uint32_t stratum, uint64_t key_col, uint64_t key_val,
                    OffT *__restrict__ Ap, int32_t *__restrict__ Aj, ValT *__restrict__ Ax) {
    const int64_t nnz = (int64_t)n_rows * K;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nnz;
         c += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t h = draw(key_col, (uint64_t)c);
        const uint32_t jitter = (uint32_t)(((h >> 32) * (uint64_t)stratum) >> 32);
        const int64_t k = c % K;
        Aj[c] = (int32_t)(k * stratum + jitter);
        Ax[c] = pm1<ValT>(draw(key_val, (uint64_t)c));
        if (k == 0) Ap[c / K] = (OffT)c;
        if (c == nnz - 1) Ap[n_rows] = (OffT)nnz;
    }
}

// R-MAT (a,b,c,d) = (0.57,0.19,0.19,0.05) as 24-bit thresholds; `scale` quadrant draws per
// edge, most significant bit first, two 24-bit draws per hash.
constexpr uint32_t RMAT_A = 9563013;    // round(0.57 * 2^24)
constexpr uint32_t RMAT_AB = 12750684;  // round(0.76 * 2^24)
constexpr uint32_t RMAT_ABC = 15938355; // round(0.95 * 2^24)
__global__ void __launch_bounds__(256)
rmat_edges_kernel(int32_t scale, uint64_t key, uint64_t first, int64_t count,
                  int32_t *__restrict__ rows, int32_t *__restrict__ cols) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t e = first + (uint64_t)i;
        uint32_t row = 0, col = 0;
        uint64_t h = 0;
        for (int level = 0; level < scale; ++level) {
            uint32_t u;
            if ((level & 1) == 0) {
                h = draw(key, e * 16 + (uint64_t)(level >> 1));
                u = (uint32_t)(h >> 40);
            } else {
                u = (uint32_t)(h >> 16) & 0xFFFFFFu;
            }
            const uint32_t rb = u >= RMAT_AB;
            const uint32_t cb = u < RMAT_A ? 0u : (u < RMAT_AB ? 1u : (u < RMAT_ABC ? 0u : 1u));
            row = (row << 1) | rb;
            col = (col << 1) | cb;
        }
        rows[i] = (int32_t)row;
        cols[i] = (int32_t)col;
    }
}

// Ap[r] = first position in the sorted row array holding a value >= r (lower bound)
template <typename OffT>
__global__ void __launch_bounds__(256)
offsets_from_sorted_rows_kernel(int32_t n_rows, int64_t nnz, const int32_t *__restrict__ sorted_rows,
                                OffT *__restrict__ Ap) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)__ldg(sorted_rows + mid) < r) lo = mid + 1;
            else hi = mid;
        }
        Ap[r] = (OffT)lo;
    }
}

// values follow their edge through the sort as a permutation index would be 8 B/edge; instead
// sort (row, original position) is avoided too: values are gathered by a second sort pass
// only when the caller supplies them.
template <typename ValT>
__global__ void __launch_bounds__(256)
gather_vals_kernel(int64_t nnz, const int64_t *__restrict__ perm, const ValT *__restrict__ vals,
                   ValT *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = vals[perm[i]];
}
__global__ void __launch_bounds__(256) iota_kernel(int64_t n, int64_t *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i;
}

inline unsigned grid_for(int64_t n, int sm_count) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

int bits_for(int64_t n_rows) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n_rows) ++b;
    return b;
}

}  // namespace

template <typename ValT>
int gen_uniform_pm1(uint64_t seed, uint32_t stream_id, uint64_t first, int64_t count, ValT *out,
                    cudaStream_t stream) {
    if (count <= 0) return SPMVB200_OK;
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    uniform_pm1_kernel<ValT><<<grid_for(count, di->sm_count), 256, 0, stream>>>(
        stream_key(seed, stream_id), first, count, out);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

template <typename OffT, typename ValT>
int gen_lap2d(int32_t n, OffT *Ap, int32_t *Aj, ValT *Ax, cudaStream_t stream) {
    if (n <= 0) return SPMVB200_ERR_INVALID;
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    lap2d_kernel<OffT, ValT><<<grid_for((int64_t)n * n + 1, di->sm_count), 256, 0, stream>>>(n, Ap, Aj, Ax);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

template <typename OffT, typename ValT>
int gen_uniform_rows(int32_t n_rows, int32_t n_cols, int32_t K, uint64_t seed, OffT *Ap, int32_t *Aj,
                     ValT *Ax, cudaStream_t stream) {
    if (n_rows <= 0 || K <= 0 || n_cols <= 0 || n_cols % K != 0) return SPMVB200_ERR_INVALID;
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    uniform_rows_kernel<OffT, ValT><<<grid_for((int64_t)n_rows * K, di->sm_count), 256, 0, stream>>>(
        n_rows, K, (uint32_t)(n_cols / K), stream_key(seed, STREAM_COL), stream_key(seed, STREAM_VAL),
        Ap, Aj, Ax);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

int gen_rmat_edges(int32_t scale, uint64_t seed, uint64_t first, int64_t count, int32_t *rows,
                   int32_t *cols, cudaStream_t stream) {
    if (scale < 1 || scale > 31) return SPMVB200_ERR_INVALID;
    if (count <= 0) return SPMVB200_OK;
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    rmat_edges_kernel<<<grid_for(count, di->sm_count), 256, 0, stream>>>(
        scale, stream_key(seed, STREAM_RMAT), first, count, rows, cols);
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

// Stable LSD radix sort of (row, col) by row with CUB (64-bit item counts), then offsets by
// lower-bound search.  rows/cols are used as one half of the sort's double buffers.
template <typename OffT, typename ValT>
int coo_to_csr(int32_t n_rows, int64_t nnz, int32_t *rows, int32_t *cols, const ValT *vals, OffT *Ap,
               int32_t *Aj, ValT *Ax, cudaStream_t stream) {
    if (n_rows < 0 || nnz < 0) return SPMVB200_ERR_INVALID;
    if (sizeof(OffT) == 4 && nnz > 0x7fffffffLL) return SPMVB200_ERR_INVALID;
    const DeviceInfo *di = nullptr;
    SPMV_TRY(current_device_info(&di));
    const int end_bit = bits_for(n_rows > 1 ? n_rows : 2);
    int32_t *rows_alt = nullptr;
    void *temp = nullptr;
    int64_t *perm = nullptr, *perm_alt = nullptr;
    int status = SPMVB200_OK;
    auto cleanup = [&]() {
        if (rows_alt) cudaFree(rows_alt);
        if (temp) cudaFree(temp);
        if (perm) cudaFree(perm);
        if (perm_alt) cudaFree(perm_alt);
    };
#define GEN_TRY(expr)                                                   \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) {                                        \
            record_cuda_error(_e, #expr, __FILE__, __LINE__);           \
            cleanup();                                                  \
            return SPMVB200_ERR_CUDA;                                   \
        }                                                               \
    } while (0)
    if (nnz > 0) {
        GEN_TRY(cudaMalloc(&rows_alt, (size_t)nnz * sizeof(int32_t)));
        size_t temp_bytes = 0;
        if (vals == nullptr) {
            // pattern: sort (row -> col); Aj doubles as the alternate value buffer
            cub::DoubleBuffer<int32_t> k(rows, rows_alt);
            cub::DoubleBuffer<int32_t> v(cols, Aj);
            GEN_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, k, v, nnz, 0, end_bit, stream));
            GEN_TRY(cudaMalloc(&temp, temp_bytes ? temp_bytes : 16));
            GEN_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, nnz, 0, end_bit, stream));
            count_launch(4);
            if (v.Current() != Aj)
                GEN_TRY(cudaMemcpyAsync(Aj, v.Current(), (size_t)nnz * sizeof(int32_t),
                                        cudaMemcpyDeviceToDevice, stream));
            offsets_from_sorted_rows_kernel<OffT><<<grid_for((int64_t)n_rows + 1, di->sm_count), 256, 0, stream>>>(
                n_rows, nnz, k.Current(), Ap);
            count_launch();
            GEN_TRY(cudaGetLastError());
        } else {
            // with values: sort (row -> original position), then gather cols and vals
            GEN_TRY(cudaMalloc(&perm, (size_t)nnz * sizeof(int64_t)));
            GEN_TRY(cudaMalloc(&perm_alt, (size_t)nnz * sizeof(int64_t)));
            iota_kernel<<<grid_for(nnz, di->sm_count), 256, 0, stream>>>(nnz, perm);
            count_launch();
            cub::DoubleBuffer<int32_t> k(rows, rows_alt);
            cub::DoubleBuffer<int64_t> v(perm, perm_alt);
            GEN_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, k, v, nnz, 0, end_bit, stream));
            GEN_TRY(cudaMalloc(&temp, temp_bytes ? temp_bytes : 16));
            GEN_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k, v, nnz, 0, end_bit, stream));
            count_launch(4);
            gather_vals_kernel<int32_t><<<grid_for(nnz, di->sm_count), 256, 0, stream>>>(nnz, v.Current(), cols, Aj);
            gather_vals_kernel<ValT><<<grid_for(nnz, di->sm_count), 256, 0, stream>>>(nnz, v.Current(), vals, Ax);
            offsets_from_sorted_rows_kernel<OffT><<<grid_for((int64_t)n_rows + 1, di->sm_count), 256, 0, stream>>>(
                n_rows, nnz, k.Current(), Ap);
            count_launch(3);
            GEN_TRY(cudaGetLastError());
        }
    } else {
        GEN_TRY(cudaMemsetAsync(Ap, 0, ((size_t)n_rows + 1) * sizeof(OffT), stream));
    }
    GEN_TRY(cudaStreamSynchronize(stream));
#undef GEN_TRY
    cleanup();
    return status;
}

// explicit instantiations used by api.cu
template int gen_uniform_pm1<float>(uint64_t, uint32_t, uint64_t, int64_t, float *, cudaStream_t);
template int gen_uniform_pm1<double>(uint64_t, uint32_t, uint64_t, int64_t, double *, cudaStream_t);
#define INST(OffT, ValT)                                                                           \
    template int gen_lap2d<OffT, ValT>(int32_t, OffT *, int32_t *, ValT *, cudaStream_t);          \
    template int gen_uniform_rows<OffT, ValT>(int32_t, int32_t, int32_t, uint64_t, OffT *,         \
                                              int32_t *, ValT *, cudaStream_t);                    \
    template int coo_to_csr<OffT, ValT>(int32_t, int64_t, int32_t *, int32_t *, const ValT *,      \
                                        OffT *, int32_t *, ValT *, cudaStream_t);
INST(int32_t, float)
INST(int32_t, double)
INST(int64_t, float)
INST(int64_t, double)
#undef INST

}  // namespace spmvb200
