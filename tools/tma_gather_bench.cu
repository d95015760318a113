// tma_gather_bench.cu -- microbenchmark: random 4-byte gathers of x through the L1TEX path
// (ld.global.nc, what the SpMV kernels do) against the TMA path
// (cp.async.bulk.tensor.2d tile::gather4 over x viewed as rows of 16 or 32 bytes, landing in
// shared memory).  Question: can the TMA unit take gather load off L1TEX, which bounds the
// gather-heavy configurations at ~0.7 gathers / cycle / SM?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tma_gather_bench.cu -o bin/tma_gather_bench
//   ./bin/tma_gather_bench [log2_n_x=24] [row_bytes=16]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

constexpr int BLOCK = 256;
constexpr int PER_THREAD = 4;  // one gather4 per thread per round

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(BLOCK)
ldg_gather_kernel(const float *__restrict__ x, const int4 *__restrict__ idx, int64_t rounds_total, float *out) {
    float acc = 0.f;
    for (int64_t r = blockIdx.x; r < rounds_total; r += gridDim.x) {
        const int4 i = __ldg(idx + r * BLOCK + threadIdx.x);
        const float a = __ldg(x + i.x), b = __ldg(x + i.y), c = __ldg(x + i.z), d = __ldg(x + i.w);
        acc += a + b + c + d;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int ROW_FLOATS>
__global__ void __launch_bounds__(BLOCK)
tma_gather_kernel(const __grid_constant__ CUtensorMap tmap, const int4 *__restrict__ idx, int64_t rounds_total,
                  float *out) {
    constexpr int STAGES = 4;
    constexpr int ROW_BYTES = ROW_FLOATS * 4;
    constexpr int THREAD_BYTES = PER_THREAD * ROW_BYTES < 128 ? 128 : PER_THREAD * ROW_BYTES;  // TMA dst: 128-byte aligned
    constexpr int STAGE_BYTES = BLOCK * THREAD_BYTES;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar[STAGES];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(BLOCK));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    int4 pend[STAGES];
    int64_t my_rounds = (rounds_total - blockIdx.x + gridDim.x - 1) / gridDim.x;
    // software pipeline: issue round k+STAGES-1 while consuming round k
    auto issue = [&](int64_t k) {
        const int s = (int)(k % STAGES);
        const int64_t r = blockIdx.x + k * gridDim.x;
        const int4 i = __ldg(idx + r * BLOCK + tid);
        pend[s] = i;
        unsigned char *dst = smem + s * STAGE_BYTES + tid * THREAD_BYTES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])),
                     "r"(PER_THREAD * ROW_BYTES)
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(dst)),
            "l"(&tmap), "r"(0), "r"(i.x / ROW_FLOATS), "r"(i.y / ROW_FLOATS), "r"(i.z / ROW_FLOATS),
            "r"(i.w / ROW_FLOATS), "r"(smem_u32(&bar[s]))
            : "memory");
    };
    for (int64_t k = 0; k < STAGES - 1 && k < my_rounds; ++k) issue(k);
    for (int64_t k = 0; k < my_rounds; ++k) {
        if (k + STAGES - 1 < my_rounds) issue(k + STAGES - 1);
        const int s = (int)(k % STAGES);
        const uint32_t parity = (uint32_t)((k / STAGES) & 1);
        uint32_t ok = 0;
        while (!ok)
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(smem_u32(&bar[s])), "r"(parity)
                : "memory");
        const float *rows = reinterpret_cast<const float *>(smem + s * STAGE_BYTES + tid * THREAD_BYTES);
        const int4 i = pend[s];
        acc += rows[0 * ROW_FLOATS + (i.x % ROW_FLOATS)] + rows[1 * ROW_FLOATS + (i.y % ROW_FLOATS)] +
               rows[2 * ROW_FLOATS + (i.z % ROW_FLOATS)] + rows[3 * ROW_FLOATS + (i.w % ROW_FLOATS)];
        __syncthreads();  // the stage may be refilled only after everybody has read it
    }
    if (acc == 123.456f) out[0] = acc;
}



typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Mixed kernel: can the two paths run side by side?  Warps [0, 8 - TMA_WARPS) of every CTA gather
// through ld.global.nc, the others through TMA gather4 (32-byte rows, per-warp mbarriers,
// STAGES rounds in flight per warp).  A round is 128 gathers for either kind of warp; LDG warps
// take rounds [0, rounds_ldg), TMA warps rounds [rounds_ldg, rounds_total).
template <int TMA_WARPS, int STAGES>
__global__ void __launch_bounds__(BLOCK)
mixed_gather_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ x,
                    const int4 *__restrict__ idx, int64_t rounds_ldg, int64_t rounds_total, float *out) {
    constexpr int ROW_FLOATS = 8;
    constexpr int WARPS = BLOCK / 32;
    constexpr int LDG_WARPS = WARPS - TMA_WARPS;
    extern __shared__ __align__(128) unsigned char smem[];  // [TMA_WARPS][STAGES][32 lanes][128 B]
    __shared__ __align__(8) uint64_t bar[TMA_WARPS > 0 ? TMA_WARPS * STAGES : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc = 0.f;
    if (warp < LDG_WARPS) {
        const int64_t w = (int64_t)blockIdx.x * LDG_WARPS + warp, nw = (int64_t)gridDim.x * LDG_WARPS;
        for (int64_t r = w; r < rounds_ldg; r += nw) {
            const int4 i = __ldg(idx + r * 32 + lane);
            const float a = __ldg(x + i.x), b = __ldg(x + i.y), c = __ldg(x + i.z), d = __ldg(x + i.w);
            acc += a + b + c + d;
        }
    } else if (TMA_WARPS > 0) {
        const int tw = warp - LDG_WARPS;
        uint64_t *mybar = bar + tw * STAGES;
        unsigned char *mysm = smem + (size_t)tw * STAGES * 32 * 128;
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mybar[s])), "r"(32));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        const int64_t w = (int64_t)blockIdx.x * TMA_WARPS + tw, nw = (int64_t)gridDim.x * TMA_WARPS;
        const int64_t n_tma = rounds_total - rounds_ldg;
        const int64_t my_rounds = w < n_tma ? (n_tma - w + nw - 1) / nw : 0;
        int4 pend[STAGES];
        auto issue = [&](int64_t k) {
            const int s = (int)(k % STAGES);
            const int64_t r = rounds_ldg + w + k * nw;
            const int4 i = __ldg(idx + r * 32 + lane);
            pend[s] = i;
            unsigned char *dst = mysm + ((size_t)s * 32 + lane) * 128;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mybar[s])), "r"(128)
                         : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
                "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(dst)),
                "l"(&tmap), "r"(0), "r"(i.x / ROW_FLOATS), "r"(i.y / ROW_FLOATS), "r"(i.z / ROW_FLOATS),
                "r"(i.w / ROW_FLOATS), "r"(smem_u32(&mybar[s]))
                : "memory");
        };
        for (int64_t k = 0; k < STAGES - 1 && k < my_rounds; ++k) issue(k);
        for (int64_t k = 0; k < my_rounds; ++k) {
            if (k + STAGES - 1 < my_rounds) issue(k + STAGES - 1);
            const int s = (int)(k % STAGES);
            const uint32_t parity = (uint32_t)((k / STAGES) & 1);
            uint32_t ok = 0;
            while (!ok)
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(smem_u32(&mybar[s])), "r"(parity)
                    : "memory");
            const float *rows = reinterpret_cast<const float *>(mysm + ((size_t)s * 32 + lane) * 128);
            const int4 i = pend[s];
            acc += rows[0 * ROW_FLOATS + (i.x % ROW_FLOATS)] + rows[1 * ROW_FLOATS + (i.y % ROW_FLOATS)] +
                   rows[2 * ROW_FLOATS + (i.z % ROW_FLOATS)] + rows[3 * ROW_FLOATS + (i.w % ROW_FLOATS)];
            __syncwarp();  // the stage may be refilled only after the whole warp has read it
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int TMA_WARPS, int STAGES>
void run_mixed(EncodeTiledFn enc, float *x, int64_t n, const int4 *idx, int64_t gathers, int grid, double tma_share,
               float *out) {
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {8, (cuuint64_t)(n / 8)};
    cuuint64_t gstr[1] = {32};
    cuuint32_t box[2] = {8, 1};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        return;
    }
    const int64_t rounds_total = gathers / 128;
    const int64_t rounds_ldg = TMA_WARPS == 0 ? rounds_total : (int64_t)((1.0 - tma_share) * rounds_total);
    const size_t smem = (size_t)(TMA_WARPS > 0 ? TMA_WARPS : 0) * STAGES * 32 * 128;
    CK(cudaFuncSetAttribute(mixed_gather_kernel<TMA_WARPS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    mixed_gather_kernel<TMA_WARPS, STAGES><<<grid, BLOCK, smem>>>(tmap, x, idx, rounds_ldg, rounds_total, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    mixed_gather_kernel<TMA_WARPS, STAGES><<<grid, BLOCK, smem>>>(tmap, x, idx, rounds_ldg, rounds_total, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("x = %lld MB  mixed: %d of 8 warps TMA (%d stages), %4.0f%% of the gathers by TMA, grid %4d: %8.3f ms  %7.1f G gathers/s\n",
           (long long)(n * 4 >> 20), TMA_WARPS, STAGES, 100.0 * (rounds_total - rounds_ldg) / rounds_total, grid, ms,
           gathers / ms / 1e6);
}

template <int ROW_FLOATS>
double run_tma(EncodeTiledFn enc, float *x, int64_t n, const int4 *idx, int64_t rounds, int grid, float *out) {
    CUtensorMap tmap;
    cuuint64_t gdim[2] = {(cuuint64_t)ROW_FLOATS, (cuuint64_t)(n / ROW_FLOATS)};
    cuuint64_t gstr[1] = {(cuuint64_t)ROW_FLOATS * 4};
    cuuint32_t box[2] = {(cuuint32_t)ROW_FLOATS, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d (row of %d floats)\n", (int)r, ROW_FLOATS);
        return -1;
    }
    const size_t smem = 4 * BLOCK * (PER_THREAD * ROW_FLOATS * 4 < 128 ? 128 : PER_THREAD * ROW_FLOATS * 4);
    CK(cudaFuncSetAttribute(tma_gather_kernel<ROW_FLOATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    tma_gather_kernel<ROW_FLOATS><<<grid, BLOCK, smem>>>(tmap, idx, rounds, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    tma_gather_kernel<ROW_FLOATS><<<grid, BLOCK, smem>>>(tmap, idx, rounds, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main(int argc, char **argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 24;
    const int64_t n = 1ll << lg;
    const int64_t gathers = 1ll << 28;
    const int64_t rounds = gathers / (BLOCK * PER_THREAD);
    float *x, *out;
    int4 *idx;
    CK(cudaMalloc(&x, n * 4));
    CK(cudaMalloc(&out, 16));
    CK(cudaMalloc(&idx, gathers * 4));
    CK(cudaMemset(x, 0, n * 4));
    {
        std::vector<int> h((size_t)gathers);
        uint64_t s = 88172645463325252ull;
        for (auto &v : h) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            v = (int)(s % (uint64_t)n);
        }
        CK(cudaMemcpy(idx, h.data(), gathers * 4, cudaMemcpyHostToDevice));
    }
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int occ : {4, 8}) {
        ldg_gather_kernel<<<sms * occ, BLOCK>>>(x, idx, rounds, out);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        ldg_gather_kernel<<<sms * occ, BLOCK>>>(x, idx, rounds, out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("x = %lld MB  LDG gather      grid %4d: %8.3f ms  %7.1f G gathers/s\n", (long long)(n * 4 >> 20), sms * occ,
               ms, gathers / ms / 1e6);
    }
    EncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q));
    if (!enc) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    for (int occ : {1, 2, 4}) {
        double ms = run_tma<4>(enc, x, n, idx, rounds, sms * occ, out);
        if (ms > 0)
            printf("x = %lld MB  TMA gather4 16B grid %4d: %8.3f ms  %7.1f G gathers/s\n", (long long)(n * 4 >> 20),
                   sms * occ, ms, gathers / ms / 1e6);
        ms = run_tma<8>(enc, x, n, idx, rounds, sms * occ, out);
        if (ms > 0)
            printf("x = %lld MB  TMA gather4 32B grid %4d: %8.3f ms  %7.1f G gathers/s\n", (long long)(n * 4 >> 20),
                   sms * occ, ms, gathers / ms / 1e6);
    }
    // the two paths side by side
    for (int occ : {4}) {
        run_mixed<0, 2>(enc, x, n, idx, gathers, sms * occ, 0.0, out);
        for (double share : {0.125, 0.2, 0.25, 0.3}) {
            run_mixed<1, 4>(enc, x, n, idx, gathers, sms * occ, share, out);
            run_mixed<2, 2>(enc, x, n, idx, gathers, sms * occ, share, out);
            run_mixed<2, 4>(enc, x, n, idx, gathers, sms * occ, share, out);
        }
    }
    return 0;
}
