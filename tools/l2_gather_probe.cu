// l2_gather_probe.cu -- where does the L2 go when SpMV gathers x?  (VERDICT r1, next #1.)
//
// The probe is SpMV with the row structure removed: every thread streams 8 consecutive column
// indices (two 128-bit loads, the Aj stream), optionally 8 values (the Ax stream), gathers
// x[col] and accumulates.  What varies:
//   * the column distribution: uniform over a footprint, uniform over ONE 32-byte sector per
//     128-byte line (a footprint in lines four times its footprint in data), or the column
//     marginal of the R-MAT generator of csrc/gen.cu (every bit 1 with probability 0.24);
//   * the gather load flavour (GMODE) and the stream load flavour (SMODE);
//   * host knobs: cudaLimitMaxL2FetchGranularity, a persisting-L2 access-policy window.
// Output: one line per experiment with the time and G gathers/s; run under
// `ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum,...` for the hit rates.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/l2_gather_probe.cu -o bin/l2_gather_probe
//   ./bin/l2_gather_probe [suite=all|cap|line|rmat|knobs|ncu] [log2_gathers=28]
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

constexpr int BLOCK = 256;
constexpr uint64_t PHI = 0x9E3779B97F4A7C15ull;
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---------------------------------------------------------------- index generators
enum Pattern { UNIFORM = 0, ONE_SECTOR_PER_LINE = 1, RMAT = 2, RMAT_SCATTER = 3, RMAT_HOTCOLD = 4 };
// UNIFORM: col uniform in [0, n).  ONE_SECTOR_PER_LINE: line uniform in [0, n/32), sector 0 of the
// line, float uniform in the sector.  RMAT: `scale` bits, each 1 with probability 0.24 (= b + d
// of (0.57, 0.19, 0.19, 0.05): the column marginal of rmat_edges_kernel, csrc/gen.cu).
// RMAT_SCATTER: the same columns with the 128-byte LINE index sent through an odd multiplier
// (a bijection that keeps every line's four sectors together and its popularity, but destroys the
// low-popcount structure of the hot addresses: tests slice / channel hot-spotting).
// RMAT_HOTCOLD: columns with popcount <= n (here: the threshold) are redirected to a dense "hot"
// array of `hot_slots` floats placed 2^27 floats into x (what a hot/cold split of x would gather).
__global__ void gen_idx_kernel(int pattern, int64_t n, int scale, uint64_t seed, int64_t count, int32_t *idx,
                               int64_t hot_slots) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t h = mix64(seed + (uint64_t)(i + 1) * PHI);
        int32_t c;
        if (pattern == UNIFORM) {
            c = (int32_t)(((h >> 32) * (uint64_t)n) >> 32);
        } else if (pattern == ONE_SECTOR_PER_LINE) {
            const uint64_t line = ((h >> 32) * (uint64_t)(n / 32)) >> 32;
            c = (int32_t)(line * 32 + (h & 7));
        } else {  // RMAT and its variants
            uint32_t col = 0;
            for (int level = 0; level < scale; ++level) {
                if ((level & 3) == 0 && level) h = mix64(h + PHI);
                const uint32_t u = (uint32_t)(h >> (16 * (level & 3))) & 0xFFFFu;
                col = (col << 1) | (u < 15729u ? 1u : 0u);  // 0.24 * 65536
            }
            c = (int32_t)col;
            if (pattern == RMAT_SCATTER) {
                const uint32_t lines = 1u << (scale - 5);
                const uint32_t line = ((col >> 5) * 0x9E3779B1u) & (lines - 1);
                c = (int32_t)((line << 5) | (col & 31u));
            } else if (pattern == RMAT_HOTCOLD && __popc(col) <= (int)n) {
                c = (int32_t)((1u << 27) + (uint32_t)(mix64(col * PHI + 12345) % (uint64_t)hot_slots));
            }
        }
        idx[i] = c;
    }
}

// ---------------------------------------------------------------- load flavours
// GMODE 9: tex1Dfetch through a texture object over x (SASS TLD: the TEX pipe of L1TEX instead
// of the LSU pipe) -- what the reference's LightSpMV does for x (LightSpMV.cuh:62-69, 286-304).
// GMODE 10: lanes alternate between the two pipes.
template <int GMODE>
__device__ __forceinline__ float gather(const float *p, uint64_t pol, cudaTextureObject_t tex = 0,
                                        const float *x0 = nullptr) {
    float v;
    if constexpr (GMODE == 9) return tex1Dfetch<float>(tex, (int)(p - x0));
    if constexpr (GMODE == 10) {
        if (threadIdx.x & 1) return tex1Dfetch<float>(tex, (int)(p - x0));
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
        return v;
    }
    if constexpr (GMODE == 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if constexpr (GMODE == 1) asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    else if constexpr (GMODE == 2) asm volatile("ld.global.nc.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if constexpr (GMODE == 3) asm volatile("ld.global.nc.L2::256B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if constexpr (GMODE == 4) asm volatile("ld.global.nc.L2::cache_hint.L2::128B.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    else if constexpr (GMODE == 5) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if constexpr (GMODE == 6) asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else if constexpr (GMODE == 7) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    else asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
static const char *kGName[] = {"nc", "nc+evict_last", "nc+L2::128B", "nc+L2::256B", "nc+evict_last+L2::128B",
                               "nc.noL1+L2::128B", "nc+L2::64B", "nc.noL1", "cg", "tex1Dfetch", "half tex / half nc"};

template <int SMODE>
__device__ __forceinline__ int4 stream4(const void *p, uint64_t pol) {
    int4 r;
    if constexpr (SMODE == 0)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    else if constexpr (SMODE == 1)
        asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if constexpr (SMODE == 2)
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else
        asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
static const char *kSName[] = {"nc.noL1+evict_first", "nc", "nc.noL1", "cs"};

// ---------------------------------------------------------------- the probe kernel
template <int GMODE, int SMODE, bool VALS, int IPT = 8>
__global__ void __launch_bounds__(BLOCK)
probe_kernel(const float *__restrict__ x, const int32_t *__restrict__ idx, const float *__restrict__ vals,
             int64_t count, float *out, cudaTextureObject_t tex) {
    uint64_t pol_first, pol_last;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    if constexpr (IPT == 4) {   // half the gathers in flight per thread
        const int64_t g = ((int64_t)blockIdx.x * BLOCK + threadIdx.x) * 4;
        if (g + 4 > count) return;
        const int4 a = stream4<SMODE>(idx + g, pol_first);
        const float v0 = gather<GMODE>(x + a.x, pol_last, tex, x), v1 = gather<GMODE>(x + a.y, pol_last, tex, x),
                    v2 = gather<GMODE>(x + a.z, pol_last, tex, x), v3 = gather<GMODE>(x + a.w, pol_last, tex, x);
        const float acc = v0 + v1 + v2 + v3;
        if (acc == 123.456f) out[0] = acc;
        return;
    }
    const int64_t g = ((int64_t)blockIdx.x * BLOCK + threadIdx.x) * 8;
    if (g + 8 > count) return;
    const int4 a = stream4<SMODE>(idx + g, pol_first), b = stream4<SMODE>(idx + g + 4, pol_first);
    float w[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if constexpr (VALS) {
        const int4 va = stream4<SMODE>(vals + g, pol_first), vb = stream4<SMODE>(vals + g + 4, pol_first);
        w[0] = __int_as_float(va.x); w[1] = __int_as_float(va.y); w[2] = __int_as_float(va.z); w[3] = __int_as_float(va.w);
        w[4] = __int_as_float(vb.x); w[5] = __int_as_float(vb.y); w[6] = __int_as_float(vb.z); w[7] = __int_as_float(vb.w);
    }
    const int c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = gather<GMODE>(x + c[k], pol_last, tex, x);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k] * w[k];
    if (acc == 123.456f) out[0] = acc;
}

struct Ctx {
    float *x = nullptr;      // 1 GB
    int32_t *idx = nullptr;  // count
    float *vals = nullptr;   // count
    float *out = nullptr;
    float *flush = nullptr;
    int64_t count = 0;
    cudaStream_t stream;
    cudaTextureObject_t tex = 0;
};

template <int G, int S, bool V, int IPT = 8>
static float run_one(const Ctx &c, int reps = 3, int carveout = -1) {
    const unsigned grid = (unsigned)((c.count / IPT + BLOCK - 1) / BLOCK);
    CK(cudaFuncSetAttribute(probe_kernel<G, S, V, IPT>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps + 1; ++r) {
        CK(cudaMemsetAsync(c.flush, 0, 512ull << 20, c.stream));  // L2 cold
        CK(cudaEventRecord(e0, c.stream));
        probe_kernel<G, S, V, IPT><<<grid, BLOCK, 0, c.stream>>>(c.x, c.idx, c.vals, c.count, c.out, c.tex);
        CK(cudaEventRecord(e1, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

using RunFn = float (*)(const Ctx &, int);
template <int S, bool V>
static float run_g(int g, const Ctx &c) {
    switch (g) {
        case 0: return run_one<0, S, V>(c);
        case 1: return run_one<1, S, V>(c);
        case 2: return run_one<2, S, V>(c);
        case 3: return run_one<3, S, V>(c);
        case 4: return run_one<4, S, V>(c);
        case 5: return run_one<5, S, V>(c);
        case 6: return run_one<6, S, V>(c);
        case 7: return run_one<7, S, V>(c);
        default: return run_one<8, S, V>(c);
    }
}
static float run(int g, int s, bool v, const Ctx &c) {
    if (v) {
        switch (s) {
            case 0: return run_g<0, true>(g, c);
            case 1: return run_g<1, true>(g, c);
            case 2: return run_g<2, true>(g, c);
            default: return run_g<3, true>(g, c);
        }
    }
    switch (s) {
        case 0: return run_g<0, false>(g, c);
        case 1: return run_g<1, false>(g, c);
        case 2: return run_g<2, false>(g, c);
        default: return run_g<3, false>(g, c);
    }
}

static void gen(const Ctx &c, int pattern, int64_t n, int scale, int64_t hot_slots = 1) {
    // the R-MAT variants share one seed: the same column draws, relocated
    const uint64_t seed = pattern >= RMAT ? 0x5eedull + 154 : 0x5eedull + (uint64_t)pattern * 77 + (uint64_t)n;
    gen_idx_kernel<<<148 * 8, 256, 0, c.stream>>>(pattern, n, scale, seed, c.count, c.idx, hot_slots);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c.stream));
}
static void report(const char *what, int g, int s, bool v, const Ctx &c, float ms) {
    printf("%-44s gather=%-24s stream=%-20s vals=%d  %8.3f ms  %7.1f G gathers/s\n", what, kGName[g], kSName[s], (int)v,
           ms, (double)c.count / ms * 1e-6);
    fflush(stdout);
}

int main(int argc, char **argv) {
    const std::string suite = argc > 1 ? argv[1] : "all";
    const int lg = argc > 2 ? atoi(argv[2]) : 28;
    Ctx c;
    c.count = 1ll << lg;
    CK(cudaSetDevice(0));
    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&c.x, 1ull << 30));
    CK(cudaMemset(c.x, 0, 1ull << 30));
    CK(cudaMalloc(&c.idx, (size_t)c.count * 4));
    CK(cudaMalloc(&c.vals, (size_t)c.count * 4));
    CK(cudaMemset(c.vals, 0, (size_t)c.count * 4));
    CK(cudaMalloc(&c.out, 64));
    CK(cudaMalloc(&c.flush, 512ull << 20));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    {   // texture object over the first 2^27 floats of x (the most a 1-D linear texture takes here)
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof rd);
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = c.x;
        rd.res.linear.desc = cudaCreateChannelDesc<float>();
        rd.res.linear.sizeInBytes = std::min((size_t)1 << 29, (size_t)prop.maxTexture1DLinear * 4);
        cudaTextureDesc td;
        memset(&td, 0, sizeof td);
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&c.tex, &rd, &td, nullptr));
        printf("# maxTexture1DLinear %d elements\n", prop.maxTexture1DLinear);
    }
    size_t gran = 0;
    CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
    printf("# %s, %d SMs, L2 %.1f MB, persisting max %.1f MB, window max %.1f MB, default max L2 fetch granularity %zu B, %lld gathers per run\n",
           prop.name, prop.multiProcessorCount, prop.l2CacheSize / 1048576.0, prop.persistingL2CacheMaxSize / 1048576.0,
           prop.accessPolicyMaxWindowSize / 1048576.0, gran, (long long)c.count);
    const bool all = suite == "all";
    char what[128];

    if (all || suite == "cap") {
        printf("# capacity: uniform columns over a footprint (index stream only), L2 flushed before each run\n");
        for (int mb : {16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256, 384, 512, 1024}) {
            gen(c, UNIFORM, (int64_t)mb << 18, 0);
            snprintf(what, sizeof what, "uniform %4d MB", mb);
            report(what, 0, 0, false, c, run(0, 0, false, c));
        }
    }
    if (all || suite == "line") {
        printf("# line granularity: ONE sector per 128-byte line; MB = footprint in lines (data touched = MB/4)\n");
        for (int mb : {32, 64, 96, 128, 160, 192, 256, 384, 512, 1024}) {
            gen(c, ONE_SECTOR_PER_LINE, (int64_t)mb << 18, 0);
            snprintf(what, sizeof what, "1-sector-per-line %4d MB of lines", mb);
            report(what, 0, 0, false, c, run(0, 0, false, c));
        }
        printf("# the same with 128-byte L2 prefetch (whole line fetched on a miss)\n");
        for (int mb : {64, 128, 256, 512}) {
            gen(c, ONE_SECTOR_PER_LINE, (int64_t)mb << 18, 0);
            snprintf(what, sizeof what, "1-sector-per-line %4d MB of lines", mb);
            report(what, 2, 0, false, c, run(2, 0, false, c));
        }
    }
    if (all || suite == "rmat") {
        printf("# R-MAT column marginal (bit = 1 w.p. 0.24), x = 4 << scale bytes; every gather flavour; index stream only\n");
        for (int scale : {24, 25, 26, 27}) {
            gen(c, RMAT, 0, scale);
            for (int g = 0; g < 9; ++g) {
                snprintf(what, sizeof what, "rmat scale %d (x %4d MB)", scale, 4 << (scale - 20));
                report(what, g, 0, false, c, run(g, 0, false, c));
            }
        }
        printf("# scale 27 with the value stream as well (8 B streamed per gather, as SpMV), stream flavours\n");
        gen(c, RMAT, 0, 27);
        for (int s = 0; s < 4; ++s)
            for (int g : {0, 1, 2, 4}) {
                report("rmat scale 27 + values", g, s, true, c, run(g, s, true, c));
            }
        printf("# uniform 512 MB, gather flavours (is a 128-byte miss as cheap as a 32-byte miss?)\n");
        gen(c, UNIFORM, 512ll << 18, 0);
        for (int g : {0, 2, 3, 6}) report("uniform  512 MB", g, 0, false, c, run(g, 0, false, c));
    }
    if (all || suite == "knobs") {
        printf("# host knobs on rmat scale 27 + values\n");
        gen(c, RMAT, 0, 27);
        for (size_t gsz : {(size_t)32, (size_t)64, (size_t)128}) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gsz);
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            snprintf(what, sizeof what, "maxL2FetchGranularity %zu (%s, reads back %zu)", gsz, cudaGetErrorName(e), got);
            report(what, 0, 0, true, c, run(0, 0, true, c));
            report(what, 2, 0, true, c, run(2, 0, true, c));
        }
        CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
        // persisting window over the head of x: the top column bits are 0 w.p. 0.76 each, so the
        // first 1/8 of x (64 MB) receives 0.76^3 = 44 % of the gathers
        for (int set_mb : {32, 64, 78}) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)set_mb << 20);
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
            for (int win_mb : {32, 64, 128}) {
                cudaStreamAttrValue v;
                memset(&v, 0, sizeof v);
                v.accessPolicyWindow.base_ptr = c.x;
                v.accessPolicyWindow.num_bytes = (size_t)win_mb << 20;
                v.accessPolicyWindow.hitRatio = std::min(1.0f, (float)got / (float)((size_t)win_mb << 20));
                v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
                cudaError_t e2 = cudaStreamSetAttribute(c.stream, cudaStreamAttributeAccessPolicyWindow, &v);
                snprintf(what, sizeof what, "persist %d MB (%s, got %.0f), window %d MB (%s)", set_mb, cudaGetErrorName(e),
                         got / 1048576.0, win_mb, cudaGetErrorName(e2));
                report(what, 0, 0, true, c, run(0, 0, true, c));
            }
        }
        cudaStreamAttrValue v;
        memset(&v, 0, sizeof v);
        cudaStreamSetAttribute(c.stream, cudaStreamAttributeAccessPolicyWindow, &v);
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    if (all || suite == "v2" || suite == "ncu2") {
        printf("# what bounds the scale-27 gathers?  (index stream only, plain ld.global.nc)\n");
        gen(c, RMAT, 0, 27);
        report("rmat 27 (reference point)", 0, 0, false, c, run_one<0, 0, false>(c));
        if (suite != "ncu2") {
            report("rmat 27, 4 gathers per thread", 0, 0, false, c, run_one<0, 0, false, 4>(c));
            for (int cv : {0, 25, 50, 75, 100}) {
                snprintf(what, sizeof what, "rmat 27, shared-memory carveout %d %%", cv);
                report(what, 0, 0, false, c, run_one<0, 0, false>(c, 3, cv));
            }
        }
        gen(c, RMAT_SCATTER, 0, 27);
        report("rmat 27, lines scattered (odd multiplier)", 0, 0, false, c, run_one<0, 0, false>(c));
        // hot/cold split: columns of popcount <= th in a dense array (sizes: exact column counts)
        const int ths[] = {6, 7, 8, 9, 10};
        const int64_t cnts[] = {397594, 1285624, 3505699, 8192524, 16628809};
        for (int i = 0; i < 5; ++i) {
            gen(c, RMAT_HOTCOLD, ths[i], 27, cnts[i]);
            snprintf(what, sizeof what, "rmat 27, popcount<=%d hot (%.1f MB dense)", ths[i], cnts[i] * 4.0 / 1048576.0);
            report(what, 0, 0, false, c, run_one<0, 0, false>(c));
            if (suite == "ncu2" && i != 2 && i != 3) continue;
            report(what, 0, 0, true, c, run_one<0, 0, true>(c));
        }
        gen(c, RMAT, 0, 26);
        report("rmat 26 (reference point)", 0, 0, false, c, run_one<0, 0, false>(c));
        gen(c, UNIFORM, 512ll << 18, 0);
        report("uniform 512 MB", 0, 0, false, c, run_one<0, 0, false>(c));
        if (suite != "ncu2") {
            for (size_t gsz : {(size_t)32, (size_t)128}) {
                cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gsz);
                snprintf(what, sizeof what, "uniform 512 MB, maxL2FetchGranularity %zu", gsz);
                report(what, 0, 0, false, c, run_one<0, 0, false>(c));
            }
            CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
        }
    }
    if (all || suite == "tex") {
        printf("# LSU pipe (ld.global.nc) against TEX pipe (tex1Dfetch) for the same gathers\n");
        struct { int pattern; int64_t n; int scale; const char *name; } cases[] = {
            {UNIFORM, 16ll << 18, 0, "uniform 16 MB"}, {UNIFORM, 64ll << 18, 0, "uniform 64 MB"},
            {RMAT, 0, 24, "rmat 24"}, {RMAT, 0, 26, "rmat 26"}, {RMAT, 0, 27, "rmat 27"}, {UNIFORM, 512ll << 18, 0, "uniform 512 MB"}};
        for (auto &cs : cases) {
            gen(c, cs.pattern, cs.n, cs.scale);
            report(cs.name, 0, 0, false, c, run_one<0, 0, false>(c));
            report(cs.name, 9, 0, false, c, run_one<9, 0, false>(c));
            report(cs.name, 10, 0, false, c, run_one<10, 0, false>(c));
            report(cs.name, 9, 0, true, c, run_one<9, 0, true>(c));
        }
    }
    if (suite == "ncu") {
        // the short list wrapped in ncu: hit rates and DRAM bytes per flavour
        printf("# ncu list: launches in this order (4 launches each: 1 warm + 3)\n");
        gen(c, RMAT, 0, 27);
        for (int g : {0, 1, 2, 4, 3}) report("rmat scale 27, index stream only", g, 0, false, c, run(g, 0, false, c));
        for (int g : {0, 2}) report("rmat scale 27 + values", g, 0, true, c, run(g, 0, true, c));
        gen(c, RMAT, 0, 24);
        for (int g : {0, 2}) report("rmat scale 24, index stream only", g, 0, false, c, run(g, 0, false, c));
        gen(c, UNIFORM, 64ll << 18, 0);
        report("uniform 64 MB", 0, 0, false, c, run(0, 0, false, c));
        gen(c, UNIFORM, 512ll << 18, 0);
        for (int g : {0, 2}) report("uniform 512 MB", g, 0, false, c, run(g, 0, false, c));
        gen(c, ONE_SECTOR_PER_LINE, 128ll << 18, 0);
        report("1-sector-per-line 128 MB of lines", 0, 0, false, c, run(0, 0, false, c));
    }
    return 0;
}
