// multi.cu -- the row-sharded power iteration of BASELINE.json's last configuration, driven from
// ONE host process over the GPUs of a box, behind the C ABI: no Python, no torch.distributed.
//
// (spmv_samples_b200/dist.py is the one-process-per-GPU form bench.py needs; this is what a C++
// caller such as main.cu --gpus N links against.  The reference is single-device: main.cu:53.)
//
//   * rows are cut at the merge path's nnz-balanced boundaries (spmvb200_row_split_*), each GPU
//     gets its row block of A (offsets rebased) by peer copies, and two full-length replicas of x;
//   * a step is, per GPU and without the host waiting for anything: the SpMV of the local rows
//     (kind "auto"), whose row stores go into the local replica of the next x AND, through peer
//     access over NVLink, into every other GPU's replica -- the all-gather is the kernel's
//     epilogue; then spmvb200_norm_exchange: sum of squares, exchange of the per-GPU sums through
//     peer-mapped mailboxes, alpha = 1/||A x||, which is also the step barrier;
//   * x is double-buffered, the scale is applied by the next SpMV through its device alpha, and
//     the host only enqueues: one thread, GPU after GPU, step after step.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

struct spmvb200_power {
    struct Gpu {
        int dev = -1;
        cudaStream_t stream = nullptr;
        void *Ap = nullptr, *Ax = nullptr;
        int32_t *Aj = nullptr;
        int64_t row_begin = 0, rows = 0, nnz = 0;
        void *xbuf[2] = {nullptr, nullptr};
        double *sumsq = nullptr;
        void *alpha = nullptr;
        int *error = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        int64_t calls = 0;  // SpMVs issued on this GPU's row block since it was uploaded
    };
    std::vector<Gpu> gpu;
    std::vector<int64_t> row_bounds;
    int offset_bits = 0, value_bits = 0, kind = SPMVB200_KIND_AUTO;
    int64_t n = 0, nnz = 0;
    size_t tail_off = 0;  // bytes from a replica's base to its mailbox
    uint64_t step = 0;
    // NVLink multicast (mcast.cu): when set, the replicas of x live in its arena and mc_buf[b] is
    // the multicast address of buffer b; otherwise they are cudaMalloc'ed and fed by peer stores
    spmvb200::McastArena *mc = nullptr;
    void *mc_buf[2] = {nullptr, nullptr};
};

namespace spmvb200 {
namespace {

template <typename OffT>
__global__ void __launch_bounds__(256)
rebase_offsets_kernel(int64_t count, OffT base, OffT *__restrict__ Ap) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        Ap[i] -= base;
}
template <typename ValT>
__global__ void __launch_bounds__(256) fill_kernel(int64_t count, ValT v, ValT *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = v;
}

struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int fill_values(int value_bits, int64_t count, double v, void *out, cudaStream_t s) {
    if (count <= 0) return SPMVB200_OK;
    const unsigned grid = (unsigned)((count + 255) / 256 < 4096 ? (count + 255) / 256 : 4096);
    if (value_bits == 32) fill_kernel<float><<<grid, 256, 0, s>>>(count, (float)v, static_cast<float *>(out));
    else fill_kernel<double><<<grid, 256, 0, s>>>(count, v, static_cast<double *>(out));
    SPMV_LAUNCH_CHECK();
    return SPMVB200_OK;
}

}  // namespace
}  // namespace spmvb200

using namespace spmvb200;

extern "C" {

void spmvb200_power_destroy(spmvb200_power_t *p) {
    if (!p) return;
    DeviceGuard guard;
    for (auto &g : p->gpu) {
        if (g.dev < 0 || cudaSetDevice(g.dev) != cudaSuccess) continue;
        if (g.stream) cudaStreamSynchronize(g.stream);
        if (g.Aj) hot_plan_drop(g.Aj);
        for (void *q : {g.Ap, (void *)g.Aj, g.Ax, (void *)g.sumsq, g.alpha, (void *)g.error})
            if (q) cudaFree(q);
        if (!p->mc)
            for (void *q : g.xbuf)
                if (q) cudaFree(q);
        if (g.ev0) cudaEventDestroy(g.ev0);
        if (g.ev1) cudaEventDestroy(g.ev1);
        if (g.stream) cudaStreamDestroy(g.stream);
    }
    mcast_arena_destroy(p->mc);
    delete p;
}

int spmvb200_power_create_from_device(int n_gpus, const int *devices, int offset_bits, int value_bits,
                                      int64_t n_rows, int64_t nnz, const void *Ap, const int32_t *Aj,
                                      const void *Ax, int kind, spmvb200_power_t **out) {
    if (!out || n_gpus < 1 || n_gpus > 8 || n_rows < 1 || n_rows > 0x7fffffffLL || nnz < 0 || !Ap ||
        (nnz > 0 && (!Aj || !Ax)))
        return SPMVB200_ERR_INVALID;
    if ((offset_bits != 32 && offset_bits != 64) || (value_bits != 32 && value_bits != 64))
        return SPMVB200_ERR_UNSUPPORTED;
    DeviceGuard guard;
    spmvb200_power *p = new (std::nothrow) spmvb200_power;
    if (!p) return SPMVB200_ERR_INVALID;
    p->offset_bits = offset_bits;
    p->value_bits = value_bits;
    p->kind = kind;
    p->n = n_rows;
    p->nnz = nnz;
    p->gpu.resize((size_t)n_gpus);
    for (int g = 0; g < n_gpus; ++g) p->gpu[(size_t)g].dev = devices ? devices[g] : g;
    const size_t ob = (size_t)offset_bits / 8, vb = (size_t)value_bits / 8;
    p->tail_off = ((size_t)n_rows * vb + 255) / 256 * 256;
    const int src_dev = p->gpu[0].dev;
    int status = SPMVB200_OK;
#define POWER_TRY(expr)                                                \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) {                                       \
            record_cuda_error(_e, #expr, __FILE__, __LINE__);          \
            spmvb200_power_destroy(p);                                 \
            return SPMVB200_ERR_CUDA;                                  \
        }                                                              \
    } while (0)
#define POWER_ST(expr)                                                 \
    do {                                                               \
        status = (expr);                                               \
        if (status != SPMVB200_OK) {                                   \
            spmvb200_power_destroy(p);                                 \
            return status;                                             \
        }                                                              \
    } while (0)

    // ---- the split (device search on the GPU that holds the matrix), and Ap at the boundaries
    POWER_TRY(cudaSetDevice(src_dev));
    p->row_bounds.assign((size_t)n_gpus + 1, 0);
    if (n_gpus > 1) {
        if (offset_bits == 32)
            POWER_ST(spmvb200_row_split_o32((int32_t)n_rows, (int32_t)nnz, static_cast<const int32_t *>(Ap), n_gpus,
                                            p->row_bounds.data(), nullptr));
        else
            POWER_ST(spmvb200_row_split_o64((int32_t)n_rows, nnz, static_cast<const int64_t *>(Ap), n_gpus,
                                            p->row_bounds.data(), nullptr));
    } else {
        p->row_bounds[1] = n_rows;
    }
    std::vector<int64_t> k_at((size_t)n_gpus + 1, 0);
    for (int g = 0; g <= n_gpus; ++g) {
        if (offset_bits == 32) {
            int32_t v = 0;
            POWER_TRY(cudaMemcpy(&v, static_cast<const int32_t *>(Ap) + p->row_bounds[(size_t)g], 4, cudaMemcpyDeviceToHost));
            k_at[(size_t)g] = v;
        } else {
            POWER_TRY(cudaMemcpy(&k_at[(size_t)g], static_cast<const int64_t *>(Ap) + p->row_bounds[(size_t)g], 8,
                                 cudaMemcpyDeviceToHost));
        }
    }

    // ---- peer access all to all (UVA: a peer's cudaMalloc pointer is then usable as is)
    for (int a = 0; a < n_gpus; ++a) {
        POWER_TRY(cudaSetDevice(p->gpu[(size_t)a].dev));
        for (int b = 0; b < n_gpus; ++b) {
            if (a == b) continue;
            int can = 0;
            POWER_TRY(cudaDeviceCanAccessPeer(&can, p->gpu[(size_t)a].dev, p->gpu[(size_t)b].dev));
            if (!can) {
                spmvb200_power_destroy(p);
                return SPMVB200_ERR_UNSUPPORTED;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(p->gpu[(size_t)b].dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) POWER_TRY(e);
            (void)cudaGetLastError();
        }
    }

    // ---- per GPU: row block of A, replicas of x, scalars
    for (int g = 0; g < n_gpus; ++g) {
        auto &G = p->gpu[(size_t)g];
        POWER_TRY(cudaSetDevice(G.dev));
        POWER_TRY(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
        POWER_TRY(cudaEventCreate(&G.ev0));
        POWER_TRY(cudaEventCreate(&G.ev1));
        G.row_begin = p->row_bounds[(size_t)g];
        G.rows = p->row_bounds[(size_t)g + 1] - G.row_begin;
        const int64_t k0 = k_at[(size_t)g];
        G.nnz = k_at[(size_t)g + 1] - k0;
        POWER_TRY(cudaMalloc(&G.Ap, (size_t)(G.rows + 1) * ob));
        POWER_TRY(cudaMalloc((void **)&G.Aj, (size_t)(G.nnz > 0 ? G.nnz : 1) * 4));
        POWER_TRY(cudaMalloc(&G.Ax, (size_t)(G.nnz > 0 ? G.nnz : 1) * vb));
        POWER_TRY(cudaMemcpyPeerAsync(G.Ap, G.dev, static_cast<const char *>(Ap) + (size_t)G.row_begin * ob, src_dev,
                                      (size_t)(G.rows + 1) * ob, G.stream));
        if (G.nnz > 0) {
            POWER_TRY(cudaMemcpyPeerAsync(G.Aj, G.dev, Aj + k0, src_dev, (size_t)G.nnz * 4, G.stream));
            POWER_TRY(cudaMemcpyPeerAsync(G.Ax, G.dev, static_cast<const char *>(Ax) + (size_t)k0 * vb, src_dev,
                                          (size_t)G.nnz * vb, G.stream));
        }
        const unsigned grid = (unsigned)((G.rows + 256) / 256 < 4096 ? (G.rows + 256) / 256 : 4096);
        if (offset_bits == 32)
            rebase_offsets_kernel<int32_t><<<grid, 256, 0, G.stream>>>(G.rows + 1, (int32_t)k0, static_cast<int32_t *>(G.Ap));
        else
            rebase_offsets_kernel<int64_t><<<grid, 256, 0, G.stream>>>(G.rows + 1, k0, static_cast<int64_t *>(G.Ap));
        count_launch();
        POWER_TRY(cudaGetLastError());
        POWER_TRY(cudaMalloc((void **)&G.sumsq, sizeof(double)));
        POWER_TRY(cudaMalloc(&G.alpha, 8));
        POWER_TRY(cudaMalloc((void **)&G.error, sizeof(int)));
        POWER_TRY(cudaMemsetAsync(G.error, 0, sizeof(int), G.stream));
    }
    // ---- the replicas of x.  Option "power_exchange": 0 = peer stores, 1 = NVLink multicast,
    // -1 = multicast where the box has it.  Rows are split by nonzeros, so on a skewed matrix one
    // GPU owns most of the rows and with peer stores sends them n_gpus - 1 times; at 8 GPUs that
    // transfer outlasts its SpMV (mcast.cu).  R-MAT scale 27, ms per step at 2 / 4 / 8 GPUs:
    // 5.71 / 2.93 / 1.60 with multicast, 5.88 / 3.09 / 2.63 with peer stores.
    const size_t stride = p->tail_off + SPMVB200_MAILBOX_BYTES;
    const int64_t xopt = option_get("power_exchange", -1);
    if (n_gpus > 1 && xopt != 0) {
        std::vector<int> devs((size_t)n_gpus);
        std::vector<void *> base((size_t)n_gpus, nullptr);
        for (int g = 0; g < n_gpus; ++g) devs[(size_t)g] = p->gpu[(size_t)g].dev;
        void *mc_base = nullptr;
        status = mcast_arena_create(devs.data(), n_gpus, 2 * stride, &p->mc, base.data(), &mc_base);
        if (status == SPMVB200_OK) {
            for (int b = 0; b < 2; ++b) {
                p->mc_buf[b] = static_cast<char *>(mc_base) + (size_t)b * stride;
                for (int g = 0; g < n_gpus; ++g)
                    p->gpu[(size_t)g].xbuf[b] = static_cast<char *>(base[(size_t)g]) + (size_t)b * stride;
            }
        } else if (status != SPMVB200_ERR_UNSUPPORTED || xopt > 0) {
            spmvb200_power_destroy(p);   // asked for by option and not available, or a real failure
            return status;
        }
    }
    if (!p->mc)
        for (auto &G : p->gpu) {
            POWER_TRY(cudaSetDevice(G.dev));
            for (int b = 0; b < 2; ++b) POWER_TRY(cudaMalloc(&G.xbuf[b], stride));
        }
    *out = p;
    status = spmvb200_power_reset(p);
    if (status != SPMVB200_OK) {
        spmvb200_power_destroy(p);
        *out = nullptr;
    }
    return status;
#undef POWER_TRY
#undef POWER_ST
}

int spmvb200_power_create(int n_gpus, const int *devices, int offset_bits, int value_bits, int64_t n_rows,
                          int64_t nnz, const void *Ap_host, const int32_t *Aj_host, const void *Ax_host,
                          int kind, spmvb200_power_t **out) {
    if (!out || n_gpus < 1 || n_rows < 1 || nnz < 0 || !Ap_host || (nnz > 0 && (!Aj_host || !Ax_host)))
        return SPMVB200_ERR_INVALID;
    if ((offset_bits != 32 && offset_bits != 64) || (value_bits != 32 && value_bits != 64))
        return SPMVB200_ERR_UNSUPPORTED;
    DeviceGuard guard;
    SPMV_CUDA_TRY(cudaSetDevice(devices ? devices[0] : 0));
    const size_t ob = (size_t)offset_bits / 8, vb = (size_t)value_bits / 8;
    void *Ap = nullptr, *Ax = nullptr;
    int32_t *Aj = nullptr;
    int status = SPMVB200_OK;
    do {
        cudaError_t e;
        if ((e = cudaMalloc(&Ap, (size_t)(n_rows + 1) * ob)) != cudaSuccess ||
            (e = cudaMalloc((void **)&Aj, (size_t)(nnz > 0 ? nnz : 1) * 4)) != cudaSuccess ||
            (e = cudaMalloc(&Ax, (size_t)(nnz > 0 ? nnz : 1) * vb)) != cudaSuccess ||
            (e = cudaMemcpy(Ap, Ap_host, (size_t)(n_rows + 1) * ob, cudaMemcpyHostToDevice)) != cudaSuccess ||
            (nnz > 0 && ((e = cudaMemcpy(Aj, Aj_host, (size_t)nnz * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
                         (e = cudaMemcpy(Ax, Ax_host, (size_t)nnz * vb, cudaMemcpyHostToDevice)) != cudaSuccess))) {
            record_cuda_error(e, "spmvb200_power_create: staging the matrix", __FILE__, __LINE__);
            status = SPMVB200_ERR_CUDA;
            break;
        }
        status = spmvb200_power_create_from_device(n_gpus, devices, offset_bits, value_bits, n_rows, nnz, Ap, Aj, Ax,
                                                   kind, out);
        if (status == SPMVB200_OK) status = spmvb200_power_sync(*out);
    } while (false);
    if (Ap) cudaFree(Ap);
    if (Aj) cudaFree(Aj);
    if (Ax) cudaFree(Ax);
    return status;
}

// x0 = 1/sqrt(n), alpha = 1, every mailbox empty; the step counter restarts
int spmvb200_power_reset(spmvb200_power_t *p) {
    if (!p) return SPMVB200_ERR_INVALID;
    DeviceGuard guard;
    SPMV_TRY(spmvb200_power_sync(p));
    const size_t vb = (size_t)p->value_bits / 8;
    for (auto &G : p->gpu) {
        SPMV_CUDA_TRY(cudaSetDevice(G.dev));
        SPMV_TRY(fill_values(p->value_bits, p->n, 1.0 / std::sqrt((double)p->n), G.xbuf[0], G.stream));
        SPMV_CUDA_TRY(cudaMemsetAsync(G.xbuf[1], 0, (size_t)p->n * vb, G.stream));
        SPMV_TRY(fill_values(64, SPMVB200_MAILBOX_BYTES / 8, -1.0, static_cast<char *>(G.xbuf[0]) + p->tail_off, G.stream));
        SPMV_TRY(fill_values(p->value_bits, 1, 1.0, G.alpha, G.stream));
        SPMV_CUDA_TRY(cudaMemsetAsync(G.sumsq, 0, sizeof(double), G.stream));
    }
    p->step = 0;
    return spmvb200_power_sync(p);   // nobody may publish before every mailbox is empty
}

int spmvb200_power_sync(spmvb200_power_t *p) {
    if (!p) return SPMVB200_ERR_INVALID;
    DeviceGuard guard;
    for (auto &G : p->gpu) {
        SPMV_CUDA_TRY(cudaSetDevice(G.dev));
        if (G.stream) SPMV_CUDA_TRY(cudaStreamSynchronize(G.stream));
    }
    return SPMVB200_OK;
}

// enqueue `steps` steps on every GPU (the host does not wait)
int spmvb200_power_steps(spmvb200_power_t *p, int steps) {
    if (!p || steps < 0) return SPMVB200_ERR_INVALID;
    DeviceGuard guard;
    const int P = (int)p->gpu.size();
    const size_t vb = (size_t)p->value_bits / 8;
    for (int s = 0; s < steps; ++s, ++p->step) {
        const int cur = (int)(p->step & 1), nxt = cur ^ 1;
        for (int g = 0; g < P; ++g) {
            auto &G = p->gpu[(size_t)g];
            SPMV_CUDA_TRY(cudaSetDevice(G.dev));
            void *peers[kMaxPeers] = {};
            void *mailboxes[8] = {};
            int np = 0;
            for (int q = 0; q < P; ++q) {
                mailboxes[q] = static_cast<char *>(p->gpu[(size_t)q].xbuf[0]) + p->tail_off;
                if (q != g && !p->mc)
                    peers[np++] = static_cast<char *>(p->gpu[(size_t)q].xbuf[nxt]) + (size_t)G.row_begin * vb;
            }
            if (p->mc) {   // one multimem.st per row, replicated by the switch (spmv_b200.h: n_peers == -1)
                peers[0] = static_cast<char *>(p->mc_buf[nxt]) + (size_t)G.row_begin * vb;
                np = -1;
            }
            spmvb200_args_t a;
            std::memset(&a, 0, sizeof(a));
            a.kind = p->kind;
            a.offset_bits = p->offset_bits;
            a.value_bits = p->value_bits;
            a.n_rows = G.rows;
            a.n_cols = p->n;
            a.nnz = G.nnz;
            a.Ap = G.Ap;
            a.Aj = G.Aj;
            a.Ax = G.Ax;
            a.x = G.xbuf[cur];
            a.y = static_cast<char *>(G.xbuf[nxt]) + (size_t)G.row_begin * vb;
            a.alpha_dev = G.alpha;
            a.y_peers = peers;
            a.n_peers = np;
            a.stream = G.stream;
            // the row block is resident and never changes: every call after the first vouches for
            // it (a reset restarts the iteration, not the matrix)
            a.flags = G.calls++ > 0 ? SPMVB200_FLAG_STATIC_PATTERN : 0;
            if (G.rows > 0) SPMV_TRY(spmvb200_spmv(&a));
            // Rows without nonzeros are never sent to the peers, so their entries must already be 0
            // in every replica: buffer 1 starts zeroed; buffer 0 held x0 and is cleared after the
            // first step's kernel has read it and before the exchange lets a peer store into it.
            if (p->step == 0 && P > 1) SPMV_CUDA_TRY(cudaMemsetAsync(G.xbuf[0], 0, (size_t)p->n * vb, G.stream));
            SPMV_TRY(spmvb200_norm_exchange(p->value_bits, G.rows, a.y, g, P, p->step, mailboxes[g], mailboxes,
                                            p->mc ? static_cast<char *>(p->mc_buf[0]) + p->tail_off : nullptr,
                                            G.sumsq, G.alpha, G.error, G.stream));
        }
    }
    return SPMVB200_OK;
}

// `steps` steps timed on the devices: events on every GPU's stream around its part, the slowest
// GPU counts.  Synchronises before and after.
int spmvb200_power_run(spmvb200_power_t *p, int steps, double *ms_per_step) {
    if (!p || steps < 1) return SPMVB200_ERR_INVALID;
    DeviceGuard guard;
    SPMV_TRY(spmvb200_power_sync(p));
    for (auto &G : p->gpu) {
        SPMV_CUDA_TRY(cudaSetDevice(G.dev));
        SPMV_CUDA_TRY(cudaEventRecord(G.ev0, G.stream));
    }
    SPMV_TRY(spmvb200_power_steps(p, steps));
    for (auto &G : p->gpu) {
        SPMV_CUDA_TRY(cudaSetDevice(G.dev));
        SPMV_CUDA_TRY(cudaEventRecord(G.ev1, G.stream));
    }
    SPMV_TRY(spmvb200_power_sync(p));
    double worst = 0.0;
    for (auto &G : p->gpu) {
        SPMV_CUDA_TRY(cudaSetDevice(G.dev));
        float ms = 0.f;
        SPMV_CUDA_TRY(cudaEventElapsedTime(&ms, G.ev0, G.ev1));
        if (ms > worst) worst = ms;
        int err = 0;
        SPMV_CUDA_TRY(cudaMemcpy(&err, G.error, sizeof(int), cudaMemcpyDeviceToHost));
        if (err) return SPMVB200_ERR_CUDA;   // a GPU did not arrive at an exchange
    }
    if (ms_per_step) *ms_per_step = worst / steps;
    return SPMVB200_OK;
}

// how the replicas of x are fed: 0 = peer stores (or one GPU), 1 = NVLink multicast
int spmvb200_power_exchange(const spmvb200_power_t *p) { return p && p->mc ? 1 : 0; }

// the current iterate (n values, not yet scaled by 1/||.||) and ||A x_k|| of the last step
int spmvb200_power_get(spmvb200_power_t *p, void *x_host, double *norm, int64_t *row_bounds) {
    if (!p) return SPMVB200_ERR_INVALID;
    DeviceGuard guard;
    SPMV_TRY(spmvb200_power_sync(p));
    auto &G = p->gpu[0];
    SPMV_CUDA_TRY(cudaSetDevice(G.dev));
    if (x_host)
        SPMV_CUDA_TRY(cudaMemcpy(x_host, G.xbuf[p->step & 1], (size_t)p->n * (size_t)p->value_bits / 8,
                                 cudaMemcpyDeviceToHost));
    if (norm) {
        double s = 0.0;
        SPMV_CUDA_TRY(cudaMemcpy(&s, G.sumsq, sizeof(double), cudaMemcpyDeviceToHost));
        *norm = std::sqrt(s);
    }
    if (row_bounds)
        for (size_t i = 0; i < p->row_bounds.size(); ++i) row_bounds[i] = p->row_bounds[i];
    return SPMVB200_OK;
}

}  // extern "C"
