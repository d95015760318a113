// mcast.cu -- memory replicated over the GPUs of one process and bound to ONE NVLink multicast
// object (NVLS): a store to the multicast address lands in every GPU's replica, replicated by the
// NVSwitch, so the sender's link carries each value once instead of once per peer.
//
// Why the power iteration needs it at 8 GPUs: rows are split by nonzeros, so on a skewed matrix
// the GPU with the short rows owns most of the ROWS (R-MAT scale 27: 54 M of 134 M), and with
// plain peer stores it sends them 7 times -- 1.5 GB per step over a 900 GB/s link, 1.7 ms, longer
// than its SpMV (1.45 ms).  Through the switch it is 0.2 GB.
//
// The driver API (cuMulticast*, cuMem*) is reached through cudaGetDriverEntryPoint, so the
// library has no link-time dependency on libcuda and still loads on a machine without a driver.
#include <cuda.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace spmvb200 {

struct McastArena {
    size_t bytes = 0;                                  // per replica, rounded up to the granularity
    std::vector<int> dev;
    CUmemGenericAllocationHandle mc_handle = 0;
    std::vector<CUmemGenericAllocationHandle> mem;     // one physical allocation per GPU
    std::vector<CUdeviceptr> va;                       // where each GPU's replica is mapped
    std::vector<char> bound;
    CUdeviceptr mc_va = 0;
    bool mc_mapped = false;
};

namespace {

struct DriverApi {
    bool ok = false;
    CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                                 unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
};

template <typename F>
bool entry(const char *name, F *fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        (void)cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

const DriverApi &driver() {
    static DriverApi d = [] {
        DriverApi a;
        a.ok = entry("cuDeviceGet", &a.DeviceGet) && entry("cuDeviceGetAttribute", &a.DeviceGetAttribute) &&
               entry("cuMulticastCreate", &a.MulticastCreate) && entry("cuMulticastAddDevice", &a.MulticastAddDevice) &&
               entry("cuMulticastBindMem", &a.MulticastBindMem) && entry("cuMulticastUnbind", &a.MulticastUnbind) &&
               entry("cuMulticastGetGranularity", &a.MulticastGetGranularity) && entry("cuMemCreate", &a.MemCreate) &&
               entry("cuMemRelease", &a.MemRelease) &&
               entry("cuMemGetAllocationGranularity", &a.MemGetAllocationGranularity) &&
               entry("cuMemAddressReserve", &a.MemAddressReserve) && entry("cuMemAddressFree", &a.MemAddressFree) &&
               entry("cuMemMap", &a.MemMap) && entry("cuMemUnmap", &a.MemUnmap) && entry("cuMemSetAccess", &a.MemSetAccess);
        return a;
    }();
    return d;
}

}  // namespace

void mcast_arena_destroy(McastArena *a) {
    if (!a) return;
    const DriverApi &D = driver();
    if (D.ok) {
        if (a->mc_mapped) D.MemUnmap(a->mc_va, a->bytes);
        if (a->mc_va) D.MemAddressFree(a->mc_va, a->bytes);
        for (size_t g = 0; g < a->dev.size(); ++g) {
            CUdevice cd = 0;
            if (g < a->bound.size() && a->bound[g] && D.DeviceGet(&cd, a->dev[g]) == CUDA_SUCCESS)
                D.MulticastUnbind(a->mc_handle, cd, 0, a->bytes);
            if (g < a->va.size() && a->va[g]) {
                D.MemUnmap(a->va[g], a->bytes);
                D.MemAddressFree(a->va[g], a->bytes);
            }
            if (g < a->mem.size() && a->mem[g]) D.MemRelease(a->mem[g]);
        }
        if (a->mc_handle) D.MemRelease(a->mc_handle);
    }
    delete a;
}

// `bytes` per GPU on each of devices[0..n): replica g at replica[g], all of them behind *mc.
// SPMVB200_ERR_UNSUPPORTED when the driver or a device has no multicast (no NVSwitch): the caller
// falls back to peer stores.  Every device must already have its primary context (the caller has
// run something on each).
int mcast_arena_create(const int *devices, int n, size_t bytes, McastArena **out, void **replica, void **mc) {
    *out = nullptr;
    const DriverApi &D = driver();
    if (!D.ok || n < 2) return SPMVB200_ERR_UNSUPPORTED;
    std::vector<CUdevice> cd((size_t)n);
    for (int g = 0; g < n; ++g) {
        int has = 0;
        if (D.DeviceGet(&cd[(size_t)g], devices[g]) != CUDA_SUCCESS ||
            D.DeviceGetAttribute(&has, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd[(size_t)g]) != CUDA_SUCCESS || !has)
            return SPMVB200_ERR_UNSUPPORTED;
    }
    McastArena *a = new (std::nothrow) McastArena;
    if (!a) return SPMVB200_ERR_INVALID;
    a->dev.assign(devices, devices + n);
    a->mem.assign((size_t)n, 0);
    a->va.assign((size_t)n, 0);
    a->bound.assign((size_t)n, 0);
#define MC_TRY(expr)                                                                   \
    do {                                                                               \
        const CUresult _r = (expr);                                                    \
        if (_r != CUDA_SUCCESS) {                                                      \
            record_cuda_error(cudaErrorUnknown, #expr, __FILE__, __LINE__);            \
            mcast_arena_destroy(a);                                                    \
            return _r == CUDA_ERROR_NOT_SUPPORTED || _r == CUDA_ERROR_NOT_PERMITTED    \
                       ? SPMVB200_ERR_UNSUPPORTED                                      \
                       : SPMVB200_ERR_CUDA;                                            \
        }                                                                              \
    } while (0)

    CUmulticastObjectProp mprop = {};
    mprop.numDevices = (unsigned)n;
    mprop.size = bytes;
    mprop.handleTypes = 0;
    size_t gran = 0;
    MC_TRY(D.MulticastGetGranularity(&gran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    for (int g = 0; g < n; ++g) {   // the physical allocations have a granularity of their own
        CUmemAllocationProp ap = {};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = devices[g];
        size_t ag = 0;
        MC_TRY(D.MemGetAllocationGranularity(&ag, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
        if (ag > gran) gran = ag;
    }
    a->bytes = (bytes + gran - 1) / gran * gran;
    mprop.size = a->bytes;
    MC_TRY(D.MulticastCreate(&a->mc_handle, &mprop));
    for (int g = 0; g < n; ++g) MC_TRY(D.MulticastAddDevice(a->mc_handle, cd[(size_t)g]));   // all, before any bind

    std::vector<CUmemAccessDesc> access((size_t)n);
    for (int g = 0; g < n; ++g) {
        access[(size_t)g].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        access[(size_t)g].location.id = devices[g];
        access[(size_t)g].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    }
    for (int g = 0; g < n; ++g) {
        CUmemAllocationProp ap = {};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = devices[g];
        MC_TRY(D.MemCreate(&a->mem[(size_t)g], a->bytes, &ap, 0));
        MC_TRY(D.MulticastBindMem(a->mc_handle, 0, a->mem[(size_t)g], 0, a->bytes, 0));
        a->bound[(size_t)g] = 1;
        MC_TRY(D.MemAddressReserve(&a->va[(size_t)g], a->bytes, gran, 0, 0));
        MC_TRY(D.MemMap(a->va[(size_t)g], a->bytes, 0, a->mem[(size_t)g], 0));
        MC_TRY(D.MemSetAccess(a->va[(size_t)g], a->bytes, access.data(), (size_t)n));   // peers store into the mailboxes
        replica[g] = reinterpret_cast<void *>(a->va[(size_t)g]);
    }
    MC_TRY(D.MemAddressReserve(&a->mc_va, a->bytes, gran, 0, 0));
    MC_TRY(D.MemMap(a->mc_va, a->bytes, 0, a->mc_handle, 0));
    a->mc_mapped = true;
    MC_TRY(D.MemSetAccess(a->mc_va, a->bytes, access.data(), (size_t)n));
#undef MC_TRY
    *mc = reinterpret_cast<void *>(a->mc_va);
    *out = a;
    return SPMVB200_OK;
}

}  // namespace spmvb200
