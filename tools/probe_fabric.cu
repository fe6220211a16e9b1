// tools/probe_fabric.cu -- what the replicate step can use on this box, measured (not assumed):
//   1. does the driver offer NVSwitch multicast objects (cuMulticast*), and does `multimem.st` land in every
//      member's memory;
//   2. store bandwidth of ONE GPU into its peers as a function of CTA count, for the three candidate primitives
//      of the leader's T5 step: per-thread 16 B `st.global.v4` (repeated per follower), `multimem.st.v4` on a
//      multicast mapping (one store, the switch fans out), and `cp.async.bulk` shared -> peer global;
//   3. can another process' file descriptor be taken with pidfd_getfd (the handle exchange a one-process-per-GPU
//      deployment needs for VMM allocations; cudaIpc* does not cover them).
// Build: make -C tools   Run: tools/probe_fabric [mode]   (each mode in its own process: a fault must not hide the rest)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/syscall.h>
#include <unistd.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char *s = nullptr; cuGetErrorString(e, &s); printf("driver error %d (%s) at line %d: %s\n", (int)e, s ? s : "?", __LINE__, #x); exit(3);} } while (0)

__device__ __forceinline__ void st16(void *p, uint4 v) { asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void mst16(void *p, uint4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                 "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}

// mode 0: st.v4 to each of nd destinations; mode 1: multimem.st to dst[0] (a multicast address)
__global__ void k_stream(uint8_t *d0, uint8_t *d1, uint8_t *d2, uint8_t *d3, uint8_t *d4, uint8_t *d5, int nd, size_t bytes, int mode, uint32_t seed)
{
    uint8_t *dst[6] = {d0, d1, d2, d3, d4, d5};
    const size_t nchunks = bytes / 16;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = make_uint4((uint32_t)c ^ seed, (uint32_t)(c >> 32) + seed, seed * 3u + 1u, (uint32_t)c * 2654435761u);
        if (mode == 1) mst16(dst[0] + 16 * c, v);
        else
            for (int f = 0; f < nd; f++) st16(dst[f] + 16 * c, v);
    }
}

// mode 2: each CTA fills a 32 KiB shared tile and pushes it with cp.async.bulk (shared -> global) to every destination
__global__ void k_bulk(uint8_t *d0, uint8_t *d1, uint8_t *d2, uint8_t *d3, uint8_t *d4, uint8_t *d5, int nd, size_t bytes, uint32_t seed)
{
    extern __shared__ __align__(128) uint8_t tile[];
    uint8_t *dst[6] = {d0, d1, d2, d3, d4, d5};
    const uint32_t TB = 32768;
    const size_t ntiles = bytes / TB;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (uint32_t c = threadIdx.x; c < TB / 16; c += blockDim.x) {
            const size_t gc = t * (TB / 16) + c;
            reinterpret_cast<uint4 *>(tile)[c] = make_uint4((uint32_t)gc ^ seed, (uint32_t)(gc >> 32) + seed, seed * 3u + 1u, (uint32_t)gc * 2654435761u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(tile);
            for (int f = 0; f < nd; f++)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst[f] + t * TB), "r"(s), "r"(TB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the tile may be overwritten
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void k_check(const uint8_t *p, size_t bytes, uint32_t seed, unsigned long long *bad)
{
    const size_t nchunks = bytes / 16;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4 *>(p)[c];
        const uint4 w = make_uint4((uint32_t)c ^ seed, (uint32_t)(c >> 32) + seed, seed * 3u + 1u, (uint32_t)c * 2654435761u);
        if (v.x != w.x || v.y != w.y || v.z != w.z || v.w != w.w) atomicAdd(bad, 1ull);
    }
}

static size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

struct Fabric {
    int nd = 0;
    size_t size = 0;
    std::vector<CUmemGenericAllocationHandle> mem;
    std::vector<CUdeviceptr> uc;       // unicast mapping of member i's memory (accessible from every device)
    CUmemGenericAllocationHandle mc = 0;
    CUdeviceptr mcptr = 0;
    bool have_mc = false;
};

static void vmm_setup(Fabric &F, int nd, size_t want, bool multicast)
{
    F.nd = nd;
    CUmemAllocationProp ap;
    memset(&ap, 0, sizeof ap);
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    ap.location.id = 0;
    CU(cuMemGetAllocationGranularity(&gran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof mp);
    size_t mgran = gran;
    if (multicast) {
        mp.numDevices = nd; mp.size = want; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CU(cuMulticastGetGranularity(&mgran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
        size_t mmin = 0;
        CU(cuMulticastGetGranularity(&mmin, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
        printf("granularity: alloc %zu, multicast min %zu recommended %zu\n", gran, mmin, mgran);
    }
    const size_t g = gran > mgran ? gran : mgran;
    F.size = round_up(want, g);
    if (multicast) {
        mp.size = F.size;
        CU(cuMulticastCreate(&F.mc, &mp));
        for (int d = 0; d < nd; d++) CU(cuMulticastAddDevice(F.mc, d));
    }
    std::vector<CUmemAccessDesc> acc(nd);
    for (int d = 0; d < nd; d++) { acc[d].location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc[d].location.id = d; acc[d].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; }
    F.mem.resize(nd); F.uc.resize(nd);
    for (int d = 0; d < nd; d++) {
        ap.location.id = d;
        CU(cuMemCreate(&F.mem[d], F.size, &ap, 0));
        CU(cuMemAddressReserve(&F.uc[d], F.size, g, 0, 0));
        CU(cuMemMap(F.uc[d], F.size, 0, F.mem[d], 0));
        CU(cuMemSetAccess(F.uc[d], F.size, acc.data(), nd));
        if (multicast) CU(cuMulticastBindMem(F.mc, 0, F.mem[d], 0, F.size, 0));
    }
    if (multicast) {
        CU(cuMemAddressReserve(&F.mcptr, F.size, g, 0, 0));
        CU(cuMemMap(F.mcptr, F.size, 0, F.mc, 0));
        CU(cuMemSetAccess(F.mcptr, F.size, acc.data(), nd));
        F.have_mc = true;
    }
}

static unsigned long long check(Fabric &F, int d, size_t bytes, uint32_t seed)
{
    CK(cudaSetDevice(d));
    unsigned long long *bad;
    CK(cudaMallocManaged(&bad, 8));
    *bad = 0;
    k_check<<<296, 512>>>((const uint8_t *)F.uc[d], bytes, seed, bad);
    CK(cudaDeviceSynchronize());
    unsigned long long b = *bad;
    cudaFree(bad);
    return b;
}

static double timed(int mode, Fabric &F, int first_dst, int ndst, size_t bytes, int ctas, uint32_t seed)
{
    CK(cudaSetDevice(0));
    uint8_t *d[6] = {0, 0, 0, 0, 0, 0};
    if (mode == 1) d[0] = (uint8_t *)F.mcptr;
    else for (int f = 0; f < ndst; f++) d[f] = (uint8_t *)F.uc[first_dst + f];
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        CK(cudaEventRecord(e0));
        if (mode == 2) k_bulk<<<ctas, 256, 32768>>>(d[0], d[1], d[2], d[3], d[4], d[5], ndst, bytes, seed);
        else k_stream<<<ctas, 512>>>(d[0], d[1], d[2], d[3], d[4], d[5], ndst, bytes, mode, seed);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it && ms < best) best = ms;
    }
    return (double)bytes / (best * 1e-3) / 1e9;        // GB/s of SOURCE bytes (one copy of the range)
}

int main(int argc, char **argv)
{
    const char *mode = argc > 1 ? argv[1] : "info";
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    CU(cuInit(0));
    for (int d = 0; d < nd; d++) { CK(cudaSetDevice(d)); CK(cudaFree(0)); }
    if (!strcmp(mode, "info")) {
        for (int d = 0; d < nd; d++) {
            int mc = 0, vmm = 0, fd = 0, fabric = 0;
            CU(cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d));
            CU(cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, d));
            CU(cuDeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, d));
            CU(cuDeviceGetAttribute(&fabric, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, d));
            printf("device %d: multicast %d, vmm %d, posix-fd handles %d, fabric handles %d\n", d, mc, vmm, fd, fabric);
        }
        // pidfd_getfd on ourselves (the cross-process form needs the same ptrace permission process_vm_writev does)
        int pfd = (int)syscall(SYS_pidfd_open, getpid(), 0);
        int got = pfd >= 0 ? (int)syscall(438 /* SYS_pidfd_getfd */, pfd, 1, 0) : -1;
        printf("pidfd_open %s, pidfd_getfd %s\n", pfd >= 0 ? "ok" : "FAILED", got >= 0 ? "ok" : "FAILED");
        return 0;
    }
    if (nd < 2) { printf("needs >= 2 GPUs\n"); return 0; }
    for (int a = 0; a < nd; a++) {
        CK(cudaSetDevice(a));
        for (int b = 0; b < nd; b++) if (a != b) { cudaError_t e = cudaDeviceEnablePeerAccess(b, 0); if (e != cudaSuccess) cudaGetLastError(); }
    }
    const size_t bytes = 256ull << 20;
    Fabric F;
    const bool want_mc = !strcmp(mode, "mc") || !strcmp(mode, "mcbulk");
    vmm_setup(F, nd, bytes, want_mc);
    printf("%d devices, %zu MiB per member, multicast %s\n", nd, F.size >> 20, F.have_mc ? "bound" : "not used");
    const int followers = nd - 1 > 6 ? 6 : nd - 1;
    if (!strcmp(mode, "mc")) {
        // correctness: one multimem.st stream from GPU 0, every member (GPU 0 included) must hold it
        timed(1, F, 0, 1, bytes, 148, 0xC0FFEEu);
        for (int d = 0; d < nd; d++) printf("multimem.st: member %d mismatching chunks: %llu\n", d, check(F, d, bytes, 0xC0FFEEu));
        for (int ctas : {1, 2, 4, 8, 16, 32, 64, 148, 296})
            printf("multimem.st.v4 from GPU 0 to %d members, %3d CTAs x 512 thr: %8.1f GB/s of source bytes (x%d members landed)\n", nd, ctas,
                   timed(1, F, 0, 1, bytes, ctas, 7u + ctas), nd);
    } else if (!strcmp(mode, "st")) {
        timed(0, F, 1, followers, bytes, 148, 0xBEEFu);
        for (int d = 1; d <= followers; d++) printf("st.v4: member %d mismatching chunks: %llu\n", d, check(F, d, bytes, 0xBEEFu));
        for (int nf : {1, followers}) {
            for (int ctas : {1, 2, 4, 8, 16, 32, 64, 148, 296})
                printf("st.global.v4 from GPU 0 to %d peer(s), %3d CTAs x 512 thr: %8.1f GB/s of source bytes, %8.1f GB/s egress\n", nf, ctas,
                       timed(0, F, 1, nf, bytes, ctas, 11u + ctas), nf * timed(0, F, 1, nf, bytes, ctas, 13u + ctas));
            if (followers == 1) break;
        }
    } else if (!strcmp(mode, "bulk")) {
        timed(2, F, 1, followers, bytes, 148, 0xABCDu);
        for (int d = 1; d <= followers; d++) printf("cp.async.bulk: member %d mismatching chunks: %llu\n", d, check(F, d, bytes, 0xABCDu));
        for (int nf : {1, followers}) {
            for (int ctas : {1, 2, 4, 8, 16, 32, 64, 148})
                printf("cp.async.bulk smem->peer from GPU 0 to %d peer(s), %3d CTAs: %8.1f GB/s of source bytes, %8.1f GB/s egress\n", nf, ctas,
                       timed(2, F, 1, nf, bytes, ctas, 17u + ctas), nf * timed(2, F, 1, nf, bytes, ctas, 19u + ctas));
            if (followers == 1) break;
        }
    } else if (!strcmp(mode, "mcbulk")) {
        // does the bulk-copy engine accept a multicast destination?
        CK(cudaSetDevice(0));
        k_bulk<<<148, 256, 32768>>>((uint8_t *)F.mcptr, 0, 0, 0, 0, 0, 1, bytes, 0x5151u);
        cudaError_t e = cudaDeviceSynchronize();
        printf("cp.async.bulk to a multicast address: %s\n", cudaGetErrorString(e));
        if (e == cudaSuccess) for (int d = 0; d < nd; d++) printf("  member %d mismatching chunks: %llu\n", d, check(F, d, bytes, 0x5151u));
    }
    return 0;
}
