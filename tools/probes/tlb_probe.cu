// tlb_probe.cu — random 64-byte gathers over buffers of growing size, allocated with cudaMalloc and with the
// virtual-memory API (cuMemCreate/cuMemMap, 512 MiB-aligned), to see where TLB reach ends and whether large
// allocations are mapped with bigger pages.  Build: nvcc -arch=sm_100a -O3 -o tlb_probe tlb_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void gather(const double* __restrict__ buf, unsigned long long nblk, unsigned long long n, double* out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    double acc = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long b = mix64(i) % nblk;
        double a0, a1, a2, a3, b0, b1, b2, b3;
        const double* p = buf + b * 8;
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(p));
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3) : "l"(p + 4));
        acc += a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3;
    }
    if (acc == 12345.678) out[0] = acc;
}

static double run(const double* buf, size_t bytes, double* out) {
    const unsigned long long n = 400000000ull;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    gather<<<148 * 8, 256>>>(buf, bytes / 64, n / 4, out);
    cudaEventRecord(e0);
    gather<<<148 * 8, 256>>>(buf, bytes / 64, n, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return n / (ms * 1e-3) / 1e9;
}

int main() {
    cuInit(0);
    cudaSetDevice(0);
    cudaFree(0);
    double* out;
    cudaMalloc(&out, 8);
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    size_t gmin = 0, grec = 0;
    cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    printf("granularity min %zu recommended %zu\n", gmin, grec);
    const size_t MB = 1ull << 20;
    size_t sizes[] = {64 * MB, 128 * MB, 256 * MB, 384 * MB, 512 * MB, 1024 * MB, 2048 * MB, 4096 * MB};
    for (size_t bytes : sizes) {
        double* a;
        cudaMalloc(&a, bytes);
        cudaMemset(a, 0, bytes);
        double r1 = run(a, bytes, out);
        cudaFree(a);
        // VMM allocation: one physical handle, VA aligned to 512 MiB
        double r2 = -1;
        size_t sz = (bytes + 512 * MB - 1) / (512 * MB) * (512 * MB);
        CUmemGenericAllocationHandle h;
        CUdeviceptr va = 0;
        if (cuMemCreate(&h, sz, &prop, 0) == CUDA_SUCCESS) {
            if (cuMemAddressReserve(&va, sz, 512 * MB, 0, 0) == CUDA_SUCCESS) {
                if (cuMemMap(va, sz, 0, h, 0) == CUDA_SUCCESS) {
                    CUmemAccessDesc acc = {};
                    acc.location = prop.location;
                    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
                    if (cuMemSetAccess(va, sz, &acc, 1) == CUDA_SUCCESS) {
                        cudaMemset((void*)va, 0, bytes);
                        r2 = run((const double*)va, bytes, out);
                    }
                    cuMemUnmap(va, sz);
                }
                cuMemAddressFree(va, sz);
            }
            cuMemRelease(h);
        }
        printf("%5zu MiB: cudaMalloc %.2f G gathers/s (%.2f TB/s)   VMM-512MiB-aligned %.2f G gathers/s\n", bytes / MB, r1,
               r1 * 64 / 1e3, r2);
        fflush(stdout);
    }
    return 0;
}
