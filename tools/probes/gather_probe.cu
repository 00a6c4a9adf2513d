// gather_probe.cu — what limits the plain-layout corner gather of a 256^3 f64 table (134 MB)?
// Modes: rows  = 4 x 16-byte loads per point at the four (x,y) rows of a random cell (plain layout pattern)
//        rows+s = same, plus a 24-byte streaming read and an 8-byte streaming write per point
//        cell  = 1 x 64-byte block per point in a 134 MB buffer (packed-like, same footprint)
//        cell+s = same plus streams
// Build: nvcc -arch=sm_100a -O3 -o gather_probe gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

template <int MODE, bool STREAMS>
__global__ void probe(const double* __restrict__ tab, int G, unsigned long long n, const double* __restrict__ sx,
                      const double* __restrict__ sy, const double* __restrict__ sz, double* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    double sink = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double acc = 0;
        if (STREAMS) acc = __ldcs(sx + i) + __ldcs(sy + i) + __ldcs(sz + i);
        const unsigned long long h = mix64(i);
        if (MODE == 0) {
            const unsigned x = (unsigned)(h % (unsigned)(G - 1)), y = (unsigned)((h >> 20) % (unsigned)(G - 1));
            const unsigned z = (unsigned)((h >> 40) % (unsigned)(G - 1)) & ~1u;
            const double* p = tab + ((size_t)x * G + y) * G + z;
            double2 a = __ldg((const double2*)p), b = __ldg((const double2*)(p + G));
            double2 c = __ldg((const double2*)(p + (size_t)G * G)), d = __ldg((const double2*)(p + (size_t)G * G + G));
            acc += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
        } else {
            const unsigned long long nblk = (unsigned long long)G * G * G / 8;
            const double* p = tab + (h % nblk) * 8;
            double a0, a1, a2, a3, b0, b1, b2, b3;
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(p));
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3) : "l"(p + 4));
            acc += a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3;
        }
        if (STREAMS) __stcs(out + i, acc);
        else sink += acc;
    }
    if (!STREAMS && sink == 12345.678) out[0] = sink;
}

template <int MODE, bool STREAMS>
double run(const double* tab, int G, unsigned long long n, const double* sx, const double* sy, const double* sz, double* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE, STREAMS><<<148 * 8, 256>>>(tab, G, n / 4, sx, sy, sz, out);
    cudaEventRecord(e0);
    probe<MODE, STREAMS><<<148 * 8, 256>>>(tab, G, n, sx, sy, sz, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return n / (ms * 1e-3) / 1e9;
}

int main() {
    const unsigned long long n = 400000000ull;
    double *sx, *sy, *sz, *out;
    cudaMalloc(&sx, n * 8);
    cudaMalloc(&sy, n * 8);
    cudaMalloc(&sz, n * 8);
    cudaMalloc(&out, n * 8);
    cudaMemset(sx, 0, n * 8);
    cudaMemset(sy, 0, n * 8);
    cudaMemset(sz, 0, n * 8);
    for (int G : {128, 160, 192, 224, 256, 288, 320}) {
        double* tab;
        const size_t bytes = (size_t)G * G * G * 8;
        cudaMalloc(&tab, bytes + 4096);
        cudaMemset(tab, 0, bytes + 4096);
        printf("G=%d (%.0f MB): rows %.1f  rows+streams %.1f  cell %.1f  cell+streams %.1f  Gpt/s\n", G, bytes / 1e6,
               run<0, false>(tab, G, n, sx, sy, sz, out), run<0, true>(tab, G, n, sx, sy, sz, out),
               run<1, false>(tab, G, n, sx, sy, sz, out), run<1, true>(tab, G, n, sx, sy, sz, out));
        fflush(stdout);
        cudaFree(tab);
    }
    return 0;
}
