// layout_probe.cu — candidate table layouts for random queries on a 256^3 f64 grid, as empty gather kernels
// (access pattern + the 32 B/point of query/result streams, no arithmetic):
//   plain   134 MB  4 rows x (16 B aligned pair, second pair when k is odd)        ~5 sectors, 4 lines / point
//   yzpair  268 MB  rows j and j+1 interleaved per k: 32 contiguous bytes per x-plane  ~3 sectors, 2 lines / point
//   packed  1 GiB   one aligned 64-byte block per cell                               2 sectors, 1 line / point
// Build: nvcc -arch=sm_100a -O3 -o layout_probe layout_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) probe(const double* __restrict__ tab, int G, unsigned long long n,
                                                const double* __restrict__ sx, const double* __restrict__ sy,
                                                const double* __restrict__ sz, double* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double acc = __ldcs(sx + i) + __ldcs(sy + i) + __ldcs(sz + i);
        const unsigned long long h = mix64(i);
        const unsigned x = (unsigned)(h % (unsigned)(G - 1)), y = (unsigned)((h >> 20) % (unsigned)(G - 1));
        const unsigned z = (unsigned)((h >> 40) % (unsigned)(G - 1));
        if (MODE == 0) {
            const double* p = tab + ((size_t)x * G + y) * G + (z & ~1u);
            const bool odd = z & 1;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double* q = p + (r & 1) * (size_t)G + (r >> 1) * (size_t)G * G;
                double2 a = __ldg((const double2*)q);
                acc += a.x + a.y;
                if (odd) {
                    double2 b = __ldg((const double2*)(q + 2));
                    acc += b.x;
                }
            }
        } else if (MODE == 1) {
            // element (i, j, k) pair-interleaved: offset ((i*G + j)*G + k) * 2, 4 doubles = cell face in one x-plane
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double* q = tab + (((size_t)(x + r) * G + y) * G + z) * 2;
                double2 a = __ldg((const double2*)q), b = __ldg((const double2*)(q + 2));
                acc += a.x + a.y + b.x + b.y;
            }
        } else {
            const double* p = tab + (((size_t)x * (G - 1) + y) * (G - 1) + z) * 8;
            double a0, a1, a2, a3, b0, b1, b2, b3;
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(p));
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3) : "l"(p + 4));
            acc += a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3;
        }
        __stcs(out + i, acc);
    }
}

template <int MODE>
double run(const double* tab, int G, unsigned long long n, const double* sx, const double* sy, const double* sz, double* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<148 * 4, 256>>>(tab, G, n / 4, sx, sy, sz, out);
    cudaEventRecord(e0);
    probe<MODE><<<148 * 4, 256>>>(tab, G, n, sx, sy, sz, out);
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
    for (int G : {128, 192, 256, 320}) {
        double* tab;
        const size_t bytes = (size_t)G * G * G * 8 * 8 + 4096;
        cudaMalloc(&tab, bytes);
        cudaMemset(tab, 0, bytes);
        printf("G=%d: plain(%.0f MB) %.1f   yzpair(%.0f MB) %.1f   packed(%.0f MB) %.1f   Gpt/s (with streams, 4 CTAs/SM)\n", G,
               (double)G * G * G * 8 / 1e6, run<0>(tab, G, n, sx, sy, sz, out), (double)G * G * G * 16 / 1e6,
               run<1>(tab, G, n, sx, sy, sz, out), (double)(G - 1) * (G - 1) * (G - 1) * 64 / 1e6, run<2>(tab, G, n, sx, sy, sz, out));
        fflush(stdout);
        cudaFree(tab);
    }
    return 0;
}
