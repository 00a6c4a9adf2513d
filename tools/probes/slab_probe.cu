// slab_probe.cu — upper bound for "two slab passes with compaction" on random queries (256^3 f64, plain layout):
// one pass = scan x of all n points + evaluate the n/2 members (all lanes busy) whose cells lie in one half of
// the table (67 MB).  Empty kernel: access pattern and streams only.
// Build: nvcc -arch=sm_100a -O3 -o slab_probe slab_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// One pass: the first `members` points are evaluated (all lanes busy, as after perfect compaction) against the slab
// of x-planes [xlo, xlo + xn); the remaining points only have their x read (the membership scan).
__global__ void __launch_bounds__(256, 4) pass(const double* __restrict__ tab, int G, int xlo, int xn, unsigned long long n,
                                               unsigned long long members, const double* __restrict__ sx,
                                               const double* __restrict__ sy, const double* __restrict__ sz,
                                               double* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    double sink = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (i >= members) {
            sink += __ldcs(sx + i);
            continue;
        }
        double acc = __ldcs(sx + i) + __ldcs(sy + i) + __ldcs(sz + i);
        const unsigned long long h = mix64(i);
        const unsigned x = xlo + (unsigned)(h % (unsigned)xn), y = (unsigned)((h >> 20) % (unsigned)(G - 1));
        const unsigned z = (unsigned)((h >> 40) % (unsigned)(G - 1));
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
        __stcs(out + i, acc);
    }
    if (sink == 12345.678) out[0] = sink;
}

int main() {
    const unsigned long long n = 400000000ull;
    const int G = 256;
    double *sx, *sy, *sz, *out, *tab;
    cudaMalloc(&sx, n * 8);
    cudaMalloc(&sy, n * 8);
    cudaMalloc(&sz, n * 8);
    cudaMalloc(&out, n * 8);
    cudaMalloc(&tab, (size_t)G * G * G * 8 + 4096);
    cudaMemset(sx, 0, n * 8);
    cudaMemset(sy, 0, n * 8);
    cudaMemset(sz, 0, n * 8);
    cudaMemset(tab, 0, (size_t)G * G * G * 8 + 4096);
    for (int W : {1, 2, 3, 4}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const int per = (G - 1 + W - 1) / W;
        // warm-up
        pass<<<148 * 4, 256>>>(tab, G, 0, per - 1, n / 4, n / 8, sx, sy, sz, out);
        cudaEventRecord(e0);
        // W passes; pass w evaluates n/W members (emulated as the first n/W*2 points' even indices)
        for (int w = 0; w < W; ++w) pass<<<148 * 4, 256>>>(tab, G, w * per, per - 1, n, n / W, sx, sy, sz, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("W=%d slabs of %d planes (%.0f MB): %.2f ms for %llu members -> %.1f Gpt/s   [%s]\n",
               W, per, (double)per * G * G * 8 / 1e6, ms, n, n / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
