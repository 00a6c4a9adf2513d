// tma_gather4_probe.cu — can TMA's tile::gather4 replace the four 16-byte row loads of a random 3-D cell?
//
// The plain-layout corner gather of a random cell costs four LDG.128 on four different lines; on the slab passes the
// L1 data pipe runs at 88 % of its wavefront rate because of them.  cp.async.bulk.tensor.2d...tile::gather4 (UTMALDG)
// fetches four rows of a 2-D tensor at one column offset into shared memory without going through the LSU: the table
// is viewed as (G*G rows) x (G columns), a point's rows are (x,y), (x,y+1), (x+1,y), (x+1,y+1) and the box is four
// doubles wide (an aligned 32-byte sector per row that holds z and z+1).
// Modes: ldg  = 4 x LDG.128 per point (as gather_probe "rows")
//        tma  = 1 x gather4 per point issued by the point's own lane, 32 per warp per mbarrier phase, double buffered
// Both read the eight corner values and add them; `+s` adds the 24 B in / 8 B out streams.
// Build: nvcc -arch=sm_100a -O3 -o tma_gather4_probe tma_gather4_probe.cu   (no -lcuda: the encoder comes from the runtime)
// Usage: tma_gather4_probe [box_rows=1|4]
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
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cell_of(unsigned long long i, int G, unsigned& x, unsigned& y, unsigned& z) {
    const unsigned long long h = mix64(i);
    x = (unsigned)(h % (unsigned)(G - 1));
    y = (unsigned)((h >> 20) % (unsigned)(G - 1));
    z = (unsigned)((h >> 40) % (unsigned)(G - 1));
}

template <bool STREAMS>
__global__ void __launch_bounds__(256) probe_ldg(const double* __restrict__ tab, int G, unsigned long long n,
                                                 const double* __restrict__ sx, const double* __restrict__ sy,
                                                 const double* __restrict__ sz, double* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    double sink = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double acc = 0;
        if (STREAMS) acc = __ldcs(sx + i) + __ldcs(sy + i) + __ldcs(sz + i);
        unsigned x, y, z;
        cell_of(i, G, x, y, z);
        z &= ~1u;
        const double* p = tab + ((size_t)x * G + y) * G + z;
        double2 a = __ldg((const double2*)p), b = __ldg((const double2*)(p + G));
        double2 c = __ldg((const double2*)(p + (size_t)G * G)), d = __ldg((const double2*)(p + (size_t)G * G + G));
        acc += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
        if (STREAMS) __stcs(out + i, acc);
        else sink += acc;
    }
    if (!STREAMS && sink == 12345.678) out[0] = sink;
}

constexpr int kWarps = 4;
template <bool STREAMS>
__global__ void __launch_bounds__(kWarps * 32) probe_tma(const __grid_constant__ CUtensorMap map, int G, unsigned long long n,
                                                         const double* __restrict__ sx, const double* __restrict__ sy,
                                                         const double* __restrict__ sz, double* __restrict__ out,
                                                         unsigned long long* __restrict__ bad) {
    __shared__ __align__(128) double buf[kWarps][2][32][16];  // 4 rows x 4 doubles per lane
    __shared__ __align__(8) unsigned long long bar[kWarps][2];
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    if (lane == 0) {
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[w][b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long iters = (n + stride - 1) / stride;  // warp-uniform trip count (tail lanes are clamped)
    unsigned phase[2] = {0u, 0u};
    auto issue = [&](unsigned long long idx, unsigned b) {
        unsigned x, y, z;
        cell_of(idx < n ? idx : n - 1, G, x, y, z);
        const int row = (int)(x * (unsigned)G + y);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[w][b])), "r"(32u * 128u) : "memory");
        __syncwarp();
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                smem_u32(&buf[w][b][lane][0])),
            "l"(&map), "r"((int)(z & ~1u)), "r"(row), "r"(row + 1), "r"(row + G), "r"(row + G + 1), "r"(smem_u32(&bar[w][b]))
            : "memory");
    };
    auto wait = [&](unsigned b) {
        for (unsigned tries = 0; tries < (1u << 22); ++tries) {
            unsigned done;
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&bar[w][b])), "r"(phase[b]) : "memory");
            if (done) { phase[b] ^= 1u; return; }
        }
        __trap();
    };
    double sink = 0;
    issue(i, 0);
    for (unsigned long long it = 0; it < iters; ++it) {
        const unsigned b = (unsigned)(it & 1u);
        if (it + 1 < iters) issue(i + stride, b ^ 1u);
        double acc = 0;
        if (STREAMS && i < n) acc = __ldcs(sx + i) + __ldcs(sy + i) + __ldcs(sz + i);
        wait(b);
        const double* v = &buf[w][b][lane][0];
        double s = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) s += v[4 * r] + v[4 * r + 1];   // z is even in this probe: columns 0 and 1
        acc += s;
        __syncwarp();  // the buffer is free for the issue of iteration it + 2
        if (i < n) {
            if (STREAMS) __stcs(out + i, acc);
            else sink += acc;
        }
        i += stride;
    }
    if (!STREAMS && sink == 12345.678) out[0] = sink;
    (void)bad;
}

// correctness of the gather4 addressing: table[i] = i, the sum of the eight corners is known in closed form
__global__ void check_tma(const __grid_constant__ CUtensorMap map, int G, unsigned long long n, unsigned long long* bad) {
    __shared__ __align__(128) double buf[32][16];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;
    for (unsigned long long i = blockIdx.x * 32ull + lane; i < n; i += gridDim.x * 32ull) {
        unsigned x, y, z;
        cell_of(i, G, x, y, z);
        z &= ~1u;
        const int row = (int)(x * (unsigned)G + y);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32u * 128u) : "memory");
        __syncwarp();
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                smem_u32(&buf[lane][0])),
            "l"(&map), "r"((int)z), "r"(row), "r"(row + 1), "r"(row + G), "r"(row + G + 1), "r"(smem_u32(&bar))
            : "memory");
        bool ok = false;
        for (unsigned tries = 0; tries < (1u << 22) && !ok; ++tries) {
            unsigned done;
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            ok = done != 0;
        }
        if (!ok) __trap();
        phase ^= 1u;
        const double base = (double)(((size_t)x * G + y) * G + z);
        const double want[4] = {base, base + G, base + (double)G * G, base + (double)G * G + G};
        for (int r = 0; r < 4; ++r)
            if (buf[lane][4 * r] != want[r] || buf[lane][4 * r + 1] != want[r] + 1) atomicAdd(bad, 1ull);
        __syncwarp();
    }
}
__global__ void fill_iota(double* t, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) t[i] = (double)i;
}

template <class F>
double timed(unsigned long long n, F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(n / 4);
    cudaEventRecord(e0);
    launch(n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    return n / (ms * 1e-3) / 1e9;
}

int main(int argc, char** argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    const unsigned long long n = 400000000ull;
    double *sx, *sy, *sz, *out;
    unsigned long long* bad;
    cudaMalloc(&sx, n * 8); cudaMalloc(&sy, n * 8); cudaMalloc(&sz, n * 8); cudaMalloc(&out, n * 8);
    cudaMalloc(&bad, 8);
    cudaMemset(sx, 0, n * 8); cudaMemset(sy, 0, n * 8); cudaMemset(sz, 0, n * 8);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encoder\n"); return 1; }
    printf("tile::gather4 probe, tensor map box = {4 doubles, %d row(s)}\n", box_rows);
    for (int G : {128, 192, 256}) {
        double* tab;
        const size_t elems = (size_t)G * G * G;
        cudaMalloc(&tab, elems * 8 + 4096);
        fill_iota<<<148 * 8, 256>>>(tab, elems);
        CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)G, (cuuint64_t)G * G};
        const cuuint64_t gstr[1] = {(cuuint64_t)G * 8};
        const cuuint32_t box[2] = {4, (cuuint32_t)box_rows};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, tab, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("G=%d: cuTensorMapEncodeTiled failed (%d)\n", G, (int)r); return 1; }
        cudaMemset(bad, 0, 8);
        check_tma<<<148, 32>>>(map, G, 1000000ull, bad);
        unsigned long long hbad = 0;
        cudaError_t e = cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("G=%d: gather4 check failed to run: %s\n", G, cudaGetErrorString(e)); return 1; }
        printf("G=%d (%.0f MB): gather4 addressing check: %llu wrong rows of 4000000\n", G, elems * 8 / 1e6, hbad);
        if (hbad) return 1;
        const double l0 = timed(n, [&](unsigned long long m) { probe_ldg<false><<<148 * 8, 256>>>(tab, G, m, sx, sy, sz, out); });
        const double l1 = timed(n, [&](unsigned long long m) { probe_ldg<true><<<148 * 8, 256>>>(tab, G, m, sx, sy, sz, out); });
        const double t0 = timed(n, [&](unsigned long long m) { probe_tma<false><<<148 * 6, kWarps * 32>>>(map, G, m, sx, sy, sz, out, bad); });
        const double t1 = timed(n, [&](unsigned long long m) { probe_tma<true><<<148 * 6, kWarps * 32>>>(map, G, m, sx, sy, sz, out, bad); });
        printf("G=%d (%.0f MB): 4 x LDG.128 %.1f  (+streams %.1f)   gather4 %.1f  (+streams %.1f)  Gpt/s\n", G, elems * 8 / 1e6, l0, l1, t0, t1);
        fflush(stdout);
        cudaFree(tab);
    }
    return 0;
}
