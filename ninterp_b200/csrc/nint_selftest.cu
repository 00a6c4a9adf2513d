// nint_selftest.cu — device self-test of div_cell(): the FMA-corrected reciprocal division used by the
// Linear kernels must be bit-identical to the IEEE division the reference performs.
#include "nint_device.cuh"

namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// Mantissa patterns that stress rounding: random, few set bits, all ones, all ones but the last.
__device__ __forceinline__ unsigned long long mantissa(unsigned long long r, int bits) {
    const unsigned long long mask = (1ull << bits) - 1ull;
    switch ((r >> 60) & 7) {
    case 0: return mask;
    case 1: return mask - 1;
    case 2: return 0;
    case 3: return 1;
    case 4: return (r & mask) & ((r >> 7) & mask) & ((r >> 13) & mask);  // sparse
    case 5: return ((r & mask) | ((r >> 7) & mask) | ((r >> 13) & mask));  // dense
    default: return r & mask;
    }
}

template <class T>
struct Fp;
template <>
struct Fp<double> {
    static __device__ double make(unsigned long long r, int emin, int emax) {
        const int e = emin + (int)((r >> 40) % (unsigned)(emax - emin + 1));
        const unsigned long long bits = ((unsigned long long)(1023 + e) << 52) | mantissa(mix64(r), 52) | ((r & 1ull) << 63);
        return __longlong_as_double((long long)bits);
    }
    static __device__ bool same(double a, double b) {
        return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
    }
    static __device__ double nudge(double x, int k) {
        return __longlong_as_double(__double_as_longlong(x) + k);
    }
};
template <>
struct Fp<float> {
    static __device__ float make(unsigned long long r, int emin, int emax) {
        const int e = emin + (int)((r >> 40) % (unsigned)(emax - emin + 1));
        const unsigned bits = ((unsigned)(127 + e) << 23) | (unsigned)mantissa(mix64(r), 23) | ((unsigned)(r & 1ull) << 31);
        return __int_as_float((int)bits);
    }
    static __device__ bool same(float a, float b) {
        return __float_as_int(a) == __float_as_int(b) || (a != a && b != b);
    }
    static __device__ float nudge(float x, int k) { return __int_as_float(__float_as_int(x) + k); }
};

template <class T>
__global__ void div_selftest_kernel(unsigned long long n, unsigned long long seed, int emin, int emax,
                                    unsigned long long* mismatches, unsigned long long* fast_taken) {
    unsigned long long bad = 0, fast = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long r0 = mix64(seed + 3 * i), r1 = mix64(seed + 3 * i + 1), r2 = mix64(seed + 3 * i + 2);
        T w = Fp<T>::make(r0, emin, emax);
        w = w < T(0) ? -w : w;  // cell widths are positive
        T a = Fp<T>::make(r1, emin, emax);
        if ((r2 & 3) == 0) {
            // quotient within a few ulp of a representable number: a = RN(q * w) nudged by -2..2 ulp
            const T q = Fp<T>::make(r2, -4, 4);
            a = Fp<T>::nudge(q * w, (int)((r2 >> 8) % 5) - 2);
        } else if ((r2 & 3) == 1) {
            // 0 <= a <= w, as for an in-range query inside its cell
            const T u = Fp<T>::make(r2, -30, -1);
            a = (u < T(0) ? -u : u) * w;
        }
        const T r = T(1) / w;  // what the host stores (correctly rounded reciprocal)
        // the host only stores r for widths inside the window; outside it stores 0
        const T lim_lo = sizeof(T) == 8 ? (T)0x1p-500 : (T)0x1p-60f;
        const T lim_hi = sizeof(T) == 8 ? (T)0x1p500 : (T)0x1p60f;
        const T r_host = (w >= lim_lo && w < lim_hi) ? r : T(0);
        const T got = nint::div_cell(a, w, r_host);
        const T want = a / w;
        if (!Fp<T>::same(got, want)) ++bad;
        const T aa = a < T(0) ? -a : a;
        if (r_host != T(0) && aa >= lim_lo && aa < lim_hi) ++fast;
    }
    if (bad) atomicAdd(mismatches, bad);
    if (fast) atomicAdd(fast_taken, fast);
}

}  // namespace

extern "C" NINT_API int nint_selftest_divide(int dtype, uint64_t n, uint64_t seed, uint64_t* mismatches,
                                             uint64_t* fast_path_taken) {
    if (!mismatches || !fast_path_taken) return NINT_E_INVALID_ARGUMENT;
    unsigned long long* d = nullptr;
    if (cudaMalloc((void**)&d, 16) != cudaSuccess) return NINT_E_CUDA;
    cudaMemset(d, 0, 16);
    // exponent range slightly wider than the fast-path window so that both sides of its edges are hit
    if (dtype == NINT_F64) div_selftest_kernel<double><<<148 * 8, 256>>>(n, seed, -520, 520, d, d + 1);
    else div_selftest_kernel<float><<<148 * 8, 256>>>(n, seed, -70, 70, d, d + 1);
    unsigned long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return NINT_E_CUDA;
    *mismatches = h[0];
    *fast_path_taken = h[1];
    return NINT_OK;
}
