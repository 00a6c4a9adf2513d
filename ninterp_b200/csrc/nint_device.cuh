// nint_device.cuh — device side of the batch engine: one query point per thread.
//
// Per-point semantics follow ninterp 0.8.1 (paths relative to the reference's src/):
//   front-end  interpolator/{one,two,three,n}/mod.rs  (bounds test + Extrapolate policy)
//   bracket    strategy/traits.rs:10-33                (find_nearest_index)
//   strategies interpolator/{one,two,three,n}/strategies.rs
// Every arithmetic step is a single IEEE round-to-nearest operation in T: this translation unit
// must be compiled with -fmad=false (Rust never contracts a*b+c) and without -use_fast_math.
//
// Data layout in HBM
//   coords   struct of arrays (one column per axis) or array of [T;N] points (coord_stride)
//   values   dense C-order table, last axis fastest: the 2^N corners of a cell are 2^(N-1)
//            pairs of adjacent elements
//   pack     per axis: the L nodes followed by an M-entry bucket table; staged once per CTA in
//            shared memory so the bracket search never leaves the SM
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/ninterp_cuda.h"

namespace nint {

constexpr int kMaxDim = NINT_MAX_NDIM;
constexpr int kBlock = 256;
constexpr int kExtrapRuntime = -1;

template <class T>
struct Params {
    const T* coords[kMaxDim];  // coordinate d of point i is coords[d][i * coord_stride]
    long long coord_stride;    // 1 for columns (SoA), ndim for points (AoS)
    T* out;
    uint8_t* status;
    nint_batch_info* info;
    const T* values;
    const unsigned char* pack;  // axis pack in global memory (nodes + bucket tables)
    unsigned pack_bytes;        // multiple of 16
    unsigned long long n;
    unsigned long long index_base;  // batch index of point 0 (host pipeline chunks)
    T fill;
    int extrap;  // policy when the kernel is instantiated with EXTRAP == kExtrapRuntime
    int L[kMaxDim];
    int M[kMaxDim];
    unsigned node_off[kMaxDim];  // byte offsets into the pack
    unsigned lut_off[kMaxDim];
    long long vstride[kMaxDim];  // element strides of `values`
    T g0[kMaxDim], gl[kMaxDim], inv_w[kMaxDim];
};

template <class T>
struct AxisRef {
    const T* g;        // nodes
    const uint2* lut;  // bucket b -> [lo, hi]: the node count below any point of the bucket
    int L, M;
    T g0, gl, inv_w;
};

__device__ __forceinline__ int to_int_sat(float x) { return __float2int_rz(x); }    // NaN -> 0
__device__ __forceinline__ int to_int_sat(double x) { return __double2int_rz(x); }  // saturating

template <class T>
__device__ __forceinline__ T quiet_nan();
template <>
__device__ __forceinline__ float quiet_nan<float>() { return __int_as_float(0x7fc00000); }
template <>
__device__ __forceinline__ double quiet_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }

// Streaming accesses: queries and results are touched once, keep them out of L1 and mark them
// evict-first in L2 so that the value table stays resident.
template <class T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
template <class T>
__device__ __forceinline__ void st_stream(T* p, T v) { __stcs(p, v); }
template <class T>
__device__ __forceinline__ T ld_table(const T* p) { return __ldg(p); }

// c = #{ i : g[i] < p } for a non-NaN p.  The bucket table bounds c to a handful of candidates
// (one bucket of slack on each side absorbs the rounding of the bucket index); the decision is
// always made by comparing against the node values themselves, so c is exact for every input.
template <class T>
__device__ __forceinline__ int count_less(const AxisRef<T>& a, T p) {
    int b = to_int_sat((p - a.g0) * a.inv_w);
    b = max(0, min(b, a.M - 1));
    const uint2 e = a.lut[b];
    int lo = (int)e.x;
    int n = (int)e.y - (int)e.x;
    while (n > 0) {
        const int half = n >> 1;
        const bool lt = a.g[lo + half] < p;
        lo = lt ? lo + half + 1 : lo;
        n = lt ? n - half - 1 : half;
    }
    return lo;
}

// lib.rs:81-83 with Rust's float rem_euclid (fmod, then `+ |b|` when negative)
template <class T>
__device__ __forceinline__ T wrap_coord(T x, T lo, T hi) {
    const T a = x - lo;
    const T b = hi - lo;
    T r = fmod(a, b);
    if (r < T(0)) r = r + fabs(b);
    return lo + r;
}

// num_traits::clamp
template <class T>
__device__ __forceinline__ T clamp_coord(T x, T lo, T hi) {
    return (x < lo) ? lo : ((x > hi) ? hi : x);
}

// ---------------------------------------------------------------------------------------------
// Strategies.  `status` is set to NINT_STATUS_PANIC where the reference would index out of
// bounds; the returned value is then unspecified (the caller writes a quiet NaN).
//
// COLLAPSE  an axis whose coordinate equals a node is *selected*, not blended
//           (one/strategies.rs:14-16 for 1-D, n/strategies.rs:36-46 for N-D).
// LEN1      InterpND returns early when a single element is left after collapsing
//           (n/strategies.rs:47-50), which also covers un-collapsed axes of length 1.
// ---------------------------------------------------------------------------------------------
template <class T, int N, bool COLLAPSE, bool LEN1>
__device__ __forceinline__ T eval_linear(const AxisRef<T> (&ax)[N], const long long (&vs)[N],
                                         const T* __restrict__ values, const T (&q)[N], unsigned& status) {
    int idx[N];
    T t[N];
    bool fixed[N];
    bool panic = false;
    bool single = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const T p = q[d];
        const bool isnan = (p != p);
        const int L = ax[d].L;
        const int c = isnan ? 0 : count_less(ax[d], p);
        bool hit = false;
        if (COLLAPSE) hit = !isnan && c < L && ax[d].g[c] == p;
        fixed[d] = hit;
        idx[d] = hit ? c : 0;
        t[d] = T(0);
        if (!hit) {
            single = single && (L == 1);
            if (isnan || L < 2) {
                // three/strategies.rs:16-24 -> traits.rs:10-33 returns L-1 for NaN, then g[L] panics;
                // a length-1 axis panics on g[1] or on `len - 2`
                panic = true;
            } else {
                const int l = (p >= ax[d].gl) ? L - 2 : max(c, 1) - 1;
                const T gl_ = ax[d].g[l];
                const T gu_ = ax[d].g[l + 1];
                idx[d] = l;
                t[d] = (p - gl_) / (gu_ - gl_);
            }
        }
    }
    long long base = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) base += (long long)idx[d] * vs[d];
    if (LEN1 && single) return ld_table(values + base);
    if (panic) {
        status = NINT_STATUS_PANIC;
        return T(0);
    }
    // gather the cell: corner bit of axis 0 is the most significant one
    T cube[1 << N];
#pragma unroll
    for (int c = 0; c < (1 << N); ++c) {
        long long off = base;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            if ((c >> (N - 1 - d)) & 1) off += fixed[d] ? 0 : vs[d];
        }
        cube[c] = ld_table(values + off);
    }
    // reduce axis 0 first, then 1, ...: v_l * (1 - t) + v_u * t, each op rounded separately
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int half = 1 << (N - 1 - d);
        const T td = t[d];
        const T sd = T(1) - td;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const T blended = cube[i] * sd + cube[half + i] * td;
            cube[i] = fixed[d] ? cube[i] : blended;
        }
    }
    return cube[0];
}

// NEAREST (N-D and hard-coded) and, for N == 1, LEFT_NEAREST / RIGHT_NEAREST
template <class T, int N, bool COLLAPSE, bool LEN1, int STRAT>
__device__ __forceinline__ T eval_index(const AxisRef<T> (&ax)[N], const long long (&vs)[N],
                                        const T* __restrict__ values, const T (&q)[N], unsigned& status) {
    int idx[N];
    bool panic = false;
    bool single = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const T p = q[d];
        const bool isnan = (p != p);
        const int L = ax[d].L;
        const int c = isnan ? 0 : count_less(ax[d], p);
        bool hit = false;
        if (COLLAPSE) hit = !isnan && c < L && ax[d].g[c] == p;
        idx[d] = hit ? c : 0;
        if (!hit) {
            single = single && (L == 1);
            if (STRAT == NINT_NEAREST) {
                // find_nearest_index: `== last` -> L-2; beyond last or NaN -> L-1, then g[L] panics
                if (isnan || L < 2 || (c >= L)) {
                    panic = true;
                } else {
                    const int l = (p == ax[d].gl) ? L - 2 : max(c, 1) - 1;
                    const T gl_ = ax[d].g[l];
                    const T gu_ = ax[d].g[l + 1];
                    idx[d] = ((p - gl_) < (gu_ - p)) ? l : l + 1;
                }
            } else {
                // one/strategies.rs:79 and :102 — values[l] / values[l + 1] with l = find_nearest_index
                const int l = (isnan || c >= L) ? L - 1 : max(c, 1) - 1;
                if (STRAT == NINT_LEFT_NEAREST) {
                    idx[d] = l;
                } else {
                    if (l + 1 >= L) panic = true;
                    else idx[d] = l + 1;
                }
            }
        }
    }
    long long off = 0;
#pragma unroll
    for (int d = 0; d < N; ++d) off += (long long)idx[d] * vs[d];
    if (LEN1 && single) return ld_table(values + off);  // un-collapsed length-1 axes sit at index 0
    if (panic) {
        status = NINT_STATUS_PANIC;
        return T(0);
    }
    return ld_table(values + off);
}

__device__ __forceinline__ unsigned long long shfl_down_u64(unsigned long long v, int delta) {
    return __shfl_down_sync(0xffffffffu, v, delta);
}

// ---------------------------------------------------------------------------------------------
// The kernel: persistent CTAs, grid-stride over the batch, one point per thread per trip.
// ---------------------------------------------------------------------------------------------
template <class T, int N, bool IS_ND, int STRAT, int EXTRAP, bool SMEM>
__global__ void __launch_bounds__(kBlock) interp_kernel(const __grid_constant__ Params<T> P) {
    extern __shared__ __align__(16) unsigned char smem_pack[];
    const unsigned char* pack;
    if (SMEM) {
        const uint4* src = reinterpret_cast<const uint4*>(P.pack);
        uint4* dst = reinterpret_cast<uint4*>(smem_pack);
        for (unsigned i = threadIdx.x; i < P.pack_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
        pack = smem_pack;
    } else {
        pack = P.pack;
    }

    AxisRef<T> ax[N];
    long long vs[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
        ax[d].g = reinterpret_cast<const T*>(pack + P.node_off[d]);
        ax[d].lut = reinterpret_cast<const uint2*>(pack + P.lut_off[d]);
        ax[d].L = P.L[d];
        ax[d].M = P.M[d];
        ax[d].g0 = P.g0[d];
        ax[d].gl = P.gl[d];
        ax[d].inv_w = P.inv_w[d];
        vs[d] = P.vstride[d];
    }
    const int extrap = (EXTRAP == kExtrapRuntime) ? P.extrap : EXTRAP;
    constexpr bool kCollapse = IS_ND || (N == 1);
    constexpr bool kLen1 = IS_ND;

    unsigned long long n_err = 0, n_panic = 0;
    unsigned long long first_err = ~0ull, first_panic = ~0ull;

    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        T q[N];
        bool oob = false;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            q[d] = ld_stream(P.coords[d] + i * P.coord_stride);
            // !(first..=last).contains(p): NaN is out of bounds (three/mod.rs:198-201)
            oob = oob || !(ax[d].g0 <= q[d] && q[d] <= ax[d].gl);
        }
        unsigned status = NINT_STATUS_OK;
        T result;
        bool done = false;
        if (oob) {
            if (extrap == NINT_EXTRAPOLATE_FILL) {  // three/mod.rs:204
                result = P.fill;
                done = true;
            } else if (extrap == NINT_EXTRAPOLATE_ERROR) {  // :225-236
                status = NINT_STATUS_EXTRAPOLATE_ERROR;
                done = true;
            } else if (extrap == NINT_EXTRAPOLATE_CLAMP) {  // :205-214, every axis
#pragma unroll
                for (int d = 0; d < N; ++d) q[d] = clamp_coord(q[d], ax[d].g0, ax[d].gl);
            } else if (extrap == NINT_EXTRAPOLATE_WRAP) {  // :215-224, every axis
#pragma unroll
                for (int d = 0; d < N; ++d) q[d] = wrap_coord(q[d], ax[d].g0, ax[d].gl);
            }
            // Enable: evaluate as is (:203)
        }
        if (!done) {
            if (STRAT == NINT_LINEAR) result = eval_linear<T, N, kCollapse, kLen1>(ax, vs, P.values, q, status);
            else result = eval_index<T, N, kCollapse, kLen1, STRAT>(ax, vs, P.values, q, status);
        }
        if (status != NINT_STATUS_OK) {
            result = quiet_nan<T>();
            if (status == NINT_STATUS_PANIC) {
                ++n_panic;
                first_panic = min(first_panic, i + P.index_base);
            } else {
                ++n_err;
                first_err = min(first_err, i + P.index_base);
            }
        }
        st_stream(P.out + i, result);
        if (P.status) P.status[i] = (uint8_t)status;
    }

    if (P.info) {
#pragma unroll
        for (int delta = 16; delta > 0; delta >>= 1) {
            n_err += shfl_down_u64(n_err, delta);
            n_panic += shfl_down_u64(n_panic, delta);
            first_err = min(first_err, shfl_down_u64(first_err, delta));
            first_panic = min(first_panic, shfl_down_u64(first_panic, delta));
        }
        if ((threadIdx.x & 31) == 0) {
            if (n_err) {
                atomicAdd((unsigned long long*)&P.info->n_error, n_err);
                atomicMin((unsigned long long*)&P.info->first_error, first_err);
            }
            if (n_panic) {
                atomicAdd((unsigned long long*)&P.info->n_panic, n_panic);
                atomicMin((unsigned long long*)&P.info->first_panic, first_panic);
            }
        }
    }
}

// 0-D InterpND (n/mod.rs:104-110): every point evaluates to the single stored value.
template <class T>
__global__ void __launch_bounds__(kBlock) constant_kernel(const T* __restrict__ values, T* __restrict__ out,
                                                          uint8_t* __restrict__ status, unsigned long long n) {
    const T v = __ldg(values);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        st_stream(out + i, v);
        if (status) status[i] = NINT_STATUS_OK;
    }
}

// Launch interface implemented by the per-(T, family) translation units.
struct LaunchCfg {
    int ndim;
    int is_nd;
    int strategy;
    int extrap;
    int smem_axes;     // stage the axis pack in shared memory
    unsigned smem_bytes;
    int sm_count;
    void* stream;
};

template <class T>
cudaError_t launch_fixed(const Params<T>& P, const LaunchCfg& cfg);
template <class T>
cudaError_t launch_nd(const Params<T>& P, const LaunchCfg& cfg);
template <class T>
cudaError_t launch_constant(const T* values, T* out, uint8_t* status, unsigned long long n, int sm_count,
                            void* stream);

}  // namespace nint
