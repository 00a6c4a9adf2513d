// nint_device.cuh — device side of the batch engine: one query point per thread.
//
// Per-point semantics follow ninterp 0.8.1 (paths relative to the reference's src/):
//   front-end  interpolator/{one,two,three,n}/mod.rs  (bounds test + Extrapolate policy)
//   bracket    strategy/traits.rs:10-33                (find_nearest_index)
//   strategies interpolator/{one,two,three,n}/strategies.rs
// Every arithmetic step of the reference is a single IEEE round-to-nearest operation in T: this
// translation unit must be compiled with -fmad=false (Rust never contracts a*b+c) and without
// -use_fast_math.  The only FMAs are explicit ones inside div_markstein(), which reproduce IEEE division.
//
// Structure of a point's evaluation
//   fast path   guess the cell on every axis, VERIFY it against the node values (g[k] < p <= g[k+1]) while
//               computing the weights, gather the cell, blend.  A verified cell implies the point is in
//               range and finite, so the fast path is the same for every Extrapolate policy.
//   general     everything else (out of range, NaN/inf, exact node hits, degenerate axes, duplicate
//               nodes, buckets holding several nodes): an out-of-line restatement of the reference's control flow.
//
// Data layout in HBM
//   coords   struct of arrays (one column per axis) or array of [T;N] points (coord_stride)
//   values   dense C-order table, last axis fastest
//   cells    optional cell-packed copy for Linear: the 2^N corners of each cell contiguous and aligned
//   pack     per axis: L+2 records {node, cell reciprocal} and, for a non-uniform axis, an M-entry guess table
//            (bucket -> cell); staged once per CTA in shared memory when small enough.  Behind them, in global
//            memory only, the {lo, hi} bucket tables the general path searches in
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/ninterp_cuda.h"

// -DNINT_BOUNDS_CHECK (python -m ninterp_b200.build with NINT_LIB_VARIANT=checked): every index the kernels form --
// axis records, guess and bucket tables, value table and cell-packed table offsets, shared-memory lists, record lists
// and tiles of the binned route -- is checked against its container and the kernel traps on a violation.  This is
// the stand-in for compute-sanitizer's memcheck, which is closed on the GPU pool: the whole
// -m gpu suite runs against this build.
#ifdef NINT_BOUNDS_CHECK
#define NINT_ASSERT(cond)                   \
    do {                                    \
        if (!(cond)) asm volatile("trap;"); \
    } while (0)
#else
#define NINT_ASSERT(cond) \
    do {                  \
    } while (0)
#endif

namespace nint {

constexpr int kMaxDim = NINT_MAX_NDIM;
#ifndef NINT_BLOCK
#define NINT_BLOCK 256
#endif
constexpr int kBlock = NINT_BLOCK;
// interp_pipe_kernel (nint_pipe.cuh)
constexpr int kPipeChunk = 512;    // points per stage = consumer threads
constexpr int kPipeStages = 4;
constexpr int kPipeThreads = kPipeChunk + 32;  // + the producer warp
// The point loop indexes with 32 bits (fewer registers, one-instruction address arithmetic).
constexpr unsigned long long kMaxLaunchPoints = 1ull << 30;

// per-axis flags (Params::flags)
constexpr int kAxisUniform = 1;  // nodes are evenly spaced to within a small fraction of a cell
constexpr int kModeCoherent = 1;  // consecutive queries mostly fall into neighbouring cells
constexpr int kModeRandom = 2;    // they do not: evaluate in slab passes

constexpr int kAxisAnalytic = 4;  // uniform axis whose nodes are, bit for bit, g[k] = g0 + k * h in T (two roundings)
constexpr int kAxisFast = 2;     // axis may use the fast path: >= 2 nodes, strictly increasing, every cell
                                 // width inside the division window (all reciprocals stored)

template <class T>
struct Params {
    const T* coords[kMaxDim];  // coordinate d of point i is coords[d][i * coord_stride]
    unsigned coord_stride;     // 1 for columns (SoA), ndim for points (AoS)
    T* out;
    uint8_t* status;
    nint_batch_info* info;
    const T* values;
    unsigned long long values_len;  // elements of `values` (the allocation has 4 more: ld_pair's slack)
    unsigned long long ncells;      // cells of `cells`
    const T* cells;              // cell-packed copy of the table (2^N corners per cell, contiguous) or NULL
    const unsigned char* pack;   // axis pack in global memory
    unsigned pack_bytes;         // multiple of 16: records + guess tables, the part staged in shared memory
    unsigned long long n;           // points in this launch: at most kMaxLaunchPoints (the host splits larger batches)
    unsigned long long index_base;  // batch index of point 0 (host pipeline chunks, split launches)
    T fill;
    int extrap;
    float l2_frac;  // fraction of the value table kept evict_last in L2 (0 = no hint)
    // Query-order modes (see "slab passes" below).  mode_flag is written by order_probe_kernel; a launch runs
    // only if *mode_flag == mode_want (NULL: always).  A slab launch evaluates the points whose axis-0 coordinate
    // lies in (win_lo_x, win_hi_x]; the first window is open below (and takes NaN), the last is open above.
    const int* mode_flag;
    int mode_want;
    int blocked;  // interp_kernel: every CTA walks one contiguous range of the batch instead of striding over it
    int wide_lines;  // coherent batches: the cell-packed table is read with an L2 prefetch size of 256 bytes
    T win_lo_x, win_hi_x;
    int win_first, win_last;
    int L[kMaxDim];
    int M[kMaxDim];
    int flags[kMaxDim];
    unsigned node_off[kMaxDim];  // byte offsets into the pack: {node, cell reciprocal} records
    unsigned lut_off[kMaxDim];   // {lo, hi} bucket tables: beyond pack_bytes, read from global memory only
    unsigned guess_off[kMaxDim]; // guess tables of the non-uniform axes (inside pack_bytes)
    long long vstride[kMaxDim];  // element strides of `values`
    long long cstride[kMaxDim];  // cell strides of `cells` (in cells)
    T g0[kMaxDim], gl[kMaxDim];
    T inv_w[kMaxDim];            // buckets per unit length (guess table and {lo, hi} table share the buckets)
    T inv_h[kMaxDim];            // cells per unit length (uniform axes)
    T h[kMaxDim];                // cell width of an analytic axis (kAxisAnalytic): g[k] == g0 + k * h exactly
    // Interp1D / 1-D InterpND Linear: fused pack {node, value} pairs + cell reciprocals (+ guess table), or NULL
    const unsigned char* fused;
    unsigned fused_bytes, fused_rcp_off, fused_guess_off;
};

template <class T>
struct AxisRef {
    const T* rec;      // L+2 records {node, rcp}: rec[2l] = g[l] (two +inf pads at the end),
                       // rec[2l+1] = RN(1 / (g[l+1] - g[l])) for cells inside the safe exponent window, else 0
    const uint2* lut;  // bucket b -> [lo, hi]: bounds on the node count below any point of the bucket (global memory)
    int L, M, flags;
    T g0, gl, inv_w, inv_h;
    __device__ __forceinline__ T g(int i) const {
        NINT_ASSERT(i >= 0 && i <= L + 1);
        return rec[2 * i];
    }
    __device__ __forceinline__ T rcp(int i) const {
        NINT_ASSERT(i >= 0 && i <= L + 1);
        return rec[2 * i + 1];
    }
};

// {g[k], rcp[k]} with one 2-element shared/global load
__device__ __forceinline__ void ld_rec(const double* p, double& g, double& r) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    g = v.x;
    r = v.y;
}
__device__ __forceinline__ void ld_rec(const float* p, float& g, float& r) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    g = v.x;
    r = v.y;
}

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

// Value-table loads carry an L2 cache policy (a fraction of the lines evict_last, the rest evict_first).
__device__ __forceinline__ unsigned long long make_table_policy(float frac) {
    unsigned long long pol;
    if (frac > 0.0f) {
        asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(frac));
    } else {
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    }
    return pol;
}
__device__ __forceinline__ double ld_table(const double* p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_table(const float* p, unsigned long long pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void ld_table2(const double* p, unsigned long long pol, double& a, double& b) {
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(a), "=d"(b) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void ld_table2(const float* p, unsigned long long pol, float& a, float& b) {
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(a), "=f"(b) : "l"(p), "l"(pol));
}

// One cell of the cell-packed table: COUNT = 2^N contiguous elements, loaded 32 bytes at a time (LDG.256 on
// sm_100) or with the widest load the block size allows; plain LRU, no cache-policy operand.
// The same with an L2 prefetch-size hint of 256 bytes: on batches the order probe found coherent the neighbouring cells'
// blocks are the next ones needed, and a first touch then brings four 64-byte blocks into L2 for one DRAM round trip.
template <int COUNT>
__device__ __forceinline__ void ld_cells_wide(const double* p, double (&v)[COUNT]) {
    if constexpr (COUNT >= 4) {
#pragma unroll
        for (int k = 0; k < COUNT; k += 4)
            asm volatile("ld.global.nc.L2::256B.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(v[k]), "=d"(v[k + 1]), "=d"(v[k + 2]), "=d"(v[k + 3]) : "l"(p + k));
    } else {
        asm volatile("ld.global.nc.L2::256B.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
    }
}
template <int COUNT>
__device__ __forceinline__ void ld_cells_wide(const float* p, float (&v)[COUNT]) {
    if constexpr (COUNT >= 8) {
#pragma unroll
        for (int k = 0; k < COUNT; k += 8)
            asm volatile("ld.global.nc.L2::256B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[k]), "=f"(v[k + 1]), "=f"(v[k + 2]), "=f"(v[k + 3]), "=f"(v[k + 4]), "=f"(v[k + 5]),
                           "=f"(v[k + 6]), "=f"(v[k + 7]) : "l"(p + k));
    } else if constexpr (COUNT == 4) {
        asm volatile("ld.global.nc.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L2::256B.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
    }
}
template <int COUNT>
__device__ __forceinline__ void ld_cells(const double* p, double (&v)[COUNT]) {
    if constexpr (COUNT >= 4) {
#pragma unroll
        for (int k = 0; k < COUNT; k += 4)
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(v[k]), "=d"(v[k + 1]), "=d"(v[k + 2]), "=d"(v[k + 3]) : "l"(p + k));
    } else {
        asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
    }
}
template <int COUNT>
__device__ __forceinline__ void ld_cells(const float* p, float (&v)[COUNT]) {
    if constexpr (COUNT >= 8) {
#pragma unroll
        for (int k = 0; k < COUNT; k += 8)
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[k]), "=f"(v[k + 1]), "=f"(v[k + 2]), "=f"(v[k + 3]), "=f"(v[k + 4]), "=f"(v[k + 5]),
                           "=f"(v[k + 6]), "=f"(v[k + 7]) : "l"(p + k));
    } else if constexpr (COUNT == 4) {
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
    } else {
        asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
    }
}

// The same with an L2 cache policy (the slab lists keep the slab being gathered from evict_last: nint_part.cuh).
template <int COUNT>
__device__ __forceinline__ void ld_cells_keep(const double* p, unsigned long long pol, double (&v)[COUNT]) {
    static_assert(COUNT >= 4, "3-D cells");
#pragma unroll
    for (int k = 0; k < COUNT; k += 4)
        asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                     : "=d"(v[k]), "=d"(v[k + 1]), "=d"(v[k + 2]), "=d"(v[k + 3]) : "l"(p + k), "l"(pol));
}
template <int COUNT>
__device__ __forceinline__ void ld_cells_keep(const float* p, unsigned long long pol, float (&v)[COUNT]) {
    static_assert(COUNT >= 8, "3-D cells");
#pragma unroll
    for (int k = 0; k < COUNT; k += 8)
        asm volatile("ld.global.nc.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=f"(v[k]), "=f"(v[k + 1]), "=f"(v[k + 2]), "=f"(v[k + 3]), "=f"(v[k + 4]), "=f"(v[k + 5]),
                       "=f"(v[k + 6]), "=f"(v[k + 7]) : "l"(p + k), "l"(pol));
}

// Two adjacent table elements row[0], row[1] with one aligned vector load; when `row` sits in the upper
// half of its aligned pair a second (predicated) vector load fetches the neighbour.  The table is
// allocated with one pair of slack so the neighbour load never leaves the allocation.
template <class T>
__device__ __forceinline__ void ld_pair(const T* row, unsigned long long pol, T& lo, T& hi) {
    const unsigned long long addr = (unsigned long long)row;
    const bool odd = (addr & sizeof(T)) != 0;
    const T* base = (const T*)(addr & ~(unsigned long long)(2 * sizeof(T) - 1));
    T a0, a1, b0 = T(0), b1 = T(0);
    ld_table2(base, pol, a0, a1);
    if (odd) ld_table2(base + 2, pol, b0, b1);
    lo = odd ? a1 : a0;
    hi = odd ? b0 : a1;
    (void)b1;
}

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
    NINT_ASSERT(b >= 0 && b < a.M && lo >= 0 && n >= 0 && lo + n <= a.L);
    while (n > 0) {
        const int half = n >> 1;
        const bool lt = a.g(lo + half) < p;
        lo = lt ? lo + half + 1 : lo;
        n = lt ? n - half - 1 : half;
    }
    return lo;
}

// a / w, correctly rounded, for a cell whose reciprocal r = RN(1/w) was prepared on the host.
// q0 = RN(a*r) is within 2 ulp of a/w; one FMA residual step makes it faithful and a second one
// (Markstein's theorem, r being the correctly rounded reciprocal) makes it the correctly rounded
// quotient, i.e. bit-identical to the IEEE division the reference performs.  Only valid inside the
// exponent window checked by in_div_window(); nint_selftest_divide() compares it with `a / w`.
template <class T>
__device__ __forceinline__ T div_markstein(T a, T w, T r);
template <>
__device__ __forceinline__ double div_markstein<double>(double a, double w, double r) {
    double q = a * r;
    double e = __fma_rn(-q, w, a);
    q = __fma_rn(e, r, q);
    e = __fma_rn(-q, w, a);
    return __fma_rn(e, r, q);
}
template <>
__device__ __forceinline__ float div_markstein<float>(float a, float w, float r) {
    float q = a * r;
    float e = __fmaf_rn(-q, w, a);
    q = __fmaf_rn(e, r, q);
    e = __fmaf_rn(-q, w, a);
    return __fmaf_rn(e, r, q);
}
// |a| in [2^-500, 2^500) for f64, [2^-60, 2^60) for f32: with w in the same window (host side, r != 0)
// every intermediate of div_markstein() stays normal.  in_div_window_pos() is the same test for an `a`
// already known to satisfy 0 < a <= w with w inside the window: only the lower bound can fail.
__device__ __forceinline__ bool in_div_window_pos(double a) { return __double2hiint(a) >= 0x20b00000; }
__device__ __forceinline__ bool in_div_window_pos(float a) { return __float_as_int(a) >= 0x21800000; }
__device__ __forceinline__ bool in_div_window(double a) {
    const unsigned hi = (unsigned)__double2hiint(a) & 0x7fffffffu;
    return (hi - 0x20b00000u) < 0x3e800000u;
}
__device__ __forceinline__ bool in_div_window(float a) {
    const unsigned bits = (unsigned)__float_as_int(a) & 0x7fffffffu;
    return (bits - 0x21800000u) < 0x3c000000u;
}
template <class T>
__device__ __forceinline__ T div_cell(T a, T w, T r) {
    if (in_div_window(a) && r != T(0)) return div_markstein<T>(a, w, r);
    return a / w;
}

// lib.rs:81-83 with Rust's float rem_euclid (fmod, then `+ |b|` when negative)
template <class T>
__device__ __forceinline__ T wrap_coord(T x, T lo, T hi) {
    const T a = x - lo;
    const T b = hi - lo;
    // fmod(a, b) is exact.  Within two spans of the grid it needs no loop: |a| < |b| leaves a, and for
    // |b| <= |a| < 2|b| the remainder |a| - |b| is an exact subtraction (Sterbenz), with the sign of a.  NaN, inf,
    // b == 0 and everything further out fail the test and take the library routine.
    const T fa = fabs(a), fb = fabs(b);
    T r;
    if (fa < fb) r = a;
    else if (fa < fb + fb) r = copysign(fa - fb, a);
    else r = fmod(a, b);
    if (r < T(0)) r = r + fb;
    return lo + r;
}

// num_traits::clamp
template <class T>
__device__ __forceinline__ T clamp_coord(T x, T lo, T hi) {
    return (x < lo) ? lo : ((x > hi) ? hi : x);
}

// Axis d as seen by the kernel.  Built at the point of use from the kernel parameters (constant bank),
// so that nothing about the axes has to stay in registers across the point loop.
template <class T>
__device__ __forceinline__ AxisRef<T> load_axis(const Params<T>& P, const unsigned char* pack, int d) {
    AxisRef<T> a;
    a.rec = reinterpret_cast<const T*>(pack + P.node_off[d]);
    a.lut = reinterpret_cast<const uint2*>(P.pack + P.lut_off[d]);
    a.L = P.L[d];
    a.M = P.M[d];
    a.flags = P.flags[d];
    a.g0 = P.g0[d];
    a.gl = P.gl[d];
    a.inv_w = P.inv_w[d];
    a.inv_h = P.inv_h[d];
    return a;
}
template <class T, int N>
__device__ __forceinline__ void load_axes(const Params<T>& P, const unsigned char* pack, AxisRef<T> (&ax)[N]) {
#pragma unroll
    for (int d = 0; d < N; ++d) ax[d] = load_axis<T>(P, pack, d);
}

template <class T, int N>
struct Point {
    T v[N];
};
template <class T>
struct Outcome {
    T value;
    unsigned status;
};

// The 2^(N-1) rows of a cell in the plain table (rows run along the last axis, stride 1).
// Row r: bit (N-2-d) of r selects the upper node of axis d, so axis 0 is the most significant bit.
template <class T, int N>
__device__ __forceinline__ void gather_rows(const T* base, const long long (&step)[N], unsigned long long pol,
                                            bool last_fixed, T (&cube)[1 << N]) {
    constexpr int kRows = 1 << (N - 1);
    const T* rowp[kRows];
    rowp[0] = base;
#pragma unroll
    for (int r = 1; r < kRows; ++r) {
        const int low = r & (-r);
        int d = 0;
#pragma unroll
        for (int k = 0; k < N - 1; ++k)
            if ((low >> (N - 2 - k)) & 1) d = k;
        rowp[r] = rowp[r & (r - 1)] + step[d];
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        T lo, hi;
        ld_pair(rowp[r], pol, lo, hi);
        cube[2 * r] = lo;
        cube[2 * r + 1] = last_fixed ? lo : hi;
    }
}

// ---------------------------------------------------------------------------------------------
// General path.  `status` is set to NINT_STATUS_PANIC where the reference would index out of
// bounds; the returned value is then unspecified (the caller writes a quiet NaN).
//
// COLLAPSE  an axis whose coordinate equals a node is *selected*, not blended
//           (one/strategies.rs:14-16 for 1-D, n/strategies.rs:36-46 for N-D).
// LEN1      InterpND returns early when a single element is left after collapsing
//           (n/strategies.rs:47-50), which also covers un-collapsed axes of length 1.
// ---------------------------------------------------------------------------------------------
template <class T, int N, bool COLLAPSE, bool LEN1>
__device__ __forceinline__ T eval_linear(const AxisRef<T> (&ax)[N], const Params<T>& P, unsigned long long pol,
                                         const T (&q)[N], unsigned& status) {
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
        if (COLLAPSE) hit = !isnan && c < L && ax[d].g(c) == p;
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
                const T gl_ = ax[d].g(l);
                const T gu_ = ax[d].g(l + 1);
                idx[d] = l;
                t[d] = div_cell(p - gl_, gu_ - gl_, ax[d].rcp(l));
            }
        }
    }
    long long base = 0;
    long long step[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
        base += (long long)idx[d] * P.vstride[d];
        step[d] = fixed[d] ? 0 : P.vstride[d];
    }
    if (LEN1 && single) return ld_table(P.values + base, pol);
    if (panic) {
        status = NINT_STATUS_PANIC;
        return T(0);
    }
    T cube[1 << N];
#ifdef NINT_BOUNDS_CHECK
    {
        long long top = base;
        for (int d = 0; d < N; ++d) top += step[d];
        NINT_ASSERT(base >= 0 && (unsigned long long)top < P.values_len);
    }
#endif
    gather_rows<T, N>(P.values + base, step, pol, fixed[N - 1], cube);
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
__device__ __forceinline__ T eval_index(const AxisRef<T> (&ax)[N], const Params<T>& P, unsigned long long pol,
                                        const T (&q)[N], unsigned& status) {
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
        if (COLLAPSE) hit = !isnan && c < L && ax[d].g(c) == p;
        idx[d] = hit ? c : 0;
        if (!hit) {
            single = single && (L == 1);
            if (STRAT == NINT_NEAREST) {
                // find_nearest_index: `== last` -> L-2; beyond last or NaN -> L-1, then g[L] panics
                if (isnan || L < 2 || (c >= L)) {
                    panic = true;
                } else {
                    const int l = (p == ax[d].gl) ? L - 2 : max(c, 1) - 1;
                    const T gl_ = ax[d].g(l);
                    const T gu_ = ax[d].g(l + 1);
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
    for (int d = 0; d < N; ++d) off += (long long)idx[d] * P.vstride[d];
    NINT_ASSERT(off >= 0 && (unsigned long long)off < P.values_len);
    if (LEN1 && single) return ld_table(P.values + off, pol);  // un-collapsed length-1 axes sit at index 0
    if (panic) {
        status = NINT_STATUS_PANIC;
        return T(0);
    }
    return ld_table(P.values + off, pol);
}

// Front-end + strategy for one point, exactly as the reference orders them.  Out of line on purpose:
// it keeps the register allocation of the fast path small.
template <class T, int N, bool IS_ND, int STRAT>
__device__ __noinline__ Outcome<T> general_point(const Params<T>& P, const unsigned char* pack, Point<T, N> qin) {
    AxisRef<T> ax[N];
    load_axes<T, N>(P, pack, ax);
    const unsigned long long pol = make_table_policy(P.l2_frac);
    constexpr bool kCollapse = IS_ND || (N == 1);
    constexpr bool kLen1 = IS_ND;
    T q[N];
    bool oob = false;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        q[d] = qin.v[d];
        // !(first..=last).contains(p): NaN is out of bounds (three/mod.rs:198-201)
        oob = oob || !(ax[d].g0 <= q[d] && q[d] <= ax[d].gl);
    }
    unsigned status = NINT_STATUS_OK;
    T result = T(0);
    bool done = false;
    if (oob) {
        const int extrap = P.extrap;
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
        if (STRAT == NINT_LINEAR) result = eval_linear<T, N, kCollapse, kLen1>(ax, P, pol, q, status);
        else result = eval_index<T, N, kCollapse, kLen1, STRAT>(ax, P, pol, q, status);
    }
    Outcome<T> o;
    o.value = result;
    o.status = status;
    return o;
}

// Extrapolate::Wrap applied to every axis (three/mod.rs:215-224); out of line because of fmod.
template <class T, int N>
__device__ __noinline__ Point<T, N> wrap_point(const Params<T>& P, Point<T, N> q) {
    Point<T, N> w;
#pragma unroll
    for (int d = 0; d < N; ++d) w.v[d] = wrap_coord(q.v[d], P.g0[d], P.gl[d]);
    return w;
}

// Fast path, phase 1: the cell guess of one axis.  Uniform axes scale the coordinate; the others read a
// bucket of the table and count the (at most two) nodes below the point.  The guess is only a guess:
// verify_cell() decides.  Returns false when the axis cannot use the fast path at all.
template <class T, bool UNI>
__device__ __forceinline__ bool guess_cell(const AxisRef<T>& a, const unsigned* guess, T p, int& k) {
    if (UNI) {
        // kernel variant for interpolators whose axes are all uniform and fast-path eligible
        k = to_int_sat((p - a.g0) * a.inv_h);
        k = max(0, min(k, a.L - 2));
        return true;
    }
    bool ok = (a.flags & kAxisFast) != 0;
    if (a.flags & kAxisUniform) {
        k = to_int_sat((p - a.g0) * a.inv_h);
    } else {
        // the cell of the bucket's left edge (already inside [0, L-2]); verify_*_retry() settles whether the point
        // is in it, in the next one, or needs the general path
        int b = to_int_sat((p - a.g0) * a.inv_w);
        b = max(0, min(b, a.M - 1));
        NINT_ASSERT(b >= 0 && b < a.M);
        k = (int)guess[b];  // bucket -> the cell holding the bucket's left edge
    }
    k = max(0, min(k, a.L - 2));
    return ok;
}

// Fast path, phase 2: g[k] < p <= g[k+1] (strictly below g[k+1] where an exact node hit must collapse the
// axis).  A verified cell is the cell the reference picks: find_nearest_index returns the cell to the left
// of an exact hit and maps `== last` to L-2, both of which satisfy the same inequality when the nodes are
// strictly increasing; and it implies that p is in range and not NaN.
// FIRST (hard-coded Nearest under Clamp only): p == g[0] in the first cell is accepted too.  Clamping produces that
// coordinate for every point below the grid, and find_nearest_index(g, g[0]) = 0 with (p - g[0]) < (g[1] - p), so the
// reference picks index 0 (three/strategies.rs:69-74) -- what the fast path computes from k = 0.  Without the clause a
// batch with a few per cent of its points below the grid sends a lane of nearly every warp to the general path.
template <class T, bool STRICT, bool FIRST = false>
__device__ __forceinline__ bool verify_cell(const AxisRef<T>& a, T p, int k, T& G0, T& G1, T& R) {
    NINT_ASSERT(k >= 0 && k <= a.L - 2);
    ld_rec(a.rec + 2 * k, G0, R);
    G1 = a.g(k + 1);
    const bool lower = (G0 < p) || (FIRST && !STRICT && k == 0 && p == G0);
    return lower && (STRICT ? (p < G1) : (p <= G1));
}

// Verification for Linear plus the division-window test on a = p - g[k].  Besides g[k] < p <= g[k+1] the
// hard-coded interpolators also accept the in-range point p == g[0] in the first cell, for which the reference
// computes t = 0/w = +0 (three/strategies.rs:16-27) -- div_markstein(0, w, r) returns exactly that.  Where axes
// collapse, an exact hit selects the node instead and is left to the general path.
// kAxisFast guarantees every cell has a stored reciprocal (width inside the window) and the verified bracket
// gives 0 <= a <= w, so only a tiny non-zero `a` can leave the window.
// FIRST enables the p == g[0] clause; only the Clamp kernels use it (clamping produces that coordinate for
// every point below the grid), elsewhere an exact lower-bound query is rare enough for the general path.
template <class T, bool STRICT, bool FIRST>
__device__ __forceinline__ bool verify_cell_linear(const AxisRef<T>& ax, T p, int k, T& a, T& w, T& R) {
    T G0, G1;
    NINT_ASSERT(k >= 0 && k <= ax.L - 2);
    ld_rec(ax.rec + 2 * k, G0, R);
    G1 = ax.g(k + 1);
    a = p - G0;
    w = G1 - G0;
    // a = p - g[k] >= 2^-500 (tested on the high word) already implies g[k] < p; NaN and +inf pass that test
    // but fail the upper bound
    if (STRICT) return (p < G1) && in_div_window_pos(a);
    if (!FIRST) return (p <= G1) && in_div_window_pos(a);
    const bool first = (k == 0) && (p == G0);
    return (in_div_window_pos(a) || first) && (p <= G1);
}

// A non-uniform axis guesses the cell that holds its bucket's left edge, so the point may also sit in the next cell
// (buckets are narrower than almost every cell): one retry there, for the lanes that need it, before the general
// path.  Uniform axes guess by formula and are right or on a node.
template <class T, bool STRICT, bool FIRST, bool UNI>
__device__ __forceinline__ bool verify_linear_retry(const AxisRef<T>& ax, T p, int& k, T& a, T& w, T& R) {
    bool ok = verify_cell_linear<T, STRICT, FIRST>(ax, p, k, a, w, R);
    if (!UNI) {
        if (!ok && !(ax.flags & kAxisUniform) && k < ax.L - 2 && p > ax.g(k + 1)) {
            ++k;
            ok = verify_cell_linear<T, STRICT, FIRST>(ax, p, k, a, w, R);
        }
    }
    return ok;
}
template <class T, bool STRICT, bool UNI, bool FIRST = false>
__device__ __forceinline__ bool verify_retry(const AxisRef<T>& ax, T p, int& k, T& G0, T& G1, T& R) {
    bool ok = verify_cell<T, STRICT, FIRST>(ax, p, k, G0, G1, R);
    if (!UNI) {
        if (!ok && !(ax.flags & kAxisUniform) && k < ax.L - 2 && p > G1) {
            ++k;
            ok = verify_cell<T, STRICT, FIRST>(ax, p, k, G0, G1, R);
        }
    }
    return ok;
}

// Cell guess, verification and weights for axes whose nodes are exactly g0 + k * h (checked bit for bit on the host,
// Params::h): the two nodes of the guessed cell are computed instead of loaded, and t = a / w is the IEEE division the
// reference performs.  No shared-memory traffic at all: where the L1 data pipe is the limit (queries that are random
// inside a small region: one wavefront per lane and load) the node records are 40 % of the wavefronts.
template <class T, int N, bool IS_ND, bool FIRST>
__device__ __forceinline__ bool fast_weights_analytic(const Params<T>& P, const T (&p)[N], int (&k)[N], T (&t)[N]) {
    constexpr bool kCollapse = IS_ND || (N == 1);
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        int kd = to_int_sat((p[d] - P.g0[d]) * P.inv_h[d]);
        kd = max(0, min(kd, P.L[d] - 2));
        const T G0 = P.g0[d] + (T)kd * P.h[d];
        const T G1 = P.g0[d] + (T)(kd + 1) * P.h[d];
        const T a = p[d] - G0;
        const T w = G1 - G0;
        // g[k] < p <= g[k+1] (strictly below where an exact hit collapses the axis); the hard-coded interpolators also
        // take p == g[0] in the first cell under Clamp, where the reference computes t = 0 / w
        const bool lower = (G0 < p[d]) || (FIRST && !kCollapse && kd == 0 && p[d] == G0);
        const bool upper = kCollapse ? (p[d] < G1) : (p[d] <= G1);
        ok = ok && lower && upper;
        t[d] = a / w;
        k[d] = kd;
    }
    return ok;
}

// The blend of the 2^N corners with the stage widths as compile-time constants.  Four and five axes: with the loop form
// below (bound 1 << (N - 1 - d) inside an unrolled loop over d) the compiler kept the 16- and 8-wide stages as loops over
// a copy of the corners in local memory (ncu: STL.128 / LDL.128 around the gather, 256 bytes of stack).
template <class T, int N, int D = 0>
__device__ __forceinline__ void blend_stages(T (&cube)[1 << N], const T (&t)[N]) {
    if constexpr (D < N) {
        constexpr int half = 1 << (N - 1 - D);
        const T td = t[D];
        const T sd = T(1) - td;
#pragma unroll
        for (int j = 0; j < half; ++j) cube[j] = cube[j] * sd + cube[half + j] * td;
        blend_stages<T, N, D + 1>(cube, t);
    }
}

// The whole fast path for one point; false when anything about it needs the general path.
// PLAIN: 0 = Linear gathers from the cell-packed table when there is one, 1 = from the plain table (slab passes),
// 2 = from the cell-packed table with the L2 policy `pol` (slab lists; 3-D only)
template <class T, int N, bool IS_ND, int STRAT, bool UNI, bool FIRST, int PLAIN = 0, bool ANA = false>
__device__ __forceinline__ bool fast_point(const Params<T>& P, const unsigned char* pack, unsigned long long pol,
                                           const T (&p)[N], T& result) {
    constexpr bool kCollapse = IS_ND || (N == 1);
    int k[N];
    bool ok = true;
    if (ANA && STRAT == NINT_LINEAR) {
        T t[N];
        ok = fast_weights_analytic<T, N, IS_ND, FIRST>(P, p, k, t);
        if (N >= 4 && !ok) return false;
        T cube[1 << N];
        if (PLAIN != 1 && P.cells != nullptr) {
            unsigned cell = 0;
#pragma unroll
            for (int d = 0; d < N; ++d) cell += (unsigned)k[d] * (unsigned)P.cstride[d];
            NINT_ASSERT((unsigned long long)cell < P.ncells);
            if constexpr (PLAIN == 2 && N == 3) ld_cells_keep<(1 << N)>(P.cells + (size_t)cell * (1 << N), pol, cube);
            else ld_cells<(1 << N)>(P.cells + (size_t)cell * (1 << N), cube);
        } else {
            long long base = 0;
            long long step[N];
#pragma unroll
            for (int d = 0; d < N; ++d) {
                base += (long long)k[d] * P.vstride[d];
                step[d] = P.vstride[d];
            }
#ifdef NINT_BOUNDS_CHECK
            {
                long long top = base;
                for (int d = 0; d < N; ++d) top += step[d];
                NINT_ASSERT(base >= 0 && (unsigned long long)top < P.values_len);
            }
#endif
            gather_rows<T, N>(P.values + base, step, pol, false, cube);
        }
#pragma unroll
        for (int d = 0; d < N; ++d) {
            const int half = 1 << (N - 1 - d);
            const T td = t[d];
            const T sd = T(1) - td;
#pragma unroll
            for (int j = 0; j < half; ++j) cube[j] = cube[j] * sd + cube[half + j] * td;
        }
        result = cube[0];
        return ok;
    }
    if (ANA && STRAT != NINT_LINEAR) {
        // Nearest family on analytic axes: the bracket g[k] < p <= g[k+1] from computed nodes, then the same choice
        // between l and l + 1 (three/strategies.rs:70-74; one/strategies.rs:79, :102)
        long long off = 0;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            int kd = to_int_sat((p[d] - P.g0[d]) * P.inv_h[d]);
            kd = max(0, min(kd, P.L[d] - 2));
            const T G0 = P.g0[d] + (T)kd * P.h[d];
            const T G1 = P.g0[d] + (T)(kd + 1) * P.h[d];
            const bool lower = (G0 < p[d]) || (FIRST && STRAT == NINT_NEAREST && !kCollapse && kd == 0 && p[d] == G0);
            ok = ok && lower && (kCollapse ? (p[d] < G1) : (p[d] <= G1));
            int idx = kd;
            if (STRAT == NINT_NEAREST) idx += ((p[d] - G0) < (G1 - p[d])) ? 0 : 1;
            if (STRAT == NINT_RIGHT_NEAREST) idx += 1;
            off += (long long)idx * P.vstride[d];
        }
        NINT_ASSERT(off >= 0 && (unsigned long long)off < P.values_len);
        result = ld_table(P.values + off, pol);  // idx <= L-1 on every axis: inside the table even if !ok
        return ok;
    }
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const unsigned* guess = UNI ? nullptr : reinterpret_cast<const unsigned*>(pack + P.guess_off[d]);
        ok = guess_cell<T, UNI>(load_axis<T>(P, pack, d), guess, p[d], k[d]) && ok;
    }
    if (!ok) return false;  // also keeps the speculative loads below inside the table (every axis has >= 2 nodes)

    if (STRAT == NINT_LINEAR) {
        // Verify and compute the weights first, then gather and blend: keeping the 2^N corner registers out of
        // the weight computation lets the 3-D f64 kernel fit 64 registers (4 CTAs/SM), and on B200 the extra
        // resident warps hide the table-load latency better than issuing the loads early does (measured).
        T t[N];
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T a, w, r;
            ok = verify_cell_linear<T, kCollapse, FIRST>(load_axis<T>(P, pack, d), p[d], k[d], a, w, r) && ok;
            t[d] = div_markstein<T>(a, w, r);
        }
        if (!UNI && !ok) {
            // A non-uniform axis guesses the cell that holds its bucket's left edge, so the point may sit in the
            // next cell (buckets are narrower than almost every cell).  Off the straight-line path: the lanes that
            // failed go round once more with that cell before the general path gets them.
            ok = true;
#pragma unroll
            for (int d = 0; d < N; ++d) {
                T a, w, r;
                ok = verify_linear_retry<T, kCollapse, FIRST, UNI>(load_axis<T>(P, pack, d), p[d], k[d], a, w, r) && ok;
                t[d] = div_markstein<T>(a, w, r);
            }
        }
        // N >= 4: no speculative gather for a point that is going to the general path anyway; with mostly
        // out-of-range batches (Fill, Error) those 2^N-corner loads would be the bulk of the table traffic
        // (32^5 f32, 87 % out of range, Fill: 16.2 vs 9.5 Gpoints/s).  Below that the branch costs the
        // cell-ordered 3-D kernel 3 % and is left out.
        if (N >= 4 && !ok) return false;
        T cube[1 << N];
        // the all-uniform kernels are only dispatched when the cell-packed table exists (nint_launch.cuh)
        if (PLAIN != 1 && (UNI || P.cells != nullptr)) {
            unsigned cell = 0;  // build_cells() only packs tables with fewer than 2^31 cells
#pragma unroll
            for (int d = 0; d < N; ++d) cell += (unsigned)k[d] * (unsigned)P.cstride[d];
            NINT_ASSERT((unsigned long long)cell < P.ncells);
            if constexpr (PLAIN == 2 && N == 3) {
                ld_cells_keep<(1 << N)>(P.cells + (size_t)cell * (1 << N), pol, cube);
            } else {
                if (P.wide_lines) ld_cells_wide<(1 << N)>(P.cells + (size_t)cell * (1 << N), cube);
                else ld_cells<(1 << N)>(P.cells + (size_t)cell * (1 << N), cube);
            }
        } else {
            long long base = 0;
            long long step[N];
#pragma unroll
            for (int d = 0; d < N; ++d) {
                base += (long long)k[d] * P.vstride[d];
                step[d] = P.vstride[d];
            }
#ifdef NINT_BOUNDS_CHECK
            {
                long long top = base;
                for (int d = 0; d < N; ++d) top += step[d];
                NINT_ASSERT(base >= 0 && (unsigned long long)top < P.values_len);
            }
#endif
            gather_rows<T, N>(P.values + base, step, pol, false, cube);
        }
        // reduce axis 0 first, then 1, ...: v_l * (1 - t) + v_u * t, each op rounded separately
        if constexpr (N == 4 || N == 5) {  // (six axes and more: 2^N corners do not fit the register file either way)
            blend_stages<T, N>(cube, t);
        } else {
#pragma unroll
            for (int d = 0; d < N; ++d) {
                const int half = 1 << (N - 1 - d);
                const T td = t[d];
                const T sd = T(1) - td;
#pragma unroll
                for (int j = 0; j < half; ++j) cube[j] = cube[j] * sd + cube[half + j] * td;
            }
        }
        result = cube[0];
        return ok;
    } else {
        long long off = 0;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T G0, G1, r;
            ok = verify_cell<T, kCollapse, FIRST && STRAT == NINT_NEAREST>(load_axis<T>(P, pack, d), p[d], k[d], G0, G1, r) && ok;
            int idx = k[d];
            if (STRAT == NINT_NEAREST) idx += ((p[d] - G0) < (G1 - p[d])) ? 0 : 1;  // tie -> upper (three/strategies.rs:70-74)
            if (STRAT == NINT_RIGHT_NEAREST) idx += 1;
            off += (long long)idx * P.vstride[d];
        }
        if (!UNI && !ok) {  // once more with the next cell of the non-uniform axes that missed (see Linear above)
            ok = true;
            off = 0;
#pragma unroll
            for (int d = 0; d < N; ++d) {
                T G0, G1, r;
                ok = verify_retry<T, kCollapse, UNI, FIRST && STRAT == NINT_NEAREST>(load_axis<T>(P, pack, d), p[d], k[d], G0, G1, r) && ok;
                int idx = k[d];
                if (STRAT == NINT_NEAREST) idx += ((p[d] - G0) < (G1 - p[d])) ? 0 : 1;
                if (STRAT == NINT_RIGHT_NEAREST) idx += 1;
                off += (long long)idx * P.vstride[d];
            }
        }
        NINT_ASSERT(off >= 0 && (unsigned long long)off < P.values_len);
        result = ld_table(P.values + off, pol);  // idx <= L-1 on every axis: inside the table even if !ok
        return ok;
    }
}

// How the Extrapolate policy enters the fast path (kernel template argument).
constexpr int kPrePlain = 0;  // nothing to do up front (Error, Fill, Enable: out-of-range points take the general path)
constexpr int kPreClamp = 1;  // Clamp: clamp every coordinate first, then bracket
constexpr int kPreWrap = 2;   // Wrap: points that fail the fast path are wrapped (if out of range) and retried

// ---------------------------------------------------------------------------------------------
// The kernel: persistent CTAs, grid-stride over the batch, one point per thread per trip, the next
// trip's coordinates prefetched while the current point is evaluated.
// ---------------------------------------------------------------------------------------------
// Resident CTAs per SM the register allocator must leave room for.  Measured on B200 (tools/exp_ctas.sh):
// with the weights computed before the table loads the 3-D f64 fast path fits 64 registers, and 4 CTAs of 256
// threads (32 warps/SM) beat 3 CTAs at 80 registers (136 vs 115 Gpoints/s on cell-ordered queries); 128
// registers leave too few warps to hide the table-load latency.  NINT_MIN_CTAS overrides for tuning.
constexpr int min_ctas(int n, int elem_bytes) {
#ifdef NINT_MIN_CTAS
    if (n <= 3) return NINT_MIN_CTAS;
#endif
    if (n <= 2) return 4;
    if (n == 3) return 4;
    // 5-D f32: 80 registers since the blend keeps the 32 corners in registers (blend_stages); three CTAs measured on
    // 32^5, 2e8 points: in-range compaction pass (Fill) 57.9 -> 63.7 Gpoints/s, single pass under Wrap 9.3 (unchanged)
#ifndef NINT_ND_CTAS
#define NINT_ND_CTAS 3
#endif
    if (n == 5 && elem_bytes == 4) return NINT_ND_CTAS;
    return (n == 4) ? 2 : 1;
}

template <class T, int N, bool IS_ND, int STRAT, bool SMEM, int PRE, bool UNI, bool ANA = false>
__global__ void __launch_bounds__(kBlock, min_ctas(N, (int)sizeof(T))) interp_kernel(const __grid_constant__ Params<T> P) {
    extern __shared__ __align__(16) unsigned char smem_pack[];
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;  // the other query-order mode runs this batch
    const unsigned char* pack;
    if (SMEM) {
        const uint4* src = reinterpret_cast<const uint4*>(P.pack);
        uint4* dst = reinterpret_cast<uint4*>(smem_pack);
        for (unsigned i = threadIdx.x; i < P.pack_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
        pack = smem_pack;
    } else {
        pack = P.pack;
    }
    const unsigned long long pol = make_table_policy(P.l2_frac);

    // per-CTA failure counters live in shared memory (failures are the rare case; no registers held)
    __shared__ unsigned long long s_info[4];  // n_error, n_panic, first_error, first_panic
    if (threadIdx.x < 4) s_info[threadIdx.x] = (threadIdx.x < 2) ? 0ull : ~0ull;
    __syncthreads();

    // Grid-stride over the batch, or -- for batches the order probe found coherent -- one contiguous range per CTA:
    // consecutive queries then share cells inside one SM's L1 instead of being dealt out to all SMs (queries grouped
    // by 8^3-cell tiles: 81 -> see DESIGN.md).  `n` is the end of this CTA's range; <= 2^30, so i + stride cannot wrap.
    unsigned n = (unsigned)P.n;
    unsigned stride = gridDim.x * blockDim.x;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (P.blocked) {
        const unsigned chunk = ((n + gridDim.x - 1) / gridDim.x + kBlock - 1) / kBlock * kBlock;
        const unsigned long long lo = (unsigned long long)blockIdx.x * chunk;
        i = (unsigned)min(lo, (unsigned long long)n) + threadIdx.x;
        n = (unsigned)min(lo + chunk, (unsigned long long)n);
        stride = blockDim.x;
    }
    Point<T, N> q;
    if (i < n) {
#pragma unroll
        for (int d = 0; d < N; ++d) q.v[d] = ld_stream(P.coords[d] + (size_t)i * P.coord_stride);
    }
    // evaluates point i (coordinates q) and stores its result
    auto process = [&](unsigned i, const Point<T, N>& q) {
        T p[N];
        bool oob = false;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            // Clamp touches every axis as soon as one is out of range and leaves in-range coordinates
            // alone, so clamping up front is the same map (three/mod.rs:205-214)
            p[d] = (PRE == kPreClamp) ? clamp_coord(q.v[d], P.g0[d], P.gl[d]) : q.v[d];
            if (PRE == kPreWrap) oob = oob || !(P.g0[d] <= q.v[d] && q.v[d] <= P.gl[d]);
        }
        if (PRE == kPreWrap) {
            // Wrap re-maps every axis as soon as one coordinate is out of range (three/mod.rs:215-224)
            if (oob) {
                const Point<T, N> w = wrap_point<T, N>(P, q);
#pragma unroll
                for (int d = 0; d < N; ++d) p[d] = w.v[d];
            }
        }
        T result;
        const bool ok = fast_point<T, N, IS_ND, STRAT, UNI, PRE == kPreClamp, false, ANA>(P, pack, pol, p, result);
        // 3-D f64 is the kernel that lives at the edge of its 64 registers: there the common case stores on its
        // own, so that no status or result register is shared with the slow path (cell-ordered queries 152 -> 157
        // Gpoints/s).  The other kernels measured faster with the single store after the join (2-D: 84 vs 72).
        constexpr bool kSplitStore = (N == 3 && sizeof(T) == 8);
        if (kSplitStore && ok) {
            st_stream(P.out + i, result);
            if (P.status) P.status[i] = (uint8_t)NINT_STATUS_OK;
            return;
        }
        unsigned status = NINT_STATUS_OK;
        if (!ok) {
            const Outcome<T> o = general_point<T, N, IS_ND, STRAT>(P, pack, q);
            result = o.value;
            status = o.status;
            if (status != NINT_STATUS_OK) {
                result = quiet_nan<T>();
                const int slot = (status == NINT_STATUS_PANIC) ? 1 : 0;
                atomicAdd(&s_info[slot], 1ull);
                atomicMin(&s_info[2 + slot], (unsigned long long)i + P.index_base);
            }
        }
        st_stream(P.out + i, result);
        if (P.status) P.status[i] = (uint8_t)status;
    };
    // requests the coordinates of the trip after point i into `next`; false when there is none
    auto look_ahead = [&](unsigned i, Point<T, N>& next) {
        const unsigned in = i + stride;
        // Nearest family: few instructions per point, so a trip is shorter than a DRAM round trip and one trip of
        // look-ahead does not cover it.  The trip after the next starts its way to L2 (one request per 128-byte
        // line, no register held): cell-ordered 3-D Nearest 146 -> 175 (f64), 177 -> 217 (f32) Gpoints/s.  Linear
        // trips are long enough; there the extra requests cost 1-2 %.
        if (STRAT != NINT_LINEAR) {
            const unsigned i2 = in + stride;
            if (i2 < n && (threadIdx.x & ((128u / (unsigned)sizeof(T)) - 1u)) == 0u) {
#pragma unroll
                for (int d = 0; d < N; ++d) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.coords[d] + (size_t)i2 * P.coord_stride));
            }
        }
        if (in >= n) return false;
#pragma unroll
        for (int d = 0; d < N; ++d) next.v[d] = ld_stream(P.coords[d] + (size_t)in * P.coord_stride);
        return true;
    };
    // Linear: two trips per turn with the roles of the coordinate registers swapped, so that there are no copies
    // at the back edge (cell-ordered 3-D f32 175 -> 195 Gpoints/s, f64 and 2-D unchanged to +2 %).  The Nearest
    // kernels are 3 % faster with the plain loop.
    constexpr bool kTwoTrips = (STRAT == NINT_LINEAR);
    Point<T, N> q2;
    while (i < n) {
        const bool more = look_ahead(i, q2);
        process(i, q);
        i += stride;
        if (!more) break;
        if (kTwoTrips) {
            const bool more2 = look_ahead(i, q);
            process(i, q2);
            i += stride;
            if (!more2) break;
        } else {
            q = q2;
        }
    }

    if (P.info) {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_info[0]) {
                atomicAdd((unsigned long long*)&P.info->n_error, s_info[0]);
                atomicMin((unsigned long long*)&P.info->first_error, s_info[2]);
            }
            if (s_info[1]) {
                atomicAdd((unsigned long long*)&P.info->n_panic, s_info[1]);
                atomicMin((unsigned long long*)&P.info->first_panic, s_info[3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 1-D Linear with everything a point needs in shared memory (one/strategies.rs:9-29).
//
// The general kernel reads, per point, a guess-table entry, the node record {g[k], 1/w}, the next node and the value
// pair v[k], v[k+1] -- the last from global memory, where 32 lanes in 32 random cells are 32 L1 wavefronts; on a
// 1000-node axis the L1 data pipe ran at 97 % of its wavefront rate (202 Gpoints/s, half the HBM roofline).  Here the
// axis and the values are interleaved as {g[l], v[l]} pairs, staged once per CTA with the reciprocals, so a point is
// two 16-byte and one 8-byte shared-memory loads after the guess.  Same verification (g[k] < p < g[k+1], exact hits
// leave for the general path, which selects the node value) and the same arithmetic as fast_point<N = 1>.
// ---------------------------------------------------------------------------------------------
template <class T, bool IS_ND, int PRE, bool UNI>
__global__ void __launch_bounds__(kBlock, 4) interp1d_fused_kernel(const __grid_constant__ Params<T> P) {
    constexpr int N = 1;
    extern __shared__ __align__(16) unsigned char smem_pack[];
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;  // the other query-order mode runs this batch
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.fused);
        uint4* dst = reinterpret_cast<uint4*>(smem_pack);
        for (unsigned i = threadIdx.x; i < P.fused_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const T* pairs = reinterpret_cast<const T*>(smem_pack);                      // {g[l], v[l]}, l = 0..L (last: +inf pad)
    const T* rcps = reinterpret_cast<const T*>(smem_pack + P.fused_rcp_off);     // RN(1 / (g[l+1] - g[l]))
    const unsigned* guess = reinterpret_cast<const unsigned*>(smem_pack + P.fused_guess_off);
    __shared__ unsigned long long s_info[4];  // n_error, n_panic, first_error, first_panic
    if (threadIdx.x < 4) s_info[threadIdx.x] = (threadIdx.x < 2) ? 0ull : ~0ull;
    __syncthreads();
    const unsigned n = (unsigned)P.n;
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const int L = P.L[0];
    const T g0 = P.g0[0], gl = P.gl[0];
    T q = T(0), q2 = T(0);
    if (i < n) q = ld_stream(P.coords[0] + (size_t)i * P.coord_stride);
    auto process = [&](unsigned i, T q) {
        T p = (PRE == kPreClamp) ? clamp_coord(q, g0, gl) : q;
        if (PRE == kPreWrap) {
            if (!(g0 <= q && q <= gl)) p = wrap_coord(q, g0, gl);  // lib.rs:81-83 on the single axis (one/mod.rs:189-196)
        }
        int k;
        if (UNI) {
            k = to_int_sat((p - g0) * P.inv_h[0]);
        } else {
            int b = to_int_sat((p - g0) * P.inv_w[0]);
            b = max(0, min(b, P.M[0] - 1));
            k = (int)guess[b];
        }
        k = max(0, min(k, L - 2));
        // rcp_div: t by the stored reciprocal with two FMA corrections (one more 8-byte shared load); otherwise by IEEE
        // division (more FP64 instructions, fewer L1 wavefronts).  Both give RN(a / w).
        const bool rcp_div = P.fused_rcp_off != 0u;
        NINT_ASSERT(k >= 0 && (unsigned)(2 * k + 3) * sizeof(T) < P.fused_bytes);
        T G0, V0, G1, V1, r = T(0);
        ld_rec(pairs + 2 * k, G0, V0);
        ld_rec(pairs + 2 * k + 2, G1, V1);
        if (rcp_div) r = rcps[k];
        T a = p - G0;
        bool ok = (p < G1) && (rcp_div ? in_div_window_pos(a) : (G0 < p));
        if (!UNI) {
            if (!ok && k < L - 2 && p > G1) {  // the bucket's left edge was in the previous cell: once more with the next
                ++k;
                G0 = G1;
                V0 = V1;
                ld_rec(pairs + 2 * k + 2, G1, V1);
                if (rcp_div) r = rcps[k];
                a = p - G0;
                ok = (p < G1) && (rcp_div ? in_div_window_pos(a) : (G0 < p));
            }
        }
        unsigned status = NINT_STATUS_OK;
        T result;
        if (ok) {
            const T t = rcp_div ? div_markstein<T>(a, G1 - G0, r) : a / (G1 - G0);
            result = V0 * (T(1) - t) + V1 * t;  // one/strategies.rs:28
        } else {
            Point<T, N> qq;
            qq.v[0] = q;
            const Outcome<T> o = general_point<T, N, IS_ND, NINT_LINEAR>(P, P.pack, qq);  // axis pack from global memory
            result = o.value;
            status = o.status;
            if (status != NINT_STATUS_OK) {
                result = quiet_nan<T>();
                const int slot = (status == NINT_STATUS_PANIC) ? 1 : 0;
                atomicAdd(&s_info[slot], 1ull);
                atomicMin(&s_info[2 + slot], (unsigned long long)i + P.index_base);
            }
        }
        st_stream(P.out + i, result);
        if (P.status) P.status[i] = (uint8_t)status;
    };
    while (i < n) {
        const unsigned in = i + stride;
        const bool more = in < n;
        if (more) q2 = ld_stream(P.coords[0] + (size_t)in * P.coord_stride);
        process(i, q);
        if (!more) break;
        i = in;
        q = q2;
    }
    if (P.info) {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_info[0]) {
                atomicAdd((unsigned long long*)&P.info->n_error, s_info[0]);
                atomicMin((unsigned long long*)&P.info->first_error, s_info[2]);
            }
            if (s_info[1]) {
                atomicAdd((unsigned long long*)&P.info->n_panic, s_info[1]);
                atomicMin((unsigned long long*)&P.info->first_panic, s_info[3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Slab passes for incoherent queries on tables larger than L2 (Linear).
//
// When consecutive queries land in unrelated cells and the table does not fit L2, every corner load is a DRAM
// (and, for the packed table, a TLB) miss.  The batch is then evaluated in W launches; launch w only evaluates
// the points whose axis-0 cell lies in slab w of the plain table, so the part of the table being gathered fits
// L2.  Warps work on their own: a warp scans a span of kSlabSpan points (axis-0 coordinate only), appends the
// members to its list in shared memory (ballot + popc, no atomics) and evaluates them 32 at a time, so every
// lane is busy whatever the membership pattern; what is left over (< 32) waits in a register for the next span.
// Which mode a batch uses is decided on the device by order_probe_kernel from a sample of the batch; both kinds
// of launch are always enqueued and the ones of the other mode return at once.
// ---------------------------------------------------------------------------------------------
constexpr int kSlabPer = 4;                   // points per lane and span (a larger list only costs L1: shared-memory carve-out)
constexpr unsigned kSlabSpan = 32u * kSlabPer;  // points a warp scans at a time


// Resident CTAs per SM for slab_kernel.  Unlike the single pass, this kernel keeps the scan state next to a whole
// point evaluation: at 64 registers (4 CTAs) it spills 60-170 bytes per thread and runs at 28.1 Gpoints/s on the
// headline workload; at 80 registers (3 CTAs, no spills) 32.7; at 96 (2 CTAs) 31.4; at 48 (5 CTAs) 25.6.
constexpr int slab_ctas(int n, int elem_bytes) { return n <= 3 ? 3 : min_ctas(n, elem_bytes); }
// Requesting the next span's scan one span ahead overlaps its DRAM round trip with the evaluation: +2 % in f32 (50.9 ->
// 52.0 Gpoints/s); in f64 the eight extra registers spill at 80 registers / 3 CTAs and it is slower (31 -> 35 ms).
#ifndef NINT_SLAB_AHEAD
#define NINT_SLAB_AHEAD (sizeof(T) == 4)
#endif
#ifndef NINT_SLAB_ANA_CTAS
#define NINT_SLAB_ANA_CTAS 3  // analytic variant (no node records): tried 4
#endif

// INR (Fill and Error only): one pass over the whole batch in which the members are the points inside the grid on every
// axis; the others are finished by the scan itself (fill value, or NaN + ExtrapolateError status: n/mod.rs:294-298,
// 326-337).  With most of a batch out of range (BASELINE config 4: 87 %) the corner gathers then run with full warps
// instead of four lanes in 32, and nobody calls the out-of-line general path for an out-of-range point.
template <class T, int N, bool IS_ND, int PRE, bool UNI, bool ANA = false, bool INR = false>
__global__ void __launch_bounds__(kBlock, (ANA && N == 3) ? NINT_SLAB_ANA_CTAS : slab_ctas(N, (int)sizeof(T))) slab_kernel(const __grid_constant__ Params<T> P) {
    extern __shared__ __align__(16) unsigned char smem_pack[];
    __shared__ unsigned short s_list[kBlock / 32][kSlabSpan];
    __shared__ unsigned long long s_info[4];  // n_error, n_panic, first_error, first_panic
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.pack);
        uint4* dst = reinterpret_cast<uint4*>(smem_pack);
        for (unsigned i = threadIdx.x; i < P.pack_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const unsigned char* pack = smem_pack;
    const unsigned long long pol = make_table_policy(1.0f);  // the slab stays, the query streams go first
    if (threadIdx.x < 4) s_info[threadIdx.x] = (threadIdx.x < 2) ? 0ull : ~0ull;
    __syncthreads();
    const unsigned n = (unsigned)P.n;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    unsigned short* list = s_list[threadIdx.x >> 5];
    const unsigned nspans = (n + kSlabSpan - 1) / kSlabSpan;
    const unsigned warps = gridDim.x * (kBlock / 32);
    // members that did not fill a group of 32 wait here, one per lane in lanes [0, ncarry)
    unsigned carry = 0, ncarry = 0;

    auto fetch = [&](unsigned i, Point<T, N>& q) {
#pragma unroll
        for (int d = 0; d < N; ++d) q.v[d] = ld_stream(P.coords[d] + (size_t)i * P.coord_stride);
    };
    auto evaluate = [&](unsigned i, const Point<T, N>& q) {
        T p[N];
#pragma unroll
        for (int d = 0; d < N; ++d) p[d] = (PRE == kPreClamp) ? clamp_coord(q.v[d], P.g0[d], P.gl[d]) : q.v[d];
        unsigned status = NINT_STATUS_OK;
        T result;
        const bool ok = fast_point<T, N, IS_ND, NINT_LINEAR, UNI, PRE == kPreClamp, !INR, ANA>(P, pack, pol, p, result);
        if (!ok) {
            const Outcome<T> o = general_point<T, N, IS_ND, NINT_LINEAR>(P, pack, q);
            result = o.value;
            status = o.status;
            if (status != NINT_STATUS_OK) {
                result = quiet_nan<T>();
                const int slot = (status == NINT_STATUS_PANIC) ? 1 : 0;
                atomicAdd(&s_info[slot], 1ull);
                atomicMin(&s_info[2 + slot], (unsigned long long)i + P.index_base);
            }
        }
        st_stream(P.out + i, result);
        if (P.status) P.status[i] = (uint8_t)status;
    };

    // axis-0 coordinates of a span, all requested before the first is used
    auto scan_span = [&](unsigned span, T (&xs)[kSlabPer]) {
#pragma unroll
        for (int r = 0; r < kSlabPer; ++r) {
            const unsigned j = span * kSlabSpan + r * 32u + lane;
            xs[r] = (span < nspans && j < n) ? ld_stream(P.coords[0] + (size_t)j * P.coord_stride) : quiet_nan<T>();
        }
    };
    const unsigned span0 = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    T xs[kSlabPer];
    if (!INR) scan_span(span0, xs);
    for (unsigned span = span0; span < nspans; span += warps) {
        const unsigned base = span * kSlabSpan;
        unsigned count = 0;
        // f32: the next span's scan is requested now, so that its DRAM round trip overlaps this span's evaluation
        T xn[kSlabPer];
        if (!INR && NINT_SLAB_AHEAD) scan_span(span + warps, xn);
        if (INR) {
            // members: inside the grid on every axis (NaN is outside: three/mod.rs:198-201).  The scan loads whole
            // points, normal priority: the members read theirs again a moment later.
#pragma unroll
            for (int r = 0; r < kSlabPer; ++r) {
                const unsigned j = base + r * 32u + lane;
                bool inside = j < n;
                if (j < n) {
#pragma unroll
                    for (int d = 0; d < N; ++d) {
                        const T c = __ldg(P.coords[d] + (size_t)j * P.coord_stride);
                        inside = inside && (P.g0[d] <= c && c <= P.gl[d]);
                    }
                }
                const bool outside = (j < n) && !inside;
                if (outside) {
                    const bool err = P.extrap == NINT_EXTRAPOLATE_ERROR;
                    st_stream(P.out + j, err ? quiet_nan<T>() : P.fill);
                    if (P.status) P.status[j] = (uint8_t)(err ? NINT_STATUS_EXTRAPOLATE_ERROR : NINT_STATUS_OK);
                }
                if (P.extrap == NINT_EXTRAPOLATE_ERROR) {  // warp-uniform
                    const unsigned eb = __ballot_sync(0xffffffffu, outside);
                    if (eb != 0u && lane == 0u) {
                        atomicAdd(&s_info[0], (unsigned long long)__popc(eb));
                        atomicMin(&s_info[2], (unsigned long long)(base + r * 32u + (unsigned)(__ffs(eb) - 1)) + P.index_base);
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, inside);
                if (inside) list[count + __popc(bal & lt)] = (unsigned short)(r * 32u + lane);
                count += __popc(bal);
            }
        } else {
            // members: axis-0 coordinate in (win_lo_x, win_hi_x]; the first window also takes everything below and
            // NaN, the last everything above.  Listed as offsets into the span, in any order.
#pragma unroll
            for (int r = 0; r < kSlabPer; ++r) {
                const unsigned j = base + r * 32u + lane;
                const T x = xs[r];
                const bool above = P.win_first ? true : (x > P.win_lo_x);
                const bool below = P.win_last ? true : !(x > P.win_hi_x);
                const bool mine = (j < n) && above && below;
                const unsigned bal = __ballot_sync(0xffffffffu, mine);
                if (mine) list[count + __popc(bal & lt)] = (unsigned short)(r * 32u + lane);
                count += __popc(bal);
            }
        }
        NINT_ASSERT(count <= kSlabSpan);
        __syncwarp();
        // groups of 32: the waiting members first, the rest from the list; the coordinates of the next group
        // are in flight while the current one is evaluated
        unsigned pos = 0;
        if (ncarry + count >= 32u) {
            unsigned i = (lane < ncarry) ? carry : base + list[lane - ncarry];
            Point<T, N> q;
            fetch(i, q);
            pos = 32u - ncarry;
            ncarry = 0;
            while (count - pos >= 32u) {
                const unsigned in = base + list[pos + lane];
                Point<T, N> qn;
                fetch(in, qn);
                evaluate(i, q);
                i = in;
                q = qn;
                pos += 32u;
            }
            evaluate(i, q);
        }
        const unsigned rem = count - pos;
        if (lane >= ncarry && lane < ncarry + rem) carry = base + list[pos + lane - ncarry];
        ncarry += rem;
        __syncwarp();
        if (!INR) {
            if (NINT_SLAB_AHEAD) {
#pragma unroll
                for (int r = 0; r < kSlabPer; ++r) xs[r] = xn[r];
            } else {
                scan_span(span + warps, xs);
            }
        }
    }
    if (lane < ncarry) {
        Point<T, N> q;
        fetch(carry, q);
        evaluate(carry, q);
    }
    if (P.info) {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_info[0]) {
                atomicAdd((unsigned long long*)&P.info->n_error, s_info[0]);
                atomicMin((unsigned long long*)&P.info->first_error, s_info[2]);
            }
            if (s_info[1]) {
                atomicAdd((unsigned long long*)&P.info->n_panic, s_info[1]);
                atomicMin((unsigned long long*)&P.info->first_panic, s_info[3]);
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

// Builds the cell-packed copy: out[cell * 2^N + corner] = values[(l + corner bits) . vstride].
template <class T>
__global__ void __launch_bounds__(kBlock) pack_cells_kernel(const T* __restrict__ values, T* __restrict__ out, int ndim,
                                                            Params<T> P, unsigned long long total) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const int corners = 1 << ndim;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int corner = (int)(e & (unsigned long long)(corners - 1));
        unsigned long long cell = e >> ndim;
        long long src = 0;
        for (int d = ndim - 1; d >= 0; --d) {
            const unsigned long long nc = (unsigned long long)(P.L[d] - 1);
            const unsigned long long l = cell % nc;
            cell /= nc;
            src += (long long)(l + ((corner >> (ndim - 1 - d)) & 1)) * P.vstride[d];
        }
        NINT_ASSERT(src >= 0 && (unsigned long long)src < P.values_len);
        out[e] = values[src];
    }
}

// Launch interface implemented by the per-(T, family) translation units.
struct LaunchCfg {
    int ndim;
    int is_nd;
    int strategy;
    int smem_axes;  // stage the axis pack in shared memory
    int all_uniform;  // every axis is uniform and fast-path eligible
    int all_analytic; // ... and its nodes are exactly g0 + k * h (launch the single-pass kernels that compute them)
    int slab_analytic; // the same for the slab passes
    int pipe;         // launch interp_pipe_kernel (whole chunks of the batch; hard-coded 2-D/3-D, axes in shared memory)
    int slab;         // launch slab_kernel (Params::win_*) instead of interp_kernel
    int inrange;      // ... in its one-pass form whose members are the in-range points (Fill, Error)
    unsigned smem_bytes;
    int sm_count;
    void* stream;
};

template <class T>
cudaError_t launch_fixed(const Params<T>& P, const LaunchCfg& cfg);
template <class T>
cudaError_t launch_nd(const Params<T>& P, const LaunchCfg& cfg);
template <class T>
cudaError_t launch_pack_cells(const Params<T>& P, int ndim, T* out, unsigned long long total, int sm_count, void* stream);
template <class T>
cudaError_t launch_constant(const T* values, T* out, uint8_t* status, unsigned long long n, int sm_count,
                            void* stream);

}  // namespace nint
