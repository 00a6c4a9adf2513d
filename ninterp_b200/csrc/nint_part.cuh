// nint_part.cuh — slab lists: the route for incoherent 3-D Linear batches on tables far beyond L2.
//
// The slab passes (slab_kernel) keep one third of the PLAIN table in L2 per pass and pay for it three times: every pass
// scans the batch again, its members re-read their coordinates with one L1 wavefront per lane and coordinate, the 8
// corners are 4-6 sixteen-byte gathers, and every pass writes a third of each 32-byte sector of `out` (a DRAM fill and a
// write-back per pass).  ncu: the L1 data pipe at 90 % of its wavefront rate, 45 GB of DRAM traffic per pass for 32 GB
// of algorithmic bytes per batch.
//
// Here the batch is split ONCE, chunk by chunk, by the slab of the CELL-PACKED table (two 32-byte loads per point, one
// 64-byte block) the point's axis-0 cell lies in:
//   part_scatter_kernel  one CTA per chunk of 4096 points: reads the coordinates (coalesced), finds every point's list
//                        (slab), places the point inside its list's run of the chunk through shared memory, and writes
//                        the chunk back as runs -- list 0's points first, then list 1's, ... -- with TMA bulk stores;
//                        also the position it gave each point (u16) and, per list, the run {start, count}.
//   part_eval_kernel     list by list (all warps are in the same list at about the same time, so the slab being gathered
//                        from stays in L2), every warp walks the runs of its chunks: coalesced record loads, full warps,
//                        the same evaluation as the single pass, results and status bytes stored at the record's place.
//   part_merge_kernel    one CTA per chunk: the chunk's results come in through shared memory and leave in query order,
//                        whole sectors; errors are counted here, where the batch index is known.
// DRAM bytes per f64 point: 24 + 24 + 2 (scatter), 24 + 8 + 1 (evaluate), 2 + 8 + 1 + 8 (merge) = 102 against 136
// measured for the slab passes, and 3.5 L1 wavefronts against 7.5.  Every point is evaluated by the same device
// functions as everywhere else (fast_point on the packed table, general_point for the rest), so the bits are the same;
// which list a point lands in only decides what its gathers find in L2.
//
// InterpND with four or five axes (kernels templated on N, WIDE scatter): the packed table is 2^N times the plain one
// (3.7 GB for 32^5 f32) and one axis-0 plane of it is already beyond L2, so a slab is a range of the flattened
// (axis-0 cell, axis-1 cell) index and there are up to 128 lists; the cells come from the axis records (count_less).
//
// What the evaluate kernel is bound by (ncu, 256^3 f64, 2.6e8 random points, 3.26 ms): the L1 data pipe.  A 64-byte
// cell block per lane is two LDG.256 = two wavefronts per point whatever the cache level that answers; with the record
// and result streams 2.7 wavefronts per point, 75 % of the pipe's rate (FP64 pipe 31 %, issue slots 46 %, DRAM 38 %).
// At one wavefront per SM and clock, two per point bound ANY f64 3-D gather kernel on this part at
// 148 SMs x 1.92 GHz / 2 = 142 Gpoints/s.  Measured and dropped on the headline batch (25.5 ms as it stands):
//   * several groups per trip, weights of all, then all gathers, then the blends (97-128 registers, 2 CTAs):
//     29.4-32.1 ms -- fewer warps cost more than the overlapped gathers give;
//   * the slab of a list requested piece by piece (cp.async.bulk.prefetch.L2) by the first warps that enter the list:
//     27.8 vs 27.8 ms without (and both slower than without the code: 8 bytes of spill at 80 registers);
//   * four record groups in flight at 2 CTAs: 29.9 ms; slabs of 28 / 52 MB instead of 40: 30.9 / 25.7 ms;
//     segments of 2^29 / 2^30 points instead of 2^28: 27.3 / 27.1 ms;
//   * scatter/merge: chunks of 2048 points with 4 CTAs of 256 threads 27.0 ms, with the third column requested early
//     26.6 ms, chunks of 4096 with the third column requested early (spills) 27.7 ms.
#pragma once
#include "nint_async.cuh"
#include "nint_device.cuh"

namespace nint {

// 1: the scatter kernel requests the next chunk's first column before the third column's round trip instead of after the
// chunk's stores (measured on the headline batch: scatter 2.61 -> 2.51 ms per 2.6e8 points, whole call 25.5 -> 25.2 ms;
// f32 16.6 -> 16.05 ms)
#ifndef NINT_PART_EARLY_X
#define NINT_PART_EARLY_X 1
#endif
constexpr unsigned kPartChunkShift = 12;
constexpr unsigned kPartChunk = 1u << kPartChunkShift;  // points per chunk: positions and run bounds fit 16 bits
constexpr int kPartThreads = 512;
constexpr int kPartPer = (int)kPartChunk / kPartThreads;  // points per thread in the scatter and merge kernels
constexpr int kPartMaxLists = 32;       // narrow scatter: one slab axis, a lane of the scan warp per list
constexpr int kPartMaxListsWide = 128;  // wide scatter (N >= 4): slabs of the flattened (axis 0, axis 1) cell index
constexpr int kPartMaxDim = 5;

template <class T>
struct PartGeom {
    int lists;    // K: slabs of the packed table along axis 0 (wide: along the flattened index of axes 0 and 1)
    int uniform;  // axis 0 is uniform: list = (x - x0) * scale; otherwise the thresholds are searched
    int wrap;     // Extrapolate::Wrap: out-of-range coordinates are wrapped before they are classified
    T x0, x1;     // first and last node of axis 0
    T scale;      // lists per unit length (uniform)
    T thr[kPartMaxLists];  // thr[k], k >= 1: the axis-0 node where list k starts
    unsigned wide_per;     // wide: flattened (axis 0, axis 1) cells per list
    unsigned wide_cells1;  // wide: cells of axis 1
    T* rec[kPartMaxDim];   // the coordinates, every chunk rearranged into runs
    T* vals;               // results, at the record's place
    uint8_t* stat;         // status bytes, at the record's place
    unsigned short* pos;   // place of point i inside its chunk
    unsigned* runs;        // runs[k * run_stride + chunk] = start | count << 16
    unsigned* ticket;      // work counter of the evaluate kernel (zeroed by the scatter kernel)
    unsigned run_stride;
    unsigned nchunks;
};

__device__ __forceinline__ void bulk_store(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// The list of a point from its axis-0 coordinate.  Any answer is correct (every list is evaluated against the whole
// table); the right one keeps the gathers inside the slab that is L2-resident during that list.  NaN -> list 0.
template <class T>
__device__ __forceinline__ int part_list(const PartGeom<T>& G, const T* s_thr, T x) {
    if (G.wrap && !(G.x0 <= x && x <= G.x1)) x = wrap_coord(x, G.x0, G.x1);
    if (G.uniform) return max(0, min(to_int_sat((x - G.x0) * G.scale), G.lists - 1));
    int s = 0;
#pragma unroll
    for (int step = kPartMaxLists / 2; step >= 1; step >>= 1) {
        const int t = s + step;
        if (t < G.lists && x >= s_thr[t]) s = t;
    }
    return s;
}

// Wide form: the cell of one coordinate on axis d, for classification only (uniform axes by formula, the others from
// the node records through the bucket table, global memory / L1), and the list of a point from its first two.
template <class T>
__device__ __forceinline__ int part_axis_cell(const Params<T>& P, int d, T x, bool wrap) {
    if (wrap && !(P.g0[d] <= x && x <= P.gl[d])) x = wrap_coord(x, P.g0[d], P.gl[d]);
    if (!(x == x)) return 0;
    int k;
    if (P.flags[d] & kAxisUniform) k = to_int_sat((x - P.g0[d]) * P.inv_h[d]);
    else k = count_less(load_axis<T>(P, P.pack, d), x) - 1;
    return max(0, min(k, P.L[d] - 2));
}
template <class T>
__device__ __forceinline__ int part_list_wide(const Params<T>& P, const PartGeom<T>& G, T x, T y) {
    const unsigned flat = (unsigned)part_axis_cell(P, 0, x, G.wrap != 0) * G.wide_cells1 + (unsigned)part_axis_cell(P, 1, y, G.wrap != 0);
    return (int)min(flat / G.wide_per, (unsigned)G.lists - 1u);
}

// Measured alternative (1e9 random points, 256^3 f64): the three columns brought in by per-thread 8-byte cp.async into
// the slots they were read from and permuted in place after a barrier -- no coordinate registers, every request out
// before the first wait -- ran the kernel at 3.02 ms per 2^28 points against 2.73 ms for this form.
template <class T>
__global__ void __launch_bounds__(kPartThreads, 2) part_scatter_kernel(const __grid_constant__ Params<T> P,
                                                                       const __grid_constant__ PartGeom<T> G) {
    extern __shared__ __align__(128) unsigned char part_smem[];
    T* s_rec = reinterpret_cast<T*>(part_smem);  // [3][kPartChunk]
    __shared__ unsigned s_cnt[kPartMaxLists];
    __shared__ unsigned s_off[kPartMaxLists];
    __shared__ T s_thr[kPartMaxLists];
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;
    const unsigned t = threadIdx.x, lane = t & 31u, lt = (1u << lane) - 1u;
    const unsigned n = (unsigned)P.n;
    if (t < (unsigned)kPartMaxLists) s_thr[t] = G.thr[t];
    if (blockIdx.x == 0 && t == 0) *G.ticket = 0u;
    // resident CTAs walk the chunks (a CTA per chunk would be 65536 CTAs per segment, which cost a coherent batch --
    // where this launch only reads the verdict and returns -- 25 us each time)
    // the first two columns of a chunk are requested during the previous chunk: x -- which the classification waits
    // for -- as soon as its registers are free (before the third column's round trip), y while the bulk stores drain
    T x[kPartPer], y[kPartPer];
    auto request_x = [&](unsigned c) {
        const unsigned base = c << kPartChunkShift;
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            x[r] = (c < G.nchunks && i < n) ? ld_stream(P.coords[0] + (size_t)i * P.coord_stride) : quiet_nan<T>();
        }
    };
    auto request_y = [&](unsigned c) {
        const unsigned base = c << kPartChunkShift;
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            y[r] = (c < G.nchunks && i < n) ? ld_stream(P.coords[1] + (size_t)i * P.coord_stride) : T(0);
        }
    };
    request_x(blockIdx.x);
    request_y(blockIdx.x);
    for (unsigned c = blockIdx.x; c < G.nchunks; c += gridDim.x) {
    const unsigned base = c << kPartChunkShift;
    if (t < (unsigned)kPartMaxLists) s_cnt[t] = 0u;
    __syncthreads();
    // place inside the list: one shared-memory atomic per group of lanes with the same list
    unsigned pk[kPartPer];  // place inside the list | list << 16, then the place inside the chunk
#pragma unroll
    for (int r = 0; r < kPartPer; ++r) {
        const unsigned i = base + r * kPartThreads + t;
        const unsigned s = (i < n) ? (unsigned)part_list(G, s_thr, x[r]) : 255u;
        const unsigned m = __match_any_sync(0xffffffffu, s);
        const int leader = __ffs(m) - 1;
        unsigned b = 0u;
        if ((int)lane == leader && s != 255u) b = atomicAdd(&s_cnt[s], (unsigned)__popc(m));
        b = __shfl_sync(0xffffffffu, b, leader);
        pk[r] = (b + __popc(m & lt)) | (s << 16);
    }
    __syncthreads();
    if (t < 32u) {
        const unsigned cnt = ((int)t < G.lists) ? s_cnt[t] : 0u;
        unsigned incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += up;
        }
        s_off[t] = incl - cnt;
        NINT_ASSERT(incl <= kPartChunk);
        if ((int)t < G.lists) G.runs[(size_t)t * G.run_stride + c] = (incl - cnt) | (cnt << 16);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kPartPer; ++r) {
        const unsigned i = base + r * kPartThreads + t;
        if (i < n) {
            pk[r] = s_off[pk[r] >> 16] + (pk[r] & 0xffffu);
            NINT_ASSERT(pk[r] < kPartChunk);
            s_rec[pk[r]] = x[r];
            s_rec[kPartChunk + pk[r]] = y[r];
            G.pos[i] = (unsigned short)pk[r];
        }
    }
#if NINT_PART_EARLY_X
    request_x(c + gridDim.x);  // the first column of the next chunk: its registers are free, and the classification is what waits for it
#endif
    {
        T z[kPartPer];
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            z[r] = (i < n) ? ld_stream(P.coords[2] + (size_t)i * P.coord_stride) : T(0);
        }
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            if (i < n) s_rec[2 * kPartChunk + pk[r]] = z[r];
        }
    }
    // whole chunks leave (the lists are allocated in whole chunks; what lies beyond a ragged chunk's points is never read)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) bulk_store(G.rec[d] + base, s_rec + d * kPartChunk, kPartChunk * (unsigned)sizeof(T));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
#if !NINT_PART_EARLY_X
    request_x(c + gridDim.x);
#endif
    request_y(c + gridDim.x);
    if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();  // the staged chunk has been read: the next one may be written
    }
}

// The same for four and five axes: N columns, up to 128 lists by the flattened (axis 0, axis 1) cell index (a copy rather
// than a template argument of the kernel above: the 3-D f64 kernel sits at 64 registers and spilled in the merged form).
template <class T, int N>
__global__ void __launch_bounds__(kPartThreads, 2) part_scatter_wide_kernel(const __grid_constant__ Params<T> P,
                                                                       const __grid_constant__ PartGeom<T> G) {
    static_assert(N >= 4 && N <= kPartMaxDim, "wide scatter: four or five axes");
    constexpr bool WIDE = true;
    constexpr int kLists = WIDE ? kPartMaxListsWide : kPartMaxLists;
    extern __shared__ __align__(128) unsigned char part_smem[];
    T* s_rec = reinterpret_cast<T*>(part_smem);  // [N][kPartChunk]
    __shared__ unsigned s_cnt[kLists];
    __shared__ unsigned s_off[kLists];
    __shared__ T s_thr[kPartMaxLists];
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;
    const unsigned t = threadIdx.x, lane = t & 31u, lt = (1u << lane) - 1u;
    const unsigned n = (unsigned)P.n;
    if (t < (unsigned)kPartMaxLists) s_thr[t] = G.thr[t];
    if (blockIdx.x == 0 && t == 0) *G.ticket = 0u;
    // resident CTAs walk the chunks (a CTA per chunk would be 65536 CTAs per segment, which cost a coherent batch --
    // where this launch only reads the verdict and returns -- 25 us each time)
    // the first two columns of a chunk are requested while the previous chunk drains (its bulk stores read the staged
    // records, then a barrier): one of the two DRAM round trips of a chunk is off the critical path
    T x[kPartPer], y[kPartPer];
    auto request_xy = [&](unsigned c) {
        const unsigned base = c << kPartChunkShift;
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            x[r] = (c < G.nchunks && i < n) ? ld_stream(P.coords[0] + (size_t)i * P.coord_stride) : quiet_nan<T>();
        }
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            y[r] = (c < G.nchunks && i < n) ? ld_stream(P.coords[1] + (size_t)i * P.coord_stride) : T(0);
        }
    };
    request_xy(blockIdx.x);
    for (unsigned c = blockIdx.x; c < G.nchunks; c += gridDim.x) {
    const unsigned base = c << kPartChunkShift;
    if (t < (unsigned)kLists) s_cnt[t] = 0u;
    __syncthreads();
    // place inside the list: one shared-memory atomic per group of lanes with the same list
    unsigned pk[kPartPer];  // place inside the list | list << 16, then the place inside the chunk
#pragma unroll
    for (int r = 0; r < kPartPer; ++r) {
        const unsigned i = base + r * kPartThreads + t;
        unsigned s;
        if constexpr (WIDE) s = (i < n) ? (unsigned)part_list_wide(P, G, x[r], y[r]) : 255u;
        else s = (i < n) ? (unsigned)part_list(G, s_thr, x[r]) : 255u;
        const unsigned m = __match_any_sync(0xffffffffu, s);
        const int leader = __ffs(m) - 1;
        unsigned b = 0u;
        if ((int)lane == leader && s != 255u) b = atomicAdd(&s_cnt[s], (unsigned)__popc(m));
        b = __shfl_sync(0xffffffffu, b, leader);
        pk[r] = (b + __popc(m & lt)) | (s << 16);
    }
    __syncthreads();
    if (t < 32u) {
        if constexpr (!WIDE) {
            const unsigned cnt = ((int)t < G.lists) ? s_cnt[t] : 0u;
            unsigned incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += up;
            }
            s_off[t] = incl - cnt;
            NINT_ASSERT(incl <= kPartChunk);
            if ((int)t < G.lists) G.runs[(size_t)t * G.run_stride + c] = (incl - cnt) | (cnt << 16);
        } else {  // four lists per lane
            constexpr int kEach = kLists / 32;
            unsigned cnt[kEach], sum = 0u;
#pragma unroll
            for (int e = 0; e < kEach; ++e) {
                const int k = (int)t * kEach + e;
                cnt[e] = (k < G.lists) ? s_cnt[k] : 0u;
                sum += cnt[e];
            }
            unsigned incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += up;
            }
            NINT_ASSERT(incl <= kPartChunk);
            unsigned start = incl - sum;
#pragma unroll
            for (int e = 0; e < kEach; ++e) {
                const int k = (int)t * kEach + e;
                s_off[k] = start;
                if (k < G.lists) G.runs[(size_t)k * G.run_stride + c] = start | (cnt[e] << 16);
                start += cnt[e];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kPartPer; ++r) {
        const unsigned i = base + r * kPartThreads + t;
        if (i < n) {
            pk[r] = s_off[pk[r] >> 16] + (pk[r] & 0xffffu);
            NINT_ASSERT(pk[r] < kPartChunk);
            s_rec[pk[r]] = x[r];
            s_rec[kPartChunk + pk[r]] = y[r];
            G.pos[i] = (unsigned short)pk[r];
        }
    }
#pragma unroll 1
    for (int d = 2; d < N; ++d) {
        T z[kPartPer];
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            z[r] = (i < n) ? ld_stream(P.coords[d] + (size_t)i * P.coord_stride) : T(0);
        }
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            if (i < n) s_rec[d * kPartChunk + pk[r]] = z[r];
        }
    }
    // whole chunks leave (the lists are allocated in whole chunks; what lies beyond a ragged chunk's points is never read)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int d = 0; d < N; ++d) bulk_store(G.rec[d] + base, s_rec + d * kPartChunk, kPartChunk * (unsigned)sizeof(T));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    request_xy(c + gridDim.x);
    if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();  // the staged chunk has been read: the next one may be written
    }
}

// One point, the way interp_kernel evaluates it (Clamp up front, Wrap re-maps all axes of an out-of-range point, the
// general path for whatever the fast path declines); errors are reported through `status`, the caller counts them.
template <class T, int N, bool IS_ND, int PRE, bool UNI, bool ANA>
__device__ __forceinline__ T part_point(const Params<T>& P, const unsigned char* pack, unsigned long long pol, const Point<T, N>& q,
                                        unsigned& status) {
    T p[N];
    bool oob = false;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        p[d] = (PRE == kPreClamp) ? clamp_coord(q.v[d], P.g0[d], P.gl[d]) : q.v[d];
        if (PRE == kPreWrap) oob = oob || !(P.g0[d] <= q.v[d] && q.v[d] <= P.gl[d]);
    }
    if (PRE == kPreWrap) {
        if (oob) {
            const Point<T, N> w = wrap_point<T, N>(P, q);
#pragma unroll
            for (int d = 0; d < N; ++d) p[d] = w.v[d];
        }
    }
    T result;
    status = NINT_STATUS_OK;
    const bool ok = fast_point<T, N, IS_ND, NINT_LINEAR, UNI, PRE == kPreClamp, 2, ANA>(P, pack, pol, p, result);
    if (!ok) {
        const Outcome<T> o = general_point<T, N, IS_ND, NINT_LINEAR>(P, pack, q);
        result = o.value;
        status = o.status;
        if (status != NINT_STATUS_OK) result = quiet_nan<T>();
    }
    return result;
}

// f64: 80 registers, 3 CTAs; f32: 64 registers, 4 CTAs (measured: 17.4 -> 16.3 ms; f64 at 64 registers spills: 29.5 ms);
// four and five axes, f32: 3 CTAs (32^5 under Wrap, 2e8 points: 15.6 -> 13.2 ms against 2 CTAs), f64: what 2^N corners leave
#ifndef NINT_PART_CTAS
#define NINT_PART_CTAS (N == 3 ? (sizeof(T) == 4 ? 4 : 3) : (sizeof(T) == 4 ? 3 : (N == 4 ? 2 : 1)))
#endif
#ifndef NINT_PART_AHEAD
#define NINT_PART_AHEAD 2
#endif

constexpr unsigned kPartTicket = 4;  // chunks per ticket of the evaluate kernel

// Work is handed out in tickets of kPartTicket chunks, list-major, from one counter: at any moment all warps are in the
// same list or at its edge whatever their speed, so one slab (and the edge of the next) is what the gathers need in L2.
// A warp holds the ticket after its current one; the run entries of a ticket are one load (lane j: chunk j of the ticket).
//
// Measured on B200 (1e9 random points, 256^3 f64, whole route): records of one group ahead in registers 29.6 ms, of two
// groups 25.6 ms, of three 27.6 ms (80 registers, 3 CTAs); 4 CTAs at 64 registers 26.9 (one ahead) / 29.5 ms (two,
// spills).  A per-warp ring in shared memory filled by TMA bulk copies (three 256-byte requests per group, mbarrier per
// stage, 4 CTAs without spills) measured 27.5 ms with 4 stages, 28.1 with 3, 38.8 with 6.
template <class T, bool IS_ND, int PRE, bool UNI, bool ANA, int N = 3>
__global__ void __launch_bounds__(kBlock, NINT_PART_CTAS) part_eval_kernel(const __grid_constant__ Params<T> P,
                                                                           const __grid_constant__ PartGeom<T> G) {
    extern __shared__ __align__(16) unsigned char smem_pack[];
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.pack);
        uint4* dst = reinterpret_cast<uint4*>(smem_pack);
        for (unsigned i = threadIdx.x; i < P.pack_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const unsigned char* pack = smem_pack;
    const unsigned long long pol = make_table_policy(1.0f);  // the slab stays, the record and result streams go first
    const unsigned lane = threadIdx.x & 31u;
    const unsigned tpl = (G.nchunks + kPartTicket - 1) / kPartTicket;  // tickets per list
    const unsigned total = tpl * (unsigned)G.lists;

    auto draw = [&]() -> unsigned { return (lane == 0u) ? atomicAdd(G.ticket, 1u) : 0u; };  // the value stays in lane 0 until needed
    auto entries = [&](unsigned t) -> unsigned {
        if (t >= total) return 0u;
        const unsigned k = t / tpl;
        const unsigned c = (t - k * tpl) * kPartTicket + lane;
        return (lane < kPartTicket && c < G.nchunks) ? __ldg(G.runs + (size_t)k * G.run_stride + c) : 0u;
    };
    unsigned raw = draw();
    unsigned t_next = __shfl_sync(0xffffffffu, raw, 0);
    unsigned ent_next = entries(t_next);
    bool pending = false;  // t_next / ent_next still to be read from `raw`
    unsigned ent = 0u, cb = 0u, j = kPartTicket - 1u;
    unsigned r = 0, end = 0, cbase = 0;
    auto advance = [&]() -> bool {
        r += 32u;
        if (pending) {  // one group after the draw: its latency is behind us
            t_next = __shfl_sync(0xffffffffu, raw, 0);
            ent_next = entries(t_next);
            pending = false;
        }
        while (r >= end) {
            if (++j == kPartTicket) {
                if (pending) {
                    t_next = __shfl_sync(0xffffffffu, raw, 0);
                    ent_next = entries(t_next);
                    pending = false;
                }
                if (t_next >= total) return false;
                ent = ent_next;
                cb = (t_next % tpl) * kPartTicket;
                j = 0u;
                raw = draw();
                pending = true;
            }
            const unsigned e = __shfl_sync(0xffffffffu, ent, j);
            r = e & 0xffffu;
            end = r + (e >> 16);
            cbase = (cb + j) << kPartChunkShift;
        }
        return true;
    };
    auto fetch = [&](unsigned idx, Point<T, N>& q) {
#pragma unroll
        for (int d = 0; d < N; ++d) q.v[d] = ld_stream(G.rec[d] + idx);
    };
    // The f64 kernel sits at the edge of its 80 registers and is sensitive to the shape of this loop.  Measured on the
    // headline batch (25.8 ms in this form): an `inflight` counter and idx == ~0u for idle lanes instead of the flags
    // 32.2 ms; the runs of a ticket walked as one sequence so that only a ticket's last group is ragged (instead of one
    // group in ten) 32.3 ms; the slots taking turns in an unrolled trip of three groups (no register moves) 31.7 ms.
    // kAhead groups of records are in flight beyond the one being evaluated (a record load is a DRAM round trip, and a
    // group's evaluation -- two dependent L2 gathers -- is shorter than that under load)
    constexpr int kAhead = NINT_PART_AHEAD;
    Point<T, N> q[kAhead + 1];
    unsigned idx[kAhead + 1];
    bool live[kAhead + 1], has[kAhead + 1];
    bool done = false;
    auto load_slot = [&](int s) {
        has[s] = !done && advance();
        done = !has[s];
        live[s] = false;
        idx[s] = 0u;
        if (has[s]) {
            idx[s] = cbase + r + lane;
            live[s] = r + lane < end;
            if (live[s]) fetch(idx[s], q[s]);
        }
    };
#pragma unroll
    for (int s = 0; s < kAhead; ++s) load_slot(s);
    while (has[0]) {
        load_slot(kAhead);
        if (live[0]) {
            unsigned status;
            const T res = part_point<T, N, IS_ND, PRE, UNI, ANA>(P, pack, pol, q[0], status);
            st_stream(G.vals + idx[0], res);
            G.stat[idx[0]] = (uint8_t)status;
        }
#pragma unroll
        for (int s = 0; s < kAhead; ++s) {
            q[s] = q[s + 1];
            idx[s] = idx[s + 1];
            live[s] = live[s + 1];
            has[s] = has[s + 1];
        }
    }
}

// Shared memory of part_merge_kernel: two buffers of a chunk's results and status bytes
template <class T>
constexpr unsigned part_merge_smem() { return 2u * kPartChunk * ((unsigned)sizeof(T) + 1u); }

template <class T>
__global__ void __launch_bounds__(kPartThreads, 3) part_merge_kernel(const __grid_constant__ Params<T> P,
                                                                     const __grid_constant__ PartGeom<T> G) {
    extern __shared__ __align__(128) unsigned char part_smem[];
    constexpr unsigned kBuf = kPartChunk * ((unsigned)sizeof(T) + 1u);  // results, then status bytes
    __shared__ unsigned long long bar[2];
    __shared__ unsigned long long s_info[4];  // n_error, n_panic, first_error, first_panic
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;
    const unsigned t = threadIdx.x;
    const unsigned n = (unsigned)P.n;
    if (t < 4) s_info[t] = (t < 2) ? 0ull : ~0ull;
    if (t == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // resident CTAs walk the chunks; the next chunk's results are on their way into the other buffer meanwhile
    auto request = [&](unsigned c, unsigned b) {
        if (t == 0 && c < G.nchunks) {
            const unsigned long long pol = l2_policy_evict_first();
            const unsigned base = c << kPartChunkShift;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // an earlier chunk was read from this buffer
            mbar_expect_tx(&bar[b], kBuf);
            bulk_load(part_smem + b * kBuf, G.vals + base, kPartChunk * (unsigned)sizeof(T), &bar[b], pol);
            bulk_load(part_smem + b * kBuf + kPartChunk * sizeof(T), G.stat + base, kPartChunk, &bar[b], pol);
        }
    };
    request(blockIdx.x, 0u);
    unsigned it = 0u;
    for (unsigned c = blockIdx.x; c < G.nchunks; c += gridDim.x, ++it) {
        const unsigned b = it & 1u;
        const unsigned base = c << kPartChunkShift;
        request(c + gridDim.x, b ^ 1u);  // (that buffer's last readers passed the barrier at the end of the previous turn)
        unsigned short ps[kPartPer];
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            ps[r] = (i < n) ? __ldcs(G.pos + i) : (unsigned short)0;
        }
        const T* s_vals = reinterpret_cast<const T*>(part_smem + b * kBuf);
        const uint8_t* s_stat = part_smem + b * kBuf + kPartChunk * sizeof(T);
        mbar_wait(&bar[b], (it >> 1) & 1u);
#pragma unroll
        for (int r = 0; r < kPartPer; ++r) {
            const unsigned i = base + r * kPartThreads + t;
            if (i < n) {
                NINT_ASSERT(ps[r] < kPartChunk);
                const T v = s_vals[ps[r]];
                const unsigned status = s_stat[ps[r]];
                st_stream(P.out + i, v);
                if (P.status) P.status[i] = (uint8_t)status;
                if (status != NINT_STATUS_OK) {
                    const int slot = (status == NINT_STATUS_PANIC) ? 1 : 0;
                    atomicAdd(&s_info[slot], 1ull);
                    atomicMin(&s_info[2 + slot], (unsigned long long)i + P.index_base);
                }
            }
        }
        __syncthreads();  // every thread has its values: this buffer may be refilled
    }
    if (P.info) {
        __syncthreads();
        if (t == 0) {
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

struct PartLaunch {
    int is_nd;
    int all_uniform;
    int all_analytic;  // every axis is uniform with nodes exactly g0 + k * h: the evaluate kernel computes them
    int ndim;             // 3; 4 or 5 with the wide scatter (InterpND)
    int sm_count;
    unsigned smem_bytes;  // axis pack
    int grid_div, merge_div;  // tuning: > 0 = one CTA per that many chunks in the scatter / merge kernel instead of resident CTAs
    void* stream;
};

template <class T>
cudaError_t launch_part(const Params<T>& P, const PartGeom<T>& G, const PartLaunch& cfg);

}  // namespace nint
