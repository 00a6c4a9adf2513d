// nint_tile.cuh — the binned route for incoherent queries on 3-D to 5-D tables beyond L2 (Linear).
//
// A random cell of a table that does not fit L2 costs 2^(N-1) DRAM sectors or more, and even with an L2-resident slab
// the gathers are bound by L2 sector bandwidth.  This route takes the gathers out of the memory
// system altogether: queries are bucketed by grid tile and every tile's value sub-block is staged once in shared
// memory by TMA (cp.async.bulk.tensor, UTMALDG), where the 2^N corner loads are LDS.
//
//   tile_bin_kernel   one point per thread: front-end, cell guess + verification and the weights t_d exactly as the
//                     single pass computes them (fast_weights == first half of fast_point); the point's record
//                     {t_0 .. t_{N-1}, index, cell inside the tile} is appended to the list of its (superchunk, tile)
//                     with one global atomic.  Points that need the general path are finished here.
//   tile_eval_kernel  persistent CTAs walk the (superchunk, tile) lists superchunk-major: TMA brings the tile's
//                     prod(B_d + 1) values into shared memory (ring of two, mbarrier), the records arrive by
//                     cp.async.bulk, are blended from shared memory with the reference's operation order
//                     (three/strategies.rs:37-49, n/strategies.rs:86-104) and the results scattered to out[index].
//                     A superchunk is 2^22 (f64) / 2^23 (f32) consecutive points, so the scatter stays inside a 32 MB
//                     window of `out` that L2 merges into full lines before it reaches DRAM.
//
// The arithmetic is the single pass's: t_d is computed once (bin) by the same inlined functions and the blend uses
// the same expression tree, so results are bit-identical whichever route a batch takes.
#pragma once
#include <cuda.h>

#include "nint_async.cuh"
#include "nint_device.cuh"

namespace nint {

constexpr unsigned kCursorStride = 8;  // u32 slots between list cursors: 32 bytes, so neighbours hit different L2 slices
constexpr int kTileMaxDim = 5;         // TMA tensor maps have at most five dimensions
constexpr int kLocBits = 6;            // bits per axis of a record's cell-inside-the-tile field (tiles of at most 64 cells per axis)

// A record: N weights in T, padding, then {point index, cell inside the tile} in the last 8 bytes.
template <class T, int N>
__host__ __device__ constexpr int tile_rec_bytes() {
    return (N * (int)sizeof(T) + 8 <= 32) ? 32 : ((N * (int)sizeof(T) + 8 <= 48) ? 48 : 64);
}

// Geometry of the binned route, shared by both kernels.
struct TileGeom {
    int ndim;
    int sh[kTileMaxDim];   // log2(cells per tile) along each axis
    int nt[kTileMaxDim];   // tiles per axis
    int ntiles;
    int box[kTileMaxDim];  // staged elements per axis: cells per tile + 1 (innermost padded to a 16-byte multiple)
    unsigned tile_bytes;   // bytes TMA delivers per tile (box volume x sizeof(T))
    unsigned buf_bytes;    // tile_bytes rounded up to 128
    unsigned rec_bytes;    // tile_rec_bytes<T, N>()
    unsigned cap;          // records per (superchunk, tile) list: rounds << page_shift
    unsigned page_shift;   // log2(records per page).  A list is stored as pages interleaved over the tiles of its
    unsigned rounds;       // superchunk: page r of every tile, then page r + 1, ...  Lists grow at the same pace when the
                           // queries are spread evenly, so the appends of a moment land in a few MB instead of one spot
                           // per list over hundreds of MB (one TLB miss per record otherwise)
    unsigned sc_shift;     // log2(points per superchunk)
    unsigned n_sc;         // superchunks in this launch
    unsigned flags;        // tuning (NINT_TILE_FLAGS): 2 = no evict-first hint on the record and tile loads, 4 = no
                           // evict-last hint on the result stores, 8 = no result stores at all (measurement only),
                           // 16 = no prefetch of the result window
    unsigned long long records_bytes;  // size of `records` (bounds-checked builds)
    unsigned* cursor;      // [n_sc * ntiles] list lengths, kCursorStride apart
    unsigned char* records;  // [n_sc][rounds][ntiles][1 << page_shift] records
};

// Byte offset of record `slot` of list (sc, tile).
__device__ __forceinline__ size_t record_offset(const TileGeom& G, unsigned sc, unsigned tile, unsigned slot) {
    const size_t page = ((size_t)sc * G.rounds + (slot >> G.page_shift)) * (size_t)G.ntiles + tile;
    return ((page << G.page_shift) + (slot & ((1u << G.page_shift) - 1u))) * (size_t)G.rec_bytes;
}

// First half of fast_point() for Linear: cell guess, verification and weights of every axis.  Same functions, same
// order, so a point is accepted here exactly when the single pass accepts it, with the same k and t.
template <class T, int N, bool IS_ND, bool UNI, bool FIRST>
__device__ __forceinline__ bool fast_weights(const Params<T>& P, const unsigned char* pack, const T (&p)[N], int (&k)[N],
                                             T (&t)[N]) {
    constexpr bool kCollapse = IS_ND || (N == 1);
    bool ok = true;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const unsigned* guess = UNI ? nullptr : reinterpret_cast<const unsigned*>(pack + P.guess_off[d]);
        ok = guess_cell<T, UNI>(load_axis<T>(P, pack, d), guess, p[d], k[d]) && ok;
    }
    if (!ok) return false;
#pragma unroll
    for (int d = 0; d < N; ++d) {
        T a, w, r;
        ok = verify_cell_linear<T, kCollapse, FIRST>(load_axis<T>(P, pack, d), p[d], k[d], a, w, r) && ok;
        t[d] = div_markstein<T>(a, w, r);
    }
    if (!UNI && !ok) {
        ok = true;
#pragma unroll
        for (int d = 0; d < N; ++d) {
            T a, w, r;
            ok = verify_linear_retry<T, kCollapse, FIRST, UNI>(load_axis<T>(P, pack, d), p[d], k[d], a, w, r) && ok;
            t[d] = div_markstein<T>(a, w, r);
        }
    }
    return ok;
}

// three/strategies.rs:37-49 (and two/, n/): reduce axis 0 first, v_l * (1 - t) + v_u * t, every op rounded separately.
template <class T, int N>
__device__ __forceinline__ T blend_cube(T (&cube)[1 << N], const T (&t)[N]) {
    if constexpr (N == 4 || N == 5) {  // stage widths as constants: the loop form below left the corners in local memory
        blend_stages<T, N>(cube, t);
        return cube[0];
    }
#pragma unroll
    for (int d = 0; d < N; ++d) {
        const int half = 1 << (N - 1 - d);
        const T td = t[d];
        const T sd = T(1) - td;
#pragma unroll
        for (int j = 0; j < half; ++j) cube[j] = cube[j] * sd + cube[half + j] * td;
    }
    return cube[0];
}

// Record I/O.  Stored with default L2 priority: the 2 KB page a list is filling stays in L2 until its lines are
// complete (evict-first sectors leave one by one and turn into partial-line DRAM writes).
template <class T, int N>
__device__ __forceinline__ void st_record(unsigned char* dst, const T (&t)[N], unsigned idx, unsigned loc) {
    constexpr int W = tile_rec_bytes<T, N>() / 8;
    unsigned long long w[W];
#pragma unroll
    for (int i = 0; i < W; ++i) w[i] = 0ull;
    if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int d = 0; d < N; ++d) w[d] = (unsigned long long)__double_as_longlong((double)t[d]);
    } else {
#pragma unroll
        for (int d = 0; d < N; ++d) w[d / 2] |= (unsigned long long)__float_as_uint((float)t[d]) << (32 * (d & 1));
    }
    w[W - 1] = (unsigned long long)idx | ((unsigned long long)loc << 32);
    if constexpr (W % 4 == 0) {  // 32- and 64-byte records are 32-byte aligned
#pragma unroll
        for (int i = 0; i < W; i += 4)
            asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(dst + 8 * i), "l"(w[i]), "l"(w[i + 1]), "l"(w[i + 2]), "l"(w[i + 3]) : "memory");
    } else {  // 48-byte records: 16-byte aligned
#pragma unroll
        for (int i = 0; i < W; i += 2)
            asm volatile("st.global.v2.b64 [%0], {%1,%2};" ::"l"(dst + 8 * i), "l"(w[i]), "l"(w[i + 1]) : "memory");
    }
}
// from a record stage in shared memory
template <class T, int N>
__device__ __forceinline__ void ld_record_shared(const unsigned char* src, T (&t)[N], unsigned& idx, unsigned& loc) {
    constexpr int W = tile_rec_bytes<T, N>() / 8;
    unsigned long long w[W];
#pragma unroll
    for (int i = 0; i < W; i += 2) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + 8 * i);
        w[i] = v.x;
        w[i + 1] = v.y;
    }
    if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int d = 0; d < N; ++d) t[d] = (T)__longlong_as_double((long long)w[d]);
    } else {
#pragma unroll
        for (int d = 0; d < N; ++d) t[d] = (T)__uint_as_float((unsigned)(w[d / 2] >> (32 * (d & 1))));
    }
    idx = (unsigned)w[W - 1];
    loc = (unsigned)(w[W - 1] >> 32);
}

// ---------------------------------------------------------------------------------------------
// Bin: the point loop of interp_kernel with the gather replaced by an append to the point's list.
// ---------------------------------------------------------------------------------------------
#ifndef NINT_BIN_ANA_CTAS
#define NINT_BIN_ANA_CTAS 4  // resident CTAs the analytic variant is compiled for; 5 (48 registers) and 6 (40) spill and
                            // measured slower: 16.3 -> 19.1 and 19.5 ms per 1e9 points
#endif
#ifndef NINT_BIN_ND_CTAS
#define NINT_BIN_ND_CTAS 4  // four and five axes: the weights need no corner registers, the rare overflow gather may spill
#endif
constexpr int bin_ctas(int n, bool ana) { return n == 3 ? (ana ? NINT_BIN_ANA_CTAS : 4) : NINT_BIN_ND_CTAS; }

template <class T, int N, bool IS_ND, int PRE, bool UNI, bool ANA = false>
__global__ void __launch_bounds__(kBlock, bin_ctas(N, ANA)) tile_bin_kernel(const __grid_constant__ Params<T> P,
                                                                            const __grid_constant__ TileGeom G) {
    extern __shared__ __align__(16) unsigned char smem_pack[];
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.pack);
        uint4* dst = reinterpret_cast<uint4*>(smem_pack);
        for (unsigned i = threadIdx.x; i < P.pack_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const unsigned char* pack = smem_pack;
    const unsigned long long pol = make_table_policy(P.l2_frac);
    __shared__ unsigned long long s_info[4];  // n_error, n_panic, first_error, first_panic
    if (threadIdx.x < 4) s_info[threadIdx.x] = (threadIdx.x < 2) ? 0ull : ~0ull;
    __syncthreads();

    const unsigned n = (unsigned)P.n;
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    Point<T, N> q;
    if (i < n) {
#pragma unroll
        for (int d = 0; d < N; ++d) q.v[d] = ld_stream(P.coords[d] + (size_t)i * P.coord_stride);
    }
    auto process = [&](unsigned i, const Point<T, N>& q) {
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
        int k[N];
        T t[N];
        const bool fast = ANA ? fast_weights_analytic<T, N, IS_ND, PRE == kPreClamp>(P, p, k, t)
                              : fast_weights<T, N, IS_ND, UNI, PRE == kPreClamp>(P, pack, p, k, t);
        if (fast) {
            unsigned tile = 0, loc = 0;
#pragma unroll
            for (int d = 0; d < N; ++d) {
                tile = tile * (unsigned)G.nt[d] + (unsigned)(k[d] >> G.sh[d]);
                loc |= (unsigned)(k[d] & ((1 << G.sh[d]) - 1)) << (kLocBits * d);
            }
            const unsigned sc = i >> G.sc_shift;
            const unsigned list = sc * (unsigned)G.ntiles + tile;
            NINT_ASSERT(sc < G.n_sc && tile < (unsigned)G.ntiles);
            const unsigned slot = atomicAdd(G.cursor + (size_t)list * kCursorStride, 1u);
            if (slot < G.cap) {
                NINT_ASSERT(record_offset(G, sc, tile, slot) + G.rec_bytes <= G.records_bytes);
                st_record<T, N>(G.records + record_offset(G, sc, tile, slot), t, i, loc);
            } else {
                // list full (queries far from uniform): finish the point here from the plain table
                long long base = 0;
                long long step[N];
#pragma unroll
                for (int d = 0; d < N; ++d) {
                    base += (long long)k[d] * P.vstride[d];
                    step[d] = P.vstride[d];
                }
                T cube[1 << N];
                gather_rows<T, N>(P.values + base, step, pol, false, cube);
                st_stream(P.out + i, blend_cube<T, N>(cube, t));
            }
            if (P.status) P.status[i] = (uint8_t)NINT_STATUS_OK;
            return;
        }
        const Outcome<T> o = general_point<T, N, IS_ND, NINT_LINEAR>(P, pack, q);
        T result = o.value;
        if (o.status != NINT_STATUS_OK) {
            result = quiet_nan<T>();
            const int slot = (o.status == NINT_STATUS_PANIC) ? 1 : 0;
            atomicAdd(&s_info[slot], 1ull);
            atomicMin(&s_info[2 + slot], (unsigned long long)i + P.index_base);
        }
        st_stream(P.out + i, result);
        if (P.status) P.status[i] = (uint8_t)o.status;
    };
    Point<T, N> q2;
    while (i < n) {
        const unsigned in = i + stride;
        const bool more = in < n;
        if (more) {
#pragma unroll
            for (int d = 0; d < N; ++d) q2.v[d] = ld_stream(P.coords[d] + (size_t)in * P.coord_stride);
        }
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
// Evaluate: TMA-staged tiles, records blended from shared memory.
// ---------------------------------------------------------------------------------------------
// One tile: the box of the table at element coordinates c (c[N-1] innermost) -> shared memory.
template <int N>
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, unsigned long long* bar, const int (&c)[N],
                                              unsigned long long pol) {
    if constexpr (N == 3) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
                smem_u32(dst)),
            "l"(map), "r"(c[2]), "r"(c[1]), "r"(c[0]), "r"(smem_u32(bar)), "l"(pol)
            : "memory");
    } else if constexpr (N == 4) {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;" ::"r"(
                smem_u32(dst)),
            "l"(map), "r"(c[3]), "r"(c[2]), "r"(c[1]), "r"(c[0]), "r"(smem_u32(bar)), "l"(pol)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;" ::"r"(
                smem_u32(dst)),
            "l"(map), "r"(c[4]), "r"(c[3]), "r"(c[2]), "r"(c[1]), "r"(c[0]), "r"(smem_u32(bar)), "l"(pol)
            : "memory");
    }
}

template <class T>
struct TileEval {
    T* out;
    unsigned n;  // points of this launch (the last superchunk's window is shorter)
    const int* mode_flag;
    int mode_want;
};

// One CTA per SM: warp 0 is the producer (one lane drives the TMA engine), kEvalConsumers threads evaluate.
//   tile ring    2 buffers   producer: UTMALDG of the item's tile        consumers: corner loads (LDS)
//   record ring  stages of kEvalChunk records, UBLKCP page by page; one record per consumer thread
// Every buffer has a `full` (transaction-count) and an `empty` (one arrival per consumer warp) mbarrier, so record
// and tile traffic of the items ahead is in flight while the current records are blended: nothing waits on DRAM
// latency.  Both roles walk the same item sequence (the lists are complete before this kernel starts).
constexpr int kEvalConsumers = 512;
constexpr int kEvalThreads = kEvalConsumers + 32;
constexpr int kEvalChunk = kEvalConsumers;  // records per stage
__host__ __device__ constexpr int eval_stages(int rec_bytes) { return rec_bytes == 32 ? 4 : 2; }  // 64 KB (48 KB) of record stages

template <class T, int N>
__global__ void __launch_bounds__(kEvalThreads, 1) tile_eval_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                     const __grid_constant__ TileEval<T> E,
                                                                     const __grid_constant__ TileGeom G) {
    constexpr int kRec = tile_rec_bytes<T, N>();
    constexpr int kStages = eval_stages(kRec);
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long tile_full[2], tile_empty[2], rec_full[kStages], rec_empty[kStages];
    if (E.mode_flag != nullptr && *E.mode_flag != E.mode_want) return;
    const unsigned tid = threadIdx.x;
    unsigned char* const tiles = smem_dyn;
    unsigned char* const stages = smem_dyn + 2u * G.buf_bytes;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tile_full[b], 1);
            mbar_init(&tile_empty[b], kEvalConsumers / 32);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&rec_full[s], 1);
            mbar_init(&rec_empty[s], kEvalConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // the only CTA-wide barrier: the roles part here

    const unsigned nitems = G.n_sc * (unsigned)G.ntiles;
    // this CTA's items: blockIdx.x, + gridDim.x, ...: superchunk-major, so that at any time all CTAs scatter into
    // the same window of `out`
    auto next_item = [&](unsigned j) {
        while (j < nitems && G.cursor[(size_t)j * kCursorStride] == 0u) j += gridDim.x;
        return j;
    };

    if (tid < 32) {
        if (tid != 0) return;
        // ---------------------------------------------------------------- producer
        // records and tiles are used once per superchunk: evict-first, so that the window of `out` the superchunk
        // scatters into stays in L2 until its sectors are complete
        const unsigned long long stream_pol = (G.flags & 2u) ? make_table_policy(0.0f) : l2_policy_evict_first();
        unsigned tb = 0, tb_phase = 0, st = 0, st_phase = 0;
        // the tile of the NEXT item is requested before the records of the current one: the record stages only free
        // up as the consumers work through the item, and a tile requested after them would land too late
        auto request_tile = [&](unsigned item) {
            unsigned tile = item % (unsigned)G.ntiles;
            int c[N];
#pragma unroll
            for (int d = N - 1; d >= 0; --d) {
                c[d] = (int)(tile % (unsigned)G.nt[d]) << G.sh[d];
                tile /= (unsigned)G.nt[d];
            }
            mbar_wait(&tile_empty[tb], tb_phase ^ 1u);
            mbar_expect_tx(&tile_full[tb], G.tile_bytes);
            tma_load_tile<N>(tiles + (size_t)tb * G.buf_bytes, &tmap, &tile_full[tb], c, stream_pol);
            tb ^= 1u;
            tb_phase ^= (tb == 0u);
        };
        unsigned item = next_item(blockIdx.x);
        if (item < nitems) request_tile(item);
        while (item < nitems) {
            const unsigned nxt = next_item(item + gridDim.x);
            if (nxt < nitems) request_tile(nxt);
            const unsigned sc = item / (unsigned)G.ntiles;
            const unsigned tile = item - sc * (unsigned)G.ntiles;
            const unsigned count = min(G.cursor[(size_t)item * kCursorStride], G.cap);
            for (unsigned base = 0; base < count; base += kEvalChunk) {
                const unsigned m = min((unsigned)kEvalChunk, count - base);
                mbar_wait(&rec_empty[st], st_phase ^ 1u);
                mbar_expect_tx(&rec_full[st], m * kRec);
                unsigned char* dst = stages + (size_t)st * (kEvalChunk * kRec);
                // page by page: a piece never crosses a page boundary
                for (unsigned slot = base; slot < base + m;) {
                    const unsigned in_page = slot & ((1u << G.page_shift) - 1u);
                    const unsigned len = min((1u << G.page_shift) - in_page, base + m - slot);
                    NINT_ASSERT(record_offset(G, sc, tile, slot) + (size_t)len * kRec <= G.records_bytes);
                    bulk_load(dst + (size_t)(slot - base) * kRec, G.records + record_offset(G, sc, tile, slot), len * kRec,
                              &rec_full[st], stream_pol);
                    slot += len;
                }
                if (++st == kStages) {
                    st = 0;
                    st_phase ^= 1u;
                }
            }
            item = nxt;
        }
        return;
    }
    // -------------------------------------------------------------------- consumers
    const unsigned ct = tid - 32u;
    const unsigned lane = tid & 31u;
    int stride[N];  // element strides of the staged tile
    stride[N - 1] = 1;
#pragma unroll
    for (int d = N - 2; d >= 0; --d) stride[d] = stride[d + 1] * G.box[d + 1];
    const unsigned long long keep_pol = (G.flags & 4u) ? make_table_policy(0.0f) : l2_policy_evict_last();
    unsigned tb = 0, tb_phase = 0, st = 0, st_phase = 0;
    unsigned window_sc = ~0u;
    for (unsigned item = next_item(blockIdx.x); item < nitems; item = next_item(item + gridDim.x)) {
        const unsigned count = min(G.cursor[(size_t)item * kCursorStride], G.cap);
        // On entering a superchunk, bring this CTA's share of its window of `out` into L2 (evict-last): an 8-byte
        // store to a sector L2 does not hold is a write miss with a 32-byte fill from DRAM, one per point; after the
        // coalesced prefetch the scattered stores of the whole superchunk are hits that merge into full lines.
        const unsigned sc_now = item / (unsigned)G.ntiles;
        if (sc_now != window_sc && !(G.flags & 16u)) {
            window_sc = sc_now;
            const size_t first = (size_t)sc_now << G.sc_shift;
            const size_t last = min((size_t)E.n, first + ((size_t)1 << G.sc_shift));
            const char* base = reinterpret_cast<const char*>(E.out + first);
            const size_t lines = ((last - first) * sizeof(T) + 127) / 128;
            const size_t per_cta = (lines + gridDim.x - 1) / gridDim.x;
            const size_t lo = per_cta * blockIdx.x, hi = min(lines, lo + per_cta);
            for (size_t l = lo + ct; l < hi; l += kEvalConsumers)
                asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + l * 128));
        }
        mbar_wait(&tile_full[tb], tb_phase);
        const T* tile = reinterpret_cast<const T*>(tiles + (size_t)tb * G.buf_bytes);
        for (unsigned base = 0; base < count; base += kEvalChunk) {
            const unsigned m = min((unsigned)kEvalChunk, count - base);
            mbar_wait(&rec_full[st], st_phase);
            if (ct < m) {
                T t[N];
                unsigned idx, loc;
                ld_record_shared<T, N>(stages + (size_t)st * (kEvalChunk * kRec) + (size_t)ct * kRec, t, idx, loc);
                int off = 0, span = 0;
#pragma unroll
                for (int d = 0; d < N; ++d) {
                    off += (int)((loc >> (kLocBits * d)) & ((1u << kLocBits) - 1u)) * stride[d];
                    span += stride[d];
                }
                NINT_ASSERT(idx < E.n && (unsigned)(off + span) * sizeof(T) < G.tile_bytes);
                const T* c = tile + off;
                T cube[1 << N];
#pragma unroll
                for (int r = 0; r < (1 << N); ++r) {  // bit (N-1-d) of r selects the upper node of axis d
                    int o = 0;
#pragma unroll
                    for (int d = 0; d < N; ++d)
                        if ((r >> (N - 1 - d)) & 1) o += stride[d];
                    cube[r] = c[o];
                }
                const T value = blend_cube<T, N>(cube, t);
                if (!(G.flags & 8u) || idx == 0xffffffffu) st_hint(E.out + idx, value, keep_pol);  // 8: measurement without the scatter
                (void)span;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&rec_empty[st]);
            if (++st == kStages) {
                st = 0;
                st_phase ^= 1u;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&tile_empty[tb]);
        tb ^= 1u;
        tb_phase ^= (tb == 0u);
    }
}

// ---------------------------------------------------------------------------------------------
// Order probe: footprint of windows of consecutive queries.
//
// kProbeWindows windows of kProbeWindowPts consecutive points, spread over the batch, one CTA each.  A window's
// footprint is the number of distinct coarse tiles (8 cells per axis) its points fall into; the batch is coherent
// when the mean footprint is at most `max_distinct` -- the host derives that bound from the number of coarse tiles
// of the gathered table that fit a share of L2, so cell-ordered, tile-ordered, Morton-ordered and axis-sorted
// batches all take the single pass, and uniformly random ones do not.
// ---------------------------------------------------------------------------------------------
constexpr int kProbeWindows = 8;
constexpr int kProbeWindowPts = 4096;
constexpr int kProbeThreads = 1024;
constexpr int kProbeBits = 1 << 16;

struct ProbeSlot {  // one per in-flight batch call (ring in the handle)
    int mode;
    unsigned sum;
    unsigned ticket;
    unsigned pad;
};

template <class T>
__global__ void __launch_bounds__(kProbeThreads) order_probe2_kernel(const __grid_constant__ Params<T> P, int ndim, int shift,
                                                                      unsigned max_distinct_total, ProbeSlot* slot) {
    __shared__ unsigned s_bits[kProbeBits / 32];
    __shared__ unsigned s_count;
    for (unsigned w = threadIdx.x; w < kProbeBits / 32; w += blockDim.x) s_bits[w] = 0u;
    if (threadIdx.x == 0) s_count = 0u;
    __syncthreads();
    const unsigned n = (unsigned)P.n;
    const unsigned span = n > (unsigned)kProbeWindowPts ? n - (unsigned)kProbeWindowPts : 0u;
    const unsigned start = (gridDim.x > 1) ? (unsigned)(((unsigned long long)span * blockIdx.x) / (gridDim.x - 1)) : 0u;
    unsigned long long total_tiles = 1;
    for (int d = 0; d < ndim; ++d) total_tiles *= (unsigned long long)(((max(P.L[d], 2) - 2) >> shift) + 1);
    for (unsigned j = threadIdx.x; j < (unsigned)kProbeWindowPts; j += blockDim.x) {
        const unsigned i = start + j;
        if (i >= n) break;
        unsigned long long lin = 0;
        for (int d = 0; d < ndim; ++d) {
            const AxisRef<T> a = load_axis<T>(P, P.pack, d);
            const unsigned* guess = reinterpret_cast<const unsigned*>(P.pack + P.guess_off[d]);
            const T x = __ldg(P.coords[d] + (size_t)i * P.coord_stride);
            int k = 0;
            if (a.L >= 2) guess_cell<T, false>(a, guess, x, k);
            lin = lin * (unsigned long long)(((max(a.L, 2) - 2) >> shift) + 1) + (unsigned long long)(k >> shift);
        }
        const unsigned bit = (total_tiles <= (unsigned long long)kProbeBits)
                                 ? (unsigned)lin
                                 : (unsigned)(((lin * 0x9E3779B97F4A7C15ull) >> 40) & (kProbeBits - 1));
        atomicOr(&s_bits[bit >> 5], 1u << (bit & 31u));
    }
    __syncthreads();
    unsigned c = 0;
    for (unsigned w = threadIdx.x; w < kProbeBits / 32; w += blockDim.x) c += __popc(s_bits[w]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31u) == 0u) atomicAdd(&s_count, c);
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&slot->sum, s_count);
        __threadfence();
        const unsigned ticket = atomicAdd(&slot->ticket, 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            const unsigned sum = atomicAdd(&slot->sum, 0u);
            slot->mode = (sum <= max_distinct_total) ? kModeCoherent : kModeRandom;
            slot->sum = 0u;
            slot->ticket = 0u;
        }
    }
}

// Launch interface (nint_tile_f32.cu, nint_tile_f64.cu).
struct TileLaunch {
    int ndim, is_nd, all_uniform, all_analytic, sm_count;
    unsigned smem_bytes;  // axis pack staged by the bin kernel
    void* stream;
};
template <class T>
cudaError_t launch_tile_bin(const Params<T>& P, const TileGeom& G, const TileLaunch& cfg);
template <class T>
cudaError_t launch_tile_eval(const CUtensorMap& tmap, const TileEval<T>& E, const TileGeom& G, const TileLaunch& cfg);
template <class T>
cudaError_t launch_order_probe2(const Params<T>& P, int ndim, int shift, unsigned max_distinct_total, ProbeSlot* slot, void* stream);

}  // namespace nint
