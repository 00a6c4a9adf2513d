// nint_pipe.cuh — the single pass with its query stream behind a TMA pipeline.
//
// interp_kernel requests a point's coordinates one trip ahead with ordinary loads; on cell-ordered batches 36 % of
// its stall samples are warps waiting for those coordinates (profiles/r2_ncu_full_interp3d_linear_f64_256cubed_
// cell_sorted.txt: the first use of each prefetched coordinate), because one trip is shorter than a DRAM round trip
// and a deeper register prefetch costs the registers the corner gather needs.  Here one producer lane per CTA keeps
// kPipeStages chunks of kPipeChunk points in flight with cp.async.bulk (UBLKCP, one copy per coordinate column and
// chunk, mbarrier transaction counts); the consumer warps read their coordinates from shared memory, release the
// stage at once and evaluate the point exactly as interp_kernel does (same fast_point / general_point, same stores).
// Every CTA walks one contiguous range of chunks.
//
// Measured on B200 (1e9 points, 256^3 f64): cell-ordered 143 -> 146 Gpoints/s, Morton order 144 -> 151, queries grouped
// by 8^3-cell tiles 84 -> 89, f32 cell-ordered 178 -> 187.  (Once the table loads carried a 256-byte L2 prefetch size the
// two kernels ended within 3 % of each other -- the ring couples the CTA's 16 warps, and ncu shows the consumers
// spinning on chunks a slower warp still holds back -- and the default became: pipelined for f32, interp_kernel for f64.)  The gain is small because the wait moves: with the
// coordinates on time a warp still meets, on nearly every trip, a cell nobody has touched yet and waits a DRAM round
// trip for its corner block (32 warps x 32 points per ~1 us and SM = 145 Gpoints/s).  Tried against that and dropped:
// a 17th warp that guesses the cells of the chunks in flight and requests their lines -- with prefetch.global the
// load/store unit stalls on the translation misses of the 1 GiB table (143 -> 115 Gpoints/s), with cp.async.bulk
// copies into a scratch slot the requests are off the LSU but the warp's own round trip gates the stage ring (68); and the same guess made by
// warp 0 of the consumers two chunks ahead, non-blocking, one point in sixteen (146 -> 138, L1 or L2 alike).  Struct-of-arrays input with 16-byte aligned columns only; the
// last n % kPipeChunk points and every other case go to interp_kernel.
#pragma once
#include "nint_async.cuh"
#include "nint_device.cuh"

namespace nint {

template <class T, int N, bool IS_ND, int STRAT, int PRE, bool UNI>
__global__ void __launch_bounds__(kPipeThreads, 2) interp_pipe_kernel(const __grid_constant__ Params<T> P) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ __align__(8) unsigned long long full[kPipeStages], empty[kPipeStages];
    __shared__ unsigned long long s_info[4];  // n_error, n_panic, first_error, first_panic
    if (P.mode_flag != nullptr && *P.mode_flag != P.mode_want) return;  // the other query-order mode runs this batch
    // dynamic shared memory: kPipeStages x N columns of kPipeChunk coordinates, then the axis pack
    T* const stage0 = reinterpret_cast<T*>(smem_dyn);
    unsigned char* const smem_pack = smem_dyn + (size_t)kPipeStages * N * kPipeChunk * sizeof(T);
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.pack);
        uint4* dst = reinterpret_cast<uint4*>(smem_pack);
        for (unsigned i = threadIdx.x; i < P.pack_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    const unsigned char* pack = smem_pack;
    const unsigned tid = threadIdx.x;
    if (tid < 4) s_info[tid] = (tid < 2) ? 0ull : ~0ull;
    if (tid == 0) {
        for (int s = 0; s < kPipeStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kPipeChunk / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // this CTA's contiguous range of whole chunks
    const unsigned nchunks = (unsigned)(P.n / kPipeChunk);
    const unsigned per = (nchunks + gridDim.x - 1) / gridDim.x;
    const unsigned c_lo = min(nchunks, blockIdx.x * per), c_hi = min(nchunks, c_lo + per);

    if (tid >= (unsigned)kPipeChunk) {
        // ---------------------------------------------------------------- producer (one lane)
        if (tid == (unsigned)kPipeChunk) {
            const unsigned long long pol = l2_policy_evict_first();  // the query stream is touched once
            unsigned st = 0, phase = 0;
            for (unsigned c = c_lo; c < c_hi; ++c) {
                mbar_wait(&empty[st], phase ^ 1u);
                mbar_expect_tx(&full[st], (unsigned)(N * kPipeChunk * sizeof(T)));
#pragma unroll
                for (int d = 0; d < N; ++d)
                    bulk_load(stage0 + ((size_t)st * N + d) * kPipeChunk, P.coords[d] + (size_t)c * kPipeChunk,
                              (unsigned)(kPipeChunk * sizeof(T)), &full[st], pol);
                if (++st == kPipeStages) {
                    st = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- consumers: one point per thread and chunk
        const unsigned long long pol = make_table_policy(P.l2_frac);
        const unsigned lane = tid & 31u;
        unsigned st = 0, phase = 0;
        for (unsigned c = c_lo; c < c_hi; ++c) {
            mbar_wait(&full[st], phase);
            Point<T, N> q;
#pragma unroll
            for (int d = 0; d < N; ++d) q.v[d] = stage0[((size_t)st * N + d) * kPipeChunk + tid];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);  // the coordinates are in registers: the stage can be refilled
            if (++st == kPipeStages) {
                st = 0;
                phase ^= 1u;
            }
            const unsigned i = c * kPipeChunk + tid;
            // from here on: interp_kernel's process()
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
            const bool ok = fast_point<T, N, IS_ND, STRAT, UNI, PRE == kPreClamp>(P, pack, pol, p, result);
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
        }
    }
    if (P.info) {
        __syncthreads();
        if (tid == 0) {
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

}  // namespace nint
