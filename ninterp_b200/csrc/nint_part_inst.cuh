// nint_part_inst.cuh — instantiates the slab-list route (nint_part.cuh) for NINT_T.
#include <algorithm>
#include <atomic>

#include "nint_part.cuh"

namespace nint {

extern std::atomic<unsigned long long> g_launch_count;

namespace {

template <class T, bool IS_ND, int PRE, bool UNI, bool ANA = false, int N = 3>
cudaError_t launch_part_eval_variant(const Params<T>& P, const PartGeom<T>& G, const PartLaunch& cfg) {
    auto kern = part_eval_kernel<T, IS_ND, PRE, UNI, ANA, N>;
    const unsigned smem = cfg.smem_bytes;
    if (smem > 40u * 1024u) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    static thread_local unsigned cached_smem = ~0u;
    static thread_local int cached_bps = 0;
    int bps = cached_bps;
    if (cached_smem != smem || bps == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, kBlock, smem);
        if (e != cudaSuccess) return e;
        if (bps < 1) bps = 1;
        cached_bps = bps;
        cached_smem = smem;
    }
    const unsigned long long want = ((unsigned long long)(G.nchunks + kPartTicket - 1) / kPartTicket + (kBlock / 32) - 1) / (kBlock / 32);  // a warp per ticket of a list at most
    const unsigned long long cap = (unsigned long long)cfg.sm_count * (unsigned long long)bps;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (grid == 0) return cudaSuccess;
    kern<<<grid, kBlock, smem, (cudaStream_t)cfg.stream>>>(P, G);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <class T, bool IS_ND, int PRE>
cudaError_t launch_part_eval_uni(const Params<T>& P, const PartGeom<T>& G, const PartLaunch& cfg) {
    if constexpr (!IS_ND) {  // computed nodes: the hard-coded interpolator only (as in the single pass and the binned tiles)
        if (cfg.all_uniform && cfg.all_analytic) return launch_part_eval_variant<T, IS_ND, PRE, true, true>(P, G, cfg);
    }
    return cfg.all_uniform ? launch_part_eval_variant<T, IS_ND, PRE, true>(P, G, cfg) : launch_part_eval_variant<T, IS_ND, PRE, false>(P, G, cfg);
}
template <class T, bool IS_ND>
cudaError_t launch_part_eval(const Params<T>& P, const PartGeom<T>& G, const PartLaunch& cfg) {
    if (P.extrap == NINT_EXTRAPOLATE_CLAMP) return launch_part_eval_uni<T, IS_ND, kPreClamp>(P, G, cfg);
    if (P.extrap == NINT_EXTRAPOLATE_WRAP) return launch_part_eval_uni<T, IS_ND, kPreWrap>(P, G, cfg);
    return launch_part_eval_uni<T, IS_ND, kPrePlain>(P, G, cfg);
}

// InterpND with four or five axes
template <class T, int N, int PRE>
cudaError_t launch_part_eval_nd_uni(const Params<T>& P, const PartGeom<T>& G, const PartLaunch& cfg) {
    return cfg.all_uniform ? launch_part_eval_variant<T, true, PRE, true, false, N>(P, G, cfg)
                           : launch_part_eval_variant<T, true, PRE, false, false, N>(P, G, cfg);
}
template <class T, int N>
cudaError_t launch_part_eval_nd(const Params<T>& P, const PartGeom<T>& G, const PartLaunch& cfg) {
    if (P.extrap == NINT_EXTRAPOLATE_CLAMP) return launch_part_eval_nd_uni<T, N, kPreClamp>(P, G, cfg);
    if (P.extrap == NINT_EXTRAPOLATE_WRAP) return launch_part_eval_nd_uni<T, N, kPreWrap>(P, G, cfg);
    return launch_part_eval_nd_uni<T, N, kPrePlain>(P, G, cfg);
}

template <class T, int N>
cudaError_t launch_part_scatter(const Params<T>& P, const PartGeom<T>& G, const PartLaunch& cfg) {
    void (*kern)(const Params<T>, const PartGeom<T>);
    if constexpr (N == 3) kern = part_scatter_kernel<T>;
    else kern = part_scatter_wide_kernel<T, N>;
    const unsigned smem = (unsigned)N * kPartChunk * (unsigned)sizeof(T);
    // per device and function: set on every launch (a host thread may drive several GPUs)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned resident = smem > 110u * 1024u ? 1u : 2u;  // __launch_bounds__(kPartThreads, 2), 96/48 KB of shared memory each for 3-D
    unsigned cap = (unsigned)cfg.sm_count * resident;
    if (cfg.grid_div > 0) cap = std::max(cap, (G.nchunks + (unsigned)cfg.grid_div - 1) / (unsigned)cfg.grid_div);
    kern<<<G.nchunks < cap ? G.nchunks : cap, kPartThreads, smem, (cudaStream_t)cfg.stream>>>(P, G);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace

// scatter, evaluate (all lists in one launch), merge: one segment of a batch (P.n points, G sized for them)
template <>
cudaError_t launch_part<NINT_T>(const Params<NINT_T>& P, const PartGeom<NINT_T>& G, const PartLaunch& cfg) {
    using T = NINT_T;
    if (G.nchunks == 0) return cudaSuccess;
    cudaStream_t stream = (cudaStream_t)cfg.stream;
    cudaError_t e;
    if (cfg.ndim == 5) e = launch_part_scatter<T, 5>(P, G, cfg);
    else if (cfg.ndim == 4) e = launch_part_scatter<T, 4>(P, G, cfg);
    else e = launch_part_scatter<T, 3>(P, G, cfg);
    if (e != cudaSuccess) return e;
    if (cfg.ndim == 5) e = launch_part_eval_nd<T, 5>(P, G, cfg);
    else if (cfg.ndim == 4) e = launch_part_eval_nd<T, 4>(P, G, cfg);
    else e = cfg.is_nd ? launch_part_eval<T, true>(P, G, cfg) : launch_part_eval<T, false>(P, G, cfg);
    if (e != cudaSuccess) return e;
    {
        auto kern = part_merge_kernel<T>;
        const unsigned smem = part_merge_smem<T>();
        cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ea != cudaSuccess) return ea;
        unsigned cap = (unsigned)cfg.sm_count * 3u;  // __launch_bounds__(kPartThreads, 3)
        if (cfg.merge_div > 0) cap = std::max(cap, (G.nchunks + (unsigned)cfg.merge_div - 1) / (unsigned)cfg.merge_div);
        kern<<<G.nchunks < cap ? G.nchunks : cap, kPartThreads, smem, stream>>>(P, G);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace nint
