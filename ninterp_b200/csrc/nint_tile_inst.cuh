// nint_tile_inst.cuh — instantiates the binned route (nint_tile.cuh) and the order probe for NINT_T.
#include <atomic>

#include "nint_tile.cuh"

namespace nint {

extern std::atomic<unsigned long long> g_launch_count;

namespace {

template <class T, int N, bool IS_ND, int PRE, bool UNI, bool ANA = false>
cudaError_t launch_bin_variant(const Params<T>& P, const TileGeom& G, const TileLaunch& cfg) {
    auto kern = tile_bin_kernel<T, N, IS_ND, PRE, UNI, ANA>;
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
    const unsigned long long want = (P.n + kBlock - 1) / kBlock;
    const unsigned long long cap = (unsigned long long)cfg.sm_count * (unsigned long long)bps;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (grid == 0) return cudaSuccess;
    kern<<<grid, kBlock, smem, (cudaStream_t)cfg.stream>>>(P, G);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <class T, int N, bool IS_ND, int PRE>
cudaError_t launch_bin_uni(const Params<T>& P, const TileGeom& G, const TileLaunch& cfg) {
    if constexpr (N == 3) {  // computed nodes: the 3-D kernels only
        if (cfg.all_uniform && cfg.all_analytic) return launch_bin_variant<T, N, IS_ND, PRE, true, true>(P, G, cfg);
    }
    return cfg.all_uniform ? launch_bin_variant<T, N, IS_ND, PRE, true>(P, G, cfg) : launch_bin_variant<T, N, IS_ND, PRE, false>(P, G, cfg);
}
template <class T, int N, bool IS_ND>
cudaError_t launch_bin_pre(const Params<T>& P, const TileGeom& G, const TileLaunch& cfg) {
    if (P.extrap == NINT_EXTRAPOLATE_CLAMP) return launch_bin_uni<T, N, IS_ND, kPreClamp>(P, G, cfg);
    if (P.extrap == NINT_EXTRAPOLATE_WRAP) return launch_bin_uni<T, N, IS_ND, kPreWrap>(P, G, cfg);
    return launch_bin_uni<T, N, IS_ND, kPrePlain>(P, G, cfg);
}

template <class T, int N>
cudaError_t launch_eval_n(const CUtensorMap& tmap, const TileEval<T>& E, const TileGeom& G, const TileLaunch& cfg) {
    auto kern = tile_eval_kernel<T, N>;
    constexpr int kRec = tile_rec_bytes<T, N>();
    const unsigned smem = 2u * G.buf_bytes + (unsigned)(eval_stages(kRec) * kEvalChunk * kRec);
    {  // per device and function: set on every launch (a host thread may drive several GPUs)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<(unsigned)cfg.sm_count, kEvalThreads, smem, (cudaStream_t)cfg.stream>>>(tmap, E, G);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace

// Interp3D, and InterpND with three to five axes
template <>
cudaError_t launch_tile_bin<NINT_T>(const Params<NINT_T>& P, const TileGeom& G, const TileLaunch& cfg) {
    if (!cfg.is_nd) return cfg.ndim == 3 ? launch_bin_pre<NINT_T, 3, false>(P, G, cfg) : cudaErrorInvalidValue;
    switch (cfg.ndim) {
    case 3: return launch_bin_pre<NINT_T, 3, true>(P, G, cfg);
    case 4: return launch_bin_pre<NINT_T, 4, true>(P, G, cfg);
    case 5: return launch_bin_pre<NINT_T, 5, true>(P, G, cfg);
    default: return cudaErrorInvalidValue;
    }
}

template <>
cudaError_t launch_tile_eval<NINT_T>(const CUtensorMap& tmap, const TileEval<NINT_T>& E, const TileGeom& G,
                                     const TileLaunch& cfg) {
    switch (cfg.ndim) {
    case 3: return launch_eval_n<NINT_T, 3>(tmap, E, G, cfg);
    case 4: return launch_eval_n<NINT_T, 4>(tmap, E, G, cfg);
    case 5: return launch_eval_n<NINT_T, 5>(tmap, E, G, cfg);
    default: return cudaErrorInvalidValue;
    }
}

template <>
cudaError_t launch_order_probe2<NINT_T>(const Params<NINT_T>& P, int ndim, int shift, unsigned max_distinct_total, ProbeSlot* slot,
                                        void* stream) {
    order_probe2_kernel<NINT_T><<<kProbeWindows, kProbeThreads, 0, (cudaStream_t)stream>>>(P, ndim, shift, max_distinct_total, slot);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace nint
