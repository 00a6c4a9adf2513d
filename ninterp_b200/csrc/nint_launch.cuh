// nint_launch.cuh — launch helper shared by the instantiation units.
#pragma once
#include <atomic>

#include "nint_device.cuh"
#include "nint_pipe.cuh"

namespace nint {

extern std::atomic<unsigned long long> g_launch_count;

template <class T, int N, bool IS_ND, int STRAT, bool SMEM, int PRE, bool UNI, bool ANA = false>
cudaError_t launch_variant(const Params<T>& P, const LaunchCfg& cfg) {
    auto kern = interp_kernel<T, N, IS_ND, STRAT, SMEM, PRE, UNI, ANA>;
    const unsigned smem = SMEM ? cfg.smem_bytes : 0u;
    if (smem > 48u * 1024u) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // resident CTAs per SM for this kernel at this shared-memory size (cached per host thread; homogeneous GPUs)
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
    kern<<<grid, kBlock, smem, (cudaStream_t)cfg.stream>>>(P);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// The pipelined form of the single pass (nint_pipe.cuh) for the whole chunks of a batch; false when it does not
// apply (then nothing was launched).  The caller runs interp_kernel on the remaining points.
template <class T, int N, bool IS_ND, int STRAT, int PRE, bool UNI>
cudaError_t launch_pipe_variant(const Params<T>& P, const LaunchCfg& cfg) {
    auto kern = interp_pipe_kernel<T, N, IS_ND, STRAT, PRE, UNI>;
    const unsigned smem = (unsigned)(kPipeStages * N * kPipeChunk * sizeof(T)) + cfg.smem_bytes;
    {  // per device and function: set on every launch (a host thread may drive several GPUs)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    static thread_local unsigned cached_smem = ~0u;
    static thread_local int cached_bps = 0;
    int bps = cached_bps;
    if (cached_smem != smem || bps == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, kPipeThreads, smem);
        if (e != cudaSuccess) return e;
        if (bps < 1) bps = 1;
        cached_bps = bps;
        cached_smem = smem;
    }
    const unsigned long long chunks = P.n / kPipeChunk;
    const unsigned long long cap = (unsigned long long)cfg.sm_count * (unsigned long long)bps;
    const unsigned grid = (unsigned)(chunks < cap ? chunks : cap);
    if (grid == 0) return cudaSuccess;
    kern<<<grid, kPipeThreads, smem, (cudaStream_t)cfg.stream>>>(P);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}
template <class T, int N, bool IS_ND, int STRAT>
cudaError_t launch_pipe(const Params<T>& P, const LaunchCfg& cfg) {
    const bool uni = cfg.all_uniform && (STRAT != NINT_LINEAR || P.cells != nullptr);
    if (P.extrap == NINT_EXTRAPOLATE_CLAMP)
        return uni ? launch_pipe_variant<T, N, IS_ND, STRAT, kPreClamp, true>(P, cfg) : launch_pipe_variant<T, N, IS_ND, STRAT, kPreClamp, false>(P, cfg);
    if (P.extrap == NINT_EXTRAPOLATE_WRAP)
        return uni ? launch_pipe_variant<T, N, IS_ND, STRAT, kPreWrap, true>(P, cfg) : launch_pipe_variant<T, N, IS_ND, STRAT, kPreWrap, false>(P, cfg);
    return uni ? launch_pipe_variant<T, N, IS_ND, STRAT, kPrePlain, true>(P, cfg) : launch_pipe_variant<T, N, IS_ND, STRAT, kPrePlain, false>(P, cfg);
}

// Clamp and Wrap enter the fast path (clamp first / wrap and retry), so they select their own kernels; so do
// interpolators whose axes are all uniform (no guess tables, no per-axis flags in the point loop), wherever
// their axis pack lives.
template <class T, int N, bool IS_ND, int STRAT, bool SMEM, int PRE>
cudaError_t launch_uni(const Params<T>& P, const LaunchCfg& cfg) {
    // Analytic axes (nodes computed, not loaded): the hard-coded 2-D/3-D interpolators get the extra kernels, opt-in
    // with NINT_ANALYTIC=1.  Measured: faster where the L1 data pipe is the limit (queries random inside
    // small tiles +15 %, a 128^3 table +24 %), slower on cell-ordered batches (-13 %) and on 2048^2 under Clamp (-24 %).
    if constexpr (!IS_ND && (N == 2 || N == 3) && (STRAT == NINT_LINEAR || STRAT == NINT_NEAREST)) {
        if (cfg.all_uniform && cfg.all_analytic) return launch_variant<T, N, IS_ND, STRAT, SMEM, PRE, true, true>(P, cfg);
    }
    if (cfg.all_uniform && (STRAT != NINT_LINEAR || P.cells != nullptr))
        return launch_variant<T, N, IS_ND, STRAT, SMEM, PRE, true>(P, cfg);
    return launch_variant<T, N, IS_ND, STRAT, SMEM, PRE, false>(P, cfg);
}
template <class T, int N, bool IS_ND, int STRAT, bool SMEM>
cudaError_t launch_one(const Params<T>& P, const LaunchCfg& cfg) {
    if (P.extrap == NINT_EXTRAPOLATE_CLAMP) return launch_uni<T, N, IS_ND, STRAT, SMEM, kPreClamp>(P, cfg);
    if (P.extrap == NINT_EXTRAPOLATE_WRAP) return launch_uni<T, N, IS_ND, STRAT, SMEM, kPreWrap>(P, cfg);
    return launch_uni<T, N, IS_ND, STRAT, SMEM, kPrePlain>(P, cfg);
}

// 1-D Linear from the fused shared-memory pack (Params::fused).
template <class T, bool IS_ND, int PRE, bool UNI>
cudaError_t launch_fused1d_variant(const Params<T>& P, const LaunchCfg& cfg) {
    auto kern = interp1d_fused_kernel<T, IS_ND, PRE, UNI>;
    const unsigned smem = P.fused_bytes;
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
    kern<<<grid, kBlock, smem, (cudaStream_t)cfg.stream>>>(P);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}
template <class T, bool IS_ND>
cudaError_t launch_fused1d(const Params<T>& P, const LaunchCfg& cfg) {
    const bool uni = cfg.all_uniform != 0;
    if (P.extrap == NINT_EXTRAPOLATE_CLAMP)
        return uni ? launch_fused1d_variant<T, IS_ND, kPreClamp, true>(P, cfg) : launch_fused1d_variant<T, IS_ND, kPreClamp, false>(P, cfg);
    if (P.extrap == NINT_EXTRAPOLATE_WRAP)
        return uni ? launch_fused1d_variant<T, IS_ND, kPreWrap, true>(P, cfg) : launch_fused1d_variant<T, IS_ND, kPreWrap, false>(P, cfg);
    return uni ? launch_fused1d_variant<T, IS_ND, kPrePlain, true>(P, cfg) : launch_fused1d_variant<T, IS_ND, kPrePlain, false>(P, cfg);
}

// Slab passes (Linear only, axes in shared memory, policy Clamp or one that leaves the fast path alone).
template <class T, int N, bool IS_ND, int PRE, bool UNI, bool ANA = false, bool INR = false>
cudaError_t launch_slab_variant(const Params<T>& P, const LaunchCfg& cfg) {
    auto kern = slab_kernel<T, N, IS_ND, PRE, UNI, ANA, INR>;
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
    const unsigned long long want = (P.n + kSlabSpan * (kBlock / 32) - 1) / (kSlabSpan * (kBlock / 32));
    const unsigned long long cap = (unsigned long long)cfg.sm_count * (unsigned long long)bps;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (grid == 0) return cudaSuccess;
    kern<<<grid, kBlock, smem, (cudaStream_t)cfg.stream>>>(P);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}
template <class T, int N, bool IS_ND>
cudaError_t launch_slab(const Params<T>& P, const LaunchCfg& cfg) {
    const bool uni = cfg.all_uniform != 0;
    if constexpr (N >= 4 && IS_ND) {  // in-range compaction: InterpND with four or more axes, Fill and Error
        if (cfg.inrange) return uni ? launch_slab_variant<T, N, IS_ND, kPrePlain, true, false, true>(P, cfg)
                                    : launch_slab_variant<T, N, IS_ND, kPrePlain, false, false, true>(P, cfg);
    }
    if constexpr (N == 3 && !IS_ND) {  // analytic axes (nodes computed, IEEE division): the hard-coded 3-D interpolator only
        if (uni && cfg.slab_analytic) {
            if (P.extrap == NINT_EXTRAPOLATE_CLAMP) return launch_slab_variant<T, N, IS_ND, kPreClamp, true, true>(P, cfg);
            return launch_slab_variant<T, N, IS_ND, kPrePlain, true, true>(P, cfg);
        }
    }
    if (P.extrap == NINT_EXTRAPOLATE_CLAMP)
        return uni ? launch_slab_variant<T, N, IS_ND, kPreClamp, true>(P, cfg) : launch_slab_variant<T, N, IS_ND, kPreClamp, false>(P, cfg);
    return uni ? launch_slab_variant<T, N, IS_ND, kPrePlain, true>(P, cfg) : launch_slab_variant<T, N, IS_ND, kPrePlain, false>(P, cfg);
}

}  // namespace nint
