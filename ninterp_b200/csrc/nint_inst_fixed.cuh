// nint_inst_fixed.cuh — instantiates the Interp1D/2D/3D kernels for NINT_T (included by
// nint_fixed_f32.cu / nint_fixed_f64.cu): one kernel per (dimension, strategy, axes in shared or
// global memory).  The Extrapolate policy only matters on the out-of-line general path.
#include "nint_launch.cuh"

namespace nint {

namespace {

template <class T, int N, int STRAT>
cudaError_t dispatch_smem(const Params<T>& P, const LaunchCfg& cfg) {
    return cfg.smem_axes ? launch_one<T, N, false, STRAT, true>(P, cfg) : launch_one<T, N, false, STRAT, false>(P, cfg);
}

template <class T, int N>
cudaError_t dispatch_strategy(const Params<T>& P, const LaunchCfg& cfg) {
    switch (cfg.strategy) {
    case NINT_LINEAR: return dispatch_smem<T, N, NINT_LINEAR>(P, cfg);
    case NINT_NEAREST: return dispatch_smem<T, N, NINT_NEAREST>(P, cfg);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace

template <>
cudaError_t launch_fixed<NINT_T>(const Params<NINT_T>& P, const LaunchCfg& cfg) {
    if (cfg.pipe) {  // pipelined single pass: Interp2D / Interp3D, Linear and Nearest
        if (cfg.ndim == 3) return cfg.strategy == NINT_LINEAR ? launch_pipe<NINT_T, 3, false, NINT_LINEAR>(P, cfg) : launch_pipe<NINT_T, 3, false, NINT_NEAREST>(P, cfg);
        if (cfg.ndim == 2) return cfg.strategy == NINT_LINEAR ? launch_pipe<NINT_T, 2, false, NINT_LINEAR>(P, cfg) : launch_pipe<NINT_T, 2, false, NINT_NEAREST>(P, cfg);
        return cudaErrorInvalidValue;
    }
    if (cfg.slab) {
        switch (cfg.ndim) {
        case 1: return launch_slab<NINT_T, 1, false>(P, cfg);
        case 2: return launch_slab<NINT_T, 2, false>(P, cfg);
        case 3: return launch_slab<NINT_T, 3, false>(P, cfg);
        default: return cudaErrorInvalidValue;
        }
    }
    if (cfg.ndim == 1 && cfg.strategy == NINT_LINEAR && P.fused != nullptr) return launch_fused1d<NINT_T, false>(P, cfg);
    switch (cfg.ndim) {
    case 1:
        if (cfg.strategy == NINT_LEFT_NEAREST) return dispatch_smem<NINT_T, 1, NINT_LEFT_NEAREST>(P, cfg);
        if (cfg.strategy == NINT_RIGHT_NEAREST) return dispatch_smem<NINT_T, 1, NINT_RIGHT_NEAREST>(P, cfg);
        return dispatch_strategy<NINT_T, 1>(P, cfg);
    case 2: return dispatch_strategy<NINT_T, 2>(P, cfg);
    case 3: return dispatch_strategy<NINT_T, 3>(P, cfg);
    default: return cudaErrorInvalidValue;
    }
}

template <>
cudaError_t launch_pack_cells<NINT_T>(const Params<NINT_T>& P, int ndim, NINT_T* out, unsigned long long total,
                                      int sm_count, void* stream) {
    if (total == 0) return cudaSuccess;
    const unsigned long long want = (total + kBlock - 1) / kBlock;
    const unsigned long long cap = (unsigned long long)sm_count * 8ull;
    pack_cells_kernel<NINT_T><<<(unsigned)(want < cap ? want : cap), kBlock, 0, (cudaStream_t)stream>>>(P.values, out, ndim, P, total);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <>
cudaError_t launch_constant<NINT_T>(const NINT_T* values, NINT_T* out, uint8_t* status, unsigned long long n,
                                    int sm_count, void* stream) {
    if (n == 0) return cudaSuccess;
    const unsigned long long want = (n + kBlock - 1) / kBlock;
    const unsigned long long cap = (unsigned long long)sm_count * 8ull;
    constant_kernel<NINT_T><<<(unsigned)(want < cap ? want : cap), kBlock, 0, (cudaStream_t)stream>>>(values, out, status, n);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

}  // namespace nint
