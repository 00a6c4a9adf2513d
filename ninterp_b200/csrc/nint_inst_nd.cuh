// nint_inst_nd.cuh — instantiates the InterpND kernels (N = 1..NINT_MAX_NDIM) for NINT_T.
#include "nint_launch.cuh"

namespace nint {

namespace {

template <class T, int N>
cudaError_t dispatch_nd(const Params<T>& P, const LaunchCfg& cfg) {
    if (cfg.strategy == NINT_LINEAR) {
        return cfg.smem_axes ? launch_one<T, N, true, NINT_LINEAR, true>(P, cfg)
                             : launch_one<T, N, true, NINT_LINEAR, false>(P, cfg);
    }
    if (cfg.strategy == NINT_NEAREST) {
        return cfg.smem_axes ? launch_one<T, N, true, NINT_NEAREST, true>(P, cfg)
                             : launch_one<T, N, true, NINT_NEAREST, false>(P, cfg);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

template <>
cudaError_t launch_nd<NINT_T>(const Params<NINT_T>& P, const LaunchCfg& cfg) {
    if (cfg.slab) {
        switch (cfg.ndim) {
        case 1: return launch_slab<NINT_T, 1, true>(P, cfg);
        case 2: return launch_slab<NINT_T, 2, true>(P, cfg);
        case 3: return launch_slab<NINT_T, 3, true>(P, cfg);
        case 4: return launch_slab<NINT_T, 4, true>(P, cfg);
        case 5: return launch_slab<NINT_T, 5, true>(P, cfg);
        case 6: return launch_slab<NINT_T, 6, true>(P, cfg);
        case 7: return launch_slab<NINT_T, 7, true>(P, cfg);
        case 8: return launch_slab<NINT_T, 8, true>(P, cfg);
        default: return cudaErrorInvalidValue;
        }
    }
    if (cfg.ndim == 1 && cfg.strategy == NINT_LINEAR && P.fused != nullptr) return launch_fused1d<NINT_T, true>(P, cfg);
    switch (cfg.ndim) {
    case 1: return dispatch_nd<NINT_T, 1>(P, cfg);
    case 2: return dispatch_nd<NINT_T, 2>(P, cfg);
    case 3: return dispatch_nd<NINT_T, 3>(P, cfg);
    case 4: return dispatch_nd<NINT_T, 4>(P, cfg);
    case 5: return dispatch_nd<NINT_T, 5>(P, cfg);
    case 6: return dispatch_nd<NINT_T, 6>(P, cfg);
    case 7: return dispatch_nd<NINT_T, 7>(P, cfg);
    case 8: return dispatch_nd<NINT_T, 8>(P, cfg);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace nint
