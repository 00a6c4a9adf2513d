/* ninterp_cuda.h — C ABI of the B200 batch engine for ninterp's `Interpolator::interpolate`.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The reference
 * crate (NREL/ninterp 0.8.1, pure Rust) has no FFI and no batch entry point today; each
 * function below names the reference item it stands in for (paths relative to the reference
 * repository).  INTEGRATION.md shows the Rust `extern "C"` block and the `interpolate_batch`
 * methods a maintainer would add on top of it.
 *
 * Conventions
 *  - every function returns an `int`: 0 = ok, negative = `nint_error`; nothing throws.
 *  - `nint_last_error_message()` / `nint_last_error_dim()` are thread-local details of the last
 *    failing call on this thread (the `dim {i}` of `ValidateError`, src/error.rs:19-23).
 *  - a handle owns private device copies of the grid and the value table (host buffers may be
 *    freed as soon as `nint_create` returns); it is immutable during batch calls, so concurrent
 *    `nint_interpolate_batch` calls on different streams are allowed.  Setters need external
 *    synchronisation, like `&mut self` in Rust.
 *  - the element type T is `float` or `double` as chosen at creation; all `void*` data are T.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *    NINT_E_CUDA.
 */
#ifndef NINTERP_CUDA_H
#define NINTERP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NINT_MAX_NDIM 8

#if defined(__GNUC__)
#define NINT_API __attribute__((visibility("default")))
#else
#define NINT_API
#endif

typedef struct nint_handle nint_handle;

typedef enum nint_dtype { NINT_F32 = 0, NINT_F64 = 1 } nint_dtype;

/* Interp1D/Interp2D/Interp3D (hard-coded, src/interpolator/{one,two,three}/mod.rs) versus
 * InterpND (src/interpolator/n/mod.rs).  They differ on points that sit exactly on a node. */
typedef enum nint_kind { NINT_KIND_FIXED = 0, NINT_KIND_ND = 1 } nint_kind;

/* strategy::{Linear, Nearest, LeftNearest, RightNearest} (src/strategy/mod.rs:14,25,33,41).
 * Left/RightNearest exist for Interp1D only (src/strategy/enums/one.rs:8-13). */
typedef enum nint_strategy {
    NINT_LINEAR = 0,
    NINT_NEAREST = 1,
    NINT_LEFT_NEAREST = 2,
    NINT_RIGHT_NEAREST = 3
} nint_strategy;

/* Extrapolate<T> (src/interpolator/mod.rs:57-72). */
typedef enum nint_extrapolate {
    NINT_EXTRAPOLATE_ENABLE = 0,
    NINT_EXTRAPOLATE_FILL = 1,
    NINT_EXTRAPOLATE_CLAMP = 2,
    NINT_EXTRAPOLATE_WRAP = 3,
    NINT_EXTRAPOLATE_ERROR = 4
} nint_extrapolate;

typedef enum nint_layout {
    NINT_LAYOUT_SOA = 0, /* ndim arrays of n coordinates each */
    NINT_LAYOUT_AOS = 1  /* n points of ndim coordinates each: a Rust `&[[T; N]]` */
} nint_layout;

typedef enum nint_error {
    NINT_OK = 0,
    NINT_E_EXTRAPOLATE_SELECTION = -1, /* ValidateError::ExtrapolateSelection  src/error.rs:17 */
    NINT_E_EMPTY_GRID = -2,            /* ValidateError::EmptyGrid(dim)        src/error.rs:19 */
    NINT_E_MONOTONICITY = -3,          /* ValidateError::Monotonicity(dim)     src/error.rs:21 */
    NINT_E_INCOMPATIBLE_SHAPES = -4,   /* ValidateError::IncompatibleShapes    src/error.rs:23 */
    NINT_E_OTHER = -5,                 /* ValidateError::Other(msg)            src/error.rs:25 */
    NINT_E_POINT_LENGTH = -6,          /* InterpolateError::PointLength(n)     src/error.rs:42 */
    NINT_E_EXTRAPOLATE = -7,           /* InterpolateError::ExtrapolateError   src/error.rs:40 */
    NINT_E_PANIC = -8,                 /* the reference would panic on this input (index out of bounds) */
    NINT_E_CUDA = -9,                  /* CUDA runtime failure, or no device */
    NINT_E_INVALID_ARGUMENT = -10,
    NINT_E_UNSUPPORTED = -11           /* e.g. ndim > NINT_MAX_NDIM, Left/RightNearest outside 1-D */
} nint_error;

/* Per-point status byte written by the batch calls (stands in for the per-call `Result`). */
#define NINT_STATUS_OK 0u
#define NINT_STATUS_EXTRAPOLATE_ERROR 1u /* Err(ExtrapolateError): three/mod.rs:225-236 */
#define NINT_STATUS_PANIC 0xFFu          /* the reference panics on this point (NaN coordinate, ...) */

/* Batch summary.  Indices are positions in the batch; UINT64_MAX when the count is 0. */
typedef struct nint_batch_info {
    uint64_t n_error;     /* points with NINT_STATUS_EXTRAPOLATE_ERROR */
    uint64_t n_panic;     /* points with NINT_STATUS_PANIC */
    uint64_t first_error; /* smallest index with an extrapolation error */
    uint64_t first_panic; /* smallest index with a panic */
} nint_batch_info;

/* Interp1D::new / Interp2D::new / Interp3D::new / InterpND::new
 * (one/mod.rs:110-124, two/mod.rs:118-133, three/mod.rs:129-145, n/mod.rs:225-239).
 *
 * device        CUDA ordinal that receives the grid and the value table.
 * ngrid         number of grid axes; axis_len[i] nodes at axes_host[i] (may be 0 / NULL).
 * nvalue_dim    rank of the value array, value_shape[i] its extents.  For NINT_KIND_FIXED it must
 *               equal ngrid (1..3); InterpND accepts a mismatch only to reject it the way
 *               n/mod.rs:71-83 does, and the 0-D case (one value, empty grids).
 * value_strides element strides of values_host, NULL = C order.  Strided `ArrayView`s
 *               (data.rs:92-97) are gathered into a dense device copy at upload.
 * fill          T*, read only for NINT_EXTRAPOLATE_FILL.
 * Runs data.validate() (data.rs:69-89, n/mod.rs:71-101) then check_extrapolate()
 * (interpolator/mod.rs:83-107) and reports their errors. */
NINT_API int nint_create(int device, int kind, int dtype, int ngrid, const size_t* axis_len,
                const void* const* axes_host, int nvalue_dim, const size_t* value_shape,
                const void* values_host, const ptrdiff_t* value_strides, int strategy, int extrapolate,
                const void* fill, nint_handle** out);

/* Drop for the interpolator. */
NINT_API void nint_destroy(nint_handle* h);

/* Interpolator::ndim (interpolator/mod.rs:27; InterpND: 0 when it holds a single value). */
NINT_API int nint_ndim(const nint_handle* h);

/* Interpolator::validate (three/mod.rs:186-191): check_extrapolate, then data.validate(). */
NINT_API int nint_validate(nint_handle* h);

/* Interpolator::set_extrapolate (three/mod.rs:240-244), including the reference's quirk that the
 * ">= 2 nodes per axis" test looks at the *current* setting (interpolator/mod.rs:97). */
NINT_API int nint_set_extrapolate(nint_handle* h, int extrapolate, const void* fill);

/* InterpXD::set_strategy for the built-in strategy enums (three/mod.rs:259-271). */
NINT_API int nint_set_strategy(nint_handle* h, int strategy);

/* The batch form of Interpolator::interpolate (interpolator/mod.rs:31): evaluates n points that
 * are already on the handle's device.  Asynchronous on `stream` (a cudaStream_t, NULL = default).
 *
 * coords    ndim device pointers, coords[d][i] = coordinate d of point i (struct of arrays).
 * out       device, n values.  Points whose status is not OK receive a quiet NaN.
 * status    device, n bytes, or NULL.
 * info_dev  device nint_batch_info accumulated by the kernel (zeroed by this call), or NULL.
 */
NINT_API int nint_interpolate_batch(nint_handle* h, const void* const* coords, size_t n, void* out,
                           uint8_t* status, nint_batch_info* info_dev, void* stream);

/* Same, for n points stored as an array of [T; ndim] on the device (what `&[[T; N]]` is). */
NINT_API int nint_interpolate_batch_aos(nint_handle* h, const void* points, size_t n, void* out, uint8_t* status,
                               nint_batch_info* info_dev, void* stream);

/* Host-buffer form: what a Rust caller holding `&[[T; N]]` / `Vec<T>` columns calls.  Streams the
 * batch through the device in chunks (H2D, kernel and D2H overlapped on three streams) and
 * returns when `out_host` is complete.  Pinned host buffers (nint_host_alloc) reach PCIe speed.
 *
 * layout NINT_LAYOUT_SOA: `data` is `const void* const*` with ndim host column pointers;
 *        NINT_LAYOUT_AOS: `data` is the host array of points.
 * Returns NINT_E_EXTRAPOLATE if any point failed under Extrapolate::Error, NINT_E_PANIC if any
 * point would make the reference panic (both after finishing the whole batch), else 0. */
NINT_API int nint_interpolate_batch_host(nint_handle* h, int layout, const void* data, size_t n, void* out_host,
                                uint8_t* status_host, nint_batch_info* info_host);

/* Host-buffer batch sharded over several GPUs of one box: `handles` are nh interpolators built from the same
 * data on different devices (the grid is replicated); the batch is split into nh contiguous index ranges that
 * run concurrently, one host thread per device, each through nint_interpolate_batch_host, so the host<->device
 * copies use every GPU's own PCIe link.  There is no collective: results land in place in out_host.  Indices
 * in info_host refer to the whole batch.  Return value as for nint_interpolate_batch_host. */
NINT_API int nint_interpolate_batch_host_multi(nint_handle* const* handles, int nh, int layout, const void* data, size_t n,
                                               void* out_host, uint8_t* status_host, nint_batch_info* info_host);

/* Interpolator::interpolate(&self, point: &[T]) -> Result<T, InterpolateError> for one point
 * (three/mod.rs:193-238): point_len != ndim gives NINT_E_POINT_LENGTH. */
NINT_API int nint_interpolate(nint_handle* h, const void* point, size_t point_len, void* out);

/* Diagnostics: how the most recent batch on this handle was classified by the on-device order probe that runs
 * for Linear on tables well beyond L2: 0 = no probe ran, 1 = coherent (single pass on the
 * cell-packed table), 2 = incoherent (slab lists, slab passes or binned tiles, see nint_last_incoherent_route).
 * Synchronises the device. */
NINT_API int nint_last_query_order(nint_handle* h);

/* Diagnostics: what the most recent batch on this handle enqueued for an incoherent verdict of the order probe:
 * 0 = nothing (no probe ran), 1 = slab passes on the plain table, 2 = binned tiles staged in shared memory by TMA,
 * 3 = slab lists (the batch split once by slab of the cell-packed table, evaluated list by list, merged back).
 * Host-side state, no synchronisation. */
NINT_API int nint_last_incoherent_route(nint_handle* h);

/* Pinned host memory for nint_interpolate_batch_host. */
NINT_API int nint_host_alloc(void** ptr, size_t bytes);
NINT_API int nint_host_free(void* ptr);

NINT_API const char* nint_last_error_message(void);
NINT_API int nint_last_error_dim(void);

/* Device self-test: n pseudo-random and adversarial (a, w) pairs, counting where the kernels'
 * FMA-corrected reciprocal division differs from IEEE `a / w` (must be 0), and how many pairs
 * took the fast path. */
NINT_API int nint_selftest_divide(int dtype, uint64_t n, uint64_t seed, uint64_t* mismatches,
                                  uint64_t* fast_path_taken);

/* Library version, and the number of kernel launches issued by this process (bench accounting). */
NINT_API const char* nint_version(void);
NINT_API uint64_t nint_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* NINTERP_CUDA_H */
