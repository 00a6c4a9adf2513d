//! Raw declarations of `include/ninterp_cuda.h`, one per exported symbol, in header order.
//! Source only: this image has no Rust toolchain, so the crate has not been compiled here; the same ABI is
//! exercised from Python (ctypes) and C++ by the repository's test-suite, and `tests/test_abi.py` checks that
//! every symbol of the header is declared below.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct nint_handle {
    _private: [u8; 0],
}

pub const NINT_MAX_NDIM: usize = 8;

pub const NINT_F32: c_int = 0;
pub const NINT_F64: c_int = 1;

pub const NINT_KIND_FIXED: c_int = 0; // Interp1D / Interp2D / Interp3D
pub const NINT_KIND_ND: c_int = 1; // InterpND

pub const NINT_LINEAR: c_int = 0;
pub const NINT_NEAREST: c_int = 1;
pub const NINT_LEFT_NEAREST: c_int = 2;
pub const NINT_RIGHT_NEAREST: c_int = 3;

pub const NINT_EXTRAPOLATE_ENABLE: c_int = 0;
pub const NINT_EXTRAPOLATE_FILL: c_int = 1;
pub const NINT_EXTRAPOLATE_CLAMP: c_int = 2;
pub const NINT_EXTRAPOLATE_WRAP: c_int = 3;
pub const NINT_EXTRAPOLATE_ERROR: c_int = 4;

pub const NINT_LAYOUT_SOA: c_int = 0;
pub const NINT_LAYOUT_AOS: c_int = 1;

pub const NINT_OK: c_int = 0;
pub const NINT_E_EXTRAPOLATE_SELECTION: c_int = -1; // ValidateError::ExtrapolateSelection
pub const NINT_E_EMPTY_GRID: c_int = -2; // ValidateError::EmptyGrid(dim)
pub const NINT_E_MONOTONICITY: c_int = -3; // ValidateError::Monotonicity(dim)
pub const NINT_E_INCOMPATIBLE_SHAPES: c_int = -4; // ValidateError::IncompatibleShapes(dim)
pub const NINT_E_OTHER: c_int = -5; // ValidateError::Other
pub const NINT_E_POINT_LENGTH: c_int = -6; // InterpolateError::PointLength
pub const NINT_E_EXTRAPOLATE: c_int = -7; // InterpolateError::ExtrapolateError
pub const NINT_E_PANIC: c_int = -8; // the scalar path would panic on some point
pub const NINT_E_CUDA: c_int = -9;
pub const NINT_E_INVALID_ARGUMENT: c_int = -10;
pub const NINT_E_UNSUPPORTED: c_int = -11;

pub const NINT_STATUS_OK: u8 = 0;
pub const NINT_STATUS_EXTRAPOLATE_ERROR: u8 = 1;
pub const NINT_STATUS_PANIC: u8 = 0xFF;

#[repr(C)]
#[derive(Default, Clone, Copy, Debug, PartialEq, Eq)]
pub struct nint_batch_info {
    pub n_error: u64,
    pub n_panic: u64,
    pub first_error: u64,
    pub first_panic: u64,
}

extern "C" {
    pub fn nint_create(
        device: c_int,
        kind: c_int,
        dtype: c_int,
        ngrid: c_int,
        axis_len: *const usize,
        axes_host: *const *const c_void,
        nvalue_dim: c_int,
        value_shape: *const usize,
        values_host: *const c_void,
        value_strides: *const isize,
        strategy: c_int,
        extrapolate: c_int,
        fill: *const c_void,
        out: *mut *mut nint_handle,
    ) -> c_int;
    pub fn nint_destroy(h: *mut nint_handle);
    pub fn nint_ndim(h: *const nint_handle) -> c_int;
    pub fn nint_validate(h: *mut nint_handle) -> c_int;
    pub fn nint_set_extrapolate(h: *mut nint_handle, extrapolate: c_int, fill: *const c_void) -> c_int;
    pub fn nint_set_strategy(h: *mut nint_handle, strategy: c_int) -> c_int;
    pub fn nint_interpolate_batch(
        h: *mut nint_handle,
        coords: *const *const c_void,
        n: usize,
        out: *mut c_void,
        status: *mut u8,
        info_dev: *mut nint_batch_info,
        stream: *mut c_void,
    ) -> c_int;
    pub fn nint_interpolate_batch_aos(
        h: *mut nint_handle,
        points: *const c_void,
        n: usize,
        out: *mut c_void,
        status: *mut u8,
        info_dev: *mut nint_batch_info,
        stream: *mut c_void,
    ) -> c_int;
    pub fn nint_interpolate_batch_host(
        h: *mut nint_handle,
        layout: c_int,
        data: *const c_void,
        n: usize,
        out_host: *mut c_void,
        status_host: *mut u8,
        info_host: *mut nint_batch_info,
    ) -> c_int;
    pub fn nint_interpolate_batch_host_multi(
        handles: *const *mut nint_handle,
        nh: c_int,
        layout: c_int,
        data: *const c_void,
        n: usize,
        out_host: *mut c_void,
        status_host: *mut u8,
        info_host: *mut nint_batch_info,
    ) -> c_int;
    pub fn nint_interpolate(h: *mut nint_handle, point: *const c_void, point_len: usize, out: *mut c_void) -> c_int;
    pub fn nint_last_query_order(h: *mut nint_handle) -> c_int;
    pub fn nint_last_incoherent_route(h: *mut nint_handle) -> c_int;
    pub fn nint_host_alloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn nint_host_free(ptr: *mut c_void) -> c_int;
    pub fn nint_last_error_message() -> *const c_char;
    pub fn nint_last_error_dim() -> c_int;
    pub fn nint_selftest_divide(dtype: c_int, n: u64, seed: u64, mismatches: *mut u64, fast_path_taken: *mut u64) -> c_int;
    pub fn nint_version() -> *const c_char;
    pub fn nint_launch_count() -> u64;
}
