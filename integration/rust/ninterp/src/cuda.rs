//! `cuda` feature module for the ninterp crate: `to_cuda` on the interpolators and `interpolate_batch` on the
//! returned handle, over `ninterp-cuda-sys`.  Shown for `Interp3D`; `Interp1D`, `Interp2D` and `InterpND` follow the
//! same shape (`kind = NINT_KIND_ND` for `InterpND`).  Source only: no Rust toolchain exists in the image this was
//! written in, so it has not been compiled; INTEGRATION.md section 2 quotes this file and tests/test_abi.py keeps
//! the two identical.
use std::fmt::Debug;
use std::os::raw::{c_int, c_void};

use ndarray::{Array1, Data, RawDataClone};
use ninterp_cuda_sys as sys;

use crate::prelude::*;

fn last_error_message() -> String {
    unsafe { std::ffi::CStr::from_ptr(sys::nint_last_error_message()) }.to_string_lossy().into_owned()
}

// ---- quoted in INTEGRATION.md section 2 ----
/// Marker for the element types the engine evaluates.
pub trait CudaElem: Copy { const DTYPE: c_int; }
impl CudaElem for f32 { const DTYPE: c_int = sys::NINT_F32; }
impl CudaElem for f64 { const DTYPE: c_int = sys::NINT_F64; }

/// Maps the built-in strategies to the C enum; custom `Box<dyn Strategy3D>` strategies have no mapping and
/// `interpolate_batch` returns `InterpolateError::Other` for them (there is no CPU fallback).
pub trait CudaStrategy { fn code(&self) -> Option<c_int>; }
impl CudaStrategy for strategy::Linear  { fn code(&self) -> Option<c_int> { Some(sys::NINT_LINEAR) } }
impl CudaStrategy for strategy::Nearest { fn code(&self) -> Option<c_int> { Some(sys::NINT_NEAREST) } }
impl CudaStrategy for strategy::enums::Strategy3DEnum {
    fn code(&self) -> Option<c_int> { Some(match self { Self::Linear(_) => sys::NINT_LINEAR, Self::Nearest(_) => sys::NINT_NEAREST }) }
}

pub struct CudaHandle(*mut sys::nint_handle);
impl Drop for CudaHandle { fn drop(&mut self) { unsafe { sys::nint_destroy(self.0) } } }
unsafe impl Send for CudaHandle {}
unsafe impl Sync for CudaHandle {}   // immutable during batch calls, like `&self`

impl<D, S> Interp3D<D, S>
where D: Data + RawDataClone + Clone, D::Elem: CudaElem + PartialOrd + Debug, S: Strategy3D<D> + CudaStrategy + Clone,
{
    /// Uploads grid and values to `device` (validation is re-run by the library: data.rs:69-89, mod.rs:83-107).
    pub fn to_cuda(&self, device: i32) -> Result<CudaHandle, ValidateError> {
        let strat = self.strategy.code().ok_or_else(|| ValidateError::Other("custom strategies cannot run on the GPU".into()))?;
        let grids: Vec<Array1<D::Elem>> = self.data.grid.iter().map(|g| g.to_owned()).collect();   // contiguous axes
        let lens: [usize; 3] = std::array::from_fn(|i| grids[i].len());
        let ptrs: [*const c_void; 3] = std::array::from_fn(|i| grids[i].as_ptr() as *const c_void);
        let shape = self.data.values.shape();
        let strides: Vec<isize> = self.data.values.strides().to_vec();
        let (ext, fill) = match &self.extrapolate {
            Extrapolate::Enable => (sys::NINT_EXTRAPOLATE_ENABLE, None), Extrapolate::Fill(v) => (sys::NINT_EXTRAPOLATE_FILL, Some(*v)),
            Extrapolate::Clamp => (sys::NINT_EXTRAPOLATE_CLAMP, None),   Extrapolate::Wrap => (sys::NINT_EXTRAPOLATE_WRAP, None),
            Extrapolate::Error => (sys::NINT_EXTRAPOLATE_ERROR, None),
        };
        let mut h = std::ptr::null_mut();
        let rc = unsafe { sys::nint_create(device, sys::NINT_KIND_FIXED, D::Elem::DTYPE, 3, lens.as_ptr(), ptrs.as_ptr(), 3,
            shape.as_ptr(), self.data.values.as_ptr() as *const c_void, strides.as_ptr(), strat, ext,
            fill.as_ref().map_or(std::ptr::null(), |v| v as *const _ as *const c_void), &mut h) };
        validate_result(rc).map(|_| CudaHandle(h))
    }
}

impl CudaHandle {
    /// The batch form of `Interpolator::interpolate`: `points` is what a caller's loop would have passed one at a
    /// time.  Mirrors `Result<T, InterpolateError>` per batch: the first out-of-bounds point under
    /// `Extrapolate::Error` gives `Err(ExtrapolateError(..))` with the message of `three/mod.rs:226-229` for that
    /// point (formatted here on the host from `first_error`).
    pub fn interpolate_batch<T: CudaElem, const N: usize>(&self, points: &[[T; N]]) -> Result<Vec<T>, InterpolateError> {
        if unsafe { sys::nint_ndim(self.0) } as usize != N { return Err(InterpolateError::PointLength(N)); }
        let mut out = Vec::<T>::with_capacity(points.len());
        let mut info = sys::nint_batch_info::default();
        let rc = unsafe { sys::nint_interpolate_batch_host(self.0, sys::NINT_LAYOUT_AOS, points.as_ptr() as *const c_void,
            points.len(), out.as_mut_ptr() as *mut c_void, std::ptr::null_mut(), &mut info) };
        match rc {
            sys::NINT_OK => { unsafe { out.set_len(points.len()) }; Ok(out) }
            sys::NINT_E_EXTRAPOLATE => Err(InterpolateError::ExtrapolateError(format!(
                "\n    point index {} is out of bounds ({} point(s) in the batch)", info.first_error, info.n_error))),
            sys::NINT_E_PANIC => panic!("index out of bounds: point {} (the scalar path panics on this input too)", info.first_panic),
            sys::NINT_E_POINT_LENGTH => Err(InterpolateError::PointLength(N)),
            _ => Err(InterpolateError::Other(last_error_message())),
        }
    }
}

fn validate_result(rc: c_int) -> Result<(), ValidateError> {
    let dim = unsafe { sys::nint_last_error_dim() } as usize;
    match rc {
        sys::NINT_OK => Ok(()),
        sys::NINT_E_EXTRAPOLATE_SELECTION => Err(ValidateError::ExtrapolateSelection(last_error_message())),
        sys::NINT_E_EMPTY_GRID => Err(ValidateError::EmptyGrid(dim)),
        sys::NINT_E_MONOTONICITY => Err(ValidateError::Monotonicity(dim)),
        sys::NINT_E_INCOMPATIBLE_SHAPES => Err(ValidateError::IncompatibleShapes(dim)),
        _ => Err(ValidateError::Other(last_error_message())),
    }
}
