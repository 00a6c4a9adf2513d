//! `cuda` feature module for the ninterp crate: `to_cuda` on `Interp1D`, `Interp2D`, `Interp3D`, `InterpND` and
//! `InterpolatorEnum`, and `interpolate_batch` on the returned handle, over `ninterp-cuda-sys`.
//! Source only: no Rust toolchain exists in the image this was written in, so it has not been compiled;
//! INTEGRATION.md section 2 quotes this file and tests/test_abi.py keeps the two identical and checks every
//! `sys::` call against include/ninterp_cuda.h.
use std::fmt::Debug;
use std::os::raw::{c_int, c_void};

use ndarray::{Array1, ArrayBase, Data, Dimension, Ix1, RawDataClone};
use ninterp_cuda_sys as sys;

use crate::prelude::*;

fn last_error_message() -> String {
    unsafe { std::ffi::CStr::from_ptr(sys::nint_last_error_message()) }.to_string_lossy().into_owned()
}

// ---- quoted in INTEGRATION.md section 2 ----
/// Marker for the element types the engine evaluates.
pub trait CudaElem: Copy + PartialOrd + Debug { const DTYPE: c_int; }
impl CudaElem for f32 { const DTYPE: c_int = sys::NINT_F32; }
impl CudaElem for f64 { const DTYPE: c_int = sys::NINT_F64; }

/// Maps the built-in strategies to the C enum; custom `Box<dyn StrategyXD>` strategies have no mapping and
/// `to_cuda` returns `ValidateError::Other` for them (there is no CPU fallback).
pub trait CudaStrategy { fn code(&self) -> Option<c_int>; }
impl CudaStrategy for strategy::Linear       { fn code(&self) -> Option<c_int> { Some(sys::NINT_LINEAR) } }
impl CudaStrategy for strategy::Nearest      { fn code(&self) -> Option<c_int> { Some(sys::NINT_NEAREST) } }
impl CudaStrategy for strategy::LeftNearest  { fn code(&self) -> Option<c_int> { Some(sys::NINT_LEFT_NEAREST) } }
impl CudaStrategy for strategy::RightNearest { fn code(&self) -> Option<c_int> { Some(sys::NINT_RIGHT_NEAREST) } }
impl CudaStrategy for strategy::enums::Strategy1DEnum {   // strategy/enums/one.rs:8-13
    fn code(&self) -> Option<c_int> {
        Some(match self {
            Self::Linear(_) => sys::NINT_LINEAR, Self::Nearest(_) => sys::NINT_NEAREST,
            Self::LeftNearest(_) => sys::NINT_LEFT_NEAREST, Self::RightNearest(_) => sys::NINT_RIGHT_NEAREST,
        })
    }
}
macro_rules! two_variant_enum { ($($E:ident),*) => { $(                     // strategy/enums/{two,three,n}.rs:8-11
    impl CudaStrategy for strategy::enums::$E {
        fn code(&self) -> Option<c_int> { Some(match self { Self::Linear(_) => sys::NINT_LINEAR, Self::Nearest(_) => sys::NINT_NEAREST }) }
    })* } }
two_variant_enum!(Strategy2DEnum, Strategy3DEnum, StrategyNDEnum);

/// A device-resident interpolator.  Keeps owned copies of the grid axes and the policy so that
/// `ExtrapolateError` can be formatted exactly as the scalar path formats it.
pub struct CudaHandle<T: CudaElem> {
    raw: *mut sys::nint_handle,
    grid: Vec<Array1<T>>,
    one_d: bool,        // Interp1D reports dim 0 only and returns at once (one/mod.rs:198-203)
}
impl<T: CudaElem> Drop for CudaHandle<T> { fn drop(&mut self) { unsafe { sys::nint_destroy(self.raw) } } }
unsafe impl<T: CudaElem> Send for CudaHandle<T> {}
// `interpolate_batch` takes `&self`; the setters below take `&mut self`, so the borrow checker provides the
// external synchronisation the C ABI asks for.
unsafe impl<T: CudaElem> Sync for CudaHandle<T> {}

fn extrapolate_code<T: Copy>(e: &Extrapolate<T>) -> (c_int, Option<T>) {
    match e {
        Extrapolate::Enable => (sys::NINT_EXTRAPOLATE_ENABLE, None), Extrapolate::Fill(v) => (sys::NINT_EXTRAPOLATE_FILL, Some(*v)),
        Extrapolate::Clamp => (sys::NINT_EXTRAPOLATE_CLAMP, None),   Extrapolate::Wrap => (sys::NINT_EXTRAPOLATE_WRAP, None),
        Extrapolate::Error => (sys::NINT_EXTRAPOLATE_ERROR, None),
    }
}

/// `InterpXD::new` on the device: uploads grid and values (validation is re-run by the library: data.rs:69-89,
/// n/mod.rs:71-101, interpolator/mod.rs:83-107).  Owned and viewed data both work: a strided `ArrayView` is
/// passed with its element strides (data.rs:92-97) and gathered at upload.
fn create<T, D, Dm>(device: i32, kind: c_int, grid: &[ArrayBase<D, Ix1>], values: &ArrayBase<D, Dm>, strat: Option<c_int>,
                    extrapolate: &Extrapolate<T>) -> Result<CudaHandle<T>, ValidateError>
where T: CudaElem, D: Data<Elem = T>, Dm: Dimension,
{
    let strat = strat.ok_or_else(|| ValidateError::Other("custom strategies cannot run on the GPU".into()))?;
    let grids: Vec<Array1<T>> = grid.iter().map(|g| g.to_owned()).collect();   // contiguous axes
    let lens: Vec<usize> = grids.iter().map(|g| g.len()).collect();
    let ptrs: Vec<*const c_void> = grids.iter().map(|g| g.as_ptr() as *const c_void).collect();
    let shape: Vec<usize> = values.shape().to_vec();
    let strides: Vec<isize> = values.strides().to_vec();
    let (ext, fill) = extrapolate_code(extrapolate);
    let mut raw = std::ptr::null_mut();
    let rc = unsafe { sys::nint_create(device, kind, T::DTYPE, grids.len() as c_int, lens.as_ptr(), ptrs.as_ptr(),
        shape.len() as c_int, shape.as_ptr(), values.as_ptr() as *const c_void, strides.as_ptr(), strat, ext,
        fill.as_ref().map_or(std::ptr::null(), |v| v as *const T as *const c_void), &mut raw) };
    validate_result(rc).map(|_| CudaHandle { raw, one_d: kind == sys::NINT_KIND_FIXED && grids.len() == 1, grid: grids })
}

macro_rules! to_cuda_impl { ($Interp:ident, $Strategy:ident, $kind:expr) => {
    impl<D, S> $Interp<D, S>
    where D: Data + RawDataClone + Clone, D::Elem: CudaElem, S: $Strategy<D> + CudaStrategy + Clone,
    {
        /// Uploads this interpolator to `device`; the result evaluates batches with the same semantics as
        /// `Interpolator::interpolate` on `self`.
        pub fn to_cuda(&self, device: i32) -> Result<CudaHandle<D::Elem>, ValidateError> {
            create(device, $kind, &self.data.grid[..], &self.data.values, self.strategy.code(), &self.extrapolate)
        }
    }
} }
to_cuda_impl!(Interp1D, Strategy1D, sys::NINT_KIND_FIXED);   // one/mod.rs:110-124
to_cuda_impl!(Interp2D, Strategy2D, sys::NINT_KIND_FIXED);   // two/mod.rs:118-133
to_cuda_impl!(Interp3D, Strategy3D, sys::NINT_KIND_FIXED);   // three/mod.rs:129-145
to_cuda_impl!(InterpND, StrategyND, sys::NINT_KIND_ND);      // n/mod.rs:225-239

/// `InterpolatorEnum` on the device (enums.rs:78-88): `Interp0D` stays a host constant (zero/mod.rs:51-56), the
/// other variants own a handle.
pub enum CudaInterpolator<T: CudaElem> { Constant(T), Device(CudaHandle<T>) }

impl<D> InterpolatorEnum<D>
where D: Data + RawDataClone + Clone, D::Elem: CudaElem + num_traits::Num + num_traits::Euclid,
{
    pub fn to_cuda(&self, device: i32) -> Result<CudaInterpolator<D::Elem>, ValidateError> {
        Ok(match self {
            InterpolatorEnum::Interp0D(i) => CudaInterpolator::Constant(i.0),
            InterpolatorEnum::Interp1D(i) => CudaInterpolator::Device(i.to_cuda(device)?),
            InterpolatorEnum::Interp2D(i) => CudaInterpolator::Device(i.to_cuda(device)?),
            InterpolatorEnum::Interp3D(i) => CudaInterpolator::Device(i.to_cuda(device)?),
            InterpolatorEnum::InterpND(i) => CudaInterpolator::Device(i.to_cuda(device)?),
        })
    }
}

impl<T: CudaElem> CudaInterpolator<T> {
    /// `Interpolator::ndim` (enums.rs:225-233).
    pub fn ndim(&self) -> usize { match self { Self::Constant(_) => 0, Self::Device(h) => h.ndim() } }
    /// The batch form of `InterpolatorEnum::interpolate` (enums.rs:247-255): `points` holds `n` points of `ndim()`
    /// coordinates each, back to back (`n` is given because a 0-D point has no coordinates).
    pub fn interpolate_batch_flat(&self, points: &[T], n: usize) -> Result<Vec<T>, InterpolateError> {
        match self {
            Self::Constant(v) => if points.is_empty() { Ok(vec![*v; n]) } else { Err(InterpolateError::PointLength(0)) },  // zero/mod.rs:52-54
            Self::Device(h) => h.interpolate_batch_flat(points, n),
        }
    }
}

impl<T: CudaElem> CudaHandle<T> {
    /// `Interpolator::ndim` (interpolator/mod.rs:27; 0 for an `InterpND` holding a single value, n/mod.rs:104-110).
    pub fn ndim(&self) -> usize { unsafe { sys::nint_ndim(self.raw) as usize } }

    /// `Interpolator::validate` (three/mod.rs:186-191).
    pub fn validate(&mut self) -> Result<(), ValidateError> { validate_result(unsafe { sys::nint_validate(self.raw) }) }

    /// `Interpolator::set_extrapolate` (three/mod.rs:240-244).
    pub fn set_extrapolate(&mut self, extrapolate: Extrapolate<T>) -> Result<(), ValidateError> {
        let (ext, fill) = extrapolate_code(&extrapolate);
        validate_result(unsafe { sys::nint_set_extrapolate(self.raw, ext, fill.as_ref().map_or(std::ptr::null(), |v| v as *const T as *const c_void)) })
    }

    /// `InterpXD::set_strategy` for the built-in strategy enums (three/mod.rs:259-271).
    pub fn set_strategy(&mut self, strategy: &impl CudaStrategy) -> Result<(), ValidateError> {
        let code = strategy.code().ok_or_else(|| ValidateError::Other("custom strategies cannot run on the GPU".into()))?;
        validate_result(unsafe { sys::nint_set_strategy(self.raw, code) })
    }

    /// The batch form of `Interpolator::interpolate` for the hard-coded interpolators: `points` is what the caller's
    /// loop (benches/benchmark.rs:149-151) would have passed one at a time.
    pub fn interpolate_batch<const N: usize>(&self, points: &[[T; N]]) -> Result<Vec<T>, InterpolateError> {
        if self.ndim() != N { return Err(InterpolateError::PointLength(self.ndim())); }   // three/mod.rs:194-196
        let flat = unsafe { std::slice::from_raw_parts(points.as_ptr() as *const T, points.len() * N) };
        self.interpolate_batch_flat(flat, points.len())
    }

    /// Runtime-N form (what `InterpND::interpolate(&[T])` takes, n/mod.rs:286-290): `n` points of `ndim()`
    /// coordinates each, back to back.  Mirrors the scalar `Result<T, InterpolateError>` for the batch: the first
    /// out-of-bounds point under `Extrapolate::Error` gives the `ExtrapolateError` the scalar call gives for it.
    pub fn interpolate_batch_flat(&self, points: &[T], n: usize) -> Result<Vec<T>, InterpolateError> {
        let nd = self.ndim();
        if points.len() != n * nd { return Err(InterpolateError::PointLength(nd)); }
        let mut out = Vec::<T>::with_capacity(n);
        let mut info = sys::nint_batch_info::default();
        let rc = unsafe { sys::nint_interpolate_batch_host(self.raw, sys::NINT_LAYOUT_AOS, points.as_ptr() as *const c_void,
            n, out.as_mut_ptr() as *mut c_void, std::ptr::null_mut(), &mut info) };
        match rc {
            sys::NINT_OK => { unsafe { out.set_len(n) }; Ok(out) }
            sys::NINT_E_EXTRAPOLATE => {
                let i = info.first_error as usize;
                Err(InterpolateError::ExtrapolateError(self.out_of_bounds_message(&points[i * nd..(i + 1) * nd])))
            }
            sys::NINT_E_PANIC => panic!("index out of bounds: point {} (the scalar path panics on this input too)", info.first_panic),
            sys::NINT_E_POINT_LENGTH => Err(InterpolateError::PointLength(nd)),
            _ => Err(InterpolateError::Other(last_error_message())),
        }
    }

    /// The text `Interpolator::interpolate` puts into `ExtrapolateError` for this point under `Extrapolate::Error`:
    /// one line per out-of-bounds dimension, in order (three/mod.rs:226-229, n/mod.rs:326-329); `Interp1D` has the
    /// single line of one/mod.rs:198-203.  The grid is printed with ndarray's `Debug`, as the reference does.
    fn out_of_bounds_message(&self, point: &[T]) -> String {
        let mut errors = Vec::new();
        for (dim, g) in self.grid.iter().enumerate() {
            let (first, last) = (g.first().unwrap(), g.last().unwrap());
            if !(first..=last).contains(&&point[dim]) {
                errors.push(format!("\n    point[{dim}] = {:?} is out of bounds for grid[{dim}] = {:?}", point[dim], g));
                if self.one_d { break; }
            }
        }
        errors.join("")
    }
}

fn validate_result(rc: c_int) -> Result<(), ValidateError> {
    let dim = unsafe { sys::nint_last_error_dim() } as usize;
    match rc {
        sys::NINT_OK => Ok(()),
        // error.rs:17 takes the Debug text of the current policy; the library's message carries it in parentheses
        sys::NINT_E_EXTRAPOLATE_SELECTION => Err(ValidateError::ExtrapolateSelection(
            last_error_message().split(['(', ')']).nth(1).unwrap_or("").to_string())),
        sys::NINT_E_EMPTY_GRID => Err(ValidateError::EmptyGrid(dim)),
        sys::NINT_E_MONOTONICITY => Err(ValidateError::Monotonicity(dim)),
        sys::NINT_E_INCOMPATIBLE_SHAPES => Err(ValidateError::IncompatibleShapes(dim)),
        _ => Err(ValidateError::Other(last_error_message())),
    }
}
