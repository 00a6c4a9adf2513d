// oracle_capi.cpp — C entry points over ninterp_oracle.hpp for ctypes (tests/oracle_binding.py).
//
// TEST INFRASTRUCTURE ONLY: the checker and the CPU baseline ("port" of ninterp 0.8.1), never
// the product path.  Built by oracle/Makefile into oracle/_build/libninterp_oracle.so.
#include "ninterp_oracle.hpp"

#include <chrono>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

template <class T>
nio::Data<T> make_data(int ngrid, const size_t* axis_len, int nshape, const size_t* shape,
                       const T* const* axes, const T* values) {
    nio::Data<T> d;
    d.ndim = nshape;
    for (int i = 0; i < ngrid; ++i) d.grid.push_back(nio::View1<T>{axes ? axes[i] : nullptr, axis_len[i]});
    for (int i = 0; i < nshape; ++i) d.shape.push_back(shape[i]);
    d.values = values;
    d.finish();
    return d;
}

template <class T>
int eval_batch(int kind, int ndim, const size_t* axis_len, const size_t* shape, const T* const* axes,
               const T* values, int strategy, int extrap, T fill, const T* const* coords, size_t n,
               T* out, uint8_t* status, int nthreads) {
    nio::Data<T> d = make_data<T>(ndim, axis_len, ndim, shape, axes, values);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long long i = 0; i < (long long)n; ++i) {
        T pt[32];
        for (int k = 0; k < ndim; ++k) pt[k] = coords[k][i];
        T v;
        uint8_t st = nio::interpolate<T>(kind, d, strategy, extrap, fill, pt, &v);
        out[i] = v;
        if (status) status[i] = st;
    }
    return 0;
}

template <class T>
int validate_all(int kind, int ngrid, const size_t* axis_len, int nshape, const size_t* shape,
                 const T* const* axes, int strategy, int extrap, int* dim) {
    // Order of `InterpXD::new`: data.validate() (one/mod.rs:24-31) then check_extrapolate (:121)
    nio::Data<T> d = make_data<T>(ngrid, axis_len, nshape, shape, axes, nullptr);
    int rc = nio::validate_data<T>(kind, d, dim);
    if (rc != nio::OK) return rc;
    return nio::check_extrapolate<T>(d, strategy, extrap, extrap, dim);
}

}  // namespace

extern "C" {

int nio_eval_f64(int kind, int ndim, const size_t* axis_len, const size_t* shape,
                 const double* const* axes, const double* values, int strategy, int extrap,
                 double fill, const double* const* coords, size_t n, double* out, uint8_t* status,
                 int nthreads) {
    return eval_batch<double>(kind, ndim, axis_len, shape, axes, values, strategy, extrap, fill,
                              coords, n, out, status, nthreads);
}

int nio_eval_f32(int kind, int ndim, const size_t* axis_len, const size_t* shape,
                 const float* const* axes, const float* values, int strategy, int extrap, float fill,
                 const float* const* coords, size_t n, float* out, uint8_t* status, int nthreads) {
    return eval_batch<float>(kind, ndim, axis_len, shape, axes, values, strategy, extrap, fill,
                             coords, n, out, status, nthreads);
}

int nio_validate_f64(int kind, int ngrid, const size_t* axis_len, int nshape, const size_t* shape,
                     const double* const* axes, int strategy, int extrap, int* dim) {
    return validate_all<double>(kind, ngrid, axis_len, nshape, shape, axes, strategy, extrap, dim);
}

int nio_validate_f32(int kind, int ngrid, const size_t* axis_len, int nshape, const size_t* shape,
                     const float* const* axes, int strategy, int extrap, int* dim) {
    return validate_all<float>(kind, ngrid, axis_len, nshape, shape, axes, strategy, extrap, dim);
}

// set_extrapolate()'s check (interpolator/mod.rs:83-107 with current != requested)
int nio_check_extrapolate(int ndim, const size_t* axis_len, int strategy, int current, int requested,
                          int* dim) {
    nio::Data<double> d;
    for (int i = 0; i < ndim; ++i) d.grid.push_back(nio::View1<double>{nullptr, axis_len[i]});
    return nio::check_extrapolate<double>(d, strategy, current, requested, dim);
}

double nio_wrap_f64(double x, double lo, double hi) { return nio::wrap<double>(x, lo, hi); }
float nio_wrap_f32(float x, float lo, float hi) { return nio::wrap<float>(x, lo, hi); }

// returns -1 where the reference would panic
long long nio_find_nearest_index_f64(const double* arr, size_t len, double target) {
    try {
        return (long long)nio::find_nearest_index<double>(nio::View1<double>{arr, len}, target);
    } catch (const nio::Panic&) {
        return -1;
    }
}
long long nio_find_nearest_index_f32(const float* arr, size_t len, float target) {
    try {
        return (long long)nio::find_nearest_index<float>(nio::View1<float>{arr, len}, target);
    } catch (const nio::Panic&) {
        return -1;
    }
}

int nio_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
