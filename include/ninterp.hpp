// ninterp.hpp — header-only C++ mirror of the reference's interpolator API over the C ABI (ninterp_cuda.h).
//
// Same names and argument meaning as ninterp 0.8.1 (paths relative to the reference's src/):
//   Interp1D/2D/3D::new    interpolator/{one,two,three}/mod.rs      -> constructors
//   InterpND::new          interpolator/n/mod.rs:225-239
//   Interpolator trait     interpolator/mod.rs:25-34                 -> ndim / validate / interpolate / set_extrapolate
//   Extrapolate<T>         interpolator/mod.rs:57-72
//   ValidateError, InterpolateError   error.rs:9-27, 38-45           -> exceptions with the same variants
// plus the batch entry points the engine adds.  Nothing here computes an interpolated value.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "ninterp_cuda.h"

namespace ninterp {

namespace strategy {
enum Strategy : int { Linear = NINT_LINEAR, Nearest = NINT_NEAREST, LeftNearest = NINT_LEFT_NEAREST, RightNearest = NINT_RIGHT_NEAREST };
}

template <class T>
struct Extrapolate {
    int code;
    T value;
    static Extrapolate Enable() { return {NINT_EXTRAPOLATE_ENABLE, T(0)}; }
    static Extrapolate Fill(T v) { return {NINT_EXTRAPOLATE_FILL, v}; }
    static Extrapolate Clamp() { return {NINT_EXTRAPOLATE_CLAMP, T(0)}; }
    static Extrapolate Wrap() { return {NINT_EXTRAPOLATE_WRAP, T(0)}; }
    static Extrapolate Error() { return {NINT_EXTRAPOLATE_ERROR, T(0)}; }
};

struct ValidateError : std::runtime_error {
    int code, dim;  // code: NINT_E_EXTRAPOLATE_SELECTION .. NINT_E_OTHER
    ValidateError(int c, int d, const std::string& m) : std::runtime_error(m), code(c), dim(d) {}
};
struct InterpolateError : std::runtime_error {
    int code;  // NINT_E_POINT_LENGTH or NINT_E_EXTRAPOLATE
    InterpolateError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct ReferencePanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct EngineError : std::runtime_error {
    int code;
    EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void check(int rc) {
    if (rc == NINT_OK) return;
    const std::string msg = nint_last_error_message();
    if (rc <= NINT_E_EXTRAPOLATE_SELECTION && rc >= NINT_E_OTHER) throw ValidateError(rc, nint_last_error_dim(), msg);
    if (rc == NINT_E_POINT_LENGTH || rc == NINT_E_EXTRAPOLATE) throw InterpolateError(rc, msg);
    if (rc == NINT_E_PANIC) throw ReferencePanic(msg);
    throw EngineError(rc, msg);
}

// Common implementation: the `Interpolator<T>` trait.
template <class T>
class Interpolator {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "f32 or f64");

public:
    Interpolator(int kind, const std::vector<std::vector<T>>& grid, const std::vector<T>& values,
                 const std::vector<size_t>& shape, strategy::Strategy s, Extrapolate<T> e, int device = 0) {
        std::vector<size_t> lens;
        std::vector<const void*> ptrs;
        for (auto& g : grid) {
            lens.push_back(g.size());
            ptrs.push_back(g.empty() ? nullptr : g.data());
        }
        check(nint_create(device, kind, std::is_same<T, double>::value ? NINT_F64 : NINT_F32, (int)grid.size(), lens.data(),
                          ptrs.data(), (int)shape.size(), shape.data(), values.data(), nullptr, (int)s, e.code, &e.value, &h_));
    }
    ~Interpolator() { nint_destroy(h_); }
    Interpolator(const Interpolator&) = delete;
    Interpolator& operator=(const Interpolator&) = delete;

    int ndim() const { return nint_ndim(h_); }
    void validate() { check(nint_validate(h_)); }
    // Interpolator::interpolate(&self, point: &[T]) -> Result<T, InterpolateError>
    T interpolate(const std::vector<T>& point) const {
        T out;
        check(nint_interpolate(h_, point.data(), point.size(), &out));
        return out;
    }
    void set_extrapolate(Extrapolate<T> e) { check(nint_set_extrapolate(h_, e.code, &e.value)); }
    void set_strategy(strategy::Strategy s) { check(nint_set_strategy(h_, (int)s)); }

    // Batch over host points stored as [T; ndim] each (what `&[[T; N]]` is); throws like the scalar call on
    // the first failing point.
    std::vector<T> interpolate_batch(const std::vector<T>& points_aos) const {
        const size_t nd = (size_t)ndim();
        const size_t n = nd ? points_aos.size() / nd : 0;
        std::vector<T> out(n);
        nint_batch_info info;
        check(nint_interpolate_batch_host(h_, NINT_LAYOUT_AOS, points_aos.data(), n, out.data(), nullptr, &info));
        return out;
    }
    // Same, never throwing for per-point failures: status bytes and summary are returned instead.
    std::vector<T> interpolate_batch(const std::vector<T>& points_aos, std::vector<uint8_t>& status, nint_batch_info& info) const {
        const size_t nd = (size_t)ndim();
        const size_t n = nd ? points_aos.size() / nd : 0;
        std::vector<T> out(n);
        status.assign(n, 0);
        const int rc = nint_interpolate_batch_host(h_, NINT_LAYOUT_AOS, points_aos.data(), n, out.data(), status.data(), &info);
        if (rc != NINT_OK && rc != NINT_E_EXTRAPOLATE && rc != NINT_E_PANIC) check(rc);
        return out;
    }
    nint_handle* raw() const { return h_; }

private:
    nint_handle* h_ = nullptr;
};

template <class T>
struct Interp1D : Interpolator<T> {
    Interp1D(const std::vector<T>& x, const std::vector<T>& f_x, strategy::Strategy s, Extrapolate<T> e, int device = 0)
        : Interpolator<T>(NINT_KIND_FIXED, {x}, f_x, {f_x.size()}, s, e, device) {}
};
template <class T>
struct Interp2D : Interpolator<T> {
    // f_xy in C order: f_xy[i * y.size() + j]
    Interp2D(const std::vector<T>& x, const std::vector<T>& y, const std::vector<T>& f_xy, strategy::Strategy s,
             Extrapolate<T> e, int device = 0)
        : Interpolator<T>(NINT_KIND_FIXED, {x, y}, f_xy, {x.size(), y.size()}, s, e, device) {}
};
template <class T>
struct Interp3D : Interpolator<T> {
    Interp3D(const std::vector<T>& x, const std::vector<T>& y, const std::vector<T>& z, const std::vector<T>& f_xyz,
             strategy::Strategy s, Extrapolate<T> e, int device = 0)
        : Interpolator<T>(NINT_KIND_FIXED, {x, y, z}, f_xyz, {x.size(), y.size(), z.size()}, s, e, device) {}
};
template <class T>
struct InterpND : Interpolator<T> {
    InterpND(const std::vector<std::vector<T>>& grid, const std::vector<T>& values, const std::vector<size_t>& shape,
             strategy::Strategy s, Extrapolate<T> e, int device = 0)
        : Interpolator<T>(NINT_KIND_ND, grid, values, shape, s, e, device) {}
};

}  // namespace ninterp
