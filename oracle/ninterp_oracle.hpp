// ninterp_oracle.hpp — CPU restatement of ninterp 0.8.1's `Interpolator::interpolate`.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ninterp_b200/, include/) may
// include, link or call this file.  It exists so that tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / `--impl reference` legs have a checker and a CPU baseline.
//
// Parity status: PINNED BY THE REFERENCE'S OWN GOLDEN VECTORS.  The Rust crate cannot be
// compiled in this image (no cargo/rustc), so there is no oracle/_ref; instead every exact
// `assert_eq!` / `assert_approx_eq!` vector held by the reference's unit tests, doctests and
// examples for this path is transcribed in tests/kat_vectors.py and this file must reproduce
// all of them (tests/test_oracle_kat.py).  Behaviour not covered by any reference test
// (NaN/inf coordinates, duplicate nodes, length-1 axes, f32, partial-dimension Clamp/Wrap) is
// derived from reading the code cited below and is cross-checked against an independently
// written pure-Python model (tests/pyref.py).
//
// Every operation is a single IEEE-754 round-to-nearest op in T; build with
// `-ffp-contract=off -fno-fast-math` (Rust never contracts a*b+c into an FMA).
//
// Where the reference would panic (index out of bounds, usize underflow, slice out of range)
// this restatement throws `Panic`, reported per point as status 0xFF.
//
// Citations are relative to /root/reference/src/.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace nio {

enum Strategy : int { LINEAR = 0, NEAREST = 1, LEFT_NEAREST = 2, RIGHT_NEAREST = 3 };
enum Extrap : int { ENABLE = 0, FILL = 1, CLAMP = 2, WRAP = 3, ERROR = 4 };
enum Kind : int { KIND_FIXED = 0, KIND_ND = 1 };  // Interp1D/2D/3D vs InterpND
enum Status : uint8_t { ST_OK = 0, ST_EXTRAPOLATE_ERROR = 1, ST_PANIC = 0xFF };

// validate()/check_extrapolate() outcomes, numbered like include/ninterp_cuda.h
enum Err : int {
    OK = 0,
    E_EXTRAPOLATE_SELECTION = -1,  // error.rs:17
    E_EMPTY_GRID = -2,             // error.rs:19
    E_MONOTONICITY = -3,           // error.rs:21
    E_INCOMPATIBLE_SHAPES = -4,    // error.rs:23
    E_OTHER = -5,                  // error.rs:25
};

struct Panic {};

// Bounds-checked 1-D view (ndarray indexing panics when out of range).
template <class T>
struct View1 {
    const T* p;
    size_t len;
    T operator[](size_t i) const {
        if (i >= len) throw Panic{};
        return p[i];
    }
    T first() const { return (*this)[0]; }
    T last() const { return (*this)[len - 1]; }
};

// usize subtraction: debug builds panic on underflow, release builds wrap and the very next
// index panics; both end in a panic, so model it as one.
inline size_t checked_sub(size_t a, size_t b) {
    if (a < b) throw Panic{};
    return a - b;
}

// strategy/traits.rs:10-33
template <class T>
size_t find_nearest_index(View1<T> arr, T target) {
    if (target == arr.last()) {              // :11
        return checked_sub(arr.len, 2);      // :12
    }
    size_t low = 0;                          // :15
    size_t high = arr.len - 1;               // :16
    while (low < high) {                     // :18
        size_t mid = low + (high - low) / 2; // :19
        if (arr[mid] >= target) {            // :21
            high = mid;
        } else {
            low = mid + 1;
        }
    }
    if (low > 0 && arr[low] >= target) {     // :28
        return low - 1;
    }
    return low;
}

// f32/f64 `rem_euclid` as reached through num_traits::Euclid (lib.rs:82): Rust std
// `let r = self % rhs; if r < 0.0 { r + rhs.abs() } else { r }`; `%` on floats is C fmod.
template <class T>
T rem_euclid(T a, T b) {
    T r = std::fmod(a, b);
    return (r < T(0)) ? r + std::fabs(b) : r;
}

// lib.rs:81-83
template <class T>
T wrap(T input, T min, T max) {
    return min + rem_euclid<T>(input - min, max - min);
}

// num_traits::clamp (call sites one/mod.rs:183, three/mod.rs:207, n/mod.rs:304)
template <class T>
T clamp(T input, T min, T max) {
    if (input < min) return min;
    if (input > max) return max;
    return input;
}

// A rectilinear grid + C-order value table (data.rs:31-44, n/mod.rs:26-35).
template <class T>
struct Data {
    int ndim = 0;
    std::vector<View1<T>> grid;
    std::vector<size_t> shape;    // values.shape()
    std::vector<size_t> stride;   // C-order element strides
    const T* values = nullptr;
    size_t values_len = 1;

    void finish() {
        stride.assign(shape.size(), 1);
        values_len = 1;
        for (int d = (int)shape.size() - 1; d >= 0; --d) {
            stride[d] = values_len;
            values_len *= shape[d];
        }
    }
    // bounds-checked multi-index load
    T at(const size_t* idx) const {
        size_t off = 0;
        for (size_t d = 0; d < shape.size(); ++d) {
            if (idx[d] >= shape[d]) throw Panic{};
            off += idx[d] * stride[d];
        }
        return values[off];
    }
};

// data.rs:69-89 (fixed) and n/mod.rs:71-110 (N-D); `dim` receives the offending axis.
template <class T>
int validate_data(int kind, const Data<T>& d, int* dim) {
    *dim = -1;
    size_t n = d.shape.size();
    if (kind == KIND_ND) {
        // n/mod.rs:104-110
        size_t nd = (d.values_len == 1) ? 0 : d.shape.size();
        bool all_empty = true;
        for (auto& g : d.grid) all_empty = all_empty && (g.len == 0);
        if (d.grid.size() != nd && !(nd == 0 && all_empty)) return E_OTHER;  // n/mod.rs:76-83
        n = nd;
    }
    for (size_t i = 0; i < n; ++i) {
        size_t len = d.grid[i].len;
        if (len == 0) { *dim = (int)i; return E_EMPTY_GRID; }                 // data.rs:76
        for (size_t k = 0; k + 1 < len; ++k) {
            if (!(d.grid[i].p[k] <= d.grid[i].p[k + 1])) {                    // data.rs:80
                *dim = (int)i;
                return E_MONOTONICITY;
            }
        }
        if (len != d.shape[i]) { *dim = (int)i; return E_INCOMPATIBLE_SHAPES; } // data.rs:84
    }
    return OK;
}

inline bool allow_extrapolate(int strategy) { return strategy == LINEAR; }  // one/strategies.rs:32,61,84,107

// interpolator/mod.rs:83-107.  `current` is self.extrapolate, `requested` the argument.
template <class T>
int check_extrapolate(const Data<T>& d, int strategy, int current, int requested, int* dim) {
    *dim = -1;
    if (requested == ENABLE && !allow_extrapolate(strategy)) return E_EXTRAPOLATE_SELECTION; // :88
    if (current == ENABLE) {                                                                // :97 (sic)
        for (size_t i = 0; i < d.grid.size(); ++i) {
            if (d.grid[i].len < 2) { *dim = (int)i; return E_OTHER; }                       // :99
        }
    }
    return OK;
}

// ---------------------------------------------------------------------------------------
// Strategies
// ---------------------------------------------------------------------------------------

// Linear's lower-index choice: one/strategies.rs:19-25, two:16-24, three:16-24, n:61-67
template <class T>
size_t linear_lower(View1<T> g, T p) {
    if (p < g.first()) return 0;
    if (p > g.last()) return checked_sub(g.len, 2);
    return find_nearest_index<T>(g, p);
}

// `iter().position(|&x| x == p)` — one/strategies.rs:14, n/strategies.rs:38-41
template <class T>
bool position(View1<T> g, T p, size_t* pos) {
    for (size_t i = 0; i < g.len; ++i) {
        if (g.p[i] == p) { *pos = i; return true; }
    }
    return false;
}

// one/strategies.rs (all four strategies)
template <class T>
T strategy_1d(const Data<T>& d, int strategy, const T* pt) {
    const View1<T>& g = d.grid[0];
    T p = pt[0];
    size_t i;
    if (position<T>(g, p, &i)) { size_t ix[1] = {i}; return d.at(ix); }  // :14-16 / :47 / :76 / :99
    switch (strategy) {
    case LINEAR: {
        size_t l = linear_lower<T>(g, p);                               // :19-25
        size_t u = l + 1;
        T diff = (p - g[l]) / (g[u] - g[l]);                            // :27
        size_t il[1] = {l}, iu[1] = {u};
        return d.at(il) * (T(1) - diff) + d.at(iu) * diff;              // :28
    }
    case NEAREST: {
        size_t l = find_nearest_index<T>(g, p);                         // :50
        size_t u = l + 1;
        size_t k = (p - g[l] < g[u] - p) ? l : u;                       // :52-56
        size_t ix[1] = {k};
        return d.at(ix);
    }
    case LEFT_NEAREST: {
        size_t ix[1] = {find_nearest_index<T>(g, p)};                   // :79
        return d.at(ix);
    }
    default: {  // RIGHT_NEAREST
        size_t ix[1] = {find_nearest_index<T>(g, p) + 1};               // :102
        return d.at(ix);
    }
    }
}

// two/strategies.rs
template <class T>
T strategy_2d(const Data<T>& d, int strategy, const T* pt) {
    if (strategy == LINEAR) {
        size_t lo[2];
        for (int k = 0; k < 2; ++k) lo[k] = linear_lower<T>(d.grid[k], pt[k]);            // :16-24
        size_t xl = lo[0], xu = xl + 1;
        T xd = (pt[0] - d.grid[0][xl]) / (d.grid[0][xu] - d.grid[0][xl]);                 // :28
        size_t yl = lo[1], yu = yl + 1;
        T yd = (pt[1] - d.grid[1][yl]) / (d.grid[1][yu] - d.grid[1][yl]);                 // :32
        size_t a[2] = {xl, yl}, b[2] = {xu, yl}, c[2] = {xl, yu}, e[2] = {xu, yu};
        T f0 = d.at(a) * (T(1) - xd) + d.at(b) * xd;                                      // :34-35
        T f1 = d.at(c) * (T(1) - xd) + d.at(e) * xd;                                      // :36-37
        return f0 * (T(1) - yd) + f1 * yd;                                                // :39
    }
    // Nearest :53-76
    size_t ix[2];
    for (int k = 0; k < 2; ++k) {
        size_t l = find_nearest_index<T>(d.grid[k], pt[k]);
        size_t u = l + 1;
        ix[k] = (pt[k] - d.grid[k][l] < d.grid[k][u] - pt[k]) ? l : u;
    }
    return d.at(ix);
}

// three/strategies.rs
template <class T>
T strategy_3d(const Data<T>& d, int strategy, const T* pt) {
    if (strategy == LINEAR) {
        size_t lo[3];
        for (int k = 0; k < 3; ++k) lo[k] = linear_lower<T>(d.grid[k], pt[k]);            // :16-24
        size_t xl = lo[0], xu = xl + 1;
        T xd = (pt[0] - d.grid[0][xl]) / (d.grid[0][xu] - d.grid[0][xl]);                 // :27
        size_t yl = lo[1], yu = yl + 1;
        T yd = (pt[1] - d.grid[1][yl]) / (d.grid[1][yu] - d.grid[1][yl]);                 // :31
        size_t zl = lo[2], zu = zl + 1;
        T zd = (pt[2] - d.grid[2][zl]) / (d.grid[2][zu] - d.grid[2][zl]);                 // :35
        size_t i000[3] = {xl, yl, zl}, i100[3] = {xu, yl, zl};
        size_t i001[3] = {xl, yl, zu}, i101[3] = {xu, yl, zu};
        size_t i010[3] = {xl, yu, zl}, i110[3] = {xu, yu, zl};
        size_t i011[3] = {xl, yu, zu}, i111[3] = {xu, yu, zu};
        T f00 = d.at(i000) * (T(1) - xd) + d.at(i100) * xd;                               // :37-38
        T f01 = d.at(i001) * (T(1) - xd) + d.at(i101) * xd;                               // :39-40
        T f10 = d.at(i010) * (T(1) - xd) + d.at(i110) * xd;                               // :41-42
        T f11 = d.at(i011) * (T(1) - xd) + d.at(i111) * xd;                               // :43-44
        T f0 = f00 * (T(1) - yd) + f10 * yd;                                              // :46
        T f1 = f01 * (T(1) - yd) + f11 * yd;                                              // :47
        return f0 * (T(1) - zd) + f1 * zd;                                                // :49
    }
    // Nearest :63-94
    size_t ix[3];
    for (int k = 0; k < 3; ++k) {
        size_t l = find_nearest_index<T>(d.grid[k], pt[k]);
        size_t u = l + 1;
        ix[k] = (pt[k] - d.grid[k][l] < d.grid[k][u] - pt[k]) ? l : u;
    }
    return d.at(ix);
}

// n/strategies.rs:22-108 (Linear) and :121-202 (Nearest).  Written with per-call vectors on
// purpose: this is where the reference spends its time, and the CPU baseline should too.
template <class T>
T strategy_nd(const Data<T>& d, int strategy, const T* pt_in) {
    size_t n = d.shape.size();
    std::vector<T> point(pt_in, pt_in + n);                                   // :33
    std::vector<View1<T>> grid(d.grid.begin(), d.grid.end());                 // :34
    // values_view: per original axis either a fixed index or "still an axis"
    std::vector<size_t> fixed(n, 0);
    std::vector<char> is_fixed(n, 0);
    std::vector<size_t> active;                                               // original axis ids, ascending
    for (size_t k = 0; k < n; ++k) active.push_back(k);
    for (size_t dim = n; dim-- > 0;) {                                        // :36 (reversed)
        size_t pos;
        if (position<T>(grid[dim], point[dim], &pos)) {                       // :38-41
            point.erase(point.begin() + dim);                                 // :43
            grid.erase(grid.begin() + dim);                                   // :44
            fixed[active[dim]] = pos;                                         // :45 index_axis_inplace
            is_fixed[active[dim]] = 1;
            active.erase(active.begin() + dim);
        }
    }
    size_t view_len = 1;
    for (size_t a : active) view_len *= d.shape[a];
    std::vector<size_t> idx(n, 0);
    for (size_t k = 0; k < n; ++k) idx[k] = is_fixed[k] ? fixed[k] : 0;
    if (view_len == 1) {                                                      // :47-50
        return d.at(idx.data());
    }
    n = active.size();                                                        // :52

    std::vector<size_t> lower(n);
    std::vector<T> diffs(n);
    std::vector<char> lower_closer(n);
    for (size_t k = 0; k < n; ++k) {
        if (strategy == LINEAR) {
            size_t l = linear_lower<T>(grid[k], point[k]);                    // :61-67
            diffs[k] = (point[k] - grid[k][l]) / (grid[k][l + 1] - grid[k][l]); // :68-69
            lower[k] = l;
        } else {
            size_t l = find_nearest_index<T>(grid[k], point[k]);              // :158
            lower_closer[k] = (point[k] - grid[k][l] < grid[k][l + 1] - point[k]); // :159-160
            lower[k] = l;
        }
    }
    // :76-81 slice_each_axis(lower..=lower+1).to_owned(): a (2,2,...,2) cube, C-order
    size_t ncorner = size_t(1) << n;
    std::vector<T> vals(ncorner);
    for (size_t c = 0; c < ncorner; ++c) {
        for (size_t k = 0; k < n; ++k) {
            size_t bit = (c >> (n - 1 - k)) & 1;  // axis 0 is the slowest-varying bit
            idx[active[k]] = lower[k] + bit;
        }
        vals[c] = d.at(idx.data());               // slice out of range -> panic
    }
    // :86-104 reduce axis 0 first: new[rest] = vals[0,rest] (op) vals[1,rest]
    for (size_t k = 0; k < n; ++k) {
        size_t half = vals.size() / 2;
        std::vector<T> next(half);
        for (size_t i = 0; i < half; ++i) {
            if (strategy == LINEAR) {
                next[i] = vals[i] * (T(1) - diffs[k]) + vals[half + i] * diffs[k];  // :101-102
            } else {
                next[i] = lower_closer[k] ? vals[i] : vals[half + i];             // :191-195
            }
        }
        vals.swap(next);
    }
    return vals[0];                                                           // :107
}

template <class T>
T run_strategy(int kind, const Data<T>& d, int strategy, const T* pt) {
    if (kind == KIND_ND) return strategy_nd<T>(d, strategy, pt);
    switch (d.shape.size()) {
    case 1: return strategy_1d<T>(d, strategy, pt);
    case 2: return strategy_2d<T>(d, strategy, pt);
    default: return strategy_3d<T>(d, strategy, pt);
    }
}

// ---------------------------------------------------------------------------------------
// Front-ends: one/mod.rs:172-207, two/mod.rs:181-226, three/mod.rs:193-238, n/mod.rs:286-340
// Returns the per-point status; *out receives the value (quiet NaN when status != OK).
// ---------------------------------------------------------------------------------------
template <class T>
uint8_t interpolate(int kind, const Data<T>& d, int strategy, int extrap, T fill, const T* pt, T* out) {
    const size_t n = (kind == KIND_ND && d.values_len == 1) ? 0 : d.shape.size();
    *out = std::numeric_limits<T>::quiet_NaN();
    if (kind == KIND_ND && d.values_len == 1) {
        // 0-D InterpND (n/mod.rs:104-110): validate() only admits empty grids here, the
        // collapse loop of n/strategies.rs:36-46 never runs its closure, and :47-50 returns
        // the single element.
        *out = d.values[0];
        return ST_OK;
    }
    try {
        bool any_error = false;
        for (size_t dim = 0; dim < n; ++dim) {
            T first = d.grid[dim].first(), last = d.grid[dim].last();
            if (!(first <= pt[dim] && pt[dim] <= last)) {        // !(first..=last).contains(p)
                switch (extrap) {
                case ENABLE: break;
                case FILL: *out = fill; return ST_OK;
                case CLAMP: {
                    std::vector<T> q(n);
                    for (size_t i = 0; i < n; ++i)
                        q[i] = clamp<T>(pt[i], d.grid[i].first(), d.grid[i].last());
                    *out = run_strategy<T>(kind, d, strategy, q.data());
                    return ST_OK;
                }
                case WRAP: {
                    std::vector<T> q(n);
                    for (size_t i = 0; i < n; ++i)
                        q[i] = wrap<T>(pt[i], d.grid[i].first(), d.grid[i].last());
                    *out = run_strategy<T>(kind, d, strategy, q.data());
                    return ST_OK;
                }
                default: any_error = true; break;                 // Error: collect (1-D returns at once)
                }
            }
        }
        if (any_error) return ST_EXTRAPOLATE_ERROR;
        *out = run_strategy<T>(kind, d, strategy, pt);
        return ST_OK;
    } catch (const Panic&) {
        *out = std::numeric_limits<T>::quiet_NaN();
        return ST_PANIC;
    }
}

}  // namespace nio
