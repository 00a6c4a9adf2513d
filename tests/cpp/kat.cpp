// kat.cpp — the reference's own tests, re-read through the C++ mirror (include/ninterp.hpp) on the GPU.
// Each block cites the reference test it transcribes.  Exit code 0 = all passed.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ninterp.hpp"

using namespace ninterp;
static int failures = 0;
#define EXPECT(cond)                                                   \
    do {                                                               \
        if (!(cond)) {                                                 \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                \
        }                                                              \
    } while (0)
static bool same_bits(double a, double b) { return std::memcmp(&a, &b, 8) == 0; }
#define EXPECT_EQ(a, b) EXPECT(same_bits((a), (b)))
#define EXPECT_APPROX(a, b) EXPECT(std::fabs((a) - (b)) <= 1e-6)

int main() {
    using E = Extrapolate<double>;
    const std::vector<double> x5{0., 1., 2., 3., 4.}, f5{0.2, 0.4, 0.6, 0.8, 1.0};
    {  // one/tests.rs:4-17 test_invalid_args
        Interp1D<double> interp(x5, f5, strategy::Linear, E::Error());
        try {
            interp.interpolate({});
            EXPECT(false);
        } catch (const InterpolateError& e) {
            EXPECT(e.code == NINT_E_POINT_LENGTH);
        }
        EXPECT_EQ(interp.interpolate({1.0}), 0.4);
    }
    {  // one/tests.rs:20-36 test_linear
        Interp1D<double> interp(x5, f5, strategy::Linear, E::Error());
        for (size_t i = 0; i < x5.size(); ++i) EXPECT_EQ(interp.interpolate({x5[i]}), f5[i]);
        EXPECT_EQ(interp.interpolate({3.00}), 0.8);
        EXPECT_EQ(interp.interpolate({3.75}), 0.95);
        EXPECT_EQ(interp.interpolate({4.00}), 1.0);
    }
    {  // one/tests.rs:39-98 left / right / nearest
        Interp1D<double> l(x5, f5, strategy::LeftNearest, E::Error()), r(x5, f5, strategy::RightNearest, E::Error()),
            n(x5, f5, strategy::Nearest, E::Error());
        EXPECT_EQ(l.interpolate({3.75}), 0.8);
        EXPECT_EQ(r.interpolate({3.25}), 1.0);
        EXPECT_EQ(n.interpolate({3.25}), 0.8);
        EXPECT_EQ(n.interpolate({3.50}), 1.0);
    }
    {  // one/tests.rs:101-174 extrapolate inputs / fill / clamp / enable
        try {
            Interp1D<double> bad(x5, f5, strategy::Nearest, E::Enable());
            EXPECT(false);
        } catch (const ValidateError& e) {
            EXPECT(e.code == NINT_E_EXTRAPOLATE_SELECTION);
        }
        Interp1D<double> err(x5, f5, strategy::Linear, E::Error());
        try {
            err.interpolate({-1.});
            EXPECT(false);
        } catch (const InterpolateError& e) {
            EXPECT(e.code == NINT_E_EXTRAPOLATE);
        }
        Interp1D<double> fill(x5, f5, strategy::Linear, E::Fill(NAN));
        EXPECT_EQ(fill.interpolate({1.5}), 0.5);
        EXPECT(std::isnan(fill.interpolate({5.})));
        Interp1D<double> clamp(x5, f5, strategy::Linear, E::Clamp());
        EXPECT_EQ(clamp.interpolate({-1.}), 0.2);
        EXPECT_EQ(clamp.interpolate({5.}), 1.0);
        Interp1D<double> en(x5, f5, strategy::Linear, E::Enable());
        EXPECT_EQ(en.interpolate({-1.}), 0.0);
        EXPECT_APPROX(en.interpolate({-0.75}), 0.05);
        EXPECT_EQ(en.interpolate({5.}), 1.2);
    }
    const std::vector<double> x3{0.05, 0.10, 0.15}, y3{0.10, 0.20, 0.30}, z3{0.20, 0.40, 0.60};
    std::vector<double> f33(9), f333(27);
    for (int i = 0; i < 9; ++i) f33[i] = i;
    for (int i = 0; i < 27; ++i) f333[i] = i;
    {  // two/tests.rs:4-24 test_linear, :66-94 test_nearest
        Interp2D<double> interp(x3, y3, f33, strategy::Linear, E::Error());
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) EXPECT_EQ(interp.interpolate({x3[i], y3[j]}), f33[i * 3 + j]);
        EXPECT_EQ(interp.interpolate({0.075, 0.25}), 3.);
        Interp2D<double> near(x3, y3, f33, strategy::Nearest, E::Error());
        EXPECT_EQ(near.interpolate({0.07, 0.15 + 0.0001}), 1.);  // "float imprecision"
        EXPECT_EQ(near.interpolate({0.14, 0.29}), 8.);
    }
    {  // three/tests.rs:4-38 test_linear, :41-79 test_linear_extrapolation, n/tests.rs:110-212 N-D == 3-D
        Interp3D<double> interp(x3, y3, z3, f333, strategy::Linear, E::Error());
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) EXPECT_EQ(interp.interpolate({x3[i], y3[j], z3[k]}), f333[(i * 3 + j) * 3 + k]);
        EXPECT_APPROX(interp.interpolate({0.075, 0.15, 0.20}), 6.);
        Interp3D<double> en(x3, y3, z3, f333, strategy::Linear, E::Enable());
        InterpND<double> nd({x3, y3, z3}, f333, {3, 3, 3}, strategy::Linear, E::Enable());
        EXPECT_APPROX(en.interpolate({0.01, 0.06, 0.17}), -8.55);
        EXPECT_APPROX(en.interpolate({0.19, 0.36, 0.65}), 35.25);
        const double pts[4][3] = {{0.01, 0.06, 0.17}, {0.02, 0.36, 0.19}, {0.17, 0.06, 0.63}, {0.19, 0.36, 0.65}};
        for (auto& p : pts) EXPECT_EQ(en.interpolate({p[0], p[1], p[2]}), nd.interpolate({p[0], p[1], p[2]}));
    }
    {  // three/mod.rs:98-127 doctest
        std::vector<double> v{0.6, 0.8, 1.0, 1.2, 0.8, 1.0, 1.2, 1.4, 1.0, 1.2, 1.4, 1.6,
                              0.8, 1.0, 1.2, 1.4, 1.0, 1.2, 1.4, 1.6, 1.2, 1.4, 1.6, 1.8};
        Interp3D<double> interp({1., 2.}, {1., 2., 3.}, {1., 2., 3., 4.}, v, strategy::Linear, E::Error());
        EXPECT_EQ(interp.interpolate({1.5, 1.5, 1.5}), 0.9);
        try {
            interp.interpolate({5.5, 5.5, 5.5});
            EXPECT(false);
        } catch (const InterpolateError& e) {
            EXPECT(e.code == NINT_E_EXTRAPOLATE);
        }
        // the batch call maps the caller's loop: same values, and the same error for a bad point
        std::vector<double> out = interp.interpolate_batch({1.5, 1.5, 1.5, 1., 1., 1., 2., 3., 4.});
        EXPECT(out.size() == 3);
        EXPECT_EQ(out[0], 0.9);
        EXPECT_EQ(out[1], 0.6);
        EXPECT_EQ(out[2], 1.8);
        std::vector<uint8_t> st;
        nint_batch_info info;
        out = interp.interpolate_batch({1.5, 1.5, 1.5, 5.5, 5.5, 5.5}, st, info);
        EXPECT(st[0] == NINT_STATUS_OK && st[1] == NINT_STATUS_EXTRAPOLATE_ERROR && info.n_error == 1 && info.first_error == 1);
    }
    {  // n/tests.rs:322-346 test_extrapolate_wrap, :349-381 test_mismatched_grid
        std::vector<double> u{0., 1.}, f{0., 1., 2., 3., 4., 5., 6., 7.};
        InterpND<double> w({u, u, u}, f, {2, 2, 2}, strategy::Linear, E::Wrap());
        EXPECT_EQ(w.interpolate({-0.25, -0.2, -0.4}), w.interpolate({0.75, 0.8, 0.6}));
        EXPECT_EQ(w.interpolate({2.5, 2.1, 2.3}), w.interpolate({0.5, 0.1, 0.3}));
        try {
            InterpND<double> bad({u, u, u}, {0., 1., 2., 3.}, {2, 2}, strategy::Linear, E::Error());
            EXPECT(false);
        } catch (const ValidateError& e) {
            EXPECT(e.code == NINT_E_OTHER);
        }
        InterpND<double> zero({{}}, {0.5}, {1}, strategy::Linear, E::Error());
        EXPECT(zero.ndim() == 0);
        EXPECT_EQ(zero.interpolate({}), 0.5);
    }
    std::printf(failures ? "%d FAILURES\n" : "all reference KATs passed through include/ninterp.hpp%.0d\n", failures);
    return failures ? 1 : 0;
}
