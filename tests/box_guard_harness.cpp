// Host build of integrands.h for tests/test_box_guard.py: the region-level guard of the rational Genz families
// (box_safe) against the per-argument range condition of the in-range reciprocal (detmath.h: dm_rcp_flag).
#include <cstdint>
#include <cstring>
#include "../cubature_b200/csrc/integrands.h"

// the condition dm_rcp_flag records in `bad` on the device (zero, subnormal, huge, inf, nan are out of range)
static bool arg_in_range(double x) {
    uint64_t u;
    std::memcpy(&u, &x, 8);
    const int32_t hx = (int32_t)(u >> 32);
    const int32_t lo = (int32_t)((uint32_t)hx + 0x300402u);
    return !(((uint32_t)lo & 0x7fffffffu) < 0x00400402u);
}

template <int DIM>
static double arg_corner(const double* a, const double* x) {
    return CubPow<DIM + 1>::of(cub_dot2<DIM>(a, x, 1.0));
}
template <int DIM>
static double arg_product(const GenzProductPeak<DIM, 0>& g, const double* x, double* d0_out, double* d1_out) {
    double d0 = 1.0, d1 = 1.0;
    for (int i = 0; i < DIM; i += 2) {
        double t = x[i] - g.u[i];
        d0 = d0 * dm_fma(t, t, g.c[i]);
    }
    for (int i = 1; i < DIM; i += 2) {
        double t = x[i] - g.u[i];
        d1 = d1 * dm_fma(t, t, g.c[i]);
    }
    *d0_out = d0;
    *d1_out = d1;
    return d0 * d1;
}

// the argument the exponential families hand to dm_exp / dm_exp_inrange
template <int DIM>
static double arg_gaussian(const GenzGaussian<DIM, 1>& g, const double* x) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = 0; i < DIM; i += 2) {
        double t = g.a[i] * (x[i] - g.u[i]);
        s0 = dm_fma(t, t, s0);
    }
    for (int i = 1; i < DIM; i += 2) {
        double t = g.a[i] * (x[i] - g.u[i]);
        s1 = dm_fma(t, t, s1);
    }
    return -(s0 + s1);
}
static bool exp_arg_in_range(double a) { return a >= -700.0 && a <= 700.0; }  // false for NaN

// bits of dm_exp(x) and dm_exp_inrange(x)
extern "C" void bg_exp_pair(double x, double* out) {
    out[0] = dm_exp(x);
    out[1] = dm_exp_inrange(x);
}

// family 3 (corner peak) / 2 (product peak) / 4 (Gaussian) / 5 (C0) / 6 (discontinuous) / 10 (isotropic Gaussian), DIM = 4: returns box_safe; *all_in_range = every sample point x[n][4] (inside
// the box) has its reciprocal argument -- and for the product peak both chain products -- in range
extern "C" int bg_check(int family, const double* params, const double* c, const double* h, const double* x, int n, int* all_in_range) {
    constexpr int DIM = 4;
    bool ok = true, safe;
    if (family == 3) {
        GenzCornerPeak<DIM, 0> g(params, 1);
        safe = g.box_safe(c, h);
        for (int k = 0; k < n; ++k) ok = ok && arg_in_range(arg_corner<DIM>(g.a, x + k * DIM));
    } else if (family == 4) {
        GenzGaussian<DIM, 1> g(params, 1);
        safe = g.box_safe(c, h);
        for (int k = 0; k < n; ++k) ok = ok && exp_arg_in_range(arg_gaussian<DIM>(g, x + k * DIM));
    } else if (family == 5) {
        GenzC0<DIM, 1> g(params, 1);
        safe = g.box_safe(c, h);
        for (int k = 0; k < n; ++k) {
            double s = 0.0;
            for (int i = 0; i < DIM; ++i) s = dm_fma(g.a[i], dm_abs(x[k * DIM + i] - g.u[i]), s);
            ok = ok && exp_arg_in_range(-s);
        }
    } else if (family == 6) {
        GenzDiscontinuous<DIM, 1> g(params, 1);
        safe = g.box_safe(c, h);
        for (int k = 0; k < n; ++k) {
            double s = 0.0;
            for (int i = 0; i < DIM; ++i) s = dm_fma(g.a[i], x[k * DIM + i], s);
            ok = ok && exp_arg_in_range(s);
        }
    } else if (family == 10) {
        GaussIso<DIM, 1> g(params, 1);
        safe = g.box_safe(c, h);
        for (int k = 0; k < n; ++k) {
            double s = 0.0;
            for (int i = 0; i < DIM; ++i) s = s + x[k * DIM + i] * x[k * DIM + i];
            ok = ok && exp_arg_in_range(-g.sigma * s);
        }
    } else {
        GenzProductPeak<DIM, 0> g(params, 1);
        safe = g.box_safe(c, h);
        for (int k = 0; k < n; ++k) {
            double d0, d1;
            const double p = arg_product<DIM>(g, x + k * DIM, &d0, &d1);
            ok = ok && arg_in_range(p) && arg_in_range(d0) && arg_in_range(d1);
        }
    }
    *all_in_range = ok ? 1 : 0;
    return safe ? 1 : 0;
}
