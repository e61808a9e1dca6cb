// integrands.h -- built-in integrand families (the "device-side plugin" of BASELINE.json).
//
// Each family is a small functor template<int DIM>: the constructor derives per-thread constants
// from the parameter vector, eval() computes all fdim components at one point x[DIM].
// The header compiles unchanged with g++ (CPU oracle), nvcc and NVRTC; all transcendental math
// goes through detmath.h so host and device produce identical bits (compile without contraction).
// Every loop runs over the dimension in ascending order: that order is part of the definition.
//
// Families
//   1..6   Genz test family on [0,1]^d; params = a[0..d) then u[0..d)
//   10     isotropic Gaussian exp(-sigma * sum x_i^2); params = {sigma}
//          (the reference's only golden vector: TestCubature.java:435-465 / README.md:224)
//   20     the reference's test integrands 0..8 (TestCubature.java:46-177); params = which[fdim]
//   30     oscillatory vector f_k(x) = cos((k+1) * sum x_i), k < fdim   (BASELINE config C2)
//   31     complex exponentials g_m(x) = exp(-(1 + i w_m) sum x_i), interleaved (re, im)
//          (BASELINE config C4; params = w[fdim/2])
#ifndef CUBATURE_INTEGRANDS_H
#define CUBATURE_INTEGRANDS_H

#include "detmath.h"

#define CUB_FAM_GENZ_OSCILLATORY 1
#define CUB_FAM_GENZ_PRODUCT_PEAK 2
#define CUB_FAM_GENZ_CORNER_PEAK 3
#define CUB_FAM_GENZ_GAUSSIAN 4
#define CUB_FAM_GENZ_C0 5
#define CUB_FAM_GENZ_DISCONTINUOUS 6
#define CUB_FAM_GAUSS_ISO 10
#define CUB_FAM_REFTEST 20
#define CUB_FAM_OSC_VEC 30
#define CUB_FAM_CEXP 31
#define CUB_FAM_USER 1000

#define CUB_TWO_PI 6.283185307179586

// ---- building blocks with a FIXED association (part of each family's definition) -----------------
// Two interleaved accumulation chains (even / odd coordinates) halve the dependent-FMA depth of a
// d-term sum on the GPU; x^N by square-and-multiply needs ~2 log2 N multiplications instead of N-1.
template <int DIM>
DM_HD double cub_dot2(const double* a, const double* x, double init) {
    double s0 = init, s1 = 0.0;
    for (int i = 0; i < DIM; i += 2) s0 = dm_fma(a[i], x[i], s0);
    for (int i = 1; i < DIM; i += 2) s1 = dm_fma(a[i], x[i], s1);
    return DIM > 1 ? s0 + s1 : s0;
}
template <int N>
struct CubPow {  // x^N, N >= 1: p = x^(N/2); N even -> p*p, N odd -> (p*p)*x
    DM_HD static double of(double x) {
        const double p = CubPow<N / 2>::of(x);
        return (N & 1) ? (p * p) * x : p * p;
    }
};
template <>
struct CubPow<1> {
    DM_HD static double of(double x) { return x; }
};

// ---- Genz family --------------------------------------------------------------------------------

template <int DIM>
struct GenzOscillatory {  // cos(2 pi u_0 + sum a_i x_i)
    static const bool kSeparable = false;
    double a[DIM], phase;
    DM_HD GenzOscillatory(const double* p, int) {
        for (int i = 0; i < DIM; ++i) a[i] = p[i];
        phase = CUB_TWO_PI * p[DIM];
    }
    DM_HD void eval(const double* x, double* f) const { f[0] = dm_cos(cub_dot2<DIM>(a, x, phase)); }
};

// The two rational families come in three variants of the final reciprocal (MODE):
//   0  IEEE division (the definition; what the oracle computes)
//   1  the branch-free in-range reciprocal, leaving its exact range is recorded in `bad` (detmath.h: dm_rcp_flag)
//   2  the in-range reciprocal without the check -- only for boxes box_safe() has accepted
// box_safe(c, h): true only if EVERY point of the box c +- h (all rule points lie inside it) gives an argument far
// inside the reciprocal's exact range, intermediate products included; conservative by hundreds of binades, so the
// rounding of the bound itself is irrelevant; any NaN makes it false.  The scalar rule kernel asks once per region and
// runs variant 2 (149 range checks per region become one), else variant 0 (rule_kernels.cuh: cub_gm_scalar).
template <int DIM, int MODE = 1>
struct GenzProductPeak {  // prod 1 / (a_i^-2 + (x_i - u_i)^2), evaluated as 1 / prod(...)
    static const bool kSeparable = false;
    typedef GenzProductPeak<DIM, 0> Exact;      // always IEEE division (rule_kernels.cuh: cub_fast_failed)
    typedef GenzProductPeak<DIM, 2> Unchecked;  // in-range reciprocal, no check (after box_safe)
    double c[DIM], u[DIM];
    mutable int bad;
    DM_HD GenzProductPeak(const double* p, int) {
        for (int i = 0; i < DIM; ++i) {
            c[i] = 1.0 / (p[i] * p[i]);
            u[i] = p[DIM + i];
        }
        bad = 0;
    }
    DM_HD void eval(const double* x, double* f) const {
        double d0 = 1.0, d1 = 1.0;  // two interleaved product chains (even / odd coordinates)
        for (int i = 0; i < DIM; i += 2) {
            double t = x[i] - u[i];
            d0 = d0 * dm_fma(t, t, c[i]);
        }
        for (int i = 1; i < DIM; i += 2) {
            double t = x[i] - u[i];
            d1 = d1 * dm_fma(t, t, c[i]);
        }
        f[0] = MODE == 1 ? dm_rcp_flag(d0 * d1, &bad) : (MODE == 2 ? dm_rcp_inrange(d0 * d1) : 1.0 / (d0 * d1));
    }
    DM_HD bool box_safe(const double* bc, const double* bh) const {
        // factor i lies in [c_i, c_i + (|bc_i - u_i| + bh_i)^2]; each chain's product must stay within 2^+-400
        double lo0 = 1.0, lo1 = 1.0, up0 = 1.0, up1 = 1.0;
        for (int i = 0; i < DIM; i += 2) {
            const double d = dm_abs(bc[i] - u[i]) + dm_abs(bh[i]);
            lo0 = lo0 * c[i];
            up0 = up0 * dm_fma(d, d, c[i]);
        }
        for (int i = 1; i < DIM; i += 2) {
            const double d = dm_abs(bc[i] - u[i]) + dm_abs(bh[i]);
            lo1 = lo1 * c[i];
            up1 = up1 * dm_fma(d, d, c[i]);
        }
        const double tiny = 3.8725919148493183e-121, huge = 2.5822498780869086e+120;  // 2^-400, 2^400
        return lo0 > tiny && lo1 > tiny && up0 < huge && up1 < huge;
    }
};

template <int DIM, int MODE = 1>
struct GenzCornerPeak {  // (1 + sum a_i x_i)^-(d+1)
    static const bool kSeparable = false;
    typedef GenzCornerPeak<DIM, 0> Exact;      // always IEEE division (rule_kernels.cuh: cub_fast_failed)
    typedef GenzCornerPeak<DIM, 2> Unchecked;  // in-range reciprocal, no check (after box_safe)
    double a[DIM];
    mutable int bad;
    DM_HD GenzCornerPeak(const double* p, int) {
        for (int i = 0; i < DIM; ++i) a[i] = p[i];
        bad = 0;
    }
    DM_HD void eval(const double* x, double* f) const {
        const double s = cub_dot2<DIM>(a, x, 1.0);
        const double p = CubPow<DIM + 1>::of(s);  // s^(d+1) by square-and-multiply (fixed association)
        f[0] = MODE == 1 ? dm_rcp_flag(p, &bad) : (MODE == 2 ? dm_rcp_inrange(p) : 1.0 / p);
    }
    DM_HD bool box_safe(const double* bc, const double* bh) const {
        // s = mid +- rad over the box, |terms| <= mag: |s| > 2^-40 mag >= 2^-40 (rounding: ~2^-50 mag) and |s| < mag, so
        // 2^-40(d+1) < |s^(d+1)| < mag^(d+1) < 2^900
        if (DIM + 1 > 22) return false;
        double mid = 1.0, rad = 0.0, mag = 1.0;
        for (int i = 0; i < DIM; ++i) {
            mid = dm_fma(a[i], bc[i], mid);
            rad = dm_fma(dm_abs(a[i]), dm_abs(bh[i]), rad);
            mag = dm_fma(dm_abs(a[i]), dm_abs(bc[i]), mag);
        }
        mag = mag + rad;
        return dm_abs(mid) - rad > 9.094947017729282e-13 * mag && mag < dm_pow2i(900 / (DIM + 1));
    }
};

// The exponential families come in two variants (MODE): 1 dm_exp (the definition), 2 dm_exp_inrange -- the same bits
// without the out-of-range handling, for boxes in which box_safe() has bounded the exponent's argument by 690 in
// magnitude (dm_exp_inrange is exact to 700; any NaN makes box_safe false).  See the rational families above.
#define CUB_EXP_SAFE 690.0
template <int DIM, int MODE = 1>
struct GenzGaussian {  // exp(-sum a_i^2 (x_i - u_i)^2)
    static const bool kSeparable = false;
    typedef GenzGaussian<DIM, 1> Exact;
    typedef GenzGaussian<DIM, 2> Unchecked;
    double a[DIM], u[DIM];
    DM_HD GenzGaussian(const double* p, int) {
        for (int i = 0; i < DIM; ++i) {
            a[i] = p[i];
            u[i] = p[DIM + i];
        }
    }
    DM_HD void eval(const double* x, double* f) const {
        double s0 = 0.0, s1 = 0.0;  // two interleaved chains (even / odd coordinates)
        for (int i = 0; i < DIM; i += 2) {
            double t = a[i] * (x[i] - u[i]);
            s0 = dm_fma(t, t, s0);
        }
        for (int i = 1; i < DIM; i += 2) {
            double t = a[i] * (x[i] - u[i]);
            s1 = dm_fma(t, t, s1);
        }
        f[0] = MODE == 2 ? dm_exp_inrange(-(s0 + s1)) : dm_exp(-(s0 + s1));
    }
    DM_HD bool box_safe(const double* bc, const double* bh) const {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) {
            const double t = dm_abs(a[i]) * (dm_abs(bc[i] - u[i]) + dm_abs(bh[i]));
            s = dm_fma(t, t, s);
        }
        return s < CUB_EXP_SAFE;
    }
};

template <int DIM, int MODE = 1>
struct GenzC0 {  // exp(-sum a_i |x_i - u_i|)
    static const bool kSeparable = false;
    typedef GenzC0<DIM, 1> Exact;
    typedef GenzC0<DIM, 2> Unchecked;
    double a[DIM], u[DIM];
    DM_HD GenzC0(const double* p, int) {
        for (int i = 0; i < DIM; ++i) {
            a[i] = p[i];
            u[i] = p[DIM + i];
        }
    }
    DM_HD void eval(const double* x, double* f) const {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) s = dm_fma(a[i], dm_abs(x[i] - u[i]), s);
        f[0] = MODE == 2 ? dm_exp_inrange(-s) : dm_exp(-s);
    }
    DM_HD bool box_safe(const double* bc, const double* bh) const {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) s = dm_fma(dm_abs(a[i]), dm_abs(bc[i] - u[i]) + dm_abs(bh[i]), s);
        return s < CUB_EXP_SAFE;
    }
};

template <int DIM, int MODE = 1>
struct GenzDiscontinuous {  // 0 if x_0 > u_0 or x_1 > u_1, else exp(sum a_i x_i)
    static const bool kSeparable = false;
    typedef GenzDiscontinuous<DIM, 1> Exact;
    typedef GenzDiscontinuous<DIM, 2> Unchecked;
    double a[DIM], u0, u1;
    DM_HD GenzDiscontinuous(const double* p, int) {
        for (int i = 0; i < DIM; ++i) a[i] = p[i];
        u0 = p[DIM];
        u1 = DIM > 1 ? p[DIM + (DIM > 1 ? 1 : 0)] : 0.0;
    }
    DM_HD void eval(const double* x, double* f) const {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) s = dm_fma(a[i], x[i], s);
        bool outside = x[0] > u0;
        if (DIM > 1) outside = outside || (x[DIM > 1 ? 1 : 0] > u1);
        f[0] = outside ? 0.0 : (MODE == 2 ? dm_exp_inrange(s) : dm_exp(s));
    }
    DM_HD bool box_safe(const double* bc, const double* bh) const {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) s = dm_fma(dm_abs(a[i]), dm_abs(bc[i]) + dm_abs(bh[i]), s);
        return s < CUB_EXP_SAFE;
    }
};

// ---- the reference's golden-vector integrand -----------------------------------------------------
// TestCubature.java:442-455: sum += x*x left to right (Java never fuses), exp(-sigma*sum).
template <int DIM, int MODE = 1>
struct GaussIso {
    static const bool kSeparable = false;
    typedef GaussIso<DIM, 1> Exact;
    typedef GaussIso<DIM, 2> Unchecked;
    double sigma;
    DM_HD GaussIso(const double* p, int) { sigma = p[0]; }
    DM_HD void eval(const double* x, double* f) const {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) s = s + x[i] * x[i];
        f[0] = MODE == 2 ? dm_exp_inrange(-sigma * s) : dm_exp(-sigma * s);
    }
    DM_HD bool box_safe(const double* bc, const double* bh) const {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) {
            const double t = dm_abs(bc[i]) + dm_abs(bh[i]);
            s = dm_fma(t, t, s);
        }
        return dm_abs(sigma) * s < CUB_EXP_SAFE;
    }
};

// ---- the reference's test integrands, TestCubature.java:46-177 ------------------------------------
// Same operation order as the Java source; Math.exp/cos/sin/pow replaced by the detmath versions
// (Math.pow with an integer exponent -> dm_powi; pow(x, 2.0) -> x*x which is what an exact pow gives).
#define CUB_REFTEST_MAX_FDIM 16
template <int DIM>
struct RefTest {
    static const bool kSeparable = false;
    int fdim;
    int which[CUB_REFTEST_MAX_FDIM];
    DM_HD RefTest(const double* p, int fdim_) {
        fdim = fdim_ < CUB_REFTEST_MAX_FDIM ? fdim_ : CUB_REFTEST_MAX_FDIM;
        for (int j = 0; j < CUB_REFTEST_MAX_FDIM; ++j) which[j] = j < fdim ? (int)p[j] : 0;
    }
    DM_HD static double one(const double* x, int w) {
        const double K_2_SQRTPI = 1.12837916709551257390;
        const double radius = 0.50124145262344534123412;
        double val = 0.0;
        switch (w) {
            case 0: {  // prod cos(x_i)                                   TestCubature.java:118-123
                val = 1.0;
                for (int i = 0; i < DIM; ++i) val = val * dm_cos(x[i]);
            } break;
            case 1: {  // exp(-z^2) remapped from (0, inf)                 TestCubature.java:124-139
                double scale = 1.0;
                bool zero = false;
                for (int i = 0; i < DIM; ++i) {
                    if (zero) continue;
                    if (x[i] > 0) {
                        double z = (1 - x[i]) / x[i];
                        val = val + z * z;
                        scale = scale * (K_2_SQRTPI / (x[i] * x[i]));
                    } else {
                        scale = 0;
                        zero = true;
                    }
                }
                val = dm_exp(-val) * scale;
            } break;
            case 2: {  // hypersphere indicator                              TestCubature.java:140-147
                for (int i = 0; i < DIM; ++i) val = val + x[i] * x[i];
                val = val < radius * radius ? 1.0 : 0.0;
            } break;
            case 3: {  // prod 2 x_i                                         TestCubature.java:46-53
                val = 1.0;
                for (int i = 0; i < DIM; ++i) val = val * (2.0 * x[i]);
            } break;
            case 4: {  // Gaussian a = 0.1 centred at 1/2                    TestCubature.java:58-66
                const double a = 0.1;
                double sum = 0.0;
                for (int i = 0; i < DIM; ++i) {
                    double dx = x[i] - 0.5;
                    sum = sum + dx * dx;
                }
                val = dm_powi(K_2_SQRTPI / (2.0 * a), DIM) * dm_exp(-sum / (a * a));
            } break;
            case 5: {  // double Gaussian                                    TestCubature.java:69-81
                const double a = 0.1;
                double sum1 = 0.0, sum2 = 0.0;
                for (int i = 0; i < DIM; ++i) {
                    double dx1 = x[i] - 1.0 / 3.0;
                    double dx2 = x[i] - 2.0 / 3.0;
                    sum1 = sum1 + dx1 * dx1;
                    sum2 = sum2 + dx2 * dx2;
                }
                val = 0.5 * dm_powi(K_2_SQRTPI / (2.0 * a), DIM) *
                      (dm_exp(-sum1 / (a * a)) + dm_exp(-sum2 / (a * a)));
            } break;
            case 6: {  // Tsuda, c = (1 + sqrt(10)) / 9                      TestCubature.java:84-91
                const double c = (1.0 + dm_sqrt(10.0)) / 9.0;
                val = 1.0;
                for (int i = 0; i < DIM; ++i) {
                    double q = (c + 1) / (c + x[i]);
                    val = val * (c / (c + 1) * (q * q));
                }
            } break;
            case 7: {  // Morokoff-Caflisch                                  TestCubature.java:98-106
                double pw = 1.0 / DIM;
                val = dm_powi(1 + pw, DIM);
                for (int i = 0; i < DIM; ++i) val = val * dm_pow(x[i], pw);
            } break;
            case 8: {  // HCubature.jl#4 (dim 3 only)                        TestCubature.java:164-171
                if (DIM == 3) {
                    val = x[0] * 0.2 * (x[DIM > 2 ? 2 : 0] - 0.5) * 0.4 *
                          dm_sin(x[DIM > 1 ? 1 : 0] * 6.283185307179586);
                    val = 1 + val * val;
                }
            } break;
            default:
                val = 0.0;
        }
        return val;
    }
    DM_HD void eval(const double* x, double* f) const {
        for (int j = 0; j < fdim; ++j) f[j] = one(x, which[j]);
    }
};

// ---- vector families ------------------------------------------------------------------------------

template <int DIM>
struct OscVec {  // f_k = cos((k+1) * sum x_i): every component independent -> separable
    static const bool kSeparable = true;
    int fdim;
    DM_HD OscVec(const double*, int fdim_) { fdim = fdim_; }
    DM_HD static double sumx(const double* x) {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) s = s + x[i];
        return s;
    }
    DM_HD double eval1(const double* x, int j) const { return dm_cos((double)(j + 1) * sumx(x)); }
    DM_HD void eval(const double* x, double* f) const {
        double s = sumx(x);
        for (int j = 0; j < fdim; ++j) f[j] = dm_cos((double)(j + 1) * s);
    }
};

template <int DIM>
struct CExp {  // g_m = exp(-(1 + i w_m) s), s = sum x_i ; f[2m] = Re, f[2m+1] = Im
    static const bool kSeparable = false;
    int fdim;
    const double* w;
    DM_HD CExp(const double* p, int fdim_) {
        fdim = fdim_;
        w = p;
    }
    DM_HD void eval(const double* x, double* f) const {
        double s = 0.0;
        for (int i = 0; i < DIM; ++i) s = s + x[i];
        double e = dm_exp(-s);
        for (int m = 0; 2 * m + 1 < fdim; ++m) {
            double ph = w[m] * s;
            f[2 * m] = e * dm_cos(ph);
            f[2 * m + 1] = -(e * dm_sin(ph));
        }
        if (fdim & 1) f[fdim - 1] = e;
    }
};

// ---- infinite-interval variable transforms (README.md:237-272, fused into point generation) -------
// code 0: x = t
// code 1: [a, inf)    x = a + t/(1-t),   t in [0,1),  dx = dt / (1-t)^2
// code 2: (-inf, b]   x = b - t/(1-t),   t in [0,1),  dx = dt / (1-t)^2
// code 3: (-inf, inf) x = t/(1-t^2),     t in (-1,1), dx = (1+t^2)/(1-t^2)^2 dt
// One division per transformed coordinate; returns the Jacobian factor of that coordinate.
#define CUB_TR_NONE 0
#define CUB_TR_SEMI_INF_UP 1
#define CUB_TR_SEMI_INF_DOWN 2
#define CUB_TR_INF 3

DM_HD double cub_transform_coord(int code, double shift, double t, double* x) {
    if (code == CUB_TR_SEMI_INF_UP || code == CUB_TR_SEMI_INF_DOWN) {
        double inv = 1.0 / (1.0 - t);
        double q = t * inv;
        *x = code == CUB_TR_SEMI_INF_UP ? shift + q : shift - q;
        return inv * inv;
    }
    if (code == CUB_TR_INF) {
        double t2 = t * t;
        double inv = 1.0 / (1.0 - t2);
        *x = t * inv;
        return ((1.0 + t2) * inv) * inv;
    }
    *x = t;
    return 1.0;
}

#endif  // CUBATURE_INTEGRANDS_H
