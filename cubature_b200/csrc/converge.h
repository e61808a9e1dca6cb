// converge.h -- the reference's convergence predicate, Rule.converged (Rule.java:162-279), as one
// host+device template so that the host-driven (HEAP_EXACT) loop and the device-side selection
// kernels evaluate literally the same expression.
//
// Mirrored quirks: INDIVIDUAL / PAIRED are satisfied as soon as ANY component
// (or pair) meets EITHER tolerance (`converged |= ...`, Rule.java:172-213); PAIRED scales the values
// by the larger SIGNED value (Rule.java:187-196); strict `<` everywhere; a NaN tolerance disables
// that test.  Not mirrored: the stray printf of the L2 branch (Rule.java:264).
//
// Extension (NOT reference behaviour): norm | CUB_NORM_STRICT (8) makes INDIVIDUAL
// and PAIRED require EVERY component (pair) to meet a tolerance -- what README.md:112-122 documents and
// what upstream C cubature implements -- instead of the reference's ANY.
//
// Acc is any type with `double val(int j) const` and `double err(int j) const`.
#ifndef CUB_CONVERGE_H
#define CUB_CONVERGE_H

#if defined(__CUDACC__)
#define CUB_HD __host__ __device__ __forceinline__
#else
#define CUB_HD inline
#endif

CUB_HD bool cub_isnan(double x) { return x != x; }
CUB_HD double cub_fabs(double x) { return x < 0 ? -x : (x == 0 ? 0.0 : x); }  // -0.0 -> +0.0 like Math.abs
CUB_HD double cub_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return sqrt(x);
#else
    return __builtin_sqrt(x);
#endif
}

template <class Acc>
CUB_HD bool cub_converged(int fdim, const Acc& ee, double absTol, double relTol, int norm) {
    const bool useAbs = !cub_isnan(absTol), useRel = !cub_isnan(relTol);
    const bool strict = (norm & 8) != 0;
    bool conv = false;
    int j;
    if (strict && (norm & 7) <= 1) {
        conv = true;
        if ((norm & 7) == 0) {
            for (j = 0; j < fdim; ++j) {
                const double e = ee.err(j);
                bool cj = false;
                if (useAbs) cj |= e < absTol;
                if (useRel) cj |= e < cub_fabs(ee.val(j)) * relTol;
                conv &= cj;
            }
        } else {
            for (j = 0; j + 1 < fdim; j += 2) {
                const double e0 = ee.err(j), e1 = ee.err(j + 1), v0 = ee.val(j), v1 = ee.val(j + 1);
                const double maxerr = e0 > e1 ? e0 : e1;
                const double maxval = v0 > v1 ? v0 : v1;
                const double serr = maxerr > 0 ? 1 / maxerr : 1;
                const double sval = maxval > 0 ? 1 / maxval : 1;
                const double err = cub_sqrt((e0 * serr) * (e0 * serr) + (e1 * serr) * (e1 * serr)) * maxerr;
                const double val = cub_sqrt((v0 * sval) * (v0 * sval) + (v1 * sval) * (v1 * sval)) * maxval;
                bool cj = false;
                if (useAbs) cj |= err < absTol;
                if (useRel) cj |= err < cub_fabs(val) * relTol;
                conv &= cj;
            }
            if (j < fdim) {
                const double e = ee.err(j);
                bool cj = false;
                if (useAbs) cj |= e < absTol;
                if (useRel) cj |= e < cub_fabs(ee.val(j)) * relTol;
                conv &= cj;
            }
        }
        return conv;
    }
    switch (norm & 7) {
        case 0:  // INDIVIDUAL
            for (j = 0; j < fdim; ++j) {
                const double e = ee.err(j);
                if (useAbs) conv |= e < absTol;
                if (useRel) conv |= e < cub_fabs(ee.val(j)) * relTol;
            }
            break;
        case 1:  // PAIRED
            for (j = 0; j + 1 < fdim; j += 2) {
                const double e0 = ee.err(j), e1 = ee.err(j + 1), v0 = ee.val(j), v1 = ee.val(j + 1);
                const double maxerr = e0 > e1 ? e0 : e1;
                const double maxval = v0 > v1 ? v0 : v1;
                const double serr = maxerr > 0 ? 1 / maxerr : 1;
                const double sval = maxval > 0 ? 1 / maxval : 1;
                const double err = cub_sqrt((e0 * serr) * (e0 * serr) + (e1 * serr) * (e1 * serr)) * maxerr;
                const double val = cub_sqrt((v0 * sval) * (v0 * sval) + (v1 * sval) * (v1 * sval)) * maxval;
                if (useAbs) conv |= err < absTol;
                if (useRel) conv |= err < cub_fabs(val) * relTol;
            }
            if (j < fdim) {
                const double e = ee.err(j);
                if (useAbs) conv |= e < absTol;
                if (useRel) conv |= e < cub_fabs(ee.val(j)) * relTol;
            }
            break;
        case 3: {  // L1
            double err = 0, val = 0;
            for (j = 0; j < fdim; ++j) {
                err += ee.err(j);
                val += cub_fabs(ee.val(j));
            }
            if (useAbs) conv |= err < absTol;
            if (useRel) conv |= err < val * relTol;
        } break;
        case 4: {  // LINF
            double err = 0, val = 0;
            for (j = 0; j < fdim; ++j) {
                const double absval = cub_fabs(ee.val(j));
                const double e = ee.err(j);
                if (e > err) err = e;
                if (absval > val) val = absval;
            }
            if (useAbs) conv |= err < absTol;
            if (useRel) conv |= err < val * relTol;
        } break;
        case 2: {  // L2
            double maxerr = 0, maxval = 0, err = 0, val = 0;
            for (j = 0; j < fdim; ++j) {
                const double absval = cub_fabs(ee.val(j));
                const double e = ee.err(j);
                if (e > maxerr) maxerr = e;
                if (absval > maxval) maxval = absval;
            }
            const double serr = maxerr > 0 ? 1 / maxerr : 1;
            const double sval = maxval > 0 ? 1 / maxval : 1;
            for (j = 0; j < fdim; ++j) {
                const double e = ee.err(j) * serr;
                const double v = cub_fabs(ee.val(j)) * sval;
                err += e * e;
                val += v * v;
            }
            err = cub_sqrt(err) * maxerr;
            val = cub_sqrt(val) * maxval;
            if (useAbs) conv |= err < absTol;
            if (useRel) conv |= err < val * relTol;
        } break;
        default:
            conv = false;
    }
    return conv;
}

#endif
