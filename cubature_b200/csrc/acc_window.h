// acc_window.h -- the per-thread window accumulator of the exact running sums (acc_layout.h): 256-bit two's-complement
// integers, bit 0 at position `base` of the full-range fixed-point number.  Host + device (nvcc, NVRTC, g++: the CPU tests
// drive exactly this code through tests/xa_harness.cpp); no standard headers (NVRTC).
#ifndef CUB_ACC_WINDOW_H
#define CUB_ACC_WINDOW_H

#include "acc_layout.h"

#if defined(__CUDACC__)
#define XAW_FN __host__ __device__ __forceinline__
#else
#define XAW_FN inline
#endif

typedef unsigned long long xaw_u64;

// (lo, hi) +/-= (alo, ahi) with the carry / borrow out of the pair
XAW_FN xaw_u64 xaw_addsub2(xaw_u64& lo, xaw_u64& hi, xaw_u64 alo, xaw_u64 ahi, bool neg) {
    xaw_u64 c;
#if defined(__CUDA_ARCH__)
    if (!neg)
        asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u64 %2, 0, 0;" : "+l"(lo), "+l"(hi), "=l"(c) : "l"(alo), "l"(ahi));
    else
        asm("sub.cc.u64 %0, %0, %3;\n\tsubc.cc.u64 %1, %1, %4;\n\tsubc.u64 %2, 0, 0;" : "+l"(lo), "+l"(hi), "=l"(c) : "l"(alo), "l"(ahi));
    return c & 1ULL;  // subc leaves 0 - borrow
#else
    if (!neg) {
        const xaw_u64 t = lo + alo, c0 = t < alo ? 1ULL : 0ULL;
        const xaw_u64 u = hi + ahi, c1 = u < ahi ? 1ULL : 0ULL;
        const xaw_u64 v = u + c0, c2 = v < c0 ? 1ULL : 0ULL;
        lo = t;
        hi = v;
        c = c1 | c2;
    } else {
        const xaw_u64 b0 = lo < alo ? 1ULL : 0ULL;
        const xaw_u64 t = lo - alo;
        const xaw_u64 b1 = hi < ahi ? 1ULL : 0ULL;
        const xaw_u64 u = hi - ahi;
        const xaw_u64 b2 = u < b0 ? 1ULL : 0ULL;
        lo = t;
        hi = u - b0;
        c = b1 | b2;
    }
    return c;
#endif
}

// window (limb k at w[k * stride], 256-bit two's complement) += (neg ? -1 : 1) * m * 2^(pos - base) for a 53-bit mantissa m.
// Returns false (window untouched) when the value does not fit: its lsb lies below `base` or more than XA_WIN_MAX_SHIFT bits
// above it.  Only the two limbs the shifted mantissa covers are touched; a carry / borrow beyond them is rare (a sign change
// of the running sum, or a limb filling up) and walks up limb by limb.
XAW_FN bool xaw_accumulate(xaw_u64* w, int stride, xaw_u64 m, int pos, int base, bool neg) {
    const int sh = pos - base;
    if ((unsigned)sh > (unsigned)XA_WIN_MAX_SHIFT) return false;
    const int q = sh >> 6, r = sh & 63;                    // q <= 2
    const xaw_u64 alo = m << r, ahi = (m >> 1) >> (63 - r);  // m * 2^r as two limbs (no shift by 64)
    xaw_u64* p = w + q * stride;
    xaw_u64 lo = p[0], hi = p[stride];
    xaw_u64 c = xaw_addsub2(lo, hi, alo, ahi, neg);
    p[0] = lo;
    p[stride] = hi;
    for (int k = q + 2; c && k < XA_WLIMBS; ++k) {
        const xaw_u64 x = w[k * stride];
        if (!neg) {
            w[k * stride] = x + 1ULL;
            c = x == ~0ULL ? 1ULL : 0ULL;
        } else {
            w[k * stride] = x - 1ULL;
            c = x == 0ULL ? 1ULL : 0ULL;
        }
    }
    return true;
}

// 32-bit half h (0 .. XA_HALVES-1) of a window as the signed cell it contributes: the top half carries the sign
// (x = limb h / 2 of the window)
XAW_FN long long xaw_cell(xaw_u64 x, int h) {
    const unsigned part = (h & 1) ? (unsigned)(x >> 32) : (unsigned)x;
    return h == XA_HALVES - 1 ? (long long)(int)part : (long long)part;
}

// a value outside the window as three signed cells of weight 2^(32 j), j = pos / 32 ...
XAW_FN void xaw_far_cells(xaw_u64 m, int pos, bool neg, int* j, long long (&c)[3]) {
    const int s = pos & 31;
    const xaw_u64 lo = m << s, hi = s > 11 ? (m >> (64 - s)) : 0ULL;  // m < 2^53: bits above 63 only when s > 11
    *j = pos >> 5;
    c[0] = (long long)(lo & 0xffffffffULL);
    c[1] = (long long)(lo >> 32);
    c[2] = (long long)hi;
    if (neg) {
        c[0] = -c[0];
        c[1] = -c[1];
        c[2] = -c[2];
    }
}

#endif
