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

// a += b + cin over four 64-bit limbs (wraps at 2^256); cin is 0 or 1
XAW_FN void xaw_add4(xaw_u64 (&a)[XA_WLIMBS], const xaw_u64 (&b)[XA_WLIMBS], xaw_u64 cin) {
#if defined(__CUDA_ARCH__)
    // one carry chain: `cin + (2^64 - 1)` raises the carry flag exactly when cin == 1
    asm("{\n\t"
        ".reg .u64 t;\n\t"
        "add.cc.u64 t, %8, 0xffffffffffffffff;\n\t"
        "addc.cc.u64 %0, %0, %4;\n\t"
        "addc.cc.u64 %1, %1, %5;\n\t"
        "addc.cc.u64 %2, %2, %6;\n\t"
        "addc.u64 %3, %3, %7;\n\t"
        "}"
        : "+l"(a[0]), "+l"(a[1]), "+l"(a[2]), "+l"(a[3])
        : "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]), "l"(cin));
#else
    xaw_u64 c = cin;
    for (int k = 0; k < XA_WLIMBS; ++k) {
        const xaw_u64 t = a[k] + b[k];
        const xaw_u64 c1 = t < b[k] ? 1ULL : 0ULL;
        const xaw_u64 u = t + c;
        const xaw_u64 c2 = u < c ? 1ULL : 0ULL;
        a[k] = u;
        c = c1 | c2;
    }
#endif
}

// a += (neg ? -1 : 1) * m * 2^(pos - base) for a 53-bit mantissa m.  Returns false (a untouched) when the value does not
// fit the window: its lsb lies below `base` or more than XA_WIN_MAX_SHIFT bits above it.
XAW_FN bool xaw_addend(xaw_u64 (&a)[XA_WLIMBS], xaw_u64 m, int pos, int base, bool neg) {
    const int sh = pos - base;
    if ((unsigned)sh > (unsigned)XA_WIN_MAX_SHIFT) return false;
    const int q = sh >> 6, r = sh & 63;                  // q <= 2
    const xaw_u64 lo = m << r, hi = (m >> 1) >> (63 - r);  // (m * 2^r) as two limbs; no shift by 64
    const xaw_u64 mask = neg ? ~0ULL : 0ULL;             // -x = ~x + 1
    xaw_u64 b[XA_WLIMBS];
    b[0] = (q == 0 ? lo : 0ULL) ^ mask;
    b[1] = (q == 0 ? hi : (q == 1 ? lo : 0ULL)) ^ mask;
    b[2] = (q == 1 ? hi : (q == 2 ? lo : 0ULL)) ^ mask;
    b[3] = (q == 2 ? hi : 0ULL) ^ mask;
    xaw_add4(a, b, neg ? 1ULL : 0ULL);
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
