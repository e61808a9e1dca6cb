// exact_acc.h -- arithmetic on the exact running sums of acc_layout.h (host + device, nvcc / g++).
//
// A bin set represents  sum_e (lo[e] + 2^32 hi[e]) * 2^(max(e,1) - 1075)  exactly.  The loop decisions of the reference
// (Rule.converged at the loop top, Rule.java:73, and after every pop, Rule.java:129) compare such a sum with a tolerance;
// XaBig holds it as a 2300-bit fixed-point number (32-bit limbs in 64-bit cells, unit 2^-1074) and converts it to a double
// either truncated toward zero -- then `double < tol` IS the exact comparison `sum < tol` -- or rounded to nearest (the
// reported totals).
#ifndef CUB_EXACT_ACC_H
#define CUB_EXACT_ACC_H

#include "acc_layout.h"
#include "converge.h"

typedef unsigned long long xa_u64;
typedef unsigned __int128 xa_u128;
typedef __int128 xa_i128;

CUB_HD xa_u64 xa_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (xa_u64)__double_as_longlong(x);
#else
    xa_u64 b;
    __builtin_memcpy(&b, &x, 8);
    return b;
#endif
}
CUB_HD double xa_from_bits(xa_u64 b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double x;
    __builtin_memcpy(&x, &b, 8);
    return x;
#endif
}
// bit position (unit 2^-1074) of the least significant mantissa bit of a double with biased exponent e
CUB_HD int xa_lsb_pos(int e) { return e > 0 ? e - 1 : 0; }
// integer mantissa of a finite double given its bits
CUB_HD xa_u64 xa_mantissa(xa_u64 b) {
    const xa_u64 fr = b & 0xfffffffffffffULL;
    return ((b >> 52) & 2047ULL) ? (fr | (1ULL << 52)) : fr;
}

struct XaBig {
    xa_u64 l[XA_LIMBS];  // limb j has weight 2^(32 j); values may exceed 2^32 until normalize()
    int jlo, jhi;        // every limb outside [jlo, jhi] is zero (the loops below only visit this range)
    CUB_HD void clear() {
        for (int j = 0; j < XA_LIMBS; ++j) l[j] = 0ULL;
        jlo = 0;
        jhi = XA_LIMBS - 1;
    }
    // += v * 2^pos, v < 2^96, pos + 96 <= 32 * (XA_LIMBS - 1)
    CUB_HD void add_shifted(xa_u128 v, int pos) {
        const int j = pos >> 5, s = pos & 31;
        const xa_u128 lo = v << s;                         // low 128 bits of v * 2^s (v < 2^96: nothing lost)
        l[j] += (xa_u64)(lo & 0xffffffffULL);
        l[j + 1] += (xa_u64)((lo >> 32) & 0xffffffffULL);
        l[j + 2] += (xa_u64)((lo >> 64) & 0xffffffffULL);
        l[j + 3] += (xa_u64)(lo >> 96);
    }
    CUB_HD void normalize() {
        xa_u64 carry = 0;
        for (int j = jlo; j <= jhi; ++j) {
            const xa_u64 t = l[j] + carry;
            l[j] = t & 0xffffffffULL;
            carry = t >> 32;
        }
    }
    // the functions below need a normalized number
    CUB_HD int top_limb() const {
        for (int j = jhi; j >= jlo; --j)
            if (l[j]) return j;
        return -1;
    }
    CUB_HD bool is_zero() const { return top_limb() < 0; }
    // compare: -1, 0, 1
    CUB_HD int cmp(const XaBig& o) const {
        const int hi = jhi > o.jhi ? jhi : o.jhi, lo = jlo < o.jlo ? jlo : o.jlo;
        for (int j = hi; j >= lo; --j) {
            if (l[j] != o.l[j]) return l[j] < o.l[j] ? -1 : 1;
        }
        return 0;
    }
    // this -= o (requires this >= o, both normalized)
    CUB_HD void sub(const XaBig& o) {
        const int hi = jhi > o.jhi ? jhi : o.jhi, lo = jlo < o.jlo ? jlo : o.jlo;
        long long borrow = 0;
        for (int j = lo; j <= hi; ++j) {
            long long t = (long long)l[j] - (long long)o.l[j] - borrow;
            borrow = 0;
            if (t < 0) {
                t += (1LL << 32);
                borrow = 1;
            }
            l[j] = (xa_u64)t;
        }
        jlo = lo;
        jhi = hi;
    }
    // bits [pos, pos + 64) as an integer, pos >= 0
    CUB_HD xa_u64 window64(int pos) const {
        const int j = pos >> 5, s = pos & 31;
        xa_u128 w = 0;
        if (j < XA_LIMBS) w |= (xa_u128)l[j];
        if (j + 1 < XA_LIMBS) w |= (xa_u128)l[j + 1] << 32;
        if (j + 2 < XA_LIMBS) w |= (xa_u128)l[j + 2] << 64;
        return (xa_u64)(w >> s);
    }
    CUB_HD int top_bit() const {  // position of the most significant set bit, -1 for zero
        const int t = top_limb();
        if (t < 0) return -1;
        int b = 31;
        while (!((l[t] >> b) & 1ULL)) --b;
        return 32 * t + b;
    }
    CUB_HD bool any_below(int pos) const {  // any set bit at a position < pos
        if (pos <= 0) return false;
        const int full = pos >> 5;
        for (int j = jlo; j < full && j <= jhi; ++j)
            if (l[j]) return true;
        if (full < XA_LIMBS && (pos & 31)) return (l[full] & ((1ULL << (pos & 31)) - 1ULL)) != 0;
        return false;
    }
    // value as a double; nearest = 0: truncated toward zero, 1: round to nearest even.  Overflow gives +Inf.
    CUB_HD double to_double(int nearest) const {
        const int t = top_bit();
        if (t < 0) return 0.0;
        if (t <= 52) return xa_from_bits(window64(0));  // subnormal or smallest normals: the bits ARE the encoding
        // normal: mantissa = bits [t-52, t], biased exponent = t - 51
        xa_u64 m = window64(t - 52) & 0x1fffffffffffffULL;
        int e = t - 51;
        if (nearest) {
            const bool half = (window64(t - 53) & 1ULL) != 0, sticky = any_below(t - 53);
            if (half && (sticky || (m & 1ULL))) {
                ++m;
                if (m >> 53) {
                    m >>= 1;
                    ++e;
                }
            }
        }
        if (e >= 2047) return xa_from_bits(0x7ff0000000000000ULL);
        return xa_from_bits(((xa_u64)e << 52) | (m & 0xfffffffffffffULL));
    }
    // += x for a finite double x >= 0
    CUB_HD void add_double(double x) {
        const xa_u64 b = xa_bits(x);
        add_shifted((xa_u128)xa_mantissa(b), xa_lsb_pos((int)((b >> 52) & 2047ULL)));
    }
};

// a += v * 2^pos for a signed v (|v| < 2^63): positive parts go to *pos_part, negative ones to *neg_part
CUB_HD void xa_add_signed(XaBig* pos_part, XaBig* neg_part, long long v, int pos) {
    if (v > 0) pos_part->add_shifted((xa_u128)(xa_u64)v, pos);
    if (v < 0) neg_part->add_shifted((xa_u128)(xa_u64)(-v), pos);
}
// the XA_HALVES delta cells of a window that starts at bit `base` (acc_layout.h): cell j has weight 2^(base + 32 j)
CUB_HD void xa_fold_cells(XaBig* pos_part, XaBig* neg_part, const long long* cell, int base) {
    for (int j = 0; j < XA_HALVES; ++j) xa_add_signed(pos_part, neg_part, cell[j], base + 32 * j);
}

// ---- 128-bit helpers for the levels below the binade (all mantissas share one exponent there) -------------------
// bin words -> integer: lo + 2^32 hi with lo UNSIGNED (every contribution adds a low part in [0, 2^32)) and hi SIGNED
// (acc_device.cuh: a removal is a negative high part)
CUB_HD xa_i128 xa_join_signed(xa_u64 lo, xa_u64 hi) { return (xa_i128)lo + (xa_i128)(long long)hi * ((xa_i128)1 << 32); }
CUB_HD double xa_u128_to_double_approx(xa_u128 x) { return (double)(xa_u64)(x >> 64) * 18446744073709551616.0 + (double)(xa_u64)x; }
CUB_HD int xa_clz64(xa_u64 x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
CUB_HD int xa_top_bit128(xa_u128 x) {  // -1 for zero
    const xa_u64 hi = (xa_u64)(x >> 64), lo = (xa_u64)x;
    if (hi) return 127 - xa_clz64(hi);
    if (lo) return 63 - xa_clz64(lo);
    return -1;
}
// Truncated (toward zero) double of  X = D * 2^pos + L,  0 <= L < 2^pos, for D > 0.  L1 holds bits [pos - 64, pos) of X
// (zero-extended below bit 0); the unit of X is 2^-1074.  Only the top 53 bits of X matter for truncation, and with D >= 1
// they lie in D and the top bits of L.
CUB_HD double xa_trunc_hi(xa_u128 D, xa_u64 L1, int pos) {
    const int t = xa_top_bit128(D);
    const int T = pos + t;  // top bit of X
    if (T <= 52) {          // X < 2^53: the integer is the encoding (subnormals and the first normal binade); pos <= 52
        const xa_u64 L = pos > 0 ? (L1 >> (64 - pos)) : 0ULL;
        return xa_from_bits((((xa_u64)D) << pos) | L);
    }
    xa_u64 mant;
    if (t >= 52) {
        mant = (xa_u64)(D >> (t - 52));
    } else {
        const int k = 52 - t;  // 1..52 more bits from L (pos > k here)
        mant = (((xa_u64)D) << k) | (L1 >> (64 - k));
    }
    const int e = T - 51;
    if (e >= 2047) return xa_from_bits(0x7ff0000000000000ULL);
    return xa_from_bits(((xa_u64)e << 52) | (mant & 0xfffffffffffffULL));
}

#endif
