// acc_device.cuh -- device-side maintenance of the exact running sums of a region pool (layout: acc_layout.h).
// Included by the NVRTC-compiled rule kernels (rule_kernels.cuh) and by the nvcc-compiled iteration kernels (iter_exact.cu).
#ifndef CUB_ACC_DEVICE_CUH
#define CUB_ACC_DEVICE_CUH

#include "acc_layout.h"

// ---- exact running sums (acc_layout.h) -----------------------------------------------------------------------------
// Integer accumulators, so the order in which lanes, warps, CTAs or GPUs contribute does not matter.  A warp first adds up
// the lanes whose values share a binade (__match_any_sync + three 18-bit redux adds), its leader lanes then add to a
// shared-memory window of XA_WIN binades, and the CTA flushes the window once at its end; binades outside the window go to
// global memory directly.  sign = +1 adds a region (RegionHeap.add, RegionHeap.java:30-36), -1 removes it (poll, :21-28).
struct CubAccWin {
    unsigned long long e[XA_WIN * 3];
    unsigned long long v[XA_WIN * 2];
    unsigned long long min_key;
    int e_base, v_base;
};
__device__ __forceinline__ void cub_acc_word(unsigned long long* g, unsigned long long* sm, int idx, int base, int stride, int field,
                                             unsigned long long x) {
    const int w = idx - base;
    if (w >= 0 && w < XA_WIN)
        atomicAdd(&sm[w * stride + field], x);
    else
        atomicAdd(&g[(long long)idx * stride + field], x);
}
// Sum over the lanes of `mask` (all of which hold the same binade) of sgn * m, m < 2^53, as a signed 64-bit integer.
__device__ __forceinline__ long long cub_acc_group_sum(unsigned mask, unsigned long long m, int sgn) {
    const int p0 = (int)(m & 0x3ffffULL) * sgn, p1 = (int)((m >> 18) & 0x3ffffULL) * sgn, p2 = (int)(m >> 36) * sgn;
    const long long s0 = __reduce_add_sync(mask, p0), s1 = __reduce_add_sync(mask, p1), s2 = __reduce_add_sync(mask, p2);
    return s0 + s1 * (1LL << 18) + s2 * (1LL << 36);
}
// Called by all 32 lanes of a warp (converged); lanes with valid == false contribute nothing.
__device__ __forceinline__ void cub_acc_warp(unsigned long long* acc, CubAccWin* win, bool valid, double val, double err, int sign) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long* misc = acc + XA_M_OFF;
    {   // error bins: count and mantissa sum per binade of errmax = max(0, err), NaN -> 0 (Rule.java:281-289)
        const unsigned long long b = (unsigned long long)__double_as_longlong(err);
        const int e = (int)((b >> 52) & 2047ULL);
        const unsigned long long fr = b & 0xfffffffffffffULL;
        int bin = e;
        unsigned long long m = e ? (fr | (1ULL << 52)) : fr;
        if (valid && e == 2047) {  // non-finite: ordered like errmax, summed through the counters
            m = 0ULL;
            if (fr) bin = 0;
            atomicAdd(&misc[fr ? XA_M_ENAN : XA_M_EINF], (unsigned long long)(long long)sign);
            atomicOr(&misc[XA_M_ESTICKY], sign > 0 ? 1ULL : 2ULL);
        } else if (b >> 63) {  // cannot happen (err is an absolute value); ordered as errmax = 0
            bin = 0;
            m = 0ULL;
        }
        const unsigned grp = __match_any_sync(0xffffffffu, valid ? bin : 4096 + (int)lane);
        const long long t = cub_acc_group_sum(grp, m, sign);
        if (valid && lane == (unsigned)(__ffs(grp) - 1)) {
            unsigned long long* g = acc + XA_E_OFF;
            cub_acc_word(g, win->e, bin, win->e_base, 3, 0, (unsigned long long)((long long)__popc(grp) * sign));
            if (t) {  // t = lo + 2^32 hi with 0 <= lo < 2^32 and hi signed
                cub_acc_word(g, win->e, bin, win->e_base, 3, 1, (unsigned long long)t & 0xffffffffULL);
                cub_acc_word(g, win->e, bin, win->e_base, 3, 2, (unsigned long long)(t >> 32));
            }
        }
    }
    {   // value bins: signed mantissa sum per binade
        const unsigned long long b = (unsigned long long)__double_as_longlong(val);
        const int e = (int)((b >> 52) & 2047ULL);
        const unsigned long long fr = b & 0xfffffffffffffULL;
        unsigned long long m = e ? (fr | (1ULL << 52)) : fr;
        bool ok = valid;
        if (valid && e == 2047) {
            atomicAdd(&misc[fr ? XA_M_VNAN : ((b >> 63) ? XA_M_VNINF : XA_M_VPINF)], (unsigned long long)(long long)sign);
            atomicOr(&misc[XA_M_VSTICKY], sign > 0 ? 1ULL : 2ULL);
            ok = false;
        }
        if (m == 0ULL) ok = false;
        const unsigned grp = __match_any_sync(0xffffffffu, ok ? e : 4096 + (int)lane);
        const long long t = cub_acc_group_sum(grp, m, (b >> 63) ? -sign : sign);
        if (ok && t && lane == (unsigned)(__ffs(grp) - 1)) {
            unsigned long long* g = acc + XA_V_OFF;
            cub_acc_word(g, win->v, e, win->v_base, 2, 0, (unsigned long long)t & 0xffffffffULL);
            cub_acc_word(g, win->v, e, win->v_base, 2, 1, (unsigned long long)(t >> 32));
        }
    }
}
// every thread of the CTA calls these (barriers inside)
__device__ __forceinline__ void cub_acc_begin(const unsigned long long* acc, CubAccWin* win) {
    for (int q = threadIdx.x; q < XA_WIN * 3; q += blockDim.x) win->e[q] = 0ULL;
    for (int q = threadIdx.x; q < XA_WIN * 2; q += blockDim.x) win->v[q] = 0ULL;
    if (threadIdx.x == 0) {  // window hints of the iteration kernel (0 before the first iteration)
        win->e_base = (int)acc[XA_M_OFF + XA_M_EBASE];
        win->v_base = (int)acc[XA_M_OFF + XA_M_VBASE];
        win->min_key = ~0ULL;
    }
    __syncthreads();
}
// my_min: bits of the smallest errmax this thread added (~0: none)
__device__ __forceinline__ void cub_acc_end(unsigned long long* acc, CubAccWin* win, unsigned long long my_min) {
    if (my_min != ~0ULL) atomicMin(&win->min_key, my_min);
    __syncthreads();
    // (lo, hi) of a window bin stand for lo + 2^32 hi with lo < 2^63 after a CTA's life; carry lo's upper half into hi so that
    // the global low words grow by less than 2^32 per flush
    for (int q = threadIdx.x; q < XA_WIN; q += blockDim.x) {
        const unsigned long long c = win->e[q * 3], lo = win->e[q * 3 + 1], hi = win->e[q * 3 + 2];
        unsigned long long* g = acc + XA_E_OFF + ((long long)win->e_base + q) * 3;
        if (c) atomicAdd(&g[0], c);
        if (lo | hi) {
            if (lo & 0xffffffffULL) atomicAdd(&g[1], lo & 0xffffffffULL);
            const unsigned long long h2 = hi + (lo >> 32);
            if (h2) atomicAdd(&g[2], h2);
        }
        const unsigned long long vlo = win->v[q * 2], vhi = win->v[q * 2 + 1];
        if (vlo | vhi) {
            unsigned long long* gv = acc + XA_V_OFF + ((long long)win->v_base + q) * 2;
            if (vlo & 0xffffffffULL) atomicAdd(&gv[0], vlo & 0xffffffffULL);
            const unsigned long long h2 = vhi + (vlo >> 32);
            if (h2) atomicAdd(&gv[1], h2);
        }
    }
    if (threadIdx.x == 0 && win->min_key != ~0ULL) atomicMin(&acc[XA_M_OFF + XA_M_MINKEY], win->min_key);
}
__device__ __forceinline__ unsigned long long cub_errmax_bits(double err) {
    return (unsigned long long)__double_as_longlong(err > 0.0 ? err : 0.0);
}

#endif
