// acc_device.cuh -- device-side maintenance of the exact running sums of a region pool (layout: acc_layout.h, arithmetic:
// acc_window.h).  Used by the iteration kernels (iter_exact.cu): the decide kernel streams over the regions the last rule
// launch wrote (RegionHeap.add, RegionHeap.java:30-36) and over the (val, err) pairs of the regions the last selection popped
// (RegionHeap.poll, :21-28).  The rule kernels themselves do not take part: their loop body fills the instruction cache and
// they run at four warps per scheduler, where every additional instruction costs its full latency (measured: +14 % run
// time for 400 integer instructions per region, profiles/r02_acc_cost.md); a streaming pass over 16 bytes per region costs
// 2 - 3 %.
//
// Every thread owns one 256-bit window accumulator per sum in shared memory (limb-major: consecutive threads touch
// consecutive 8-byte words) and adds a value with one two-limb carry chain: no atomics, no shuffles.  When the CTA is done
// the windows are summed as 32-bit cells and added to one replica of the delta cells in global memory.  No floating-point
// addition is involved anywhere, so the result does not depend on how the regions are spread over lanes, warps, CTAs or GPUs.
#ifndef CUB_ACC_DEVICE_CUH
#define CUB_ACC_DEVICE_CUH

#include "acc_layout.h"
#include "acc_window.h"

template <int T>
struct CubAccWin {
    unsigned long long w[2][XA_WLIMBS][T];  // [0] error sum, [1] value sum: thread t owns w[.][.][t]
    unsigned long long key[3][T];           // per thread: smallest errmax, largest errmax, largest |val| (bits)
    int cnt[T];                             // regions added minus removed
    unsigned long long far[2][XA_LIMBS];    // values outside the window: signed cells of weight 2^(32 j) (shared atomics)
    unsigned long long red[T / 32][20];     // per-warp partial results of the final reduction
    int e_base, v_base, far_used;
};

__device__ __forceinline__ unsigned long long cub_errmax_bits(double err) {
    return (unsigned long long)__double_as_longlong(err > 0.0 ? err : 0.0);
}
__device__ __forceinline__ unsigned long long cub_abs_bits(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffULL;
    return b >= 0x7ff0000000000000ULL ? 0ULL : b;  // non-finite values give no hint
}

// every thread of the CTA calls these (barriers inside); blockDim.x == T
template <int T>
__device__ __forceinline__ void cub_acc_begin(const unsigned long long* acc, CubAccWin<T>* win) {
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < XA_WLIMBS; ++k) win->w[0][k][t] = win->w[1][k][t] = 0ULL;
    win->key[0][t] = ~0ULL;
    win->key[1][t] = 0ULL;
    win->key[2][t] = 0ULL;
    win->cnt[t] = 0;
    for (int q = t; q < 2 * XA_LIMBS; q += T) (&win->far[0][0])[q] = 0ULL;
    if (t == 0) {  // window hints of the iteration kernel (0 before the first iteration)
        win->e_base = (int)acc[XA_M_OFF + XA_M_EBASE];
        win->v_base = (int)acc[XA_M_OFF + XA_M_VBASE];
        win->far_used = 0;
    }
    __syncthreads();
}

// the rare cases, out of line: a non-finite value (counted, never summed), or a finite one outside the window (far cells)
template <int T>
__device__ __noinline__ void cub_acc_cold(CubAccWin<T>* win, unsigned long long* acc, int kind, unsigned long long b, int sgn) {
    const int e = (int)((b >> 52) & 2047ULL);
    const unsigned long long fr = b & 0xfffffffffffffULL;
    if (e == 2047) {
        unsigned long long* misc = acc + XA_M_OFF;
        if (kind == 0) {
            atomicAdd(&misc[fr ? XA_M_ENAN : XA_M_EINF], (unsigned long long)(long long)sgn);
            atomicOr(&misc[XA_M_ESTICKY], sgn > 0 ? 1ULL : 2ULL);
        } else {
            atomicAdd(&misc[fr ? XA_M_VNAN : ((b >> 63) ? XA_M_VNINF : XA_M_VPINF)], (unsigned long long)(long long)sgn);
            atomicOr(&misc[XA_M_VSTICKY], sgn > 0 ? 1ULL : 2ULL);
        }
        return;
    }
    const bool neg = (sgn < 0) != ((b >> 63) != 0ULL);
    int j;
    long long c[3];
    xaw_far_cells(e ? (fr | (1ULL << 52)) : fr, e > 0 ? e - 1 : 0, neg, &j, c);
    unsigned long long* far = win->far[kind];
    atomicAdd(&far[j], (unsigned long long)c[0]);
    if (c[1]) atomicAdd(&far[j + 1], (unsigned long long)c[1]);
    if (c[2]) atomicAdd(&far[j + 2], (unsigned long long)c[2]);
    win->far_used = 1;
}
// sum KIND += sgn * x.  KIND 0: x is an error estimate (an absolute value; a negative one counts as errmax = 0,
// Rule.java:281-289), KIND 1: a (signed) integral estimate.  The hot path is straight-line: decode, shift, two-limb add.
template <int T, int KIND>
__device__ __forceinline__ void cub_acc_term(CubAccWin<T>* win, unsigned long long* acc, unsigned long long* w, int base, double x, int sgn) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    const unsigned hi32 = (unsigned)(b >> 32);
    const int e = (int)((hi32 >> 20) & 2047u);
    const unsigned long long m = (b & 0xfffffffffffffULL) | (e ? (1ULL << 52) : 0ULL);
    const bool minus = (hi32 >> 31) != 0u;
    const int sh = (e > 0 ? e - 1 : 0) - base;
    if ((unsigned)sh > (unsigned)XA_WIN_MAX_SHIFT || e == 2047 || (KIND == 0 && minus)) {  // rare
        if ((m != 0ULL || e == 2047) && !(KIND == 0 && minus && e != 2047)) cub_acc_cold(win, acc, KIND, b, sgn);
        return;
    }
    const int q = sh >> 6, r = sh & 63;
    const unsigned long long alo = m << r, ahi = (m >> 1) >> (63 - r);
    unsigned long long* p = w + q * T;
    unsigned long long lo = p[0], hi = p[T];
    const bool neg = KIND == 0 ? sgn < 0 : ((sgn < 0) != minus);
    unsigned long long c = xaw_addsub2(lo, hi, alo, ahi, neg);
    p[0] = lo;
    p[T] = hi;
    if (c) {  // a carry / borrow beyond the two limbs: rare (the running sum changes sign, or a limb fills up)
        for (int k = q + 2; c && k < XA_WLIMBS; ++k) {
            const unsigned long long v = w[k * T];
            w[k * T] = neg ? v - 1ULL : v + 1ULL;
            c = (neg ? v == 0ULL : v == ~0ULL) ? 1ULL : 0ULL;
        }
    }
}

// One region enters (sgn = +1) or leaves (-1) the sums.  Called by the threads that hold a region (no warp-level
// cooperation).  keys: the caller's running {smallest errmax, largest errmax, largest |val|} bits of the regions added.
struct CubAccKeys {
    unsigned long long mn, ex, vx;
    int cnt;
};
__device__ __forceinline__ void cub_acc_keys_init(CubAccKeys& k) {
    k.mn = ~0ULL;
    k.ex = 0ULL;
    k.vx = 0ULL;
    k.cnt = 0;
}
template <int T>
__device__ __forceinline__ void cub_acc_region(CubAccWin<T>* win, unsigned long long* acc, CubAccKeys& k, double val, double err, int sgn) {
    const int t = threadIdx.x;
    cub_acc_term<T, 0>(win, acc, &win->w[0][0][t], win->e_base, err, sgn);
    cub_acc_term<T, 1>(win, acc, &win->w[1][0][t], win->v_base, val, sgn);
    k.cnt += sgn;
    if (sgn > 0) {
        const unsigned long long kb = cub_errmax_bits(err), vb = cub_abs_bits(val);
        k.mn = kb < k.mn ? kb : k.mn;
        k.ex = kb > k.ex ? kb : k.ex;
        k.vx = vb > k.vx ? vb : k.vx;
    }
}
// before cub_acc_end: the thread's keys go to shared memory
template <int T>
__device__ __forceinline__ void cub_acc_keys_store(CubAccWin<T>* win, const CubAccKeys& k) {
    const int t = threadIdx.x;
    win->key[0][t] = k.mn;
    win->key[1][t] = k.ex;
    win->key[2][t] = k.vx;
    win->cnt[t] = k.cnt;
}

template <int T>
__device__ __forceinline__ void cub_acc_end(unsigned long long* acc, CubAccWin<T>* win) {
    __syncthreads();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    {   // cell c = t % 16 of the threads t / 16, t / 16 + T / 16, ...
        const int c = t & 15, kind = c >> 3, h = c & 7;
        long long s = 0;
        for (int j = t >> 4; j < T; j += T / 16) s += xaw_cell(win->w[kind][h >> 1][j], h);
        s += __shfl_down_sync(0xffffffffu, s, 16);
        if (lane < 16) win->red[warp][c] = (unsigned long long)s;
    }
    {
        unsigned long long mn = win->key[0][t], ex = win->key[1][t], vx = win->key[2][t];
        int cn = win->cnt[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, ex, o),
                                     c = __shfl_xor_sync(0xffffffffu, vx, o);
            mn = a < mn ? a : mn;
            ex = b > ex ? b : ex;
            vx = c > vx ? c : vx;
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
        }
        if (lane == 0) {
            win->red[warp][16] = mn;
            win->red[warp][17] = ex;
            win->red[warp][18] = vx;
            win->red[warp][19] = (unsigned long long)(long long)cn;
        }
    }
    __syncthreads();
    if (t < 20) {
        unsigned long long r = win->red[0][t];
        for (int w = 1; w < T / 32; ++w) {
            const unsigned long long x = win->red[w][t];
            if (t == 16)
                r = x < r ? x : r;
            else if (t == 17 || t == 18)
                r = x > r ? x : r;
            else
                r += x;
        }
        unsigned long long* cells = acc + XA_D_OFF + (unsigned long long)(blockIdx.x % XA_REPL) * XA_CELLS;
        unsigned long long* misc = acc + XA_M_OFF;
        if (t < 16) {
            if (r) atomicAdd(&cells[t], r);
        } else if (t == 19) {
            if (r) atomicAdd(&cells[XA_C_COUNT], r);
        } else if (t == 16) {
            if (r != ~0ULL) atomicMin(&misc[XA_M_MINKEY], r);
        } else if (t == 17) {
            if (r) {
                atomicMax(&misc[XA_M_EMAXKEY], r);
                atomicMax(&misc[XA_M_TOPKEY], r);
            }
        } else if (r) {
            atomicMax(&misc[XA_M_VMAXKEY], r);
        }
    }
    if (win->far_used) {
        for (int q = t; q < 2 * XA_LIMBS; q += T) {
            const unsigned long long v = (&win->far[0][0])[q];
            if (v) atomicAdd(&acc[XA_FAR_E_OFF + q], v);
        }
    }
}

#endif
