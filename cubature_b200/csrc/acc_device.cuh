// acc_device.cuh -- device-side maintenance of the exact running sums of a region pool (layout: acc_layout.h, arithmetic:
// acc_window.h).  Included by the NVRTC-compiled rule kernels (rule_kernels.cuh) and by the nvcc-compiled iteration
// kernels (iter_exact.cu).
//
// Every thread owns one 256-bit window accumulator per sum in shared memory (limb-major: consecutive threads touch
// consecutive 8-byte words).  Per region it builds its net contribution -- its child in, and for a left child the parent out
// (RegionHeap.add / poll, RegionHeap.java:21-36) -- in registers and adds it to its window with one carry chain: no atomics,
// no shuffles, about a hundred integer instructions next to the few thousand FP64 instructions of the rule.  When the CTA
// retires the windows are summed as 32-bit cells and added to one replica of the delta cells in global memory.  No
// floating-point addition is involved anywhere, so the result does not depend on how the regions are spread over lanes,
// warps, CTAs or GPUs.
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
// a += sgn * x, kind 0: x is an error estimate (an absolute value; a negative one counts as errmax = 0, Rule.java:281-289),
// kind 1: a (signed) integral estimate.
template <int T>
__device__ __forceinline__ void cub_acc_term(CubAccWin<T>* win, unsigned long long* acc, int kind, xaw_u64 (&a)[XA_WLIMBS], double x, int sgn) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    const int e = (int)((b >> 52) & 2047ULL);
    const unsigned long long fr = b & 0xfffffffffffffULL;
    bool neg = sgn < 0;
    if (b >> 63) {
        if (kind == 0 && e != 2047) return;
        neg = !neg;
    }
    const unsigned long long m = e ? (fr | (1ULL << 52)) : fr;
    if (m == 0ULL) return;
    if (e == 2047 || !xaw_addend(a, m, e > 0 ? e - 1 : 0, kind ? win->v_base : win->e_base, neg)) cub_acc_cold(win, acc, kind, b, sgn);
}

// One region enters the sums (val, err); when has_parent, the parent it replaces (pval, perr) leaves them.  Called by the
// threads that hold a region (no warp-level cooperation).  Out of line: the rule's register allocation is not disturbed.
template <int T>
__device__ __noinline__ void cub_acc_thread(CubAccWin<T>* win, unsigned long long* acc, int has_parent, double pval, double perr, double val,
                                            double err) {
    const int t = threadIdx.x;
#pragma unroll
    for (int kind = 0; kind < 2; ++kind) {
        xaw_u64 a[XA_WLIMBS] = {0ULL, 0ULL, 0ULL, 0ULL};
        cub_acc_term(win, acc, kind, a, kind ? val : err, 1);
        if (has_parent) cub_acc_term(win, acc, kind, a, kind ? pval : perr, -1);
        xaw_u64 s[XA_WLIMBS];
#pragma unroll
        for (int k = 0; k < XA_WLIMBS; ++k) s[k] = win->w[kind][k][t];
        xaw_add4(s, a, 0ULL);
#pragma unroll
        for (int k = 0; k < XA_WLIMBS; ++k) win->w[kind][k][t] = s[k];
    }
    const unsigned long long kb = cub_errmax_bits(err), vb = cub_abs_bits(val);
    if (kb < win->key[0][t]) win->key[0][t] = kb;
    if (kb > win->key[1][t]) win->key[1][t] = kb;
    if (vb > win->key[2][t]) win->key[2][t] = vb;
    if (!has_parent) win->cnt[t] += 1;
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
