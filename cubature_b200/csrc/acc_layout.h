// acc_layout.h -- layout of the exact running sums of a region pool (plain #defines: included by the NVRTC-compiled rule
// kernels, by the nvcc-compiled iteration kernels and by the host).
//
// The reference keeps Sum(val), Sum(err) over its heap as drifting incremental doubles (RegionHeap.java:21-36).  Here they
// are INTEGERS: a double x = m * 2^pos (53-bit integer mantissa m, pos = bit position of its least significant bit in
// units of 2^-1074) is added to a fixed-point number.  Integer addition is associative, so the sums are exact, independent
// of the order in which threads, CTAs or GPUs contribute, and a region that is popped is subtracted again without drift.
//
// The rule kernels do not touch the totals directly.  Every THREAD of a rule launch owns a 256-bit two's-complement
// accumulator per sum (acc_window.h), a window of the full range that starts at bit `base`; the bases are hints the
// iteration kernel derives from the largest values of the previous launch.  A thread adds its NET contribution (its child
// in, the parent it replaces out) with one carry chain, no atomics and no warp traffic.  Values outside the window go to a
// full-range set of 32-bit-weight cells in shared memory (rare).  When a CTA retires, the threads' windows are summed as
// XA_HALVES signed cells of weight 2^(base + 32 j) and added to one of XA_REPL replicas in global memory; the iteration
// kernel folds replicas and far cells into the totals (arbitrary-precision integers) and clears them.
#ifndef CUB_ACC_LAYOUT_H
#define CUB_ACC_LAYOUT_H

#define XA_NBINS 2048   // binary exponents of a double
#define XA_LIMBS 72     // 32-bit limbs of a full-range fixed-point number (2300 bits, unit 2^-1074)
#define XA_WLIMBS 4     // 64-bit limbs of a thread's window accumulator
#define XA_HALVES 8     // 32-bit halves of a window = delta cells per sum
#define XA_WIN_MAX_SHIFT 159  // values whose lsb lies up to 159 bits above the window base fit (top bit <= 211: 44 bits of head room)
#define XA_REPL 64      // replicas of the delta cells
#define XA_CELLS 32     // words per replica: error cells [0, 8), value cells [8, 16), [16] regions added minus removed
#define XA_C_E 0
#define XA_C_V XA_HALVES
#define XA_C_COUNT (2 * XA_HALVES)

// 8-byte words
#define XA_D_OFF 0                            // [XA_REPL][XA_CELLS] signed delta cells
#define XA_M_OFF (XA_REPL * XA_CELLS)         // [32] misc
#define XA_FAR_E_OFF (XA_M_OFF + 32)          // [XA_LIMBS] signed cells of weight 2^(32 j): error values outside the window
#define XA_FAR_V_OFF (XA_FAR_E_OFF + XA_LIMBS)
#define XA_TOT_OFF (XA_FAR_V_OFF + XA_LIMBS)  // totals owned by the iteration kernel: E limbs[72], V magnitude limbs[72], misc[8]
#define XA_TOT_E 0
#define XA_TOT_V XA_LIMBS
#define XA_TOT_MISC (2 * XA_LIMBS)            // [0] V is negative, [1] regions counted, [2] E limb range lo, [3] hi, [4] V lo, [5] hi
#define XA_WORDS (XA_TOT_OFF + 2 * XA_LIMBS + 8)
// misc words
#define XA_M_ENAN 0     // regions currently in the pool whose err is NaN
#define XA_M_EINF 1     //   ... +Inf
#define XA_M_VNAN 2     // regions whose val is NaN
#define XA_M_VPINF 3
#define XA_M_VNINF 4
#define XA_M_ESTICKY 5  // bit 0: a non-finite err was ever added, bit 1: ever removed (Inf - Inf: the reference's sum is NaN from then on)
#define XA_M_VSTICKY 6
#define XA_M_EBASE 8    // bit position of the error window's bit 0 (hint, written by the iteration kernel)
#define XA_M_VBASE 9
#define XA_M_MINKEY 10  // bits of the smallest errmax added since the last reset (~0: none); atomicMin
#define XA_M_EMAXKEY 11 // bits of the largest errmax / |val| added since the last reset (window hints); atomicMax
#define XA_M_VMAXKEY 12
#define XA_M_TOPKEY 13  // bits of the largest errmax ever added (never reset: an upper bound of the pool's largest errmax)

#endif
