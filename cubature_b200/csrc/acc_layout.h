// acc_layout.h -- layout of the exact running sums of a region pool (plain #defines: included by the NVRTC-compiled rule
// kernels, by the nvcc-compiled pool kernels and by the host).
//
// The reference keeps Sum(val), Sum(err) over the heap as drifting incremental doubles (RegionHeap.java:21-36).  Here
// they are INTEGER accumulators, one bin per binary exponent: a double x = m * 2^(e' - 1075) with the 53-bit integer
// mantissa m (implicit bit included) and e' = max(1, biased exponent) adds 1 to cnt[e], (m & 0xffffffff) to lo[e] and
// (m >> 32) to hi[e].  Integer addition is associative, so the sums are exact, independent of the order in which threads,
// CTAs or GPUs contribute, and a region that is popped is subtracted again without drift.  The error bins double as the
// level-0 histogram of the batch selection (one bucket per binade of errmax).
#ifndef CUB_ACC_LAYOUT_H
#define CUB_ACC_LAYOUT_H

#define XA_NBINS 2048
// 8-byte words of one accumulator set:
//   E bins  [XA_NBINS][3]: cnt, lo, hi      regions by the binade of errmax = max(0, err) (NaN counts as 0, Rule.java:281-289)
//   V bins  [XA_NBINS][2]: lo, hi           two's complement: negative values subtract their mantissa
//   misc    [32]
#define XA_E_OFF 0
#define XA_V_OFF (3 * XA_NBINS)
#define XA_M_OFF (5 * XA_NBINS)
#define XA_WORDS (5 * XA_NBINS + 32)
// misc words
#define XA_M_ENAN 0     // regions currently in the pool whose err is NaN
#define XA_M_EINF 1     //   ... +Inf
#define XA_M_VNAN 2     // regions whose val is NaN
#define XA_M_VPINF 3
#define XA_M_VNINF 4
#define XA_M_ESTICKY 5  // bit 0: a non-finite err was ever added, bit 1: ever removed (Inf - Inf: the reference's sum is NaN from then on)
#define XA_M_VSTICKY 6
#define XA_M_EBASE 8    // window hints for the rule kernels (written by the iteration kernel): first binade of the shared-memory
#define XA_M_VBASE 9    //   window of the error / value bins
#define XA_M_MINKEY 10  // bits of the smallest errmax added since the last reset (~0: none); atomicMin

// window of binades the rule kernels pre-aggregate in shared memory (values outside go to global memory directly)
#define XA_WIN 64

#endif
