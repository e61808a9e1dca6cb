// rule_kernels.cuh -- the FP64 rule kernels (K1 Genz-Malik 7/5, K2 Gauss-Kronrod 15), fused with
// bisection, point generation, the variable transform, the integrand, the rule sums, the error
// estimate and the split-dimension heuristic.
//
// This header is compiled at run time by NVRTC for sm_100a, once per (integrand, dim, fdim,
// transform on/off): the translation unit that includes it defines
//     CUB_DIM, CUB_FDIM, CUB_NPARAMS, CUB_HAS_TRANSFORM, CUB_INTEGRAND (a functor type, integrands.h)
// so every loop over the dimension is fully unrolled and every point coordinate lives in a register.
// It is also compiled by nvcc in csrc/static_check.cu for `-Xptxas -v` / cuobjdump inspection.
//
// Results must be bit-identical to the reference's (operation order of SURVEY.md Appendix A):
//   Rule75GenzMalik.java:61-184 (evalError), :237-326 (point generators, Gray-code order),
//   RuleGaussKronrod_1d.java:56-164, Region.java:31-38 (cut), Hypercube.java:36-41 (volume).
// Hence no FMA contraction (-fmad=false), serial left-to-right sums, and `c - r` / `c + r` with
// r = h*lambda rounded first.
//
// Work decomposition
//   gm scalar (fdim == 1): one thread per region; the P = 2^d + 2d^2 + 2d + 1 integrand values are
//       consumed as they are produced (no staging), sums stay in registers in reference order.
//   gk separable        : one thread per (region, component), 15 values in registers.
//   staged (any fdim)   : a CTA stages the [P][fdim] value tile of a few regions in shared memory,
//       then one thread per (region, component) runs the serial reference-order sums.
// Pool layout (structure of arrays in HBM, `cap` = capacity in regions):
//   center[d][cap], halfw[d][cap], val[f][cap], err[f][cap], errmax[cap], splitdim[cap]
// so consecutive threads (= consecutive regions) read and write consecutive addresses.
#ifndef CUBATURE_RULE_KERNELS_CUH
#define CUBATURE_RULE_KERNELS_CUH

#include "detmath.h"
#include "integrands.h"

#ifndef CUB_DIM
#error "CUB_DIM must be defined"
#endif

struct CubPoolView {
    double* center;   // [CUB_DIM][cap]
    double* halfw;    // [CUB_DIM][cap]
    double* val;      // [CUB_FDIM][cap]
    double* err;      // [CUB_FDIM][cap]
    double* errmax;   // [cap]
    int* splitdim;    // [cap]
    long long cap;
};

// Which regions a launch evaluates (n_items of them).
//   mode 0 PLAIN   : item i is slot  (list ? list[i] : base + i).
//   mode 1 BISECT  : item i is a CHILD of parent p = (list ? list[i/2] : i/2); the thread bisects the
//                    parent along its split dimension (Region.java:31-38) itself.  The left child
//                    (i even) replaces the parent in slot p, the right child (i odd) goes to slot
//                    base + i/2.  Only the one-thread-per-region kernels use this (K4 fused into K1).
//   mode 2 CHILDREN: same slot arithmetic as mode 1, but the child boxes were already written by
//                    the stand-alone bisect kernel (pool_kernels.cu) -- used when several threads
//                    share one region (fdim > 1), where the in-place update would race.
// When `dyn` is non-null the launch is sized by the DEVICE (no host round trip between the selection
// kernels and the rule kernel): dyn[0] = n_items, dyn[1] = 0 drops the list (pop-all iteration),
// dyn[2] = base.  The host then launches a grid for the largest possible batch and CTAs beyond the actual
// one retire at once.
#define CUB_MODE_PLAIN 0
#define CUB_MODE_BISECT 1
#define CUB_MODE_CHILDREN 2
struct CubWork {
    const int* list;
    const long long* dyn;
    long long n_items;
    long long base;
    int mode;
    // optional dense copies of the results in item order (host-driven exact mode reads them back)
    double* stage_val;  // [CUB_FDIM][stage_stride]
    double* stage_err;  // [CUB_FDIM][stage_stride]
    long long stage_stride;
    // device-sized launches only: this kernel takes the batch only if dyn_lo <= n_items < dyn_hi (dyn_hi == 0: always).
    // The host enqueues the one-thread-per-region kernel for [T, inf) and the point-parallel staged kernel for [0, T):
    // a small batch does not pay the serial latency of a whole region per thread.
    long long dyn_lo, dyn_hi;
};

struct CubParams {
#if CUB_NPARAMS > 0
    double v[CUB_NPARAMS];
#else
    double v[1];
#endif
};

struct CubTransform {
    int code[CUB_DIM];
    double shift[CUB_DIM];
};

// The always-exact variant of an integrand type: F::Exact where the family has a flagged fast path (integrands.h),
// F itself otherwise.  Only cub_gm_scalar runs the fast variant (it checks the flag and re-evaluates).
template <class F>
__device__ auto cub_exact_probe(int) -> typename F::Exact;
template <class F>
__device__ F cub_exact_probe(long);
typedef decltype(cub_exact_probe<CUB_INTEGRAND>(0)) CubExactIntegrand;

#define CUB_GM_POINTS (1 + 4 * CUB_DIM + 2 * CUB_DIM * (CUB_DIM - 1) + (1 << CUB_DIM))

__device__ __forceinline__ CubWork cub_resolve(const CubWork& w) {
    CubWork r = w;
    if (w.dyn) {
        r.n_items = w.dyn[0];
        if (!w.dyn[1]) r.list = nullptr;
        r.base = w.dyn[2];
        if (w.dyn_hi > 0 && (r.n_items < w.dyn_lo || r.n_items >= w.dyn_hi)) r.n_items = 0;  // another kernel's batch
    }
    return r;
}

__device__ __forceinline__ long long cub_item_slot(const CubWork& w, long long i) {
    if (w.mode == CUB_MODE_PLAIN) return w.list ? (long long)w.list[i] : w.base + i;
    const long long pi = i >> 1;
    return (i & 1) ? w.base + pi : (w.list ? (long long)w.list[pi] : pi);
}

// Loads the geometry of work item `i` (bisecting the parent when asked to) and returns its slot.
// All loads of a warp happen before any of its stores (__syncwarp) because in BISECT mode the two
// children of one parent sit in adjacent lanes and the left child overwrites what the right reads.
__device__ __forceinline__ long long cub_load_item(const CubPoolView& pool, const CubWork& w, long long i, bool active,
                                                   double (&c)[CUB_DIM], double (&h)[CUB_DIM]) {
    long long slot = 0;
    if (active) {
        slot = cub_item_slot(w, i);
        if (w.mode != CUB_MODE_BISECT) {
#pragma unroll
            for (int d = 0; d < CUB_DIM; ++d) {
                c[d] = pool.center[(long long)d * pool.cap + slot];
                h[d] = pool.halfw[(long long)d * pool.cap + slot];
            }
        } else {
            const long long pi = i >> 1;
            const long long parent = w.list ? (long long)w.list[pi] : pi;
            const int right = (int)(i & 1);
            const int s = pool.splitdim[parent];
#pragma unroll
            for (int d = 0; d < CUB_DIM; ++d) {
                c[d] = pool.center[(long long)d * pool.cap + parent];
                h[d] = pool.halfw[(long long)d * pool.cap + parent];
            }
#pragma unroll
            for (int d = 0; d < CUB_DIM; ++d) {
                if (d == s) {
                    h[d] = h[d] * 0.5;                          // Region.java:32
                    c[d] = right ? c[d] + h[d] : c[d] - h[d];   // Region.java:35-36
                }
            }
        }
    }
    if (w.mode == CUB_MODE_BISECT) {
        __syncwarp();
        if (active) {
#pragma unroll
            for (int d = 0; d < CUB_DIM; ++d) {
                pool.center[(long long)d * pool.cap + slot] = c[d];
                pool.halfw[(long long)d * pool.cap + slot] = h[d];
            }
        }
    }
    return slot;
}

// One integrand evaluation at rule point t (t-space), all fdim components, transform fused.
template <class F>
__device__ __forceinline__ void cub_eval_point(const F& fn, const CubTransform& tr, const double (&t)[CUB_DIM], double* f) {
#if CUB_HAS_TRANSFORM
    double x[CUB_DIM];
    double jac = 1.0;
#pragma unroll
    for (int d = 0; d < CUB_DIM; ++d) jac = jac * cub_transform_coord(tr.code[d], tr.shift[d], t[d], &x[d]);
    fn.eval(x, f);
#pragma unroll 1
    for (int j = 0; j < CUB_FDIM; ++j) f[j] = f[j] * jac;
#else
    fn.eval(t, f);
#endif
}

#if CUB_DIM >= 2
// ---- Genz-Malik constants (Rule75GenzMalik.java:44-53, :66-70) -----------------------------------
#define CUB_L2 0.3585685828003180919906451539079374954541
#define CUB_L4 0.9486832980505137995996680633298155601160
#define CUB_L5 0.6882472016116852977216287342936235251269
#define CUB_W1 ((12824 - 9120 * CUB_DIM + 400 * CUB_DIM * CUB_DIM) / 19683.0)
#define CUB_W2 (980.0 / 6561.0)
#define CUB_W3 ((1820 - 400 * CUB_DIM) / 19683.0)
#define CUB_W4 (200.0 / 19683.0)
#define CUB_W5 (6859.0 / (19683.0 * (1 << CUB_DIM)))
#define CUB_E1 ((729 - 950 * CUB_DIM + 50 * CUB_DIM * CUB_DIM) / 729.0)
#define CUB_E2 (245.0 / 486.0)
#define CUB_E3 ((265 - 100 * CUB_DIM) / 1458.0)
#define CUB_E4 (25.0 / 729.0)

#if CUB_FDIM == 1
// One region of the scalar Genz-Malik rule: all P integrand values are consumed as they are produced, sums in
// registers in reference order.  F is the integrand functor (fast or exact variant, see cub_fast_failed).
template <class F>
__device__ __forceinline__ void cub_gm_region(const F& fn, const CubTransform& tr, const double (&c)[CUB_DIM], const double (&h)[CUB_DIM],
                                              double& result, double& err, int& split_out) {
    const double ratio = (CUB_L2 * CUB_L2) / (CUB_L4 * CUB_L4);

    double vol = 1.0;  // Hypercube.java:36-41
#pragma unroll
    for (int d = 0; d < CUB_DIM; ++d) vol = vol * (h[d] * 2.0);

    double x[CUB_DIM];
#pragma unroll
    for (int d = 0; d < CUB_DIM; ++d) x[d] = c[d];

    double val0, sum2 = 0.0, sum3 = 0.0, sum4 = 0.0, sum5 = 0.0, fv;
    cub_eval_point(fn, tr, x, &val0);  // centre, Rule75GenzMalik.java:303-305

    // (lambda2,0..0) and (lambda4,0..0) orbits + fourth differences, Rule75GenzMalik.java:307-325,138-150
    double maxdiff = 0.0;
    int split = 0;
    const double twoval0 = 2 * val0;
#pragma unroll
    for (int d = 0; d < CUB_DIM; ++d) {
        const double r2 = h[d] * CUB_L2;
        const double r4 = h[d] * CUB_L4;
        double v0, v1, v2, v3;
        x[d] = c[d] - r2;
        cub_eval_point(fn, tr, x, &v0);
        x[d] = c[d] + r2;
        cub_eval_point(fn, tr, x, &v1);
        x[d] = c[d] - r4;
        cub_eval_point(fn, tr, x, &v2);
        x[d] = c[d] + r4;
        cub_eval_point(fn, tr, x, &v3);
        x[d] = c[d];
        const double s01 = v0 + v1, s23 = v2 + v3;
        sum2 = sum2 + s01;
        sum3 = sum3 + s23;
        const double diff = dm_abs((s01 - twoval0) - ratio * (s23 - twoval0));
        if (diff > maxdiff) {  // first strictly larger wins, Rule75GenzMalik.java:172-183
            maxdiff = diff;
            split = d;
        }
    }

    // (lambda4,lambda4,0..0) orbit, Rule75GenzMalik.java:271-297
    {
        double cm[CUB_DIM], cp[CUB_DIM];
#pragma unroll
        for (int d = 0; d < CUB_DIM; ++d) {
            const double r4 = h[d] * CUB_L4;
            cm[d] = c[d] - r4;
            cp[d] = c[d] + r4;
        }
#pragma unroll
        for (int a = 0; a < CUB_DIM - 1; ++a) {
#pragma unroll
            for (int b = a + 1; b < CUB_DIM; ++b) {
                x[a] = cm[a];
                x[b] = cm[b];
                cub_eval_point(fn, tr, x, &fv);
                sum4 = sum4 + fv;
                x[a] = cp[a];
                cub_eval_point(fn, tr, x, &fv);
                sum4 = sum4 + fv;
                x[b] = cp[b];
                cub_eval_point(fn, tr, x, &fv);
                sum4 = sum4 + fv;
                x[a] = cm[a];
                cub_eval_point(fn, tr, x, &fv);
                sum4 = sum4 + fv;
                x[b] = c[b];
            }
            x[a] = c[a];
        }
    }

    // (lambda5,..,lambda5) orbit in Gray-code order, Rule75GenzMalik.java:237-269.  Point n has sign
    // mask g(n) = n ^ (n >> 1) (bit set = minus); consecutive points differ in bit ctz(n + 1).
    {
        double cm[CUB_DIM], cp[CUB_DIM];
#pragma unroll
        for (int d = 0; d < CUB_DIM; ++d) {
            const double r5 = h[d] * CUB_L5;
            cm[d] = c[d] - r5;
            cp[d] = c[d] + r5;
            x[d] = cp[d];
        }
#ifndef CUB_GM_GRAY_LOW
#define CUB_GM_GRAY_LOW 4  // (3: +2.9 % rule time on the d = 6 corner peak; 5, 6: spills, profiles/r02/summary.md section 6)
#endif
        constexpr int LOW = CUB_DIM < CUB_GM_GRAY_LOW ? CUB_DIM : CUB_GM_GRAY_LOW;  // bits flipped inside one unrolled block
        constexpr int BLOCK = 1 << LOW;
#pragma unroll 1
        for (int blk = 0; blk < (1 << CUB_DIM) / BLOCK; ++blk) {
            const unsigned g = (unsigned)(blk * BLOCK) ^ ((unsigned)(blk * BLOCK) >> 1);
            // coordinates >= LOW keep their sign for the whole block
#pragma unroll
            for (int d = LOW; d < CUB_DIM; ++d) x[d] = ((g >> d) & 1u) ? cm[d] : cp[d];
            // low coordinates: bit LOW-1 of g(n) also depends on bit LOW of n (= parity of blk)
            const unsigned flip_top = (unsigned)blk & 1u;
#pragma unroll
            for (int j = 0; j < BLOCK; ++j) {
                const unsigned gl = (unsigned)j ^ ((unsigned)j >> 1);  // compile-time after unrolling
#pragma unroll
                for (int d = 0; d < LOW; ++d) {
                    unsigned bit = (gl >> d) & 1u;
                    if (d == LOW - 1) bit ^= flip_top;
                    x[d] = bit ? cm[d] : cp[d];
                }
                cub_eval_point(fn, tr, x, &fv);
                sum5 = sum5 + fv;
            }
        }
    }

    // Rule75GenzMalik.java:163-167
    result = vol * (CUB_W1 * val0 + CUB_W2 * sum2 + CUB_W3 * sum3 + CUB_W4 * sum4 + CUB_W5 * sum5);
    const double res5th = vol * (CUB_E1 * val0 + CUB_E2 * sum2 + CUB_E3 * sum3 + CUB_E4 * sum4);
    err = dm_abs(res5th - result);
    split_out = split;
}

// Integrands may carry a branch-free fast path that is exact only on a stated argument range (detmath.h:
// dm_rcp_flag); they record leaving it in a member `bad` and name their always-exact variant `Exact`.  The kernel
// then re-evaluates that region with the exact variant (out of line: never fetched unless needed).
template <class F>
__device__ __forceinline__ auto cub_fast_failed(const F& f, int) -> decltype(f.bad != 0) { return f.bad != 0; }
template <class F>
__device__ __forceinline__ bool cub_fast_failed(const F&, long) { return false; }
struct CubRegionResult {
    double result, err;
    int split;
};
template <class F, class E = typename F::Exact>
__device__ __noinline__ CubRegionResult cub_gm_region_exact(const CubPoolView& pool, long long slot, const CubParams& prm,
                                                            const CubTransform& tr, int) {
    double c[CUB_DIM], h[CUB_DIM];
#pragma unroll
    for (int d = 0; d < CUB_DIM; ++d) {
        c[d] = pool.center[(long long)d * pool.cap + slot];
        h[d] = pool.halfw[(long long)d * pool.cap + slot];
    }
    const E fe(prm.v, 1);
    CubRegionResult r;
    cub_gm_region(fe, tr, c, h, r.result, r.err, r.split);
    return r;
}
template <class F>
__device__ __forceinline__ CubRegionResult cub_gm_region_exact(const CubPoolView&, long long, const CubParams&, const CubTransform&, long) {
    return CubRegionResult();
}

// Integrands whose fast path can be proved safe for a whole box (integrands.h: box_safe / Unchecked): the kernel asks
// once per region and then runs the variant without any per-point check (the range checks were 1.5 of the ~21
// instructions per integrand value); a region that cannot be proved safe goes to the exact variant.  Not used with a
// variable transform (the box is in t-space, the integrand's argument is not).
#ifndef CUB_GM_BOX_GUARD
#define CUB_GM_BOX_GUARD 1
#endif
template <class F>
__device__ auto cub_unchecked_probe(int) -> typename F::Unchecked;
template <class F>
__device__ F cub_unchecked_probe(long);
typedef decltype(cub_unchecked_probe<CUB_INTEGRAND>(0)) CubUncheckedIntegrand;
template <class A, class B>
struct CubSameType {
    static const bool value = false;
};
template <class A>
struct CubSameType<A, A> {
    static const bool value = true;
};
template <class F>
__device__ __forceinline__ auto cub_box_safe(const F& f, const double* c, const double* h, int) -> decltype(f.box_safe(c, h)) {
    return f.box_safe(c, h);
}
template <class F>
__device__ __forceinline__ bool cub_box_safe(const F&, const double*, const double*, long) {
    return false;
}
#define CUB_GM_HAS_BOX_GUARD (CUB_GM_BOX_GUARD && !CUB_HAS_TRANSFORM && !CubSameType<CubUncheckedIntegrand, CUB_INTEGRAND>::value)

// K1, scalar integrand: one thread per region.
#ifndef CUB_GM_MIN_BLOCKS
#define CUB_GM_MIN_BLOCKS 5
#endif
// (tried and dropped: per-thread coordinate tables in shared memory to raise occupancy -- slower at every
//  CTA count, profiles/r01_summary.md)
// Geometry pipeline (CUB_GM_PREFETCH, default on).  A tile starts with a chain of two dependent global loads (list
// entry -> parent geometry) in front of ~2 800 instructions of FP64 work; with four warps per scheduler that chain
// showed up as 15 % of the warps' time in "long scoreboard" (profiles/r02/ncu_gm_scalar_raw.csv).  Every thread
// therefore fetches the geometry of ITS item of the CTA's next tile, and the list entry of the tile after that, with
// cp.async into thread-private shared-memory cells (two buffers, [field][thread]: conflict-free) while it evaluates
// the current tile; nothing but the tile counter stays live in registers across the evaluation.  The cells are
// private to a thread, so no CTA barrier is needed, only cp.async.wait_all.  In BISECT mode the parents of the next
// tile are written by nobody but that tile itself, so reading them ahead is safe.
#ifndef CUB_GM_THREADS
#define CUB_GM_THREADS 128
#endif
#ifndef CUB_GM_PREFETCH  // (the cells must fit well inside the 48 KB of static shared memory: d <= 10 at 128 threads)
#define CUB_GM_PREFETCH ((2 * 2 * CUB_DIM * 8 + 3 * 2 * 4) * CUB_GM_THREADS <= 44 * 1024)
#endif
__device__ __forceinline__ void cub_cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cub_cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cub_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cub_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// (CUB_GM_MAXNREG: an explicit register cap instead of an occupancy target -- tuning only; the host then needs
//  CUBATURE_GM_THREADS, because the kernel no longer carries its CTA size as an attribute)
#if defined(CUB_GM_MAXNREG)
#define CUB_GM_BOUNDS __maxnreg__(CUB_GM_MAXNREG)
#elif CUB_GM_MIN_BLOCKS > 0
#define CUB_GM_BOUNDS __launch_bounds__(CUB_GM_THREADS, CUB_GM_MIN_BLOCKS)
#else
#define CUB_GM_BOUNDS __launch_bounds__(CUB_GM_THREADS)
#endif
extern "C" __global__ void CUB_GM_BOUNDS
cub_gm_scalar(CubPoolView pool, CubWork work_in, const __grid_constant__ CubParams prm, const __grid_constant__ CubTransform tr) {
    // Persistent: the grid covers the machine once and every CTA walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    // so a launch needs no host knowledge of the batch size (device-sized launches: CubWork::dyn).
    const CubWork work = cub_resolve(work_in);
    const long long n = work.n_items;
#if CUB_GM_PREFETCH
    __shared__ double s_geo[2][2 * CUB_DIM][CUB_GM_THREADS];  // centre[d], halfwidth[d] of the item's source region
    __shared__ int s_split[2][CUB_GM_THREADS];                 // its split dimension (BISECT)
    __shared__ int s_src[2][CUB_GM_THREADS];                   // the source region's slot
    __shared__ int s_list[2][CUB_GM_THREADS];                  // list entry of the tile after the next
    const int tid = threadIdx.x;
    // item of this thread in the CTA's k-th tile, and where its geometry comes from: the slot itself (PLAIN) or the
    // parent it is a child of (BISECT); `list` holds one entry per slot / per parent
#define CUB_GM_ITEM(k) ((((long long)blockIdx.x + (long long)(k) * gridDim.x) * CUB_GM_THREADS) + tid)
#define CUB_GM_LIST_INDEX(i) (work.mode == CUB_MODE_PLAIN ? (i) : ((i) >> 1))
#define CUB_GM_SRC_NOLIST(i) (work.mode == CUB_MODE_PLAIN ? (int)(work.base + (i)) : (int)((i) >> 1))
    // (CHILDREN mode -- boxes already written, not used with this kernel today -- reads the child's own slot)
#define CUB_GM_SRC_FIX(src, i) \
    if (work.mode == CUB_MODE_CHILDREN && ((i) & 1)) src = (int)(work.base + ((i) >> 1))
#define CUB_GM_FETCH_GEOMETRY(src, buf)                                                                       \
    do {                                                                                                      \
        _Pragma("unroll") for (int d = 0; d < CUB_DIM; ++d) {                                                 \
            cub_cp_async8(&s_geo[buf][d][tid], &pool.center[(long long)d * pool.cap + (src)]);                \
            cub_cp_async8(&s_geo[buf][CUB_DIM + d][tid], &pool.halfw[(long long)d * pool.cap + (src)]);       \
        }                                                                                                     \
        if (work.mode == CUB_MODE_BISECT) cub_cp_async4(&s_split[buf][tid], &pool.splitdim[(src)]);           \
    } while (0)
    {  // prologue: tile 0's geometry and tile 1's list entry
        const long long i0 = CUB_GM_ITEM(0);
        if (i0 < n) {
            int src0 = work.list ? work.list[CUB_GM_LIST_INDEX(i0)] : CUB_GM_SRC_NOLIST(i0);
            CUB_GM_SRC_FIX(src0, i0);
            s_src[0][tid] = src0;
            CUB_GM_FETCH_GEOMETRY(src0, 0);
        }
        const long long i1 = CUB_GM_ITEM(1);
        if (work.list && i1 < n) cub_cp_async4(&s_list[1][tid], &work.list[CUB_GM_LIST_INDEX(i1)]);
        cub_cp_async_commit();
    }
    for (int k = 0; ((long long)blockIdx.x + (long long)k * gridDim.x) * CUB_GM_THREADS < n; ++k) {
        const int b = k & 1;
        const long long i = CUB_GM_ITEM(k);
        const bool active = i < n;
        cub_cp_async_wait_all();
        double c[CUB_DIM], h[CUB_DIM];
        int src = 0, s = 0;
#pragma unroll
        for (int d = 0; d < CUB_DIM; ++d) c[d] = h[d] = 0.0;
        if (active) {
#pragma unroll
            for (int d = 0; d < CUB_DIM; ++d) {
                c[d] = s_geo[b][d][tid];
                h[d] = s_geo[b][CUB_DIM + d][tid];
            }
            src = s_src[b][tid];
            if (work.mode == CUB_MODE_BISECT) s = s_split[b][tid];
        }
        {  // next tile's geometry, list entry of the one after
            const long long i1 = CUB_GM_ITEM(k + 1);
            if (i1 < n) {
                int src1 = work.list ? s_list[b ^ 1][tid] : CUB_GM_SRC_NOLIST(i1);
                CUB_GM_SRC_FIX(src1, i1);
                s_src[b ^ 1][tid] = src1;
                CUB_GM_FETCH_GEOMETRY(src1, b ^ 1);
            }
            const long long i2 = CUB_GM_ITEM(k + 2);
            if (work.list && i2 < n) cub_cp_async4(&s_list[b][tid], &work.list[CUB_GM_LIST_INDEX(i2)]);
            cub_cp_async_commit();
        }
        long long slot = src;
        if (work.mode == CUB_MODE_BISECT) {
            // Region.java:31-38; the left child replaces the parent, the right one goes to base + parent index.  Both
            // children (adjacent lanes) hold their copy of the parent before either writes (__syncwarp).
            const int right = (int)(i & 1);
            if (active) {
#pragma unroll
                for (int d = 0; d < CUB_DIM; ++d) {
                    if (d == s) {
                        h[d] = h[d] * 0.5;
                        c[d] = right ? c[d] + h[d] : c[d] - h[d];
                    }
                }
                if (right) slot = work.base + (i >> 1);
            }
            __syncwarp();
            if (active) {
#pragma unroll
                for (int d = 0; d < CUB_DIM; ++d) {
                    pool.center[(long long)d * pool.cap + slot] = c[d];
                    pool.halfw[(long long)d * pool.cap + slot] = h[d];
                }
            }
        }
#else
    for (long long tile = blockIdx.x; tile * CUB_GM_THREADS < n; tile += gridDim.x) {
        const long long i = tile * CUB_GM_THREADS + threadIdx.x;
        const bool active = i < n;
        double c[CUB_DIM], h[CUB_DIM];
#pragma unroll
        for (int d = 0; d < CUB_DIM; ++d) c[d] = h[d] = 0.0;
        const long long slot = cub_load_item(pool, work, i, active, c, h);
#endif
        double result = 0.0, err = 0.0;
        int split = 0;
        if (active) {
            const CUB_INTEGRAND fn(prm.v, 1);
            bool redo;
            if (CUB_GM_HAS_BOX_GUARD) {  // (compile-time)
                redo = !cub_box_safe(fn, c, h, 0);
                if (!redo) {
                    const CubUncheckedIntegrand fu(prm.v, 1);
                    cub_gm_region(fu, tr, c, h, result, err, split);
                }
            } else {
                cub_gm_region(fn, tr, c, h, result, err, split);
                redo = cub_fast_failed(fn, 0);  // an argument left the fast path's exact range
            }
            if (redo) {  // redo with IEEE division
                const CubRegionResult r = cub_gm_region_exact<CUB_INTEGRAND>(pool, slot, prm, tr, 0);
                result = r.result;
                err = r.err;
                split = r.split;
            }
        }
        if (active) {
            double emax = 0.0;  // Rule.java:281-289
            if (err > emax) emax = err;
            pool.val[slot] = result;
            pool.err[slot] = err;
            pool.errmax[slot] = emax;
            pool.splitdim[slot] = split;
            if (work.stage_val) {
                work.stage_val[i] = result;
                work.stage_err[i] = err;
            }
        }
    }
}
#endif  // CUB_FDIM == 1
#endif  // CUB_DIM >= 2

#if CUB_DIM == 1
// ---- Gauss-Kronrod 15 tables, RuleGaussKronrod_1d.java:20-49 --------------------------------------
__device__ __forceinline__ void cub_gk15_sums(const double (&q)[15], double halfwidth, double* val_out, double* err_out) {
    const double xw_wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                             0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
    const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                           0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                           0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                           0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
    const double DBL_EPS = 2.220446049250313e-16;
    const double MIN_NORMAL = 2.2250738585072014e-308;
    // RuleGaussKronrod_1d.java:98-118
    double rg = q[0] * xw_wg[3];
    double rk = q[0] * wgk[7];
    double rabs = dm_abs(rk);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int k = 1 + 2 * j;
        const double v = q[k] + q[k + 1];
        rg = rg + xw_wg[j] * v;
        rk = rk + wgk[2 * j + 1] * v;
        rabs = rabs + wgk[2 * j + 1] * (dm_abs(q[k]) + dm_abs(q[k + 1]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = 7 + 2 * j;
        rk = rk + wgk[2 * j] * (q[k] + q[k + 1]);
        rabs = rabs + wgk[2 * j] * (dm_abs(q[k]) + dm_abs(q[k + 1]));
    }
    *val_out = rk * halfwidth;  // :121
    // :128-157
    const double mean = rk * 0.5;
    double rasc = wgk[7] * dm_abs(q[0] - mean);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int k = 1 + 2 * j;
        rasc = rasc + wgk[2 * j + 1] * (dm_abs(q[k] - mean) + dm_abs(q[k + 1] - mean));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = 7 + 2 * j;
        rasc = rasc + wgk[2 * j] * (dm_abs(q[k] - mean) + dm_abs(q[k + 1] - mean));
    }
    double err = dm_abs(rk - rg) * halfwidth;
    rabs = rabs * halfwidth;
    rasc = rasc * halfwidth;
    if (rasc != 0 && err != 0) {
        const double base = 200 * err / rasc;
        const double scale = base * dm_sqrt(base);  // deterministic stand-in for Math.pow(base, 1.5)
        err = (scale < 1) ? rasc * scale : rasc;
    }
    if (rabs > MIN_NORMAL / (50 * DBL_EPS)) {
        const double min_err = 50 * DBL_EPS * rabs;
        if (min_err > err) err = min_err;
    }
    *err_out = err;
}

// node k of the 15-point rule in the reference's order (RuleGaussKronrod_1d.java:60-84)
__device__ __forceinline__ double cub_gk15_node(int k, double c, double h) {
    const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                           0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                           0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                           0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
    if (k == 0) return c;
    // k = 1..6: Gauss nodes xgk[1], xgk[3], xgk[5]; k = 7..14: Kronrod nodes xgk[0], xgk[2], xgk[4], xgk[6]
    const int idx = k < 7 ? 2 * ((k - 1) >> 1) + 1 : 2 * ((k - 7) >> 1);
    const double w = h * xgk[idx];
    return (k & 1) ? c - w : c + w;
}

#ifndef CUB_SEPARABLE
#define CUB_SEPARABLE 0
#endif
#if CUB_FDIM == 1 || CUB_SEPARABLE
// K2, separable integrand (or fdim == 1): one thread per (component, region); regions fastest, so
// lanes of a warp hold consecutive regions and all stores are coalesced.  With fdim > 1 the host
// uses mode CHILDREN (several threads share a region), with fdim == 1 the fused BISECT mode.
extern "C" __global__ void __launch_bounds__(128)
cub_gk_separable(CubPoolView pool, CubWork work_in, const __grid_constant__ CubParams prm, const __grid_constant__ CubTransform tr) {
    const CubWork work = cub_resolve(work_in);
    const long long n = work.n_items;
    for (long long tile = blockIdx.x; tile * 128 < n * CUB_FDIM; tile += gridDim.x) {  // persistent, see cub_gm_scalar
        const long long t = tile * 128 + threadIdx.x;
        const bool active = t < n * CUB_FDIM;
        const long long i = active ? t % n : 0;
        const int j = active ? (int)(t / n) : 0;
        double cc[1] = {0.0}, hh[1] = {0.0};
        const long long slot = cub_load_item(pool, work, i, active, cc, hh);
        double val = 0.0, err = 0.0;
        if (active) {
            const double c = cc[0], h = hh[0];
            const CubExactIntegrand fn(prm.v, CUB_FDIM);
            double q[15];
#pragma unroll
            for (int k = 0; k < 15; ++k) {
                double tt[1] = {cub_gk15_node(k, c, h)};
#if CUB_HAS_TRANSFORM
                double xx[1];
                const double jac = 1.0 * cub_transform_coord(tr.code[0], tr.shift[0], tt[0], &xx[0]);
#if CUB_FDIM == 1
                fn.eval(xx, &q[k]);
#else
                q[k] = fn.eval1(xx, j);
#endif
                q[k] = q[k] * jac;
#else
#if CUB_FDIM == 1
                fn.eval(tt, &q[k]);
#else
                q[k] = fn.eval1(tt, j);
#endif
#endif
            }
            cub_gk15_sums(q, h, &val, &err);
        }
        if (active) {
            pool.val[(long long)j * pool.cap + slot] = val;
            pool.err[(long long)j * pool.cap + slot] = err;
#if CUB_FDIM == 1
            double emax = 0.0;
            if (err > emax) emax = err;
            pool.errmax[slot] = emax;
            pool.splitdim[slot] = 0;
#endif
            // fdim > 1: errmax / splitdim are produced by cub_errmax_rows (pool_kernels.cu) after this launch
            if (work.stage_val) {
                work.stage_val[(long long)j * work.stage_stride + i] = val;
                work.stage_err[(long long)j * work.stage_stride + i] = err;
            }
        }
    }
}
#endif  // CUB_FDIM == 1 || CUB_SEPARABLE
#endif  // CUB_DIM == 1


// ---- staged kernel: any fdim, any dim -------------------------------------------------------------
// A CTA processes `rpb` regions per pass.  Phase 1: one thread per (region, point) evaluates all
// fdim components into the shared tile sval[r][k][0..fdim) (row stride FS = fdim rounded up to odd:
// conflict-free for the column reads of phase 2).  Phase 2: one thread per (region, component) runs
// the reference-order serial sums over the P values of its column.  Phase 3 (Genz-Malik): one
// thread per (region, dimension) accumulates the fourth differences over the components in
// component order (Rule75GenzMalik.java:121,147-149: diff[k] is summed over all fdim).  Phase 4: one
// thread per region picks the split dimension and errmax (Rule.java:281-289).
#if CUB_DIM >= 2
#define CUB_RULE_POINTS CUB_GM_POINTS
// rule point k of box (c, h), in the reference's order; k is a run-time
// value here, so coordinates are picked with selects instead of register renaming.
__device__ __forceinline__ void cub_gm_point(int k, const double* c, const double* h, double (&x)[CUB_DIM]) {
#pragma unroll
    for (int d = 0; d < CUB_DIM; ++d) x[d] = c[d];
    if (k == 0) return;
    k -= 1;
    if (k < 4 * CUB_DIM) {
        const int i = k >> 2, q = k & 3;
        const double lam = q < 2 ? CUB_L2 : CUB_L4;
#pragma unroll
        for (int d = 0; d < CUB_DIM; ++d) {
            if (d == i) {
                const double r = h[d] * lam;
                x[d] = (q & 1) ? c[d] + r : c[d] - r;
            }
        }
        return;
    }
    k -= 4 * CUB_DIM;
    if (k < 2 * CUB_DIM * (CUB_DIM - 1)) {
        int pair = k >> 2;
        const int q = k & 3;
        int a = 0;
        while (pair >= CUB_DIM - 1 - a) {
            pair -= CUB_DIM - 1 - a;
            ++a;
        }
        const int b = a + 1 + pair;
        const bool a_minus = (q == 0) || (q == 3);  // (-,-) (+,-) (+,+) (-,+), Rule75GenzMalik.java:274-296
        const bool b_minus = q < 2;
#pragma unroll
        for (int d = 0; d < CUB_DIM; ++d) {
            if (d == a || d == b) {
                const double r = h[d] * CUB_L4;
                const bool minus = d == a ? a_minus : b_minus;
                x[d] = minus ? c[d] - r : c[d] + r;
            }
        }
        return;
    }
    k -= 2 * CUB_DIM * (CUB_DIM - 1);
    const unsigned g = (unsigned)k ^ ((unsigned)k >> 1);
#pragma unroll
    for (int d = 0; d < CUB_DIM; ++d) {
        const double r = h[d] * CUB_L5;
        x[d] = ((g >> d) & 1u) ? c[d] - r : c[d] + r;
    }
}
#else
#define CUB_RULE_POINTS 15
#endif

#define CUB_FS ((CUB_FDIM) | 1)

extern "C" __global__ void __launch_bounds__(256)
cub_rule_staged(CubPoolView pool, CubWork work_in, const __grid_constant__ CubParams prm, const __grid_constant__ CubTransform tr, int rpb) {
    const CubWork work = cub_resolve(work_in);
    extern __shared__ double cub_smem[];
    double* sval = cub_smem;                                        // [rpb][P][FS]
    double* sc = sval + (long long)rpb * CUB_RULE_POINTS * CUB_FS;  // [rpb][DIM]
    double* sh = sc + rpb * CUB_DIM;                                // [rpb][DIM]
    double* sdiff = sh + rpb * CUB_DIM;                             // [rpb][DIM]
    double* serr = sdiff + rpb * CUB_DIM;                           // [rpb][FDIM]
    long long* sslot = (long long*)(serr + (long long)rpb * CUB_FDIM);  // [rpb] the item's slot
    long long* ssrc = sslot + rpb;                                      // [rpb] where its geometry is read (BISECT: the parent)
    int* ssplit = (int*)(ssrc + rpb);                                   // [rpb] the parent's split dimension (BISECT)

    const long long n = work.n_items;
    const CubExactIntegrand fn(prm.v, CUB_FDIM);
    const int tid = threadIdx.x, nt = blockDim.x;

    for (long long first = (long long)blockIdx.x * rpb; first < n; first += (long long)gridDim.x * rpb) {
        const int nr = (int)(n - first < rpb ? n - first : rpb);
        for (int r = tid; r < nr; r += nt) {
            const long long slot = cub_item_slot(work, first + r);
            sslot[r] = slot;
            if (work.mode == CUB_MODE_BISECT) {
                const long long pi = (first + r) >> 1;
                const long long parent = work.list ? (long long)work.list[pi] : pi;
                ssrc[r] = parent;
                ssplit[r] = pool.splitdim[parent];
            } else {
                ssrc[r] = slot;
            }
        }
        __syncthreads();
        for (int idx = tid; idx < nr * CUB_DIM; idx += nt) {
            const int r = idx / CUB_DIM, d = idx - r * CUB_DIM;
            sc[idx] = pool.center[(long long)d * pool.cap + ssrc[r]];
            sh[idx] = pool.halfw[(long long)d * pool.cap + ssrc[r]];
        }
        __syncthreads();
        if (work.mode == CUB_MODE_BISECT) {
            // Region.java:31-38 in shared memory; both children of a parent are items 2k, 2k + 1 of the SAME pass (rpb is
            // even, the host's contract), and every read of the parent precedes the barrier above: the left child may
            // now overwrite the parent's slot
            for (int idx = tid; idx < nr * CUB_DIM; idx += nt) {
                const int r = idx / CUB_DIM, d = idx - r * CUB_DIM;
                double c = sc[idx], hw = sh[idx];
                if (d == ssplit[r]) {
                    hw = hw * 0.5;
                    c = ((first + r) & 1) ? c + hw : c - hw;
                    sc[idx] = c;
                    sh[idx] = hw;
                }
                pool.center[(long long)d * pool.cap + sslot[r]] = c;
                pool.halfw[(long long)d * pool.cap + sslot[r]] = hw;
            }
            __syncthreads();
        }
        // phase 1
        for (int idx = tid; idx < nr * CUB_RULE_POINTS; idx += nt) {
            const int r = idx / CUB_RULE_POINTS, k = idx - r * CUB_RULE_POINTS;
            double t[CUB_DIM];
#if CUB_DIM >= 2
            cub_gm_point(k, sc + r * CUB_DIM, sh + r * CUB_DIM, t);
#else
            t[0] = cub_gk15_node(k, sc[r], sh[r]);
#endif
            cub_eval_point(fn, tr, t, sval + (long long)idx * CUB_FS);
        }
        __syncthreads();
        // phase 2
        for (int idx = tid; idx < nr * CUB_FDIM; idx += nt) {
            const int r = idx / CUB_FDIM, j = idx - r * CUB_FDIM;
            const double* v = sval + (long long)r * CUB_RULE_POINTS * CUB_FS + j;  // v[k*FS]
            double val, err;
#if CUB_DIM >= 2
            const double* h = sh + r * CUB_DIM;
            double vol = 1.0;
#pragma unroll
            for (int d = 0; d < CUB_DIM; ++d) vol = vol * (h[d] * 2.0);
            const double val0 = v[0];
            double sum2 = 0.0, sum3 = 0.0, sum4 = 0.0, sum5 = 0.0;
            int k0 = 1;
#pragma unroll
            for (int d = 0; d < CUB_DIM; ++d) {
                sum2 = sum2 + (v[(k0 + 4 * d) * CUB_FS] + v[(k0 + 4 * d + 1) * CUB_FS]);
                sum3 = sum3 + (v[(k0 + 4 * d + 2) * CUB_FS] + v[(k0 + 4 * d + 3) * CUB_FS]);
            }
            k0 += 4 * CUB_DIM;
            for (int k = 0; k < 2 * CUB_DIM * (CUB_DIM - 1); ++k) sum4 = sum4 + v[(k0 + k) * CUB_FS];
            k0 += 2 * CUB_DIM * (CUB_DIM - 1);
            for (int k = 0; k < (1 << CUB_DIM); ++k) sum5 = sum5 + v[(k0 + k) * CUB_FS];
            val = vol * (CUB_W1 * val0 + CUB_W2 * sum2 + CUB_W3 * sum3 + CUB_W4 * sum4 + CUB_W5 * sum5);
            const double res5th = vol * (CUB_E1 * val0 + CUB_E2 * sum2 + CUB_E3 * sum3 + CUB_E4 * sum4);
            err = dm_abs(res5th - val);
#else
            double q[15];
#pragma unroll
            for (int k = 0; k < 15; ++k) q[k] = v[k * CUB_FS];
            cub_gk15_sums(q, sh[r], &val, &err);
#endif
            serr[idx] = err;
            const long long slot = sslot[r];
            pool.val[(long long)j * pool.cap + slot] = val;
            pool.err[(long long)j * pool.cap + slot] = err;
            if (work.stage_val) {
                work.stage_val[(long long)j * work.stage_stride + first + r] = val;
                work.stage_err[(long long)j * work.stage_stride + first + r] = err;
            }
        }
#if CUB_DIM >= 2
        // phase 3
        for (int idx = tid; idx < nr * CUB_DIM; idx += nt) {
            const int r = idx / CUB_DIM, d = idx - r * CUB_DIM;
            const double* v = sval + (long long)r * CUB_RULE_POINTS * CUB_FS;
            const double ratio = (CUB_L2 * CUB_L2) / (CUB_L4 * CUB_L4);
            double diff = 0.0;
            for (int j = 0; j < CUB_FDIM; ++j) {
                const double val0 = v[j];
                const double v0 = v[(1 + 4 * d) * CUB_FS + j], v1 = v[(2 + 4 * d) * CUB_FS + j];
                const double v2 = v[(3 + 4 * d) * CUB_FS + j], v3 = v[(4 + 4 * d) * CUB_FS + j];
                diff = diff + dm_abs(((v0 + v1) - 2 * val0) - ratio * ((v2 + v3) - 2 * val0));
            }
            sdiff[idx] = diff;
        }
#endif
        __syncthreads();
        // phase 4
        for (int r = tid; r < nr; r += nt) {
            double emax = 0.0;
            for (int j = 0; j < CUB_FDIM; ++j) {
                const double e = serr[r * CUB_FDIM + j];
                if (e > emax) emax = e;
            }
            int split = 0;
#if CUB_DIM >= 2
            double maxdiff = 0.0;
#pragma unroll
            for (int d = 0; d < CUB_DIM; ++d) {
                if (sdiff[r * CUB_DIM + d] > maxdiff) {
                    maxdiff = sdiff[r * CUB_DIM + d];
                    split = d;
                }
            }
#endif
            pool.errmax[sslot[r]] = emax;
            pool.splitdim[sslot[r]] = split;
        }
        __syncthreads();
    }
}

// Integrand alone at caller-supplied points x[dim][n] -> out[fdim][n] (bit-parity tests of detmath
// and of the built-in families against the g++ build of the same header).
extern "C" __global__ void cub_eval_points(const double* x, double* out, long long n, const __grid_constant__ CubParams prm,
                                           const __grid_constant__ CubTransform tr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const CubExactIntegrand fn(prm.v, CUB_FDIM);
    double t[CUB_DIM];
#pragma unroll
    for (int d = 0; d < CUB_DIM; ++d) t[d] = x[(long long)d * n + i];
    double f[CUB_FDIM];  // local array; large fdim spills (tests only)
    cub_eval_point(fn, tr, t, f);
    for (int j = 0; j < CUB_FDIM; ++j) out[(long long)j * n + i] = f[j];
}

#endif  // CUBATURE_RULE_KERNELS_CUH
