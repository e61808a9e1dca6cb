// iter_exact.cu -- the device-resident refinement loop of scalar integrands (fdim == 1): convergence test, Gladwell batch
// selection and compaction of Rule.cubature (Rule.java:72-142) on EXACT running sums.
//
// The reference keeps Sum(val), Sum(err) of its heap as incremental doubles (RegionHeap.java:21-36) and decides how many
// regions to pop from the sequential residual `ee.err -= popped.err` (Rule.java:122-124, :129).  Here both sums are integer
// accumulators (acc_layout.h), maintained by the decide kernel (acc_device.cuh), so every decision is the one the reference's
// loop takes when its sums carry no rounding error: independent of the order in which threads, CTAs or GPUs contribute, hence
// identical for any grid size and any number of ranks by construction.  The comparison `residual < tolerance` is exact: the
// residual is a big integer (exact_acc.h) converted to a double truncated toward zero, for which `double < tol` is the exact
// test; the value sum is rounded to nearest.  The oracle's exact-sums variant (oracle/cubature_oracle.cpp:
// orc_set_exact_sums) takes literally these decisions; where they differ from the reference's drifting sums the difference
// is an artefact of the reference's accumulated rounding (reported as cub_stats_t.ambiguous_decisions).
//
//   k_exact_decide (grid)      phase 1, all CTAs: stream over the children the last rule launch wrote and over the (val, err)
//                              pairs of the parents the last selection popped, and add / subtract them (acc_device.cuh);
//                              phase 2, the CTA that finishes last: totals, loop condition and convergence (Rule.java:72-79),
//                              evaluation budget, the pop-everything shortcut (exact smallest errmax)
//   k_exact_select (coop grid) walk down the binades, radix select inside the threshold binade on the 52 fraction bits
//                              (8+8+4x9), integer bucket sums, stable compaction of the popped slots (and a copy of their
//                              val / err for the next decide); retires at once when there is nothing to cut
//   rule kernel (NVRTC)        bisects + evaluates the batch; knows nothing about the sums
// Sharded pools (R > 1): the ranks exchange their bins / histograms through peer-mapped buffers inside these kernels
// (push over NVLink, flag, spin); integer sums need no fixed rank order.
#include <cooperative_groups.h>

#include "acc_device.cuh"
#include "exact_acc.h"
#include "iter_exact.h"

namespace cg = cooperative_groups;

#define IX_THREADS 256
#define IX_WARPS (IX_THREADS / 32)
constexpr int IX_TILE_ITEMS = 8;
constexpr int IX_TILE = IX_THREADS * IX_TILE_ITEMS;

struct ScalarAccX {
    double v, e;
    __host__ __device__ double val(int) const { return v; }
    __host__ __device__ double err(int) const { return e; }
};
__device__ __forceinline__ bool ix_conv(double V, double e, const IterXParams& p) {
    return cub_converged(1, ScalarAccX{V, e}, p.absTol, p.relTol, p.norm);
}
__device__ __forceinline__ unsigned long long ix_bits(double x) { return (unsigned long long)__double_as_longlong(x); }
__device__ __forceinline__ unsigned long long ix_key_of_err(double e) { return ix_bits(e > 0.0 ? e : 0.0); }  // Rule.java:281-289

// ---- peer exchange ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ix_st_release_sys(unsigned long long* ptr, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ix_ld_acquire_sys(const unsigned long long* ptr) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ix_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long* ix_pay(unsigned long long* buf, unsigned long long seq, int src_rank) {
    return buf + PK_XB_DATA + ((size_t)(seq & 1ULL) * PK_MAX_RANKS + (size_t)src_rank) * PK_XB_PAY;
}
__device__ __forceinline__ unsigned long long* ix_flag(unsigned long long* buf, unsigned long long seq, int src_rank, int slice) {
    return buf + PK_XB_FLAGS + ((size_t)(seq & 1ULL) * PK_MAX_RANKS + (size_t)src_rank) * PK_XB_SLICES + (size_t)slice;
}
// All IX_THREADS threads of ONE CTA: push payload words [w0, w1) of this rank (src is indexed by payload word) into slot
// [parity][rank] of every rank's buffer, raise flag `slice`, wait for the same slice of every rank.  Afterwards the words
// [w0, w1) of all R payloads in the own buffer are valid for this CTA.  Returns 0, or LE_* when a peer timed out / aborted.
// seq counts the exchanges of the communicator's life; a rank can be at most one exchange ahead of its slowest peer, so two
// payload slots (parity) suffice.
__device__ int ix_exchange(const IterXParams& p, unsigned long long seq, int slice, const unsigned long long* src, int w0, int w1) {
    __shared__ int s_fail;
    const int tid = threadIdx.x;
    unsigned long long* own = p.xbuf[p.rank];
    if (tid == 0) s_fail = 0;
    for (int r = 0; r < p.R; ++r) {
        unsigned long long* dst = ix_pay(p.xbuf[r], seq, p.rank);
        for (int w = w0 + tid; w < w1; w += IX_THREADS) dst[w] = src[w];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < p.R) {
        ix_st_release_sys(ix_flag(p.xbuf[tid], seq, p.rank, slice), seq);
        const unsigned long long* f = ix_flag(own, seq, tid, slice);
        const unsigned long long t0 = ix_timer_ns();
        unsigned spins = 0;
        while (ix_ld_acquire_sys(f) < seq) {
            if ((++spins & 63u) == 0u) {
                if (ix_ld_acquire_sys(own + PK_XB_ABORT)) {
                    s_fail = LE_PEER_ABORT;
                    break;
                }
                if (ix_timer_ns() - t0 > PK_XCHG_TIMEOUT_NS) {
                    s_fail = LE_PEER_TIMEOUT;
                    break;
                }
            }
        }
    }
    __syncthreads();
    __threadfence_system();
    return s_fail;
}

// ---- decide ------------------------------------------------------------------------------------------------------------
// Folds the deltas the rule kernels left (acc_layout.h) into the totals, takes the loop decisions of Rule.java:72-79 and
// the pop-everything shortcut; when a batch has to be cut it hands the totals to the select kernel.  One CTA; the
// arbitrary-precision arithmetic (a handful of limbs) is one thread's work.
struct DecideShared {
    XaBig E, V, P, Q;
    long long dcell[XA_CELLS];
    long long far_e[XA_LIMBS], far_v[XA_LIMBS];
    unsigned long long hdr[PK_MAX_RANKS][16];
    int fail;
};

__device__ __forceinline__ void ix_big_add_signed(XaBig* pos_part, XaBig* neg_part, long long v, int pos) { xa_add_signed(pos_part, neg_part, v, pos); }
// a -= v * 2^pos for a normalized a >= v * 2^pos, v < 2^96
__device__ void ix_big_sub(XaBig* a, xa_u128 v, int pos) {
    const int j = pos >> 5, sh = pos & 31;
    const xa_u128 lo = v << sh;
    long long borrow = 0;
    for (int q = 0; q < 4 || borrow; ++q) {
        if (j + q >= XA_LIMBS) break;
        const long long piece = q < 4 ? (long long)(xa_u64)((lo >> (32 * q)) & 0xffffffffULL) : 0LL;
        long long t = (long long)a->l[j + q] - piece - borrow;
        borrow = 0;
        if (t < 0) {
            t += (1LL << 32);
            borrow = 1;
        }
        a->l[j + q] = (xa_u64)t;
    }
}
__device__ void ix_big_full_range(XaBig* a) {
    a->jlo = 0;
    a->jhi = XA_LIMBS - 1;
}
__device__ void ix_big_tighten(XaBig* a) {  // normalized number, zero outside [jlo, jhi]: shrink the range to the non-zero limbs
    int lo = a->jlo, hi = a->jhi;
    while (hi > lo && a->l[hi] == 0ULL) --hi;
    while (lo < hi && a->l[lo] == 0ULL) ++lo;
    a->jlo = lo;
    a->jhi = hi + 2 < XA_LIMBS ? hi + 2 : XA_LIMBS - 1;  // head room for carries of later additions
}

__device__ void ix_finish(const IterXParams& p, int state, int err) {
    p.ctl[LC_STAMP + 3] = (long long)ix_timer_ns();
    p.ctl[LC_STATE] = state;
    p.ctl[LC_ERR] = err;
    p.ctl[LC_DYN_N] = 0;
    p.ctl[LC_SELECT] = 0;
}

// layout of the totals the decide kernel publishes for the select kernel (p.gbins): E limbs [XA_LIMBS] | jlo | jhi
#define IX_PUB_JLO XA_LIMBS
#define IX_PUB_JHI (XA_LIMBS + 1)
#define IX_PUB_TOPE (XA_LIMBS + 2)

// Two builds of the same code: <4> for large pools (phase 1 wants occupancy; phase 2's single thread spills, which does not
// matter next to a millisecond of streaming) and <1> for small ones (no spills: phase 2 is the latency of the iteration).
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(IX_THREADS, MIN_BLOCKS) k_exact_decide(IterXParams p) {
    __shared__ DecideShared S;
    __shared__ CubAccWin<IX_THREADS> win;
    __shared__ int s_last;
    const int tid = threadIdx.x;
    long long* ctl = p.ctl;
    if (ctl[LC_STATE] != LS_RUN) {  // finished or paused: this launch and the ones queued behind it do nothing
        if (blockIdx.x == 0 && tid == 0) {
            ctl[LC_DYN_N] = 0;
            ctl[LC_SELECT] = 0;
        }
        return;
    }
    if (blockIdx.x == 0 && tid == 0) ctl[LC_STAMP + 4] = (long long)ix_timer_ns();
    // ---- phase 1 (every CTA): what happened to the pool since the last decision ---------------------------------------
    // the children of the last batch enter the sums (item j of the rule launch: slot arithmetic of CUB_MODE_BISECT), the
    // popped parents leave them (their val / err were copied by the select kernel before the left children overwrote them)
    {
        const long long n_add = ctl[LC_DYN_N], use_list = ctl[LC_DYN_LIST], base = ctl[LC_DYN_BASE], n_sub = ctl[LC_SUB_K], n_readd = ctl[LC_READD];
        const long long n_pairs = n_add >> 1;  // (left, right) children of one parent; an odd item is the root region
        const long long first = (long long)blockIdx.x * IX_THREADS, stride = (long long)gridDim.x * IX_THREADS;
        if (first < n_pairs + (n_add & 1) || first < n_sub || first < n_readd) {
            cub_acc_begin(p.acc, &win);
            CubAccKeys keys;
            cub_acc_keys_init(keys);
            const double* __restrict__ val = p.pool.val;
            const double* __restrict__ err = p.pool.err;
            const double* __restrict__ rval = val + base;
            const double* __restrict__ rerr = err + base;
            for (long long i0 = first + tid; i0 < n_pairs; i0 += 2 * stride) {
                const long long i1 = i0 + stride;
                const bool two = i1 < n_pairs;
                const long long l0 = use_list ? (long long)p.list[i0] : i0, l1 = two ? (use_list ? (long long)p.list[i1] : i1) : l0;
                const double v0 = val[l0], e0 = err[l0], v1 = rval[i0], e1 = rerr[i0];
                double v2 = 0.0, e2 = 0.0, v3 = 0.0, e3 = 0.0;
                if (two) {
                    v2 = val[l1];
                    e2 = err[l1];
                    v3 = rval[i1];
                    e3 = rerr[i1];
                }
                cub_acc_region(&win, p.acc, keys, v0, e0, 1);
                cub_acc_region(&win, p.acc, keys, v1, e1, 1);
                if (two) {
                    cub_acc_region(&win, p.acc, keys, v2, e2, 1);
                    cub_acc_region(&win, p.acc, keys, v3, e3, 1);
                }
            }
            if ((n_add & 1) && first + tid == n_pairs) {
                const long long l0 = use_list ? (long long)p.list[n_pairs] : n_pairs;
                cub_acc_region(&win, p.acc, keys, val[l0], err[l0], 1);
            }
            for (long long i0 = first + tid; i0 < n_sub; i0 += 2 * stride) {
                const long long i1 = i0 + stride;
                const bool two = i1 < n_sub;
                const double v0 = p.stash_val[i0], e0 = p.stash_err[i0];
                const double v1 = two ? p.stash_val[i1] : 0.0, e1 = two ? p.stash_err[i1] : 0.0;
                cub_acc_region(&win, p.acc, keys, v0, e0, -1);
                if (two) cub_acc_region(&win, p.acc, keys, v1, e1, -1);
            }
            // a selection that popped almost everything has reset the sums: the few regions it kept enter them again
            for (long long i0 = first + tid; i0 < n_readd; i0 += stride) {
                const long long sl = (long long)p.kept[i0];
                cub_acc_region(&win, p.acc, keys, val[sl], err[sl], 1);
            }
            cub_acc_keys_store(&win, keys);
            cub_acc_end(p.acc, &win);
        }
    }
    // ---- the CTA that arrives last goes on (the others are done) ----------------------------------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned long long prev = atomicAdd(&p.bar[1], 1ULL);
        s_last = prev + 1ULL == (unsigned long long)gridDim.x ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) {
        p.bar[1] = 0ULL;
        ctl[LC_SUB_K] = 0;
        ctl[LC_READD] = 0;
        ctl[LC_STAMP] = (long long)ix_timer_ns();
    }
    const long long N = ctl[LC_N], Nglobal = ctl[LC_NGLOBAL], numEval = ctl[LC_NUMEVAL], maxEval = ctl[LC_MAXEVAL], P = ctl[LC_P];
    // (everything the decisions below read, issued together: one L2 round trip instead of one per branch)
    const long long c_stop_at = ctl[LC_STOP_AT], c_cap = ctl[LC_CAP], c_iter = ctl[LC_ITER], c_maxiter = ctl[LC_MAXITER], c_rebal_min = ctl[LC_REBAL_MIN];
    const bool budget_left = maxEval == 0 || numEval < maxEval;  // Rule.java:72
    // pops the evaluation budget allows (checked after every pop, Rule.java:133)
    long long Kcap = Nglobal;
    if (maxEval > 0 && budget_left) {
        const long long room = maxEval - numEval, k = (room + 2 * P - 1) / (2 * P);
        Kcap = k < 1 ? 1 : k;
        if (Kcap > Nglobal) Kcap = Nglobal;
    }
    unsigned long long* acc = p.acc;
    unsigned long long* tot = acc + XA_TOT_OFF;
    // ---- collect the deltas (and clear them for the next launch) -----------------------------------------------------
    if (tid < XA_CELLS) {
        // (sixteen loads in flight, then the stores: interleaved, every load waited for the store before it -- a quarter of
        // this kernel's latency on small pools)
        long long s = 0;
#pragma unroll 1
        for (int r0 = 0; r0 < XA_REPL; r0 += 16) {
            unsigned long long v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = __ldcg(acc + XA_D_OFF + (size_t)(r0 + r) * XA_CELLS + tid);  // written by the other CTAs' atomics: read at L2
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                s += (long long)v[r];
                if (v[r]) acc[XA_D_OFF + (size_t)(r0 + r) * XA_CELLS + tid] = 0ULL;
            }
        }
        S.dcell[tid] = s;
    }
    for (int j = tid; j < XA_LIMBS; j += IX_THREADS) {
        S.far_e[j] = (long long)__ldcg(&acc[XA_FAR_E_OFF + j]);
        S.far_v[j] = (long long)__ldcg(&acc[XA_FAR_V_OFF + j]);
        acc[XA_FAR_E_OFF + j] = 0ULL;
        acc[XA_FAR_V_OFF + j] = 0ULL;
        S.E.l[j] = tot[XA_TOT_E + j];
        S.V.l[j] = tot[XA_TOT_V + j];
        S.P.l[j] = 0ULL;
        S.Q.l[j] = 0ULL;
    }
    bool far_mine = false;
    for (int j = tid; j < XA_LIMBS; j += IX_THREADS) far_mine = far_mine || S.far_e[j] != 0 || S.far_v[j] != 0;
    const bool far_any = __syncthreads_or(far_mine ? 1 : 0) != 0;
    unsigned long long misc[8], my_misc[8];
    unsigned long long minkey = 0, my_count = 0, topkey = 0;
    int v_neg = 0;
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) misc[q] = my_misc[q] = __ldcg(&acc[XA_M_OFF + q]);
        minkey = __ldcg(&acc[XA_M_OFF + XA_M_MINKEY]);
        topkey = __ldcg(&acc[XA_M_OFF + XA_M_TOPKEY]);
        const int ebase = (int)__ldcg(&acc[XA_M_OFF + XA_M_EBASE]), vbase = (int)__ldcg(&acc[XA_M_OFF + XA_M_VBASE]);
        v_neg = (int)tot[XA_TOT_MISC + 0];
        my_count = tot[XA_TOT_MISC + 1] + (unsigned long long)S.dcell[XA_C_COUNT];
        // limbs that can be touched: the stored totals, the delta window (256 bits + 64-bit cells), carries
        int e_lo = (int)tot[XA_TOT_MISC + 2], e_hi = (int)tot[XA_TOT_MISC + 3], v_lo = (int)tot[XA_TOT_MISC + 4], v_hi = (int)tot[XA_TOT_MISC + 5];
        e_lo = min(e_lo, ebase >> 5);
        e_hi = max(e_hi, ((ebase + 32 * XA_HALVES + 64) >> 5) + 2);
        v_lo = min(v_lo, vbase >> 5);
        v_hi = max(v_hi, ((vbase + 32 * XA_HALVES + 64) >> 5) + 2);
        if (far_any) {
            e_lo = v_lo = 0;
            e_hi = v_hi = XA_LIMBS - 1;
        }
        e_hi = min(e_hi, XA_LIMBS - 1);
        v_hi = min(v_hi, XA_LIMBS - 1);
        // error sum: E += (positive part of the delta) - (negative part)
        S.P.jlo = S.Q.jlo = S.E.jlo = e_lo;
        S.P.jhi = S.Q.jhi = S.E.jhi = e_hi;
        xa_fold_cells(&S.P, &S.Q, &S.dcell[XA_C_E], ebase);
        if (far_any)
            for (int j = 0; j < XA_LIMBS - 3; ++j) ix_big_add_signed(&S.P, &S.Q, S.far_e[j], 32 * j);
        for (int j = e_lo; j <= e_hi; ++j) S.E.l[j] += S.P.l[j];
        S.E.normalize();
        S.Q.normalize();
        S.E.sub(S.Q);  // the regions removed were added before: never negative
        ix_big_tighten(&S.E);
        // value sum, sign and magnitude: A = (V if V >= 0) + positive part, B = (|V| if V < 0) + negative part, V = A - B
        for (int j = e_lo; j <= e_hi; ++j) S.P.l[j] = S.Q.l[j] = 0ULL;
        S.P.jlo = S.Q.jlo = v_lo;
        S.P.jhi = S.Q.jhi = v_hi;
        xa_fold_cells(&S.P, &S.Q, &S.dcell[XA_C_V], vbase);
        if (far_any)
            for (int j = 0; j < XA_LIMBS - 3; ++j) ix_big_add_signed(&S.P, &S.Q, S.far_v[j], 32 * j);
        XaBig* same = v_neg ? &S.Q : &S.P;
        for (int j = v_lo; j <= v_hi; ++j) same->l[j] += S.V.l[j];
        S.P.normalize();
        S.Q.normalize();
        const int c = S.P.cmp(S.Q);
        if (c >= 0) {
            S.P.sub(S.Q);
            for (int j = v_lo; j <= v_hi; ++j) S.V.l[j] = S.P.l[j];
            v_neg = 0;
        } else {
            S.Q.sub(S.P);
            for (int j = v_lo; j <= v_hi; ++j) S.V.l[j] = S.Q.l[j];
            v_neg = 1;
        }
        S.V.jlo = v_lo;
        S.V.jhi = v_hi;
        ix_big_tighten(&S.V);
        // store this rank's totals
        tot[XA_TOT_MISC + 0] = (unsigned long long)v_neg;
        tot[XA_TOT_MISC + 1] = my_count;
        tot[XA_TOT_MISC + 2] = (unsigned long long)S.E.jlo;
        tot[XA_TOT_MISC + 3] = (unsigned long long)S.E.jhi;
        tot[XA_TOT_MISC + 4] = (unsigned long long)S.V.jlo;
        tot[XA_TOT_MISC + 5] = (unsigned long long)S.V.jhi;
        // window hints for the next rule launch from the largest values of the last one
        const unsigned long long ek = __ldcg(&acc[XA_M_OFF + XA_M_EMAXKEY]), vk = __ldcg(&acc[XA_M_OFF + XA_M_VMAXKEY]);
        if (ek) {
            const int b = xa_lsb_pos((int)(ek >> 52)) + 8 - XA_WIN_MAX_SHIFT;
            acc[XA_M_OFF + XA_M_EBASE] = (unsigned long long)(b < 0 ? 0 : b);
            acc[XA_M_OFF + XA_M_EMAXKEY] = 0ULL;
        }
        if (vk) {
            const int b = xa_lsb_pos((int)(vk >> 52)) + 8 - XA_WIN_MAX_SHIFT;
            acc[XA_M_OFF + XA_M_VBASE] = (unsigned long long)(b < 0 ? 0 : b);
            acc[XA_M_OFF + XA_M_VMAXKEY] = 0ULL;
        }
        ctl[LC_STAMP + 1] = (long long)ix_timer_ns();
    }
    __syncthreads();
    for (int j = tid; j < XA_LIMBS; j += IX_THREADS) {
        tot[XA_TOT_E + j] = S.E.l[j];
        tot[XA_TOT_V + j] = S.V.l[j];
    }
    __syncthreads();  // (thread 0 may reset the totals below)
    if (budget_left) {
        // pauses are taken before any exchange, so that a resumed launch starts the iteration from scratch on every rank (the
        // sums are up to date: a resumed launch finds nothing left to fold)
        if (c_stop_at > 0 && Nglobal >= c_stop_at) {
            if (tid == 0) ix_finish(p, LS_HANDOFF, 0);
            return;
        }
        const long long kmax = N < Kcap ? N : Kcap;
        if (N + kmax > c_cap) {
            if (tid == 0) ix_finish(p, LS_GROW, 0);
            return;
        }
    }
    unsigned long long Ntot = my_count;
    if (p.R > 1) {
        // ---- exchange the ranks' totals: header[16] | E limbs | V limbs; every rank adds them up -------------------
        unsigned long long* stage = p.gbins + 256;
        if (tid == 0) {
            stage[0] = my_count;
            stage[1] = minkey;
            for (int q = 0; q < 7; ++q) stage[2 + q] = misc[q];
            stage[9] = (unsigned long long)v_neg;
            stage[10] = topkey;
            for (int q = 11; q < 16; ++q) stage[q] = 0ULL;
        }
        for (int j = tid; j < XA_LIMBS; j += IX_THREADS) {
            stage[16 + j] = S.E.l[j];
            stage[16 + XA_LIMBS + j] = S.V.l[j];
        }
        __syncthreads();
        const unsigned long long seq = (unsigned long long)ctl[LC_SEQ];
        const int fail = ix_exchange(p, seq, 0, stage, 0, 16 + 2 * XA_LIMBS);
        if (tid == 0) ctl[LC_SEQ] = (long long)(seq + 1ULL);
        if (fail) {
            if (tid == 0) ix_finish(p, LS_ERROR, fail);
            return;
        }
        unsigned long long* own = p.xbuf[p.rank];
        if (tid < 16)
            for (int r = 0; r < p.R; ++r) S.hdr[r][tid] = __ldcg(&ix_pay(own, seq, r)[tid]);
        // limb-wise sums over the ranks (not normalized): E, and V split by sign
        for (int j = tid; j < XA_LIMBS; j += IX_THREADS) {
            unsigned long long e = 0, vp = 0, vn = 0;
            for (int r = 0; r < p.R; ++r) {
                const unsigned long long* pay = ix_pay(own, seq, r);
                e += __ldcg(&pay[16 + j]);
                const unsigned long long v = __ldcg(&pay[16 + XA_LIMBS + j]);
                if (__ldcg(&pay[9]))
                    vn += v;
                else
                    vp += v;
            }
            S.E.l[j] = e;
            S.P.l[j] = vp;
            S.Q.l[j] = vn;
        }
        __syncthreads();
        if (tid == 0) {
            for (int q = 0; q < 8; ++q) misc[q] = 0ULL;
            minkey = ~0ULL;
            Ntot = 0;
            for (int r = 0; r < p.R; ++r) {
                Ntot += S.hdr[r][0];
                ctl[LC_RANKN + r] = (long long)S.hdr[r][0];
                minkey = S.hdr[r][1] < minkey ? S.hdr[r][1] : minkey;
                topkey = S.hdr[r][10] > topkey ? S.hdr[r][10] : topkey;
                for (int q = 0; q < 5; ++q) misc[q] += S.hdr[r][2 + q];
                misc[XA_M_ESTICKY] |= S.hdr[r][2 + XA_M_ESTICKY];
                misc[XA_M_VSTICKY] |= S.hdr[r][2 + XA_M_VSTICKY];
            }
            ix_big_full_range(&S.E);
            S.E.normalize();
            ix_big_tighten(&S.E);
            ix_big_full_range(&S.P);
            ix_big_full_range(&S.Q);
            S.P.normalize();
            S.Q.normalize();
            if (S.P.cmp(S.Q) >= 0) {
                S.P.sub(S.Q);
                for (int j = 0; j < XA_LIMBS; ++j) S.V.l[j] = S.P.l[j];
                v_neg = 0;
            } else {
                S.Q.sub(S.P);
                for (int j = 0; j < XA_LIMBS; ++j) S.V.l[j] = S.Q.l[j];
                v_neg = 1;
            }
            ix_big_full_range(&S.V);
            ix_big_tighten(&S.V);
        }
        __syncthreads();
    }
    // the partition's error sum for the select kernel
    for (int j = tid; j < XA_LIMBS; j += IX_THREADS) p.gbins[j] = S.E.l[j];
    if (tid != 0) return;  // the rest is one thread's work
    p.gbins[IX_PUB_JLO] = (unsigned long long)S.E.jlo;
    p.gbins[IX_PUB_JHI] = (unsigned long long)S.E.jhi;
    p.gbins[IX_PUB_TOPE] = topkey >> 52;  // no region of the partition has a larger binary exponent of errmax

    ctl[LC_STAMP + 2] = (long long)ix_timer_ns();
    // ---- totals: E truncated (exact `<` tests), V rounded to nearest -------------------------------------------------
    const double inf = __longlong_as_double(0x7ff0000000000000LL), nan = __longlong_as_double(0x7ff8000000000000LL);
    const bool e_nonfinite = misc[XA_M_ESTICKY] != 0ULL || misc[XA_M_ENAN] != 0ULL || misc[XA_M_EINF] != 0ULL;
    double E = S.E.to_double(0), E_report = S.E.to_double(1);  // decisions use the truncated sum, the caller gets the nearest double
    if (e_nonfinite) {
        E = ((misc[XA_M_ESTICKY] & 2ULL) || misc[XA_M_ENAN]) ? nan : inf;
        E_report = E;
    }
    double V = S.V.to_double(1);
    if (v_neg) V = -V;
    if (misc[XA_M_VSTICKY] | misc[XA_M_VNAN] | misc[XA_M_VPINF] | misc[XA_M_VNINF]) {
        if ((misc[XA_M_VSTICKY] & 2ULL) || misc[XA_M_VNAN] || (misc[XA_M_VPINF] && misc[XA_M_VNINF]))
            V = nan;
        else if (misc[XA_M_VPINF])
            V = inf;
        else if (misc[XA_M_VNINF])
            V = -inf;
    }
    const bool count_only = e_nonfinite;
    const double eps = 2.220446049250313e-16;
    const double band = 4.0 * eps * sqrt(3.0 * (double)Nglobal);  // rounding band of the reference's drifting sums (diagnostic)
    ctl[LC_V] = (long long)ix_bits(V);
    ctl[LC_E] = (long long)ix_bits(E_report);
    if ((long long)Ntot != Nglobal) {  // the sums lost track of the pool: never decide on them
        ctl[LC_DBG0] = (long long)Ntot;
        ix_finish(p, LS_ERROR, LE_INTERNAL);
        return;
    }
    if (!budget_left) {
        ix_finish(p, LS_MAXEVAL, 0);
        return;
    }
    if (c_iter >= c_maxiter) {
        ix_finish(p, LS_GUARD, 0);
        return;
    }
    const bool c0 = !count_only && ix_conv(V, E, p);  // Rule.java:73
    if (!count_only) {
        const bool c1 = ix_conv(V, E * (1.0 + band), p), c2 = ix_conv(V, E * (1.0 - band), p);
        if (c0 != c1 || c0 != c2) ctl[LC_AMBIG] += 1;
    }
    if (c0) {
        ix_finish(p, LS_CONVERGED, 0);
        return;
    }
    // ---- sharded pool: the loop pauses for a rebalance when the fullest shard exceeds 1.25 x the mean.  Every
    // rank sees the same counts, so every rank pauses at the same iteration; the host moves the surplus (cubature_cuda.cpp:
    // rebalance_shards), rebuilds the sums and resumes -- this launch has folded everything, a resumed one starts from scratch
    if (p.R > 1 && c_rebal_min > 0 && Nglobal >= c_rebal_min) {
        unsigned long long nmax = 0;
        for (int r = 0; r < p.R; ++r) nmax = S.hdr[r][0] > nmax ? S.hdr[r][0] : nmax;
        if (4ULL * nmax * (unsigned long long)p.R > 5ULL * Ntot) {
            ix_finish(p, LS_REBALANCE, 0);
            return;
        }
    }
    // ---- the whole pool is popped when the residual left by the smallest region alone does not pass (monotone) -------
    const double err_min = __longlong_as_double((long long)minkey);
    bool all = count_only || minkey == ~0ULL || !ix_conv(V, err_min, p);
    if (!count_only && minkey != ~0ULL) {
        const bool a1 = !ix_conv(V, err_min + band * E, p), a2 = !ix_conv(V, err_min - band * E, p);
        if (a1 != all || a2 != all) ctl[LC_AMBIG] += 1;
    }
    if (Nglobal == 1) all = true;
    all = all && Kcap >= Nglobal;
    if (all) {
        ctl[LC_DYN_N] = 2 * N;
        ctl[LC_DYN_LIST] = 0;
        ctl[LC_DYN_BASE] = N;
        ctl[LC_SELECT] = 0;
        ctl[LC_K] = N;
        ctl[LC_KGLOBAL] = Nglobal;
        ctl[LC_N] = 2 * N;
        ctl[LC_NGLOBAL] = 2 * Nglobal;
        ctl[LC_NUMEVAL] = numEval + 2 * P * Nglobal;
        ctl[LC_EVAL_LOCAL] += 2 * N;
        ctl[LC_WORDS + (ctl[LC_NBATCH] % LC_LOG_CAP)] = 2 * Nglobal;
        ctl[LC_NBATCH] += 1;
        ctl[LC_ITER] = c_iter + 1;
        // every region leaves the pool (one RegionHeap.poll per region): this rank's sums restart from zero; a non-finite
        // value that leaves turns the reference's running sum into NaN for good (Inf - Inf)
        acc[XA_M_OFF + XA_M_MINKEY] = ~0ULL;
        for (int j = 0; j < 2 * XA_LIMBS; ++j) tot[j] = 0ULL;
        tot[XA_TOT_MISC + 0] = 0ULL;
        tot[XA_TOT_MISC + 1] = 0ULL;
        tot[XA_TOT_MISC + 2] = tot[XA_TOT_MISC + 4] = (unsigned long long)(XA_LIMBS - 1);
        tot[XA_TOT_MISC + 3] = tot[XA_TOT_MISC + 5] = 0ULL;
        if (my_misc[XA_M_ENAN] | my_misc[XA_M_EINF]) acc[XA_M_OFF + XA_M_ESTICKY] = my_misc[XA_M_ESTICKY] | 2ULL;
        if (my_misc[XA_M_VNAN] | my_misc[XA_M_VPINF] | my_misc[XA_M_VNINF]) acc[XA_M_OFF + XA_M_VSTICKY] = my_misc[XA_M_VSTICKY] | 2ULL;
        for (int q = 0; q < 5; ++q) acc[XA_M_OFF + q] = 0ULL;
        ctl[LC_STAMP + 3] = (long long)ix_timer_ns();
        return;
    }
    // ---- a batch has to be cut: the select kernel finds the threshold ------------------------------------------------
    SelX* sx = p.selx;
    sx->count_only = count_only ? 1 : 0;
    sx->Kcap = (unsigned long long)Kcap;
    sx->V = V;
    sx->E = E;
    sx->band = band;
    sx->thr_key = 0ULL;
    sx->tie_take = 0ULL;
    sx->K = 0ULL;
    *p.bar = 0ULL;  // grid barrier of the select kernel
    ctl[LC_SELECT] = 1;
    ctl[LC_DYN_N] = 0;  // the select kernel sizes the batch
    ctl[LC_STAMP + 3] = (long long)ix_timer_ns();
}

// ---- select ------------------------------------------------------------------------------------------------------------
// Radix select inside the threshold binade on the 52 fraction bits, most significant first: two levels of 8 bits that
// stream the rank's chunk of the pool (256 buckets: four private copies per CTA keep the shared-memory atomics apart, and a
// CTA adds at most 256 x 3 words to the global histogram), then four levels of 9 bits that only touch the candidates the
// second scan extracted (the regions that match the 27 bits decided so far, binade included).
struct PickShared {
    unsigned long long frac;      // decided fraction bits of the threshold key
    unsigned long long Pm_lo, Pm_hi;  // mantissa sum of the regions popped inside the binade so far
    unsigned long long Cc;        // regions popped so far (higher binades included)
    unsigned long long cb, Sb_lo, Sb_hi;  // chosen bucket: count, mantissa sum
    unsigned long long tie_take, K, thr_key, ambiguous;
};
__device__ __forceinline__ int ix_level_bits(int level) { return level < 2 ? 8 : 9; }
__device__ __forceinline__ int ix_level_shift(int level) { return level < 2 ? 44 - 8 * level : 27 - 9 * (level - 2); }
__device__ __forceinline__ int ix_level_nb(int level) { return 1 << ix_level_bits(level); }

__device__ __forceinline__ double ix_residual(const SelX& sx, xa_u128 H, xa_u128 popped) {
    const xa_u128 D = H - popped;  // >= 0: the popped regions are part of H
    if (D == 0) return sx.Ldbl;
    return xa_trunc_hi(D, sx.L1, sx.pos);
}

// barrier over the first n_ctas CTAs of a cooperative launch (all resident); *bar counts arrivals since the decide kernel
// cleared it, target = (barriers so far) * n_ctas
__device__ __forceinline__ void ix_gsync(unsigned long long* bar, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1ULL);
        unsigned long long v;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// Block-wide pick of one level (all threads): buckets in descending order of the mantissa; the first bucket at which the
// budget runs out or the exact residual passes.  gh: this level's histogram of the partition (global memory:
// cnt[IX_NB] | low words [IX_NB] | high words [IX_NB]).  Returns 1 when the batch size is final (last level, or the chosen
// bucket holds one region).
__device__ int ix_pick(const IterXParams& p, const SelX& sx, PickShared* ps, const unsigned long long* gh, int level,
                       unsigned long long* shc, unsigned long long* shl, unsigned long long* shh) {
    const int tid = threadIdx.x;
    const int NB = ix_level_nb(level), shift = ix_level_shift(level);
    const int per_t = NB / IX_THREADS;  // 1 or 2
    const unsigned long long implicit = sx.binade ? (1ULL << 52) : 0ULL;
    const unsigned long long frac = ps->frac, Cc = ps->Cc;
    const xa_u128 Pm = ((xa_u128)ps->Pm_hi << 64) | ps->Pm_lo;
    const xa_u128 H = ((xa_u128)sx.H_hi << 64) | sx.H_lo;
    unsigned long long c_in[2], cl[2];
    xa_u128 s_in[2], sl[2];
    unsigned long long ct = 0;
    xa_u128 st = 0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        c_in[j] = 0ULL;
        s_in[j] = 0;
        if (j < per_t) {
            const int jb = NB - 1 - (tid * per_t + j);
            const unsigned long long c = __ldcg(&gh[jb]), r_lo = __ldcg(&gh[IX_NB + jb]), r_hi = __ldcg(&gh[2 * IX_NB + jb]);  // (together)
            c_in[j] = c;
            if (c) {
                const unsigned long long base = implicit | frac | ((unsigned long long)jb << shift);
                const xa_u128 r = (xa_u128)r_lo + ((xa_u128)r_hi << 32);
                s_in[j] = (xa_u128)c * base + r;
                ct += c;
                st += s_in[j];
            }
        }
        cl[j] = ct;
        sl[j] = st;
    }
    __syncthreads();  // everybody has read *ps
    // inclusive scan over the threads: shuffles inside a warp, the warps' totals through shared memory
    const unsigned lane = tid & 31u, warp = tid >> 5;
    unsigned long long sc = ct, sl_ = (unsigned long long)st, sh_ = (unsigned long long)(st >> 64);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long ta = __shfl_up_sync(0xffffffffu, sc, o), tl = __shfl_up_sync(0xffffffffu, sl_, o), th = __shfl_up_sync(0xffffffffu, sh_, o);
        if ((int)lane >= o) {
            sc += ta;
            const unsigned long long nl = sl_ + tl;
            sh_ += th + (nl < tl ? 1ULL : 0ULL);
            sl_ = nl;
        }
    }
    if (lane == 31u) {
        shc[warp] = sc;
        shl[warp] = sl_;
        shh[warp] = sh_;
    }
    __syncthreads();
    for (unsigned w = 0; w < warp; ++w) {
        sc += shc[w];
        const unsigned long long tl = shl[w], nl = sl_ + tl;
        sh_ += shh[w] + (nl < tl ? 1ULL : 0ULL);
        sl_ = nl;
    }
    const unsigned long long c_off = sc - ct;
    const xa_u128 s_off = (((xa_u128)sh_ << 64) | sl_) - st;
    unsigned first_stop = ~0u, last_nz = 0u;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (j < per_t && c_in[j]) {
            const unsigned q = (unsigned)(tid * per_t + j);
            last_nz = q + 1u;
            const unsigned long long Cincl = Cc + c_off + cl[j];
            bool stop = Cincl >= sx.Kcap;
            if (!stop && !sx.count_only) stop = ix_conv(sx.V, ix_residual(sx, H, Pm + s_off + sl[j]), p);
            if (stop && first_stop == ~0u) first_stop = q;
        }
    }
    first_stop = __reduce_min_sync(0xffffffffu, first_stop);
    last_nz = __reduce_max_sync(0xffffffffu, last_nz);
    __syncthreads();  // the warps' scan totals have been read
    if (lane == 0u) {
        shc[warp] = first_stop;
        shl[warp] = last_nz;
    }
    __syncthreads();
    unsigned long long any_stop = ~0ULL, nz = 0ULL;
    {
        unsigned fs = ~0u, ln = 0u;
        for (int w = 0; w < IX_WARPS; ++w) {
            fs = (unsigned)shc[w] < fs ? (unsigned)shc[w] : fs;
            ln = (unsigned)shl[w] > ln ? (unsigned)shl[w] : ln;
        }
        any_stop = fs == ~0u ? ~0ULL : (unsigned long long)fs;
        nz = ln;
    }
    __syncthreads();
    if (nz == 0ULL) {  // cannot happen: the binade holds at least one region
        if (tid == 0) {
            ps->K = 0ULL;
            ps->tie_take = 0ULL;
            ps->thr_key = ~0ULL;
        }
        __syncthreads();
        return 1;
    }
    const int chosen = any_stop != ~0ULL ? (int)any_stop : (int)(nz - 1ULL);
    if (chosen >= tid * per_t && chosen < tid * per_t + per_t) {  // the owner publishes the descent
        const int j = chosen - tid * per_t;
        const unsigned long long c_prev = j > 0 ? cl[0] : 0ULL;
        const xa_u128 s_prev = j > 0 ? sl[0] : (xa_u128)0;
        const unsigned long long cb = j > 0 ? c_in[1] : c_in[0];
        const xa_u128 Sb = j > 0 ? s_in[1] : s_in[0];
        const int jb = NB - 1 - chosen;
        const xa_u128 Pn = Pm + s_off + s_prev;
        ps->frac = frac | ((unsigned long long)jb << shift);
        ps->Cc = Cc + c_off + c_prev;
        ps->Pm_lo = (unsigned long long)Pn;
        ps->Pm_hi = (unsigned long long)(Pn >> 64);
        ps->cb = cb;
        ps->Sb_lo = (unsigned long long)Sb;
        ps->Sb_hi = (unsigned long long)(Sb >> 64);
    }
    __syncthreads();
    const unsigned long long cb = ps->cb;
    int fin = 0;
    if (level == IX_LEVELS - 1 || cb == 1ULL) {
        fin = 1;
        if (tid == 0) {
            // the bucket is a single key: cb regions with identical errmax (at the last level), or one region whose
            // mantissa is the bucket's sum
            const xa_u128 Sb = ((xa_u128)ps->Sb_hi << 64) | ps->Sb_lo;
            const unsigned long long m = cb == 1ULL ? (unsigned long long)Sb : (implicit | ps->frac);
            const xa_u128 Pn = ((xa_u128)ps->Pm_hi << 64) | ps->Pm_lo;
            const unsigned long long cc = ps->Cc;
            const unsigned long long room = sx.Kcap - cc;  // >= 1
            const unsigned long long mmax = cb < room ? cb : room;
            unsigned long long lo = 1, hi = mmax;  // smallest t in [1, mmax] whose residual passes (monotone), else mmax
            if (!sx.count_only && ix_conv(sx.V, ix_residual(sx, H, Pn + (xa_u128)mmax * m), p)) {
                while (lo < hi) {
                    const unsigned long long mid = lo + (hi - lo) / 2;
                    if (ix_conv(sx.V, ix_residual(sx, H, Pn + (xa_u128)mid * m), p))
                        hi = mid;
                    else
                        lo = mid + 1;
                }
            } else {
                lo = mmax;
            }
            ps->tie_take = lo;
            ps->K = cc + lo;
            ps->thr_key = ((unsigned long long)sx.binade << 52) | (m & 0xfffffffffffffULL);
            // sensitivity of this decision to the rounding the reference's running sums carry (diagnostic only)
            unsigned long long amb = 0ULL;
            if (!sx.count_only) {
                const double rK = ix_residual(sx, H, Pn + (xa_u128)lo * m), e1 = __longlong_as_double((long long)ps->thr_key), bb = sx.band * sx.E;
                const bool a0 = ix_conv(sx.V, rK + bb, p) != ix_conv(sx.V, rK - bb, p);
                const bool a1 = ix_conv(sx.V, rK + e1 + bb, p) != ix_conv(sx.V, rK + e1 - bb, p);
                amb = (a0 || a1) ? 1ULL : 0ULL;
            }
            ps->ambiguous = amb;
        }
    }
    __syncthreads();
    return fin;
}

// ---- shared-memory histograms of the select kernel ------------------------------------------------------------------
// hist[5][1024] 32-bit cells: plane 0 counts, planes 1..4 the 16-bit digits of the summed values (mantissas / remainders,
// < 2^53 each), bucket q of private copy cp at [plane * 1024 + cp * NB + q] (bucket-major for the binade histogram).  32-bit cells because shared memory has native
// 32-bit atomic adds only (64-bit ones are compare-and-swap loops: the binade histogram of 2e5 regions took 56 us with
// them).  A cell takes fewer than 2^16 additions of less than 2^16 each between two flushes (IX_ROUND elements per CTA).
#define IX_HPLANE 1024
#define IX_BW 64  // binary exponents per window of the binade walk
#define IX_ROUND 65536
// a thread's run of `run_c` equal digits with value sum `run` (< 2^63) -> histogram copy `h`
__device__ __forceinline__ void ix_hist_flush(unsigned* h, unsigned cur, unsigned run_c, unsigned long long run) {
    if (run_c) {
        atomicAdd(&h[cur], run_c);
        const unsigned p0 = (unsigned)(run & 0xffffULL), p1 = (unsigned)((run >> 16) & 0xffffULL), p2 = (unsigned)((run >> 32) & 0xffffULL),
                       p3 = (unsigned)(run >> 48);
        if (p0) atomicAdd(&h[IX_HPLANE + cur], p0);
        if (p1) atomicAdd(&h[2 * IX_HPLANE + cur], p1);
        if (p2) atomicAdd(&h[3 * IX_HPLANE + cur], p2);
        if (p3) atomicAdd(&h[4 * IX_HPLANE + cur], p3);
    }
}
__device__ __forceinline__ void ix_hist_clear(unsigned* hist) {
    for (int q = threadIdx.x; q < 5 * IX_HPLANE; q += IX_THREADS) hist[q] = 0u;
    __syncthreads();
}
// all threads: adds the CTA's histogram (NB buckets, `copies` private copies; bucket q of copy cp at cp * cs + q * qs) to the
// global one (count | low 32 bits | rest of the value sums, IX_NB words each) and clears it
__device__ void ix_hist_to_global(unsigned* hist, int NB, int copies, int cs, int qs, unsigned long long* gh) {
    __syncthreads();
    for (int q = threadIdx.x; q < NB; q += IX_THREADS) {
        unsigned c = 0;
        unsigned long long pc[4] = {0ULL, 0ULL, 0ULL, 0ULL};
        for (int cp = 0; cp < copies; ++cp) {
            const int at = cp * cs + q * qs;
            c += hist[at];
            hist[at] = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                pc[k] += hist[(1 + k) * IX_HPLANE + at];
                hist[(1 + k) * IX_HPLANE + at] = 0u;
            }
        }
        if (c) {
            const xa_u128 v = (xa_u128)pc[0] + ((xa_u128)pc[1] << 16) + ((xa_u128)pc[2] << 32) + ((xa_u128)pc[3] << 48);
            atomicAdd(&gh[q], (unsigned long long)c);
            const unsigned long long lo = (unsigned long long)(v & 0xffffffffULL), hi = (unsigned long long)(v >> 32);
            if (lo) atomicAdd(&gh[IX_NB + q], lo);
            if (hi) atomicAdd(&gh[2 * IX_NB + q], hi);
        }
    }
    __syncthreads();
}

#define IX_HCOPIES 4  // private histogram copies of the 256-bucket levels (one per pair of warps)

// block-wide sums of two counters (all threads; the sums come back to every thread): shuffles, then the warps' totals
__device__ __forceinline__ void ix_block_sum2(unsigned long long& a, unsigned long long& c, unsigned long long* sha, unsigned long long* shb) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    __syncthreads();  // earlier readers of sha / shb are done
    if ((threadIdx.x & 31u) == 0u) {
        sha[threadIdx.x >> 5] = a;
        shb[threadIdx.x >> 5] = c;
    }
    __syncthreads();
    a = 0ULL;
    c = 0ULL;
    for (int w = 0; w < IX_WARPS; ++w) {
        a += sha[w];
        c += shb[w];
    }
}

#define IX_SEL_STAMP(k)                                                                 \
    do {                                                                                \
        if (blockIdx.x == 0 && threadIdx.x == 0) p.ctl[LC_STAMP + 5 + (k)] = (long long)ix_timer_ns(); \
    } while (0)

__global__ void __launch_bounds__(IX_THREADS, 4) k_exact_select(IterXParams p) {
    long long* ctl = p.ctl;
    if (ctl[LC_STATE] != LS_RUN || !ctl[LC_SELECT]) return;  // uniform over the grid
    IX_SEL_STAMP(0);
    const long long n = ctl[LC_N], c_readd_max = ctl[LC_READD_MAX];
    // CTAs that take part: one per 4096 regions (the grid is sized for the largest pool the host could not rule out)
    long long G_ll = (n + 4095) / 4096;
    if (G_ll > (long long)gridDim.x) G_ll = (long long)gridDim.x;
    if (G_ll < 1) G_ll = 1;
    const int G = (int)G_ll, b = blockIdx.x, tid = threadIdx.x;
    if (b >= G) return;
    __shared__ __align__(16) unsigned hist[5 * IX_HPLANE];
    unsigned long long* const stage = (unsigned long long*)hist;  // 3 x IX_NB words: the partition's bins of a window (hist is dead then)
    __shared__ unsigned long long shc[IX_THREADS], shl[IX_THREADS], shh[IX_THREADS];
    __shared__ PickShared ps;
    __shared__ SelX sx;
    const unsigned lane = tid & 31u;
    const bool sharded = p.R > 1;
    unsigned long long xseq = (unsigned long long)ctl[LC_SEQ];
    unsigned long long* const bar = p.bar;
    unsigned long long nsync = 0;
    if (tid == 0) {
        sx = *p.selx;
        ps.frac = 0ULL;
        ps.Pm_lo = ps.Pm_hi = 0ULL;
        ps.Cc = 0ULL;
        ps.cb = ps.Sb_lo = ps.Sb_hi = 0ULL;
        ps.tie_take = ps.K = ps.thr_key = ps.ambiguous = 0ULL;
    }
    __syncthreads();
    const double* err = p.pool.err;
    const long long per = (n + G - 1) / G;
    const long long lo = per * b;
    long long hi = lo + per;
    if (hi > n) hi = n;
    if (lo > hi) hi = lo;
    unsigned long long* const cand_count = p.hset + (size_t)(IX_LEVELS + 1) * 3 * IX_NB;
    int failed = 0, levels_run = 0;
    // ---- binary exponents: histogram (count, mantissa sum) of windows of IX_BW binades from the top, then every CTA walks
    // down the bins, popping whole binades, until the evaluation budget is used up or the exact residual passes
    // (Rule.java:109-133 at binade granularity): the batch ends inside that binade ------------------------------------
    __shared__ XaBig sE;
    __shared__ unsigned long long s_walk[4];  // found, binade, regions popped above it
    for (int j = tid; j < XA_LIMBS; j += IX_THREADS) sE.l[j] = __ldcg(&p.gbins[j]);
    if (tid == 0) {
        sE.jlo = (int)__ldcg(&p.gbins[IX_PUB_JLO]);
        sE.jhi = (int)__ldcg(&p.gbins[IX_PUB_JHI]);
        s_walk[0] = 0ULL;
        s_walk[1] = 0ULL;
        s_walk[2] = 0ULL;
    }
    __syncthreads();
    const int e_hi_all = (int)__ldcg(&p.gbins[IX_PUB_TOPE]);
    const unsigned long long Nglobal_now = (unsigned long long)ctl[LC_NGLOBAL];
    for (int wdx = 0; !failed; ++wdx) {
        const int e_top_w = e_hi_all - IX_BW * wdx;
        if (e_top_w < 0) break;  // (cannot happen: the walk stops at the last non-empty bin)
        ix_hist_clear(hist);
        unsigned long long* gh0 = p.hset;
        {
            // binary exponents cluster on a few values: a histogram shared by the lanes of a warp would serialise on them.  Every
            // lane pair (l, l + 16) owns a private copy; the 16 copies of a bucket sit in consecutive banks
            unsigned* const my = hist + (lane & 15u);
            for (long long r0 = lo; r0 < hi; r0 += IX_ROUND) {
                const long long r1 = r0 + IX_ROUND < hi ? r0 + IX_ROUND : hi;
                for (long long i0 = r0 + tid; i0 < r1; i0 += 8 * IX_THREADS) {
                    double a[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = i0 + u * IX_THREADS < r1 ? err[i0 + u * IX_THREADS] : -1.0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (i0 + u * IX_THREADS < r1) {
                            const unsigned long long k = ix_key_of_err(a[u]);
                            const int e = (int)(k >> 52), d = e_top_w - e;
                            if (d >= 0 && d < IX_BW) ix_hist_flush(my, (unsigned)d * 16u, 1u, (e ? (1ULL << 52) : 0ULL) | (k & 0xfffffffffffffULL));
                        }
                    }
                }
                ix_hist_to_global(hist, IX_BW, 16, 1, 16, gh0);
            }
        }
        levels_run = 1;
        ix_gsync(bar, ++nsync * (unsigned long long)G);
        IX_SEL_STAMP(1);
        if (sharded) {
            if (b == 0) {
                const int f = ix_exchange(p, xseq, 0, gh0, 0, 3 * IX_NB);
                if (f) atomicExch((unsigned long long*)&ctl[LC_DBG1], (unsigned long long)f);
                unsigned long long* own = p.xbuf[p.rank];
                for (int w = tid; w < 3 * IX_NB; w += IX_THREADS) {
                    unsigned long long sacc = 0;
                    for (int r = 0; r < p.R; ++r) sacc += __ldcg(&ix_pay(own, xseq, r)[w]);
                    gh0[w] = sacc;
                }
            }
            ++xseq;
            ix_gsync(bar, ++nsync * (unsigned long long)G);
            if (__ldcg((const unsigned long long*)&ctl[LC_DBG1])) {
                failed = (int)ctl[LC_DBG1];
                break;
            }
        }
        // the window's bins of the partition -> shared memory (the private histograms are dead): counts, low words, high words
        for (int q = tid; q < 3 * IX_NB; q += IX_THREADS) stage[q] = __ldcg(&gh0[q]);
        __syncthreads();
        if (tid == 0) {
            unsigned long long Cabove = s_walk[2];
            for (int d = 0; d < IX_BW; ++d) {
                const unsigned long long c = stage[d];
                if (!c) continue;
                const int e = e_top_w - d;
                const xa_u128 mv = (xa_u128)stage[IX_NB + d] + ((xa_u128)stage[2 * IX_NB + d] << 32);
                bool stop = Cabove + c >= sx.Kcap || Cabove + c >= Nglobal_now;  // the budget ends here / the lowest bin
                if (!stop && !sx.count_only) {
                    // residual after popping this bin as well; put the bin back when that passes
                    const bool has = mv > 0 && e < 2047;
                    if (has) ix_big_sub(&sE, mv, xa_lsb_pos(e));
                    if (ix_conv(sx.V, sE.to_double(0), p)) {
                        if (has) {
                            sE.add_shifted(mv, xa_lsb_pos(e));
                            sE.normalize();
                        }
                        stop = true;
                    }
                }
                if (stop) {
                    s_walk[0] = 1ULL;
                    s_walk[1] = (unsigned long long)e;
                    break;
                }
                Cabove += c;
            }
            s_walk[2] = Cabove;
            if (s_walk[0]) {  // sE is the residual before the first pop inside the binade
                const int bstar = (int)s_walk[1], pos = xa_lsb_pos(bstar);
                sx.binade = bstar;
                sx.pos = pos;
                sx.H_lo = sE.window64(pos);
                sx.H_hi = sE.window64(pos + 64);
                sx.L1 = pos >= 64 ? sE.window64(pos - 64) : (pos > 0 ? (sE.window64(0) << (64 - pos)) : 0ULL);
                // the part below pos as a truncated double (only used when everything above it is popped)
                for (int j = sE.jlo; j <= sE.jhi; ++j) {
                    const int lo_bit = 32 * j;
                    if (lo_bit >= pos)
                        sE.l[j] = 0ULL;
                    else if (lo_bit + 32 > pos)
                        sE.l[j] &= (1ULL << (pos - lo_bit)) - 1ULL;
                }
                sx.Ldbl = sE.to_double(0);
                sx.Cabove = Cabove;
                ps.Cc = Cabove;
            }
        }
        __syncthreads();
        IX_SEL_STAMP(2);
        if (s_walk[0]) break;
        // the batch reaches below this window (more than IX_BW binades under the top): clear the histogram and go on
        ix_gsync(bar, ++nsync * (unsigned long long)G);
        for (long long q = (long long)b * IX_THREADS + tid; q < 3 * IX_NB; q += (long long)G * IX_THREADS) p.hset[q] = 0ULL;
        ix_gsync(bar, ++nsync * (unsigned long long)G);
    }
    if (failed) {
        if (b == 0 && tid == 0) {
            ix_finish(p, LS_ERROR, failed);
            ctl[LC_SEQ] = (long long)xseq;
        }
        return;
    }
    const unsigned long long top_mask = 0xfffULL << 52;  // sign + exponent
    const unsigned long long top_prefix = (unsigned long long)sx.binade << 52;
    bool have_cand = false, gathered = false;
    unsigned long long above_def = 0;  // (thread-local) regions of this chunk that are certainly popped
    long long M = 0, Mg = 0;
    for (int level = 0; level < IX_LEVELS; ++level) {
        const int shift = ix_level_shift(level), NB = ix_level_nb(level);
        const int copies = level <= 1 ? IX_HCOPIES : 1;
        ix_hist_clear(hist);
        // keys decided so far: sign, exponent and the fraction bits above this level's digit
        const unsigned long long mask = top_mask | (0xfffffffffffffULL & ~((1ULL << (shift + ix_level_bits(level))) - 1ULL));
        const unsigned long long prefix = top_prefix | ps.frac;
        const unsigned long long rmask = (1ULL << shift) - 1ULL;
        unsigned long long* gh = p.hset + (size_t)(level + 1) * 3 * IX_NB;
        levels_run = level + 2;
        if (level <= 1) {
            // full scan of the chunk; the second scan also extracts the candidates and counts the certain pops
            unsigned* const my = hist + (tid >> 6) * 256;
            for (long long r0 = lo; r0 < hi; r0 += IX_ROUND) {
                const long long r1 = r0 + IX_ROUND < hi ? r0 + IX_ROUND : hi;
                unsigned cur = 0, run_c = 0;
                unsigned long long run = 0;
                for (long long w0 = r0 + (long long)(tid - (int)lane); w0 < r1; w0 += 4 * IX_THREADS) {
                    const long long i0 = w0 + lane;
                    double a[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) a[u] = i0 + u * IX_THREADS < r1 ? err[i0 + u * IX_THREADS] : 0.0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const long long i = i0 + u * IX_THREADS;
                        const bool in = i < r1;
                        const unsigned long long k = ix_key_of_err(a[u]);
                        const bool match = in && (k & mask) == prefix;
                        if (level == 1) {
                            above_def += (in && (k & mask) > prefix) ? 1ULL : 0ULL;
                            const unsigned bal = __ballot_sync(0xffffffffu, match);  // warp-aggregated append
                            if (bal) {
                                unsigned long long base = 0;
                                if (lane == (unsigned)(__ffs(bal) - 1)) base = atomicAdd(cand_count, (unsigned long long)__popc(bal));
                                base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
                                if (match) {
                                    const unsigned long long pos = base + (unsigned long long)__popc(bal & ((1u << lane) - 1u));
                                    p.cand_key[pos] = k;
                                    p.cand_idx[pos] = (int)i;
                                }
                            }
                        }
                        if (match) {
                            const unsigned d = (unsigned)(k >> shift) & (unsigned)(NB - 1);
                            if (d != cur) {
                                ix_hist_flush(my, cur, run_c, run);
                                cur = d;
                                run_c = 0;
                                run = 0;
                            }
                            ++run_c;
                            run += k & rmask;
                        }
                    }
                }
                ix_hist_flush(my, cur, run_c, run);
                ix_hist_to_global(hist, NB, copies, NB, 1, gh);
            }
            if (level == 1) have_cand = true;
        } else {
            // the candidates: this rank's own, or (sharded pool, after the gather below) those of the whole partition
            const unsigned long long* const ckeys = gathered ? p.gcand : p.cand_key;
            const long long Mlist = gathered ? Mg : M;
            const long long cstride = (long long)G * IX_THREADS;
            for (long long r0 = 0; r0 < Mlist; r0 += cstride * (IX_ROUND / IX_THREADS)) {  // a CTA takes IX_ROUND candidates per round
                const long long r1 = r0 + cstride * (IX_ROUND / IX_THREADS) < Mlist ? r0 + cstride * (IX_ROUND / IX_THREADS) : Mlist;
                unsigned cur = 0, run_c = 0;
                unsigned long long run = 0;
                for (long long q = r0 + (long long)b * IX_THREADS + tid; q < r1; q += cstride) {
                    const unsigned long long k = gathered ? __ldcg(&ckeys[q]) : ckeys[q];
                    if ((k & mask) == prefix) {
                        const unsigned d = (unsigned)(k >> shift) & (unsigned)(NB - 1);
                        if (d != cur) {
                            ix_hist_flush(hist, cur, run_c, run);
                            cur = d;
                            run_c = 0;
                            run = 0;
                        }
                        ++run_c;
                        run += k & rmask;
                    }
                }
                ix_hist_flush(hist, cur, run_c, run);
                ix_hist_to_global(hist, NB, copies, NB, 1, gh);
            }
        }
        if (level == 1) {  // certain pops of this chunk -> per-CTA counters of the compaction
            unsigned long long sa = above_def, sb = 0ULL;
            ix_block_sum2(sa, sb, shc, shl);
            if (tid == 0) {
                p.blk[b] = sa;
                p.blk[G + b] = 0ULL;
            }
        }
        ix_gsync(bar, ++nsync * (unsigned long long)G);
        if (sharded && !gathered) {  // this level's histogram of the whole partition: sum of the R local ones
            // (one CTA: the ranks' grids differ in size, and 1 536 words per peer are not worth slicing)
            const int words = 3 * IX_NB;
            if (b == 0) {
                const int w0 = 0, w1 = words;
                const int f = ix_exchange(p, xseq, 0, gh, w0, w1);
                if (f) atomicExch((unsigned long long*)&ctl[LC_DBG1], (unsigned long long)f);
                unsigned long long* own = p.xbuf[p.rank];
                for (int w = w0 + tid; w < w1; w += IX_THREADS) {
                    unsigned long long s = 0;
                    for (int r = 0; r < p.R; ++r) s += __ldcg(&ix_pay(own, xseq, r)[w]);
                    gh[w] = s;
                }
            }
            ++xseq;
            ix_gsync(bar, ++nsync * (unsigned long long)G);
            if (__ldcg((const unsigned long long*)&ctl[LC_DBG1])) {
                failed = (int)ctl[LC_DBG1];
                break;
            }
        }
        if (level == 1) M = (long long)__ldcg(cand_count);
        const int fin = ix_pick(p, sx, &ps, gh, level, shc, shl, shh);
        if (level < 6) IX_SEL_STAMP(3 + level);
        if (fin) break;
        if (sharded && level == 1 && p.gcand) {
            // ---- sharded pool: instead of one histogram exchange per remaining level (up to four), the ranks all-gather
            // their candidates once (the regions that match the 27 key bits decided so far: a few hundred per rank) and
            // every rank finishes the select on the identical list.  Falls back to per-level exchanges when a rank holds
            // more candidates than a payload slot takes.
            unsigned long long* const stagec = p.gcand + (size_t)PK_MAX_RANKS * IX_GCAP;  // [16 + IX_GCAP] staging: header | keys
            if (b == 0) {
                const long long mc = M <= IX_GCAP ? M : 0;
                if (tid < 16) stagec[tid] = tid == 0 ? (unsigned long long)M : 0ULL;
                for (long long q = tid; q < mc; q += IX_THREADS) stagec[16 + q] = p.cand_key[q];
                __threadfence();
                __syncthreads();
                const int f = ix_exchange(p, xseq, 0, stagec, 0, 16 + (int)mc);
                if (f) atomicExch((unsigned long long*)&ctl[LC_DBG1], (unsigned long long)f);
            }
            ++xseq;
            ix_gsync(bar, ++nsync * (unsigned long long)G);
            if (__ldcg((const unsigned long long*)&ctl[LC_DBG1])) {
                failed = (int)ctl[LC_DBG1];
                break;
            }
            {
                const unsigned long long* own = p.xbuf[p.rank];
                bool fits = true;
                long long tot = 0;
                for (int r = 0; r < p.R; ++r) {
                    const long long mr = (long long)__ldcg(&ix_pay((unsigned long long*)own, xseq - 1ULL, r)[0]);
                    fits = fits && mr <= IX_GCAP;
                    tot += mr;
                }
                if (fits) {  // uniform over the grid and over the ranks (same headers everywhere)
                    long long base_r = 0;
                    for (int r = 0; r < p.R; ++r) {  // concatenation in rank order
                        const unsigned long long* pay = ix_pay((unsigned long long*)own, xseq - 1ULL, r);
                        const long long mr = (long long)__ldcg(&pay[0]);
                        for (long long q = (long long)b * IX_THREADS + tid; q < mr; q += (long long)G * IX_THREADS) p.gcand[base_r + q] = __ldcg(&pay[16 + q]);
                        base_r += mr;
                    }
                    gathered = true;
                    Mg = tot;
                    ix_gsync(bar, ++nsync * (unsigned long long)G);
                }
            }
        }
    }
    if (failed) {
        if (b == 0 && tid == 0) {
            ix_finish(p, LS_ERROR, failed);
            ctl[LC_SEQ] = (long long)xseq;
        }
        return;
    }
    // ---- stable compaction of the popped slots (every CTA holds the same selection state: integer arithmetic on the
    // same global histograms) -----------------------------------------------------------------------------------------
    const unsigned long long thr = ps.thr_key, tie_take = ps.tie_take;
    if (have_cand) {
        // per-CTA counts: certain pops (second scan) + the candidates above / at the threshold key
        for (long long q = (long long)b * IX_THREADS + tid; q < M; q += (long long)G * IX_THREADS) {
            const unsigned long long k = p.cand_key[q];
            if (k >= thr) {
                const long long owner = (long long)p.cand_idx[q] / per;
                atomicAdd(&p.blk[(k > thr ? 0 : 1) * (long long)G + owner], 1ULL);
            }
        }
    } else {
        unsigned long long my_gt = 0, my_eq = 0;
        for (long long i = lo + tid; i < hi; i += IX_THREADS) {
            const unsigned long long k = ix_key_of_err(err[i]);
            my_gt += k > thr;
            my_eq += k == thr;
        }
        ix_block_sum2(my_gt, my_eq, shc, shl);
        if (tid == 0) {
            p.blk[b] = my_gt;
            p.blk[G + b] = my_eq;
        }
    }
    ix_gsync(bar, ++nsync * (unsigned long long)G);
    IX_SEL_STAMP(9);
    // the histograms are not read any more: clear the levels that were used (and the candidate counter) for the next launch
    for (long long q = (long long)b * IX_THREADS + tid; q < (long long)levels_run * 3 * IX_NB; q += (long long)G * IX_THREADS) p.hset[q] = 0ULL;
    if (b == 0 && tid < 8) cand_count[tid] = 0ULL;
    unsigned long long tie_take_local = tie_take, K_local = ps.K;
    if (sharded) {
        // regions equal to the threshold key are taken in (rank, slot) order: every rank publishes how many of its regions are
        // above / at the threshold and takes its share of the tie budget
        if (b == 0) {
            unsigned long long a = 0, c = 0;
            for (int i = tid; i < G; i += IX_THREADS) {
                a += __ldcg(&p.blk[i]);
                c += __ldcg(&p.blk[G + i]);
            }
            ix_block_sum2(a, c, shc, shl);
            if (tid < 16) shh[tid] = tid == 0 ? a : (tid == 1 ? c : 0ULL);
            __syncthreads();
            const int f = ix_exchange(p, xseq, 0, shh, 0, 16);
            if (tid == 0) {
                unsigned long long eq_before = 0ULL, my_gt = 0ULL, my_eq = 0ULL;
                for (int r = 0; r < p.R; ++r) {
                    const unsigned long long* pay = ix_pay(p.xbuf[p.rank], xseq, r);
                    if (r < p.rank) eq_before += __ldcg(&pay[1]);
                    if (r == p.rank) {
                        my_gt = __ldcg(&pay[0]);
                        my_eq = __ldcg(&pay[1]);
                    }
                }
                unsigned long long share = tie_take > eq_before ? tie_take - eq_before : 0ULL;
                if (share > my_eq) share = my_eq;
                p.blk[2 * G] = share;
                p.blk[2 * G + 1] = my_gt + share;
                p.blk[2 * G + 2] = (unsigned long long)f;
            }
            __threadfence();
        }
        ++xseq;
        ix_gsync(bar, ++nsync * (unsigned long long)G);
        if (__ldcg(&p.blk[2 * G + 2])) {
            if (b == 0 && tid == 0) {
                ix_finish(p, LS_ERROR, (int)p.blk[2 * G + 2]);
                ctl[LC_SEQ] = (long long)xseq;
            }
            return;
        }
        tie_take_local = __ldcg(&p.blk[2 * G]);
        K_local = __ldcg(&p.blk[2 * G + 1]);
    }
    // A batch that takes almost the whole shard (BASELINE config C5: every cut leaves a handful of regions whose errors are
    // below the unreachable tolerance): instead of copying (val, err) of the K popped parents for the next decide kernel to
    // subtract (32 B written and 24 B read per region), this rank's sums are reset and the few kept regions are listed
    // and added again.  Exact sums: the result is the same integers.  Only when the pool holds no non-finite value (their
    // counters and sticky flags follow the reference's Inf - Inf = NaN history, which a reset would lose).
    bool readd = false;
    {
        const long long kept_n = n - (long long)K_local, readd_max = c_readd_max;
        if (readd_max > 0 && kept_n * 16 <= n * readd_max) {  // at most LC_READD_MAX / 16 of the shard stays
            unsigned long long nonfinite = 0ULL;
#pragma unroll
            for (int q = 0; q < 5; ++q) nonfinite |= __ldcg(&p.acc[XA_M_OFF + q]);  // (five loads in flight)
            readd = nonfinite == 0ULL;
        }
    }
    unsigned long long gt_base = 0, eq_base = 0;
    {
        unsigned long long a = 0, c = 0;
        for (int i = tid; i < b; i += IX_THREADS) {
            a += __ldcg(&p.blk[i]);
            c += __ldcg(&p.blk[G + i]);
        }
        ix_block_sum2(a, c, shc, shl);
        gt_base = a;
        eq_base = c;
    }
    __syncthreads();
    // tiles of IX_TILE consecutive slots, element (u, thread) at t_lo + u * IX_THREADS + thread (coalesced); the rank of an
    // element among the popped ones follows from warp ballots and one scan over the tile's 64 (u, warp) counts
    unsigned* const sgt = hist;
    unsigned* const seq = hist + 80;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int warp = tid >> 5;
    for (long long t_lo = lo; t_lo < hi; t_lo += IX_TILE) {
        unsigned bg[IX_TILE_ITEMS], be[IX_TILE_ITEMS];
#pragma unroll
        for (int u = 0; u < IX_TILE_ITEMS; ++u) {
            const long long i = t_lo + (long long)u * IX_THREADS + tid;
            const unsigned long long k = i < hi ? ix_key_of_err(err[i]) : 0ULL;
            bg[u] = __ballot_sync(0xffffffffu, i < hi && k > thr);
            be[u] = __ballot_sync(0xffffffffu, i < hi && k == thr);
            if (lane == 0) {
                sgt[u * IX_WARPS + warp] = (unsigned)__popc(bg[u]);
                seq[u * IX_WARPS + warp] = (unsigned)__popc(be[u]);
            }
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 64 counts (two per lane), totals behind them
            const unsigned g0 = sgt[2 * lane], g1 = sgt[2 * lane + 1], q0 = seq[2 * lane], q1 = seq[2 * lane + 1];
            unsigned gs = g0 + g1, qs = q0 + q1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned tg = __shfl_up_sync(0xffffffffu, gs, o), tq = __shfl_up_sync(0xffffffffu, qs, o);
                if ((int)lane >= o) {
                    gs += tg;
                    qs += tq;
                }
            }
            sgt[2 * lane] = gs - g0 - g1;
            sgt[2 * lane + 1] = gs - g1;
            seq[2 * lane] = qs - q0 - q1;
            seq[2 * lane + 1] = qs - q1;
            if (lane == 31) {
                sgt[64] = gs;
                seq[64] = qs;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < IX_TILE_ITEMS; ++u) {
            const bool is_gt = (bg[u] >> lane) & 1u, is_eq = (be[u] >> lane) & 1u;
            const long long i = t_lo + (long long)u * IX_THREADS + tid;
            // popped regions ahead of this one in slot order
            const unsigned long long gt_before = gt_base + sgt[u * IX_WARPS + warp] + (unsigned)__popc(bg[u] & lt_mask);
            const unsigned long long eq_before = eq_base + seq[u * IX_WARPS + warp] + (unsigned)__popc(be[u] & lt_mask);
            const unsigned long long pos = gt_before + (eq_before < tie_take_local ? eq_before : tie_take_local);
            if (is_gt || (is_eq && eq_before < tie_take_local)) {
                p.list[pos] = (int)i;
                if (!readd) {
                    // the parent leaves the running sums (RegionHeap.poll, RegionHeap.java:21-28): the next decide kernel subtracts
                    // this copy, the slot itself is about to be overwritten by the left child
                    p.stash_val[pos] = p.pool.val[i];
                    p.stash_err[pos] = err[i];
                }
            } else if (readd && i < hi) {
                p.kept[(unsigned long long)i - pos] = (int)i;  // kept regions in slot order: i minus the popped ones ahead of it
            }
        }
        gt_base += sgt[64];
        eq_base += seq[64];
        __syncthreads();
    }
    IX_SEL_STAMP(10);
    if (b == 0 && tid == 0) {
        const long long Kl = (long long)K_local, Kg = (long long)ps.K, Nglobal = ctl[LC_NGLOBAL];
        p.selx->thr_key = ps.thr_key;
        p.selx->tie_take = ps.tie_take;
        p.selx->K = ps.K;
        ctl[LC_DYN_N] = 2 * Kl;
        ctl[LC_DYN_LIST] = 1;
        ctl[LC_DYN_BASE] = n;
        ctl[LC_K] = Kl;
        ctl[LC_SUB_K] = readd ? 0 : Kl;
        ctl[LC_READD] = readd ? n - Kl : 0;
        if (readd) {  // this rank's sums restart from zero (as when the whole pool is popped); the kept regions follow in the next decide
            unsigned long long* tot = p.acc + XA_TOT_OFF;
            for (int j = 0; j < 2 * XA_LIMBS; ++j) tot[j] = 0ULL;
            tot[XA_TOT_MISC + 0] = 0ULL;
            tot[XA_TOT_MISC + 1] = 0ULL;
            tot[XA_TOT_MISC + 2] = tot[XA_TOT_MISC + 4] = (unsigned long long)(XA_LIMBS - 1);
            tot[XA_TOT_MISC + 3] = tot[XA_TOT_MISC + 5] = 0ULL;
            p.acc[XA_M_OFF + XA_M_MINKEY] = ~0ULL;
        }
        ctl[LC_KGLOBAL] = Kg;
        ctl[LC_N] = n + Kl;
        ctl[LC_NGLOBAL] = Nglobal + Kg;
        ctl[LC_NUMEVAL] += 2 * ctl[LC_P] * Kg;
        ctl[LC_EVAL_LOCAL] += 2 * Kl;
        ctl[LC_WORDS + (ctl[LC_NBATCH] % LC_LOG_CAP)] = 2 * Kg;
        ctl[LC_NBATCH] += 1;
        ctl[LC_ITER] += 1;
        ctl[LC_NSELECT] += 1;
        ctl[LC_AMBIG] += (long long)ps.ambiguous;
        ctl[LC_SEQ] = (long long)xseq;
        ctl[LC_SELECT] = 0;
        if (Kg >= Nglobal) p.acc[XA_M_OFF + XA_M_MINKEY] = ~0ULL;  // every region leaves the pool
    }
}

// ---- rebuild the sums from the pool ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(IX_THREADS) k_acc_rebuild(PoolView pool, long long n, unsigned long long* acc) {
    __shared__ CubAccWin<IX_THREADS> win;
    cub_acc_begin(acc, &win);
    CubAccKeys keys;
    cub_acc_keys_init(keys);
    for (long long i = (long long)blockIdx.x * IX_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * IX_THREADS)
        cub_acc_region(&win, acc, keys, pool.val[i], pool.err[i], 1);
    cub_acc_keys_store(&win, keys);
    cub_acc_end(acc, &win);
}

__global__ void k_abort_peers(int R, IterXParams p) {
    if (threadIdx.x < R && p.xbuf[threadIdx.x]) {
        p.xbuf[threadIdx.x][PK_XB_ABORT] = 1ULL;
        __threadfence_system();
    }
}

// ---- host launchers ----------------------------------------------------------------------------------------------------
static int g_decide_blocks_per_sm = -1;
int ix_decide_max_blocks(int n_sms) {
    if (g_decide_blocks_per_sm < 0) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_exact_decide<4>, IX_THREADS, 0);
        g_decide_blocks_per_sm = nb > 0 ? (nb > 4 ? 4 : nb) : 1;
    }
    return n_sms * g_decide_blocks_per_sm;
}
cudaError_t ix_decide(const IterXParams& p, int n_blocks, int n_sms, cudaStream_t s) {
    if (n_blocks < 1) n_blocks = 1;
    if (n_blocks <= n_sms)
        k_exact_decide<1><<<(unsigned)n_blocks, IX_THREADS, 0, s>>>(p);
    else
        k_exact_decide<4><<<(unsigned)n_blocks, IX_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}
static int g_select_blocks_per_sm = -1, g_select_max_blocks = 1;
int ix_select_max_blocks(int n_sms) {
    if (g_select_blocks_per_sm < 0) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_exact_select, IX_THREADS, 0);
        g_select_blocks_per_sm = nb > 0 ? (nb > 4 ? 4 : nb) : 1;
    }
    g_select_max_blocks = n_sms * g_select_blocks_per_sm;
    return g_select_max_blocks;
}
cudaError_t ix_select(const IterXParams& p, int n_blocks, cudaStream_t s) {
    IterXParams q = p;
    void* args[1] = {&q};
    return cudaLaunchCooperativeKernel((const void*)k_exact_select, dim3((unsigned)n_blocks), dim3(IX_THREADS), args, 0, s);
}
void ix_rebuild(const PoolView& pool, long long n, unsigned long long* acc, int n_sms, cudaStream_t s) {
    // everything but the window hints, the sticky flags and the upper bound of the largest errmax starts from zero
    cudaMemsetAsync(acc + XA_D_OFF, 0, (size_t)XA_REPL * XA_CELLS * 8, s);
    cudaMemsetAsync(acc + XA_M_OFF, 0, 5 * 8, s);
    cudaMemsetAsync(acc + XA_M_OFF + XA_M_MINKEY, 0xff, 8, s);
    cudaMemsetAsync(acc + XA_M_OFF + XA_M_EMAXKEY, 0, 2 * 8, s);
    cudaMemsetAsync(acc + XA_FAR_E_OFF, 0, (size_t)(XA_WORDS - XA_FAR_E_OFF) * 8, s);
    long long g = (n + IX_THREADS - 1) / IX_THREADS;
    if (g > (long long)n_sms * 8) g = (long long)n_sms * 8;
    if (g < 1) g = 1;
    k_acc_rebuild<<<(unsigned)g, IX_THREADS, 0, s>>>(pool, n, acc);
}
void ix_abort_peers(int R, void* const* xbuf, cudaStream_t s) {
    IterXParams p{};
    for (int r = 0; r < R && r < PK_MAX_RANKS; ++r) p.xbuf[r] = (unsigned long long*)xbuf[r];
    k_abort_peers<<<1, 32, 0, s>>>(R, p);
}
