// iter_vec.cu -- the device-resident refinement loop of VECTOR integrands (fdim > 1): one cooperative kernel per
// iteration of Rule.cubature (Rule.java:72-142).
//
//   A  totals of val / err per component in a fixed order (lane-strided partial sums, shuffle tree, CTA order), smallest /
//      largest errmax; for the separable GK15 path also errmax of the regions the last rule launch wrote (Rule.java:137-139)
//   B  loop condition, Rule.converged on the totals (Rule.java:72-79), evaluation budget (Rule.java:133), the pop-everything
//      shortcut (the residual the smallest region alone leaves does not pass: no smaller batch does, the predicate is
//      monotone in the batch size)
//   C  the Gladwell batch cut (Rule.java:109-133): the reference pops regions in descending errmax and stops as soon as
//      converged(val, err - popped err) holds for the vector of per-component residuals (Rule.java:122-124, :129).  Here:
//      radix descent over the composite key (errmax bits, then lower slot first) 4 bits at a time, starting at the highest
//      bit in which the pool's keys differ; every level sums err[j] per bucket and component (registers, fixed order: no
//      atomics, bit-reproducible), the bucket in which the predicate first holds (or the budget ends) is entered, the
//      buckets above it join the popped sum.  Once the bucket holds <= IV_CMAX regions, CTA 0 sorts them (bitonic, shared
//      memory), forms the running residuals and evaluates the predicate for every batch size in parallel (+- the rounding
//      band of the reference's drifting sums: cub_stats_t.ambiguous_decisions).
//   D  stable compaction of the popped slots, bisection (Region.cut, Region.java:31-38), control block for the rule launch.
//
// Summation order: the reference subtracts the popped errors one by one from a drifting running sum; here the popped sum is
// a tree (buckets, chunks) -- same value up to rounding, identical decisions outside the rounding band.
#include "converge.h"
#include "iter_vec.h"

namespace {

typedef unsigned __int128 iv_u128;
#define IV_WARPS (IV_THREADS / 32)
#define IV_TILE 2048
enum { IV_ACT_FINISH = 0, IV_ACT_POPALL = 1, IV_ACT_CUT = 2 };

__device__ __forceinline__ unsigned long long iv_bits(double x) { return (unsigned long long)__double_as_longlong(x); }
__device__ __forceinline__ unsigned long long iv_key(double errmax) { return iv_bits(errmax > 0.0 ? errmax : 0.0); }  // Rule.java:281-289
__device__ __forceinline__ iv_u128 iv_comp(unsigned long long key, long long slot) {
    return ((iv_u128)key << 32) | (iv_u128)(0xffffffffu - (unsigned)slot);
}

__device__ __forceinline__ unsigned long long iv_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// phase time stamps of the last iteration (CUBATURE_TIMING): CTA 0, thread 0
#define IV_STAMP(k)                                                                         \
    do {                                                                                    \
        if (blockIdx.x == 0 && threadIdx.x == 0) p.ctl[LC_STAMP + (k)] = (long long)iv_timer_ns(); \
    } while (0)

__device__ __forceinline__ void iv_gsync(unsigned long long* bar, unsigned long long target) {
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

// One accessor type for every residual the kernel tests, and ONE out-of-line instantiation of Rule.converged for it: the
// predicate is evaluated at eleven places, and inlined per accessor type the kernel grew to 46 000 instructions (740 KB) --
// every phase of every iteration then started with instruction-cache misses (51 us for a pass that computes for 5 us).
struct IvAcc {
    int kind, j0, f, q;
    const double* v;  // Sum val (shared)
    const double* a;
    const double* b;
    const double* c;
    double s, t;
    __device__ double val(int j) const { return v[j0 + j]; }
    __device__ double err(int jj) const {
        const int j = j0 + jj;
        switch (kind) {
            case 0:  // a_j * s + t * b_j (shared vectors)
                return a[j] * s + t * b[j];
            case 1: {  // residual after popping the buckets q .. IV_NB-1 of a level: a = Sum err (shared), b = popped sum, c = bucket sums (global)
                double x = 0.0;
                for (int u = IV_NB - 1; u >= q; --u) x += __ldcg(&c[(long long)u * f + j]);
                return a[j] - (__ldcg(&b[j]) + x);
            }
            case 2:  // the same from suffix sums staged in shared memory: b = popped sum, c = suffix sum of the buckets q ..
                return a[j] - (b[j] + c[j]);
            default:  // residual after popping the sorted candidates 0 .. k: b = residual before k's chunk, c = chunk-local running sum (global)
                return (b[j] - c[j]) + t * a[j];
        }
    }
};
__device__ __forceinline__ IvAcc iv_band_acc(const double* V, const double* e, const double* base, double scale, double add) {
    return IvAcc{0, 0, 0, 0, V, e, base, nullptr, scale, add};
}
__device__ __forceinline__ IvAcc iv_bucket_acc(const double* V, const double* E, const double* sab, const double* bsum, int f, int q) {
    return IvAcc{1, 0, f, q, V, E, sab, bsum, 0.0, 0.0};
}
__device__ __forceinline__ IvAcc iv_suffix_acc(const double* V, const double* E, const double* sab, const double* ssq) {
    return IvAcc{2, 0, 0, 0, V, E, sab, ssq, 0.0, 0.0};
}
__device__ __forceinline__ IvAcc iv_cand_acc(const double* V, const double* E, const double* cbase, const double* crow, double add) {
    return IvAcc{3, 0, 0, 0, V, E, cbase, crow, 0.0, add};
}
__device__ __noinline__ bool iv_conv(int f, IvAcc acc, double absTol, double relTol, int norm) { return cub_converged(f, acc, absTol, relTol, norm); }

// Rule.converged evaluated by a whole warp (every lane calls with the same arguments, every lane gets the result).
// INDIVIDUAL / PAIRED test the components (pairs) independently and combine them with OR (the reference's ANY, Rule.java:
// 172-213) or AND (strict-ALL extension): lanes take the units in turn and evaluate each with cub_converged itself on a
// one-unit view, so the arithmetic is literally the reference's.  The other norms carry ordered sums: lane 0 evaluates.
struct IvArr {  // flat vectors
    const double* v;
    const double* e;
    __device__ double val(int j) const { return v[j]; }
    __device__ double err(int j) const { return e[j]; }
};
__device__ __noinline__ bool iv_conv_arr(int f, const double* v, const double* e, double absTol, double relTol, int norm) {
    return cub_converged(f, IvArr{v, e}, absTol, relTol, norm);
}
// scratch: this warp's [f] doubles in global memory (the norms with ordered sums: the lanes materialise the residual vector,
// lane 0 walks the flat array -- through the generic accessor the walk took 90 ns per element)
__device__ __forceinline__ bool iv_conv_warp(int f, IvAcc ee, double absTol, double relTol, int norm, double* scratch) {
    const unsigned lane = threadIdx.x & 31u;
    const int n = norm & 7;
    if (n <= 1) {
        const int width = n == 0 ? 1 : 2;
        const int units = (f + width - 1) / width;
        const bool strict = (norm & 8) != 0;
        bool mine = strict;
        for (int u = (int)lane; u < units; u += 32) {
            IvAcc one = ee;
            one.j0 = u * width;
            const int len = f - one.j0 < width ? f - one.j0 : width;
            const bool c = iv_conv(len, one, absTol, relTol, norm);
            mine = strict ? (mine && c) : (mine || c);
        }
        return strict ? __all_sync(0xffffffffu, mine) != 0 : __any_sync(0xffffffffu, mine) != 0;
    }
    for (int j = (int)lane; j < f; j += 32) scratch[j] = ee.err(j);
    __syncwarp();
    int r = 0;
    if (lane == 0) r = iv_conv_arr(f, ee.v + ee.j0, scratch, absTol, relTol, norm) ? 1 : 0;
    r = __shfl_sync(0xffffffffu, r, 0);
    __syncwarp();  // the scratch vector is free again
    return r != 0;
}

struct IvShared {
    unsigned char sidx[IV_TILE];    // digits of the tile's matching regions ...
    unsigned short midx[IV_TILE];   // ... and their index inside the tile, in slot order
    unsigned long long red_a[IV_THREADS], red_b[IV_THREADS], red_c[IV_THREADS];
    unsigned s_cnt[IV_NB];
    unsigned long long bc[IV_NB];  // bucket counts of the partition at the current level
    int s_pred[8];
    int s_stop[IV_NB];
    int action, state, count_only, chosen, amb;
    long long Kcap;
    unsigned long long kmin[3];
    unsigned wcnt[IV_WARPS + 1];
};

__device__ __forceinline__ void iv_bisect_one(const PoolView& pool, int dim, long long parent, long long right) {
    const int s = pool.splitdim[parent];
    for (int d = 0; d < dim; ++d) {
        double c = pool.center[(long long)d * pool.cap + parent];
        double h = pool.halfw[(long long)d * pool.cap + parent];
        if (d == s) {
            h = h * 0.5;
            pool.halfw[(long long)d * pool.cap + parent] = h;
            pool.center[(long long)d * pool.cap + parent] = c - h;
            pool.center[(long long)d * pool.cap + right] = c + h;
        } else {
            pool.center[(long long)d * pool.cap + right] = c;
        }
        pool.halfw[(long long)d * pool.cap + right] = h;
    }
}

__global__ void __launch_bounds__(IV_THREADS, 2) k_vec_iter(IterVParams p) {
    extern __shared__ double tv[];  // [4 f]: Sum val, Sum err, val / err of the smallest region; then [IV_NB + 1][f] if f <= IV_SS_FMAX
    const bool use_ss = p.fdim <= IV_SS_FMAX;
    double* const ss = tv + 4 * p.fdim;
    __shared__ IvShared S;
    long long* ctl = p.ctl;
    const int tid = threadIdx.x, b = blockIdx.x, Gs = gridDim.x, f = p.fdim;
    const unsigned lane = tid & 31u;
    const int warp = tid >> 5;
    double* const wscr = p.wscratch + ((long long)b * IV_WARPS + warp) * f;  // this warp's scratch vector
    if (ctl[LC_STATE] != LS_RUN) {  // finished or paused: this launch and the ones queued behind it do nothing
        if (b == 0 && tid == 0) ctl[LC_DYN_N] = 0;
        return;
    }
    // the loop state as it is at launch: CTA 0 changes the control block at the end of this launch, so every CTA reads what
    // it needs before the first grid barrier
    const long long N = ctl[LC_N], c_numEval = ctl[LC_NUMEVAL], c_maxEval = ctl[LC_MAXEVAL], c_P = ctl[LC_P], c_cap = ctl[LC_CAP],
                    c_iter = ctl[LC_ITER], c_maxiter = ctl[LC_MAXITER], c_stop_at = ctl[LC_STOP_AT];
    const PoolView& pool = p.pool;
    unsigned long long* const bar = p.bar;
    IV_STAMP(0);
    // ---- errmax of the regions the last rule launch wrote (separable GK15 path), by the whole grid ---------------------
    if (p.need_errmax) {  // one warp per item: lanes over the components
        const long long n_items = ctl[LC_DYN_N], use_list = ctl[LC_DYN_LIST], base = ctl[LC_DYN_BASE];
        for (long long i = (long long)b * IV_WARPS + warp; i < n_items; i += (long long)Gs * IV_WARPS) {
            const long long pi = i >> 1;
            const long long slot = (i & 1) ? base + pi : (use_list ? (long long)p.list[pi] : pi);
            double emax = 0.0;
            for (int j = (int)lane; j < f; j += 32) {
                const double e = pool.err[(long long)j * pool.cap + slot];
                if (e > emax) emax = e;  // NaN never wins (Rule.java:281-289)
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double e = __shfl_xor_sync(0xffffffffu, emax, o);
                if (e > emax) emax = e;
            }
            if (lane == 0) {
                pool.errmax[slot] = emax;
                pool.splitdim[slot] = 0;
            }
        }
    }
    // every CTA of the grid has read the loop state before anybody changes it
    iv_gsync(bar, (unsigned long long)Gs);
    IV_STAMP(1);
    long long Gp_ll = (N + 1023) / 1024;  // CTAs that take part: one per 1 024 regions, at least one warp per row of the totals
    if (Gp_ll < (2 * f + IV_WARPS - 1) / IV_WARPS) Gp_ll = (2 * f + IV_WARPS - 1) / IV_WARPS;
    if (Gp_ll > Gs) Gp_ll = Gs;
    if (Gp_ll < 1) Gp_ll = 1;
    const int Gp = (int)Gp_ll;
    unsigned long long nsync = 0;  // barriers among the Gp participants
#define IV_SYNC() iv_gsync(bar, (unsigned long long)Gs + (++nsync) * (unsigned long long)Gp)
    if (b < Gp) {
        const long long per = (N + Gp - 1) / Gp;
        const long long lo = per * b;
        long long hi = lo + per;
        if (hi > N) hi = N;
        if (lo > hi) hi = lo;
        // ---- A: partial sums of (row, chunk) items, one warp per item; the chunking depends on N alone --------------------
        long long nch_ll = (N + 1023) / 1024;
        if (nch_ll > Gp) nch_ll = Gp;
        const int nch = (int)nch_ll;
        const long long perA = (N + nch - 1) / nch;
        for (long long item = (long long)b * IV_WARPS + warp; item < (long long)2 * f * nch; item += (long long)Gp * IV_WARPS) {
            const int row = (int)(item / nch), c = (int)(item - (long long)row * nch);
            const double* src = row < f ? pool.val + (long long)row * pool.cap : pool.err + (long long)(row - f) * pool.cap;
            const long long a_lo = perA * c;
            long long a_hi = a_lo + perA;
            if (a_hi > N) a_hi = N;
            double s = 0.0;  // fixed order: lane l adds elements l, l + 32, ... of the chunk, then a shuffle tree
            long long i = a_lo + lane;
            for (; i + 224 < a_hi; i += 256) {
                double a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = src[i + 32 * u];
#pragma unroll
                for (int u = 0; u < 8; ++u) s += a[u];
            }
            for (; i < a_hi; i += 32) s += src[i];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) p.partial[(long long)row * Gs + c] = s;
        }
        {
            unsigned long long mn = ~0ULL, ms = 0ULL, mx = 0ULL;  // smallest errmax (highest slot among ties), largest
            for (long long i = lo + tid; i < hi; i += IV_THREADS) {
                const unsigned long long k = iv_key(pool.errmax[i]);
                if (k <= mn) {
                    mn = k;
                    ms = (unsigned long long)i;
                }
                mx = k > mx ? k : mx;
            }
            S.red_a[tid] = mn;
            S.red_b[tid] = ms;
            S.red_c[tid] = mx;
            __syncthreads();
            for (int q = IV_THREADS / 2; q > 0; q >>= 1) {
                if (tid < q) {
                    const unsigned long long k2 = S.red_a[tid + q], s2 = S.red_b[tid + q];
                    if (k2 < S.red_a[tid] || (k2 == S.red_a[tid] && s2 > S.red_b[tid])) {
                        S.red_a[tid] = k2;
                        S.red_b[tid] = s2;
                    }
                    if (S.red_c[tid + q] > S.red_c[tid]) S.red_c[tid] = S.red_c[tid + q];
                }
                __syncthreads();
            }
            if (tid == 0) {
                p.blk[b] = S.red_a[0];
                p.blk[Gs + b] = S.red_b[0];
                p.blk[2 * Gs + b] = S.red_c[0];
            }
        }
        IV_SYNC();
        IV_STAMP(2);
        // ---- totals in CTA order (spread over the grid), the smallest region's vectors (CTA 0) ---------------------------
        for (long long g = (long long)b * IV_WARPS + warp; g < 2 * f; g += (long long)Gp * IV_WARPS) {
            double s = 0.0;  // lane l adds chunks l, l + 32, ..., then a shuffle tree: fixed order
            for (int c = (int)lane; c < nch; c += 32) s += __ldcg(&p.partial[g * Gs + c]);
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) p.totals[g] = s;
        }
        if (b == 0) {
            unsigned long long mn = ~0ULL, ms = 0ULL, mx = 0ULL;
            for (int i = tid; i < Gp; i += IV_THREADS) {
                const unsigned long long k = __ldcg(&p.blk[i]), sl = __ldcg(&p.blk[Gs + i]), m2 = __ldcg(&p.blk[2 * Gs + i]);
                if (k < mn || (k == mn && sl > ms)) {
                    mn = k;
                    ms = sl;
                }
                mx = m2 > mx ? m2 : mx;
            }
            __syncthreads();
            S.red_a[tid] = mn;
            S.red_b[tid] = ms;
            S.red_c[tid] = mx;
            __syncthreads();
            for (int q = IV_THREADS / 2; q > 0; q >>= 1) {
                if (tid < q) {
                    const unsigned long long k2 = S.red_a[tid + q], s2 = S.red_b[tid + q];
                    if (k2 < S.red_a[tid] || (k2 == S.red_a[tid] && s2 > S.red_b[tid])) {
                        S.red_a[tid] = k2;
                        S.red_b[tid] = s2;
                    }
                    if (S.red_c[tid + q] > S.red_c[tid]) S.red_c[tid] = S.red_c[tid + q];
                }
                __syncthreads();
            }
            const long long argmin = (long long)S.red_b[0];
            for (int j = tid; j < f; j += IV_THREADS) {
                p.totals[2 * f + j] = pool.val[(long long)j * pool.cap + argmin];
                p.totals[3 * f + j] = pool.err[(long long)j * pool.cap + argmin];
                p.sabove[j] = 0.0;
            }
            if (tid == 0) {
                p.blk[4 * Gs + 0] = S.red_a[0];
                p.blk[4 * Gs + 1] = S.red_c[0];
                p.state->M = 0ULL;
                p.state->K = 0ULL;
                p.state->fail = 0ULL;
                p.state->kmin[0] = p.state->kmin[1] = p.state->kmin[2] = ~0ULL;
            }
        }
        IV_SYNC();
        IV_STAMP(3);
        // ---- B: decisions, redundantly in every CTA (identical inputs) ----------------------------------------------------
        for (int j = tid; j < 4 * f; j += IV_THREADS) tv[j] = __ldcg(&p.totals[j]);
        __syncthreads();
        const double* tV = tv;
        const double* tE = tv + f;
        const double* le = tv + 3 * f;
        const unsigned long long kmin = __ldcg(&p.blk[4 * Gs + 0]), kmax = __ldcg(&p.blk[4 * Gs + 1]);
        const long long numEval = c_numEval, maxEval = c_maxEval, P = c_P;
        const double eps = 2.220446049250313e-16;
        const double band = 4.0 * eps * sqrt(3.0 * (double)N);  // rounding band of the reference's drifting sums (diagnostic)
        if (warp < 6) {  // six predicate evaluations side by side, one per warp
            bool r;
            if (warp == 0)
                r = iv_conv_warp(f, iv_band_acc(tV, tE, tE, 1.0, 0.0), p.absTol, p.relTol, p.norm, wscr);
            else if (warp == 1)
                r = iv_conv_warp(f, iv_band_acc(tV, tE, tE, 1.0 + band, 0.0), p.absTol, p.relTol, p.norm, wscr);
            else if (warp == 2)
                r = iv_conv_warp(f, iv_band_acc(tV, tE, tE, 1.0 - band, 0.0), p.absTol, p.relTol, p.norm, wscr);
            else if (warp == 3)
                r = iv_conv_warp(f, iv_band_acc(tV, le, tE, 1.0, 0.0), p.absTol, p.relTol, p.norm, wscr);
            else if (warp == 4)
                r = iv_conv_warp(f, iv_band_acc(tV, le, tE, 1.0, band), p.absTol, p.relTol, p.norm, wscr);
            else
                r = iv_conv_warp(f, iv_band_acc(tV, le, tE, 1.0, -band), p.absTol, p.relTol, p.norm, wscr);
            if (lane == 0) S.s_pred[warp] = r ? 1 : 0;
        }
        __syncthreads();
        if (tid == 0) {
            const bool budget_left = maxEval == 0 || numEval < maxEval;  // Rule.java:72
            long long Kcap = N;
            if (maxEval > 0 && budget_left) {
                const long long room = maxEval - numEval, k = (room + 2 * P - 1) / (2 * P);
                Kcap = k < 1 ? 1 : k;
                if (Kcap > N) Kcap = N;
            }
            int action = IV_ACT_CUT, state = LS_RUN, amb = 0, count_only = 0;
            const bool c0 = S.s_pred[0] != 0;
            if (!budget_left) {
                action = IV_ACT_FINISH;
                state = LS_MAXEVAL;
            } else if (c_stop_at > 0 && N >= c_stop_at) {  // multi-GPU: the host deals the pool out to the ranks
                action = IV_ACT_FINISH;
                state = LS_HANDOFF;
            } else if (N + (N < Kcap ? N : Kcap) > c_cap) {
                action = IV_ACT_FINISH;
                state = LS_GROW;
            } else if (c_iter >= c_maxiter) {
                action = IV_ACT_FINISH;
                state = LS_GUARD;
            } else {
                if (c0 != (S.s_pred[1] != 0) || c0 != (S.s_pred[2] != 0)) amb += 1;
                if (c0) {  // Rule.java:73
                    action = IV_ACT_FINISH;
                    state = LS_CONVERGED;
                } else {
                    bool all = !S.s_pred[3];
                    if ((!S.s_pred[4]) != all || (!S.s_pred[5]) != all) amb += 1;
                    if (N == 1) all = true;
                    if (all && Kcap >= N)
                        action = IV_ACT_POPALL;
                    else
                        count_only = all ? 1 : 0;  // the budget alone ends the batch
                }
            }
            S.action = action;
            S.state = state;
            S.amb = amb;
            S.count_only = count_only;
            S.Kcap = Kcap;
        }
        __syncthreads();
        const int action = S.action;
        IV_STAMP(4);
        long long K = 0;
        int use_list = 0, amb_total = S.amb;
        if (action == IV_ACT_POPALL) {
            K = N;
        } else if (action == IV_ACT_CUT) {
            // ---- C: the batch cut ------------------------------------------------------------------------------------
            const long long Kcap = S.Kcap;
            const int count_only = S.count_only;
            if (b == 0 && tid == 0) p.ctl[LC_STAMP + 13] = 0;
            unsigned long long Cabove = 0;
            int sab_par = 0;
            bool match_all = N <= IV_CMAX;
            iv_u128 prefix = 0;
            int sh = 0;  // candidates: (comp >> sh) == prefix
            if (!match_all) {
                const unsigned long long diff = kmax ^ kmin;
                const int hb = diff ? 32 + (63 - __clzll((long long)diff)) : 31;
                int s = (hb / 4) * 4;
                prefix = iv_comp(kmax, 0) >> (s + 4);
                int Gc = Gp < p.gc_max ? Gp : p.gc_max;
                if (Gc < 1) Gc = 1;
                const long long perC = (N + Gc - 1) / Gc;
                for (;;) {
                    IV_STAMP(15);
                    // -- per-CTA bucket sums of the regions that match the prefix --
                    if (b < Gc) {
                        const long long clo = perC * b;
                        long long chi = clo + perC;
                        if (chi > N) chi = N;
                        if (clo > chi) chi = clo;
                        double* part = p.bsum_part + (long long)b * IV_NB * f;
                        for (int x = tid; x < IV_NB * f; x += IV_THREADS) part[x] = 0.0;
                        if (tid < IV_NB) S.s_cnt[tid] = 0u;
                        __syncthreads();
                        for (long long t_lo = clo; t_lo < chi; t_lo += IV_TILE) {
                            const int tn = (int)(chi - t_lo < IV_TILE ? chi - t_lo : IV_TILE);
                            // the tile's matching regions, in slot order: local index and digit (ordered compaction: ballots + a
                            // scan over the warps' counts per round of IV_THREADS elements)
                            int nm = 0;
                            for (int r0 = 0; r0 < tn; r0 += IV_THREADS) {
                                const int i = r0 + tid;
                                unsigned d = 255u;
                                if (i < tn) {
                                    const iv_u128 c = iv_comp(iv_key(pool.errmax[t_lo + i]), t_lo + i);
                                    if ((c >> (s + 4)) == prefix) {
                                        d = (unsigned)(c >> s) & (IV_NB - 1);
                                        atomicAdd(&S.s_cnt[d], 1u);
                                    }
                                }
                                const unsigned bal = __ballot_sync(0xffffffffu, d != 255u);
                                if (lane == 0) S.wcnt[warp] = (unsigned)__popc(bal);
                                __syncthreads();
                                unsigned off = 0, tot = 0;
                                for (int w = 0; w < IV_WARPS; ++w) {
                                    if (w < warp) off += S.wcnt[w];
                                    tot += S.wcnt[w];
                                }
                                if (d != 255u) {
                                    const int at = nm + (int)off + __popc(bal & ((1u << lane) - 1u));
                                    S.midx[at] = (unsigned short)i;
                                    S.sidx[at] = (unsigned char)d;
                                }
                                nm += (int)tot;
                                __syncthreads();
                            }
                            if (nm > 0) {
                                for (int j = warp; j < f; j += IV_WARPS) {
                                    const double* src = pool.err + (long long)j * pool.cap + t_lo;
                                    double acc[IV_NB];
#pragma unroll
                                    for (int q = 0; q < IV_NB; ++q) acc[q] = 0.0;
                                    for (int m0 = 0; m0 < nm; m0 += 8 * 32) {  // lane l takes the matches l, l + 32, ...: eight loads in flight
                                        double e[8];
                                        unsigned dg[8];
#pragma unroll
                                        for (int u = 0; u < 8; ++u) {
                                            const int m = m0 + 32 * u + (int)lane;
                                            const bool in = m < nm;
                                            dg[u] = in ? (unsigned)S.sidx[m] : 255u;
                                            e[u] = in ? src[S.midx[m]] : 0.0;
                                        }
#pragma unroll
                                        for (int u = 0; u < 8; ++u) {
#pragma unroll
                                            for (int q = 0; q < IV_NB; ++q) acc[q] += (dg[u] == (unsigned)q) ? e[u] : 0.0;
                                        }
                                    }
#pragma unroll
                                    for (int o = 16; o > 0; o >>= 1) {  // sixteen shuffle trees side by side
#pragma unroll
                                        for (int q = 0; q < IV_NB; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
                                    }
                                    double mine = 0.0;
#pragma unroll
                                    for (int q = 0; q < IV_NB; ++q)
                                        if ((int)lane == q) mine = acc[q];
                                    if (lane < IV_NB) part[(long long)lane * f + j] += mine;  // cell owned by this warp
                                }
                            }
                            __syncthreads();
                        }
                        if (tid < IV_NB) p.bcnt_part[(long long)b * IV_NB + tid] = (unsigned long long)S.s_cnt[tid];
                    }
                    IV_SYNC();
                    IV_STAMP(5);
                    // -- bucket sums of the partition, in CTA order --
                    for (long long x = (long long)b * IV_WARPS + warp; x < (long long)IV_NB * f + IV_NB; x += (long long)Gp * IV_WARPS) {
                        if (x < (long long)IV_NB * f) {
                            double sacc = 0.0;  // lane l adds CTAs l, l + 32, ..., then a shuffle tree: fixed order
                            for (int c = (int)lane; c < Gc; c += 32) sacc += __ldcg(&p.bsum_part[(long long)c * IV_NB * f + x]);
                            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, o);
                            if (lane == 0) p.bsum[x] = sacc;
                        } else {
                            const int q = (int)(x - (long long)IV_NB * f);
                            unsigned long long cacc = 0;
                            for (int c = (int)lane; c < Gc; c += 32) cacc += __ldcg(&p.bcnt_part[(long long)c * IV_NB + q]);
                            for (int o = 16; o > 0; o >>= 1) cacc += __shfl_down_sync(0xffffffffu, cacc, o);
                            if (lane == 0) p.bcnt[q] = cacc;
                        }
                    }
                    IV_SYNC();
                    IV_STAMP(6);
                    // -- the bucket in which the batch ends: buckets in descending order, the first non-empty one at which
                    //    the budget runs out or the residual passes (every CTA, same numbers) --
                    const double* sab = p.sabove + (long long)sab_par * f;
                    if (tid < IV_NB) S.bc[tid] = __ldcg(&p.bcnt[tid]);
                    if (!use_ss) __syncthreads();
                    if (use_ss) {  // suffix sums over the buckets (the order of IvAcc kind 1) and the popped sum, in shared memory
                        for (int j = tid; j < f; j += IV_THREADS) {
                            double bq[IV_NB];
#pragma unroll
                            for (int t = 0; t < IV_NB; ++t) bq[t] = __ldcg(&p.bsum[(long long)t * f + j]);
                            double sacc = 0.0;
#pragma unroll
                            for (int t = IV_NB - 1; t >= 0; --t) {
                                sacc += bq[t];
                                ss[t * f + j] = sacc;
                            }
                            ss[IV_NB * f + j] = __ldcg(&sab[j]);
                        }
                        __syncthreads();
                    }
                    for (int q = warp; q < IV_NB; q += IV_WARPS) {  // one warp per bucket
                        unsigned long long cq = Cabove;
                        for (int t = IV_NB - 1; t >= q; --t) cq += S.bc[t];
                        const bool nonempty = S.bc[q] > 0ULL;
                        bool stop = cq >= (unsigned long long)Kcap;
                        if (!stop && !count_only && nonempty)
                            stop = use_ss ? iv_conv_warp(f, iv_suffix_acc(tV, tE, ss + IV_NB * f, ss + q * f), p.absTol, p.relTol, p.norm, wscr)
                                          : iv_conv_warp(f, iv_bucket_acc(tV, tE, sab, p.bsum, f, q), p.absTol, p.relTol, p.norm, wscr);
                        if (lane == 0) S.s_stop[q] = (stop && nonempty) ? 1 : 0;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        int chosen = -1, lowest = -1;
                        for (int q = IV_NB - 1; q >= 0; --q) {
                            if (S.bc[q] > 0ULL) lowest = q;
                            if (chosen < 0 && S.s_stop[q]) chosen = q;
                        }
                        if (chosen < 0) chosen = lowest;
                        S.chosen = chosen;
                    }
                    __syncthreads();
                    const int chosen = S.chosen;
                    if (chosen < 0) {  // cannot happen: the prefix matches at least one region
                        if (b == 0 && tid == 0) p.state->fail = 1ULL;
                        break;
                    }
                    for (int t = IV_NB - 1; t > chosen; --t) Cabove += S.bc[t];
                    const unsigned long long cnt_in = S.bc[chosen];
                    if (b == 0) {  // the buckets above the chosen one join the popped sum (same order as IvAcc kind 1)
                        double* nsab = p.sabove + (long long)(sab_par ^ 1) * f;
                        for (int j = tid; j < f; j += IV_THREADS) {
                            double sacc = 0.0;
                            if (use_ss) {
                                if (chosen < IV_NB - 1) sacc = ss[(chosen + 1) * f + j];
                            } else {
                                for (int t = IV_NB - 1; t > chosen; --t) sacc += __ldcg(&p.bsum[(long long)t * f + j]);
                            }
                            nsab[j] = __ldcg(&sab[j]) + sacc;
                        }
                    }
                    sab_par ^= 1;
                    IV_STAMP(7);
                    if (b == 0 && tid == 0) p.ctl[LC_STAMP + 13] += 1;  // narrowing levels run
                    prefix = (prefix << 4) | (iv_u128)chosen;
                    sh = s;
                    if (cnt_in <= (unsigned long long)IV_CMAX || s == 0) break;
                    s -= 4;
                }
            }
            // -- gather the candidates --
            {
                for (long long i = lo + tid; i < hi; i += IV_THREADS) {
                    const iv_u128 c = iv_comp(iv_key(pool.errmax[i]), i);
                    if (match_all || (c >> sh) == prefix) {
                        const unsigned long long pos = atomicAdd(&p.state->M, 1ULL);
                        if (pos < (unsigned long long)IV_CMAX) {
                            p.cand_hi[pos] = (unsigned long long)(c >> 32);
                            p.cand_lo[pos] = (unsigned)c;
                        }
                    }
                }
            }
            IV_SYNC();
            IV_STAMP(8);
            // -- sort the candidates by rank counting, one warp per candidate over the whole grid (the composite keys are
            //    distinct): candidate i goes to position #{keys greater than its own} of the sorted arrays --
            const unsigned long long Mll = __ldcg(&p.state->M);
            const int M = (int)(Mll < (unsigned long long)IV_CMAX ? Mll : (unsigned long long)IV_CMAX);
            for (int i = warp * Gp + b; i < M; i += IV_WARPS * Gp) {
                const unsigned long long mh = __ldcg(&p.cand_hi[i]);
                const unsigned ml = __ldcg(&p.cand_lo[i]);
                int greater = 0;
                for (int q = (int)lane; q < M; q += 32) {
                    const unsigned long long qh = __ldcg(&p.cand_hi[q]);
                    const unsigned ql = __ldcg(&p.cand_lo[q]);
                    greater += (qh > mh || (qh == mh && ql > ml)) ? 1 : 0;
                }
                for (int o = 16; o > 0; o >>= 1) greater += __shfl_xor_sync(0xffffffffu, greater, o);
                if (lane == 0) {
                    p.sort_hi[greater] = mh;
                    p.sort_lo[greater] = ml;
                }
            }
            if (b == 0 && tid == 0 && (Mll > (unsigned long long)IV_CMAX || M < 1)) p.state->fail = 2ULL;
            IV_SYNC();
            IV_STAMP(14);
            const double* sab_f = p.sabove + (long long)sab_par * f;
            const int cs = (M + IV_CHUNKS - 1) / IV_CHUNKS > 0 ? (M + IV_CHUNKS - 1) / IV_CHUNKS : 1;  // candidates per chunk
            if (!count_only) {
                // -- chunk-local running sums of err over the sorted candidates: one thread per (chunk, component), whole grid --
                const int jblocks = (f + 31) / 32;
                for (long long g = (long long)warp * Gp + b; g < (long long)IV_CHUNKS * jblocks; g += (long long)IV_WARPS * Gp) {  // one warp per (chunk, 32 components)
                    const int w = (int)(g / jblocks), j = (int)(g - (long long)w * jblocks) * 32 + (int)lane;
                    if (j >= f) continue;
                    const int k0 = w * cs, k1 = (k0 + cs < M ? k0 + cs : M);
                    const double* src = pool.err + (long long)j * pool.cap;
                    double run = 0.0;
                    int k = k0;
                    for (; k + 8 <= k1; k += 8) {
                        double a[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) a[u] = src[0xffffffffu - __ldcg(&p.sort_lo[k + u])];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            run += a[u];
                            p.cres[(long long)(k + u) * f + j] = run;
                        }
                    }
                    for (; k < k1; ++k) {
                        run += src[0xffffffffu - __ldcg(&p.sort_lo[k])];
                        p.cres[(long long)k * f + j] = run;
                    }
                    p.cbase[(long long)w * f + j] = run;  // the chunk's total for now
                }
                IV_SYNC();
                // -- residual before the first candidate of every chunk --
                for (long long j = (long long)b * IV_THREADS + tid; j < f; j += (long long)Gp * IV_THREADS) {
                    double popped = __ldcg(&sab_f[j]);
                    for (int w = 0; w < IV_CHUNKS; ++w) {
                        const double t = __ldcg(&p.cbase[(long long)w * f + j]);
                        p.cbase[(long long)w * f + j] = tE[j] - popped;
                        popped += t;
                    }
                }
                IV_SYNC();
                // -- the predicate for every batch size (and its sensitivity to the rounding band), spread over the grid --
                for (long long item = (long long)warp * Gp + b; item < (long long)3 * M; item += (long long)IV_WARPS * Gp) {  // one warp per item
                    const int k = (int)(item / 3), v = (int)(item - (long long)k * 3);
                    const int w = k / cs;
                    const double add = v == 0 ? 0.0 : (v == 1 ? band : -band);
                    const bool c = iv_conv_warp(f, iv_cand_acc(tV, tE, p.cbase + (long long)w * f, p.cres + (long long)k * f, add), p.absTol, p.relTol, p.norm, wscr);
                    if (c && lane == 0) atomicMin(&p.state->kmin[v], (unsigned long long)k);
                }
                IV_SYNC();
            }
            if (b == 0 && tid == 0 && M >= 1) {
                const unsigned long long room = (unsigned long long)Kcap - Cabove;  // >= 1: the budget did not end above
                unsigned long long kk[3];
                for (int q = 0; q < 3; ++q) {
                    const unsigned long long km = __ldcg(&p.state->kmin[q]);
                    unsigned long long take = km == ~0ULL ? (unsigned long long)M : km + 1ULL;
                    if (take > room) take = room;
                    if (take > (unsigned long long)M) take = (unsigned long long)M;
                    if (take < 1ULL) take = 1ULL;
                    kk[q] = take;
                }
                const int kp = (int)kk[0] - 1;
                p.state->thr_hi = __ldcg(&p.sort_hi[kp]);
                p.state->thr_lo = (unsigned long long)__ldcg(&p.sort_lo[kp]);
                p.state->K = Cabove + kk[0];
                p.state->pad[0] = (kk[1] != kk[0] || kk[2] != kk[0]) ? 1ULL : 0ULL;
            }
            IV_SYNC();
            IV_STAMP(9);
            if (__ldcg(&p.state->fail)) {
                if (b == 0 && tid == 0) {
                    ctl[LC_DBG0] = (long long)__ldcg(&p.state->fail);
                    ctl[LC_STATE] = LS_ERROR;
                    ctl[LC_ERR] = LE_INTERNAL;
                    ctl[LC_DYN_N] = 0;
                }
                goto done;
            }
            // ---- D: stable compaction of the popped slots -----------------------------------------------------------------
            const iv_u128 thr = ((iv_u128)__ldcg(&p.state->thr_hi) << 32) | (iv_u128)__ldcg(&p.state->thr_lo);
            K = (long long)__ldcg(&p.state->K);
            amb_total += (int)__ldcg(&p.state->pad[0]);
            use_list = 1;
            {
                unsigned long long mine = 0;
                for (long long i = lo + tid; i < hi; i += IV_THREADS) mine += iv_comp(iv_key(pool.errmax[i]), i) >= thr ? 1ULL : 0ULL;
                S.red_a[tid] = mine;
                __syncthreads();
                for (int q = IV_THREADS / 2; q > 0; q >>= 1) {
                    if (tid < q) S.red_a[tid] += S.red_a[tid + q];
                    __syncthreads();
                }
                if (tid == 0) p.blk[3 * Gs + b] = S.red_a[0];
            }
            IV_SYNC();
            IV_STAMP(10);
            {
                unsigned long long before = 0;
                for (int c = tid; c < b; c += IV_THREADS) before += __ldcg(&p.blk[3 * Gs + c]);
                __syncthreads();
                S.red_a[tid] = before;
                __syncthreads();
                for (int q = IV_THREADS / 2; q > 0; q >>= 1) {
                    if (tid < q) S.red_a[tid] += S.red_a[tid + q];
                    __syncthreads();
                }
                unsigned long long base_pos = S.red_a[0];
                __syncthreads();
                for (long long t_lo = lo; t_lo < hi; t_lo += IV_THREADS) {
                    const long long i = t_lo + tid;
                    const bool pop = i < hi && iv_comp(iv_key(pool.errmax[i]), i) >= thr;
                    const unsigned bal = __ballot_sync(0xffffffffu, pop);
                    if (lane == 0) S.wcnt[warp] = (unsigned)__popc(bal);
                    __syncthreads();
                    unsigned off = 0, tot = 0;
                    for (int w = 0; w < IV_WARPS; ++w) {
                        if (w < warp) off += S.wcnt[w];
                        tot += S.wcnt[w];
                    }
                    if (pop) p.list[base_pos + off + (unsigned)__popc(bal & ((1u << lane) - 1u))] = (int)i;
                    base_pos += tot;
                    __syncthreads();
                }
                if (b == Gp - 1 && tid == 0 && base_pos != (unsigned long long)K) {  // the list and the batch size disagree
                    ctl[LC_DBG0] = (long long)base_pos;
                    p.state->fail = 3ULL;
                }
            }
            IV_SYNC();
            IV_STAMP(11);
            if (__ldcg(&p.state->fail)) {
                if (b == 0 && tid == 0) {
                    ctl[LC_STATE] = LS_ERROR;
                    ctl[LC_ERR] = LE_INTERNAL;
                    ctl[LC_DYN_N] = 0;
                }
                goto done;
            }
        }
        if (action == IV_ACT_FINISH) {
            if (b == 0 && tid == 0) {
                ctl[LC_STATE] = S.state;
                ctl[LC_ERR] = 0;
                ctl[LC_DYN_N] = 0;
                ctl[LC_AMBIG] += amb_total;
            }
            goto done;
        }
        // ---- bisection (Region.cut): the left child takes the parent's slot, the right child slot N + i -------------------------
        for (long long i = (long long)b * IV_THREADS + tid; i < K; i += (long long)Gp * IV_THREADS)
            iv_bisect_one(pool, p.dim, use_list ? (long long)p.list[i] : i, N + i);
        IV_STAMP(12);
        if (b == 0 && tid == 0) {
            ctl[LC_DYN_N] = 2 * K;
            ctl[LC_DYN_LIST] = use_list;
            ctl[LC_DYN_BASE] = N;
            ctl[LC_K] = K;
            ctl[LC_KGLOBAL] = K;
            ctl[LC_N] = N + K;
            ctl[LC_NGLOBAL] = N + K;
            ctl[LC_NUMEVAL] += 2 * P * K;
            ctl[LC_EVAL_LOCAL] += 2 * K;
            ctl[LC_WORDS + (ctl[LC_NBATCH] % LC_LOG_CAP)] = 2 * K;
            ctl[LC_NBATCH] += 1;
            ctl[LC_ITER] += 1;
            ctl[LC_NSELECT] += use_list;
            ctl[LC_AMBIG] += amb_total;
        }
    }
done:
    // the last CTA out clears the barrier words for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned long long prev = atomicAdd(&bar[1], 1ULL);
        if (prev + 1ULL == (unsigned long long)Gs) {
            bar[0] = 0ULL;
            bar[1] = 0ULL;
        }
    }
#undef IV_SYNC
}

}  // namespace

size_t iv_dyn_smem(int fdim) { return ((size_t)4 * fdim + (fdim <= IV_SS_FMAX ? (size_t)(IV_NB + 1) * fdim : 0)) * sizeof(double); }

static int g_iv_blocks_per_sm = -1, g_iv_fdim = -1;
int iv_max_blocks(int n_sms, int fdim) {
    if (g_iv_blocks_per_sm < 0 || g_iv_fdim != fdim) {
        int nb = 0;
        cudaFuncSetAttribute(k_vec_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iv_dyn_smem(fdim));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_vec_iter, IV_THREADS, iv_dyn_smem(fdim));
        g_iv_blocks_per_sm = nb > 0 ? (nb > 2 ? 2 : nb) : 1;
        g_iv_fdim = fdim;
    }
    return n_sms * g_iv_blocks_per_sm;
}

cudaError_t iv_iter(const IterVParams& p, int n_blocks, cudaStream_t s) {
    IterVParams q = p;
    void* args[1] = {&q};
    return cudaLaunchCooperativeKernel((const void*)k_vec_iter, dim3((unsigned)n_blocks), dim3(IV_THREADS), args, iv_dyn_smem(p.fdim), s);
}
