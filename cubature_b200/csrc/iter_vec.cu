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

struct IvBandAcc {  // err_j * scale + add * base_j (shared-memory vectors)
    const double* v;
    const double* e;
    const double* base;
    double scale, add;
    __device__ double val(int j) const { return v[j]; }
    __device__ double err(int j) const { return e[j] * scale + add * base[j]; }
};
// residual after popping the buckets q .. IV_NB-1 of a narrowing level on top of `sab`
struct IvBucketAcc {
    const double* v;      // shared
    const double* E;      // shared
    const double* sab;    // global
    const double* bsum;   // global [IV_NB][f]
    int f, q;
    __device__ double val(int j) const { return v[j]; }
    __device__ double err(int j) const {
        double s = 0.0;
        for (int t = IV_NB - 1; t >= q; --t) s += __ldcg(&bsum[(long long)t * f + j]);
        return E[j] - (__ldcg(&sab[j]) + s);
    }
};
// residual after popping the sorted candidates 0 .. k
struct IvCandAcc {
    const double* v;      // shared
    const double* E;      // shared
    const double* cbase;  // global: residual before the first candidate of k's chunk
    const double* crow;   // global: chunk-local running sum, row k
    double add;
    __device__ double val(int j) const { return v[j]; }
    __device__ double err(int j) const { return (cbase[j] - crow[j]) + add * E[j]; }
};

struct IvShared {
    unsigned long long k_hi[IV_CMAX];
    unsigned k_lo[IV_CMAX];
    unsigned char sidx[IV_TILE];
    unsigned long long red_a[IV_THREADS], red_b[IV_THREADS], red_c[IV_THREADS];
    unsigned s_cnt[IV_NB];
    int s_pred[8];
    int s_stop[IV_NB];
    int action, state, count_only, chosen, amb, tile_any;
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
    extern __shared__ double tv[];  // [4 f]: Sum val, Sum err, val / err of the smallest region
    __shared__ IvShared S;
    long long* ctl = p.ctl;
    const int tid = threadIdx.x, b = blockIdx.x, Gs = gridDim.x, f = p.fdim;
    const unsigned lane = tid & 31u;
    const int warp = tid >> 5;
    if (ctl[LC_STATE] != LS_RUN) {  // finished or paused: this launch and the ones queued behind it do nothing
        if (b == 0 && tid == 0) ctl[LC_DYN_N] = 0;
        return;
    }
    // the loop state as it is at launch: CTA 0 changes the control block at the end of this launch, so every CTA reads what
    // it needs before the first grid barrier
    const long long N = ctl[LC_N], c_numEval = ctl[LC_NUMEVAL], c_maxEval = ctl[LC_MAXEVAL], c_P = ctl[LC_P], c_cap = ctl[LC_CAP],
                    c_iter = ctl[LC_ITER], c_maxiter = ctl[LC_MAXITER];
    const PoolView& pool = p.pool;
    unsigned long long* const bar = p.bar;
    // ---- errmax of the regions the last rule launch wrote (separable GK15 path), by the whole grid ---------------------
    if (p.need_errmax) {
        const long long n_items = ctl[LC_DYN_N], use_list = ctl[LC_DYN_LIST], base = ctl[LC_DYN_BASE];
        for (long long i = (long long)b * IV_THREADS + tid; i < n_items; i += (long long)Gs * IV_THREADS) {
            const long long pi = i >> 1;
            const long long slot = (i & 1) ? base + pi : (use_list ? (long long)p.list[pi] : pi);
            double emax = 0.0;
            for (int j = 0; j < f; ++j) {
                const double e = pool.err[(long long)j * pool.cap + slot];
                if (e > emax) emax = e;
            }
            pool.errmax[slot] = emax;
            pool.splitdim[slot] = 0;
        }
    }
    // every CTA of the grid has read the loop state before anybody changes it
    iv_gsync(bar, (unsigned long long)Gs);
    long long Gp_ll = (N + 1023) / 1024;
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
        // ---- A: chunk partials ------------------------------------------------------------------------------------
        for (int row = warp; row < 2 * f; row += IV_WARPS) {
            const double* src = row < f ? pool.val + (long long)row * pool.cap : pool.err + (long long)(row - f) * pool.cap;
            double s = 0.0;  // fixed order: lane l adds elements l, l + 32, ... of the chunk, then a shuffle tree
            long long i = lo + lane;
            for (; i + 96 < hi; i += 128) {
                const double a0 = src[i], a1 = src[i + 32], a2 = src[i + 64], a3 = src[i + 96];
                s += a0;
                s += a1;
                s += a2;
                s += a3;
            }
            for (; i < hi; i += 32) s += src[i];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) p.partial[(long long)row * Gs + b] = s;
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
        // ---- totals in CTA order (spread over the grid), the smallest region's vectors (CTA 0) ---------------------------
        for (long long g = (long long)b * IV_THREADS + tid; g < 2 * f; g += (long long)Gp * IV_THREADS) {
            double s = 0.0;
            for (int c = 0; c < Gp; ++c) s += __ldcg(&p.partial[g * Gs + c]);
            p.totals[g] = s;
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
            }
        }
        IV_SYNC();
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
        if (lane == 0 && warp < 6) {  // six predicate evaluations side by side, one per warp
            bool r;
            if (warp == 0)
                r = cub_converged(f, IvBandAcc{tV, tE, tE, 1.0, 0.0}, p.absTol, p.relTol, p.norm);
            else if (warp == 1)
                r = cub_converged(f, IvBandAcc{tV, tE, tE, 1.0 + band, 0.0}, p.absTol, p.relTol, p.norm);
            else if (warp == 2)
                r = cub_converged(f, IvBandAcc{tV, tE, tE, 1.0 - band, 0.0}, p.absTol, p.relTol, p.norm);
            else if (warp == 3)
                r = cub_converged(f, IvBandAcc{tV, le, tE, 1.0, 0.0}, p.absTol, p.relTol, p.norm);
            else if (warp == 4)
                r = cub_converged(f, IvBandAcc{tV, le, tE, 1.0, band}, p.absTol, p.relTol, p.norm);
            else
                r = cub_converged(f, IvBandAcc{tV, le, tE, 1.0, -band}, p.absTol, p.relTol, p.norm);
            S.s_pred[warp] = r ? 1 : 0;
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
        long long K = 0;
        int use_list = 0, amb_total = S.amb;
        if (action == IV_ACT_POPALL) {
            K = N;
        } else if (action == IV_ACT_CUT) {
            // ---- C: the batch cut ------------------------------------------------------------------------------------
            const long long Kcap = S.Kcap;
            const int count_only = S.count_only;
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
                            if (tid == 0) S.tile_any = 0;
                            __syncthreads();
                            bool any = false;
                            for (int i = tid; i < tn; i += IV_THREADS) {
                                const iv_u128 c = iv_comp(iv_key(pool.errmax[t_lo + i]), t_lo + i);
                                unsigned char d = 255;
                                if ((c >> (s + 4)) == prefix) {
                                    d = (unsigned char)((unsigned)(c >> s) & (IV_NB - 1));
                                    atomicAdd(&S.s_cnt[d], 1u);
                                    any = true;
                                }
                                S.sidx[i] = d;
                            }
                            if (any) S.tile_any = 1;
                            __syncthreads();
                            if (S.tile_any) {
                                for (int j = warp; j < f; j += IV_WARPS) {
                                    const double* src = pool.err + (long long)j * pool.cap + t_lo;
                                    double acc[IV_NB];
#pragma unroll
                                    for (int q = 0; q < IV_NB; ++q) acc[q] = 0.0;
                                    for (int i = (int)lane; i < tn; i += 32) {
                                        const unsigned d = S.sidx[i];
                                        if (d != 255u) {
                                            const double e = src[i];
#pragma unroll
                                            for (int q = 0; q < IV_NB; ++q) acc[q] += (d == (unsigned)q) ? e : 0.0;
                                        }
                                    }
                                    double mine = 0.0;
#pragma unroll
                                    for (int q = 0; q < IV_NB; ++q) {
                                        double v = acc[q];
                                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                                        if ((int)lane == q) mine = v;
                                    }
                                    if (lane < IV_NB) part[(long long)lane * f + j] += mine;  // cell owned by this warp
                                }
                            }
                            __syncthreads();
                        }
                        if (tid < IV_NB) p.bcnt_part[(long long)b * IV_NB + tid] = (unsigned long long)S.s_cnt[tid];
                    }
                    IV_SYNC();
                    // -- bucket sums of the partition, in CTA order --
                    for (long long x = (long long)b * IV_THREADS + tid; x < (long long)IV_NB * f + IV_NB; x += (long long)Gp * IV_THREADS) {
                        if (x < (long long)IV_NB * f) {
                            double sacc = 0.0;
                            for (int c = 0; c < Gc; ++c) sacc += __ldcg(&p.bsum_part[(long long)c * IV_NB * f + x]);
                            p.bsum[x] = sacc;
                        } else {
                            const int q = (int)(x - (long long)IV_NB * f);
                            unsigned long long cacc = 0;
                            for (int c = 0; c < Gc; ++c) cacc += __ldcg(&p.bcnt_part[(long long)c * IV_NB + q]);
                            p.bcnt[q] = cacc;
                        }
                    }
                    IV_SYNC();
                    // -- the bucket in which the batch ends: buckets in descending order, the first non-empty one at which
                    //    the budget runs out or the residual passes (every CTA, same numbers) --
                    const double* sab = p.sabove + (long long)sab_par * f;
                    if (tid < IV_NB) {
                        const int q = tid;
                        unsigned long long cq = Cabove;
                        for (int t = IV_NB - 1; t >= q; --t) cq += __ldcg(&p.bcnt[t]);
                        bool stop = cq >= (unsigned long long)Kcap;
                        if (!stop && !count_only && __ldcg(&p.bcnt[q]) > 0ULL)
                            stop = cub_converged(f, IvBucketAcc{tV, tE, sab, p.bsum, f, q}, p.absTol, p.relTol, p.norm);
                        S.s_stop[q] = (stop && __ldcg(&p.bcnt[q]) > 0ULL) ? 1 : 0;
                    }
                    __syncthreads();
                    if (tid == 0) {
                        int chosen = -1, lowest = -1;
                        for (int q = IV_NB - 1; q >= 0; --q) {
                            if (__ldcg(&p.bcnt[q]) > 0ULL) lowest = q;
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
                    for (int t = IV_NB - 1; t > chosen; --t) Cabove += __ldcg(&p.bcnt[t]);
                    const unsigned long long cnt_in = __ldcg(&p.bcnt[chosen]);
                    if (b == 0) {  // the buckets above the chosen one join the popped sum (same order as IvBucketAcc)
                        double* nsab = p.sabove + (long long)(sab_par ^ 1) * f;
                        for (int j = tid; j < f; j += IV_THREADS) {
                            double sacc = 0.0;
                            for (int t = IV_NB - 1; t > chosen; --t) sacc += __ldcg(&p.bsum[(long long)t * f + j]);
                            nsab[j] = __ldcg(&sab[j]) + sacc;
                        }
                    }
                    sab_par ^= 1;
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
            // -- CTA 0: sort, running residuals, the predicate for every batch size --
            if (b == 0) {
                const double* sab = p.sabove + (long long)sab_par * f;
                const unsigned long long Mll = __ldcg(&p.state->M);
                const int M = (int)(Mll < (unsigned long long)IV_CMAX ? Mll : (unsigned long long)IV_CMAX);
                int M2 = 2;
                while (M2 < M) M2 <<= 1;
                for (int i = tid; i < M2; i += IV_THREADS) {
                    S.k_hi[i] = i < M ? __ldcg(&p.cand_hi[i]) : 0ULL;
                    S.k_lo[i] = i < M ? __ldcg(&p.cand_lo[i]) : 0u;
                }
                __syncthreads();
                for (int kk = 2; kk <= M2; kk <<= 1) {  // bitonic sort, descending (padding keys are the smallest)
                    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                        for (int i = tid; i < M2; i += IV_THREADS) {
                            const int l = i ^ jj;
                            if (l > i) {
                                const unsigned long long ah = S.k_hi[i], bh = S.k_hi[l];
                                const unsigned al = S.k_lo[i], bl = S.k_lo[l];
                                const bool a_lt_b = ah < bh || (ah == bh && al < bl);
                                const bool desc = (i & kk) == 0;
                                if (desc ? a_lt_b : !a_lt_b && !(ah == bh && al == bl)) {
                                    S.k_hi[i] = bh;
                                    S.k_lo[i] = bl;
                                    S.k_hi[l] = ah;
                                    S.k_lo[l] = al;
                                }
                            }
                        }
                        __syncthreads();
                    }
                }
                const int cs = (M + IV_CHUNKS - 1) / IV_CHUNKS;  // candidates per chunk (>= 1 when M >= 1)
                {   // chunk-local running sums of err over the sorted candidates, one chunk per warp
                    const int k0 = warp * cs, k1 = (k0 + cs < M ? k0 + cs : M);
                    for (int j = (int)lane; j < f; j += 32) {
                        const double* src = pool.err + (long long)j * pool.cap;
                        double run = 0.0;
                        for (int k = k0; k < k1; ++k) {
                            run += src[0xffffffffu - S.k_lo[k]];
                            p.cres[(long long)k * f + j] = run;
                        }
                        p.cbase[(long long)warp * f + j] = run;  // the chunk's total for now
                    }
                }
                __syncthreads();
                for (int j = tid; j < f; j += IV_THREADS) {  // residual before the first candidate of every chunk
                    double popped = __ldcg(&sab[j]);
                    for (int w = 0; w < IV_CHUNKS; ++w) {
                        const double t = p.cbase[(long long)w * f + j];
                        p.cbase[(long long)w * f + j] = tE[j] - popped;
                        popped += t;
                    }
                }
                if (tid < 3) S.kmin[tid] = ~0ULL;
                __syncthreads();
                if (!count_only) {
                    for (int k = tid; k < M; k += IV_THREADS) {
                        const int w = k / cs;
                        const double* cb = p.cbase + (long long)w * f;
                        const double* cr = p.cres + (long long)k * f;
                        if (cub_converged(f, IvCandAcc{tV, tE, cb, cr, 0.0}, p.absTol, p.relTol, p.norm)) atomicMin(&S.kmin[0], (unsigned long long)k);
                        if (cub_converged(f, IvCandAcc{tV, tE, cb, cr, band}, p.absTol, p.relTol, p.norm)) atomicMin(&S.kmin[1], (unsigned long long)k);
                        if (cub_converged(f, IvCandAcc{tV, tE, cb, cr, -band}, p.absTol, p.relTol, p.norm)) atomicMin(&S.kmin[2], (unsigned long long)k);
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    const unsigned long long room = (unsigned long long)Kcap - Cabove;  // >= 1: the budget did not end above
                    unsigned long long kk[3];
                    for (int q = 0; q < 3; ++q) {
                        unsigned long long take = S.kmin[q] == ~0ULL ? (unsigned long long)M : S.kmin[q] + 1ULL;
                        if (take > room) take = room;
                        if (take > (unsigned long long)M) take = (unsigned long long)M;
                        if (take < 1ULL) take = 1ULL;
                        kk[q] = take;
                    }
                    if (kk[1] != kk[0] || kk[2] != kk[0]) S.amb = 1;
                    else S.amb = 0;
                    const int kp = (int)kk[0] - 1;
                    p.state->thr_hi = S.k_hi[kp];
                    p.state->thr_lo = (unsigned long long)S.k_lo[kp];
                    p.state->K = Cabove + kk[0];
                    p.state->pad[0] = (unsigned long long)S.amb;
                    if (Mll > (unsigned long long)IV_CMAX || M < 1) p.state->fail = 2ULL;
                }
            }
            IV_SYNC();
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

size_t iv_dyn_smem(int fdim) { return (size_t)4 * fdim * sizeof(double); }

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
