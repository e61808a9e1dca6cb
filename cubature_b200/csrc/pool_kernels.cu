// pool_kernels.cu -- region-pool kernels that do not depend on the integrand (compiled by nvcc for
// sm_100a, -fmad=false): stand-alone bisection (K4), errmax rows, pool totals (K5), and the
// device-side batch selection (K3): stable radix sort of the regions by errmax, per-component prefix
// sums in that order, and the parallel evaluation of the reference's residual test for every k.
//
// All of these are HBM-bound passes over the structure-of-arrays pool; they use coalesced 8-byte
// accesses, fixed-shape reductions (bit-reproducible run to run) and integer atomics only where the
// result is order-independent.
//
// Reference semantics reproduced (paths under src/main/java/de/labathome/cubature/):
//   Region.cut                  Region.java:31-38           -> k_bisect
//   Rule.errMax                 Rule.java:281-289           -> k_errmax_rows
//   RegionHeap running sums     RegionHeap.java:21-36       -> k_totals_* (recomputed, not incremental)
//   Gladwell batch pop          Rule.java:101-133           -> k_sort_*, k_scan_*, k_find_k
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "converge.h"
#include "pool_kernels.h"

namespace {

__device__ __forceinline__ unsigned long long dbits(double x) { return (unsigned long long)__double_as_longlong(x); }

// ---- K4: stand-alone bisection (used when several threads share a region, fdim > 1) ---------------
__global__ void k_bisect(PoolView pool, int dim, const int* list, long long K, long long base) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const long long parent = list ? (long long)list[i] : i;
    const long long right = base + i;
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

// ---- errmax over the components of each evaluated item (fdim > 1 separable GK path) ---------------
__global__ void k_errmax_rows(PoolView pool, int fdim, const int* list, long long n_items, long long base, int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    long long slot;
    if (mode == 0) {
        slot = list ? (long long)list[i] : base + i;
    } else {
        const long long pi = i >> 1;
        slot = (i & 1) ? base + pi : (list ? (long long)list[pi] : pi);
    }
    double emax = 0.0;
    for (int j = 0; j < fdim; ++j) {
        const double e = pool.err[(long long)j * pool.cap + slot];
        if (e > emax) emax = e;
    }
    pool.errmax[slot] = emax;
    pool.splitdim[slot] = 0;
}

// ---- K5: totals over the active regions [0, n) ----------------------------------------------------
// grid = (TOT_BLOCKS, 2*fdim): blockIdx.y < fdim sums val row j, else err row j - fdim.
// Fixed shape (TOT_BLOCKS x 256 threads, strided then tree) => bit-reproducible for a given n.
constexpr int TOT_THREADS = 256;

__global__ void k_totals_partial(PoolView pool, int fdim, long long n, double* partial /*[2f][gridDim.x]*/) {
    const int row = blockIdx.y;
    const double* src = row < fdim ? pool.val + (long long)row * pool.cap : pool.err + (long long)(row - fdim) * pool.cap;
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = per * blockIdx.x;
    long long hi = lo + per;
    if (hi > n) hi = n;
    double s = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += TOT_THREADS) s += src[i];
    __shared__ double sh[TOT_THREADS];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = TOT_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(long long)row * gridDim.x + blockIdx.x] = sh[0];
}

__global__ void k_totals_final(const double* partial, int nblk, double* totals /*[2f]*/) {
    const int row = blockIdx.x;
    __shared__ double sh[TOT_THREADS];
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += TOT_THREADS) s += partial[(long long)row * nblk + i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = TOT_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[row] = sh[0];
}

// min / max errmax key over [0, n) and, among the regions holding the minimum, the highest slot
// (the region popped last under the "lower slot first" tie rule).  Integer atomics: order-independent.
__global__ void k_minmax_key(const double* errmax, long long n, unsigned long long* out /*[0]=min,[1]=max*/) {
    unsigned long long mn = ~0ULL, mx = 0ULL;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = dbits(errmax[i]);
        mn = k < mn ? k : mn;
        mx = k > mx ? k : mx;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], mn);
        atomicMax(&out[1], mx);
    }
}
__global__ void k_argmin_slot(const double* errmax, long long n, const unsigned long long* minmax, unsigned long long* out_slot) {
    const unsigned long long mn = minmax[0];
    unsigned long long best = 0;
    bool any = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (dbits(errmax[i]) == mn) {
            best = (unsigned long long)i;  // ascending i: the last one seen is the largest
            any = true;
        }
    }
    if (any) atomicMax(out_slot, best);
}

// gathers one column (slot) of val / err into a dense [2f] vector
__global__ void k_gather_slot(PoolView pool, int fdim, const unsigned long long* slot_ptr, double* out /*[2f]*/) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= fdim) return;
    const long long s = (long long)*slot_ptr;
    out[j] = pool.val[(long long)j * pool.cap + s];
    out[fdim + j] = pool.err[(long long)j * pool.cap + s];
}

// ---- K5 fused: one launch per iteration computes everything the loop control needs ------------------
// Every CTA reduces a contiguous chunk (same fixed shape as k_totals_partial: bit-reproducible) and its
// min / max errmax key; the CTA that finishes last (ticket) adds the partials in CTA order, finds the
// smallest-errmax region (highest slot among ties: the region popped last), gathers its val/err column,
// evaluates the reference's convergence test (Rule.java:73) and the select-all shortcut, writes the
// device control block and clears the scratch of the selection kernels.  Control block (8-byte words):
//   [0] n_items  [1] use_list  [2] base        <- read by the rule kernel (CubWork::dyn)
//   [3] K        [4] done      [5] all   [6] ambiguous   [8] min key  [9] max key  [10] argmin slot
//   [16 ..) totals val[f], err[f], then the smallest region's val[f], err[f]
struct BandAcc {  // err_j * scale + add * base_j
    const double* v;
    const double* e;
    const double* base;
    double scale, add;
    __device__ double val(int j) const { return v[j]; }
    __device__ double err(int j) const { return e[j] * scale + add * base[j]; }
};

__global__ void __launch_bounds__(TOT_THREADS)
k_pool_stats(PoolView pool, int fdim, long long n, double* partial /*[2f][G]*/, unsigned long long* blk /*[3][G]: min key, its slot, max key*/,
             unsigned* ticket, long long* ctl, long long Kcap, double absTol, double relTol, int norm, double band,
             unsigned long long* sel_zero, int sel_zero_words) {
    __shared__ double sh[TOT_THREADS];
    __shared__ unsigned long long shk[TOT_THREADS], shs[TOT_THREADS], shm[TOT_THREADS];
    __shared__ bool s_last;
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const long long per = (n + G - 1) / G;
    const long long lo = per * b;
    long long hi = lo + per;
    if (hi > n) hi = n;
    for (int row = 0; row < 2 * fdim; ++row) {
        const double* src = row < fdim ? pool.val + (long long)row * pool.cap : pool.err + (long long)(row - fdim) * pool.cap;
        double s = 0.0;
        for (long long i = lo + tid; i < hi; i += TOT_THREADS) s += src[i];
        sh[tid] = s;
        __syncthreads();
        for (int w = TOT_THREADS / 2; w > 0; w >>= 1) {
            if (tid < w) sh[tid] += sh[tid + w];
            __syncthreads();
        }
        if (tid == 0) partial[(long long)row * G + b] = sh[0];
        __syncthreads();
    }
    {
        unsigned long long mn = ~0ULL, ms = 0ULL, mx = 0ULL;
        for (long long i = lo + tid; i < hi; i += TOT_THREADS) {
            const unsigned long long k = dbits(pool.errmax[i]);
            if (k <= mn) {  // ascending i: among equal keys the highest slot wins
                mn = k;
                ms = (unsigned long long)i;
            }
            mx = k > mx ? k : mx;
        }
        shk[tid] = mn;
        shs[tid] = ms;
        shm[tid] = mx;
        __syncthreads();
        for (int w = TOT_THREADS / 2; w > 0; w >>= 1) {
            if (tid < w) {
                const unsigned long long k2 = shk[tid + w], s2 = shs[tid + w];
                if (k2 < shk[tid] || (k2 == shk[tid] && s2 > shs[tid])) {
                    shk[tid] = k2;
                    shs[tid] = s2;
                }
                if (shm[tid + w] > shm[tid]) shm[tid] = shm[tid + w];
            }
            __syncthreads();
        }
        if (tid == 0) {
            blk[b] = shk[0];
            blk[G + b] = shs[0];
            blk[2 * G + b] = shm[0];
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == (unsigned)G - 1;
    __syncthreads();
    if (!s_last) return;
    // ---- last CTA -------------------------------------------------------------------------------
    double* totals = (double*)(ctl + 16);
    double* minvec = totals + 2 * fdim;
    for (int row = 0; row < 2 * fdim; ++row) {
        double s = 0.0;
        for (int i = tid; i < G; i += TOT_THREADS) s += __ldcg(&partial[(long long)row * G + i]);
        sh[tid] = s;
        __syncthreads();
        for (int w = TOT_THREADS / 2; w > 0; w >>= 1) {
            if (tid < w) sh[tid] += sh[tid + w];
            __syncthreads();
        }
        if (tid == 0) totals[row] = sh[0];
        __syncthreads();
    }
    {
        unsigned long long mn = ~0ULL, ms = 0ULL, mx = 0ULL;
        for (int i = tid; i < G; i += TOT_THREADS) {
            const unsigned long long k = __ldcg(&blk[i]), sl = __ldcg(&blk[G + i]), m2 = __ldcg(&blk[2 * G + i]);
            if (k < mn || (k == mn && sl > ms)) {
                mn = k;
                ms = sl;
            }
            mx = m2 > mx ? m2 : mx;
        }
        shk[tid] = mn;
        shs[tid] = ms;
        shm[tid] = mx;
        __syncthreads();
        for (int w = TOT_THREADS / 2; w > 0; w >>= 1) {
            if (tid < w) {
                const unsigned long long k2 = shk[tid + w], s2 = shs[tid + w];
                if (k2 < shk[tid] || (k2 == shk[tid] && s2 > shs[tid])) {
                    shk[tid] = k2;
                    shs[tid] = s2;
                }
                if (shm[tid + w] > shm[tid]) shm[tid] = shm[tid + w];
            }
            __syncthreads();
        }
    }
    const long long argmin = (long long)shs[0];
    for (int j = tid; j < fdim; j += TOT_THREADS) {
        minvec[j] = pool.val[(long long)j * pool.cap + argmin];
        minvec[fdim + j] = pool.err[(long long)j * pool.cap + argmin];
    }
    for (int i = tid; i < sel_zero_words; i += TOT_THREADS) sel_zero[i] = 0ULL;
    __syncthreads();
    if (tid != 0) return;
    *ticket = 0u;
    ((unsigned long long*)ctl)[8] = shk[0];
    ((unsigned long long*)ctl)[9] = shm[0];
    ctl[10] = argmin;
    const double* tv = totals;
    const double* te = totals + fdim;
    long long amb = 0;
    const bool c0 = cub_converged(fdim, BandAcc{tv, te, te, 1.0, 0.0}, absTol, relTol, norm);
    const bool c1 = cub_converged(fdim, BandAcc{tv, te, te, 1.0 + band, 0.0}, absTol, relTol, norm);
    const bool c2 = cub_converged(fdim, BandAcc{tv, te, te, 1.0 - band, 0.0}, absTol, relTol, norm);
    if (c0 != c1 || c0 != c2) amb += 1;
    ctl[0] = 0;
    ctl[1] = 0;
    ctl[2] = n;
    ctl[3] = 0;
    ctl[4] = c0 ? 1 : 0;
    ctl[5] = 0;
    if (!c0) {
        // select-all shortcut: the residual after popping all but the smallest-errmax region is that
        // region's own error vector; if that is not converged, no smaller k is (monotone predicate)
        const double* le = minvec + fdim;
        bool all = !cub_converged(fdim, BandAcc{tv, le, te, 1.0, 0.0}, absTol, relTol, norm);
        const bool a1 = !cub_converged(fdim, BandAcc{tv, le, te, 1.0, band}, absTol, relTol, norm);
        const bool a2 = !cub_converged(fdim, BandAcc{tv, le, te, 1.0, -band}, absTol, relTol, norm);
        if (a1 != all || a2 != all) amb += 1;
        if (n == 1) all = true;
        if (all && Kcap >= n) {
            ctl[0] = 2 * n;
            ctl[3] = n;
            ctl[5] = 1;
        }
    }
    ctl[6] = amb;
}

// ---- K3a: stable LSD radix sort of (key = ~bits(errmax), slot), 8 bits per pass --------------------
// Ascending order of ~bits == descending errmax; stability keeps equal keys in ascending slot order.
// One warp owns a contiguous chunk of SORT_CHUNK elements; per-(bin, chunk) counts are scanned per
// bin (256 blocks) and across bins (one tiny block), then each warp scatters its chunk in order.
constexpr int SORT_WARPS = 8;
constexpr int SORT_CHUNK = 4096;

__global__ void k_sort_prepare(const double* errmax, long long n, unsigned long long* keys, int* idx,
                               unsigned long long* or_and /*[0]=OR,[1]=AND*/) {
    unsigned long long o = 0ULL, a = ~0ULL;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k = ~dbits(errmax[i]);
        keys[i] = k;
        idx[i] = (int)i;
        o |= k;
        a &= k;
    }
    for (int s = 16; s > 0; s >>= 1) {
        o |= __shfl_xor_sync(0xffffffffu, o, s);
        a &= __shfl_xor_sync(0xffffffffu, a, s);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicOr(&or_and[0], o);
        atomicAnd(&or_and[1], a);
    }
}

__global__ void __launch_bounds__(SORT_WARPS * 32)
k_sort_hist(const unsigned long long* keys, long long n, int shift, long long nchunks, unsigned* hist /*[256][nchunks]*/) {
    __shared__ unsigned sh[SORT_WARPS][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = lane; b < 256; b += 32) sh[w][b] = 0;
    __syncwarp();
    const long long chunk = (long long)blockIdx.x * SORT_WARPS + w;
    if (chunk < nchunks) {
        const long long lo = chunk * SORT_CHUNK;
        long long hi = lo + SORT_CHUNK;
        if (hi > n) hi = n;
        for (long long i = lo + lane; i < hi; i += 32) atomicAdd(&sh[w][(unsigned)(keys[i] >> shift) & 255u], 1u);
        __syncwarp();
        for (int b = lane; b < 256; b += 32) hist[(long long)b * nchunks + chunk] = sh[w][b];
    }
}

// exclusive scan of each bin's row (one block per bin); totals[bin] = row sum
__global__ void k_sort_scan_rows(unsigned* hist, long long nchunks, unsigned long long* totals) {
    const int bin = blockIdx.x;
    unsigned* row = hist + (long long)bin * nchunks;
    __shared__ unsigned long long sh[256];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nchunks; base += 256) {
        const long long i = base + threadIdx.x;
        const unsigned v = i < nchunks ? row[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {  // Hillis-Steele inclusive scan
            unsigned long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0ULL;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nchunks) row[i] = (unsigned)(carry + sh[threadIdx.x] - v);
        __syncthreads();
        if (threadIdx.x == 255) carry += sh[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[bin] = carry;
}
__global__ void k_sort_scan_bins(const unsigned long long* totals, unsigned long long* binbase) {
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int b = 0; b < 256; ++b) {
            binbase[b] = s;
            s += totals[b];
        }
    }
}

__global__ void __launch_bounds__(SORT_WARPS * 32)
k_sort_scatter(const unsigned long long* keys, const int* idx, long long n, int shift, long long nchunks,
               const unsigned* hist, const unsigned long long* binbase, unsigned long long* out_keys, int* out_idx) {
    __shared__ unsigned long long off[SORT_WARPS][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chunk = (long long)blockIdx.x * SORT_WARPS + w;
    if (chunk >= nchunks) return;
    for (int b = lane; b < 256; b += 32) off[w][b] = binbase[b] + hist[(long long)b * nchunks + chunk];
    __syncwarp();
    const long long lo = chunk * SORT_CHUNK;
    long long hi = lo + SORT_CHUNK;
    if (hi > n) hi = n;
    const unsigned lt = (1u << lane) - 1u;
    for (long long base = lo; base < hi; base += 32) {
        const long long i = base + lane;
        const bool active = i < hi;
        unsigned long long k = 0;
        int v = 0;
        unsigned d = 256u + (unsigned)lane;  // inactive lanes never match anybody
        if (active) {
            k = keys[i];
            v = idx[i];
            d = (unsigned)(k >> shift) & 255u;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned long long pos = 0;
        if (active) pos = off[w][d] + (unsigned long long)__popc(peers & lt);
        __syncwarp();
        if (active && (peers & lt) == 0u) off[w][d] += (unsigned long long)__popc(peers);  // lowest peer lane updates
        __syncwarp();
        if (active) {
            out_keys[pos] = k;
            out_idx[pos] = v;
        }
    }
}

// ---- K3b: per-component inclusive prefix sums of err in sorted order ------------------------------
// tile = SCAN_THREADS * SCAN_ITEMS consecutive sorted positions; local inclusive prefix + tile totals,
// tile offsets by a second single-block pass per component.  Fixed shape => reproducible.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(PoolView pool, const int* perm, long long n, double* prefix /*[f][n]*/, long long ntiles, double* tile_tot /*[f][ntiles]*/) {
    const int j = blockIdx.y;
    const long long tile = blockIdx.x;
    const double* src = pool.err + (long long)j * pool.cap;
    const long long lo = tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    double v[SCAN_ITEMS];
    double run = 0.0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) {
        const long long k = lo + t;
        const double e = k < n ? src[perm[k]] : 0.0;
        run += e;
        v[t] = run;
    }
    __shared__ double sh[SCAN_THREADS];
    sh[threadIdx.x] = run;
    __syncthreads();
    for (int o = 1; o < SCAN_THREADS; o <<= 1) {
        const double t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0.0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    const double before = threadIdx.x > 0 ? sh[threadIdx.x - 1] : 0.0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) {
        const long long k = lo + t;
        if (k < n) prefix[(long long)j * n + k] = before + v[t];
    }
    if (threadIdx.x == SCAN_THREADS - 1) tile_tot[(long long)j * ntiles + tile] = sh[SCAN_THREADS - 1];
}

__global__ void k_scan_tile_offsets(double* tile_tot, long long ntiles) {  // exclusive, in place, one block per component
    double* row = tile_tot + (long long)blockIdx.x * ntiles;
    __shared__ double sh[256];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (long long base = 0; base < ntiles; base += 256) {
        const long long i = base + threadIdx.x;
        const double v = i < ntiles ? row[i] : 0.0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const double t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0.0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < ntiles) row[i] = carry + (sh[threadIdx.x] - v);
        __syncthreads();
        if (threadIdx.x == 255) carry += sh[255];
        __syncthreads();
    }
}

// ---- K3c: smallest k such that the residual after popping the k largest regions is converged ------
// Rule.java:121-133: ee.err -= popped.err ; converged(ee) ? break.  Thread k evaluates the predicate
// on  err_total - prefix[k]  (val_total unchanged, Rule.java:122-124 touches only err); three variants
// (exact, residual +band, residual -band) give the decision and its sensitivity to the rounding of the
// reference's running sums.  out[0..2] = min k+1 over converged threads (initialised to n+1 by the host).
struct ResidualAcc {
    const double* tot;  // [2f]: val then err
    const double* prefix;
    const double* tile_off;
    long long n, ntiles, k;
    int fdim;
    double band;  // relative band applied to the residual: r +- band * err_total
    __device__ double val(int j) const { return tot[j]; }
    __device__ double err(int j) const {
        const double p = tile_off[(long long)j * ntiles + k / SCAN_TILE] + prefix[(long long)j * n + k];
        return (tot[fdim + j] - p) + band * tot[fdim + j];
    }
};

__global__ void k_find_k(const double* totals, const double* prefix, const double* tile_off, long long n, long long ntiles,
                         int fdim, double absTol, double relTol, int norm, double band, unsigned long long* out) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    ResidualAcc acc{totals, prefix, tile_off, n, ntiles, k, fdim, 0.0};
    if (cub_converged(fdim, acc, absTol, relTol, norm)) atomicMin(&out[0], (unsigned long long)(k + 1));
    acc.band = band;
    if (cub_converged(fdim, acc, absTol, relTol, norm)) atomicMin(&out[1], (unsigned long long)(k + 1));
    acc.band = -band;
    if (cub_converged(fdim, acc, absTol, relTol, norm)) atomicMin(&out[2], (unsigned long long)(k + 1));
}

// ---- K3 for scalar integrands: weighted radix select (no sort) ------------------------------------
// With fdim == 1 the residual test depends on ONE running sum, so the batch size K (Rule.java:109-133)
// can be found by descending the 64-bit key space 8 bits at a time: every pass histograms the still
// undecided regions by their next digit (count and sum of err per bucket, largest errmax first), a
// one-thread kernel walks the 256 buckets accumulating the popped error until either the residual
// test passes inside a bucket or the pop budget Kcap (maxEval, Rule.java:133) runs out inside it, and
// descends into that bucket.  After 8 passes the threshold key is exact; regions with exactly that
// errmax are taken in ascending slot order.  8 light passes over errmax/err (16 B per region) replace
// a full sort (~250 B per region); the list of popped slots is then produced by a stable compaction.
struct SelState {
    unsigned long long prefix;     // decided high bits of the threshold key (key = ~bits(errmax))
    unsigned long long mask;       // which bits are decided
    unsigned long long count_cum;  // regions strictly ahead of the current bucket (all popped)
    unsigned long long K;          // result: number of regions to pop
    unsigned long long tie_take;   // how many of the regions with key == prefix are popped
    unsigned long long ambiguous;  // decision within the rounding band
    double P_cum;                  // sum of err of the count_cum regions
};

struct ScalarAcc {
    double v, e;
    __device__ double val(int) const { return v; }
    __device__ double err(int) const { return e; }
};

constexpr int SEL_THREADS = 256;
constexpr int SEL_WARPS = SEL_THREADS / 32;
constexpr int SEL_UNROLL = 4;

// Digit schedule of the descent over key = ~bits(errmax): eight 8-bit digits, most significant first.
// Every pass streams errmax of the whole pool (8 B per region, SEL_UNROLL independent loads in flight
// per thread) but only the regions inside the bucket chosen so far take part in the histogram (after
// two passes: about 1 % of the pool), so passes 2..7 are pure streaming reads.
__device__ __forceinline__ void sel_flush(unsigned* scnt, double* ssum, unsigned cur, unsigned run_c, double run_s) {
    if (run_c) {
        atomicAdd(&scnt[cur], run_c);
        atomicAdd(&ssum[cur], run_s);
    }
}

// Walks the 256 buckets of one pass (largest errmax first), accumulating the popped error until the
// residual test passes inside a bucket or the pop budget runs out inside it; descends into that bucket.
// returns 1 when the batch size is final (last pass, or the chosen bucket holds a single region: then only
// that region's complete key is still missing, see k_iter_coop), else 0
__device__ int sel_pick_body(SelState* st, const unsigned long long* scnt, const double* ssum, const unsigned* nonzero /*[8] bit masks*/,
                             int shift, int last, double V, double E, unsigned long long Kcap, double absTol, double relTol, int norm,
                             double band, int early_exit = 0) {
    unsigned long long count_cum = st->count_cum;
    double P = st->P_cum;
    int chosen = -1;
    bool stop = false;
    for (int w = 0; w < 8 && !stop; ++w) {  // only the non-empty buckets, in ascending order
        unsigned m = nonzero[w];
        while (m) {
            const int d = w * 32 + (__ffs(m) - 1);
            m &= m - 1;
            const unsigned long long c = scnt[d];
            const double s = ssum[d];
            chosen = d;  // the last non-empty bucket is chosen if nothing stops earlier
            stop = (count_cum + c >= Kcap) || cub_converged(1, ScalarAcc{V, E - (P + s)}, absTol, relTol, norm);
            if (stop) break;
            count_cum += c;
            P += s;
        }
    }
    if (chosen < 0) {  // cannot happen for n >= 1
        st->K = 0;
        return 1;
    }
    const unsigned long long c_b = scnt[chosen];
    const double s_b = ssum[chosen];
    st->prefix |= (unsigned long long)chosen << shift;
    st->mask |= 255ULL << shift;
    st->count_cum = count_cum;
    st->P_cum = P;
    if (last || (early_exit && c_b == 1)) {
        // the bucket is a single key: c_b regions with identical errmax; each carries e = s_b / c_b
        const double e = s_b / (double)c_b;
        unsigned long long room = Kcap - count_cum;  // >= 1
        unsigned long long mmax = c_b < room ? c_b : room;
        // smallest m in [1, mmax] whose residual passes (monotone), else mmax
        unsigned long long lo = 1, hi = mmax;
        if (cub_converged(1, ScalarAcc{V, E - (P + (double)mmax * e)}, absTol, relTol, norm)) {
            while (lo < hi) {
                const unsigned long long mid = lo + (hi - lo) / 2;
                if (cub_converged(1, ScalarAcc{V, E - (P + (double)mid * e)}, absTol, relTol, norm))
                    hi = mid;
                else
                    lo = mid + 1;
            }
        } else {
            lo = mmax;
        }
        st->tie_take = lo;
        st->K = count_cum + lo;
        // sensitivity of the decision to the rounding of the reference's running sums
        const double rK = E - (P + (double)lo * e), rKm1 = rK + e, b = band * E;
        const bool a0 = cub_converged(1, ScalarAcc{V, rK + b}, absTol, relTol, norm) != cub_converged(1, ScalarAcc{V, rK - b}, absTol, relTol, norm);
        const bool a1 = cub_converged(1, ScalarAcc{V, rKm1 + b}, absTol, relTol, norm) != cub_converged(1, ScalarAcc{V, rKm1 - b}, absTol, relTol, norm);
        st->ambiguous = (a0 || a1) ? 1ULL : 0ULL;
        return 1;
    }
    return 0;
}

// One pass: histogram of the pool regions that match the decided prefix by the digit (key >> shift) & 255.
// Every warp owns a private copy of the histogram in shared memory (double atomicAdd on shared memory is
// a compare-and-swap loop; with one copy per CTA 256 threads fight over the ~20 hot buckets).
// With fuse_pick the CTA that finishes last (ticket counter) also runs the pick, so a pass is one launch;
// the multi-GPU driver passes fuse_pick = 0 and puts the all-gather between histogram and pick.
__global__ void __launch_bounds__(SEL_THREADS)
k_sel_pass(const double* __restrict__ errmax, const double* __restrict__ err, long long n, SelState* st, int shift, int last,
           unsigned long long* cnt, double* sum, unsigned* ticket, int fuse_pick, const double* totals, unsigned long long Kcap,
           double absTol, double relTol, int norm, double band, const long long* ctl) {
    if (ctl && (ctl[4] | ctl[5])) return;  // converged, or the whole pool is popped: nothing to select
    __shared__ unsigned scnt[SEL_WARPS][256];
    __shared__ double ssum[SEL_WARPS][256];
    __shared__ bool s_last;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    for (int b = lane; b < 256; b += 32) {
        scnt[w][b] = 0;
        ssum[w][b] = 0.0;
    }
    __syncwarp();
    const unsigned long long prefix = st->prefix, mask = st->mask;
    // run-length aggregation in registers: in the first pass nearly every region shares the digit, so a
    // thread touches shared memory only when its digit changes
    unsigned cur = 0, run_c = 0;
    double run_s = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * SEL_UNROLL) {
        unsigned long long k[SEL_UNROLL];
#pragma unroll
        for (int u = 0; u < SEL_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            k[u] = i < n ? ~dbits(errmax[i]) : 0ULL;  // key 0 never matches: bit 63 of a real key is 1
        }
#pragma unroll
        for (int u = 0; u < SEL_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < n && (k[u] & mask) == prefix) {
                const unsigned d = (unsigned)(k[u] >> shift) & 255u;
                if (d != cur) {
                    sel_flush(scnt[w], ssum[w], cur, run_c, run_s);
                    cur = d;
                    run_c = 0;
                    run_s = 0.0;
                }
                ++run_c;
                run_s += err[i];
            }
        }
    }
    sel_flush(scnt[w], ssum[w], cur, run_c, run_s);
    __syncthreads();
    {
        const int b = threadIdx.x;
        unsigned c = 0;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < SEL_WARPS; ++q) {
            c += scnt[q][b];
            s += ssum[q][b];
        }
        if (c) {
            atomicAdd(&cnt[b], (unsigned long long)c);
            atomicAdd(&sum[b], s);
        }
    }
    if (!fuse_pick) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    // last CTA: stage the global histogram (L2 loads), clear it for the next pass, thread 0 picks
    unsigned long long* pc = (unsigned long long*)&ssum[0][0];  // reuse shared memory: 256 u64 + 256 f64
    double* ps = &ssum[2][0];
    unsigned* nzm = &scnt[0][0];
    {
        const unsigned long long c = __ldcg(&cnt[threadIdx.x]);
        pc[threadIdx.x] = c;
        ps[threadIdx.x] = __ldcg(&sum[threadIdx.x]);
        cnt[threadIdx.x] = 0ULL;
        sum[threadIdx.x] = 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, c != 0ULL);
        if (lane == 0) nzm[w] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0u;
        sel_pick_body(st, pc, ps, nzm, shift, last, __ldcg(&totals[0]), __ldcg(&totals[1]), Kcap, absTol, relTol, norm, band);
    }
}

// stand-alone pick (multi-GPU: after the all-gather and the rank-ordered sum of the histograms)
__global__ void __launch_bounds__(SEL_THREADS)
k_sel_pick(SelState* st, unsigned long long* cnt, double* sum, int shift, int last, const double* totals, unsigned long long Kcap,
           double absTol, double relTol, int norm, double band) {
    __shared__ unsigned long long pc[256];
    __shared__ double ps[256];
    __shared__ unsigned nzm[8];
    const unsigned long long c = cnt[threadIdx.x];
    pc[threadIdx.x] = c;
    ps[threadIdx.x] = sum[threadIdx.x];
    cnt[threadIdx.x] = 0ULL;
    sum[threadIdx.x] = 0.0;
    const unsigned m = __ballot_sync(0xffffffffu, c != 0ULL);
    if ((threadIdx.x & 31) == 0) nzm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) sel_pick_body(st, pc, ps, nzm, shift, last, totals[0], totals[1], Kcap, absTol, relTol, norm, band);
}

// stable compaction of the popped slots: slot i is popped iff key_i < thr, or key_i == thr and it is
// among the first tie_take such slots.  Tiles of SEL_TILE slots; two counts per tile, scanned by one block.
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_THREADS * SEL_ITEMS;

__global__ void __launch_bounds__(SEL_THREADS)
k_sel_tile_counts(const double* errmax, long long n, const SelState* st, unsigned long long* tile_less, unsigned long long* tile_eq,
                  const long long* ctl) {
    if (ctl && (ctl[4] | ctl[5])) return;
    const unsigned long long thr = st->prefix;
    const long long lo = (long long)blockIdx.x * SEL_TILE;
    unsigned less = 0, eq = 0;
    for (int t = 0; t < SEL_ITEMS; ++t) {
        const long long i = lo + (long long)t * SEL_THREADS + threadIdx.x;
        if (i < n) {
            const unsigned long long k = ~dbits(errmax[i]);
            less += k < thr;
            eq += k == thr;
        }
    }
    __shared__ unsigned sl[SEL_THREADS], se[SEL_THREADS];
    sl[threadIdx.x] = less;
    se[threadIdx.x] = eq;
    __syncthreads();
    for (int w = SEL_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            sl[threadIdx.x] += sl[threadIdx.x + w];
            se[threadIdx.x] += se[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_less[blockIdx.x] = sl[0];
        tile_eq[blockIdx.x] = se[0];
    }
}

__global__ void k_sel_tile_scan(unsigned long long* tile_less, unsigned long long* tile_eq, long long ntiles,
                                unsigned long long* totals_out /*[2]: regions below the threshold key, regions equal to it*/,
                                const SelState* st, long long* ctl, long long n) {
    if (ctl && (ctl[4] | ctl[5])) return;
    // exclusive scans in place, one block, 256 threads
    __shared__ unsigned long long sh[2][256];
    __shared__ unsigned long long carry[2];
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (long long base = 0; base < ntiles; base += 256) {
        const long long i = base + threadIdx.x;
        const unsigned long long a = i < ntiles ? tile_less[i] : 0ULL, b = i < ntiles ? tile_eq[i] : 0ULL;
        sh[0][threadIdx.x] = a;
        sh[1][threadIdx.x] = b;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const unsigned long long ta = threadIdx.x >= o ? sh[0][threadIdx.x - o] : 0ULL;
            const unsigned long long tb = threadIdx.x >= o ? sh[1][threadIdx.x - o] : 0ULL;
            __syncthreads();
            sh[0][threadIdx.x] += ta;
            sh[1][threadIdx.x] += tb;
            __syncthreads();
        }
        if (i < ntiles) {
            tile_less[i] = carry[0] + sh[0][threadIdx.x] - a;
            tile_eq[i] = carry[1] + sh[1][threadIdx.x] - b;
        }
        __syncthreads();
        if (threadIdx.x == 255) {
            carry[0] += sh[0][255];
            carry[1] += sh[1][255];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && totals_out) {
        totals_out[0] = carry[0];
        totals_out[1] = carry[1];
    }
    if (threadIdx.x == 0 && ctl) {  // hand the batch to the rule kernel without a host round trip
        ctl[0] = 2 * (long long)st->K;
        ctl[1] = 1;
        ctl[2] = n;
        ctl[3] = (long long)st->K;
        ctl[6] += (long long)st->ambiguous;
    }
}

__global__ void __launch_bounds__(SEL_THREADS)
k_sel_compact(const double* errmax, long long n, const SelState* st, const unsigned long long* tile_less,
              const unsigned long long* tile_eq, long long tie_take_arg, int* list, const long long* ctl) {
    if (ctl && (ctl[4] | ctl[5])) return;
    // tie_take_arg < 0: take SelState::tie_take (single GPU); else this rank's share of the tied regions
    const unsigned long long thr = st->prefix, tie_take = tie_take_arg < 0 ? st->tie_take : (unsigned long long)tie_take_arg;
    // thread owns SEL_ITEMS CONSECUTIVE slots so that positions follow slot order
    const long long lo = (long long)blockIdx.x * SEL_TILE + (long long)threadIdx.x * SEL_ITEMS;
    unsigned fl = 0, fe = 0;  // bit t set: slot lo+t is less / equal
    unsigned less = 0, eq = 0;
#pragma unroll
    for (int t = 0; t < SEL_ITEMS; ++t) {
        const long long i = lo + t;
        if (i < n) {
            const unsigned long long k = ~dbits(errmax[i]);
            if (k < thr) {
                fl |= 1u << t;
                ++less;
            } else if (k == thr) {
                fe |= 1u << t;
                ++eq;
            }
        }
    }
    __shared__ unsigned sl[SEL_THREADS], se[SEL_THREADS];
    sl[threadIdx.x] = less;
    se[threadIdx.x] = eq;
    __syncthreads();
    for (int o = 1; o < SEL_THREADS; o <<= 1) {
        const unsigned ta = threadIdx.x >= o ? sl[threadIdx.x - o] : 0u, tb = threadIdx.x >= o ? se[threadIdx.x - o] : 0u;
        __syncthreads();
        sl[threadIdx.x] += ta;
        se[threadIdx.x] += tb;
        __syncthreads();
    }
    unsigned long long less_before = tile_less[blockIdx.x] + (sl[threadIdx.x] - less);
    unsigned long long eq_before = tile_eq[blockIdx.x] + (se[threadIdx.x] - eq);
#pragma unroll
    for (int t = 0; t < SEL_ITEMS; ++t) {
        const bool is_less = (fl >> t) & 1u, is_eq = (fe >> t) & 1u;
        if (is_less || (is_eq && eq_before < tie_take)) {
            const unsigned long long pos = less_before + (eq_before < tie_take ? eq_before : tie_take);
            list[pos] = (int)(lo + t);
        }
        less_before += is_less;
        eq_before += is_eq;
    }
}

// ---- one cooperative launch per iteration: K5 + loop control + K3 (scalar integrands) ----------------
// The whole host-free part of an iteration in ONE kernel, phases separated by grid-wide barriers
// (cooperative launch, every CTA co-resident).  Every CTA owns one contiguous chunk of the pool.
//   A  chunk partial sums of val/err per component, min/max errmax key, and (scalar integrands) the level-0
//      histogram of the weighted select (top 8 key bits), fused into the same read
//   B  (after barrier 1) every CTA redundantly adds the partials in CTA order -- identical bits in every
//      CTA, no second barrier -- and takes the loop decisions: converged? pop everything? (control block)
//   C  weighted radix select, coarse to fine.  key = ~bits(errmax); digits: 8 bits, then 11 bits per level.
//      Level 1 and level 2 stream the chunk again (8 B per region); the level-2 scan also copies the regions
//      that match the 19 key bits decided so far (about 0.2 % of the pool) into a candidate list and counts
//      the regions that are certainly popped; all further levels, and the per-CTA offsets of the compaction,
//      only touch the candidates.  After each level's barrier every CTA picks the bucket redundantly with a
//      block-wide prefix scan over the buckets (no serial walk).
//   D  stable compaction of the popped slots (one more read of the chunk, 4 B written per popped region)
// HBM traffic per active region and iteration: 16 B (A) + 8 B + 8 B (C) + 12 B (D) instead of one 8 B read per
// radix pass (8) plus two compaction reads.  For scalar integrands errmax is derived from err (errmax ==
// max(0, err) bit for bit in every rule kernel, Rule.java:281-289), so errmax[] is not read at all.
// The fdim > 1 path stops after B.
namespace cg = cooperative_groups;

constexpr int SEL_LEVELS = 7;
constexpr int SEL_LB = 2048;  // buckets of levels 1..5
__device__ __forceinline__ int sel_shift(int level) { return level == 0 ? 56 : (level <= 5 ? 56 - 11 * level : 0); }
// (the last level re-reads bits 7..1, which are already decided, so that it has one bucket per thread like level 0;
//  only the two buckets that match the prefix can be occupied)
#ifndef SEL_LAST_BUCKETS
#define SEL_LAST_BUCKETS 256  // (2 reproduces the observation in DESIGN.md section 9, item 5)
#endif
__device__ __forceinline__ int sel_nbuckets(int level) { return level == 0 ? 256 : (level <= 5 ? SEL_LB : SEL_LAST_BUCKETS); }
__device__ __forceinline__ unsigned long long sel_key_of_err(double e) { return ~dbits(e > 0.0 ? e : 0.0); }

struct IterParams {
    PoolView pool;
    int fdim, norm, do_select;
    long long n, Kcap;
    double absTol, relTol, band;
    double* partial;             // [2f][G]
    unsigned long long* blk;     // [3][G] min key, its slot, max key ; [2][G] less / equal counts ; [5G] unique key
    long long* ctl;              // control block (see k_pool_stats)
    SelState* st;                // final selection state (written by CTA 0)
    unsigned long long* hset;    // this launch's histogram set (PK_COOP_SET_WORDS words, zero on entry)
    unsigned long long* hzero;   // the other set: cleared by this launch for the next one
    unsigned long long* cand_key;  // [n] candidates of the level-2 scan
    int* cand_idx;                 // [n]
    int* list;
    // sharded pool (R > 1): this launch handles one shard; the per-level histograms, the totals and the tie
    // counts of the R ranks are exchanged through peer-mapped buffers over NVLink (no NCCL call, no host)
    int R, rank;
    unsigned long long* xbuf[PK_MAX_RANKS];  // [rank] = own buffer, others = peer mappings (layout: pool_kernels.h)
    unsigned long long* gblock;              // [16] local global scratch: combined totals / decisions of CTA 0
    unsigned long long seq0;                 // sequence number of this launch's first exchange
    int xsliced;                             // 1: the first CTAs share an exchange (one 256-bucket slice each); 0: CTA 0 alone
};
// histogram set layout (8-byte words): cnt[SEL_LEVELS][SEL_LB] | sum[SEL_LEVELS][SEL_LB] | misc[8] ([0] candidate count)
__device__ __forceinline__ unsigned long long* set_cnt(unsigned long long* set, int level) { return set + (size_t)level * SEL_LB; }
__device__ __forceinline__ double* set_sum(unsigned long long* set, int level) {
    return (double*)(set + (size_t)SEL_LEVELS * SEL_LB) + (size_t)level * SEL_LB;
}
__device__ __forceinline__ unsigned long long* set_misc(unsigned long long* set) { return set + (size_t)2 * SEL_LEVELS * SEL_LB; }

// ---- exchange between the ranks of a sharded pool ------------------------------------------------------------
// Push model, sliced over CTAs: CTA c < nslices owns the histogram buckets [256 c, 256 c + 256) (CTA 0 also the 16-word
// header).  It writes its slice of this rank's payload into slot [parity][rank] of EVERY rank's buffer (plain stores over
// NVLink), fences, raises flag [parity][rank][c] = seq in every buffer, waits for the R flags [parity][*][c] of its own
// buffer, then adds the R slices in rank order (identical bits on every rank) and stores the sum as this rank's
// "global" histogram.  seq counts the exchanges of the communicator's life (host-tracked: cub_comm::seq), parity =
// seq & 1: a rank can run at most one exchange ahead of its slowest peer (it needs the peer's flags of exchange e to
// finish e, and its own CTAs meet at a grid barrier after every exchange), so two payload slots suffice.  All ranks
// execute the same exchanges in the same order (their control flow depends on combined data only).  A wait longer
// than PK_XCHG_TIMEOUT_NS marks the launch failed (gblock[7]) instead of hanging the GPU.
__device__ __forceinline__ void st_release_sys(unsigned long long* ptr, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* ptr) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long* xb_payload(unsigned long long* buf, unsigned long long seq, int src_rank) {
    return buf + PK_XB_DATA + ((size_t)(seq & 1ULL) * PK_MAX_RANKS + (size_t)src_rank) * PK_XB_PAY;
}
__device__ __forceinline__ unsigned long long* xb_flag(unsigned long long* buf, unsigned long long seq, int src_rank, int slice) {
    return buf + PK_XB_FLAGS + ((size_t)(seq & 1ULL) * PK_MAX_RANKS + (size_t)src_rank) * PK_XB_SLICES + (size_t)slice;
}
// Called by every CTA (all SEL_THREADS threads); CTAs with blockIdx.x >= nslices return at once.  hdr: shared memory of
// CTA 0 (16 words).  n_cnt buckets (0: header only); cnt/sum in: this rank's histogram, out: the partition's.
__device__ void xchg_slice_body(const IterParams& p, unsigned long long seq, const unsigned long long* hdr, unsigned long long* cnt, double* sum,
                                int n_cnt, int slice);
__device__ void xchg_slice(const IterParams& p, unsigned long long seq, const unsigned long long* hdr, unsigned long long* cnt, double* sum,
                           int n_cnt) {
    const int nslices = n_cnt > SEL_THREADS ? n_cnt / SEL_THREADS : 1;
    if (p.xsliced) {
        if ((int)blockIdx.x < nslices) xchg_slice_body(p, seq, hdr, cnt, sum, n_cnt, blockIdx.x);
    } else if (blockIdx.x == 0) {  // CTA 0 runs the slices one after the other
        for (int sl = 0; sl < nslices; ++sl) xchg_slice_body(p, seq, hdr, cnt, sum, n_cnt, sl);
    }
}
__device__ void xchg_slice_body(const IterParams& p, unsigned long long seq, const unsigned long long* hdr, unsigned long long* cnt, double* sum,
                                int n_cnt, int slice) {
    const int tid = threadIdx.x;
    unsigned long long* own = p.xbuf[p.rank];
    const int q = slice * SEL_THREADS + tid;
    const bool have = q < n_cnt;
    unsigned long long c_val = 0ULL, s_val = 0ULL;
    if (have) {
        c_val = __ldcg(&cnt[q]);
        s_val = dbits(__ldcg(&sum[q]));
    }
    for (int r = 0; r < p.R; ++r) {
        unsigned long long* dst = xb_payload(p.xbuf[r], seq, p.rank);
        if (slice == 0 && tid < 16) dst[tid] = hdr[tid];
        if (have) {
            dst[16 + q] = c_val;
            dst[16 + n_cnt + q] = s_val;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid < p.R) {
        st_release_sys(xb_flag(p.xbuf[tid], seq, p.rank, slice), seq);
        const unsigned long long* f = xb_flag(own, seq, tid, slice);
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < seq) {
            if (global_timer_ns() - t0 > PK_XCHG_TIMEOUT_NS) {
                atomicExch(&p.gblock[7], 1ULL);
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
    if (have) {
        unsigned long long c = 0ULL;
        double sacc = 0.0;
        for (int r = 0; r < p.R; ++r) {
            const unsigned long long* pay = xb_payload(own, seq, r);
            c += __ldcg(&pay[16 + q]);
            sacc += __longlong_as_double((long long)__ldcg(&pay[16 + n_cnt + q]));
        }
        cnt[q] = c;
        sum[q] = sacc;
    }
}

// Block-wide pick of one level (all SEL_THREADS threads): the first bucket (ascending key = descending errmax)
// at which the pop budget runs out or the residual test passes, found with a prefix scan over the buckets.
// pc / ps: this level's global histogram staged in shared memory.  Returns 1 when the batch size is final
// (last level, or the chosen bucket holds a single region: then only that region's full key is still missing).
__device__ int sel_pick_parallel(volatile SelState* lst, const unsigned long long* pc, const double* ps, int NB, int shift, int last, double V,
                                 double E, unsigned long long Kcap, double absTol, double relTol, int norm, double band, double* sh,
                                 unsigned long long* shk, unsigned long long* shs, volatile double* s_dbl /*[2]*/, volatile long long* s_ll /*[2]*/) {
    const int tid = threadIdx.x;
    const int per_t = NB >= SEL_THREADS ? NB / SEL_THREADS : 1;
    const int first = tid * per_t;
    const unsigned long long count_cum = lst->count_cum;
    const double P_cum = lst->P_cum;
    unsigned long long c_loc[8];
    double s_loc[8];
    unsigned long long ct = 0;
    double stt = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < per_t && first + j < NB) {
            ct += pc[first + j];
            stt += ps[first + j];
        }
        c_loc[j] = ct;
        s_loc[j] = stt;
    }
    __syncthreads();  // everybody has read lst
    shk[tid] = ct;
    sh[tid] = stt;
    __syncthreads();
    for (int o = 1; o < SEL_THREADS; o <<= 1) {  // inclusive scan over the threads, fixed order
        const unsigned long long ta = tid >= o ? shk[tid - o] : 0ULL;
        const double tb = tid >= o ? sh[tid - o] : 0.0;
        __syncthreads();
        shk[tid] += ta;
        if (tid >= o) sh[tid] = tb + sh[tid];
        __syncthreads();
    }
    const unsigned long long c_off = shk[tid] - ct;
    const double s_off = tid > 0 ? sh[tid - 1] : 0.0;
    unsigned long long first_stop = ~0ULL, last_nz = 0ULL;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < per_t && first + j < NB) {
            const unsigned long long c = pc[first + j];
            if (c) {
                last_nz = (unsigned long long)(first + j) + 1ULL;
                const unsigned long long Cincl = count_cum + c_off + c_loc[j];
                const double Sincl = P_cum + (s_off + s_loc[j]);
                const bool stop = Cincl >= Kcap || cub_converged(1, ScalarAcc{V, E - Sincl}, absTol, relTol, norm);
                if (stop && first_stop == ~0ULL) first_stop = (unsigned long long)(first + j);
            }
        }
    }
    __syncthreads();
    shk[tid] = first_stop;
    shs[tid] = last_nz;
    __syncthreads();
    for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
        if (tid < q) {
            if (shk[tid + q] < shk[tid]) shk[tid] = shk[tid + q];
            if (shs[tid + q] > shs[tid]) shs[tid] = shs[tid + q];
        }
        __syncthreads();
    }
    const unsigned long long any_stop = shk[0], nz = shs[0];
    __syncthreads();
    if (nz == 0ULL) {  // cannot happen for n >= 1
        if (tid == 0) lst->K = 0;
        __syncthreads();
        return 1;
    }
    const int chosen = any_stop != ~0ULL ? (int)any_stop : (int)(nz - 1ULL);
    if (chosen >= first && chosen < first + per_t) {  // the owner publishes the descent
        const int j = chosen - first;
        unsigned long long c_prev = 0;
        double s_prev = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q == j - 1) {
                c_prev = c_loc[q];
                s_prev = s_loc[q];
            }
        lst->prefix = lst->prefix | ((unsigned long long)chosen << shift);
        lst->mask = lst->mask | ((unsigned long long)(NB - 1) << shift);
        lst->count_cum = count_cum + c_off + c_prev;
        lst->P_cum = j > 0 || tid > 0 ? P_cum + (s_off + s_prev) : P_cum;
        s_ll[0] = (long long)pc[chosen];
        s_dbl[0] = ps[chosen];

    }
    __syncthreads();
    const unsigned long long c_b = (unsigned long long)s_ll[0];
    int fin = 0;
    if (last || c_b == 1ULL) {
        fin = 1;
        if (tid == 0) {
            // the bucket is a single key: c_b regions with identical errmax; each carries e = s_b / c_b
            const double s_b = s_dbl[0], P = lst->P_cum;
            const unsigned long long cc = lst->count_cum;
            const double e = s_b / (double)c_b;
            const unsigned long long room = Kcap - cc;  // >= 1
            const unsigned long long mmax = c_b < room ? c_b : room;
            unsigned long long lo = 1, hi = mmax;  // smallest m in [1, mmax] whose residual passes (monotone), else mmax
            if (cub_converged(1, ScalarAcc{V, E - (P + (double)mmax * e)}, absTol, relTol, norm)) {
                while (lo < hi) {
                    const unsigned long long mid = lo + (hi - lo) / 2;
                    if (cub_converged(1, ScalarAcc{V, E - (P + (double)mid * e)}, absTol, relTol, norm))
                        hi = mid;
                    else
                        lo = mid + 1;
                }
            } else {
                lo = mmax;
            }
            lst->tie_take = lo;
            lst->K = cc + lo;
            // sensitivity of the decision to the rounding of the reference's running sums
            const double rK = E - (P + (double)lo * e), rKm1 = rK + e, bb = band * E;
            const bool a0 = cub_converged(1, ScalarAcc{V, rK + bb}, absTol, relTol, norm) != cub_converged(1, ScalarAcc{V, rK - bb}, absTol, relTol, norm);
            const bool a1 = cub_converged(1, ScalarAcc{V, rKm1 + bb}, absTol, relTol, norm) != cub_converged(1, ScalarAcc{V, rKm1 - bb}, absTol, relTol, norm);
            lst->ambiguous = (a0 || a1) ? 1ULL : 0ULL;
        }
    }
    __syncthreads();
    return fin;
}

__global__ void __launch_bounds__(SEL_THREADS, 4) k_iter_coop(IterParams p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[SEL_THREADS];
    __shared__ unsigned long long shk[SEL_THREADS], shs[SEL_THREADS], shm[SEL_THREADS];
    __shared__ unsigned long long pc[SEL_LB];  // staged global counts; its first half doubles as the CTA histogram counts
    __shared__ double ps[SEL_LB];              // staged global sums; doubles as the CTA histogram sums
    __shared__ SelState lst_storage;  // written by single threads, read by all: every access is volatile
    volatile SelState& lst = lst_storage;
    __shared__ long long s_flags[2];
    __shared__ double s_dbl_storage[2];
    __shared__ long long s_ll_storage[2];
    volatile double* const s_dbl = s_dbl_storage;
    volatile long long* const s_ll = s_ll_storage;
    unsigned* const scnt = (unsigned*)pc;  // [SEL_WARPS][256] at level 0, [SEL_LB] at levels >= 1
    double* const ssum = ps;
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, fdim = p.fdim;
    const unsigned lane = tid & 31u, w = tid >> 5;
    const long long n = p.n;
    const PoolView& pool = p.pool;
    const bool scalar_select = fdim == 1 && p.do_select;
    const bool sharded = p.R > 1;
    unsigned long long xseq = p.seq0;  // next exchange (uniform over the grid and over the ranks)
    __shared__ unsigned long long s_hdr[16];
    if (b == 0 && tid == 0) p.blk[5 * G] = 0ULL;  // "unique key not found here" (a real key has bit 63 set)
    // clear the OTHER histogram set for the next launch (nobody reads or writes it during this one)
    if (p.hzero)
        for (long long q = (long long)b * SEL_THREADS + tid; q < PK_COOP_SET_WORDS; q += (long long)G * SEL_THREADS) p.hzero[q] = 0ULL;
    // ---- A: chunk partials (+ level-0 histogram) -----------------------------------------------------
    const long long per = (n + G - 1) / G;
    const long long lo = per * b;
    long long hi = lo + per;
    if (hi > n) hi = n;
    if (scalar_select) {
        for (int q = tid; q < SEL_WARPS * 256; q += SEL_THREADS) {
            scnt[q] = 0u;
            ssum[q] = 0.0;
        }
        __syncthreads();
    }
    unsigned long long mn = ~0ULL, ms = 0ULL, mx = 0ULL;  // smallest errmax bits (highest slot among ties), largest
    const bool warp_rows = fdim > 4;  // many components: one warp per row (no block barrier per row), else one block per row
    if (warp_rows) {
        for (int row = (int)w; row < 2 * fdim; row += SEL_WARPS) {
            const double* src = row < fdim ? pool.val + (long long)row * pool.cap : pool.err + (long long)(row - fdim) * pool.cap;
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
            if (lane == 0) p.partial[(long long)row * G + b] = s;
        }
        __syncthreads();
    }
    for (int row = 0; row < (warp_rows ? 0 : 2 * fdim); ++row) {
        const double* src = row < fdim ? pool.val + (long long)row * pool.cap : pool.err + (long long)(row - fdim) * pool.cap;
        // fixed summation order (thread tid adds elements tid, tid+256, ... in order); the loads of four
        // consecutive steps are issued together
        double s = 0.0;
        long long i = lo + tid;
        if (scalar_select && row == 1) {
            // err row of a scalar integrand: the same read also feeds min/max and the level-0 histogram
            unsigned cur = 0, run_c = 0;
            double run_s = 0.0;
            unsigned* const myc = scnt + w * 256;
            double* const mys = ssum + w * 256;
            for (; i < hi; i += 4 * SEL_THREADS) {
                double a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = i + u * SEL_THREADS < hi ? src[i + u * SEL_THREADS] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (i + u * SEL_THREADS < hi) {
                        s += a[u];
                        const unsigned long long kb = dbits(a[u] > 0.0 ? a[u] : 0.0);
                        if (kb <= mn) {
                            mn = kb;
                            ms = (unsigned long long)(i + u * SEL_THREADS);
                        }
                        mx = kb > mx ? kb : mx;
                        const unsigned d = (unsigned)((~kb) >> 56);
                        if (d != cur) {
                            sel_flush(myc, mys, cur, run_c, run_s);
                            cur = d;
                            run_c = 0;
                            run_s = 0.0;
                        }
                        ++run_c;
                        run_s += a[u];
                    }
                }
            }
            sel_flush(myc, mys, cur, run_c, run_s);
        } else {
            for (; i + 3 * SEL_THREADS < hi; i += 4 * SEL_THREADS) {
                const double a0 = src[i], a1 = src[i + SEL_THREADS], a2 = src[i + 2 * SEL_THREADS], a3 = src[i + 3 * SEL_THREADS];
                s += a0;
                s += a1;
                s += a2;
                s += a3;
            }
            for (; i < hi; i += SEL_THREADS) s += src[i];
        }
        sh[tid] = s;
        __syncthreads();
        for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
            if (tid < q) sh[tid] += sh[tid + q];
            __syncthreads();
        }
        if (tid == 0) p.partial[(long long)row * G + b] = sh[0];
        __syncthreads();
    }
    if (!scalar_select) {  // vector integrands (and plain totals): errmax[] holds max over the components
        long long i = lo + tid;
        for (; i + 3 * SEL_THREADS < hi; i += 4 * SEL_THREADS) {
            unsigned long long k[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) k[u] = dbits(pool.errmax[i + u * SEL_THREADS]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k[u] <= mn) {
                    mn = k[u];
                    ms = (unsigned long long)(i + u * SEL_THREADS);
                }
                mx = k[u] > mx ? k[u] : mx;
            }
        }
        for (; i < hi; i += SEL_THREADS) {
            const unsigned long long k = dbits(pool.errmax[i]);
            if (k <= mn) {
                mn = k;
                ms = (unsigned long long)i;
            }
            mx = k > mx ? k : mx;
        }
    } else {
        // flush the level-0 histogram (per-warp copies) to the global one
        unsigned c = 0;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < SEL_WARPS; ++q) {
            c += scnt[q * 256 + tid];
            s += ssum[q * 256 + tid];
        }
        if (c) {
            atomicAdd(&set_cnt(p.hset, 0)[tid], (unsigned long long)c);
            atomicAdd(&set_sum(p.hset, 0)[tid], s);
        }
    }
    {
        shk[tid] = mn;
        shs[tid] = ms;
        shm[tid] = mx;
        __syncthreads();
        for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
            if (tid < q) {
                const unsigned long long k2 = shk[tid + q], s2 = shs[tid + q];
                if (k2 < shk[tid] || (k2 == shk[tid] && s2 > shs[tid])) {
                    shk[tid] = k2;
                    shs[tid] = s2;
                }
                if (shm[tid + q] > shm[tid]) shm[tid] = shm[tid + q];
            }
            __syncthreads();
        }
        if (tid == 0) {
            p.blk[b] = shk[0];
            p.blk[G + b] = shs[0];
            p.blk[2 * G + b] = shm[0];
        }
    }
    grid.sync();
    // ---- B: totals and decisions, redundantly in every CTA --------------------------------------------
    double* totals = (double*)(p.ctl + 16);  // CTA 0 publishes; every CTA keeps V, E for fdim == 1 in registers
    double* minvec = totals + 2 * fdim;
    double V = 0.0, E = 0.0;
    if (warp_rows) {  // vector integrands: only CTA 0 needs the totals (it alone takes the decisions below)
        if (b == 0) {
            for (int row = (int)w; row < 2 * fdim; row += SEL_WARPS) {
                double s = 0.0;
                for (int i = (int)lane; i < G; i += 32) s += __ldcg(&p.partial[(long long)row * G + i]);
                for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
                if (lane == 0) totals[row] = s;
            }
        }
        __syncthreads();
    }
    for (int row = 0; row < (warp_rows ? 0 : 2 * fdim); ++row) {
        double s = 0.0;
        for (int i = tid; i < G; i += SEL_THREADS) s += __ldcg(&p.partial[(long long)row * G + i]);
        sh[tid] = s;
        __syncthreads();
        for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
            if (tid < q) sh[tid] += sh[tid + q];
            __syncthreads();
        }
        if (row == 0) V = sh[0];
        if (row == fdim) E = sh[0];
        if (b == 0 && tid == 0) totals[row] = sh[0];
        __syncthreads();
    }
    {
        mn = ~0ULL, ms = 0ULL, mx = 0ULL;
        for (int i = tid; i < G; i += SEL_THREADS) {
            const unsigned long long k = __ldcg(&p.blk[i]), sl = __ldcg(&p.blk[G + i]), m2 = __ldcg(&p.blk[2 * G + i]);
            if (k < mn || (k == mn && sl > ms)) {
                mn = k;
                ms = sl;
            }
            mx = m2 > mx ? m2 : mx;
        }
        shk[tid] = mn;
        shs[tid] = ms;
        shm[tid] = mx;
        __syncthreads();
        for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
            if (tid < q) {
                const unsigned long long k2 = shk[tid + q], s2 = shs[tid + q];
                if (k2 < shk[tid] || (k2 == shk[tid] && s2 > shs[tid])) {
                    shk[tid] = k2;
                    shs[tid] = s2;
                }
                if (shm[tid + q] > shm[tid]) shm[tid] = shm[tid + q];
            }
            __syncthreads();
        }
    }
    const long long argmin = (long long)shs[0];
    long long n_glob = n;  // regions of the whole partition
    if (fdim == 1) {
        double e_last = 0.0, v_last = 0.0;
        unsigned long long kmin = shk[0], kmax = shm[0];
        if (sharded) {
            // CTA 0 exchanges {V, E, smallest-errmax region, N} and the level-0 histogram with the other ranks, adds the
            // blocks in rank order (identical bits on every rank) and publishes the result to the other CTAs
            if (b == 0) {
                if (tid == 0) {
                    s_hdr[0] = dbits(V);
                    s_hdr[1] = dbits(E);
                    s_hdr[2] = n > 0 ? shk[0] : ~0ULL;
                    s_hdr[3] = shm[0];
                    s_hdr[4] = dbits(n > 0 ? pool.val[argmin] : 0.0);
                    s_hdr[5] = dbits(n > 0 ? pool.err[argmin] : 0.0);
                    s_hdr[6] = (unsigned long long)n;
                    for (int q = 7; q < 16; ++q) s_hdr[q] = 0ULL;
                }
                __syncthreads();
            }
            const unsigned long long seq = xseq++;
            xchg_slice(p, seq, s_hdr, set_cnt(p.hset, 0), set_sum(p.hset, 0), scalar_select ? 256 : 0);
            if (b == 0) {
                if (tid == 0) {
                    unsigned long long* own = p.xbuf[p.rank];
                    double gV = 0.0, gE = 0.0, gvl = 0.0, gel = 0.0;
                    unsigned long long gmin = ~0ULL, gmax = 0ULL, gN = 0ULL;
                    bool any = false;
                    for (int r = 0; r < p.R; ++r) {
                        const unsigned long long* pay = xb_payload(own, seq, r);
                        const unsigned long long nr = __ldcg(&pay[6]);
                        if (nr == 0ULL) continue;  // a rank without regions contributes nothing
                        gV += __longlong_as_double((long long)__ldcg(&pay[0]));
                        gE += __longlong_as_double((long long)__ldcg(&pay[1]));
                        gN += nr;
                        const unsigned long long km = __ldcg(&pay[2]), kx = __ldcg(&pay[3]);
                        if (!any || km <= gmin) {  // ties: the highest rank pops last
                            gmin = km;
                            gvl = __longlong_as_double((long long)__ldcg(&pay[4]));
                            gel = __longlong_as_double((long long)__ldcg(&pay[5]));
                            any = true;
                        }
                        gmax = kx > gmax ? kx : gmax;
                    }
                    p.gblock[0] = dbits(gV);
                    p.gblock[1] = dbits(gE);
                    p.gblock[2] = gmin;
                    p.gblock[3] = gmax;
                    p.gblock[4] = dbits(gvl);
                    p.gblock[5] = dbits(gel);
                    p.gblock[6] = gN;
                    totals[0] = gV;
                    totals[1] = gE;
                }
                __threadfence();
            }
            grid.sync();
            if (__ldcg(&p.gblock[7])) {  // a peer did not answer (timeout)
                if (b == 0 && tid == 0) {
                    p.ctl[11] = 1;
                    p.ctl[13] = (long long)(xseq - p.seq0);
                }
                return;
            }
            V = __longlong_as_double((long long)__ldcg(&p.gblock[0]));
            E = __longlong_as_double((long long)__ldcg(&p.gblock[1]));
            kmin = __ldcg(&p.gblock[2]);
            kmax = __ldcg(&p.gblock[3]);
            v_last = __longlong_as_double((long long)__ldcg(&p.gblock[4]));
            e_last = __longlong_as_double((long long)__ldcg(&p.gblock[5]));
            n_glob = (long long)__ldcg(&p.gblock[6]);
        }
        // scalar: every CTA decides on its own registers (no dependence on CTA 0's stores)
        if (tid == 0) {
            if (!sharded) {
                e_last = pool.err[argmin];
                v_last = pool.val[argmin];
            }
            long long amb = 0;
            const bool c0 = cub_converged(1, ScalarAcc{V, E}, p.absTol, p.relTol, p.norm);
            const bool c1 = cub_converged(1, ScalarAcc{V, E * (1.0 + p.band)}, p.absTol, p.relTol, p.norm);
            const bool c2 = cub_converged(1, ScalarAcc{V, E * (1.0 - p.band)}, p.absTol, p.relTol, p.norm);
            if (c0 != c1 || c0 != c2) amb += 1;
            bool all = false;
            if (!c0) {
                all = !cub_converged(1, ScalarAcc{V, e_last}, p.absTol, p.relTol, p.norm);
                const bool a1 = !cub_converged(1, ScalarAcc{V, e_last + p.band * E}, p.absTol, p.relTol, p.norm);
                const bool a2 = !cub_converged(1, ScalarAcc{V, e_last - p.band * E}, p.absTol, p.relTol, p.norm);
                if (a1 != all || a2 != all) amb += 1;
                if (n_glob == 1) all = true;
                all = all && p.Kcap >= n_glob;
            }
            s_flags[0] = c0 ? 1 : 0;
            s_flags[1] = all ? 1 : 0;
            if (b == 0) {
                minvec[0] = v_last;
                minvec[1] = e_last;
                ((unsigned long long*)p.ctl)[8] = kmin;
                ((unsigned long long*)p.ctl)[9] = kmax;
                p.ctl[10] = argmin;
                p.ctl[0] = all ? 2 * n : 0;
                p.ctl[1] = 0;
                p.ctl[2] = n;
                p.ctl[3] = all ? n : 0;
                p.ctl[4] = c0 ? 1 : 0;
                p.ctl[5] = all ? 1 : 0;
                p.ctl[6] = amb;
                p.ctl[7] = all ? n_glob : 0;  // pops of the whole partition
                p.ctl[12] = n_glob;
                p.ctl[13] = (long long)(xseq - p.seq0);  // exchanges so far (final unless the select runs)
            }
        }
        __syncthreads();
    } else {
        // vector integrands: CTA 0 alone publishes the decisions; the host continues (sort-based selection)
        if (b == 0) {
            for (int j = tid; j < fdim; j += SEL_THREADS) {
                minvec[j] = pool.val[(long long)j * pool.cap + argmin];
                minvec[fdim + j] = pool.err[(long long)j * pool.cap + argmin];
            }
            __syncthreads();
            if (tid == 0) {
                const double* tv = totals;
                const double* te = totals + fdim;
                long long amb = 0;
                const bool c0 = cub_converged(fdim, BandAcc{tv, te, te, 1.0, 0.0}, p.absTol, p.relTol, p.norm);
                const bool c1 = cub_converged(fdim, BandAcc{tv, te, te, 1.0 + p.band, 0.0}, p.absTol, p.relTol, p.norm);
                const bool c2 = cub_converged(fdim, BandAcc{tv, te, te, 1.0 - p.band, 0.0}, p.absTol, p.relTol, p.norm);
                if (c0 != c1 || c0 != c2) amb += 1;
                bool all = false;
                if (!c0) {
                    const double* le = minvec + fdim;
                    all = !cub_converged(fdim, BandAcc{tv, le, te, 1.0, 0.0}, p.absTol, p.relTol, p.norm);
                    const bool a1 = !cub_converged(fdim, BandAcc{tv, le, te, 1.0, p.band}, p.absTol, p.relTol, p.norm);
                    const bool a2 = !cub_converged(fdim, BandAcc{tv, le, te, 1.0, -p.band}, p.absTol, p.relTol, p.norm);
                    if (a1 != all || a2 != all) amb += 1;
                    if (n == 1) all = true;
                    all = all && p.Kcap >= n;
                }
                ((unsigned long long*)p.ctl)[8] = shk[0];
                ((unsigned long long*)p.ctl)[9] = shm[0];
                p.ctl[10] = argmin;
                p.ctl[0] = all ? 2 * n : 0;
                p.ctl[1] = 0;
                p.ctl[2] = n;
                p.ctl[3] = all ? n : 0;
                p.ctl[4] = c0 ? 1 : 0;
                p.ctl[5] = all ? 1 : 0;
                p.ctl[6] = amb;
            }
        }
        return;
    }
    if (s_flags[0] | s_flags[1]) return;  // converged, or the whole pool is popped (uniform over the grid)
    if (!p.do_select) return;
    // ---- C: weighted radix select, coarse to fine -----------------------------------------------------
    if (tid == 0) {
        lst.prefix = 0ULL;
        lst.mask = 0ULL;
        lst.count_cum = 0ULL;
        lst.K = 0ULL;
        lst.tie_take = 0ULL;
        lst.ambiguous = 0ULL;
        lst.P_cum = 0.0;
    }
    __syncthreads();
    const double* err = pool.err;
    bool have_cand = false;
    unsigned long long less_def = 0;  // (thread-local) regions of this chunk that are certainly popped
    long long M = 0;                  // candidates
    int fin = 0;
    for (int level = 0; level < SEL_LEVELS; ++level) {
        const int shift = sel_shift(level), NB = sel_nbuckets(level);
        if (level >= 1) {
            for (int q = tid; q < NB; q += SEL_THREADS) {
                scnt[q] = 0u;
                ssum[q] = 0.0;
            }
            __syncthreads();
            const unsigned long long prefix = lst.prefix, mask = lst.mask;
            unsigned cur = 0, run_c = 0;
            double run_s = 0.0;
            if (level <= 2) {
                // full scan of the chunk; level 2 also extracts the candidates and counts the certain pops
                unsigned long long* const cand_count = set_misc(p.hset);
                // (loop bounds are warp-uniform: the level-2 body votes with the full mask)
                for (long long w0 = lo + (long long)(tid - (int)lane); w0 < hi; w0 += 4 * SEL_THREADS) {
                    const long long i0 = w0 + lane;
                    double a[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) a[u] = i0 + u * SEL_THREADS < hi ? err[i0 + u * SEL_THREADS] : 0.0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const long long i = i0 + u * SEL_THREADS;
                        const bool in = i < hi;
                        const unsigned long long k = sel_key_of_err(a[u]);
                        const bool match = in && (k & mask) == prefix;
                        if (level == 2) {
                            less_def += (in && (k & mask) < prefix) ? 1ULL : 0ULL;
                            // warp-aggregated append
                            const unsigned bal = __ballot_sync(0xffffffffu, match);
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
                                sel_flush(scnt, ssum, cur, run_c, run_s);
                                cur = d;
                                run_c = 0;
                                run_s = 0.0;
                            }
                            ++run_c;
                            run_s += a[u];
                        }
                    }
                }
                if (level == 2) have_cand = true;
            } else {
                // candidates only
                for (long long q = (long long)b * SEL_THREADS + tid; q < M; q += (long long)G * SEL_THREADS) {
                    const unsigned long long k = p.cand_key[q];
                    if ((k & mask) == prefix) {
                        const unsigned d = (unsigned)(k >> shift) & (unsigned)(NB - 1);
                        if (d != cur) {
                            sel_flush(scnt, ssum, cur, run_c, run_s);
                            cur = d;
                            run_c = 0;
                            run_s = 0.0;
                        }
                        ++run_c;
                        run_s += __longlong_as_double((long long)~k);  // err == errmax for a candidate (errmax > 0 or both 0)
                    }
                }
            }
            sel_flush(scnt, ssum, cur, run_c, run_s);
            __syncthreads();
            if (G > 1 || sharded) {
                unsigned long long* gc = set_cnt(p.hset, level);
                double* gs = set_sum(p.hset, level);
                for (int q = tid; q < NB; q += SEL_THREADS) {
                    const unsigned c = scnt[q];
                    if (c) {
                        atomicAdd(&gc[q], (unsigned long long)c);
                        atomicAdd(&gs[q], ssum[q]);
                    }
                }
                if (level == 2) {  // certain pops of this chunk -> per-CTA counters of the compaction
                    shk[tid] = less_def;
                    __syncthreads();
                    for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
                        if (tid < q) shk[tid] += shk[tid + q];
                        __syncthreads();
                    }
                    if (tid == 0) {
                        p.blk[3 * G + b] = shk[0];
                        p.blk[4 * G + b] = 0ULL;
                    }
                }
                __threadfence();
                grid.sync();
                if (sharded) {  // this level's histogram of the whole partition: rank-ordered sum of the R local ones
                    if (b == 0) {
                        if (tid < 16) s_hdr[tid] = 0ULL;
                        __syncthreads();
                    }
                    xchg_slice(p, xseq++, s_hdr, gc, gs, NB);
                    __threadfence();
                    grid.sync();
                    if (__ldcg(&p.gblock[7])) {
                        if (b == 0 && tid == 0) {
                            p.ctl[11] = 1;
                            p.ctl[13] = (long long)(xseq - p.seq0);
                        }
                        return;
                    }
                }
            }
        }
        // stage this level's global histogram and pick (every CTA, same inputs, same result)
        if (G > 1 || sharded || level == 0) {
            const unsigned long long* gc = set_cnt(p.hset, level);
            const double* gs = set_sum(p.hset, level);
            __syncthreads();
            for (int q = tid; q < NB; q += SEL_THREADS) {
                pc[q] = __ldcg(&gc[q]);
                ps[q] = __ldcg(&gs[q]);
            }
        } else {
            // one CTA holds the whole pool: its own histogram is the global one (widen the counts in place, top down)
            unsigned cc[SEL_LB / SEL_THREADS];
#pragma unroll
            for (int j = 0; j < SEL_LB / SEL_THREADS; ++j) cc[j] = tid + j * SEL_THREADS < NB ? scnt[tid + j * SEL_THREADS] : 0u;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < SEL_LB / SEL_THREADS; ++j)
                if (tid + j * SEL_THREADS < NB) pc[tid + j * SEL_THREADS] = cc[j];
            if (level == 2) {
                shk[tid] = less_def;
                __syncthreads();
                for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
                    if (tid < q) shk[tid] += shk[tid + q];
                    __syncthreads();
                }
                if (tid == 0) {
                    p.blk[3 * G + b] = shk[0];
                    p.blk[4 * G + b] = 0ULL;
                }
            }
        }
        __syncthreads();
        if (level == 2) M = (long long)__ldcg(set_misc(p.hset));
        fin = sel_pick_parallel(&lst, pc, ps, NB, shift, level == SEL_LEVELS - 1, V, E, (unsigned long long)p.Kcap, p.absTol, p.relTol, p.norm,
                                p.band, sh, shk, shs, s_dbl, s_ll);
        if (fin) {
            if (level < SEL_LEVELS - 1) {
                // the chosen bucket holds ONE region: the batch size is final, only the low bits of the threshold
                // key are missing.  Whoever sees that region publishes its key (one more barrier instead of the
                // remaining levels).
                const unsigned long long prefix2 = lst.prefix, mask2 = lst.mask;
                if (have_cand) {
                    for (long long q = (long long)b * SEL_THREADS + tid; q < M; q += (long long)G * SEL_THREADS) {
                        const unsigned long long k = p.cand_key[q];
                        if ((k & mask2) == prefix2) p.blk[5 * G] = k;
                    }
                } else {
                    for (long long i = lo + tid; i < hi; i += SEL_THREADS) {
                        const unsigned long long k = sel_key_of_err(err[i]);
                        if ((k & mask2) == prefix2) p.blk[5 * G] = k;
                    }
                }
                if (G > 1 || sharded) {
                    __threadfence();
                    grid.sync();
                } else {
                    __syncthreads();
                }
                if (sharded) {  // exactly one rank owns that region
                    const unsigned long long seq = xseq++;
                    if (b == 0) {
                        if (tid < 16) s_hdr[tid] = tid == 0 ? __ldcg(&p.blk[5 * G]) : 0ULL;
                        __syncthreads();
                        xchg_slice(p, seq, s_hdr, nullptr, nullptr, 0);
                        if (tid == 0) {
                            unsigned long long key = 0ULL;
                            for (int r = 0; r < p.R; ++r) {
                                const unsigned long long k = __ldcg(&xb_payload(p.xbuf[p.rank], seq, r)[0]);
                                if (k) key = k;
                            }
                            p.blk[5 * G] = key;
                        }
                        __threadfence();
                    }
                    grid.sync();
                    if (__ldcg(&p.gblock[7])) {
                        if (b == 0 && tid == 0) {
                            p.ctl[11] = 1;
                            p.ctl[13] = (long long)(xseq - p.seq0);
                        }
                        return;
                    }
                }
                if (tid == 0) {
                    lst.prefix = __ldcg(&p.blk[5 * G]);
                    lst.mask = ~0ULL;
                }
                __syncthreads();
            }
            break;
        }
    }
    // ---- D: stable compaction of the popped slots -----------------------------------------------------
    // Every CTA has computed the selection state redundantly; CTA 0's copy is the one that counts (it publishes K):
    // broadcast it so that the compaction below uses one threshold everywhere, and count disagreeing CTAs.
    const volatile SelState* const vlst = &lst;  // (same object)
    if (G > 1) {
        if (b == 0 && tid == 0) {
            p.blk[5 * G + 1] = vlst->prefix;
            p.blk[5 * G + 2] = vlst->tie_take;
            p.blk[5 * G + 3] = vlst->K;
        }
        __threadfence();
        grid.sync();
        if (tid == 0) {
            const unsigned long long g_prefix = __ldcg(&p.blk[5 * G + 1]), g_tie = __ldcg(&p.blk[5 * G + 2]), g_K = __ldcg(&p.blk[5 * G + 3]);
            if (g_prefix != vlst->prefix || g_tie != vlst->tie_take || g_K != vlst->K) atomicAdd((unsigned long long*)&p.ctl[14], 1ULL);
            lst.prefix = g_prefix;
            lst.tie_take = g_tie;
            lst.K = g_K;
        }
        __syncthreads();
    }
    const unsigned long long thr = vlst->prefix, tie_take = vlst->tie_take;
    if (have_cand) {
        // per-CTA counts: certain pops (level-2 scan) + the candidates below / at the threshold key
        for (long long q = (long long)b * SEL_THREADS + tid; q < M; q += (long long)G * SEL_THREADS) {
            const unsigned long long k = p.cand_key[q];
            if (k <= thr) {
                const long long owner = (long long)p.cand_idx[q] / per;
                atomicAdd(&p.blk[(k < thr ? 3 : 4) * (long long)G + owner], 1ULL);
            }
        }
    } else {
        unsigned long long my_less = 0, my_eq = 0;
        for (long long i = lo + tid; i < hi; i += SEL_THREADS) {
            const unsigned long long k = sel_key_of_err(err[i]);
            my_less += k < thr;
            my_eq += k == thr;
        }
        shk[tid] = my_less;
        shs[tid] = my_eq;
        __syncthreads();
        for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
            if (tid < q) {
                shk[tid] += shk[tid + q];
                shs[tid] += shs[tid + q];
            }
            __syncthreads();
        }
        if (tid == 0) {
            p.blk[3 * G + b] = shk[0];
            p.blk[4 * G + b] = shs[0];
        }
    }
    if (G > 1 || sharded) {
        __threadfence();
        grid.sync();
    } else {
        __syncthreads();
    }
    unsigned long long tie_take_local = tie_take, K_local = vlst->K;
    if (sharded) {
        const unsigned long long seq = xseq++;
        // regions equal to the threshold key are taken in (rank, slot) order: every rank publishes how many of its
        // regions are below / at the threshold and takes its share of the global tie budget
        if (b == 0) {
            unsigned long long a = 0, c = 0;
            for (int i = tid; i < G; i += SEL_THREADS) {
                a += __ldcg(&p.blk[3 * G + i]);
                c += __ldcg(&p.blk[4 * G + i]);
            }
            shk[tid] = a;
            shs[tid] = c;
            __syncthreads();
            for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
                if (tid < q) {
                    shk[tid] += shk[tid + q];
                    shs[tid] += shs[tid + q];
                }
                __syncthreads();
            }
            if (tid < 16) s_hdr[tid] = tid == 0 ? shk[0] : (tid == 1 ? shs[0] : 0ULL);
            __syncthreads();
            xchg_slice(p, seq, s_hdr, nullptr, nullptr, 0);
            if (tid == 0) {
                unsigned long long eq_before = 0ULL, my_less = 0ULL, my_eq = 0ULL;
                for (int r = 0; r < p.R; ++r) {
                    const unsigned long long* pay = xb_payload(p.xbuf[p.rank], seq, r);
                    if (r < p.rank) eq_before += __ldcg(&pay[1]);
                    if (r == p.rank) {
                        my_less = __ldcg(&pay[0]);
                        my_eq = __ldcg(&pay[1]);
                    }
                }
                unsigned long long share = tie_take > eq_before ? tie_take - eq_before : 0ULL;
                if (share > my_eq) share = my_eq;
                p.gblock[8] = share;
                p.gblock[9] = my_less + share;

            }
            __threadfence();
        }
        grid.sync();
        if (__ldcg(&p.gblock[7])) {
            if (b == 0 && tid == 0) {
                p.ctl[11] = 1;
                p.ctl[13] = (long long)(xseq - p.seq0);
            }
            return;
        }
        tie_take_local = __ldcg(&p.gblock[8]);
        K_local = __ldcg(&p.gblock[9]);
    }
    {
        unsigned long long a = 0, c = 0;
        for (int i = tid; i < b; i += SEL_THREADS) {
            a += __ldcg(&p.blk[3 * G + i]);
            c += __ldcg(&p.blk[4 * G + i]);
        }
        shk[tid] = a;
        shs[tid] = c;
        __syncthreads();
        for (int q = SEL_THREADS / 2; q > 0; q >>= 1) {
            if (tid < q) {
                shk[tid] += shk[tid + q];
                shs[tid] += shs[tid + q];
            }
            __syncthreads();
        }
    }
    unsigned long long less_base = shk[0], eq_base = shs[0];
    __syncthreads();
    unsigned* const sl = (unsigned*)pc;
    unsigned* const se = sl + SEL_THREADS;
    for (long long t_lo = lo; t_lo < hi; t_lo += SEL_TILE) {
        const long long i_lo = t_lo + (long long)tid * SEL_ITEMS;  // SEL_ITEMS consecutive slots per thread
        unsigned fl = 0, fe = 0, less = 0, eq = 0;
#pragma unroll
        for (int u = 0; u < SEL_ITEMS; ++u) {
            const long long i = i_lo + u;
            if (i < hi) {
                const unsigned long long k = sel_key_of_err(err[i]);
                if (k < thr) {
                    fl |= 1u << u;
                    ++less;
                } else if (k == thr) {
                    fe |= 1u << u;
                    ++eq;
                }
            }
        }
        sl[tid] = less;
        se[tid] = eq;
        __syncthreads();
        for (int o = 1; o < SEL_THREADS; o <<= 1) {
            const unsigned ta = tid >= o ? sl[tid - o] : 0u, tb = tid >= o ? se[tid - o] : 0u;
            __syncthreads();
            sl[tid] += ta;
            se[tid] += tb;
            __syncthreads();
        }
        unsigned long long less_before = less_base + (sl[tid] - less);
        unsigned long long eq_before = eq_base + (se[tid] - eq);
        const unsigned tile_less = sl[SEL_THREADS - 1], tile_eq = se[SEL_THREADS - 1];
#pragma unroll
        for (int u = 0; u < SEL_ITEMS; ++u) {
            const bool is_less = (fl >> u) & 1u, is_eq = (fe >> u) & 1u;
            if (is_less || (is_eq && eq_before < tie_take_local)) {
                const unsigned long long pos = less_before + (eq_before < tie_take_local ? eq_before : tie_take_local);
                p.list[pos] = (int)(i_lo + u);
            }
            less_before += is_less;
            eq_before += is_eq;
        }
        less_base += tile_less;
        eq_base += tile_eq;
        __syncthreads();
    }
    if (b == 0 && tid == 0) {
        p.st->prefix = lst.prefix;
        p.st->mask = lst.mask;
        p.st->count_cum = lst.count_cum;
        p.st->K = lst.K;
        p.st->tie_take = lst.tie_take;
        p.st->ambiguous = lst.ambiguous;
        p.st->P_cum = lst.P_cum;
        p.ctl[0] = 2 * (long long)K_local;
        p.ctl[1] = 1;
        p.ctl[2] = n;
        p.ctl[3] = (long long)K_local;
        p.ctl[6] += (long long)lst.ambiguous;
        p.ctl[7] = (long long)lst.K;  // pops of the whole partition
        p.ctl[13] = (long long)(xseq - p.seq0);

    }
}

// ---- FP64 peak microbenchmark: independent DFMA chains, no memory traffic --------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // keeps the chains alive without a store in the common case
}

}  // namespace

// ---- host launchers --------------------------------------------------------------------------------
static inline unsigned nblocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

void pk_bisect(const PoolView& pool, int dim, const int* list, long long K, long long base, cudaStream_t s) {
    if (K <= 0) return;
    k_bisect<<<nblocks(K, 128), 128, 0, s>>>(pool, dim, list, K, base);
}
void pk_errmax_rows(const PoolView& pool, int fdim, const int* list, long long n_items, long long base, int mode, cudaStream_t s) {
    if (n_items <= 0) return;
    k_errmax_rows<<<nblocks(n_items, 128), 128, 0, s>>>(pool, fdim, list, n_items, base, mode);
}
int pk_totals_blocks(long long n) {
    long long b = (n + 2047) / 2048;
    if (b < 1) b = 1;
    if (b > 1184) b = 1184;  // 8 x 148 SMs
    return (int)b;
}
void pk_totals(const PoolView& pool, int fdim, long long n, double* partial, double* totals, cudaStream_t s) {
    const int nb = pk_totals_blocks(n);
    dim3 grid(nb, 2 * fdim);
    k_totals_partial<<<grid, TOT_THREADS, 0, s>>>(pool, fdim, n, partial);
    k_totals_final<<<2 * fdim, TOT_THREADS, 0, s>>>(partial, nb, totals);
}
void pk_min_region(const PoolView& pool, int fdim, long long n, unsigned long long* scratch /*[3]*/, double* out_vec /*[2f]*/,
                   cudaStream_t s) {
    const unsigned long long init[3] = {~0ULL, 0ULL, 0ULL};
    cudaMemcpyAsync(scratch, init, sizeof init, cudaMemcpyHostToDevice, s);
    long long b = (n + 255) / 256;
    if (b > 1184) b = 1184;
    k_minmax_key<<<(unsigned)b, 256, 0, s>>>(pool.errmax, n, scratch);
    k_argmin_slot<<<(unsigned)b, 256, 0, s>>>(pool.errmax, n, scratch, scratch + 2);
    k_gather_slot<<<nblocks(fdim, 128), 128, 0, s>>>(pool, fdim, scratch + 2, out_vec);
}

long long pk_sort_chunks(long long n) { return (n + SORT_CHUNK - 1) / SORT_CHUNK; }
long long pk_scan_tiles(long long n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

void pk_sort_prepare(const double* errmax, long long n, unsigned long long* keys, int* idx, unsigned long long* or_and,
                     cudaStream_t s) {
    const unsigned long long init[2] = {0ULL, ~0ULL};
    cudaMemcpyAsync(or_and, init, sizeof init, cudaMemcpyHostToDevice, s);
    long long b = (n + 255) / 256;
    if (b > 2368) b = 2368;
    k_sort_prepare<<<(unsigned)b, 256, 0, s>>>(errmax, n, keys, idx, or_and);
}
void pk_sort_pass(const unsigned long long* keys_in, const int* idx_in, unsigned long long* keys_out, int* idx_out, long long n,
                  int shift, unsigned* hist, unsigned long long* bin_totals, unsigned long long* binbase, cudaStream_t s) {
    const long long nchunks = pk_sort_chunks(n);
    const unsigned blocks = (unsigned)((nchunks + SORT_WARPS - 1) / SORT_WARPS);
    k_sort_hist<<<blocks, SORT_WARPS * 32, 0, s>>>(keys_in, n, shift, nchunks, hist);
    k_sort_scan_rows<<<256, 256, 0, s>>>(hist, nchunks, bin_totals);
    k_sort_scan_bins<<<1, 32, 0, s>>>(bin_totals, binbase);
    k_sort_scatter<<<blocks, SORT_WARPS * 32, 0, s>>>(keys_in, idx_in, n, shift, nchunks, hist, binbase, keys_out, idx_out);
}
void pk_prefix(const PoolView& pool, int fdim, const int* perm, long long n, double* prefix, double* tile_tot, cudaStream_t s) {
    const long long ntiles = pk_scan_tiles(n);
    dim3 grid((unsigned)ntiles, fdim);
    k_scan_tiles<<<grid, SCAN_THREADS, 0, s>>>(pool, perm, n, prefix, ntiles, tile_tot);
    k_scan_tile_offsets<<<fdim, 256, 0, s>>>(tile_tot, ntiles);
}
void pk_find_k(const double* totals, const double* prefix, const double* tile_off, long long n, int fdim, double absTol,
               double relTol, int norm, double band, unsigned long long* out /*[3]*/, cudaStream_t s) {
    const unsigned long long init[3] = {(unsigned long long)(n + 1), (unsigned long long)(n + 1), (unsigned long long)(n + 1)};
    cudaMemcpyAsync(out, init, sizeof init, cudaMemcpyHostToDevice, s);
    k_find_k<<<nblocks(n, 128), 128, 0, s>>>(totals, prefix, tile_off, n, pk_scan_tiles(n), fdim, absTol, relTol, norm, band, out);
}
long long pk_sel_tiles(long long n) { return (n + SEL_TILE - 1) / SEL_TILE; }
size_t pk_sel_state_bytes() { return sizeof(SelState); }

// Weighted radix select + compaction (fdim == 1), in pieces so that the multi-GPU driver can put a
// collective between them.  scratch: see SelScratch (pool_kernels.h); cnt and sum are contiguous
// (u64[2048] then double[2048]) so that one all-gather ships both.
// SelState: K at byte 24, tie_take at 32, ambiguous at 40.
void pk_sel_init(const SelScratch& sc, cudaStream_t s) {
    cudaMemsetAsync(sc.state, 0, 512 + PK_SEL_HIST_BYTES, s);  // SelState, ticket, cnt[256], sum[256]
}
int pk_sel_passes() { return 8; }
// pass p decides bits [63 - 8p, 56 - 8p]; fuse_pick = 1: histogram + pick in one launch (single GPU)
void pk_sel_pass(const PoolView& pool, long long n, const SelScratch& sc, int pass, int fuse_pick, const double* totals, long long Kcap,
                 double absTol, double relTol, int norm, double band, const long long* ctl, cudaStream_t s) {
    long long b = (n + (long long)SEL_THREADS * SEL_UNROLL - 1) / ((long long)SEL_THREADS * SEL_UNROLL);
    if (b > 148 * 8) b = 148 * 8;
    if (b < 1) b = 1;
    k_sel_pass<<<(unsigned)b, SEL_THREADS, 0, s>>>(pool.errmax, pool.err, n, (SelState*)sc.state, 56 - 8 * pass, pass == 7, sc.cnt, sc.sum,
                                                  sc.ticket, fuse_pick, totals, (unsigned long long)Kcap, absTol, relTol, norm, band, ctl);
}
void pk_sel_pick(const SelScratch& sc, int pass, const double* totals, long long Kcap, double absTol, double relTol, int norm,
                 double band, cudaStream_t s) {
    k_sel_pick<<<1, SEL_THREADS, 0, s>>>((SelState*)sc.state, sc.cnt, sc.sum, 56 - 8 * pass, pass == 7, totals, (unsigned long long)Kcap,
                                        absTol, relTol, norm, band);
}
void pk_sel_tile_counts(const PoolView& pool, long long n, const void* state, unsigned long long* tile_less,
                        unsigned long long* tile_eq, unsigned long long* totals_out, long long* ctl, cudaStream_t s) {
    const long long ntiles = pk_sel_tiles(n);
    if (ntiles > 0) k_sel_tile_counts<<<(unsigned)ntiles, SEL_THREADS, 0, s>>>(pool.errmax, n, (const SelState*)state, tile_less, tile_eq, ctl);
    k_sel_tile_scan<<<1, 256, 0, s>>>(tile_less, tile_eq, ntiles, totals_out, (const SelState*)state, ctl, n);
}
void pk_sel_compact(const PoolView& pool, long long n, const void* state, const unsigned long long* tile_less,
                    const unsigned long long* tile_eq, long long tie_take, int* list, const long long* ctl, cudaStream_t s) {
    const long long ntiles = pk_sel_tiles(n);
    if (ntiles > 0)
        k_sel_compact<<<(unsigned)ntiles, SEL_THREADS, 0, s>>>(pool.errmax, n, (const SelState*)state, tile_less, tile_eq, tie_take, list, ctl);
}
// One launch: totals, extreme keys, convergence test, select-all shortcut -> device control block.
// scratch: partial [2f][G] doubles, blk [3][G] u64, ticket; the selection scratch header is cleared.
void pk_pool_stats(const PoolView& pool, int fdim, long long n, double* partial, unsigned long long* blk, unsigned* ticket, long long* ctl,
                   long long Kcap, double absTol, double relTol, int norm, double band, unsigned long long* sel_zero, int sel_zero_words,
                   cudaStream_t s) {
    const int G = pk_totals_blocks(n);
    k_pool_stats<<<G, TOT_THREADS, 0, s>>>(pool, fdim, n, partial, blk, ticket, ctl, Kcap, absTol, relTol, norm, band, sel_zero, sel_zero_words);
}
int pk_select_scalar(const PoolView& pool, long long n, const double* totals, long long Kcap, double absTol, double relTol,
                     int norm, double band, const SelScratch& sc, int* list, cudaStream_t s) {
    // (the selection scratch header was cleared by k_pool_stats when ctl != nullptr)
    if (!sc.ctl) pk_sel_init(sc, s);
    for (int pass = 0; pass < 8; ++pass) pk_sel_pass(pool, n, sc, pass, 1, totals, Kcap, absTol, relTol, norm, band, sc.ctl, s);
    pk_sel_tile_counts(pool, n, sc.state, sc.tile_less, sc.tile_eq, nullptr, sc.ctl, s);
    pk_sel_compact(pool, n, sc.state, sc.tile_less, sc.tile_eq, -1, list, sc.ctl, s);
    return 8 + 3;
}

// adds the per-rank histogram blocks [R][cnt u64 x256 | sum f64 x256] in rank order
__global__ void k_sum_hist_ranks(const unsigned char* gathered, int R, unsigned long long* cnt, double* sum) {
    const int b = threadIdx.x;  // 256 threads
    unsigned long long c = 0;
    double s = 0.0;
    for (int r = 0; r < R; ++r) {
        const unsigned long long* pc = (const unsigned long long*)(gathered + (size_t)r * PK_SEL_HIST_BYTES);
        const double* ps = (const double*)(gathered + (size_t)r * PK_SEL_HIST_BYTES + 2048);
        c += pc[b];
        s += ps[b];
    }
    cnt[b] = c;
    sum[b] = s;
}
void pk_sum_hist_ranks(const unsigned char* gathered, int R, unsigned long long* cnt, double* sum, cudaStream_t s) {
    k_sum_hist_ranks<<<1, 256, 0, s>>>(gathered, R, cnt, sum);
}

// ---- sharded vector select: this rank's part of the first K entries of the global error ranking ------------------
// sorted[i] are global region numbers (rank-major: rank r owns [lo_r, hi_r)) in descending-errmax order; the ranks'
// local lists are the stable sub-sequences of sorted[0..K) that fall into their range.  Three small launches:
// per-tile counts, a one-block scan of the tile counts, the stable write.
constexpr int FLT_THREADS = 256;
constexpr int FLT_ITEMS = 8;
constexpr int FLT_TILE = FLT_THREADS * FLT_ITEMS;
__global__ void __launch_bounds__(FLT_THREADS) k_filter_count(const int* sorted, long long K, long long lo, long long hi, unsigned long long* tile_cnt) {
    const long long base = (long long)blockIdx.x * FLT_TILE + (long long)threadIdx.x * FLT_ITEMS;
    unsigned c = 0;
    for (int u = 0; u < FLT_ITEMS; ++u) {
        const long long i = base + u;
        if (i < K) {
            const long long g = sorted[i];
            c += (g >= lo && g < hi) ? 1u : 0u;
        }
    }
    __shared__ unsigned sh[FLT_THREADS];
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int w = FLT_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = sh[0];
}
__global__ void k_filter_scan(unsigned long long* tile_cnt, long long ntiles, unsigned long long* total_out) {  // exclusive, in place
    __shared__ unsigned long long sh[256];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < ntiles; base += 256) {
        const long long i = base + threadIdx.x;
        const unsigned long long a = i < ntiles ? tile_cnt[i] : 0ULL;
        sh[threadIdx.x] = a;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const unsigned long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0ULL;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < ntiles) tile_cnt[i] = carry + sh[threadIdx.x] - a;
        __syncthreads();
        if (threadIdx.x == 255) carry += sh[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(FLT_THREADS) k_filter_write(const int* sorted, long long K, long long lo, long long hi,
                                                              const unsigned long long* tile_off, int* list) {
    const long long base = (long long)blockIdx.x * FLT_TILE + (long long)threadIdx.x * FLT_ITEMS;
    unsigned flags = 0, c = 0;
    for (int u = 0; u < FLT_ITEMS; ++u) {
        const long long i = base + u;
        if (i < K) {
            const long long g = sorted[i];
            if (g >= lo && g < hi) {
                flags |= 1u << u;
                ++c;
            }
        }
    }
    __shared__ unsigned sh[FLT_THREADS];
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int o = 1; o < FLT_THREADS; o <<= 1) {
        const unsigned t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    unsigned long long pos = tile_off[blockIdx.x] + (sh[threadIdx.x] - c);
    for (int u = 0; u < FLT_ITEMS; ++u)
        if ((flags >> u) & 1u) list[pos++] = (int)((long long)sorted[base + u] - lo);
}

// deals the replicated pool out to the ranks: with `perm` = the regions in descending-errmax order,
// rank r keeps perm[r], perm[r + R], ... (round robin over the error ranking, so every rank gets the
// same share of the regions that will be refined most); without perm, slots r, r + R, ...
__global__ void k_shard_gather(PoolView src, PoolView dst, int dim, int fdim, long long n_local, int R, int rank, const int* perm) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_local) return;
    const long long s = perm ? (long long)perm[j * R + rank] : j * R + rank;
    for (int d = 0; d < dim; ++d) {
        dst.center[(long long)d * dst.cap + j] = src.center[(long long)d * src.cap + s];
        dst.halfw[(long long)d * dst.cap + j] = src.halfw[(long long)d * src.cap + s];
    }
    for (int k = 0; k < fdim; ++k) {
        dst.val[(long long)k * dst.cap + j] = src.val[(long long)k * src.cap + s];
        dst.err[(long long)k * dst.cap + j] = src.err[(long long)k * src.cap + s];
    }
    dst.errmax[j] = src.errmax[s];
    dst.splitdim[j] = src.splitdim[s];
}
void pk_shard_gather(const PoolView& src, const PoolView& dst, int dim, int fdim, long long n_local, int R, int rank, const int* perm,
                     cudaStream_t s) {
    if (n_local > 0) k_shard_gather<<<nblocks(n_local, 128), 128, 0, s>>>(src, dst, dim, fdim, n_local, R, rank, perm);
}

long long pk_filter_tiles(long long K) { return (K + FLT_TILE - 1) / FLT_TILE; }
// list[0..*count) = local slots (g - lo) of the entries of sorted[0..K) inside [lo, hi), in their sorted order
void pk_filter_range(const int* sorted, long long K, long long lo, long long hi, unsigned long long* tile_scratch /*[pk_filter_tiles(K)]*/,
                     unsigned long long* count_out, int* list, cudaStream_t s) {
    const long long nt = pk_filter_tiles(K);
    if (nt <= 0) {
        cudaMemsetAsync(count_out, 0, 8, s);
        return;
    }
    k_filter_count<<<(unsigned)nt, FLT_THREADS, 0, s>>>(sorted, K, lo, hi, tile_scratch);
    k_filter_scan<<<1, 256, 0, s>>>(tile_scratch, nt, count_out);
    k_filter_write<<<(unsigned)nt, FLT_THREADS, 0, s>>>(sorted, K, lo, hi, tile_scratch, list);
}

// Cooperative launch of one whole iteration (everything but the rule kernel).  hsets: two histogram sets of
// PK_COOP_SET_BYTES each (both zero before the first launch of a loop; launch k uses set k & 1 and clears the other
// one for launch k + 1).  cand_key / cand_idx: candidate scratch of the select, n entries each.
// Returns cudaSuccess or the launch error.
static int g_coop_blocks_per_sm = -1;
cudaError_t pk_iter_coop(const PoolView& pool, int fdim, long long n, long long Kcap, double absTol, double relTol, int norm, double band,
                         int do_select, double* partial, unsigned long long* blk, long long* ctl, void* state, void* hsets, int parity,
                         unsigned long long* cand_key, int* cand_idx, int* list, int n_sms, cudaStream_t s, const PkShard* shard) {
    if (g_coop_blocks_per_sm < 0) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_iter_coop, SEL_THREADS, 0);
        g_coop_blocks_per_sm = nb > 0 ? nb : 1;
    }
    const int per_sm = g_coop_blocks_per_sm < 4 ? g_coop_blocks_per_sm : 4;
    long long G = (n + (long long)SEL_THREADS * 16 - 1) / ((long long)SEL_THREADS * 16);
    if (G > (long long)n_sms * per_sm) G = (long long)n_sms * per_sm;
    if (G < 1) G = 1;
    IterParams p;
    p.pool = pool;
    p.fdim = fdim;
    p.norm = norm;
    p.do_select = do_select;
    p.n = n;
    p.Kcap = Kcap;
    p.absTol = absTol;
    p.relTol = relTol;
    p.band = band;
    p.partial = partial;
    p.blk = blk;
    p.ctl = ctl;
    p.st = (SelState*)state;
    p.hset = (unsigned long long*)((unsigned char*)hsets + (size_t)(parity & 1) * PK_COOP_SET_BYTES);
    p.hzero = (unsigned long long*)((unsigned char*)hsets + (size_t)((parity + 1) & 1) * PK_COOP_SET_BYTES);
    p.cand_key = cand_key;
    p.cand_idx = cand_idx;
    p.list = list;
    p.R = 1;
    p.rank = 0;
    for (int r = 0; r < PK_MAX_RANKS; ++r) p.xbuf[r] = nullptr;
    p.gblock = nullptr;
    p.seq0 = 0;
    p.xsliced = 0;
    if (shard && shard->R > 1) {
        p.R = shard->R;
        p.rank = shard->rank;
        for (int r = 0; r < shard->R; ++r) p.xbuf[r] = (unsigned long long*)shard->xbuf[r];
        p.gblock = shard->gblock;
        p.seq0 = shard->seq0;
        // CUBATURE_XCHG_SLICED=0: CTA 0 alone runs every exchange (slower: 22.3 vs 19.5 ms per C5 step on 8 B200)
        static const int sliced = [] { const char* e = getenv("CUBATURE_XCHG_SLICED"); return e && e[0] == '0' ? 0 : 1; }();
        p.xsliced = sliced;
        if (sliced && G < PK_XB_SLICES) G = PK_XB_SLICES;  // the exchange is sliced over the first CTAs
    }
    void* args[1] = {&p};
    return cudaLaunchCooperativeKernel((const void*)k_iter_coop, dim3((unsigned)G), dim3(SEL_THREADS), args, 0, s);
}
int pk_iter_coop_max_blocks(int n_sms) { return n_sms * 4; }

void pk_dfma_peak(double* out, int blocks, int iters, cudaStream_t s) {
    k_dfma_peak<<<blocks, 256, 0, s>>>(out, iters, 0.999999, 1e-6);
}
