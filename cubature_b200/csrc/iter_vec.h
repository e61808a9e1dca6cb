// iter_vec.h -- the device-resident refinement loop for VECTOR integrands (fdim > 1; iter_vec.cu).  One iteration of
// Rule.cubature (Rule.java:72-142) is ONE cooperative launch (totals, Rule.converged, the Gladwell batch cut with the
// per-component residual of Rule.java:122-124, compaction and bisection) plus the rule launch it sizes through the control
// block, so the host enqueues several iterations per visit exactly as for scalar integrands (iter_exact.h: same control
// block, same states).
#ifndef CUB_ITER_VEC_H
#define CUB_ITER_VEC_H

#include <cuda_runtime.h>

#include "iter_exact.h"
#include "pool_kernels.h"

#define IV_THREADS 256
#define IV_NB 16          // buckets per narrowing level (4 bits of the composite key)
#define IV_CMAX 1024      // candidates the final stage sorts in one CTA
#define IV_CHUNKS 32      // k-chunks of the candidates' running sums
#define IV_SS_FMAX 256     // up to this fdim the bucket suffix sums of a level are staged in shared memory
#define IV_GC_ROWS 16384  // narrowing: at most this many (CTA, component) pairs of per-bucket partial sums per bucket

// selection state shared by the CTAs through global memory (written by CTA 0, read by all after a grid barrier)
struct IvState {
    unsigned long long thr_hi;  // threshold composite key: errmax bits ...
    unsigned long long thr_lo;  //   ... and 0xffffffff - slot (regions with a composite key >= threshold are popped)
    unsigned long long K;       // batch size
    unsigned long long M;       // candidates gathered
    unsigned long long fail;
    unsigned long long kmin[3];  // smallest candidate index at which the predicate holds (exact, + band, - band)
    unsigned long long pad[3];   // [0] the decision lies within the rounding band
};

struct IterVParams {
    PoolView pool;
    int dim, fdim;
    int need_errmax;          // 1: the rule kernel leaves errmax / splitdim to this kernel (separable GK15 path)
    int gc_max;               // CTAs that take part in the narrowing passes
    long long* ctl;           // [LC_WORDS + LC_LOG_CAP] (iter_exact.h)
    double* partial;          // [2 f][maxG] per-CTA partial totals
    double* totals;           // [4 f]: Sum val, Sum err, then val / err of the smallest-errmax region
    unsigned long long* blk;  // [4 maxG + 8] per-CTA min key / its slot / max key / popped count
    unsigned long long* bar;  // [0] arrivals at the grid barrier, [1] exit tickets (the last CTA out clears both)
    double* bsum_part;        // [gc][IV_NB][f] per-CTA bucket sums of one narrowing level
    unsigned long long* bcnt_part;  // [gc][IV_NB]
    double* bsum;             // [IV_NB][f] bucket sums of the partition
    unsigned long long* bcnt; // [IV_NB]
    double* sabove;           // [2][f] err sums of the regions above the current bucket (double-buffered by level)
    unsigned long long* cand_hi;  // [IV_CMAX] candidates (composite keys), unsorted
    unsigned* cand_lo;        // [IV_CMAX]
    unsigned long long* sort_hi;  // [IV_CMAX] the candidates in descending order
    unsigned* sort_lo;        // [IV_CMAX]
    double* cres;             // [IV_CMAX][f] chunk-local running sums of err over the sorted candidates
    double* cbase;            // [IV_CHUNKS][f] residual before the first candidate of a chunk
    double* wscratch;         // [maxG][8 warps][f] one residual vector per warp (norms with ordered sums)
    IvState* state;
    int* list;                // [cap] popped slots, ascending
    double absTol, relTol;
    int norm;
};

cudaError_t iv_iter(const IterVParams& p, int n_blocks, cudaStream_t s);
int iv_max_blocks(int n_sms, int fdim);
size_t iv_dyn_smem(int fdim);

#endif
