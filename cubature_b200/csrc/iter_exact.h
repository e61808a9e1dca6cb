// iter_exact.h -- the device-resident refinement loop for scalar integrands (iter_exact.cu): loop-control block, selection
// state and host launchers.  One iteration of Rule.cubature (Rule.java:72-142) is three launches that need no host
// decision in between -- decide (streams over what the last iteration changed, then one CTA decides), select (cooperative
// grid; retires at once when the batch is the whole pool), bisect + rule (persistent grid, sized by the control block) --
// so the host enqueues several iterations per visit.
#ifndef CUB_ITER_EXACT_H
#define CUB_ITER_EXACT_H

#include <cuda_runtime.h>

#include "acc_layout.h"
#include "pool_kernels.h"

// ---- loop-control block: 8-byte words in device memory ---------------------------------------------------------------
enum {
    LC_DYN_N = 0,     // rule kernel: number of work items (CubWork::dyn[0]); 0 = nothing to do.  The next decide kernel adds
                      //   exactly these items to the running sums (the host seeds 1: the root region)
    LC_DYN_LIST = 1,  //   1: parents come from the popped list, 0: identity (the whole pool is popped)
    LC_DYN_BASE = 2,  //   first slot of the right children
    LC_N = 3,         // regions in this rank's pool
    LC_NGLOBAL = 4,   // regions of the partition
    LC_NUMEVAL = 5,   // Rule.java:54,69,127
    LC_ITER = 6,      // iterations done
    LC_STATE = 7,     // LS_*
    LC_ERR = 8,       // detail of LS_ERROR
    LC_CAP = 9,       // pool capacity (host)
    LC_MAXEVAL = 10,  // host
    LC_P = 11,        // points per region (host)
    LC_SELECT = 12,   // 1: the select kernel of this iteration has work
    LC_AMBIG = 13,    // decisions within the rounding band of the reference's drifting sums (diagnostic)
    LC_SEQ = 14,      // sequence number of the next peer exchange (sharded pool)
    LC_K = 15,        // regions popped by this rank / by the partition in the last iteration
    LC_KGLOBAL = 16,
    LC_NBATCH = 17,   // entries written to the batch log
    LC_STOP_AT = 18,  // pause (LS_HANDOFF) when the partition holds this many regions (0: never)
    LC_V = 19,        // totals (double bits) at the last decision
    LC_E = 20,
    LC_MAXITER = 21,  // iteration guard (host)
    LC_NSELECT = 22,  // iterations that had to cut a batch
    LC_EVAL_LOCAL = 23,  // regions evaluated by this rank
    LC_DBG0 = 24,
    LC_DBG1 = 25,
    LC_SUB_K = 26,  // popped parents whose (val, err) copies the next decide kernel has to subtract from the sums
    LC_REBAL_MIN = 27,    // sharded pool: pause for a rebalance (LS_REBALANCE) when the partition holds at least this many regions
                          //   and the fullest shard exceeds 1.25 x the mean (0: never; host)
    LC_REBAL_COUNT = 28,  // rebalances so far (host)
    LC_READD = 29,        // regions the last selection KEPT, listed in IterXParams::kept: the select kernel has reset this rank's sums
                          //   and the next decide kernel adds these regions again instead of subtracting the popped ones
    LC_READD_MAX = 30,    // a selection takes that path when it keeps at most LC_READD_MAX / 16 of the shard (0: never; host)
    LC_STAMP = 32,  // [16] time stamps (CUBATURE_TIMING): [0..4] decide kernel, [5..15] select kernel (CTA 0)
    LC_RANKN = 48,  // [PK_MAX_RANKS] regions of every rank's shard at the last decision (sharded pool)
    LC_WORDS = 64,
};
#define LC_LOG_CAP 1024  // batch-size log [LC_WORDS .. LC_WORDS + LC_LOG_CAP): 2 K of iteration i at i % LC_LOG_CAP
enum { LS_RUN = 0, LS_CONVERGED = 1, LS_MAXEVAL = 2, LS_GROW = 3, LS_ERROR = 4, LS_GUARD = 5, LS_HANDOFF = 6, LS_REBALANCE = 7 };
enum { LE_NONE = 0, LE_PEER_TIMEOUT = 1, LE_PEER_ABORT = 2, LE_INTERNAL = 3 };

// selection state handed from the decide kernel to the select kernel (device memory)
struct SelX {
    int binade;         // binary exponent (biased) of the threshold region
    int pos;            // bit position of that binade's mantissa lsb (unit 2^-1074)
    int count_only;     // the error sum is not finite: batches are cut by the evaluation budget alone
    int pad;
    unsigned long long H_lo, H_hi;  // residual before the first pop of this binade, >> pos
    unsigned long long L1;          // its 64 bits below pos
    double Ldbl;                    // its part below pos as a (truncated) double
    unsigned long long Cabove;      // regions in higher binades: all popped
    unsigned long long Kcap;        // pops the evaluation budget allows (Rule.java:133)
    double V, E, band;
    unsigned long long bin_cnt;     // regions of the partition in this binade
    // results of the select kernel
    unsigned long long thr_key, tie_take, K;
};

// histogram set of the select kernel: [IX_LEVELS][3][IX_NB] words (count, low / high part of the mantissa remainders) + misc[8]
// levels: 8 + 8 bits (streamed), then 4 x 9 bits (candidates only) of the 52 fraction bits
#define IX_LEVELS 6
#define IX_NB 512
#define IX_GCAP 8192  // sharded pool: candidates per rank that one payload slot of the peer exchange takes (16 + IX_GCAP <= PK_XB_PAY)
#define IX_HSET_WORDS ((IX_LEVELS + 1) * 3 * IX_NB + 8)  // slot 0: binary exponents, slots 1..IX_LEVELS: fraction digits

struct IterXParams {
    PoolView pool;
    unsigned long long* acc;   // [XA_WORDS] exact running sums of this rank's pool (acc_layout.h)
    long long* ctl;            // [LC_WORDS + LC_LOG_CAP]
    SelX* selx;
    unsigned long long* hset;  // [IX_HSET_WORDS], zero between launches
    unsigned long long* blk;   // [3 * maxG + 8] per-CTA counters of the compaction
    unsigned long long* bar;   // [0] arrival counter of the select kernel's grid barrier (cleared by the decide kernel),
                               // [1] arrival counter of the decide kernel's CTAs
    unsigned long long* gbins; // [XA_WORDS] combined bins of the partition (sharded pool)
    unsigned long long* gcand; // [PK_MAX_RANKS * IX_GCAP + 16 + IX_GCAP] sharded pool: the partition's candidates, staging (or null)
    unsigned long long* cand_key;  // [cap]
    int* cand_idx;                 // [cap]
    int* list;                     // [cap] popped slots
    int* kept;                     // [cap] slots the last selection kept (LC_READD; shares cand_idx's storage)
    double* stash_val;             // [cap] val / err of the popped regions in list order (stash_err shares cand_key's storage:
    double* stash_err;             //   the candidates are dead when the compaction writes it)
    double absTol, relTol;
    int norm;
    int R, rank;
    unsigned long long* xbuf[PK_MAX_RANKS];
};

cudaError_t ix_decide(const IterXParams& p, int n_blocks, int n_sms, cudaStream_t s);
int ix_decide_max_blocks(int n_sms);
cudaError_t ix_select(const IterXParams& p, int n_blocks, cudaStream_t s);
int ix_select_max_blocks(int n_sms);
// acc := exact sums of the first n regions of the pool (after dealing a pool out to the ranks; consistency checks)
void ix_rebuild(const PoolView& pool, long long n, unsigned long long* acc, int n_sms, cudaStream_t s);
// raises the abort word in every rank's exchange buffer: peers spinning in an exchange fail at once
void ix_abort_peers(int R, void* const* xbuf, cudaStream_t s);

#endif
