// pool_kernels.h -- host launchers of the integrand-independent pool kernels (pool_kernels.cu).
#ifndef CUB_POOL_KERNELS_H
#define CUB_POOL_KERNELS_H

#include <cuda_runtime.h>

// Structure-of-arrays region pool in HBM; same bytes as CubPoolView in rule_kernels.cuh.
struct PoolView {
    double* center;  // [dim][cap]
    double* halfw;   // [dim][cap]
    double* val;     // [fdim][cap]
    double* err;     // [fdim][cap]
    double* errmax;  // [cap]
    int* splitdim;   // [cap]
    long long cap;
};

void pk_bisect(const PoolView& pool, int dim, const int* list, long long K, long long base, cudaStream_t s);
void pk_errmax_rows(const PoolView& pool, int fdim, const int* list, long long n_items, long long base, int mode, cudaStream_t s);

int pk_totals_blocks(long long n);
void pk_totals(const PoolView& pool, int fdim, long long n, double* partial, double* totals, cudaStream_t s);
void pk_min_region(const PoolView& pool, int fdim, long long n, unsigned long long* scratch, double* out_vec, cudaStream_t s);

long long pk_sort_chunks(long long n);
long long pk_scan_tiles(long long n);
void pk_sort_prepare(const double* errmax, long long n, unsigned long long* keys, int* idx, unsigned long long* or_and,
                     cudaStream_t s);
void pk_sort_pass(const unsigned long long* keys_in, const int* idx_in, unsigned long long* keys_out, int* idx_out, long long n,
                  int shift, unsigned* hist, unsigned long long* bin_totals, unsigned long long* binbase, cudaStream_t s);
void pk_prefix(const PoolView& pool, int fdim, const int* perm, long long n, double* prefix, double* tile_tot, cudaStream_t s);
void pk_find_k(const double* totals, const double* prefix, const double* tile_off, long long n, int fdim, double absTol,
               double relTol, int norm, double band, unsigned long long* out, cudaStream_t s);
long long pk_sel_tiles(long long n);
size_t pk_sel_state_bytes();
// device scratch of the scalar select
struct SelScratch {
    void* state;                    // SelState (56 B) at byte 0 of a zero-initialised 512 B header
    unsigned* ticket;               // last-CTA ticket counter (byte 128 of the header)
    unsigned long long* cnt;        // [256] at byte 512
    double* sum;                    // [256] directly after cnt (one all-gather ships both)
    unsigned long long* tile_less;  // [pk_sel_tiles(capacity)]
    unsigned long long* tile_eq;    // [pk_sel_tiles(capacity)]
    long long* ctl;                 // device control block of the fused loop, or null
};
#define PK_CTL_WORDS 16  // 8-byte header words of the control block; totals[2f] and minvec[2f] follow
#define PK_SEL_HIST_BYTES (256 * 16)
int pk_sel_passes();
void pk_sel_init(const SelScratch& sc, cudaStream_t s);
void pk_sel_pass(const PoolView& pool, long long n, const SelScratch& sc, int pass, int fuse_pick, const double* totals, long long Kcap,
                 double absTol, double relTol, int norm, double band, const long long* ctl, cudaStream_t s);
void pk_pool_stats(const PoolView& pool, int fdim, long long n, double* partial, unsigned long long* blk, unsigned* ticket, long long* ctl,
                   long long Kcap, double absTol, double relTol, int norm, double band, unsigned long long* sel_zero, int sel_zero_words,
                   cudaStream_t s);
void pk_sel_pick(const SelScratch& sc, int pass, const double* totals, long long Kcap, double absTol, double relTol, int norm,
                 double band, cudaStream_t s);
void pk_sel_tile_counts(const PoolView& pool, long long n, const void* state, unsigned long long* tile_less,
                        unsigned long long* tile_eq, unsigned long long* totals_out, long long* ctl, cudaStream_t s);
void pk_sel_compact(const PoolView& pool, long long n, const void* state, const unsigned long long* tile_less,
                    const unsigned long long* tile_eq, long long tie_take, int* list, const long long* ctl, cudaStream_t s);
void pk_sum_hist_ranks(const unsigned char* gathered, int R, unsigned long long* cnt, double* sum, cudaStream_t s);
void pk_shard_gather(const PoolView& src, const PoolView& dst, int dim, int fdim, long long n_local, int R, int rank, const int* perm,
                     cudaStream_t s);
long long pk_filter_tiles(long long K);
void pk_filter_range(const int* sorted, long long K, long long lo, long long hi, unsigned long long* tile_scratch, unsigned long long* count_out,
                     int* list, cudaStream_t s);
int pk_select_scalar(const PoolView& pool, long long n, const double* totals, long long Kcap, double absTol, double relTol,
                     int norm, double band, const SelScratch& sc, int* list, cudaStream_t s);
// peer exchange buffer of a sharded pool (one per rank, mapped into every peer; 8-byte words):
//   [PK_XB_FLAGS + (parity * PK_MAX_RANKS + src) * PK_XB_SLICES + slice]  flag: sequence number of src's latest
//                                                exchange of that parity, one flag per 256-bucket slice of the payload
//   [PK_XB_DATA + (parity * PK_MAX_RANKS + src) * PK_XB_PAY ..]  payload of src: header[16] | cnt[n] | sum[n]
#define PK_MAX_RANKS 16
#define PK_XB_FLAGS 0
#define PK_XB_SLICES 8
#define PK_XB_DATA 512
#define PK_XB_ABORT 300  // != 0: a rank gave up (iter_exact.cu: ix_abort_peers); spinning exchanges fail at once
#define PK_XB_PAY (5 * 2048 + 64)  // header[16] + the largest payload: all error / value bins of a rank (iter_exact.cu)
#define PK_XB_WORDS (PK_XB_DATA + 2 * PK_MAX_RANKS * PK_XB_PAY)
#define PK_XB_BYTES (PK_XB_WORDS * 8)
#define PK_XCHG_TIMEOUT_NS 20000000000ULL
// sharded pool: this rank's view of the peer exchange buffers (nullptr / R == 1: single GPU)
struct PkShard {
    int R, rank;
    void* xbuf[PK_MAX_RANKS];
    unsigned long long* gblock;  // [16] words of local device scratch
    unsigned long long seq0;     // sequence number of the launch's first exchange (cub_comm::seq)
};
cudaError_t pk_iter_coop(const PoolView& pool, int fdim, long long n, long long Kcap, double absTol, double relTol, int norm, double band,
                         int do_select, double* partial, unsigned long long* blk, long long* ctl, void* state, void* hsets, int parity,
                         unsigned long long* cand_key, int* cand_idx, int* list, int n_sms, cudaStream_t s, const PkShard* shard = nullptr);
int pk_iter_coop_max_blocks(int n_sms);
// one histogram set of the cooperative iteration kernel: cnt[7][2048] u64 | sum[7][2048] f64 | misc[8]
#define PK_COOP_SET_WORDS (2 * 7 * 2048 + 8)
#define PK_COOP_SET_BYTES (PK_COOP_SET_WORDS * 8)
#define PK_COOP_HIST_BYTES (2 * PK_COOP_SET_BYTES)
void pk_dfma_peak(double* out, int blocks, int iters, cudaStream_t s);

#endif
