/* cubature_cuda.h -- C ABI of libcubature_cuda.so, the B200 (sm_100a) drop-in for the h-adaptive
 * integration loop of jonathanschilling/Cubature.
 *
 * This is the boundary a host-language binding targets (Java Panama FFM / JNI, Python ctypes, C):
 * plain pointers, sizes and scalars only, caller-owned buffers, no exceptions or callbacks across
 * the ABI, negative return value = error (cubature_cuda_strerror).  INTEGRATION.md shows the
 * reference-side binding a maintainer would add.
 *
 * Reference interfaces replaced (paths under src/main/java/de/labathome/cubature/ of the reference):
 *   cubature_cuda_integrate        <- Cubature.integrate(...)              Cubature.java:108-170
 *                                     (+ Rule.cubature, Rule.java:51-160; evalError of
 *                                     Rule75GenzMalik.java:61-184 and RuleGaussKronrod_1d.java:56-164;
 *                                     RegionHeap.java:5-37; Region.cut, Region.java:31-38)
 *   cubature_cuda_integrand_*      <- the UnaryOperator<double[][]> integrand plugin, Cubature.java:98-99
 *                                     (the integrand moves onto the device: built-in family or CUDA C
 *                                     source compiled for sm_100a with NVRTC; fdim is declared instead
 *                                     of probed at the box centre, Cubature.java:132-145)
 *   norm argument                  <- CubatureError ordinals, CubatureError.java:9-29
 *   cubature_cuda_eval_regions     <- Rule.evalError(integrand, Region[], nR), Rule.java:36
 * Return layout of `out`: [2][fdim] row-major, values then error estimates (Cubature.java:169).
 */
#ifndef CUBATURE_CUDA_H
#define CUBATURE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUBATURE_CUDA_ABI_VERSION 1

/* ---- status codes ---------------------------------------------------------------------------- */
#define CUB_OK 0
#define CUB_ERR_BAD_ARGS (-1)       /* null/empty xmin, xmax, out ...          Cubature.java:119-128 */
#define CUB_ERR_BAD_TOL (-2)        /* relTol and absTol both NaN              Rule.java:164-167     */
#define CUB_ERR_BAD_DIM (-3)        /* dim < 1, dim >= 32 (Rule75GenzMalik.java:30-42) or > kernel limit */
#define CUB_ERR_BAD_NORM (-4)       /* norm outside 0..4                       Rule.java:273-276     */
#define CUB_ERR_NO_DEVICE (-5)      /* no usable CUDA device / driver                                 */
#define CUB_ERR_CUDA (-6)           /* a CUDA runtime call failed (message in cub_stats_t)            */
#define CUB_ERR_NVRTC (-7)          /* integrand source failed to compile (log returned)              */
#define CUB_ERR_POOL_EXHAUSTED (-8) /* region pool capacity reached before convergence / maxEval      */
#define CUB_ERR_NCCL (-9)           /* NCCL unavailable or a collective failed                        */
#define CUB_ERR_UNSUPPORTED (-10)   /* valid request this build cannot serve (e.g. fdim too large)    */
#define CUB_ERR_BAD_FAMILY (-11)    /* unknown built-in integrand family / wrong parameter count      */

/* ---- error norms: ordinals of the reference's enum (CubatureError.java:9-29) ------------------- */
#define CUB_NORM_INDIVIDUAL 0
#define CUB_NORM_PAIRED 1
#define CUB_NORM_L2 2
#define CUB_NORM_L1 3
#define CUB_NORM_LINF 4
#define CUB_NORM_STRICT 8   /* flag for cubature_cuda_converged: INDIVIDUAL|STRICT, PAIRED|STRICT (see cub_options_t.strict_all) */

/* ---- built-in integrand families (cubature_b200/csrc/integrands.h) ---------------------------- */
#define CUB_FAMILY_GENZ_OSCILLATORY 1   /* params: a[dim], u[dim]               */
#define CUB_FAMILY_GENZ_PRODUCT_PEAK 2
#define CUB_FAMILY_GENZ_CORNER_PEAK 3
#define CUB_FAMILY_GENZ_GAUSSIAN 4
#define CUB_FAMILY_GENZ_C0 5
#define CUB_FAMILY_GENZ_DISCONTINUOUS 6
#define CUB_FAMILY_GAUSS_ISO 10         /* params: sigma            (TestCubature.java:435-465)   */
#define CUB_FAMILY_REFTEST 20           /* params: which[fdim] in 0..8 (TestCubature.java:109-177) */
#define CUB_FAMILY_OSC_VEC 30           /* f_k = cos((k+1) sum x), no params                      */
#define CUB_FAMILY_CEXP 31              /* params: w[fdim/2]; exp(-(1+i w_m) sum x), (re,im) pairs */

/* ---- per-coordinate variable transforms (README.md:237-272), fused into point generation ------ */
#define CUB_TRANSFORM_NONE 0
#define CUB_TRANSFORM_SEMI_INF_UP 1     /* [xmin, +inf):  x = xmin + t/(1-t), t in [0,1)          */
#define CUB_TRANSFORM_SEMI_INF_DOWN 2   /* (-inf, xmax]:  x = xmax - t/(1-t), t in [0,1)          */
#define CUB_TRANSFORM_INF 3             /* (-inf, +inf):  x = t/(1-t^2),      t in (-1,1)         */

/* ---- region-selection policy ------------------------------------------------------------------ */
#define CUB_SELECT_DEVICE 0      /* everything on the GPU: convergence, top-K selection, bisection.
                                    Ties in errmax are broken by lower slot index first.            */
#define CUB_SELECT_HEAP_EXACT 1  /* the host keeps a java.util.PriorityQueue-compatible heap of
                                    (errmax, slot) and the reference's running sums, so batch sizes,
                                    tie order and the final summation order are the reference's,
                                    bit for bit; the GPU does rule evaluation and bisection.        */

typedef struct cub_integrand cub_integrand_t; /* opaque: compiled device plugin + parameters        */
typedef struct cub_comm cub_comm_t;           /* opaque: one rank of a multi-GPU communicator       */

/* Optional trace: caller-owned buffers, filled up to their capacities; n_* report the true counts. */
typedef struct cub_trace {
    int64_t batch_cap;   /* capacity of batches[]                                                   */
    int64_t* batches;    /* nR (= 2K) of every refinement iteration (Rule.java:135)                 */
    int64_t n_batches;   /* out                                                                     */
    int64_t region_cap;  /* capacity (in regions) of the five arrays below                          */
    double* centers;     /* [n][dim]  final regions (this rank's shard), heap-array order in         */
    double* halfwidths;  /* [n][dim]  HEAP_EXACT mode, slot order in DEVICE mode                     */
    int32_t* split_dim;  /* [n]                                                                     */
    double* val;         /* [n][fdim]                                                               */
    double* err;         /* [n][fdim]                                                               */
    int64_t n_regions;   /* out                                                                     */
} cub_trace_t;

typedef struct cub_options {
    int32_t struct_size;    /* = sizeof(cub_options_t)                                              */
    int32_t select;         /* CUB_SELECT_*                                                         */
    int32_t device;         /* CUDA device ordinal; -1 = current device                             */
    int32_t strict_all;     /* 1: INDIVIDUAL / PAIRED converge only when EVERY component (pair) meets a tolerance
                               (what README.md:112-122 documents); 0: the reference's ANY semantics (Rule.java:172-213) */
    int64_t pool_capacity;  /* regions; 0 = derive from maxEval and free memory                     */
    cub_trace_t* trace;     /* NULL = no trace                                                      */
    cub_comm_t* comm;       /* NULL = single GPU; else shard the pool over comm's ranks (DEVICE mode)*/
    int64_t shard_min_regions; /* replicate the loop on every rank until the pool holds this many
                                  regions, then shard (0 = default: 32768 * nranks when maxEval allows more than
                                  2^24 regions, else 2^23 -- small jobs stay replicated)               */
    const char* checkpoint_save; /* NULL, or a path: when the call returns (converged, or stopped by maxEval) the region
                                    pool and the loop counters are written to this file (DEVICE mode, one GPU)          */
    const char* checkpoint_load; /* NULL, or a checkpoint written by an earlier call with the same integrand, box and
                                    transform: the loop continues from that pool instead of the root box (Rule.java:62-69);
                                    relTol / absTol / norm / maxEval are this call's (maxEval counts the evaluations of
                                    the checkpointed run too)                                                           */
} cub_options_t;

/* What a pool checkpoint holds (host-only readers: no GPU needed). */
typedef struct cub_checkpoint_info {
    int32_t dim, fdim, family, reserved;
    int64_t n_regions, num_eval, iterations;
    int32_t transform[12];
    double shift[12];
} cub_checkpoint_info_t;

typedef struct cub_stats {
    int32_t struct_size;     /* = sizeof(cub_stats_t)                                               */
    int32_t converged;       /* 1 = tolerance met; 0 = stopped by maxEval (not an error: the reference
                                prints a message and returns the best estimate, Rule.java:144-159)   */
    int64_t num_eval;        /* the reference's numEval accounting (Rule.java:69,127), all ranks     */
    int64_t num_regions;     /* regions in the final partition, all ranks                            */
    int64_t iterations;      /* refinement iterations (evalError calls after the first)              */
    int64_t regions_evaluated; /* = num_eval / points-per-region                                     */
    int64_t ambiguous_decisions; /* DEVICE mode: iterations whose batch size was decided within the
                                    rounding band of the reference's running sums    */
    int64_t gpu_launches;    /* kernels launched by this call                                        */
    int64_t h2d_bytes;       /* bytes copied host->device by this call (geometry, lists, scalars)    */
    int64_t d2h_bytes;       /* bytes copied device->host by this call (control scalars, results)    */
    double device_ms;        /* CUDA-event time of the whole loop on the library's stream            */
    double rule_ms;          /* CUDA-event time summed over the rule-kernel launches only            */
    double wall_ms;          /* host wall clock of the call                                          */
    double jit_ms;           /* NVRTC compile + module load done inside this call (0 when cached)    */
    char message[256];       /* human-readable detail for errors                                     */
    int64_t rebalances;      /* sharded pool: how often surplus regions were moved between the ranks' shards */
} cub_stats_t;

/* Library / ABI version: major*1000 + minor. */
int cubature_cuda_version(void);
const char* cubature_cuda_strerror(int status);

/* Number of rule points per region: 15 for dim == 1, else 2^dim + 2 dim^2 + 2 dim + 1
 * (Cubature.java:153-165).  Returns a negative status for an unsupported dim. */
int64_t cubature_cuda_points_per_region(int dim);

/* The convergence predicate alone, Rule.converged (Rule.java:162-279), on host vectors val[fdim],
 * err[fdim]: returns 1 (converged), 0, or CUB_ERR_BAD_TOL / CUB_ERR_BAD_NORM.  Host-only utility. */
int cubature_cuda_converged(int fdim, const double* val, const double* err, double absTol, double relTol, int norm);

/* The rule points of one box in the reference's order (Rule75GenzMalik.java:237-326,
 * RuleGaussKronrod_1d.java:60-84): points[P][dim] row-major, P = cubature_cuda_points_per_region(dim).  Host-only utility
 * for plotting evaluation positions (the reference's examples/PlotEvalPositions.java) from a cub_trace_t region dump. */
int cubature_cuda_rule_points(int dim, const double* center, const double* halfwidth, double* points);

/* Pool checkpoints (cub_options_t.checkpoint_save / checkpoint_load): header of a file, and regions [first, first + count)
 * of it in the layout of cub_trace_t (centers, halfwidths [count][dim]; val, err [count][fdim]; split_dim [count]; any of
 * the five may be NULL).  checkpoint_regions returns the number of regions copied or a negative status.  Together with
 * cubature_cuda_rule_points this reproduces examples/PlotEvalPositions.java:28-54 (tools/eval_positions.py). */
int cubature_cuda_checkpoint_info(const char* path, cub_checkpoint_info_t* info);
int64_t cubature_cuda_checkpoint_regions(const char* path, int64_t first, int64_t count, double* centers, double* halfwidths,
                                         double* val, double* err, int32_t* split_dim);

/* Built-in integrand family -> device plugin. */
int cubature_cuda_integrand_builtin(int family_id, int dim, int fdim, const double* params, size_t nparams,
                                    cub_integrand_t** out);

/* CUDA C source -> device plugin, compiled with NVRTC for sm_100a (no FMA contraction).
 * The source must define
 *     __device__ void <entry>(const double* x, double* f, const double* p);
 * reading x[CUB_DIM] and writing f[CUB_FDIM]; CUB_DIM / CUB_FDIM are predefined macros and the
 * deterministic math of detmath.h (dm_exp, dm_sin, dm_cos, dm_log, dm_pow, dm_powi, dm_fma, dm_sqrt,
 * dm_abs) is pre-included.  On failure the NVRTC log is copied to `log` (NUL-terminated). */
int cubature_cuda_integrand_nvrtc(const char* cuda_src, const char* entry, int dim, int fdim, const double* params,
                                  size_t nparams, cub_integrand_t** out, char* log, size_t log_cap);

void cubature_cuda_integrand_free(cub_integrand_t* integrand);

/* Copies the sm_100a cubin of the integrand's rule kernels (compiling it if necessary; needs no GPU)
 * into buf; returns its size in bytes or a negative status.  For cuobjdump / ncu source views. */
int64_t cubature_cuda_integrand_cubin(cub_integrand_t* integrand, int has_transform, void* buf, size_t buf_cap);

/* The h-adaptive integration, Cubature.integrate(integrand, xmin, xmax, relTol, absTol, norm, maxEval).
 *   transform: NULL (infinite bounds are then mapped automatically) or dim codes CUB_TRANSFORM_*
 *   maxEval:   0 = unlimited (Rule.java:72); int64 because BASELINE config 5 needs 1e11
 *   out:       [2][fdim] values then errors
 *   options, stats: may be NULL */
int cubature_cuda_integrate(cub_integrand_t* integrand, int dim, const double* xmin, const double* xmax,
                            const int* transform, double relTol, double absTol, int norm, int64_t maxEval,
                            const cub_options_t* options, double* out, cub_stats_t* stats);

/* Rule evaluation alone (Rule.evalError) on nR caller-supplied boxes, host buffers:
 *   centers, halfwidths [nR][dim] (in t-space when a transform is given; shift[dim] = finite bound)
 *   val, err [nR][fdim]; split_dim [nR].  stats may be NULL. */
int cubature_cuda_eval_regions(cub_integrand_t* integrand, int dim, const int* transform, const double* shift,
                               int64_t nR, const double* centers, const double* halfwidths, double* val, double* err,
                               int32_t* split_dim, cub_stats_t* stats);

/* Integrand values alone at n points x[dim][n] -> f[fdim][n] (host buffers; bit-parity tests). */
int cubature_cuda_eval_points(cub_integrand_t* integrand, int dim, int64_t n, const double* x, double* f);

/* Throughput benchmark of the rule kernel alone: evaluates n_regions boxes resident in HBM
 * (a uniform grid partition of [xmin,xmax]) `repeats` times; returns the CUDA-event time per launch. */
int cubature_cuda_bench_rule(cub_integrand_t* integrand, int dim, const double* xmin, const double* xmax,
                             int64_t n_regions, int repeats, double* ms_per_launch, double* checksum);

/* Device and pinned buffers are cached per device between calls; this frees the idle ones. */
void cubature_cuda_release_workspace(void);

/* Measured FP64 FMA peak of the current device (dependent-chain DFMA microbenchmark), TFLOP/s. */
int cubature_cuda_fp64_peak(int device, double* tflops, double* ms);
/* The same probe launched back to back for window_ms: burst = best single launch, sustained = rate over the second half of
 * the window (the SM clock drops under a sustained FP64 load once the board reaches its power cap). */
int cubature_cuda_fp64_peak_sustained(int device, double window_ms, double* tflops_burst, double* tflops_sustained);

/* Multi-GPU: one rank per process (or per thread).  Rank 0 creates the id and ships the 128 bytes to
 * the other ranks by any means (torch.distributed, MPI, a file); every rank then calls comm_init.
 * comm_init also maps one exchange buffer per rank into every peer (CUDA IPC over NVLink): with it the sharded
 * refinement loop of a scalar integrand runs without any NCCL call (the iteration kernel exchanges through peer
 * memory).  CUBATURE_PEER_EXCHANGE=0 / CUBATURE_PEER_LOOP=0 fall back to NCCL all-gathers driven by the host. */
#define CUB_COMM_ID_BYTES 128
int cubature_cuda_comm_unique_id(void* id_out);
int cubature_cuda_comm_init(const void* id, int nranks, int rank, int device, cub_comm_t** out);
/* Host-side sharding arithmetic of the multi-GPU path (pure functions; exported so that the N > 1
 * logic can be tested without GPUs):
 *   shard_count: how many of n replicated regions rank keeps (slots rank, rank+nranks, ...)
 *   tie_share:   how many of its regions tied at the selection threshold rank pops, given the global
 *                budget tie_take and every rank's tie count (ties are popped in (rank, slot) order)   */
int64_t cubature_cuda_shard_count(int64_t n, int nranks, int rank);
int64_t cubature_cuda_tie_share(int64_t tie_take, const int64_t* ties_per_rank, int nranks, int rank);

/* Ranks as threads of ONE process (a single-process host driving several GPUs, or several shards on one
 * GPU for testing): every rank calls this with the same group_key; the exchange goes through host memory. */
int cubature_cuda_comm_init_local(uint64_t group_key, int nranks, int rank, int device, cub_comm_t** out);
void cubature_cuda_comm_free(cub_comm_t* comm);

#ifdef __cplusplus
}
#endif
#endif /* CUBATURE_CUDA_H */
