// cub_internal.h -- declarations shared by the host-side translation units of libcubature_cuda.so.
#ifndef CUB_INTERNAL_H
#define CUB_INTERNAL_H

#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cubature_cuda.h"

#define CUB_MAX_DIM 12            // kernel limit (2^dim Gray-code points; the reference allows < 32)
#define CUB_MAX_PARAMS_BY_VALUE 448  // doubles passed in the kernel parameter block (4 KB limit)

// ---- device-side argument blocks: must match rule_kernels.cuh byte for byte -----------------------
struct HostPoolView {
    double* center;
    double* halfw;
    double* val;
    double* err;
    double* errmax;
    int* splitdim;
    long long cap;
};
struct HostWork {
    const int* list;
    const long long* dyn;  // device-resident {n_items, use_list, base} or null
    long long n_items;
    long long base;
    int mode;
    double* stage_val;
    double* stage_err;
    long long stage_stride;
    long long dyn_lo, dyn_hi;  // device-sized launches: batch sizes this kernel takes (rule_kernels.cuh: CubWork)
};
enum { MODE_PLAIN = 0, MODE_BISECT = 1, MODE_CHILDREN = 2 };

// ---- JIT-compiled rule kernels of one integrand variant on one device -----------------------------
struct RuleKernels {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t gm_scalar = nullptr;     // dim >= 2, fdim == 1
    cudaKernel_t gk_separable = nullptr;  // dim == 1, fdim == 1 or separable family
    cudaKernel_t staged = nullptr;        // any
    cudaKernel_t eval_points = nullptr;
    int staged_max_smem = 0;
    int gm_threads = 128;                 // block size gm_scalar was compiled for
    int gm_blocks_per_sm = 4;             // resident CTAs per SM (occupancy query): the persistent grid is n_sms times this
    int gk_blocks_per_sm = 8;
};

struct cub_integrand {
    int family = 0;
    int dim = 0, fdim = 0;
    bool separable = false;
    std::vector<double> params;
    std::string user_src, entry;
    std::mutex mu;
    std::vector<char> cubin[2];                  // [has_transform]
    std::map<int, RuleKernels> loaded[2];        // device ordinal -> kernels
    double jit_ms_pending = 0.0;                 // compile/load time not yet reported in a cub_stats_t
};

// jit.cpp
int cub_jit_compile(cub_integrand* integ, int has_transform, std::string* log);
int cub_jit_get_kernels(cub_integrand* integ, int has_transform, int device, RuleKernels** out, std::string* log);
std::string cub_jit_translation_unit(const cub_integrand* integ, int has_transform);

// nccl_dyn.cpp
struct cub_local_group;
#include "pool_kernels.h"
#define CUB_MAX_RANKS PK_MAX_RANKS
#define CUB_XBUF_BYTES PK_XB_BYTES
struct cub_comm {
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    cub_local_group* local = nullptr;  // in-process (thread-per-rank) communicator, else NCCL
    unsigned long long local_key = 0;
    // NVLink peer exchange (NCCL communicators only): every rank owns one exchange buffer in its HBM, mapped
    // into all peers with CUDA IPC; the sharded iteration kernel publishes and gathers through these
    // directly (pool_kernels.cu: k_iter_coop with R > 1), so the refinement loop itself makes no NCCL call.
    bool peer_ok = false;
    bool poisoned = false;  // a call failed inside a peer exchange: the ranks no longer agree on seq; later calls fail fast
    unsigned long long seq = 1;  // sequence number of the next exchange (every launch reports how many it ran)
    void* xbuf[CUB_MAX_RANKS] = {nullptr};  // xbuf[rank] = own buffer, others = IPC mappings
};
int cub_comm_allgather(cub_comm* comm, const void* dev_send, size_t bytes, void* dev_recv, cudaStream_t stream);

int64_t cub_points_per_region(int dim);

#define CUB_CUDA_CHECK(expr, stats_msg)                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            cub_set_message(stats_msg, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                               \
            return CUB_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

void cub_set_message(char* msg, const char* fmt, ...);

#endif
