// cubature_cuda.cpp -- host side of libcubature_cuda.so: the C ABI (include/cubature_cuda.h) and the
// two drivers of the h-adaptive loop.
//
//   DEVICE mode (default): the region pool lives in HBM as a structure of arrays and never leaves it.
//     Per refinement iteration the GPU computes the pool totals (K5), decides the batch of regions to
//     bisect (K3: select-all shortcut, else stable radix sort by errmax + prefix sums + the reference's
//     residual test evaluated for every k in parallel), and runs the fused bisect + rule kernel (K4+K1/K2).
//     The host only reads back a few scalars per iteration to size the next launches.
//   HEAP_EXACT mode: the host replays Rule.cubature literally (Rule.java:51-160) with a
//     java.util.PriorityQueue-compatible heap of (errmax, slot) and the reference's incremental
//     running sums (RegionHeap.java:21-36), so batch sizes, tie order and the final summation order
//     are the reference's bit for bit; the GPU evaluates the rule and bisects.
//
// Nothing in this file (or anywhere in cubature_b200/) touches oracle/: there is no CPU fallback.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <climits>
#include <limits>
#include <memory>
#include <mutex>

#include "converge.h"
#include "cub_internal.h"
#include "integrands.h"
#include "pool_kernels.h"
#include "iter_exact.h"
#include "iter_vec.h"
#include "checkpoint.h"

// ---- host-side sharding arithmetic of the multi-GPU path (also exported for CPU-only tests) --------
extern "C" int64_t cubature_cuda_shard_count(int64_t n, int nranks, int rank) {
    // rank r keeps slots r, r + R, r + 2R, ... of a replicated pool of n regions
    if (nranks < 1 || rank < 0 || rank >= nranks || n < 0) return CUB_ERR_BAD_ARGS;
    return n > rank ? (n - rank + nranks - 1) / nranks : 0;
}
extern "C" int64_t cubature_cuda_tie_share(int64_t tie_take, const int64_t* ties_per_rank, int nranks, int rank) {
    // regions whose errmax equals the selection threshold are popped in (rank, slot) order:
    // rank r takes what is left of the global budget after the lower ranks, at most its own tie count
    if (nranks < 1 || rank < 0 || rank >= nranks || !ties_per_rank) return CUB_ERR_BAD_ARGS;
    int64_t before = 0;
    for (int r = 0; r < rank; ++r) before += ties_per_rank[r];
    const int64_t left = tie_take - before;
    if (left <= 0) return 0;
    return left < ties_per_rank[rank] ? left : ties_per_rank[rank];
}

int64_t cub_points_per_region(int dim) {
    if (dim == 1) return 15;
    if (dim < 1 || dim > CUB_MAX_DIM) return CUB_ERR_BAD_DIM;
    return 1 + 4 * (int64_t)dim + 2 * (int64_t)dim * (dim - 1) + ((int64_t)1 << dim);
}

namespace {

struct VecAcc {  // accessor over dense val[f], err[f] for cub_converged
    const double* v;
    const double* e;
    CUB_HD double val(int j) const { return v[j]; }
    CUB_HD double err(int j) const { return e[j]; }
};

// cudaMemGetInfo is a driver-wide query: 0.1 ms on an idle process, but 8 - 30 ms per call once the process holds tens of
// GB or another thread polls NVML (measured inside the bench: 30 ms of a 112 ms step, profiles/r02/summary.md).  The
// free-memory figure only bounds the pool capacity, so it is cached per device and refreshed after this library's own
// allocations / frees (a stale figure costs at most a failed cudaMalloc, which is reported as CUB_ERR_POOL_EXHAUSTED).
struct MemInfoCache {
    std::atomic<unsigned long long> epoch{1};
    std::mutex mu;
    struct Entry { unsigned long long epoch = 0; size_t free_b = 0, total_b = 0; } dev[64];
    static MemInfoCache& get() {
        static MemInfoCache c;
        return c;
    }
    void touch() { epoch.fetch_add(1, std::memory_order_relaxed); }
    cudaError_t query(size_t* free_b, size_t* total_b) {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return cudaMemGetInfo(free_b, total_b);
        std::lock_guard<std::mutex> lk(mu);
        const unsigned long long now = epoch.load(std::memory_order_relaxed);
        if (dev[d].epoch != now) {
            cudaError_t e = cudaMemGetInfo(&dev[d].free_b, &dev[d].total_b);
            if (e != cudaSuccess) return e;
            dev[d].epoch = now;
        }
        *free_b = dev[d].free_b;
        *total_b = dev[d].total_b;
        return cudaSuccess;
    }
};

struct DevBuf {  // RAII device allocation
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            cudaFree(p);
            MemInfoCache::get().touch();
        }
        p = nullptr;
        bytes = 0;
    }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, n);
        MemInfoCache::get().touch();
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <class T>
    T* as() const { return (T*)p; }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~PinnedBuf() {
        if (p) cudaFreeHost(p);
    }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMallocHost(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <class T>
    T* as() const { return (T*)p; }
};

struct Pool {
    DevBuf center, halfw, val, err, errmax, splitdim;
    long long cap = 0;
    int dim = 0, fdim = 0;
    PoolView view() const {
        PoolView v;
        v.center = center.as<double>();
        v.halfw = halfw.as<double>();
        v.val = val.as<double>();
        v.err = err.as<double>();
        v.errmax = errmax.as<double>();
        v.splitdim = splitdim.as<int>();
        v.cap = cap;
        return v;
    }
    static size_t bytes_per_region(int dim, int fdim) { return (size_t)8 * (2 * dim + 2 * fdim + 1) + 4; }
    static constexpr size_t kPresizeBytes = (size_t)512 << 20;
    // cap_: capacity the caller would like (an upper bound derived from maxEval, or an explicit request when `exact`).
    // Unless exact, at most kPresizeBytes of pool are allocated up front -- the loop grows the pool on demand -- and
    // buffers cached from an earlier call are used as they are: a call never pays a multi-GB cudaFree / cudaMalloc
    // because the free-memory estimate behind cap_ moved a little.
    cudaError_t alloc(int dim_, int fdim_, long long cap_, bool exact = false) {
        dim = dim_;
        fdim = fdim_;
        cap = cap_;
        if (!exact) {
            size_t have = center.bytes / ((size_t)8 * dim);
            have = std::min(have, halfw.bytes / ((size_t)8 * dim));
            have = std::min(have, val.bytes / ((size_t)8 * fdim));
            have = std::min(have, err.bytes / ((size_t)8 * fdim));
            have = std::min(have, errmax.bytes / 8);
            have = std::min(have, splitdim.bytes / 4);
            const long long presize = std::max<long long>(1 << 12, (long long)(kPresizeBytes / bytes_per_region(dim, fdim)));
            cap = std::min<long long>(cap_, std::max<long long>(presize, (long long)have));
        }
        cudaError_t e;
        if ((e = center.ensure((size_t)8 * dim * cap)) != cudaSuccess) return e;
        if ((e = halfw.ensure((size_t)8 * dim * cap)) != cudaSuccess) return e;
        if ((e = val.ensure((size_t)8 * fdim * cap)) != cudaSuccess) return e;
        if ((e = err.ensure((size_t)8 * fdim * cap)) != cudaSuccess) return e;
        if ((e = errmax.ensure((size_t)8 * cap)) != cudaSuccess) return e;
        if ((e = splitdim.ensure((size_t)4 * cap)) != cudaSuccess) return e;
        // buffers cached from an earlier call may hold more: use all of it (saves the re-striding growth steps, and
        // cudaMalloc / cudaFree are slow once peer mappings exist)
        size_t fit = center.bytes / ((size_t)8 * dim);
        fit = std::min(fit, halfw.bytes / ((size_t)8 * dim));
        fit = std::min(fit, val.bytes / ((size_t)8 * fdim));
        fit = std::min(fit, err.bytes / ((size_t)8 * fdim));
        fit = std::min(fit, errmax.bytes / 8);
        fit = std::min(fit, splitdim.bytes / 4);
        if ((long long)fit > cap) cap = (long long)std::min<size_t>(fit, (size_t)std::numeric_limits<int>::max());
        return cudaSuccess;
    }
    // re-stride every row into a pool of larger capacity (n live regions)
    cudaError_t grow(long long new_cap, long long n, cudaStream_t s) {
        Pool np;
        cudaError_t e = np.alloc(dim, fdim, new_cap, true);
        if (e != cudaSuccess) return e;
        auto rows = [&](DevBuf& dst, DevBuf& src, int nrows, size_t elt) {
            return cudaMemcpy2DAsync(dst.p, (size_t)new_cap * elt, src.p, (size_t)cap * elt, (size_t)n * elt, nrows,
                                     cudaMemcpyDeviceToDevice, s);
        };
        if ((e = rows(np.center, center, dim, 8)) != cudaSuccess) return e;
        if ((e = rows(np.halfw, halfw, dim, 8)) != cudaSuccess) return e;
        if ((e = rows(np.val, val, fdim, 8)) != cudaSuccess) return e;
        if ((e = rows(np.err, err, fdim, 8)) != cudaSuccess) return e;
        if ((e = rows(np.errmax, errmax, 1, 8)) != cudaSuccess) return e;
        if ((e = rows(np.splitdim, splitdim, 1, 4)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
        std::swap(center, np.center);
        std::swap(halfw, np.halfw);
        std::swap(val, np.val);
        std::swap(err, np.err);
        std::swap(errmax, np.errmax);
        std::swap(splitdim, np.splitdim);
        cap = new_cap;
        return cudaSuccess;
    }
};

// java.util.PriorityQueue semantics over slot ids keyed by errmax[slot];
// comparator = Region.compareTo (Region.java:41-55): equal keys compare 0, larger errmax first.
struct SlotHeap {
    std::vector<int> q;
    const std::vector<double>* key = nullptr;
    int cmp(int a, int b) const {
        const double ka = (*key)[a], kb = (*key)[b];
        if (ka == kb) return 0;
        return ka < kb ? 1 : -1;
    }
    void add(int x) {
        size_t k = q.size();
        q.push_back(x);
        while (k > 0) {
            const size_t parent = (k - 1) >> 1;
            if (cmp(x, q[parent]) >= 0) break;
            q[k] = q[parent];
            k = parent;
        }
        q[k] = x;
    }
    int poll() {
        const int top = q[0];
        const int x = q.back();
        q.pop_back();
        const size_t n = q.size();
        if (n > 0) {
            size_t k = 0;
            const size_t half = n >> 1;
            while (k < half) {
                size_t child = 2 * k + 1;
                if (child + 1 < n && cmp(q[child], q[child + 1]) > 0) ++child;
                if (cmp(x, q[child]) <= 0) break;
                q[k] = q[child];
                k = child;
            }
            q[k] = x;
        }
        return top;
    }
};

struct TransformSpec {
    bool any = false;
    int code[CUB_MAX_DIM];
    double shift[CUB_MAX_DIM];
    double tmin[CUB_MAX_DIM], tmax[CUB_MAX_DIM];
};

int make_transform(int dim, const double* xmin, const double* xmax, const int* transform, TransformSpec* t) {
    t->any = false;
    for (int d = 0; d < dim; ++d) {
        int code;
        if (transform) {
            code = transform[d];
            if (code < 0 || code > 3) return CUB_ERR_BAD_ARGS;
        } else {
            const bool lo_inf = std::isinf(xmin[d]) && xmin[d] < 0, hi_inf = std::isinf(xmax[d]) && xmax[d] > 0;
            code = lo_inf && hi_inf ? CUB_TRANSFORM_INF : hi_inf ? CUB_TRANSFORM_SEMI_INF_UP : lo_inf ? CUB_TRANSFORM_SEMI_INF_DOWN : 0;
        }
        t->code[d] = code;
        t->shift[d] = 0.0;
        t->tmin[d] = xmin[d];
        t->tmax[d] = xmax[d];
        if (code == CUB_TRANSFORM_SEMI_INF_UP) {
            t->shift[d] = xmin[d];
            t->tmin[d] = 0.0;
            t->tmax[d] = 1.0;
        } else if (code == CUB_TRANSFORM_SEMI_INF_DOWN) {
            t->shift[d] = xmax[d];
            t->tmin[d] = 0.0;
            t->tmax[d] = 1.0;
        } else if (code == CUB_TRANSFORM_INF) {
            t->tmin[d] = -1.0;
            t->tmax[d] = 1.0;
        }
        if (code) t->any = true;
    }
    return CUB_OK;
}

struct SelectBuffers {
    DevBuf keysA, keysB, idxA, idxB, hist, bin_totals, binbase, prefix, tile_tot, small;
    long long cap = 0;
    int fdim = 0;
    cudaError_t ensure_scalar(long long n) {  // fdim == 1 select path: list, tile counters, state
        cudaError_t e;
        if ((e = idxA.ensure((size_t)4 * n)) != cudaSuccess) return e;
        if ((e = keysA.ensure((size_t)8 * n)) != cudaSuccess) return e;  // candidate keys / slots of the cooperative select
        if ((e = keysB.ensure((size_t)8 * n)) != cudaSuccess) return e;  // val of the popped regions (iter_exact.h: stash_val)
        if ((e = idxB.ensure((size_t)4 * n)) != cudaSuccess) return e;
        if ((e = small.ensure(512 + PK_COOP_HIST_BYTES)) != cudaSuccess) return e;
        if ((e = hist.ensure((size_t)16 * pk_sel_tiles(n) + 64)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    SelScratch scratch(long long n) const {
        SelScratch sc;
        unsigned char* sm = small.as<unsigned char>();
        sc.state = sm;
        sc.ticket = (unsigned*)(sm + 128);
        sc.cnt = (unsigned long long*)(sm + 512);
        sc.sum = (double*)(sm + 512 + 256 * 8);
        sc.tile_less = hist.as<unsigned long long>();
        sc.tile_eq = sc.tile_less + pk_sel_tiles(n);
        sc.ctl = nullptr;
        return sc;
    }
    cudaError_t ensure(long long n, int f) {
        if (n <= cap && f == fdim) return cudaSuccess;
        const long long c = std::max<long long>(n, 1024);
        cudaError_t e;
        if ((e = keysA.ensure((size_t)8 * c)) != cudaSuccess) return e;
        if ((e = keysB.ensure((size_t)8 * c)) != cudaSuccess) return e;
        if ((e = idxA.ensure((size_t)4 * c)) != cudaSuccess) return e;
        if ((e = idxB.ensure((size_t)4 * c)) != cudaSuccess) return e;
        if ((e = hist.ensure((size_t)4 * 256 * pk_sort_chunks(c))) != cudaSuccess) return e;
        if ((e = bin_totals.ensure(8 * 256)) != cudaSuccess) return e;
        if ((e = binbase.ensure(8 * 256)) != cudaSuccess) return e;
        if ((e = prefix.ensure((size_t)8 * f * c)) != cudaSuccess) return e;
        if ((e = tile_tot.ensure((size_t)8 * f * pk_scan_tiles(c))) != cudaSuccess) return e;
        cap = c;
        fdim = f;
        return cudaSuccess;
    }
};

// Device / pinned buffers of one integrate call, cached per device between calls so that repeated
// integrations (the normal use: many integrals, one process) do not pay cudaMalloc / cudaFree each time.
// cubature_cuda_release_workspace() returns the memory.
struct Workspace {
    int device = -1;
    bool busy = false;
    Pool pool;
    SelectBuffers sel;
    DevBuf d_list, d_stage_val, d_stage_err, d_partial, d_small;
    DevBuf g_send, g_recv, g_dense;  // sharded vector select: padded send block, gathered blocks, dense [f + 1][Nglobal]
    DevBuf d_exact;                  // device-resident loop (iter_exact.h): control block, exact sums, histograms
    DevBuf d_vec;                    // device-resident loop of vector integrands (iter_vec.h)
    PinnedBuf h_list, h_stage, h_small, h_exact;
    long long stage_stride = 0;
};
std::mutex g_ws_mu;
std::vector<std::unique_ptr<Workspace>> g_ws;

Workspace* acquire_workspace(int device) {
    std::lock_guard<std::mutex> lock(g_ws_mu);
    for (auto& w : g_ws)
        if (!w->busy && w->device == device) {
            w->busy = true;
            return w.get();
        }
    g_ws.emplace_back(new Workspace());
    g_ws.back()->device = device;
    g_ws.back()->busy = true;
    return g_ws.back().get();
}
void release_workspace(Workspace* w) {
    if (!w) return;
    std::lock_guard<std::mutex> lock(g_ws_mu);
    w->busy = false;
    // keep at most two idle workspaces per device
    int idle = 0;
    for (size_t i = 0; i < g_ws.size();) {
        if (!g_ws[i]->busy && g_ws[i]->device == w->device && ++idle > 2) {
            cudaSetDevice(g_ws[i]->device);
            g_ws.erase(g_ws.begin() + (long)i);
        } else {
            ++i;
        }
    }
}

// One integrate / eval call: device, stream, kernels, launch helpers.
struct Session {
    cub_integrand* integ = nullptr;
    int dim = 0, fdim = 0, device = 0;
    int64_t P = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> evpool;  // (start, stop) pairs around the rule launches not yet added to rule_ms
    size_t ev_used = 0;
    RuleKernels* kern = nullptr;
    TransformSpec tr;
    std::vector<unsigned char> tr_arg;   // device-side CubTransform bytes
    std::vector<double> prm_arg;         // device-side CubParams
    int64_t launches = 0;
    int n_sms = 148;
    double rule_ms = 0.0;
    char* msg = nullptr;
    char msgbuf[256];
    Workspace* ws = nullptr;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int64_t local_regions_evaluated = 1;  // regions this rank evaluated (the root box counts for everyone)

    ~Session() {
        release_workspace(ws);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (cudaEvent_t e : evpool) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }

    int open(cub_integrand* in, int dim_, int device_, const TransformSpec& t, cub_stats_t* stats) {
        integ = in;
        dim = dim_;
        fdim = in->fdim;
        tr = t;
        msgbuf[0] = 0;
        msg = stats ? stats->message : msgbuf;
        P = cub_points_per_region(dim);
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            cub_set_message(msg, "no CUDA device available (libcubature_cuda has no CPU fallback)");
            return CUB_ERR_NO_DEVICE;
        }
        if (device_ < 0) {
            CUB_CUDA_CHECK(cudaGetDevice(&device), msg);
        } else {
            device = device_;
            CUB_CUDA_CHECK(cudaSetDevice(device), msg);
        }
        ws = acquire_workspace(device);
        cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device);
        CUB_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), msg);
        CUB_CUDA_CHECK(cudaEventCreate(&ev0), msg);
        CUB_CUDA_CHECK(cudaEventCreate(&ev1), msg);
        std::string log;
        int rc = cub_jit_get_kernels(integ, tr.any ? 1 : 0, device, &kern, &log);
        if (rc != CUB_OK) {
            cub_set_message(msg, "%s", log.c_str());
            return rc;
        }
        // CubTransform { int code[DIM]; double shift[DIM]; }
        const size_t off = ((size_t)4 * dim + 7) & ~(size_t)7;
        tr_arg.assign(off + (size_t)8 * dim, 0);
        memcpy(tr_arg.data(), tr.code, (size_t)4 * dim);
        memcpy(tr_arg.data() + off, tr.shift, (size_t)8 * dim);
        prm_arg = integ->params;
        if (prm_arg.empty()) prm_arg.push_back(0.0);
        return CUB_OK;
    }

    // Launches the rule kernel for `work`; for fdim > 1 a BISECT request is split into the stand-alone
    // bisect kernel + CHILDREN mode.  Rule-kernel time is accumulated with a pair of events.
    int launch_rule(const PoolView& pool, HostWork work, bool errmax_rows = true) {
        if (work.n_items <= 0) return CUB_OK;
        if (ev_used + 2 > evpool.size()) {
            if (evpool.size() >= 512) {
                flush_rule_timer();
            } else {
                cudaEvent_t a = nullptr, b = nullptr;
                if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
                    cub_set_message(msg, "cudaEventCreate failed");
                    return CUB_ERR_CUDA;
                }
                evpool.push_back(a);
                evpool.push_back(b);
            }
        }
        cudaEvent_t evr0 = evpool[ev_used], evr1 = evpool[ev_used + 1];
        ev_used += 2;
        const bool one_thread_per_region = (dim >= 2 && fdim == 1) || (dim == 1 && fdim == 1);
        if (work.mode == MODE_BISECT && !one_thread_per_region) {
            pk_bisect(pool, dim, work.list, work.n_items / 2, work.base, stream);
            ++launches;
            work.mode = MODE_CHILDREN;
        }
        PoolView pv = pool;
        void* args[5] = {&pv, &work, prm_arg.data(), tr_arg.data(), nullptr};
        cudaEventRecord(evr0, stream);
        cudaError_t e;
        if (dim >= 2 && fdim == 1) {
            // persistent tiles: a bounded number of CTAs, each walks the batch with the grid as its stride (the device-resident
            // loop sizes launches by an upper bound of the batch).  16 CTAs per resident slot: with a single wave the four CTAs
            // of an SM run their load / compute / store phases in step and the kernel is 10 % slower (6.5e9 vs 7.2e9 regions/s,
            // profiles/r02_acc_cost.md); CUBATURE_RULE_WAVES overrides (0 = one CTA per tile).
            const int threads = kern->gm_threads;  // the kernel's __launch_bounds__ (CUB_GM_THREADS)
            const long long tiles = (work.n_items + threads - 1) / threads;
            static const int waves = [] { const char* e = getenv("CUBATURE_RULE_WAVES"); return e ? atoi(e) : 16; }();
            const unsigned grid = (unsigned)(waves > 0 ? std::min<long long>(tiles, (long long)n_sms * kern->gm_blocks_per_sm * waves) : tiles);
            // Experiment, off by default (CUBATURE_SMALL_BATCH=n enables it for batches below n items): a small device-sized
            // bisecting batch goes to the point-parallel staged kernel (one thread per (region, point), then the serial sums)
            // instead of paying the serial latency of a whole region per thread.  Both kernels are enqueued; each reads the
            // batch size from the control block and the other one retires at once.  Results are bit-identical (GPU suite green
            // with n = 2048), but the second launch per iteration costs more (4-5 us) than the small batches gain: C3 time to
            // tolerance +0.2-0.3 ms for five of six families (profiles/r02/late/small_batch_time_to_tol.log).
            const char* small_env = getenv("CUBATURE_SMALL_BATCH");  // (read per launch: tools toggle it within one process)
            const long long small = small_env ? atoll(small_env) : 0LL;
            const bool split_by_size = small >= 2 && work.dyn && work.mode == MODE_BISECT && !work.stage_val && kern->staged;
            if (split_by_size) {
                work.dyn_lo = small;
                work.dyn_hi = LLONG_MAX;
            }
            e = cudaLaunchKernel((const void*)kern->gm_scalar, dim3(grid), dim3(threads), args, 0, stream);
            ++launches;
            if (split_by_size && e == cudaSuccess) {
                work.dyn_lo = 0;
                work.dyn_hi = small;
                int rpb = 2;  // even: both children of a parent in one pass (rule_kernels.cuh: cub_rule_staged, BISECT)
                const size_t per_region = ((size_t)P * 1 + 3 * (size_t)dim + 1) * 8 + 24;
                const long long groups = (std::min<long long>(work.n_items, small) + rpb - 1) / rpb;
                const unsigned sgrid = (unsigned)std::max<long long>(1, std::min<long long>(groups, (long long)n_sms * 8));
                args[4] = &rpb;
                e = cudaLaunchKernel((const void*)kern->staged, dim3(sgrid), dim3(256), args, per_region * rpb, stream);
                ++launches;
            }
        } else if (dim == 1 && (fdim == 1 || integ->separable)) {
            const long long threads = work.n_items * fdim;
            const unsigned grid = (unsigned)std::min<long long>((threads + 127) / 128, (long long)n_sms * kern->gk_blocks_per_sm);
            e = cudaLaunchKernel((const void*)kern->gk_separable, dim3(grid), dim3(128), args, 0, stream);
            ++launches;
            if (e == cudaSuccess && fdim > 1 && errmax_rows) {
                pk_errmax_rows(pool, fdim, work.list, work.n_items, work.base, work.mode, stream);
                ++launches;
            }
        } else {
            const int FS = fdim | 1;
            const size_t per_region = ((size_t)P * FS + 3 * (size_t)dim + (size_t)fdim) * 8 + 24;
            const size_t max_smem = (size_t)kern->staged_max_smem;
            if (per_region > max_smem) {
                cub_set_message(msg, "fdim=%d at dim=%d needs %zu bytes of shared memory per region (max %zu)", fdim, dim,
                                per_region, max_smem);
                return CUB_ERR_UNSUPPORTED;
            }
            size_t budget = 100 * 1024;  // two CTAs per SM
            if (per_region > budget) budget = max_smem;
            int rpb = (int)std::min<size_t>(budget / per_region, 16);
            if (rpb < 1) rpb = 1;
            const long long groups = (work.n_items + rpb - 1) / rpb;
            const unsigned grid = (unsigned)std::min<long long>(groups, 148 * 8);
            args[4] = &rpb;
            e = cudaLaunchKernel((const void*)kern->staged, dim3(grid), dim3(256), args, per_region * rpb, stream);
            ++launches;
        }
        cudaEventRecord(evr1, stream);
        if (e != cudaSuccess) {
            cub_set_message(msg, "rule kernel launch failed: %s", cudaGetErrorString(e));
            return CUB_ERR_CUDA;
        }
        return CUB_OK;
    }
    void flush_rule_timer() {
        if (!ev_used) return;
        if (cudaEventSynchronize(evpool[ev_used - 1]) == cudaSuccess) {
            for (size_t i = 0; i + 1 < ev_used; i += 2) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, evpool[i], evpool[i + 1]) == cudaSuccess) rule_ms += ms;
            }
        }
        ev_used = 0;
    }
};

int set_root(Session& S, Pool& pool) {
    // Hypercube.initFromRanges, Hypercube.java:19-30 (in t-space)
    std::vector<double> c(S.dim), h(S.dim);
    for (int d = 0; d < S.dim; ++d) {
        c[d] = (S.tr.tmax[d] + S.tr.tmin[d]) / 2.0;
        h[d] = (S.tr.tmax[d] - S.tr.tmin[d]) / 2.0;
    }
    CUB_CUDA_CHECK(cudaMemcpy2DAsync(pool.center.p, (size_t)pool.cap * 8, c.data(), 8, 8, S.dim, cudaMemcpyHostToDevice, S.stream), S.msg);
    CUB_CUDA_CHECK(cudaMemcpy2DAsync(pool.halfw.p, (size_t)pool.cap * 8, h.data(), 8, 8, S.dim, cudaMemcpyHostToDevice, S.stream), S.msg);
    CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
    S.h2d_bytes += 16 * S.dim;
    return CUB_OK;
}

long long auto_capacity(int dim, int fdim, int64_t P, int64_t maxEval, long long requested, size_t extra_per_region) {
    if (requested > 0) return requested;
    // unknown amount of work: start with about 256 MB of pool (at least 4 096, at most 2^21 regions) and grow on demand
    long long cap = std::min<long long>(1 << 21, std::max<long long>(1 << 12, (long long)(((size_t)256 << 20) / Pool::bytes_per_region(dim, fdim))));
    if (maxEval > 0) {
        const long long need = maxEval / (2 * P) + 8;
        cap = std::max<long long>(need, 1024);
    }
    size_t free_b = 0, total_b = 0;
    if (MemInfoCache::get().query(&free_b, &total_b) == cudaSuccess) {
        const size_t per = Pool::bytes_per_region(dim, fdim) + extra_per_region;
        const long long fit = (long long)((free_b * 0.85) / per);
        cap = std::min(cap, std::max<long long>(fit, 1024));
    }
    return cap;
}

void fill_trace_regions(Session& S, Pool& pool, const std::vector<int>* order, long long n, cub_trace_t* tr) {
    if (!tr) return;
    tr->n_regions = n;
    const long long m = std::min<long long>(n, tr->region_cap);
    if (m <= 0) return;
    // pull the SoA rows [0, pool_n) to the host and re-pack as [n][dim] / [n][fdim]
    long long span = 0;
    if (order) {
        for (long long i = 0; i < m; ++i) span = std::max<long long>(span, (*order)[i] + 1);
    } else {
        span = m;
    }
    std::vector<double> row((size_t)span);
    std::vector<int> irow((size_t)span);
    auto slot = [&](long long i) { return order ? (long long)(*order)[i] : i; };
    for (int d = 0; d < S.dim; ++d) {
        if (tr->centers) {
            cudaMemcpy(row.data(), pool.center.as<double>() + (size_t)d * pool.cap, (size_t)span * 8, cudaMemcpyDeviceToHost);
            for (long long i = 0; i < m; ++i) tr->centers[i * S.dim + d] = row[(size_t)slot(i)];
        }
        if (tr->halfwidths) {
            cudaMemcpy(row.data(), pool.halfw.as<double>() + (size_t)d * pool.cap, (size_t)span * 8, cudaMemcpyDeviceToHost);
            for (long long i = 0; i < m; ++i) tr->halfwidths[i * S.dim + d] = row[(size_t)slot(i)];
        }
    }
    for (int j = 0; j < S.fdim; ++j) {
        if (tr->val) {
            cudaMemcpy(row.data(), pool.val.as<double>() + (size_t)j * pool.cap, (size_t)span * 8, cudaMemcpyDeviceToHost);
            for (long long i = 0; i < m; ++i) tr->val[i * S.fdim + j] = row[(size_t)slot(i)];
        }
        if (tr->err) {
            cudaMemcpy(row.data(), pool.err.as<double>() + (size_t)j * pool.cap, (size_t)span * 8, cudaMemcpyDeviceToHost);
            for (long long i = 0; i < m; ++i) tr->err[i * S.fdim + j] = row[(size_t)slot(i)];
        }
    }
    if (tr->split_dim) {
        cudaMemcpy(irow.data(), pool.splitdim.as<int>(), (size_t)span * 4, cudaMemcpyDeviceToHost);
        for (long long i = 0; i < m; ++i) tr->split_dim[i] = irow[(size_t)slot(i)];
    }
}

// Iteration guard: with maxEval == 0 and an unreachable tolerance (e.g. err == 0 == val, strict `<`)
// the reference loops forever.  Every iteration at least doubles nothing, so the
// guard is generous; hitting it returns the best estimate with converged = 0.
const int64_t kMaxIterations = 100000;
const long long kDeviceSizedMaxRegions = 1 << 20;  // below this the rule kernel is sized on the device (no round trip)

struct LoopResult {
    int converged = 0;
    int64_t num_eval = 0, num_regions = 0, iterations = 0, ambiguous = 0, rebalances = 0;
};

// state handed from the replicated (single-GPU style) phase to the sharded phase of a multi-GPU run
struct LoopState {
    long long N = 1;
    int64_t numEval = 0;
    double ops = 1.0;
    bool handoff = false;
    long long cap_hint = 0;  // capacity maxEval can need on this rank (upper bound for pool growth)
};
int integrate_device_fused(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                           LoopResult* res, long long stop_at_regions = 0, LoopState* state_out = nullptr, long long cap_override = 0);

int integrate_sharded_peer(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                           LoopResult* res, const LoopState& st0);

// When is a pool dealt out to the ranks?  Dealing costs a ranked sort of the pool (1 ms at 1e5 regions, 8 ms at 8e6) and every
// sharded iteration a few peer exchanges (50 - 100 us), which only pays once a rule launch is longer than that: above about
// 1e6 regions per iteration.  A job whose evaluation budget says it will grow far beyond that (BASELINE config C5: 3.4e8
// regions) is dealt out early, while the pool is small and the deal is cheap; a job without such a budget (tolerance-driven,
// config C3: 1e3 ... 4e6 regions) stays replicated -- every rank computes the identical result, as fast as one GPU -- until
// it has proven large (2^23 regions).  Measured on 2 B200 (profiles/r02/summary.md): C5 59.4 ms either way; C3 corner peak
// 10.9 ms dealt out at 65 536 regions, 8.2 ms replicated (one GPU: 8.1 ms).
long long default_shard_min(int R, int64_t maxEval, int64_t P) {
    const bool known_large = maxEval > 0 && maxEval / (2 * P) >= (1LL << 24);
    return known_large ? 32768LL * R : (1LL << 23);
}

// CUBATURE_LOOP=legacy: the round-1 loop (k_iter_coop on floating-point sums, one host visit per iteration)
bool legacy_loop() {
    static const bool v = [] { const char* e = getenv("CUBATURE_LOOP"); return e && strcmp(e, "legacy") == 0; }();
    return v;
}

void push_batch(cub_trace_t* tr, int64_t nR) {
    if (!tr) return;
    if (tr->batches && tr->n_batches < tr->batch_cap) tr->batches[tr->n_batches] = nR;
    tr->n_batches += 1;
}

// ---- HEAP_EXACT: Rule.cubature replayed on the host, rule evaluation + bisection on the GPU --------
struct LoopState;
int replicated_phase_exact(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                           LoopResult* res, long long stop_at_regions, LoopState* st0, long long cap);
int integrate_device_vec(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                         LoopResult* res, long long stop_at_regions = 0, LoopState* state_out = nullptr, long long cap_override = 0);
int integrate_heap_exact(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt,
                         double* out, LoopResult* res) {
    const int f = S.fdim;
    const int64_t P = S.P;
    cub_trace_t* trace = opt ? opt->trace : nullptr;
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    long long cap = auto_capacity(S.dim, f, P, maxEval, opt ? opt->pool_capacity : 0, (size_t)16 * f + 4);
    if (!(opt && opt->pool_capacity > 0)) cap = std::min<long long>(cap, 1 << 18);  // grows on demand
    CUB_CUDA_CHECK(pool.alloc(S.dim, f, cap), S.msg);
    int rc = set_root(S, pool);
    if (rc != CUB_OK) return rc;

    DevBuf &d_list = W.d_list, &d_stage_val = W.d_stage_val, &d_stage_err = W.d_stage_err;
    PinnedBuf &h_list = W.h_list, &h_stage = W.h_stage;
    long long& stage_stride = W.stage_stride;
    stage_stride = 0;  // the cached buffers may have been sized for another fdim: re-derive the stride
    auto ensure_stage = [&](long long n_items) -> int {
        if (n_items <= stage_stride) return CUB_OK;
        long long ns = std::max<long long>(n_items * 2, 1024);
        CUB_CUDA_CHECK(d_stage_val.ensure((size_t)8 * f * ns), S.msg);
        CUB_CUDA_CHECK(d_stage_err.ensure((size_t)8 * f * ns), S.msg);
        CUB_CUDA_CHECK(h_stage.ensure((size_t)16 * f * ns), S.msg);
        CUB_CUDA_CHECK(d_list.ensure((size_t)4 * ns), S.msg);
        CUB_CUDA_CHECK(h_list.ensure((size_t)4 * ns), S.msg);
        stage_stride = ns;
        return CUB_OK;
    };

    std::vector<double> hval, herr, herrmax;  // per slot, host mirror ([slot][f])
    auto reserve_slots = [&](long long n) {
        if ((long long)herrmax.size() < n) {
            const size_t m = (size_t)std::max<long long>(n, (long long)herrmax.size() * 2);
            hval.resize(m * f);
            herr.resize(m * f);
            herrmax.resize(m);
        }
    };
    SlotHeap heap;
    heap.key = &herrmax;
    std::vector<double> sum_val(f, 0.0), sum_err(f, 0.0);  // RegionHeap.ee (running, incremental)
    auto heap_add = [&](int slot) {  // RegionHeap.java:30-36
        for (int j = 0; j < f; ++j) {
            sum_val[j] += hval[(size_t)slot * f + j];
            sum_err[j] += herr[(size_t)slot * f + j];
        }
        heap.add(slot);
    };
    // evaluates n_items work items and returns their val/err in h_stage: [val f rows][err f rows], row stride = n_items
    auto run_rule = [&](HostWork w) -> int {
        int r = ensure_stage(w.n_items);
        if (r != CUB_OK) return r;
        w.stage_val = d_stage_val.as<double>();
        w.stage_err = d_stage_err.as<double>();
        w.stage_stride = stage_stride;
        r = S.launch_rule(pool.view(), w);
        if (r != CUB_OK) return r;
        const size_t width = (size_t)w.n_items * 8;
        CUB_CUDA_CHECK(cudaMemcpy2DAsync(h_stage.p, width, d_stage_val.p, (size_t)stage_stride * 8, width, f,
                                         cudaMemcpyDeviceToHost, S.stream), S.msg);
        CUB_CUDA_CHECK(cudaMemcpy2DAsync(h_stage.as<double>() + (size_t)f * w.n_items, width, d_stage_err.p,
                                         (size_t)stage_stride * 8, width, f, cudaMemcpyDeviceToHost, S.stream), S.msg);
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
        S.d2h_bytes += (int64_t)2 * width * f;
        return CUB_OK;
    };
    auto store_item = [&](long long item, long long n_items, int slot) {
        const double* sv = h_stage.as<double>();
        const double* se = sv + (size_t)f * n_items;
        double emax = 0.0;  // Rule.errMax, Rule.java:281-289
        for (int j = 0; j < f; ++j) {
            const double v = sv[(size_t)j * n_items + item], e = se[(size_t)j * n_items + item];
            hval[(size_t)slot * f + j] = v;
            herr[(size_t)slot * f + j] = e;
            if (e > emax) emax = e;
        }
        herrmax[(size_t)slot] = emax;
    };

    // Rule.java:62-69: the whole box
    reserve_slots(1);
    {
        HostWork w{};
        w.n_items = 1;
        w.base = 0;
        w.mode = MODE_PLAIN;
        rc = run_rule(w);
        if (rc != CUB_OK) return rc;
        store_item(0, 1, 0);
        heap_add(0);
    }
    long long N = 1;
    int64_t numEval = P;
    bool conv = false;
    std::vector<double> res_err(f), res_val(f);
    std::vector<int> list;
    while (numEval < maxEval || maxEval == 0) {  // Rule.java:72
        if (cub_converged(f, VecAcc{sum_val.data(), sum_err.data()}, absTol, relTol, norm)) {
            conv = true;
            break;
        }
        for (int j = 0; j < f; ++j) {  // Rule.java:102-107: a copy of BOTH sums; only err is updated below
            res_err[j] = sum_err[j];
            res_val[j] = sum_val[j];
        }
        list.clear();
        do {  // Rule.java:109-133
            const int s = heap.poll();
            for (int j = 0; j < f; ++j) {  // RegionHeap.java:21-28
                sum_val[j] -= hval[(size_t)s * f + j];
                sum_err[j] -= herr[(size_t)s * f + j];
            }
            for (int j = 0; j < f; ++j) res_err[j] -= herr[(size_t)s * f + j];
            list.push_back(s);
            numEval += 2 * P;
            if (cub_converged(f, VecAcc{res_val.data(), res_err.data()}, absTol, relTol, norm)) break;
        } while (!heap.q.empty() && (numEval < maxEval || maxEval == 0));
        const long long K = (long long)list.size();
        if (N + K > pool.cap) {
            long long nc = std::max<long long>(pool.cap * 2, N + K);
            cudaError_t e = pool.grow(nc, N, S.stream);
            if (e != cudaSuccess) {
                cudaGetLastError();
                cub_set_message(S.msg, "region pool exhausted at %lld regions (%s)", N, cudaGetErrorString(e));
                return CUB_ERR_POOL_EXHAUSTED;
            }
        }
        if (N + K > (long long)std::numeric_limits<int>::max()) {
            cub_set_message(S.msg, "region pool exceeds 2^31 slots");
            return CUB_ERR_POOL_EXHAUSTED;
        }
        rc = ensure_stage(2 * K);
        if (rc != CUB_OK) return rc;
        memcpy(h_list.p, list.data(), (size_t)4 * K);
        CUB_CUDA_CHECK(cudaMemcpyAsync(d_list.p, h_list.p, (size_t)4 * K, cudaMemcpyHostToDevice, S.stream), S.msg);
        S.h2d_bytes += 4 * K;
        HostWork w{};
        w.list = d_list.as<int>();
        w.n_items = 2 * K;
        w.base = N;
        w.mode = MODE_BISECT;
        rc = run_rule(w);
        if (rc != CUB_OK) return rc;
        reserve_slots(N + K);
        for (long long i = 0; i < 2 * K; ++i) {  // Rule.java:137-140: left0, right0, left1, right1, ...
            const int slot = (i & 1) ? (int)(N + (i >> 1)) : list[(size_t)(i >> 1)];
            store_item(i, 2 * K, slot);
            heap_add(slot);
        }
        N += K;
        res->iterations += 1;
        push_batch(trace, 2 * K);
        if (res->iterations >= kMaxIterations) {  // the reference would spin forever (App. C); we stop and say so
            cub_set_message(S.msg, "stopped after %lld iterations without convergence", (long long)res->iterations);
            break;
        }
    }
    // Rule.java:148-159: re-sum in heap ARRAY order
    for (int j = 0; j < f; ++j) out[j] = out[f + j] = 0;
    for (size_t i = 0; i < heap.q.size(); ++i) {
        const int s = heap.q[i];
        for (int j = 0; j < f; ++j) {
            out[j] += hval[(size_t)s * f + j];
            out[f + j] += herr[(size_t)s * f + j];
        }
    }
    S.flush_rule_timer();
    res->converged = conv ? 1 : 0;
    res->num_eval = numEval;
    res->num_regions = (int64_t)heap.q.size();
    fill_trace_regions(S, pool, &heap.q, (long long)heap.q.size(), trace);
    return CUB_OK;
}

// ---- DEVICE mode ------------------------------------------------------------------------------------
// The pool can be sharded over the ranks of a communicator (options->comm).  Until the pool holds
// shard_min_regions regions every rank runs the same deterministic iterations on a full copy (no
// communication, identical bits everywhere); then rank r keeps the regions in slots r, r+R, r+2R, ...
// and from there on refines only its own regions.  Per iteration the ranks exchange one small block
// (local sums of val/err, the smallest-errmax region, the local region count) with ONE all-gather and
// add the blocks in rank order, so every rank takes the same decision; the distributed batch selection
// exchanges one 4 KB histogram per radix pass.  Children stay with the rank that owns the parent.
int integrate_device(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                     LoopResult* res) {
    const int f = S.fdim;
    const int64_t P = S.P;
    cub_trace_t* trace = opt ? opt->trace : nullptr;
    cub_comm* comm = opt ? opt->comm : nullptr;
    const int R = comm ? comm->nranks : 1, rank = comm ? comm->rank : 0;
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    const size_t extra = 24 + (size_t)8 * f + 1;  // sort keys/indices + prefix sums per region
    long long cap = auto_capacity(S.dim, f, P, maxEval, opt ? opt->pool_capacity : 0, extra);
    // (measured on 4 and 8 B200: sharding later, at 2^21 regions, is slower -- ranking 2^21 regions for the
    // deal costs more than the cheaper replicated iterations save)
    long long shard_min = opt && opt->shard_min_regions > 0 ? opt->shard_min_regions : default_shard_min(R, maxEval, P);
    if (R > 1 && !(opt && opt->pool_capacity > 0) && maxEval > 0) {
        // a rank holds about 1/R of the final pool (plus imbalance head-room), but at least the replicated phase
        const long long total = maxEval / (2 * P) + 8;
        cap = std::min(cap, std::max<long long>(total / R + total / (4 * R) + 1024, 4 * shard_min));
    }
    // replicated phase: every rank runs the single-GPU device-driven loop on a full copy until the pool is
    // large enough to be dealt out (identical, deterministic kernels => identical pools on every rank)
    LoopState st0;
    int rc;
    if (legacy_loop()) {
        rc = integrate_device_fused(S, relTol, absTol, norm, maxEval, opt, out, res, R > 1 ? shard_min : 0, &st0, cap);
    } else if (f == 1) {  // the device-resident loops (the round-1 kernel k_iter_coop is not on any default path any more)
        rc = replicated_phase_exact(S, relTol, absTol, norm, maxEval, opt, out, res, R > 1 ? shard_min : 0, &st0, cap);
    } else {
        rc = integrate_device_vec(S, relTol, absTol, norm, maxEval, opt, out, res, R > 1 ? shard_min : 0, &st0, cap);
    }
    if (rc != CUB_OK || !st0.handoff) return rc;  // finished (converged / maxEval) before sharding: result already complete
    {
        const char* env = getenv("CUBATURE_PEER_LOOP");
        if (R > 1 && comm->peer_ok && f == 1 && !(env && env[0] == '0'))
            return integrate_sharded_peer(S, relTol, absTol, norm, maxEval, opt, out, res, st0);
    }

    SelectBuffers& sel = W.sel;
    DevBuf &d_partial = W.d_partial, &d_small = W.d_small;
    PinnedBuf& h_small = W.h_small;
    // exchange block (doubles): totals[2f] | minvec[2f] | u64: min key, max key, argmin slot, local N
    const size_t blk = (size_t)4 * f + 4;
    CUB_CUDA_CHECK(d_partial.ensure((size_t)8 * 2 * f * 1184), S.msg);
    CUB_CUDA_CHECK(d_small.ensure(blk * 8 * (size_t)(R + 1) + 64 + (size_t)R * PK_SEL_HIST_BYTES), S.msg);
    CUB_CUDA_CHECK(h_small.ensure(blk * 8 * (size_t)(R + 1) + 64), S.msg);
    double* d_totals = d_small.as<double>();
    double* d_minvec = d_totals + 2 * f;
    unsigned long long* d_u64 = (unsigned long long*)(d_minvec + 2 * f);  // [0..3] part of the block, [4..7] scratch
    double* d_gather = d_totals + blk + 8;                                 // [R][blk]
    unsigned char* d_histgather = (unsigned char*)(d_gather + blk * R);    // [R][PK_SEL_HIST_BYTES]
    double* h_block = h_small.as<double>();                                // [R][blk] after the exchange
    unsigned long long* h_u64 = (unsigned long long*)(h_block + blk * R);  // scratch [8]

    long long N = st0.N;          // regions held by this rank
    long long Nglobal = st0.N;    // regions of the whole partition
    bool sharded = false;
    int64_t numEval = st0.numEval;
    bool conv = false;
    double ops = st0.ops;  // additions the reference's running sums have seen so far (rounding band)
    std::vector<double> tmp_err(f), tot(2 * f), minvec(2 * f);
    std::vector<long long> Nr((size_t)R, 0);  // regions per rank (from the per-iteration exchange)
    const double eps = 2.220446049250313e-16;
    const bool timing = getenv("CUBATURE_TIMING") != nullptr;
    double t_stats = 0, t_select = 0, t_rule = 0, t_shard = 0;
    int n_sel_iters = 0, n_iters = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };

    while (numEval < maxEval || maxEval == 0) {
        auto t_iter = now();
        if (R > 1 && !sharded && N >= shard_min) {
            // deal the replicated pool out: rank r keeps slots r, r+R, ...
            const long long n_local = cubature_cuda_shard_count(N, R, rank);
            // gather into a small temporary (the pool is still tiny here), then copy the rows back in place
            Pool tmp;
            CUB_CUDA_CHECK(tmp.alloc(S.dim, f, std::max<long long>(n_local, 1), true), S.msg);
            // rank the regions by errmax (stable radix sort, identical on every rank) and deal round robin
            CUB_CUDA_CHECK(sel.ensure(N, f), S.msg);
            unsigned long long* ka = sel.keysA.as<unsigned long long>();
            unsigned long long* kb = sel.keysB.as<unsigned long long>();
            int* ia = sel.idxA.as<int>();
            int* ib = sel.idxB.as<int>();
            pk_sort_prepare(pool.view().errmax, N, ka, ia, d_u64 + 4, S.stream);
            CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64 + 4, 16, cudaMemcpyDeviceToHost, S.stream), S.msg);
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            const unsigned long long vary0 = h_u64[0] ^ h_u64[1];  // digits shared by every key need no pass
            for (int shift = 0; shift < 64; shift += 8) {
                if (((vary0 >> shift) & 0xffULL) == 0) continue;
                pk_sort_pass(ka, ia, kb, ib, N, shift, sel.hist.as<unsigned>(), sel.bin_totals.as<unsigned long long>(),
                             sel.binbase.as<unsigned long long>(), S.stream);
                std::swap(ka, kb);
                std::swap(ia, ib);
            }
            pk_shard_gather(pool.view(), tmp.view(), S.dim, f, n_local, R, rank, ia, S.stream);
            S.launches += 34;
            auto rows = [&](DevBuf& dst, DevBuf& src, int nrows, size_t elt) {
                return cudaMemcpy2DAsync(dst.p, (size_t)pool.cap * elt, src.p, (size_t)tmp.cap * elt, (size_t)n_local * elt, nrows,
                                         cudaMemcpyDeviceToDevice, S.stream);
            };
            if (n_local > 0) {
                CUB_CUDA_CHECK(rows(pool.center, tmp.center, S.dim, 8), S.msg);
                CUB_CUDA_CHECK(rows(pool.halfw, tmp.halfw, S.dim, 8), S.msg);
                CUB_CUDA_CHECK(rows(pool.val, tmp.val, f, 8), S.msg);
                CUB_CUDA_CHECK(rows(pool.err, tmp.err, f, 8), S.msg);
                CUB_CUDA_CHECK(rows(pool.errmax, tmp.errmax, 1, 8), S.msg);
                CUB_CUDA_CHECK(rows(pool.splitdim, tmp.splitdim, 1, 4), S.msg);
            }
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            N = n_local;
            sharded = true;
            t_shard += ms_since(t_iter);
            t_iter = now();
        }
        const PoolView pv = pool.view();
        pk_totals(pv, f, N, d_partial.as<double>(), d_totals, S.stream);
        pk_min_region(pv, f, N, d_u64, d_minvec, S.stream);
        S.launches += 5;
        S.h2d_bytes += 24;
        unsigned long long min_key = 0;
        if (!sharded) {
            CUB_CUDA_CHECK(cudaMemcpyAsync(h_block, d_totals, blk * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
            S.d2h_bytes += (int64_t)blk * 8;
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            for (int j = 0; j < 2 * f; ++j) {
                tot[j] = h_block[j];
                minvec[j] = h_block[2 * f + j];
            }
            min_key = ((unsigned long long*)(h_block + 4 * f))[0];
        } else {
            const unsigned long long nloc = (unsigned long long)N;
            CUB_CUDA_CHECK(cudaMemcpyAsync(d_u64 + 3, &nloc, 8, cudaMemcpyHostToDevice, S.stream), S.msg);
            rc = cub_comm_allgather(comm, d_totals, blk * 8, d_gather, S.stream);
            if (rc != CUB_OK) return rc;
            CUB_CUDA_CHECK(cudaMemcpyAsync(h_block, d_gather, blk * 8 * R, cudaMemcpyDeviceToHost, S.stream), S.msg);
            S.d2h_bytes += (int64_t)blk * 8 * R;
            S.h2d_bytes += 8;
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            // add the per-rank blocks in rank order: identical bits on every rank
            for (int j = 0; j < 2 * f; ++j) tot[j] = 0.0;
            Nglobal = 0;
            int best = -1;
            for (int r = 0; r < R; ++r) {
                const double* b = h_block + (size_t)r * blk;
                const unsigned long long* u = (const unsigned long long*)(b + 4 * f);
                Nr[(size_t)r] = (long long)u[3];
                if (u[3] == 0) continue;  // a rank without regions contributes nothing
                for (int j = 0; j < 2 * f; ++j) tot[j] += b[j];
                Nglobal += (long long)u[3];
                if (best < 0 || u[0] <= min_key) {  // ties: the highest rank pops last
                    best = r;
                    min_key = u[0];
                }
            }
            for (int j = 0; j < 2 * f; ++j) minvec[j] = h_block[(size_t)best * blk + 2 * f + j];
            // the selection kernels read the GLOBAL totals
            CUB_CUDA_CHECK(cudaMemcpyAsync(d_totals, tot.data(), (size_t)16 * f, cudaMemcpyHostToDevice, S.stream), S.msg);
            S.h2d_bytes += 16 * f;
        }
        if (!sharded) Nglobal = N;
        t_stats += ms_since(t_iter);
        t_iter = now();
        ++n_iters;
        const double band = 4.0 * eps * std::sqrt(ops);
        const double* tv = tot.data();
        const double* te = tot.data() + f;
        {
            const bool c0 = cub_converged(f, VecAcc{tv, te}, absTol, relTol, norm);
            for (int j = 0; j < f; ++j) tmp_err[j] = te[j] * (1.0 + band);
            const bool c1 = cub_converged(f, VecAcc{tv, tmp_err.data()}, absTol, relTol, norm);
            for (int j = 0; j < f; ++j) tmp_err[j] = te[j] * (1.0 - band);
            const bool c2 = cub_converged(f, VecAcc{tv, tmp_err.data()}, absTol, relTol, norm);
            if (c0 != c1 || c0 != c2) res->ambiguous += 1;
            if (c0) {
                conv = true;
                break;
            }
        }
        long long Kcap = Nglobal;  // pops allowed by maxEval (Rule.java:133), over all ranks
        if (maxEval > 0) {
            const int64_t room = maxEval - numEval;  // > 0 here
            Kcap = std::min<long long>(Nglobal, std::max<int64_t>(1, (room + 2 * P - 1) / (2 * P)));
        }
        // select-all shortcut: the residual after popping all but the smallest-errmax region is that
        // region's own error vector; if that is not converged, no smaller k is (monotone predicate).
        const double* last_err = minvec.data() + f;
        bool all = !cub_converged(f, VecAcc{tv, last_err}, absTol, relTol, norm);
        {
            for (int j = 0; j < f; ++j) tmp_err[j] = last_err[j] + band * te[j];
            const bool a1 = !cub_converged(f, VecAcc{tv, tmp_err.data()}, absTol, relTol, norm);
            for (int j = 0; j < f; ++j) tmp_err[j] = last_err[j] - band * te[j];
            const bool a2 = !cub_converged(f, VecAcc{tv, tmp_err.data()}, absTol, relTol, norm);
            if (a1 != all || a2 != all) res->ambiguous += 1;
        }
        if (Nglobal == 1) all = true;
        long long K = N;            // pops of this rank
        long long Kglobal = Nglobal;
        const int* list = nullptr;
        if (!(all && Kcap == Nglobal) && f == 1) {
            // scalar integrand: weighted radix select on the device (no sort), then a stable compaction
            CUB_CUDA_CHECK(sel.ensure_scalar(pool.cap), S.msg);
            const SelScratch sc = sel.scratch(pool.cap);
            unsigned char* sm = (unsigned char*)sc.state;
            if (!sharded) {
                S.launches += pk_select_scalar(pv, N, d_totals, Kcap, absTol, relTol, norm, band, sc, sel.idxA.as<int>(), S.stream);
                CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, sm + 24, 24, cudaMemcpyDeviceToHost, S.stream), S.msg);
                S.d2h_bytes += 24;
                CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
                K = Kglobal = (long long)h_u64[0];  // SelState::K ; [1] = tie_take ; [2] = ambiguous
                if (h_u64[2]) res->ambiguous += 1;
            } else {
                // distributed: local histogram -> all-gather -> rank-ordered sum -> identical pick on every rank
                pk_sel_init(sc, S.stream);
                for (int pass = 0; pass < pk_sel_passes(); ++pass) {
                    pk_sel_pass(pv, N, sc, pass, 0, d_totals, Kcap, absTol, relTol, norm, band, nullptr, S.stream);
                    rc = cub_comm_allgather(comm, sc.cnt, PK_SEL_HIST_BYTES, d_histgather, S.stream);
                    if (rc != CUB_OK) return rc;
                    pk_sum_hist_ranks(d_histgather, R, sc.cnt, sc.sum, S.stream);
                    pk_sel_pick(sc, pass, d_totals, Kcap, absTol, relTol, norm, band, S.stream);
                    S.launches += 3;
                }
                // regions equal to the threshold key are taken in (rank, slot) order
                pk_sel_tile_counts(pv, N, sc.state, sc.tile_less, sc.tile_eq, d_u64 + 4, nullptr, S.stream);
                S.launches += 2;
                rc = cub_comm_allgather(comm, d_u64 + 4, 16, d_histgather, S.stream);
                if (rc != CUB_OK) return rc;
                CUB_CUDA_CHECK(cudaMemcpyAsync(h_block, d_histgather, (size_t)16 * R, cudaMemcpyDeviceToHost, S.stream), S.msg);
                CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, sm + 24, 24, cudaMemcpyDeviceToHost, S.stream), S.msg);
                S.d2h_bytes += 16 * R + 24;
                CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
                Kglobal = (long long)h_u64[0];
                const long long tie_take_global = (long long)h_u64[1];
                if (h_u64[2]) res->ambiguous += 1;
                const unsigned long long* le = (const unsigned long long*)h_block;  // [R][2]: less, equal
                std::vector<int64_t> ties((size_t)R);
                for (int r = 0; r < R; ++r) ties[(size_t)r] = (int64_t)le[2 * r + 1];
                const long long take = cubature_cuda_tie_share(tie_take_global, ties.data(), R, rank);
                K = (long long)le[2 * rank] + take;
                pk_sel_compact(pv, N, sc.state, sc.tile_less, sc.tile_eq, take, sel.idxA.as<int>(), nullptr, S.stream);
                S.launches += 1;
            }
            list = sel.idxA.as<int>();
            if (Kglobal < 1 || Kglobal > Nglobal || K < 0 || K > N) {
                cub_set_message(S.msg, "internal: device selection returned K=%lld of N=%lld (global %lld of %lld)", K, N, Kglobal, Nglobal);
                return CUB_ERR_CUDA;
            }
        } else if (!(all && Kcap == Nglobal) && sharded) {
            // Vector integrand, sharded pool, the batch has to be cut: every rank gathers errmax and err[f] of ALL regions
            // (one all-gather of (f + 1) rows padded to the largest shard), runs the single-GPU selection on the gathered
            // arrays (stable sort by errmax -- ties in (rank, slot) order --, per-component prefix sums, residual test for
            // every k: identical inputs, identical K on every rank) and keeps its own part of the first K regions.
            long long Nmax = 0, off = 0;
            for (int r = 0; r < R; ++r) {
                Nmax = std::max(Nmax, Nr[(size_t)r]);
                if (r < rank) off += Nr[(size_t)r];
            }
            if (Nglobal > (long long)std::numeric_limits<int>::max()) {
                cub_set_message(S.msg, "sharded vector selection: more than 2^31 regions in the partition");
                return CUB_ERR_UNSUPPORTED;
            }
            const size_t rows = (size_t)f + 1;
            CUB_CUDA_CHECK(W.g_send.ensure(rows * (size_t)Nmax * 8), S.msg);
            CUB_CUDA_CHECK(W.g_recv.ensure(rows * (size_t)Nmax * 8 * (size_t)R), S.msg);
            CUB_CUDA_CHECK(W.g_dense.ensure(rows * (size_t)Nglobal * 8), S.msg);
            if (N > 0) {
                CUB_CUDA_CHECK(cudaMemcpy2DAsync(W.g_send.p, (size_t)Nmax * 8, pool.err.p, (size_t)pool.cap * 8, (size_t)N * 8, (size_t)f,
                                                 cudaMemcpyDeviceToDevice, S.stream), S.msg);
                CUB_CUDA_CHECK(cudaMemcpyAsync(W.g_send.as<double>() + (size_t)f * Nmax, pool.errmax.p, (size_t)N * 8, cudaMemcpyDeviceToDevice,
                                               S.stream), S.msg);
            }
            rc = cub_comm_allgather(comm, W.g_send.p, rows * (size_t)Nmax * 8, W.g_recv.p, S.stream);
            if (rc != CUB_OK) return rc;
            {
                long long o = 0;  // drop the padding: dense [f + 1][Nglobal], rank-major region numbering
                for (int r = 0; r < R; ++r) {
                    if (Nr[(size_t)r] > 0)
                        CUB_CUDA_CHECK(cudaMemcpy2DAsync(W.g_dense.as<double>() + o, (size_t)Nglobal * 8,
                                                         W.g_recv.as<double>() + (size_t)r * rows * (size_t)Nmax, (size_t)Nmax * 8,
                                                         (size_t)Nr[(size_t)r] * 8, rows, cudaMemcpyDeviceToDevice, S.stream), S.msg);
                    o += Nr[(size_t)r];
                }
            }
            PoolView gv = pv;  // only err / errmax / cap are read by the selection kernels
            gv.err = W.g_dense.as<double>();
            gv.errmax = W.g_dense.as<double>() + (size_t)f * Nglobal;
            gv.cap = Nglobal;
            CUB_CUDA_CHECK(sel.ensure(std::max<long long>(Nglobal, pool.cap), f), S.msg);
            pk_sort_prepare(gv.errmax, Nglobal, sel.keysA.as<unsigned long long>(), sel.idxA.as<int>(), d_u64 + 4, S.stream);
            S.launches += 1;
            CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64 + 4, 16, cudaMemcpyDeviceToHost, S.stream), S.msg);
            S.d2h_bytes += 16;
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            const unsigned long long vary = h_u64[0] ^ h_u64[1];
            unsigned long long* ka = sel.keysA.as<unsigned long long>();
            unsigned long long* kb = sel.keysB.as<unsigned long long>();
            int* ia = sel.idxA.as<int>();
            int* ib = sel.idxB.as<int>();
            for (int shift = 0; shift < 64; shift += 8) {
                if (((vary >> shift) & 0xffULL) == 0) continue;
                pk_sort_pass(ka, ia, kb, ib, Nglobal, shift, sel.hist.as<unsigned>(), sel.bin_totals.as<unsigned long long>(),
                             sel.binbase.as<unsigned long long>(), S.stream);
                S.launches += 4;
                std::swap(ka, kb);
                std::swap(ia, ib);
            }
            if (all) {
                Kglobal = Kcap;  // maxEval truncation: the Kcap largest
            } else {
                pk_prefix(gv, f, ia, Nglobal, sel.prefix.as<double>(), sel.tile_tot.as<double>(), S.stream);
                pk_find_k(d_totals, sel.prefix.as<double>(), sel.tile_tot.as<double>(), Nglobal, f, absTol, relTol, norm, band, d_u64 + 4,
                          S.stream);
                S.launches += 3;
                CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64 + 4, 24, cudaMemcpyDeviceToHost, S.stream), S.msg);
                S.d2h_bytes += 24;
                CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
                long long k0 = (long long)h_u64[0], k1 = (long long)h_u64[1], k2 = (long long)h_u64[2];
                if (k0 > Nglobal) k0 = Nglobal;
                if (k1 > Nglobal) k1 = Nglobal;
                if (k2 > Nglobal) k2 = Nglobal;
                if (std::min(k1, Kcap) != std::min(k0, Kcap) || std::min(k2, Kcap) != std::min(k0, Kcap)) res->ambiguous += 1;
                Kglobal = std::min(k0, Kcap);
            }
            // this rank's part of the batch, in ranking order (the other sort buffer holds the list)
            CUB_CUDA_CHECK(W.g_send.ensure((size_t)8 * (size_t)(pk_filter_tiles(Kglobal) + 8)), S.msg);
            pk_filter_range(ia, Kglobal, off, off + N, W.g_send.as<unsigned long long>(), d_u64 + 4, ib, S.stream);
            S.launches += 3;
            CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64 + 4, 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
            S.d2h_bytes += 8;
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            K = (long long)h_u64[0];
            list = ib;
            if (Kglobal < 1 || Kglobal > Nglobal || K < 0 || K > N) {
                cub_set_message(S.msg, "internal: sharded vector selection returned K=%lld of N=%lld (global %lld of %lld)", K, N, Kglobal, Nglobal);
                return CUB_ERR_CUDA;
            }
        } else if (!(all && Kcap == Nglobal)) {
            // sort by errmax (descending, ties by ascending slot), then prefix sums and the residual test
            CUB_CUDA_CHECK(sel.ensure(pool.cap, f), S.msg);
            pk_sort_prepare(pv.errmax, N, sel.keysA.as<unsigned long long>(), sel.idxA.as<int>(), d_u64 + 4, S.stream);
            S.launches += 1;
            CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64 + 4, 16, cudaMemcpyDeviceToHost, S.stream), S.msg);
            S.d2h_bytes += 16;
            S.h2d_bytes += 16;
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            const unsigned long long vary = h_u64[0] ^ h_u64[1];  // bits that differ between keys
            unsigned long long* ka = sel.keysA.as<unsigned long long>();
            unsigned long long* kb = sel.keysB.as<unsigned long long>();
            int* ia = sel.idxA.as<int>();
            int* ib = sel.idxB.as<int>();
            for (int shift = 0; shift < 64; shift += 8) {
                if (((vary >> shift) & 0xffULL) == 0) continue;  // every key has the same digit here
                pk_sort_pass(ka, ia, kb, ib, N, shift, sel.hist.as<unsigned>(), sel.bin_totals.as<unsigned long long>(),
                             sel.binbase.as<unsigned long long>(), S.stream);
                S.launches += 4;
                std::swap(ka, kb);
                std::swap(ia, ib);
            }
            list = ia;
            if (all) {
                K = Kcap;  // maxEval truncation: the Kcap largest
            } else {
                pk_prefix(pv, f, ia, N, sel.prefix.as<double>(), sel.tile_tot.as<double>(), S.stream);
                pk_find_k(d_totals, sel.prefix.as<double>(), sel.tile_tot.as<double>(), N, f, absTol, relTol, norm, band, d_u64 + 4,
                          S.stream);
                S.launches += 3;
                CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64 + 4, 24, cudaMemcpyDeviceToHost, S.stream), S.msg);
                S.d2h_bytes += 24;
                S.h2d_bytes += 24;
                CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
                long long k0 = (long long)h_u64[0], k1 = (long long)h_u64[1], k2 = (long long)h_u64[2];
                if (k0 > N) k0 = N;  // never converged: pop until the heap is empty (Rule.java:133)
                if (k1 > N) k1 = N;
                if (k2 > N) k2 = N;
                if (std::min(k1, Kcap) != std::min(k0, Kcap) || std::min(k2, Kcap) != std::min(k0, Kcap)) res->ambiguous += 1;
                K = std::min(k0, Kcap);
            }
            Kglobal = K;
        }
        if (N + K > pool.cap) {
            long long nc = std::max<long long>(pool.cap * 2, N + K);
            cudaError_t e = pool.grow(nc, N, S.stream);
            if (e != cudaSuccess) {
                cudaGetLastError();
                cub_set_message(S.msg, "region pool exhausted at %lld regions (%s)", N, cudaGetErrorString(e));
                return CUB_ERR_POOL_EXHAUSTED;
            }
        }
        if (N + K > (long long)std::numeric_limits<int>::max()) {
            cub_set_message(S.msg, "region pool exceeds 2^31 slots");
            return CUB_ERR_POOL_EXHAUSTED;
        }
        if (list) ++n_sel_iters;
        t_select += ms_since(t_iter);
        t_iter = now();
        HostWork w{};
        w.list = list;
        w.n_items = 2 * K;
        w.base = N;
        w.mode = MODE_BISECT;
        rc = S.launch_rule(pool.view(), w);
        if (rc != CUB_OK) return rc;
        if (timing) cudaStreamSynchronize(S.stream);
        t_rule += ms_since(t_iter);
        N += K;
        Nglobal += Kglobal;
        numEval += 2 * P * Kglobal;
        S.local_regions_evaluated += 2 * K;
        ops += 3.0 * (double)Kglobal;
        res->iterations += 1;
        push_batch(trace, 2 * Kglobal);
        if (res->iterations >= kMaxIterations) {
            cub_set_message(S.msg, "stopped after %lld iterations without convergence", (long long)res->iterations);
            break;
        }
    }
    // final result: K5 over the pool (rank-ordered sum of the shards)
    pk_totals(pool.view(), f, N, d_partial.as<double>(), d_totals, S.stream);
    S.launches += 2;
    if (!sharded) {
        CUB_CUDA_CHECK(cudaMemcpyAsync(h_block, d_totals, (size_t)16 * f, cudaMemcpyDeviceToHost, S.stream), S.msg);
        S.d2h_bytes += 16 * f;
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
        for (int j = 0; j < 2 * f; ++j) out[j] = h_block[j];
    } else {
        rc = cub_comm_allgather(comm, d_totals, (size_t)16 * f, d_gather, S.stream);
        if (rc != CUB_OK) return rc;
        CUB_CUDA_CHECK(cudaMemcpyAsync(h_block, d_gather, (size_t)16 * f * R, cudaMemcpyDeviceToHost, S.stream), S.msg);
        S.d2h_bytes += 16 * f * R;
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
        for (int j = 0; j < 2 * f; ++j) {
            double s = 0.0;
            for (int r = 0; r < R; ++r) s += h_block[(size_t)r * 2 * f + j];
            out[j] = s;
        }
    }
    S.flush_rule_timer();
    if (timing)
        fprintf(stderr, "[cubature rank %d] sharded phase: %d iterations (%d with selection): shard %.2f ms, stats+exchange %.2f ms, select %.2f ms, rule %.2f ms\n",
                rank, n_iters, n_sel_iters, t_shard, t_stats, t_select, t_rule);
    res->converged = conv ? 1 : 0;
    res->num_eval = numEval;
    res->num_regions = Nglobal;
    fill_trace_regions(S, pool, nullptr, N, trace);
    return CUB_OK;
}

// Deals a replicated pool of N regions out to the R ranks: the regions are ranked by errmax (stable radix sort, identical on
// every rank) and rank r keeps ranks r, r + R, ... of that order -- every rank owns the same share of the regions that will
// be refined most.  Gathered into the free upper part of the pool, then moved down.  d_u64 / h_u64: 16 bytes of scratch.
int deal_pool_ranked(Session& S, long long& N, int R, int rank, unsigned long long* d_u64, unsigned long long* h_u64) {
    const int f = S.fdim;
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    SelectBuffers& sel = W.sel;
    const long long n_local = cubature_cuda_shard_count(N, R, rank);
    {
        if (N + n_local > pool.cap) {
            cudaError_t e = pool.grow(std::max<long long>(pool.cap * 2, N + n_local), N, S.stream);
            if (e != cudaSuccess) {
                cudaGetLastError();
                cub_set_message(S.msg, "region pool exhausted at %lld regions (%s)", N, cudaGetErrorString(e));
                return CUB_ERR_POOL_EXHAUSTED;
            }
        }
        CUB_CUDA_CHECK(sel.ensure(N, f), S.msg);
        unsigned long long* ka = sel.keysA.as<unsigned long long>();
        unsigned long long* kb = sel.keysB.as<unsigned long long>();
        int* ia = sel.idxA.as<int>();
        int* ib = sel.idxB.as<int>();
        pk_sort_prepare(pool.view().errmax, N, ka, ia, d_u64, S.stream);
        CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64, 16, cudaMemcpyDeviceToHost, S.stream), S.msg);
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
        const unsigned long long vary0 = h_u64[0] ^ h_u64[1];  // digits shared by every key need no pass
        S.launches += 1;
        for (int shift = 0; shift < 64; shift += 8) {
            if (((vary0 >> shift) & 0xffULL) == 0) continue;
            pk_sort_pass(ka, ia, kb, ib, N, shift, sel.hist.as<unsigned>(), sel.bin_totals.as<unsigned long long>(),
                         sel.binbase.as<unsigned long long>(), S.stream);
            std::swap(ka, kb);
            std::swap(ia, ib);
            S.launches += 4;
        }
        const PoolView src = pool.view();
        PoolView dst = src;  // same arrays, shifted by N slots
        dst.center += N;
        dst.halfw += N;
        dst.val += N;
        dst.err += N;
        dst.errmax += N;
        dst.splitdim += N;
        pk_shard_gather(src, dst, S.dim, f, n_local, R, rank, ia, S.stream);
        S.launches += 1;
        auto rows = [&](DevBuf& buf, int nrows, size_t elt) {
            return cudaMemcpy2DAsync(buf.p, (size_t)pool.cap * elt, (const unsigned char*)buf.p + (size_t)N * elt, (size_t)pool.cap * elt,
                                     (size_t)n_local * elt, nrows, cudaMemcpyDeviceToDevice, S.stream);
        };
        if (n_local > 0) {
            CUB_CUDA_CHECK(rows(pool.center, S.dim, 8), S.msg);
            CUB_CUDA_CHECK(rows(pool.halfw, S.dim, 8), S.msg);
            CUB_CUDA_CHECK(rows(pool.val, f, 8), S.msg);
            CUB_CUDA_CHECK(rows(pool.err, f, 8), S.msg);
            CUB_CUDA_CHECK(rows(pool.errmax, 1, 8), S.msg);
            CUB_CUDA_CHECK(rows(pool.splitdim, 1, 4), S.msg);
        }
        N = n_local;
    }
    return CUB_OK;
}

// ---- DEVICE mode, sharded pool, scalar integrand, NVLink peer exchange ---------------------------------------
// After the replicated phase every rank ranks the (identical) pool by errmax, keeps every R-th region and then
// runs the same device-driven loop as a single GPU: per iteration ONE cooperative kernel (k_iter_coop with R > 1:
// shard totals and histograms, exchanged with the other ranks through peer-mapped buffers inside the kernel,
// identical decisions on every rank, compaction of this rank's share of the batch) and the fused bisect + rule
// kernel on the rank's own regions, then one 200-byte read-back.  No NCCL call and no host-side exchange in the loop.
int integrate_sharded_peer(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                           LoopResult* res, const LoopState& st0) {
    const int f = S.fdim;  // == 1
    const int64_t P = S.P;
    cub_trace_t* trace = opt ? opt->trace : nullptr;
    cub_comm* comm = opt->comm;
    const int R = comm->nranks, rank = comm->rank;
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    SelectBuffers& sel = W.sel;
    const bool timing = getenv("CUBATURE_TIMING") != nullptr;
    const bool debug_peer = getenv("CUBATURE_DEBUG_PEER") != nullptr;
    auto t_begin = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };

    long long N = st0.N;
    long long Nglobal = st0.N;
    int64_t numEval = st0.numEval;
    double ops = st0.ops;
    const double eps = 2.220446049250313e-16;

    const int G = 1184;
    const size_t ctl_words = PK_CTL_WORDS + (size_t)4 * f;
    CUB_CUDA_CHECK(W.d_partial.ensure((size_t)8 * 2 * f * G + (size_t)8 * (5 * G + 8)), S.msg);
    CUB_CUDA_CHECK(W.d_small.ensure(ctl_words * 8 + 256 + 128), S.msg);
    CUB_CUDA_CHECK(W.h_small.ensure(ctl_words * 8 + 256), S.msg);
    double* d_part = W.d_partial.as<double>();
    unsigned long long* d_blk = (unsigned long long*)(d_part + (size_t)2 * f * G);
    long long* d_ctl = W.d_small.as<long long>();
    unsigned long long* d_u64 = (unsigned long long*)(d_ctl + ctl_words) + 8;
    unsigned long long* d_gblock = (unsigned long long*)((unsigned char*)W.d_small.p + ctl_words * 8 + 256);
    long long* h_ctl = W.h_small.as<long long>();
    const double* h_totals = (const double*)(h_ctl + PK_CTL_WORDS);
    unsigned long long* h_u64 = (unsigned long long*)(h_ctl + ctl_words) + 8;

    {
        const int rc_deal = deal_pool_ranked(S, N, R, rank, d_u64, h_u64);
        if (rc_deal != CUB_OK) return rc_deal;
    }
    const double t_shard = ms_since(t_begin);

    // One NCCL all-gather (outside the loop) lines the ranks up before the first in-kernel exchange: a rank that is still
    // compiling its integrand or growing its pool would otherwise be waited for inside a spinning kernel.
    {
        int rc0 = cub_comm_allgather(comm, d_u64, 8, d_u64 + 8, S.stream);
        if (rc0 != CUB_OK) return rc0;
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
    }
    PkShard shard;
    shard.R = R;
    shard.rank = rank;
    for (int r = 0; r < PK_MAX_RANKS; ++r) shard.xbuf[r] = r < R ? comm->xbuf[r] : nullptr;
    shard.gblock = d_gblock;
    CUB_CUDA_CHECK(sel.ensure_scalar(pool.cap), S.msg);
    CUB_CUDA_CHECK(cudaMemsetAsync(sel.small.p, 0, 512 + PK_COOP_HIST_BYTES, S.stream), S.msg);
    CUB_CUDA_CHECK(cudaMemsetAsync(W.d_small.p, 0, ctl_words * 8 + 256 + 128, S.stream), S.msg);
    int coop_parity = 0;
    bool conv = false, have_totals = false;
    int n_iters = 0;

    while (numEval < maxEval || maxEval == 0) {
        long long Kcap = Nglobal;  // pops allowed by maxEval (Rule.java:133), over all ranks
        if (maxEval > 0) {
            const int64_t room = maxEval - numEval;  // > 0 here
            Kcap = std::min<long long>(Nglobal, std::max<int64_t>(1, (room + 2 * P - 1) / (2 * P)));
        }
        const long long Klocal_max = std::min<long long>(N, Kcap);
        if (N + Klocal_max > pool.cap) {
            long long nc = std::max<long long>(pool.cap * 2, N + Klocal_max);
            if (st0.cap_hint > N + Klocal_max) nc = std::min(nc, st0.cap_hint);
            cudaError_t e = pool.grow(nc, N, S.stream);
            if (e != cudaSuccess) {
                cudaGetLastError();
                cub_set_message(S.msg, "region pool exhausted at %lld regions (%s)", N, cudaGetErrorString(e));
                return CUB_ERR_POOL_EXHAUSTED;
            }
            CUB_CUDA_CHECK(sel.ensure_scalar(pool.cap), S.msg);
        }
        if (N + Klocal_max > (long long)std::numeric_limits<int>::max()) {
            cub_set_message(S.msg, "region pool exceeds 2^31 slots");
            return CUB_ERR_POOL_EXHAUSTED;
        }
        const PoolView pv = pool.view();
        const double band = 4.0 * eps * std::sqrt(ops);
        SelScratch sc = sel.scratch(pool.cap);
        {
            shard.seq0 = comm->seq;
            cudaError_t e = pk_iter_coop(pv, f, N, Kcap, absTol, relTol, norm, band, 1, d_part, d_blk, d_ctl, sc.state,
                                         (unsigned char*)sc.state + 512, coop_parity++, sel.keysA.as<unsigned long long>(), sel.idxB.as<int>(),
                                         sel.idxA.as<int>(), S.n_sms, S.stream, &shard);
            if (e != cudaSuccess) {
                cub_set_message(S.msg, "cooperative launch failed: %s", cudaGetErrorString(e));
                return CUB_ERR_CUDA;
            }
        }
        S.launches += 1;
        const bool device_sized = N <= kDeviceSizedMaxRegions;
        HostWork w{};
        w.list = sel.idxA.as<int>();
        w.base = N;
        w.mode = MODE_BISECT;
        int rc;
        if (device_sized && N > 0) {
            w.dyn = d_ctl;
            w.n_items = 2 * Klocal_max;  // upper bound: sizes the grid only
            rc = S.launch_rule(pool.view(), w);
            if (rc != CUB_OK) return rc;
        }
        CUB_CUDA_CHECK(cudaMemcpyAsync(h_ctl, d_ctl, ctl_words * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
        S.d2h_bytes += (int64_t)ctl_words * 8;
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
        if (debug_peer)
            fprintf(stderr, "[peer rank %d] iter %d seq0 %llu exchanges %lld N %lld Nglobal %lld K %lld Kglobal %lld conv %lld all %lld totals %.17g %.17g disagreeing CTAs %lld\n", rank,
                    n_iters, comm->seq, h_ctl[13], N, h_ctl[12], h_ctl[3], h_ctl[7], h_ctl[4], h_ctl[5], h_totals[0], h_totals[1], h_ctl[14]);
        comm->seq += (unsigned long long)h_ctl[13];
        if (h_ctl[11]) {
            cub_set_message(S.msg, "peer exchange timed out: a rank of the communicator did not take part in the iteration");
            return CUB_ERR_NCCL;
        }
        res->ambiguous += h_ctl[6];
        if (h_ctl[14]) {
            cub_set_message(S.msg, "internal: %lld CTAs of the iteration kernel disagree on the selection state", h_ctl[14]);
            return CUB_ERR_CUDA;
        }
        Nglobal = h_ctl[12];
        ++n_iters;
        if (h_ctl[4]) {
            conv = true;
            have_totals = true;
            break;
        }
        const long long K = h_ctl[3], Kglobal = h_ctl[7];
        if (K < 0 || K > Klocal_max || Kglobal < 1 || Kglobal > Kcap) {
            cub_set_message(S.msg, "internal: sharded selection returned K=%lld of N=%lld (global %lld of %lld, cap %lld)", K, N, Kglobal,
                            Nglobal, Kcap);
            return CUB_ERR_CUDA;
        }
        if (!(device_sized && N > 0) && K >= 1) {
            if (h_ctl[5]) w.list = nullptr;  // pop-all: identity list, coalesced parent reads
            w.n_items = 2 * K;
            rc = S.launch_rule(pool.view(), w);
            if (rc != CUB_OK) return rc;
        }
        N += K;
        Nglobal += Kglobal;
        numEval += 2 * P * Kglobal;
        S.local_regions_evaluated += 2 * K;
        ops += 3.0 * (double)Kglobal;
        res->iterations += 1;
        push_batch(trace, 2 * Kglobal);
        if (res->iterations >= kMaxIterations) {
            cub_set_message(S.msg, "stopped after %lld iterations without convergence", (long long)res->iterations);
            break;
        }
    }
    if (!have_totals) {  // stopped by maxEval: K5 over the final shards, exchanged and added in rank order
        SelScratch sc = sel.scratch(pool.cap);
        shard.seq0 = comm->seq;
        cudaError_t e = pk_iter_coop(pool.view(), f, N, Nglobal, absTol, relTol, norm, 0.0, 0, d_part, d_blk, d_ctl, sc.state,
                                     (unsigned char*)sc.state + 512, coop_parity++, sel.keysA.as<unsigned long long>(), sel.idxB.as<int>(),
                                     sel.idxA.as<int>(), S.n_sms, S.stream, &shard);
        if (e != cudaSuccess) {
            cub_set_message(S.msg, "cooperative launch failed: %s", cudaGetErrorString(e));
            return CUB_ERR_CUDA;
        }
        S.launches += 1;
        CUB_CUDA_CHECK(cudaMemcpyAsync(h_ctl, d_ctl, ctl_words * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
        S.d2h_bytes += (int64_t)ctl_words * 8;
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
        if (debug_peer)
            fprintf(stderr, "[peer rank %d] final seq0 %llu exchanges %lld N %lld Nglobal %lld totals %.17g %.17g\n", rank, comm->seq, h_ctl[13], N,
                    h_ctl[12], h_totals[0], h_totals[1]);
        comm->seq += (unsigned long long)h_ctl[13];
        if (h_ctl[11]) {
            cub_set_message(S.msg, "peer exchange timed out: a rank of the communicator did not take part in the final sum");
            return CUB_ERR_NCCL;
        }
    }
    for (int j = 0; j < 2 * f; ++j) out[j] = h_totals[j];
    S.flush_rule_timer();
    if (timing)
        fprintf(stderr, "[cubature rank %d] sharded peer loop: %d iterations, deal %.2f ms, total %.2f ms, rule %.2f ms\n", rank, n_iters, t_shard,
                ms_since(t_begin), S.rule_ms);
    res->converged = conv ? 1 : 0;
    res->num_eval = numEval;
    res->num_regions = Nglobal;
    fill_trace_regions(S, pool, nullptr, N, trace);
    return CUB_OK;
}

// ---- DEVICE mode, one GPU: device-driven iterations -----------------------------------------------------
// Per iteration the host enqueues, without waiting in between: k_pool_stats (totals, extreme keys,
// convergence test, select-all shortcut -> device control block), the eight select passes + compaction
// (scalar integrands; they retire at once when the control block says "converged" or "pop everything"),
// and the fused bisect + rule kernel, which reads its work size from the control block (CubWork::dyn).
// Then ONE small device->host copy tells the host how many regions were popped.  Integrands with
// fdim > 1 use the same stats kernel and fall back to the sort-based selection when a batch has to be cut.
int integrate_device_fused(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                           LoopResult* res, long long stop_at_regions, LoopState* state_out, long long cap_override) {
    const int f = S.fdim;
    const int64_t P = S.P;
    cub_trace_t* trace = opt ? opt->trace : nullptr;
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    const size_t extra = 24 + (size_t)8 * f + 1;
    long long cap = cap_override > 0 ? cap_override : auto_capacity(S.dim, f, P, maxEval, opt ? opt->pool_capacity : 0, extra);
    CUB_CUDA_CHECK(pool.alloc(S.dim, f, cap, opt && opt->pool_capacity > 0), S.msg);
    int rc = set_root(S, pool);
    if (rc != CUB_OK) return rc;

    SelectBuffers& sel = W.sel;
    DevBuf &d_partial = W.d_partial, &d_small = W.d_small;
    PinnedBuf& h_small = W.h_small;
    const int G = 1184;
    const size_t ctl_words = PK_CTL_WORDS + (size_t)4 * f;
    CUB_CUDA_CHECK(d_partial.ensure((size_t)8 * 2 * f * G + (size_t)8 * (5 * G + 8)), S.msg);
    CUB_CUDA_CHECK(d_small.ensure(ctl_words * 8 + 256), S.msg);
    CUB_CUDA_CHECK(h_small.ensure(ctl_words * 8 + 256), S.msg);
    double* d_part = d_partial.as<double>();
    unsigned long long* d_blk = (unsigned long long*)(d_part + (size_t)2 * f * G);
    long long* d_ctl = d_small.as<long long>();
    unsigned long long* d_u64 = (unsigned long long*)(d_ctl + ctl_words) + 8;  // scratch for the fdim > 1 path
    double* d_totals = (double*)(d_ctl + PK_CTL_WORDS);
    long long* h_ctl = h_small.as<long long>();
    const double* h_totals = (const double*)(h_ctl + PK_CTL_WORDS);
    unsigned long long* h_u64 = (unsigned long long*)(h_ctl + ctl_words) + 8;
    CUB_CUDA_CHECK(cudaMemsetAsync(d_small.p, 0, ctl_words * 8 + 256, S.stream), S.msg);
    CUB_CUDA_CHECK(sel.ensure_scalar(pool.cap), S.msg);
    // both histogram sets of the cooperative kernel start out zero; launch k then clears the set of launch k + 1
    CUB_CUDA_CHECK(cudaMemsetAsync(sel.small.p, 0, 512 + PK_COOP_HIST_BYTES, S.stream), S.msg);
    int coop_parity = 0;

    {
        HostWork w{};
        w.n_items = 1;
        w.base = 0;
        w.mode = MODE_PLAIN;
        rc = S.launch_rule(pool.view(), w);
        if (rc != CUB_OK) return rc;
    }
    long long N = 1;
    int64_t numEval = P;
    bool conv = false, have_totals = false;
    double ops = 1.0;
    const double eps = 2.220446049250313e-16;
    const bool timing = getenv("CUBATURE_TIMING") != nullptr;
    double t_coop = 0, t_rule = 0, t_wait = 0, t_grow = 0;
    int n_grow = 0;
    auto tick = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
    };
    const auto t_loop = tick();

    while (numEval < maxEval || maxEval == 0) {
        if (stop_at_regions > 0 && N >= stop_at_regions) {  // multi-GPU: time to deal the pool out to the ranks
            state_out->N = N;
            state_out->numEval = numEval;
            state_out->ops = ops;
            state_out->cap_hint = cap;
            state_out->handoff = true;
            return CUB_OK;
        }
        long long Kcap = N;
        if (maxEval > 0) {
            const int64_t room = maxEval - numEval;  // > 0 here
            Kcap = std::min<long long>(N, std::max<int64_t>(1, (room + 2 * P - 1) / (2 * P)));
        }
        // the rule kernel is sized on the device: make room for the largest batch this iteration can pop
        if (N + Kcap > pool.cap) {
            const auto tg = tick();
            ++n_grow;
            long long nc = std::max<long long>(pool.cap * 2, N + Kcap);
            if (cap > N + Kcap) nc = std::min(nc, cap);  // never beyond what maxEval can need
            cudaError_t e = pool.grow(nc, N, S.stream);
            if (e != cudaSuccess) {
                cudaGetLastError();
                cub_set_message(S.msg, "region pool exhausted at %lld regions (%s)", N, cudaGetErrorString(e));
                return CUB_ERR_POOL_EXHAUSTED;
            }
            CUB_CUDA_CHECK(sel.ensure_scalar(pool.cap), S.msg);
            t_grow += ms_since(tg);
        }
        if (N + Kcap > (long long)std::numeric_limits<int>::max()) {
            cub_set_message(S.msg, "region pool exceeds 2^31 slots");
            return CUB_ERR_POOL_EXHAUSTED;
        }
        const PoolView pv = pool.view();
        const double band = 4.0 * eps * std::sqrt(ops);
        SelScratch sc = sel.scratch(pool.cap);
        sc.ctl = d_ctl;
        // ONE cooperative launch: totals, loop decisions and (scalar integrands) the whole batch selection
        {
            const auto tc = tick();
            cudaError_t e = pk_iter_coop(pv, f, N, Kcap, absTol, relTol, norm, band, f == 1 ? 1 : 0, d_part, d_blk, d_ctl, sc.state,
                                         (unsigned char*)sc.state + 512, coop_parity++, sel.keysA.as<unsigned long long>(), sel.idxB.as<int>(),
                                         sel.idxA.as<int>(), S.n_sms, S.stream);
            if (e != cudaSuccess) {
                cub_set_message(S.msg, "cooperative launch failed: %s", cudaGetErrorString(e));
                return CUB_ERR_CUDA;
            }
            t_coop += ms_since(tc);
        }
        S.launches += 1;
        long long K = 0;
        if (f == 1) {
            // small pools: the rule kernel is sized on the device (grid for the largest possible batch, surplus
            // CTAs retire at once) and the host learns K afterwards; large pools: one extra round trip buys an
            // exactly sized launch
            const bool device_sized = N <= kDeviceSizedMaxRegions;
            HostWork w{};
            w.list = sel.idxA.as<int>();
            w.base = N;
            w.mode = MODE_BISECT;
            if (device_sized) {
                const auto tr0 = tick();
                w.dyn = d_ctl;
                w.n_items = 2 * Kcap;  // upper bound: sizes the grid only
                rc = S.launch_rule(pool.view(), w);
                if (rc != CUB_OK) return rc;
                t_rule += ms_since(tr0);
            }
            const auto tw = tick();
            CUB_CUDA_CHECK(cudaMemcpyAsync(h_ctl, d_ctl, ctl_words * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
            S.d2h_bytes += (int64_t)ctl_words * 8;
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            t_wait += ms_since(tw);
            res->ambiguous += h_ctl[6];
            if (h_ctl[14]) {
                cub_set_message(S.msg, "internal: %lld CTAs of the iteration kernel disagree on the selection state", h_ctl[14]);
                return CUB_ERR_CUDA;
            }
            if (h_ctl[4]) {
                conv = true;
                have_totals = true;
                break;
            }
            K = h_ctl[3];
            if (!device_sized && K >= 1 && K <= Kcap) {
                if (h_ctl[5]) w.list = nullptr;  // pop-all: identity list, coalesced parent reads
                w.n_items = 2 * K;
                rc = S.launch_rule(pool.view(), w);
                if (rc != CUB_OK) return rc;
            }
            if (K < 1 || K > Kcap) {
                cub_set_message(S.msg, "internal: device selection returned K=%lld of N=%lld (cap %lld)", K, N, Kcap);
                return CUB_ERR_CUDA;
            }
        } else {
            CUB_CUDA_CHECK(cudaMemcpyAsync(h_ctl, d_ctl, ctl_words * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
            S.d2h_bytes += (int64_t)ctl_words * 8;
            CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
            res->ambiguous += h_ctl[6];
            if (h_ctl[4]) {
                conv = true;
                have_totals = true;
                break;
            }
            const int* list = nullptr;
            if (h_ctl[5]) {
                K = N;
            } else {
                // the batch has to be cut: stable sort by errmax, prefix sums, residual test for every k
                const bool all = (!cub_converged(f, VecAcc{h_totals, h_totals + 3 * f}, absTol, relTol, norm)) || N == 1;
                CUB_CUDA_CHECK(sel.ensure(pool.cap, f), S.msg);
                pk_sort_prepare(pv.errmax, N, sel.keysA.as<unsigned long long>(), sel.idxA.as<int>(), d_u64, S.stream);
                S.launches += 1;
                CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64, 16, cudaMemcpyDeviceToHost, S.stream), S.msg);
                S.d2h_bytes += 16;
                S.h2d_bytes += 16;
                CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
                const unsigned long long vary = h_u64[0] ^ h_u64[1];
                unsigned long long* ka = sel.keysA.as<unsigned long long>();
                unsigned long long* kb = sel.keysB.as<unsigned long long>();
                int* ia = sel.idxA.as<int>();
                int* ib = sel.idxB.as<int>();
                for (int shift = 0; shift < 64; shift += 8) {
                    if (((vary >> shift) & 0xffULL) == 0) continue;
                    pk_sort_pass(ka, ia, kb, ib, N, shift, sel.hist.as<unsigned>(), sel.bin_totals.as<unsigned long long>(),
                                 sel.binbase.as<unsigned long long>(), S.stream);
                    S.launches += 4;
                    std::swap(ka, kb);
                    std::swap(ia, ib);
                }
                list = ia;
                if (all) {
                    K = Kcap;
                } else {
                    pk_prefix(pv, f, ia, N, sel.prefix.as<double>(), sel.tile_tot.as<double>(), S.stream);
                    pk_find_k(d_totals, sel.prefix.as<double>(), sel.tile_tot.as<double>(), N, f, absTol, relTol, norm, band, d_u64, S.stream);
                    S.launches += 3;
                    CUB_CUDA_CHECK(cudaMemcpyAsync(h_u64, d_u64, 24, cudaMemcpyDeviceToHost, S.stream), S.msg);
                    S.d2h_bytes += 24;
                    S.h2d_bytes += 24;
                    CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
                    long long k0 = (long long)h_u64[0], k1 = (long long)h_u64[1], k2 = (long long)h_u64[2];
                    if (k0 > N) k0 = N;
                    if (k1 > N) k1 = N;
                    if (k2 > N) k2 = N;
                    if (std::min(k1, Kcap) != std::min(k0, Kcap) || std::min(k2, Kcap) != std::min(k0, Kcap)) res->ambiguous += 1;
                    K = std::min(k0, Kcap);
                }
            }
            HostWork w{};
            w.list = list;
            w.n_items = 2 * K;
            w.base = N;
            w.mode = MODE_BISECT;
            rc = S.launch_rule(pool.view(), w);
            if (rc != CUB_OK) return rc;
        }
        N += K;
        numEval += 2 * P * K;
        S.local_regions_evaluated += 2 * K;
        ops += 3.0 * (double)K;
        res->iterations += 1;
        push_batch(trace, 2 * K);
        if (res->iterations >= kMaxIterations) {
            cub_set_message(S.msg, "stopped after %lld iterations without convergence", (long long)res->iterations);
            break;
        }
    }
    if (!have_totals) {  // stopped by maxEval: K5 over the final pool
        SelScratch sc = sel.scratch(pool.cap);
        cudaError_t e = pk_iter_coop(pool.view(), f, N, N, absTol, relTol, norm, 0.0, 0, d_part, d_blk, d_ctl, sc.state,
                                     (unsigned char*)sc.state + 512, coop_parity++, sel.keysA.as<unsigned long long>(), sel.idxB.as<int>(),
                                     sel.idxA.as<int>(), S.n_sms, S.stream);
        if (e != cudaSuccess) {
            cub_set_message(S.msg, "cooperative launch failed: %s", cudaGetErrorString(e));
            return CUB_ERR_CUDA;
        }
        S.launches += 1;
        CUB_CUDA_CHECK(cudaMemcpyAsync(h_ctl, d_ctl, ctl_words * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
        S.d2h_bytes += (int64_t)ctl_words * 8;
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
    }
    for (int j = 0; j < 2 * f; ++j) out[j] = h_totals[j];
    S.flush_rule_timer();
    if (timing)
        fprintf(stderr, "[cubature] device loop: %lld iterations, %.2f ms: host time in coop launch %.2f, rule launch %.2f, read-back wait %.2f, %d pool growths %.2f ms; rule kernels %.2f ms\n",
                (long long)res->iterations, ms_since(t_loop), t_coop, t_rule, t_wait, n_grow, t_grow, S.rule_ms);
    res->converged = conv ? 1 : 0;
    res->num_eval = numEval;
    res->num_regions = N;
    fill_trace_regions(S, pool, nullptr, N, trace);
    return CUB_OK;
}

// ---- pool checkpoint (checkpoint.h): dump / upload of the structure-of-arrays pool ------------------------------------------
int checkpoint_write(Session& S, Pool& pool, long long n, int64_t num_eval, int64_t iterations, const char* path) {
    FILE* fp = fopen(path, "wb");
    if (!fp) {
        cub_set_message(S.msg, "cannot write the pool checkpoint %s", path);
        return CUB_ERR_BAD_ARGS;
    }
    CkptHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "CUBCKPT1", 8);
    h.version = 1;
    h.dim = S.dim;
    h.fdim = S.fdim;
    h.family = S.integ->family;
    h.n_regions = n;
    h.num_eval = num_eval;
    h.iterations = iterations;
    for (int d = 0; d < S.dim; ++d) {
        h.transform[d] = S.tr.code[d];
        h.shift[d] = S.tr.shift[d];
        h.tmin[d] = S.tr.tmin[d];
        h.tmax[d] = S.tr.tmax[d];
    }
    h.integrand_hash = ckpt_integrand_hash(S.integ);
    bool ok = fwrite(&h, sizeof h, 1, fp) == 1;  // placeholder: rewritten with the payload checksum
    std::vector<double> row((size_t)n);
    uint64_t sum = 0x452821e638d01377ULL;
    auto dump = [&](const DevBuf& buf, int nrows, size_t elt) {
        for (int r = 0; r < nrows && ok; ++r) {
            if (cudaMemcpy(row.data(), (const char*)buf.p + (size_t)r * pool.cap * elt, (size_t)n * elt, cudaMemcpyDeviceToHost) != cudaSuccess) {
                ok = false;
                break;
            }
            sum = ckpt_mix(sum, row.data(), (size_t)n * elt);
            ok = fwrite(row.data(), elt, (size_t)n, fp) == (size_t)n;
            S.d2h_bytes += (int64_t)n * (int64_t)elt;
        }
    };
    dump(pool.center, S.dim, 8);
    dump(pool.halfw, S.dim, 8);
    dump(pool.val, S.fdim, 8);
    dump(pool.err, S.fdim, 8);
    dump(pool.errmax, 1, 8);
    dump(pool.splitdim, 1, 4);
    h.payload_checksum = sum;
    ckpt_seal_header(&h);
    ok = ok && fseek(fp, 0, SEEK_SET) == 0 && fwrite(&h, sizeof h, 1, fp) == 1;
    ok = (fclose(fp) == 0) && ok;
    if (!ok) {
        cudaGetLastError();
        cub_set_message(S.msg, "writing the pool checkpoint %s failed", path);
        return CUB_ERR_CUDA;
    }
    return CUB_OK;
}

// Reads a checkpoint of this very problem (integrand, box, transform) into the pool (allocated for at least min_cap regions).
int checkpoint_load(Session& S, Pool& pool, const char* path, long long min_cap, bool exact_cap, CkptHeader* h) {
    FILE* fp = fopen(path, "rb");
    if (!fp) {
        cub_set_message(S.msg, "cannot read the pool checkpoint %s", path);
        return CUB_ERR_BAD_ARGS;
    }
    int rc = ckpt_read_header(fp, h, S.msg);
    if (rc == CUB_OK) {
        bool same = h->dim == S.dim && h->fdim == S.fdim && h->integrand_hash == ckpt_integrand_hash(S.integ);
        for (int d = 0; d < S.dim && same; ++d)
            same = h->transform[d] == S.tr.code[d] && h->shift[d] == S.tr.shift[d] && h->tmin[d] == S.tr.tmin[d] && h->tmax[d] == S.tr.tmax[d];
        if (!same) {
            cub_set_message(S.msg, "the pool checkpoint belongs to another problem (integrand, box or transform differ)");
            rc = CUB_ERR_BAD_ARGS;
        }
    }
    if (rc == CUB_OK && h->n_regions >= (int64_t)std::numeric_limits<int>::max() / 2) {
        cub_set_message(S.msg, "the pool checkpoint holds %lld regions: more than one GPU's slot range", (long long)h->n_regions);
        rc = CUB_ERR_POOL_EXHAUSTED;
    }
    if (rc != CUB_OK) {
        fclose(fp);
        return rc;
    }
    const long long n = h->n_regions;
    cudaError_t e = pool.alloc(S.dim, S.fdim, std::max<long long>(min_cap, n + 1), exact_cap);
    if (e == cudaSuccess && pool.cap < n + 1) e = pool.alloc(S.dim, S.fdim, n + 1, true);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fclose(fp);
        cub_set_message(S.msg, "region pool exhausted: the checkpoint's %lld regions do not fit (%s)", n, cudaGetErrorString(e));
        return CUB_ERR_POOL_EXHAUSTED;
    }
    std::vector<double> row((size_t)n);
    uint64_t sum = 0x452821e638d01377ULL;
    bool ok = true;
    auto load = [&](DevBuf& buf, int nrows, size_t elt) {
        for (int r = 0; r < nrows && ok; ++r) {
            ok = fread(row.data(), elt, (size_t)n, fp) == (size_t)n;
            if (!ok) break;
            sum = ckpt_mix(sum, row.data(), (size_t)n * elt);
            ok = cudaMemcpy((char*)buf.p + (size_t)r * pool.cap * elt, row.data(), (size_t)n * elt, cudaMemcpyHostToDevice) == cudaSuccess;
            S.h2d_bytes += (int64_t)n * (int64_t)elt;
        }
    };
    load(pool.center, S.dim, 8);
    load(pool.halfw, S.dim, 8);
    load(pool.val, S.fdim, 8);
    load(pool.err, S.fdim, 8);
    load(pool.errmax, 1, 8);
    load(pool.splitdim, 1, 4);
    fclose(fp);
    if (!ok || sum != h->payload_checksum) {
        cudaGetLastError();
        cub_set_message(S.msg, "the pool checkpoint %s is truncated or damaged", path);
        return CUB_ERR_BAD_ARGS;
    }
    return CUB_OK;
}

// ---- DEVICE mode, scalar integrands: device-resident loop on exact running sums (iter_exact.cu) -------------------------
// One iteration of Rule.cubature (Rule.java:72-142) = three launches with no host decision in between: k_exact_decide,
// k_exact_select, the bisect + rule kernel (sized by the control block).  The host enqueues `ahead` iterations, then reads
// the 256-byte control block once; launches queued behind the end of the loop (or a pause) retire at once.  Pauses:
// LS_GROW (the pool has to grow: host work), LS_HANDOFF (multi-GPU: time to deal the pool out to the ranks).
struct LoopCtl {  // the control block of a device-resident loop (iter_exact.h) as the host sees it
    long long* d_ctl = nullptr;
    long long* h_ctl = nullptr;            // pinned: the control block as the host last wrote / finally read it
    long long* h_snap[2] = {nullptr, nullptr};  // pinned: snapshots taken behind the batches in flight
    int64_t batches_logged = 0;
};
struct ExactBufs : LoopCtl {
    IterXParams p{};
    int max_select_blocks = 1, max_decide_blocks = 1;
};

int exact_setup(Session& S, ExactBufs& X, double relTol, double absTol, int norm) {
    Workspace& W = *S.ws;
    const int maxG = ix_select_max_blocks(S.n_sms);
    const size_t ctl_b = (size_t)(LC_WORDS + LC_LOG_CAP) * 8, acc_b = (size_t)XA_WORDS * 8, hset_b = (size_t)IX_HSET_WORDS * 8,
                 blk_b = ((size_t)3 * maxG + 8 + 7) / 8 * 64, gcand_b = ((size_t)PK_MAX_RANKS * IX_GCAP + 16 + IX_GCAP) * 8;
    static_assert(16 + IX_GCAP <= PK_XB_PAY, "a payload slot of the peer exchange must take a rank's candidates");
    CUB_CUDA_CHECK(W.d_exact.ensure(ctl_b + 2 * acc_b + hset_b + 512 + blk_b + gcand_b), S.msg);
    CUB_CUDA_CHECK(W.h_exact.ensure(3 * ctl_b), S.msg);
    unsigned char* base = W.d_exact.as<unsigned char>();
    CUB_CUDA_CHECK(cudaMemsetAsync(base, 0, ctl_b + 2 * acc_b + hset_b + 512 + blk_b, S.stream), S.msg);
    X.d_ctl = (long long*)base;
    X.h_ctl = W.h_exact.as<long long>();
    X.h_snap[0] = X.h_ctl + (LC_WORDS + LC_LOG_CAP);
    X.h_snap[1] = X.h_ctl + 2 * (LC_WORDS + LC_LOG_CAP);
    X.p.ctl = X.d_ctl;
    X.p.acc = (unsigned long long*)(base + ctl_b);
    X.p.gbins = (unsigned long long*)(base + ctl_b + acc_b);
    X.p.hset = (unsigned long long*)(base + ctl_b + 2 * acc_b);
    X.p.selx = (SelX*)(base + ctl_b + 2 * acc_b + hset_b);
    X.p.blk = (unsigned long long*)(base + ctl_b + 2 * acc_b + hset_b + 512);
    X.p.bar = X.p.blk + (size_t)3 * maxG + 4;
    X.p.gcand = (unsigned long long*)(base + ctl_b + 2 * acc_b + hset_b + 512 + blk_b);
    if (getenv("CUBATURE_GATHER_CANDIDATES") && getenv("CUBATURE_GATHER_CANDIDATES")[0] == '0') X.p.gcand = nullptr;  // A/B: one exchange per level
    X.p.absTol = absTol;
    X.p.relTol = relTol;
    X.p.norm = norm;
    X.p.R = 1;
    X.p.rank = 0;
    for (int r = 0; r < PK_MAX_RANKS; ++r) X.p.xbuf[r] = nullptr;
    X.max_select_blocks = maxG;
    X.max_decide_blocks = ix_decide_max_blocks(S.n_sms);
    CUB_CUDA_CHECK(cudaMemsetAsync(X.p.acc + XA_M_OFF + XA_M_MINKEY, 0xff, 8, S.stream), S.msg);
    return CUB_OK;
}

void exact_bind_pool(Session& S, ExactBufs& X) {
    Workspace& W = *S.ws;
    X.p.pool = W.pool.view();
    X.p.cand_key = W.sel.keysA.as<unsigned long long>();
    X.p.cand_idx = W.sel.idxB.as<int>();
    X.p.list = W.sel.idxA.as<int>();
    X.p.kept = W.sel.idxB.as<int>();  // (the candidates are dead when the compaction lists the kept regions)
    X.p.stash_val = W.sel.keysB.as<double>();
    X.p.stash_err = W.sel.keysA.as<double>();
}

// Runs iterations until the loop ends or pauses for a reason the caller handles (LS_HANDOFF); grows the pool on LS_GROW.
// On entry X.h_ctl holds the control block the host has just written to the device; on return the final one.
// The host never makes the device wait: while it looks at the control block as it was after batch i, batch i + 1 is already
// queued (two snapshot buffers, one event each).  A host thread that is descheduled for milliseconds -- seen on shared boxes
// as 10 - 40 ms added to single steps of a 118 ms job when every visit was a stream synchronisation -- then costs nothing;
// the price is at most two batches of launches that find the loop finished and retire at once.
// enqueue_iter(n_upper): queue one iteration (selection kernels + rule launch) for a pool of at most n_upper regions;
// rebind(): the pool has been re-allocated.
int loop_run(Session& S, LoopCtl& X, cub_trace_t* trace, long long cap_limit, const std::function<int(long long)>& enqueue_iter,
             const std::function<int()>& rebind, bool scalar_stamps) {
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    const bool timing = getenv("CUBATURE_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    static const int batch = [] { const char* e = getenv("CUBATURE_AHEAD"); const int v = e ? atoi(e) : 0; return v > 0 ? std::min(v, 40) : 8; }();
    int visits = 0, grows = 0;
    long long n_upper = 0;  // upper bound of this rank's region count (doubles at most per iteration)
    const size_t read_words = trace ? (size_t)(LC_WORDS + LC_LOG_CAP) : (size_t)LC_WORDS;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) CUB_CUDA_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming), S.msg);
    struct EvGuard {
        cudaEvent_t* e;
        ~EvGuard() {
            for (int i = 0; i < 2; ++i)
                if (e[i]) cudaEventDestroy(e[i]);
        }
    } guard{ev};
    auto log_batches = [&](const long long* ctl) {  // batch sizes since the last look (ring log)
        if (!trace) return;
        const int64_t nb = ctl[LC_NBATCH];
        for (int64_t i = std::max<int64_t>(X.batches_logged, nb - LC_LOG_CAP); i < nb; ++i) push_batch(trace, ctl[LC_WORDS + (i % LC_LOG_CAP)]);
        X.batches_logged = nb;
    };
    auto enqueue_batch = [&](int slot) -> int {
        for (int it = 0; it < batch; ++it) {
            cudaGetLastError();
            const int rc = enqueue_iter(n_upper);
            if (rc != CUB_OK) return rc;
            n_upper = std::min<long long>(2 * n_upper, pool.cap);
        }
        CUB_CUDA_CHECK(cudaMemcpyAsync(X.h_snap[slot], X.d_ctl, read_words * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
        CUB_CUDA_CHECK(cudaEventRecord(ev[slot], S.stream), S.msg);
        S.d2h_bytes += (int64_t)read_words * 8;
        return CUB_OK;
    };
    for (;;) {
        n_upper = X.h_ctl[LC_N];
        int cur = 0;
        int rc = enqueue_batch(0);
        if (rc != CUB_OK) return rc;
        for (;;) {
            rc = enqueue_batch(cur ^ 1);
            if (rc != CUB_OK) return rc;
            CUB_CUDA_CHECK(cudaEventSynchronize(ev[cur]), S.msg);
            ++visits;
            const long long* snap = X.h_snap[cur];
            log_batches(snap);
            if (snap[LC_STATE] != LS_RUN) break;
            // regions after batch `cur`; the batch queued behind it can at most double them per iteration
            long long nb = snap[LC_N];
            for (int it = 0; it < batch && nb < pool.cap; ++it) nb *= 2;
            n_upper = std::min<long long>(nb, pool.cap);
            cur ^= 1;
        }
        // the loop has stopped on the device: what is queued behind retires at once; the final control block
        CUB_CUDA_CHECK(cudaMemcpyAsync(X.h_ctl, X.d_ctl, read_words * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
        S.d2h_bytes += (int64_t)read_words * 8;
        log_batches(X.h_ctl);
        const long long state = X.h_ctl[LC_STATE];
        if (state == LS_GROW) {
            const long long N = X.h_ctl[LC_N];
            long long need = 2 * N;
            if (X.h_ctl[LC_MAXEVAL] > 0) {
                const long long room = X.h_ctl[LC_MAXEVAL] - X.h_ctl[LC_NUMEVAL];
                const long long kcap = std::max<long long>(1, (room + 2 * S.P - 1) / (2 * S.P));
                need = N + std::min(N, kcap);
            }
            long long nc = std::max<long long>(pool.cap * 2, need);
            if (cap_limit > need) nc = std::min(nc, cap_limit);  // never beyond what maxEval can need
            if (nc > (long long)std::numeric_limits<int>::max()) nc = (long long)std::numeric_limits<int>::max();
            if (nc < need) {
                cub_set_message(S.msg, "region pool exceeds 2^31 slots per GPU (%lld regions, %lld needed): shard the pool over more GPUs", N, need);
                return CUB_ERR_POOL_EXHAUSTED;
            }
            cudaError_t e = pool.grow(nc, N, S.stream);
            if (e != cudaSuccess) {
                cudaGetLastError();
                cub_set_message(S.msg, "region pool exhausted at %lld regions (%s)", N, cudaGetErrorString(e));
                return CUB_ERR_POOL_EXHAUSTED;
            }
            ++grows;
            {
                const int rc2 = rebind();
                if (rc2 != CUB_OK) return rc2;
            }
            X.h_ctl[LC_CAP] = pool.cap;
            X.h_ctl[LC_STATE] = LS_RUN;
            CUB_CUDA_CHECK(cudaMemcpyAsync(X.d_ctl + LC_CAP, X.h_ctl + LC_CAP, 8, cudaMemcpyHostToDevice, S.stream), S.msg);
            CUB_CUDA_CHECK(cudaMemcpyAsync(X.d_ctl + LC_STATE, X.h_ctl + LC_STATE, 8, cudaMemcpyHostToDevice, S.stream), S.msg);
            S.h2d_bytes += 16;
            continue;
        }
        break;
    }
    if (timing && scalar_stamps)
        fprintf(stderr, "[cubature] decide kernel of the last iteration (us): streaming + arrival %.1f, fold %.1f, exchange %.1f, decision %.1f\n",
                (X.h_ctl[LC_STAMP] - X.h_ctl[LC_STAMP + 4]) * 1e-3, (X.h_ctl[LC_STAMP + 1] - X.h_ctl[LC_STAMP]) * 1e-3,
                (X.h_ctl[LC_STAMP + 2] - X.h_ctl[LC_STAMP + 1]) * 1e-3, (X.h_ctl[LC_STAMP + 3] - X.h_ctl[LC_STAMP + 2]) * 1e-3);
    if (timing && scalar_stamps) {
        const long long* t = X.h_ctl + LC_STAMP + 5;
        fprintf(stderr, "[cubature] select kernel of the last cut (us, CTA 0): binade histogram %.1f, walk %.1f, levels", (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3);
        long long prev = t[2];
        for (int l = 0; l < 6; ++l)
            if (t[3 + l] > prev) {
                fprintf(stderr, " %.1f", (t[3 + l] - prev) * 1e-3);
                prev = t[3 + l];
            }
        fprintf(stderr, ", counts %.1f, compaction %.1f\n", (t[9] - prev) * 1e-3, (t[10] - t[9]) * 1e-3);
    }
    if (timing)
        fprintf(stderr, "[cubature] device-resident loop: %lld iterations (%lld cut a batch), %d host visits, %d pool growths, %.2f ms\n",
                (long long)X.h_ctl[LC_ITER], (long long)X.h_ctl[LC_NSELECT], visits, grows,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    return CUB_OK;
}

int exact_run(Session& S, ExactBufs& X, cub_trace_t* trace, long long cap_limit) {
    Pool& pool = S.ws->pool;
    auto enqueue_iter = [&](long long n_upper) -> int {
        // the decide kernel first folds what the previous iteration changed: at most 2 n_upper children (written by the
        // rule launch before it; n_upper already counts them) and half as many popped parents, 8 per thread
        long long Gd = (3 * n_upper / 2 + 8 * 256 - 1) / (8 * 256);
        Gd = std::max<long long>(1, std::min<long long>(Gd, X.max_decide_blocks));
        cudaError_t e = ix_decide(X.p, (int)Gd, S.n_sms, S.stream);
        if (e != cudaSuccess) {
            cub_set_message(S.msg, "decide kernel launch failed: %s", cudaGetErrorString(e));
            return CUB_ERR_CUDA;
        }
        long long G = (n_upper + 4095) / 4096;
        G = std::max<long long>(1, std::min<long long>(G, X.max_select_blocks));
        e = ix_select(X.p, (int)G, S.stream);
        if (e != cudaSuccess) {
            cub_set_message(S.msg, "select kernel launch failed (%lld CTAs of at most %d): %s", G, X.max_select_blocks, cudaGetErrorString(e));
            return CUB_ERR_CUDA;
        }
        S.launches += 2;
        HostWork w{};
        w.list = X.p.list;
        w.dyn = X.d_ctl;
        w.n_items = 2 * std::min<long long>(n_upper, pool.cap);  // upper bound: sizes the grid only
        w.base = 0;
        w.mode = MODE_BISECT;
        return S.launch_rule(pool.view(), w);
    };
    auto rebind = [&]() -> int {
        cudaError_t e = S.ws->sel.ensure_scalar(pool.cap);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cub_set_message(S.msg, "region pool exhausted at %lld regions (%s)", (long long)X.h_ctl[LC_N], cudaGetErrorString(e));
            return CUB_ERR_POOL_EXHAUSTED;
        }
        exact_bind_pool(S, X);
        return CUB_OK;
    };
    return loop_run(S, X, trace, cap_limit, enqueue_iter, rebind, true);
}

int exact_finish(Session& S, ExactBufs& X, double* out, LoopResult* res, cub_trace_t* trace) {
    const long long state = X.h_ctl[LC_STATE];
    if (state == LS_ERROR) {
        const long long why = X.h_ctl[LC_ERR];
        if (why == LE_PEER_TIMEOUT)
            cub_set_message(S.msg, "peer exchange timed out: a rank of the communicator did not take part in the iteration");
        else if (why == LE_PEER_ABORT)
            cub_set_message(S.msg, "peer exchange aborted: another rank of the communicator failed");
        else
            cub_set_message(S.msg, "internal: the running sums hold %lld regions, the pool %lld", (long long)X.h_ctl[LC_DBG0], (long long)X.h_ctl[LC_NGLOBAL]);
        return why == LE_INTERNAL ? CUB_ERR_CUDA : CUB_ERR_NCCL;
    }
    if (state == LS_GUARD) cub_set_message(S.msg, "stopped after %lld iterations without convergence", (long long)X.h_ctl[LC_ITER]);
    double v, e;
    memcpy(&v, &X.h_ctl[LC_V], 8);
    memcpy(&e, &X.h_ctl[LC_E], 8);
    out[0] = v;
    out[1] = e;
    S.flush_rule_timer();
    res->converged = state == LS_CONVERGED ? 1 : 0;
    res->num_eval = X.h_ctl[LC_NUMEVAL];
    res->num_regions = X.h_ctl[LC_NGLOBAL];
    res->iterations = X.h_ctl[LC_ITER];
    res->ambiguous = X.h_ctl[LC_AMBIG];
    S.local_regions_evaluated = 1 + X.h_ctl[LC_EVAL_LOCAL];
    fill_trace_regions(S, S.ws->pool, nullptr, X.h_ctl[LC_N], trace);
    return CUB_OK;
}

// One GPU (or the replicated phase of a multi-GPU run: stop_at_regions > 0 pauses the loop for the deal).
int integrate_device_exact(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                           LoopResult* res, ExactBufs& X, long long stop_at_regions, long long cap_override, bool* handoff) {
    const int f = S.fdim;  // == 1
    const int64_t P = S.P;
    cub_trace_t* trace = opt ? opt->trace : nullptr;
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    const bool timing = getenv("CUBATURE_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    const size_t extra = 24 + (size_t)8 * f + 1;
    const long long cap = cap_override > 0 ? cap_override : auto_capacity(S.dim, f, P, maxEval, opt ? opt->pool_capacity : 0, extra);
    const double t_cap = since();
    const char* ck_load = opt ? opt->checkpoint_load : nullptr;
    CkptHeader ck;
    int rc = CUB_OK;
    if (ck_load) {
        rc = checkpoint_load(S, pool, ck_load, cap, opt && opt->pool_capacity > 0, &ck);
        if (rc != CUB_OK) return rc;
    } else {
        CUB_CUDA_CHECK(pool.alloc(S.dim, f, cap, opt && opt->pool_capacity > 0), S.msg);
    }
    CUB_CUDA_CHECK(W.sel.ensure_scalar(pool.cap), S.msg);
    const double t_alloc = since();
    if (!ck_load) rc = set_root(S, pool);
    if (rc != CUB_OK) return rc;
    const double t_root = since();
    rc = exact_setup(S, X, relTol, absTol, norm);
    if (rc != CUB_OK) return rc;
    exact_bind_pool(S, X);
    if (timing) fprintf(stderr, "[cubature] start-up (host ms): capacity %.3f, buffers %.3f, root box %.3f, loop state %.3f\n", t_cap, t_alloc - t_cap, t_root - t_alloc, since() - t_root);
    if (!ck_load) {
        HostWork w{};
        w.n_items = 1;
        w.base = 0;
        w.mode = MODE_PLAIN;
        rc = S.launch_rule(pool.view(), w);  // Rule.java:64-69: the root region
        if (rc != CUB_OK) return rc;
    }
    long long* h = X.h_ctl;
    memset(h, 0, (size_t)LC_WORDS * 8);
    h[LC_DYN_N] = 1;  // the first decide kernel adds the root region (item 0 of an identity batch = slot 0) to the sums
    h[LC_N] = 1;
    h[LC_NGLOBAL] = 1;
    h[LC_NUMEVAL] = P;
    if (ck_load) {  // resume: the pool comes from the file, the exact sums are rebuilt from it, nothing is pending
        ix_rebuild(pool.view(), ck.n_regions, X.p.acc, S.n_sms, S.stream);
        S.launches += 1;
        h[LC_DYN_N] = 0;
        h[LC_N] = h[LC_NGLOBAL] = ck.n_regions;
        h[LC_NUMEVAL] = ck.num_eval;
        h[LC_ITER] = ck.iterations;
    }
    h[LC_CAP] = pool.cap;
    h[LC_MAXEVAL] = maxEval;
    h[LC_P] = P;
    {   // batches that keep at most this many sixteenths of the shard reset the sums and re-add the kept regions instead of
        // copying and subtracting the popped ones (CUBATURE_READD_16THS; 0 = never)
        static const long long readd_max = [] { const char* e = getenv("CUBATURE_READD_16THS"); return e ? atoll(e) : 10LL; }();
        h[LC_READD_MAX] = readd_max;
    }
    h[LC_STOP_AT] = stop_at_regions;
    h[LC_MAXITER] = kMaxIterations;
    h[LC_SEQ] = 1;
    CUB_CUDA_CHECK(cudaMemcpyAsync(X.d_ctl, h, (size_t)LC_WORDS * 8, cudaMemcpyHostToDevice, S.stream), S.msg);
    S.h2d_bytes += LC_WORDS * 8;
    const double t_run0 = since();
    rc = exact_run(S, X, trace, cap);
    if (rc != CUB_OK) return rc;
    if (X.h_ctl[LC_STATE] == LS_HANDOFF) {
        *handoff = true;
        return CUB_OK;
    }
    if (handoff) *handoff = false;
    const double t_run1 = since();
    rc = exact_finish(S, X, out, res, trace);
    if (rc == CUB_OK && opt && opt->checkpoint_save) rc = checkpoint_write(S, pool, X.h_ctl[LC_N], X.h_ctl[LC_NUMEVAL], X.h_ctl[LC_ITER], opt->checkpoint_save);
    if (timing) fprintf(stderr, "[cubature] root launch + control block %.3f ms, loop %.3f ms, finish %.3f ms (host clock)\n", t_run0 - t_root, t_run1 - t_run0, since() - t_run1);
    return rc;
}

// The replicated phase of a sharded run on the exact loop: complete result, or the state at the hand-off.
int replicated_phase_exact(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                           LoopResult* res, long long stop_at_regions, LoopState* st0, long long cap) {
    ExactBufs X;
    bool handoff = false;
    const int rc = integrate_device_exact(S, relTol, absTol, norm, maxEval, opt, out, res, X, stop_at_regions, cap, &handoff);
    if (rc == CUB_OK && handoff) {
        st0->N = X.h_ctl[LC_N];
        st0->numEval = X.h_ctl[LC_NUMEVAL];
        st0->ops = 1.0 + 3.0 * (double)(X.h_ctl[LC_N] - 1);
        st0->cap_hint = cap;
        st0->handoff = true;
        res->iterations = X.h_ctl[LC_ITER];
        res->ambiguous = X.h_ctl[LC_AMBIG];
        S.local_regions_evaluated = 1 + X.h_ctl[LC_EVAL_LOCAL];
        S.flush_rule_timer();
    }
    return rc;
}

// ---- DEVICE mode, vector integrands (fdim > 1): device-resident loop (iter_vec.cu) ------------------------------------------
// One iteration of Rule.cubature (Rule.java:72-142) = the cooperative iteration kernel (totals, convergence, batch cut on the
// per-component residual, compaction, bisection) + the rule launch it sizes; same control block, same enqueue-ahead host
// loop as the scalar path.
struct VecBufs : LoopCtl {
    IterVParams p{};
    int max_blocks = 1;
};

int vec_setup(Session& S, VecBufs& X, double relTol, double absTol, int norm) {
    Workspace& W = *S.ws;
    const int f = S.fdim;
    const int maxG = iv_max_blocks(S.n_sms, f);
    int gc = IV_GC_ROWS / f;
    gc = std::max(1, std::min(gc, maxG));
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t ctl_b = up((size_t)(LC_WORDS + LC_LOG_CAP) * 8), partial_b = up((size_t)2 * f * maxG * 8), totals_b = up((size_t)4 * f * 8),
                 blk_b = up(((size_t)4 * maxG + 8) * 8), bar_b = 256, bpart_b = up((size_t)gc * IV_NB * f * 8), bcpart_b = up((size_t)gc * IV_NB * 8),
                 bsum_b = up((size_t)IV_NB * f * 8), bcnt_b = up((size_t)IV_NB * 8), sab_b = up((size_t)2 * f * 8), chi_b = up((size_t)IV_CMAX * 8),
                 clo_b = up((size_t)IV_CMAX * 4), shi_b = chi_b, slo_b = clo_b, cres_b = up((size_t)IV_CMAX * f * 8), cbase_b = up((size_t)IV_CHUNKS * f * 8), st_b = 256,
                 wscr_b = up((size_t)maxG * (IV_THREADS / 32) * f * 8);
    const size_t total = ctl_b + partial_b + totals_b + blk_b + bar_b + bpart_b + bcpart_b + bsum_b + bcnt_b + sab_b + chi_b + clo_b + shi_b + slo_b + cres_b + cbase_b + st_b + wscr_b;
    CUB_CUDA_CHECK(W.d_vec.ensure(total), S.msg);
    CUB_CUDA_CHECK(W.h_exact.ensure(3 * (size_t)(LC_WORDS + LC_LOG_CAP) * 8 + (size_t)2 * f * 8), S.msg);
    unsigned char* q = W.d_vec.as<unsigned char>();
    // the control block, the barrier words and the selection state start from zero; the rest is written before it is read
    CUB_CUDA_CHECK(cudaMemsetAsync(q, 0, ctl_b, S.stream), S.msg);
    X.d_ctl = (long long*)q;
    q += ctl_b;
    X.p.partial = (double*)q;
    q += partial_b;
    X.p.totals = (double*)q;
    q += totals_b;
    X.p.blk = (unsigned long long*)q;
    q += blk_b;
    X.p.bar = (unsigned long long*)q;
    CUB_CUDA_CHECK(cudaMemsetAsync(q, 0, bar_b, S.stream), S.msg);
    q += bar_b;
    X.p.bsum_part = (double*)q;
    q += bpart_b;
    X.p.bcnt_part = (unsigned long long*)q;
    q += bcpart_b;
    X.p.bsum = (double*)q;
    q += bsum_b;
    X.p.bcnt = (unsigned long long*)q;
    q += bcnt_b;
    X.p.sabove = (double*)q;
    q += sab_b;
    X.p.cand_hi = (unsigned long long*)q;
    q += chi_b;
    X.p.cand_lo = (unsigned*)q;
    q += clo_b;
    X.p.sort_hi = (unsigned long long*)q;
    q += shi_b;
    X.p.sort_lo = (unsigned*)q;
    q += slo_b;
    X.p.cres = (double*)q;
    q += cres_b;
    X.p.cbase = (double*)q;
    q += cbase_b;
    X.p.state = (IvState*)q;
    CUB_CUDA_CHECK(cudaMemsetAsync(q, 0, st_b, S.stream), S.msg);
    q += st_b;
    X.p.wscratch = (double*)q;
    X.h_ctl = W.h_exact.as<long long>();
    X.h_snap[0] = X.h_ctl + (LC_WORDS + LC_LOG_CAP);
    X.h_snap[1] = X.h_ctl + 2 * (LC_WORDS + LC_LOG_CAP);
    X.p.ctl = X.d_ctl;
    X.p.dim = S.dim;
    X.p.fdim = f;
    X.p.need_errmax = (S.dim == 1 && S.integ->separable) ? 1 : 0;  // cub_gk_separable leaves errmax / splitdim to the loop
    X.p.gc_max = gc;
    X.p.absTol = absTol;
    X.p.relTol = relTol;
    X.p.norm = norm;
    X.max_blocks = maxG;
    return CUB_OK;
}

int vec_bind_pool(Session& S, VecBufs& X) {
    Workspace& W = *S.ws;
    cudaError_t e = W.sel.idxA.ensure((size_t)4 * W.pool.cap);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cub_set_message(S.msg, "region pool exhausted (%s)", cudaGetErrorString(e));
        return CUB_ERR_POOL_EXHAUSTED;
    }
    X.p.pool = W.pool.view();
    X.p.list = W.sel.idxA.as<int>();
    return CUB_OK;
}

int integrate_device_vec(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                         LoopResult* res, long long stop_at_regions, LoopState* state_out, long long cap_override) {
    const int f = S.fdim;
    const int64_t P = S.P;
    cub_trace_t* trace = opt ? opt->trace : nullptr;
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    const size_t extra = 4;
    const long long cap = cap_override > 0 ? cap_override : auto_capacity(S.dim, f, P, maxEval, opt ? opt->pool_capacity : 0, extra);
    const char* ck_load = opt ? opt->checkpoint_load : nullptr;
    CkptHeader ck;
    int rc = CUB_OK;
    if (ck_load) {
        rc = checkpoint_load(S, pool, ck_load, cap, opt && opt->pool_capacity > 0, &ck);
    } else {
        CUB_CUDA_CHECK(pool.alloc(S.dim, f, cap, opt && opt->pool_capacity > 0), S.msg);
        rc = set_root(S, pool);
    }
    if (rc != CUB_OK) return rc;
    VecBufs X;
    rc = vec_setup(S, X, relTol, absTol, norm);
    if (rc != CUB_OK) return rc;
    rc = vec_bind_pool(S, X);
    if (rc != CUB_OK) return rc;
    if (!ck_load) {
        HostWork w{};
        w.n_items = 1;
        w.base = 0;
        w.mode = MODE_PLAIN;
        rc = S.launch_rule(pool.view(), w, false);  // Rule.java:64-69: the root region
        if (rc != CUB_OK) return rc;
    }
    long long* h = X.h_ctl;
    memset(h, 0, (size_t)LC_WORDS * 8);
    h[LC_DYN_N] = 1;  // the first iteration kernel finds the root region as item 0 of an identity batch (slot 0)
    h[LC_N] = 1;
    h[LC_NGLOBAL] = 1;
    h[LC_NUMEVAL] = P;
    if (ck_load) {  // resume: the pool (errmax included) comes from the file
        h[LC_DYN_N] = 0;
        h[LC_N] = h[LC_NGLOBAL] = ck.n_regions;
        h[LC_NUMEVAL] = ck.num_eval;
        h[LC_ITER] = ck.iterations;
    }
    h[LC_CAP] = pool.cap;
    h[LC_MAXEVAL] = maxEval;
    h[LC_P] = P;
    h[LC_MAXITER] = kMaxIterations;
    h[LC_STOP_AT] = stop_at_regions;
    CUB_CUDA_CHECK(cudaMemcpyAsync(X.d_ctl, h, (size_t)LC_WORDS * 8, cudaMemcpyHostToDevice, S.stream), S.msg);
    S.h2d_bytes += LC_WORDS * 8;
    auto enqueue_iter = [&](long long n_upper) -> int {
        // one CTA per 1 024 regions, and enough warps for one row of the totals each (2 f rows, 8 warps per CTA)
        long long G = std::max<long long>((n_upper + 1023) / 1024, (2 * (long long)f + 7) / 8);
        G = std::max<long long>(1, std::min<long long>(G, X.max_blocks));
        cudaError_t e = iv_iter(X.p, (int)G, S.stream);
        if (e != cudaSuccess) {
            cub_set_message(S.msg, "iteration kernel launch failed (%lld CTAs of at most %d): %s", G, X.max_blocks, cudaGetErrorString(e));
            return CUB_ERR_CUDA;
        }
        S.launches += 1;
        HostWork w{};
        w.list = X.p.list;
        w.dyn = X.d_ctl;
        w.n_items = 2 * std::min<long long>(n_upper, pool.cap);  // upper bound: sizes the grid only
        w.base = 0;
        w.mode = MODE_CHILDREN;  // the iteration kernel has bisected the parents
        return S.launch_rule(pool.view(), w, false);
    };
    auto rebind = [&]() -> int { return vec_bind_pool(S, X); };
    rc = loop_run(S, X, trace, cap, enqueue_iter, rebind, false);
    if (rc != CUB_OK) return rc;
    if (getenv("CUBATURE_TIMING")) {
        const long long* t = X.h_ctl + LC_STAMP;
        auto us = [&](int a, int b2) { return t[b2] > t[a] ? (t[b2] - t[a]) * 1e-3 : 0.0; };
        fprintf(stderr, "[cubature] vector iteration kernel, last iteration / last cut (us, CTA 0): errmax + barrier %.1f, partials %.1f, totals %.1f, decisions %.1f; "
                        "last of %lld narrowing levels: sums %.1f, reduce %.1f, pick %.1f; gather %.1f, sort %.1f, residuals + predicate %.1f, counts %.1f, list %.1f, bisect %.1f\n",
                us(0, 1), us(1, 2), us(2, 3), us(3, 4), (long long)t[13], us(15, 5), us(5, 6), us(6, 7), t[13] ? us(7, 8) : us(4, 8), us(8, 14), us(14, 9), us(9, 10), us(10, 11), us(11, 12));
    }
    const long long state = X.h_ctl[LC_STATE];
    if (state == LS_HANDOFF && state_out) {  // multi-GPU: time to deal the pool out to the ranks (the caller goes on)
        state_out->N = X.h_ctl[LC_N];
        state_out->numEval = X.h_ctl[LC_NUMEVAL];
        state_out->ops = 1.0 + 3.0 * (double)(X.h_ctl[LC_N] - 1);
        state_out->cap_hint = cap;
        state_out->handoff = true;
        res->iterations = X.h_ctl[LC_ITER];
        res->ambiguous = X.h_ctl[LC_AMBIG];
        S.local_regions_evaluated = 1 + X.h_ctl[LC_EVAL_LOCAL];
        S.flush_rule_timer();
        return CUB_OK;
    }
    if (state == LS_ERROR) {
        cub_set_message(S.msg, "internal: the batch selection of the vector loop failed (code %lld)", (long long)X.h_ctl[LC_DBG0]);
        return CUB_ERR_CUDA;
    }
    if (state == LS_GUARD) cub_set_message(S.msg, "stopped after %lld iterations without convergence", (long long)X.h_ctl[LC_ITER]);
    // the totals of the last decision: val[f] then err[f]
    double* h_tot = (double*)(X.h_ctl + 3 * (LC_WORDS + LC_LOG_CAP));
    CUB_CUDA_CHECK(cudaMemcpyAsync(h_tot, X.p.totals, (size_t)2 * f * 8, cudaMemcpyDeviceToHost, S.stream), S.msg);
    CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
    S.d2h_bytes += (int64_t)2 * f * 8;
    for (int j = 0; j < 2 * f; ++j) out[j] = h_tot[j];
    S.flush_rule_timer();
    res->converged = state == LS_CONVERGED ? 1 : 0;
    res->num_eval = X.h_ctl[LC_NUMEVAL];
    res->num_regions = X.h_ctl[LC_NGLOBAL];
    res->iterations = X.h_ctl[LC_ITER];
    res->ambiguous = X.h_ctl[LC_AMBIG];
    S.local_regions_evaluated = 1 + X.h_ctl[LC_EVAL_LOCAL];
    fill_trace_regions(S, pool, nullptr, X.h_ctl[LC_N], trace);
    if (opt && opt->checkpoint_save) return checkpoint_write(S, pool, X.h_ctl[LC_N], X.h_ctl[LC_NUMEVAL], X.h_ctl[LC_ITER], opt->checkpoint_save);
    return CUB_OK;
}

// ---- periodic rebalancing of a sharded pool -----------------------------------------------------------
// The decide kernel has paused every rank at the same iteration (LS_REBALANCE) with every rank's region count in the control
// block.  Every rank derives the same plan from the counts: shards above their target (Ntot / R, the first Ntot % R ranks one
// more) give away the regions in their last slots, shards below it append what they are missing; the surplus regions form one
// stream in donor-rank order, the receivers take consecutive pieces of it in rank order.  Transport: NCCL all-gather of
// bounded SoA blocks (the only collective the communicator loads; a rebalance is rare and outside the loop).  Regions are
// self-contained (box, val, err, errmax, split dimension), children are addable to any owner: the partition, the exact sums
// and therefore every later decision are unchanged; only the (rank, slot) order of exactly tied regions can differ.
int rebalance_shards(Session& S, ExactBufs& X, cub_comm* comm, int R, int rank) {
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    const int f = S.fdim, dim = S.dim;
    std::vector<long long> cnt(R), target(R), surplus(R), deficit(R), sur_before(R + 1, 0), def_before(R + 1, 0);
    long long Ntot = 0, maxS = 0;
    for (int r = 0; r < R; ++r) Ntot += (cnt[r] = X.h_ctl[LC_RANKN + r]);
    for (int r = 0; r < R; ++r) {
        target[r] = Ntot / R + (r < Ntot % R ? 1 : 0);
        surplus[r] = std::max<long long>(0, cnt[r] - target[r]);
        deficit[r] = std::max<long long>(0, target[r] - cnt[r]);
        sur_before[r + 1] = sur_before[r] + surplus[r];
        def_before[r + 1] = def_before[r] + deficit[r];
        maxS = std::max(maxS, surplus[r]);
    }
    if (cnt[rank] != X.h_ctl[LC_N]) {
        cub_set_message(S.msg, "internal: rebalance: this rank holds %lld regions, the exchange says %lld", (long long)X.h_ctl[LC_N], cnt[rank]);
        return CUB_ERR_CUDA;
    }
    long long N = cnt[rank];
    const long long N_new = N - surplus[rank] + deficit[rank];
    if (N_new + 1 > pool.cap) {
        cudaError_t e = pool.grow(std::max<long long>(pool.cap + pool.cap / 4, N_new + 1), N, S.stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cub_set_message(S.msg, "region pool exhausted while rebalancing (%lld regions, %s)", N_new, cudaGetErrorString(e));
            return CUB_ERR_POOL_EXHAUSTED;
        }
    }
    const int drows = 2 * dim + 2 * f + 1;                     // double rows of a region
    const long long CH = std::min<long long>(maxS, 1 << 18);   // regions per rank and round
    const size_t blk = (size_t)CH * ((size_t)drows * 8 + 4);   // one rank's block: drows rows of CH doubles, then CH ints
    CUB_CUDA_CHECK(W.g_send.ensure(blk), S.msg);
    CUB_CUDA_CHECK(W.g_recv.ensure(blk * (size_t)R), S.msg);
    struct Row { DevBuf* buf; int nrows; size_t elt; };
    Row rows[6] = {{&pool.center, dim, 8}, {&pool.halfw, dim, 8}, {&pool.val, f, 8}, {&pool.err, f, 8}, {&pool.errmax, 1, 8}, {&pool.splitdim, 1, 4}};
    const long long my_tail = N - surplus[rank];  // first slot this rank gives away
    for (long long c0 = 0; c0 < maxS; c0 += CH) {
        const long long mine = std::max<long long>(0, std::min<long long>(CH, surplus[rank] - c0));  // elements of this round I contribute
        {   // pack: row-major block, row pitch CH
            size_t off = 0;
            for (const Row& rw : rows) {
                if (mine > 0)
                    CUB_CUDA_CHECK(cudaMemcpy2DAsync(W.g_send.as<unsigned char>() + off, (size_t)CH * rw.elt,
                                                     (const unsigned char*)rw.buf->p + (size_t)(my_tail + c0) * rw.elt, (size_t)pool.cap * rw.elt,
                                                     (size_t)mine * rw.elt, rw.nrows, cudaMemcpyDeviceToDevice, S.stream), S.msg);
                off += (size_t)CH * rw.elt * rw.nrows;
            }
        }
        int rc = cub_comm_allgather(comm, W.g_send.p, blk, W.g_recv.p, S.stream);
        if (rc != CUB_OK) return rc;
        if (deficit[rank] > 0) {
            // my piece of the surplus stream: [def_before[rank], def_before[rank] + deficit[rank]); this round carries the
            // elements [c0, c0 + CH) of every donor's tail = stream positions sur_before[d] + c0 ...
            const long long want_lo = def_before[rank], want_hi = want_lo + deficit[rank];
            for (int d = 0; d < R; ++d) {
                const long long have = std::max<long long>(0, std::min<long long>(CH, surplus[d] - c0));
                if (have <= 0) continue;
                const long long s_lo = sur_before[d] + c0, s_hi = s_lo + have;
                const long long lo = std::max(s_lo, want_lo), hi = std::min(s_hi, want_hi);
                if (lo >= hi) continue;
                const long long src_e = lo - s_lo;                       // element inside donor d's block
                const long long dst_slot = (N - surplus[rank]) + (lo - want_lo);  // receivers have no surplus: N + position in my piece
                size_t off = 0;
                for (const Row& rw : rows) {
                    CUB_CUDA_CHECK(cudaMemcpy2DAsync((unsigned char*)rw.buf->p + (size_t)dst_slot * rw.elt, (size_t)pool.cap * rw.elt,
                                                     W.g_recv.as<const unsigned char>() + (size_t)d * blk + off + (size_t)src_e * rw.elt, (size_t)CH * rw.elt,
                                                     (size_t)(hi - lo) * rw.elt, rw.nrows, cudaMemcpyDeviceToDevice, S.stream), S.msg);
                    off += (size_t)CH * rw.elt * rw.nrows;
                }
            }
        }
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);  // the send block is reused by the next round
    }
    S.launches += 0;
    CUB_CUDA_CHECK(W.sel.ensure_scalar(pool.cap), S.msg);
    exact_bind_pool(S, X);
    ix_rebuild(pool.view(), N_new, X.p.acc, S.n_sms, S.stream);
    S.launches += 1;
    X.h_ctl[LC_N] = N_new;
    X.h_ctl[LC_DYN_N] = 0;
    X.h_ctl[LC_SUB_K] = 0;
    X.h_ctl[LC_CAP] = pool.cap;
    X.h_ctl[LC_STATE] = LS_RUN;
    X.h_ctl[LC_REBAL_COUNT] += 1;
    CUB_CUDA_CHECK(cudaMemcpyAsync(X.d_ctl, X.h_ctl, (size_t)LC_WORDS * 8, cudaMemcpyHostToDevice, S.stream), S.msg);
    S.h2d_bytes += LC_WORDS * 8;
    if (getenv("CUBATURE_TIMING"))
        fprintf(stderr, "[cubature rank %d] rebalance %lld: %lld -> %lld regions (partition %lld, fullest shard %lld)\n", rank, (long long)X.h_ctl[LC_REBAL_COUNT], N, N_new,
                Ntot, *std::max_element(cnt.begin(), cnt.end()));
    return CUB_OK;
}

// Multi-GPU (section 8e), scalar integrands, NVLink peers: replicated phase on every rank, ranked deal, then the same
// device-resident loop on the rank's shard; the decide / select kernels exchange bins and histograms with the other ranks
// through peer-mapped buffers (iter_exact.cu), no NCCL call and no host decision inside the loop.
int integrate_sharded_exact(Session& S, double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* opt, double* out,
                            LoopResult* res) {
    const int f = S.fdim;  // == 1
    const int64_t P = S.P;
    cub_trace_t* trace = opt ? opt->trace : nullptr;
    cub_comm* comm = opt->comm;
    const int R = comm->nranks, rank = comm->rank;
    Workspace& W = *S.ws;
    Pool& pool = W.pool;
    if (comm->poisoned) {
        cub_set_message(S.msg, "communicator unusable: an earlier call on it failed inside a peer exchange; create a new one");
        return CUB_ERR_NCCL;
    }
    const size_t extra = 24 + (size_t)8 * f + 1;
    long long cap = auto_capacity(S.dim, f, P, maxEval, opt->pool_capacity, extra);
    const long long shard_min = opt->shard_min_regions > 0 ? opt->shard_min_regions : default_shard_min(R, maxEval, S.P);
    if (!(opt->pool_capacity > 0) && maxEval > 0) {
        // a rank holds about 1/R of the final pool (plus imbalance head-room), but at least the replicated phase
        const long long total = maxEval / (2 * P) + 8;
        cap = std::min(cap, std::max<long long>(total / R + total / (4 * R) + 1024, 4 * shard_min));
    }
    ExactBufs X;
    bool handoff = false;
    int rc = integrate_device_exact(S, relTol, absTol, norm, maxEval, opt, out, res, X, shard_min, cap, &handoff);
    if (rc != CUB_OK || !handoff) return rc;  // finished before sharding: every rank holds the complete result
    const auto t_begin = std::chrono::steady_clock::now();
    long long N = X.h_ctl[LC_N];
    CUB_CUDA_CHECK(W.d_small.ensure(256), S.msg);
    CUB_CUDA_CHECK(W.h_small.ensure(256), S.msg);
    unsigned long long* d_u64 = W.d_small.as<unsigned long long>();
    unsigned long long* h_u64 = W.h_small.as<unsigned long long>();
    rc = deal_pool_ranked(S, N, R, rank, d_u64, h_u64);
    if (rc == CUB_OK && W.sel.ensure_scalar(pool.cap) != cudaSuccess) {
        cudaGetLastError();
        cub_set_message(S.msg, "out of device memory for the selection buffers");
        rc = CUB_ERR_POOL_EXHAUSTED;
    }
    if (rc == CUB_OK) {
        exact_bind_pool(S, X);
        ix_rebuild(pool.view(), N, X.p.acc, S.n_sms, S.stream);  // running sums of this rank's shard
        S.launches += 1;
        // one NCCL all-gather (outside the loop) lines the ranks up before the first in-kernel exchange
        rc = cub_comm_allgather(comm, d_u64, 8, d_u64 + 8, S.stream);
        if (rc == CUB_OK && cudaStreamSynchronize(S.stream) != cudaSuccess) rc = CUB_ERR_CUDA;
    }
    if (rc == CUB_OK) {
        X.p.R = R;
        X.p.rank = rank;
        for (int r = 0; r < PK_MAX_RANKS; ++r) X.p.xbuf[r] = r < R ? (unsigned long long*)comm->xbuf[r] : nullptr;
        X.h_ctl[LC_N] = N;
        X.h_ctl[LC_DYN_N] = 0;  // the sums were rebuilt from the shard: nothing pending
        X.h_ctl[LC_SUB_K] = 0;
        X.h_ctl[LC_CAP] = pool.cap;
        X.h_ctl[LC_STOP_AT] = 0;
        X.h_ctl[LC_STATE] = LS_RUN;
        {   // rebalance trigger: partitions of at least 2^20 regions per rank (CUBATURE_REBALANCE_MIN overrides; 0 = never)
            const char* e = getenv("CUBATURE_REBALANCE_MIN");
            X.h_ctl[LC_REBAL_MIN] = e ? atoll(e) : (long long)R << 20;
        }
        X.h_ctl[LC_SEQ] = (long long)comm->seq;
        X.h_ctl[LC_DBG1] = 0;
        if (cudaMemcpyAsync(X.d_ctl, X.h_ctl, (size_t)LC_WORDS * 8, cudaMemcpyHostToDevice, S.stream) != cudaSuccess) rc = CUB_ERR_CUDA;
        S.h2d_bytes += LC_WORDS * 8;
    }
    const double t_deal = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    if (rc == CUB_OK) rc = exact_run(S, X, trace, cap);
    while (rc == CUB_OK && X.h_ctl[LC_STATE] == LS_REBALANCE) {
        rc = rebalance_shards(S, X, comm, R, rank);
        if (rc == CUB_OK) rc = exact_run(S, X, trace, cap);
    }
    res->rebalances = X.h_ctl[LC_REBAL_COUNT];
    if (rc == CUB_OK && X.h_ctl[LC_STATE] == LS_ERROR) {
        comm->poisoned = true;  // the ranks no longer agree on the exchange sequence
        return exact_finish(S, X, out, res, trace);
    }
    if (rc != CUB_OK) {
        // this rank gives up: release the peers that wait for it inside an exchange
        comm->poisoned = true;
        ix_abort_peers(R, comm->xbuf, S.stream);
        cudaStreamSynchronize(S.stream);
        return rc;
    }
    comm->seq = (unsigned long long)X.h_ctl[LC_SEQ];
    if (getenv("CUBATURE_TIMING"))
        fprintf(stderr, "[cubature rank %d] sharded: deal %.2f ms, %lld local of %lld regions\n", rank, t_deal, (long long)X.h_ctl[LC_N],
                (long long)X.h_ctl[LC_NGLOBAL]);
    return exact_finish(S, X, out, res, trace);
}

void init_stats(cub_stats_t* st) {
    if (!st) return;
    const int32_t sz = st->struct_size;
    memset(st, 0, sizeof *st);
    st->struct_size = sz ? sz : (int32_t)sizeof(cub_stats_t);
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------------
extern "C" {

int cubature_cuda_version(void) { return 1000 * CUBATURE_CUDA_ABI_VERSION + 0; }

const char* cubature_cuda_strerror(int status) {
    switch (status) {
        case CUB_OK: return "ok";
        case CUB_ERR_BAD_ARGS: return "bad arguments (xmin/xmax/out null or empty, sizes inconsistent)";
        case CUB_ERR_BAD_TOL: return "either relTol or absTol or both have to be not NaN";
        case CUB_ERR_BAD_DIM: return "unsupported dimension";
        case CUB_ERR_BAD_NORM: return "unknown error norm";
        case CUB_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case CUB_ERR_CUDA: return "CUDA runtime error";
        case CUB_ERR_NVRTC: return "NVRTC compilation of the integrand failed";
        case CUB_ERR_POOL_EXHAUSTED: return "region pool exhausted";
        case CUB_ERR_NCCL: return "NCCL error";
        case CUB_ERR_UNSUPPORTED: return "unsupported configuration";
        case CUB_ERR_BAD_FAMILY: return "unknown integrand family or wrong parameter count";
    }
    return "unknown status";
}

int64_t cubature_cuda_points_per_region(int dim) { return cub_points_per_region(dim); }

int cubature_cuda_converged(int fdim, const double* val, const double* err, double absTol, double relTol, int norm) {
    if (fdim < 1 || !val || !err) return CUB_ERR_BAD_ARGS;
    if (std::isnan(relTol) && std::isnan(absTol)) return CUB_ERR_BAD_TOL;
    if ((norm & ~8) < 0 || (norm & ~8) > 4 || ((norm & 8) && (norm & 7) > 1)) return CUB_ERR_BAD_NORM;
    return cub_converged(fdim, VecAcc{val, err}, absTol, relTol, norm) ? 1 : 0;
}

int cubature_cuda_rule_points(int dim, const double* c, const double* h, double* pts) {
    if (!c || !h || !pts) return CUB_ERR_BAD_ARGS;
    if (dim < 1 || dim > CUB_MAX_DIM) return CUB_ERR_BAD_DIM;
    auto put = [&](int64_t k, const std::vector<double>& x) {
        for (int d = 0; d < dim; ++d) pts[k * dim + d] = x[(size_t)d];
    };
    std::vector<double> x(c, c + dim);
    int64_t k = 0;
    if (dim == 1) {  // RuleGaussKronrod_1d.java:60-84: centre, Gauss nodes 1,3,5, then Kronrod nodes 0,2,4,6
        static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                      0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                      0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                      0.207784955007898467600689403773245, 0.0};
        pts[k++] = c[0];
        for (int j = 1; j < 7; j += 2) {
            const double w = h[0] * xgk[j];
            pts[k++] = c[0] - w;
            pts[k++] = c[0] + w;
        }
        for (int j = 0; j < 8; j += 2) {
            const double w = h[0] * xgk[j];
            pts[k++] = c[0] - w;
            pts[k++] = c[0] + w;
        }
        return CUB_OK;
    }
    const double l2 = 0.3585685828003180919906451539079374954541, l4 = 0.9486832980505137995996680633298155601160,
                 l5 = 0.6882472016116852977216287342936235251269;
    put(k++, x);
    for (int i = 0; i < dim; ++i) {  // Rule75GenzMalik.java:307-325
        const double r2 = h[i] * l2, r4 = h[i] * l4;
        x[(size_t)i] = c[i] - r2; put(k++, x);
        x[(size_t)i] = c[i] + r2; put(k++, x);
        x[(size_t)i] = c[i] - r4; put(k++, x);
        x[(size_t)i] = c[i] + r4; put(k++, x);
        x[(size_t)i] = c[i];
    }
    for (int i = 0; i < dim - 1; ++i) {  // Rule75GenzMalik.java:274-296
        for (int j = i + 1; j < dim; ++j) {
            const double ri = h[i] * l4, rj = h[j] * l4;
            x[(size_t)i] = c[i] - ri; x[(size_t)j] = c[j] - rj; put(k++, x);
            x[(size_t)i] = c[i] + ri; put(k++, x);
            x[(size_t)j] = c[j] + rj; put(k++, x);
            x[(size_t)i] = c[i] - ri; put(k++, x);
            x[(size_t)j] = c[j];
        }
        x[(size_t)i] = c[i];
    }
    for (unsigned n = 0; n < (1u << dim); ++n) {  // Gray-code order, Rule75GenzMalik.java:237-269
        const unsigned g = n ^ (n >> 1);
        for (int d = 0; d < dim; ++d) {
            const double r = h[d] * l5;
            x[(size_t)d] = ((g >> d) & 1u) ? c[d] - r : c[d] + r;
        }
        put(k++, x);
    }
    return CUB_OK;
}

int cubature_cuda_integrand_builtin(int family_id, int dim, int fdim, const double* params, size_t nparams, cub_integrand_t** out) {
    if (!out) return CUB_ERR_BAD_ARGS;
    *out = nullptr;
    if (dim < 1 || dim > CUB_MAX_DIM) return CUB_ERR_BAD_DIM;
    if (fdim < 1) return CUB_ERR_BAD_ARGS;
    if (nparams > 0 && !params) return CUB_ERR_BAD_ARGS;
    size_t need = 0;
    bool separable = false;
    int need_fdim = 0;
    switch (family_id) {
        case CUB_FAMILY_GENZ_OSCILLATORY:
        case CUB_FAMILY_GENZ_PRODUCT_PEAK:
        case CUB_FAMILY_GENZ_CORNER_PEAK:
        case CUB_FAMILY_GENZ_GAUSSIAN:
        case CUB_FAMILY_GENZ_C0:
        case CUB_FAMILY_GENZ_DISCONTINUOUS:
            need = 2 * (size_t)dim;
            need_fdim = 1;
            break;
        case CUB_FAMILY_GAUSS_ISO:
            need = 1;
            need_fdim = 1;
            break;
        case CUB_FAMILY_REFTEST:
            need = (size_t)fdim;
            if (fdim > CUB_REFTEST_MAX_FDIM) return CUB_ERR_UNSUPPORTED;
            break;
        case CUB_FAMILY_OSC_VEC:
            need = 0;
            separable = true;
            break;
        case CUB_FAMILY_CEXP:
            need = (size_t)(fdim / 2);
            break;
        default:
            return CUB_ERR_BAD_FAMILY;
    }
    if (nparams != need) return CUB_ERR_BAD_FAMILY;
    if (need_fdim && fdim != need_fdim) return CUB_ERR_BAD_FAMILY;
    if (nparams > CUB_MAX_PARAMS_BY_VALUE) return CUB_ERR_UNSUPPORTED;
    cub_integrand* in = new cub_integrand();
    in->family = family_id;
    in->dim = dim;
    in->fdim = fdim;
    in->separable = separable;
    if (nparams) in->params.assign(params, params + nparams);
    *out = in;
    return CUB_OK;
}

int cubature_cuda_integrand_nvrtc(const char* cuda_src, const char* entry, int dim, int fdim, const double* params, size_t nparams,
                                  cub_integrand_t** out, char* log, size_t log_cap) {
    if (log && log_cap) log[0] = 0;
    if (!out || !cuda_src || !entry || !*entry) return CUB_ERR_BAD_ARGS;
    *out = nullptr;
    if (dim < 1 || dim > CUB_MAX_DIM) return CUB_ERR_BAD_DIM;
    if (fdim < 1) return CUB_ERR_BAD_ARGS;
    if (nparams > 0 && !params) return CUB_ERR_BAD_ARGS;
    if (nparams > CUB_MAX_PARAMS_BY_VALUE) return CUB_ERR_UNSUPPORTED;
    std::unique_ptr<cub_integrand> in(new cub_integrand());
    in->family = CUB_FAM_USER;
    in->dim = dim;
    in->fdim = fdim;
    in->user_src = cuda_src;
    in->entry = entry;
    if (nparams) in->params.assign(params, params + nparams);
    std::string l;
    int rc = cub_jit_compile(in.get(), 0, &l);  // compile now so that source errors surface here
    if (log && log_cap) {
        strncpy(log, l.c_str(), log_cap - 1);
        log[log_cap - 1] = 0;
    }
    if (rc != CUB_OK) return rc;
    *out = in.release();
    return CUB_OK;
}

void cubature_cuda_integrand_free(cub_integrand_t* integrand) {
    if (!integrand) return;
    for (int v = 0; v < 2; ++v)
        for (auto& kv : integrand->loaded[v])
            if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
    cudaGetLastError();
    delete integrand;
}

int64_t cubature_cuda_integrand_cubin(cub_integrand_t* integrand, int has_transform, void* buf, size_t buf_cap) {
    if (!integrand) return CUB_ERR_BAD_ARGS;
    std::lock_guard<std::mutex> lock(integrand->mu);
    std::string log;
    int rc = cub_jit_compile(integrand, has_transform, &log);
    if (rc != CUB_OK) {
        fprintf(stderr, "cubature_cuda: %s\n", log.c_str());
        return rc;
    }
    const std::vector<char>& c = integrand->cubin[has_transform ? 1 : 0];
    if (buf && buf_cap >= c.size()) memcpy(buf, c.data(), c.size());
    return (int64_t)c.size();
}

}  // extern "C"
namespace {
int integrate_impl(cub_integrand_t* integrand, int dim, const double* xmin, const double* xmax, const int* transform,
                            double relTol, double absTol, int norm, int64_t maxEval, const cub_options_t* options, double* out,
                            cub_stats_t* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    init_stats(stats);
    // argument checks mirror the reference's exceptions (Cubature.java:119-128, Rule.java:164-167)
    if (!integrand || !xmin || !xmax || !out) return CUB_ERR_BAD_ARGS;
    if (dim < 1 || dim != integrand->dim) return dim < 1 ? CUB_ERR_BAD_ARGS : CUB_ERR_BAD_DIM;
    if (dim > CUB_MAX_DIM) return CUB_ERR_BAD_DIM;
    if (std::isnan(relTol) && std::isnan(absTol)) return CUB_ERR_BAD_TOL;
    if (norm < 0 || norm > 4) return CUB_ERR_BAD_NORM;
    if (maxEval < 0) return CUB_ERR_BAD_ARGS;
    TransformSpec tr;
    int rc = make_transform(dim, xmin, xmax, transform, &tr);
    if (rc != CUB_OK) return rc;
    const int select = options ? options->select : CUB_SELECT_DEVICE;
    if (select != CUB_SELECT_DEVICE && select != CUB_SELECT_HEAP_EXACT) return CUB_ERR_BAD_ARGS;
    if (options && options->strict_all && norm <= CUB_NORM_PAIRED) norm |= CUB_NORM_STRICT;  // section 8f-3 extension
    if (options && options->trace) {
        options->trace->n_batches = 0;
        options->trace->n_regions = 0;
    }
    const auto t_call = std::chrono::steady_clock::now();
    auto call_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
    Session S;
    rc = S.open(integrand, dim, options ? options->device : -1, tr, stats);
    if (rc != CUB_OK) return rc;
    const double t_open = call_ms();
    LoopResult res;
    cudaEventRecord(S.ev0, S.stream);
    const bool wants_ckpt = options && (options->checkpoint_save || options->checkpoint_load);
    if (wants_ckpt && (select == CUB_SELECT_HEAP_EXACT || (options->comm && options->comm->nranks > 1) || legacy_loop())) {
        cub_set_message(S.msg, "pool checkpoints need the device-resident loop on one GPU (select = DEVICE, no communicator)");
        return CUB_ERR_UNSUPPORTED;
    }
    if (select == CUB_SELECT_HEAP_EXACT)
        rc = integrate_heap_exact(S, relTol, absTol, norm, maxEval, options, out, &res);
    else if (options && options->comm && options->comm->nranks > 1) {
        const char* env = getenv("CUBATURE_PEER_LOOP");
        if (S.fdim == 1 && options->comm->peer_ok && !legacy_loop() && !(env && env[0] == '0'))
            rc = integrate_sharded_exact(S, relTol, absTol, norm, maxEval, options, out, &res);
        else
            rc = integrate_device(S, relTol, absTol, norm, maxEval, options, out, &res);
    }
    else if (S.fdim == 1 && !legacy_loop()) {
        ExactBufs X;
        bool handoff = false;
        rc = integrate_device_exact(S, relTol, absTol, norm, maxEval, options, out, &res, X, 0, 0, &handoff);
    } else if (S.fdim > 1 && !legacy_loop())
        rc = integrate_device_vec(S, relTol, absTol, norm, maxEval, options, out, &res);
    else
        rc = integrate_device_fused(S, relTol, absTol, norm, maxEval, options, out, &res);
    const double t_body = call_ms();
    cudaEventRecord(S.ev1, S.stream);
    cudaEventSynchronize(S.ev1);
    if (getenv("CUBATURE_TIMING")) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, S.ev0, S.ev1);
        fprintf(stderr, "[cubature] call (host ms): session %.3f, body %.3f, closing event %.3f; device events %.3f ms\n", t_open, t_body - t_open,
                call_ms() - t_body, ms);
    }
    if (stats) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, S.ev0, S.ev1) == cudaSuccess) stats->device_ms = ms;
        stats->converged = res.converged;
        stats->num_eval = res.num_eval;
        stats->num_regions = res.num_regions;
        stats->iterations = res.iterations;
        stats->regions_evaluated = select == CUB_SELECT_DEVICE ? S.local_regions_evaluated : (S.P > 0 ? res.num_eval / S.P : 0);
        stats->ambiguous_decisions = res.ambiguous;
        stats->rebalances = res.rebalances;
        stats->gpu_launches = S.launches;
        stats->h2d_bytes = S.h2d_bytes;
        stats->d2h_bytes = S.d2h_bytes;
        stats->rule_ms = S.rule_ms;
        stats->jit_ms = integrand->jit_ms_pending;
        integrand->jit_ms_pending = 0.0;
        stats->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    cudaGetLastError();
    return rc;
}

// struct_size versioning of the two ABI structs: the library works on full-size private copies.  Of the caller's options
// only the first struct_size bytes are read (fields beyond them keep their defaults); of the statistics only the first
// struct_size bytes are written back.
struct StatsBridge {
    cub_stats_t local;
    cub_stats_t* user;
    int32_t user_size;
    bool ok;
    explicit StatsBridge(cub_stats_t* u) : user(u), user_size(u ? u->struct_size : 0), ok(true) {
        memset(&local, 0, sizeof local);
        local.struct_size = (int32_t)sizeof local;
        if (u && user_size < (int32_t)(2 * sizeof(int32_t))) ok = false;  // cannot even hold struct_size and the converged flag
    }
    cub_stats_t* get() { return user && ok ? &local : nullptr; }
    ~StatsBridge() {
        if (!user || !ok) return;
        const size_t n = std::min<size_t>((size_t)user_size, sizeof local);
        memcpy(user, &local, n);
        user->struct_size = user_size;
    }
};
}  // namespace

extern "C" {

int cubature_cuda_integrate(cub_integrand_t* integrand, int dim, const double* xmin, const double* xmax, const int* transform, double relTol,
                            double absTol, int norm, int64_t maxEval, const cub_options_t* options, double* out, cub_stats_t* stats) {
    StatsBridge sb(stats);
    if (!sb.ok) return CUB_ERR_BAD_ARGS;
    cub_options_t opt;
    memset(&opt, 0, sizeof opt);
    opt.struct_size = (int32_t)sizeof opt;
    opt.select = CUB_SELECT_DEVICE;
    opt.device = -1;
    if (options) {
        const int32_t sz = options->struct_size;
        if (sz < (int32_t)(2 * sizeof(int32_t))) {
            if (sb.get()) cub_set_message(sb.get()->message, "cub_options_t.struct_size = %d: set it to sizeof(cub_options_t)", (int)sz);
            return CUB_ERR_BAD_ARGS;
        }
        memcpy(&opt, options, std::min<size_t>((size_t)sz, sizeof opt));
        opt.struct_size = (int32_t)sizeof opt;
    }
    return integrate_impl(integrand, dim, xmin, xmax, transform, relTol, absTol, norm, maxEval, options ? &opt : nullptr, out, sb.get());
}

int cubature_cuda_eval_regions(cub_integrand_t* integrand, int dim, const int* transform, const double* shift, int64_t nR,
                               const double* centers, const double* halfwidths, double* val, double* err, int32_t* split_dim,
                               cub_stats_t* stats_user) {
    StatsBridge sb(stats_user);
    if (!sb.ok) return CUB_ERR_BAD_ARGS;
    cub_stats_t* stats = sb.get();
    if (!integrand || !centers || !halfwidths || !val || !err || !split_dim || nR < 0) return CUB_ERR_BAD_ARGS;
    if (dim != integrand->dim) return CUB_ERR_BAD_DIM;
    if (nR == 0) return CUB_OK;
    TransformSpec tr;
    tr.any = false;
    for (int d = 0; d < dim; ++d) {
        tr.code[d] = transform ? transform[d] : 0;
        tr.shift[d] = shift ? shift[d] : 0.0;
        tr.tmin[d] = tr.tmax[d] = 0.0;
        if (tr.code[d]) tr.any = true;
    }
    Session S;
    int rc = S.open(integrand, dim, -1, tr, stats);
    if (rc != CUB_OK) return rc;
    const int f = S.fdim;
    Pool& pool = S.ws->pool;
    CUB_CUDA_CHECK(pool.alloc(dim, f, nR, true), S.msg);
    pool.cap = nR;  // dense rows: the host copies below assume stride nR (the cached buffers may be larger)
    // [nR][dim] -> [dim][nR]
    std::vector<double> tmp((size_t)nR * std::max(dim, f));
    for (int pass = 0; pass < 2; ++pass) {
        const double* src = pass ? halfwidths : centers;
        for (int64_t i = 0; i < nR; ++i)
            for (int d = 0; d < dim; ++d) tmp[(size_t)d * nR + i] = src[i * dim + d];
        CUB_CUDA_CHECK(cudaMemcpy(pass ? pool.halfw.p : pool.center.p, tmp.data(), (size_t)8 * dim * nR, cudaMemcpyHostToDevice), S.msg);
    }
    HostWork w{};
    w.n_items = nR;
    w.base = 0;
    w.mode = MODE_PLAIN;
    cudaEventRecord(S.ev0, S.stream);
    rc = S.launch_rule(pool.view(), w);
    cudaEventRecord(S.ev1, S.stream);
    if (rc != CUB_OK) return rc;
    CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
    for (int pass = 0; pass < 2; ++pass) {
        CUB_CUDA_CHECK(cudaMemcpy(tmp.data(), pass ? pool.err.p : pool.val.p, (size_t)8 * f * nR, cudaMemcpyDeviceToHost), S.msg);
        double* dst = pass ? err : val;
        for (int64_t i = 0; i < nR; ++i)
            for (int j = 0; j < f; ++j) dst[i * f + j] = tmp[(size_t)j * nR + i];
    }
    CUB_CUDA_CHECK(cudaMemcpy(split_dim, pool.splitdim.p, (size_t)4 * nR, cudaMemcpyDeviceToHost), S.msg);
    S.flush_rule_timer();
    if (stats) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, S.ev0, S.ev1) == cudaSuccess) stats->device_ms = ms;
        stats->rule_ms = S.rule_ms;
        stats->gpu_launches = S.launches;
        stats->regions_evaluated = nR;
        stats->num_eval = nR * S.P;
        stats->jit_ms = integrand->jit_ms_pending;
        integrand->jit_ms_pending = 0.0;
    }
    return CUB_OK;
}

int cubature_cuda_eval_points(cub_integrand_t* integrand, int dim, int64_t n, const double* x, double* fout) {
    if (!integrand || !x || !fout || n < 0) return CUB_ERR_BAD_ARGS;
    if (dim != integrand->dim) return CUB_ERR_BAD_DIM;
    if (n == 0) return CUB_OK;
    TransformSpec tr;
    tr.any = false;
    for (int d = 0; d < dim; ++d) {
        tr.code[d] = 0;
        tr.shift[d] = tr.tmin[d] = tr.tmax[d] = 0.0;
    }
    Session S;
    int rc = S.open(integrand, dim, -1, tr, nullptr);
    if (rc != CUB_OK) {
        fprintf(stderr, "cubature_cuda_eval_points: %s\n", S.msgbuf);
        return rc;
    }
    DevBuf dx, df;
    CUB_CUDA_CHECK(dx.ensure((size_t)8 * dim * n), S.msg);
    CUB_CUDA_CHECK(df.ensure((size_t)8 * S.fdim * n), S.msg);
    CUB_CUDA_CHECK(cudaMemcpy(dx.p, x, (size_t)8 * dim * n, cudaMemcpyHostToDevice), S.msg);
    const double* xp = dx.as<double>();
    double* fp = df.as<double>();
    long long nn = n;
    void* args[5] = {&xp, &fp, &nn, S.prm_arg.data(), S.tr_arg.data()};
    CUB_CUDA_CHECK(cudaLaunchKernel((const void*)S.kern->eval_points, dim3((unsigned)((n + 127) / 128)), dim3(128), args, 0, S.stream), S.msg);
    CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
    CUB_CUDA_CHECK(cudaMemcpy(fout, df.p, (size_t)8 * S.fdim * n, cudaMemcpyDeviceToHost), S.msg);
    return CUB_OK;
}

int cubature_cuda_bench_rule(cub_integrand_t* integrand, int dim, const double* xmin, const double* xmax, int64_t n_regions,
                             int repeats, double* ms_per_launch, double* checksum) {
    if (!integrand || !xmin || !xmax || n_regions < 1 || repeats < 1 || !ms_per_launch) return CUB_ERR_BAD_ARGS;
    if (dim != integrand->dim) return CUB_ERR_BAD_DIM;
    TransformSpec tr;
    int rc = make_transform(dim, xmin, xmax, nullptr, &tr);
    if (rc != CUB_OK) return rc;
    Session S;
    rc = S.open(integrand, dim, -1, tr, nullptr);
    if (rc != CUB_OK) {
        fprintf(stderr, "cubature_cuda_bench_rule: %s\n", S.msgbuf);
        return rc;
    }
    const int f = S.fdim;
    Pool& pool = S.ws->pool;
    // CUBATURE_BENCH_MODE=bisect: time the fused bisect + rule launch the adaptive loop uses (n_regions parents ->
    // 2 n_regions children per launch, every launch halves the boxes again) instead of the plain evaluation
    const char* bm = getenv("CUBATURE_BENCH_MODE");
    const bool bisect = bm && strcmp(bm, "bisect") == 0 && f == 1 && dim >= 2;
    CUB_CUDA_CHECK(pool.alloc(dim, f, bisect ? 2 * n_regions : n_regions, true), S.msg);
    // uniform slabs along dimension 0: region i covers [xmin0 + i*w, xmin0 + (i+1)*w] x the full box elsewhere
    {
        std::vector<double> row((size_t)n_regions);
        for (int d = 0; d < dim; ++d) {
            const double lo = tr.tmin[d], hi = tr.tmax[d];
            for (int64_t i = 0; i < n_regions; ++i) {
                if (d == 0) {
                    const double w = (hi - lo) / (double)n_regions;
                    row[(size_t)i] = lo + (i + 0.5) * w;
                } else {
                    row[(size_t)i] = 0.5 * (lo + hi);
                }
            }
            CUB_CUDA_CHECK(cudaMemcpy(pool.center.as<double>() + (size_t)d * pool.cap, row.data(), (size_t)8 * n_regions, cudaMemcpyHostToDevice), S.msg);
            for (int64_t i = 0; i < n_regions; ++i) row[(size_t)i] = d == 0 ? 0.5 * (hi - lo) / (double)n_regions : 0.5 * (hi - lo);
            CUB_CUDA_CHECK(cudaMemcpy(pool.halfw.as<double>() + (size_t)d * pool.cap, row.data(), (size_t)8 * n_regions, cudaMemcpyHostToDevice), S.msg);
        }
    }
    HostWork w{};
    w.n_items = n_regions;
    w.mode = MODE_PLAIN;
    rc = S.launch_rule(pool.view(), w);  // warm-up
    if (rc != CUB_OK) return rc;
    CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
    S.flush_rule_timer();
    S.rule_ms = 0.0;
    if (bisect) {
        w.n_items = 2 * n_regions;
        w.base = n_regions;
        w.mode = MODE_BISECT;
    }
    for (int r = 0; r < repeats; ++r) {
        rc = S.launch_rule(pool.view(), w);
        if (rc != CUB_OK) return rc;
    }
    CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
    S.flush_rule_timer();
    *ms_per_launch = S.rule_ms / repeats / (bisect ? 2.0 : 1.0);  // per n_regions evaluated
    if (checksum) {
        DevBuf part, tot;
        CUB_CUDA_CHECK(part.ensure((size_t)8 * 2 * f * 1184), S.msg);
        CUB_CUDA_CHECK(tot.ensure((size_t)16 * f), S.msg);
        pk_totals(pool.view(), f, n_regions, part.as<double>(), tot.as<double>(), S.stream);
        std::vector<double> h((size_t)2 * f);
        CUB_CUDA_CHECK(cudaMemcpyAsync(h.data(), tot.p, (size_t)16 * f, cudaMemcpyDeviceToHost, S.stream), S.msg);
        CUB_CUDA_CHECK(cudaStreamSynchronize(S.stream), S.msg);
        *checksum = h[0];
    }
    return CUB_OK;
}

void cubature_cuda_release_workspace(void) {
    std::lock_guard<std::mutex> lock(g_ws_mu);
    for (size_t i = 0; i < g_ws.size();) {
        if (!g_ws[i]->busy) {
            cudaSetDevice(g_ws[i]->device);
            g_ws.erase(g_ws.begin() + (long)i);
        } else {
            ++i;
        }
    }
    cudaGetLastError();
}

int cubature_cuda_fp64_peak(int device, double* tflops, double* ms_out) {
    if (!tflops) return CUB_ERR_BAD_ARGS;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return CUB_ERR_NO_DEVICE;
    }
    char msg[256];
    if (device >= 0) CUB_CUDA_CHECK(cudaSetDevice(device), msg);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t s;
    CUB_CUDA_CHECK(cudaStreamCreate(&s), msg);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    DevBuf out;
    CUB_CUDA_CHECK(out.ensure(64), msg);
    const int blocks = sms * 8, iters = 1 << 16;
    pk_dfma_peak(out.as<double>(), blocks, 1024, s);  // warm-up
    cudaStreamSynchronize(s);
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a, s);
        pk_dfma_peak(out.as<double>(), blocks, iters, s);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaStreamDestroy(s);
    return cudaGetLastError() == cudaSuccess ? CUB_OK : CUB_ERR_CUDA;
}

// The same probe run back to back for `window_ms`: a B200 at its 1000 W power cap lowers the SM clock under a sustained
// FP64 load, so a rule kernel timed inside a long step is bounded by the SUSTAINED rate (the second half of the window),
// not by the burst rate of a 10 ms probe.
int cubature_cuda_fp64_peak_sustained(int device, double window_ms, double* tflops_burst, double* tflops_sustained) {
    if (!tflops_burst || !tflops_sustained || !(window_ms > 0)) return CUB_ERR_BAD_ARGS;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return CUB_ERR_NO_DEVICE;
    }
    char msg[256];
    if (device >= 0) CUB_CUDA_CHECK(cudaSetDevice(device), msg);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t s;
    CUB_CUDA_CHECK(cudaStreamCreate(&s), msg);
    DevBuf out;
    CUB_CUDA_CHECK(out.ensure(64), msg);
    const int blocks = sms * 8, iters = 1 << 15;  // about 5 ms per launch
    const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    const int max_launches = 2048;
    std::vector<cudaEvent_t> ev;
    pk_dfma_peak(out.as<double>(), blocks, 1024, s);  // warm-up
    cudaStreamSynchronize(s);
    cudaEvent_t e0;
    cudaEventCreate(&e0);
    cudaEventRecord(e0, s);
    ev.push_back(e0);
    const auto t0 = std::chrono::steady_clock::now();
    int n = 0;
    while (n < max_launches) {
        pk_dfma_peak(out.as<double>(), blocks, iters, s);
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.push_back(e);
        ++n;
        if (n % 4 == 0) {  // keep at most a few launches queued ahead of the clock
            cudaEventSynchronize(ev[(size_t)n - 2]);
            if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() >= window_ms) break;
        }
    }
    cudaStreamSynchronize(s);
    double best = 1e30;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[(size_t)i], ev[(size_t)i + 1]);
        if (ms < best) best = ms;
    }
    float half = 0.f;
    cudaEventElapsedTime(&half, ev[(size_t)n / 2], ev[(size_t)n]);
    *tflops_burst = flops / (best * 1e-3) / 1e12;
    *tflops_sustained = flops * (double)(n - n / 2) / (half * 1e-3) / 1e12;
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    cudaStreamDestroy(s);
    return cudaGetLastError() == cudaSuccess ? CUB_OK : CUB_ERR_CUDA;
}

}  // extern "C"
