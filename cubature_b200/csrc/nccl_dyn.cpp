// nccl_dyn.cpp -- NCCL bound at run time with dlopen("libnccl.so.2").
//
// Why dlopen: (1) libcubature_cuda.so must load on machines without NCCL (single-GPU use, CPU-only
// symbol checks); (2) inside a PyTorch process the already-loaded torch-bundled libnccl is reused
// (same soname), so there is exactly one NCCL in the address space.
// One cub_comm_t = one rank (one GPU) of a communicator; ranks live in separate processes
// (torchrun) or separate threads of one host process (the Java host of BASELINE.json).
#include <dlfcn.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "cub_internal.h"

namespace {
struct NcclId {
    char internal[128];
};
struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};
Nccl* nccl() {
    static Nccl api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("CUBATURE_NCCL_PATH");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = "cannot dlopen libnccl.so.2 (set CUBATURE_NCCL_PATH)";
            return;
        }
#define LOAD(field, sym)                                       \
    *(void**)(&api.field) = dlsym(api.handle, sym);           \
    if (!api.field) {                                          \
        api.error = std::string("libnccl lacks symbol ") + sym; \
        return;                                                \
    }
        LOAD(GetUniqueId, "ncclGetUniqueId")
        LOAD(CommInitRank, "ncclCommInitRank")
        LOAD(CommDestroy, "ncclCommDestroy")
        LOAD(AllReduce, "ncclAllReduce")
        LOAD(AllGather, "ncclAllGather")
        LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    });
    return &api;
}
}  // namespace

// In-process communicator: the ranks are threads of one host process (the single-process Java host of
// BASELINE.json, or the multi-rank tests on one GPU).  The exchange goes through host memory with two
// barriers; no kernel ever waits for another rank.
struct cub_local_group {
    std::mutex mu;
    std::condition_variable cv;
    int nranks = 0, arrived = 0, refs = 0;
    unsigned long long generation = 0;
    std::vector<std::vector<char>> slots;
    void barrier() {
        std::unique_lock<std::mutex> lock(mu);
        const unsigned long long gen = generation;
        if (++arrived == nranks) {
            arrived = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lock, [&] { return generation != gen; });
        }
    }
};
namespace {
std::mutex g_groups_mu;
std::map<unsigned long long, cub_local_group*> g_groups;
}  // namespace

// The one collective the driver needs: every rank contributes `bytes` and receives nranks*bytes in rank
// order.  Reductions are done afterwards by a device kernel that adds the blocks in rank order, so the
// result is bit-identical on every rank and independent of NCCL's choice of algorithm.
int cub_comm_allgather(cub_comm* comm, const void* dev_send, size_t bytes, void* dev_recv, cudaStream_t stream) {
    if (!comm || comm->nranks == 1) {
        return cudaMemcpyAsync(dev_recv, dev_send, bytes, cudaMemcpyDeviceToDevice, stream) == cudaSuccess ? CUB_OK : CUB_ERR_CUDA;
    }
    if (comm->local) {
        cub_local_group* g = comm->local;
        std::vector<char>& mine = g->slots[(size_t)comm->rank];
        mine.resize(bytes);
        if (cudaMemcpyAsync(mine.data(), dev_send, bytes, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return CUB_ERR_CUDA;
        if (cudaStreamSynchronize(stream) != cudaSuccess) return CUB_ERR_CUDA;
        g->barrier();
        for (int r = 0; r < comm->nranks; ++r) {
            if (g->slots[(size_t)r].size() != bytes) return CUB_ERR_NCCL;
            if (cudaMemcpyAsync((char*)dev_recv + (size_t)r * bytes, g->slots[(size_t)r].data(), bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess)
                return CUB_ERR_CUDA;
        }
        if (cudaStreamSynchronize(stream) != cudaSuccess) return CUB_ERR_CUDA;
        g->barrier();
        return CUB_OK;
    }
    Nccl* n = nccl();
    const int rc = n->AllGather(dev_send, dev_recv, bytes, /*ncclChar*/ 0, comm->nccl_comm, stream);
    if (rc != 0) {
        fprintf(stderr, "cubature_cuda: ncclAllGather: %s\n", n->GetErrorString(rc));
        return CUB_ERR_NCCL;
    }
    return CUB_OK;
}

extern "C" {

int cubature_cuda_comm_unique_id(void* id_out) {
    if (!id_out) return CUB_ERR_BAD_ARGS;
    Nccl* n = nccl();
    if (!n->error.empty()) {
        fprintf(stderr, "cubature_cuda: %s\n", n->error.c_str());
        return CUB_ERR_NCCL;
    }
    NcclId id;
    if (n->GetUniqueId(&id) != 0) return CUB_ERR_NCCL;
    memcpy(id_out, &id, sizeof id);
    return CUB_OK;
}

int cubature_cuda_comm_init(const void* id, int nranks, int rank, int device, cub_comm_t** out) {
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return CUB_ERR_BAD_ARGS;
    *out = nullptr;
    cub_comm* c = new cub_comm();
    c->nranks = nranks;
    c->rank = rank;
    if (device >= 0) {
        if (cudaSetDevice(device) != cudaSuccess) {
            cudaGetLastError();
            delete c;
            return CUB_ERR_NO_DEVICE;
        }
        c->device = device;
    } else {
        cudaGetDevice(&c->device);
    }
    if (nranks > 1) {
        if (!id) {
            delete c;
            return CUB_ERR_BAD_ARGS;
        }
        Nccl* n = nccl();
        if (!n->error.empty()) {
            fprintf(stderr, "cubature_cuda: %s\n", n->error.c_str());
            delete c;
            return CUB_ERR_NCCL;
        }
        NcclId nid;
        memcpy(&nid, id, sizeof nid);
        int rc = n->CommInitRank(&c->nccl_comm, nranks, nid, rank);
        if (rc != 0) {
            fprintf(stderr, "cubature_cuda: ncclCommInitRank: %s\n", n->GetErrorString(rc));
            delete c;
            return CUB_ERR_NCCL;
        }
        // NVLink peer exchange buffers (best effort: without them the driver falls back to NCCL all-gathers).
        // CUBATURE_PEER_EXCHANGE=0 disables.
        const char* env = getenv("CUBATURE_PEER_EXCHANGE");
        if (nranks <= CUB_MAX_RANKS && !(env && env[0] == '0')) {
            bool ok = true;
            void* mine = nullptr;
            cudaIpcMemHandle_t h;
            cudaIpcMemHandle_t* d_handles = nullptr;
            std::vector<cudaIpcMemHandle_t> all((size_t)nranks);
            ok = ok && cudaMalloc(&mine, CUB_XBUF_BYTES) == cudaSuccess;
            ok = ok && cudaMemset(mine, 0, CUB_XBUF_BYTES) == cudaSuccess;
            ok = ok && cudaIpcGetMemHandle(&h, mine) == cudaSuccess;
            ok = ok && cudaMalloc((void**)&d_handles, sizeof(h) * (size_t)(nranks + 1)) == cudaSuccess;
            // every rank takes part in the all-gather even if its own setup failed (a zero handle then)
            if (!ok) memset(&h, 0, sizeof h);
            if (d_handles) {
                cudaMemcpy(d_handles + nranks, &h, sizeof h, cudaMemcpyHostToDevice);
                int rc2 = n->AllGather(d_handles + nranks, d_handles, sizeof h, /*ncclChar*/ 0, c->nccl_comm, (cudaStream_t)0);
                ok = ok && rc2 == 0 && cudaDeviceSynchronize() == cudaSuccess;
                ok = ok && cudaMemcpy(all.data(), d_handles, sizeof(h) * (size_t)nranks, cudaMemcpyDeviceToHost) == cudaSuccess;
                cudaFree(d_handles);
            }
            cudaIpcMemHandle_t zero;
            memset(&zero, 0, sizeof zero);
            for (int r = 0; ok && r < nranks; ++r) {
                if (memcmp(&all[(size_t)r], &zero, sizeof zero) == 0) ok = false;
            }
            for (int r = 0; ok && r < nranks; ++r) {
                if (r == rank) {
                    c->xbuf[r] = mine;
                } else if (cudaIpcOpenMemHandle(&c->xbuf[r], all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    ok = false;
                }
            }
            cudaGetLastError();
            // all ranks must agree: exchange the outcome (min over ranks) through one more all-gather
            int* d_ok = nullptr;
            std::vector<int> oks((size_t)nranks, 0);
            if (cudaMalloc((void**)&d_ok, sizeof(int) * (size_t)(nranks + 1)) == cudaSuccess) {
                const int mine_ok = ok ? 1 : 0;
                cudaMemcpy(d_ok + nranks, &mine_ok, sizeof(int), cudaMemcpyHostToDevice);
                if (n->AllGather(d_ok + nranks, d_ok, sizeof(int), 0, c->nccl_comm, (cudaStream_t)0) == 0 && cudaDeviceSynchronize() == cudaSuccess)
                    cudaMemcpy(oks.data(), d_ok, sizeof(int) * (size_t)nranks, cudaMemcpyDeviceToHost);
                cudaFree(d_ok);
            }
            bool all_ok = true;
            for (int r = 0; r < nranks; ++r) all_ok = all_ok && oks[(size_t)r] == 1;
            c->peer_ok = all_ok;
            if (!all_ok) {
                for (int r = 0; r < nranks; ++r)
                    if (r != rank && c->xbuf[r]) cudaIpcCloseMemHandle(c->xbuf[r]);
                if (mine) cudaFree(mine);
                for (int r = 0; r < nranks; ++r) c->xbuf[r] = nullptr;
                cudaGetLastError();
            }
        }
    }
    *out = c;
    return CUB_OK;
}

int cubature_cuda_comm_init_local(uint64_t group_key, int nranks, int rank, int device, cub_comm_t** out) {
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return CUB_ERR_BAD_ARGS;
    *out = nullptr;
    cub_comm* c = new cub_comm();
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    std::lock_guard<std::mutex> lock(g_groups_mu);
    cub_local_group*& g = g_groups[group_key];
    if (!g) {
        g = new cub_local_group();
        g->nranks = nranks;
        g->slots.resize((size_t)nranks);
    }
    if (g->nranks != nranks) {
        delete c;
        return CUB_ERR_BAD_ARGS;
    }
    g->refs += 1;
    c->local = g;
    c->local_key = group_key;
    *out = c;
    return CUB_OK;
}

void cubature_cuda_comm_free(cub_comm_t* comm) {
    if (!comm) return;
    if (comm->peer_ok) {
        for (int r = 0; r < comm->nranks; ++r) {
            if (!comm->xbuf[r]) continue;
            if (r == comm->rank)
                cudaFree(comm->xbuf[r]);
            else
                cudaIpcCloseMemHandle(comm->xbuf[r]);
        }
        cudaGetLastError();
    }
    if (comm->nccl_comm) nccl()->CommDestroy(comm->nccl_comm);
    if (comm->local) {
        std::lock_guard<std::mutex> lock(g_groups_mu);
        if (--comm->local->refs == 0) {
            g_groups.erase(comm->local_key);
            delete comm->local;
        }
    }
    delete comm;
}

}  // extern "C"
