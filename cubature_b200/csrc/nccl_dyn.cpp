// nccl_dyn.cpp -- NCCL bound at run time with dlopen("libnccl.so.2").
//
// Why dlopen: (1) libcubature_cuda.so must load on machines without NCCL (single-GPU use, CPU-only
// symbol checks); (2) inside a PyTorch process the already-loaded torch-bundled libnccl is reused
// (same soname), so there is exactly one NCCL in the address space.
// One cub_comm_t = one rank (one GPU) of a communicator; ranks live in separate processes
// (torchrun) or separate threads of one host process (the Java host of BASELINE.json).
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

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
        LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    });
    return &api;
}
}  // namespace

int cub_comm_allreduce_sum_f64(cub_comm* comm, double* dev_buf, size_t count, cudaStream_t stream) {
    if (!comm || comm->nranks == 1) return CUB_OK;
    Nccl* n = nccl();
    int rc = n->AllReduce(dev_buf, dev_buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm->nccl_comm, stream);
    return rc == 0 ? CUB_OK : CUB_ERR_NCCL;
}
int cub_comm_allreduce_max_u64(cub_comm* comm, unsigned long long* dev_buf, size_t count, cudaStream_t stream) {
    if (!comm || comm->nranks == 1) return CUB_OK;
    Nccl* n = nccl();
    int rc = n->AllReduce(dev_buf, dev_buf, count, /*ncclUint64*/ 5, /*ncclMax*/ 2, comm->nccl_comm, stream);
    return rc == 0 ? CUB_OK : CUB_ERR_NCCL;
}
int cub_comm_allreduce_sum_u64(cub_comm* comm, unsigned long long* dev_buf, size_t count, cudaStream_t stream) {
    if (!comm || comm->nranks == 1) return CUB_OK;
    Nccl* n = nccl();
    int rc = n->AllReduce(dev_buf, dev_buf, count, /*ncclUint64*/ 5, /*ncclSum*/ 0, comm->nccl_comm, stream);
    return rc == 0 ? CUB_OK : CUB_ERR_NCCL;
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
    }
    *out = c;
    return CUB_OK;
}

void cubature_cuda_comm_free(cub_comm_t* comm) {
    if (!comm) return;
    if (comm->nccl_comm) nccl()->CommDestroy(comm->nccl_comm);
    delete comm;
}

}  // extern "C"
