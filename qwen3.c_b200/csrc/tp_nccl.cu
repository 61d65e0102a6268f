// tp_nccl.cu -- tensor-parallel plumbing: NCCL (loaded with dlopen, so single-GPU users need
// no NCCL at all) for the two all-reduces per layer and the logits all-gather
// (SURVEY.md section 8e). One process per GPU; the unique id travels through the caller's
// own rendezvous (bench.py / tests use torch.distributed for that).
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

int qw_mega_tp_connect(QwenCudaCtx* c);

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
} g;

int load_nccl() {
    if (g.h) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g.h) break;
    }
    if (!g.h) {
        qw_set_error("tensor parallelism needs NCCL: %s", dlerror());
        return -1;
    }
    g.GetUniqueId = (decltype(g.GetUniqueId)) dlsym(g.h, "ncclGetUniqueId");
    g.CommInitRank = (decltype(g.CommInitRank)) dlsym(g.h, "ncclCommInitRank");
    g.CommDestroy = (decltype(g.CommDestroy)) dlsym(g.h, "ncclCommDestroy");
    g.AllReduce = (decltype(g.AllReduce)) dlsym(g.h, "ncclAllReduce");
    g.AllGather = (decltype(g.AllGather)) dlsym(g.h, "ncclAllGather");
    g.GetErrorString = (decltype(g.GetErrorString)) dlsym(g.h, "ncclGetErrorString");
    if (!g.GetUniqueId || !g.CommInitRank || !g.AllReduce || !g.AllGather) {
        qw_set_error("libnccl is missing expected symbols");
        return -1;
    }
    return 0;
}
} // namespace

extern "C" int qwen_cuda_tp_unique_id(void* out128) {
    if (load_nccl()) return -1;
    ncclUniqueId id;
    const int rc = g.GetUniqueId(&id);
    if (rc != ncclSuccess) {
        qw_set_error("ncclGetUniqueId: %s", g.GetErrorString ? g.GetErrorString(rc) : "?");
        return -1;
    }
    memcpy(out128, &id, sizeof id);
    return 0;
}

extern "C" int qwen_cuda_tp_init(QwenCudaCtx* c, const void* id128) {
    if (!c || !id128) return -2;
    if (c->tp_size == 1) return 0;
    if (load_nccl()) return -1;
    QW_CUDA(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    const int rc = g.CommInitRank(&comm, c->tp_size, id, c->tp_rank);
    if (rc != ncclSuccess) {
        qw_set_error("ncclCommInitRank: %s", g.GetErrorString ? g.GetErrorString(rc) : "?");
        return -1;
    }
    c->nccl_comm = comm;
    // map the peers' flow arenas: from here on the persistent kernel does the all-reduce itself (decode_mega.cu)
    return qw_mega_tp_connect(c);
}

void qw_tp_free(QwenCudaCtx* c) {
    if (c->nccl_comm && g.CommDestroy) g.CommDestroy((ncclComm_t) c->nccl_comm);
    c->nccl_comm = nullptr;
}

int qw_tp_allreduce(QwenCudaCtx* c, float* buf, size_t n) {
    if (c->tp_size == 1) return 0;
    if (!c->nccl_comm) {
        qw_set_error("tensor-parallel context used before qwen_cuda_tp_init");
        return -1;
    }
    const int rc = g.AllReduce(buf, buf, n, ncclFloat32, ncclSum, (ncclComm_t) c->nccl_comm, c->stream);
    if (rc != ncclSuccess) {
        qw_set_error("ncclAllReduce: %s", g.GetErrorString ? g.GetErrorString(rc) : "?");
        return -1;
    }
    return 0;
}

int qw_tp_allgather(QwenCudaCtx* c, const float* src, float* dst, size_t n_per_rank) {
    if (c->tp_size == 1) return 0;
    if (!c->nccl_comm) {
        qw_set_error("tensor-parallel context used before qwen_cuda_tp_init");
        return -1;
    }
    const int rc = g.AllGather(src, dst, n_per_rank, ncclFloat32, (ncclComm_t) c->nccl_comm, c->stream);
    if (rc != ncclSuccess) {
        qw_set_error("ncclAllGather: %s", g.GetErrorString ? g.GetErrorString(rc) : "?");
        return -1;
    }
    return 0;
}
