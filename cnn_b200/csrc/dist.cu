// dist.cu -- the one exchange step of the data-parallel path: the batch mean of the weight
// gradients (conv2d.cpp:148,157, linear.cpp:62,70) over ranks = ONE sum all-reduce of the flat gradient
// slab, issued on the context's stream by the library itself so that the whole step
// (forward + backward + all-reduce + SGD) is a single CUDA-graph launch per rank.
//
// One process per GPU.  NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process this
// is the copy torch already loaded, otherwise the system library), so a single-GPU user needs no NCCL at
// all.  The 128-byte ncclUniqueId travels through whatever the launcher has (torch.distributed broadcast
// in cnn_b200/dist.py, MPI, a file, ...).
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    bool ok = false;
};

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    return api.ok ? &api : nullptr;
}

int nccl_fail(NcclApi* a, ncclResult_t r, const char* what) {
    cnn_set_error("NCCL error %d (%s) in %s", (int)r, a->GetErrorString(r), what);
    return CNN_ERR_NCCL;
}

}  // namespace

extern "C" {

int cnn_dist_unique_id(void* out128) {
    CNN_REQUIRE(out128, "cnn_dist_unique_id: NULL argument");
    NcclApi* a = nccl();
    if (!a) { cnn_set_error("libnccl.so.2 not found (%s)", dlerror()); return CNN_ERR_NCCL; }
    ncclUniqueId id;
    const ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(a, r, "ncclGetUniqueId");
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, sizeof(id));
    return CNN_OK;
}

int cnn_dist_init(cnn_ctx* ctx, int rank, int world, const void* id128) {
    CNN_REQUIRE(ctx && id128 && world >= 1 && rank >= 0 && rank < world, "cnn_dist_init: bad argument");
    CNN_REQUIRE(!ctx->nccl_comm, "cnn_dist_init: already initialised");
    NcclApi* a = nccl();
    if (!a) { cnn_set_error("libnccl.so.2 not found (%s)", dlerror()); return CNN_ERR_NCCL; }
    CNN_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t r = a->CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) return nccl_fail(a, r, "ncclCommInitRank");
    ctx->nccl_comm = comm;
    ctx->dist_rank = rank;
    ctx->dist_world = world;
    return CNN_OK;
}

int cnn_dist_set_sync_bn(cnn_ctx* ctx, int enable) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    ctx->sync_bn = enable != 0;
    return CNN_OK;
}

int cnn_dist_world(const cnn_ctx* ctx) { return ctx && ctx->nccl_comm ? ctx->dist_world : 1; }

int cnn_dist_allreduce_sum(cnn_ctx* ctx, float* buf, size_t n) {
    CNN_REQUIRE(ctx && buf, "cnn_dist_allreduce_sum: NULL argument");
    if (!ctx->nccl_comm || ctx->dist_world == 1 || n == 0) return CNN_OK;   // one rank: the sum is the buffer
    NcclApi* a = nccl();
    const ncclResult_t r = a->AllReduce(buf, buf, n, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(a, r, "ncclAllReduce");
    ++ctx->launches;
    return CNN_OK;
}

int cnn_dist_finalize(cnn_ctx* ctx) {
    if (!ctx || !ctx->nccl_comm) return CNN_OK;
    cudaStreamSynchronize(ctx->stream);
    NcclApi* a = nccl();
    if (a) a->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->dist_world = 1;
    ctx->dist_rank = 0;
    return CNN_OK;
}

}  // extern "C"
