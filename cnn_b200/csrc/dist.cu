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

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
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

// ---- one-shot gradient exchange fused with the SGD step, over NVLink peer memory ---------------------------------
// The slab that has to be summed is tiny (445 KB for the reference net: latency, not bandwidth), so instead of a
// ring / tree collective every rank READS all peers' slabs directly (cudaIpc-mapped pointers, NVSwitch gives every
// pair full bandwidth), adds them in rank order -- the same order on every rank, so the replicas stay bit-identical
// without a broadcast -- and applies `p -= lr * g` in the same pass (conv2d.cpp:205-217 after the batch mean of
// :148,157).  Two flag rounds in peer memory replace the collective's synchronisation: "my slab is complete"
// before the reads, "I am done reading yours" before a slab may be rewritten.
constexpr int kMaxPeers = 16;

struct PeerArgs {
    const float* g[kMaxPeers];      // every rank's gradient slab (own pointer at [rank])
    uint32_t* flag[kMaxPeers];      // every rank's flag block: [0, 16) arrive, [16, 32) done, [32, 48) arrive of the early (bulk) phase
    uint32_t* state;                // local: [0] epoch of the last finished exchange, [1] CTA counter
    float* gsum;                    // local: reduced slab
    float* params;
    float* grads;                   // local slab (receives the reduced values in the finish kernel)
    size_t n, P;                    // slab length (P + 1: loss tail), parameter count
    int world, rank, do_sgd;
    const float* lr;                // device scalar (a captured step serves every learning rate)
    // one launch reduces slab elements [lo, hi) (lo a multiple of 4).  bulk = 1: the early phase on the side stream (the
    // gradients of every layer but the first, final before the first layer's backward runs) -- own arrive flags, no
    // "done" round; bulk = 0: the closing phase, which also tells the peers that this rank is done reading.
    size_t lo, hi;
    int bulk;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// peer memory changes between launches: never served from a cache line of ours; 16 bytes per request
__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// Thread = four consecutive slab elements: the loads from all ranks are issued back to back (one NVLink round trip
// of latency for the whole exchange, not one per rank), then added in rank order.
__global__ void __launch_bounds__(256) peer_allreduce_sgd_kernel(const PeerArgs a) {
    __shared__ uint32_t ep_s;
    if (threadIdx.x == 0) ep_s = *reinterpret_cast<volatile uint32_t*>(a.state) + 1;
    __syncthreads();
    const uint32_t ep = ep_s;
    const float lr = a.do_sgd ? *a.lr : 0.f;
    const int slot = a.bulk ? 2 * kMaxPeers : 0;
    // everything this rank wrote to [lo, hi) of its slab was written by earlier kernels of this stream: tell every peer
    if (blockIdx.x == 0 && (int)threadIdx.x < a.world) {
        __threadfence_system();
        st_release_sys(a.flag[threadIdx.x] + slot + a.rank, ep);
    }
    if ((int)threadIdx.x < a.world)
        while (ld_acquire_sys(a.flag[a.rank] + slot + threadIdx.x) < ep) {
        }
    __syncthreads();
    const size_t lo4 = a.lo / 4, hi4 = a.hi / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = lo4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
        float4 v[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < a.world) v[r] = ld_peer4(a.g[r] + 4 * i);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < a.world) {   // rank order: identical on every rank
                s.x = __fadd_rn(s.x, v[r].x); s.y = __fadd_rn(s.y, v[r].y); s.z = __fadd_rn(s.z, v[r].z); s.w = __fadd_rn(s.w, v[r].w);
            }
        *reinterpret_cast<float4*>(a.gsum + 4 * i) = s;
        if (a.do_sgd) {
            const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (4 * i + j < a.P) a.params[4 * i + j] = __fsub_rn(a.params[4 * i + j], __fmul_rn(lr, sv[j]));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.hi & 3)) {   // the last few elements of the range (the loss tail lives at the slab's end)
        const size_t i = hi4 * 4 + threadIdx.x;
        float s = 0.f;
        for (int r = 0; r < a.world; ++r) s = __fadd_rn(s, ld_peer(a.g[r] + i));
        a.gsum[i] = s;
        if (a.do_sgd && i < a.P) a.params[i] = __fsub_rn(a.params[i], __fmul_rn(lr, s));
    }
    if (a.bulk) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.state + 1, 1u) == gridDim.x - 1) {   // last CTA: this rank has read everything it needs (both phases)
            a.state[1] = 0;
            __threadfence_system();
            for (int r = 0; r < a.world; ++r) st_release_sys(a.flag[r] + kMaxPeers + a.rank, ep);
            *reinterpret_cast<volatile uint32_t*>(a.state) = ep;
        }
    }
}

// every peer is done reading this rank's slab: it may take the reduced values (what an in-place all-reduce leaves)
__global__ void __launch_bounds__(256) peer_allreduce_finish_kernel(const PeerArgs a) {
    const uint32_t ep = *reinterpret_cast<volatile uint32_t*>(a.state);
    if ((int)threadIdx.x < a.world)
        while (ld_acquire_sys(a.flag[a.rank] + kMaxPeers + threadIdx.x) < ep) {
        }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) a.grads[i] = a.gsum[i];
}

struct PeerState {
    PeerArgs a{};
    void* opened[2 * kMaxPeers] = {};
    int n_opened = 0;
    void* owned[3] = {};
};

}  // namespace

// Collective over the ranks of cnn_dist_init: maps every rank's gradient slab and flag block into every other rank
// (cudaIpc, exchanged through the NCCL communicator).  Returns CNN_ERR_UNSUPPORTED where peer mapping is not possible.
int cnn_peer_exchange_setup(cnn_ctx* ctx, float* grads, float* params, size_t P, void** state_out) {
    CNN_REQUIRE(ctx && grads && params && state_out, "cnn_peer_exchange_setup: NULL argument");
    *state_out = nullptr;
    if (!ctx->nccl_comm || ctx->dist_world < 2) return CNN_ERR_UNSUPPORTED;
    if (ctx->dist_world > kMaxPeers) return CNN_ERR_UNSUPPORTED;
    NcclApi* a = nccl();
    if (!a || !a->AllGather) return CNN_ERR_UNSUPPORTED;
    PeerState* st = new PeerState;
    const int world = ctx->dist_world, rank = ctx->dist_rank;
    uint32_t* flags = nullptr;
    uint32_t* state = nullptr;
    float* gsum = nullptr;
    struct Handles { cudaIpcMemHandle_t g, f; int ok; int pad[15]; };
    static_assert(sizeof(Handles) % 16 == 0, "handle record size");
    Handles mine{};
    Handles* d_all = nullptr;
    std::vector<Handles> all(world);
    bool ok = cudaMalloc(&flags, 3 * kMaxPeers * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc(&state, 2 * sizeof(uint32_t)) == cudaSuccess && cudaMalloc(&gsum, (P + 1) * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&d_all, sizeof(Handles) * world) == cudaSuccess;
    if (ok) {
        cudaMemset(flags, 0, 3 * kMaxPeers * sizeof(uint32_t));
        cudaMemset(state, 0, 2 * sizeof(uint32_t));
        ok = cudaIpcGetMemHandle(&mine.g, grads) == cudaSuccess && cudaIpcGetMemHandle(&mine.f, flags) == cudaSuccess;
    }
    mine.ok = ok ? 1 : 0;
    // the handle exchange is collective: every rank takes part even if its own preparation failed
    bool xok = d_all && cudaMemcpy(d_all + rank, &mine, sizeof(mine), cudaMemcpyHostToDevice) == cudaSuccess;
    if (d_all) {
        cudaStreamSynchronize(ctx->stream);
        xok = xok && a->AllGather(d_all + rank, d_all, sizeof(Handles), ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream) == ncclSuccess &&
              cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
              cudaMemcpy(all.data(), d_all, sizeof(Handles) * world, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    ok = ok && xok;
    for (int r = 0; ok && r < world; ++r) ok = all[r].ok == 1;
    for (int r = 0; ok && r < world; ++r) {
        if (r == rank) {
            st->a.g[r] = grads;
            st->a.flag[r] = flags;
            continue;
        }
        void *pg = nullptr, *pf = nullptr;
        ok = cudaIpcOpenMemHandle(&pg, all[r].g, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        if (ok) st->opened[st->n_opened++] = pg;
        ok = ok && cudaIpcOpenMemHandle(&pf, all[r].f, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        if (ok) st->opened[st->n_opened++] = pf;
        st->a.g[r] = static_cast<const float*>(pg);
        st->a.flag[r] = static_cast<uint32_t*>(pf);
    }
    if (d_all) cudaFree(d_all);
    // all ranks must agree: one failure anywhere keeps everybody on the NCCL path
    {
        float* d_ok = nullptr;
        float h_ok = ok ? 0.f : 1.f;
        if (cudaMalloc(&d_ok, sizeof(float)) == cudaSuccess) {
            cudaMemcpy(d_ok, &h_ok, sizeof(float), cudaMemcpyHostToDevice);
            a->AllReduce(d_ok, d_ok, 1, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            cudaMemcpy(&h_ok, d_ok, sizeof(float), cudaMemcpyDeviceToHost);
            cudaFree(d_ok);
            ok = h_ok == 0.f;
        } else {
            ok = false;
        }
    }
    cudaGetLastError();   // a failed IPC call must not poison later launches
    st->owned[0] = flags; st->owned[1] = state; st->owned[2] = gsum;
    if (!ok) {
        cnn_peer_exchange_destroy(st);
        return CNN_ERR_UNSUPPORTED;
    }
    st->a.state = state; st->a.gsum = gsum; st->a.params = params; st->a.grads = grads;
    st->a.n = P + 1; st->a.P = P; st->a.world = world; st->a.rank = rank;
    *state_out = st;
    return CNN_OK;
}

void cnn_peer_exchange_destroy(void* state) {
    PeerState* st = static_cast<PeerState*>(state);
    if (!st) return;
    for (int i = 0; i < st->n_opened; ++i) cudaIpcCloseMemHandle(st->opened[i]);
    for (void* p : st->owned) if (p) cudaFree(p);
    delete st;
}

namespace {
int peer_grid(const cnn_ctx* ctx, size_t elems, int per_sm) {
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count * per_sm, (elems / 4 + 255) / 256));
}
}  // namespace

// Early phase: slab elements [lo, P + 1) -- every gradient that is final before the first layer's backward runs, and
// the loss tail.  Launched on the caller's current stream (the engine's side stream: it overlaps the first layer's
// weight gradient); cnn_peer_exchange_step(..., lo) must follow on a stream ordered after it.
int cnn_peer_exchange_bulk(cnn_ctx* ctx, void* state, const float* lr_dev, int do_sgd, size_t lo) {
    PeerState* st = static_cast<PeerState*>(state);
    CNN_REQUIRE(ctx && st, "cnn_peer_exchange_bulk: NULL argument");
    CNN_REQUIRE((lo & 3) == 0 && lo < st->a.n, "cnn_peer_exchange_bulk: bad split");
    PeerArgs a = st->a;
    a.lr = lr_dev; a.do_sgd = do_sgd; a.lo = lo; a.hi = a.n; a.bulk = 1;
    CNN_LAUNCH(ctx, peer_allreduce_sgd_kernel, peer_grid(ctx, a.hi - a.lo, 1), 256, 0, a);
    return CNN_OK;
}

// Closing phase: slab elements [0, hi) (hi = P + 1 when there was no early phase), then the "done reading" round and
// the copy of the reduced slab.
int cnn_peer_exchange_step(cnn_ctx* ctx, void* state, const float* lr_dev, int do_sgd, size_t hi) {
    PeerState* st = static_cast<PeerState*>(state);
    CNN_REQUIRE(ctx && st, "cnn_peer_exchange_step: NULL argument");
    PeerArgs a = st->a;
    if (hi == 0 || hi > a.n) hi = a.n;
    a.lr = lr_dev; a.do_sgd = do_sgd; a.lo = 0; a.hi = hi; a.bulk = 0;
    CNN_LAUNCH(ctx, peer_allreduce_sgd_kernel, peer_grid(ctx, hi, 2), 256, 0, a);
    CNN_LAUNCH(ctx, peer_allreduce_finish_kernel, peer_grid(ctx, a.n, 2), 256, 0, a);
    return CNN_OK;
}

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
