// ctx.cu -- context, error reporting and memory entry points of the C ABI.
#include <sched.h>

#include <cctype>
#include <cstring>

#include <mutex>

#include "common.cuh"

static thread_local char g_err[512] = "";

std::recursive_mutex& cnn_global_mutex() {
    static std::recursive_mutex m;
    return m;
}

void cnn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cnn_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    cnn_set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return CNN_ERR_CUDA;
}

float* cnn_scratch(cnn_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return ctx->scratch;
    // stream-ordered growth would break graph capture; grow eagerly and generously instead
    if (ctx->scratch) cudaFree(ctx->scratch);
    size_t want = bytes < (size_t(8) << 20) ? (size_t(8) << 20) : bytes;
    if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        return nullptr;
    }
    ctx->scratch_bytes = want;
    return ctx->scratch;
}

void* cnn_arena(cnn_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->arena_bytes) return ctx->arena;
    if (ctx->arena) cudaFree(ctx->arena);   // synchronises: nothing in flight still reads the old arena
    ctx->arena = nullptr;
    ctx->arena_bytes = 0;
    if (cudaMalloc(&ctx->arena, bytes) != cudaSuccess) return nullptr;
    ctx->arena_bytes = bytes;
    return ctx->arena;
}

// cuTensorMapEncodeTiled is a driver-API entry point: resolved through the runtime (no -lcuda at link time)
int cnn_tmap_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    {
        std::lock_guard<std::recursive_mutex> lk(cnn_global_mutex());
        if (!fn) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
            if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
                cnn_set_error("cuTensorMapEncodeTiled is not available from this driver");
                return CNN_ERR_CUDA;
            }
            fn = reinterpret_cast<EncodeFn>(p);
        }
    }
    CNN_REQUIRE(rank >= 2 && rank <= 5, "cnn_tmap_encode: rank %d", rank);
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    // 4-byte elements moved verbatim (fp32 data and packed bf16 pairs alike); out-of-bounds box elements read as zero
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cnn_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return CNN_ERR_CUDA;
    }
    return CNN_OK;
}

int cnn_tmap_encode_3d(CUtensorMap* map, const void* base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                       const uint32_t box[3]) {
    return cnn_tmap_encode(map, base, 3, dims, strides_bytes, box);
}

void cnn_prof_mark(cnn_ctx* ctx, const char* name) {
    cnn_prof* p = ctx->prof;
    if (!p || p->n >= cnn_prof::kMax) return;
    cudaEventRecord(p->ev[p->n], ctx->stream);
    p->tag[p->n] = ctx->prof_tag;
    if (name[0] == '(') ++name;   // CNN_LAUNCH(ctx, (kernel<a, b>), ...) stringifies with its parentheses
    p->name[p->n++] = name;
}

extern "C" {

const char* cnn_last_error(void) { return g_err; }
const char* cnn_version(void) { return "cnn_b200 0.1 (sm_100a)"; }

int cnn_ctx_create(int device, void* stream, cnn_ctx** out) {
    CNN_REQUIRE(out, "cnn_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cnn_set_error("no CUDA device available (%s); libcnn_b200 has no CPU fallback",
                      e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
        return CNN_ERR_CUDA;
    }
    CNN_REQUIRE(device >= 0 && device < n, "cnn_ctx_create: device %d out of range (%d)", device, n);
    CNN_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CNN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        cnn_set_error("device %d is sm_%d%d; libcnn_b200 is built for sm_100a only", device, prop.major,
                      prop.minor);
        return CNN_ERR_UNSUPPORTED;
    }
    cnn_ctx* c = new cnn_ctx;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return cnn_cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
        }
        c->own_stream = true;
    }
    c->thin_slot = conv_thin_acquire_slot(device);
    if (!cnn_scratch(c, size_t(8) << 20)) {
        conv_thin_release_slot(device, c->thin_slot);
        delete c;
        cnn_set_error("scratch allocation failed");
        return CNN_ERR_CUDA;
    }
    *out = c;
    return CNN_OK;
}

int cnn_ctx_destroy(cnn_ctx* ctx) {
    if (!ctx) return CNN_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cnn_dist_finalize(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    conv_thin_release_slot(ctx->device, ctx->thin_slot);
    if (ctx->prof) {
        for (int i = 0; i <= cnn_prof::kMax; ++i) cudaEventDestroy(ctx->prof->ev[i]);
        delete ctx->prof;
    }
    delete ctx;
    return CNN_OK;
}

int cnn_ctx_set_stream(cnn_ctx* ctx, void* stream) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    if (ctx->own_stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    ctx->stream = (cudaStream_t)stream;
    return CNN_OK;
}

void* cnn_ctx_stream(cnn_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int cnn_ctx_set_conv_algo(cnn_ctx* ctx, int algo) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    CNN_REQUIRE(algo >= CNN_CONV_AUTO && algo <= CNN_CONV_TCGEN05, "unknown conv algo %d", algo);
    ctx->conv_algo = algo;
    return CNN_OK;
}

int cnn_ctx_set_tc_precision(cnn_ctx* ctx, int mode) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    CNN_REQUIRE(mode == CNN_TC_TF32X3 || mode == CNN_TC_BF16X3 || mode == CNN_TC_MIXED || mode == CNN_TC_BF16X1,
                "unknown tensor-core precision mode %d", mode);
    ctx->tc_precision = mode;
    return CNN_OK;
}

// One process per GPU: keep the calling thread (and with it the pages it touches first -- pinned staging buffers) on
// the NUMA node the GPU hangs off.  On a two-socket 8-GPU box the host->device copies of all ranks otherwise meet in
// one socket's memory controllers.  Linux sysfs only; a no-op (CNN_OK) where the topology cannot be read.
int cnn_ctx_bind_numa(cnn_ctx* ctx) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), ctx->device) != cudaSuccess) return CNN_OK;
    for (char* p = bus; *p; ++p) *p = (char)tolower(*p);
    char path[256];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return CNN_OK;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return CNN_OK;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return CNN_OK;
    char list[4096] = {0};
    const bool ok = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!ok) return CNN_OK;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k >= 1)
            for (int c2 = a; c2 <= b && c2 < CPU_SETSIZE; ++c2) { CPU_SET(c2, &set); ++n; }
    }
    if (n > 0) sched_setaffinity(0, sizeof(set), &set);
    return CNN_OK;
}

int cnn_sync(cnn_ctx* ctx) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    return CNN_OK;
}

long long cnn_launch_count(cnn_ctx* ctx) { return ctx ? ctx->launches : 0; }

int cnn_prof_begin(cnn_ctx* ctx) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    if (!ctx->prof) {
        ctx->prof = new cnn_prof;
        for (int i = 0; i <= cnn_prof::kMax; ++i) CNN_CUDA(cudaEventCreate(&ctx->prof->ev[i]));
    }
    ctx->prof->n = 0;
    ctx->prof->on = true;
    return CNN_OK;
}

int cnn_prof_end(cnn_ctx* ctx, char* names, size_t names_cap, float* us, int max_entries, int* n_out) {
    CNN_REQUIRE(ctx && ctx->prof && n_out, "cnn_prof_end: no profile in progress");
    cnn_prof* p = ctx->prof;
    p->on = false;
    CNN_CUDA(cudaEventRecord(p->ev[p->n], ctx->stream));
    CNN_CUDA(cudaEventSynchronize(p->ev[p->n]));
    const int n = p->n < max_entries ? p->n : max_entries;
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        CNN_CUDA(cudaEventElapsedTime(&ms, p->ev[i], p->ev[i + 1]));
        if (us) us[i] = ms * 1e3f;
        if (names) {
            char buf[160];
            if (p->tag[i] >= 0) snprintf(buf, sizeof(buf), "L%d%c:%s", p->tag[i] >> 2, "fbu?"[p->tag[i] & 3], p->name[i]);
            else snprintf(buf, sizeof(buf), "%s", p->name[i]);
            size_t len = strlen(buf);
            if (len && buf[len - 1] == ')') buf[--len] = 0;
            if (off + len + 2 > names_cap) break;
            memcpy(names + off, buf, len);
            off += len;
            names[off++] = '\n';
        }
    }
    if (names && names_cap) names[off < names_cap ? off : names_cap - 1] = 0;
    *n_out = n;
    return CNN_OK;
}

int cnn_malloc(cnn_ctx* ctx, size_t bytes, void** dptr) {
    CNN_REQUIRE(ctx && dptr, "cnn_malloc: NULL argument");
    CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_CUDA(cudaMalloc(dptr, bytes ? bytes : 4));
    return CNN_OK;
}

int cnn_free(cnn_ctx* ctx, void* dptr) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    if (dptr) CNN_CUDA(cudaFree(dptr));
    return CNN_OK;
}

int cnn_host_alloc(cnn_ctx* ctx, size_t bytes, void** hptr) {
    CNN_REQUIRE(ctx && hptr, "cnn_host_alloc: NULL argument");
    CNN_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 4, cudaHostAllocDefault));
    return CNN_OK;
}

int cnn_host_free(cnn_ctx* ctx, void* hptr) {
    CNN_REQUIRE(ctx, "ctx is NULL");
    if (hptr) CNN_CUDA(cudaFreeHost(hptr));
    return CNN_OK;
}

int cnn_memset(cnn_ctx* ctx, void* dptr, int byte, size_t bytes) {
    CNN_REQUIRE(ctx && dptr, "cnn_memset: NULL argument");
    CNN_CUDA(cudaMemsetAsync(dptr, byte, bytes, ctx->stream));
    return CNN_OK;
}

int cnn_h2d(cnn_ctx* ctx, void* dst, const void* host_src, size_t bytes) {
    CNN_REQUIRE(ctx && dst && host_src, "cnn_h2d: NULL argument");
    CNN_CUDA(cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return CNN_OK;
}

int cnn_d2h(cnn_ctx* ctx, void* host_dst, const void* src, size_t bytes) {
    CNN_REQUIRE(ctx && host_dst && src, "cnn_d2h: NULL argument");
    CNN_CUDA(cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    return CNN_OK;
}

int cnn_d2d(cnn_ctx* ctx, void* dst, const void* src, size_t bytes) {
    CNN_REQUIRE(ctx && dst && src, "cnn_d2d: NULL argument");
    CNN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return CNN_OK;
}

}  // extern "C"
