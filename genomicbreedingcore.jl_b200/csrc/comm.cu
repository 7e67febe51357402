// comm.cu — the communicator of a multi-GPU context (one process per GPU, NCCL over NVLink / NVSwitch).
//
// The reference's functions are single calls (grmsimple(genomes), src/grm/calc.jl:115); for the B200 backend to shard
// one call over several GPUs the exchange steps have to live inside the library, behind the same entry points.  A
// context is given a communicator once (grm_cuda_comm_init: the 128-byte NCCL unique id is created on one rank with
// grm_cuda_comm_unique_id and handed to the others by whatever the host application already has — MPI, Distributed.jl,
// torch.distributed, a file); afterwards grm_cuda_simple / grm_cuda_ploidyaware with opts.distributed = 1 run the
// whole loci-split pipeline (run.cu).
//
// NCCL is loaded with dlopen on first use, so single-GPU users need no NCCL at all.  Which copy: the one the host named
// (grm_cuda_comm_set_library), else one that is ALREADY in the process (RTLD_NOLOAD — PyTorch brings its own), else the
// system's libnccl.so.2.  The order matters in someone else's process: the dynamic loader matches DT_NEEDED entries by
// soname, so an older system NCCL loaded here first would later be handed to a PyTorch that needs a newer one
// ("undefined symbol ncclDevCommCreate").  Hosts that load another NCCL afterwards should therefore name it up front;
// the Python wrapper does (the nvidia-nccl wheel PyTorch links against).
//
// Exchange of bulk data that only has to MOVE (the partial tiles of the K-split) does not go through NCCL kernels: they
// need SMs, which the persistent SYRK owns.  It goes through the copy engines into peer-mapped receive buffers
// (grm_comm_peer_buffer: cudaMalloc + CUDA IPC handles, exchanged once through the communicator).
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

struct GrmNccl {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr;  // optional
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;  // optional
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

constexpr int GRM_MAX_WORLD = 64;

struct GrmComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // copy-engine exchange
    void *recv = nullptr;  // own peer-visible buffer (cudaMalloc)
    size_t recv_bytes = 0;
    void *peers[GRM_MAX_WORLD] = {};  // peers[q] = rank q's buffer as mapped in this process
    uint32_t epoch = 0;
    cudaStream_t xs[GRM_XSTREAMS] = {};
    cudaEvent_t xe[GRM_XSTREAMS + 1] = {};
};

static GrmNccl g_nccl;
static char g_nccl_path[1024] = "";

static int32_t nccl_load() {
    if (g_nccl.handle) return GRM_CUDA_OK;
    void *h = nullptr;
    if (g_nccl_path[0]) {
        h = dlopen(g_nccl_path, RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            grm_set_error("dlopen(%s) failed: %s", g_nccl_path, dlerror());
            return GRM_CUDA_ERR_COMM;
        }
    }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);  // the host process already has one
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        grm_set_error("multi-GPU support needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
        return GRM_CUDA_ERR_COMM;
    }
    GrmNccl t;
    t.handle = h;
#define GRM_NCCL_SYM(field, name)                                                   \
    do {                                                                            \
        *reinterpret_cast<void **>(&t.field) = dlsym(h, name);                      \
        if (!t.field) {                                                             \
            grm_set_error("NCCL symbol %s not found", name);                        \
            return GRM_CUDA_ERR_COMM;                                               \
        }                                                                           \
    } while (0)
    GRM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    GRM_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    GRM_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    GRM_NCCL_SYM(AllReduce, "ncclAllReduce");
    GRM_NCCL_SYM(ReduceScatter, "ncclReduceScatter");
    GRM_NCCL_SYM(AllGather, "ncclAllGather");
    GRM_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    GRM_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef GRM_NCCL_SYM
    *reinterpret_cast<void **>(&t.CommInitRankConfig) = dlsym(h, "ncclCommInitRankConfig");
    *reinterpret_cast<void **>(&t.CommAbort) = dlsym(h, "ncclCommAbort");
    g_nccl = t;
    return GRM_CUDA_OK;
}

#define GRM_NCCL_TRY(expr)                                                                                  \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess) {                                                                            \
            grm_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, g_nccl.GetErrorString(_r));  \
            return GRM_CUDA_ERR_COMM;                                                                       \
        }                                                                                                   \
    } while (0)

extern "C" int32_t grm_cuda_comm_set_library(const char *path) try {
    if (g_nccl.handle) {
        grm_set_error("comm_set_library: NCCL is already loaded");
        return GRM_CUDA_ERR_COMM;
    }
    if (!path || strlen(path) >= sizeof(g_nccl_path)) {
        grm_set_error("comm_set_library: bad path");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    strcpy(g_nccl_path, path);
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_comm_unique_id(uint8_t *id128) try {
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id size");
    if (!id128) {
        grm_set_error("comm_unique_id: null pointer");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_TRY(nccl_load());
    ncclUniqueId id;
    GRM_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

static void peer_release(GrmComm *c) {
    for (int q = 0; q < c->world; ++q)
        if (c->peers[q] && q != c->rank) cudaIpcCloseMemHandle(c->peers[q]);
    for (auto &pp : c->peers) pp = nullptr;
    if (c->recv) cudaFree(c->recv);
    c->recv = nullptr;
    c->recv_bytes = 0;
}

// Context teardown without a prior grm_cuda_comm_destroy (an error path, an interpreter exiting after an exception): the
// peers may never enter the collectives this rank still has in flight, so they are aborted rather than waited for.
int32_t grm_comm_abort(grm_cuda_ctx *ctx) {
    if (ctx && ctx->comm && ctx->comm->comm && g_nccl.CommAbort) {
        g_nccl.CommAbort(ctx->comm->comm);
        ctx->comm->comm = nullptr;
    }
    return GRM_CUDA_OK;
}

int32_t grm_comm_destroy(grm_cuda_ctx *ctx) {
    if (ctx && ctx->comm) {
        peer_release(ctx->comm);
        for (auto &x : ctx->comm->xs)
            if (x) cudaStreamDestroy(x);
        for (auto &e : ctx->comm->xe)
            if (e) cudaEventDestroy(e);
        if (ctx->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_comm_init(grm_cuda_ctx *ctx, const uint8_t *id128, int32_t rank, int32_t world) try {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) {
        grm_set_error("comm_init: bad argument (rank=%d world=%d)", rank, world);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_TRY(nccl_load());
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    grm_comm_destroy(ctx);
    GrmComm *c = new (std::nothrow) GrmComm();
    if (!c) return GRM_CUDA_ERR_OOM;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    // tuning key "comm_ctas" (> 0, set before this call) caps NCCL's CTAs — for hosts that run the library next to
    // their own kernels.  Off by default: measured on C3 / 8 GPUs, 8 CTAs move ~80 GB/s (profiles/README.md).
    ncclResult_t r;
    const int ctas = ctx->tune.comm_ctas;
    if (world > GRM_MAX_WORLD) {
        grm_set_error("comm_init: world %d exceeds %d", world, GRM_MAX_WORLD);
        delete c;
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (ctas > 0 && g_nccl.CommInitRankConfig) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.maxCTAs = ctas;
        cfg.minCTAs = std::min(ctas, 4);
        r = g_nccl.CommInitRankConfig(&c->comm, world, id, rank, &cfg);
    } else {
        r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    }
    if (r != ncclSuccess) {
        grm_set_error("ncclCommInitRank(rank %d of %d) failed: %s", rank, world, g_nccl.GetErrorString(r));
        delete c;
        return GRM_CUDA_ERR_COMM;
    }
    c->rank = rank;
    c->world = world;
    ctx->comm = c;
    int prio_least = 0, prio_greatest = 0;
    GRM_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    for (auto &x : c->xs) GRM_CUDA_TRY(cudaStreamCreateWithPriority(&x, cudaStreamNonBlocking, prio_greatest));
    for (auto &e : c->xe) GRM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

// ---------------------------------------------------------------------------------------------------------------
// peer-mapped receive buffer
// ---------------------------------------------------------------------------------------------------------------
int32_t grm_comm_peer_buffer(grm_cuda_ctx *ctx, size_t bytes, void ***peers) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    GrmComm *c = ctx->comm;
    if (!c) {
        grm_set_error("no communicator on this context (grm_cuda_comm_init)");
        return GRM_CUDA_ERR_COMM;
    }
    if (c->recv && c->recv_bytes >= bytes) {
        *peers = c->peers;
        return GRM_CUDA_OK;
    }
    // (re)allocation: every rank gets here in the same call (same n, same world), so this is a collective
    GRM_CUDA_TRY(cudaDeviceSynchronize());
    peer_release(c);
    GRM_CUDA_TRY(cudaMalloc(&c->recv, bytes));
    GRM_CUDA_TRY(cudaMemset(c->recv, 0, GRM_PEER_HEADER));  // flag words start below every epoch
    c->recv_bytes = bytes;
    cudaIpcMemHandle_t mine;
    GRM_CUDA_TRY(cudaIpcGetMemHandle(&mine, c->recv));
    GRM_TRY(ctx->cons.reserve(8192 + 64 * static_cast<size_t>(c->world + 1)));
    uint8_t *d = ctx->cons.as<uint8_t>() + 8192;
    GRM_CUDA_TRY(cudaMemcpyAsync(d, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
    GRM_TRY(grm_comm_allgather(ctx, d, d + 64, 64, GRM_DT_U8, ctx->stream));
    std::vector<cudaIpcMemHandle_t> all(c->world);
    GRM_CUDA_TRY(cudaMemcpyAsync(all.data(), d + 64, 64 * static_cast<size_t>(c->world), cudaMemcpyDeviceToHost, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < c->world; ++q) {
        if (q == c->rank) {
            c->peers[q] = c->recv;
            continue;
        }
        const cudaError_t e = cudaIpcOpenMemHandle(&c->peers[q], all[q], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            grm_set_error("cudaIpcOpenMemHandle(rank %d's receive buffer) failed: %s", q, cudaGetErrorString(e));
            c->peers[q] = nullptr;
            return GRM_CUDA_ERR_COMM;
        }
    }
    // nobody may start storing into a buffer whose owner has not finished setting it up
    GRM_TRY(grm_comm_allreduce(ctx, d, d, 1, GRM_DT_U8, 1, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *peers = c->peers;
    return GRM_CUDA_OK;
}

uint32_t grm_comm_next_epoch(grm_cuda_ctx *ctx) { return ++ctx->comm->epoch; }
cudaStream_t grm_comm_xstream(grm_cuda_ctx *ctx, int i) { return ctx->comm->xs[i % GRM_XSTREAMS]; }
cudaEvent_t grm_comm_xevent(grm_cuda_ctx *ctx, int i) { return ctx->comm->xe[i % (GRM_XSTREAMS + 1)]; }

extern "C" int32_t grm_cuda_comm_destroy(grm_cuda_ctx *ctx) try {
    if (!ctx) return GRM_CUDA_ERR_BAD_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    return grm_comm_destroy(ctx);
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_comm_info(grm_cuda_ctx *ctx, int32_t *rank, int32_t *world, int32_t *nccl_version) try {
    if (!ctx) return GRM_CUDA_ERR_BAD_ARGUMENT;
    if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
    if (world) *world = ctx->comm ? ctx->comm->world : 1;
    if (nccl_version) {
        int v = 0;
        if (g_nccl.GetVersion) g_nccl.GetVersion(&v);
        *nccl_version = v;
    }
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

int grm_comm_rank(const grm_cuda_ctx *ctx) { return ctx->comm ? ctx->comm->rank : 0; }
int grm_comm_world(const grm_cuda_ctx *ctx) { return ctx->comm ? ctx->comm->world : 1; }

static ncclDataType_t nccl_type(int t) {
    switch (t) {
        case GRM_DT_I32: return ncclInt32;
        case GRM_DT_I64: return ncclInt64;
        case GRM_DT_U8: return ncclUint8;
        default: return ncclFloat64;
    }
}

int32_t grm_comm_allreduce(grm_cuda_ctx *ctx, const void *send, void *recv, size_t count, int dtype, int op_max,
                           cudaStream_t s) {
    if (!ctx->comm) {
        grm_set_error("no communicator on this context (grm_cuda_comm_init)");
        return GRM_CUDA_ERR_COMM;
    }
    GRM_NCCL_TRY(g_nccl.AllReduce(send, recv, count, nccl_type(dtype), op_max ? ncclMax : ncclSum, ctx->comm->comm, s));
    return GRM_CUDA_OK;
}

int32_t grm_comm_reduce_scatter(grm_cuda_ctx *ctx, const void *send, void *recv, size_t recvcount, int dtype,
                                cudaStream_t s) {
    if (!ctx->comm) {
        grm_set_error("no communicator on this context (grm_cuda_comm_init)");
        return GRM_CUDA_ERR_COMM;
    }
    GRM_NCCL_TRY(g_nccl.ReduceScatter(send, recv, recvcount, nccl_type(dtype), ncclSum, ctx->comm->comm, s));
    return GRM_CUDA_OK;
}

int32_t grm_comm_allgather(grm_cuda_ctx *ctx, const void *send, void *recv, size_t sendcount, int dtype, cudaStream_t s) {
    if (!ctx->comm) {
        grm_set_error("no communicator on this context (grm_cuda_comm_init)");
        return GRM_CUDA_ERR_COMM;
    }
    GRM_NCCL_TRY(g_nccl.AllGather(send, recv, sendcount, nccl_type(dtype), ctx->comm->comm, s));
    return GRM_CUDA_OK;
}
