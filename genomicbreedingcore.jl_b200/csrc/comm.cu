// comm.cu — the communicator of a multi-GPU context (one process per GPU, NCCL over NVLink / NVSwitch).
//
// The reference's functions are single calls (grmsimple(genomes), src/grm/calc.jl:115); for the B200 backend to shard
// one call over several GPUs the exchange steps have to live inside the library, behind the same entry points.  A
// context is given a communicator once (grm_cuda_comm_init: the 128-byte NCCL unique id is created on one rank with
// grm_cuda_comm_unique_id and handed to the others by whatever the host application already has — MPI, Distributed.jl,
// torch.distributed, a file); afterwards grm_cuda_simple / grm_cuda_ploidyaware with opts.distributed = 1 run the
// whole loci-split pipeline (run.cu).
//
// NCCL is loaded with dlopen on first use, so single-GPU users need no NCCL at all and, inside a process that already
// carries an NCCL (PyTorch), that same copy is used (dlopen by soname returns the loaded object).
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"

struct GrmNccl {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

struct GrmComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

static GrmNccl g_nccl;

static int32_t nccl_load() {
    if (g_nccl.handle) return GRM_CUDA_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
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

extern "C" int32_t grm_cuda_comm_unique_id(uint8_t *id128) {
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
}

int32_t grm_comm_destroy(grm_cuda_ctx *ctx) {
    if (ctx && ctx->comm) {
        if (ctx->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_comm_init(grm_cuda_ctx *ctx, const uint8_t *id128, int32_t rank, int32_t world) {
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
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        grm_set_error("ncclCommInitRank(rank %d of %d) failed: %s", rank, world, g_nccl.GetErrorString(r));
        delete c;
        return GRM_CUDA_ERR_COMM;
    }
    c->rank = rank;
    c->world = world;
    ctx->comm = c;
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_comm_destroy(grm_cuda_ctx *ctx) {
    if (!ctx) return GRM_CUDA_ERR_BAD_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    return grm_comm_destroy(ctx);
}

extern "C" int32_t grm_cuda_comm_info(grm_cuda_ctx *ctx, int32_t *rank, int32_t *world, int32_t *nccl_version) {
    if (!ctx) return GRM_CUDA_ERR_BAD_ARGUMENT;
    if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
    if (world) *world = ctx->comm ? ctx->comm->world : 1;
    if (nccl_version) {
        int v = 0;
        if (g_nccl.GetVersion) g_nccl.GetVersion(&v);
        *nccl_version = v;
    }
    return GRM_CUDA_OK;
}

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
