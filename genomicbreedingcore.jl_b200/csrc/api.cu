// api.cu — context, error plumbing, tuning knobs and the tile partition of libgrm_cuda.  The reference-facing entry points
// (grmsimple / grmploidyaware / inflatediagonals!, src/grm/calc.jl:115-142, :189-217, :41-70) live in run.cu.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "common.cuh"

#include "ingest_host.h"

int32_t grm_comm_destroy(grm_cuda_ctx *ctx);  // comm.cu
int32_t grm_comm_abort(grm_cuda_ctx *ctx);

// ---------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void grm_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *grm_cuda_last_error(void) { return g_err; }
extern "C" int32_t grm_cuda_version(void) { return GRM_CUDA_VERSION; }

extern "C" const char *grm_cuda_status_string(int32_t s) {
    switch (s) {
        case GRM_CUDA_OK: return "ok";
        case GRM_CUDA_ERR_BAD_ARGUMENT: return "bad argument";
        case GRM_CUDA_ERR_MISSING: return "allele_frequencies contains missing (NaN) values";
        case GRM_CUDA_ERR_NONFINITE: return "allele_frequencies contains non-finite values";
        case GRM_CUDA_ERR_NOT_SYMMETRIC: return "The input matrix is not symmetric.";
        case GRM_CUDA_ERR_NAN_MATRIX: return "The input matrix contains NaN values.";
        case GRM_CUDA_ERR_INF_MATRIX: return "The input matrix contains Inf values.";
        case GRM_CUDA_ERR_CUDA: return "CUDA error";
        case GRM_CUDA_ERR_OOM: return "out of memory";
        case GRM_CUDA_ERR_NOT_DOSAGE: return "input is not dosage-quantised";
        case GRM_CUDA_ERR_OVERFLOW: return "int32 accumulator overflow";
        case GRM_CUDA_ERR_ALL_FILTERED: return "All loci filtered out";
        case GRM_CUDA_ERR_NOT_POSDEF: return "matrix is not positive definite";
        case GRM_CUDA_ERR_COMM: return "communicator (NCCL) error";
        default: return "unknown status";
    }
}

extern "C" void grm_cuda_options_default(grm_cuda_options *o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->struct_size = static_cast<int32_t>(sizeof(*o));
    o->path = GRM_CUDA_PATH_AUTO;
    o->max_iter = 1000;  // calc.jl:41,115,189
    o->world = 1;
    o->maf = 0.0;
    o->max_locus_sparsity = 1.0;
}

int32_t grm_tmap_encoder(GrmEncodeTiledFn *out) {
    static GrmEncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GRM_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) {
            grm_set_error("cuTensorMapEncodeTiled not available from the driver");
            return GRM_CUDA_ERR_CUDA;
        }
        encode = reinterpret_cast<GrmEncodeTiledFn>(fn);
    }
    *out = encode;
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// buffers / context
// ---------------------------------------------------------------------------------------------------------------
int32_t DevBuf::reserve(size_t bytes) {
    if (bytes <= cap && ptr) return GRM_CUDA_OK;
    if (ptr) {
        cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    GRM_CUDA_TRY(cudaMalloc(&ptr, bytes));
    cap = bytes;
    return GRM_CUDA_OK;
}
void DevBuf::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

extern "C" int32_t grm_cuda_ctx_create(int32_t device, grm_cuda_ctx **out) try {
    if (!out) {
        grm_set_error("ctx_create: null output pointer");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        grm_set_error("no CUDA device available (%s); libgrm_cuda has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return GRM_CUDA_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        grm_set_error("ctx_create: device %d out of range [0,%d)", device, count);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    GRM_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        grm_set_error("device %d is sm_%d%d; libgrm_cuda is built for sm_100a (B200) only", device, prop.major,
                      prop.minor);
        return GRM_CUDA_ERR_CUDA;
    }
    grm_cuda_ctx *ctx = new (std::nothrow) grm_cuda_ctx();
    if (!ctx) return GRM_CUDA_ERR_OOM;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    int prio_least = 0, prio_greatest = 0;
    GRM_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    GRM_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest));
    GRM_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    GRM_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_greatest));
    for (auto &ev : ctx->ev) GRM_CUDA_TRY(cudaEventCreate(&ev));
    for (auto &ev : ctx->wave_ev) GRM_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) {
        // waited for by host threads (ingest ring, result copy-out) while the worker pool needs every core: block, do not spin
        GRM_CUDA_TRY(cudaEventCreateWithFlags(&ctx->copy_done[i], cudaEventDisableTiming | cudaEventBlockingSync));
        GRM_CUDA_TRY(cudaEventCreateWithFlags(&ctx->stage_free[i], cudaEventDisableTiming));
    }
    *out = ctx;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

int32_t grm_cusolver_destroy(void *h);
int32_t grm_cusolver_params_destroy(void *p);

extern "C" int32_t grm_cuda_ctx_destroy(grm_cuda_ctx *ctx) try {
    if (!ctx) return GRM_CUDA_OK;
    cudaSetDevice(ctx->device);
    if (ctx->comm) grm_comm_abort(ctx);  // never wait for collectives the peers may not enter any more
    cudaDeviceSynchronize();
    if (ctx->comm) grm_comm_destroy(ctx);
    if (ctx->cusolver) grm_cusolver_destroy(ctx->cusolver);
    if (ctx->cusolver_params) grm_cusolver_params_destroy(ctx->cusolver_params);
    for (auto &h : ctx->cusolver_lane)
        if (h) grm_cusolver_destroy(h);
    for (auto &x : ctx->lane_stream)
        if (x) cudaStreamDestroy(x);
    for (auto &e : ctx->lane_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->lane_pin) cudaFreeHost(ctx->lane_pin);
    if (ctx->pool) grm_pool_destroy(ctx->pool);
    for (auto &pp : ctx->pin)
        if (pp) cudaFreeHost(pp);
    DevBuf *bufs[] = {&ctx->codes, &ctx->colsum, &ctx->flags, &ctx->q, &ctx->w, &ctx->u, &ctx->stats, &ctx->partials,
                      &ctx->G, &ctx->C, &ctx->stage[0], &ctx->stage[1], &ctx->stage[2], &ctx->stage[3], &ctx->lu, &ctx->lu2, &ctx->lu3, &ctx->lane_work[0], &ctx->lane_work[1], &ctx->lane_ipiv[0],
                      &ctx->lane_ipiv[1],
                      &ctx->lu_work, &ctx->ipiv, &ctx->misc, &ctx->misc2, &ctx->planes, &ctx->rT, &ctx->cnt, &ctx->keep,
                      &ctx->progress, &ctx->packed, &ctx->reduced, &ctx->gtiles, &ctx->gtiles_all, &ctx->codes_all,
                      &ctx->cons, &ctx->tiles_i8.buf, &ctx->tiles_i8_2cta.buf, &ctx->tiles_f64.buf, &ctx->tiles_packed.buf,
                      &ctx->tiles_wide.buf, &ctx->tiles_f64_all.buf, &ctx->packed_plan[0].buf, &ctx->packed_plan[1].buf};
    for (DevBuf *b : bufs) b->release();
    for (auto &ev : ctx->ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->wave_ev) cudaEventDestroy(ev);
    for (int i = 0; i < 4; ++i) {
        cudaEventDestroy(ctx->copy_done[i]);
        cudaEventDestroy(ctx->stage_free[i]);
    }
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->comm_stream);
    delete ctx;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" void *grm_cuda_ctx_stream(grm_cuda_ctx *ctx) { return ctx ? ctx->stream : nullptr; }

extern "C" int32_t grm_cuda_ctx_synchronize(grm_cuda_ctx *ctx) try {
    if (!ctx) return GRM_CUDA_ERR_BAD_ARGUMENT;
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->comm_stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

// ---------------------------------------------------------------------------------------------------------------
// per-context kernel attributes and tuning knobs
// ---------------------------------------------------------------------------------------------------------------
int32_t grm_func_smem(grm_cuda_ctx *ctx, const void *func, size_t bytes) {
    for (const void *f : ctx->attr_done)
        if (f == func) return GRM_CUDA_OK;
    GRM_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    ctx->attr_done.push_back(func);
    return GRM_CUDA_OK;
}

static int *tuning_slot(GrmTuning &t, const char *key) {
    struct Entry {
        const char *name;
        int GrmTuning::*field;
    };
    static const Entry table[] = {
        {"syrk_i8", &GrmTuning::syrk_i8},           {"syrk_packed", &GrmTuning::syrk_packed},
        {"syrk_sync", &GrmTuning::syrk_sync},       {"syrk_window", &GrmTuning::syrk_window},
        {"pack", &GrmTuning::pack},                 {"ieee_div", &GrmTuning::ieee_div},
        {"tile_band", &GrmTuning::tile_band},       {"host_threads", &GrmTuning::host_threads},
        {"ingest", &GrmTuning::ingest},             {"ksplit_waves", &GrmTuning::ksplit_waves},
        {"f64_kernel", &GrmTuning::f64_kernel},     {"ring_mb", &GrmTuning::ring_mb},
        {"comm_ctas", &GrmTuning::comm_ctas},       {"exchange", &GrmTuning::exchange},
        {"repair", &GrmTuning::repair},             {"repair_det", &GrmTuning::repair_det},
    };
    for (const Entry &e : table)
        if (!strcmp(e.name, key)) return &(t.*(e.field));
    return nullptr;
}

extern "C" int32_t grm_cuda_ctx_set_tuning(grm_cuda_ctx *ctx, const char *key, int64_t value) try {
    int *slot = (ctx && key) ? tuning_slot(ctx->tune, key) : nullptr;
    if (!slot || value < INT32_MIN || value > INT32_MAX) {
        grm_set_error("ctx_set_tuning: unknown key or bad value (%s)", key ? key : "null");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    *slot = static_cast<int>(value);
    if (!strcmp(key, "host_threads") && ctx->pool) {  // the pool is rebuilt with the new size on the next host call
        grm_pool_destroy(ctx->pool);
        ctx->pool = nullptr;
    }
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_ctx_get_tuning(grm_cuda_ctx *ctx, const char *key, int64_t *value) try {
    int *slot = (ctx && key) ? tuning_slot(ctx->tune, key) : nullptr;
    if (!slot || !value) {
        grm_set_error("ctx_get_tuning: unknown key (%s)", key ? key : "null");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    *value = *slot;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

// ---------------------------------------------------------------------------------------------------------------
// tile partition: 256 x 256 tiles of the upper triangle, enumerated in bands of `band` tile rows so that tiles that are
// in flight together (and the contiguous range a rank owns) share few operand panels -> L2 reuse of the codes.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GRM_TILE_BAND_DEFAULT = 8;

void grm_tile_order(int64_t n, int band, std::vector<int2> &tiles) {
    const int T = static_cast<int>((n + GRM_TILE - 1) / GRM_TILE);
    const int BAND = band > 0 ? band : GRM_TILE_BAND_DEFAULT;
    tiles.clear();
    tiles.reserve(static_cast<size_t>(T) * (T + 1) / 2);
    for (int r0 = 0; r0 < T; r0 += BAND) {
        const int r1 = std::min(T, r0 + BAND);
        for (int tj = r0; tj < T; ++tj)
            for (int ti = r0; ti < std::min(r1, tj + 1); ++ti) tiles.push_back(make_int2(ti, tj));
    }
}

void grm_tile_range(int64_t total, int32_t rank, int32_t world, int64_t *begin, int64_t *end) {
    *begin = total * rank / world;
    *end = total * (rank + 1) / world;
}

extern "C" int32_t grm_cuda_tile_edge(void) { return GRM_TILE; }

extern "C" int64_t grm_cuda_tile_count(int64_t n, int32_t rank, int32_t world) try {
    if (n < 1 || world < 1 || rank < 0 || rank >= world) return -1;
    const int64_t T = (n + GRM_TILE - 1) / GRM_TILE;
    int64_t b, e;
    grm_tile_range(T * (T + 1) / 2, rank, world, &b, &e);
    return e - b;
} GRM_CATCH_I64

extern "C" int32_t grm_cuda_tile_owner(int64_t n, int64_t ti, int64_t tj, int32_t world) try {
    if (n < 1 || world < 1 || ti < 0 || tj < ti) return -1;
    std::vector<int2> order;
    grm_tile_order(n, 0, order);
    const int64_t total = static_cast<int64_t>(order.size());
    for (int64_t t = 0; t < total; ++t)
        if (order[t].x == ti && order[t].y == tj) {
            for (int32_t r = 0; r < world; ++r) {
                int64_t b, e;
                grm_tile_range(total, r, world, &b, &e);
                if (t >= b && t < e) return r;
            }
        }
    return -1;
} GRM_CATCH_I64

// position t of the rank's tile list -> (ti, tj); lets a host build gather plans without a device
extern "C" int32_t grm_cuda_tile_at(int64_t n, int32_t rank, int32_t world, int64_t t, int64_t *ti, int64_t *tj) try {
    if (n < 1 || world < 1 || rank < 0 || rank >= world || !ti || !tj) return GRM_CUDA_ERR_BAD_ARGUMENT;
    std::vector<int2> order;
    grm_tile_order(n, 0, order);
    int64_t b, e;
    grm_tile_range(static_cast<int64_t>(order.size()), rank, world, &b, &e);
    if (t < 0 || b + t >= e) return GRM_CUDA_ERR_BAD_ARGUMENT;
    *ti = order[b + t].x;
    *tj = order[b + t].y;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

