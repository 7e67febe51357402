// api.cu — context, error plumbing, tile partition and the three reference-facing entry points of libgrm_cuda
// (grmsimple / grmploidyaware / inflatediagonals!, src/grm/calc.jl:115-142, :189-217, :41-70).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "common.cuh"

int32_t grm_repair_device(grm_cuda_ctx *ctx, double *X, int64_t n, int64_t ldx, int32_t max_iter, int32_t *iters,
                          double *eps_total);
extern "C" int32_t grm_cuda_scale_mirror(grm_cuda_ctx *ctx, double *d_G, int64_t n, int64_t ldg, double denom);
extern "C" int32_t grm_cuda_locus_stats_f64(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda,
                                            double *d_q, double *d_stats, uint32_t *d_flags, int32_t accumulate,
                                            int32_t skip_missing);
extern "C" int32_t grm_cuda_syrk_f64(grm_cuda_ctx *ctx, const double *dZ, int64_t n, int64_t p, int64_t ldz,
                                     int32_t accumulate, double *d_G, int64_t ldg, int32_t rank, int32_t world);
extern "C" int32_t grm_cuda_centre_f64(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t pc, int64_t lda,
                                       int32_t centred, double ploidy, const double *d_q, int32_t fill_missing,
                                       const uint8_t *d_keep, double *d_Z, int64_t ldz);
extern "C" int32_t grm_cuda_locus_qc(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda, double maf,
                                     double max_locus_sparsity, int32_t error_on_missing, double *d_q, uint8_t *d_keep,
                                     uint64_t *d_kept, double *d_stats, int32_t accumulate_stats, uint32_t *d_flags);
extern "C" int32_t grm_cuda_dosage_stats(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                         const int32_t *d_colsum, int32_t m, uint64_t *d_sums, int64_t *d_r,
                                         int64_t *d_T, int32_t accumulate);
extern "C" int32_t grm_cuda_dosage_centre(grm_cuda_ctx *ctx, const uint64_t *d_sums, const int64_t *d_r,
                                          const int64_t *d_T, int64_t n, int64_t p, int32_t m, int32_t ploidy,
                                          double *d_u, double *scalars);

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

extern "C" int32_t grm_cuda_ctx_create(int32_t device, grm_cuda_ctx **out) {
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
    // the compute stream outranks the auxiliary one: when a persistent SYRK and a pack kernel are both runnable
    // (pipelined flow, run_grm) the SYRK CTAs are placed first and the pack fills what is left of each SM
    int prio_least = 0, prio_greatest = 0;
    GRM_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    GRM_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest));
    GRM_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    GRM_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio_least));
    for (auto &ev : ctx->ev) GRM_CUDA_TRY(cudaEventCreate(&ev));
    for (int i = 0; i < 40; ++i)
        GRM_CUDA_TRY(i < 16 ? cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming)
                            : cudaEventCreate(&ctx->pipe_ev[i]));
    for (int i = 0; i < 2; ++i) {
        GRM_CUDA_TRY(cudaEventCreateWithFlags(&ctx->copy_done[i], cudaEventDisableTiming));
        GRM_CUDA_TRY(cudaEventCreateWithFlags(&ctx->stage_free[i], cudaEventDisableTiming));
    }
    *out = ctx;
    return GRM_CUDA_OK;
}

int32_t grm_cusolver_destroy(void *h);

extern "C" int32_t grm_cuda_ctx_destroy(grm_cuda_ctx *ctx) {
    if (!ctx) return GRM_CUDA_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->cusolver) grm_cusolver_destroy(ctx->cusolver);
    DevBuf *bufs[] = {&ctx->codes, &ctx->colsum, &ctx->flags, &ctx->q, &ctx->w, &ctx->u, &ctx->stats, &ctx->partials,
                      &ctx->G, &ctx->C, &ctx->stage[0], &ctx->stage[1], &ctx->lu, &ctx->lu_work, &ctx->ipiv, &ctx->misc,
                      &ctx->tiles_i8.buf, &ctx->tiles_i8_2cta.buf, &ctx->tiles_f64.buf, &ctx->tiles_packed.buf, &ctx->tiles_wide.buf, &ctx->planes,
                      &ctx->rT, &ctx->cnt, &ctx->keep, &ctx->misc2, &ctx->progress, &ctx->acc};
    for (DevBuf *b : bufs) b->release();
    for (auto &ev : ctx->ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->pipe_ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(ctx->copy_done[i]);
        cudaEventDestroy(ctx->stage_free[i]);
    }
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->aux_stream);
    delete ctx;
    return GRM_CUDA_OK;
}

extern "C" void *grm_cuda_ctx_stream(grm_cuda_ctx *ctx) { return ctx ? ctx->stream : nullptr; }

// ---------------------------------------------------------------------------------------------------------------
// Peer-visible device buffers for the one-process-per-GPU exchange (grm_cuda_syrk_i8_scatter): plain cudaMalloc
// allocations exported / opened through CUDA IPC handles (64 opaque bytes the host processes exchange themselves).
// ---------------------------------------------------------------------------------------------------------------
extern "C" int32_t grm_cuda_ipc_alloc(grm_cuda_ctx *ctx, int64_t bytes, void **dptr, uint8_t *handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!ctx || bytes < 1 || !dptr || !handle64) {
        grm_set_error("ipc_alloc: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    void *p = nullptr;
    const cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
    if (e != cudaSuccess) {
        cudaGetLastError();
        grm_set_error("ipc_alloc: cudaMalloc of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        return GRM_CUDA_ERR_OOM;
    }
    cudaIpcMemHandle_t h;
    const cudaError_t e2 = cudaIpcGetMemHandle(&h, p);
    if (e2 != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
        grm_set_error("ipc_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e2));
        return GRM_CUDA_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    *dptr = p;
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_ipc_open(grm_cuda_ctx *ctx, const uint8_t *handle64, void **dptr) {
    if (!ctx || !handle64 || !dptr) {
        grm_set_error("ipc_open: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    GRM_CUDA_TRY(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_ipc_close(grm_cuda_ctx *ctx, void *dptr) {
    if (!ctx || !dptr) return GRM_CUDA_ERR_BAD_ARGUMENT;
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_CUDA_TRY(cudaIpcCloseMemHandle(dptr));
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_ipc_free(grm_cuda_ctx *ctx, void *dptr) {
    if (!ctx || !dptr) return GRM_CUDA_ERR_BAD_ARGUMENT;
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_CUDA_TRY(cudaFree(dptr));
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_ctx_synchronize(grm_cuda_ctx *ctx) {
    if (!ctx) return GRM_CUDA_ERR_BAD_ARGUMENT;
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->aux_stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// tile partition: 256 x 256 tiles of the upper triangle, enumerated in bands of 8 tile rows so that tiles that are
// in flight together (and the contiguous range a rank owns) share few operand panels -> L2 reuse of the codes.
// ---------------------------------------------------------------------------------------------------------------
void grm_tile_order(int64_t n, std::vector<int2> &tiles) {
    const int T = static_cast<int>((n + GRM_TILE - 1) / GRM_TILE);
    constexpr int BAND = 8;
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

extern "C" int64_t grm_cuda_tile_count(int64_t n, int32_t rank, int32_t world) {
    if (n < 1 || world < 1 || rank < 0 || rank >= world) return -1;
    const int64_t T = (n + GRM_TILE - 1) / GRM_TILE;
    int64_t b, e;
    grm_tile_range(T * (T + 1) / 2, rank, world, &b, &e);
    return e - b;
}

extern "C" int32_t grm_cuda_tile_owner(int64_t n, int64_t ti, int64_t tj, int32_t world) {
    if (n < 1 || world < 1 || ti < 0 || tj < ti) return -1;
    std::vector<int2> order;
    grm_tile_order(n, order);
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
}

// position t of the rank's tile list -> (ti, tj); lets a host build gather plans without a device
extern "C" int32_t grm_cuda_tile_at(int64_t n, int32_t rank, int32_t world, int64_t t, int64_t *ti, int64_t *tj) {
    if (n < 1 || world < 1 || rank < 0 || rank >= world || !ti || !tj) return GRM_CUDA_ERR_BAD_ARGUMENT;
    std::vector<int2> order;
    grm_tile_order(n, order);
    int64_t b, e;
    grm_tile_range(static_cast<int64_t>(order.size()), rank, world, &b, &e);
    if (t < 0 || b + t >= e) return GRM_CUDA_ERR_BAD_ARGUMENT;
    *ti = order[b + t].x;
    *tj = order[b + t].y;
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// host <-> device streaming of loci chunks (a chunk = a contiguous column range = contiguous memory in Julia's
// column-major layout).  Two staging buffers; copies run on copy_stream and overlap the pack / SYRK of the
// previous chunk on the compute stream.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct ChunkFeeder {
    grm_cuda_ctx *ctx;
    const double *A;
    int64_t n, p, lda;
    bool host;
    int64_t chunk;
    int64_t nchunks;
    int issued = 0;  // host chunks whose copy has been enqueued

    int32_t init(grm_cuda_ctx *c, const double *A_, int64_t n_, int64_t p_, int64_t lda_, bool host_, int64_t want) {
        ctx = c;
        A = A_;
        n = n_;
        p = p_;
        lda = lda_;
        host = host_;
        if (!host) {
            chunk = p;
            nchunks = 1;
            return GRM_CUDA_OK;
        }
        int64_t ch = want;
        if (ch <= 0) ch = std::max<int64_t>(128, ((256ll << 20) / (n * 8)) / 128 * 128);
        ch = std::max<int64_t>(128, ch / 128 * 128);
        chunk = std::min(ch, (p + 127) / 128 * 128);
        nchunks = (p + chunk - 1) / chunk;
        const size_t bytes = static_cast<size_t>(n) * static_cast<size_t>(std::min(chunk, p)) * sizeof(double);
        GRM_TRY(ctx->stage[0].reserve(bytes));
        if (nchunks > 1) GRM_TRY(ctx->stage[1].reserve(bytes));
        return GRM_CUDA_OK;
    }
    int64_t k0(int64_t c) const { return c * chunk; }
    int64_t len(int64_t c) const { return std::min(chunk, p - c * chunk); }
    int64_t ld_dev() const { return host ? n : lda; }

    int32_t enqueue_copy(int64_t c) {
        const int b = static_cast<int>(c & 1);
        if (c >= 2) GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[b], 0));
        const double *src = A + k0(c) * lda;
        double *dst = ctx->stage[b].as<double>();
        if (lda == n)
            GRM_CUDA_TRY(cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * len(c) * sizeof(double),
                                         cudaMemcpyHostToDevice, ctx->copy_stream));
        else
            GRM_CUDA_TRY(cudaMemcpy2DAsync(dst, n * sizeof(double), src, lda * sizeof(double), n * sizeof(double),
                                           static_cast<size_t>(len(c)), cudaMemcpyHostToDevice, ctx->copy_stream));
        GRM_CUDA_TRY(cudaEventRecord(ctx->copy_done[b], ctx->copy_stream));
        return GRM_CUDA_OK;
    }
    // device pointer of chunk c, valid on the compute stream; prefetches chunk c+1
    int32_t acquire(int64_t c, const double **d) {
        if (!host) {
            *d = A;
            return GRM_CUDA_OK;
        }
        while (issued <= c + 1 && issued < nchunks) {
            GRM_TRY(enqueue_copy(issued));
            ++issued;
        }
        const int b = static_cast<int>(c & 1);
        GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->copy_done[b], 0));
        *d = ctx->stage[b].as<double>();
        return GRM_CUDA_OK;
    }
    int32_t release(int64_t c) {
        if (!host) return GRM_CUDA_OK;
        GRM_CUDA_TRY(cudaEventRecord(ctx->stage_free[c & 1], ctx->stream));
        return GRM_CUDA_OK;
    }
};

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    float ms() const {
        return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

int32_t flags_to_status(uint32_t f) {
    if (f & 8u) {
        grm_set_error("a locus has no valid (non-missing) cell: its mean fill is undefined");
        return GRM_CUDA_ERR_MISSING;
    }
    if (f & 1u) {
        grm_set_error("allele_frequencies contains missing (NaN) values; the reference GRM functions throw on missing data "
                      "(src/grm/calc.jl:127,205) — filter or impute first");
        return GRM_CUDA_ERR_MISSING;
    }
    if (f & 2u) {
        grm_set_error("allele_frequencies contains +-Inf");
        return GRM_CUDA_ERR_NONFINITE;
    }
    return GRM_CUDA_OK;
}

// launches of the staged entry points go to ctx->stream; the pipelined flow redirects some of them to the auxiliary
// stream for the lifetime of this guard
struct StreamSwap {
    grm_cuda_ctx *ctx;
    cudaStream_t saved;
    bool saved_small;
    StreamSwap(grm_cuda_ctx *c, cudaStream_t s, bool small_pack) : ctx(c), saved(c->stream), saved_small(c->pack_small_footprint) {
        ctx->stream = s;
        ctx->pack_small_footprint = small_pack;
    }
    ~StreamSwap() {
        ctx->stream = saved;
        ctx->pack_small_footprint = saved_small;
    }
};

int32_t run_grm(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda, bool centred, int32_t ploidy,
                double *G, int64_t ldg, const grm_cuda_options *uopts, grm_cuda_info *info) {
    Timer wall;
    grm_cuda_options o;
    grm_cuda_options_default(&o);
    if (uopts) memcpy(&o, uopts, std::min<size_t>(sizeof(o), static_cast<size_t>(std::max(uopts->struct_size, 0))));
    if (!ctx || !A || !G || n < 1 || p < 1 || lda < n || ldg < n || o.world < 1 || o.rank < 0 || o.rank >= o.world ||
        o.max_iter < 0) {
        grm_set_error("bad argument (n=%lld p=%lld lda=%lld ldg=%lld rank=%d world=%d)", (long long)n, (long long)p,
                      (long long)lda, (long long)ldg, o.rank, o.world);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (!o.skip_repair && o.world == 1 && n > 46000) {
        grm_set_error("inflatediagonals! on n = %lld exceeds the single-GPU cuSOLVER range (n <= 46000); pass skip_repair "
                      "to obtain the raw GRM", (long long)n);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (n > INT32_MAX - 512 || p > INT32_MAX - 512) {
        grm_set_error("n and p must fit TMA's 32-bit coordinates");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    const bool in_host = o.input_location == GRM_CUDA_LOC_HOST;
    const bool out_host = o.output_location == GRM_CUDA_LOC_HOST;
    const bool sharded = o.world > 1;
    const bool meanfill = o.missing_policy == GRM_CUDA_MISSING_MEANFILL;
    const bool masked = o.maf > 0.0 || o.max_locus_sparsity < 1.0;
    if (!(o.maf >= 0.0 && o.maf <= 0.5) || !(o.max_locus_sparsity >= 0.0 && o.max_locus_sparsity <= 1.0)) {
        grm_set_error("maf must be in [0, 0.5] and max_locus_sparsity in [0, 1]");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    uint8_t *d_keep = nullptr;
    uint64_t *d_kept = nullptr;
    if (masked) {
        GRM_TRY(ctx->keep.reserve(static_cast<size_t>(p) + 64));
        d_keep = ctx->keep.as<uint8_t>() + 16;
        d_kept = ctx->keep.as<uint64_t>();
        GRM_TRY(ctx->q.reserve(static_cast<size_t>((p + 127) / 128 * 128) * sizeof(double)));
    }
    int64_t p_eff = p;  // loci that enter the GRM
    cudaStream_t st = ctx->stream;
    cudaEvent_t *ev = ctx->ev;

    grm_cuda_info inf;
    memset(&inf, 0, sizeof(inf));
    inf.struct_size = static_cast<int32_t>(sizeof(inf));

    // output buffer on the device
    double *dG = G;
    int64_t ldG = ldg;
    if (out_host) {
        GRM_TRY(ctx->G.reserve(static_cast<size_t>(n) * n * sizeof(double)));
        dG = ctx->G.as<double>();
        ldG = n;
        if (sharded) GRM_CUDA_TRY(cudaMemsetAsync(dG, 0, static_cast<size_t>(n) * n * sizeof(double), st));
    }

    ChunkFeeder feed;
    GRM_TRY(feed.init(ctx, A, n, p, lda, in_host, o.chunk_loci));

    GRM_TRY(ctx->flags.reserve(64));
    uint32_t *d_flags = ctx->flags.as<uint32_t>();
    GRM_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 16 * sizeof(uint32_t), st));
    GRM_TRY(ctx->stats.reserve(8 * sizeof(double)));
    double *d_stats = ctx->stats.as<double>();

    GRM_CUDA_TRY(cudaEventRecord(ev[0], st));

    // ---- path decision: dosage denominator "detected on load" (first chunk), unless given ---------------------
    int32_t m = 0;
    int path = o.path;
    if (path != GRM_CUDA_PATH_F64) {
        m = o.denominator;
        if (m == 0) {
            const double *d0;
            GRM_TRY(feed.acquire(0, &d0));
            // a sample decides (the first 2^22 cells); the pack pass verifies every cell against it and the loop
            // below retries with m = 64 / falls back to FP64 if a later locus disagrees
            GRM_TRY(grm_cuda_detect_denominator(ctx, d0, n, feed.len(0), feed.ld_dev(), 1ll << 22, masked ? 1 : 0, &m));
        } else if (m < 1 || m > 64 || (m & (m - 1)) != 0) {
            grm_set_error("denominator must be a power of two in [1,64], got %d", m);
            return GRM_CUDA_ERR_BAD_ARGUMENT;
        }
        if (m == 0) {
            if (path == GRM_CUDA_PATH_I8) {
                grm_set_error("int8 path requested but the allele frequencies are not multiples of 1/m, m = 2^k <= 64");
                return GRM_CUDA_ERR_NOT_DOSAGE;
            }
            path = GRM_CUDA_PATH_F64;
        } else {
            path = GRM_CUDA_PATH_I8;
        }
    }

    // pipelined int8 flow, opt-in (GRM_CUDA_PIPELINE = number of loci chunks, 2..8) for device-resident, unmasked,
    // single-GPU calls.  Off by default: measured on C3 it is bit-identical and 2.7 % faster in short bursts (71.1 vs
    // 73.1 ms) but 3 % slower in a sustained, power-capped loop (76.1 vs 73.8 ms) — a pack CTA sharing an SM with a
    // persistent SYRK CTA runs at 1/3 of its stand-alone rate (the MMA operand reads own the shared-memory / L1 data
    // path), so 3/4 of the pack ends up on the critical path anyway (profiles/README.md).
    int pipe_chunks = 1, pipelined_done = 0;
    if (!in_host && !masked && !meanfill && !sharded && grm_syrk_i8_pairs() && p >= 8 * 128) {
        const char *e = getenv("GRM_CUDA_PIPELINE");
        pipe_chunks = e ? std::max(1, std::min(atoi(e), 8)) : 1;
    }

    double denom = 0.0;
    bool done = false;
    while (!done) {
        if (path == GRM_CUDA_PATH_I8) {
            if (static_cast<double>(m) * m * static_cast<double>(p) >= 2147483648.0) {
                grm_set_error("m^2 * p = %d^2 * %lld overflows the int32 accumulators", m, (long long)p);
                return GRM_CUDA_ERR_OVERFLOW;
            }
            const int64_t ldc = (p + 127) / 128 * 128;
            GRM_TRY(ctx->codes.reserve(static_cast<size_t>(n) * ldc));
            GRM_TRY(ctx->colsum.reserve(static_cast<size_t>(ldc) * sizeof(int32_t)));
            int8_t *d_codes = ctx->codes.as<int8_t>();
            int32_t *d_colsum = ctx->colsum.as<int32_t>();
            GRM_CUDA_TRY(cudaMemsetAsync(d_colsum, 0, static_cast<size_t>(ldc) * sizeof(int32_t), st));
            if (masked) GRM_CUDA_TRY(cudaMemsetAsync(d_kept, 0, sizeof(uint64_t), st));
            if (pipe_chunks > 1) {
                // ---- pipelined flow (device-resident input): the loci are cut into chunks; the pack of chunk c + 1
                // (and, at the end, the integer statistics) runs on the low-priority auxiliary stream while the
                // tensor cores contract chunk c.  Partial sums live as packed int32 tiles; the last chunk's launch adds
                // them in its fused FP64 epilogue, so G is still written once.  Same integers, same epilogue
                // arithmetic: bit-identical to the serial flow.
                const int nk = pipe_chunks;
                cudaStream_t aux = ctx->aux_stream;
                cudaEvent_t *pe = ctx->pipe_ev;
                GRM_TRY(ctx->acc.reserve(static_cast<size_t>(grm_syrk_i8_partial_bytes(n))));
                int32_t *d_acc = ctx->acc.as<int32_t>();
                GRM_CUDA_TRY(cudaEventRecord(pe[15], st));
                GRM_CUDA_TRY(cudaStreamWaitEvent(aux, pe[15], 0));
                auto cut = [&](int c) { return c >= nk ? p : (p * c / nk) / 128 * 128; };
                int32_t rc = GRM_CUDA_OK;
                for (int c = 0; c < nk && rc == GRM_CUDA_OK; ++c) {
                    const int64_t k0 = cut(c), len = cut(c + 1) - k0;
                    {
                        StreamSwap on_aux(ctx, aux, /*small footprint pack*/ c > 0);
                        if (c == 0) GRM_CUDA_TRY(cudaEventRecord(pe[32], aux));
                        rc = grm_cuda_pack_dosage(ctx, A + k0 * lda, n, len, lda, m, d_codes + k0, ldc, d_colsum + k0,
                                                  d_flags, nullptr);
                        if (rc != GRM_CUDA_OK) break;
                        GRM_CUDA_TRY(cudaEventRecord(pe[c], aux));
                        if (c == nk - 1) GRM_CUDA_TRY(cudaEventRecord(pe[33], aux));
                    }
                    if (c < nk - 1) {
                        GRM_CUDA_TRY(cudaStreamWaitEvent(st, pe[c], 0));
                        GRM_CUDA_TRY(cudaEventRecord(pe[16 + 2 * c], st));
                        rc = grm_syrk_i8_partial(ctx, d_codes + k0, n, len, ldc, d_acc, c > 0);
                        GRM_CUDA_TRY(cudaEventRecord(pe[17 + 2 * c], st));
                    }
                }
                uint32_t h_flags = 0;
                double sc[4] = {1.0 / (static_cast<double>(m) * m), 0.0, 0.0, static_cast<double>(p)};
                if (rc == GRM_CUDA_OK) {
                    StreamSwap on_aux(ctx, aux, false);
                    GRM_CUDA_TRY(cudaMemcpyAsync(&h_flags, d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, aux));
                    GRM_CUDA_TRY(cudaEventRecord(pe[34], aux));
                    if (centred) {
                        // exact integer statistics -> u, W2, d, hidden under the SYRK of chunk nk - 2
                        rc = ctx->u.reserve(static_cast<size_t>(n) * sizeof(double));
                        if (rc == GRM_CUDA_OK) rc = ctx->rT.reserve(static_cast<size_t>(2 * n + 2) * sizeof(int64_t));
                        uint64_t *d_sums = ctx->rT.as<uint64_t>();
                        int64_t *d_r = ctx->rT.as<int64_t>() + 2, *d_T = d_r + n;
                        if (rc == GRM_CUDA_OK)
                            rc = grm_cuda_dosage_stats(ctx, d_codes, n, p, ldc, d_colsum, m, d_sums, d_r, d_T, 0);
                        if (rc == GRM_CUDA_OK)
                            rc = grm_cuda_dosage_centre(ctx, d_sums, d_r, d_T, n, p, m, ploidy, ctx->u.as<double>(), sc);
                    }
                    GRM_CUDA_TRY(cudaEventRecord(pe[35], aux));
                    GRM_CUDA_TRY(cudaStreamSynchronize(aux));
                }
                if (rc == GRM_CUDA_OK) rc = flags_to_status(h_flags);
                if (rc != GRM_CUDA_OK) {
                    cudaStreamSynchronize(aux);
                    cudaStreamSynchronize(st);
                    return rc;
                }
                if (h_flags & 4u) {
                    // a later locus broke the denominator guessed from the sample: same policy as the serial flow
                    if (o.path == GRM_CUDA_PATH_I8 || o.denominator != 0) {
                        cudaStreamSynchronize(st);
                        grm_set_error("allele frequencies are not all multiples of 1/%d", m);
                        return GRM_CUDA_ERR_NOT_DOSAGE;
                    }
                    if (m < 64) m = 64; else path = GRM_CUDA_PATH_F64;
                    GRM_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 16 * sizeof(uint32_t), st));
                    continue;
                }
                denom = sc[3];  // calc.jl:127 / :201
                const int64_t k0 = cut(nk - 1), len = p - k0;
                GRM_CUDA_TRY(cudaEventRecord(ev[1], st));
                GRM_CUDA_TRY(cudaEventRecord(ev[2], st));
                GRM_CUDA_TRY(cudaEventRecord(pe[16 + 2 * (nk - 1)], st));
                GRM_TRY(grm_syrk_i8_f64_prev(ctx, d_codes + k0, n, len, ldc, 1, sc[0], sc[1], sc[2], denom,
                                             centred ? ctx->u.as<double>() : nullptr, dG, ldG, 0, 1, d_acc));
                GRM_CUDA_TRY(cudaEventRecord(pe[17 + 2 * (nk - 1)], st));
                GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
                GRM_TRY(grm_cuda_scale_mirror(ctx, dG, n, ldG, 1.0));
                pipelined_done = nk;
                done = true;
                continue;
            }
            for (int64_t c = 0; c < feed.nchunks; ++c) {
                const double *dc;
                GRM_TRY(feed.acquire(c, &dc));
                if (masked)  // the mask of a locus depends on its own column only: decide it while the chunk is staged
                    GRM_TRY(grm_cuda_locus_qc(ctx, dc, n, feed.len(c), feed.ld_dev(), o.maf, o.max_locus_sparsity, 0,
                                              ctx->q.as<double>() + feed.k0(c), d_keep + feed.k0(c), d_kept, nullptr, 0,
                                              d_flags));
                GRM_TRY(grm_cuda_pack_dosage(ctx, dc, n, feed.len(c), feed.ld_dev(), m, d_codes + feed.k0(c), ldc,
                                             d_colsum + feed.k0(c), d_flags, masked ? d_keep + feed.k0(c) : nullptr));
                GRM_TRY(feed.release(c));
            }
            GRM_CUDA_TRY(cudaEventRecord(ev[1], st));
            uint32_t h_flags = 0;
            GRM_CUDA_TRY(cudaMemcpyAsync(&h_flags, d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            if (masked) {
                uint64_t h_kept = 0;
                GRM_CUDA_TRY(cudaMemcpyAsync(&h_kept, d_kept, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
                GRM_CUDA_TRY(cudaStreamSynchronize(st));
                p_eff = static_cast<int64_t>(h_kept);
                if (p_eff == 0) {
                    grm_set_error("All loci filtered out at maf = %g, maximum locus sparsity = %g.", o.maf,
                                  o.max_locus_sparsity);
                    return GRM_CUDA_ERR_ALL_FILTERED;
                }
            }
            GRM_CUDA_TRY(cudaStreamSynchronize(st));
            if (meanfill && (h_flags & 1u) && !(h_flags & 2u)) {
                // missing cells + mean-fill policy: the filled values are not dosages -> FP64 path
                if (o.path == GRM_CUDA_PATH_I8) {
                    grm_set_error("int8 path requested but missing cells must be mean-filled (FP64 path)");
                    return GRM_CUDA_ERR_NOT_DOSAGE;
                }
                path = GRM_CUDA_PATH_F64;
                GRM_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 16 * sizeof(uint32_t), st));
                feed.issued = 0;
                continue;
            }
            GRM_TRY(flags_to_status(h_flags));
            if (h_flags & 4u) {
                // a later chunk broke the denominator guessed from the first one
                if (o.path == GRM_CUDA_PATH_I8 || o.denominator != 0) {
                    grm_set_error("allele frequencies are not all multiples of 1/%d", m);
                    return GRM_CUDA_ERR_NOT_DOSAGE;
                }
                if (m < 64) {
                    m = 64;  // retry once with the finest supported grid, then give up on int8
                } else {
                    path = GRM_CUDA_PATH_F64;
                }
                GRM_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 16 * sizeof(uint32_t), st));
                feed.issued = 0;
                continue;
            }
            if (!centred) {
                denom = static_cast<double>(p_eff);  // calc.jl:127
                GRM_CUDA_TRY(cudaEventRecord(ev[2], st));
                GRM_TRY(grm_cuda_syrk_i8_f64(ctx, d_codes, n, p, ldc, 1, 1.0 / (static_cast<double>(m) * m), 0.0, 0.0,
                                             denom, nullptr, dG, ldG, o.rank, o.world));
            } else {
                // exact integer statistics -> u, W2, d: no FP64 pass over the loci at all
                GRM_TRY(ctx->u.reserve(static_cast<size_t>(n) * sizeof(double)));
                GRM_TRY(ctx->rT.reserve(static_cast<size_t>(2 * n + 2) * sizeof(int64_t)));
                uint64_t *d_sums = ctx->rT.as<uint64_t>();
                int64_t *d_r = ctx->rT.as<int64_t>() + 2, *d_T = d_r + n;
                GRM_TRY(grm_cuda_dosage_stats(ctx, d_codes, n, p, ldc, d_colsum, m, d_sums, d_r, d_T, 0));
                double sc[4];
                GRM_TRY(grm_cuda_dosage_centre(ctx, d_sums, d_r, d_T, n, p_eff, m, ploidy, ctx->u.as<double>(), sc));
                denom = sc[3];  // calc.jl:201
                GRM_CUDA_TRY(cudaEventRecord(ev[2], st));
                GRM_TRY(grm_cuda_syrk_i8_f64(ctx, d_codes, n, p, ldc, 1, sc[0], sc[1], sc[2], denom, ctx->u.as<double>(),
                                             dG, ldG, o.rank, o.world));
            }
            GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
            if (!sharded) GRM_TRY(grm_cuda_scale_mirror(ctx, dG, n, ldG, 1.0));
            done = true;
        } else {
            // ---- FP64 path: per chunk, column means (+ validity) then accumulate Z_c * Z_c' into G ------------
            GRM_TRY(ctx->q.reserve(static_cast<size_t>((p + 127) / 128 * 128) * sizeof(double)));
            if (masked) GRM_CUDA_TRY(cudaMemsetAsync(d_kept, 0, sizeof(uint64_t), st));
            for (int64_t c = 0; c < feed.nchunks; ++c) {
                const double *dc;
                GRM_TRY(feed.acquire(c, &dc));
                if (masked)
                    GRM_TRY(grm_cuda_locus_qc(ctx, dc, n, feed.len(c), feed.ld_dev(), o.maf, o.max_locus_sparsity,
                                              meanfill ? 0 : 1, ctx->q.as<double>() + feed.k0(c), d_keep + feed.k0(c),
                                              d_kept, d_stats, c > 0, d_flags));
                else
                    GRM_TRY(grm_cuda_locus_stats_f64(ctx, dc, n, feed.len(c), feed.ld_dev(),
                                                     ctx->q.as<double>() + feed.k0(c), d_stats, d_flags, c > 0,
                                                     meanfill ? 1 : 0));
                const double *qc = ctx->q.as<double>() + feed.k0(c);
                const uint8_t *kc = masked ? d_keep + feed.k0(c) : nullptr;
                if (!(centred || meanfill || masked)) {
                    // grmsimple on clean data: the operands are the input itself
                    GRM_TRY(grm_cuda_syrk_f64(ctx, dc, n, feed.len(c), feed.ld_dev(), c > 0, dG, ldG, o.rank, o.world));
                } else if (in_host) {
                    // the staged chunk is ours: centre / fill / mask it in place, then a pure-DMMA SYRK
                    double *z = const_cast<double *>(dc);
                    GRM_TRY(grm_cuda_centre_f64(ctx, dc, n, feed.len(c), n, centred ? 1 : 0, static_cast<double>(ploidy),
                                                qc, meanfill ? 1 : 0, kc, z, n));
                    GRM_TRY(grm_cuda_syrk_f64(ctx, z, n, feed.len(c), n, c > 0, dG, ldG, o.rank, o.world));
                } else {
                    // device-resident caller buffer (read-only): transformed sub-chunks go through a scratch buffer
                    const int64_t zc = std::max<int64_t>(32, std::min<int64_t>(feed.len(c), (512ll << 20) / (n * 8)));
                    GRM_TRY(ctx->stage[0].reserve(static_cast<size_t>(n) * zc * sizeof(double)));
                    double *z = ctx->stage[0].as<double>();
                    for (int64_t s0 = 0; s0 < feed.len(c); s0 += zc) {
                        const int64_t sl = std::min(zc, feed.len(c) - s0);
                        GRM_TRY(grm_cuda_centre_f64(ctx, dc + s0 * feed.ld_dev(), n, sl, feed.ld_dev(), centred ? 1 : 0,
                                                    static_cast<double>(ploidy), qc + s0, meanfill ? 1 : 0,
                                                    kc ? kc + s0 : nullptr, z, n));
                        GRM_TRY(grm_cuda_syrk_f64(ctx, z, n, sl, n, (c > 0 || s0 > 0) ? 1 : 0, dG, ldG, o.rank, o.world));
                    }
                }
                GRM_TRY(feed.release(c));
            }
            GRM_CUDA_TRY(cudaEventRecord(ev[1], st));
            GRM_CUDA_TRY(cudaEventRecord(ev[2], st));
            uint32_t h_flags = 0;
            double h_stats[3];
            GRM_CUDA_TRY(cudaMemcpyAsync(&h_flags, d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            GRM_CUDA_TRY(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
            uint64_t h_kept = static_cast<uint64_t>(p);
            if (masked) GRM_CUDA_TRY(cudaMemcpyAsync(&h_kept, d_kept, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            GRM_CUDA_TRY(cudaStreamSynchronize(st));
            p_eff = static_cast<int64_t>(h_kept);
            if (masked && p_eff == 0) {
                grm_set_error("All loci filtered out at maf = %g, maximum locus sparsity = %g.", o.maf,
                              o.max_locus_sparsity);
                return GRM_CUDA_ERR_ALL_FILTERED;
            }
            GRM_TRY(flags_to_status(h_flags));
            denom = centred ? static_cast<double>(ploidy) * h_stats[0] : static_cast<double>(p_eff);
            GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
            if (!sharded) GRM_TRY(grm_cuda_scale_mirror(ctx, dG, n, ldG, denom));
            // sharded FP64 calls return raw sums scaled by the caller (it owns the exchange of d)
            done = true;
        }
    }
    GRM_CUDA_TRY(cudaEventRecord(ev[4], st));

    // ---- inflatediagonals! (calc.jl:133 / :211) -----------------------------------------------------------------
    // The repair only ever touches the diagonal, so the bulk device->host copy of G runs on the copy stream WHILE the
    // O(n^3) factorisations run; if a round of inflation happened, the n diagonal values are copied again afterwards.
    const bool do_repair = !o.skip_repair && !sharded;
    bool bulk_started = false;
    auto copy_out = [&](cudaStream_t s) -> int32_t {
        if (ldg == n)
            GRM_CUDA_TRY(cudaMemcpyAsync(G, dG, static_cast<size_t>(n) * n * sizeof(double), cudaMemcpyDeviceToHost, s));
        else
            GRM_CUDA_TRY(cudaMemcpy2DAsync(G, ldg * sizeof(double), dG, n * sizeof(double), n * sizeof(double),
                                           static_cast<size_t>(n), cudaMemcpyDeviceToHost, s));
        return GRM_CUDA_OK;
    };
    if (out_host && do_repair) {
        GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ev[4], 0));
        GRM_CUDA_TRY(cudaEventRecord(ev[7], ctx->copy_stream));
        GRM_TRY(copy_out(ctx->copy_stream));
        GRM_CUDA_TRY(cudaEventRecord(ev[8], ctx->copy_stream));
        bulk_started = true;
    }
    if (do_repair) {
        const int32_t rst = grm_repair_device(ctx, dG, n, ldG, o.max_iter, &inf.repair_iters, &inf.eps_total);
        if (rst != GRM_CUDA_OK) {
            cudaStreamSynchronize(ctx->copy_stream);  // never leave a copy into the caller's buffer in flight
            return rst;
        }
    }
    GRM_CUDA_TRY(cudaEventRecord(ev[5], st));
    if (out_host) {
        if (bulk_started) {
            if (inf.repair_iters > 0) {
                // diagonal only: n elements with a pitch of (ld + 1) doubles on both sides
                GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ev[5], 0));
                GRM_CUDA_TRY(cudaMemcpy2DAsync(G, (ldg + 1) * sizeof(double), dG, (ldG + 1) * sizeof(double),
                                               sizeof(double), static_cast<size_t>(n), cudaMemcpyDeviceToHost,
                                               ctx->copy_stream));
            }
            GRM_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
        } else {
            GRM_CUDA_TRY(cudaEventRecord(ev[7], st));
            GRM_TRY(copy_out(st));
            GRM_CUDA_TRY(cudaEventRecord(ev[8], st));
        }
    }
    GRM_CUDA_TRY(cudaEventRecord(ev[6], st));
    GRM_CUDA_TRY(cudaStreamSynchronize(st));

    inf.path = path;
    inf.denominator = path == GRM_CUDA_PATH_I8 ? m : 0;
    inf.scale = denom;
    inf.tiles = grm_cuda_tile_count(n, o.rank, o.world);
    cudaEventElapsedTime(&inf.ms_pack, ev[0], ev[1]);
    cudaEventElapsedTime(&inf.ms_rowdot, ev[1], ev[2]);
    cudaEventElapsedTime(&inf.ms_syrk, ev[2], ev[3]);
    if (pipelined_done > 0) {
        // overlapping stages: each is timed on its own stream (pack and statistics on the auxiliary stream, the SYRK
        // launches on the compute stream), so the stage times sum to more than the call
        cudaEventElapsedTime(&inf.ms_pack, ctx->pipe_ev[32], ctx->pipe_ev[33]);
        cudaEventElapsedTime(&inf.ms_rowdot, ctx->pipe_ev[34], ctx->pipe_ev[35]);
        inf.ms_syrk = 0.f;
        for (int c = 0; c < pipelined_done; ++c) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ctx->pipe_ev[16 + 2 * c], ctx->pipe_ev[17 + 2 * c]);
            inf.ms_syrk += t;
        }
    }
    cudaEventElapsedTime(&inf.ms_finalize, ev[3], ev[4]);
    cudaEventElapsedTime(&inf.ms_repair, ev[4], ev[5]);
    if (out_host) cudaEventElapsedTime(&inf.ms_d2h, ev[7], ev[8]);
    inf.ms_h2d = in_host ? inf.ms_pack : 0.f;  // host input: the ingest stage is the H2D stream with the pack hidden under it
    inf.ms_total = wall.ms();
    inf.kernel_launches = ctx->launches;
    inf.loci_kept = static_cast<int32_t>(p_eff);
    if (info) memcpy(info, &inf, std::min<size_t>(sizeof(inf), static_cast<size_t>(std::max(info->struct_size, 0))));
    return GRM_CUDA_OK;
}

}  // namespace

extern "C" int32_t grm_cuda_simple(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda, double *G,
                                   int64_t ldg, const grm_cuda_options *opts, grm_cuda_info *info) {
    return run_grm(ctx, A, n, p, lda, false, 0, G, ldg, opts, info);
}

extern "C" int32_t grm_cuda_ploidyaware(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda,
                                        int32_t ploidy, double *G, int64_t ldg, const grm_cuda_options *opts,
                                        grm_cuda_info *info) {
    return run_grm(ctx, A, n, p, lda, true, ploidy, G, ldg, opts, info);
}

extern "C" int32_t grm_cuda_inflate_diagonals(grm_cuda_ctx *ctx, double *X, int64_t n, int64_t ldx, int32_t max_iter,
                                              int32_t location, int32_t *iters, double *eps_total) {
    if (!ctx || !X || n < 1 || ldx < n || max_iter < 0) {
        grm_set_error("inflate_diagonals: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    int32_t it = 0;
    double tot = 0.0;
    int32_t status;
    if (location == GRM_CUDA_LOC_DEVICE) {
        status = grm_repair_device(ctx, X, n, ldx, max_iter, &it, &tot);
    } else {
        GRM_TRY(ctx->G.reserve(static_cast<size_t>(n) * n * sizeof(double)));
        double *d = ctx->G.as<double>();
        GRM_CUDA_TRY(cudaMemcpy2DAsync(d, n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double),
                                       static_cast<size_t>(n), cudaMemcpyHostToDevice, ctx->stream));
        status = grm_repair_device(ctx, d, n, n, max_iter, &it, &tot);
        if (status == GRM_CUDA_OK && it > 0) {
            GRM_CUDA_TRY(cudaMemcpy2DAsync(X, ldx * sizeof(double), d, n * sizeof(double), n * sizeof(double),
                                           static_cast<size_t>(n), cudaMemcpyDeviceToHost, ctx->stream));
        }
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    if (iters) *iters = it;
    if (eps_total) *eps_total = tot;
    return status;
}
