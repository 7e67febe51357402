// run.cu — the reference-facing entry points of libgrm_cuda:
//   grmsimple(genomes)        src/grm/calc.jl:115-142      grm_cuda_simple / _simple_union / _from_codes
//   grmploidyaware(genomes)   src/grm/calc.jl:189-217      grm_cuda_ploidyaware / _ploidyaware_union / _from_codes
//   inflatediagonals!(X)      src/grm/calc.jl:41-70        grm_cuda_inflate_diagonals
//
// One call = ingest (host threads -> pinned ring -> PCIe -> pack, or device-resident input) -> per-locus statistics ->
// SYRK -> exchange (distributed calls) -> inflatediagonals! -> output.  Every stage is enqueued on the context's
// streams; the host only blocks where the reference's control flow needs a value (path decision, error flags, det).
//
// Distributed calls (opts.distributed, one process per GPU, the communicator of comm.cu): every decision that changes
// the sequence of collectives (dosage denominator, retry with m = 64, fall back to FP64, error flags) is taken on
// values agreed by an all-reduce, so all ranks always walk the same branch.
#include <float.h>
#include <math.h>
#include <sched.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <chrono>
#include <thread>

#include "common.cuh"
#include "ingest_host.h"

int32_t grm_repair_device(grm_cuda_ctx *ctx, double *X, int64_t n, int64_t ldx, int32_t max_iter, int32_t *iters,
                          double *eps_total);

namespace {

constexpr uint32_t F_MISSING = 1u, F_INF = 2u, F_NOT_DOSAGE = 4u, F_EMPTY_LOCUS = 8u, F_NAN_PAYLOAD = 16u;
constexpr int RING = 4;  // pinned host slots / device staging buffers of the ingest

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    float ms() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// where the allele frequencies of THIS RANK's loci range come from (pointers already offset to its first locus)
struct Source {
    const double *A = nullptr;      // FP64 cells, host or device
    const uint8_t *sel = nullptr;   // host selector bytes of Matrix{Union{Float64,Missing}}, or null
    int float_tag = 0;
    const int8_t *codes = nullptr;  // pre-packed host codes (grm_cuda_from_codes)
    int64_t lda = 0;
    bool host = false;
};

int host_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int c = CPU_COUNT(&set);
        if (c > 0) return c;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? static_cast<int>(h) : 1;
}

int32_t ensure_pool(grm_cuda_ctx *ctx, int world) {
    if (ctx->pool) return GRM_CUDA_OK;
    int t = ctx->tune.host_threads;
    if (t <= 0) t = std::max(1, std::min(64, host_cpus() / std::max(1, world)));  // ranks of one node share its cores
    ctx->pool = grm_pool_create(t);
    if (!ctx->pool) {
        grm_set_error("could not start %d host ingest threads", t);
        return GRM_CUDA_ERR_OOM;
    }
    return GRM_CUDA_OK;
}

int32_t ensure_ring(grm_cuda_ctx *ctx, size_t slot_bytes) {
    if (ctx->pin[0] && ctx->pin_cap >= slot_bytes) return GRM_CUDA_OK;
    for (auto &pp : ctx->pin) {
        if (pp) cudaFreeHost(pp);
        pp = nullptr;
    }
    ctx->pin_cap = 0;
    for (auto &pp : ctx->pin) {
        const cudaError_t e = cudaHostAlloc(&pp, slot_bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            grm_set_error("cudaHostAlloc of a %zu-byte ingest slot failed: %s", slot_bytes, cudaGetErrorString(e));
            return GRM_CUDA_ERR_OOM;
        }
    }
    ctx->pin_cap = slot_bytes;
    return GRM_CUDA_OK;
}

bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// ---------------------------------------------------------------------------------------------------------------
// The host half of the ingest as ONE continuous job for the whole pass: the loci are cut into ring chunks and every
// chunk into small units (a few loci, ~1.5 MB of source); the workers pull units from a global counter in chunk order,
// so there is no barrier between chunks and a slow core simply takes fewer units.  A worker may only write into the
// ring slot of chunk c once the copy of chunk c - RING out of that slot has finished (`writable`, advanced by the thread
// that drives the GPU); the workers therefore run up to RING - 1 chunks ahead of the device.
// ---------------------------------------------------------------------------------------------------------------
struct StreamJob {
    const Source *src = nullptr;
    int64_t n = 0, pc = 0, chunk = 0, nchunks = 0, ldo = 0, grain = 1, total_units = 0;
    int m = 0, mode = 0;  // mode: 0 quantise FP64 -> int8, 1 copy pre-packed codes, 2 copy FP64
    size_t elt = 1;
    void *pin[RING] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<int64_t> unit_base;  // first unit of chunk c; unit_base[nchunks] = total
    std::atomic<int64_t> next{0};
    std::unique_ptr<std::atomic<int32_t>[]> done;  // units finished per chunk
    std::atomic<int64_t> writable{RING};
    std::atomic<uint32_t> flags{0};
    std::atomic<int> abort{0};

    int64_t chunk_len(int64_t c) const { return std::min(chunk, pc - c * chunk); }
    int32_t units_of(int64_t c) const { return static_cast<int32_t>(unit_base[c + 1] - unit_base[c]); }
};

void stream_worker(void *arg, int, int) {
    StreamJob *j = static_cast<StreamJob *>(arg);
    const Source &s = *j->src;
    for (;;) {
        const int64_t u = j->next.fetch_add(1, std::memory_order_relaxed);
        if (u >= j->total_units || j->abort.load(std::memory_order_relaxed)) return;
        const int64_t c = std::upper_bound(j->unit_base.begin(), j->unit_base.end(), u) - j->unit_base.begin() - 1;
        while (c >= j->writable.load(std::memory_order_acquire)) {
            if (j->abort.load(std::memory_order_relaxed)) return;
            std::this_thread::yield();
        }
        const int64_t c0 = c * j->chunk;
        const int64_t a = c0 + (u - j->unit_base[c]) * j->grain, b = std::min(a + j->grain, c0 + j->chunk_len(c));
        char *out = static_cast<char *>(j->pin[c % RING]) + static_cast<size_t>(a - c0) * j->ldo * j->elt;
        uint32_t f;
        if (j->mode == 0)
            f = grm_host_quantise(s.A, s.sel, s.float_tag, j->n, s.lda, a, b, j->m, reinterpret_cast<int8_t *>(out), j->ldo);
        else if (j->mode == 1)
            f = grm_host_copy_codes(s.codes, j->n, s.lda, a, b, j->m, reinterpret_cast<int8_t *>(out), j->ldo);
        else
            f = grm_host_copy(s.A, s.sel, s.float_tag, j->n, s.lda, a, b, reinterpret_cast<double *>(out), j->ldo);
        if (f) j->flags.fetch_or(f, std::memory_order_relaxed);
        j->done[c].fetch_add(1, std::memory_order_release);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Feeder: the rank's loci [0, pc) in chunks, each made available as a DEVICE buffer on the compute stream.
//   DEVICE   the caller's device buffer, one chunk, no copy
//   DIRECT   pinned host FP64 without selectors: asynchronous copies straight from the caller's memory
//   RING_F64 host FP64 (pageable and / or Union layout): worker threads copy into the pinned ring (missing -> NaN)
//   RING_I8  host dosages: worker threads quantise (or copy pre-packed codes) into the pinned ring, 1 byte per cell
// The ring has RING slots (pinned host + device staging); chunks must be fetched in order.
// ---------------------------------------------------------------------------------------------------------------
struct Feeder {
    enum Kind { DEVICE, DIRECT, RING_F64, RING_I8 };
    grm_cuda_ctx *ctx = nullptr;
    const Source *src = nullptr;
    Kind kind = DEVICE;
    int64_t n = 0, pc = 0, chunk = 0, nchunks = 0, ld = 0;  // ld: leading dimension (elements) of a staged chunk
    int m = 0;
    size_t elt = 8;
    bool used[RING] = {false, false, false, false};
    int cur = 0;
    float host_ms = 0.f;
    uint32_t host_flags = 0;
    StreamJob job;
    bool running = false;
    int64_t fetched = 0;  // chunks handed to the device so far
    Timer t_job;

    int32_t init(grm_cuda_ctx *c, const Source *s, Kind k, int64_t n_, int64_t pc_, int m_, int64_t want_chunk, int world) {
        ctx = c;
        src = s;
        kind = k;
        n = n_;
        pc = pc_;
        m = m_;
        if (kind == DEVICE) {
            chunk = std::max<int64_t>(pc, 1);
            nchunks = pc > 0 ? 1 : 0;
            ld = s->lda;
            return GRM_CUDA_OK;
        }
        elt = kind == RING_I8 ? 1 : 8;
        ld = kind == RING_I8 ? (n + 3) / 4 * 4 : (n + 1) / 2 * 2;
        int64_t ch = want_chunk;
        if (ch <= 0) {
            // FP64 chunks feed one SYRK launch each, with a fixed cost of ~0.8 ms per launch (column means, centring, the
            // read-modify-write of G): at C4's n = 10,000 a 64 MiB slot is 838 loci = 3.5 ms per chunk, a 256 MiB slot
            // brought the end-to-end time from 4.22 s to 3.47 s against 3.28 s for the device-resident step
            // (profiles/bench_r2_c4_n1*.json).  The int8 chunks only feed the transpose: 32 MiB keeps the ring short.
            const int64_t slot_mb = ctx->tune.ring_mb > 0 ? ctx->tune.ring_mb : (kind == RING_I8 ? 32 : 256);
            ch = (slot_mb << 20) / (ld * static_cast<int64_t>(elt));
        }
        ch = std::max<int64_t>(128, ch / 128 * 128);
        chunk = std::min(ch, (std::max<int64_t>(pc, 1) + 127) / 128 * 128);
        nchunks = (pc + chunk - 1) / chunk;
        const size_t slot = static_cast<size_t>(ld) * static_cast<size_t>(std::min(chunk, std::max<int64_t>(pc, 1))) * elt;
        const int slots = static_cast<int>(std::min<int64_t>(RING, std::max<int64_t>(nchunks, 1)));
        for (int i = 0; i < slots; ++i) GRM_TRY(ctx->stage[i].reserve(slot));
        if (kind == DIRECT || nchunks == 0) return GRM_CUDA_OK;
        GRM_TRY(ensure_pool(ctx, world));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));  // nothing of an earlier pass may still read the ring
        GRM_TRY(ensure_ring(ctx, slot));
        // the whole pass as one job
        job.src = src;
        job.n = n;
        job.pc = pc;
        job.chunk = chunk;
        job.nchunks = nchunks;
        job.ldo = ld;
        job.m = m;
        job.mode = kind == RING_F64 ? 2 : (src->codes ? 1 : 0);
        job.elt = elt;
        for (int i = 0; i < RING; ++i) job.pin[i] = ctx->pin[i];
        const int64_t src_per_locus = n * (job.mode == 1 ? 1 : 9);
        job.grain = std::max<int64_t>(1, std::min<int64_t>(chunk, (3ll << 19) / std::max<int64_t>(src_per_locus, 1)));
        job.unit_base.assign(nchunks + 1, 0);
        for (int64_t cc = 0; cc < nchunks; ++cc)
            job.unit_base[cc + 1] = job.unit_base[cc] + (job.chunk_len(cc) + job.grain - 1) / job.grain;
        job.total_units = job.unit_base[nchunks];
        job.done.reset(new std::atomic<int32_t>[nchunks]);
        for (int64_t cc = 0; cc < nchunks; ++cc) job.done[cc].store(0, std::memory_order_relaxed);
        job.writable.store(RING, std::memory_order_relaxed);
        t_job = Timer();
        grm_pool_start(ctx->pool, stream_worker, &job);
        running = true;
        return GRM_CUDA_OK;
    }
    int64_t k0(int64_t c) const { return c * chunk; }
    int64_t len(int64_t c) const { return std::min(chunk, pc - c * chunk); }

    // device pointer of chunk c, valid on the compute stream (chunks in order)
    int32_t get(int64_t c, const void **d) {
        if (kind == DEVICE) {
            *d = src->A;
            return GRM_CUDA_OK;
        }
        const int s = static_cast<int>(c % RING);
        const int64_t a = k0(c), l = len(c);
        void *dst = ctx->stage[s].ptr;
        if (kind == DIRECT) {
            if (used[s]) GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[s], 0));
            GRM_CUDA_TRY(cudaMemcpy2DAsync(dst, ld * sizeof(double), src->A + a * src->lda, src->lda * sizeof(double),
                                           n * sizeof(double), static_cast<size_t>(l), cudaMemcpyHostToDevice,
                                           ctx->copy_stream));
        } else {
            if (c >= 1) {
                // the copy of chunk c - 1 has left its pinned slot: chunk c - 1 + RING may be written
                GRM_CUDA_TRY(cudaEventSynchronize(ctx->copy_done[(c - 1) % RING]));
                job.writable.store(c + RING, std::memory_order_release);
            }
            const int32_t want = job.units_of(c);
            for (unsigned spins = 0; job.done[c].load(std::memory_order_acquire) < want; ++spins) {
                if (spins < 64) std::this_thread::yield();
                else std::this_thread::sleep_for(std::chrono::microseconds(20));
            }
            host_flags |= job.flags.load(std::memory_order_relaxed);
            if (used[s]) GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[s], 0));
            GRM_CUDA_TRY(cudaMemcpyAsync(dst, ctx->pin[s], static_cast<size_t>(l) * ld * elt, cudaMemcpyHostToDevice,
                                         ctx->copy_stream));
        }
        GRM_CUDA_TRY(cudaEventRecord(ctx->copy_done[s], ctx->copy_stream));
        GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->copy_done[s], 0));
        used[s] = true;
        cur = s;
        fetched = c + 1;
        *d = dst;
        return GRM_CUDA_OK;
    }
    int32_t done() {
        if (kind == DEVICE) return GRM_CUDA_OK;
        GRM_CUDA_TRY(cudaEventRecord(ctx->stage_free[cur], ctx->stream));
        return GRM_CUDA_OK;
    }
    // end of the pass (also an early end): stop the workers and collect their flags
    void finish() {
        if (!running) return;
        // every fetched chunk was complete; after an early end a worker may still wait for a ring slot that will never be
        // released, so the decision is taken on the chunks, not on the unit counter
        if (fetched < nchunks) job.abort.store(1);
        grm_pool_wait(ctx->pool);
        running = false;
        host_ms += t_job.ms();
        host_flags |= job.flags.load();
    }
    ~Feeder() { finish(); }
};

// ---------------------------------------------------------------------------------------------------------------
// Way out for a PAGEABLE result (Julia's Matrix{Float64}): an asynchronous device->host copy into pageable memory
// blocks the calling thread until it is done, which would put the 8 n^2 bytes in front of the repair instead of under
// it.  A helper thread drains the matrix through the pinned ring (idle once the ingest is over) — copy engine -> pinned
// slot, then the worker pool spreads the memcpy into the caller's pages — while the main thread drives the
// factorisations.  The repair only ever changes the diagonal, which is patched afterwards.
// ---------------------------------------------------------------------------------------------------------------
struct OutCopier {
    std::thread th;
    int32_t status = GRM_CUDA_OK;
    char err[256] = "";
    bool active = false;

    struct CopyJob {
        const double *src;
        double *dst;
        int64_t n, ldg, cols;
    };
    static void copy_job(void *arg, int tid, int nt) {
        const CopyJob *j = static_cast<const CopyJob *>(arg);
        const int64_t a = j->cols * tid / nt, b = j->cols * (tid + 1) / nt;
        for (int64_t c = a; c < b; ++c) memcpy(j->dst + c * j->ldg, j->src + c * j->n, static_cast<size_t>(j->n) * sizeof(double));
    }

    void fail(const char *what, cudaError_t e) {
        status = GRM_CUDA_ERR_CUDA;
        snprintf(err, sizeof(err), "result copy-out: %s failed: %s", what, cudaGetErrorString(e));
    }

    // ready: event on which dG is complete.  Uses ctx->copy_stream, ctx->copy_done[] and the pinned ring.
    // `n` rows (contiguous) x `ncols` columns: the dense matrix (n x n) or a rank's packed tiles (65536 x count).
    void start(grm_cuda_ctx *ctx, const double *dG, int64_t ldG, int64_t n, int64_t ncols, double *G, int64_t ldg,
               cudaEvent_t ready) {
        active = true;
        th = std::thread([=]() {
            cudaError_t e = cudaSetDevice(ctx->device);
            if (e != cudaSuccess) return fail("cudaSetDevice", e);
            cudaStream_t cs = ctx->copy_stream;
            if ((e = cudaStreamWaitEvent(cs, ready, 0)) != cudaSuccess) return fail("cudaStreamWaitEvent", e);
            const int64_t cols = std::max<int64_t>(1, static_cast<int64_t>(ctx->pin_cap / (static_cast<size_t>(n) * sizeof(double))));
            const int64_t nchunks = (ncols + cols - 1) / cols;
            auto issue = [&](int64_t c) -> cudaError_t {
                const int s = static_cast<int>(c % RING);
                const int64_t c0 = c * cols, l = std::min(cols, ncols - c0);
                cudaError_t r = cudaMemcpy2DAsync(ctx->pin[s], n * sizeof(double), dG + c0 * ldG, ldG * sizeof(double),
                                                  n * sizeof(double), static_cast<size_t>(l), cudaMemcpyDeviceToHost, cs);
                if (r != cudaSuccess) return r;
                return cudaEventRecord(ctx->copy_done[s], cs);
            };
            for (int64_t c = 0; c < std::min<int64_t>(RING, nchunks); ++c)
                if ((e = issue(c)) != cudaSuccess) return fail("cudaMemcpy2DAsync", e);
            for (int64_t c = 0; c < nchunks; ++c) {
                const int s = static_cast<int>(c % RING);
                if ((e = cudaEventSynchronize(ctx->copy_done[s])) != cudaSuccess) return fail("cudaEventSynchronize", e);
                const int64_t c0 = c * cols, l = std::min(cols, ncols - c0);
                CopyJob job{static_cast<const double *>(ctx->pin[s]), G + c0 * ldg, n, ldg, l};
                grm_pool_run(ctx->pool, copy_job, &job);
                if (c + RING < nchunks && (e = issue(c + RING)) != cudaSuccess) return fail("cudaMemcpy2DAsync", e);
            }
        });
    }
    int32_t join() {
        if (!active) return GRM_CUDA_OK;
        th.join();
        active = false;
        if (status != GRM_CUDA_OK) grm_set_error("%s", err);
        return status;
    }
    ~OutCopier() {
        if (active && th.joinable()) th.join();
    }
};

// precedence as in the reference: any `missing` makes grmsimple / grmploidyaware throw (calc.jl:127,205) whatever else
// the matrix holds; a NaN payload only surfaces later, in inflatediagonals! (calc.jl:43)
int32_t flags_to_status(uint32_t f, bool skip_repair) {
    if (f & F_EMPTY_LOCUS) {
        grm_set_error("a locus has no valid (non-missing) cell: its mean fill is undefined");
        return GRM_CUDA_ERR_MISSING;
    }
    if (f & F_MISSING) {
        grm_set_error("allele_frequencies contains missing values; the reference GRM functions throw on missing data "
                      "(src/grm/calc.jl:127,205) — filter or impute first");
        return GRM_CUDA_ERR_MISSING;
    }
    if (f & F_INF) {
        grm_set_error("allele_frequencies contains +-Inf");
        return GRM_CUDA_ERR_NONFINITE;
    }
    if (f & F_NAN_PAYLOAD) {
        if (skip_repair) {
            grm_set_error("allele_frequencies contains NaN (as a Float64 value, not `missing`): the GRM would contain NaN");
            return GRM_CUDA_ERR_NAN_MATRIX;
        }
        grm_set_error("The input matrix is not symmetric.");  // NaN != NaN: what calc.jl:43 reports for such a GRM
        return GRM_CUDA_ERR_NOT_SYMMETRIC;
    }
    return GRM_CUDA_OK;
}

int denominator_from_bits(uint32_t bits) {
    // every code64 is a multiple of 2^tz  =>  x = (code64 >> tz) / (64 >> tz)
    int tz = 6;
    if (bits != 0) {
        tz = 0;
        while (((bits >> tz) & 1u) == 0) ++tz;
    }
    return 64 >> tz;
}

struct Run {
    grm_cuda_ctx *ctx = nullptr;
    grm_cuda_options o;
    Source src;  // offset to this rank's first locus
    int64_t n = 0, p = 0;  // p: total loci
    bool centred = false;
    int32_t ploidy = 0;
    double *G = nullptr;
    int64_t ldg = 0;
    grm_cuda_info *info = nullptr;

    bool dist = false, sharded = false, out_host = false, out_tiles = false, masked = false, meanfill = false;
    int rank = 0, world = 1;      // ranks of the communicator (distributed) — NOT the legacy o.rank / o.world
    int64_t L = 0, k0 = 0, pc = 0;  // loci per rank, first locus and count of this rank
    grm_cuda_info inf;

    // ---- small collectives on agreed values --------------------------------------------------------------
    int32_t agree_or(uint32_t *word) {
        if (!dist) return GRM_CUDA_OK;
        int32_t h[32];
        for (int i = 0; i < 32; ++i) h[i] = static_cast<int32_t>((*word >> i) & 1u);
        GRM_TRY(ctx->cons.reserve(4096));
        int32_t *d = ctx->cons.as<int32_t>();
        GRM_CUDA_TRY(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
        GRM_TRY(grm_comm_allreduce(ctx, d, d, 32, GRM_DT_I32, 1, ctx->stream));
        GRM_CUDA_TRY(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        uint32_t w = 0;
        for (int i = 0; i < 32; ++i) w |= (h[i] ? 1u : 0u) << i;
        *word = w;
        return GRM_CUDA_OK;
    }
    int32_t agree_sum(int64_t *vals, int count) {
        if (!dist) return GRM_CUDA_OK;
        GRM_TRY(ctx->cons.reserve(4096));
        int64_t *d = ctx->cons.as<int64_t>() + 64;
        GRM_CUDA_TRY(cudaMemcpyAsync(d, vals, count * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        GRM_TRY(grm_comm_allreduce(ctx, d, d, static_cast<size_t>(count), GRM_DT_I64, 0, ctx->stream));
        GRM_CUDA_TRY(cudaMemcpyAsync(vals, d, count * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return GRM_CUDA_OK;
    }
    // every rank's `count` doubles, in rank order (deterministic sums on the host)
    int32_t gather_f64(const double *mine, int count, std::vector<double> &all) {
        all.assign(static_cast<size_t>(world) * count, 0.0);
        if (!dist) {
            std::copy(mine, mine + count, all.begin());
            return GRM_CUDA_OK;
        }
        GRM_TRY(ctx->cons.reserve(4096 + static_cast<size_t>(world + 1) * count * sizeof(double)));
        double *d = reinterpret_cast<double *>(ctx->cons.as<char>() + 4096);
        GRM_CUDA_TRY(cudaMemcpyAsync(d, mine, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        GRM_TRY(grm_comm_allgather(ctx, d, d + count, static_cast<size_t>(count), GRM_DT_F64, ctx->stream));
        GRM_CUDA_TRY(cudaMemcpyAsync(all.data(), d + count, all.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return GRM_CUDA_OK;
    }

    int32_t detect(int32_t *m_out);
    int32_t go();
    int32_t int8_path(int32_t m, double *dG, int64_t ldG, double **d_tiles_out, uint32_t *retry, int64_t *p_eff,
                      double *denom, Feeder::Kind kind);
    int32_t f64_path(double *dG, int64_t ldG, double **d_tiles_out, int64_t *p_eff, double *denom, Feeder::Kind kind);
    int32_t gather_dense(const double *d_own_tiles, double *dG, int64_t ldG);
    int32_t repair_parallel(double *dG, int64_t ldG);
};

// dosage denominator "detected on load": a sample of the rank's first loci decides (2^22 cells); the full pack pass
// verifies every cell against it and the caller retries with m = 64 / falls back to FP64 if a later locus disagrees
int32_t Run::detect(int32_t *m_out) {
    uint32_t word = 0;  // bits 0..6: OR of round(64 x); bit 8: sample holds a cell that is no dosage / no number
    if (pc > 0) {
        if (src.host) {
            uint32_t bits = 0, fl = 0;
            grm_host_detect(src.A, src.sel, src.float_tag, n, pc, src.lda, 1ll << 22, masked ? 1 : 0, &bits, &fl);
            word = (bits & 0x7fu) | (fl ? 0x100u : 0u);
        } else {
            GRM_TRY(ctx->misc.reserve(64));
            uint32_t *d_out = ctx->misc.as<uint32_t>();
            GRM_TRY(grm_launch_detect(ctx, src.A, n, pc, src.lda, 1ll << 22, d_out, masked ? 1 : 0));
            uint32_t h[2];
            GRM_CUDA_TRY(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
            GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            word = (h[0] & 0x7fu) | (h[1] ? 0x100u : 0u);
        }
    }
    GRM_TRY(agree_or(&word));
    *m_out = (word & 0x100u) ? 0 : denominator_from_bits(word & 0x7fu);
    return GRM_CUDA_OK;
}

// gathered copy of the finished tiles: all-gather of every rank's packed FP64 tiles, then unpack into the dense
// matrix with an exact mirror (calc.jl:128-129)
int32_t Run::gather_dense(const double *d_own_tiles, double *dG, int64_t ldG) {
    const int64_t tmax = grm_cuda_packed_tiles_max(n, world);
    const size_t per_rank = static_cast<size_t>(tmax) * 65536;
    GRM_TRY(ctx->gtiles_all.reserve(per_rank * world * sizeof(double)));
    GRM_TRY(grm_comm_allgather(ctx, d_own_tiles, ctx->gtiles_all.ptr, per_rank, GRM_DT_F64, ctx->stream));
    GRM_TRY(grm_unpack_tiles(ctx, ctx->gtiles_all.as<double>(), n, world, dG, ldG));
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// int8 path.  *retry: 0 = done, 1 = retry with m = 64, 2 = fall back to the FP64 path.
// ---------------------------------------------------------------------------------------------------------------
int32_t Run::int8_path(int32_t m, double *dG, int64_t ldG, double **d_tiles_out, uint32_t *retry, int64_t *p_eff,
                       double *denom_out, Feeder::Kind kind) {
    cudaStream_t st = ctx->stream;
    cudaEvent_t *ev = ctx->ev;
    *retry = 0;
    const bool forced = o.path == GRM_CUDA_PATH_I8 || o.denominator != 0 || src.codes != nullptr;
    // limits of the integer arithmetic: int32 Gram entries, the byte planes / 64-bit sums of the statistics
    const double mnf = static_cast<double>(m) * static_cast<double>(n);
    const bool gram_overflow = static_cast<double>(m) * m * static_cast<double>(p) >= 2147483648.0;
    const bool stats_overflow = centred && (n >= (1ll << 18) || static_cast<double>(p) * mnf * mnf >= 9.0e18);
    if (gram_overflow || stats_overflow) {
        if (!forced) {  // the reference computes these inputs: so do we, on the FP64 path
            *retry = 2;
            return GRM_CUDA_OK;
        }
        grm_set_error("m^2 * p = %d^2 * %lld (n = %lld) exceeds the exact integer range of the int8 path", m, (long long)p,
                      (long long)n);
        return GRM_CUDA_ERR_OVERFLOW;
    }

    const int64_t cols = dist ? L : p;              // loci held by one rank
    const int64_t ldc = (cols + 127) / 128 * 128;
    const bool ksplit = dist && p > 4 * n;
    GRM_TRY(ctx->codes.reserve(static_cast<size_t>(n) * ldc));
    GRM_TRY(ctx->colsum.reserve(static_cast<size_t>(ldc) * sizeof(int32_t)));
    int8_t *d_codes = ctx->codes.as<int8_t>();
    int32_t *d_colsum = ctx->colsum.as<int32_t>();
    uint32_t *d_flags = ctx->flags.as<uint32_t>();
    GRM_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 16 * sizeof(uint32_t), st));
    GRM_CUDA_TRY(cudaMemsetAsync(d_colsum, 0, static_cast<size_t>(ldc) * sizeof(int32_t), st));
    if (dist && pc < L)  // a short (or empty) last slab: the loci it does not hold must read as zeros
        GRM_CUDA_TRY(cudaMemsetAsync(d_codes, 0, static_cast<size_t>(n) * ldc, st));
    int32_t *d_miss = nullptr;
    uint8_t *d_keep = nullptr;
    uint64_t *d_kept = nullptr;
    if (masked) {
        GRM_TRY(ctx->cnt.reserve(static_cast<size_t>(ldc) * sizeof(int32_t)));
        GRM_TRY(ctx->keep.reserve(static_cast<size_t>(ldc) + 64));
        d_miss = ctx->cnt.as<int32_t>();
        d_keep = ctx->keep.as<uint8_t>() + 16;
        d_kept = ctx->keep.as<uint64_t>();
        GRM_CUDA_TRY(cudaMemsetAsync(d_miss, 0, static_cast<size_t>(ldc) * sizeof(int32_t), st));
        GRM_CUDA_TRY(cudaMemsetAsync(d_kept, 0, sizeof(uint64_t), st));
    }

    // ---- ingest: quantise + transpose + column sums, one pass -------------------------------------------------
    Feeder feed;
    GRM_TRY(feed.init(ctx, &src, kind, n, pc, m, o.chunk_loci, world));
    for (int64_t c = 0; c < feed.nchunks; ++c) {
        const void *dc;
        GRM_TRY(feed.get(c, &dc));
        const int64_t a = feed.k0(c), l = feed.len(c);
        if (kind == Feeder::RING_I8)
            GRM_TRY(grm_cuda_pack_codes(ctx, static_cast<const int8_t *>(dc), n, l, feed.ld, d_codes + a, ldc, d_colsum + a,
                                        d_flags, masked ? d_miss + a : nullptr));
        else
            GRM_TRY(grm_cuda_pack_dosage(ctx, static_cast<const double *>(dc), n, l, feed.ld, m, d_codes + a, ldc,
                                         d_colsum + a, d_flags, masked ? d_miss + a : nullptr));
        GRM_TRY(feed.done());
        // a cell that breaks the guessed denominator ends the pass early (the whole pass is repeated anyway)
        if (feed.host_flags & (F_NOT_DOSAGE | F_INF | F_NAN_PAYLOAD)) break;
    }
    feed.finish();
    inf.ms_host_busy += feed.host_ms;
    if (masked && pc > 0) {
        // a KEPT locus that still holds missing cells always raises the flag: under the "error" policy that is the
        // reference's throw, under mean-fill the filled values are no dosages and the call moves to the FP64 path
        GRM_TRY(grm_cuda_locus_keep(ctx, d_colsum, d_miss, n, pc, m, o.maf, o.max_locus_sparsity, 1, d_keep, d_kept, d_flags));
        GRM_TRY(grm_cuda_mask_codes(ctx, d_codes, n, pc, ldc, d_keep));
    }
    GRM_CUDA_TRY(cudaEventRecord(ev[1], st));

    // ---- flags / kept loci (agreed by all ranks) ---------------------------------------------------------------
    uint32_t h_flags = 0;
    uint64_t h_kept = static_cast<uint64_t>(pc);
    GRM_CUDA_TRY(cudaMemcpyAsync(&h_flags, d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (masked) GRM_CUDA_TRY(cudaMemcpyAsync(&h_kept, d_kept, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    GRM_CUDA_TRY(cudaStreamSynchronize(st));
    h_flags |= feed.host_flags & (masked ? ~F_MISSING : ~0u);
    GRM_TRY(agree_or(&h_flags));
    int64_t kept = static_cast<int64_t>(h_kept);
    if (masked) GRM_TRY(agree_sum(&kept, 1));
    *p_eff = masked ? kept : p;
    if ((h_flags & F_MISSING) && meanfill && !(h_flags & (F_INF | F_NAN_PAYLOAD))) {
        // missing cells + mean-fill policy: the filled values are not dosages -> FP64 path
        if (o.path == GRM_CUDA_PATH_I8 || src.codes) {
            grm_set_error("int8 path requested but missing cells must be mean-filled (FP64 path)");
            return GRM_CUDA_ERR_NOT_DOSAGE;
        }
        *retry = 2;
        return GRM_CUDA_OK;
    }
    GRM_TRY(flags_to_status(h_flags, o.skip_repair != 0));
    if (h_flags & F_NOT_DOSAGE) {
        if (forced) {
            grm_set_error("allele frequencies are not all multiples of 1/%d", m);
            return GRM_CUDA_ERR_NOT_DOSAGE;
        }
        *retry = m < 64 ? 1 : 2;  // once more with the finest supported grid, then give up on int8
        return GRM_CUDA_OK;
    }
    if (masked && *p_eff == 0) {
        grm_set_error("All loci filtered out at maf = %g, maximum locus sparsity = %g.", o.maf, o.max_locus_sparsity);
        return GRM_CUDA_ERR_ALL_FILTERED;
    }

    // ---- centring statistics: exact integers, additive over loci ------------------------------
    double sc[4] = {1.0 / (static_cast<double>(m) * m), 0.0, 0.0, static_cast<double>(*p_eff)};  // calc.jl:127
    const double *d_u = nullptr;
    if (centred) {
        GRM_TRY(ctx->u.reserve(static_cast<size_t>(n) * sizeof(double)));
        GRM_TRY(ctx->rT.reserve(static_cast<size_t>(2 * n + 2) * sizeof(int64_t)));
        uint64_t *d_sums = ctx->rT.as<uint64_t>();
        int64_t *d_r = ctx->rT.as<int64_t>() + 2, *d_T = d_r + n;
        if (pc > 0)
            GRM_TRY(grm_cuda_dosage_stats(ctx, d_codes, n, pc, ldc, d_colsum, m, d_sums, d_r, d_T, 0));
        else
            GRM_CUDA_TRY(cudaMemsetAsync(d_sums, 0, static_cast<size_t>(2 * n + 2) * sizeof(int64_t), st));
        if (dist)
            GRM_TRY(grm_comm_allreduce(ctx, d_sums, d_sums, static_cast<size_t>(2 * n + 2), GRM_DT_I64, 0, st));
        GRM_TRY(grm_cuda_dosage_centre(ctx, d_sums, d_r, d_T, n, *p_eff, m, ploidy, ctx->u.as<double>(), sc));
        d_u = ctx->u.as<double>();  // sc[3] = d, calc.jl:201
    }
    *denom_out = sc[3];
    GRM_CUDA_TRY(cudaEventRecord(ev[2], st));

    // ---- contraction ---------------------------------------------------------------------------------------------
    const int64_t pk = std::max<int64_t>(pc, 1);  // an empty slab contracts one zero K block
    if (!dist) {
        if (out_tiles) {
            const int64_t cnt = grm_cuda_tile_count(n, o.rank, o.world);
            GRM_TRY(ctx->gtiles.reserve(static_cast<size_t>(std::max<int64_t>(cnt, 1)) * 65536 * sizeof(double)));
            GRM_TRY(grm_cuda_syrk_i8_f64_tiles(ctx, d_codes, n, pk, ldc, 1, sc[0], sc[1], sc[2], sc[3], d_u,
                                               ctx->gtiles.as<double>(), o.rank, o.world));
            *d_tiles_out = ctx->gtiles.as<double>();
        } else {
            GRM_TRY(grm_cuda_syrk_i8_f64(ctx, d_codes, n, pk, ldc, 1, sc[0], sc[1], sc[2], sc[3], d_u, dG, ldG, o.rank,
                                         o.world));
        }
        inf.syrk_window = ctx->last_syrk_window;
        GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
        GRM_CUDA_TRY(cudaEventRecord(ev[10], st));
        GRM_CUDA_TRY(cudaEventRecord(ev[11], st));
        if (!sharded && !out_tiles) GRM_TRY(grm_cuda_scale_mirror(ctx, dG, n, ldG, 1.0));
        return GRM_CUDA_OK;
    }

    const int64_t tmax = grm_cuda_packed_tiles_max(n, world);
    GRM_TRY(ctx->gtiles.reserve(static_cast<size_t>(tmax) * 65536 * sizeof(double)));
    double *d_gt = ctx->gtiles.as<double>();
    if (ksplit) {
        // loci-split: every rank contracts ITS loci over ALL tiles; the int32 partial tiles are summed by a reduce-scatter
        // cut into waves, wave w travelling (comm stream) under the contraction of wave w + 1 (compute stream)
        int waves = ctx->tune.ksplit_waves > 0 ? ctx->tune.ksplit_waves : (world > 1 ? 4 : 1);
        waves = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(waves, 8), tmax)));
        inf.exchange_waves = waves;
        const int exchange = ctx->tune.exchange;  // 0: one launch per wave + NCCL; 1: streamed + NCCL; 2: streamed + copy engines
        GRM_TRY(ctx->packed.reserve(static_cast<size_t>(tmax) * world * 65536 * sizeof(int32_t)));
        int32_t *d_packed = ctx->packed.as<int32_t>();
        cudaStream_t cs = ctx->comm_stream;
        // ONE persistent launch contracts every tile in wave order and counts the finished tiles per wave; a gate kernel
        // per wave on the communication stream waits for that count and releases the exchange of the wave while the
        // tensor cores work on the later waves
        GRM_TRY(ctx->progress.reserve(4096));
        uint32_t *d_done = ctx->progress.as<uint32_t>() + 512;  // clear of the K-window progress words
        uint32_t *d_epoch = d_done + 16;
        GRM_CUDA_TRY(cudaMemsetAsync(d_done, 0, 8 * sizeof(uint32_t), st));
        if (exchange == 0) {
            // default: one launch per wave; the reduce-scatter (int32 sum, exact) and the finalisation of wave w go to the
            // communication stream and slot in next to the launch of wave w + 1
            GRM_TRY(ctx->reduced.reserve(static_cast<size_t>(tmax) * 65536 * sizeof(int32_t)));
            int32_t *d_red = ctx->reduced.as<int32_t>();
            for (int w = 0; w < waves; ++w) {
                int64_t s0, s1;
                grm_packed_wave_range(n, world, waves, w, &s0, &s1);
                GRM_TRY(grm_syrk_i8_packed_wave(ctx, d_codes, n, pk, ldc, world, waves, w, d_packed));
                GRM_CUDA_TRY(cudaEventRecord(ctx->wave_ev[w], st));
                if (s1 <= s0) continue;
                GRM_CUDA_TRY(cudaStreamWaitEvent(cs, ctx->wave_ev[w], 0));
                GRM_TRY(grm_comm_reduce_scatter(ctx, d_packed + static_cast<size_t>(world) * s0 * 65536, d_red + s0 * 65536,
                                                static_cast<size_t>(s1 - s0) * 65536, GRM_DT_I32, cs));
                GRM_TRY(grm_finalize_tiles_range(ctx, d_red, n, sc[0], sc[1], sc[2], sc[3], d_u, nullptr, 0, d_gt, rank, world, s0,
                                                 s1 - s0, cs));
            }
            GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
        } else if (exchange == 1) {
            // streamed, NCCL reduce-scatter: its kernels need SMs, so the SYRK leaves `comm_ctas` of them free
            GRM_TRY(ctx->reduced.reserve(static_cast<size_t>(tmax) * 65536 * sizeof(int32_t)));
            int32_t *d_red = ctx->reduced.as<int32_t>();
            GRM_CUDA_TRY(cudaEventRecord(ctx->wave_ev[9], st));
            GRM_CUDA_TRY(cudaStreamWaitEvent(cs, ctx->wave_ev[9], 0));  // codes, u and the zeroed counters are ready
            GRM_TRY(grm_syrk_i8_packed_stream(ctx, d_codes, n, pk, ldc, world, waves, d_packed, d_done,
                                              waves > 1 ? std::max(ctx->tune.comm_ctas, 0) : 0));
            GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
            for (int w = 0; w < waves; ++w) {
                int64_t s0, s1;
                grm_packed_wave_range(n, world, waves, w, &s0, &s1);
                if (s1 <= s0) continue;
                GRM_TRY(grm_wave_gate(ctx, d_done, n, world, waves, w, cs));
                GRM_TRY(grm_comm_reduce_scatter(ctx, d_packed + static_cast<size_t>(world) * s0 * 65536, d_red + s0 * 65536,
                                                static_cast<size_t>(s1 - s0) * 65536, GRM_DT_I32, cs));
                GRM_TRY(grm_finalize_tiles_range(ctx, d_red, n, sc[0], sc[1], sc[2], sc[3], d_u, nullptr, 0, d_gt, rank, world, s0,
                                                 s1 - s0, cs));
            }
        } else {
            // streamed, exchange by the COPY ENGINES over NVLink peer mappings (tuning "exchange" = 2) — no SM is taken from the
            // contraction, but with every GPU sending to every other at once the engines move only 80 GB/s per sender in
            // 16 MiB pieces and ~340 GB/s in 128 MiB ones (profiles/p2p_copy_engine_probe_r2.jsonl) against ~530 GB/s for
            // NCCL's reduce-scatter, so it is not the default.  Rank r's
            // receive buffer holds one part per source rank; when wave w is stored, this rank copies, for every owner q,
            // its partial tiles of q's slots [s0, s1) into part `rank` of q's buffer and then a 4-byte epoch flag on the
            // same stream (so the flag lands after the data).  The owner's gate waits for the flags of all sources, then
            // the parts are summed (int32, exact, fixed order) inside the finalisation of the wave's tiles.
            void **peers = nullptr;
            GRM_TRY(grm_comm_peer_buffer(ctx, GRM_PEER_HEADER + static_cast<size_t>(world) * tmax * 65536 * sizeof(int32_t), &peers));
            const uint32_t epoch = grm_comm_next_epoch(ctx);
            GRM_TRY(grm_set_u32(ctx, d_epoch, epoch, st));
            GRM_CUDA_TRY(cudaEventRecord(ctx->wave_ev[9], st));
            GRM_CUDA_TRY(cudaStreamWaitEvent(cs, ctx->wave_ev[9], 0));  // codes, u, the counters and the epoch word are ready
            GRM_TRY(grm_syrk_i8_packed_stream(ctx, d_codes, n, pk, ldc, world, waves, d_packed, d_done, 0));
            GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
            auto data_of = [&](int q) { return reinterpret_cast<int32_t *>(static_cast<char *>(peers[q]) + GRM_PEER_HEADER); };
            auto flag_of = [&](int q, int w, int src) { return static_cast<uint32_t *>(peers[q]) + w * 64 + src; };
            cudaEvent_t released = grm_comm_xevent(ctx, GRM_XSTREAMS);
            for (int w = 0; w < waves; ++w) {
                int64_t s0, s1;
                grm_packed_wave_range(n, world, waves, w, &s0, &s1);
                if (s1 <= s0) continue;
                const int64_t len = s1 - s0;
                GRM_TRY(grm_wave_gate(ctx, d_done, n, world, waves, w, cs));
                GRM_CUDA_TRY(cudaEventRecord(released, cs));
                for (int i = 0; i < std::min(GRM_XSTREAMS, world); ++i)
                    GRM_CUDA_TRY(cudaStreamWaitEvent(grm_comm_xstream(ctx, i), released, 0));
                for (int k = 0; k < world; ++k) {
                    const int q = (rank + 1 + k) % world;  // every rank starts with another owner: the links share the load
                    cudaStream_t xs = grm_comm_xstream(ctx, k);
                    GRM_CUDA_TRY(cudaMemcpyAsync(data_of(q) + (static_cast<size_t>(rank) * tmax + s0) * 65536,
                                                 d_packed + (static_cast<size_t>(world) * s0 + static_cast<size_t>(q) * len) * 65536,
                                                 static_cast<size_t>(len) * 65536 * sizeof(int32_t), cudaMemcpyDeviceToDevice, xs));
                    GRM_CUDA_TRY(cudaMemcpyAsync(flag_of(q, w, rank), d_epoch, sizeof(uint32_t), cudaMemcpyDeviceToDevice, xs));
                }
                GRM_TRY(grm_flags_gate(ctx, flag_of(rank, w, 0), world, epoch, cs));
                GRM_TRY(grm_finalize_parts_range(ctx, data_of(rank), n, sc[0], sc[1], sc[2], sc[3], d_u, d_gt, rank, world, s0, len, cs));
            }
            // the copies read d_packed: the compute stream (hence the end of the call) waits for them too
            for (int i = 0; i < std::min(GRM_XSTREAMS, world); ++i) {
                GRM_CUDA_TRY(cudaEventRecord(grm_comm_xevent(ctx, i), grm_comm_xstream(ctx, i)));
                GRM_CUDA_TRY(cudaStreamWaitEvent(st, grm_comm_xevent(ctx, i), 0));
            }
        }
        GRM_CUDA_TRY(cudaEventRecord(ctx->wave_ev[8], cs));
        GRM_CUDA_TRY(cudaStreamWaitEvent(st, ctx->wave_ev[8], 0));
        GRM_CUDA_TRY(cudaEventRecord(ev[10], st));
    } else {
        // tile-split: the packed codes are all-gathered (n p bytes in total) and every rank contracts all loci over
        // the tiles it owns (3-D tensor map over [slab][entry][locus])
        GRM_TRY(ctx->codes_all.reserve(static_cast<size_t>(world) * n * ldc));
        GRM_TRY(grm_comm_allgather(ctx, d_codes, ctx->codes_all.ptr, static_cast<size_t>(n) * ldc, GRM_DT_U8, st));
        GRM_CUDA_TRY(cudaEventRecord(ev[12], st));
        GRM_TRY(grm_cuda_syrk_i8_f64_tiles(ctx, ctx->codes_all.as<int8_t>(), n, std::min(L, p), ldc, world, sc[0], sc[1], sc[2],
                                           sc[3], d_u, d_gt, rank, world));
        GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
        GRM_CUDA_TRY(cudaEventRecord(ev[10], st));
    }
    if (out_tiles) {
        *d_tiles_out = d_gt;
        GRM_CUDA_TRY(cudaEventRecord(ev[11], st));
        return GRM_CUDA_OK;
    }
    GRM_TRY(gather_dense(d_gt, dG, ldG));
    GRM_CUDA_TRY(cudaEventRecord(ev[11], st));
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// FP64 path: per chunk, column means (+ validity, QC mask), operand Z, then G += Z Z' (upper triangle, DMMA)
// ---------------------------------------------------------------------------------------------------------------
int32_t Run::f64_path(double *dG, int64_t ldG, double **d_tiles_out, int64_t *p_eff, double *denom_out, Feeder::Kind kind) {
    cudaStream_t st = ctx->stream;
    cudaEvent_t *ev = ctx->ev;
    uint32_t *d_flags = ctx->flags.as<uint32_t>();
    GRM_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 16 * sizeof(uint32_t), st));
    GRM_TRY(ctx->stats.reserve(8 * sizeof(double)));
    double *d_stats = ctx->stats.as<double>();
    GRM_CUDA_TRY(cudaMemsetAsync(d_stats, 0, 8 * sizeof(double), st));
    const int64_t cols = std::max<int64_t>(dist ? L : p, 1);
    GRM_TRY(ctx->q.reserve(static_cast<size_t>((cols + 127) / 128 * 128) * sizeof(double)));
    uint8_t *d_keep = nullptr;
    uint64_t *d_kept = nullptr;
    if (masked) {
        GRM_TRY(ctx->keep.reserve(static_cast<size_t>(cols) + 64));
        d_keep = ctx->keep.as<uint8_t>() + 16;
        d_kept = ctx->keep.as<uint64_t>();
        GRM_CUDA_TRY(cudaMemsetAsync(d_kept, 0, sizeof(uint64_t), st));
    }
    // distributed: every rank accumulates ITS loci over ALL tiles (a plain single-GPU SYRK into its local matrix)
    const int32_t srank = dist ? 0 : o.rank, sworld = dist ? 1 : o.world;
    Feeder feed;
    GRM_TRY(feed.init(ctx, &src, kind, n, pc, 0, o.chunk_loci, world));
    if (feed.nchunks == 0) GRM_CUDA_TRY(cudaMemset2DAsync(dG, ldG * sizeof(double), 0, n * sizeof(double), n, st));
    const bool staged = kind != Feeder::DEVICE;
    for (int64_t c = 0; c < feed.nchunks; ++c) {
        const void *dv;
        GRM_TRY(feed.get(c, &dv));
        const double *dc = static_cast<const double *>(dv);
        const int64_t a = feed.k0(c), l = feed.len(c), ld = feed.ld;
        double *qc = ctx->q.as<double>() + a;
        if (masked)
            GRM_TRY(grm_cuda_locus_qc(ctx, dc, n, l, ld, o.maf, o.max_locus_sparsity, meanfill ? 0 : 1, qc, d_keep + a, d_kept,
                                      d_stats, c > 0, d_flags));
        else
            GRM_TRY(grm_cuda_locus_stats_f64(ctx, dc, n, l, ld, qc, d_stats, d_flags, c > 0, meanfill ? 1 : 0));
        const uint8_t *kc = masked ? d_keep + a : nullptr;
        if (!(centred || meanfill || masked)) {
            // grmsimple on clean data: the operands are the input itself
            GRM_TRY(grm_cuda_syrk_f64(ctx, dc, n, l, ld, c > 0, dG, ldG, srank, sworld));
        } else if (staged) {
            // the staged chunk is ours: centre / fill / mask it in place, then a pure-DMMA SYRK
            double *z = const_cast<double *>(dc);
            GRM_TRY(grm_cuda_centre_f64(ctx, dc, n, l, ld, centred ? 1 : 0, static_cast<double>(ploidy), qc, meanfill ? 1 : 0,
                                        kc, z, ld));
            GRM_TRY(grm_cuda_syrk_f64(ctx, z, n, l, ld, c > 0, dG, ldG, srank, sworld));
        } else {
            // device-resident caller buffer (read-only): transformed sub-chunks go through a scratch buffer
            const int64_t zc = std::max<int64_t>(32, std::min<int64_t>(l, (512ll << 20) / (n * 8)));
            GRM_TRY(ctx->stage[0].reserve(static_cast<size_t>(n) * zc * sizeof(double)));
            double *z = ctx->stage[0].as<double>();
            for (int64_t s0 = 0; s0 < l; s0 += zc) {
                const int64_t sl = std::min(zc, l - s0);
                GRM_TRY(grm_cuda_centre_f64(ctx, dc + s0 * ld, n, sl, ld, centred ? 1 : 0, static_cast<double>(ploidy), qc + s0,
                                            meanfill ? 1 : 0, kc ? kc + s0 : nullptr, z, n));
                GRM_TRY(grm_cuda_syrk_f64(ctx, z, n, sl, n, (c > 0 || s0 > 0) ? 1 : 0, dG, ldG, srank, sworld));
            }
        }
        GRM_TRY(feed.done());
    }
    feed.finish();
    inf.ms_host_busy += feed.host_ms;
    GRM_CUDA_TRY(cudaEventRecord(ev[1], st));
    GRM_CUDA_TRY(cudaEventRecord(ev[2], st));
    uint32_t h_flags = 0;
    double h_stats[3] = {0.0, 0.0, 0.0};
    uint64_t h_kept = static_cast<uint64_t>(pc);
    GRM_CUDA_TRY(cudaMemcpyAsync(&h_flags, d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    GRM_CUDA_TRY(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
    if (masked) GRM_CUDA_TRY(cudaMemcpyAsync(&h_kept, d_kept, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    GRM_CUDA_TRY(cudaStreamSynchronize(st));
    // with selector bytes a NaN payload is a value, not `missing`; the staged copy cannot tell them apart any more
    if (feed.host_flags & F_NAN_PAYLOAD) h_flags |= F_NAN_PAYLOAD;
    GRM_TRY(agree_or(&h_flags));
    int64_t kept = static_cast<int64_t>(h_kept);
    if (masked) GRM_TRY(agree_sum(&kept, 1));
    *p_eff = masked ? kept : p;
    if (h_flags & F_NAN_PAYLOAD) h_flags &= ~F_MISSING;
    GRM_TRY(flags_to_status(h_flags, o.skip_repair != 0));
    if (masked && *p_eff == 0) {
        grm_set_error("All loci filtered out at maf = %g, maximum locus sparsity = %g.", o.maf, o.max_locus_sparsity);
        return GRM_CUDA_ERR_ALL_FILTERED;
    }
    double sum_q1q = h_stats[0];
    if (dist) {  // sum q(1-q) over the ranks in rank order: the same bits on every rank and in every run
        std::vector<double> all;
        GRM_TRY(gather_f64(h_stats, 1, all));
        sum_q1q = 0.0;
        for (int r = 0; r < world; ++r) sum_q1q += all[r];
    }
    const double denom = centred ? static_cast<double>(ploidy) * sum_q1q : static_cast<double>(*p_eff);  // calc.jl:201, :127
    *denom_out = denom;
    GRM_CUDA_TRY(cudaEventRecord(ev[3], st));
    if (!dist) {
        GRM_CUDA_TRY(cudaEventRecord(ev[10], st));
        GRM_CUDA_TRY(cudaEventRecord(ev[11], st));
        if (sharded)
            GRM_TRY(grm_scale_dense_tiles(ctx, dG, n, ldG, o.rank, o.world, denom));
        else
            GRM_TRY(grm_cuda_scale_mirror(ctx, dG, n, ldG, denom));
        return GRM_CUDA_OK;
    }
    // loci-split exchange: upper-triangle tiles -> reduce-scatter (FP64 sum) -> / d on the owner -> all-gather
    const int64_t tmax = grm_cuda_packed_tiles_max(n, world);
    GRM_TRY(ctx->packed.reserve(static_cast<size_t>(tmax) * world * 65536 * sizeof(double)));
    GRM_TRY(ctx->gtiles.reserve(static_cast<size_t>(tmax) * 65536 * sizeof(double)));
    double *d_packed = ctx->packed.as<double>(), *d_gt = ctx->gtiles.as<double>();
    GRM_TRY(grm_pack_tiles_f64(ctx, dG, n, ldG, world, d_packed));
    GRM_TRY(grm_comm_reduce_scatter(ctx, d_packed, d_gt, static_cast<size_t>(tmax) * 65536, GRM_DT_F64, st));
    GRM_TRY(grm_scale_tiles_f64(ctx, d_gt, n, rank, world, denom));
    GRM_CUDA_TRY(cudaEventRecord(ev[10], st));
    if (out_tiles) {
        *d_tiles_out = d_gt;
        GRM_CUDA_TRY(cudaEventRecord(ev[11], st));
        return GRM_CUDA_OK;
    }
    GRM_TRY(gather_dense(d_gt, dG, ldG));
    GRM_CUDA_TRY(cudaEventRecord(ev[11], st));
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// inflatediagonals! with the matrix replicated on every rank (calc.jl:41-70): the loop test of round t only looks at
// X with its diagonal inflated t times, so rank r probes round base + r (isposdef, and the determinant when it is
// positive definite: repair.cu) and a batch of probes costs one round's factorisation of wall time and covers `world`
// rounds.  Each probe performs the arithmetic of the single-GPU loop and the decision is taken on the all-gathered
// results, hence the same outcome on every rank as on one GPU.
// ---------------------------------------------------------------------------------------------------------------
int32_t Run::repair_parallel(double *dG, int64_t ldG) {
    const double eps = DBL_EPSILON;
    int64_t base = 0;
    int32_t iters = -1;
    double maxabs = 0.0;
    while (iters < 0) {
        const int64_t r = base + rank;
        double mine[4] = {NAN, -1.0, 0.0, 0.0};
        if (r <= o.max_iter) {
            double det = NAN, mx = 0.0;
            int32_t pd = -1;
            uint32_t fl = 0;
            GRM_TRY(grm_cuda_repair_probe(ctx, dG, n, ldG, static_cast<int32_t>(r), &det, &pd, &fl, &mx));
            mine[0] = det;
            mine[1] = static_cast<double>(pd);
            mine[2] = static_cast<double>(fl);
            mine[3] = mx;
        }
        std::vector<double> all;
        GRM_TRY(gather_f64(mine, 4, all));
        if (base == 0) {
            const uint32_t flags = static_cast<uint32_t>(all[2]);
            maxabs = all[3];
            if (flags & 1u) {
                grm_set_error("The input matrix is not symmetric.");  // calc.jl:43-45
                return GRM_CUDA_ERR_NOT_SYMMETRIC;
            }
            if (all[0] > eps && all[1] == 1.0) {  // calc.jl:49-51
                if (rank == 0) ctx->factor_n = n;
                inf.repair_iters = 0;
                inf.eps_total = 0.0;
                return GRM_CUDA_OK;
            }
            if (flags & 2u) {
                grm_set_error("The input matrix contains NaN values.");
                return GRM_CUDA_ERR_NAN_MATRIX;
            }
            if (flags & 4u) {
                grm_set_error("The input matrix contains Inf values.");
                return GRM_CUDA_ERR_INF_MATRIX;
            }
        }
        for (int i = 0; i < world && iters < 0; ++i) {
            const int64_t t = base + i;
            const double det = all[4 * i], pd = all[4 * i + 1];
            // calc.jl:61 (a NaN det compares false, then isposdef decides; a half that was not evaluated is decided by the other)
            const bool need = (det < eps) || (pd != 1.0);
            if (!need || t >= o.max_iter) {
                iters = static_cast<int32_t>(t);
                if (!need && rank == i) ctx->factor_n = n;  // this rank's Cholesky is of the result
            }
        }
        base += world;
    }
    double tot = 0.0;
    GRM_TRY(grm_cuda_repair_apply(ctx, dG, n, ldG, iters, maxabs, &tot));
    inf.repair_iters = iters;
    inf.eps_total = tot;
    return GRM_CUDA_OK;
}

int32_t Run::go() {
    Timer wall;
    if (!ctx || (!src.A && !src.codes) || n < 1 || p < 1 || src.lda < n || o.max_iter < 0 || o.world < 1 || o.rank < 0 ||
        o.rank >= o.world) {
        grm_set_error("bad argument (n=%lld p=%lld lda=%lld ldg=%lld rank=%d world=%d)", (long long)n, (long long)p,
                      (long long)src.lda, (long long)ldg, o.rank, o.world);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    dist = o.distributed != 0;
    sharded = o.world > 1;
    out_tiles = o.output_mode == GRM_CUDA_OUT_TILES;
    out_host = o.output_location == GRM_CUDA_LOC_HOST;
    meanfill = o.missing_policy == GRM_CUDA_MISSING_MEANFILL;
    masked = o.maf > 0.0 || o.max_locus_sparsity < 1.0;
    if (dist && sharded) {
        grm_set_error("opts.distributed and the communicator-less tile sharding (opts.rank / opts.world) exclude each other");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (dist && !ctx->comm) {
        grm_set_error("opts.distributed needs a communicator on the context (grm_cuda_comm_init)");
        return GRM_CUDA_ERR_COMM;
    }
    rank = dist ? grm_comm_rank(ctx) : 0;
    world = dist ? grm_comm_world(ctx) : 1;
    if (!G && !(dist && !out_tiles)) {
        grm_set_error("null output pointer");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (G && !out_tiles && ldg < n) {
        grm_set_error("bad argument (ldg=%lld < n=%lld)", (long long)ldg, (long long)n);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (!(o.maf >= 0.0 && o.maf <= 0.5) || !(o.max_locus_sparsity >= 0.0 && o.max_locus_sparsity <= 1.0)) {
        grm_set_error("maf must be in [0, 0.5] and max_locus_sparsity in [0, 1]");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (o.output_mode != GRM_CUDA_OUT_DENSE && o.output_mode != GRM_CUDA_OUT_TILES) {
        grm_set_error("unknown output_mode %d", o.output_mode);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    const bool do_repair = !o.skip_repair && !sharded && !out_tiles;
    if (n > INT32_MAX - 512 || p > INT32_MAX - 512) {
        grm_set_error("n and p must fit TMA's 32-bit coordinates");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    ctx->repair_lu_count = 0;
    ctx->last_syrk_window = 0;
    memset(&inf, 0, sizeof(inf));
    inf.struct_size = static_cast<int32_t>(sizeof(inf));
    inf.world = world;

    // this rank's loci
    L = dist ? (p + world - 1) / world : p;
    k0 = dist ? std::min<int64_t>(p, static_cast<int64_t>(rank) * L) : 0;
    pc = dist ? std::min<int64_t>(p, k0 + L) - k0 : p;
    if (dist && !o.input_slab) {
        if (src.A) src.A += k0 * src.lda;
        if (src.sel) src.sel += k0 * src.lda;
        if (src.codes) src.codes += k0 * src.lda;
    }

    cudaStream_t st = ctx->stream;
    cudaEvent_t *ev = ctx->ev;
    GRM_TRY(ctx->flags.reserve(64));
    GRM_CUDA_TRY(cudaEventRecord(ev[0], st));

    // output buffer on the device
    double *dG = nullptr;
    int64_t ldG = n;
    if (!out_tiles) {
        if (out_host || !G) {
            GRM_TRY(ctx->G.reserve(static_cast<size_t>(n) * n * sizeof(double)));
            dG = ctx->G.as<double>();
            if (sharded) GRM_CUDA_TRY(cudaMemsetAsync(dG, 0, static_cast<size_t>(n) * n * sizeof(double), st));
        } else {
            dG = G;
            ldG = ldg;
        }
    }

    // ---- path decision ---------------------------------------------------------------------------------------
    int path = src.codes ? GRM_CUDA_PATH_I8 : o.path;
    int32_t m = 0;
    if (path != GRM_CUDA_PATH_F64) {
        m = o.denominator;
        if (m == 0) {
            GRM_TRY(detect(&m));
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

    // how host input travels (info.ingest)
    auto feeder_kind = [&](int pth) -> Feeder::Kind {
        if (!src.host) return Feeder::DEVICE;
        if (src.codes) return Feeder::RING_I8;
        const int want = ctx->tune.ingest;
        if (pth == GRM_CUDA_PATH_I8 && want != GRM_CUDA_INGEST_STREAM) return Feeder::RING_I8;
        if (!src.sel && is_pinned(src.A)) return Feeder::DIRECT;
        return Feeder::RING_F64;
    };

    double denom = 0.0;
    int64_t p_eff = p;
    double *d_tiles = nullptr;
    for (bool done = false; !done;) {
        const Feeder::Kind kind = feeder_kind(path);
        inf.ingest = kind == Feeder::DEVICE ? 0 : (kind == Feeder::RING_I8 ? GRM_CUDA_INGEST_HOSTQ : GRM_CUDA_INGEST_STREAM);
        if (path == GRM_CUDA_PATH_I8) {
            uint32_t retry = 0;
            GRM_TRY(int8_path(m, dG, ldG, &d_tiles, &retry, &p_eff, &denom, kind));
            if (retry == 1) {
                m = 64;
            } else if (retry == 2) {
                path = GRM_CUDA_PATH_F64;
            } else {
                done = true;
            }
        } else {
            if (out_tiles && !dist) {
                grm_set_error("output_mode TILES is available on the int8 path and on distributed calls");
                return GRM_CUDA_ERR_BAD_ARGUMENT;
            }
            if (!dG) {  // tiles out: the FP64 loci-split still accumulates into a local dense matrix before the exchange
                GRM_TRY(ctx->G.reserve(static_cast<size_t>(n) * n * sizeof(double)));
                dG = ctx->G.as<double>();
                ldG = n;
            }
            GRM_TRY(f64_path(dG, ldG, &d_tiles, &p_eff, &denom, kind));
            done = true;
        }
    }
    GRM_CUDA_TRY(cudaEventRecord(ev[4], st));

    // ---- inflatediagonals! (calc.jl:133 / :211) and the way out -------------------------------------------------
    // The repair only ever touches the diagonal, so the bulk device->host copy of G runs on the copy stream WHILE the
    // O(n^3) factorisations run; if a round of inflation happened, the n diagonal values are copied again afterwards.
    const bool want_copy = out_host && G != nullptr;
    // a pageable dense result leaves through the helper thread (OutCopier); pinned memory is copied to directly
    const int64_t tiles_owned = dist ? grm_cuda_tile_count(n, rank, world) : grm_cuda_tile_count(n, o.rank, o.world);
    const bool via_ring = want_copy && !is_pinned(G) && (!out_tiles || tiles_owned > 0);
    OutCopier drain;
    bool bulk_started = false;
    auto copy_out = [&](cudaStream_t s) -> int32_t {
        if (out_tiles) {
            const int64_t cnt = dist ? grm_cuda_tile_count(n, rank, world) : grm_cuda_tile_count(n, o.rank, o.world);
            if (cnt > 0)
                GRM_CUDA_TRY(cudaMemcpyAsync(G, d_tiles, static_cast<size_t>(cnt) * 65536 * sizeof(double), cudaMemcpyDeviceToHost, s));
        } else if (ldg == n) {
            GRM_CUDA_TRY(cudaMemcpyAsync(G, dG, static_cast<size_t>(n) * n * sizeof(double), cudaMemcpyDeviceToHost, s));
        } else {
            GRM_CUDA_TRY(cudaMemcpy2DAsync(G, ldg * sizeof(double), dG, ldG * sizeof(double), n * sizeof(double),
                                           static_cast<size_t>(n), cudaMemcpyDeviceToHost, s));
        }
        return GRM_CUDA_OK;
    };
    if (via_ring) {
        GRM_TRY(ensure_pool(ctx, world));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));  // the ring and its events belong to the finished ingest
        GRM_TRY(ensure_ring(ctx, std::max<size_t>(ctx->pin_cap, std::max<size_t>(32u << 20, static_cast<size_t>(out_tiles ? 65536 : n) * sizeof(double)))));
        GRM_CUDA_TRY(cudaEventRecord(ev[7], st));
        if (out_tiles)
            drain.start(ctx, d_tiles, 65536, 65536, tiles_owned, G, 65536, ev[4]);
        else
            drain.start(ctx, dG, ldG, n, n, G, ldg, ev[4]);
        bulk_started = true;
    } else if (want_copy && do_repair) {
        GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ev[4], 0));
        GRM_CUDA_TRY(cudaEventRecord(ev[7], ctx->copy_stream));
        GRM_TRY(copy_out(ctx->copy_stream));
        GRM_CUDA_TRY(cudaEventRecord(ev[8], ctx->copy_stream));
        bulk_started = true;
    }
    if (do_repair) {
        const int32_t rst = dist ? repair_parallel(dG, ldG)
                                 : grm_repair_device(ctx, dG, n, ldG, o.max_iter, &inf.repair_iters, &inf.eps_total);
        if (rst != GRM_CUDA_OK) {
            drain.join();
            cudaStreamSynchronize(ctx->copy_stream);  // never leave a copy into the caller's buffer in flight
            return rst;
        }
    }
    GRM_CUDA_TRY(cudaEventRecord(ev[5], st));
    if (via_ring) {
        Timer td;
        GRM_TRY(drain.join());
        GRM_CUDA_TRY(cudaEventRecord(ev[8], st));
        inf.ms_d2h = td.ms();  // what the copy-out adds after the repair (the rest ran underneath it)
        if (inf.repair_iters > 0) {
            // the repaired diagonal: n doubles through a pinned slot, scattered by the host
            double *pd = static_cast<double *>(ctx->pin[0]);
            GRM_CUDA_TRY(cudaMemcpy2DAsync(pd, sizeof(double), dG, (ldG + 1) * sizeof(double), sizeof(double),
                                           static_cast<size_t>(n), cudaMemcpyDeviceToHost, st));
            GRM_CUDA_TRY(cudaStreamSynchronize(st));
            for (int64_t i = 0; i < n; ++i) G[i * (ldg + 1)] = pd[i];
        }
    } else if (want_copy) {
        if (bulk_started) {
            if (inf.repair_iters > 0) {
                // diagonal only: n elements with a pitch of (ld + 1) doubles on both sides
                GRM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ev[5], 0));
                GRM_CUDA_TRY(cudaMemcpy2DAsync(G, (ldg + 1) * sizeof(double), dG, (ldG + 1) * sizeof(double), sizeof(double),
                                               static_cast<size_t>(n), cudaMemcpyDeviceToHost, ctx->copy_stream));
            }
            GRM_CUDA_TRY(cudaStreamSynchronize(ctx->copy_stream));
        } else {
            GRM_CUDA_TRY(cudaEventRecord(ev[7], st));
            GRM_TRY(copy_out(st));
            GRM_CUDA_TRY(cudaEventRecord(ev[8], st));
        }
    } else if (!out_host && G) {
        if (out_tiles) {
            const int64_t cnt = dist ? grm_cuda_tile_count(n, rank, world) : grm_cuda_tile_count(n, o.rank, o.world);
            if (cnt > 0 && d_tiles != G)
                GRM_CUDA_TRY(cudaMemcpyAsync(G, d_tiles, static_cast<size_t>(cnt) * 65536 * sizeof(double),
                                             cudaMemcpyDeviceToDevice, st));
        } else if (dG != G) {
            GRM_CUDA_TRY(cudaMemcpy2DAsync(G, ldg * sizeof(double), dG, ldG * sizeof(double), n * sizeof(double),
                                           static_cast<size_t>(n), cudaMemcpyDeviceToDevice, st));
        }
    }
    GRM_CUDA_TRY(cudaEventRecord(ev[6], st));
    GRM_CUDA_TRY(cudaStreamSynchronize(st));

    inf.path = path;
    inf.denominator = path == GRM_CUDA_PATH_I8 ? m : 0;
    inf.scale = denom;
    inf.tiles = dist ? (path == GRM_CUDA_PATH_I8 && p <= 4 * n ? grm_cuda_tile_count(n, rank, world)
                                                               : grm_cuda_tile_count(n, 0, 1))
                     : grm_cuda_tile_count(n, o.rank, o.world);
    cudaEventElapsedTime(&inf.ms_pack, ev[0], ev[1]);
    cudaEventElapsedTime(&inf.ms_rowdot, ev[1], ev[2]);
    cudaEventElapsedTime(&inf.ms_syrk, ev[2], ev[3]);
    cudaEventElapsedTime(&inf.ms_exchange, ev[3], ev[10]);
    cudaEventElapsedTime(&inf.ms_gather, ev[10], ev[11]);
    cudaEventElapsedTime(&inf.ms_finalize, ev[11], ev[4]);
    cudaEventElapsedTime(&inf.ms_repair, ev[4], ev[5]);
    if (want_copy && !via_ring) cudaEventElapsedTime(&inf.ms_d2h, ev[7], ev[8]);
    inf.ms_h2d = src.host ? inf.ms_pack : 0.f;  // host input: the ingest stage is host work, PCIe and pack overlapped
    inf.host_threads = (src.host && ctx->pool) ? grm_pool_threads(ctx->pool) : 0;
    inf.ms_total = wall.ms();
    inf.kernel_launches = ctx->launches;
    inf.repair_lu = ctx->repair_lu_count;
    inf.loci_kept = static_cast<int32_t>(p_eff);
    if (info) memcpy(info, &inf, std::min<size_t>(sizeof(inf), static_cast<size_t>(std::max(info->struct_size, 0))));
    return GRM_CUDA_OK;
}

int32_t start(Run &r, grm_cuda_ctx *ctx, int64_t n, int64_t p, bool centred, int32_t ploidy, double *G, int64_t ldg,
              const grm_cuda_options *uopts, grm_cuda_info *info) {
    grm_cuda_options_default(&r.o);
    if (uopts) memcpy(&r.o, uopts, std::min<size_t>(sizeof(r.o), static_cast<size_t>(std::max(uopts->struct_size, 0))));
    r.ctx = ctx;
    r.n = n;
    r.p = p;
    r.centred = centred;
    r.ploidy = ploidy;
    r.G = G;
    r.ldg = ldg;
    r.info = info;
    return GRM_CUDA_OK;
}

}  // namespace

extern "C" int32_t grm_cuda_simple(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda, double *G,
                                   int64_t ldg, const grm_cuda_options *opts, grm_cuda_info *info) try {
    Run r;
    start(r, ctx, n, p, false, 0, G, ldg, opts, info);
    r.src.A = A;
    r.src.lda = lda;
    r.src.host = r.o.input_location == GRM_CUDA_LOC_HOST;
    return r.go();
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_ploidyaware(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda,
                                        int32_t ploidy, double *G, int64_t ldg, const grm_cuda_options *opts,
                                        grm_cuda_info *info) try {
    Run r;
    start(r, ctx, n, p, true, ploidy, G, ldg, opts, info);
    r.src.A = A;
    r.src.lda = lda;
    r.src.host = r.o.input_location == GRM_CUDA_LOC_HOST;
    return r.go();
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_simple_union(grm_cuda_ctx *ctx, const double *payload, const uint8_t *selectors,
                                         int32_t float_tag, int64_t n, int64_t p, int64_t lda, double *G, int64_t ldg,
                                         const grm_cuda_options *opts, grm_cuda_info *info) try {
    Run r;
    start(r, ctx, n, p, false, 0, G, ldg, opts, info);
    r.src.A = payload;
    r.src.sel = selectors;
    r.src.float_tag = float_tag;
    r.src.lda = lda;
    r.src.host = true;
    return r.go();
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_ploidyaware_union(grm_cuda_ctx *ctx, const double *payload, const uint8_t *selectors,
                                              int32_t float_tag, int64_t n, int64_t p, int64_t lda, int32_t ploidy,
                                              double *G, int64_t ldg, const grm_cuda_options *opts, grm_cuda_info *info) try {
    Run r;
    start(r, ctx, n, p, true, ploidy, G, ldg, opts, info);
    r.src.A = payload;
    r.src.sel = selectors;
    r.src.float_tag = float_tag;
    r.src.lda = lda;
    r.src.host = true;
    return r.go();
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_from_codes(grm_cuda_ctx *ctx, const int8_t *codes, int64_t n, int64_t p, int64_t ldc, int32_t m,
                                       int32_t ploidy, double *G, int64_t ldg, const grm_cuda_options *opts,
                                       grm_cuda_info *info) try {
    Run r;
    start(r, ctx, n, p, ploidy >= 0, ploidy >= 0 ? ploidy : 0, G, ldg, opts, info);
    if (m < 1 || m > 64 || (m & (m - 1)) != 0) {
        grm_set_error("from_codes: denominator must be a power of two in [1,64], got %d", m);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    r.o.denominator = m;
    r.o.path = GRM_CUDA_PATH_I8;
    r.src.codes = codes;
    r.src.lda = ldc;
    r.src.host = true;
    return r.go();
} GRM_CATCH_STATUS

// ---------------------------------------------------------------------------------------------------------------
// host utilities (no GPU involved): the quantiser of the ingest, callable on its own — it writes the int8 payload of
// the wire format (io.py) and lets the tests compare host-packed codes with the device pack kernel cell by cell
// ---------------------------------------------------------------------------------------------------------------
extern "C" int32_t grm_cuda_host_quantise(const double *payload, const uint8_t *selectors, int32_t float_tag, int64_t n,
                                          int64_t p, int64_t lda, int32_t m, int8_t *codes, int64_t ldo, int32_t threads,
                                          uint32_t *flags) try {
    if (!payload || !codes || n < 1 || p < 1 || lda < n || ldo < n || m < 1 || m > 64 || (m & (m - 1)) != 0) {
        grm_set_error("host_quantise: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    int nt = threads > 0 ? threads : host_cpus();
    nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(nt, 256), p)));
    std::vector<uint32_t> fl(nt, 0u);
    auto work = [&](int t) {
        const int64_t a = p * t / nt, b = p * (t + 1) / nt;
        if (a < b) fl[t] = grm_host_quantise(payload, selectors, float_tag, n, lda, a, b, m, codes + a * ldo, ldo);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    uint32_t f = 0;
    for (uint32_t v : fl) f |= v;
    if (flags) *flags = f;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_host_copy(const double *payload, const uint8_t *selectors, int32_t float_tag, int64_t n, int64_t p,
                                      int64_t lda, double *out, int64_t ldo, int32_t threads, uint32_t *flags) try {
    if (!payload || !out || n < 1 || p < 1 || lda < n || ldo < n) {
        grm_set_error("host_copy: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    int nt = threads > 0 ? threads : host_cpus();
    nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::min(nt, 256), p)));
    std::vector<uint32_t> fl(nt, 0u);
    auto work = [&](int t) {
        const int64_t a = p * t / nt, b = p * (t + 1) / nt;
        if (a < b) fl[t] = grm_host_copy(payload, selectors, float_tag, n, lda, a, b, out + a * ldo, ldo);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    uint32_t f = 0;
    for (uint32_t v : fl) f |= v;
    if (flags) *flags = f;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_host_detect(const double *payload, const uint8_t *selectors, int32_t float_tag, int64_t n,
                                        int64_t p, int64_t lda, int64_t max_cells, int32_t ignore_missing, int32_t *m_out) try {
    if (!payload || !m_out || n < 1 || p < 1 || lda < n) {
        grm_set_error("host_detect: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    uint32_t bits = 0, fl = 0;
    grm_host_detect(payload, selectors, float_tag, n, p, lda, max_cells > 0 ? max_cells : n * p, ignore_missing, &bits, &fl);
    *m_out = fl ? 0 : denominator_from_bits(bits & 0x7fu);
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" const char *grm_cuda_host_simd(void) { return grm_host_simd(); }

extern "C" int32_t grm_cuda_host_tune(const char *key, int64_t value) try {
    // a negative value restores the default of the key
    if (key && !strcmp(key, "prefetch")) {
        grm_host_set_prefetch(static_cast<int>(std::max<int64_t>(-1, std::min<int64_t>(value, 1 << 20))));
        return GRM_CUDA_OK;
    }
    if (key && !strcmp(key, "prefetch_hint")) {
        grm_host_set_prefetch_hint(static_cast<int>(std::max<int64_t>(-1, std::min<int64_t>(value, 3))));
        return GRM_CUDA_OK;
    }
    if (key && !strcmp(key, "stream_stores")) {
        grm_host_set_stream_stores(value < 0 ? -1 : (value != 0));
        return GRM_CUDA_OK;
    }
    grm_set_error("host_tune: unknown key (%s)", key ? key : "null");
    return GRM_CUDA_ERR_BAD_ARGUMENT;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_inflate_diagonals(grm_cuda_ctx *ctx, double *X, int64_t n, int64_t ldx, int32_t max_iter,
                                              int32_t location, int32_t *iters, double *eps_total) try {
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
} GRM_CATCH_STATUS
