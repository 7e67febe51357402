// common.cuh — shared declarations of libgrm_cuda (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <exception>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/grm_cuda.h"

// ---------------------------------------------------------------------------------------------------------
// error plumbing: no exception crosses the C ABI; every internal failure becomes a status + thread-local text
// ---------------------------------------------------------------------------------------------------------
void grm_set_error(const char *fmt, ...);

#define GRM_CUDA_TRY(expr)                                                                           \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            grm_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? GRM_CUDA_ERR_OOM : GRM_CUDA_ERR_CUDA;         \
        }                                                                                            \
    } while (0)

#define GRM_TRY(expr)                        \
    do {                                     \
        int32_t _s = (expr);                 \
        if (_s != GRM_CUDA_OK) return _s;    \
    } while (0)

// Every extern "C" entry point is a function-try-block closed by one of these: a C++ exception from the host runtime
// (std::bad_alloc of a vector, std::system_error of a thread) ends as a status, never in the caller's frames — Julia's
// ccall, like any C caller, cannot unwind through them.
#define GRM_CATCH_STATUS                                                    \
    catch (const std::bad_alloc &) {                                        \
        grm_set_error("out of host memory");                                \
        return GRM_CUDA_ERR_OOM;                                            \
    }                                                                       \
    catch (const std::exception &e) {                                       \
        grm_set_error("internal error: %s", e.what());                      \
        return GRM_CUDA_ERR_CUDA;                                           \
    }                                                                       \
    catch (...) {                                                           \
        grm_set_error("internal error: unknown C++ exception");             \
        return GRM_CUDA_ERR_CUDA;                                           \
    }
#define GRM_CATCH_I64                                                       \
    catch (...) {                                                           \
        return -1;                                                          \
    }

// grow-only device buffer
struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    int32_t reserve(size_t bytes);
    void release();
    template <class T> T *as() { return reinterpret_cast<T *>(ptr); }
};

struct TileListCache {
    int64_t n = -1;
    int32_t rank = -1, world = -1, kind = -1;
    int32_t count = 0;
    DevBuf buf;
};

// Explicit per-context tuning knobs (grm_cuda_ctx_set_tuning): the library never reads the environment.  Defaults are
// the measured best (DESIGN.md / profiles/README.md); tests use the knobs to force every kernel variant on small shapes.
struct GrmTuning {
    int syrk_i8 = 0;        // fused-epilogue int8 SYRK: 0 auto, 1 single-CTA, 2 CTA pair, 3 wide (two tiles per pair)
    int syrk_packed = 0;    // K-split packed-tile SYRK: 0 auto (CTA pair), 2 CTA pair, 3 wide
    int syrk_sync = -1;     // K-window synchronisation of the persistent pairs: -1 auto, 0 off, 1 forced
    int syrk_window = 64;   // K blocks a pair may run ahead of the slowest one
    int pack = 0;           // FP64 -> int8 pack kernel: 0 auto (TMA-staged when aligned), 1 register-staged
    int ieee_div = 0;       // 1: IEEE division in the epilogue instead of reciprocal + 2 FMA corrections (bit-identical)
    int tile_band = 0;      // band height (tile rows) of the tile enumeration; 0 = default
    int host_threads = 0;   // host ingest threads; 0 = automatic
    int ingest = 0;         // host input: 0 auto, 1 stream FP64 and pack on the device, 2 quantise on the host
    int ksplit_waves = 0;   // K-split exchange waves (reduce-scatter of wave w under the SYRK of the later waves); 0 = auto
    int comm_ctas = 0;      // > 0: cap NCCL's CTAs (and leave as many SMs free during the K-split SYRK when NCCL carries it)
    int repair = 0;         // single-GPU inflatediagonals!: 0 auto (two rounds evaluated concurrently from n = 3072), 1 sequential
    int repair_det = 0;     // determinant of the repair: 0 auto (pivots of the Cholesky when they can vouch for det < eps,
                            // else LU), 1 always LU first, Cholesky only when !(det < eps) — the literal order of calc.jl:49,61
    int exchange = 0;       // K-split partial tiles: 0 auto = NCCL reduce-scatter, 2 = copy engines into peer-mapped buffers
    int f64_kernel = 0;     // FP64 SYRK operand staging: 0 auto (TMA when aligned), 1 register-staged
    int ring_mb = 0;        // bytes (MiB) of one pinned ingest slot; 0 = automatic
};

// tile lists of the K-split contraction, per exchange wave (syrk_i8.cu)
struct PackedPlanCache {
    int64_t n = -1;
    int32_t world = -1, waves = -1, band = -1;
    std::vector<int32_t> tile_off, tile_cnt, unit_off, unit_cnt;  // per wave, in elements of the lists below
    int32_t total_tiles = 0, slots_padded = 0;
    DevBuf buf;  // [tiles: int2 x total][slots: int32 x total, padded to even][units of the wide kernel: int2]
};

struct GrmHostPool;  // ingest_host.cpp
struct GrmComm;      // comm.cu

struct grm_cuda_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;   // compute
    cudaStream_t copy_stream = nullptr;
    cudaStream_t comm_stream = nullptr;  // NCCL collectives that overlap the compute stream
    cudaEvent_t ev[16] = {};
    cudaEvent_t wave_ev[16] = {};        // K-split waves: SYRK of wave w finished / exchange of wave w finished
    cudaEvent_t copy_done[4] = {}, stage_free[4] = {};
    void *cusolver = nullptr;  // cusolverDnHandle_t
    void *cusolver_params = nullptr;  // cusolverDnParams_t of the 64-bit entry points (n > 46,000)
    // concurrent lanes of the speculative repair (repair.cu): extra handles, streams, events, pinned result buffers
    void *cusolver_lane[2] = {nullptr, nullptr};
    cudaStream_t lane_stream[2] = {nullptr, nullptr};
    cudaEvent_t lane_ev[4] = {};
    void *lane_pin = nullptr;
    size_t lane_pin_cap = 0;
    // workspaces (grow-only; reused across calls so a warm call does no cudaMalloc)
    DevBuf codes, colsum, flags, q, w, u, stats, partials, G, C, stage[4], lu, lu2, lu3, lu_work, lane_work[2], lane_ipiv[2], ipiv, misc, misc2, planes, rT, cnt,
        keep, progress, packed, reduced, gtiles, gtiles_all, codes_all, cons;
    TileListCache tiles_i8, tiles_i8_2cta, tiles_f64, tiles_packed, tiles_wide, tiles_f64_all;
    int64_t factor_n = -1;  // >= 0: a LOWER Cholesky factor of the n x n matrix the last repair returned is cached ...
    int factor_buf = 0;     // ... in ctx->lu (0) or ctx->lu2 (1)
    int32_t repair_lu_count = 0;  // LU factorisations the repair ran since the counter was last reset
    int32_t launches = 0;  // kernels launched since last reset
    GrmTuning tune;
    std::vector<const void *> attr_done;  // kernels whose dynamic shared memory limit was raised ON THIS CONTEXT'S DEVICE
    std::vector<std::pair<const void *, int>> max_clusters;  // co-resident CTA pairs of a kernel on the idle device
    int32_t last_syrk_window = 0;  // 1 if the last pair-kernel launch used the K-window synchronisation
    PackedPlanCache packed_plan[2];  // [0]: waves > 1 (K-split exchange waves), [1]: waves == 1
    // pinned host ring of the threaded ingest (grow-only) and the worker pool
    void *pin[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t pin_cap = 0;
    GrmHostPool *pool = nullptr;
    GrmComm *comm = nullptr;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE property: remember it per context, not per process
// (a second context on another GPU of the same process must raise the limit again).
int32_t grm_func_smem(grm_cuda_ctx *ctx, const void *func, size_t bytes);

// tile partition (tiles.cpp part of api.cu)
constexpr int GRM_TILE = 256;  // edge, in entries, of the unit of multi-GPU sharding
// all upper-triangular (ti <= tj) tiles in an L2-friendly order: bands of `band` tile rows (<= 0: default)
void grm_tile_order(int64_t n, int band, std::vector<int2> &tiles);
void grm_tile_range(int64_t total, int32_t rank, int32_t world, int64_t *begin, int64_t *end);

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time libcuda dependency)
typedef CUresult (*GrmEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int32_t grm_tmap_encoder(GrmEncodeTiledFn *fn);

// communicator (comm.cu): thin wrappers over the dlopen'ed NCCL; dtype = GRM_DT_*
constexpr int GRM_DT_F64 = 0, GRM_DT_I32 = 1, GRM_DT_I64 = 2, GRM_DT_U8 = 3;
int grm_comm_rank(const grm_cuda_ctx *ctx);
int grm_comm_world(const grm_cuda_ctx *ctx);
int32_t grm_comm_allreduce(grm_cuda_ctx *ctx, const void *send, void *recv, size_t count, int dtype, int op_max,
                           cudaStream_t s);
int32_t grm_comm_reduce_scatter(grm_cuda_ctx *ctx, const void *send, void *recv, size_t recvcount, int dtype,
                                cudaStream_t s);
int32_t grm_comm_allgather(grm_cuda_ctx *ctx, const void *send, void *recv, size_t sendcount, int dtype, cudaStream_t s);
// Peer-mapped receive buffer of the copy-engine exchange: every rank owns `bytes` (grow-only; COLLECTIVE when it has to
// grow: handles travel through the communicator) that all peers can address.  peers[q] = rank q's buffer as mapped here
// (peers[rank] = the own one).  The first GRM_PEER_HEADER bytes are flag words, the data follows.
constexpr size_t GRM_PEER_HEADER = 65536;
constexpr int GRM_XSTREAMS = 4;
int32_t grm_comm_peer_buffer(grm_cuda_ctx *ctx, size_t bytes, void ***peers);
uint32_t grm_comm_next_epoch(grm_cuda_ctx *ctx);
cudaStream_t grm_comm_xstream(grm_cuda_ctx *ctx, int i);
cudaEvent_t grm_comm_xevent(grm_cuda_ctx *ctx, int i);

// tile-layout helpers of the distributed paths (tiles.cu)
int32_t grm_packed_plan(grm_cuda_ctx *ctx, int64_t n, int32_t world, int32_t waves, PackedPlanCache **out);
// gathered FP64 tiles [world][tiles_max][256 x 256] -> dense n x n matrix, both triangles (exact mirror)
int32_t grm_unpack_tiles(grm_cuda_ctx *ctx, const double *d_tiles_all, int64_t n, int32_t world, double *d_G, int64_t ldg);
// dense upper triangle -> reduce-scatter layout [world][tiles_max][256 x 256] (FP64 K-split)
int32_t grm_pack_tiles_f64(grm_cuda_ctx *ctx, const double *d_G, int64_t n, int64_t ldg, int32_t world, double *d_packed);
// owned reduced FP64 tiles: t /= denom, zeros outside the matrix
int32_t grm_scale_tiles_f64(grm_cuda_ctx *ctx, double *d_tiles, int64_t n, int32_t rank, int32_t world, double denom);
// owned tiles of a dense matrix: upper-triangle elements /= denom
int32_t grm_scale_dense_tiles(grm_cuda_ctx *ctx, double *d_G, int64_t n, int64_t ldg, int32_t rank, int32_t world,
                              double denom);
// packed owned tiles [count][256 x 256] -> this rank's tiles inside a dense matrix (upper triangle only)
int32_t grm_scatter_tiles(grm_cuda_ctx *ctx, const double *d_tiles, int64_t n, int32_t rank, int32_t world, double *d_G,
                          int64_t ldg);

// K-split contraction in exchange waves and the finalisation of reduced tiles (syrk_i8.cu)
void grm_packed_wave_range(int64_t n, int32_t world, int32_t waves, int32_t w, int64_t *s0, int64_t *s1);
int32_t grm_syrk_i8_packed_wave(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc, int32_t world,
                                int32_t waves, int32_t wave, int32_t *d_packed);
int32_t grm_syrk_i8_packed_stream(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc, int32_t world,
                                  int32_t waves, int32_t *d_packed, uint32_t *d_wave_done, int32_t reserve_sms);
int32_t grm_wave_gate(grm_cuda_ctx *ctx, const uint32_t *d_wave_done, int64_t n, int32_t world, int32_t waves, int32_t wave,
                      cudaStream_t stream);
int32_t grm_flags_gate(grm_cuda_ctx *ctx, const uint32_t *d_flags, int count, uint32_t epoch, cudaStream_t stream);
int32_t grm_set_u32(grm_cuda_ctx *ctx, uint32_t *d_word, uint32_t value, cudaStream_t stream);
int32_t grm_finalize_parts_range(grm_cuda_ctx *ctx, const int32_t *d_parts, int64_t n, double alpha, double beta, double gamma,
                                 double denom, const double *d_u, double *d_Gt, int32_t rank, int32_t world, int64_t t0,
                                 int64_t cnt, cudaStream_t stream);
int32_t grm_finalize_tiles_range(grm_cuda_ctx *ctx, const int32_t *d_reduced, int64_t n, double alpha, double beta,
                                 double gamma, double denom, const double *d_u, double *d_G, int64_t ldg, double *d_Gt,
                                 int32_t rank, int32_t world, int64_t t0, int64_t cnt, cudaStream_t stream);

// kernels' host launchers (implemented across the .cu files)
int32_t grm_launch_detect(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda, int64_t max_cells,
                          uint32_t *d_out /* [0]=OR of code64, [1]=flags */, int ignore_nan);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier, TMA, tcgen05).  sm_100a only.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    // bounded spin: a protocol bug traps (error 719) instead of hanging the GPU.  2^26 polls is seconds, orders of
    // magnitude beyond any legitimate wait in these kernels.
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2D tiled load global -> shared, completion on an mbarrier (bytes)
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const void *tmap, uint64_t *bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 3D tiled load (loci, entries, slab)
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const void *tmap, uint64_t *bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t NCOLS> __device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread t gets lane (base_lane + t), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// ---- 2-CTA (cta_group::2) variants: a CTA pair on one TPC shares one UMMA of M = 256 ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <uint32_t NCOLS> __device__ __forceinline__ void tmem_alloc_2cta(uint32_t *smem_dst) {  // one warp per CTA
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS> __device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit of the address cleared)
__device__ __forceinline__ void tma_load_3d_2cta(void *smem_dst, const void *tmap, uint64_t *bar, int32_t c0, int32_t c1,
                                                 int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// The same load multicast to the CTAs of `mask` (same shared-memory offset in each); each destination's bytes are
// credited to the mbarrier at the same offset in the LEADER of the destination's CTA pair
__device__ __forceinline__ void tma_load_3d_2cta_mc(void *smem_dst, const void *tmap, uint64_t *bar, int32_t c0, int32_t c1,
                                                    int32_t c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_i8_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in every CTA of `mask` once the issued MMAs completed
__device__ __forceinline__ void umma_commit_2cta(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// arrive on the mbarrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}

// shared-memory matrix descriptor: K-major operand tile, 128-byte swizzle, rows of 128 bytes, 8-row groups
// 1024 bytes apart (the layout a SWIZZLE_128B TMA box {128 B, rows} lands in).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address        [0,14)
    d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset   [16,30) (unused for SW128 K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset    [32,46)
    d |= static_cast<uint64_t>(1) << 46;                      // descriptor version 1 (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::i8: S8 x S8 -> S32, both operands K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4)                                  // D format S32
           | (1u << 7)                                // A format: signed 8-bit
           | (1u << 10)                               // B format: signed 8-bit
           | (static_cast<uint32_t>(N >> 3) << 17)    // N
           | (static_cast<uint32_t>(M >> 4) << 24);   // M
}
#endif  // __CUDACC__
