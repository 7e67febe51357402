// syrk_i8.cu — exact integer Gram matrix C = c * c' of the packed dosage codes on tcgen05 tensor cores.
//
// Replaces the pair loop `dot(A[i,:], A[j,:])` of the reference (src/grm/calc.jl:124-131 and :202-209) for
// dosage-quantised input: with x = c/m the reference's FP64 sum is exactly C_ij/m^2, so the
// contraction is done in int8 x int8 -> int32 and the FP64 scaling / rank-1 centring is applied
// in the epilogue.
//
// Kernel shape (persistent, one CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer: operand tiles (K-major int8, 128-byte swizzle) global -> shared ring
//   warp 1      MMA issuer:   one thread issues tcgen05.mma.kind::i8, accumulators (int32) in TMEM
//   warps 2..5  epilogue:     tcgen05.ld TMEM -> registers -> (int32 | FP64 formula) -> global, overlapped with the
//                             next tile's main loop through two TMEM accumulator stages (2 x 256 columns)
// Three variants of the same pipeline (chosen by launch(); grm_cuda_ctx_set_tuning("syrk_i8", ...) forces one):
//   syrk_i8_2cta_kernel (default)  CTA pairs, cta_group::2, UMMA 256 x 256 x 32, 256 x 256 tiles, 6 x 32 KB stages
//   syrk_i8_wide_kernel (long K)   CTA pairs on two tiles that share their column panel, 4 x 48 KB stages
//   syrk_i8_kernel                 single CTA, UMMA 128 x 256 x 32, 128 x 256 tiles, 4 x 48 KB stages
// Work items come from a host-built list that covers only the upper triangle.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;       // rows of A per CTA tile  (UMMA M)
constexpr int BN = 256;       // rows of B per CTA tile  (UMMA N)
constexpr int BK = 128;       // bytes (= int8 elements) of K per pipeline stage = one 128-byte swizzle atom
constexpr int UMMA_K = 32;    // K per tcgen05.mma for 8-bit operands
constexpr int STAGES = 4;
constexpr uint32_t A_BYTES = BM * BK;               // 16 KiB
constexpr uint32_t B_BYTES = BN * BK;               // 32 KiB
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES; // 48 KiB
constexpr int NUM_THREADS = 192;
constexpr uint32_t TMEM_COLS = 2 * BN;              // two accumulator stages of 256 int32 columns
constexpr size_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct EpiParams {
    int64_t n;
    int32_t *C;      // mode 0
    int64_t ldC;
    double *G;       // modes 1, 2
    int64_t ldg;
    double alpha, beta, gamma, denom;
    double rcp;  // RN(1/denom) when the fast exact division applies, else 0
    const double *u;
    const int32_t *slots;  // mode 3: slot of tile t in the packed int32 tile buffer C ([slot][256 cols][256 rows])
    double *Gt;            // modes 1, 2 (2-CTA kernel): if set, tile t is stored packed at Gt + t*65536 ([col][row],
                           // zeros outside n) instead of into the dense matrix G: the sharded-output form
    // mode 3, streamed exchange: tiles [wave_end[w-1], wave_end[w]) of the list form exchange wave w; every epilogue warp
    // that has stored its quarter of a tile adds 1 to wave_done[w] (8 per tile), so a gate kernel on the communication
    // stream can release the reduce-scatter of wave w while the contraction of the later waves is still running
    uint32_t *wave_done;
    int32_t wave_end[8];
    int32_t nwaves;
};

// Correctly rounded v / d for a launch-constant divisor: q0 = RN(v*rcp), then two FMA residual corrections
// (Markstein).  After the first correction q is within 1/2 ulp + 2^-51 ulp of v/d, i.e. faithful; the second one
// then yields RN(v/d).  For grmsimple (v = C/m^2 with C < 2^31, d = p < 2^31) a quotient is at least 2^-32 ulp away
// from any rounding midpoint, far more than the 2^-54 ulp the corrected value can be off — the result is
// bit-identical to IEEE division (checked against `/` in tests/test_gpu_parity.py).  5 FP64 ops instead of ~30.
__device__ __forceinline__ double div_const(double v, double d, double rcp) {
    double q = v * rcp;
    double r = fma(-q, d, v);
    q = fma(r, rcp, q);
    r = fma(-q, d, v);
    return fma(r, rcp, q);
}

// EPI: 0 = store int32 C;  1 = G = (alpha*C)/denom;  2 = G = ((alpha*C - beta*(u_i+u_j)) + gamma)/denom
// One thread = one output row, 32 consecutive columns: for each column the warp writes 32 consecutive rows
// (column-major G => one coalesced 256-byte store per column).
template <int EPI>
__device__ __forceinline__ void epilogue_store(const uint32_t (&r)[32], int64_t row, bool row_ok, int64_t col0, double ui,
                                               const EpiParams &ep) {
    if (EPI == 3) return;  // packed tiles are stored by epilogue_store_packed
    if (row_ok && col0 < ep.n) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int64_t col = col0 + j;
            if (col < ep.n) {
                const int32_t c = static_cast<int32_t>(r[j]);
                if (EPI == 0) {
                    ep.C[row + col * ep.ldC] = c;
                } else {
                    double v = static_cast<double>(c) * ep.alpha;
                    if (EPI == 2) {
                        const double uj = __ldg(ep.u + col);
                        v = (v - ep.beta * (ui + uj)) + ep.gamma;
                    }
                    ep.G[row + col * ep.ldg] = (ep.rcp != 0.0) ? div_const(v, ep.denom, ep.rcp) : v / ep.denom;
                }
            }
        }
    }
}

// mode 3: the raw int32 tile, contiguous ([col][row] inside the tile; rows/cols beyond n hold exact zeros), for the
// loci-split multi-GPU path where partial tiles of different GPUs are summed by a reduce-scatter
__device__ __forceinline__ void epilogue_store_packed(const uint32_t (&r)[32], int32_t *tile, int row_local, int c0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) tile[(c0 + j) * 256 + row_local] = static_cast<int32_t>(r[j]);
}

template <int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
syrk_i8_kernel(const __grid_constant__ CUtensorMap tmap, const int2 *__restrict__ tiles, int num_tiles, int num_kb,
               int kb_per_slab, EpiParams ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full = empty_bar + STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int2 tc = tiles[t];
                int slab = 0, kin = 0;  // loci are stored as [slab][entry][locus-in-slab]
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *sA = smem + stage * STAGE_BYTES;
                    uint8_t *sB = sA + A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_3d(sA, &tmap, &full_bar[stage], kin * BK, tc.x, slab);
                    tma_load_3d(sB, &tmap, &full_bar[stage], kin * BK, tc.y, slab);
                    tma_load_3d(sB + 128 * BK, &tmap, &full_bar[stage], kin * BK, tc.y + 128, slab);
                    if (++kin == kb_per_slab) {
                        kin = 0;
                        ++slab;
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++local) {
                const int acc = local & 1;
                const uint32_t acc_phase = (local >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t adesc = umma_desc_k_sw128(sA);
                    const uint64_t bdesc = umma_desc_k_sw128(sA + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing K by 32 bytes inside the 128-byte swizzle atom = +2 in the 16-byte address field
                        umma_i8(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        int local = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++local) {
            const int acc = local & 1;
            const uint32_t acc_phase = (local >> 1) & 1;
            const int2 tc = tiles[t];
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int64_t row = static_cast<int64_t>(tc.x) + quarter * 32 + lane;
            const bool row_ok = row < ep.n;
            double ui = 0.0;
            if (EPI == 2 && row_ok) ui = ep.u[row];
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + acc * BN + c0 + (static_cast<uint32_t>(quarter * 32) << 16), r);
                tmem_wait_ld();
                epilogue_store<EPI>(r, row, row_ok, static_cast<int64_t>(tc.y) + c0, ui, ep);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2-CTA variant: a cluster of two CTAs (one TPC) computes one 256 x 256 tile with tcgen05.mma.cta_group::2
// (UMMA M = 256, N = 256).  CTA r stages rows [row0 + 128 r, +128) as A and rows [col0 + 128 r, +128) as its half of
// B, so every operand byte is read from L2 and from shared memory once per PAIR: 1/3 less L2 traffic and 1/3 less
// shared-memory traffic per MAC than the 128 x 256 single-CTA tile, and a 6-stage 32 KB ring instead of 4 x 48 KB.
// Only the leader CTA (rank 0) issues MMAs; its tcgen05.commit multicasts to the barriers of both CTAs.
// ---------------------------------------------------------------------------------------------------------------
constexpr int STAGES2 = 6;
constexpr int NUM_THREADS2 = 224;  // warps 0..5 as in the single-CTA kernel + warp 6: K-window watcher
constexpr uint32_t STAGE2_BYTES = 2 * A_BYTES;  // A (16 KiB) + this CTA's half of B (16 KiB)
constexpr size_t SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256;

// K-window synchronisation of the persistent CTA pairs.  All pairs stream the loci (K) of their current tile at the same
// nominal pace, and tiles that are in flight together share operand panels — but only while they read the SAME part of
// K: left alone, the pairs drift apart over the ~40 tiles each of them processes and the L2 working set grows from
// (panels in flight) x (a few K blocks) to (panels in flight) x (all of K): ncu measured 383 GB of DRAM reads for the
// 10 GB operand of C3 (L2 hit rate 51 %, HBM at 69 % of peak while the tensor pipe waits for power).  The leader's
// producer publishes its global K step every SYNC_PUBLISH blocks; a watcher warp (warp 6 of the leader) keeps
// min(progress of all pairs) + SYNC_WINDOW in shared memory, and the producer does not issue a load beyond it.  The
// slowest pair never waits, and with a static tile assignment the fast pairs would only have idled at the end anyway.
// The launch only enables it when the whole grid fits the idle device (cudaOccupancyMaxActiveClusters) and launches
// cooperatively so that every pair is co-resident; the wait itself is bounded and degrades to "no throttle"
// (k_window_wait), so a foreign kernel holding SMs can cost bandwidth but never hangs or traps.
// Measured on C3: 102 ms -> 66 ms.
constexpr uint32_t SYNC_PUBLISH = 8, SYNC_WINDOW_DEFAULT = 64;
constexpr bool SYRK_WIDE_DEFAULT = true;  // two tiles per CTA pair sharing the column panel (long contractions)
constexpr int SYRK_WIDE_MIN_KB = 2048;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Epilogue of one 256 x 256 tile for one warp of a CTA pair: CTA `cta` (0 / 1) owns rows 128 cta .. 128 cta + 127 of the
// tile, warp `quarter` 32 of them, one thread per row; the accumulator starts at TMEM column address `tmem_acc`.
template <int EPI>
__device__ __forceinline__ void drain_tile(const EpiParams &ep, int t, int2 tc, uint32_t tmem_acc, uint32_t cta, int quarter,
                                           int lane) {
    const int64_t row = static_cast<int64_t>(tc.x) + 128 * cta + quarter * 32 + lane;
    const bool row_ok = row < ep.n;
    double ui = 0.0;
    if (EPI == 2 && row_ok) ui = ep.u[row];
    const int64_t slot = ep.slots ? static_cast<int64_t>(ep.slots[t]) : static_cast<int64_t>(t);
    int32_t *packed = nullptr;
    if (EPI == 3) packed = ep.C + slot * 65536;
    const int row_local = 128 * static_cast<int>(cta) + quarter * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_acc + c0 + (static_cast<uint32_t>(quarter * 32) << 16), r);
        tmem_wait_ld();
        if (EPI == 3) {
            epilogue_store_packed(r, packed, row_local, c0);
        } else if ((EPI == 1 || EPI == 2) && ep.Gt != nullptr) {
            // sharded output: the same arithmetic, addressed tile-locally (leading dimension 256)
            EpiParams lt = ep;
            lt.G = ep.Gt + static_cast<int64_t>(t) * 65536 - (static_cast<int64_t>(tc.x) + static_cast<int64_t>(tc.y) * 256);
            lt.ldg = 256;
            epilogue_store<EPI>(r, row, row_ok, static_cast<int64_t>(tc.y) + c0, ui, lt);
        } else {
            epilogue_store<EPI>(r, row, row_ok, static_cast<int64_t>(tc.y) + c0, ui, ep);
        }
    }
}

// K-window throttle of a producer: publish the K step, then hold back while more than the window ahead of the slowest
// pair.  The wait is bounded: a pair that is not running (the grid is not fully co-resident because some other kernel of
// the host application holds SMs) must not stall the others for ever, let alone trap — after the bound this producer
// simply stops throttling for the rest of the kernel (the window is a bandwidth optimisation, never a correctness
// requirement).
__device__ __forceinline__ void k_window_wait(uint32_t *progress, int pair, uint32_t step, volatile uint32_t *sync_allowed,
                                              bool &throttle) {
    if ((step % SYNC_PUBLISH) == 0) st_relaxed_u32(&progress[pair], step);
    uint32_t spins = 0;
    while (step > *sync_allowed) {
        if (++spins > (1u << 22)) {  // ~10 ms: far beyond any drift between co-resident pairs
            throttle = false;
            break;
        }
    }
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
syrk_i8_2cta_kernel(const __grid_constant__ CUtensorMap tmap, const int2 *__restrict__ tiles, int num_tiles, int num_kb,
                    int kb_per_slab, EpiParams ep, uint32_t *__restrict__ progress, uint32_t sync_window) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES2 * STAGE2_BYTES);
    uint64_t *empty_bar = full_bar + STAGES2;
    uint64_t *tmem_full = empty_bar + STAGES2;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    volatile uint32_t *sync_allowed = tmem_slot + 1;  // highest K step the producer may issue
    volatile uint32_t *sync_done = tmem_slot + 2;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta = cluster_ctarank();  // rank inside the CTA pair (0 = leader: issues the MMAs)
    const bool leader = cta == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < STAGES2; ++s) {
            mbar_init(&full_bar[s], 1);   // the leader's arrive.expect_tx covers the bytes of BOTH CTAs (leader only)
            mbar_init(&empty_bar[s], 1);  // one multicast tcgen05.commit per use, in each CTA
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);   // multicast commit, in each CTA
            mbar_init(&tmem_empty[a], 8);  // 4 epilogue warps of each CTA (used in the leader only)
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
    if (threadIdx.x == 0) {
        *sync_allowed = sync_window;
        *sync_done = 0;
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, step = 0;
            const bool publish = progress != nullptr && leader;
            bool throttle = publish;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                const int2 tc = tiles[t];
                int slab = 0, kin = 0;
                for (int kb = 0; kb < num_kb; ++kb, ++step) {
                    if (throttle) k_window_wait(progress, pair, step, sync_allowed, throttle);
                    else if (publish && (step % SYNC_PUBLISH) == 0) st_relaxed_u32(&progress[pair], step);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *sA = smem + stage * STAGE2_BYTES;
                    uint8_t *sB = sA + A_BYTES;
                    // The peer's bytes may land before the leader arms the barrier: the tx-count goes transiently
                    // negative, the phase cannot complete before the leader's (only) arrival.
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);
                    tma_load_3d_2cta(sA, &tmap, &full_bar[stage], kin * BK, tc.x + 128 * static_cast<int>(cta), slab);
                    tma_load_3d_2cta(sB, &tmap, &full_bar[stage], kin * BK, tc.y + 128 * static_cast<int>(cta), slab);
                    if (++kin == kb_per_slab) {
                        kin = 0;
                        ++slab;
                    }
                    if (++stage == STAGES2) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (publish) {
                st_relaxed_u32(&progress[pair], 0xffffffffu);  // done: nobody waits for this pair
                *sync_done = 1;
            }
        }
    } else if (warp == 6) {
        // ===================== K-window watcher (leader CTA) =====================
        if (progress != nullptr && leader) {
            while (*sync_done == 0) {
                uint32_t mn = 0xffffffffu;
                for (int i = lane; i < num_pairs; i += 32) mn = min(mn, ld_relaxed_u32(&progress[i]));
                mn = __reduce_min_sync(0xffffffffu, mn);
                if (lane == 0) *sync_allowed = (mn > 0xffffffffu - sync_window) ? 0xffffffffu : mn + sync_window;
                __nanosleep(2000);  // ~8 K blocks: coarse on purpose, 74 watchers poll the same three cache lines
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(256, 256);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int t = pair; t < num_tiles; t += num_pairs, ++local) {
                const int acc = local & 1;
                const uint32_t acc_phase = (local >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * STAGE2_BYTES);
                    const uint64_t adesc = umma_desc_k_sw128(sA);
                    const uint64_t bdesc = umma_desc_k_sw128(sA + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_i8_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_2cta(&empty_bar[stage], 0b11);
                    if (++stage == STAGES2) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_2cta(&tmem_full[acc], 0b11);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 of both CTAs; CTA r owns rows 128 r .. 128 r + 127) ===========
        const int quarter = warp & 3;
        int local = 0;
        for (int t = pair; t < num_tiles; t += num_pairs, ++local) {
            const int acc = local & 1;
            const uint32_t acc_phase = (local >> 1) & 1;
            const int2 tc = tiles[t];
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            drain_tile<EPI>(ep, t, tc, tmem_base + acc * 256, cta, quarter, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
            if (EPI == 3 && ep.wave_done != nullptr) {
                __threadfence();  // this warp's stores of the tile are visible device-wide before it is counted
                if (lane == 0) {
                    int w = 0;
                    while (w + 1 < ep.nwaves && t >= ep.wave_end[w]) ++w;
                    atomicAdd(ep.wave_done + w, 1u);
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer's shared memory and TMEM stay alive until the leader's last MMA has retired
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// "Wide" CTA-pair kernel for long contractions: a pair works on TWO tiles that share their column panel at once
// (units from make_wide_units: consecutive tiles (ti, tj), (ti+1, tj) of the list).  Per K block a CTA then stages its
// 128 rows of both row panels and its half of the ONE column panel: 48 KB for two tiles' worth of MMAs instead of
// 2 x 32 KB — 25 % fewer bytes through the TMA / L2 path that limits the pair kernel — with no coupling between pairs
// (the quad variant pays for its multicast with a lock-step).  Both 256-column accumulators are live, so the epilogue
// cannot overlap the next unit's main loop: ~2 x 10 us per unit, which is why this form is only chosen when a unit's
// main loop is milliseconds long (launch()).  4 stages of 48 KB; everything else as in syrk_pair_body.
// ---------------------------------------------------------------------------------------------------------------
constexpr int STAGESW = 4;
constexpr uint32_t STAGEW_BYTES = 3 * A_BYTES;  // row panel 0, row panel 1, half of the column panel: 48 KiB
constexpr size_t SMEMW_BYTES = STAGESW * STAGEW_BYTES + 1024 + 256;

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
syrk_i8_wide_kernel(const __grid_constant__ CUtensorMap tmap, const int2 *__restrict__ tiles, const int2 *__restrict__ units,
                    int num_units, int num_kb, int kb_per_slab, EpiParams ep, uint32_t *__restrict__ progress,
                    uint32_t sync_window) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGESW * STAGEW_BYTES);
    uint64_t *empty_bar = full_bar + STAGESW;
    uint64_t *tmem_full = empty_bar + STAGESW;   // one accumulator generation at a time
    uint64_t *tmem_empty = tmem_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 1);
    volatile uint32_t *sync_allowed = tmem_slot + 1;
    volatile uint32_t *sync_done = tmem_slot + 2;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta = cluster_ctarank();
    const bool leader = cta == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < STAGESW; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 8);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
    if (threadIdx.x == 0) {
        *sync_allowed = sync_window;
        *sync_done = 0;
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, step = 0;
            const bool publish = progress != nullptr && leader;
            bool throttle = publish;
            for (int u = pair; u < num_units; u += num_pairs) {
                const int2 un = units[u];
                const bool two = un.y >= 0;
                const int2 t0 = tiles[un.x];
                const int2 t1 = two ? tiles[un.y] : t0;
                const uint32_t bytes = 2 * (two ? STAGEW_BYTES : 2 * A_BYTES);
                int slab = 0, kin = 0;
                for (int kb = 0; kb < num_kb; ++kb, ++step) {
                    if (throttle) k_window_wait(progress, pair, step, sync_allowed, throttle);
                    else if (publish && (step % SYNC_PUBLISH) == 0) st_relaxed_u32(&progress[pair], step);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *sA0 = smem + stage * STAGEW_BYTES;
                    uint8_t *sA1 = sA0 + A_BYTES;
                    uint8_t *sB = sA1 + A_BYTES;
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], bytes);
                    tma_load_3d_2cta(sA0, &tmap, &full_bar[stage], kin * BK, t0.x + 128 * static_cast<int>(cta), slab);
                    if (two) tma_load_3d_2cta(sA1, &tmap, &full_bar[stage], kin * BK, t1.x + 128 * static_cast<int>(cta), slab);
                    tma_load_3d_2cta(sB, &tmap, &full_bar[stage], kin * BK, t0.y + 128 * static_cast<int>(cta), slab);
                    if (++kin == kb_per_slab) {
                        kin = 0;
                        ++slab;
                    }
                    if (++stage == STAGESW) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            if (publish) {
                st_relaxed_u32(&progress[pair], 0xffffffffu);
                *sync_done = 1;
            }
        }
    } else if (warp == 6) {
        // ===================== K-window watcher (leader CTA) =====================
        if (progress != nullptr && leader) {
            while (*sync_done == 0) {
                uint32_t mn = 0xffffffffu;
                for (int i = lane; i < num_pairs; i += 32) mn = min(mn, ld_relaxed_u32(&progress[i]));
                mn = __reduce_min_sync(0xffffffffu, mn);
                if (lane == 0) *sync_allowed = (mn > 0xffffffffu - sync_window) ? 0xffffffffu : mn + sync_window;
                __nanosleep(2000);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(256, 256);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int u = pair; u < num_units; u += num_pairs, ++local) {
                const bool two = units[u].y >= 0;
                mbar_wait(tmem_empty, (local & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sA0 = smem_u32(smem + stage * STAGEW_BYTES);
                    const uint64_t a0 = umma_desc_k_sw128(sA0);
                    const uint64_t a1 = umma_desc_k_sw128(sA0 + A_BYTES);
                    const uint64_t bd = umma_desc_k_sw128(sA0 + 2 * A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_i8_2cta(tmem_base, a0 + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                        if (two) umma_i8_2cta(tmem_base + 256, a1 + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit_2cta(&empty_bar[stage], 0b11);
                    if (++stage == STAGESW) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_2cta(tmem_full, 0b11);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 of both CTAs) =====================
        const int quarter = warp & 3;
        int local = 0;
        for (int u = pair; u < num_units; u += num_pairs, ++local) {
            const int2 un = units[u];
            mbar_wait(tmem_full, local & 1);
            tc_fence_after();
            drain_tile<EPI>(ep, un.x, tiles[un.x], tmem_base, cta, quarter, lane);
            if (un.y >= 0) drain_tile<EPI>(ep, un.y, tiles[un.y], tmem_base + 256, cta, quarter, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty, 0);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
    }
}

int32_t make_codes_tmap(const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc, int64_t nslabs, CUtensorMap *out) {
    GrmEncodeTiledFn encode = nullptr;
    GRM_TRY(grm_tmap_encoder(&encode));
    // dim0 = loci of one slab (bytes, contiguous), dim1 = entries, dim2 = slab.  Out-of-range rows/loci are
    // zero-filled by TMA, so ragged n (not a multiple of 128) and p (not a multiple of 128) contribute exact zeros
    // to the integer sums.
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(p), static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(nslabs)};
    cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ldc), static_cast<cuuint64_t>(ldc) * static_cast<cuuint64_t>(n)};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(BK), 128u, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t *>(d_codes), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        grm_set_error("cuTensorMapEncodeTiled failed with CUresult %d (n=%lld p=%lld ldc=%lld)", (int)r, (long long)n,
                      (long long)p, (long long)ldc);
        return GRM_CUDA_ERR_CUDA;
    }
    return GRM_CUDA_OK;
}

int32_t upload_list(grm_cuda_ctx *ctx, TileListCache &tc, const void *data, size_t bytes) {
    GRM_TRY(tc.buf.reserve(std::max<size_t>(bytes, 16)));
    if (bytes) {
        // pageable -> device copy is synchronous with respect to the host buffer, safe to free afterwards
        GRM_CUDA_TRY(cudaMemcpyAsync(tc.buf.ptr, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return GRM_CUDA_OK;
}

bool list_cached(const grm_cuda_ctx *ctx, TileListCache &tc, int64_t n, int32_t rank, int32_t world) {
    return tc.n == n && tc.rank == rank && tc.world == world && tc.kind == ctx->tune.tile_band;
}
void list_stamp(const grm_cuda_ctx *ctx, TileListCache &tc, int64_t n, int32_t rank, int32_t world, size_t count) {
    tc.n = n;
    tc.rank = rank;
    tc.world = world;
    tc.kind = ctx->tune.tile_band;
    tc.count = static_cast<int32_t>(count);
}

// 128 x 256 CTA tiles of the owned 256 x 256 sharding tiles, in the L2-friendly order of grm_tile_order().
int32_t get_tile_list(grm_cuda_ctx *ctx, int64_t n, int32_t rank, int32_t world, const int2 **d_tiles, int *count) {
    TileListCache &tc = ctx->tiles_i8;
    if (!list_cached(ctx, tc, n, rank, world)) {
        std::vector<int2> order;
        grm_tile_order(n, ctx->tune.tile_band, order);
        int64_t b, e;
        grm_tile_range(static_cast<int64_t>(order.size()), rank, world, &b, &e);
        std::vector<int2> cta;
        cta.reserve(2 * (e - b));
        for (int64_t t = b; t < e; ++t) {
            const int row0 = order[t].x * GRM_TILE, col0 = order[t].y * GRM_TILE;
            cta.push_back(make_int2(row0, col0));
            if (row0 + BM < n) cta.push_back(make_int2(row0 + BM, col0));
        }
        GRM_TRY(upload_list(ctx, tc, cta.data(), cta.size() * sizeof(int2)));
        list_stamp(ctx, tc, n, rank, world, cta.size());
    }
    *d_tiles = tc.buf.as<int2>();
    *count = tc.count;
    return GRM_CUDA_OK;
}

// 256 x 256 tiles (one per CTA pair) of the owned sharding tiles
int32_t get_tile_list_2cta(grm_cuda_ctx *ctx, int64_t n, int32_t rank, int32_t world, const int2 **d_tiles, int *count) {
    TileListCache &tc = ctx->tiles_i8_2cta;
    if (!list_cached(ctx, tc, n, rank, world)) {
        std::vector<int2> order;
        grm_tile_order(n, ctx->tune.tile_band, order);
        int64_t b, e;
        grm_tile_range(static_cast<int64_t>(order.size()), rank, world, &b, &e);
        std::vector<int2> tl;
        for (int64_t t = b; t < e; ++t) tl.push_back(make_int2(order[t].x * GRM_TILE, order[t].y * GRM_TILE));
        GRM_TRY(upload_list(ctx, tc, tl.data(), tl.size() * sizeof(int2)));
        list_stamp(ctx, tc, n, rank, world, tl.size());
    }
    *d_tiles = tc.buf.as<int2>();
    *count = tc.count;
    return GRM_CUDA_OK;
}

// units of the wide kernel: two consecutive tiles of a list that share their column panel, or one leftover tile
void make_wide_units(const std::vector<int2> &tiles, size_t t0, size_t t1, std::vector<int2> &units) {
    for (size_t t = t0; t < t1;) {
        if (t + 1 < t1 && tiles[t].y == tiles[t + 1].y) {
            units.push_back(make_int2(static_cast<int>(t), static_cast<int>(t + 1)));
            t += 2;
        } else {
            units.push_back(make_int2(static_cast<int>(t), -1));
            t += 1;
        }
    }
}

int32_t get_wide_units(grm_cuda_ctx *ctx, int64_t n, int32_t rank, int32_t world, const int2 **d_units, int *num_units) {
    TileListCache &tc = ctx->tiles_wide;
    if (!list_cached(ctx, tc, n, rank, world)) {
        std::vector<int2> order;
        grm_tile_order(n, ctx->tune.tile_band, order);
        int64_t b, e;
        grm_tile_range(static_cast<int64_t>(order.size()), rank, world, &b, &e);
        std::vector<int2> own(order.begin() + b, order.begin() + e), units;
        make_wide_units(own, 0, own.size(), units);
        GRM_TRY(upload_list(ctx, tc, units.data(), units.size() * sizeof(int2)));
        list_stamp(ctx, tc, n, rank, world, units.size());
    }
    *d_units = tc.buf.as<int2>();
    *num_units = tc.count;
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Launch of a persistent CTA-pair kernel, with or without the K-window synchronisation.
//
// The window pays when drift can build up AND the operand is far larger than L2: many tiles per pair and a multi-GB
// operand (C3 on one GPU: 10 GB, 43 tiles per pair: 73 -> 54 ms).  Measured neutral at 5 GB (C3 on 2 GPUs) and a loss
// for short launches (C2: +5 %; C3 on 4 / 8 GPUs: +11 %), so it is enabled from 6 GB up (tune.syrk_sync forces it
// on / off).  It is only used when every pair of the grid can be co-resident: the capacity of the idle device is
// checked (cudaOccupancyMaxActiveClusters) and the launch is made cooperative, so the driver starts the grid only
// once all of it fits next to whatever else the host application is running.  If the cooperative launch is refused
// the kernel is launched without the window.
// ---------------------------------------------------------------------------------------------------------------
bool want_window(const grm_cuda_ctx *ctx, int pairs, int units, int64_t operand_bytes) {
    if (ctx->tune.syrk_sync == 0) return false;
    if (ctx->tune.syrk_sync == 1) return true;
    return units >= 6 * pairs && operand_bytes >= (6ll << 30);
}

uint32_t window_blocks(const grm_cuda_ctx *ctx) {
    const int v = ctx->tune.syrk_window;
    return (v > 0 && v < 65536) ? static_cast<uint32_t>(v) : SYNC_WINDOW_DEFAULT;
}

int32_t max_coresident_pairs(grm_cuda_ctx *ctx, const void *kernel, size_t smem, int *out) {
    for (auto &e : ctx->max_clusters)
        if (e.first == kernel) {
            *out = e.second;
            return GRM_CUDA_OK;
        }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(ctx->num_sms / 2 * 2));
    cfg.blockDim = dim3(NUM_THREADS2);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int mc = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&mc, kernel, &cfg);
    if (e != cudaSuccess) {
        cudaGetLastError();
        mc = 0;  // unknown: never enable the window
    }
    ctx->max_clusters.emplace_back(kernel, mc);
    *out = mc;
    return GRM_CUDA_OK;
}

// go(cfg, progress) launches the kernel with cudaLaunchKernelEx(cfg, kernel, ..., progress, window)
template <class Go>
int32_t launch_pairs(grm_cuda_ctx *ctx, const void *kernel, size_t smem, int pairs, int units, int64_t operand_bytes,
                     Go &&go) {
    GRM_TRY(grm_func_smem(ctx, kernel, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs));
    cfg.blockDim = dim3(NUM_THREADS2);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    bool window = want_window(ctx, pairs, units, operand_bytes);
    if (window) {
        int cap = 0;
        GRM_TRY(max_coresident_pairs(ctx, kernel, smem, &cap));
        if (cap < pairs) window = false;
    }
    if (window) {
        GRM_TRY(ctx->progress.reserve(static_cast<size_t>(pairs) * sizeof(uint32_t) + 64));
        GRM_CUDA_TRY(cudaMemsetAsync(ctx->progress.ptr, 0, static_cast<size_t>(pairs) * sizeof(uint32_t), ctx->stream));
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (go(&cfg, ctx->progress.as<uint32_t>()) == cudaSuccess) {
            ctx->launches++;
            ctx->last_syrk_window = 1;
            return GRM_CUDA_OK;
        }
        cudaGetLastError();  // refused (not co-schedulable / unsupported combination): run without the window
    }
    ctx->last_syrk_window = 0;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    GRM_CUDA_TRY(go(&cfg, nullptr));
    ctx->launches++;
    return GRM_CUDA_OK;
}

// wide kernel: used when a unit's main loop is long enough to dwarf its un-overlapped epilogue (>= SYRK_WIDE_MIN_KB K
// blocks) and there are enough tiles to keep every pair busy with two-tile units.  Measured: C3 (3907 K blocks, 3160
// tiles) SYRK 56.5 -> 51.7 ms; C2 (781 K blocks, 210 tiles) 0.76 -> 0.97 ms, hence the thresholds.
bool use_wide(const grm_cuda_ctx *ctx, int num_kb, int num_tiles) {
    if (ctx->tune.syrk_i8 != 0) return ctx->tune.syrk_i8 == 3;
    return SYRK_WIDE_DEFAULT && num_kb >= SYRK_WIDE_MIN_KB && num_tiles >= 4 * (ctx->num_sms / 2);
}

template <int EPI>
int32_t launch(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc, int64_t nslabs,
               const EpiParams &ep, int32_t rank, int32_t world) {
    if (!ctx || !d_codes || n < 1 || p < 1 || ldc < p || (ldc % 16) != 0 || nslabs < 1 || world < 1 || rank < 0 ||
        rank >= world) {
        grm_set_error("syrk_i8: bad argument (n=%lld p=%lld ldc=%lld rank=%d world=%d)", (long long)n, (long long)p,
                      (long long)ldc, rank, world);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    CUtensorMap tmap;
    GRM_TRY(make_codes_tmap(d_codes, n, p, ldc, nslabs, &tmap));
    const int kbs = static_cast<int>((p + BK - 1) / BK);
    const int num_kb = kbs * static_cast<int>(nslabs);
    const int64_t operand = n * ldc * nslabs;
    if (ctx->tune.syrk_i8 == 1) {  // single-CTA variant (kept as the reference point of the pair kernels)
        const int2 *d_tiles;
        int count;
        GRM_TRY(get_tile_list(ctx, n, rank, world, &d_tiles, &count));
        if (count == 0) return GRM_CUDA_OK;
        GRM_TRY(grm_func_smem(ctx, reinterpret_cast<const void *>(syrk_i8_kernel<EPI>), SMEM_BYTES));
        const int grid = std::min(count, ctx->num_sms);
        syrk_i8_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, ctx->stream>>>(tmap, d_tiles, count, num_kb, kbs, ep);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        return GRM_CUDA_OK;
    }
    const int2 *d_tiles;
    int count;
    GRM_TRY(get_tile_list_2cta(ctx, n, rank, world, &d_tiles, &count));
    if (count == 0) return GRM_CUDA_OK;
    const uint32_t win = window_blocks(ctx);
    if (use_wide(ctx, num_kb, count)) {
        const int2 *d_units;
        int nunits;
        GRM_TRY(get_wide_units(ctx, n, rank, world, &d_units, &nunits));
        const int pairs = std::min(nunits, ctx->num_sms / 2);
        return launch_pairs(ctx, reinterpret_cast<const void *>(syrk_i8_wide_kernel<EPI>), SMEMW_BYTES, pairs, nunits,
                            operand, [&](const cudaLaunchConfig_t *cfg, uint32_t *prog) {
                                return cudaLaunchKernelEx(cfg, syrk_i8_wide_kernel<EPI>, tmap, d_tiles, d_units, nunits,
                                                          num_kb, kbs, ep, prog, win);
                            });
    }
    const int pairs = std::min(count, ctx->num_sms / 2);
    return launch_pairs(ctx, reinterpret_cast<const void *>(syrk_i8_2cta_kernel<EPI>), SMEM2_BYTES, pairs, count, operand,
                        [&](const cudaLaunchConfig_t *cfg, uint32_t *prog) {
                            return cudaLaunchKernelEx(cfg, syrk_i8_2cta_kernel<EPI>, tmap, d_tiles, count, num_kb, kbs,
                                                      ep, prog, win);
                        });
}

void set_division(const grm_cuda_ctx *ctx, EpiParams &ep, double denom) {
    // fast exact division only for ordinary divisors; d = 0 (all loci fixed -> NaN GRM, calc.jl:52-54), Inf, NaN
    // or extreme magnitudes take the IEEE divide so that the special values come out exactly as in the reference
    const double a = fabs(denom);
    ep.denom = denom;
    ep.rcp = (a > 1e-200 && a < 1e200 && !(ctx && ctx->tune.ieee_div)) ? 1.0 / denom : 0.0;
}

}  // namespace

extern "C" int32_t grm_cuda_syrk_i8(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                    int64_t nslabs, int32_t *d_C, int64_t ldC, int32_t rank, int32_t world) try {
    if (!d_C || ldC < n) {
        grm_set_error("syrk_i8: bad output (ldC=%lld n=%lld)", (long long)ldC, (long long)n);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    EpiParams ep{};
    ep.n = n;
    ep.C = d_C;
    ep.ldC = ldC;
    return launch<0>(ctx, d_codes, n, p, ldc, nslabs, ep, rank, world);
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_syrk_i8_f64(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                        int64_t nslabs, double alpha, double beta, double gamma, double denom, const double *d_u,
                                        double *d_G, int64_t ldg, int32_t rank, int32_t world) try {
    if (!d_G || ldg < n || (beta != 0.0 && !d_u)) {
        grm_set_error("syrk_i8_f64: bad output or missing u (ldg=%lld n=%lld)", (long long)ldg, (long long)n);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    EpiParams ep{};
    ep.n = n;
    ep.G = d_G;
    ep.ldg = ldg;
    ep.alpha = alpha;
    ep.beta = beta;
    ep.gamma = gamma;
    set_division(ctx, ep, denom);
    ep.u = d_u;
    if (beta == 0.0 && gamma == 0.0) return launch<1>(ctx, d_codes, n, p, ldc, nslabs, ep, rank, world);
    return launch<2>(ctx, d_codes, n, p, ldc, nslabs, ep, rank, world);
} GRM_CATCH_STATUS

// Sharded-output form of grm_cuda_syrk_i8_f64: the tiles owned by (rank, world) are written as packed 256 x 256 FP64
// tiles, d_tiles[t][col][row] for the t-th owned tile (grm_cuda_tile_at), zeros outside the n x n matrix.  For outputs
// that do not fit one device as a dense matrix (BASELINE C5: n = 100,000 -> 80 GB, 10 GB per GPU on 8 GPUs).
extern "C" int32_t grm_cuda_syrk_i8_f64_tiles(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                              int64_t nslabs, double alpha, double beta, double gamma, double denom,
                                              const double *d_u, double *d_tiles, int32_t rank, int32_t world) try {
    if (!ctx || !d_tiles || (beta != 0.0 && !d_u) || n < 1 || world < 1 || rank < 0 || rank >= world) {
        grm_set_error("syrk_i8_f64_tiles: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t count = grm_cuda_tile_count(n, rank, world);
    if (count == 0) return GRM_CUDA_OK;
    // tiles at the ragged edge are only partly written by the epilogue: define the rest as zeros
    GRM_CUDA_TRY(cudaMemsetAsync(d_tiles, 0, static_cast<size_t>(count) * 65536 * sizeof(double), ctx->stream));
    EpiParams ep{};
    ep.n = n;
    ep.Gt = d_tiles;
    ep.alpha = alpha;
    ep.beta = beta;
    ep.gamma = gamma;
    set_division(ctx, ep, denom);
    ep.u = d_u;
    // the packed-tile addressing lives in the pair kernels' epilogue (drain_tile): never the single-CTA variant
    const int saved = ctx->tune.syrk_i8;
    if (saved == 1) ctx->tune.syrk_i8 = 2;
    const int32_t st = (beta == 0.0 && gamma == 0.0) ? launch<1>(ctx, d_codes, n, p, ldc, nslabs, ep, rank, world)
                                                     : launch<2>(ctx, d_codes, n, p, ldc, nslabs, ep, rank, world);
    ctx->tune.syrk_i8 = saved;
    return st;
} GRM_CATCH_STATUS

// ---------------------------------------------------------------------------------------------------------------
// loci-split (K-split) multi-GPU support: every GPU contracts ITS loci over ALL upper-triangular tiles and writes raw
// int32 tiles into a buffer laid out for a reduce-scatter; after the exchange each GPU finalises the tiles it owns.
// Exchange volume 4 n^2 / 2 bytes per GPU instead of n p for the all-gather of the codes — the better choice whenever
// p > 4 n (BASELINE C2, C3).
//
// Waves: the slots 0 .. tiles_max-1 every rank owns are cut into `waves` equal ranges [s_w, s_{w+1}); wave w of the packed
// buffer is the contiguous block [world][s_{w+1} - s_w][256 x 256] starting at element world * s_w * 65536, so the
// reduce-scatter of wave w is one plain call.  Default flow (run.cu): one launch per wave (grm_syrk_i8_packed_wave), the
// reduce-scatter of wave w on a second stream next to the launch of wave w + 1.  Streamed flow (tuning "exchange" = 1 / 2):
// ONE persistent launch contracts all tiles in wave order (grm_syrk_i8_packed_stream) and counts finished tiles per
// wave; a gate kernel on the communication stream (grm_wave_gate) releases the exchange of wave w as soon as its tiles
// are stored.  Measured on C3 / 8 GPUs the streamed flow loses: the persistent grid owns every SM, so NCCL's kernels
// only run once it retires (10.4 ms per step against 9.85), reserving SMs for a CTA-capped NCCL costs more than it hides
// (11.3 ms), and the copy engines are slow when all GPUs send at once (10.3 ms).  waves = 1 is the layout
// [world][tiles_max][256 x 256] of the single-exchange form.
// ---------------------------------------------------------------------------------------------------------------
void grm_packed_wave_range(int64_t n, int32_t world, int32_t waves, int32_t w, int64_t *s0, int64_t *s1) {
    // equal waves: each is a kernel launch of its own in the default flow, so each should fill the CTA pairs as evenly as
    // the others (shrinking waves were tried with the single-launch flow and do not pay there either, profiles/README.md)
    const int64_t tmax = grm_cuda_packed_tiles_max(n, world);
    *s0 = tmax * w / waves;
    *s1 = tmax * (w + 1) / waves;
}

namespace {

int32_t get_packed_plan_impl(grm_cuda_ctx *ctx, int64_t n, int32_t world, int32_t waves, PackedPlanCache **out) {
    PackedPlanCache &pc = ctx->packed_plan[waves == 1 ? 1 : 0];
    if (pc.n != n || pc.world != world || pc.waves != waves || pc.band != ctx->tune.tile_band) {
        std::vector<int2> order;
        grm_tile_order(n, ctx->tune.tile_band, order);
        const int64_t total = static_cast<int64_t>(order.size());
        std::vector<int2> tiles, units;
        std::vector<int32_t> slots;
        pc.tile_off.assign(waves, 0);
        pc.tile_cnt.assign(waves, 0);
        pc.unit_off.assign(waves, 0);
        pc.unit_cnt.assign(waves, 0);
        for (int32_t w = 0; w < waves; ++w) {
            int64_t s0, s1;
            grm_packed_wave_range(n, world, waves, w, &s0, &s1);
            const int64_t len = s1 - s0;
            pc.tile_off[w] = static_cast<int32_t>(tiles.size());
            for (int32_t r = 0; r < world; ++r) {
                int64_t b, e;
                grm_tile_range(total, r, world, &b, &e);
                for (int64_t j = s0; j < s1 && b + j < e; ++j) {
                    tiles.push_back(make_int2(order[b + j].x * GRM_TILE, order[b + j].y * GRM_TILE));
                    slots.push_back(static_cast<int32_t>(world * s0 + r * len + (j - s0)));
                }
            }
            pc.tile_cnt[w] = static_cast<int32_t>(tiles.size()) - pc.tile_off[w];
            // units of the wide kernel: index the wave's own tile sub-list
            std::vector<int2> sub(tiles.begin() + pc.tile_off[w], tiles.end()), u;
            make_wide_units(sub, 0, sub.size(), u);
            pc.unit_off[w] = static_cast<int32_t>(units.size());
            pc.unit_cnt[w] = static_cast<int32_t>(u.size());
            units.insert(units.end(), u.begin(), u.end());
        }
        const size_t nt = tiles.size(), slots_padded = (nt + 1) / 2 * 2;
        pc.total_tiles = static_cast<int32_t>(nt);
        pc.slots_padded = static_cast<int32_t>(slots_padded);
        // layout: [tiles: nt int2][slots: nt int32, padded to a multiple of 2][units: nu int2]
        GRM_TRY(pc.buf.reserve(std::max<size_t>(16, nt * sizeof(int2) + slots_padded * sizeof(int32_t) + units.size() * sizeof(int2))));
        if (nt) {
            GRM_CUDA_TRY(cudaMemcpyAsync(pc.buf.ptr, tiles.data(), nt * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
            GRM_CUDA_TRY(cudaMemcpyAsync(pc.buf.as<int2>() + nt, slots.data(), nt * sizeof(int32_t), cudaMemcpyHostToDevice,
                                         ctx->stream));
            GRM_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<int32_t *>(pc.buf.as<int2>() + nt) + slots_padded, units.data(),
                                         units.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
            GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        pc.n = n;
        pc.world = world;
        pc.waves = waves;
        pc.band = ctx->tune.tile_band;
    }
    *out = &pc;
    return GRM_CUDA_OK;
}

// G tile <- formula(reduced int32 tile): block = one owned tile, same arithmetic as the fused SYRK epilogue.
// Gt != nullptr: the tile is written packed ([col][row], zeros outside n) at Gt + (t0 + blockIdx.x) * 65536 instead
// of into the dense matrix (the form the all-gather of the finished tiles moves).
template <int EPI>
__global__ void __launch_bounds__(256)
finalize_tiles_kernel(const int32_t *__restrict__ packed, const int2 *__restrict__ tiles, EpiParams ep, int t0) {
    const int t = t0 + blockIdx.x;
    const int2 tc = tiles[t];
    const int32_t *tile = packed + static_cast<int64_t>(t) * 65536;
    double *gt = ep.Gt ? ep.Gt + static_cast<int64_t>(t) * 65536 : nullptr;
    for (int idx = threadIdx.x; idx < 65536; idx += 256) {
        const int cl = idx >> 8, rl = idx & 255;
        const int64_t row = tc.x + rl, col = tc.y + cl;
        double out = 0.0;
        const bool inside = row < ep.n && col < ep.n;
        if (inside) {
            const int32_t c = tile[idx];
            double v = static_cast<double>(c) * ep.alpha;
            if (EPI == 2) v = (v - ep.beta * (__ldg(ep.u + row) + __ldg(ep.u + col))) + ep.gamma;
            out = (ep.rcp != 0.0) ? div_const(v, ep.denom, ep.rcp) : v / ep.denom;
        }
        if (gt) gt[idx] = out;
        else if (inside) ep.G[row + col * ep.ldg] = out;
    }
}

}  // namespace

int32_t grm_packed_plan(grm_cuda_ctx *ctx, int64_t n, int32_t world, int32_t waves, PackedPlanCache **out) {
    return get_packed_plan_impl(ctx, n, world, waves, out);
}

extern "C" int64_t grm_cuda_packed_tiles_max(int64_t n, int32_t world) try {
    if (n < 1 || world < 1) return -1;
    const int64_t T = (n + GRM_TILE - 1) / GRM_TILE;
    return (T * (T + 1) / 2 + world - 1) / world;
} GRM_CATCH_I64

// one wave of the K-split contraction (waves = 1, wave = 0: everything)
int32_t grm_syrk_i8_packed_wave(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc, int32_t world,
                                int32_t waves, int32_t wave, int32_t *d_packed) {
    if (!ctx || !d_codes || !d_packed || n < 1 || p < 1 || ldc < p || (ldc % 16) != 0 || world < 1 || waves < 1 ||
        wave < 0 || wave >= waves) {
        grm_set_error("syrk_i8_packed: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    CUtensorMap tmap;
    GRM_TRY(make_codes_tmap(d_codes, n, p, ldc, 1, &tmap));
    PackedPlanCache *pc;
    GRM_TRY(grm_packed_plan(ctx, n, world, waves, &pc));
    const int64_t tiles_max = grm_cuda_packed_tiles_max(n, world);
    if (wave == 0 && static_cast<int64_t>(pc->total_tiles) != tiles_max * world) {
        // unused slots of the reduce-scatter layout must be defined (they are summed too)
        GRM_CUDA_TRY(cudaMemsetAsync(d_packed, 0, static_cast<size_t>(tiles_max) * world * 65536 * sizeof(int32_t),
                                     ctx->stream));
    }
    const int count = pc->tile_cnt[wave];
    if (count == 0) return GRM_CUDA_OK;
    const int2 *d_tiles = pc->buf.as<int2>() + pc->tile_off[wave];
    EpiParams ep{};
    ep.n = n;
    ep.C = d_packed;
    ep.slots = reinterpret_cast<const int32_t *>(pc->buf.as<int2>() + pc->total_tiles) + pc->tile_off[wave];
    const int kbs = static_cast<int>((p + BK - 1) / BK);
    const uint32_t win = window_blocks(ctx);
    if (ctx->tune.syrk_packed == 3) {
        // wide form on the packed path: measured no gain at 2 GPUs (profiles/README.md), kept selectable for the tests
        const int nunits = pc->unit_cnt[wave];
        const int2 *d_units = reinterpret_cast<const int2 *>(reinterpret_cast<const int32_t *>(pc->buf.as<int2>() + pc->total_tiles) +
                                                             pc->slots_padded) + pc->unit_off[wave];
        const int wpairs = std::min(nunits, ctx->num_sms / 2);
        return launch_pairs(ctx, reinterpret_cast<const void *>(syrk_i8_wide_kernel<3>), SMEMW_BYTES, wpairs, nunits, n * ldc,
                            [&](const cudaLaunchConfig_t *cfg, uint32_t *prog) {
                                return cudaLaunchKernelEx(cfg, syrk_i8_wide_kernel<3>, tmap, d_tiles, d_units, nunits, kbs,
                                                          kbs, ep, prog, win);
                            });
    }
    const int pairs = std::min(count, ctx->num_sms / 2);
    return launch_pairs(ctx, reinterpret_cast<const void *>(syrk_i8_2cta_kernel<3>), SMEM2_BYTES, pairs, count, n * ldc,
                        [&](const cudaLaunchConfig_t *cfg, uint32_t *prog) {
                            return cudaLaunchKernelEx(cfg, syrk_i8_2cta_kernel<3>, tmap, d_tiles, count, kbs, kbs, ep, prog,
                                                      win);
                        });
}

extern "C" int32_t grm_cuda_syrk_i8_packed(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                           int32_t world, int32_t *d_packed) try {
    return grm_syrk_i8_packed_wave(ctx, d_codes, n, p, ldc, world, 1, 0, d_packed);
} GRM_CATCH_STATUS

namespace {
__global__ void wave_gate_kernel(const uint32_t *counter, uint32_t target) {
    // bounded: if the producer never gets there (its launch failed) the stream must still drain — the caller already
    // holds the launch error.  ~2^22 x 1 us.
    if (threadIdx.x == 0) {
        for (uint32_t spins = 0; spins < (1u << 22); ++spins) {
            uint32_t v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            __nanosleep(1000);
        }
    }
}
}  // namespace

namespace {
// all `count` flags (32-bit words, written by the peers' copy engines after their data) have reached `epoch`
__global__ void flags_gate_kernel(const uint32_t *flags, int count, uint32_t epoch) {
    if (static_cast<int>(threadIdx.x) < count) {
        for (uint32_t spins = 0; spins < (1u << 22); ++spins) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
            if (static_cast<int32_t>(v - epoch) >= 0) break;  // epochs only grow (wrap-safe comparison)
            __nanosleep(500);
        }
    }
}

__global__ void set_u32_kernel(uint32_t *p, uint32_t v) { *p = v; }

// G tile <- formula(sum over the `world` parts of the owned tile t): parts[q][t][256 x 256] int32, exact sum in the
// fixed order q = 0 .. world-1, then the arithmetic of the fused SYRK epilogue; written packed at Gt + t * 65536
template <int EPI>
__global__ void __launch_bounds__(256)
finalize_parts_kernel(const int32_t *__restrict__ parts, int64_t part_stride, int world, const int2 *__restrict__ tiles, EpiParams ep,
                      int t0) {
    const int t = t0 + blockIdx.x;
    const int2 tc = tiles[t];
    const int32_t *base = parts + static_cast<int64_t>(t) * 65536;
    double *gt = ep.Gt + static_cast<int64_t>(t) * 65536;
    for (int idx = threadIdx.x; idx < 65536; idx += 256) {
        const int cl = idx >> 8, rl = idx & 255;
        const int64_t row = tc.x + rl, col = tc.y + cl;
        double out = 0.0;
        if (row < ep.n && col < ep.n) {
            int32_t c = 0;
            for (int q = 0; q < world; ++q) c += base[q * part_stride + idx];
            double v = static_cast<double>(c) * ep.alpha;
            if (EPI == 2) v = (v - ep.beta * (__ldg(ep.u + row) + __ldg(ep.u + col))) + ep.gamma;
            out = (ep.rcp != 0.0) ? div_const(v, ep.denom, ep.rcp) : v / ep.denom;
        }
        gt[idx] = out;
    }
}
}  // namespace

int32_t grm_flags_gate(grm_cuda_ctx *ctx, const uint32_t *d_flags, int count, uint32_t epoch, cudaStream_t stream) {
    if (count < 1 || count > 1024) {
        grm_set_error("flags_gate: bad count %d", count);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    flags_gate_kernel<<<1, (count + 31) / 32 * 32, 0, stream>>>(d_flags, count, epoch);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

int32_t grm_set_u32(grm_cuda_ctx *ctx, uint32_t *d_word, uint32_t value, cudaStream_t stream) {
    set_u32_kernel<<<1, 1, 0, stream>>>(d_word, value);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

// owned tiles [t0, t0 + cnt) from the `world` received parts ([q][tiles_max][256 x 256] int32) -> packed FP64 tiles
int32_t grm_finalize_parts_range(grm_cuda_ctx *ctx, const int32_t *d_parts, int64_t n, double alpha, double beta, double gamma,
                                 double denom, const double *d_u, double *d_Gt, int32_t rank, int32_t world, int64_t t0,
                                 int64_t cnt, cudaStream_t stream) {
    if (!ctx || !d_parts || !d_Gt || n < 1 || world < 1 || rank < 0 || rank >= world || (beta != 0.0 && !d_u)) {
        grm_set_error("finalize_parts: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    const int2 *d_tiles;
    int count;
    GRM_TRY(get_tile_list_2cta(ctx, n, rank, world, &d_tiles, &count));
    if (t0 < 0) t0 = 0;
    cnt = std::min<int64_t>(cnt, count - t0);
    if (cnt <= 0) return GRM_CUDA_OK;
    EpiParams ep{};
    ep.n = n;
    ep.Gt = d_Gt;
    ep.alpha = alpha;
    ep.beta = beta;
    ep.gamma = gamma;
    set_division(ctx, ep, denom);
    ep.u = d_u;
    const int64_t stride = grm_cuda_packed_tiles_max(n, world) * 65536;
    if (beta == 0.0 && gamma == 0.0)
        finalize_parts_kernel<1><<<static_cast<unsigned>(cnt), 256, 0, stream>>>(d_parts, stride, world, d_tiles, ep, static_cast<int>(t0));
    else
        finalize_parts_kernel<2><<<static_cast<unsigned>(cnt), 256, 0, stream>>>(d_parts, stride, world, d_tiles, ep, static_cast<int>(t0));
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

// blocks `stream` until wave `wave` of the streamed contraction has been stored (8 counts per tile)
int32_t grm_wave_gate(grm_cuda_ctx *ctx, const uint32_t *d_wave_done, int64_t n, int32_t world, int32_t waves, int32_t wave,
                      cudaStream_t stream) {
    PackedPlanCache *pc;
    GRM_TRY(grm_packed_plan(ctx, n, world, waves, &pc));
    const uint32_t target = 8u * static_cast<uint32_t>(pc->tile_cnt[wave]);
    if (target == 0) return GRM_CUDA_OK;
    wave_gate_kernel<<<1, 32, 0, stream>>>(d_wave_done + wave, target);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

// all waves of the K-split contraction in ONE persistent launch, tiles in wave order, per-wave completion counters in
// d_wave_done[waves] (zeroed by the caller); `reserve_sms` SMs are left to the communication kernels
int32_t grm_syrk_i8_packed_stream(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc, int32_t world,
                                  int32_t waves, int32_t *d_packed, uint32_t *d_wave_done, int32_t reserve_sms) {
    if (!ctx || !d_codes || !d_packed || !d_wave_done || n < 1 || p < 1 || ldc < p || (ldc % 16) != 0 || world < 1 ||
        waves < 1 || waves > 8) {
        grm_set_error("syrk_i8_packed_stream: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    CUtensorMap tmap;
    GRM_TRY(make_codes_tmap(d_codes, n, p, ldc, 1, &tmap));
    PackedPlanCache *pc;
    GRM_TRY(grm_packed_plan(ctx, n, world, waves, &pc));
    const int64_t tiles_max = grm_cuda_packed_tiles_max(n, world);
    if (static_cast<int64_t>(pc->total_tiles) != tiles_max * world)
        GRM_CUDA_TRY(cudaMemsetAsync(d_packed, 0, static_cast<size_t>(tiles_max) * world * 65536 * sizeof(int32_t), ctx->stream));
    const int count = pc->total_tiles;
    if (count == 0) return GRM_CUDA_OK;
    EpiParams ep{};
    ep.n = n;
    ep.C = d_packed;
    ep.slots = reinterpret_cast<const int32_t *>(pc->buf.as<int2>() + pc->total_tiles);
    ep.wave_done = d_wave_done;
    ep.nwaves = waves;
    for (int w = 0; w < 8; ++w) ep.wave_end[w] = w < waves ? pc->tile_off[w] + pc->tile_cnt[w] : count;
    const int kbs = static_cast<int>((p + BK - 1) / BK);
    const int pairs = std::max(1, std::min(count, ctx->num_sms / 2 - std::max(0, (reserve_sms + 1) / 2)));
    const int saved_sync = ctx->tune.syrk_sync;
    ctx->tune.syrk_sync = 0;  // a grid that shares the device with the exchange kernels must not wait on co-residency
    const int32_t st = launch_pairs(ctx, reinterpret_cast<const void *>(syrk_i8_2cta_kernel<3>), SMEM2_BYTES, pairs, count, n * ldc,
                                    [&](const cudaLaunchConfig_t *cfg, uint32_t *prog) {
                                        return cudaLaunchKernelEx(cfg, syrk_i8_2cta_kernel<3>, tmap, pc->buf.as<int2>(), count, kbs,
                                                                  kbs, ep, prog, window_blocks(ctx));
                                    });
    ctx->tune.syrk_sync = saved_sync;
    return st;
}

// finalise owned tiles [t0, t0 + cnt) of the reduced buffer ([tiles_max][256 x 256] int32, slot = position in the
// rank's tile list) into the dense matrix d_G, or (d_Gt) into packed FP64 tiles at the same slots
int32_t grm_finalize_tiles_range(grm_cuda_ctx *ctx, const int32_t *d_reduced, int64_t n, double alpha, double beta,
                                 double gamma, double denom, const double *d_u, double *d_G, int64_t ldg, double *d_Gt,
                                 int32_t rank, int32_t world, int64_t t0, int64_t cnt, cudaStream_t stream) {
    if (!ctx || !d_reduced || (!d_G && !d_Gt) || n < 1 || (d_G && ldg < n) || world < 1 || rank < 0 || rank >= world ||
        (beta != 0.0 && !d_u)) {
        grm_set_error("finalize_tiles: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    const int2 *d_tiles;
    int count;
    GRM_TRY(get_tile_list_2cta(ctx, n, rank, world, &d_tiles, &count));
    if (t0 < 0) t0 = 0;
    cnt = std::min<int64_t>(cnt, count - t0);
    if (cnt <= 0) return GRM_CUDA_OK;
    EpiParams ep{};
    ep.n = n;
    ep.G = d_G;
    ep.ldg = ldg;
    ep.Gt = d_Gt;
    ep.alpha = alpha;
    ep.beta = beta;
    ep.gamma = gamma;
    set_division(ctx, ep, denom);
    ep.u = d_u;
    if (beta == 0.0 && gamma == 0.0)
        finalize_tiles_kernel<1><<<static_cast<unsigned>(cnt), 256, 0, stream>>>(d_reduced, d_tiles, ep, static_cast<int>(t0));
    else
        finalize_tiles_kernel<2><<<static_cast<unsigned>(cnt), 256, 0, stream>>>(d_reduced, d_tiles, ep, static_cast<int>(t0));
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_finalize_tiles(grm_cuda_ctx *ctx, const int32_t *d_reduced, int64_t n, double alpha,
                                           double beta, double gamma, double denom, const double *d_u, double *d_G,
                                           int64_t ldg, int32_t rank, int32_t world) try {
    if (!ctx) {
        grm_set_error("finalize_tiles: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    return grm_finalize_tiles_range(ctx, d_reduced, n, alpha, beta, gamma, denom, d_u, d_G, ldg, nullptr, rank, world, 0,
                                    INT32_MAX, ctx->stream);
} GRM_CATCH_STATUS
