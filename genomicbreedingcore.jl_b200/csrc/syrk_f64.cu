// syrk_f64.cu — FP64 upper-triangular SYRK for continuous (pool-seq) allele frequencies, plus the
// scale-and-mirror pass that makes the result exactly symmetric.
//
// Replaces the pair loop of src/grm/calc.jl:124-131 (grmsimple, z = x) and :202-209 (grmploidyaware,
// z_ik = fl(fl(ploidy*fl(x_ik - 0.5)) - q_k), calc.jl:198,205).  The operand Z is produced once per loci chunk by
// centre_f64_kernel (reference rounding sequence, optional mean-fill and QC mask); the SYRK main loop is pure DMMA.
//
// tcgen05 has no FP64 kind, so the contraction runs on the FP64 tensor path that exists on sm_100:
// mma.sync.m16n8k8.f64 (DMMA).  Block tile 128 x 128 entries, 8 consumer warps (2 x 4), warp tile 64 x 32.  Only tiles
// that touch the upper triangle run.  Two operand feeds:
//   syrk_f64_tma_kernel (default when the operand rows are 16-byte aligned): an elected lane keeps a 3-stage ring of
//     TMA boxes (16 entries x 32 loci, SWIZZLE_128B) in flight; no block-wide barrier in the main loop, a diagonal tile
//     loads its panel once.
//   syrk_f64_kernel: register-staged __ldg -> st.shared into a padded layout, double buffered (tuning f64_kernel = 1,
//     and for operands TMA cannot describe).
#include "common.cuh"

namespace {

constexpr int FB = 128;          // block tile edge (entries)
constexpr int FK = 32;           // loci per slab (staged in two halves of 16: one barrier per 32 loci)
constexpr int FH = 16;           // loci per half-slab
constexpr int FLD = FB + 4;      // padded row length: 4a + b (a,b in 0..3) hits 16 distinct 8-byte banks
constexpr size_t F64_SMEM = 2ull * 2 * FK * FLD * sizeof(double);

// D(16x8) += A(16x8, row) * B(8x8, col), FP64.  Fragments (g = lane >> 2, t = lane & 3):
//   a0:(g, t) a1:(g+8, t) a2:(g, t+4) a3:(g+8, t+4)   b0:(k = t, n = g) b1:(k = t+4, n = g)
//   c0,c1:(g, 2t + {0,1})   c2,c3:(g+8, 2t + {0,1})
__device__ __forceinline__ void dmma_m16n8k8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__global__ void __launch_bounds__(256, 1)
syrk_f64_kernel(const double *__restrict__ A, int64_t n, int64_t p, int64_t lda, const int2 *__restrict__ tiles,
                double *__restrict__ G, int64_t ldg, int accumulate) {
    extern __shared__ double sm[];
    double(*sA)[FK][FLD] = reinterpret_cast<double(*)[FK][FLD]>(sm);
    double(*sB)[FK][FLD] = reinterpret_cast<double(*)[FK][FLD]>(sm + 2 * FK * FLD);

    const int2 tc = tiles[blockIdx.x];
    const int64_t row0 = tc.x, col0 = tc.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;

    // staging map: thread -> entry e (0..127) and 8 consecutive loci of the slab
    const int e = tid & 127, kh = (tid >> 7) * 8;  // 2 x 8 loci of a 16-locus half-slab
    const int64_t ra = row0 + e, rb = col0 + e;
    const bool va = ra < n, vb = rb < n;

    double acc[4][4][4];  // [m16 tile][n8 tile][fragment]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[i][j][f] = 0.0;

    double pa[8], pb[8];
    // half-slab h of the slab starting at k0: loci k0 + 16 h + kh + (0..7) for this thread
    auto fetch = [&](int64_t k0, int h) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t k = k0 + h * FH + kh + j;
            const bool vk = k < p;
            const double xa = (vk && va) ? __ldg(A + ra + k * lda) : 0.0;
            const double xb = (vk && vb) ? __ldg(A + rb + k * lda) : 0.0;
            pa[j] = xa;
            pb[j] = xb;
        }
    };
    auto stash = [&](int buf, int h) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sA[buf][h * FH + kh + j][e] = pa[j];
            sB[buf][h * FH + kh + j][e] = pb[j];
        }
    };
    const int fr = lane >> 2, fk = lane & 3;
    auto compute = [&](int buf, int h) {
#pragma unroll
        for (int kk = 0; kk < FH / 8; ++kk) {
            const int kb = h * FH + kk * 8;
            double a[4][4], b[4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const int r = wm * 64 + mi * 16 + fr;
                a[mi][0] = sA[buf][kb + fk][r];
                a[mi][1] = sA[buf][kb + fk][r + 8];
                a[mi][2] = sA[buf][kb + fk + 4][r];
                a[mi][3] = sA[buf][kb + fk + 4][r + 8];
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int c = wn * 32 + ni * 8 + fr;
                b[ni][0] = sB[buf][kb + fk][c];
                b[ni][1] = sB[buf][kb + fk + 4][c];
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma_m16n8k8(acc[mi][ni], a[mi], b[ni]);
        }
    };

    const int64_t nslab = (p + FK - 1) / FK;
    fetch(0, 0);
    stash(0, 0);
    fetch(0, 1);
    stash(0, 1);
    __syncthreads();

    for (int64_t s = 0; s < nslab; ++s) {
        const int buf = static_cast<int>(s & 1);
        const bool more = s + 1 < nslab;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (more) fetch((s + 1) * FK, h);
            compute(buf, h);
            if (more) stash(buf ^ 1, h);
        }
        __syncthreads();
    }

#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
            const int64_t row = row0 + wm * 64 + mi * 16 + fr + 8 * hr;
            if (row >= n) continue;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t col = col0 + wn * 32 + ni * 8 + 2 * fk + h;
                    if (col < n) {
                        double *dst = G + row + col * ldg;
                        const double v = acc[mi][ni][2 * hr + h];
                        *dst = accumulate ? (*dst + v) : v;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-fed variant.  The operand Z is column-major (entry fastest), which is already the [k][entry] order the DMMA
// fragments want, so a 2-D box {16 entries, 32 loci} lands ready to use: 32 rows of 128 bytes.  Under SWIZZLE_128B
// the 16-byte chunk c of row r sits at chunk c ^ (r & 7).  A fragment load touches, per half-warp, 4 entries-pairs x 4
// values of k; with k = t -> row t the four rows would fall on the same banks (2-way conflict).  A sum over loci does not
// care which physical locus plays which k of the MMA, so within each group of 8 loci logical k = t reads row 2t and
// k = t + 4 reads row 2t + 1 (the same mapping for A and B): the XOR then spreads the 16 threads of a half-warp over
// all 8 chunks — every fragment load is conflict-free.  r & 7 does not depend on the 8-locus step, so the swizzled
// column offsets are four per-thread constants.
// ---------------------------------------------------------------------------------------------------------------
// shared-space load on a 32-bit shared address (the 1 KiB alignment cast hides the address space from the compiler, which
// would otherwise emit generic loads); volatile: stays after the mbarrier wait that publishes the stage
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

constexpr int TS = 3;                      // ring stages
constexpr int TBOX = 16;                   // entries per box = 128 bytes = the swizzle span
constexpr int TBOX_DBL = TBOX * FK;        // 512 doubles = 4 KiB per box
constexpr int TPANEL_DBL = (FB / TBOX) * TBOX_DBL;  // 128 entries x 32 loci = 32 KiB
constexpr size_t F64T_SMEM = static_cast<size_t>(TS) * 2 * TPANEL_DBL * sizeof(double) + 1024 /*alignment*/ + 128 /*barriers*/;

// 8 warps (the register file of an SM is split over 4 sub-cores of 16 K registers: a ninth warp would put three warps of
// ~200 registers on one of them, which does not fit — so there is no dedicated producer warp).  Lane 0 of warp 0 issues
// the TMA loads: slabs 0 .. TS-2 up front, then slab s + TS - 1 right after warp 0 has finished slab s; the stage it
// refills was consumed during slab s - 1, so the wait for "stage empty" is normally over before it is asked.
__global__ void __launch_bounds__(256, 1)
syrk_f64_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n, int64_t p, const int2 *__restrict__ tiles,
                    double *__restrict__ G, int64_t ldg, int accumulate) {
    extern __shared__ uint8_t f64t_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(f64t_raw) + 1023) & ~uintptr_t(1023));
    double *ring = reinterpret_cast<double *>(base);  // [TS][A panel | B panel]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + static_cast<size_t>(TS) * 2 * TPANEL_DBL * sizeof(double));
    uint64_t *empty_bar = full_bar + TS;

    const int2 tc = tiles[blockIdx.x];
    const int64_t row0 = tc.x, col0 = tc.y;
    const bool diag = row0 == col0;  // both operands are the same panel: load it once
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < TS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 8);  // one arrival per warp
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int nslab = static_cast<int>((p + FK - 1) / FK);

    // producer state (used by thread 0 only)
    int pstage = 0;
    uint32_t pphase = 0;
    const uint32_t stage_bytes = (diag ? 1u : 2u) * TPANEL_DBL * sizeof(double);
    auto produce = [&](int t) {
        mbar_wait(&empty_bar[pstage], pphase ^ 1);  // passes at once for the first use of a stage
        mbar_arrive_expect_tx(&full_bar[pstage], stage_bytes);
        double *dA = ring + static_cast<size_t>(pstage) * 2 * TPANEL_DBL;
        // entries beyond n and loci beyond p are zero-filled by TMA: ragged edges add exact zeros
#pragma unroll
        for (int b = 0; b < FB / TBOX; ++b)
            tma_load_2d(dA + b * TBOX_DBL, &tmap, &full_bar[pstage], static_cast<int32_t>(row0) + TBOX * b, t * FK);
        if (!diag) {
#pragma unroll
            for (int b = 0; b < FB / TBOX; ++b)
                tma_load_2d(dA + TPANEL_DBL + b * TBOX_DBL, &tmap, &full_bar[pstage], static_cast<int32_t>(col0) + TBOX * b, t * FK);
        }
        if (++pstage == TS) {
            pstage = 0;
            pphase ^= 1;
        }
    };
    if (tid == 0)
        for (int t = 0; t < TS - 1 && t < nslab; ++t) produce(t);

    // ===================== consumers: 8 warps, warp tile 64 x 32 =====================
    const int wm = warp >> 2, wn = warp & 3;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[4][4][4];  // [m16 tile][n8 tile][fragment]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[i][j][f] = 0.0;
    // element (entry e of the box, row r) lives at r * 16 + (((e >> 1) ^ (r & 7)) << 1) + (e & 1); rows used by this
    // thread are 8 ks + 2 fk (+ 1), so r & 7 = 2 fk (+ 1) for every ks
    const int re = 2 * fk, ro = 2 * fk + 1;
    const int c_lo_e = (((fr >> 1) ^ re) << 1) + (fr & 1), c_hi_e = ((((fr + 8) >> 1) ^ re) << 1) + (fr & 1);
    const int c_lo_o = (((fr >> 1) ^ ro) << 1) + (fr & 1), c_hi_o = ((((fr + 8) >> 1) ^ ro) << 1) + (fr & 1);

    const uint32_t ring_s = smem_u32(ring);
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < nslab; ++s) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t pA = ring_s + static_cast<uint32_t>(stage) * 2 * TPANEL_DBL * 8;
        const uint32_t pB = diag ? pA : pA + TPANEL_DBL * 8;
#pragma unroll
        for (int ks = 0; ks < FK / 8; ++ks) {
            const int row_e = (ks * 8 + re) * TBOX, row_o = (ks * 8 + ro) * TBOX;
            double a[4][4], b[4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const uint32_t box = pA + (wm * 4 + mi) * TBOX_DBL * 8;
                a[mi][0] = lds_f64(box + (row_e + c_lo_e) * 8);
                a[mi][1] = lds_f64(box + (row_e + c_hi_e) * 8);
                a[mi][2] = lds_f64(box + (row_o + c_lo_o) * 8);
                a[mi][3] = lds_f64(box + (row_o + c_hi_o) * 8);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const uint32_t box = pB + (wn * 2 + (ni >> 1)) * TBOX_DBL * 8;
                b[ni][0] = lds_f64(box + (row_e + ((ni & 1) ? c_hi_e : c_lo_e)) * 8);
                b[ni][1] = lds_f64(box + (row_o + ((ni & 1) ? c_hi_o : c_lo_o)) * 8);
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma_m16n8k8(acc[mi][ni], a[mi], b[ni]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);  // this warp has read the stage
        if (tid == 0 && s + TS - 1 < nslab) produce(s + TS - 1);
        if (++stage == TS) {
            stage = 0;
            phase ^= 1;
        }
    }

#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
            const int64_t row = row0 + wm * 64 + mi * 16 + fr + 8 * hr;
            if (row >= n) continue;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t col = col0 + wn * 32 + ni * 8 + 2 * fk + h;
                    if (col < n) {
                        double *dst = G + row + col * ldg;
                        const double v = acc[mi][ni][2 * hr + h];
                        *dst = accumulate ? (*dst + v) : v;
                    }
                }
            }
        }
    }
}

// G[i,j] = G[i,j] / denom for i <= j, then G[j,i] = G[i,j]: 32 x 32 tiles through shared memory so both the read
// of the upper tile and the write of its mirror are coalesced.
__global__ void __launch_bounds__(256) scale_mirror_kernel(double *__restrict__ G, int64_t n, int64_t ldg, double denom,
                                                           int divide) {
    __shared__ double t[32][33];
    const int64_t bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t i0 = bi * 32, j0 = bj * 32;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int jj = ty + 8 * r;
        const int64_t i = i0 + tx, j = j0 + jj;
        double v = 0.0;
        if (i < n && j < n) {
            v = G[i + j * ldg];
            if (divide) {
                v = v / denom;
                if (i <= j) G[i + j * ldg] = v;
            }
        }
        t[jj][tx] = v;  // t[j_local][i_local]
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int ii = ty + 8 * r;
        // mirror element: row (j0 + tx), column (i0 + ii)  <-  upper element (i0 + ii, j0 + tx)
        const int64_t i = i0 + ii, j = j0 + tx;
        if (i < n && j < n && i < j) G[j + i * ldg] = t[tx][ii];
    }
}

// Z = centred / mean-filled / masked copy of a loci chunk, written once so that the SYRK main loop is pure DMMA:
// FP64 ALU work interleaved with DMMA in the operand-staging path costs far more than its instruction count suggests
// (measured: fused centring 14.5 TFLOP/s, fused centring + fill 11.5, plain operands 25.4), while this pass is HBM-bound
// and its traffic (16 bytes per cell) is negligible next to the n/16 flop per byte of the contraction.
//   z = fl(fl(ploidy * fl(x - 0.5)) - q_k)   (calc.jl:198,205; explicit round-to-nearest ops, never contracted)
__global__ void __launch_bounds__(256)
centre_f64_kernel(const double *__restrict__ A, int64_t n, int64_t pc, int64_t lda, int centred, double ploidy,
                  const double *__restrict__ q, int fill, const uint8_t *__restrict__ keep, double *__restrict__ Z,
                  int64_t ldz) {
    const int64_t k = blockIdx.y;
    const bool kp = keep == nullptr || keep[k] != 0;
    const double qk = (centred || fill) ? q[k] : 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double x = A[i + k * lda];
        if (fill && x != x) x = qk;
        if (centred) x = __dsub_rn(__dmul_rn(ploidy, __dsub_rn(x, 0.5)), qk);
        Z[i + k * ldz] = kp ? x : 0.0;
    }
}

int32_t get_tile_list_f64(grm_cuda_ctx *ctx, int64_t n, int32_t rank, int32_t world, const int2 **d_tiles, int *count) {
    TileListCache &tc = ctx->tiles_f64;
    if (tc.n != n || tc.rank != rank || tc.world != world || tc.kind != ctx->tune.tile_band) {
        std::vector<int2> order;
        grm_tile_order(n, ctx->tune.tile_band, order);
        int64_t b, e;
        grm_tile_range(static_cast<int64_t>(order.size()), rank, world, &b, &e);
        std::vector<int2> blk;
        for (int64_t t = b; t < e; ++t) {
            const int r0 = order[t].x * GRM_TILE, c0 = order[t].y * GRM_TILE;
            for (int dr = 0; dr < GRM_TILE; dr += FB)
                for (int dc = 0; dc < GRM_TILE; dc += FB) {
                    const int r = r0 + dr, c = c0 + dc;
                    if (r >= n || c >= n || r > c) continue;  // outside, or strictly below the diagonal
                    blk.push_back(make_int2(r, c));
                }
        }
        GRM_TRY(tc.buf.reserve(std::max<size_t>(blk.size(), 1) * sizeof(int2)));
        GRM_CUDA_TRY(cudaMemcpyAsync(tc.buf.ptr, blk.data(), blk.size() * sizeof(int2), cudaMemcpyHostToDevice,
                                     ctx->stream));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        tc.n = n;
        tc.rank = rank;
        tc.world = world;
        tc.kind = ctx->tune.tile_band;
        tc.count = static_cast<int32_t>(blk.size());
    }
    *d_tiles = tc.buf.as<int2>();
    *count = tc.count;
    return GRM_CUDA_OK;
}

}  // namespace

extern "C" int32_t grm_cuda_syrk_f64(grm_cuda_ctx *ctx, const double *dZ, int64_t n, int64_t p, int64_t ldz,
                                     int32_t accumulate, double *d_G, int64_t ldg, int32_t rank, int32_t world) try {
    if (!ctx || !dZ || !d_G || n < 1 || p < 1 || ldz < n || ldg < n || world < 1 || rank < 0 || rank >= world) {
        grm_set_error("syrk_f64: bad argument (n=%lld p=%lld ldz=%lld ldg=%lld)", (long long)n, (long long)p,
                      (long long)ldz, (long long)ldg);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    const int2 *d_tiles;
    int count;
    GRM_TRY(get_tile_list_f64(ctx, n, rank, world, &d_tiles, &count));
    if (count == 0) return GRM_CUDA_OK;
    if (ctx->tune.f64_kernel != 1 && (ldz % 2) == 0 && (reinterpret_cast<uintptr_t>(dZ) % 16) == 0 && n < (1ll << 31) - 256 &&
        p < (1ll << 31) - 64) {
        // TMA-fed variant: needs 16-byte aligned operand rows (global strides are multiples of 16 bytes)
        GrmEncodeTiledFn encode = nullptr;
        GRM_TRY(grm_tmap_encoder(&encode));
        CUtensorMap tmap;
        cuuint64_t gdim[2] = {static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(p)};
        cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ldz) * sizeof(double)};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(TBOX), static_cast<cuuint32_t>(FK)};
        cuuint32_t estr[2] = {1u, 1u};
        const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(dZ), gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            grm_set_error("cuTensorMapEncodeTiled (FP64 SYRK operand) failed with CUresult %d", static_cast<int>(r));
            return GRM_CUDA_ERR_CUDA;
        }
        GRM_TRY(grm_func_smem(ctx, reinterpret_cast<const void *>(syrk_f64_tma_kernel), F64T_SMEM));
        syrk_f64_tma_kernel<<<count, 256, F64T_SMEM, ctx->stream>>>(tmap, n, p, d_tiles, d_G, ldg, accumulate);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        return GRM_CUDA_OK;
    }
    GRM_TRY(grm_func_smem(ctx, reinterpret_cast<const void *>(syrk_f64_kernel), F64_SMEM));
    syrk_f64_kernel<<<count, 256, F64_SMEM, ctx->stream>>>(dZ, n, p, ldz, d_tiles, d_G, ldg, accumulate);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_scale_mirror(grm_cuda_ctx *ctx, double *d_G, int64_t n, int64_t ldg, double denom) try {
    if (!ctx || !d_G || n < 1 || ldg < n) {
        grm_set_error("scale_mirror: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    const unsigned nb = static_cast<unsigned>((n + 31) / 32);
    scale_mirror_kernel<<<dim3(nb, nb), 256, 0, ctx->stream>>>(d_G, n, ldg, denom, denom != 1.0 ? 1 : 0);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_mirror(grm_cuda_ctx *ctx, double *d_G, int64_t n, int64_t ldg) try {
    return grm_cuda_scale_mirror(ctx, d_G, n, ldg, 1.0);
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_centre_f64(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t pc, int64_t lda,
                                       int32_t centred, double ploidy, const double *d_q, int32_t fill_missing,
                                       const uint8_t *d_keep, double *d_Z, int64_t ldz) try {
    if (!ctx || !dA || !d_Z || n < 1 || pc < 1 || lda < n || ldz < n || ((centred || fill_missing) && !d_q) ||
        pc > 2147483647ll) {
        grm_set_error("centre_f64: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    // grid.y is limited to 65535: walk the loci in slices
    for (int64_t k0 = 0; k0 < pc; k0 += 65535) {
        const int64_t len = std::min<int64_t>(65535, pc - k0);
        const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 64));
        centre_f64_kernel<<<dim3(gx, static_cast<unsigned>(len)), 256, 0, ctx->stream>>>(
            dA + k0 * lda, n, len, lda, centred, ploidy, d_q ? d_q + k0 : nullptr, fill_missing,
            d_keep ? d_keep + k0 : nullptr, d_Z + k0 * ldz, ldz);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
    }
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS
