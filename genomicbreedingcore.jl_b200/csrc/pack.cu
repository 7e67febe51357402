// pack.cu — the HBM-bound pass over genomes.allele_frequencies (n x p FP64, column-major) and the small
// per-locus / per-entry vector kernels around it.
//
//   detect_kernel        dosage denominator detection ("detected on load")
//   pack_dosage_kernel   per-locus integer column sums + validity flags + FP64 -> int8 quantise + transpose to the
//                        K-major layout the tcgen05 SYRK consumes, one coalesced pass (replaces mean(A, dims=1),
//                        calc.jl:197, and the implicit "any missing => throw", calc.jl:127)
//   locus_stats_*        q_k, w_k = ploidy/2 + q_k, sum q(1-q) (calc.jl:197,201), fixed-order reductions
//   rowdot_i8_kernel     u_i = sum_k c_ik w_k for the rank-1 centring correction
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr uint32_t FLAG_NAN = 1u, FLAG_INF = 2u, FLAG_NOT_DOSAGE = 4u, FLAG_NAN_PAYLOAD = 16u;

// -------------------------------------------------------------------------------------------------------------
// detection: OR of round(64*x) over the sampled cells; any cell that is not k/64 in [0,1] raises a flag.
// -------------------------------------------------------------------------------------------------------------
__global__ void detect_kernel(const double *__restrict__ A, int64_t n, int64_t p, int64_t lda, int64_t cells,
                              uint32_t *__restrict__ out, int ignore_nan) {
    uint32_t bits = 0, flags = 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < cells; idx += stride) {
        const int64_t k = idx / n, i = idx - k * n;
        if (k >= p) break;
        const double x = A[i + k * lda];
        const double t = x * 64.0;
        const int c = __double2int_rn(t);
        if (static_cast<double>(c) == t && c >= 0 && c <= 64) {
            bits |= static_cast<uint32_t>(c);
        } else if (!(ignore_nan && x != x)) {
            flags |= (x != x) ? FLAG_NAN : (isinf(x) ? FLAG_INF : FLAG_NOT_DOSAGE);
        }
    }
    bits = __reduce_or_sync(0xffffffffu, bits);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if ((threadIdx.x & 31) == 0) {
        if (bits) atomicOr(&out[0], bits);
        if (flags) atomicOr(&out[1], flags);
    }
}

// -------------------------------------------------------------------------------------------------------------
// pack: block = 8 warps, tile = 128 loci x (64*EITERS) entries.
//   read : warp w owns loci [16w, 16w+16) of the tile; lane l reads entries l and l+32 of each 64-entry slab
//          (two fully coalesced 256-byte warp loads per locus), 4 loci at a time -> one packed 32-bit word
//   smem : word tile [64 entries][32 words + 1 pad]  (conflict-free both ways)
//   write: 8 threads x 16 bytes = one 128-byte line per entry row of the K-major code matrix
// -------------------------------------------------------------------------------------------------------------
// streaming 8-byte load: read-only path, do not allocate in L1 (every byte of A is touched exactly once)
__device__ __forceinline__ double ld_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

constexpr int PK_LOCI = 128;
constexpr int PK_SLAB = 64;

template <int EITERS, bool MASK>
__global__ void __launch_bounds__(256, 2) pack_dosage_kernel(const double *__restrict__ A, int64_t n, int64_t pc,
                                                          int64_t lda, int m, int8_t *__restrict__ codes, int64_t ldc,
                                                          int32_t *__restrict__ colsum, uint32_t *__restrict__ flags,
                                                          int32_t *__restrict__ miss) {
    // MASK (QC mode): a NaN cell is not an error here — it is packed as 0 and counted per locus in miss[]; the keep
    // decision (locus_keep_kernel) then flags the missing cells of KEPT loci only
    __shared__ uint32_t tile[PK_SLAB][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t k0 = static_cast<int64_t>(blockIdx.x) * PK_LOCI;
    const int64_t e_base = static_cast<int64_t>(blockIdx.y) * (PK_SLAB * EITERS);
    const double mf = static_cast<double>(m);

    int32_t acc[16], mis[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = mis[i] = 0;
    uint32_t bad = 0;

#pragma unroll 1
    for (int it = 0; it < EITERS; ++it) {
        const int64_t e0 = e_base + static_cast<int64_t>(it) * PK_SLAB;
        if (e0 >= n) break;
        const int64_t ea = e0 + lane, eb = e0 + lane + 32;
        const bool va = ea < n, vb = eb < n;
        // all 32 loads of this slab first (256 bytes in flight per thread), then the arithmetic
        double xa[16], xb[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int64_t k = k0 + warp * 16 + j;
            const bool vk = k < pc;
            xa[j] = (vk && va) ? ld_stream(A + ea + k * lda) : 0.0;
            xb[j] = (vk && vb) ? ld_stream(A + eb + k * lda) : 0.0;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t wa = 0, wb = 0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const double va_ = xa[g * 4 + kk], vb_ = xb[g * 4 + kk];
                const double ta = va_ * mf, tb = vb_ * mf;
                int ca = __double2int_rn(ta), cb = __double2int_rn(tb);
                if (!(static_cast<double>(ca) == ta && ca >= 0 && ca <= m)) {
                    if (MASK && va_ != va_) ++mis[g * 4 + kk];
                    else bad |= (va_ != va_) ? FLAG_NAN : (isinf(va_) ? FLAG_INF : FLAG_NOT_DOSAGE);
                    ca = 0;
                }
                if (!(static_cast<double>(cb) == tb && cb >= 0 && cb <= m)) {
                    if (MASK && vb_ != vb_) ++mis[g * 4 + kk];
                    else bad |= (vb_ != vb_) ? FLAG_NAN : (isinf(vb_) ? FLAG_INF : FLAG_NOT_DOSAGE);
                    cb = 0;
                }
                wa |= static_cast<uint32_t>(ca) << (8 * kk);
                wb |= static_cast<uint32_t>(cb) << (8 * kk);
                acc[g * 4 + kk] += ca + cb;
            }
            tile[lane][warp * 4 + g] = wa;
            tile[lane + 32][warp * 4 + g] = wb;
        }
        __syncthreads();
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int r = pass * 32 + (threadIdx.x >> 3);
            const int c = (threadIdx.x & 7) * 4;
            const int64_t e = e0 + r;
            if (e < n) {
                uint4 v = make_uint4(tile[r][c], tile[r][c + 1], tile[r][c + 2], tile[r][c + 3]);
                *reinterpret_cast<uint4 *>(codes + e * ldc + k0 + c * 4) = v;
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int32_t s = __reduce_add_sync(0xffffffffu, acc[i]);
        const int64_t k = k0 + warp * 16 + i;
        if (lane == 0 && k < pc && s != 0) atomicAdd(&colsum[k], s);
        if (MASK) {
            const int32_t ms = __reduce_add_sync(0xffffffffu, mis[i]);
            if (lane == 0 && k < pc && ms != 0) atomicAdd(&miss[k], ms);
        }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (lane == 0 && bad) atomicOr(flags, bad);
}

// -------------------------------------------------------------------------------------------------------------
// pack, TMA-staged variant (used whenever the rows of A are 16-byte aligned: lda even).
// Persistent CTAs; a producer warp keeps a 3-stage ring of 32 KB boxes (32 entries x 128 loci of FP64) in flight
// with cp.async.bulk.tensor, so HBM requests are issued continuously instead of in per-block load phases; eight
// consumer warps quantise from shared memory (warp = 16 loci, lane = entry: conflict-free 256-byte rows), transpose
// through a double-buffered 32 x 33 word tile and write one 128-byte line per entry row.
// -------------------------------------------------------------------------------------------------------------
constexpr int PT_LOCI = 128, PT_ENT = 32, PT_STAGES = 3;
constexpr uint32_t PT_STAGE_BYTES = PT_LOCI * PT_ENT * sizeof(double);  // 32 KiB
constexpr size_t PT_SMEM = PT_STAGES * PT_STAGE_BYTES + 2 * PT_ENT * 33 * sizeof(uint32_t) + 64 + 128;

template <bool MASK>
__global__ void __launch_bounds__(288, 2)
pack_dosage_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n, int64_t pc, int m, int8_t *__restrict__ codes,
                       int64_t ldc, int32_t *__restrict__ colsum, uint32_t *__restrict__ flags,
                       int32_t *__restrict__ miss, int64_t ent_blocks, int64_t units) {
    extern __shared__ uint8_t pt_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(pt_raw) + 127) & ~uintptr_t(127));
    double *stage = reinterpret_cast<double *>(base);
    uint32_t(*tile)[PT_ENT][33] = reinterpret_cast<uint32_t(*)[PT_ENT][33]>(base + PT_STAGES * PT_STAGE_BYTES);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(base + PT_STAGES * PT_STAGE_BYTES + 2 * PT_ENT * 33 * sizeof(uint32_t));
    uint64_t *empty_bar = full_bar + PT_STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // contiguous run of (locus block, entry block) units, entry block fastest: column sums stay in registers
    // until the locus block changes
    const int64_t u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < PT_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 8);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 8) {
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (int64_t u = u0; u < u1; ++u) {
                const int64_t kb = u / ent_blocks, eb = u - kb * ent_blocks;
                mbar_wait(&empty_bar[s], phase ^ 1);
                mbar_arrive_expect_tx(&full_bar[s], PT_STAGE_BYTES);
                tma_load_2d(stage + static_cast<size_t>(s) * (PT_STAGE_BYTES / sizeof(double)), &tmap, &full_bar[s],
                            static_cast<int32_t>(eb * PT_ENT), static_cast<int32_t>(kb * PT_LOCI));
                if (++s == PT_STAGES) {
                    s = 0;
                    phase ^= 1;
                }
            }
        }
        return;
    }

    const double mf = static_cast<double>(m);
    int32_t acc[16], mis[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = mis[i] = 0;
    uint32_t bad = 0;
    int64_t cur_kb = -1;
    int s = 0;
    uint32_t phase = 0;
    int tb = 0;
    auto flush = [&](int64_t kb) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int32_t sum = __reduce_add_sync(0xffffffffu, acc[i]);
            const int64_t k = kb * PT_LOCI + warp * 16 + i;
            if (lane == 0 && k < pc && sum != 0) atomicAdd(&colsum[k], sum);
            acc[i] = 0;
            if (MASK) {
                const int32_t ms = __reduce_add_sync(0xffffffffu, mis[i]);
                if (lane == 0 && k < pc && ms != 0) atomicAdd(&miss[k], ms);
                mis[i] = 0;
            }
        }
    };
    for (int64_t u = u0; u < u1; ++u) {
        const int64_t kb = u / ent_blocks, eb = u - kb * ent_blocks;
        if (kb != cur_kb) {
            if (cur_kb >= 0) flush(cur_kb);
            cur_kb = kb;
        }
        const int64_t k0 = kb * PT_LOCI, e0 = eb * PT_ENT;
        mbar_wait(&full_bar[s], phase);
        const double *st = stage + static_cast<size_t>(s) * (PT_STAGE_BYTES / sizeof(double));
        double x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = st[(warp * 16 + j) * PT_ENT + lane];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);  // this warp's 16 rows of the stage are in registers
        if (++s == PT_STAGES) {
            s = 0;
            phase ^= 1;
        }
        const bool ve = e0 + lane < n;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t w = 0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int64_t k = k0 + warp * 16 + g * 4 + kk;
                const bool vk = k < pc && ve;  // out-of-range cells were zero-filled by TMA
                const double v = x[g * 4 + kk];
                const double t = v * mf;
                int c = __double2int_rn(t);
                if (!vk) {
                    c = 0;
                } else if (!(static_cast<double>(c) == t && c >= 0 && c <= m)) {
                    if (MASK && v != v) ++mis[g * 4 + kk];  // QC mode: counted per locus, judged by locus_keep_kernel
                    else bad |= (v != v) ? FLAG_NAN : (isinf(v) ? FLAG_INF : FLAG_NOT_DOSAGE);
                    c = 0;
                }
                w |= static_cast<uint32_t>(c) << (8 * kk);
                acc[g * 4 + kk] += c;
            }
            tile[tb][lane][warp * 4 + g] = w;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 consumer warps only
        {
            const int r = threadIdx.x >> 3, c = (threadIdx.x & 7) * 4;
            const int64_t e = e0 + r;
            if (e < n) {
                const uint4 v = make_uint4(tile[tb][r][c], tile[tb][r][c + 1], tile[tb][r][c + 2], tile[tb][r][c + 3]);
                *reinterpret_cast<uint4 *>(codes + e * ldc + k0 + c * 4) = v;
            }
        }
        tb ^= 1;  // double-buffered tile: the next unit's writes cannot overtake this unit's reads (see pack.cu notes)
    }
    if (cur_kb >= 0) flush(cur_kb);
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (lane == 0 && bad) atomicOr(flags, bad);
}

// -------------------------------------------------------------------------------------------------------------
// fixed-order block reduction of up to 3 doubles per thread (bit-reproducible for a fixed launch shape)
// -------------------------------------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 256;
constexpr int RED_THREADS = 256;

template <int NV> __device__ __forceinline__ void block_reduce(double (&v)[NV], double *out /* NV doubles */) {
    __shared__ double sh[NV][RED_THREADS / 32];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[i][warp] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < RED_THREADS / 32; ++w) s += sh[i][w];
            out[i] = s;
        }
    }
}

// q, w and per-block partials {sum q(1-q), sum w^2, sum q}; block b owns a contiguous range of loci.
__global__ void __launch_bounds__(RED_THREADS)
locus_stats_i8_kernel(const int32_t *__restrict__ colsum, int64_t p, double n, double m, double half_ploidy,
                      double *__restrict__ q, double *__restrict__ w, double *__restrict__ partials) {
    const int64_t per = (p + gridDim.x - 1) / gridDim.x;
    const int64_t b = static_cast<int64_t>(blockIdx.x) * per, e = min(p, b + per);
    double v[3] = {0.0, 0.0, 0.0};
    for (int64_t k = b + threadIdx.x; k < e; k += blockDim.x) {
        const double qk = (static_cast<double>(colsum[k]) / m) / n;  // fl(S_k / n), S_k = colsum/m exact (calc.jl:197)
        const double wk = half_ploidy + qk;
        q[k] = qk;
        w[k] = wk;
        v[0] += qk * (1.0 - qk);                                     // calc.jl:201
        v[1] += wk * wk;
        v[2] += qk;
    }
    block_reduce<3>(v, partials + 3 * blockIdx.x);
}

// one warp per locus column, lanes stride over entries with a fixed shuffle tree: bit-reproducible column means
__global__ void __launch_bounds__(256)
colmean_f64_kernel(const double *__restrict__ A, int64_t n, int64_t p, int64_t lda, double *__restrict__ q,
                   uint32_t *__restrict__ flags, int skip_missing, int32_t *__restrict__ cnt_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t k = static_cast<int64_t>(blockIdx.x) * 8 + warp;
    if (k >= p) return;
    const double *col = A + k * lda;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    uint32_t bad = 0;
    int cnt = 0;  // valid cells seen by this lane (skip_missing only)
    int64_t i = lane;
    if (skip_missing) {
        // mean(skipmissing(column)) (src/genomes/filter.jl:632-635): a NaN cell contributes nothing
        for (; i < n; i += 32) {
            const double a = __ldg(col + i);
            if (a == a) {
                s0 += a;
                ++cnt;
            }
        }
    } else {
        for (; i + 96 < n; i += 128) {
            const double a = __ldg(col + i), b = __ldg(col + i + 32), c = __ldg(col + i + 64), d = __ldg(col + i + 96);
            s0 += a;
            s1 += b;
            s2 += c;
            s3 += d;
        }
        for (; i < n; i += 32) s0 += __ldg(col + i);
    }
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) {
        if (s != s) bad |= FLAG_NAN;  // NaN anywhere (or Inf-Inf) poisons the column sum
        else if (isinf(s)) bad |= FLAG_INF;
        if (skip_missing && cnt == 0 && cnt_out == nullptr) bad |= 8u;  // no valid cell: the fill value is undefined
        if (cnt_out) {
            cnt_out[k] = cnt;
            bad = 0;  // QC mode: a non-finite mean fails the MAF test and drops the locus; the mask kernel reports
                      // missing cells of KEPT loci
        }
        q[k] = s / static_cast<double>(skip_missing ? cnt : n);
        if (bad) atomicOr(flags, bad);
    }
}

// QC locus mask, the reference's filter semantics (src/genomes/filter.jl):
//   sparsity_k = mean(ismissing(column))            keep iff sparsity_k <= max_locus_sparsity     (:202-213, :361-368)
//   q_k = mean(skipmissing(column))                 keep iff maf <= q_k <= 1 - maf                (:632-635)
// A kept locus that still holds missing cells raises FLAG_NAN when the missing policy is "error".
__global__ void locus_mask_kernel(const double *__restrict__ q, const int32_t *__restrict__ cnt, int64_t n, int64_t p,
                                  double maf, double max_sparsity, int error_on_missing, uint8_t *__restrict__ keep,
                                  unsigned long long *__restrict__ kept, uint32_t *__restrict__ flags) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int kp = 0;
    uint32_t bad = 0;
    if (k < p) {
        const double sparsity = static_cast<double>(n - cnt[k]) / static_cast<double>(n);
        const double qk = q[k];
        kp = (sparsity <= max_sparsity) && (qk >= maf) && (qk <= (1.0 - maf));
        keep[k] = static_cast<uint8_t>(kp);
        if (kp && error_on_missing && cnt[k] < n) bad = FLAG_NAN;
    }
    const unsigned m = __ballot_sync(0xffffffffu, kp);
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (m) atomicAdd(kept, static_cast<unsigned long long>(__popc(m)));
        if (bad) atomicOr(flags, bad);
    }
}

__global__ void __launch_bounds__(RED_THREADS)
locus_stats_f64_kernel(const double *__restrict__ q, int64_t p, double *__restrict__ partials,
                       const uint8_t *__restrict__ keep) {
    const int64_t per = (p + gridDim.x - 1) / gridDim.x;
    const int64_t b = static_cast<int64_t>(blockIdx.x) * per, e = min(p, b + per);
    double v[3] = {0.0, 0.0, 0.0};
    for (int64_t k = b + threadIdx.x; k < e; k += blockDim.x) {
        if (keep && !keep[k]) continue;
        const double qk = q[k];
        v[0] += qk * (1.0 - qk);
        v[2] += qk;
    }
    block_reduce<3>(v, partials + 3 * blockIdx.x);
}

// stats[j] (+)= sum_b partials[b][j] in block order
__global__ void final_stats_kernel(const double *__restrict__ partials, int blocks, double *__restrict__ stats,
                                   int accumulate) {
    if (threadIdx.x < 3) {
        double s = accumulate ? stats[threadIdx.x] : 0.0;
        for (int b = 0; b < blocks; ++b) s += partials[3 * b + threadIdx.x];
        stats[threadIdx.x] = s;
    }
}

// -------------------------------------------------------------------------------------------------------------
// u_i = sum_k c_ik w_k : block = 256 threads x 8 entry rows; thread t owns 16-locus groups t, t+256, ...
// (w is read once per 8 rows; codes once).  Per-thread sequential FMA chain + fixed tree => reproducible.
// -------------------------------------------------------------------------------------------------------------
constexpr int RD_ROWS = 8;

__global__ void __launch_bounds__(256)
rowdot_i8_kernel(const int8_t *__restrict__ codes, int64_t n, int64_t p, int64_t ldc, const double *__restrict__ w,
                 double *__restrict__ u) {
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * RD_ROWS;
    double acc[RD_ROWS];
#pragma unroll
    for (int r = 0; r < RD_ROWS; ++r) acc[r] = 0.0;
    const int64_t groups = (p + 15) / 16;
    for (int64_t g = threadIdx.x; g < groups; g += blockDim.x) {
        const int64_t k = g * 16;
        double wv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) wv[j] = (k + j < p) ? __ldg(w + k + j) : 0.0;
#pragma unroll
        for (int r = 0; r < RD_ROWS; ++r) {
            const int64_t row = r0 + r;
            if (row < n) {
                // ldc is a multiple of 128 and the padding columns are zero, so the 16-byte read is in bounds
                const uint4 cv = __ldg(reinterpret_cast<const uint4 *>(codes + row * ldc + k));
                const uint32_t words[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = static_cast<int8_t>((words[j >> 2] >> (8 * (j & 3))) & 0xffu);
                    acc[r] = fma(static_cast<double>(c), wv[j], acc[r]);
                }
            }
        }
    }
    __shared__ double sh[RD_ROWS][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < RD_ROWS; ++r) {
        double v = acc[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[r][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < RD_ROWS && r0 + threadIdx.x < n) {
        double s = 0.0;
        for (int wi = 0; wi < 8; ++wi) s += sh[threadIdx.x][wi];
        u[r0 + threadIdx.x] = s;
    }
}


// -------------------------------------------------------------------------------------------------------------
// Exact integer statistics of the dosage codes (grmploidyaware on the int8 path).
//   S_k = sum_i c_ik (column sums, from the pack pass)      r_i = sum_k c_ik      T_i = sum_k c_ik S_k
//   sums = { sum_k S_k, sum_k S_k^2 }
// With q_k = S_k/(m n) and w_k = ploidy/2 + q_k (calc.jl:197-198,205) every term of the rank-1 expansion
//  is a ratio of these integers:  u_i = sum_k c_ik w_k = h r_i + T_i/(m n),
// W2 = sum_k w_k^2 = p h^2 + 2 h sum(S)/(m n) + sum(S^2)/(m n)^2,  d = ploidy (sum(S)/(m n) - sum(S^2)/(m n)^2).
// Integers are order-independent, so the result is bit-identical for any chunking or GPU count.
// -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_planes_kernel(const int32_t *__restrict__ colsum, int64_t p, int64_t ldc, uint8_t *__restrict__ planes,
                     unsigned long long *__restrict__ sums) {
    unsigned long long s1 = 0, s2 = 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < ldc; k += stride) {
        const uint32_t S = k < p ? static_cast<uint32_t>(colsum[k]) : 0u;
        planes[k] = static_cast<uint8_t>(S & 255u);            // S = b0 + 256 b1 + 65536 b2  (S < 2^24)
        planes[ldc + k] = static_cast<uint8_t>((S >> 8) & 255u);
        planes[2 * ldc + k] = static_cast<uint8_t>((S >> 16) & 255u);
        s1 += S;
        s2 += static_cast<unsigned long long>(S) * S;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[0], s1);
        atomicAdd(&sums[1], s2);
    }
}

// block = 256 threads x 8 entry rows; thread t owns 16-locus groups t, t+256, ...; 16 dp4a per row and group
__global__ void __launch_bounds__(256, 2)
rowstats_i8_kernel(const int8_t *__restrict__ codes, int64_t n, int64_t p, int64_t ldc,
                   const uint8_t *__restrict__ planes, long long *__restrict__ r_out, long long *__restrict__ T_out,
                   int accumulate) {
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * RD_ROWS;
    uint32_t t0[RD_ROWS], t1[RD_ROWS], t2[RD_ROWS], rs[RD_ROWS];
#pragma unroll
    for (int r = 0; r < RD_ROWS; ++r) t0[r] = t1[r] = t2[r] = rs[r] = 0u;
    const int64_t groups = (p + 15) / 16;
#pragma unroll 2
    for (int64_t g = threadIdx.x; g < groups; g += blockDim.x) {
        const int64_t k = g * 16;
        const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(planes + k));
        const uint4 b1 = __ldg(reinterpret_cast<const uint4 *>(planes + ldc + k));
        const uint4 b2 = __ldg(reinterpret_cast<const uint4 *>(planes + 2 * ldc + k));
        // all row loads of the group first (unconditional: rows beyond n are clamped onto the last row and their
        // sums discarded below), then the arithmetic — a conditional load per row left one load in flight per thread
        // (ncu: long-scoreboard stalls on the first dp4a of every row, HBM at 37 %)
        uint4 c[RD_ROWS];
#pragma unroll
        for (int r = 0; r < RD_ROWS; ++r) {
            const int64_t row = (r0 + r < n) ? r0 + r : n - 1;
            c[r] = __ldg(reinterpret_cast<const uint4 *>(codes + row * ldc + k));
        }
#pragma unroll
        for (int r = 0; r < RD_ROWS; ++r) {
            t0[r] = __dp4a(c[r].x, b0.x, __dp4a(c[r].y, b0.y, __dp4a(c[r].z, b0.z, __dp4a(c[r].w, b0.w, t0[r]))));
            t1[r] = __dp4a(c[r].x, b1.x, __dp4a(c[r].y, b1.y, __dp4a(c[r].z, b1.z, __dp4a(c[r].w, b1.w, t1[r]))));
            t2[r] = __dp4a(c[r].x, b2.x, __dp4a(c[r].y, b2.y, __dp4a(c[r].z, b2.z, __dp4a(c[r].w, b2.w, t2[r]))));
            rs[r] = __dp4a(c[r].x, 0x01010101u, __dp4a(c[r].y, 0x01010101u, __dp4a(c[r].z, 0x01010101u,
                                                                                      __dp4a(c[r].w, 0x01010101u, rs[r]))));
        }
    }
    __shared__ unsigned long long shT[RD_ROWS][8], shR[RD_ROWS][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < RD_ROWS; ++r) {
        unsigned long long T = static_cast<unsigned long long>(t0[r]) + (static_cast<unsigned long long>(t1[r]) << 8) +
                               (static_cast<unsigned long long>(t2[r]) << 16);
        unsigned long long R = rs[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            T += __shfl_down_sync(0xffffffffu, T, o);
            R += __shfl_down_sync(0xffffffffu, R, o);
        }
        if (lane == 0) {
            shT[r][warp] = T;
            shR[r][warp] = R;
        }
    }
    __syncthreads();
    if (threadIdx.x < RD_ROWS && r0 + threadIdx.x < n) {
        unsigned long long T = 0, R = 0;
        for (int wi = 0; wi < 8; ++wi) {
            T += shT[threadIdx.x][wi];
            R += shR[threadIdx.x][wi];
        }
        const int64_t row = r0 + threadIdx.x;
        if (accumulate) {
            T_out[row] += static_cast<long long>(T);
            r_out[row] += static_cast<long long>(R);
        } else {
            T_out[row] = static_cast<long long>(T);
            r_out[row] = static_cast<long long>(R);
        }
    }
}

// u_i = h r_i + T_i/(m n)
__global__ void centre_rows_kernel(const long long *__restrict__ r, const long long *__restrict__ T, int64_t n, double h,
                                   double mn, double *__restrict__ u) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) u[i] = h * static_cast<double>(r[i]) + static_cast<double>(T[i]) / mn;
}


// -------------------------------------------------------------------------------------------------------------
// pack of HOST-quantised chunks (ingest_host.cpp): int8 codes arrive entry-major ([locus][entry], row stride lds, the
// layout of the caller's matrix) and leave K-major ([entry][locus], the tcgen05 operand layout) — a byte transpose at
// 2 bytes of HBM traffic per cell, plus the integer column sums and the decoding of the sentinels the host wrote
// for cells that are not dosage codes.  Unit = 128 loci x 128 entries; a CTA walks a contiguous run of units (entry
// block fastest), so column sums stay in registers until the locus block changes.
//   load : warp w owns locus quads 4w .. 4w+3; lane l reads the 32-bit word of entries 4l .. 4l+3 of each of the 4
//          loci of a quad (one coalesced 128-byte line per locus), then a 4 x 4 byte transpose in registers
//   smem : tile[entry][locus word ^ swizzle] (33-word rows): conflict-free in both directions
//   write: 8 threads x 16 bytes = one 128-byte line per entry row
// -------------------------------------------------------------------------------------------------------------
constexpr int PC_TILE = 128;

template <bool MASK>
__global__ void __launch_bounds__(256, 3)
pack_codes_kernel(const int8_t *__restrict__ src, int64_t lds, int64_t n, int64_t pc, int8_t *__restrict__ codes,
                  int64_t ldc, int32_t *__restrict__ colsum, int32_t *__restrict__ miss, uint32_t *__restrict__ flags,
                  int64_t ent_blocks, int64_t units) {
    __shared__ uint32_t tile[PC_TILE][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
    int32_t acc[16], mis[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = mis[i] = 0;
    uint32_t bad = 0;
    int64_t cur_kb = -1;
    auto flush = [&](int64_t kb) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int32_t sum = __reduce_add_sync(0xffffffffu, acc[i]);
            const int64_t k = kb * PC_TILE + warp * 16 + i;
            if (lane == 0 && k < pc && sum != 0) atomicAdd(&colsum[k], sum);
            acc[i] = 0;
            if (MASK) {
                const int32_t ms = __reduce_add_sync(0xffffffffu, mis[i]);
                if (lane == 0 && k < pc && ms != 0) atomicAdd(&miss[k], ms);
                mis[i] = 0;
            }
        }
    };
    for (int64_t u = u0; u < u1; ++u) {
        const int64_t kb = u / ent_blocks, eb = u - kb * ent_blocks;
        if (kb != cur_kb) {
            if (cur_kb >= 0) flush(cur_kb);
            cur_kb = kb;
        }
        const int64_t k0 = kb * PC_TILE, e0 = eb * PC_TILE;
        const int64_t e = e0 + 4 * lane;
        // bytes of this lane's word that are real entries (the staging rows are only written up to n)
        const uint32_t emask = e + 3 < n ? 0xffffffffu : (e >= n ? 0u : (1u << (8 * static_cast<int>(n - e))) - 1u);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int64_t k = k0 + warp * 16 + i;
            w[i] = (k < pc && emask) ? (__ldg(reinterpret_cast<const uint32_t *>(src + k * lds + e)) & emask) : 0u;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (w[i] & 0x80808080u) {  // sentinels (negative bytes): decode, then pack the cell as 0
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t v = (w[i] >> (8 * b)) & 0xffu;
                    if (v & 0x80u) {
                        if (v == 0x80u) {
                            if (MASK) ++mis[i];
                            else bad |= FLAG_NAN;
                        } else {
                            bad |= v == 0x81u ? FLAG_NAN_PAYLOAD : (v == 0x82u ? FLAG_INF : FLAG_NOT_DOSAGE);
                        }
                        w[i] &= ~(0xffu << (8 * b));
                    }
                }
            }
            acc[i] = __dp4a(static_cast<int>(w[i]), 0x01010101, acc[i]);  // bytes are 0..64 here: sentinels were cleared
        }
        __syncthreads();  // the previous unit's tile has been written out
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            const uint32_t a = w[4 * qq], b = w[4 * qq + 1], c = w[4 * qq + 2], d = w[4 * qq + 3];
            const uint32_t ab_lo = __byte_perm(a, b, 0x5140), ab_hi = __byte_perm(a, b, 0x7362);
            const uint32_t cd_lo = __byte_perm(c, d, 0x5140), cd_hi = __byte_perm(c, d, 0x7362);
            const int q = warp * 4 + qq, sw = lane >> 3;
            tile[4 * lane + 0][q ^ sw] = __byte_perm(ab_lo, cd_lo, 0x5410);  // entry 4l:   loci 4q .. 4q+3
            tile[4 * lane + 1][q ^ sw] = __byte_perm(ab_lo, cd_lo, 0x7632);
            tile[4 * lane + 2][q ^ sw] = __byte_perm(ab_hi, cd_hi, 0x5410);
            tile[4 * lane + 3][q ^ sw] = __byte_perm(ab_hi, cd_hi, 0x7632);
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = it * 256 + threadIdx.x;
            const int r = idx >> 3, g = idx & 7, sw = (r >> 5) & 3;
            if (e0 + r < n) {
                const uint4 v = make_uint4(tile[r][4 * g + (0 ^ sw)], tile[r][4 * g + (1 ^ sw)], tile[r][4 * g + (2 ^ sw)],
                                           tile[r][4 * g + (3 ^ sw)]);
                *reinterpret_cast<uint4 *>(codes + (e0 + r) * ldc + k0 + 16 * g) = v;
            }
        }
    }
    if (cur_kb >= 0) flush(cur_kb);
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (lane == 0 && bad) atomicOr(flags, bad);
}

// QC locus mask from the integers of the pack pass — the reference's filter semantics (src/genomes/filter.jl):
//   sparsity_k = mean(ismissing(column))   keep iff sparsity_k <= max_locus_sparsity                 (:202-213, :361-368)
//   q_k = mean(skipmissing(column))        keep iff maf <= q_k <= 1 - maf                            (:632-635)
// For dosages sum(skipmissing(column)) = colsum/m exactly (any order), so q_k = fl(fl(colsum/m) / cnt) is the very
// number the FP64 pass (colmean_f64_kernel) computes.  A dropped locus gets colsum 0 (its codes are zeroed by
// mask_codes_kernel); a KEPT locus that still holds missing cells raises FLAG_NAN when the policy is "error".
__global__ void locus_keep_kernel(int32_t *__restrict__ colsum, const int32_t *__restrict__ miss, int64_t n, int64_t p,
                                  double m, double maf, double max_sparsity, int error_on_missing,
                                  uint8_t *__restrict__ keep, unsigned long long *__restrict__ kept,
                                  uint32_t *__restrict__ flags) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int kp = 0;
    uint32_t bad = 0;
    if (k < p) {
        const int32_t cnt = static_cast<int32_t>(n) - miss[k];
        const double sparsity = static_cast<double>(n - cnt) / static_cast<double>(n);
        const double qk = (static_cast<double>(colsum[k]) / m) / static_cast<double>(cnt);
        kp = (sparsity <= max_sparsity) && (qk >= maf) && (qk <= (1.0 - maf));
        keep[k] = static_cast<uint8_t>(kp);
        if (!kp) colsum[k] = 0;
        if (kp && error_on_missing && cnt < n) bad = FLAG_NAN;
    }
    const unsigned mk = __ballot_sync(0xffffffffu, kp);
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (mk) atomicAdd(kept, static_cast<unsigned long long>(__popc(mk)));
        if (bad) atomicOr(flags, bad);
    }
}

// codes[i, k] = 0 for every dropped locus k: thread = 16 loci of one entry row; rows whose 16 loci are all kept
// are not touched (keep is p bytes: L2-resident)
__global__ void __launch_bounds__(256)
mask_codes_kernel(int8_t *__restrict__ codes, int64_t n, int64_t p, int64_t ldc, const uint8_t *__restrict__ keep) {
    const int64_t groups = (p + 15) / 16;
    const int64_t g = static_cast<int64_t>(blockIdx.x) * 32 + (threadIdx.x & 31);
    if (g >= groups) return;
    const int64_t k = g * 16;
    uint32_t mk[4];
#pragma unroll
    for (int wv = 0; wv < 4; ++wv) {
        uint32_t mw = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t kk = k + 4 * wv + b;
            if (kk >= p || keep[kk]) mw |= 0xffu << (8 * b);
        }
        mk[wv] = mw;
    }
    if ((mk[0] & mk[1] & mk[2] & mk[3]) == 0xffffffffu) return;
    for (int64_t i = static_cast<int64_t>(blockIdx.y) * 8 + (threadIdx.x >> 5); i < n;
         i += static_cast<int64_t>(gridDim.y) * 8) {
        uint4 *ptr = reinterpret_cast<uint4 *>(codes + i * ldc + k);
        uint4 v = *ptr;
        v.x &= mk[0];
        v.y &= mk[1];
        v.z &= mk[2];
        v.w &= mk[3];
        *ptr = v;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host launchers / C ABI
// ---------------------------------------------------------------------------------------------------------------
int32_t grm_launch_detect(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda, int64_t max_cells,
                          uint32_t *d_out, int ignore_nan) {
    int64_t cells = n * p;
    if (max_cells > 0 && max_cells < cells) cells = max_cells;
    GRM_CUDA_TRY(cudaMemsetAsync(d_out, 0, 2 * sizeof(uint32_t), ctx->stream));
    const int blocks = static_cast<int>(std::min<int64_t>((cells + 255) / 256, static_cast<int64_t>(ctx->num_sms) * 16));
    detect_kernel<<<blocks, 256, 0, ctx->stream>>>(dA, n, p, lda, cells, d_out, ignore_nan);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_detect_denominator(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda,
                                               int64_t max_cells, int32_t ignore_missing, int32_t *m_out) try {
    if (!ctx || !dA || !m_out || n < 1 || p < 1 || lda < n) {
        grm_set_error("detect_denominator: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_TRY(ctx->misc.reserve(64));
    uint32_t *d_out = ctx->misc.as<uint32_t>();
    GRM_TRY(grm_launch_detect(ctx, dA, n, p, lda, max_cells, d_out, ignore_missing));
    uint32_t h[2];
    GRM_CUDA_TRY(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h[1] != 0) {
        *m_out = 0;  // continuous (or missing / non-finite: the caller's full pass reports which)
        return GRM_CUDA_OK;
    }
    // every code64 is a multiple of 2^tz  =>  x = (code64 >> tz) / (64 >> tz)
    int tz = 6;
    if (h[0] != 0) {
        tz = 0;
        while (((h[0] >> tz) & 1u) == 0) ++tz;
    }
    *m_out = 64 >> tz;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_pack_dosage(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t pc, int64_t lda,
                                        int32_t m, int8_t *d_codes, int64_t ldc, int32_t *d_colsum, uint32_t *d_flags,
                                        int32_t *d_miss) try {
    if (!ctx || !dA || !d_codes || !d_colsum || !d_flags || n < 1 || pc < 1 || lda < n || m < 1 || m > 64 ||
        (m & (m - 1)) != 0 || (ldc % 128) != 0 || ldc < ((pc + 127) / 128) * 128) {
        grm_set_error("pack_dosage: bad argument (n=%lld pc=%lld lda=%lld m=%d ldc=%lld)", (long long)n, (long long)pc,
                      (long long)lda, m, (long long)ldc);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    const bool use_tma = ctx->tune.pack != 1;
    if (use_tma && (lda % 2) == 0 && (reinterpret_cast<uintptr_t>(dA) % 16) == 0 && n < (1ll << 31) && pc < (1ll << 31)) {
        // TMA-staged variant: needs 16-byte aligned rows of A (global strides are multiples of 16 bytes)
        GrmEncodeTiledFn encode = nullptr;
        GRM_TRY(grm_tmap_encoder(&encode));
        CUtensorMap tmap;
        cuuint64_t gdim[2] = {static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(pc)};
        cuuint64_t gstride[1] = {static_cast<cuuint64_t>(lda) * sizeof(double)};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(PT_ENT), static_cast<cuuint32_t>(PT_LOCI)};
        cuuint32_t estr[2] = {1u, 1u};
        const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(dA), gdim, gstride, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            grm_set_error("cuTensorMapEncodeTiled (FP64 input) failed with CUresult %d", static_cast<int>(r));
            return GRM_CUDA_ERR_CUDA;
        }
        const int64_t ent_blocks = (n + PT_ENT - 1) / PT_ENT, loc_blocks = (pc + PT_LOCI - 1) / PT_LOCI;
        const int64_t units = ent_blocks * loc_blocks;
        const int grid = static_cast<int>(std::min<int64_t>(units, 2ll * ctx->num_sms));
        if (d_miss) {
            GRM_TRY(grm_func_smem(ctx, reinterpret_cast<const void *>(pack_dosage_tma_kernel<true>), PT_SMEM));
            pack_dosage_tma_kernel<true><<<grid, 288, PT_SMEM, ctx->stream>>>(tmap, n, pc, m, d_codes, ldc, d_colsum,
                                                                             d_flags, d_miss, ent_blocks, units);
        } else {
            GRM_TRY(grm_func_smem(ctx, reinterpret_cast<const void *>(pack_dosage_tma_kernel<false>), PT_SMEM));
            pack_dosage_tma_kernel<false><<<grid, 288, PT_SMEM, ctx->stream>>>(tmap, n, pc, m, d_codes, ldc, d_colsum,
                                                                              d_flags, nullptr, ent_blocks, units);
        }
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        return GRM_CUDA_OK;
    }
    constexpr int EITERS = 4;
    dim3 grid(static_cast<unsigned>((pc + PK_LOCI - 1) / PK_LOCI),
              static_cast<unsigned>((n + PK_SLAB * EITERS - 1) / (PK_SLAB * EITERS)));
    if (d_miss)
        pack_dosage_kernel<EITERS, true><<<grid, 256, 0, ctx->stream>>>(dA, n, pc, lda, m, d_codes, ldc, d_colsum,
                                                                        d_flags, d_miss);
    else
        pack_dosage_kernel<EITERS, false><<<grid, 256, 0, ctx->stream>>>(dA, n, pc, lda, m, d_codes, ldc, d_colsum,
                                                                         d_flags, nullptr);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_pack_codes(grm_cuda_ctx *ctx, const int8_t *d_src, int64_t n, int64_t pc, int64_t lds,
                                       int8_t *d_codes, int64_t ldc, int32_t *d_colsum, uint32_t *d_flags,
                                       int32_t *d_miss) try {
    if (!ctx || !d_src || !d_codes || !d_colsum || !d_flags || n < 1 || pc < 1 || lds < n || (lds % 4) != 0 ||
        (reinterpret_cast<uintptr_t>(d_src) % 4) != 0 || (ldc % 128) != 0 || ldc < ((pc + 127) / 128) * 128) {
        grm_set_error("pack_codes: bad argument (n=%lld pc=%lld lds=%lld ldc=%lld)", (long long)n, (long long)pc,
                      (long long)lds, (long long)ldc);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t ent_blocks = (n + PC_TILE - 1) / PC_TILE, loc_blocks = (pc + PC_TILE - 1) / PC_TILE;
    const int64_t units = ent_blocks * loc_blocks;
    const int grid = static_cast<int>(std::min<int64_t>(units, 3ll * ctx->num_sms));
    if (d_miss)
        pack_codes_kernel<true><<<grid, 256, 0, ctx->stream>>>(d_src, lds, n, pc, d_codes, ldc, d_colsum, d_miss, d_flags,
                                                               ent_blocks, units);
    else
        pack_codes_kernel<false><<<grid, 256, 0, ctx->stream>>>(d_src, lds, n, pc, d_codes, ldc, d_colsum, nullptr,
                                                                d_flags, ent_blocks, units);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_locus_keep(grm_cuda_ctx *ctx, int32_t *d_colsum, const int32_t *d_miss, int64_t n, int64_t p,
                                       int32_t m, double maf, double max_locus_sparsity, int32_t error_on_missing,
                                       uint8_t *d_keep, uint64_t *d_kept, uint32_t *d_flags) try {
    if (!ctx || !d_colsum || !d_miss || !d_keep || !d_kept || !d_flags || n < 1 || p < 1 || m < 1) {
        grm_set_error("locus_keep: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    locus_keep_kernel<<<static_cast<unsigned>((p + 255) / 256), 256, 0, ctx->stream>>>(
        d_colsum, d_miss, n, p, static_cast<double>(m), maf, max_locus_sparsity, error_on_missing, d_keep,
        reinterpret_cast<unsigned long long *>(d_kept), d_flags);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_mask_codes(grm_cuda_ctx *ctx, int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                       const uint8_t *d_keep) try {
    if (!ctx || !d_codes || !d_keep || n < 1 || p < 1 || (ldc % 128) != 0 || ldc < p) {
        grm_set_error("mask_codes: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t groups = (p + 15) / 16;
    dim3 grid(static_cast<unsigned>((groups + 31) / 32), static_cast<unsigned>(std::min<int64_t>((n + 7) / 8, 64)));
    mask_codes_kernel<<<grid, 256, 0, ctx->stream>>>(d_codes, n, p, ldc, d_keep);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_locus_stats_i8(grm_cuda_ctx *ctx, const int32_t *d_colsum, int64_t p, int64_t n, int32_t m,
                                           int32_t ploidy, double *d_q, double *d_w, double *d_stats) try {
    if (!ctx || !d_colsum || !d_q || !d_w || !d_stats || p < 1 || n < 1 || m < 1) {
        grm_set_error("locus_stats_i8: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_TRY(ctx->partials.reserve(3 * RED_BLOCKS * sizeof(double)));
    locus_stats_i8_kernel<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(d_colsum, p, static_cast<double>(n),
                                                                       static_cast<double>(m), 0.5 * ploidy, d_q, d_w,
                                                                       ctx->partials.as<double>());
    GRM_CUDA_TRY(cudaGetLastError());
    final_stats_kernel<<<1, 32, 0, ctx->stream>>>(ctx->partials.as<double>(), RED_BLOCKS, d_stats, 0);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_locus_stats_f64(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda,
                                            double *d_q, double *d_stats, uint32_t *d_flags, int32_t accumulate,
                                            int32_t skip_missing) try {
    if (!ctx || !dA || !d_q || !d_stats || !d_flags || p < 1 || n < 1 || lda < n) {
        grm_set_error("locus_stats_f64: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_TRY(ctx->partials.reserve(3 * RED_BLOCKS * sizeof(double)));
    colmean_f64_kernel<<<static_cast<unsigned>((p + 7) / 8), 256, 0, ctx->stream>>>(dA, n, p, lda, d_q, d_flags,
                                                                                    skip_missing, nullptr);
    GRM_CUDA_TRY(cudaGetLastError());
    locus_stats_f64_kernel<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(d_q, p, ctx->partials.as<double>(), nullptr);
    GRM_CUDA_TRY(cudaGetLastError());
    final_stats_kernel<<<1, 32, 0, ctx->stream>>>(ctx->partials.as<double>(), RED_BLOCKS, d_stats, accumulate);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches += 3;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_rowdot_i8(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                      const double *d_w, double *d_u) try {
    if (!ctx || !d_codes || !d_w || !d_u || n < 1 || p < 1 || (ldc % 128) != 0 || ldc < p) {
        grm_set_error("rowdot_i8: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    rowdot_i8_kernel<<<static_cast<unsigned>((n + RD_ROWS - 1) / RD_ROWS), 256, 0, ctx->stream>>>(d_codes, n, p, ldc,
                                                                                                 d_w, d_u);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_dosage_stats(grm_cuda_ctx *ctx, const int8_t *d_codes, int64_t n, int64_t p, int64_t ldc,
                                         const int32_t *d_colsum, int32_t m, uint64_t *d_sums, int64_t *d_r,
                                         int64_t *d_T, int32_t accumulate) try {
    if (!ctx || !d_codes || !d_colsum || !d_sums || !d_r || !d_T || n < 1 || p < 1 || (ldc % 128) != 0 || ldc < p ||
        m < 1 || m > 64) {
        grm_set_error("dosage_stats: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (n >= (1ll << 18)) {  // S_k <= 64 n must fit the three byte planes
        grm_set_error("dosage_stats: n = %lld exceeds 2^18 entries", (long long)n);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    // sum_k S_k^2 <= p (m n)^2 is accumulated in 64 bits: refuse shapes where even the bound could wrap
    const double mnf = static_cast<double>(m) * static_cast<double>(n);
    if (static_cast<double>(p) * mnf * mnf >= 9.0e18) {
        grm_set_error("dosage_stats: p * (m n)^2 = %.3g may overflow the 64-bit sum of squared column sums",
                      static_cast<double>(p) * mnf * mnf);
        return GRM_CUDA_ERR_OVERFLOW;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_TRY(ctx->planes.reserve(static_cast<size_t>(3 * ldc)));
    uint8_t *planes = ctx->planes.as<uint8_t>();
    if (!accumulate) GRM_CUDA_TRY(cudaMemsetAsync(d_sums, 0, 2 * sizeof(uint64_t), ctx->stream));
    const int blocks = static_cast<int>(std::min<int64_t>((ldc + 255) / 256, 1024));
    colsum_planes_kernel<<<blocks, 256, 0, ctx->stream>>>(d_colsum, p, ldc, planes,
                                                          reinterpret_cast<unsigned long long *>(d_sums));
    GRM_CUDA_TRY(cudaGetLastError());
    rowstats_i8_kernel<<<static_cast<unsigned>((n + RD_ROWS - 1) / RD_ROWS), 256, 0, ctx->stream>>>(
        d_codes, n, p, ldc, planes, reinterpret_cast<long long *>(d_r), reinterpret_cast<long long *>(d_T), accumulate);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_dosage_centre(grm_cuda_ctx *ctx, const uint64_t *d_sums, const int64_t *d_r,
                                          const int64_t *d_T, int64_t n, int64_t p, int32_t m, int32_t ploidy,
                                          double *d_u, double *scalars /* host: alpha, beta, gamma, denom */) try {
    if (!ctx || !d_sums || !d_r || !d_T || !d_u || !scalars || n < 1 || p < 1 || m < 1) {
        grm_set_error("dosage_centre: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t h_sums[2];
    GRM_CUDA_TRY(cudaMemcpyAsync(h_sums, d_sums, sizeof(h_sums), cudaMemcpyDeviceToHost, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const double mn = static_cast<double>(m) * static_cast<double>(n);  // exact (< 2^24)
    const double h = 0.5 * ploidy;
    const double s = static_cast<double>(ploidy) / static_cast<double>(m);
    // d = ploidy * (mn*sum(S) - sum(S^2)) / mn^2 with the numerator exact in 128-bit integers (calc.jl:201)
    const unsigned __int128 num = static_cast<unsigned __int128>(static_cast<uint64_t>(m) * static_cast<uint64_t>(n)) *
                                      h_sums[0] - static_cast<unsigned __int128>(h_sums[1]);
    const double S1 = static_cast<double>(h_sums[0]), S2 = static_cast<double>(h_sums[1]);
    const double denom = static_cast<double>(ploidy) * (static_cast<double>(num) / (mn * mn));
    const double W2 = (static_cast<double>(p) * h * h + 2.0 * h * (S1 / mn)) + S2 / (mn * mn);
    scalars[0] = s * s;
    scalars[1] = s;
    scalars[2] = W2;
    scalars[3] = denom;
    centre_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(
        reinterpret_cast<const long long *>(d_r), reinterpret_cast<const long long *>(d_T), n, h, mn, d_u);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

// QC locus mask in one call: q[k] = mean(skipmissing(column k)), keep[k] per the reference's filter rules, the number
// of kept loci added to *d_kept, and (d_stats != NULL) {sum q(1-q), 0, sum q} over the kept loci.
extern "C" int32_t grm_cuda_locus_qc(grm_cuda_ctx *ctx, const double *dA, int64_t n, int64_t p, int64_t lda, double maf,
                                     double max_locus_sparsity, int32_t error_on_missing, double *d_q, uint8_t *d_keep,
                                     uint64_t *d_kept, double *d_stats, int32_t accumulate_stats, uint32_t *d_flags) try {
    if (!ctx || !dA || !d_q || !d_keep || !d_kept || !d_flags || n < 1 || p < 1 || lda < n) {
        grm_set_error("locus_qc: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    GRM_TRY(ctx->cnt.reserve(static_cast<size_t>(p) * sizeof(int32_t)));
    GRM_TRY(ctx->partials.reserve(3 * RED_BLOCKS * sizeof(double)));
    colmean_f64_kernel<<<static_cast<unsigned>((p + 7) / 8), 256, 0, ctx->stream>>>(dA, n, p, lda, d_q, d_flags, 1,
                                                                                    ctx->cnt.as<int32_t>());
    GRM_CUDA_TRY(cudaGetLastError());
    locus_mask_kernel<<<static_cast<unsigned>((p + 255) / 256), 256, 0, ctx->stream>>>(
        d_q, ctx->cnt.as<int32_t>(), n, p, maf, max_locus_sparsity, error_on_missing, d_keep,
        reinterpret_cast<unsigned long long *>(d_kept), d_flags);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches += 2;
    if (d_stats) {
        locus_stats_f64_kernel<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(d_q, p, ctx->partials.as<double>(), d_keep);
        GRM_CUDA_TRY(cudaGetLastError());
        final_stats_kernel<<<1, 32, 0, ctx->stream>>>(ctx->partials.as<double>(), RED_BLOCKS, d_stats, accumulate_stats);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches += 2;
    }
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS
