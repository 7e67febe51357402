// distances.cu — pairwise distance / correlation matrices of distances(genomes).
//
// Reference: src/genomes/genomes.jl:454-654.  For a sorted, unique subset L of the loci-alleles (t of them) the
// reference builds, per requested metric, a t x t matrix over pairs of loci (vectors of length n = entries) and an
// n x n matrix over pairs of entries (vectors of length t), each pair restricted to the positions where BOTH vectors
// are non-missing, non-NaN and finite; pairs with fewer than 2 such positions keep the fill value (+Inf, or -Inf for
// "correlation").  Metrics (genomes.jl:556-569, :617-631):
//     euclidean   sqrt(sum((y1 - y2).^2))
//     correlation cor(y1, y2), skipped (-Inf) when var(y1) < 1e-7 or var(y2) < 1e-7
//     mad         mean(abs.(y1 - y2))
//     rmsd        sqrt(mean((y1 - y2).^2))
//     χ²          sum((y1 - y2).^2 ./ (y2 .+ eps(Float64)))          (not symmetric in y1, y2)
// Loci pairs are evaluated for i <= j only and mirrored (D[i,j] = D[j,i], so χ² uses the lower-indexed locus as y1),
// and their counts fill the upper triangle only; entry pairs are evaluated for every ordered (i, j).
//
// This is pairwise FP64 ALU work on a small operand (the reference's default is 100 sampled loci), not a GEMM: the
// per-pair index set depends on both vectors, and |a - b| and the per-pair means of cor are not bilinear.  One kernel
// serves both orientations: a 32 x 32 tile of pairs per CTA, operands staged through shared memory in slabs of 32
// positions, every thread accumulating 2 x 2 pairs sequentially over the positions (a fixed order: results do not
// depend on the launch shape).  cor is evaluated like Statistics.corm (means first, then centred sums, then
// xy / max(xx,yy) / sqrt(min(xx,yy) / max(xx,yy)), clamped to [-1, 1]) in a second sweep over the positions.
// When the pair grid is too small to fill the GPU (the default case: 100 x 100 loci over all entries = 16 tiles) the
// positions are sliced over grid.z and the partial sums are added in slice order by small reduction kernels.
#include <float.h>
#include <math.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int DT = 32;  // pairs tile edge
constexpr int DK = 32;  // positions per slab

// S[i + k*n] = A[i + L[k]*lda]: the selected loci as a compact n x t matrix
__global__ void gather_loci_kernel(const double *__restrict__ A, int64_t n, int64_t lda, const int64_t *__restrict__ L,
                                   double *__restrict__ S) {
    const int64_t k = blockIdx.y;
    const double *col = A + L[k] * lda;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        S[i + k * n] = col[i];
}

struct DistOut {
    double *euclid, *corr, *mad, *rmsd, *chi2, *counts;
    int64_t ld;
};

// accumulators of one pair: pass 1 = {cnt, ss, sa, sx, sy, c12, c21}, pass 2 = {xx, yy, xy}
constexpr int NACC1 = 7, NACC2 = 3;

struct PairArgs {
    const double *X;      // rows: r vectors of length c; element (row i, position k) at X[i*rs + k*ks]
    int64_t r, c, rs, ks;
    int upper_rule;       // loci: pair (i, j) with i > j is evaluated as (j, i); counts only for i <= j
    int want_corr, want_chi2;
    int64_t chunk;        // positions per grid.z slice (>= c: no split)
    double *part1;        // split mode: [z][pair][NACC1] partial sums, pair = i + j*r
    double *part2;        // split mode: [z][pair][NACC2]
    const double *means;  // split mode, pass 2: [pair][2]
};

// staging: 256 threads fill DK x DT slabs of both row blocks; the thread index runs along whichever direction is
// contiguous in memory.  Unusable cells (NaN = missing, +-Inf) are stored as NaN, so the pair loop tests x == x only;
// for chi2 the reciprocal 1/(x + eps) is formed once per staged cell instead of one FP64 division per pair and
// position (each term then differs from the quotient by at most one rounding).
__device__ __forceinline__ void stage_slab(const PairArgs &a, int64_t i0, int64_t j0, int64_t k0, int64_t kend,
                                           double (*sI)[DT + 1], double (*sJ)[DT + 1], double (*wI)[DT + 1],
                                           double (*wJ)[DT + 1]) {
#pragma unroll
    for (int it = 0; it < (DK * DT) / 256; ++it) {
        const int idx = it * 256 + threadIdx.x;
        int kk, rr;
        if (a.rs == 1) {
            rr = idx % DT;
            kk = idx / DT;
        } else {
            kk = idx % DK;
            rr = idx / DK;
        }
        const int64_t k = k0 + kk;
        const int64_t ri = i0 + rr, rj = j0 + rr;
        double xi = (k < kend && ri < a.r) ? a.X[ri * a.rs + k * a.ks] : NAN;
        double xj = (k < kend && rj < a.r) ? a.X[rj * a.rs + k * a.ks] : NAN;
        if (fabs(xi) == INFINITY) xi = NAN;
        if (fabs(xj) == INFINITY) xj = NAN;
        sI[kk][rr] = xi;
        sJ[kk][rr] = xj;
        if (a.want_chi2) {
            wJ[kk][rr] = 1.0 / (xj + DBL_EPSILON);
            if (a.upper_rule) wI[kk][rr] = 1.0 / (xi + DBL_EPSILON);
        }
    }
}

// PASS 1: cnt, sum d^2, sum |d|, sum x, sum y, chi2 both ways.  PASS 2: centred sums for cor (means known).
template <int PASS, bool CHI2>
__device__ __forceinline__ void sweep(const PairArgs &a, int64_t i0, int64_t j0, int64_t kbeg, int64_t kend, int tx, int ty,
                                      const double (&mx)[2][2], const double (&my)[2][2], double (&acc)[2][2][NACC1],
                                      double (*sm)[DK][DT + 1]) {
    double(*sI)[DT + 1] = sm[0], (*sJ)[DT + 1] = sm[1], (*wI)[DT + 1] = sm[2], (*wJ)[DT + 1] = sm[3];
    for (int64_t k0 = kbeg; k0 < kend; k0 += DK) {
        __syncthreads();
        stage_slab(a, i0, j0, k0, kend, sI, sJ, wI, wJ);
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < DK; ++kk) {
            const double xi[2] = {sI[kk][2 * tx], sI[kk][2 * tx + 1]};
            const double xj[2] = {sJ[kk][2 * ty], sJ[kk][2 * ty + 1]};
            double vi[2] = {0.0, 0.0}, vj[2] = {0.0, 0.0};
            if (PASS == 1 && CHI2) {
                vj[0] = wJ[kk][2 * ty];
                vj[1] = wJ[kk][2 * ty + 1];
                if (a.upper_rule) {
                    vi[0] = wI[kk][2 * tx];
                    vi[1] = wI[kk][2 * tx + 1];
                }
            }
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (xi[p] == xi[p] && xj[q] == xj[q]) {
                        double *s = acc[p][q];
                        if (PASS == 1) {
                            const double d = xi[p] - xj[q];
                            const double d2 = d * d;
                            s[0] += 1.0;
                            s[1] += d2;
                            s[2] += fabs(d);
                            s[3] += xi[p];
                            s[4] += xj[q];
                            if (CHI2) {
                                s[5] += d2 * vj[q];  // y1 = row i, y2 = row j
                                s[6] += d2 * vi[p];  // y1 = row j, y2 = row i (needed under the upper rule only)
                            }
                        } else {
                            const double u = xi[p] - mx[p][q], v = xj[q] - my[p][q];
                            s[0] += u * u;
                            s[1] += v * v;
                            s[2] += u * v;
                        }
                    }
                }
        }
    }
}

__device__ __forceinline__ void finish_pair(const DistOut &out, int64_t i, int64_t j, int upper_rule, const double *s1,
                                            const double *s2) {
    const int64_t o = i + j * out.ld;
    const double N = s1[0];
    const bool ok = N >= 2.0;  // genomes.jl:545-547, :611-613
    const bool swapped = upper_rule && i > j;
    if (out.counts) out.counts[o] = (ok && !swapped) ? N : 0.0;
    if (out.euclid) out.euclid[o] = ok ? sqrt(s1[1]) : INFINITY;
    if (out.mad) out.mad[o] = ok ? s1[2] / N : INFINITY;
    if (out.rmsd) out.rmsd[o] = ok ? sqrt(s1[1] / N) : INFINITY;
    if (out.chi2) out.chi2[o] = ok ? (swapped ? s1[6] : s1[5]) : INFINITY;
    if (out.corr) {
        double v = -INFINITY;
        if (ok) {
            const double xx = s2[0], yy = s2[1], xy = s2[2];
            const double vx = xx / (N - 1.0), vy = yy / (N - 1.0);  // var(): corrected
            if (!(vx < 1e-7) && !(vy < 1e-7)) {
                const double hi = fmax(xx, yy), lo = fmin(xx, yy);
                v = xy / hi / sqrt(lo / hi);   // Statistics.corm
                v = fmin(1.0, fmax(-1.0, v));  // clampcor
            }
        }
        out.corr[o] = v;
    }
}

// fused form: one CTA sees all positions of its 32 x 32 pairs
template <bool CHI2>
__global__ void __launch_bounds__(256) pair_kernel(PairArgs a, DistOut out) {
    __shared__ double sm[4][DK][DT + 1];  // row block i, row block j, and their chi2 reciprocals
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 2 x 2 pairs each
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * DT, j0 = static_cast<int64_t>(blockIdx.y) * DT;
    double acc1[2][2][NACC1], acc2[2][2][NACC1], mx[2][2], my[2][2];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            mx[p][q] = my[p][q] = 0.0;
#pragma unroll
            for (int f = 0; f < NACC1; ++f) acc1[p][q][f] = acc2[p][q][f] = 0.0;
        }
    sweep<1, CHI2>(a, i0, j0, 0, a.c, tx, ty, mx, my, acc1, sm);
    if (a.want_corr) {
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                mx[p][q] = acc1[p][q][3] / acc1[p][q][0];
                my[p][q] = acc1[p][q][4] / acc1[p][q][0];
            }
        sweep<2, false>(a, i0, j0, 0, a.c, tx, ty, mx, my, acc2, sm);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int64_t i = i0 + 2 * tx + p, j = j0 + 2 * ty + q;
            if (i < a.r && j < a.r) finish_pair(out, i, j, a.upper_rule, acc1[p][q], acc2[p][q]);
        }
}

// split form (few pairs, long vectors: the default 100 loci x 100 loci over all entries is 16 tiles): grid.z slices of
// the positions write partial sums; they are added in slice order by reduce kernels, so the result is deterministic.
template <int PASS, bool CHI2>
__global__ void __launch_bounds__(256) pair_partial_kernel(PairArgs a) {
    __shared__ double sm[4][DK][DT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * DT, j0 = static_cast<int64_t>(blockIdx.y) * DT;
    const int64_t kbeg = static_cast<int64_t>(blockIdx.z) * a.chunk, kend = kbeg + a.chunk < a.c ? kbeg + a.chunk : a.c;
    double acc[2][2][NACC1], mx[2][2], my[2][2];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int64_t i = i0 + 2 * tx + p, j = j0 + 2 * ty + q;
            const bool in = i < a.r && j < a.r;
            mx[p][q] = (PASS == 2 && in) ? a.means[2 * (i + j * a.r)] : 0.0;
            my[p][q] = (PASS == 2 && in) ? a.means[2 * (i + j * a.r) + 1] : 0.0;
#pragma unroll
            for (int f = 0; f < NACC1; ++f) acc[p][q][f] = 0.0;
        }
    sweep<PASS, CHI2>(a, i0, j0, kbeg, kend, tx, ty, mx, my, acc, sm);
    constexpr int NA = PASS == 1 ? NACC1 : NACC2;
    double *part = (PASS == 1 ? a.part1 : a.part2) + static_cast<int64_t>(blockIdx.z) * a.r * a.r * NA;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int64_t i = i0 + 2 * tx + p, j = j0 + 2 * ty + q;
            if (i < a.r && j < a.r)
#pragma unroll
                for (int f = 0; f < NA; ++f) part[(i + j * a.r) * NA + f] = acc[p][q][f];
        }
}

// pass-1 partials -> slice 0 (in place, fixed order) and the per-pair means for pass 2
__global__ void reduce1_kernel(double *part1, int64_t pairs, int Z, double *means) {
    const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (o >= pairs) return;
    double s[NACC1];
#pragma unroll
    for (int f = 0; f < NACC1; ++f) s[f] = part1[o * NACC1 + f];
    for (int z = 1; z < Z; ++z)
#pragma unroll
        for (int f = 0; f < NACC1; ++f) s[f] += part1[(static_cast<int64_t>(z) * pairs + o) * NACC1 + f];
#pragma unroll
    for (int f = 0; f < NACC1; ++f) part1[o * NACC1 + f] = s[f];
    means[2 * o] = s[3] / s[0];
    means[2 * o + 1] = s[4] / s[0];
}

__global__ void finish_split_kernel(const double *part1, const double *part2, int64_t r, int Z, int upper_rule, int want_corr,
                                    DistOut out) {
    const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t pairs = r * r;
    if (o >= pairs) return;
    double s2[NACC2] = {0.0, 0.0, 0.0};
    if (want_corr)
        for (int z = 0; z < Z; ++z)
#pragma unroll
            for (int f = 0; f < NACC2; ++f) s2[f] += part2[(static_cast<int64_t>(z) * pairs + o) * NACC2 + f];
    finish_pair(out, o % r, o / r, upper_rule, part1 + o * NACC1, s2);
}

}  // namespace

extern "C" int32_t grm_cuda_distances(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda,
                                      int32_t a_location, const int64_t *idx_loci, int64_t t, int32_t dimension,
                                      double *euclidean, double *correlation, double *mad, double *rmsd, double *chi2,
                                      double *counts, int64_t ldo, int32_t out_location) try {
    const int64_t r = dimension == GRM_CUDA_DIM_ENTRIES ? n : t;
    if (!ctx || !A || !idx_loci || n < 1 || p < 1 || lda < n || t < 1 || t > p ||
        (dimension != GRM_CUDA_DIM_ENTRIES && dimension != GRM_CUDA_DIM_LOCI) || ldo < r ||
        (a_location != GRM_CUDA_LOC_HOST && a_location != GRM_CUDA_LOC_DEVICE) ||
        (out_location != GRM_CUDA_LOC_HOST && out_location != GRM_CUDA_LOC_DEVICE) ||
        (!euclidean && !correlation && !mad && !rmsd && !chi2 && !counts)) {
        grm_set_error("distances: bad argument (n=%lld p=%lld t=%lld dimension=%d ldo=%lld)", (long long)n, (long long)p,
                      (long long)t, dimension, (long long)ldo);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    for (int64_t k = 0; k < t; ++k) {
        if (idx_loci[k] < 0 || idx_loci[k] >= p || (k > 0 && idx_loci[k] <= idx_loci[k - 1])) {
            // the reference sorts and deduplicates (genomes.jl:500) after its range check (:497-499)
            grm_set_error("We accept `idx_loci_alleles` from 1 to `n_loci_alleles` of `genomes` (sorted, unique).");
            return GRM_CUDA_ERR_BAD_ARGUMENT;
        }
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    // the selected loci as a compact n x t device matrix
    GRM_TRY(ctx->misc2.reserve(static_cast<size_t>(n) * t * sizeof(double)));
    double *S = ctx->misc2.as<double>();
    if (a_location == GRM_CUDA_LOC_HOST) {
        for (int64_t k = 0; k < t; ++k)  // a locus is one contiguous column of the caller's matrix
            GRM_CUDA_TRY(cudaMemcpyAsync(S + k * n, A + idx_loci[k] * lda, static_cast<size_t>(n) * sizeof(double),
                                         cudaMemcpyHostToDevice, st));
    } else {
        GRM_TRY(ctx->keep.reserve(static_cast<size_t>(t) * sizeof(int64_t)));
        GRM_CUDA_TRY(cudaMemcpyAsync(ctx->keep.ptr, idx_loci, static_cast<size_t>(t) * sizeof(int64_t),
                                     cudaMemcpyHostToDevice, st));
        const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 64));
        for (int64_t k0 = 0; k0 < t; k0 += 65535) {
            const int64_t len = std::min<int64_t>(65535, t - k0);
            gather_loci_kernel<<<dim3(gx, static_cast<unsigned>(len)), 256, 0, st>>>(A, n, lda, ctx->keep.as<int64_t>() + k0,
                                                                                    S + k0 * n);
            GRM_CUDA_TRY(cudaGetLastError());
            ctx->launches++;
        }
    }

    // outputs: straight into the caller's device matrices, or into a device scratch that is copied back
    double *host_dst[6] = {euclidean, correlation, mad, rmsd, chi2, counts};
    double *dev_dst[6] = {euclidean, correlation, mad, rmsd, chi2, counts};
    int64_t ldd = ldo;
    if (out_location == GRM_CUDA_LOC_HOST) {
        int wanted = 0;
        for (double *q : host_dst) wanted += q != nullptr;
        GRM_TRY(ctx->G.reserve(static_cast<size_t>(wanted) * r * r * sizeof(double)));
        double *base = ctx->G.as<double>();
        int slot = 0;
        for (int m = 0; m < 6; ++m) dev_dst[m] = host_dst[m] ? base + static_cast<int64_t>(slot++) * r * r : nullptr;
        ldd = r;
    }
    DistOut out{dev_dst[0], dev_dst[1], dev_dst[2], dev_dst[3], dev_dst[4], dev_dst[5], ldd};
    const int64_t tiles = (r + DT - 1) / DT;
    if (tiles > 65535) {
        grm_set_error("distances: %lld rows exceed the supported pair grid", (long long)r);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    const bool entries = dimension == GRM_CUDA_DIM_ENTRIES;
    PairArgs pa{};
    pa.X = S;
    pa.r = r;
    pa.c = entries ? t : n;
    pa.rs = entries ? 1 : n;
    pa.ks = entries ? n : 1;
    pa.upper_rule = entries ? 0 : 1;
    pa.want_corr = correlation != nullptr;
    pa.want_chi2 = chi2 != nullptr;
    pa.chunk = pa.c;
    // few pairs and long vectors: slice the positions over grid.z so that the GPU is filled (>= 2 CTAs per SM)
    int Z = 1;
    if (tiles * tiles < 2 * ctx->num_sms && pa.c >= 512) {
        Z = static_cast<int>(std::min<int64_t>((2 * ctx->num_sms + tiles * tiles - 1) / (tiles * tiles), pa.c / 256));
        Z = std::max(1, std::min(Z, 1024));
    }
    const dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(tiles), static_cast<unsigned>(Z));
    if (Z == 1) {
        if (pa.want_chi2) pair_kernel<true><<<grid, 256, 0, st>>>(pa, out);
        else pair_kernel<false><<<grid, 256, 0, st>>>(pa, out);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
    } else {
        pa.chunk = ((pa.c + Z - 1) / Z + DK - 1) / DK * DK;
        Z = static_cast<int>((pa.c + pa.chunk - 1) / pa.chunk);
        const int64_t pairs = r * r;
        GRM_TRY(ctx->lu.reserve(static_cast<size_t>(Z) * pairs * (NACC1 + NACC2) * sizeof(double) +
                                static_cast<size_t>(pairs) * 2 * sizeof(double)));
        ctx->factor_n = -1;  // the factor cache lives in the same workspace
        pa.part1 = ctx->lu.as<double>();
        pa.part2 = pa.part1 + static_cast<int64_t>(Z) * pairs * NACC1;
        double *means = pa.part2 + static_cast<int64_t>(Z) * pairs * NACC2;
        pa.means = means;
        const dim3 gz(static_cast<unsigned>(tiles), static_cast<unsigned>(tiles), static_cast<unsigned>(Z));
        const unsigned rb = static_cast<unsigned>((pairs + 255) / 256);
        if (pa.want_chi2) pair_partial_kernel<1, true><<<gz, 256, 0, st>>>(pa);
        else pair_partial_kernel<1, false><<<gz, 256, 0, st>>>(pa);
        GRM_CUDA_TRY(cudaGetLastError());
        reduce1_kernel<<<rb, 256, 0, st>>>(pa.part1, pairs, Z, means);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches += 2;
        if (pa.want_corr) {
            pair_partial_kernel<2, false><<<gz, 256, 0, st>>>(pa);
            GRM_CUDA_TRY(cudaGetLastError());
            ctx->launches++;
        }
        finish_split_kernel<<<rb, 256, 0, st>>>(pa.part1, pa.part2, r, Z, pa.upper_rule, pa.want_corr, out);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
    }
    if (out_location == GRM_CUDA_LOC_HOST) {
        for (int m = 0; m < 6; ++m)
            if (host_dst[m])
                GRM_CUDA_TRY(cudaMemcpy2DAsync(host_dst[m], ldo * sizeof(double), dev_dst[m], r * sizeof(double),
                                               r * sizeof(double), static_cast<size_t>(r), cudaMemcpyDeviceToHost, st));
    }
    GRM_CUDA_TRY(cudaStreamSynchronize(st));
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS
