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

__device__ __forceinline__ bool usable(double x) { return x == x && fabs(x) != INFINITY; }

// rows: r vectors of length c; element (row i, position k) at X[i*rs + k*ks]
// upper_rule (loci): pair (i, j) with i > j is evaluated as (j, i); counts only for i <= j
__global__ void __launch_bounds__(256)
pair_kernel(const double *__restrict__ X, int64_t r, int64_t c, int64_t rs, int64_t ks, int upper_rule, int want_corr,
            int want_chi2, DistOut out) {
    __shared__ double sI[DK][DT + 1], sJ[DK][DT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 2 x 2 pairs each
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * DT, j0 = static_cast<int64_t>(blockIdx.y) * DT;

    double cnt[2][2], ss[2][2], sa[2][2], sx[2][2], sy[2][2], c12[2][2], c21[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) cnt[a][b] = ss[a][b] = sa[a][b] = sx[a][b] = sy[a][b] = c12[a][b] = c21[a][b] = 0.0;

    // staging: 256 threads fill two DK x DT slabs; the thread index runs along whichever direction is contiguous
    auto stage = [&](int64_t k0) {
#pragma unroll
        for (int it = 0; it < (DK * DT) / 256; ++it) {
            const int idx = it * 256 + threadIdx.x;
            int kk, rr;
            if (rs == 1) {
                rr = idx % DT;
                kk = idx / DT;
            } else {
                kk = idx % DK;
                rr = idx / DK;
            }
            const int64_t k = k0 + kk;
            const int64_t ri = i0 + rr, rj = j0 + rr;
            sI[kk][rr] = (k < c && ri < r) ? X[ri * rs + k * ks] : NAN;
            sJ[kk][rr] = (k < c && rj < r) ? X[rj * rs + k * ks] : NAN;
        }
    };

    for (int64_t k0 = 0; k0 < c; k0 += DK) {
        __syncthreads();
        stage(k0);
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < DK; ++kk) {
            double xi[2] = {sI[kk][2 * tx], sI[kk][2 * tx + 1]};
            double xj[2] = {sJ[kk][2 * ty], sJ[kk][2 * ty + 1]};
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    if (usable(xi[a]) && usable(xj[b])) {
                        const double d = xi[a] - xj[b];
                        const double d2 = d * d;
                        cnt[a][b] += 1.0;
                        ss[a][b] += d2;
                        sa[a][b] += fabs(d);
                        sx[a][b] += xi[a];
                        sy[a][b] += xj[b];
                        if (want_chi2) {
                            c12[a][b] += d2 / (xj[b] + DBL_EPSILON);  // y1 = row i, y2 = row j
                            c21[a][b] += d2 / (xi[a] + DBL_EPSILON);  // y1 = row j, y2 = row i
                        }
                    }
                }
        }
    }

    double xx[2][2], yy[2][2], xy[2][2];
    if (want_corr) {
        double mx[2][2], my[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                mx[a][b] = sx[a][b] / cnt[a][b];
                my[a][b] = sy[a][b] / cnt[a][b];
                xx[a][b] = yy[a][b] = xy[a][b] = 0.0;
            }
        for (int64_t k0 = 0; k0 < c; k0 += DK) {
            __syncthreads();
            stage(k0);
            __syncthreads();
#pragma unroll 4
            for (int kk = 0; kk < DK; ++kk) {
                double xi[2] = {sI[kk][2 * tx], sI[kk][2 * tx + 1]};
                double xj[2] = {sJ[kk][2 * ty], sJ[kk][2 * ty + 1]};
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        if (usable(xi[a]) && usable(xj[b])) {
                            const double u = xi[a] - mx[a][b], v = xj[b] - my[a][b];
                            xx[a][b] += u * u;
                            yy[a][b] += v * v;
                            xy[a][b] += u * v;
                        }
                    }
            }
        }
    }

#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t i = i0 + 2 * tx + a, j = j0 + 2 * ty + b;
            if (i >= r || j >= r) continue;
            const int64_t o = i + j * out.ld;
            const double N = cnt[a][b];
            const bool ok = N >= 2.0;  // genomes.jl:545-547, :611-613
            const bool swapped = upper_rule && i > j;
            if (out.counts) out.counts[o] = (ok && !swapped) ? N : 0.0;
            if (out.euclid) out.euclid[o] = ok ? sqrt(ss[a][b]) : INFINITY;
            if (out.mad) out.mad[o] = ok ? sa[a][b] / N : INFINITY;
            if (out.rmsd) out.rmsd[o] = ok ? sqrt(ss[a][b] / N) : INFINITY;
            if (out.chi2) out.chi2[o] = ok ? (swapped ? c21[a][b] : c12[a][b]) : INFINITY;
            if (out.corr) {
                double v = -INFINITY;
                if (ok) {
                    const double vx = xx[a][b] / (N - 1.0), vy = yy[a][b] / (N - 1.0);  // var(): corrected
                    if (!(vx < 1e-7) && !(vy < 1e-7)) {
                        const double hi = fmax(xx[a][b], yy[a][b]), lo = fmin(xx[a][b], yy[a][b]);
                        v = xy[a][b] / hi / sqrt(lo / hi);  // Statistics.corm
                        v = fmin(1.0, fmax(-1.0, v));       // clampcor
                    }
                }
                out.corr[o] = v;
            }
        }
}

}  // namespace

extern "C" int32_t grm_cuda_distances(grm_cuda_ctx *ctx, const double *A, int64_t n, int64_t p, int64_t lda,
                                      int32_t a_location, const int64_t *idx_loci, int64_t t, int32_t dimension,
                                      double *euclidean, double *correlation, double *mad, double *rmsd, double *chi2,
                                      double *counts, int64_t ldo, int32_t out_location) {
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
    pair_kernel<<<dim3(static_cast<unsigned>(tiles), static_cast<unsigned>(tiles)), 256, 0, st>>>(
        S, r, entries ? t : n, entries ? 1 : n, entries ? n : 1, entries ? 0 : 1, correlation != nullptr, chi2 != nullptr,
        out);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    if (out_location == GRM_CUDA_LOC_HOST) {
        for (int m = 0; m < 6; ++m)
            if (host_dst[m])
                GRM_CUDA_TRY(cudaMemcpy2DAsync(host_dst[m], ldo * sizeof(double), dev_dst[m], r * sizeof(double),
                                               r * sizeof(double), static_cast<size_t>(r), cudaMemcpyDeviceToHost, st));
    }
    GRM_CUDA_TRY(cudaStreamSynchronize(st));
    return GRM_CUDA_OK;
}
