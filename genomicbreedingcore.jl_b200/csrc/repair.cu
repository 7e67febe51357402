// repair.cu — inflatediagonals! (src/grm/calc.jl:41-70) on the device.
//
// The reference decides with det(X) (LU with partial pivoting, then the SEQUENTIAL product of U's diagonal, which
// under/overflows for n >~ 500) and isposdef(X) (Cholesky): it returns early iff det > eps && isposdef (calc.jl:49) and
// inflates again iff det < eps || !isposdef (calc.jl:61).  Both tests are pure functions of the matrix, and the outcome
// is a function of (isposdef, det if isposdef): a matrix that is not positive definite is inflated whatever its
// determinant.  So the Cholesky goes first, and when it succeeds it already holds the determinant's pivots: for a
// positive definite X whose factor satisfies |L_jk| < L_kk below the diagonal, elimination with partial pivoting never
// interchanges rows and its pivots are u_k = L_kk^2 (repair_host.h).  The n diagonal entries come back to the host and are
// multiplied in pivot order in plain binary64 (gradual underflow, saturation at 0 / Inf: SURVEY.md A.4) exactly as the
// LU's pivots would be.  The real LU (cuSOLVER Dgetrf) still runs whenever that shortcut cannot vouch for the decision:
// a column where partial pivoting might interchange, an ill-determined pivot, or a product that ends anywhere near eps
// (grm_det_chol_pivots).  Tuning "repair_det" = 1 restores the literal transcription (LU first, Cholesky only when
// !(det < eps)); the tests run every case both ways and require identical rounds, totals and matrices.
// Measured at n = 20,000 on a B200 (scripts/potrf_probe.cu): Dgetrf 279 ms, Dpotrf LOWER 87 ms (UPPER 117 ms), so the
// factorisations work on the lower triangle; X is symmetric (checked first, calc.jl:43), and the factor is handed out
// transposed, as the U of Julia's cholesky(X).
// This step is O(n^3) FP64 library work and is reported separately from the GRM itself.
#include <cusolverDn.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>

#include "common.cuh"
#include "repair_host.h"

#define GRM_SOLVER_TRY(expr)                                                                  \
    do {                                                                                      \
        cusolverStatus_t _s = (expr);                                                         \
        if (_s != CUSOLVER_STATUS_SUCCESS) {                                                  \
            grm_set_error("%s failed at %s:%d: cusolver status %d", #expr, __FILE__, __LINE__, (int)_s); \
            return (_s == CUSOLVER_STATUS_ALLOC_FAILED) ? GRM_CUDA_ERR_OOM : GRM_CUDA_ERR_CUDA; \
        }                                                                                     \
    } while (0)

namespace {

// out[0] |= 1 not symmetric (X[i,j] != X[j,i], NaN counts: calc.jl:43 uses issymmetric), 2 NaN, 4 Inf;
// out64[1] = bit pattern of max |X| (non-negative doubles order like unsigned integers)
__global__ void __launch_bounds__(256) scan_kernel(const double *__restrict__ X, int64_t n, int64_t ldx,
                                                   unsigned long long *__restrict__ out) {
    unsigned flags = 0;
    unsigned long long mx = 0;
    for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
            const double a = X[i + j * ldx];
            if (i <= j) {
                const double b = X[j + i * ldx];
                if (!(a == b)) flags |= 1u;
            }
            if (a != a) flags |= 2u;
            else if (isinf(a)) flags |= 4u;
            else {
                const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(fabs(a)));
                mx = bits > mx ? bits : mx;
            }
        }
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, mx, o);
        mx = other > mx ? other : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (flags) atomicOr(&out[0], static_cast<unsigned long long>(flags));
        if (mx) atomicMax(&out[1], mx);
    }
}

__global__ void copy_matrix_kernel(const double *__restrict__ X, int64_t n, int64_t ldx, double *__restrict__ Y) {
    for (int64_t j = blockIdx.y; j < n; j += gridDim.y)
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x)
            Y[i + j * n] = X[i + j * ldx];
}

// columns per launch: gridDim.y is limited to 65535
unsigned grid_cols(int64_t n) { return static_cast<unsigned>(std::min<int64_t>(n, 65535)); }

__global__ void diag_extract_kernel(const double *__restrict__ LU, int64_t n, double *__restrict__ d) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) d[i] = LU[i + i * n];
}

// X[i,i] after `rounds` rounds of the reference's loop: ((x + e_0) + e_1) + ... with e_0 = eps0, e_{t+1} = fl(e_t (1+eps))
// — the sequence of roundings of calc.jl:63-65
__global__ void diag_add_rounds_kernel(double *__restrict__ X, int64_t n, int64_t ldx, double eps0, int rounds) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        double d = X[i + i * ldx], e = eps0;
        for (int t = 0; t < rounds; ++t) {
            d = __dadd_rn(d, e);
            e = __dmul_rn(e, 1.0 + DBL_EPSILON);
        }
        X[i + i * ldx] = d;
    }
}

// After potrf LOWER: diag[k] = L_kk, and *risk |= 1 unless every column satisfies |L_jk| <= L_kk (1 - 2^-20) for j > k,
// i.e. unless partial pivoting provably keeps the diagonal entry as the pivot of every elimination step (a near-tie
// could go either way in a real LU, so it counts as a risk; so does anything that is not a positive finite number).
__global__ void __launch_bounds__(256) chol_check_kernel(const double *__restrict__ F, int64_t n, double *__restrict__ diag,
                                                         unsigned *__restrict__ risk) {
    for (int64_t k = blockIdx.x; k < n; k += gridDim.x) {
        const double *col = F + k * n;
        const double lkk = col[k];
        const double thr = lkk * (1.0 - 0x1p-20);
        int bad = !(lkk > 0.0) || !(lkk <= DBL_MAX);
        for (int64_t j = k + 1 + threadIdx.x; j < n; j += blockDim.x)
            if (!(fabs(col[j]) <= thr)) bad = 1;
        bad = __syncthreads_or(bad);
        if (threadIdx.x == 0) {
            diag[k] = lkk;
            if (bad) atomicOr(risk, 1u);
        }
    }
}

// F[j + i n] = X[i + j ldx] for i <= j: the lower triangle of F becomes the transpose of the UPPER triangle of X (the only
// part of X that is read, like Julia's cholesky(Symmetric(X))); the strict upper triangle of F is left alone — potrf
// LOWER never looks at it.  Block (32, 8), one 32 x 32 tile of X per block.
__global__ void upper_to_lower_kernel(const double *__restrict__ X, int64_t n, int64_t ldx, double *__restrict__ F) {
    __shared__ double tile[32][33];
    const int64_t bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;
    for (int c = threadIdx.y; c < 32; c += 8) {
        const int64_t i = bi * 32 + threadIdx.x, j = bj * 32 + c;
        if (i < n && j < n) tile[c][threadIdx.x] = X[i + j * ldx];
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += 8) {
        const int64_t j = bj * 32 + threadIdx.x, i = bi * 32 + c;
        if (i < n && j < n && i <= j) F[j + i * n] = tile[threadIdx.x][c];
    }
}

// U[i + j ldu] = L[j,i] = F[j + i n] for i <= j, 0 below the diagonal: the lower factor handed out as Julia's cholesky(X).U
__global__ void lower_to_upper_kernel(const double *__restrict__ F, int64_t n, double *__restrict__ U, int64_t ldu) {
    __shared__ double tile[32][33];
    const int64_t bi = blockIdx.x, bj = blockIdx.y;  // tile of U: rows bi, columns bj
    if (bi <= bj) {
        for (int c = threadIdx.y; c < 32; c += 8) {
            const int64_t j = bj * 32 + threadIdx.x, i = bi * 32 + c;
            if (i < n && j < n) tile[c][threadIdx.x] = F[j + i * n];
        }
    }
    __syncthreads();
    for (int c = threadIdx.y; c < 32; c += 8) {
        const int64_t i = bi * 32 + threadIdx.x, j = bj * 32 + c;
        if (i < n && j < n) U[i + j * ldu] = (i <= j) ? tile[threadIdx.x][c] : 0.0;
    }
}

double bits_to_double(unsigned long long bits) {
    double d;
    memcpy(&d, &bits, sizeof(double));
    return d;
}

// cuSOLVER's legacy Dpotrf / Dgetrf index with 32-bit integers (n * n has to stay below 2^31, workspace lengths are ints);
// beyond n = 46,000 the 64-bit entry points take over.  The legacy ones stay for the sizes they cover: that is where the
// measurements were taken (scripts/potrf_probe.cu).
constexpr int64_t LEGACY_MAX_N = 46000;

int32_t solver_params(grm_cuda_ctx *ctx, cusolverDnParams_t *out) {
    if (!ctx->cusolver_params) {
        cusolverDnParams_t pr;
        GRM_SOLVER_TRY(cusolverDnCreateParams(&pr));
        ctx->cusolver_params = pr;
    }
    *out = static_cast<cusolverDnParams_t>(ctx->cusolver_params);
    return GRM_CUDA_OK;
}

struct FactorPlan {
    bool big = false;
    cusolverDnParams_t params = nullptr;
    size_t dev_bytes = 8, host_bytes = 0;  // workspaces
};

int32_t potrf_plan(grm_cuda_ctx *ctx, cusolverDnHandle_t h, int64_t n, double *A, FactorPlan *pl) {
    pl->big = n > LEGACY_MAX_N;
    if (pl->big) {
        GRM_TRY(solver_params(ctx, &pl->params));
        GRM_SOLVER_TRY(cusolverDnXpotrf_bufferSize(h, pl->params, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F,
                                                   &pl->dev_bytes, &pl->host_bytes));
        pl->dev_bytes = std::max<size_t>(pl->dev_bytes, 8);
    } else {
        int lw = 0;
        GRM_SOLVER_TRY(cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, static_cast<int>(n), A, static_cast<int>(n), &lw));
        pl->dev_bytes = static_cast<size_t>(std::max(lw, 1)) * sizeof(double);
        pl->host_bytes = 0;
    }
    return GRM_CUDA_OK;
}

// lower Cholesky factor in place, asynchronous on the handle's stream; *d_info as LAPACK's info
int32_t potrf_run(cusolverDnHandle_t h, const FactorPlan &pl, int64_t n, double *A, void *d_work, void *h_work, int *d_info) {
    if (pl.big)
        GRM_SOLVER_TRY(cusolverDnXpotrf(h, pl.params, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, d_work,
                                        pl.dev_bytes, h_work, pl.host_bytes, d_info));
    else
        GRM_SOLVER_TRY(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_LOWER, static_cast<int>(n), A, static_cast<int>(n),
                                        static_cast<double *>(d_work), static_cast<int>(pl.dev_bytes / sizeof(double)), d_info));
    return GRM_CUDA_OK;
}

int32_t getrf_plan(grm_cuda_ctx *ctx, cusolverDnHandle_t h, int64_t n, double *A, FactorPlan *pl) {
    pl->big = n > LEGACY_MAX_N;
    if (pl->big) {
        GRM_TRY(solver_params(ctx, &pl->params));
        GRM_SOLVER_TRY(cusolverDnXgetrf_bufferSize(h, pl->params, n, n, CUDA_R_64F, A, n, CUDA_R_64F, &pl->dev_bytes,
                                                   &pl->host_bytes));
        pl->dev_bytes = std::max<size_t>(pl->dev_bytes, 8);
    } else {
        int lw = 0;
        GRM_SOLVER_TRY(cusolverDnDgetrf_bufferSize(h, static_cast<int>(n), static_cast<int>(n), A, static_cast<int>(n), &lw));
        pl->dev_bytes = static_cast<size_t>(std::max(lw, 1)) * sizeof(double);
        pl->host_bytes = 0;
    }
    return GRM_CUDA_OK;
}

// One factorisation lane: a cuSOLVER handle on its own stream, an n x n work matrix, and pinned slots for what comes back.
struct Lane {
    cusolverDnHandle_t h = nullptr;
    cudaStream_t stream = nullptr;
    double *mat = nullptr;   // copy of the (inflated) matrix, then its lower Cholesky factor
    void *work = nullptr;
    FactorPlan plan;
    std::vector<char> h_work;  // host workspace of the 64-bit entry point (usually empty)
    int *d_info = nullptr;
    unsigned *d_risk = nullptr;
    double *d_diag = nullptr;
    int *h_info = nullptr;  // pinned: [info][risk][pad][diag: n doubles]
    unsigned *h_risk = nullptr;
    double *h_diag = nullptr;
    int buf = 0;  // 0: mat is ctx->lu, 1: ctx->lu2 (what grm_cuda_last_factor reads)
};

// The tests of the reference's loop on X after `rounds` rounds of inflation; X itself stays untouched until the outcome
// is known (grm_cuda_repair_apply).
struct Engine {
    grm_cuda_ctx *ctx = nullptr;
    int64_t n = 0;
    int nlanes = 1;
    Lane L[2];
    // the LU (fallback, or tuning repair_det = 1): set up on first use
    bool lu_ready = false;
    double *lu_mat = nullptr, *lu_diag = nullptr;
    void *lu_work = nullptr, *lu_ipiv = nullptr;  // ipiv: n ints (legacy) or n int64 (64-bit entry point)
    int *lu_info = nullptr;
    FactorPlan lu_plan;
    std::vector<char> lu_h_work;
    std::vector<double> h_diag;
    std::vector<int> h_ipiv;
    std::vector<int64_t> h_ipiv64;

    int32_t setup(grm_cuda_ctx *c, int64_t n_, int lanes) {
        ctx = c;
        n = n_;
        nlanes = lanes;
        if (!ctx->cusolver) {
            cusolverDnHandle_t h;
            GRM_SOLVER_TRY(cusolverDnCreate(&h));
            ctx->cusolver = h;
        }
        GRM_SOLVER_TRY(cusolverDnSetStream(static_cast<cusolverDnHandle_t>(ctx->cusolver), ctx->stream));
        if (nlanes > 1 && !ctx->cusolver_lane[0]) {
            cusolverDnHandle_t h;
            GRM_SOLVER_TRY(cusolverDnCreate(&h));
            ctx->cusolver_lane[0] = h;
            int lo = 0, hi = 0;
            GRM_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            GRM_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->lane_stream[0], cudaStreamNonBlocking, hi));
            GRM_SOLVER_TRY(cusolverDnSetStream(h, ctx->lane_stream[0]));
        }
        if (nlanes > 1 && !ctx->lane_ev[0]) GRM_CUDA_TRY(cudaEventCreateWithFlags(&ctx->lane_ev[0], cudaEventDisableTiming));
        const size_t mat = static_cast<size_t>(n) * n * sizeof(double);
        GRM_TRY(ctx->lu.reserve(mat));
        if (nlanes > 1) GRM_TRY(ctx->lu2.reserve(mat));
        L[0].h = static_cast<cusolverDnHandle_t>(ctx->cusolver);
        L[0].stream = ctx->stream;
        L[0].mat = ctx->lu.as<double>();
        L[0].buf = 0;
        if (nlanes > 1) {
            L[1].h = static_cast<cusolverDnHandle_t>(ctx->cusolver_lane[0]);
            L[1].stream = ctx->lane_stream[0];
            L[1].mat = ctx->lu2.as<double>();
            L[1].buf = 1;
        }
        FactorPlan plan;
        GRM_TRY(potrf_plan(ctx, L[0].h, n, L[0].mat, &plan));
        // per lane: info | risk | pad to 16 bytes | diag (n doubles); the same layout on the device and in pinned memory
        const size_t per = (16 + static_cast<size_t>(n) * sizeof(double) + 63) & ~size_t(63);
        if (ctx->lane_pin_cap < 2 * per) {
            if (ctx->lane_pin) cudaFreeHost(ctx->lane_pin);
            ctx->lane_pin = nullptr;
            ctx->lane_pin_cap = 0;
            GRM_CUDA_TRY(cudaHostAlloc(&ctx->lane_pin, 2 * per, cudaHostAllocDefault));
            ctx->lane_pin_cap = 2 * per;
        }
        DevBuf *small[2] = {&ctx->ipiv, &ctx->lane_ipiv[0]};
        DevBuf *works[2] = {&ctx->lu_work, &ctx->lane_work[0]};
        for (int i = 0; i < nlanes; ++i) {
            GRM_TRY(works[i]->reserve(plan.dev_bytes));
            GRM_TRY(small[i]->reserve(per));
            L[i].work = works[i]->ptr;
            L[i].plan = plan;
            L[i].h_work.resize(plan.host_bytes);
            char *dp = small[i]->as<char>();
            L[i].d_info = reinterpret_cast<int *>(dp);
            L[i].d_risk = reinterpret_cast<unsigned *>(dp + 4);
            L[i].d_diag = reinterpret_cast<double *>(dp + 16);
            char *hp = static_cast<char *>(ctx->lane_pin) + static_cast<size_t>(i) * per;
            L[i].h_info = reinterpret_cast<int *>(hp);
            L[i].h_risk = reinterpret_cast<unsigned *>(hp + 4);
            L[i].h_diag = reinterpret_cast<double *>(hp + 16);
        }
        return GRM_CUDA_OK;
    }

    int32_t copy_in(cudaStream_t s, const double *X, int64_t ldx, double *dst, int rounds, double eps0) {
        const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 64));
        copy_matrix_kernel<<<dim3(gx, grid_cols(n)), 256, 0, s>>>(X, n, ldx, dst);
        GRM_CUDA_TRY(cudaGetLastError());
        if (rounds > 0) {
            diag_add_rounds_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(dst, n, n, eps0, rounds);
            GRM_CUDA_TRY(cudaGetLastError());
        }
        return GRM_CUDA_OK;
    }

    // Cholesky of X after `rounds` rounds + the pivoting check, all asynchronous on the lane's stream.  Does not touch
    // ctx->launches (a lane may be driven by a side thread); the caller adds chol_launches(rounds).
    int32_t launch_chol(int lane, const double *X, int64_t ldx, int rounds, double eps0) {
        Lane &l = L[lane];
        GRM_TRY(copy_in(l.stream, X, ldx, l.mat, rounds, eps0));
        GRM_TRY(potrf_run(l.h, l.plan, n, l.mat, l.work, l.h_work.data(), l.d_info));
        GRM_CUDA_TRY(cudaMemsetAsync(l.d_risk, 0, sizeof(unsigned), l.stream));
        const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(n, 8 * 148));
        chol_check_kernel<<<blocks, 256, 0, l.stream>>>(l.mat, n, l.d_diag, l.d_risk);
        GRM_CUDA_TRY(cudaGetLastError());
        // info, risk and the diagonal are contiguous on both sides
        GRM_CUDA_TRY(cudaMemcpyAsync(l.h_info, l.d_info, 16 + static_cast<size_t>(n) * sizeof(double), cudaMemcpyDeviceToHost,
                                     l.stream));
        return GRM_CUDA_OK;
    }
    static int chol_launches(int rounds) { return rounds > 0 ? 3 : 2; }

    // Julia's det(lu(X; check = false)) of X after `rounds` rounds, on the context's stream
    int32_t det_lu(const double *X, int64_t ldx, int rounds, double eps0, double *out) {
        cusolverDnHandle_t h = static_cast<cusolverDnHandle_t>(ctx->cusolver);
        if (!lu_ready) {
            GRM_TRY(ctx->lu3.reserve(static_cast<size_t>(n) * n * sizeof(double)));
            lu_mat = ctx->lu3.as<double>();
            GRM_TRY(getrf_plan(ctx, h, n, lu_mat, &lu_plan));
            GRM_TRY(ctx->lane_work[1].reserve(lu_plan.dev_bytes));
            lu_work = ctx->lane_work[1].ptr;
            lu_h_work.resize(lu_plan.host_bytes);
            // ipiv (n ints, or n int64 for the 64-bit entry point) | info | pad to 16 bytes | diag(U) (n doubles)
            const size_t piv = static_cast<size_t>(n) * (lu_plan.big ? sizeof(int64_t) : sizeof(int));
            const size_t head = (piv + sizeof(int) + 15) & ~size_t(15);
            GRM_TRY(ctx->lane_ipiv[1].reserve(head + static_cast<size_t>(n) * sizeof(double)));
            lu_ipiv = ctx->lane_ipiv[1].ptr;
            lu_info = reinterpret_cast<int *>(ctx->lane_ipiv[1].as<char>() + piv);
            lu_diag = reinterpret_cast<double *>(ctx->lane_ipiv[1].as<char>() + head);
            h_diag.resize(n);
            if (lu_plan.big) h_ipiv64.resize(n);
            else h_ipiv.resize(n);
            lu_ready = true;
        }
        GRM_TRY(copy_in(ctx->stream, X, ldx, lu_mat, rounds, eps0));
        if (lu_plan.big)
            GRM_SOLVER_TRY(cusolverDnXgetrf(h, lu_plan.params, n, n, CUDA_R_64F, lu_mat, n, static_cast<int64_t *>(lu_ipiv),
                                            CUDA_R_64F, lu_work, lu_plan.dev_bytes, lu_h_work.data(), lu_plan.host_bytes, lu_info));
        else
            GRM_SOLVER_TRY(cusolverDnDgetrf(h, static_cast<int>(n), static_cast<int>(n), lu_mat, static_cast<int>(n),
                                            static_cast<double *>(lu_work), static_cast<int *>(lu_ipiv), lu_info));
        diag_extract_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(lu_mat, n, lu_diag);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches += 2 + (rounds > 0 ? 1 : 0);
        ctx->repair_lu_count++;
        int h_info = 0;
        GRM_CUDA_TRY(cudaMemcpyAsync(h_diag.data(), lu_diag, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (lu_plan.big)
            GRM_CUDA_TRY(cudaMemcpyAsync(h_ipiv64.data(), lu_ipiv, n * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        else
            GRM_CUDA_TRY(cudaMemcpyAsync(h_ipiv.data(), lu_ipiv, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaMemcpyAsync(&h_info, lu_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (h_info < 0) {
            grm_set_error("cusolver getrf: illegal argument %d", -h_info);
            return GRM_CUDA_ERR_CUDA;
        }
        if (lu_plan.big) {  // the interchanges as the 32-bit flags grm_det_lu_pivots counts: ipiv[i] != i + 1
            h_ipiv.resize(n);
            for (int64_t i = 0; i < n; ++i) h_ipiv[i] = h_ipiv64[i] == i + 1 ? static_cast<int>(i + 1) : 0;
        }
        // an exactly zero pivot: Julia's det returns 0 for a singular factorisation
        *out = h_info > 0 ? 0.0 : grm_det_lu_pivots(h_diag.data(), h_ipiv.data(), n);
        return GRM_CUDA_OK;
    }

    // Outcome of a launched lane: pd = isposdef; det = the determinant when pd (NaN otherwise: a matrix that is not
    // positive definite is inflated whatever its determinant, calc.jl:49,61).  Falls back to the LU when the Cholesky's
    // pivots cannot vouch for `det < eps`.
    int32_t finish_chol(int lane, const double *X, int64_t ldx, int rounds, double eps0, bool *pd, double *det) {
        const Lane &l = L[lane];
        GRM_CUDA_TRY(cudaStreamSynchronize(l.stream));
        *pd = false;
        *det = NAN;
        if (*l.h_info < 0) {
            grm_set_error("cusolverDnDpotrf: illegal argument %d", -*l.h_info);
            return GRM_CUDA_ERR_CUDA;
        }
        if (*l.h_info > 0) return GRM_CUDA_OK;
        *pd = true;
        // max|X| of the inflated matrix: for a positive definite matrix it sits on the diagonal
        const double scale = eps0 * (1.0 + static_cast<double>(rounds));
        if (*l.h_risk == 0 && grm_det_chol_pivots(l.h_diag, n, scale, det) == GRM_CHOLDET_OK) return GRM_CUDA_OK;
        return det_lu(X, ldx, rounds, eps0, det);
    }

    // one round, sequentially, on lane 0
    int32_t evaluate(const double *X, int64_t ldx, int rounds, double eps0, bool *pd, double *det) {
        if (ctx->tune.repair_det == 1) {
            // the literal order of calc.jl:49,61: det first, isposdef only when !(det < eps)
            GRM_TRY(det_lu(X, ldx, rounds, eps0, det));
            *pd = false;
            if (*det < DBL_EPSILON) return GRM_CUDA_OK;
            GRM_TRY(launch_chol(0, X, ldx, rounds, eps0));
            ctx->launches += chol_launches(rounds);
            GRM_CUDA_TRY(cudaStreamSynchronize(L[0].stream));
            if (*L[0].h_info < 0) {
                grm_set_error("cusolverDnDpotrf: illegal argument %d", -*L[0].h_info);
                return GRM_CUDA_ERR_CUDA;
            }
            *pd = *L[0].h_info == 0;
            return GRM_CUDA_OK;
        }
        GRM_TRY(launch_chol(0, X, ldx, rounds, eps0));
        ctx->launches += chol_launches(rounds);
        return finish_chol(0, X, ldx, rounds, eps0, pd, det);
    }

    // Rounds t and t + 1 at once.  The loop test of round t only looks at X with its diagonal inflated t times, so the
    // tests of consecutive rounds are independent (what the multi-GPU probes exploit across GPUs); here they run on two
    // streams with a handle, a work matrix and a workspace each, and the panel phases of one factorisation hide under
    // the trailing updates of the other.  Same routines on the same inputs as the sequential loop, hence the same values
    // and decisions.  Lane 1 is driven by a host thread of its own: cuSOLVER keeps its calling thread busy while it
    // issues a factorisation's launches.
    int32_t evaluate_pair(const double *X, int64_t ldx, int rounds, double eps0, bool both, bool *pd, double *det) {
        if (!both) return evaluate(X, ldx, rounds, eps0, &pd[0], &det[0]);
        GRM_CUDA_TRY(cudaEventRecord(ctx->lane_ev[0], ctx->stream));  // X is final on the context's stream
        GRM_CUDA_TRY(cudaStreamWaitEvent(L[1].stream, ctx->lane_ev[0], 0));
        int32_t st_side = GRM_CUDA_OK;
        std::string err_side;
        std::thread side([&]() {
            int32_t rc = cudaSetDevice(ctx->device) == cudaSuccess ? launch_chol(1, X, ldx, rounds + 1, eps0) : GRM_CUDA_ERR_CUDA;
            if (rc == GRM_CUDA_OK && cudaStreamSynchronize(L[1].stream) != cudaSuccess) rc = GRM_CUDA_ERR_CUDA;
            if (rc != GRM_CUDA_OK) err_side = grm_cuda_last_error();  // the message is thread-local
            st_side = rc;
        });
        int32_t rc0 = launch_chol(0, X, ldx, rounds, eps0);
        if (rc0 == GRM_CUDA_OK && cudaStreamSynchronize(L[0].stream) != cudaSuccess) {
            grm_set_error("cudaStreamSynchronize failed in the repair");
            rc0 = GRM_CUDA_ERR_CUDA;
        }
        side.join();
        ctx->launches += chol_launches(rounds) + chol_launches(rounds + 1);
        GRM_TRY(rc0);
        if (st_side != GRM_CUDA_OK) {
            grm_set_error("%s", err_side.empty() ? "repair lane failed" : err_side.c_str());
            return st_side;
        }
        GRM_TRY(finish_chol(0, X, ldx, rounds, eps0, &pd[0], &det[0]));
        return finish_chol(1, X, ldx, rounds + 1, eps0, &pd[1], &det[1]);
    }
};

// symmetry / NaN / Inf / max|X| scan (h_scan[0] = flags, h_scan[1] = bits of max|X|): calc.jl:43-48 and the first epsilon
int32_t repair_scan(grm_cuda_ctx *ctx, const double *X, int64_t n, int64_t ldx, unsigned long long *h_scan) {
    ctx->factor_n = -1;  // the work matrices are about to be overwritten
    GRM_TRY(ctx->misc.reserve(64));
    unsigned long long *d_scan = ctx->misc.as<unsigned long long>();
    GRM_CUDA_TRY(cudaMemsetAsync(d_scan, 0, 2 * sizeof(unsigned long long), ctx->stream));
    const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 64));
    scan_kernel<<<dim3(gx, grid_cols(n)), 256, 0, ctx->stream>>>(X, n, ldx, d_scan);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    GRM_CUDA_TRY(cudaMemcpyAsync(h_scan, d_scan, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GRM_CUDA_OK;
}

}  // namespace

int32_t grm_cusolver_destroy(void *h) {
    cusolverDnDestroy(static_cast<cusolverDnHandle_t>(h));
    return GRM_CUDA_OK;
}

int32_t grm_cusolver_params_destroy(void *p) {
    cusolverDnDestroyParams(static_cast<cusolverDnParams_t>(p));
    return GRM_CUDA_OK;
}

// device-resident implementation; X is a device pointer
int32_t grm_repair_device(grm_cuda_ctx *ctx, double *X, int64_t n, int64_t ldx, int32_t max_iter, int32_t *iters,
                          double *eps_total) {
    *iters = 0;
    *eps_total = 0.0;
    unsigned long long h_scan[2];
    GRM_TRY(repair_scan(ctx, X, n, ldx, h_scan));
    if (h_scan[0] & 1ull) {
        grm_set_error("The input matrix is not symmetric.");
        return GRM_CUDA_ERR_NOT_SYMMETRIC;
    }
    const double eps = DBL_EPSILON;
    const double maxabs = bits_to_double(h_scan[1]);
    // two rounds in flight from n = 3072 (below, a factorisation is too short to be worth a second stream and thread)
    // (and only while three work matrices are small next to the 180 GB: the 64-bit range runs one round at a time)
    const bool pairs = ctx->tune.repair != 1 && ctx->tune.repair_det != 1 && n >= 3072 && n <= LEGACY_MAX_N;
    Engine e;
    GRM_TRY(e.setup(ctx, n, pairs ? 2 : 1));
    int64_t b = -1;  // first round of the batch whose results are held
    bool pd_b[2] = {false, false};
    double det_b[2] = {NAN, NAN};
    bool have_second = false;
    int lane_t = 0;  // lane that evaluated the round being looked at
    auto eval = [&](int64_t t, bool *pd, double *det) -> int32_t {
        if (b < 0 || t < b || t > b + (have_second ? 1 : 0)) {
            have_second = pairs && t + 1 <= max_iter;  // round t + 1 can only matter if the loop may reach it
            GRM_TRY(e.evaluate_pair(X, ldx, static_cast<int>(t), maxabs, have_second, pd_b, det_b));
            b = t;
        }
        lane_t = static_cast<int>(t - b);
        *pd = pd_b[lane_t];
        *det = det_b[lane_t];
        return GRM_CUDA_OK;
    };
    auto keep_factor = [&](bool pd) {
        // the Cholesky factor of the matrix being returned sits in the work matrix of the lane that evaluated its round
        if (pd) {
            ctx->factor_n = n;
            ctx->factor_buf = e.L[lane_t].buf;
        }
    };

    // calc.jl:49-51 — early return iff det(X) > eps && isposdef(X)
    bool pd = false;
    double det = NAN;
    GRM_TRY(eval(0, &pd, &det));
    if (pd && det > eps) {
        keep_factor(true);
        return GRM_CUDA_OK;
    }
    // calc.jl:52-57
    if (h_scan[0] & 2ull) {
        grm_set_error("The input matrix contains NaN values.");
        return GRM_CUDA_ERR_NAN_MATRIX;
    }
    if (h_scan[0] & 4ull) {
        grm_set_error("The input matrix contains Inf values.");
        return GRM_CUDA_ERR_INF_MATRIX;
    }
    // calc.jl:58-66.  The first evaluation of the loop condition looks at the SAME matrix as the early test, so its
    // results are reused (same input, same factorisations, same values).
    int64_t t = 0;
    for (;;) {
        const bool need = det < eps || !pd;  // calc.jl:61 (a NaN det compares false, then isposdef decides)
        if (!need || t >= max_iter) {
            if (!need) keep_factor(pd);
            break;
        }
        ++t;
        GRM_TRY(eval(t, &pd, &det));
    }
    *iters = static_cast<int32_t>(t);
    return grm_cuda_repair_apply(ctx, X, n, ldx, *iters, maxabs, eps_total);
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU inflatediagonals!: the loop tests of different rounds are independent given max|X| (round t looks at X
// with its diagonal inflated t times), so with the matrix replicated on every GPU, rank r evaluates round r: the
// wall time of the repair becomes one round's factorisation instead of rounds + 1 of them.  The arithmetic of each
// probe is the arithmetic of the single-GPU loop, so the outcome is identical.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int32_t grm_cuda_repair_probe(grm_cuda_ctx *ctx, const double *dX, int64_t n, int64_t ldx, int32_t rounds,
                                         double *det_out, int32_t *posdef_out, uint32_t *scan_flags, double *maxabs) try {
    if (!ctx || !dX || n < 1 || ldx < n || rounds < 0 || !det_out || !posdef_out || !scan_flags || !maxabs) {
        grm_set_error("repair_probe: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned long long h_scan[2];
    GRM_TRY(repair_scan(ctx, dX, n, ldx, h_scan));
    *scan_flags = static_cast<uint32_t>(h_scan[0]);
    const double mx = bits_to_double(h_scan[1]);
    *maxabs = mx;
    *det_out = NAN;
    *posdef_out = -1;
    if (h_scan[0] & 1ull) return GRM_CUDA_OK;  // not symmetric: the caller raises before any factorisation (calc.jl:43)
    Engine e;
    GRM_TRY(e.setup(ctx, n, 1));
    bool pd = false;
    GRM_TRY(e.evaluate(dX, ldx, rounds, mx, &pd, det_out));
    // -1: isposdef not evaluated because det < eps already decides (only with tuning repair_det = 1)
    *posdef_out = (ctx->tune.repair_det == 1 && *det_out < DBL_EPSILON) ? -1 : (pd ? 1 : 0);
    ctx->factor_buf = 0;  // a caller that finds this probe to be the deciding one sets ctx->factor_n
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_repair_apply(grm_cuda_ctx *ctx, double *dX, int64_t n, int64_t ldx, int32_t rounds,
                                         double maxabs, double *eps_total) try {
    if (!ctx || !dX || n < 1 || ldx < n || rounds < 0) {
        grm_set_error("repair_apply: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    if (rounds > 0) {
        diag_add_rounds_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(dX, n, ldx, maxabs, rounds);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
    }
    if (eps_total) {  // calc.jl:64-65 on the host: the same sequence of roundings
        volatile double e = maxabs, total = 0.0;
        for (int t = 0; t < rounds; ++t) {
            total = total + e;
            e = e * (1.0 + DBL_EPSILON);
        }
        *eps_total = total;
    }
    return GRM_CUDA_OK;
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_host_chol_det(const double *diag_l, int64_t n, double scale, double *det_out) try {
    if (!diag_l || n < 1 || !det_out) return -1;
    return grm_det_chol_pivots(diag_l, n, scale, det_out);
} GRM_CATCH_STATUS

extern "C" double grm_cuda_host_lu_det(const double *diag_u, const int32_t *ipiv, int64_t n) {
    if (!diag_u || !ipiv || n < 1) return NAN;
    return grm_det_lu_pivots(diag_u, ipiv, n);
}

// ---------------------------------------------------------------------------------------------------------------
// Consumer side of the GRM: simulateeffects builds MvNormal(mu, Sigma) from
// simulatecovariancekinship's matrix (src/simulations/simulate_effects.jl:344-351, :479-481), i.e. PDMat(Sigma) ->
// cholesky(Sigma) = LAPACK potrf!('U', copy(Sigma)).  The repair above ends with a Cholesky factorisation of the matrix
// it returns, so the factor can be handed out instead of being recomputed (1/3 n^3 flop saved).
// ---------------------------------------------------------------------------------------------------------------
namespace {

int32_t emit_factor(grm_cuda_ctx *ctx, int64_t n, double *U, int64_t ldu, int32_t location) {
    const double *F = ctx->factor_buf ? ctx->lu2.as<double>() : ctx->lu.as<double>();
    double *dU = U;
    int64_t ld = ldu;
    if (location == GRM_CUDA_LOC_HOST) {
        GRM_TRY(ctx->misc2.reserve(static_cast<size_t>(n) * n * sizeof(double)));
        dU = ctx->misc2.as<double>();
        ld = n;
    }
    const unsigned nt = static_cast<unsigned>((n + 31) / 32);
    lower_to_upper_kernel<<<dim3(nt, nt), dim3(32, 8), 0, ctx->stream>>>(F, n, dU, ld);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    if (location == GRM_CUDA_LOC_HOST)
        GRM_CUDA_TRY(cudaMemcpy2DAsync(U, ldu * sizeof(double), dU, n * sizeof(double), n * sizeof(double),
                                       static_cast<size_t>(n), cudaMemcpyDeviceToHost, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GRM_CUDA_OK;
}

}  // namespace

extern "C" int32_t grm_cuda_last_factor(grm_cuda_ctx *ctx, int64_t n, double *U, int64_t ldu, int32_t location) try {
    if (!ctx || !U || n < 1 || ldu < n || (location != GRM_CUDA_LOC_HOST && location != GRM_CUDA_LOC_DEVICE)) {
        grm_set_error("last_factor: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (ctx->factor_n != n) {
        grm_set_error("last_factor: no Cholesky factor of a %lld x %lld matrix is cached on this context (the last repair "
                      "must have run, on this context, and ended positive definite)", (long long)n, (long long)n);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    return emit_factor(ctx, n, U, ldu, location);
} GRM_CATCH_STATUS

extern "C" int32_t grm_cuda_cholesky(grm_cuda_ctx *ctx, const double *X, int64_t n, int64_t ldx, int32_t x_location,
                                     double *U, int64_t ldu, int32_t u_location) try {
    if (!ctx || !X || !U || n < 1 || ldx < n || ldu < n ||
        (x_location != GRM_CUDA_LOC_HOST && x_location != GRM_CUDA_LOC_DEVICE) ||
        (u_location != GRM_CUDA_LOC_HOST && u_location != GRM_CUDA_LOC_DEVICE)) {
        grm_set_error("cholesky: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->cusolver) {
        cusolverDnHandle_t h;
        GRM_SOLVER_TRY(cusolverDnCreate(&h));
        ctx->cusolver = h;
    }
    cusolverDnHandle_t h = static_cast<cusolverDnHandle_t>(ctx->cusolver);
    GRM_SOLVER_TRY(cusolverDnSetStream(h, ctx->stream));
    ctx->factor_n = -1;
    GRM_TRY(ctx->lu.reserve(static_cast<size_t>(n) * n * sizeof(double)));
    double *F = ctx->lu.as<double>();
    // like Julia's cholesky(Symmetric(X)), only the upper triangle of X is read: it is transposed into the lower triangle
    // of the work matrix, where cuSOLVER's faster LOWER factorisation finds it (bit-identical to the factor the repair
    // leaves behind for the same, symmetric, matrix: same routine on the same lower triangle)
    const double *dX = X;
    int64_t ldd = ldx;
    if (x_location == GRM_CUDA_LOC_HOST) {
        GRM_TRY(ctx->misc2.reserve(static_cast<size_t>(n) * n * sizeof(double)));
        GRM_CUDA_TRY(cudaMemcpy2DAsync(ctx->misc2.as<double>(), n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double),
                                       static_cast<size_t>(n), cudaMemcpyHostToDevice, ctx->stream));
        dX = ctx->misc2.as<double>();
        ldd = n;
    }
    const unsigned nt = static_cast<unsigned>((n + 31) / 32);
    upper_to_lower_kernel<<<dim3(nt, nt), dim3(32, 8), 0, ctx->stream>>>(dX, n, ldd, F);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    FactorPlan plan;
    GRM_TRY(potrf_plan(ctx, h, n, F, &plan));
    GRM_TRY(ctx->lu_work.reserve(plan.dev_bytes));
    std::vector<char> h_work(plan.host_bytes);
    GRM_TRY(ctx->misc.reserve(64));
    int *info = ctx->misc.as<int>();
    GRM_TRY(potrf_run(h, plan, n, F, ctx->lu_work.ptr, h_work.data(), info));
    int h_info = 0;
    GRM_CUDA_TRY(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_info < 0) {
        grm_set_error("cusolverDnDpotrf: illegal argument %d", -h_info);
        return GRM_CUDA_ERR_CUDA;
    }
    if (h_info > 0) {
        // Julia: PosDefException(info) — "matrix is not positive definite; Factorization failed."
        grm_set_error("matrix is not positive definite; Cholesky factorization failed (leading minor of order %d).", h_info);
        return GRM_CUDA_ERR_NOT_POSDEF;
    }
    ctx->factor_n = n;
    ctx->factor_buf = 0;
    return emit_factor(ctx, n, U, ldu, u_location);
} GRM_CATCH_STATUS
