// repair.cu — inflatediagonals! (src/grm/calc.jl:41-70) on the device.
//
// The reference decides with det(X) (LU with partial pivoting, then the SEQUENTIAL product of U's diagonal, which
// under/overflows for n >~ 500) and isposdef(X) (Cholesky).  The decision therefore depends on IEEE under/overflow
// of a running product, so it is reproduced literally: cuSOLVER Dgetrf/Dpotrf factorise on the device, the n pivots
// come back to the host and are multiplied in pivot order in plain binary64 (gradual underflow, no log-domain
// shortcut) — SURVEY.md A.4.  This step is O(n^3) FP64 library work and is reported separately from the GRM itself.
//
// Order of evaluation: the reference's tests are pure functions of the matrix — early return iff det > eps && isposdef
// (calc.jl:49), another round iff det < eps || !isposdef (calc.jl:61) — so either factorisation may be skipped once the
// other has decided.  The LU goes first, as in the reference: for a large GRM it is the determinant test that fails
// (the product of n pivots under- or overflows: C3's matrix is numerically positive definite, its determinant
// underflows to 0), and det < eps then makes the Cholesky unnecessary.  Measured at n = 20,000 on a B200
// (scripts/potrf_probe.cu): Dgetrf 279 ms, Dpotrf 117 ms; Cholesky-first cost 817 ms per C3 repair against 689 ms.
#include <cusolverDn.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include <string>
#include <thread>

#include "common.cuh"

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
    const int64_t j = blockIdx.y;
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
    const int64_t j = blockIdx.y;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        Y[i + j * n] = X[i + j * ldx];
}

__global__ void diag_extract_kernel(const double *__restrict__ LU, int64_t n, double *__restrict__ d) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) d[i] = LU[i + i * n];
}

// X[i,i] = fl(X[i,i] + eps)   (calc.jl:63)
__global__ void diag_add_kernel(double *__restrict__ X, int64_t n, int64_t ldx, double eps) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) X[i + i * ldx] = X[i + i * ldx] + eps;
}

// X[i,i] after `rounds` rounds of the reference's loop: ((x + e_0) + e_1) + ... with e_0 = eps0, e_{t+1} = fl(e_t (1+eps))
// — the same sequence of roundings as `rounds` calls of diag_add_kernel (calc.jl:63-65)
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

struct Repair {
    grm_cuda_ctx *ctx;
    cusolverDnHandle_t h;
    double *X;
    int64_t n, ldx;
    double *work_chol;  // n x n copy the Cholesky destroys (ctx->lu: what grm_cuda_last_factor hands out)
    double *work_lu;    // n x n copy the LU destroys (its own buffer, so the LU never clobbers a cached factor)
    double *lwork;
    int lwork_len;
    int *ipiv, *info;
    double *d_diag;
    int shift_rounds = 0;  // probe mode: factorise X after this many rounds of inflation (X itself untouched)
    double shift_eps0 = 0.0;
    std::vector<double> h_diag;
    std::vector<int> h_ipiv;

    int32_t copy_in(double *dst) {
        const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 64));
        copy_matrix_kernel<<<dim3(gx, static_cast<unsigned>(n)), 256, 0, ctx->stream>>>(X, n, ldx, dst);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        if (shift_rounds > 0) {
            diag_add_rounds_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(dst, n, n, shift_eps0,
                                                                                                shift_rounds);
            GRM_CUDA_TRY(cudaGetLastError());
            ctx->launches++;
        }
        return GRM_CUDA_OK;
    }

    // Julia: det(X) = det(lu(X; check=false)): 0 if the factorisation hit an exact zero pivot, otherwise the
    // sequential product of diag(U) times (-1)^(number of row interchanges).
    int32_t det(double *out) {
        GRM_TRY(copy_in(work_lu));
        GRM_SOLVER_TRY(cusolverDnDgetrf(h, static_cast<int>(n), static_cast<int>(n), work_lu, static_cast<int>(n), lwork, ipiv,
                                        info));
        diag_extract_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(work_lu, n, d_diag);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        int h_info = 0;
        GRM_CUDA_TRY(cudaMemcpyAsync(h_diag.data(), d_diag, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaMemcpyAsync(h_ipiv.data(), ipiv, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (h_info < 0) {
            grm_set_error("cusolverDnDgetrf: illegal argument %d", -h_info);
            return GRM_CUDA_ERR_CUDA;
        }
        if (h_info > 0) {
            *out = 0.0;
            return GRM_CUDA_OK;
        }
        volatile double P = 1.0;  // volatile: one IEEE binary64 rounding per multiply, no reassociation
        int swaps = 0;
        for (int64_t i = 0; i < n; ++i) {
            P = P * h_diag[i];
            if (static_cast<int64_t>(h_ipiv[i]) != i + 1) ++swaps;
        }
        *out = (swaps & 1) ? -P : P;
        return GRM_CUDA_OK;
    }

    // Julia: isposdef(X) = ishermitian(X) && LAPACK.potrf!('U', copy(X)) succeeded (symmetry already checked)
    int32_t posdef(bool *out) {
        ctx->factor_n = -1;
        GRM_TRY(copy_in(work_chol));
        GRM_SOLVER_TRY(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_UPPER, static_cast<int>(n), work_chol, static_cast<int>(n), lwork,
                                        lwork_len, info));
        int h_info = 0;
        GRM_CUDA_TRY(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (h_info < 0) {
            grm_set_error("cusolverDnDpotrf: illegal argument %d", -h_info);
            return GRM_CUDA_ERR_CUDA;
        }
        *out = (h_info == 0);
        return GRM_CUDA_OK;
    }

    // one evaluation of the reference's pair of tests on the current matrix: dt = det(X); pd = isposdef(X) unless
    // det < eps already decides (then pd = false: not evaluated, the outcome does not depend on it)
    int32_t evaluate(bool *pd, double *dt) {
        GRM_TRY(det(dt));
        *pd = false;
        if (!(*dt < DBL_EPSILON)) GRM_TRY(posdef(pd));
        return GRM_CUDA_OK;
    }
};

}  // namespace

int32_t grm_cusolver_destroy(void *h) {
    cusolverDnDestroy(static_cast<cusolverDnHandle_t>(h));
    return GRM_CUDA_OK;
}

// cuSOLVER handle, workspaces and the symmetry / NaN / Inf / max|X| scan (h_scan[0] = flags, h_scan[1] = bits of max|X|)
static int32_t repair_setup(grm_cuda_ctx *ctx, double *X, int64_t n, int64_t ldx, Repair &r,
                            unsigned long long *h_scan) {
    if (n > 46000) {  // 32-bit indexing of the legacy potrf/getrf entry points: n*n must stay below 2^31
        grm_set_error("inflate_diagonals: n=%lld exceeds the single-GPU cuSOLVER range (n <= 46000)", (long long)n);
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (!ctx->cusolver) {
        cusolverDnHandle_t h;
        GRM_SOLVER_TRY(cusolverDnCreate(&h));
        ctx->cusolver = h;
    }
    cusolverDnHandle_t h = static_cast<cusolverDnHandle_t>(ctx->cusolver);
    GRM_SOLVER_TRY(cusolverDnSetStream(h, ctx->stream));
    ctx->factor_n = -1;  // the work matrix is about to be overwritten

    // calc.jl:43-48 — symmetry (squareness is implied by the C signature), plus the NaN/Inf/max|X| scan
    GRM_TRY(ctx->misc.reserve(64));
    unsigned long long *d_scan = ctx->misc.as<unsigned long long>();
    GRM_CUDA_TRY(cudaMemsetAsync(d_scan, 0, 2 * sizeof(unsigned long long), ctx->stream));
    {
        const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 64));
        scan_kernel<<<dim3(gx, static_cast<unsigned>(n)), 256, 0, ctx->stream>>>(X, n, ldx, d_scan);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
    }
    GRM_CUDA_TRY(cudaMemcpyAsync(h_scan, d_scan, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));

    r.ctx = ctx;
    r.h = h;
    r.X = X;
    r.n = n;
    r.ldx = ldx;
    GRM_TRY(ctx->lu.reserve(static_cast<size_t>(n) * n * sizeof(double)));
    GRM_TRY(ctx->lu2.reserve(static_cast<size_t>(n) * n * sizeof(double)));
    r.work_chol = ctx->lu.as<double>();
    r.work_lu = ctx->lu2.as<double>();
    int lw1 = 0, lw2 = 0;
    GRM_SOLVER_TRY(cusolverDnDgetrf_bufferSize(h, static_cast<int>(n), static_cast<int>(n), r.work_chol,
                                               static_cast<int>(n), &lw1));
    GRM_SOLVER_TRY(cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_UPPER, static_cast<int>(n), r.work_chol,
                                               static_cast<int>(n), &lw2));
    r.lwork_len = std::max(lw1, lw2);
    GRM_TRY(ctx->lu_work.reserve(std::max<size_t>(static_cast<size_t>(r.lwork_len), 1) * sizeof(double)));
    r.lwork = ctx->lu_work.as<double>();
    // ipiv (n ints) | info (1 int) | padding to 8 bytes | diag(U) (n doubles)
    const int64_t int_slots = (n + 16 + 1) & ~int64_t(1);
    GRM_TRY(ctx->ipiv.reserve(static_cast<size_t>(int_slots) * sizeof(int) + static_cast<size_t>(n) * sizeof(double)));
    r.ipiv = ctx->ipiv.as<int>();
    r.info = r.ipiv + n;
    r.d_diag = reinterpret_cast<double *>(r.ipiv + int_slots);
    r.h_diag.resize(n);
    r.h_ipiv.resize(n);
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Speculative repair on ONE GPU.  The loop test of round t only looks at X with its diagonal inflated t times, so the
// tests of rounds t and t + 1 are independent — what the multi-GPU probes exploit across GPUs.  cuSOLVER's getrf runs at
// about half of dgemm's rate (its panel phases leave the GPU mostly idle), so here det(t), det(t + 1) and the Cholesky of
// round t + 1 are put on three streams with a handle, a work matrix and a workspace each and evaluated CONCURRENTLY: the
// panels of one factorisation hide under the trailing updates of the others.  Same routines on the same inputs as the
// sequential loop, hence the same values and the same decisions; X itself stays untouched until the outcome is known.
// The Cholesky of round t is evaluated on demand (only when !(det_t < eps): for a large GRM the determinant is what fails).
// Results travel through pinned host memory: a copy into pageable memory would block the host and serialise the lanes.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Lane {
    cusolverDnHandle_t h = nullptr;
    cudaStream_t stream = nullptr;
    double *mat = nullptr, *work = nullptr;
    int lwork = 0;
    int *d_ipiv = nullptr, *d_info = nullptr;
    double *d_diag = nullptr;
    // pinned results
    double *h_diag = nullptr;
    int *h_ipiv = nullptr, *h_info = nullptr;
};

int32_t lane_copy_in(grm_cuda_ctx *ctx, const Lane &l, const double *X, int64_t n, int64_t ldx, int rounds, double eps0) {
    const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 64));
    copy_matrix_kernel<<<dim3(gx, static_cast<unsigned>(n)), 256, 0, l.stream>>>(X, n, ldx, l.mat);
    GRM_CUDA_TRY(cudaGetLastError());
    if (rounds > 0) {
        diag_add_rounds_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, l.stream>>>(l.mat, n, n, eps0, rounds);
        GRM_CUDA_TRY(cudaGetLastError());
        }
    return GRM_CUDA_OK;
}

int32_t lane_launch_det(grm_cuda_ctx *ctx, const Lane &l, const double *X, int64_t n, int64_t ldx, int rounds, double eps0) {
    GRM_TRY(lane_copy_in(ctx, l, X, n, ldx, rounds, eps0));
    GRM_SOLVER_TRY(cusolverDnDgetrf(l.h, static_cast<int>(n), static_cast<int>(n), l.mat, static_cast<int>(n), l.work, l.d_ipiv,
                                    l.d_info));
    diag_extract_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, l.stream>>>(l.mat, n, l.d_diag);
    GRM_CUDA_TRY(cudaGetLastError());
    GRM_CUDA_TRY(cudaMemcpyAsync(l.h_diag, l.d_diag, n * sizeof(double), cudaMemcpyDeviceToHost, l.stream));
    GRM_CUDA_TRY(cudaMemcpyAsync(l.h_ipiv, l.d_ipiv, n * sizeof(int), cudaMemcpyDeviceToHost, l.stream));
    GRM_CUDA_TRY(cudaMemcpyAsync(l.h_info, l.d_info, sizeof(int), cudaMemcpyDeviceToHost, l.stream));
    return GRM_CUDA_OK;
}

int32_t lane_launch_chol(grm_cuda_ctx *ctx, const Lane &l, const double *X, int64_t n, int64_t ldx, int rounds, double eps0) {
    GRM_TRY(lane_copy_in(ctx, l, X, n, ldx, rounds, eps0));
    GRM_SOLVER_TRY(cusolverDnDpotrf(l.h, CUBLAS_FILL_MODE_UPPER, static_cast<int>(n), l.mat, static_cast<int>(n), l.work, l.lwork,
                                    l.d_info));
    GRM_CUDA_TRY(cudaMemcpyAsync(l.h_info, l.d_info, sizeof(int), cudaMemcpyDeviceToHost, l.stream));
    return GRM_CUDA_OK;
}

// Julia's det(lu(X; check = false)) from a finished lane (the stream has been synchronised)
int32_t lane_det_value(const Lane &l, int64_t n, double *out) {
    if (*l.h_info < 0) {
        grm_set_error("cusolverDnDgetrf: illegal argument %d", -*l.h_info);
        return GRM_CUDA_ERR_CUDA;
    }
    if (*l.h_info > 0) {
        *out = 0.0;
        return GRM_CUDA_OK;
    }
    volatile double P = 1.0;  // one IEEE binary64 rounding per multiply, no reassociation
    int swaps = 0;
    for (int64_t i = 0; i < n; ++i) {
        P = P * l.h_diag[i];
        if (static_cast<int64_t>(l.h_ipiv[i]) != i + 1) ++swaps;
    }
    *out = (swaps & 1) ? -P : P;
    return GRM_CUDA_OK;
}

int32_t lanes_setup(grm_cuda_ctx *ctx, int64_t n, Lane (&L)[3]) {
    for (int i = 0; i < 2; ++i) {
        if (!ctx->cusolver_lane[i]) {
            cusolverDnHandle_t h;
            GRM_SOLVER_TRY(cusolverDnCreate(&h));
            ctx->cusolver_lane[i] = h;
            int lo = 0, hi = 0;
            GRM_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            GRM_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->lane_stream[i], cudaStreamNonBlocking, hi));
            GRM_SOLVER_TRY(cusolverDnSetStream(h, ctx->lane_stream[i]));
        }
    }
    for (auto &e : ctx->lane_ev)
        if (!e) GRM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const size_t mat = static_cast<size_t>(n) * n * sizeof(double);
    GRM_TRY(ctx->lu.reserve(mat));
    GRM_TRY(ctx->lu2.reserve(mat));
    GRM_TRY(ctx->lu3.reserve(mat));
    L[0].h = static_cast<cusolverDnHandle_t>(ctx->cusolver);  // det(t): the context's own handle and stream
    L[0].stream = ctx->stream;
    L[0].mat = ctx->lu2.as<double>();
    L[1].h = static_cast<cusolverDnHandle_t>(ctx->cusolver_lane[0]);  // det(t + 1)
    L[1].stream = ctx->lane_stream[0];
    L[1].mat = ctx->lu3.as<double>();
    L[2].h = static_cast<cusolverDnHandle_t>(ctx->cusolver_lane[1]);  // Cholesky(t + 1), into the buffer last_factor reads
    L[2].stream = ctx->lane_stream[1];
    L[2].mat = ctx->lu.as<double>();
    int lw1 = 0, lw2 = 0;
    GRM_SOLVER_TRY(cusolverDnDgetrf_bufferSize(L[0].h, static_cast<int>(n), static_cast<int>(n), L[0].mat, static_cast<int>(n), &lw1));
    GRM_SOLVER_TRY(cusolverDnDpotrf_bufferSize(L[0].h, CUBLAS_FILL_MODE_UPPER, static_cast<int>(n), L[0].mat, static_cast<int>(n), &lw2));
    const int lw = std::max(std::max(lw1, lw2), 1);
    GRM_TRY(ctx->lu_work.reserve(static_cast<size_t>(lw) * sizeof(double)));
    GRM_TRY(ctx->lane_work[0].reserve(static_cast<size_t>(lw) * sizeof(double)));
    GRM_TRY(ctx->lane_work[1].reserve(static_cast<size_t>(lw) * sizeof(double)));
    // per lane on the device: ipiv (n ints) | info | pad | diag (n doubles); pinned mirror of the same size
    const int64_t int_slots = (n + 16 + 1) & ~int64_t(1);
    const size_t per = static_cast<size_t>(int_slots) * sizeof(int) + static_cast<size_t>(n) * sizeof(double);
    GRM_TRY(ctx->ipiv.reserve(per));
    GRM_TRY(ctx->lane_ipiv[0].reserve(per));
    GRM_TRY(ctx->lane_ipiv[1].reserve(per));
    if (ctx->lane_pin_cap < 3 * per) {
        if (ctx->lane_pin) cudaFreeHost(ctx->lane_pin);
        ctx->lane_pin = nullptr;
        ctx->lane_pin_cap = 0;
        GRM_CUDA_TRY(cudaHostAlloc(&ctx->lane_pin, 3 * per, cudaHostAllocDefault));
        ctx->lane_pin_cap = 3 * per;
    }
    DevBuf *dev[3] = {&ctx->ipiv, &ctx->lane_ipiv[0], &ctx->lane_ipiv[1]};
    double *works[3] = {ctx->lu_work.as<double>(), ctx->lane_work[0].as<double>(), ctx->lane_work[1].as<double>()};
    for (int i = 0; i < 3; ++i) {
        L[i].work = works[i];
        L[i].lwork = lw;
        L[i].d_ipiv = dev[i]->as<int>();
        L[i].d_info = L[i].d_ipiv + n;
        L[i].d_diag = reinterpret_cast<double *>(L[i].d_ipiv + int_slots);
        char *hp = static_cast<char *>(ctx->lane_pin) + static_cast<size_t>(i) * per;
        L[i].h_ipiv = reinterpret_cast<int *>(hp);
        L[i].h_info = L[i].h_ipiv + n;
        L[i].h_diag = reinterpret_cast<double *>(L[i].h_ipiv + int_slots);
    }
    return GRM_CUDA_OK;
}

// inflatediagonals! (calc.jl:41-70) with two rounds in flight; `Repair r` (set up by the caller) supplies the sequential
// single-factorisation path for the on-demand Cholesky of round t
int32_t repair_speculative(grm_cuda_ctx *ctx, Repair &r, double *X, int64_t n, int64_t ldx, int32_t max_iter,
                           const unsigned long long *h_scan, int32_t *iters, double *eps_total) {
    const double eps = DBL_EPSILON;
    double maxabs = 0.0;
    {
        const long long bits = static_cast<long long>(h_scan[1]);
        memcpy(&maxabs, &bits, sizeof(double));
    }
    Lane L[3];
    GRM_TRY(lanes_setup(ctx, n, L));
    GRM_SOLVER_TRY(cusolverDnSetStream(L[0].h, ctx->stream));
    // results of the batch in flight: det of rounds b and b + 1, Cholesky of round b + 1
    int64_t b = -1;
    double det_b[2] = {NAN, NAN};
    bool pd_next = false;
    int64_t chol_round = -1;  // round whose Cholesky factor currently sits in ctx->lu (valid only if it succeeded)
    bool chol_ok = false;
    auto batch = [&](int64_t t) -> int32_t {
        // X is final on the context's stream: the two side lanes start after everything enqueued there so far
        GRM_CUDA_TRY(cudaEventRecord(ctx->lane_ev[0], ctx->stream));
        GRM_CUDA_TRY(cudaStreamWaitEvent(L[1].stream, ctx->lane_ev[0], 0));
        GRM_CUDA_TRY(cudaStreamWaitEvent(L[2].stream, ctx->lane_ev[0], 0));
        // cuSOLVER's getrf keeps its calling thread busy for most of the factorisation (issued from one thread the lanes
        // ran one after another: 694 -> 658 ms on C3), so every lane is driven by a host thread of its own
        int32_t st_side[2] = {GRM_CUDA_OK, GRM_CUDA_OK};
        std::string err_side[2];
        auto side = [&](int which) {
            if (cudaSetDevice(ctx->device) != cudaSuccess) {
                st_side[which] = GRM_CUDA_ERR_CUDA;
                return;
            }
            const Lane &l = L[1 + which];
            int32_t rc = which == 0 ? lane_launch_det(ctx, l, X, n, ldx, static_cast<int>(t + 1), maxabs)
                                    : lane_launch_chol(ctx, l, X, n, ldx, static_cast<int>(t + 1), maxabs);
            if (rc == GRM_CUDA_OK && cudaStreamSynchronize(l.stream) != cudaSuccess) rc = GRM_CUDA_ERR_CUDA;
            if (rc != GRM_CUDA_OK) err_side[which] = grm_cuda_last_error();  // the message is thread-local
            st_side[which] = rc;
        };
        std::thread th1(side, 0), th2(side, 1);
        int32_t rc0 = lane_launch_det(ctx, L[0], X, n, ldx, static_cast<int>(t), maxabs);
        if (rc0 == GRM_CUDA_OK && cudaStreamSynchronize(L[0].stream) != cudaSuccess) {
            grm_set_error("cudaStreamSynchronize failed in the repair");
            rc0 = GRM_CUDA_ERR_CUDA;
        }
        th1.join();
        th2.join();
        ctx->launches += 7 + (t > 0 ? 1 : 0);  // counted here: the lanes run on three host threads (copy, shift, extract)
        GRM_TRY(rc0);
        for (int w = 0; w < 2; ++w)
            if (st_side[w] != GRM_CUDA_OK) {
                grm_set_error("%s", err_side[w].empty() ? "repair lane failed" : err_side[w].c_str());
                return st_side[w];
            }
        GRM_TRY(lane_det_value(L[0], n, &det_b[0]));
        GRM_TRY(lane_det_value(L[1], n, &det_b[1]));
        if (*L[2].h_info < 0) {
            grm_set_error("cusolverDnDpotrf: illegal argument %d", -*L[2].h_info);
            return GRM_CUDA_ERR_CUDA;
        }
        pd_next = *L[2].h_info == 0;
        chol_round = t + 1;
        chol_ok = pd_next;
        b = t;
        return GRM_CUDA_OK;
    };
    auto det_of = [&](int64_t t, double *out) -> int32_t {
        if (b < 0 || t < b || t > b + 1) GRM_TRY(batch(t));
        *out = det_b[t - b];
        return GRM_CUDA_OK;
    };
    auto pd_of = [&](int64_t t, bool *out) -> int32_t {
        if (t == chol_round) {  // evaluated with its batch (or on demand before)
            *out = chol_ok;
            return GRM_CUDA_OK;
        }
        r.shift_rounds = static_cast<int>(t);
        r.shift_eps0 = maxabs;
        GRM_TRY(r.posdef(out));  // sequential, into ctx->lu
        chol_round = t;
        chol_ok = *out;
        return GRM_CUDA_OK;
    };

    // calc.jl:49-51 — early return iff det(X) > eps && isposdef(X)
    double det_t = NAN;
    bool pd_t = false;
    GRM_TRY(det_of(0, &det_t));
    if (!(det_t < eps)) GRM_TRY(pd_of(0, &pd_t));
    if (pd_t && det_t > eps) {
        ctx->factor_n = n;
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
    // calc.jl:58-66
    int64_t t = 0;
    for (;;) {
        const bool need = det_t < eps || !pd_t;  // a NaN det compares false, then isposdef decides
        if (!need || t >= max_iter) {
            ctx->factor_n = (!need && chol_round == t && chol_ok) ? n : -1;
            break;
        }
        ++t;
        GRM_TRY(det_of(t, &det_t));
        pd_t = false;
        if (!(det_t < eps)) GRM_TRY(pd_of(t, &pd_t));
    }
    *iters = static_cast<int32_t>(t);
    return grm_cuda_repair_apply(ctx, X, n, ldx, *iters, maxabs, eps_total);
}

}  // namespace

// device-resident implementation; X is a device pointer
int32_t grm_repair_device(grm_cuda_ctx *ctx, double *X, int64_t n, int64_t ldx, int32_t max_iter, int32_t *iters,
                          double *eps_total) {
    *iters = 0;
    *eps_total = 0.0;
    Repair r;
    unsigned long long h_scan[2];
    GRM_TRY(repair_setup(ctx, X, n, ldx, r, h_scan));
    if (h_scan[0] & 1ull) {
        grm_set_error("The input matrix is not symmetric.");
        return GRM_CUDA_ERR_NOT_SYMMETRIC;
    }
    if (ctx->tune.repair != 1 && n >= 3072) return repair_speculative(ctx, r, X, n, ldx, max_iter, h_scan, iters, eps_total);

    const double eps = DBL_EPSILON;
    // calc.jl:49-51 — early return iff det(X) > eps && isposdef(X)
    bool pd = false;
    double det_cur = NAN;
    GRM_TRY(r.evaluate(&pd, &det_cur));
    if (pd && det_cur > eps) {
        ctx->factor_n = n;  // ctx->lu holds chol(X).U (grm_cuda_last_factor)
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
    double e = 0.0;
    {
        const long long bits = static_cast<long long>(h_scan[1]);
        memcpy(&e, &bits, sizeof(double));
    }
    double total = 0.0;
    int iter = 0;
    for (;;) {
        // calc.jl:61: det < eps || !isposdef  (a NaN det compares false, then isposdef decides)
        const bool need = det_cur < eps || !pd;
        if (!need) ctx->factor_n = n;  // the last factorisation into ctx->lu is the Cholesky of the matrix being returned
        if (!need || iter >= max_iter) break;
        ++iter;
        diag_add_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(X, n, ldx, e);
        GRM_CUDA_TRY(cudaGetLastError());
        ctx->launches++;
        total += e;
        e *= (1.0 + eps);
        GRM_TRY(r.evaluate(&pd, &det_cur));
    }
    *iters = iter;
    *eps_total = total;
    return GRM_CUDA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU inflatediagonals!: the loop tests of different rounds are independent given max|X| (round t looks at X
// with its diagonal inflated t times), so with the matrix replicated on every GPU, rank r evaluates round r: the
// wall time of the repair becomes one LU (+ one Cholesky) instead of rounds + 1 of them.  The arithmetic of each probe
// is the arithmetic of the sequential loop, so the outcome is identical.
// ---------------------------------------------------------------------------------------------------------------
// which: bit 0 = the determinant, bit 1 = isposdef; 3 = both with the reference's short-circuit (isposdef only when
// !(det < eps)).  A part that was not evaluated comes back as det = NaN / posdef = -1.
int32_t grm_repair_probe_part(grm_cuda_ctx *ctx, const double *dX, int64_t n, int64_t ldx, int32_t rounds, int32_t which,
                              double *det_out, int32_t *posdef_out, uint32_t *scan_flags, double *maxabs) {
    GRM_CUDA_TRY(cudaSetDevice(ctx->device));
    Repair r;
    unsigned long long h_scan[2];
    GRM_TRY(repair_setup(ctx, const_cast<double *>(dX), n, ldx, r, h_scan));
    *scan_flags = static_cast<uint32_t>(h_scan[0]);
    double mx = 0.0;
    {
        const long long bits = static_cast<long long>(h_scan[1]);
        memcpy(&mx, &bits, sizeof(double));
    }
    *maxabs = mx;
    *det_out = NAN;
    *posdef_out = -1;
    if (h_scan[0] & 1ull) return GRM_CUDA_OK;  // not symmetric: the caller raises before any factorisation (calc.jl:43)
    r.shift_rounds = rounds;
    r.shift_eps0 = mx;
    bool pd = false;
    if (which == 3) {
        GRM_TRY(r.evaluate(&pd, det_out));
        *posdef_out = (*det_out < DBL_EPSILON) ? -1 : (pd ? 1 : 0);  // -1: not evaluated, det < eps already decides
        return GRM_CUDA_OK;
    }
    if (which & 1) GRM_TRY(r.det(det_out));
    if (which & 2) {
        GRM_TRY(r.posdef(&pd));
        *posdef_out = pd ? 1 : 0;
    }
    return GRM_CUDA_OK;
}

extern "C" int32_t grm_cuda_repair_probe(grm_cuda_ctx *ctx, const double *dX, int64_t n, int64_t ldx, int32_t rounds,
                                         double *det_out, int32_t *posdef_out, uint32_t *scan_flags, double *maxabs) {
    if (!ctx || !dX || n < 1 || ldx < n || rounds < 0 || !det_out || !posdef_out || !scan_flags || !maxabs) {
        grm_set_error("repair_probe: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    return grm_repair_probe_part(ctx, dX, n, ldx, rounds, 3, det_out, posdef_out, scan_flags, maxabs);
}

extern "C" int32_t grm_cuda_repair_apply(grm_cuda_ctx *ctx, double *dX, int64_t n, int64_t ldx, int32_t rounds,
                                         double maxabs, double *eps_total) {
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
}

// ---------------------------------------------------------------------------------------------------------------
// Consumer side of the GRM: simulateeffects builds MvNormal(mu, Sigma) from
// simulatecovariancekinship's matrix (src/simulations/simulate_effects.jl:344-351, :479-481), i.e. PDMat(Sigma) ->
// cholesky(Sigma) = LAPACK potrf!('U', copy(Sigma)).  The repair above ends with exactly that factorisation of the
// matrix it returns, so the factor can be handed out instead of being recomputed (1/3 n^3 flop saved).
// ---------------------------------------------------------------------------------------------------------------
namespace {

// U[i,j] = F[i,j] for i <= j, 0 below the diagonal (potrf leaves the other triangle of its input untouched)
__global__ void triu_copy_kernel(const double *__restrict__ F, int64_t n, double *__restrict__ U, int64_t ldu) {
    const int64_t j = blockIdx.y;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        U[i + j * ldu] = (i <= j) ? F[i + j * n] : 0.0;
}

int32_t emit_factor(grm_cuda_ctx *ctx, int64_t n, double *U, int64_t ldu, int32_t location) {
    const double *F = ctx->lu.as<double>();
    double *dU = U;
    int64_t ld = ldu;
    if (location == GRM_CUDA_LOC_HOST) {
        GRM_TRY(ctx->misc2.reserve(static_cast<size_t>(n) * n * sizeof(double)));
        dU = ctx->misc2.as<double>();
        ld = n;
    }
    const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 64));
    triu_copy_kernel<<<dim3(gx, static_cast<unsigned>(n)), 256, 0, ctx->stream>>>(F, n, dU, ld);
    GRM_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    if (location == GRM_CUDA_LOC_HOST)
        GRM_CUDA_TRY(cudaMemcpy2DAsync(U, ldu * sizeof(double), dU, n * sizeof(double), n * sizeof(double),
                                       static_cast<size_t>(n), cudaMemcpyDeviceToHost, ctx->stream));
    GRM_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GRM_CUDA_OK;
}

}  // namespace

extern "C" int32_t grm_cuda_last_factor(grm_cuda_ctx *ctx, int64_t n, double *U, int64_t ldu, int32_t location) {
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
}

extern "C" int32_t grm_cuda_cholesky(grm_cuda_ctx *ctx, const double *X, int64_t n, int64_t ldx, int32_t x_location,
                                     double *U, int64_t ldu, int32_t u_location) {
    if (!ctx || !X || !U || n < 1 || ldx < n || ldu < n ||
        (x_location != GRM_CUDA_LOC_HOST && x_location != GRM_CUDA_LOC_DEVICE) ||
        (u_location != GRM_CUDA_LOC_HOST && u_location != GRM_CUDA_LOC_DEVICE)) {
        grm_set_error("cholesky: bad argument");
        return GRM_CUDA_ERR_BAD_ARGUMENT;
    }
    if (n > 46000) {
        grm_set_error("cholesky: n=%lld exceeds the single-GPU cuSOLVER range (n <= 46000)", (long long)n);
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
    GRM_CUDA_TRY(cudaMemcpy2DAsync(F, n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double), static_cast<size_t>(n),
                                   x_location == GRM_CUDA_LOC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                   ctx->stream));
    int lw = 0;
    GRM_SOLVER_TRY(cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_UPPER, static_cast<int>(n), F, static_cast<int>(n), &lw));
    GRM_TRY(ctx->lu_work.reserve(std::max<size_t>(static_cast<size_t>(lw), 1) * sizeof(double)));
    GRM_TRY(ctx->misc.reserve(64));
    int *info = ctx->misc.as<int>();
    // like Julia's cholesky(Symmetric(X)), only the upper triangle of X is read
    GRM_SOLVER_TRY(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_UPPER, static_cast<int>(n), F, static_cast<int>(n),
                                    ctx->lu_work.as<double>(), lw, info));
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
    return emit_factor(ctx, n, U, ldu, u_location);
}
