/*
 * grm_oracle.c — CPU restatement of the reference's GRM path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product (libgrm_cuda) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference (GenomicBreedingCore.jl) is pure Julia, Julia is not installed in the build image
 * or on the GPU boxes, and the reference's own tests hold no numeric golden vectors for this path (only the
 * properties size == (n,n), issymmetric, det > eps(Float64): src/grm/calc.jl:30-39, :103-113, :177-187).  This
 * restatement follows the reference source line by line (citations below, paths relative to the reference root)
 * and the Julia stdlib semantics documented in SURVEY.md Appendix A; those three properties are asserted in
 * tests/test_oracle.py.  Arithmetic that lives in Julia's stdlib rather than in the reference:
 *   LinearAlgebra.dot (generic fallback: s = 0.0; s += x*y in index order), Statistics.mean(A, dims=1)
 *   (column sum ./ n), Base.sum (pairwise, 1024-element sequential base case), LinearAlgebra.det (LU with partial
 *   pivoting, sequential product of U's diagonal, sign from the pivot count, 0 if the factorisation hit a zero
 *   pivot), LinearAlgebra.isposdef (ishermitian && LAPACK potrf 'U' succeeds) — Julia 1.12 / LinearAlgebra 1.11.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -pthread (see oracle/Makefile).
 * -ffp-contract=off matters: Julia never fuses a*b+c, so neither may this file.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_ERR_ARG 1
#define ORACLE_ERR_MISSING 2
#define ORACLE_ERR_NOT_SYMMETRIC 4
#define ORACLE_ERR_NAN 5
#define ORACLE_ERR_INF 6

/* ------------------------------------------------------------------------------------------------------------ */
/* the pair loop: calc.jl:124-131 (simple) and :202-209 (ploidy-aware)                                            */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const double *A;
    int64_t n, p, lda;
    double *G;
    int64_t ldg;
    int64_t i_begin, i_end; /* this thread's contiguous chunk of i, like Threads.@threads */
    int64_t j_limit;        /* pairs (i, j) with i <= j < j_limit (== n for the full matrix) */
    int centred;
    double ploidy;
    const double *q;
    double denom;
    int missing;
} pair_job;

/* A[i, :] — the reference materialises a fresh row vector per pair (calc.jl:127), stride lda */
static void gather_row(const double *A, int64_t i, int64_t p, int64_t lda, double *out) {
    for (int64_t k = 0; k < p; ++k) out[k] = A[i + k * lda];
}

static void *pair_worker(void *arg) {
    pair_job *jb = (pair_job *)arg;
    const int64_t p = jb->p;
    double *xi = (double *)malloc(sizeof(double) * (size_t)p);
    double *xj = (double *)malloc(sizeof(double) * (size_t)p);
    for (int64_t i = jb->i_begin; i < jb->i_end; ++i) {
        for (int64_t j = i; j < jb->j_limit; ++j) {
            gather_row(jb->A, i, p, jb->lda, xi);
            gather_row(jb->A, j, p, jb->lda, xj);
            if (jb->centred) {
                /* G_star[i,:] - q with G_star = ploidy .* (A .- 0.5)   (calc.jl:198, :205) */
                for (int64_t k = 0; k < p; ++k) {
                    xi[k] = (jb->ploidy * (xi[k] - 0.5)) - jb->q[k];
                    xj[k] = (jb->ploidy * (xj[k] - 0.5)) - jb->q[k];
                }
            }
            /* LinearAlgebra.dot generic fallback: s = 0.0; s += x[k]*y[k] in order */
            double s = 0.0;
            for (int64_t k = 0; k < p; ++k) {
                if (xi[k] != xi[k] || xj[k] != xj[k]) jb->missing = 1;
                s += xi[k] * xj[k];
            }
            const double a = s / jb->denom; /* "/ p" (calc.jl:127) or "/ d" (calc.jl:205) */
            jb->G[i + j * jb->ldg] = a;     /* calc.jl:128-129 */
            jb->G[j + i * jb->ldg] = a;
        }
    }
    free(xi);
    free(xj);
    return NULL;
}

static int run_pairs(const double *A, int64_t n, int64_t p, int64_t lda, double *G, int64_t ldg, int64_t i_limit,
                     int centred, double ploidy, const double *q, double denom, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    pair_job jobs[256];
    /* Threads.@threads: nthreads contiguous chunks of the outer index */
    for (int t = 0; t < nthreads; ++t) {
        pair_job *jb = &jobs[t];
        jb->A = A;
        jb->n = n;
        jb->p = p;
        jb->lda = lda;
        jb->G = G;
        jb->ldg = ldg;
        jb->i_begin = i_limit * t / nthreads;
        jb->i_end = i_limit * (t + 1) / nthreads;
        jb->j_limit = n;
        jb->centred = centred;
        jb->ploidy = ploidy;
        jb->q = q;
        jb->denom = denom;
        jb->missing = 0;
    }
    if (nthreads == 1) {
        pair_worker(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, pair_worker, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    for (int t = 0; t < nthreads; ++t)
        if (jobs[t].missing) return ORACLE_ERR_MISSING; /* the reference throws: Missing cannot be stored in Matrix{Float64} */
    return ORACLE_OK;
}

/* grmsimple before inflatediagonals! (calc.jl:122-131).  rows_limit < n computes only pairs with i < rows_limit
 * (bounded CPU-baseline sample); pass n for the full matrix. */
int oracle_grmsimple_raw(const double *A, int64_t n, int64_t p, int64_t lda, double *G, int64_t ldg,
                         int64_t rows_limit, int nthreads) {
    if (!A || !G || n < 1 || p < 1 || lda < n || ldg < n) return ORACLE_ERR_ARG;
    if (rows_limit < 0 || rows_limit > n) rows_limit = n;
    return run_pairs(A, n, p, lda, G, ldg, rows_limit, 0, 0.0, NULL, (double)p, nthreads);
}

/* Base.sum pairwise reduction (mapreduce_impl, block size 1024) of q_k * (1 - q_k) */
static double pairwise_q1mq(const double *q, int64_t first, int64_t last) {
    if (first == last) return q[first] * (1.0 - q[first]);
    if (last - first < 1024) {
        double v = (q[first] * (1.0 - q[first])) + (q[first + 1] * (1.0 - q[first + 1]));
        for (int64_t i = first + 2; i <= last; ++i) v = v + (q[i] * (1.0 - q[i]));
        return v;
    }
    const int64_t mid = first + ((last - first) >> 1);
    const double v1 = pairwise_q1mq(q, first, mid);
    const double v2 = pairwise_q1mq(q, mid + 1, last);
    return v1 + v2;
}

/* q = mean(A, dims=1)[1,:] (calc.jl:197): column sum (sequential) ./ n ;  d = ploidy * sum(q .* (1 .- q)) (:201) */
int oracle_locus_stats(const double *A, int64_t n, int64_t p, int64_t lda, int ploidy, double *q, double *d) {
    if (!A || !q || !d || n < 1 || p < 1 || lda < n) return ORACLE_ERR_ARG;
    for (int64_t k = 0; k < p; ++k) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += A[i + k * lda];
        q[k] = s / (double)n;
    }
    *d = (double)ploidy * pairwise_q1mq(q, 0, p - 1);
    return ORACLE_OK;
}

/* grmploidyaware before inflatediagonals! (calc.jl:195-209) */
int oracle_grmploidyaware_raw(const double *A, int64_t n, int64_t p, int64_t lda, int ploidy, double *G, int64_t ldg,
                              double *q_out, double *d_out, int64_t rows_limit, int nthreads) {
    if (!A || !G || n < 1 || p < 1 || lda < n || ldg < n) return ORACLE_ERR_ARG;
    double *q = (double *)malloc(sizeof(double) * (size_t)p);
    double d = 0.0;
    int st = oracle_locus_stats(A, n, p, lda, ploidy, q, &d);
    if (st == ORACLE_OK) {
        if (rows_limit < 0 || rows_limit > n) rows_limit = n;
        st = run_pairs(A, n, p, lda, G, ldg, rows_limit, 1, (double)ploidy, q, d, nthreads);
    }
    if (q_out) memcpy(q_out, q, sizeof(double) * (size_t)p);
    if (d_out) *d_out = d;
    free(q);
    return st;
}

/* exact integer Gram matrix of row-major int8 codes: C[i + j*n] = sum_k c[i,k] c[j,k]  (checker for the int8 SYRK) */
int oracle_gram_i8(const int8_t *codes, int64_t n, int64_t p, int64_t ldc, int64_t *C) {
    if (!codes || !C || n < 1 || p < 1 || ldc < p) return ORACLE_ERR_ARG;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i; j < n; ++j) {
            int64_t s = 0;
            const int8_t *a = codes + i * ldc, *b = codes + j * ldc;
            for (int64_t k = 0; k < p; ++k) s += (int64_t)a[k] * (int64_t)b[k];
            C[i + j * n] = s;
            C[j + i * n] = s;
        }
    return ORACLE_OK;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* inflatediagonals! (calc.jl:41-70)                                                                              */
/* ------------------------------------------------------------------------------------------------------------ */
/* det(X) as Julia computes it: lu(X; check=false) = LAPACK getrf (partial pivoting, first maximal |a|), then
 * P = 1.0; P *= U[i,i] for i = 1..n; sign = (-1)^(number of i with ipiv[i] != i); singular factorisation -> 0. */
double oracle_det(const double *X, int64_t n, int64_t ldx) {
    double *a = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < n; ++i) a[i + j * n] = X[i + j * ldx];
    int swaps = 0, singular = 0;
    for (int64_t k = 0; k < n; ++k) {
        int64_t piv = k;
        double best = fabs(a[k + k * n]);
        for (int64_t i = k + 1; i < n; ++i) {
            const double v = fabs(a[i + k * n]);
            if (v > best) { /* NaN never wins, like idamax */
                best = v;
                piv = i;
            }
        }
        if (a[piv + k * n] == 0.0) {
            singular = 1; /* getrf: info = k, factorisation continues; det(::LU) returns 0 when !issuccess */
            continue;
        }
        if (piv != k) {
            ++swaps;
            for (int64_t j = 0; j < n; ++j) {
                const double t = a[k + j * n];
                a[k + j * n] = a[piv + j * n];
                a[piv + j * n] = t;
            }
        }
        const double r = 1.0 / a[k + k * n];
        for (int64_t i = k + 1; i < n; ++i) a[i + k * n] *= r;
        for (int64_t j = k + 1; j < n; ++j) {
            const double ukj = a[k + j * n];
            if (ukj != 0.0)
                for (int64_t i = k + 1; i < n; ++i) a[i + j * n] -= a[i + k * n] * ukj;
        }
    }
    double P = 1.0;
    if (singular) {
        P = 0.0;
    } else {
        for (int64_t i = 0; i < n; ++i) P = P * a[i + i * n];
        if (swaps & 1) P = -P;
    }
    free(a);
    return P;
}

/* isposdef(X) for a symmetric X: LAPACK potrf('U') succeeds, i.e. every pivot a_jj - sum_k u_kj^2 is > 0 (and not NaN) */
int oracle_isposdef(const double *X, int64_t n, int64_t ldx) {
    double *u = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < n; ++i) u[i + j * n] = X[i + j * ldx];
    int ok = 1;
    for (int64_t j = 0; j < n && ok; ++j) {
        double ajj = u[j + j * n];
        for (int64_t k = 0; k < j; ++k) ajj -= u[k + j * n] * u[k + j * n];
        if (!(ajj > 0.0)) {
            ok = 0;
            break;
        }
        ajj = sqrt(ajj);
        u[j + j * n] = ajj;
        for (int64_t c = j + 1; c < n; ++c) {
            double s = u[j + c * n];
            for (int64_t k = 0; k < j; ++k) s -= u[k + j * n] * u[k + c * n];
            u[j + c * n] = s / ajj;
        }
    }
    free(u);
    return ok;
}

int oracle_inflatediagonals(double *X, int64_t n, int64_t ldx, int max_iter, int *iters, double *eps_total) {
    if (!X || n < 1 || ldx < n) return ORACLE_ERR_ARG;
    if (iters) *iters = 0;
    if (eps_total) *eps_total = 0.0;
    /* calc.jl:43-45  issymmetric: X[i,j] == X[j,i] for all i <= j (NaN != NaN, so NaN anywhere => not symmetric) */
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i <= j; ++i)
            if (!(X[i + j * ldx] == X[j + i * ldx])) return ORACLE_ERR_NOT_SYMMETRIC;
    /* calc.jl:49-51 */
    const double eps = DBL_EPSILON;
    if (oracle_det(X, n, ldx) > eps && oracle_isposdef(X, n, ldx)) return ORACLE_OK;
    /* calc.jl:52-57 */
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < n; ++i)
            if (X[i + j * ldx] != X[i + j * ldx]) return ORACLE_ERR_NAN;
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < n; ++i)
            if (isinf(X[i + j * ldx])) return ORACLE_ERR_INF;
    /* calc.jl:58-66 */
    int iter = 0;
    double e = 0.0;
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < n; ++i) {
            const double v = fabs(X[i + j * ldx]);
            if (v > e) e = v;
        }
    double total = 0.0;
    while (((oracle_det(X, n, ldx) < eps) || !oracle_isposdef(X, n, ldx)) && (iter < max_iter)) {
        iter += 1;
        for (int64_t i = 0; i < n; ++i) X[i + i * ldx] = X[i + i * ldx] + e;
        total += e;
        e *= (1.0 + eps);
    }
    if (iters) *iters = iter;
    if (eps_total) *eps_total = total;
    return ORACLE_OK;
}
