// repair_host.h — host half of inflatediagonals! (src/grm/calc.jl:41-70): the determinant the reference tests with
// `det(X) < eps(Float64)` (calc.jl:49,61), rebuilt from the pivots a device factorisation hands back.  Plain C++, no CUDA.
#pragma once
#include <stdint.h>

// Julia's det(lu(X)): SEQUENTIAL product of diag(U) in pivot order, one IEEE binary64 rounding per multiply (gradual
// underflow, saturation at 0 / Inf are part of the semantics), sign flipped once per row interchange
// (ipiv is LAPACK's 1-based pivot vector).
double grm_det_lu_pivots(const double *diag_u, const int *ipiv, int64_t n);

// Why grm_det_chol_pivots did not decide
constexpr int GRM_CHOLDET_OK = 0;         // *det_out is the value an LU would multiply up, as far as `det < eps` goes
constexpr int GRM_CHOLDET_ILLCOND = 1;    // a pivot below 2^-26 * scale (or NaN): cancellation decides it, not the algorithm
constexpr int GRM_CHOLDET_BORDERLINE = 2; // the product ends within 2^+-64 of eps
constexpr int GRM_CHOLDET_UNDERFLOW = 3;  // the product passed through subnormals and what follows could bring it back
constexpr int GRM_CHOLDET_OVERFLOW = 4;   // the product came near the overflow threshold and what follows could bring it down

// The same determinant from the diagonal of a LOWER Cholesky factor L of a symmetric positive definite X whose columns
// satisfy |L_jk| < L_kk (j > k) — the caller checks that on the device.  Under that condition Gaussian elimination with
// partial pivoting performs no row interchange on X (column k of the k-th Schur complement is L_kk * L_:k, so its
// largest entry is the diagonal one) and its pivots are u_k = L_kk^2: the product below is the product det(lu(X))
// forms, up to the rounding differences between two factorisation routines (cuSOLVER's getrf and the reference's
// OpenBLAS getrf differ by as much).  The function multiplies fl(L_kk^2) sequentially, exactly like
// grm_det_lu_pivots, and returns GRM_CHOLDET_OK only when the outcome of `det < eps` cannot depend on such
// differences: no ill-determined pivot, and the end value far from eps along every nearby trajectory (bounds are carried
// through the subnormal and the near-overflow range, where a running product forgets relative accuracy).  Any other
// return value asks the caller for the real LU.  scale = max|X| of the factorised matrix (> 0).
int grm_det_chol_pivots(const double *diag_l, int64_t n, double scale, double *det_out);
