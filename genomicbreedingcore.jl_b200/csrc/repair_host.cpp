// repair_host.cpp — see repair_host.h.  Compiled by g++ (no fast-math, no FMA contraction: x86-64 baseline), and every
// running product is `volatile`: one IEEE binary64 rounding per multiply, in pivot order, as Julia's det(::LU) does.
#include "repair_host.h"

#include <float.h>
#include <math.h>

#include <algorithm>

double grm_det_lu_pivots(const double *diag_u, const int *ipiv, int64_t n) {
    volatile double P = 1.0;
    int swaps = 0;
    for (int64_t i = 0; i < n; ++i) {
        P = P * diag_u[i];
        if (static_cast<int64_t>(ipiv[i]) != i + 1) ++swaps;
    }
    return (swaps & 1) ? -P : P;
}

int grm_det_chol_pivots(const double *diag_l, int64_t n, double scale, double *det_out) {
    const double eps = DBL_EPSILON;
    // Two factorisation routines agree on a pivot u_k to about k * 2^-53 * scale / u_k (the Schur complement is a
    // difference of numbers of size `scale`).  Pivots below 2^-26 * scale are refused; for the others the bounds below
    // allow a relative difference of 2^-12 PER PIVOT, far more than that estimate, and ask for 2^64 of room around eps.
    const double pivot_floor = ldexp(scale, -26);
    const double up = 1.0 + ldexp(1.0, -12), down = 1.0 - ldexp(1.0, -12);
    const double LOW = ldexp(1.0, -1000), HIGH = ldexp(1.0, 1000);
    const double W_CAP = ldexp(1.0, 1000), V_CAP = ldexp(1.0, 999);
    volatile double P = 1.0;
    bool illcond = !(scale > 0.0) || !(pivot_floor > 0.0);
    bool went_low = false, went_high = false;
    double pmax = 1.0;
    double W = 0.0;  // after the product first fell below LOW: upper bound of every nearby trajectory, in units of 2^-1075
    double V = 0.0;  // after it first rose above HIGH: lower bound of every nearby trajectory that has not overflowed
    for (int64_t k = 0; k < n; ++k) {
        volatile double u = diag_l[k] * diag_l[k];  // the pivot of the elimination, one rounding
        const double uk = u;
        if (!(uk >= pivot_floor) || !(uk <= DBL_MAX)) illcond = true;
        P = P * uk;
        const double p = P;
        // a subnormal result is off by at most half a subnormal ulp (2^-1075) in absolute terms, a normal one by 2^-53
        // relative: next <= cur * u * (1 + slack) + 2^-1075
        if (went_low) W = std::min(W * uk * up + 1.0, W_CAP);
        else if (p < LOW) {
            went_low = true;
            W = ldexp(1.0, 1075 - 999);  // nearby trajectories are below 2 * LOW here
        }
        if (went_high) V = std::min(V * uk * down, V_CAP);
        else if (p > HIGH) {
            went_high = true;
            V = V_CAP;  // nearby trajectories are above HIGH / 2 here (or have overflowed: Inf stays Inf, pivots are > 0)
        }
        if (p > pmax) pmax = p;
    }
    *det_out = P;
    if (illcond) return GRM_CHOLDET_ILLCOND;
    if (went_low) {
        // decided "< eps" only if nothing nearby can climb back to eps * 2^-64 = 2^-116, and nothing nearby overflowed before
        if (W < ldexp(1.0, 1075 - 116) && pmax < ldexp(1.0, 1020)) return GRM_CHOLDET_OK;
        return GRM_CHOLDET_UNDERFLOW;
    }
    if (went_high) {
        // decided ">= eps" (possibly Inf) only if nothing nearby can fall to eps * 2^64 = 2^12
        if (V > ldexp(1.0, 12)) return GRM_CHOLDET_OK;
        return GRM_CHOLDET_OVERFLOW;
    }
    const double p = P;
    if (p >= eps * ldexp(1.0, -64) && p <= eps * ldexp(1.0, 64)) return GRM_CHOLDET_BORDERLINE;
    return GRM_CHOLDET_OK;
}
