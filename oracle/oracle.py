"""CPU oracle for the GRM path — TEST INFRASTRUCTURE ONLY (see grm_oracle.c for the full header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
PARITY UNPINNED by the reference: no Julia here, and the reference's tests pin only size / issymmetric / det > eps
for this path (src/grm/calc.jl:30-39, :103-113, :177-187).

Two independent restatements of src/grm/calc.jl:
  * `c_*`  — the C restatement (oracle/grm_oracle.c): reference-faithful loop order and rounding (sequential
             mul-then-add dot, per-pair row gathers, Threads.@threads-style chunks), own LU/Cholesky.
  * `np_*` — NumPy/SciPy: exact integer arithmetic for dosage inputs, extended precision ("truth") for
             continuous inputs, and LAPACK getrf/potrf through SciPy for inflatediagonals! — the same LAPACK
             routines Julia's det/isposdef call.
tests/test_oracle.py checks them against each other, against tests/golden/*.npz and against the reference's
doctest properties.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgrm_oracle.so")

OK, ERR_ARG, ERR_MISSING, ERR_NOT_SYMMETRIC, ERR_NAN, ERR_INF = 0, 1, 2, 4, 5, 6
EPS = float(np.finfo(np.float64).eps)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "grm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = build() if os.path.exists(os.path.join(_HERE, "grm_oracle.c")) and _have_gcc() else _SO
        l = C.CDLL(so)
        vp, i64, i32, f64 = C.c_void_p, C.c_int64, C.c_int, C.c_double
        l.oracle_grmsimple_raw.argtypes = [vp, i64, i64, i64, vp, i64, i64, i32]
        l.oracle_grmploidyaware_raw.argtypes = [vp, i64, i64, i64, i32, vp, i64, vp, vp, i64, i32]
        l.oracle_locus_stats.argtypes = [vp, i64, i64, i64, i32, vp, vp]
        l.oracle_gram_i8.argtypes = [vp, i64, i64, i64, vp]
        l.oracle_det.argtypes = [vp, i64, i64]
        l.oracle_det.restype = f64
        l.oracle_isposdef.argtypes = [vp, i64, i64]
        l.oracle_inflatediagonals.argtypes = [vp, i64, i64, i32, C.POINTER(i32), C.POINTER(f64)]
        _lib = l
    return _lib


def _have_gcc() -> bool:
    from shutil import which

    return which("gcc") is not None and which("make") is not None


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


# ---------------------------------------------------------------------------------------------------------------
# C restatement
# ---------------------------------------------------------------------------------------------------------------
def c_grmsimple_raw(A, rows_limit=None, nthreads=1):
    """calc.jl:122-131, before inflatediagonals!.  Returns (status, G)."""
    A = _f(A)
    n, p = A.shape
    G = np.zeros((n, n), order="F")
    st = lib().oracle_grmsimple_raw(A.ctypes.data, n, p, n, G.ctypes.data, n, n if rows_limit is None else rows_limit,
                                    nthreads)
    return st, G


def c_grmploidyaware_raw(A, ploidy=2, rows_limit=None, nthreads=1):
    """calc.jl:195-209, before inflatediagonals!.  Returns (status, G, q, d)."""
    A = _f(A)
    n, p = A.shape
    G = np.zeros((n, n), order="F")
    q = np.zeros(p)
    d = C.c_double(0.0)
    st = lib().oracle_grmploidyaware_raw(A.ctypes.data, n, p, n, int(ploidy), G.ctypes.data, n, q.ctypes.data,
                                         C.byref(d), n if rows_limit is None else rows_limit, nthreads)
    return st, G, q, d.value


def c_det(X):
    X = _f(X)
    return lib().oracle_det(X.ctypes.data, X.shape[0], X.shape[0])


def c_isposdef(X):
    X = _f(X)
    return bool(lib().oracle_isposdef(X.ctypes.data, X.shape[0], X.shape[0]))


def c_inflatediagonals(X, max_iter=1000):
    """calc.jl:41-70 on a copy.  Returns (status, X_out, iters, eps_total)."""
    X = np.array(X, dtype=np.float64, order="F", copy=True)
    it = C.c_int(0)
    tot = C.c_double(0.0)
    st = lib().oracle_inflatediagonals(X.ctypes.data, X.shape[0], X.shape[0], int(max_iter), C.byref(it), C.byref(tot))
    return st, X, it.value, tot.value


def c_gram_i8(codes):
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    n, p = codes.shape
    out = np.zeros((n, n), dtype=np.int64, order="F")
    st = lib().oracle_gram_i8(codes.ctypes.data, n, p, p, out.ctypes.data)
    assert st == OK
    return out


def c_grmsimple(A, max_iter=1000, nthreads=1):
    st, G = c_grmsimple_raw(A, nthreads=nthreads)
    if st != OK:
        return st, G, 0, 0.0
    return c_inflatediagonals(G, max_iter)


def c_grmploidyaware(A, ploidy=2, max_iter=1000, nthreads=1):
    st, G, _, _ = c_grmploidyaware_raw(A, ploidy, nthreads=nthreads)
    if st != OK:
        return st, G, 0, 0.0
    return c_inflatediagonals(G, max_iter)


# ---------------------------------------------------------------------------------------------------------------
# NumPy / SciPy restatement
# ---------------------------------------------------------------------------------------------------------------
def dosage_codes(A, m):
    c = np.rint(np.asarray(A) * m).astype(np.int64)
    if not np.array_equal(c / float(m), A):
        raise ValueError(f"not multiples of 1/{m}")
    return c


def np_gram_int(codes):
    """Exact integer c*c' (int64)."""
    c = np.asarray(codes, dtype=np.int64)
    return c @ c.T


def np_grmsimple_raw_dosage(A, m):
    """SURVEY A.1 exactness lemma: for x = c/m, m a power of two, the reference's sequential FP64 dot is exactly
    C_int/m^2, so G = fl(fl(C_int / m^2) / p) bit for bit."""
    c = dosage_codes(A, m)
    Cint = np_gram_int(c)
    return (Cint.astype(np.float64) / float(m * m)) / float(A.shape[1])


def np_grmsimple_raw_seq(A):
    """Sequential mul-then-add dot through cumsum (NumPy's cumsum is a plain left-to-right loop)."""
    A = np.asarray(A, dtype=np.float64)
    n, p = A.shape
    G = np.empty((n, n))
    for i in range(n):
        prod = A[i][None, :] * A[i:]
        G[i, i:] = np.cumsum(prod, axis=1)[:, -1] / p
        G[i:, i] = G[i, i:]
    return G


def np_truth_simple(A):
    """Extended-precision (x87 long double) reference value of A*A'/p: says which side of a mismatch is closer."""
    L = np.asarray(A, dtype=np.longdouble)
    return np.asarray((L @ L.T) / np.longdouble(A.shape[1]), dtype=np.float64)


def np_truth_ploidyaware(A, ploidy=2):
    """Extended-precision evaluation of the reference's formula with the reference's (rounded) q and z."""
    A = np.asarray(A, dtype=np.float64)
    n, p = A.shape
    q = A.sum(axis=0) / n
    Z = (ploidy * (A - 0.5)) - q  # each op rounds once, as in calc.jl:198,205
    d = ploidy * float(np.sum((q * (1.0 - q)).astype(np.longdouble)))
    L = Z.astype(np.longdouble)
    return np.asarray((L @ L.T) / np.longdouble(d), dtype=np.float64), q, d


def np_meanfill(A):
    """Extension oracle: replace each missing (NaN) cell by mean(skipmissing(column)) — the per-locus mean the
    reference's filterbymaf uses (src/genomes/filter.jl:632-635).  The mean-fill policy of the CUDA path must equal
    the reference GRM functions applied to this matrix."""
    A = np.array(A, dtype=np.float64, order="F", copy=True)
    valid = ~np.isnan(A)
    cnt = valid.sum(axis=0)
    s = np.where(valid, A, 0.0).sum(axis=0)
    fill = s / cnt
    idx = np.where(~valid)
    A[idx] = fill[idx[1]]
    return A


def np_qc_keep(A, maf=0.0, max_locus_sparsity=1.0):
    """Loci the reference's QC filters keep: filterbysparsity (src/genomes/filter.jl:202-213, :361-368:
    mean(ismissing(col)) <= max_locus_sparsity) and filterbymaf (:632-635: maf <= mean(skipmissing(col)) <= 1-maf).
    The masked CUDA call must equal the reference GRM functions applied to A[:, keep]."""
    A = np.asarray(A, dtype=np.float64)
    miss = np.isnan(A)
    sparsity = miss.mean(axis=0)
    cnt = (~miss).sum(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        q = np.where(miss, 0.0, A).sum(axis=0) / cnt
        keep = (sparsity <= max_locus_sparsity) & (q >= maf) & (q <= (1.0 - maf))
    return keep


def np_det_julia(X):
    """det(X) the way Julia evaluates it: LAPACK getrf, sequential product of diag(U), sign from the pivots."""
    from scipy.linalg import lapack

    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    lu, piv, info = lapack.dgetrf(np.asfortranarray(X))
    if info > 0:
        return 0.0
    P = np.float64(1.0)
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        for i in range(n):
            P = P * lu[i, i]
    swaps = int(np.count_nonzero(piv != np.arange(n)))
    return float(-P if swaps % 2 else P)


def np_isposdef(X):
    from scipy.linalg import lapack

    _, info = lapack.dpotrf(np.asfortranarray(X), lower=0)
    return info == 0


def np_inflatediagonals(X, max_iter=1000):
    """calc.jl:41-70 with SciPy's LAPACK.  Returns (status, X_out, iters, eps_total)."""
    X = np.array(X, dtype=np.float64, copy=True)
    n = X.shape[0]
    iu = np.triu_indices(n)
    if not np.all(X[iu] == X.T[iu]):
        return ERR_NOT_SYMMETRIC, X, 0, 0.0
    if np_det_julia(X) > EPS and np_isposdef(X):
        return OK, X, 0, 0.0
    if np.isnan(X).any():
        return ERR_NAN, X, 0, 0.0
    if np.isinf(X).any():
        return ERR_INF, X, 0, 0.0
    it = 0
    e = float(np.max(np.abs(X)))
    tot = 0.0
    di = np.diag_indices(n)
    while ((np_det_julia(X) < EPS) or not np_isposdef(X)) and it < max_iter:
        it += 1
        X[di] = X[di] + e
        tot += e
        e *= 1.0 + EPS
    return OK, X, it, tot


# ---------------------------------------------------------------------------------------------------------------
# distances(genomes) (src/genomes/genomes.jl:454-654), restated pair by pair
# ---------------------------------------------------------------------------------------------------------------
def _pair_metric(metric, y1, y2):
    """genomes.jl:556-569 / :617-631 for one pair restricted to its jointly usable positions; None = `continue`."""
    if metric == "euclidean":
        return float(np.sqrt(np.sum((y1 - y2) ** 2)))
    if metric == "correlation":
        if np.var(y1, ddof=1) < 1e-7 or np.var(y2, ddof=1) < 1e-7:
            return None
        u, v = y1 - np.mean(y1), y2 - np.mean(y2)
        xx, yy, xy = float(np.sum(u * u)), float(np.sum(v * v)), float(np.sum(u * v))
        hi, lo = max(xx, yy), min(xx, yy)
        return float(min(1.0, max(-1.0, xy / hi / np.sqrt(lo / hi))))  # Statistics.corm + clampcor
    if metric == "mad":
        return float(np.mean(np.abs(y1 - y2)))
    if metric == "rmsd":
        return float(np.sqrt(np.mean((y1 - y2) ** 2)))
    if metric == "χ²":
        return float(np.sum((y1 - y2) ** 2 / (y2 + EPS)))
    raise ValueError(metric)


def np_distances(A, idx_loci0, metrics, include_counts=True, dims=("loci_alleles", "entries")):
    """Dict "dimension|metric" -> matrix for the loci idx_loci0 (0-based, sorted, unique) of A (n x p, NaN = missing),
    following the reference's loops: loci pairs i <= j mirrored (counts upper triangle only), entry pairs all (i, j);
    fill +Inf (-Inf for correlation); pairs with fewer than 2 jointly usable positions keep the fill."""
    A = np.asarray(A, dtype=np.float64)
    S = A[:, idx_loci0]
    n, t = S.shape
    ok = np.isfinite(S)
    out = {}
    if t > 1 and "loci_alleles" in dims:
        for m in metrics:
            D = np.full((t, t), -np.inf if m == "correlation" else np.inf)
            cnt = np.zeros((t, t))
            for i in range(t):
                for j in range(i, t):
                    use = ok[:, i] & ok[:, j]
                    if use.sum() < 2:
                        continue
                    cnt[i, j] = use.sum()
                    v = _pair_metric(m, S[use, i], S[use, j])
                    if v is not None:
                        D[i, j] = D[j, i] = v
            out[f"loci_alleles|{m}"] = D
            if include_counts:
                out["loci_alleles|counts"] = cnt
    if n > 1 and "entries" in dims:
        for m in metrics:
            D = np.full((n, n), -np.inf if m == "correlation" else np.inf)
            cnt = np.zeros((n, n))
            for i in range(n):
                for j in range(n):
                    use = ok[i] & ok[j]
                    if use.sum() < 2:
                        continue
                    cnt[i, j] = use.sum()
                    v = _pair_metric(m, S[i, use], S[j, use])
                    if v is not None:
                        D[i, j] = v
            out[f"entries|{m}"] = D
            if include_counts:
                out["entries|counts"] = cnt
    return out
