"""Host half of inflatediagonals! (csrc/repair_host.cpp), no GPU needed.

The repair forms det(X) — Julia's det(lu(X)), the sequential product of the pivots (calc.jl:49,61) — from the diagonal of
the Cholesky factor whenever partial pivoting cannot interchange rows, and falls back to the LU otherwise.  These tests
pin (i) the claim itself against LAPACK (no interchanges, pivots = L_kk^2), (ii) the guards that decide when the
shortcut may be trusted, and (iii) the whole decision flow of csrc/repair.cu, emulated with SciPy's potrf + the C
function, against the oracle's literal LU-first loop on every kind of outcome."""
import ctypes as C

import numpy as np
import pytest
from scipy.linalg import lapack

import grm_b200
from grm_b200 import _lib, synth
from oracle import oracle as O

EPS = np.finfo(np.float64).eps
OK, ILLCOND, BORDERLINE, UNDERFLOW, OVERFLOW = 0, 1, 2, 3, 4


def chol_det(diag_l, scale):
    d = np.ascontiguousarray(diag_l, dtype=np.float64)
    out = C.c_double(0.0)
    reason = _lib.load().grm_cuda_host_chol_det(d.ctypes.data, d.size, float(scale), C.byref(out))
    return reason, out.value


def lu_det(diag_u, ipiv):
    d = np.ascontiguousarray(diag_u, dtype=np.float64)
    ip = np.ascontiguousarray(ipiv, dtype=np.int32)
    return _lib.load().grm_cuda_host_lu_det(d.ctypes.data, ip.ctypes.data, d.size)


def pivot_risk(L):
    """What chol_check_kernel computes: some column with |L_jk| > L_kk (1 - 2^-20) below the diagonal."""
    d = np.diag(L)
    below = np.abs(np.tril(L, -1)).max(axis=0)
    return bool(np.any(~(d > 0.0)) or np.any(below > d * (1.0 - 2.0 ** -20)))


def grm_like(n, p, seed, kind="ploidy"):
    A = synth.dosage(n, p, 4, seed=seed)
    st, G, *_ = (O.c_grmploidyaware_raw(A, 4) if kind == "ploidy" else O.c_grmsimple_raw(A))
    assert st == O.OK
    return G


def test_lu_det_is_julias_sequential_product():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((60, 60))
    lu, piv, info = lapack.dgetrf(np.asfortranarray(X))
    assert info == 0
    assert lu_det(np.diag(lu), piv + 1) == O.np_det_julia(X)
    # under / overflow are part of the semantics: 0.5^1200 = 0, 10^400 = Inf, and Inf * 0 = NaN
    ident = np.arange(1, 1201, dtype=np.int32)
    assert lu_det(np.full(1200, 0.5), ident) == 0.0
    assert lu_det(np.full(400, 10.0), ident[:400]) == np.inf
    assert np.isnan(lu_det(np.r_[np.full(400, 10.0), 0.0], ident[:401]))
    swapped = ident[:3].copy()
    swapped[0] = 2  # one interchange flips the sign
    assert lu_det(np.array([2.0, 3.0, 4.0]), swapped) == -24.0


@pytest.mark.parametrize("n,p,kind", [(200, 3000, "ploidy"), (300, 2000, "simple"), (600, 6000, "ploidy")])
def test_cholesky_pivots_are_the_lu_pivots_when_no_column_can_interchange(n, p, kind):
    G = grm_like(n, p, 11, kind)
    for shift in (0.0, float(np.abs(G).max())):
        X = G + shift * np.eye(n)
        L, info = lapack.dpotrf(np.asfortranarray(X), lower=1)
        assert info == 0
        assert not pivot_risk(np.tril(L))
        lu, piv, info = lapack.dgetrf(np.asfortranarray(X))
        assert np.array_equal(piv, np.arange(n)), "partial pivoting kept every diagonal pivot"
        u = np.diag(lu)
        assert np.max(np.abs(np.diag(L) ** 2 - u) / u) < 1e-9
        reason, det = chol_det(np.diag(L), np.abs(X).max())
        ref = lu_det(u, piv + 1)
        if reason == OK:
            assert (det < EPS) == (ref < EPS) and (det > EPS) == (ref > EPS)
            if 0.0 < ref < np.inf:
                assert abs(det - ref) <= 1e-6 * abs(ref)


def test_guards():
    # decided: far below / far above eps
    r, d = chol_det(np.sqrt(np.full(400, 0.5)), 0.5)
    assert r == OK and 0.0 < d < 1e-100
    assert chol_det(np.sqrt(np.full(400, 1e3)), 1e3) == (OK, np.inf)
    # a product that sticks at the smallest subnormal (pivots in (0.5, 1): SURVEY B.1's 5e-324) or dies (pivots < 0.5)
    r, d = chol_det(np.sqrt(np.full(20000, 0.7)), 0.7)
    assert (r, d) == (OK, 5e-324)
    r, d = chol_det(np.sqrt(np.full(20000, 0.4)), 0.4)
    assert (r, d) == (OK, 0.0)
    # within 2^64 of eps: the LU decides
    assert chol_det(np.sqrt(np.r_[np.full(15, 0.1), 1.0]), 1.0)[0] == BORDERLINE
    assert chol_det(np.sqrt(np.full(10, 0.1)), 0.1)[0] == BORDERLINE
    # an ill-determined pivot (cancellation below 2^-26 of the scale), a NaN, a zero, a non-positive scale
    assert chol_det(np.sqrt(np.r_[np.full(50, 0.5), 1e-9]), 0.5)[0] == ILLCOND
    assert chol_det(np.array([1.0, np.nan, 1.0]), 1.0)[0] == ILLCOND
    assert chol_det(np.array([1.0, 0.0, 1.0]), 1.0)[0] == ILLCOND
    assert chol_det(np.array([1.0, 1.0]), 0.0)[0] == ILLCOND
    # through the subnormals and back up: a running product forgets its relative accuracy there
    back = np.r_[np.full(1100, 0.5), np.full(1100, 2.0) * 1.0]
    r, d = chol_det(np.sqrt(back), 2.0)
    assert r == UNDERFLOW
    # ... but a modest recovery that cannot reach eps stays decided
    r, d = chol_det(np.sqrt(np.r_[np.full(1200, 0.5), np.full(100, 2.0)]), 2.0)
    assert r == OK and d < EPS
    # up to the overflow threshold and back down
    r, d = chol_det(np.sqrt(np.r_[np.full(1030, 2.0), np.full(1100, 0.5)]), 2.0)
    assert r == OVERFLOW
    r, d = chol_det(np.sqrt(np.r_[np.full(1030, 2.0), np.full(100, 0.5)]), 2.0)
    assert r == OK and d == np.inf
    # high, then all the way down: nearby trajectories may have overflowed
    r, d = chol_det(np.sqrt(np.r_[np.full(1022, 2.0), np.full(2300, 0.5)]), 2.0)
    assert r in (UNDERFLOW, OVERFLOW)
    lib = _lib.load()
    assert lib.grm_cuda_host_chol_det(None, 3, 1.0, C.byref(C.c_double())) == -1


def emulate_device_repair(X, max_iter=1000):
    """csrc/repair.cu's flow with SciPy's LAPACK in place of cuSOLVER: Cholesky first; the determinant only for a
    positive definite matrix, from the factor's diagonal when the guards allow, else by LU.  -> (iters, eps_total, LUs)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    maxabs = float(np.max(np.abs(X)))
    lus = 0

    def shifted(t):
        Y = np.array(X, order="F", copy=True)
        d, e = np.diag(Y).copy(), maxabs
        for _ in range(t):
            d = d + e
            e = e * (1.0 + EPS)
        Y[np.diag_indices(n)] = d
        return Y

    def evaluate(t):
        nonlocal lus
        Y = shifted(t)
        L, info = lapack.dpotrf(Y, lower=1)
        if info > 0:
            return False, float("nan")
        L = np.tril(L)
        if not pivot_risk(L):
            reason, det = chol_det(np.diag(L), maxabs * (1.0 + t))
            if reason == OK:
                return True, det
        lus += 1
        return True, O.np_det_julia(Y)

    pd, det = evaluate(0)
    if pd and det > EPS:
        return 0, 0.0, lus
    t = 0
    while (det < EPS or not pd) and t < max_iter:
        t += 1
        pd, det = evaluate(t)
    tot, e = 0.0, maxabs
    for _ in range(t):
        tot += e
        e *= 1.0 + EPS
    return t, tot, lus


def test_cholesky_first_flow_takes_the_reference_decisions():
    rng = np.random.default_rng(5)
    W = rng.standard_normal((300, 20))
    cases = {
        "ploidy-aware, det underflows": grm_like(600, 6000, 3),
        "grmsimple, several rounds": grm_like(300, 2000, 4, "simple"),
        "small, det in the normal range": grm_like(100, 10_000, 5),
        "continuous": O.c_grmsimple_raw(synth.continuous(300, 2000, seed=7))[1],
        "0.5 I": np.eye(400) * 0.5,
        "1e3 I": np.eye(400) * 1e3,
        "1e-3 I": np.eye(400) * 1e-3,
        "indefinite": np.eye(200) - 2.0 * np.diag((np.arange(200) % 7 == 0).astype(float)),
        "low rank": W @ W.T / 20.0,
        "-I": -np.eye(4),
        "zero": np.zeros((3, 3)),
        # positive definite, but the second column WOULD interchange under partial pivoting: the LU must decide
        "needs pivoting": np.array([[1.0, 0.9, 0.0], [0.9, 1.0, 0.0], [0.0, 0.0, 1e-17]]) + np.diag([0.0, 0.0, 1e-3]),
    }
    saw_lu = saw_chol_only = False
    for name, X in cases.items():
        for max_iter in (1000, 1, 0):
            st, Xo, it, tot = O.np_inflatediagonals(X, max_iter)
            assert st == O.OK
            it2, tot2, lus = emulate_device_repair(X, max_iter)
            assert (it2, tot2) == (it, tot), (name, max_iter, it, it2, lus)
            saw_lu |= lus > 0
            saw_chol_only |= lus == 0 and it > 0
    assert saw_lu and saw_chol_only  # both branches were exercised


def test_interchange_is_detected():
    # SPD with a sub-diagonal entry larger than the diagonal of its Schur complement column: LAPACK interchanges rows
    X = np.array([[1.0, 2.0, 0.0], [2.0, 5.0, 0.0], [0.0, 0.0, 1.0]])
    L, info = lapack.dpotrf(np.asfortranarray(X), lower=1)
    assert info == 0 and pivot_risk(np.tril(L))
    lu, piv, _ = lapack.dgetrf(np.asfortranarray(X))
    assert not np.array_equal(piv, np.arange(3))
