"""CPU: the oracle against the committed golden vectors, its two restatements against each other, and the
reference's own doctest properties for the GRM path (src/grm/calc.jl:30-39, :103-113, :177-187)."""
import numpy as np
import pytest

from grm_b200 import synth
from oracle import oracle as O

CASES = ["dip", "tet", "dip_as_tet", "cont", "unif", "wide"]
EPS = np.finfo(np.float64).eps


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_reproduces_golden(golden, name):
    A = np.asfortranarray(golden[f"{name}_A"])
    ploidy = int(golden[f"{name}_ploidy"])
    st, Gs = O.c_grmsimple_raw(A)
    assert st == O.OK and np.array_equal(Gs, golden[f"{name}_simple_raw"])
    st, Gp, q, d = O.c_grmploidyaware_raw(A, ploidy)
    assert st == O.OK and np.array_equal(Gp, golden[f"{name}_ploidy_raw"])
    assert np.array_equal(q, golden[f"{name}_q"]) and d == float(golden[f"{name}_d"])
    st, F, it, tot = O.c_grmsimple(A)
    assert np.array_equal(F, golden[f"{name}_simple_final"])
    assert [it, tot] == list(golden[f"{name}_simple_repair"])
    st, F, it, tot = O.c_grmploidyaware(A, ploidy)
    assert np.array_equal(F, golden[f"{name}_ploidy_final"])
    assert [it, tot] == list(golden[f"{name}_ploidy_repair"])


@pytest.mark.parametrize("name", CASES)
def test_numpy_restatement_agrees_with_golden(golden, name):
    A = golden[f"{name}_A"]
    ploidy = int(golden[f"{name}_ploidy"])
    # sequential-dot restatement: bit-identical to the C loop
    assert np.array_equal(O.np_grmsimple_raw_seq(A), golden[f"{name}_simple_raw"])
    # extended-precision truth: both sides within a few ulp * sqrt(p)
    T = O.np_truth_simple(A)
    assert np.abs(golden[f"{name}_simple_raw"] - T).max() <= 1e-13 * np.abs(T).max()
    T, q, d = O.np_truth_ploidyaware(A, ploidy)
    assert np.abs(golden[f"{name}_ploidy_raw"] - T).max() <= 1e-12 * np.abs(T).max()
    assert np.allclose(q, golden[f"{name}_q"], rtol=0, atol=4 * EPS)
    assert abs(d - float(golden[f"{name}_d"])) <= 1e-14 * abs(d)
    # repair: SciPy's LAPACK (what Julia calls) and the oracle's own LU/Cholesky take the same decisions
    for kind in ("simple", "ploidy"):
        st, X, it, tot = O.np_inflatediagonals(golden[f"{name}_{kind}_raw"])
        assert st == O.OK
        assert [it, tot] == list(golden[f"{name}_{kind}_repair"])
        assert np.array_equal(X, golden[f"{name}_{kind}_final"])


@pytest.mark.parametrize("name,m", [("dip", 2), ("tet", 4), ("dip_as_tet", 2), ("wide", 4)])
def test_exactness_lemma_for_power_of_two_dosages(golden, name, m):
    """for x = c/m the reference's sequential FP64 dot equals C_int/m^2 exactly."""
    A = golden[f"{name}_A"]
    assert np.array_equal(O.np_grmsimple_raw_dosage(A, m), golden[f"{name}_simple_raw"])
    c = O.dosage_codes(A, m)
    assert np.array_equal(O.c_gram_i8(c.astype(np.int8)), O.np_gram_int(c))


@pytest.mark.parametrize("gen", ["continuous", "dosage"])
@pytest.mark.parametrize("p", [1000])
def test_reference_doctest_properties(gen, p):
    """calc.jl:103-113 / :177-187: size == (100,100), issymmetric, det > eps after the repair, on
    simulategenomes(l=1_000)-like input (n = 100 is the simulator's default)."""
    A = synth.continuous(100, p, seed=42) if gen == "continuous" else synth.dosage(100, p, 2, seed=42)
    for st, G, it, tot in (O.c_grmsimple(A, nthreads=4), O.c_grmploidyaware(A, 2, nthreads=4)):
        assert st == O.OK
        assert G.shape == (100, 100)
        assert np.array_equal(G, G.T)
        assert O.np_det_julia(G) > EPS
        assert np.linalg.det(G) > EPS


def test_inflatediagonals_docstring_example(golden):
    """calc.jl:30-39: x = rand(10); X = x*x'; inflatediagonals!(X); det(X) > eps."""
    X = golden["rank1_X"]
    st, Xo, it, tot = O.c_inflatediagonals(X)
    assert st == O.OK and it >= 1
    assert np.array_equal(Xo, golden["rank1_out"]) and [it, tot] == list(golden["rank1_repair"])
    assert O.np_det_julia(Xo) > EPS
    off = ~np.eye(10, dtype=bool)
    assert np.array_equal(Xo[off], X[off])  # only the diagonal moves
    st2, X2, it2, tot2 = O.np_inflatediagonals(X)
    assert (st2, it2, tot2) == (st, it, tot) and np.array_equal(X2, Xo)


def test_inflatediagonals_error_order_and_max_iter():
    rng = np.random.default_rng(5)
    B = rng.random((6, 6))
    for fn in (O.c_inflatediagonals, O.np_inflatediagonals):
        assert fn(B)[0] == O.ERR_NOT_SYMMETRIC  # calc.jl:43-45
        S = B + B.T
        S[2, 3] = S[3, 2] = np.nan
        assert fn(S)[0] == O.ERR_NOT_SYMMETRIC  # NaN != NaN: issymmetric fails before the NaN check
        S = -(B @ B.T)
        S[1, 4] = S[4, 1] = np.inf
        assert fn(S)[0] == O.ERR_INF  # calc.jl:55-57
        S = -np.eye(4)
        st, X, it, tot = fn(S, max_iter=0)  # loop never runs: matrix returned unchanged, no error (calc.jl:61)
        assert st == O.OK and it == 0 and np.array_equal(X, S)
        st, X, it, tot = fn(S, max_iter=1000)
        assert st == O.OK and it >= 1 and O.np_isposdef(X)


def test_det_under_and_overflow_follow_the_sequential_product():
    """det saturates at 0 / Inf for large n; the repair decision follows that, not log-det."""
    n = 400
    tiny = np.eye(n) * 1e-3  # product underflows to 0 although the matrix is SPD
    assert O.c_det(tiny) == 0.0 and O.np_det_julia(tiny) == 0.0
    half = np.eye(n) * 0.5  # 0.5^400 = 3.9e-121 < eps although SPD; one round of +max|X| makes it the identity
    for fn in (O.c_inflatediagonals, O.np_inflatediagonals):
        st, X, it, tot = fn(half)
        assert st == O.OK and it == 1 and tot == 0.5 and np.array_equal(X, np.eye(n))
    big = np.eye(n) * 1e3
    assert np.isinf(O.c_det(big)) and np.isinf(O.np_det_julia(big))
    st, X, it, tot = O.c_inflatediagonals(big)
    assert it == 0  # Inf > eps and SPD: returned untouched


def test_missing_values_are_an_error():
    A = synth.add_missing(synth.dosage(8, 30, 2, seed=3), 0.05, seed=4)
    assert np.isnan(A).any()
    assert O.c_grmsimple_raw(A)[0] == O.ERR_MISSING
    assert O.c_grmploidyaware_raw(A, 2)[0] == O.ERR_MISSING


def test_threaded_oracle_equals_single_thread():
    A = synth.continuous(37, 211, seed=9)
    assert np.array_equal(O.c_grmsimple_raw(A, nthreads=1)[1], O.c_grmsimple_raw(A, nthreads=5)[1])
    assert np.array_equal(O.c_grmploidyaware_raw(A, 4, nthreads=1)[1], O.c_grmploidyaware_raw(A, 4, nthreads=3)[1])


def test_qc_keep_and_meanfill_follow_the_filter_rules():
    """np_qc_keep / np_meanfill restate filterbysparsity (src/genomes/filter.jl:202-213, :361-368) and filterbymaf
    (:632-635); they are the oracles of the mask / mean-fill extensions."""
    A = np.array([[0.0, 0.5, np.nan, 1.0, np.nan],
                  [0.0, 1.0, 0.5, 1.0, np.nan],
                  [0.0, 0.0, 0.5, 1.0, np.nan],
                  [0.0, 0.5, 1.0, np.nan, np.nan]])
    # column means over valid cells: 0, 0.5, 2/3, 1, NaN ; sparsities: 0, 0, .25, .25, 1
    assert O.np_qc_keep(A, 0.0, 1.0).tolist() == [True, True, True, True, False]  # all-missing locus never passes
    assert O.np_qc_keep(A, 0.01, 1.0).tolist() == [False, True, True, False, False]
    assert O.np_qc_keep(A, 0.01, 0.0).tolist() == [False, True, False, False, False]
    assert O.np_qc_keep(A, 0.0, 0.25).tolist() == [True, True, True, True, False]
    assert O.np_qc_keep(A, 0.0, 0.2).tolist() == [True, True, False, False, False]
    F = O.np_meanfill(A[:, :4])
    assert not np.isnan(F).any()
    assert F[0, 2] == (0.5 + 0.5 + 1.0) / 3 and F[3, 3] == 1.0
    assert np.array_equal(F[~np.isnan(A[:, :4])], A[:, :4][~np.isnan(A[:, :4])])


def test_distances_restatement_reproduces_the_reference_doctest():
    """src/genomes/genomes.jl:440-452: keys of the returned Dict, correlation diagonal == 1, χ² diagonal == 0."""
    from grm_b200 import synth

    A = synth.continuous(100, 1000, seed=42)
    idx = np.sort(np.random.default_rng(1).choice(1000, 100, replace=False))
    d = O.np_distances(A, idx, ["correlation", "χ²"])
    assert sorted(d) == ["entries|correlation", "entries|counts", "entries|χ²", "loci_alleles|correlation",
                         "loci_alleles|counts", "loci_alleles|χ²"]
    assert np.array_equal(np.diag(d["entries|correlation"]), np.ones(100))
    assert np.array_equal(np.diag(d["loci_alleles|χ²"]), np.zeros(100))
    # pairs with fewer than 2 jointly usable positions keep the fill; a constant vector has no correlation
    B = A[:6, :8].copy()
    B[:, 3] = 0.5
    B[0, :] = np.nan
    B[1, 1:] = np.nan
    e = O.np_distances(B, np.arange(8), ["correlation", "euclidean"])
    assert np.all(np.isneginf(e["entries|correlation"][0])) and np.all(np.isposinf(e["entries|euclidean"][0]))
    assert np.isposinf(e["entries|euclidean"][1, 2]) and e["entries|counts"][1, 2] == 0
    assert np.all(np.isneginf(e["loci_alleles|correlation"][3]))  # the constant locus
    assert np.isfinite(e["loci_alleles|euclidean"][3, 4])
