"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs,
against the committed golden vectors, and — at BASELINE.json's full sizes — through size-independent properties.

Bars (north_star): integer X X' bit-exact; final Float64 GRM within 1e-12 relative on the int8 path and 1e-10 on the
FP64 path (relative to max |G|); for power-of-two dosages grmsimple is bit-identical.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import grm_b200
from grm_b200 import (ArgumentError, Genomes, MissingDataError, _lib, grm_call, grmploidyaware, grmsimple,
                      inflatediagonals, synth)
from oracle import oracle as O

RTOL_I8 = 1e-12   # north_star: final GRM within 1e-12 relative (int8 path)
RTOL_F64 = 1e-10  # north_star: 1e-10 relative (FP64 path)
EPS = np.finfo(np.float64).eps


def relerr(G, ref):
    return np.abs(G - ref).max() / np.abs(ref).max()


def dev():
    import torch

    return torch.device("cuda", 0)


def to_dev(A):
    import torch

    return torch.from_numpy(np.ascontiguousarray(np.asarray(A).T)).to(dev())


# ---------------------------------------------------------------------------------------------------------------
# golden vectors
# ---------------------------------------------------------------------------------------------------------------
DOSAGE_CASES = ["dip", "tet", "dip_as_tet", "wide"]
CONT_CASES = ["cont", "unif"]


@pytest.mark.parametrize("name", DOSAGE_CASES + CONT_CASES)
def test_golden_raw_and_final(golden, name):
    A = golden[f"{name}_A"]
    ploidy = int(golden[f"{name}_ploidy"])
    is_dosage = name in DOSAGE_CASES
    tol = RTOL_I8 if is_dosage else RTOL_F64
    G, info = grm_call(A, skip_repair=True)
    assert info["path"] == (_lib.PATH_I8 if is_dosage else _lib.PATH_F64)
    ref = golden[f"{name}_simple_raw"]
    if is_dosage:
        assert np.array_equal(G, ref), "grmsimple on power-of-two dosages must be bit-identical"
    assert relerr(G, ref) <= tol
    assert np.array_equal(G, G.T)
    G, info = grm_call(A, ploidy, skip_repair=True)
    assert relerr(G, golden[f"{name}_ploidy_raw"]) <= tol
    assert abs(info["scale"] - float(golden[f"{name}_d"])) <= 1e-14 * info["scale"]
    assert np.array_equal(G, G.T)
    # with the repair: same number of rounds, same total, same matrix
    for pl, kind in ((None, "simple"), (ploidy, "ploidy")):
        G, info = grm_call(A, pl)
        it, tot = golden[f"{name}_{kind}_repair"]
        assert info["repair_iters"] == int(it)
        assert abs(info["eps_total"] - tot) <= tol * max(1.0, abs(tot))
        assert relerr(G, golden[f"{name}_{kind}_final"]) <= tol
        assert np.array_equal(G, G.T)


def test_golden_inflatediagonals_docstring_example(golden):
    X = np.array(golden["rank1_X"], order="F")
    inflatediagonals(X)
    it, tot = golden["rank1_repair"]
    assert relerr(X, golden["rank1_out"]) <= 1e-15
    assert O.np_det_julia(X) > EPS  # calc.jl:37-38


# ---------------------------------------------------------------------------------------------------------------
# P2/P3: integer Gram exactness, ragged shapes (non-multiples of the 128/256 tiles and of K = 32 / 128)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,p,m", [(2, 2, 1), (2, 33, 2), (100, 10_000, 2), (257, 33, 4), (257, 10_000, 4),
                                   (300, 4097, 2), (513, 129, 64), (1, 17, 2), (1000, 2, 4)])
def test_integer_gram_bit_exact(ctx, n, p, m):
    from grm_b200 import device as D

    A = synth.dosage(n, p, m, seed=1000 + n + p)
    codes_ref = O.dosage_codes(A, m)
    At = to_dev(A)
    assert D.detect_denominator(ctx, At) in {1, 2, 4, 8, 16, 32, 64} and m % D.detect_denominator(ctx, At) == 0
    codes, colsum, flags = D.pack_dosage(ctx, At, m)
    C = D.syrk_i8(ctx, codes, p)
    ctx.synchronize()
    assert int(flags[0]) == 0
    assert np.array_equal(codes.cpu().numpy()[:, :p], codes_ref.astype(np.int8))
    assert not codes.cpu().numpy()[:, p:].any()
    assert np.array_equal(colsum.cpu().numpy()[:p], codes_ref.sum(0))
    ref = O.np_gram_int(codes_ref)
    got = C.cpu().numpy().T.astype(np.int64)  # [i, j] = C[i + j*n]
    iu = np.triu_indices(n)
    assert np.array_equal(got[iu], ref[iu])
    # P3: final grmsimple through the C ABI is bit-identical to the sequential-dot reference
    G, info = grm_call(A, skip_repair=True)
    assert info["path"] == _lib.PATH_I8
    assert np.array_equal(G, (ref.astype(np.float64) / (m * m)) / p)
    if n * p <= 3_000_000:
        st, Gc = O.c_grmsimple_raw(A, nthreads=4)
        assert st == O.OK and np.array_equal(G, Gc)


# P4: grmploidyaware on dosages, including ploidy != denominator
@pytest.mark.parametrize("n,p,m,ploidy", [(100, 10_000, 2, 2), (257, 3000, 4, 4), (64, 5000, 2, 4), (130, 777, 4, 2),
                                          (50, 400, 2, 6)])
def test_ploidyaware_dosage(ctx, n, p, m, ploidy):
    from grm_b200 import device as D

    A = synth.dosage(n, p, m, seed=2000 + n)
    st, Gc, qc, dc = O.c_grmploidyaware_raw(A, ploidy, nthreads=4)
    G, info = grm_call(A, ploidy, skip_repair=True)
    assert info["path"] == _lib.PATH_I8
    assert relerr(G, Gc) <= RTOL_I8
    T, _, _ = O.np_truth_ploidyaware(A, ploidy)
    assert relerr(G, T) <= RTOL_I8
    assert abs(info["scale"] - dc) <= 1e-14 * dc
    # q bit-exact (column sums of power-of-two dosages are exact in any order)
    At = to_dev(A)
    codes, colsum, flags = D.pack_dosage(ctx, At, info["denominator"])
    q, w, stats = D.locus_stats_i8(ctx, colsum, p, n, info["denominator"], ploidy)
    ctx.synchronize()
    assert np.array_equal(q.cpu().numpy(), qc)


def test_epilogue_division_is_correctly_rounded(ctx):
    """The epilogue divides by a launch constant with a reciprocal + two FMA corrections; it must equal IEEE `/`."""
    from grm_b200 import device as D

    n, p, m = 513, 1900, 4
    A = synth.dosage(n, p, m, seed=41)
    c = O.dosage_codes(A, m)
    Cint = O.np_gram_int(c).astype(np.float64)
    codes, colsum, flags = D.pack_dosage(ctx, to_dev(A), m)
    iu = np.triu_indices(n)
    for alpha, denom in [(1.0, 3.0), (1.0 / 16, float(p)), (1.0, 99991.0), (0.25, 100000.0), (1.0, np.pi), (1.0, 1.0 / 3),
                         (1.0, 7e-3), (1.0, 2.0 ** 31 - 1), (1.0, -1237.0)]:
        G = D.syrk_i8_f64(ctx, codes, p, alpha, 0.0, 0.0, denom)
        ctx.synchronize()
        want = (Cint * alpha) / denom
        assert np.array_equal(G.cpu().numpy().T[iu], want[iu]), (alpha, denom)
    # rank-1 mode with arbitrary FP64 terms
    rng = np.random.default_rng(3)
    import torch

    u = rng.standard_normal(n) * 1e3
    for beta, gamma, denom in [(0.5, 123.456, 24944.479067879984), (1.0, -7.0, 0.37)]:
        G = D.syrk_i8_f64(ctx, codes, p, 0.25, beta, gamma, denom, u=torch.from_numpy(u).cuda())
        ctx.synchronize()
        want = ((Cint * 0.25 - beta * (u[:, None] + u[None, :])) + gamma) / denom
        assert np.array_equal(G.cpu().numpy().T[iu], want[iu]), (beta, gamma, denom)


def test_integer_centring_statistics(ctx):
    """r_i, T_i, sum S, sum S^2 are exact integers; u, W2 and d follow from them."""
    from grm_b200 import device as D

    n, p, m, ploidy = 300, 4097, 4, 4
    A = synth.dosage(n, p, m, seed=31)
    c = O.dosage_codes(A, m)
    S = c.sum(0)
    codes, colsum, flags = D.pack_dosage(ctx, to_dev(A), m)
    sums, r, T = D.dosage_stats(ctx, codes, p, colsum)
    u, (alpha, beta, gamma, denom) = D.dosage_centre(ctx, sums, r, T, p, m, ploidy)
    ctx.synchronize()
    assert np.array_equal(r.cpu().numpy(), c.sum(1))
    assert np.array_equal(T.cpu().numpy(), c @ S)
    assert sums.cpu().numpy().tolist() == [int(S.sum()), int((S * S).sum())]
    q = S / (m * n)
    w = ploidy / 2 + q
    assert np.allclose(u.cpu().numpy(), c @ w, rtol=1e-14, atol=0)
    assert (alpha, beta) == ((ploidy / m) ** 2, ploidy / m)
    assert abs(gamma - float((w * w).sum())) <= 1e-14 * gamma
    assert abs(denom - ploidy * float((q * (1 - q)).sum())) <= 1e-14 * denom
    # accumulate over two slabs == one pass
    h = 2048
    c1, cs1, _ = D.pack_dosage(ctx, to_dev(A[:, :h]), m)
    c2, cs2, _ = D.pack_dosage(ctx, to_dev(A[:, h:]), m)
    s2, r2, T2 = D.dosage_stats(ctx, c1, h, cs1)
    s2, r2, T2 = D.dosage_stats(ctx, c2, p - h, cs2, s2, r2, T2, accumulate=True)
    ctx.synchronize()
    import torch

    assert torch.equal(s2, sums) and torch.equal(r2, r) and torch.equal(T2, T)


# P5: continuous FP64 path
@pytest.mark.parametrize("n,p,gen", [(100, 10_000, "continuous"), (257, 1000, "continuous"), (123, 5000, "uniform"),
                                     (130, 17, "continuous")])
def test_fp64_path(n, p, gen):
    A = getattr(synth, gen)(n, p, seed=3000 + n)
    G, info = grm_call(A, skip_repair=True)
    assert info["path"] == _lib.PATH_F64
    st, Gc = O.c_grmsimple_raw(A, nthreads=4)
    assert relerr(G, Gc) <= RTOL_F64
    assert relerr(G, O.np_truth_simple(A)) <= 1e-13
    for ploidy in (2, 4):
        G, info = grm_call(A, ploidy, skip_repair=True)
        st, Gc, qc, dc = O.c_grmploidyaware_raw(A, ploidy, nthreads=4)
        assert relerr(G, Gc) <= RTOL_F64
        assert relerr(G, O.np_truth_ploidyaware(A, ploidy)[0]) <= 1e-12
        assert abs(info["scale"] - dc) <= 1e-13 * dc
        assert np.array_equal(G, G.T)


# P1: the reference's doctest properties through the reference-shaped API
@pytest.mark.parametrize("gen", ["continuous", "dosage"])
def test_reference_doctest_properties(gen):
    A = synth.continuous(100, 1000, seed=42) if gen == "continuous" else synth.dosage(100, 1000, 2, seed=42)
    g = Genomes.from_matrix(A)
    for grm in (grmsimple(g), grmploidyaware(g)):
        G = grm.genomic_relationship_matrix
        assert G.shape == (100, 100) and np.array_equal(G, G.T)  # calc.jl:108-109
        assert O.np_det_julia(G) > EPS  # calc.jl:111-112
        assert grm.entries is g.entries and grm.loci_alleles is g.loci_alleles  # aliased, calc.jl:123


# P6: repair decisions incl. det under/overflow
def test_repair_matches_oracle_on_config_like_inputs():
    for gen, n, p in (("continuous", 300, 2000), ("dosage", 300, 2000), ("continuous", 600, 3000)):
        A = synth.continuous(n, p, seed=7) if gen == "continuous" else synth.dosage(n, p, 2, seed=7)
        for ploidy in (None, 2):
            Graw, _ = grm_call(A, ploidy, skip_repair=True)
            st, X, it, tot = O.np_inflatediagonals(Graw)
            G, info = grm_call(A, ploidy)
            assert info["repair_iters"] == it
            assert abs(info["eps_total"] - tot) <= 1e-15 * max(tot, 1.0)
            assert np.array_equal(G, X), "same rounds on the same raw matrix => identical diagonal, untouched rest"


def test_repair_edge_cases():
    n = 400
    half = np.asfortranarray(np.eye(n) * 0.5)  # SPD but det underflows below eps: one round
    inflatediagonals(half)
    assert np.array_equal(half, np.eye(n))
    big = np.asfortranarray(np.eye(n) * 1e3)  # det overflows to Inf > eps: untouched
    inflatediagonals(big)
    assert np.array_equal(big, np.eye(n) * 1e3)
    rng = np.random.default_rng(5)
    B = rng.random((6, 6))
    with pytest.raises(ArgumentError, match="not symmetric"):
        inflatediagonals(np.asfortranarray(B))
    S = B + B.T
    S[2, 3] = S[3, 2] = np.nan
    with pytest.raises(ArgumentError, match="not symmetric"):  # NaN != NaN, calc.jl:43
        inflatediagonals(np.asfortranarray(S))
    S = -(B @ B.T)
    S[1, 4] = S[4, 1] = np.inf
    with pytest.raises(ArgumentError, match="Inf"):
        inflatediagonals(np.asfortranarray(S))
    S = np.asfortranarray(-np.eye(4))
    inflatediagonals(S, max_iter=0)
    assert np.array_equal(S, -np.eye(4))
    inflatediagonals(S)
    assert O.np_isposdef(S)
    one = np.array([[0.0]])
    inflatediagonals(one)  # ϵ = max|X| = 0: the loop adds 0 until max_iter; no error (calc.jl:61)
    assert one[0, 0] == 0.0


# P7: error paths
def test_error_paths():
    A = synth.dosage(40, 300, 2, seed=11)
    g = Genomes.from_matrix(synth.add_missing(A, 0.01, seed=1))
    for fn in (grmsimple, grmploidyaware):
        with pytest.raises(MissingDataError):
            fn(g)
    Ac = synth.add_missing(synth.continuous(40, 300, seed=11), 0.01, seed=1)
    with pytest.raises(MissingDataError):
        grmsimple(Genomes.from_matrix(Ac))
    Ai = np.array(A, order="F")
    Ai[3, 5] = np.inf
    with pytest.raises(ArgumentError):
        grmsimple(Genomes.from_matrix(Ai))
    # all loci fixed: q(1-q) = 0 => d = 0 => NaN matrix => the reference's issymmetric check throws (calc.jl:43)
    fixed = np.ones((5, 40), order="F")
    with pytest.raises(ArgumentError, match="not symmetric"):
        grmploidyaware(Genomes.from_matrix(fixed))
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call(synth.continuous(10, 50, seed=1), path=_lib.PATH_I8)
    assert e.value.status == _lib.ERR_NOT_DOSAGE
    # n = 1
    G, info = grm_call(synth.dosage(1, 50, 2, seed=2), skip_repair=True)
    assert G.shape == (1, 1)


# Extension (BASELINE config 4: pool-seq with missing cells): mean-fill policy == reference GRM of the imputed matrix
@pytest.mark.parametrize("gen,sparsity", [("continuous", 0.01), ("continuous", 0.2), ("dosage", 0.05)])
def test_meanfill_policy_matches_reference_on_mean_imputed_matrix(gen, sparsity):
    base = synth.continuous(150, 2000, seed=51) if gen == "continuous" else synth.dosage(150, 2000, 2, seed=51)
    A = synth.add_missing(base, sparsity, seed=52)
    F = O.np_meanfill(A)
    assert np.isnan(A).any() and not np.isnan(F).any()
    for ploidy in (None, 2):
        with pytest.raises(_lib.GrmCudaError) as e:  # default policy = the reference's behaviour: throw
            grm_call(A, ploidy, skip_repair=True)
        assert e.value.status == _lib.ERR_MISSING
        G, info = grm_call(A, ploidy, skip_repair=True, missing="meanfill")
        assert info["path"] == _lib.PATH_F64
        ref = O.c_grmsimple_raw(F, nthreads=4)[1] if ploidy is None else O.c_grmploidyaware_raw(F, ploidy, nthreads=4)[1]
        assert relerr(G, ref) <= RTOL_F64
        assert np.array_equal(G, G.T)
        G2, _ = grm_call(A, ploidy, skip_repair=True, missing="meanfill", chunk_loci=256)
        assert relerr(G2, G) <= 1e-13
        Gf, infof = grm_call(A, ploidy, missing="meanfill")  # with the repair
        st, X, it, tot = O.np_inflatediagonals(G)
        assert infof["repair_iters"] == it and np.array_equal(Gf, X)
    # a locus with no valid cell has no defined fill
    B = np.array(A, order="F")
    B[:, 7] = np.nan
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call(B, 2, skip_repair=True, missing="meanfill")
    assert e.value.status == _lib.ERR_MISSING
    # without missing cells the policy changes nothing (dosages stay on the int8 path)
    if gen == "dosage":
        G0, i0 = grm_call(base, 2, skip_repair=True)
        G1, i1 = grm_call(base, 2, skip_repair=True, missing="meanfill")
        assert i1["path"] == _lib.PATH_I8 and np.array_equal(G0, G1)


# Row (f1): fused QC locus mask == reference GRM of slice(genomes, idx_loci_alleles = kept)
def _qc_input(kind, n=140, p=1500):
    rng = np.random.default_rng(71)
    A = synth.dosage(n, p, 2, seed=72) if kind == "dosage" else synth.continuous(n, p, seed=72)
    A[:, rng.choice(p, 40, replace=False)] = 0.0   # fixed loci: fail the MAF rule
    A[:, rng.choice(p, 25, replace=False)] = 1.0
    for k in rng.choice(p, 60, replace=False):      # loci with missing cells of varying sparsity
        A[rng.choice(n, rng.integers(1, n // 3), replace=False), k] = np.nan
    A[:, 5] = np.nan                                 # one locus without any valid cell
    return np.asfortranarray(A)


@pytest.mark.parametrize("kind", ["dosage", "continuous"])
def test_qc_locus_mask_equals_reference_on_sliced_genomes(kind):
    A = _qc_input(kind)
    tol = RTOL_I8 if kind == "dosage" else RTOL_F64
    # (1) drop every locus with a missing cell (max_locus_sparsity = 0) and rare alleles: no missing data remains
    keep = O.np_qc_keep(A, maf=0.05, max_locus_sparsity=0.0)
    assert 0 < keep.sum() < A.shape[1] and not np.isnan(A[:, keep]).any()
    S = np.asfortranarray(A[:, keep])
    for ploidy in (None, 2):
        G, info = grm_call(A, ploidy, skip_repair=True, maf=0.05, max_locus_sparsity=0.0)
        assert info["loci_kept"] == int(keep.sum())
        assert info["path"] == (_lib.PATH_I8 if kind == "dosage" else _lib.PATH_F64)
        ref = O.c_grmsimple_raw(S, nthreads=4)[1] if ploidy is None else O.c_grmploidyaware_raw(S, ploidy, nthreads=4)[1]
        assert relerr(G, ref) <= tol
        if kind == "dosage" and ploidy is None:
            assert np.array_equal(G, ref)
        G2, i2 = grm_call(A, ploidy, skip_repair=True, maf=0.05, max_locus_sparsity=0.0, chunk_loci=256)
        assert i2["loci_kept"] == info["loci_kept"] and relerr(G2, G) <= 1e-13
        Gr, ir = grm_call(A, ploidy, maf=0.05, max_locus_sparsity=0.0)  # with the repair
        st, X, it, tot = O.np_inflatediagonals(G)
        assert ir["repair_iters"] == it and np.array_equal(Gr, X)
    # (2) tolerate up to 10 % missing per locus: kept loci still hold missing cells
    keep = O.np_qc_keep(A, maf=0.05, max_locus_sparsity=0.1)
    assert np.isnan(A[:, keep]).any()
    with pytest.raises(_lib.GrmCudaError) as e:  # reference behaviour on the sliced genomes: missing => throw
        grm_call(A, 2, skip_repair=True, maf=0.05, max_locus_sparsity=0.1)
    assert e.value.status == _lib.ERR_MISSING
    G, info = grm_call(A, 2, skip_repair=True, maf=0.05, max_locus_sparsity=0.1, missing="meanfill")
    assert info["loci_kept"] == int(keep.sum()) and info["path"] == _lib.PATH_F64
    ref = O.c_grmploidyaware_raw(O.np_meanfill(np.asfortranarray(A[:, keep])), 2, nthreads=4)[1]
    assert relerr(G, ref) <= RTOL_F64
    # (3) everything filtered: the reference's filters throw ErrorException
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call(np.asfortranarray(np.ones((6, 40))), 2, maf=0.01)
    assert e.value.status == _lib.ERR_ALL_FILTERED


def test_qc_mask_kernel_matches_filter_rules(ctx):
    from grm_b200 import device as D

    A = _qc_input("continuous")
    for maf, tau in [(0.01, 1.0), (0.05, 0.0), (0.2, 0.1), (0.0, 0.05)]:
        q, keep, kept, stats, flags = D.locus_qc(ctx, to_dev(A), maf, tau, error_on_missing=False)
        ctx.synchronize()
        want = O.np_qc_keep(A, maf, tau)
        assert np.array_equal(keep.cpu().numpy().astype(bool), want)
        assert int(kept[0]) == int(want.sum())


def test_context_reuse_across_sizes_and_bad_arguments():
    """One context, grow-only workspaces: small -> large -> small calls stay correct; a second context is independent;
    argument errors come back as statuses."""
    ctx1, ctx2 = grm_b200.Context(0), grm_b200.Context(0)
    small = synth.dosage(50, 300, 2, seed=81)
    large = synth.dosage(700, 9000, 4, seed=82)
    cont = synth.continuous(90, 700, seed=83)
    ref_small = O.np_grmsimple_raw_dosage(small, 2)
    ref_large = O.np_grmsimple_raw_dosage(large, 4)
    for ctx in (ctx1, ctx2, ctx1):
        assert np.array_equal(grm_call(small, ctx=ctx, skip_repair=True)[0], ref_small)
        assert np.array_equal(grm_call(large, ctx=ctx, skip_repair=True)[0], ref_large)
        assert relerr(grm_call(cont, 2, ctx=ctx, skip_repair=True)[0], O.c_grmploidyaware_raw(cont, 2)[1]) <= RTOL_F64
        assert np.array_equal(grm_call(small, ctx=ctx, skip_repair=True)[0], ref_small)
    lib = ctx1.lib
    G = np.zeros((50, 50), order="F")
    o = _lib.default_options()
    info = _lib.new_info()
    import ctypes as C

    assert lib.grm_cuda_simple(ctx1.handle, small.ctypes.data, 50, 300, 49, G.ctypes.data, 50, C.byref(o), C.byref(info)) \
        == _lib.ERR_BAD_ARGUMENT  # lda < n
    assert lib.grm_cuda_simple(ctx1.handle, small.ctypes.data, 50, 0, 50, G.ctypes.data, 50, C.byref(o), C.byref(info)) \
        == _lib.ERR_BAD_ARGUMENT  # p < 1
    o.denominator = 3
    assert lib.grm_cuda_simple(ctx1.handle, small.ctypes.data, 50, 300, 50, G.ctypes.data, 50, C.byref(o), C.byref(info)) \
        == _lib.ERR_BAD_ARGUMENT  # denominator must be a power of two
    o = _lib.default_options()
    o.maf = 0.7
    assert lib.grm_cuda_simple(ctx1.handle, small.ctypes.data, 50, 300, 50, G.ctypes.data, 50, C.byref(o), C.byref(info)) \
        == _lib.ERR_BAD_ARGUMENT
    # a wrong denominator hint is detected by the per-cell verification
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call(large, ctx=ctx1, denominator=2, skip_repair=True)
    assert e.value.status == _lib.ERR_NOT_DOSAGE
    # m^2 p must fit the int32 accumulators: detected up front (p is only a count here, no such matrix is allocated)
    ctx1.close()
    ctx2.close()


# P9: ordering follows the input order of entries
def test_row_order_follows_entries():
    A = synth.dosage(60, 500, 4, seed=13)
    perm = np.random.default_rng(0).permutation(60)
    G, _ = grm_call(A, 4, skip_repair=True)
    Gp, _ = grm_call(np.asfortranarray(A[perm]), 4, skip_repair=True)
    assert np.array_equal(Gp, G[np.ix_(perm, perm)])


def test_leading_dimensions_and_chunked_streaming():
    big = np.zeros((150, 900), order="F")
    A = synth.dosage(130, 900, 2, seed=17)
    big[:130] = A
    view = big[:130]  # lda = 150
    G0, _ = grm_call(A, 2, skip_repair=True)
    G1, _ = grm_call(view, 2, skip_repair=True)
    assert np.array_equal(G0, G1)
    G2, i2 = grm_call(A, 2, skip_repair=True, chunk_loci=128)  # 8 chunks, double-buffered H2D
    assert np.array_equal(G0, G2)
    out = np.zeros((140, 130), order="F")[:130]  # ldg = 140
    G3, _ = grm_call(A, 2, skip_repair=True, out=out)
    assert np.array_equal(G0, G3)
    Ac = synth.continuous(130, 900, seed=18)
    F0, _ = grm_call(Ac, 2, skip_repair=True)
    F1, _ = grm_call(Ac, 2, skip_repair=True, chunk_loci=128)
    assert relerr(F1, F0) <= 1e-13


def test_denominator_detection_fallbacks():
    # first chunk looks diploid, a later chunk holds quarter dosages: retried at m = 64, still exact
    A = np.asfortranarray(np.hstack([synth.dosage(70, 256, 2, seed=1), synth.dosage(70, 256, 4, seed=2)]))
    G, info = grm_call(A, skip_repair=True, chunk_loci=128)
    assert info["path"] == _lib.PATH_I8
    assert np.array_equal(G, O.np_grmsimple_raw_dosage(A, 4))
    # first chunk dosage, later chunk continuous: falls back to the FP64 path
    B = np.asfortranarray(np.hstack([synth.dosage(70, 256, 2, seed=1), synth.continuous(70, 256, seed=2)]))
    G, info = grm_call(B, skip_repair=True, chunk_loci=128)
    assert info["path"] == _lib.PATH_F64
    assert relerr(G, O.c_grmsimple_raw(B)[1]) <= RTOL_F64


# P8: tile sharding, emulated ranks on one GPU
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_tiles_union_is_bit_identical(world):
    A = synth.dosage(1100, 700, 4, seed=21)  # 5 x 5 tiles of 256
    full, _ = grm_call(A, 4, skip_repair=True)
    acc = np.zeros_like(full)
    for r in range(world):
        part, info = grm_call(A, 4, skip_repair=True, rank=r, world=world)
        assert info["tiles"] == _lib.load().grm_cuda_tile_count(1100, r, world)
        assert not (np.triu(acc != 0) & np.triu(part != 0)).any(), "tiles of different ranks overlap"
        acc += part
    iu = np.triu_indices(1100)
    assert np.array_equal(acc[iu], full[iu])  # bit-identical: integer Gram + integer centring statistics
    Ac = synth.continuous(600, 500, seed=22)
    fullc, info = grm_call(Ac, None, skip_repair=True)
    accc = np.zeros_like(fullc)
    for r in range(world):
        part, _ = grm_call(Ac, None, skip_repair=True, rank=r, world=world)
        accc += part
    iu = np.triu_indices(600)
    assert relerr(accc[iu], fullc[iu]) <= 1e-13  # sharded FP64 calls return their tiles already divided, like the int8 path


@pytest.mark.parametrize("kernel", ["2cta", "wide"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("ploidy", [None, 4])
def test_loci_split_packed_tiles_emulated_ranks(ctx, world, ploidy, kernel):
    """The "ksplit" multi-GPU decomposition with the ranks emulated on one GPU: every rank contracts its own loci
    over all tiles (raw int32 tiles in reduce-scatter layout), the tile buffers are summed (what the NCCL
    reduce-scatter does), every rank finalises the tiles it owns; the union must be bit-identical to one call.
    `kernel`: the CTA-pair kernel or the two-tiles-per-pair kernel on the packed path (tuning key "syrk_packed")."""
    import torch

    from grm_b200 import device as D, dist as GD

    ctx.set_tuning("syrk_packed", {"2cta": 2, "wide": 3}[kernel])

    n, p, m = 1100, 3001, 4
    A = synth.dosage(n, p, m, seed=61)
    full, info = grm_call(A, ploidy, skip_repair=True)
    At = to_dev(A)
    tmax = ctx.lib.grm_cuda_packed_tiles_max(n, world)
    total = torch.zeros((world * tmax, 65536), dtype=torch.int32, device=dev())
    sums = r = T = None
    for rk in range(world):
        k0, k1 = GD.loci_range(p, rk, world)
        codes, colsum, flags = D.pack_dosage(ctx, At[k0:k1].contiguous(), m)
        total += D.syrk_i8_packed(ctx, codes, k1 - k0, world)
        sums, r, T = D.dosage_stats(ctx, codes, k1 - k0, colsum, sums, r, T, accumulate=sums is not None)
        ctx.synchronize()
    if ploidy is None:
        u, (alpha, beta, gamma, denom) = None, (1.0 / (m * m), 0.0, 0.0, float(p))
    else:
        u, (alpha, beta, gamma, denom) = D.dosage_centre(ctx, sums, r, T, p, m, ploidy)
    G = torch.zeros((n, n), dtype=torch.float64, device=dev())
    for rk in range(world):
        D.finalize_tiles(ctx, total[rk * tmax:(rk + 1) * tmax].contiguous(), n, alpha, beta, gamma, denom, u, G, rk, world)
    D.scale_mirror(ctx, G, 1.0)
    ctx.synchronize()
    ctx.set_tuning("syrk_packed", 0)
    assert np.array_equal(G.cpu().numpy(), full)


@pytest.mark.parametrize("world", [1, 3])
def test_sharded_output_tiles_match_dense_result(ctx, world):
    """BASELINE config 5's output form (tiles stay on their owners, never a dense n x n matrix) at a reduced size."""
    import ctypes as C

    from grm_b200 import device as D

    n, p, m, ploidy = 1100, 2000, 2, 2
    A = synth.dosage(n, p, m, seed=91)
    full, info = grm_call(A, ploidy, skip_repair=True)
    codes, colsum, flags = D.pack_dosage(ctx, to_dev(A), m)
    sums, r, T = D.dosage_stats(ctx, codes, p, colsum)
    u, (alpha, beta, gamma, denom) = D.dosage_centre(ctx, sums, r, T, p, m, ploidy)
    ti, tj = C.c_int64(), C.c_int64()
    seen = 0
    for rk in range(world):
        tiles = D.syrk_i8_f64_tiles(ctx, codes, p, alpha, beta, gamma, denom, u, rank=rk, world=world)
        ctx.synchronize()
        th = tiles.cpu().numpy()
        for t in range(th.shape[0]):
            assert ctx.lib.grm_cuda_tile_at(n, rk, world, t, C.byref(ti), C.byref(tj)) == 0
            r0, c0 = ti.value * 256, tj.value * 256
            blk = np.zeros((256, 256))
            sub = full[r0:r0 + 256, c0:c0 + 256]
            blk[:sub.shape[0], :sub.shape[1]] = sub
            assert np.array_equal(th[t].T, blk)  # tile is [col][row]
            seen += 1
    assert seen == 15  # 5 x 5 tiles, upper triangle


@pytest.mark.parametrize("world", [1, 2, 4])
def test_parallel_repair_equals_sequential_loop(ctx, world):
    """Multi-GPU inflatediagonals!: rounds probed in parallel (ranks emulated one after another on one GPU) must give
    the sequential loop's rounds, total and matrix, bit for bit — including the det-underflow / early-return cases."""
    import torch

    from grm_b200 import device as D, dist as GD  # noqa: F401

    cases = []
    for gen, n, p, ploidy in (("dosage", 300, 2000, None), ("continuous", 300, 2000, 2), ("dosage", 600, 3000, 2)):
        A = synth.dosage(n, p, 2, seed=7) if gen == "dosage" else synth.continuous(n, p, seed=7)
        cases.append(grm_call(A, ploidy, skip_repair=True)[0])
    cases.append(np.asfortranarray(np.eye(400) * 0.5))   # SPD but det underflows: one round
    cases.append(np.asfortranarray(np.eye(400) * 1e3))   # det overflows to Inf: untouched
    cases.append(np.asfortranarray(-np.eye(4)))
    for X in cases:
        st, Xs, it, tot = O.np_inflatediagonals(X)
        Xd0 = torch.from_numpy(np.ascontiguousarray(X)).to(dev())  # stays untouched: what every rank holds
        calls = {"base": 0}

        def exchange(mine):
            # the emulated all-gather: the probes of all ranks of this batch, evaluated one after another
            out = [list(map(float, GD.repair_probe(ctx, Xd0, calls["base"] + rk))) for rk in range(world)]
            calls["base"] += world
            return out

        Xw = Xd0.clone()
        it2, tot2 = GD.inflatediagonals_parallel(ctx, Xw, 1000, world=world, rank=0, exchange=exchange)
        ctx.synchronize()
        assert (it2, tot2) == (it, tot)
        assert np.array_equal(Xw.cpu().numpy(), Xs)


def test_repair_variants_agree():
    """Three ways through inflatediagonals! (repair.cu) must give the same rounds, eps_total, every bit of the matrix and
    the same cached Cholesky factor, on every kind of outcome:
      lit  — tuning "repair_det" = 1: the literal order of calc.jl:49,61 (LU first, Cholesky only when !(det < eps));
      seq  — the default decision rule, one round at a time ("repair" = 1): Cholesky first, the determinant from the
             factor's diagonal when partial pivoting cannot interchange rows, LU otherwise;
      spec — the default from n = 3072: the same rule with two rounds in flight on two streams."""
    n = 3200
    rng = np.random.default_rng(5)
    cases = {
        "grm needing rounds": grm_call(synth.dosage(n, 6000, 2, seed=3), 2, skip_repair=True)[0],
        "grmsimple": grm_call(synth.dosage(n, 6000, 4, seed=4), None, skip_repair=True)[0],
        "det underflows, one round": np.asfortranarray(np.eye(n) * 0.5),
        "det overflows, untouched": np.asfortranarray(np.eye(n) * 1e3),
        "several rounds": np.asfortranarray(np.eye(n) * 0.3),  # 0.3^n, 0.6^n, 0.9^n underflow; 1.2^n overflows: 3 rounds
        "indefinite": np.asfortranarray(np.eye(n) - 2.0 * np.diag((np.arange(n) % 7 == 0).astype(float))),
    }
    W = rng.standard_normal((n, 40))
    cases["low rank"] = np.asfortranarray(W @ W.T / 40.0)
    # positive definite, but partial pivoting WOULD interchange rows 0 and 1: the pivots of the Cholesky are not the LU's
    swap = np.eye(n)
    swap[0, 1] = swap[1, 0] = 2.0
    swap[1, 1] = 5.0
    cases["needs pivoting"] = np.asfortranarray(swap)
    # det = 1e-15: within reach of eps, the LU has the last word
    near = np.eye(n)
    near[np.arange(15), np.arange(15)] = 0.1
    cases["det near eps"] = np.asfortranarray(near)
    lit, seq, spec = grm_b200.Context(0), grm_b200.Context(0), grm_b200.Context(0)
    try:
        lit.set_tuning("repair_det", 1)
        seq.set_tuning("repair", 1)
        for name, X in cases.items():
            for max_iter in (1000, 1, 0):
                out = {}
                for tag, c in (("lit", lit), ("seq", seq), ("spec", spec)):
                    Y = np.array(X, order="F", copy=True)
                    it, tot = C_inflate(c, Y, max_iter)
                    try:
                        U = grm_b200.last_factor(n, ctx=c)
                    except ArgumentError:
                        U = None
                    out[tag] = (it, tot, Y, U)
                i0, t0, Y0, U0 = out["lit"]
                for tag in ("seq", "spec"):
                    i1, t1, Y1, U1 = out[tag]
                    assert (i0, t0) == (i1, t1), (name, tag, max_iter, i0, i1, t0, t1)
                    assert np.array_equal(Y0, Y1), (name, tag, max_iter)
                    assert (U0 is None) == (U1 is None), (name, tag, max_iter)
                    if U0 is not None:
                        assert np.array_equal(U0, U1), (name, tag, max_iter)
                if U0 is not None:  # U'U is the returned matrix, U upper triangular
                    assert np.array_equal(np.tril(U0, -1), np.zeros_like(U0))
                    assert relerr(U0.T @ U0, Y0) < 1e-13
        # the error paths keep their order (calc.jl:43-57)
        bad = np.array(cases["grmsimple"], order="F", copy=True)
        bad[3, 5] = bad[5, 3] + 1.0
        for c in (lit, seq, spec):
            with pytest.raises(_lib.GrmCudaError) as e:
                C_inflate(c, bad.copy(order="F"), 10)
            assert e.value.status == _lib.ERR_NOT_SYMMETRIC
        # an Inf that the early return does not swallow: det underflows (0.5 I), so calc.jl:52-57 get their turn
        inf = np.asfortranarray(np.eye(n) * 0.5)
        inf[3, 5] = inf[5, 3] = np.inf
        for c in (lit, seq, spec):
            with pytest.raises(_lib.GrmCudaError) as e:
                C_inflate(c, inf.copy(order="F"), 10)
            assert e.value.status == _lib.ERR_INF_MATRIX
    finally:
        for c in (lit, seq, spec):
            c.close()


def test_repair_reports_how_the_determinant_was_formed():
    """info.repair_lu: a GRM whose determinant under- or overflows is decided by the Cholesky alone (no LU); the literal
    order runs one LU per evaluated round; a determinant within reach of eps sends the default rule to the LU too.  Same
    outcome each time."""
    A = synth.dosage(1200, 9000, 4, seed=9)  # det(raw) = 1e-52, det(inflated) = Inf: far from eps both times
    c_def, c_lit = grm_b200.Context(0), grm_b200.Context(0)
    try:
        c_lit.set_tuning("repair_det", 1)
        G0, i0 = grm_call(A, 4, ctx=c_def)
        G1, i1 = grm_call(A, 4, ctx=c_lit)
        assert np.array_equal(G0, G1) and (i0["repair_iters"], i0["eps_total"]) == (i1["repair_iters"], i1["eps_total"])
        assert i0["repair_iters"] == 1 and i0["repair_lu"] == 0 and i1["repair_lu"] == 2
        # 700 x 9000 with this seed: det(raw) = 4.5e-16, twice eps — no inflation, and only the LU may say so
        B = synth.dosage(700, 9000, 4, seed=9)
        G0, i0 = grm_call(B, 4, ctx=c_def)
        G1, i1 = grm_call(B, 4, ctx=c_lit)
        assert np.array_equal(G0, G1) and i0["repair_iters"] == i1["repair_iters"] == 0
        assert i0["repair_lu"] == 1 and i1["repair_lu"] == 1
    finally:
        c_def.close()
        c_lit.close()


def test_repair_beyond_the_32_bit_range_of_the_legacy_cusolver_entry_points():
    """n > 46,000: n * n no longer fits the 32-bit indexing of cusolverDnDpotrf / Dgetrf, the 64-bit entry points (Xpotrf /
    Xgetrf) take over (repair.cu: LEGACY_MAX_N).  Device-resident matrix, default rule against the literal LU-first order;
    the cached factor is checked through U'U = X on a few rows."""
    import ctypes as C

    import torch

    n = 46_080
    free, _total = torch.cuda.mem_get_info(0)
    if free < 8 * n * n * 7:
        pytest.skip("needs about 120 GB of free device memory")
    g = torch.Generator(device=dev()).manual_seed(3)
    X = torch.rand((n, n), dtype=torch.float64, device=dev(), generator=g)
    X = (X + X.T) * (0.25 / n)          # small symmetric off-diagonal part: row sums below 0.5
    X.diagonal().fill_(0.6)             # diagonally dominant => positive definite; det ~ 0.6^n underflows => one round,
    torch.cuda.synchronize()            # after which det ~ 1.2^n overflows: both far from eps, no LU needed by default
    torch.cuda.empty_cache()            # the library allocates with cudaMalloc, next to torch's cache
    out = {}
    for tag, knob in (("default", 0), ("lu_first", 1)):
        c = grm_b200.Context(0)
        try:
            c.set_tuning("repair_det", knob)
            Y = X.clone()
            torch.cuda.synchronize()
            it, tot = C.c_int32(0), C.c_double(0.0)
            _lib.check(c.lib.grm_cuda_inflate_diagonals(c.handle, C.c_void_p(Y.data_ptr()), n, n, 1000, _lib.LOC_DEVICE,
                                                        C.byref(it), C.byref(tot)))
            c.synchronize()
            out[tag] = (it.value, tot.value, Y)
            if tag == "default":
                U = torch.empty((n, n), dtype=torch.float64, device=dev())  # column-major n x n == U.T in torch's view
                _lib.check(c.lib.grm_cuda_last_factor(c.handle, n, C.c_void_p(U.data_ptr()), n, _lib.LOC_DEVICE))
                c.synchronize()
                Ut = U  # Ut[j, i] = U_colmajor[i, j]
                for i in (0, 1, 12_345, n - 1):
                    row = Ut @ Ut[i]  # (U'U)[:, i] = sum_k U[k, :] U[k, i]; in torch's view: Ut[:, k] Ut[i, k]
                    assert torch.allclose(row, Y[i], rtol=0.0, atol=1e-12)
                assert float(Ut[5, 7]) == 0.0 and float(Ut[7, 5]) != 0.0  # U[7,5] (below the diagonal) is zero, U[5,7] is not
                del U, Ut
        finally:
            c.close()
    (i0, t0, Y0), (i1, t1, Y1) = out["default"], out["lu_first"]
    assert (i0, t0) == (i1, t1) == (1, 0.6)
    assert torch.equal(Y0, Y1)
    assert torch.equal(Y0.diagonal(), torch.full((n,), 0.6 + 0.6, dtype=torch.float64, device=dev()))


def C_inflate(ctx, Y, max_iter):
    """grm_cuda_inflate_diagonals on a host matrix through a given context -> (iters, eps_total); Y updated in place."""
    import ctypes as C

    it, tot = C.c_int32(0), C.c_double(0.0)
    n = Y.shape[0]
    _lib.check(ctx.lib.grm_cuda_inflate_diagonals(ctx.handle, Y.ctypes.data, n, n, int(max_iter), _lib.LOC_HOST, C.byref(it),
                                                  C.byref(tot)))
    return it.value, tot.value


# ---------------------------------------------------------------------------------------------------------------
# full-size (BASELINE C2) properties: a checksum of checksums for the integer Gram matrix
#   sum_j C_ij = sum_k c_ik S_k ,  C_ii = sum_k c_ik^2 ,  sum_ij C_ij = sum_k S_k^2      (S_k = column sums)
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_c2_integer_properties(ctx):
    import torch

    from grm_b200 import device as D

    n, p, m = 5000, 100_000, 2
    At = synth.dosage_device(n, p, m, 7, dev())
    torch.cuda.synchronize()
    codes, colsum, flags = D.pack_dosage(ctx, At, m)
    C = D.syrk_i8(ctx, codes, p)
    G, info = D.grm_device(ctx, At, None, skip_repair=True)
    ctx.synchronize()
    assert int(flags[0]) == 0 and info["path"] == _lib.PATH_I8 and info["denominator"] == 2
    assert torch.equal(codes[:, :p].T.double() / m, At)  # round trip of the quantiser
    Cfull = torch.triu(C.T.long()) + torch.triu(C.T.long(), 1).T  # mirror the upper triangle
    S = colsum[:p].long()
    assert torch.equal(S, codes[:, :p].long().sum(0))
    cl = codes[:, :p].double()
    rows = (cl @ S.double()).long()  # exact: all partial sums < 2^53
    assert torch.equal(Cfull.sum(1), rows)
    assert torch.equal(torch.diagonal(Cfull), (cl * cl).sum(1).long())
    assert int(Cfull.sum()) == int((S * S).sum())
    # independent library cross-check of every element (cuBLASLt int8 GEMM; comparator only)
    ref = torch._int_mm(codes[:, :p].contiguous(), codes[:, :p].T.contiguous())
    assert torch.equal(torch.triu(C.T), torch.triu(ref))
    # final matrix: exactly fl(fl(C/m^2)/p), exactly symmetric.  (Tensor divisors: torch turns division by a Python
    # scalar into a multiplication by its reciprocal on CUDA, which is not correctly rounded.)
    four = torch.tensor(float(m * m), dtype=torch.float64, device=dev())
    pp = torch.tensor(float(p), dtype=torch.float64, device=dev())
    assert torch.equal(G, torch.div(torch.div(Cfull.double(), four), pp))
    assert torch.equal(G, G.T)
    sub = Cfull[:64, :64].cpu().numpy().astype(np.float64)
    assert np.array_equal(G[:64, :64].cpu().numpy(), (sub / (m * m)) / p)


# ---------------------------------------------------------------------------------------------------------------
# consumer side: the Cholesky factor MvNormal(mu, Sigma) computes, and the by-name lookup
# ---------------------------------------------------------------------------------------------------------------
RTOL_CHOL = 1e-10  # FP64 library factorisation (cuSOLVER potrf vs LAPACK potrf): same algorithm class, other blocking


@pytest.mark.parametrize("n,p,kind", [(100, 10_000, "dosage"), (257, 2000, "continuous"), (1, 10, "dosage")])
def test_cholesky_factor_matches_lapack_and_is_reused_from_the_repair(n, p, kind):
    from scipy.linalg import cholesky as sp_chol

    A = synth.dosage(n, p, 2, seed=21) if kind == "dosage" else synth.continuous(n, p, seed=22)
    ctx = grm_b200.default_context()
    grm = grmsimple(Genomes.from_matrix(A), ctx=ctx)
    G = grm.genomic_relationship_matrix
    U_cached = grm_b200.last_factor(n, ctx=ctx)        # left behind by the repair's final isposdef
    U = grm_b200.cholesky(G, ctx=ctx)                  # computed afresh
    assert np.array_equal(U, U_cached)                  # same routine on the same matrix
    ref = sp_chol(G, lower=False)                       # LAPACK dpotrf('U'), what Julia calls
    assert np.array_equal(np.tril(U, -1), np.zeros_like(U))
    assert relerr(U, ref) < RTOL_CHOL
    assert relerr(U.T @ U, G) < 1e-13
    assert grm_b200.isposdef(G, ctx=ctx)
    # only the upper triangle is read, as with Julia's cholesky(Symmetric(X)) inside PDMat
    G2 = np.array(G, order="F", copy=True)
    G2[np.tril_indices(n, -1)] = 123.0
    assert np.array_equal(grm_b200.cholesky(G2, ctx=ctx), U)
    # not positive definite -> PosDefException; and the cached factor is gone after a failed factorisation
    bad = -np.eye(max(n, 2))
    with pytest.raises(grm_b200.PosDefException):
        grm_b200.cholesky(bad, ctx=ctx)
    assert not grm_b200.isposdef(bad, ctx=ctx)
    with pytest.raises(ArgumentError):
        grm_b200.last_factor(max(n, 2), ctx=ctx)
    # repair skipped -> nothing cached
    grmsimple(Genomes.from_matrix(A), ctx=ctx, skip_repair=True)
    with pytest.raises(ArgumentError):
        grm_b200.last_factor(n, ctx=ctx)


def test_simulatecovariancekinship_and_lookup_by_entry_name():
    n, p = 60, 3000
    g = Genomes.from_matrix(synth.continuous(n, p, seed=5))
    S = grm_b200.simulatecovariancekinship(n, g)
    st, ref, _, _ = O.c_grmsimple(g.allele_frequencies)
    assert st == O.OK and relerr(S, ref) < RTOL_F64
    assert S.shape == (n, n) and np.var(S) > 0.0 and abs(np.linalg.det(S)) > 0.0  # simulate_effects.jl:334-342
    with pytest.raises(ArgumentError):
        grm_b200.simulatecovariancekinship(n + 1, g)
    # analyseviaBLR's lookup (bayes.jl:1314-1349): coefficient names "entries entry_XX" in another order, a subset
    grm = grmsimple(g)
    pick = [7, 3, 59, 0, 21]
    names = [f"entries {g.entries[i]}" for i in pick]
    Sig = grm_b200.covariance_from_grm(grm, names)
    sub = np.array(grm.genomic_relationship_matrix[np.ix_(pick, pick)], order="F")
    ref_sub = sub.copy()
    k = 0
    while not O.np_isposdef(ref_sub) and k <= 10:
        _, ref_sub, _, _ = O.np_inflatediagonals(ref_sub)
        k += 1
    assert np.allclose(Sig, ref_sub, rtol=1e-12, atol=0)
    assert grm_b200.isposdef(Sig)
    # interaction names that differ in more than one token stay zero
    names2 = [f"site_1 {g.entries[0]}", f"site_2 {g.entries[1]}", f"site_1 {g.entries[1]}"]
    Sig2 = grm_b200.covariance_from_grm(grm, names2)
    assert Sig2[0, 1] == 0.0 and Sig2[0, 2] != 0.0 and Sig2[1, 2] != 0.0


# ---------------------------------------------------------------------------------------------------------------
# distances(genomes): every metric, both dimensions, against the pair-by-pair restatement
# ---------------------------------------------------------------------------------------------------------------
RTOL_DIST = 1e-12  # FP64 sums of <= a few thousand terms in a different order than Julia's SIMD sum


def _assert_dist_equal(got, ref, key):
    assert got.shape == ref.shape, key
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), f"{key}: fill pattern differs"
    assert np.array_equal(got[~fin], ref[~fin]), f"{key}: fill values differ"  # +Inf / -Inf placed identically
    if "counts" in key:
        assert np.array_equal(got, ref), key
    elif fin.any():
        scale = max(np.abs(ref[fin]).max(), 1e-300)
        assert np.abs(got[fin] - ref[fin]).max() <= RTOL_DIST * scale, (key, np.abs(got[fin] - ref[fin]).max(), scale)


@pytest.mark.parametrize("n,p,t,missing", [(40, 300, 25, 0.0), (70, 500, 100, 0.15), (33, 64, 64, 0.5), (5, 200, 3, 0.0)])
def test_distances_match_the_pairwise_restatement(n, p, t, missing):
    rng = np.random.default_rng(100 + n)
    A = synth.continuous(n, p, seed=n)
    if missing > 0:
        A[rng.random(A.shape) < missing] = np.nan
        A[0, :] = np.nan                      # an entry with no data at all: every pair with it keeps the fill
        A[1, 2:] = np.nan                     # an entry with at most 2 usable loci
        A[:, 5] = 0.25                        # a constant locus: var < 1e-7 -> correlation stays -Inf
        A[3, 7] = np.inf                      # non-finite cells are excluded like missing ones
    idx1 = np.sort(rng.choice(p, t, replace=False)) + 1          # 1-based, like the Julia argument
    if missing > 0:
        idx1 = np.unique(np.concatenate([idx1, [6, 8]]))          # make sure the special loci are in
    g = Genomes.from_matrix(A)
    metrics = ["euclidean", "correlation", "mad", "rmsd", "χ²"]
    names, entries, got = grm_b200.distances(g, metrics, idx1.tolist())
    ref = O.np_distances(A, idx1 - 1, metrics)
    assert sorted(got) == sorted(ref)
    assert names == [g.loci_alleles[k - 1] for k in idx1] and entries == g.entries
    for key in ref:
        _assert_dist_equal(got[key], ref[key], key)
    # the reference's doctest properties (genomes.jl:440-452)
    if missing == 0.0:
        C = got["entries|correlation"]
        assert np.array_equal(np.diag(C), np.ones(n))
        X = got["loci_alleles|χ²"]
        assert np.array_equal(np.diag(X), np.zeros(len(idx1)))
    # entries|χ² is not symmetric, loci_alleles|χ² is (mirrored upper triangle); loci counts are upper-triangular
    assert np.array_equal(got["loci_alleles|χ²"], got["loci_alleles|χ²"].T)
    assert np.array_equal(np.tril(got["loci_alleles|counts"], -1), np.zeros_like(got["loci_alleles|counts"]))


@pytest.mark.parametrize("n,p,t,dims", [(3000, 20, 20, ("loci_alleles",)), (40, 1200, 1000, ("loci_alleles", "entries")),
                                         (700, 33, 33, ("loci_alleles",))])
def test_distances_split_form_for_few_pairs_and_long_vectors(n, p, t, dims):
    """Few pairs over long vectors (the reference's default: 100 loci x 100 loci over all entries) run as position
    slices + fixed-order reductions; same results as the pair-by-pair restatement."""
    rng = np.random.default_rng(n + t)
    A = synth.continuous(n, p, seed=n + 1)
    A[rng.random(A.shape) < 0.05] = np.nan
    A[:, 3] = 0.75
    idx1 = np.sort(rng.choice(p, t, replace=False)) + 1
    g = Genomes.from_matrix(A)
    metrics = ["euclidean", "correlation", "mad", "rmsd", "χ²"]
    _, _, got = grm_b200.distances(g, metrics, idx1.tolist(), include_entries="entries" in dims)
    ref = O.np_distances(A, idx1 - 1, metrics, dims=dims)
    assert sorted(got) == sorted(ref)
    for key in ref:
        _assert_dist_equal(got[key], ref[key], key)
    # deterministic: a second call gives the same bits
    _, _, again = grm_b200.distances(g, metrics, idx1.tolist(), include_entries="entries" in dims)
    for key in got:
        assert np.array_equal(got[key], again[key], equal_nan=True), key


def test_distances_argument_handling_and_device_buffers(ctx):
    import ctypes as C

    import torch

    g = Genomes.from_matrix(synth.continuous(20, 150, seed=2))
    with pytest.raises(ArgumentError):
        grm_b200.distances(g, ["euclidean", "covariance"], [1, 2, 3])   # not a recognised metric (genomes.jl:470)
    with pytest.raises(ArgumentError):
        grm_b200.distances(g, [], [1, 2, 3])
    with pytest.raises(ArgumentError):
        grm_b200.distances(g, ["mad"], [0, 1])                           # 1-based range check (genomes.jl:497-499)
    with pytest.raises(grm_b200.ErrorException):
        grm_b200.distances(g, ["mad"], [1, 2], include_loci_alleles=False, include_entries=False)
    # duplicates and order are normalised like unique(sort(idx)); subsets of the outputs
    _, _, d = grm_b200.distances(g, ["mad"], [9, 3, 3, 7], include_entries=False, include_counts=False)
    assert sorted(d) == ["loci_alleles|mad"] and d["loci_alleles|mad"].shape == (3, 3)
    _, _, d100 = grm_b200.distances(g, ["rmsd"])                         # default: 100 sampled loci
    assert d100["loci_alleles|rmsd"].shape == (100, 100) and d100["entries|counts"].max() == 100
    # device-resident input and output through the C ABI, leading dimension larger than the matrix
    A = g.allele_frequencies
    n, p = A.shape
    At = to_dev(A)
    idx0 = np.array([2, 5, 11, 40, 41, 149], dtype=np.int64)
    out = torch.full((n, n + 3), -7.0, dtype=torch.float64, device=dev())  # column-major n x n with ld = n + 3
    torch.cuda.synchronize()
    _lib.check(ctx.lib.grm_cuda_distances(ctx.handle, C.c_void_p(At.data_ptr()), n, p, n, _lib.LOC_DEVICE,
                                          idx0.ctypes.data, len(idx0), _lib.DIM_ENTRIES, C.c_void_p(out.data_ptr()),
                                          None, None, None, None, None, n + 3, _lib.LOC_DEVICE))
    got = out.cpu().numpy()  # (n, ld) row-major == column-major ld x n: got[j, i] is element (i, j)
    ref = O.np_distances(A, idx0, ["euclidean"])["entries|euclidean"]
    assert np.abs(got[:, :n].T - ref).max() <= RTOL_DIST * ref.max()
    assert np.all(got[:, n:] == -7.0)                                    # padding rows untouched


def test_estimateld_per_chromosome_correlation_matrices():
    """estimateld (src/genomes/impute.jl:172-279): one loci x loci correlation matrix per chromosome / scaffold label."""
    n, per = 60, [40, 1, 75]
    A = synth.continuous(n, sum(per), seed=77)
    A[np.random.default_rng(5).random(A.shape) < 0.3] = np.nan      # the doctest uses sparsity = 0.3
    names = [f"chrom_{c + 1}\t{k + 1}\tA|T\tA" for c, m in enumerate(per) for k in range(m)]
    g = Genomes.from_matrix(A, loci_alleles=names)
    chroms, LDs = grm_b200.estimateld(g)
    assert chroms == ["chrom_1", "chrom_2", "chrom_3"] and len(LDs) == 3
    assert LDs[1] is None                                            # fewer than 2 loci: skipped (impute.jl:252-254)
    off = 0
    for c, m in enumerate(per):
        if m >= 2:
            ref = O.np_distances(A, np.arange(off, off + m), ["correlation"], dims=("loci_alleles",))
            _assert_dist_equal(LDs[c], ref["loci_alleles|correlation"], f"chrom_{c + 1}")
            dg = np.diag(LDs[c])
            assert np.all((dg == 1.0) | np.isneginf(dg))               # -Inf: a locus with var < 1e-7 (genomes.jl:560)
        off += m
    # caller-defined scaffolds (divideintomockscaffolds-style labels)
    labels = [f"mock_{k // 29}" for k in range(sum(per))]
    chroms2, LDs2 = grm_b200.estimateld(g, chromosomes=labels)
    assert chroms2 == sorted(set(labels)) and [x.shape[0] for x in LDs2] == [labels.count(c) for c in chroms2]


@pytest.mark.parametrize("mode", ["wide", "single", "window"])
@pytest.mark.parametrize("n,p,m,ploidy", [(300, 1000, 2, None), (700, 3001, 4, 4), (1100, 777, 2, None), (2049, 5001, 4, 2)])
def test_syrk_kernel_variants_are_bit_identical_to_the_pair_kernel(mode, n, p, m, ploidy):
    """Tuning key "syrk_i8" = 3: a CTA pair works on two tiles that share their column panel (the default for long
    contractions with many tiles, i.e. C3; forced here on small ragged shapes incl. leftover single-tile units);
    = 1: the single-CTA kernel; "syrk_sync" = 1: the pair kernel with the K-window synchronisation (cooperative launch).
    All must reproduce the plain pair kernel bit for bit.  A fresh context per run: the knobs are per context."""
    import torch

    from grm_b200 import device as D

    At = synth.dosage_device(n, p, m, 31, dev())
    torch.cuda.synchronize()
    c0, c1 = grm_b200.Context(0), grm_b200.Context(0)
    try:
        c0.set_tuning("syrk_i8", 2)
        c0.set_tuning("syrk_sync", 0)
        G0, i0 = D.grm_device(c0, At, ploidy, denominator=m, skip_repair=True)
        if mode == "window":
            c1.set_tuning("syrk_i8", 2)
            c1.set_tuning("syrk_sync", 1)
        else:
            c1.set_tuning("syrk_i8", {"wide": 3, "single": 1}[mode])
        assert c1.get_tuning("syrk_i8") in (1, 2, 3)
        G1, i1 = D.grm_device(c1, At, ploidy, denominator=m, skip_repair=True)
        c0.synchronize()
        c1.synchronize()
        assert i0["syrk_window"] == 0
        if mode == "window":
            assert i1["syrk_window"] == 1, "the cooperative launch of the K-window kernel was refused on an idle device"
        assert torch.equal(G0, G1)
    finally:
        c0.close()
        c1.close()
    with pytest.raises(_lib.GrmCudaError):
        grm_b200.default_context(0).set_tuning("no_such_knob", 1)
