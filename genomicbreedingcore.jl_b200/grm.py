"""grmsimple / grmploidyaware / inflatediagonals — the reference's operator interface for the GRM path
(src/grm/calc.jl:115-142, :189-217, :41-70), backed by libgrm_cuda on a B200.

Same names, argument meaning and error behaviour as the Julia functions:
    grmsimple(genomes; max_iter=1000, verbose=false)::GRM
    grmploidyaware(genomes; ploidy=2, max_iter=1000, verbose=false)::GRM
    inflatediagonals!(X; max_iter=1000, verbose=false)::Nothing      (here: inflatediagonals(X, ...), in place)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .structs import GRM, ArgumentError, ErrorException, Genomes, MissingDataError, PosDefException, checkdims


def _raise_for(err: _lib.GrmCudaError):
    s = err.status
    if s == _lib.ERR_MISSING:
        raise MissingDataError(err.message) from None
    if s == _lib.ERR_ALL_FILTERED:
        raise ErrorException(err.message) from None  # filter.jl:363-365, :644-646 throw ErrorException
    if s == _lib.ERR_NOT_POSDEF:
        raise PosDefException(err.message) from None
    if s in (_lib.ERR_BAD_ARGUMENT, _lib.ERR_NOT_SYMMETRIC, _lib.ERR_NAN_MATRIX, _lib.ERR_INF_MATRIX,
             _lib.ERR_NONFINITE, _lib.ERR_NOT_DOSAGE, _lib.ERR_OVERFLOW):
        raise ArgumentError(err.message) from None
    raise err


def grm_call(A: np.ndarray, ploidy: Optional[int] = None, *, ctx: Optional[_lib.Context] = None,
             path: int = _lib.PATH_AUTO, denominator: int = 0, max_iter: int = 1_000, skip_repair: bool = False,
             chunk_loci: int = 0, rank: int = 0, world: int = 1, out: Optional[np.ndarray] = None,
             missing: str = "error", maf: float = 0.0, max_locus_sparsity: float = 1.0, distributed: bool = False,
             input_slab: bool = False, p_total: Optional[int] = None, tiles: bool = False):
    """The C-ABI call a Julia `ccall` shim makes, on host buffers: grm_cuda_simple (ploidy is None) or
    grm_cuda_ploidyaware.  `A` is n x p (NaN = missing); a strided view of a larger Fortran array is passed with its
    leading dimension, no copy.  Returns (G, info dict).  Raises GrmCudaError with the library status.

    distributed=True (the context carries a communicator, every rank makes the same call): `A` is the whole matrix
    (the library reads only this rank's loci range of it) or, with input_slab=True, the rank's own loci range of a
    matrix of p_total loci; every rank receives the complete, repaired GRM.  tiles=True returns this rank's packed
    [tiles_owned, 256, 256] tiles instead (tile t is [col][row])."""
    ctx = ctx or _lib.default_context()
    A = np.asarray(A)
    if A.dtype != np.float64 or A.ndim != 2 or A.strides[0] != 8 or (A.shape[1] > 1 and A.strides[1] % 8) \
            or (A.shape[1] > 1 and A.strides[1] < 8 * A.shape[0]):
        A = np.asfortranarray(A, dtype=np.float64)
    n, p = A.shape
    lda = A.strides[1] // 8 if p > 1 else n
    o = _lib.default_options()
    trank, tworld = rank, world
    if distributed:
        o.distributed, o.input_slab = 1, 1 if input_slab else 0
        if input_slab:
            p = int(p_total)
        trank, tworld, _ = ctx.comm_info()
    if tiles:
        o.output_mode = _lib.OUT_TILES
    if out is None:
        if tiles:
            out = np.empty((max(ctx.lib.grm_cuda_tile_count(n, trank, tworld), 1), 256, 256), dtype=np.float64)
        else:
            out = np.empty((n, n), dtype=np.float64, order="F")  # calc.jl:123
            if world > 1:
                out[...] = 0.0
    ldg = n if tiles else (out.strides[1] // 8 if n > 1 else n)
    o.path, o.denominator, o.max_iter = path, denominator, int(max_iter)
    o.skip_repair = 1 if skip_repair else 0
    o.chunk_loci = int(chunk_loci)
    o.rank, o.world = rank, world
    o.missing_policy = {"error": _lib.MISSING_ERROR, "meanfill": _lib.MISSING_MEANFILL}[missing]
    o.maf, o.max_locus_sparsity = float(maf), float(max_locus_sparsity)
    info = _lib.new_info()
    if ploidy is None:
        st = ctx.lib.grm_cuda_simple(ctx.handle, A.ctypes.data, n, p, lda, out.ctypes.data, ldg, C.byref(o),
                                     C.byref(info))
    else:
        st = ctx.lib.grm_cuda_ploidyaware(ctx.handle, A.ctypes.data, n, p, lda, int(ploidy), out.ctypes.data, ldg,
                                          C.byref(o), C.byref(info))
    _lib.check(st)
    return out, info.as_dict()


def _finish_call(ctx, fn, args, o, info):
    st = fn(ctx.handle, *args, C.byref(o), C.byref(info))
    _lib.check(st)
    return info.as_dict()


def _options(path, denominator, max_iter, skip_repair, chunk_loci, missing, maf, max_locus_sparsity):
    o = _lib.default_options()
    o.path, o.denominator, o.max_iter = path, denominator, int(max_iter)
    o.skip_repair = 1 if skip_repair else 0
    o.chunk_loci = int(chunk_loci)
    o.missing_policy = {"error": _lib.MISSING_ERROR, "meanfill": _lib.MISSING_MEANFILL}[missing]
    o.maf, o.max_locus_sparsity = float(maf), float(max_locus_sparsity)
    return o


def grm_call_union(payload: np.ndarray, selectors: Optional[np.ndarray], float_tag: int, ploidy: Optional[int] = None, *,
                   ctx: Optional[_lib.Context] = None, path: int = _lib.PATH_AUTO, denominator: int = 0,
                   max_iter: int = 1_000, skip_repair: bool = False, chunk_loci: int = 0, missing: str = "error",
                   maf: float = 0.0, max_locus_sparsity: float = 1.0):
    """grm_cuda_simple_union / grm_cuda_ploidyaware_union: the call julia/GRMCuda.jl makes on
    genomes.allele_frequencies as Julia stores it (Matrix{Union{Float64,Missing}}: payloads + selector bytes, pageable
    memory, nothing converted on the caller's side).  `payload` (n, p) Fortran-ordered float64, `selectors` (n, p)
    Fortran-ordered uint8 or None (plain Matrix{Float64})."""
    ctx = ctx or _lib.default_context()
    n, p = payload.shape
    if payload.dtype != np.float64 or not payload.flags.f_contiguous:
        raise TypeError("payload must be a Fortran-ordered float64 matrix")
    if selectors is not None and (selectors.dtype != np.uint8 or selectors.shape != payload.shape
                                  or not selectors.flags.f_contiguous):
        raise TypeError("selectors must be a Fortran-ordered uint8 matrix of the payload's shape")
    out = np.empty((n, n), dtype=np.float64, order="F")
    o = _options(path, denominator, max_iter, skip_repair, chunk_loci, missing, maf, max_locus_sparsity)
    info = _lib.new_info()
    sel_ptr = selectors.ctypes.data if selectors is not None else None
    if ploidy is None:
        d = _finish_call(ctx, ctx.lib.grm_cuda_simple_union, (payload.ctypes.data, sel_ptr, int(float_tag), n, p, n,
                                                              out.ctypes.data, n), o, info)
    else:
        d = _finish_call(ctx, ctx.lib.grm_cuda_ploidyaware_union, (payload.ctypes.data, sel_ptr, int(float_tag), n, p, n,
                                                                   int(ploidy), out.ctypes.data, n), o, info)
    return out, d


def grm_call_codes(codes: np.ndarray, m: int, ploidy: Optional[int] = None, *, ctx: Optional[_lib.Context] = None,
                   max_iter: int = 1_000, skip_repair: bool = False, chunk_loci: int = 0, maf: float = 0.0,
                   max_locus_sparsity: float = 1.0):
    """grm_cuda_from_codes: the GRM of pre-packed int8 dosage codes (io.read_genomes of an int8 wire file), codes[i, k] =
    m * x_ik, -128 = missing.  ploidy None -> grmsimple."""
    ctx = ctx or _lib.default_context()
    if codes.dtype != np.int8 or codes.ndim != 2 or not codes.flags.f_contiguous:
        codes = np.asfortranarray(codes, dtype=np.int8)
    n, p = codes.shape
    out = np.empty((n, n), dtype=np.float64, order="F")
    o = _options(_lib.PATH_I8, int(m), max_iter, skip_repair, chunk_loci, "error", maf, max_locus_sparsity)
    info = _lib.new_info()
    d = _finish_call(ctx, ctx.lib.grm_cuda_from_codes, (codes.ctypes.data, n, p, n, int(m), -1 if ploidy is None else int(ploidy),
                                                        out.ctypes.data, n), o, info)
    return out, d


def grm_from_wire(path: str, ploidy: Optional[int] = None, **kw):
    """GRM straight from a wire-format file (io.write_genomes): int8 payloads go through grm_cuda_from_codes, Float64
    payloads through the plain entry points."""
    from . import io

    w = io.read_genomes(path)
    if w.dtype == io.DTYPE_I8:
        return grm_call_codes(w.data, w.m, ploidy, **kw)
    return grm_call(w.data, ploidy, **kw)


def _run(genomes: Genomes, centred: bool, ploidy: int, max_iter: int, verbose: bool, ctx, path, denominator,
         skip_repair, missing="error") -> GRM:
    if not checkdims(genomes):
        raise ArgumentError("The Genomes struct is corrupted ☹.")  # calc.jl:118-120, :191-193
    try:
        G, info = grm_call(genomes.allele_frequencies, ploidy if centred else None, ctx=ctx, path=path,
                           denominator=denominator, max_iter=max_iter, skip_repair=skip_repair, missing=missing)
    except _lib.GrmCudaError as e:
        _raise_for(e)
    if verbose:  # calc.jl:67-69
        print(f"Performed {info['repair_iters']} iterations of diagonal inflation to ensure matrix invertibility "
              f"(ϵ_total = {info['eps_total']}).")
    # entries / loci_alleles alias the input vectors, like calc.jl:123,196
    grm = GRM(genomes.entries, genomes.loci_alleles, G, info=info)
    if not centred and not checkdims(grm):
        raise ErrorException("Error computing a simple GRM.")  # calc.jl:135-137
    return grm


def grmsimple(genomes: Genomes, max_iter: int = 1_000, verbose: bool = False, *, ctx: Optional[_lib.Context] = None,
              path: int = _lib.PATH_AUTO, denominator: int = 0, skip_repair: bool = False,
              missing: str = "error") -> GRM:
    """G = A*A'/p then inflatediagonals! (calc.jl:115-142).  Keyword-only extras select the backend path;
    missing="meanfill" (extension) fills missing cells with mean(skipmissing(locus)) instead of raising."""
    return _run(genomes, False, 0, max_iter, verbose, ctx, path, denominator, skip_repair, missing)


def grmploidyaware(genomes: Genomes, ploidy: int = 2, max_iter: int = 1_000, verbose: bool = False, *,
                   ctx: Optional[_lib.Context] = None, path: int = _lib.PATH_AUTO, denominator: int = 0,
                   skip_repair: bool = False, missing: str = "error") -> GRM:
    """q = colmean(A); Z = ploidy*(A-0.5) - 1q'; G = Z*Z'/(ploidy*sum(q(1-q))) then inflatediagonals!
    (calc.jl:189-217)."""
    return _run(genomes, True, ploidy, max_iter, verbose, ctx, path, denominator, skip_repair, missing)


def inflatediagonals(X: np.ndarray, max_iter: int = 1_000, verbose: bool = False, *,
                     ctx: Optional[_lib.Context] = None):
    """inflatediagonals!(X) (calc.jl:41-70): modifies X in place, returns None.  X must be a float64 square matrix;
    a C-ordered symmetric matrix is its own transpose, so either memory order is accepted."""
    if not isinstance(X, np.ndarray) or X.dtype != np.float64 or X.ndim != 2:
        raise TypeError("inflatediagonals expects a 2-D float64 numpy array (Matrix{Float64})")
    if X.shape[0] != X.shape[1]:
        # the reference checks symmetry first (calc.jl:43), so a non-square matrix reports "not symmetric"
        raise ArgumentError("The input matrix is not symmetric.")
    if not (X.flags.f_contiguous or X.flags.c_contiguous):
        raise TypeError("inflatediagonals needs a contiguous matrix to update in place")
    ctx = ctx or _lib.default_context()
    n = X.shape[0]
    iters = C.c_int32(0)
    tot = C.c_double(0.0)
    try:
        _lib.check(ctx.lib.grm_cuda_inflate_diagonals(ctx.handle, X.ctypes.data, n, n, int(max_iter), _lib.LOC_HOST,
                                                      C.byref(iters), C.byref(tot)))
    except _lib.GrmCudaError as e:
        _raise_for(e)
    if verbose:
        print(f"Performed {iters.value} iterations of diagonal inflation to ensure matrix invertibility "
              f"(ϵ_total = {tot.value}).")
    return None


# ---------------------------------------------------------------------------------------------------------------
# consumer side of the GRM
# ---------------------------------------------------------------------------------------------------------------
def cholesky(X: np.ndarray, *, ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """cholesky(X).U as MvNormal(mu, X) computes it (src/simulations/simulate_effects.jl:479-481 -> PDMat ->
    LAPACK potrf!('U', copy(X))): upper-triangular U with U'U = X; only the upper triangle of X is read.
    Raises PosDefException where Julia does."""
    if not isinstance(X, np.ndarray) or X.dtype != np.float64 or X.ndim != 2 or X.shape[0] != X.shape[1]:
        raise TypeError("cholesky expects a square 2-D float64 numpy array (Matrix{Float64})")
    ctx = ctx or _lib.default_context()
    Xf = np.asfortranarray(X)
    n = Xf.shape[0]
    U = np.empty((n, n), dtype=np.float64, order="F")
    try:
        _lib.check(ctx.lib.grm_cuda_cholesky(ctx.handle, Xf.ctypes.data, n, n, _lib.LOC_HOST, U.ctypes.data, n,
                                             _lib.LOC_HOST))
    except _lib.GrmCudaError as e:
        _raise_for(e)
    return U


def isposdef(X: np.ndarray, *, ctx: Optional[_lib.Context] = None) -> bool:
    """isposdef(X) = ishermitian(X) && the Cholesky factorisation succeeds (LinearAlgebra; used at
    src/grm/calc.jl:49,61 and src/tebv/bayes.jl:1334,1340)."""
    X = np.asarray(X)
    if X.ndim != 2 or X.shape[0] != X.shape[1] or not np.array_equal(X, X.T):
        return False
    try:
        cholesky(np.ascontiguousarray(X, dtype=np.float64), ctx=ctx)
    except PosDefException:
        return False
    return True


def last_factor(n: int, *, ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """The Cholesky factor U (U'U = G) of the n x n matrix that the last grmsimple / grmploidyaware /
    inflatediagonals call on `ctx` returned: the repair ends with this factorisation (calc.jl:49,61), so the consumer
    (MvNormal) does not have to redo it.  Raises ArgumentError if no such factor is cached."""
    ctx = ctx or _lib.default_context()
    U = np.empty((n, n), dtype=np.float64, order="F")
    try:
        _lib.check(ctx.lib.grm_cuda_last_factor(ctx.handle, n, U.ctypes.data, n, _lib.LOC_HOST))
    except _lib.GrmCudaError as e:
        _raise_for(e)
    return U


def simulatecovariancekinship(p: int, genomes: Genomes, *, ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """simulatecovariancekinship(p, genomes)::Matrix{Float64} (src/simulations/simulate_effects.jl:344-351):
    the grmsimple matrix of `genomes`, after checking that p is its number of entries."""
    if p != len(genomes.entries):
        raise ArgumentError("simulatecovariancekinship: p must match the number of entries in genomes.")
    return grmsimple(genomes, ctx=ctx).genomic_relationship_matrix


def covariance_from_grm(grm: GRM, coefficient_names, *, verbose: bool = False,
                        ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """The entries-by-name lookup analyseviaBLR applies to a GRM (src/tebv/bayes.jl:1314-1349): Sigma[i, j] =
    G[k, l] for the entries named in coefficient names i and j (the LAST space-separated token starting "entry_";
    the LAST matching row of grm.entries), left at zero when the two names differ in more than one token; then up to
    11 rounds of `inflatediagonals!` while the matrix is not positive definite, ErrorException if it still is not."""
    names = [str(x) for x in coefficient_names]
    n = len(names)
    last = {}
    for idx, e in enumerate(grm.entries):
        last[e] = idx  # findall(grm.entries .== entry)[end]
    G = grm.genomic_relationship_matrix
    toks = [nm.split(" ") for nm in names]
    ent = []
    for t in toks:
        hits = [w for w in t if "entry_" in w]
        if not hits:
            raise ArgumentError(f"coefficient name {' '.join(t)!r} names no entry")
        if hits[-1] not in last:
            raise ArgumentError(f"entry {hits[-1]!r} is not in the GRM")
        ent.append(last[hits[-1]])
    S = np.zeros((n, n), dtype=np.float64, order="F")
    for i in range(n):
        for j in range(n):
            a, b = toks[i], toks[j]
            if len(a) == len(b):
                same = sum(1 for x, y in zip(a, b) if x == y)
            else:
                raise ArgumentError("coefficient names of one factor must have the same number of tokens")
            if same < len(a) - 1:
                continue
            S[i, j] = G[ent[i], ent[j]]
    counter = 0
    while not isposdef(S, ctx=ctx):
        if counter > 10:
            break
        inflatediagonals(S, verbose=verbose, ctx=ctx)
        counter += 1
    if not isposdef(S, ctx=ctx):
        raise ErrorException("The variance-covariance matrix for the entries remains non-positive definite after 10 "
                             "inflation attempts.")
    return S


# ---------------------------------------------------------------------------------------------------------------
# distances(genomes)
# ---------------------------------------------------------------------------------------------------------------
RECOGNISED_DISTANCE_METRICS = ["euclidean", "correlation", "mad", "rmsd", "χ²"]  # genomes.jl:470


def distances(genomes: Genomes, distance_metrics=None, idx_loci_alleles=None, include_loci_alleles: bool = True,
              include_entries: bool = True, include_counts: bool = True, verbose: bool = False, *, seed: int = 42,
              ctx: Optional[_lib.Context] = None):
    """distances(genomes; distance_metrics, idx_loci_alleles, include_loci_alleles, include_entries, include_counts)
    (src/genomes/genomes.jl:454-654).  Returns (loci_alleles_names, entries, Dict "dimension|metric" -> matrix) with the
    reference's keys ("loci_alleles|euclidean", "entries|counts", ...).

    `idx_loci_alleles` is 1-based like the Julia argument.  Differences that cannot be avoided: the reference's default
    metric list contains "covariance", which its own validation rejects (:455 vs :470), so the default here is the list
    of recognised metrics; and `idx_loci_alleles = nothing` samples 100 loci with Julia's global RNG — here with
    NumPy's `default_rng(seed)`."""
    if not checkdims(genomes):
        raise ArgumentError("The genomes struct is corrupted ☹.")
    metrics = list(dict.fromkeys(RECOGNISED_DISTANCE_METRICS if distance_metrics is None else distance_metrics))
    if len(metrics) < 1:
        raise ArgumentError("Please supply at least 1 distance metric. Choose from:\n\t‣ " +
                            "\n\t‣ ".join(RECOGNISED_DISTANCE_METRICS))
    bad = [m for m in metrics if m not in RECOGNISED_DISTANCE_METRICS]
    if bad:
        raise ArgumentError("Unrecognised metric/s:\n\t‣ " + "\n\t‣ ".join(bad) + "\nPlease choose from:\n\t‣ " +
                            "\n\t‣ ".join(RECOGNISED_DISTANCE_METRICS))
    A = genomes.allele_frequencies
    n, p = A.shape
    if idx_loci_alleles is None:
        if verbose:
            print("Warning: Randomly sampling 100 loci-alleles")
        if p < 100:
            raise ArgumentError("Cannot take a sample of 100 loci-alleles without replacement from fewer than 100.")
        idx = np.sort(np.random.default_rng(seed).choice(p, 100, replace=False)) + 1
    else:
        idx = np.asarray(idx_loci_alleles, dtype=np.int64)
        if idx.size == 0 or idx.min() < 1 or idx.max() > p:
            raise ArgumentError("We accept `idx_loci_alleles` from 1 to `n_loci_alleles` of `genomes`.")
        idx = np.unique(idx)  # unique(sort(...)), genomes.jl:500
    t = int(idx.size)
    idx0 = np.ascontiguousarray(idx - 1, dtype=np.int64)
    ctx = ctx or _lib.default_context()
    if A.dtype != np.float64 or A.strides[0] != 8 or (p > 1 and (A.strides[1] % 8 or A.strides[1] < 8 * n)):
        A = np.asfortranarray(A, dtype=np.float64)
    lda = A.strides[1] // 8 if p > 1 else n
    slot = {"euclidean": 0, "correlation": 1, "mad": 2, "rmsd": 3, "χ²": 4}
    out = {}
    for dim_name, dim, r, on in (("loci_alleles", _lib.DIM_LOCI, t, include_loci_alleles and t > 1),
                                 ("entries", _lib.DIM_ENTRIES, n, include_entries and n > 1)):
        if not on:
            continue
        mats = [None] * 6
        for m in metrics:
            mats[slot[m]] = np.empty((r, r), dtype=np.float64, order="F")
        if include_counts:
            mats[5] = np.empty((r, r), dtype=np.float64, order="F")
        ptrs = [x.ctypes.data if x is not None else None for x in mats]
        try:
            _lib.check(ctx.lib.grm_cuda_distances(ctx.handle, A.ctypes.data, n, p, lda, _lib.LOC_HOST, idx0.ctypes.data, t,
                                                  dim, *ptrs, r, _lib.LOC_HOST))
        except _lib.GrmCudaError as e:
            _raise_for(e)
        for m in metrics:
            out[f"{dim_name}|{m}"] = mats[slot[m]]
        if include_counts:
            out[f"{dim_name}|counts"] = mats[5]
    if not out:
        raise ErrorException("Genomes struct is too sparse. No distance matrix was calculated.")
    return [genomes.loci_alleles[k] for k in idx0], list(genomes.entries), out


def estimateld(genomes: Genomes, chromosomes=None, verbose: bool = False, *, ctx: Optional[_lib.Context] = None):
    """estimateld(genomes; chromosomes, verbose) (src/genomes/impute.jl:172-279): per chromosome (or per caller-defined
    scaffold label of each locus-allele) the loci x loci correlation matrix `distances(...)["loci_alleles|correlation"]`
    over all its loci.  Returns (sorted unique chromosomes, list of matrices; None where a chromosome has fewer than 2
    loci — the reference leaves that slot undefined).  The reference's JLD2 checkpoint / resume files are storage, not
    arithmetic, and are not reproduced."""
    if not checkdims(genomes):
        raise ArgumentError("The Genomes struct is corrupted ☹.")
    if chromosomes is None:
        chromosomes = [x.split("\t")[0] for x in genomes.loci_alleles]
    chromosomes = [str(c) for c in chromosomes]
    if len(chromosomes) != len(genomes.loci_alleles):
        raise ArgumentError("`chromosomes` must label every locus-allele of `genomes`.")
    chroms_uniq = sorted(set(chromosomes))
    lab = np.asarray(chromosomes)
    LDs = []
    for i, chrom in enumerate(chroms_uniq):
        if verbose:
            print(f"Esimtating LD of chromosome: {chrom} ({i + 1} of {len(chroms_uniq)})")
        idx = np.flatnonzero(lab == chrom) + 1
        if idx.size < 2:
            LDs.append(None)
            continue
        _, _, corr = distances(genomes, ["correlation"], idx.tolist(), include_loci_alleles=True, include_counts=False,
                               include_entries=False, verbose=verbose, ctx=ctx)
        LDs.append(corr["loci_alleles|correlation"])
    return chroms_uniq, LDs
