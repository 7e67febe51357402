"""GPU: the host ingest against the device path it replaces.

genomes.allele_frequencies is a Matrix{Union{Float64,Missing}} in pageable memory (src/all_structs.jl:66).  The
grm_cuda_*_union entry points read it in place: worker threads quantise to int8 into a pinned ring (dosage input) or
copy FP64 (continuous input).  Everything here is compared with the plain double* entry points on the NaN-dense copy
of the same matrix — bit-identical on the int8 path (the quantisation rule is the same on both sides), <= 1e-13 on the
FP64 path (same kernels, other chunk boundaries) — and, for the statuses, with what the reference does
(src/grm/calc.jl:127,205 throw on `missing`; a NaN value surfaces in inflatediagonals!, calc.jl:43).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import grm_b200
from grm_b200 import _lib, grm_call, grm_call_codes, grm_call_union, grm_from_wire, io, synth
from oracle import oracle as O


def dev():
    import torch

    return torch.device("cuda", 0)


def to_dev(A):
    import torch

    return torch.from_numpy(np.ascontiguousarray(np.asarray(A).T)).to(dev())


def relerr(G, ref):
    return np.abs(G - ref).max() / np.abs(ref).max()


def host_codes(A, m, sel=None, tag=1):
    lib = _lib.load()
    n, p = A.shape
    lds = (n + 3) // 4 * 4
    codes = np.zeros((p, lds), dtype=np.int8)
    flags = C.c_uint32(0)
    assert lib.grm_cuda_host_quantise(A.ctypes.data, sel.ctypes.data if sel is not None else None, tag, n, p, n, m,
                                      codes.ctypes.data, lds, 0, C.byref(flags)) == 0
    return codes, flags.value


@pytest.mark.parametrize("n,p,m", [(1001, 3001, 4), (130, 257, 2), (4, 128, 1), (2050, 1000, 64), (257, 129, 8)])
def test_host_packed_codes_equal_the_device_pack_kernel(ctx, n, p, m):
    """host quantise + pack_codes_kernel (byte transpose)  ==  pack_dosage_tma_kernel on the FP64 cells: codes, column
    sums and flags, on clean data and with every kind of non-code cell planted."""
    import torch

    from grm_b200 import device as D

    A = synth.dosage(n, p, m, seed=n + p)
    rng = np.random.default_rng(5)
    for plant in (False, True):
        if plant:
            for v in (np.nan, np.inf, 0.3, -0.25, 1.0 + 1.0 / m):
                A[rng.integers(n), rng.integers(p)] = v
        hc, hflags = host_codes(A, m)
        c1, s1, f1 = D.pack_codes(ctx, torch.from_numpy(hc).to(dev()),
                                  codes=torch.zeros((n, D.padded_loci(p)), dtype=torch.int8, device=dev()))
        c2, s2, f2 = D.pack_dosage(ctx, to_dev(A), m)
        ctx.synchronize()
        assert torch.equal(c1, c2), "host-packed codes differ from the device pack kernel"
        assert torch.equal(s1, s2)
        assert int(f1[0]) == int(f2[0]) == hflags
        assert (int(f1[0]) != 0) == plant
    # QC mode: missing cells are counted per locus, not flagged, identically on both routes
    B = synth.add_missing(synth.dosage(n, p, m, seed=3), 0.01, seed=4)
    hc, hflags = host_codes(B, m)
    z = lambda: torch.zeros((D.padded_loci(p),), dtype=torch.int32, device=dev())
    m1, m2 = z(), z()
    c1, s1, f1 = D.pack_codes(ctx, torch.from_numpy(hc).to(dev()),
                              codes=torch.zeros((n, D.padded_loci(p)), dtype=torch.int8, device=dev()), miss=m1)
    c2, s2, f2 = D.pack_dosage(ctx, to_dev(B), m, miss=m2)
    ctx.synchronize()
    assert torch.equal(c1, c2) and torch.equal(s1, s2) and torch.equal(m1, m2)
    assert int(f1[0]) == int(f2[0]) == 0 and hflags == 1
    assert int(m1.sum()) == int(np.isnan(B).sum())


@pytest.mark.parametrize("kind,ploidy", [("dosage", None), ("dosage", 4), ("continuous", None), ("continuous", 2)])
def test_union_entry_points_equal_the_dense_call(kind, ploidy):
    n, p = 301, 2500
    A = synth.dosage(n, p, 4, seed=11) if kind == "dosage" else synth.continuous(n, p, seed=12)
    buf, payload, sel = io.union_layout(A)
    ref, iref = grm_call(A, ploidy)
    G, info = grm_call_union(payload, sel, io.FLOAT_TAG, ploidy)
    assert info["ingest"] == (_lib.INGEST_HOSTQ if kind == "dosage" else _lib.INGEST_STREAM)
    assert info["host_threads"] >= 1 and info["path"] == iref["path"]
    assert info["repair_iters"] == iref["repair_iters"] and info["eps_total"] == iref["eps_total"]
    assert np.array_equal(G, ref)
    # no selectors = a plain Matrix{Float64}
    G, info = grm_call_union(payload, None, 0, ploidy)
    assert np.array_equal(G, ref)
    # many small chunks through the ring (more chunks than ring slots): same integers, same FP64 sums up to reassociation
    G, info = grm_call_union(payload, sel, io.FLOAT_TAG, ploidy, chunk_loci=128)
    if kind == "dosage":
        assert np.array_equal(G, ref)
    else:
        assert relerr(G, ref) <= 1e-13
    # forcing the FP64 stream for dosage input (tuning key "ingest"): the device pack sees the same cells
    c = grm_b200.Context(0)
    try:
        c.set_tuning("ingest", _lib.INGEST_STREAM)
        c.set_tuning("host_threads", 3)
        G, info = grm_call_union(payload, sel, io.FLOAT_TAG, ploidy, ctx=c)
        assert info["ingest"] == _lib.INGEST_STREAM and info["host_threads"] == 3
        assert np.array_equal(G, ref)
    finally:
        c.close()


def test_union_missing_and_special_cells_give_the_reference_statuses():
    n, p = 120, 900
    A = synth.dosage(n, p, 2, seed=21)
    Am = np.array(A, order="F")
    Am[5, 40] = Am[77, 800] = np.nan
    # the payload under a `missing` selector is garbage (here: not even a dosage) and must never be looked at
    buf, payload, sel = io.union_layout(Am, garbage=0.3)
    with pytest.raises(_lib.GrmCudaError) as e:  # calc.jl:127,205: missing => throw
        grm_call_union(payload, sel, io.FLOAT_TAG, 2)
    assert e.value.status == _lib.ERR_MISSING
    # the QC mask drops the two loci: equals the dense call with the same mask, and the oracle on the sliced matrix
    G, info = grm_call_union(payload, sel, io.FLOAT_TAG, 2, skip_repair=True, max_locus_sparsity=0.0)
    Gd, idd = grm_call(Am, 2, skip_repair=True, max_locus_sparsity=0.0)
    keep = O.np_qc_keep(Am, maf=0.0, max_locus_sparsity=0.0)
    assert info["path"] == _lib.PATH_I8 and info["loci_kept"] == idd["loci_kept"] == int(keep.sum()) == p - 2
    assert np.array_equal(G, Gd)
    assert relerr(G, O.c_grmploidyaware_raw(np.asfortranarray(Am[:, keep]), 2, nthreads=4)[1]) <= 1e-12
    # mean-fill policy: FP64 path, equal to the dense call
    G, info = grm_call_union(payload, sel, io.FLOAT_TAG, 2, skip_repair=True, missing="meanfill")
    Gd, _ = grm_call(Am, 2, skip_repair=True, missing="meanfill")
    assert info["path"] == _lib.PATH_F64 and relerr(G, Gd) <= 1e-13
    # a Float64 cell that holds NaN is a VALUE: the reference computes a GRM with NaN in it and inflatediagonals!
    # reports "not symmetric" (calc.jl:43)
    for kind in ("dosage", "continuous"):
        B = synth.dosage(n, p, 2, seed=22) if kind == "dosage" else synth.continuous(n, p, seed=22)
        buf, payload, sel = io.union_layout(B)
        payload[9, 500] = np.nan
        with pytest.raises(_lib.GrmCudaError) as e:
            grm_call_union(payload, sel, io.FLOAT_TAG, None)
        assert e.value.status == _lib.ERR_NOT_SYMMETRIC and "not symmetric" in e.value.message
        with pytest.raises(_lib.GrmCudaError) as e:
            grm_call_union(payload, sel, io.FLOAT_TAG, None, skip_repair=True)
        assert e.value.status == _lib.ERR_NAN_MATRIX
        payload[9, 500] = np.inf
        with pytest.raises(_lib.GrmCudaError) as e:
            grm_call_union(payload, sel, io.FLOAT_TAG, None)
        assert e.value.status == _lib.ERR_NONFINITE
    # one non-dosage cell far beyond the detection sample: AUTO retries / falls back to FP64, a forced int8 path refuses
    n2, p2 = 2100, 2600  # > 2^22 cells
    B = synth.dosage(n2, p2, 2, seed=23)
    B[17, p2 - 1] = 0.3
    buf, payload, sel = io.union_layout(B)
    G, info = grm_call_union(payload, sel, io.FLOAT_TAG, None, skip_repair=True)
    Gd, idd = grm_call(B, None, skip_repair=True)
    assert info["path"] == idd["path"] == _lib.PATH_F64 and relerr(G, Gd) <= 1e-13
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call_union(payload, sel, io.FLOAT_TAG, None, skip_repair=True, path=_lib.PATH_I8)
    assert e.value.status == _lib.ERR_NOT_DOSAGE
    # a finer grid that only shows up late: retried with m = 64, still the int8 path, bit-identical
    B = synth.dosage(n2, p2, 2, seed=24)
    B[3, p2 - 5] = 3.0 / 64.0
    buf, payload, sel = io.union_layout(B)
    G, info = grm_call_union(payload, sel, io.FLOAT_TAG, 2, skip_repair=True)
    Gd, idd = grm_call(B, 2, skip_repair=True)
    assert info["path"] == _lib.PATH_I8 and info["denominator"] == idd["denominator"] == 64
    assert np.array_equal(G, Gd)


def test_prepacked_codes_and_the_wire_file(tmp_path):
    n, p, m = 257, 1900, 4
    A = synth.dosage(n, p, m, seed=31)
    codes, mm = io.quantise(A)
    assert mm == m and codes.dtype == np.int8 and codes.flags.f_contiguous
    for ploidy in (None, 4):
        ref, iref = grm_call(A, ploidy)
        G, info = grm_call_codes(codes, m, ploidy)
        assert info["path"] == _lib.PATH_I8 and info["denominator"] == m and info["ingest"] == _lib.INGEST_HOSTQ
        assert np.array_equal(G, ref) and info["repair_iters"] == iref["repair_iters"]
    f = str(tmp_path / "g.grw")
    assert io.write_genomes(f, A, ploidy=4) == io.DTYPE_I8
    G, info = grm_from_wire(f, 4)
    assert np.array_equal(G, grm_call(A, 4)[0])
    B = synth.continuous(90, 700, seed=32)
    assert io.write_genomes(f, B) == io.DTYPE_F64
    G, info = grm_from_wire(f, None)
    assert info["path"] == _lib.PATH_F64 and np.array_equal(G, grm_call(B, None)[0])
    # missing marker and values that are no codes
    bad = np.array(codes, order="F")
    bad[4, 7] = io.CODE_MISSING
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call_codes(bad, m, None)
    assert e.value.status == _lib.ERR_MISSING
    G, info = grm_call_codes(bad, m, None, skip_repair=True, max_locus_sparsity=0.0)  # the mask drops that locus
    keep = np.ones(p, bool)
    keep[7] = False
    assert info["loci_kept"] == p - 1 and np.array_equal(G, grm_call(np.asfortranarray(A[:, keep]), None, skip_repair=True)[0])
    bad = np.array(codes, order="F")
    bad[4, 7] = m + 1
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call_codes(bad, m, None)
    assert e.value.status == _lib.ERR_NOT_DOSAGE
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call_codes(codes, 3, None)
    assert e.value.status == _lib.ERR_BAD_ARGUMENT


def test_auto_path_falls_back_to_fp64_beyond_the_int32_range():
    """m^2 p >= 2^31: the reference computes such inputs, so the automatic path must too (FP64 SYRK) instead of
    returning an overflow error; only an explicitly forced int8 path refuses.  (api.cu:542)"""
    n, p = 40, 530_000  # 64^2 * 530,000 = 2.17e9 >= 2^31
    rng = np.random.default_rng(41)
    A = np.asfortranarray(rng.integers(0, 65, (n, p)) / 64.0)
    G, info = grm_call(A, None, skip_repair=True)
    assert info["path"] == _lib.PATH_F64
    ref = (A @ A.T) / p
    assert relerr(G, ref) <= 1e-10
    G2, info2 = grm_call(A, 2, skip_repair=True)
    assert info2["path"] == _lib.PATH_F64
    assert relerr(G2, O.c_grmploidyaware_raw(A, 2, nthreads=4)[1]) <= 1e-10
    for kw in (dict(path=_lib.PATH_I8), dict(denominator=64)):
        with pytest.raises(_lib.GrmCudaError) as e:
            grm_call(A, None, skip_repair=True, **kw)
        assert e.value.status == _lib.ERR_OVERFLOW


def test_pinned_input_is_copied_directly_and_equals_the_pageable_ring():
    """FP64 stream from a pinned caller buffer: asynchronous copies straight out of it (no host threads); from pageable
    memory: worker threads fill the pinned ring.  Same cells reach the same kernels."""
    import torch

    n, p = 300, 3000
    for gen in ("continuous", "dosage"):
        A = synth.continuous(n, p, seed=51) if gen == "continuous" else synth.dosage(n, p, 2, seed=52)
        pinned = torch.empty((p, n), dtype=torch.float64).pin_memory()
        pinned.copy_(torch.from_numpy(np.ascontiguousarray(A.T)))
        Ap = pinned.numpy().T  # (n, p) Fortran-ordered view of the pinned buffer
        assert Ap.flags.f_contiguous
        c = grm_b200.Context(0)
        try:
            c.set_tuning("ingest", _lib.INGEST_STREAM)
            G1, i1 = grm_call(Ap, 2, ctx=c, chunk_loci=256)
            G2, i2 = grm_call(A, 2, ctx=c, chunk_loci=256)
            assert i1["ingest"] == i2["ingest"] == _lib.INGEST_STREAM
            assert i1["ms_host_busy"] == 0.0 and i2["ms_host_busy"] > 0.0
            assert np.array_equal(G1, G2)
        finally:
            c.close()


@pytest.mark.parametrize("n", [999, 1000, 1001, 1002])
def test_odd_entry_counts_through_the_ring(n):
    """The ring pads staged rows (4 bytes for codes, 16 for FP64) so the aligned fast kernels still apply: every residue
    of n must give the dense call's bits."""
    p = 700
    for A, pl in ((synth.dosage(n, p, 2, seed=n), 2), (synth.continuous(n, p, seed=n), None)):
        buf, payload, sel = io.union_layout(A)
        G, info = grm_call_union(payload, sel, io.FLOAT_TAG, pl, skip_repair=True, chunk_loci=256)
        import torch

        from grm_b200 import device as D

        Gd, _ = D.grm_device(grm_b200.default_context(0), to_dev(A), pl, skip_repair=True)
        torch.cuda.synchronize()
        ref = Gd.cpu().numpy()
        if info["path"] == _lib.PATH_I8:
            assert np.array_equal(G, ref)
        else:
            assert relerr(G, ref) <= 1e-13
