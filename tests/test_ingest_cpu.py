"""CPU: the host half of the ingest (csrc/ingest_host.cpp, reached through the C ABI's GPU-free utilities) and the wire
format (io.py).  The quantiser is what the worker threads of grm_cuda_*_union run on Julia's
Matrix{Union{Float64,Missing}} (src/all_structs.jl:66); here it is checked cell by cell against a numpy restatement of
the device pack kernel's test: c = rn(m x) is a code iff float(c) == m x and 0 <= c <= m."""
import ctypes as C

import numpy as np
import pytest

from grm_b200 import _lib, io, synth

MISSING, NAN_PAYLOAD, INF, BAD = -128, -127, -126, -125
F_MISSING, F_INF, F_NOT_DOSAGE, F_NAN_PAYLOAD = 1, 2, 4, 16


def host_quantise(A, sel, tag, m, threads=0, ldo=None):
    lib = _lib.load()
    n, p = A.shape
    ldo = ldo or n
    codes = np.full((p, ldo), 99, dtype=np.int8)  # row k = locus k (column-major n x p with leading dimension ldo)
    flags = C.c_uint32(0)
    st = lib.grm_cuda_host_quantise(A.ctypes.data, sel.ctypes.data if sel is not None else None, tag, n, p, n, m,
                                    codes.ctypes.data, ldo, threads, C.byref(flags))
    assert st == _lib.OK
    return codes, flags.value


def np_quantise(A, sel, tag, m):
    """numpy restatement: (codes (n, p) int8, flags)."""
    n, p = A.shape
    out = np.zeros((n, p), dtype=np.int8)
    flags = 0
    is_float = np.ones((n, p), bool) if sel is None else (sel == tag)
    with np.errstate(invalid="ignore", over="ignore"):
        t = A * float(m)
        c = np.rint(t)
        ok = is_float & (c == t) & (c >= 0) & (c <= m)
    out[ok] = c[ok].astype(np.int8)
    miss = ~is_float
    nanp = is_float & np.isnan(A)
    inf = is_float & np.isinf(A)
    bad = is_float & ~ok & ~nanp & ~inf
    if sel is None:  # plain double*: NaN is the missing marker
        miss, nanp = nanp, np.zeros_like(nanp)
    out[miss], out[nanp], out[inf], out[bad] = MISSING, NAN_PAYLOAD, INF, BAD
    flags |= F_MISSING * bool(miss.any()) | F_NAN_PAYLOAD * bool(nanp.any()) | F_INF * bool(inf.any()) | F_NOT_DOSAGE * bool(bad.any())
    return out, flags


@pytest.mark.parametrize("n,p,m", [(1, 1, 1), (7, 13, 2), (16, 5, 4), (17, 129, 4), (1001, 257, 64), (33, 40, 8)])
@pytest.mark.parametrize("with_sel", [False, True])
def test_host_quantiser_matches_the_pack_rule_cell_by_cell(n, p, m, with_sel):
    rng = np.random.default_rng(n * 1000 + p + m)
    A = np.asfortranarray(rng.integers(0, m + 1, (n, p)) / float(m))
    sel = None
    if with_sel:
        sel = np.ones((n, p), dtype=np.uint8, order="F")
    for _ in range(min(n * p // 3, 25)):  # plant every kind of cell that is not a code, at random places
        i, k, kind = rng.integers(n), rng.integers(p), rng.integers(6)
        A[i, k] = [np.nan, np.inf, -np.inf, 0.3, 1.0 + 1.0 / m, -1.0 / m][kind]
    if with_sel:
        for _ in range(min(n * p // 4, 10)):
            i, k = rng.integers(n), rng.integers(p)
            sel[i, k] = 0
            A[i, k] = rng.choice([0.123, np.nan, 0.5])  # the payload of a missing cell is garbage and must not matter
    want, wflags = np_quantise(A, sel, 1, m)
    for threads in (1, 3):
        got, flags = host_quantise(A, sel, 1, m, threads)
        assert np.array_equal(got.T, want)
        assert flags == wflags
    # leading dimension of the output (the pinned ring pads rows to 4 bytes): padding bytes are left untouched
    ldo = (n + 3) // 4 * 4 + 4
    got, flags = host_quantise(A, sel, 1, m, 2, ldo=ldo)
    assert np.array_equal(got[:, :n].T, want) and np.all(got[:, n:] == 99)


@pytest.mark.parametrize("hint", [0, 1, 2, 3])
@pytest.mark.parametrize("stream", [0, 1])
def test_host_quantiser_prefetch_hint_and_streaming_stores_change_nothing(hint, stream):
    """grm_cuda_host_tune("prefetch_hint" / "stream_stores"): memory-system knobs only — same codes, same flags, also
    when a column of the output is not 16-byte aligned (streaming stores are then not used for it)."""
    lib = _lib.load()
    rng = np.random.default_rng(17)
    try:
        assert lib.grm_cuda_host_tune(b"prefetch_hint", hint) == _lib.OK
        assert lib.grm_cuda_host_tune(b"stream_stores", stream) == _lib.OK
        for n in (64, 1000, 1001):  # ldo = n: aligned columns, columns aligned every other time, never aligned
            A = np.asfortranarray(rng.integers(0, 5, (n, 37)) / 4.0)
            sel = np.ones((n, 37), dtype=np.uint8, order="F")
            A[5, 3], A[n - 1, 36], A[17, 20] = np.nan, 0.3, np.inf
            sel[9, 9] = 0
            want, wflags = np_quantise(A, sel, 1, 4)
            got, flags = host_quantise(A, sel, 1, 4, 3)
            assert np.array_equal(got.T, want) and flags == wflags
    finally:
        lib.grm_cuda_host_tune(b"prefetch_hint", -1)  # back to the defaults
        lib.grm_cuda_host_tune(b"stream_stores", -1)
    assert lib.grm_cuda_host_tune(b"no_such_knob", 1) != _lib.OK


@pytest.mark.parametrize("n", [1, 7, 8, 64, 1000, 1001])
@pytest.mark.parametrize("with_sel", [False, True])
@pytest.mark.parametrize("stream", [0, 1])
def test_host_copy_of_continuous_data(n, with_sel, stream):
    """grm_cuda_host_copy — what the worker threads run for continuous input (FP64 ring): every cell as Float64, `missing`
    -> NaN whatever its payload, flags as the reference-facing statuses need them; streaming stores change nothing, also
    with rows of the output that are not 64-byte aligned."""
    lib = _lib.load()
    rng = np.random.default_rng(n + 7)
    p = 23
    A = np.asfortranarray(rng.random((n, p)))
    sel = np.ones((n, p), dtype=np.uint8, order="F") if with_sel else None
    want_flags = 0
    if n >= 7:
        A[3, 5] = np.nan   # a Float64 cell holding NaN
        want_flags |= F_NAN_PAYLOAD if with_sel else F_MISSING
        if with_sel:
            sel[n - 1, 0] = sel[0, p - 1] = sel[4, 11] = 0
            A[4, 11] = np.nan  # garbage under a missing selector must not count as a NaN payload
            want_flags |= F_MISSING
    want = A.copy()
    if with_sel:
        want[sel != 1] = np.nan
    try:
        assert lib.grm_cuda_host_tune(b"stream_stores", stream) == _lib.OK
        for ldo in (n, (n + 7) // 8 * 8 + 8):
            for threads in (1, 3):
                out = np.full((p, ldo), -7.0)
                flags = C.c_uint32(0)
                st = lib.grm_cuda_host_copy(A.ctypes.data, sel.ctypes.data if with_sel else None, 1, n, p, n, out.ctypes.data, ldo,
                                            threads, C.byref(flags))
                assert st == _lib.OK and flags.value == want_flags
                got = out[:, :n].T
                assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(want)], want[~np.isnan(want)])
                assert np.all(out[:, n:] == -7.0)  # padding untouched
    finally:
        lib.grm_cuda_host_tune(b"stream_stores", -1)


def test_host_quantiser_other_float_tag_and_simd_level():
    lib = _lib.load()
    assert lib.grm_cuda_host_simd() in (b"avx512", b"avx2", b"scalar")
    A = np.asfortranarray(np.full((40, 3), 0.5))
    sel = np.zeros((40, 3), dtype=np.uint8, order="F")  # a union whose Float64 component has selector 0
    sel[7, 1] = 1
    got, flags = host_quantise(A, sel, 0, 2)
    assert flags == F_MISSING and got[1, 7] == MISSING and (got.T[sel == 0] == 1).all()


def test_host_detect_denominator():
    lib = _lib.load()
    m = C.c_int32(-1)
    for ploidy in (1, 2, 4, 8, 64):
        A = synth.dosage(50, 400, ploidy, seed=ploidy)
        assert lib.grm_cuda_host_detect(A.ctypes.data, None, 0, 50, 400, 50, 0, 0, C.byref(m)) == _lib.OK
        assert m.value == ploidy
    A = synth.dosage(50, 400, 2, seed=1)
    A[:] = np.where(A == 0.5, 1.0, A)  # only 0 and 1 left: haploid-looking diploids
    lib.grm_cuda_host_detect(A.ctypes.data, None, 0, 50, 400, 50, 0, 0, C.byref(m))
    assert m.value == 1
    B = synth.continuous(50, 400, seed=2)
    lib.grm_cuda_host_detect(B.ctypes.data, None, 0, 50, 400, 50, 0, 0, C.byref(m))
    assert m.value == 0
    A = synth.dosage(50, 400, 4, seed=3)
    A[3, 0] = np.nan
    lib.grm_cuda_host_detect(A.ctypes.data, None, 0, 50, 400, 50, 0, 0, C.byref(m))
    assert m.value == 0  # a missing cell in the sample: the caller's full pass reports it
    lib.grm_cuda_host_detect(A.ctypes.data, None, 0, 50, 400, 50, 0, 1, C.byref(m))
    assert m.value == 4  # ... unless the QC mask is going to deal with it
    lib.grm_cuda_host_detect(A.ctypes.data, None, 0, 50, 400, 50, 3, 0, C.byref(m))  # sample ends before the NaN
    assert m.value in (1, 2, 4)
    assert lib.grm_cuda_host_detect(None, None, 0, 50, 400, 50, 0, 0, C.byref(m)) == _lib.ERR_BAD_ARGUMENT


def test_union_layout_emulation():
    A = synth.dosage(9, 11, 2, seed=4)
    A[2, 3] = A[8, 10] = np.nan
    buf, payload, sel = io.union_layout(A, garbage=0.3)
    assert buf.size == 9 * 11 * 9 and payload.flags.f_contiguous and sel.flags.f_contiguous
    assert sel.ctypes.data == payload.ctypes.data + 9 * 11 * 8  # selectors directly after the payloads
    assert sel[2, 3] == 0 and sel[8, 10] == 0 and sel.sum() == 9 * 11 - 2
    assert payload[2, 3] == 0.3 and not np.isnan(payload).any()


def test_wire_format_round_trip_and_header(tmp_path):
    import struct

    A = synth.dosage(37, 211, 4, seed=3)
    A[3, 5] = np.nan
    f = str(tmp_path / "a.grw")
    assert io.write_genomes(f, A, ploidy=4) == io.DTYPE_I8
    raw = open(f, "rb").read()
    assert len(raw) == 64 + 37 * 211
    magic, ver, dt, n, p, m, pl = struct.unpack("<8sIIqqii", raw[:40])
    assert (magic, ver, dt, n, p, m, pl) == (b"GRMWIRE1", 1, 1, 37, 211, 4, 4) and raw[40:64] == b"\0" * 24
    assert np.frombuffer(raw[64:64 + 37], dtype=np.int8).tolist() == (A[:, 0] * 4).astype(int).tolist()  # column-major
    w = io.read_genomes(f)
    assert (w.dtype, w.n, w.p, w.m, w.ploidy) == (io.DTYPE_I8, 37, 211, 4, 4)
    assert w.data[3, 5] == io.CODE_MISSING and np.array_equal(w.to_float(), A, equal_nan=True)
    # continuous data stays Float64; insisting on codes is an error
    B = synth.continuous(10, 20, seed=1)
    assert io.write_genomes(f, B) == io.DTYPE_F64
    w = io.read_genomes(f)
    assert w.dtype == io.DTYPE_F64 and w.m == 0 and np.array_equal(w.data, B) and len(open(f, "rb").read()) == 64 + 1600
    with pytest.raises(ValueError):
        io.write_genomes(f, B, codes=True)
    # a denominator that is too coarse for the data
    with pytest.raises(ValueError):
        io.write_genomes(f, synth.dosage(10, 20, 4, seed=2), codes=True, m=2)
    # corrupt files are rejected
    open(f, "wb").write(b"NOTAWIRE" + raw[8:])
    with pytest.raises(ValueError):
        io.read_genomes(f)
    open(f, "wb").write(raw[:100])
    with pytest.raises(ValueError):
        io.read_genomes(f)
