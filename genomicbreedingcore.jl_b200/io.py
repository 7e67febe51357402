"""Binary formats either side of the GRM path (row f2 of SURVEY.md §8).

1. RAW matrices (`write_matrix` / `read_matrix`), shared with julia/selftest.jl so that a machine WITH Julia can dump
   `genomes.allele_frequencies` and the reference's own GRMs for replay here as reference-generated golden vectors:
   int64 rows, int64 cols (little-endian), then rows*cols Float64 in COLUMN-MAJOR order — exactly what Julia's
   `write(io, Int64(size(M,1)), Int64(size(M,2))); write(io, M)` produces.  NaN stands for `missing`.

2. The WIRE format of allele frequencies (`write_genomes` / `read_genomes`): a 64-byte header followed by one payload,
   column-major like Julia's matrix (entry fastest):

       offset  0  char[8]  magic   "GRMWIRE1"
               8  uint32   version (1)
              12  uint32   dtype   0 = Float64 cells (NaN = missing), 1 = int8 dosage codes
              16  int64    n       entries
              24  int64    p       loci
              32  int32    m       dosage denominator of the codes (power of two <= 64); 0 for Float64
              36  int32    ploidy  informational (0 = unknown)
              40  ...      zero padding to 64 bytes
              64  payload  n*p Float64, or n*p int8 with code = m*x in 0..m and -128 = missing

   The int8 payload is what `grm_cuda_from_codes` consumes: 1/8 of the bytes on disk and over PCIe, and neither the
   Union{Float64,Missing} -> Float64 conversion nor the quantisation pass is paid again at GRM time.  It is written by the
   library's own host quantiser (`grm_cuda_host_quantise`), the code the threaded ingest runs.

3. `union_layout`: an in-memory emulation of how Julia stores `Matrix{Union{Float64,Missing}}`
   (src/all_structs.jl:66) — one allocation of n*p Float64 payloads followed by n*p selector bytes — for exercising the
   `grm_cuda_*_union` entry points without Julia.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass
from typing import Optional

import numpy as np

MAGIC = b"GRMWIRE1"
VERSION = 1
DTYPE_F64, DTYPE_I8 = 0, 1
HEADER_BYTES = 64
CODE_MISSING = -128
_HEADER = struct.Struct("<8sIIqqii")


def write_matrix(path: str, M: np.ndarray) -> None:
    M = np.asfortranarray(M, dtype="<f8")
    with open(path, "wb") as f:
        np.array(M.shape, dtype="<i8").tofile(f)
        f.write(M.tobytes(order="F"))


def read_matrix(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        rows, cols = np.fromfile(f, dtype="<i8", count=2)
        data = np.fromfile(f, dtype="<f8", count=int(rows) * int(cols))
    if data.size != rows * cols:
        raise ValueError(f"{path}: expected {rows}x{cols} Float64 values, found {data.size}")
    return data.reshape((int(rows), int(cols)), order="F")


# ---------------------------------------------------------------------------------------------------------------
# wire format
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class WireGenomes:
    dtype: int            # DTYPE_F64 / DTYPE_I8
    n: int
    p: int
    m: int                # denominator of the codes, 0 for Float64
    ploidy: int
    data: np.ndarray      # (n, p) Fortran-ordered float64, or int8 codes

    def to_float(self) -> np.ndarray:
        """Float64 allele frequencies (missing -> NaN): the matrix the codes stand for, exactly (x = code / m)."""
        if self.dtype == DTYPE_F64:
            return self.data
        out = np.asfortranarray(self.data.astype(np.float64) / float(self.m))
        out[self.data == CODE_MISSING] = np.nan
        return out


def quantise(A: np.ndarray, m: int = 0, threads: int = 0):
    """(codes int8 (n, p) Fortran-ordered, m) through the library's host quantiser; m = 0 detects the denominator.
    Raises ValueError if the matrix is not dosage-quantised (or holds Inf).  NaN cells become CODE_MISSING."""
    from . import _lib

    lib = _lib.load()
    A = np.asfortranarray(A, dtype=np.float64)
    n, p = A.shape
    if m == 0:
        mm = C.c_int32(0)
        _lib.check(lib.grm_cuda_host_detect(A.ctypes.data, None, 0, n, p, n, 0, 1, C.byref(mm)))
        m = mm.value
        if m == 0:
            raise ValueError("allele frequencies are not multiples of 1/m for a power of two m <= 64")
    codes = np.empty((n, p), dtype=np.int8, order="F")
    flags = C.c_uint32(0)
    _lib.check(lib.grm_cuda_host_quantise(A.ctypes.data, None, 0, n, p, n, m, codes.ctypes.data, n, threads,
                                          C.byref(flags)))
    if flags.value & ~1:
        raise ValueError(f"allele frequencies are not all multiples of 1/{m} (flags {flags.value})")
    return codes, m


def write_genomes(path: str, A: np.ndarray, *, codes: Optional[bool] = None, m: int = 0, ploidy: int = 0) -> int:
    """Write allele frequencies in the wire format; returns the dtype written.  codes=None: int8 codes when the matrix
    is dosage-quantised, Float64 otherwise; codes=True insists (ValueError if impossible)."""
    A = np.asarray(A)
    n, p = A.shape
    payload, dtype, mm = None, DTYPE_F64, 0
    if codes is None or codes:
        try:
            payload, mm = quantise(A, m)
            dtype = DTYPE_I8
        except ValueError:
            if codes:
                raise
    if payload is None:
        payload = np.asfortranarray(A, dtype="<f8")
    with open(path, "wb") as f:
        f.write(_HEADER.pack(MAGIC, VERSION, dtype, n, p, mm, int(ploidy)).ljust(HEADER_BYTES, b"\0"))
        f.write(payload.tobytes(order="F"))
    return dtype


def read_genomes(path: str) -> WireGenomes:
    with open(path, "rb") as f:
        head = f.read(HEADER_BYTES)
        if len(head) != HEADER_BYTES:
            raise ValueError(f"{path}: truncated header")
        magic, version, dtype, n, p, m, ploidy = _HEADER.unpack(head[:_HEADER.size])
        if magic != MAGIC or version != VERSION or dtype not in (DTYPE_F64, DTYPE_I8) or n < 1 or p < 1:
            raise ValueError(f"{path}: not a GRMWIRE1 file (magic {magic!r}, version {version}, dtype {dtype})")
        if dtype == DTYPE_I8 and (m < 1 or m > 64 or m & (m - 1)):
            raise ValueError(f"{path}: bad dosage denominator {m}")
        data = np.fromfile(f, dtype="<f8" if dtype == DTYPE_F64 else np.int8, count=n * p)
    if data.size != n * p:
        raise ValueError(f"{path}: expected {n}x{p} cells, found {data.size}")
    return WireGenomes(dtype, n, p, m, ploidy, data.reshape((n, p), order="F"))


# ---------------------------------------------------------------------------------------------------------------
# Julia's Matrix{Union{Float64,Missing}} in memory
# ---------------------------------------------------------------------------------------------------------------
FLOAT_TAG = 1  # Base.uniontypes(Union{Float64,Missing}) == [Missing, Float64]: selector 0 = missing, 1 = Float64


def union_layout(A: np.ndarray, float_tag: int = FLOAT_TAG, garbage: float = 0.0):
    """(buffer, payload view, selector view) of an isbits-Union array holding A (NaN -> missing): n*p Float64 payloads
    (column-major) followed by n*p selector bytes, in ordinary pageable memory.  The payload of a missing cell is
    `garbage` (Julia leaves whatever was there)."""
    A = np.asfortranarray(A, dtype=np.float64)
    n, p = A.shape
    buf = np.empty(n * p * 9, dtype=np.uint8)
    payload = buf[:n * p * 8].view(np.float64).reshape((n, p), order="F")
    sel = buf[n * p * 8:].reshape((n, p), order="F")
    miss = np.isnan(A)
    payload[...] = A
    payload[miss] = garbage
    sel[...] = float_tag
    sel[miss] = 1 - float_tag if float_tag in (0, 1) else 0
    return buf, payload, sel
