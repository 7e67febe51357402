"""Raw binary matrix format shared with julia/selftest.jl (row f2 of SURVEY.md §8: a wire format that lets a machine
WITH Julia dump `genomes.allele_frequencies` and the reference's own GRMs, so they can be replayed here as
reference-generated golden vectors).

Layout: int64 rows, int64 cols (little-endian), then rows*cols Float64 in COLUMN-MAJOR order — exactly what Julia's
`write(io, Int64(size(M,1)), Int64(size(M,2))); write(io, M)` produces.  NaN stands for `missing`.
"""
from __future__ import annotations

import numpy as np


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
