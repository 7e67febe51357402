"""Host-side mirror of the two reference structs the GRM path touches.

`Genomes` follows src/all_structs.jl:62-72 (only the fields the hot path reads or validates), `GRM` follows
src/all_structs.jl:564-568.  `checkdims` follows src/genomes/genomes.jl:150-175 and src/grm/grm.jl:131-141;
`clone`, `hash` and `==` of a GRM follow src/grm/grm.jl:27-29, :55-61, :94-96.

Arrays are NumPy; `allele_frequencies` is n entries x p loci with NaN standing for Julia's `missing`.  Keep it
Fortran-ordered (`np.asfortranarray`) to hand it to the library without a copy — that is Julia's native layout.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


class ArgumentError(ValueError):
    """Julia's ArgumentError."""


class ErrorException(RuntimeError):
    """Julia's ErrorException."""


class PosDefException(ArithmeticError):
    """Julia's PosDefException (what cholesky / MvNormal throw for a matrix that is not positive definite)."""


class MissingDataError(TypeError):
    """What the reference raises (a MethodError inside a TaskFailedException) when `missing` reaches `dot`."""


@dataclass
class Genomes:
    entries: List[str]
    populations: List[str]
    loci_alleles: List[str]
    allele_frequencies: np.ndarray  # (n, p) float64, NaN = missing
    allele_frequencies_homologous_chroms: Optional[np.ndarray] = None
    mask: Optional[np.ndarray] = None  # (n, p) bool

    def __post_init__(self):
        if self.mask is None:
            self.mask = np.ones(self.allele_frequencies.shape, dtype=bool)

    @classmethod
    def from_matrix(cls, allele_frequencies: np.ndarray, entries=None, loci_alleles=None, populations=None):
        """Container around an existing matrix with simulategenomes-style names
        (src/simulations/simulate_genomes.jl:731: "entry_" * lpad(i, ndigits(n), "0"))."""
        n, p = allele_frequencies.shape
        w = len(str(n))
        if entries is None:
            entries = [f"entry_{i:0{w}d}" for i in range(1, n + 1)]
        if loci_alleles is None:
            loci_alleles = [f"chrom_1\t{k}\tA|T\tA" for k in range(1, p + 1)]
        if populations is None:
            populations = ["pop_1"] * n
        return cls(list(entries), list(populations), list(loci_alleles), allele_frequencies)


@dataclass
class GRM:
    entries: List[str]
    loci_alleles: List[str]
    genomic_relationship_matrix: np.ndarray  # (n, n) float64
    info: dict = field(default_factory=dict, compare=False, repr=False)  # backend diagnostics (not in the reference)

    def __hash__(self):
        h = hashlib.blake2b(digest_size=8)
        for e in self.entries:
            h.update(e.encode() + b"\0")
        h.update(b"\1")
        for e in self.loci_alleles:
            h.update(e.encode() + b"\0")
        h.update(b"\1")
        m = np.ascontiguousarray(self.genomic_relationship_matrix, dtype=np.float64)
        h.update(str(m.shape).encode())
        h.update(m.tobytes())
        return int.from_bytes(h.digest(), "little")

    def __eq__(self, other):
        return isinstance(other, GRM) and hash(self) == hash(other)


def clone(x: GRM) -> GRM:
    return GRM(list(x.entries), list(x.loci_alleles), np.array(x.genomic_relationship_matrix, copy=True))


def checkdims(x) -> bool:
    if isinstance(x, GRM):
        n = len(x.entries)
        g = x.genomic_relationship_matrix
        return g.ndim == 2 and n == g.shape[0] and n == g.shape[1]
    if isinstance(x, Genomes):
        a = x.allele_frequencies
        if a.ndim != 2:
            return False
        n, p = a.shape
        if (
            n != len(x.entries)
            or n != len(set(x.entries))
            or n != len(x.populations)
            or p != len(x.loci_alleles)
            or p != len(set(x.loci_alleles))
            or (
                x.allele_frequencies_homologous_chroms is not None
                and tuple(x.allele_frequencies_homologous_chroms.shape) != (n, p)
            )
            or tuple(x.mask.shape) != (n, p)
        ):
            return False
        return True
    raise TypeError(f"checkdims: unsupported type {type(x)!r}")
