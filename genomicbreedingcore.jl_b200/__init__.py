"""grm-b200: a B200-native (sm_100a) backend for the genomic relationship matrix path of GenomicBreedingCore.jl
(src/grm/calc.jl: grmsimple, grmploidyaware, inflatediagonals!).

The directory name contains a dot, so import it through the repo-root shim:  `import grm_b200`.
"""
from . import _lib, io, synth
from ._lib import Context, GrmCudaError, default_context
from .grm import (cholesky, covariance_from_grm, distances, estimateld, grm_call, grm_call_codes, grm_call_union, grm_from_wire,
                  grmploidyaware, grmsimple, inflatediagonals, isposdef, last_factor, simulatecovariancekinship)
from .structs import GRM, ArgumentError, ErrorException, Genomes, MissingDataError, PosDefException, checkdims, clone

__all__ = [
    "Context", "GrmCudaError", "default_context", "grm_call", "grm_call_union", "grm_call_codes", "grm_from_wire", "grmsimple", "grmploidyaware", "inflatediagonals", "GRM", "Genomes",
    "ArgumentError", "ErrorException", "MissingDataError", "PosDefException", "checkdims", "clone", "synth", "io", "cholesky",
    "isposdef", "last_factor", "simulatecovariancekinship", "covariance_from_grm", "distances", "estimateld",
]
