"""CPU: the C-ABI library loads and exports every symbol include/grm_cuda.h declares; host-only entry points work;
compute entry points fail loudly (never fall back) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from grm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "grm_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grm_cuda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/grm_cuda.h but not exported"
    # and the Python binding table covers the header exactly
    assert sorted(_lib.SIGNATURES) == syms


def test_struct_layouts_match_the_header():
    assert C.sizeof(_lib.Options) == 80
    assert C.sizeof(_lib.Info) == 120
    o = _lib.default_options()
    assert (o.struct_size, o.path, o.max_iter, o.world, o.rank, o.skip_repair) == (80, 0, 1000, 1, 0, 0)
    assert (o.maf, o.max_locus_sparsity, o.missing_policy) == (0.0, 1.0, 0)


def test_version_and_status_strings():
    lib = _lib.load()
    assert lib.grm_cuda_version() == 200
    assert lib.grm_cuda_status_string(4) == b"The input matrix is not symmetric."  # calc.jl:44
    assert lib.grm_cuda_status_string(5) == b"The input matrix contains NaN values."  # calc.jl:53
    assert lib.grm_cuda_status_string(6) == b"The input matrix contains Inf values."  # calc.jl:56


@pytest.mark.parametrize("n", [1, 255, 256, 257, 5000, 20000])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tile_partition_is_a_balanced_cover(n, world):
    lib = _lib.load()
    T = (n + 255) // 256
    total = T * (T + 1) // 2
    counts = [lib.grm_cuda_tile_count(n, r, world) for r in range(world)]
    assert sum(counts) == total
    assert max(counts) - min(counts) <= 1
    if total <= 600:  # every upper-triangular tile is owned exactly once and tile_at inverts tile_owner
        seen = set()
        ti, tj = C.c_int64(), C.c_int64()
        for r in range(world):
            for t in range(counts[r]):
                assert lib.grm_cuda_tile_at(n, r, world, t, C.byref(ti), C.byref(tj)) == 0
                assert 0 <= ti.value <= tj.value < T
                assert lib.grm_cuda_tile_owner(n, ti.value, tj.value, world) == r
                seen.add((ti.value, tj.value))
        assert len(seen) == total


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GrmCudaError) as e:
        _lib.Context(0)
    assert e.value.status == _lib.ERR_CUDA
    assert "no CPU fallback" in e.value.message


def test_bad_arguments_are_rejected_before_any_device_work():
    """Null handles / pointers and impossible shapes come back as GRM_CUDA_ERR_BAD_ARGUMENT (no crash, no GPU needed)."""
    import numpy as np

    lib = _lib.load()
    A = np.zeros((4, 6), order="F")
    G = np.zeros((4, 4), order="F")
    o = _lib.default_options()
    info = _lib.new_info()
    assert lib.grm_cuda_ctx_create(0, None) == _lib.ERR_BAD_ARGUMENT
    assert lib.grm_cuda_simple(None, A.ctypes.data, 4, 6, 4, G.ctypes.data, 4, C.byref(o), C.byref(info)) == _lib.ERR_BAD_ARGUMENT
    assert lib.grm_cuda_ploidyaware(None, A.ctypes.data, 4, 6, 4, 2, G.ctypes.data, 4, C.byref(o), C.byref(info)) == _lib.ERR_BAD_ARGUMENT
    assert lib.grm_cuda_inflate_diagonals(None, G.ctypes.data, 4, 4, 10, 0, None, None) == _lib.ERR_BAD_ARGUMENT
    assert lib.grm_cuda_ctx_destroy(None) == _lib.OK
    assert lib.grm_cuda_ctx_synchronize(None) == _lib.ERR_BAD_ARGUMENT
    assert lib.grm_cuda_tile_count(0, 0, 1) == -1 and lib.grm_cuda_tile_count(10, 2, 2) == -1
    assert lib.grm_cuda_tile_owner(1000, 3, 2, 4) == -1  # ti > tj is not an upper-triangular tile
    assert lib.grm_cuda_packed_tiles_max(1000, 3) == 4  # 4 x 4 tiles -> 10 upper tiles over 3 ranks
    assert b"bad argument" in lib.grm_cuda_last_error() or lib.grm_cuda_last_error() != b""
