"""Reference-generated goldens (julia/selftest.jl -> tests/golden/julia/*.bin), if any have been committed.
Without Julia in the image the directory is empty and these tests skip; the raw format itself is always tested."""
import glob
import os

import numpy as np
import pytest

import grm_b200
from grm_b200 import io

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "julia")
INPUTS = sorted(glob.glob(os.path.join(HERE, "golden_*_A.bin")))


def test_raw_format_round_trip(tmp_path):
    M = np.asfortranarray(np.random.default_rng(0).random((7, 5)))
    M[2, 3] = np.nan
    p = str(tmp_path / "m.bin")
    io.write_matrix(p, M)
    raw = open(p, "rb").read()
    assert len(raw) == 16 + 7 * 5 * 8
    assert np.frombuffer(raw[:16], dtype="<i8").tolist() == [7, 5]
    assert np.frombuffer(raw[16:24], dtype="<f8")[0] == M[0, 0] and np.frombuffer(raw[24:32], dtype="<f8")[0] == M[1, 0]
    back = io.read_matrix(p)
    assert back.flags.f_contiguous and np.array_equal(back, M, equal_nan=True)


@pytest.mark.skipif(not INPUTS, reason="no reference-generated goldens committed (Julia is not available here)")
@pytest.mark.parametrize("path", INPUTS)
def test_oracle_against_reference_goldens(path):
    from oracle import oracle as O

    A = io.read_matrix(path)
    ref_s = io.read_matrix(path.replace("_A.bin", "_grmsimple.bin"))
    ref_p = io.read_matrix(path.replace("_A.bin", "_grmploidyaware4.bin"))
    st, G, it, tot = O.c_grmsimple(A, nthreads=4)
    assert st == 0 and np.abs(G - ref_s).max() <= 1e-12 * np.abs(ref_s).max()
    st, G, it, tot = O.c_grmploidyaware(A, 4, nthreads=4)
    assert st == 0 and np.abs(G - ref_p).max() <= 1e-12 * np.abs(ref_p).max()


@pytest.mark.gpu
@pytest.mark.skipif(not INPUTS, reason="no reference-generated goldens committed (Julia is not available here)")
@pytest.mark.parametrize("path", INPUTS)
def test_cuda_against_reference_goldens(path):
    from grm_b200 import grm_call

    A = io.read_matrix(path)
    ref_s = io.read_matrix(path.replace("_A.bin", "_grmsimple.bin"))
    ref_p = io.read_matrix(path.replace("_A.bin", "_grmploidyaware4.bin"))
    G, info = grm_call(A)
    tol = 1e-12 if info["path"] == 1 else 1e-10
    assert np.abs(G - ref_s).max() <= tol * np.abs(ref_s).max()
    G, info = grm_call(A, 4)
    assert np.abs(G - ref_p).max() <= tol * np.abs(ref_p).max()
