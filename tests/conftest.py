import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")
    # the tests load the in-tree libgrm_cuda.so (ABI checks on CPU, everything on GPU); build it if a fresh checkout is
    # tested before __graft_entry__.build() ran (nvcc cross-compiles sm_100a without a GPU)
    lib = os.path.join(ROOT, "genomicbreedingcore.jl_b200", "libgrm_cuda.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess

        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "genomicbreedingcore.jl_b200", "csrc"), "-j8", "-s"], check=False)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "grm_golden.npz"))


@pytest.fixture(scope="session")
def ctx():
    import grm_b200

    return grm_b200.default_context(0)
