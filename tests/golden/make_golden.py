"""Regenerates tests/golden/grm_golden.npz.

The reference holds no numeric golden vectors for the GRM path and Julia is not installed, so these fixtures are
outputs of the CPU oracle (oracle/grm_oracle.c) on small seeded inputs, written once and committed; the NumPy/SciPy
restatement must agree with them (tests/test_oracle.py) and the CUDA path is compared against them through the C ABI
(tests/test_gpu_parity.py).  Inputs are stored too, so the fixtures do not depend on NumPy's RNG streams.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import grm_b200  # noqa: E402
from grm_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = {}


def case(name, A, ploidy):
    st, Gs = O.c_grmsimple_raw(A)
    assert st == 0
    st, Gp, q, d = O.c_grmploidyaware_raw(A, ploidy)
    assert st == 0
    _, Fs, its, tots = O.c_inflatediagonals(Gs)
    _, Fp, itp, totp = O.c_inflatediagonals(Gp)
    out[f"{name}_A"] = np.ascontiguousarray(A)
    out[f"{name}_ploidy"] = np.int64(ploidy)
    out[f"{name}_simple_raw"] = np.ascontiguousarray(Gs)
    out[f"{name}_simple_final"] = np.ascontiguousarray(Fs)
    out[f"{name}_simple_repair"] = np.array([its, tots])
    out[f"{name}_ploidy_raw"] = np.ascontiguousarray(Gp)
    out[f"{name}_ploidy_final"] = np.ascontiguousarray(Fp)
    out[f"{name}_ploidy_repair"] = np.array([itp, totp])
    out[f"{name}_q"] = q
    out[f"{name}_d"] = np.float64(d)


case("dip", synth.dosage(12, 40, 2, seed=101), 2)
case("tet", synth.dosage(9, 33, 4, seed=102), 4)
case("dip_as_tet", synth.dosage(11, 64, 2, seed=103), 4)  # ploidy != dosage denominator
case("cont", synth.continuous(10, 50, seed=104), 2)
case("unif", synth.uniform(7, 20, seed=105), 2)
case("wide", synth.dosage(5, 300, 4, seed=106), 4)  # p > 2 K-blocks, n << tile

# inflatediagonals! on the docstring example shape (calc.jl:30-39): rank-1 10 x 10
x = np.random.default_rng(107).random(10)
X = np.outer(x, x)
st, Xo, it, tot = O.c_inflatediagonals(X)
out["rank1_X"] = X
out["rank1_out"] = Xo
out["rank1_repair"] = np.array([it, tot])

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "grm_golden.npz"), **out)
print("wrote", len(out), "arrays")
