"""BASELINE config 4 at full size on one GPU: 10,000 pools x 1,000,000 loci continuous pool-seq frequencies with
missing cells (mean-fill policy), grmploidyaware, FP64 DMMA SYRK.  Checks: exact symmetry; sampled G_ij against a
direct FP64 evaluation of the reference formula on the mean-imputed columns (torch, comparator only)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import grm_b200
from grm_b200 import _lib, device as D, synth

dev = torch.device("cuda", 0)
ctx = grm_b200.Context(0)
n, p, ploidy = int(os.environ.get("C4_N", 10_000)), int(os.environ.get("C4_P", 1_000_000)), 2
sparsity = 0.01
A = synth.continuous_device(n, p, 42, dev, chunk=4096)
g = torch.Generator(device=dev)
g.manual_seed(7)
for k0 in range(0, p, 65536):  # ~1 % of the cells become missing
    k1 = min(p, k0 + 65536)
    A[k0:k1][torch.rand((k1 - k0, n), generator=g, device=dev) < sparsity] = float("nan")
torch.cuda.synchronize()
o = dict(path=_lib.PATH_AUTO, skip_repair=True)
G = torch.empty((n, n), dtype=torch.float64, device=dev)


def call():
    oo = _lib.default_options()
    oo.skip_repair = 1
    oo.input_location = oo.output_location = _lib.LOC_DEVICE
    oo.missing_policy = _lib.MISSING_MEANFILL
    info = _lib.new_info()
    import ctypes as C

    _lib.check(ctx.lib.grm_cuda_ploidyaware(ctx.handle, A.data_ptr(), n, p, n, ploidy, G.data_ptr(), n, C.byref(oo),
                                            C.byref(info)))
    return info.as_dict()


info = call()
t0 = time.perf_counter()
info = call()
dt = time.perf_counter() - t0
ctx.synchronize()
# spot check on sampled entries: the reference formula on the mean-imputed columns
idx = torch.tensor([0, 1, 17, n // 2, n - 1], device=dev)
valid = ~torch.isnan(A)
cnt = valid.sum(1, keepdim=True).double()
q = torch.where(valid, A, torch.zeros((), dtype=torch.float64, device=dev)).sum(1, keepdim=True) / cnt  # [p, 1]
d = ploidy * float((q * (1 - q)).sum())
rows = A[:, idx]
rows = torch.where(torch.isnan(rows), q.expand_as(rows), rows)
Z = ploidy * (rows - 0.5) - q                       # [p, 5]
ref = (Z.T @ Z) / d
got = G[idx][:, idx]
rel = float((got - ref).abs().max() / ref.abs().max())
sym = bool(torch.equal(G, G.T))
macs = n * n * p / 2
print(json.dumps({"config": f"C4 {n} x {p} continuous, {sparsity:.0%} missing, mean-fill, grmploidyaware (FP64 path)",
                  "path": info["path"], "seconds": dt, "mac_per_s": macs / dt, "tflops_useful": 2 * macs / dt / 1e12,
                  "ms_syrk_stage": info["ms_pack"], "symmetric": sym, "spot_rel_err": rel, "scale_d": info["scale"],
                  "d_check_rel": abs(info["scale"] - d) / d}))
sys.exit(0 if (sym and rel < 1e-10) else 1)
