"""Host-only sweep of the ingest quantiser (no GPU work): threads x software-prefetch distance on this machine's CPU.
    python scripts/host_ingest_sweep.py [n p]        -> JSON lines, GB/s of source read (8 B payload + 1 B selector per cell)"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grm_b200  # noqa: E402,F401
from grm_b200 import _lib  # noqa: E402

lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
A = np.empty((p, n))
for k in range(0, p, 2000):
    A[k:k + 2000] = np.random.default_rng(k).integers(0, 5, (min(2000, p - k), n)) / 4.0
sel = np.ones((p, n), dtype=np.uint8)
codes = np.zeros((p, n), dtype=np.int8)
fl = C.c_uint32()
ncpu = os.cpu_count() or 1
print(json.dumps({"simd": lib.grm_cuda_host_simd().decode(), "cpus": ncpu, "n": n, "p": p}))
for pf in (0, 1024, 2048, 4096, 8192, 16384):
    lib.grm_cuda_host_tune(b"prefetch", pf)
    for thr in sorted({max(1, ncpu // 4), max(1, ncpu // 2), ncpu, ncpu * 3 // 2, 2 * ncpu}):
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            lib.grm_cuda_host_quantise(A.ctypes.data, sel.ctypes.data, 1, n, p, n, 4, codes.ctypes.data, n, thr, C.byref(fl))
            best = min(best, time.perf_counter() - t)
        print(json.dumps({"prefetch": pf, "threads": thr, "ms": round(best * 1e3, 1), "src_gbs": round(9.0 * n * p / best / 1e9, 1)}))
