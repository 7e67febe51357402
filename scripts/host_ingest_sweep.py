"""Host-only sweep of the ingest quantiser (no GPU work) on this machine's CPU: threads x software-prefetch distance, then
prefetch hint x streaming stores x distance at one thread per CPU.
    python scripts/host_ingest_sweep.py [n p] [--hints]   -> JSON lines, GB/s of source read (8 B payload + 1 B selector per cell)"""
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
hints_only = "--hints" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(argv[0]) if len(argv) > 0 else 20000
p = int(argv[1]) if len(argv) > 1 else 16000
A = np.empty((p, n))
for k in range(0, p, 2000):
    A[k:k + 2000] = np.random.default_rng(k).integers(0, 5, (min(2000, p - k), n)) / 4.0
sel = np.ones((p, n), dtype=np.uint8)
codes = np.zeros((p, n), dtype=np.int8)
fl = C.c_uint32()
ncpu = os.cpu_count() or 1
print(json.dumps({"simd": lib.grm_cuda_host_simd().decode(), "cpus": ncpu, "n": n, "p": p}))

def timed(thr, reps=3):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        lib.grm_cuda_host_quantise(A.ctypes.data, sel.ctypes.data, 1, n, p, n, 4, codes.ctypes.data, n, thr, C.byref(fl))
        best = min(best, time.perf_counter() - t)
    return best


for hint in (0, 1, 2, 3):
    for stream in (0, 1):
        for pf in (4096, 8192, 16384, 32768):
            lib.grm_cuda_host_tune(b"prefetch_hint", hint)
            lib.grm_cuda_host_tune(b"stream_stores", stream)
            lib.grm_cuda_host_tune(b"prefetch", pf)
            best = timed(ncpu, 4)
            print(json.dumps({"hint": ["T0", "T1", "T2", "NTA"][hint], "stream_stores": stream, "prefetch": pf, "threads": ncpu,
                              "ms": round(best * 1e3, 1), "src_gbs": round(9.0 * n * p / best / 1e9, 1)}), flush=True)
for key in (b"prefetch_hint", b"stream_stores", b"prefetch"):
    lib.grm_cuda_host_tune(key, -1)  # defaults
for pf in (() if hints_only else (0, 1024, 2048, 4096, 8192, 16384)):
    lib.grm_cuda_host_tune(b"prefetch", pf)
    for thr in sorted({max(1, ncpu // 4), max(1, ncpu // 2), ncpu, ncpu * 3 // 2, 2 * ncpu}):
        best = 1e9
        for _ in range(3):
            t = time.perf_counter()
            lib.grm_cuda_host_quantise(A.ctypes.data, sel.ctypes.data, 1, n, p, n, 4, codes.ctypes.data, n, thr, C.byref(fl))
            best = min(best, time.perf_counter() - t)
        print(json.dumps({"prefetch": pf, "threads": thr, "ms": round(best * 1e3, 1), "src_gbs": round(9.0 * n * p / best / 1e9, 1)}))
