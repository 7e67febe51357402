"""GPU bring-up script: exercises every kernel once with diagnostics.  Not a test (tests/ holds those) — it prints
as much as possible per gpurun call.  Usage: python scripts/first_light.py [stage ...]"""
import os
import subprocess
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import grm_b200
from grm_b200 import _lib, device as D, synth

dev = torch.device("cuda", 0)


def sh(cmd):
    try:
        print(subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip())
    except Exception as e:
        print("sh failed", cmd, e)


def ev_time(fn, ctx, reps=3):
    with D.on_ctx_stream(ctx):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
    return min(ts), ts


def stage_env():
    sh("nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv")
    sh("nvidia-smi topo -m | head -12")


def to_dev_colmajor(A):
    # (n,p) F-ordered numpy -> (p,n) contiguous torch
    return torch.from_numpy(np.ascontiguousarray(A.T)).to(dev)


def check_int(ctx, n, p, m, seed=1):
    A = synth.dosage(n, p, ploidy=m, seed=seed)
    codes_ref = np.rint(A * m).astype(np.int64)
    At = to_dev_colmajor(A)
    md = D.detect_denominator(ctx, At)
    codes, colsum, flags = D.pack_dosage(ctx, At, m)
    ctx.synchronize()
    c_h = codes.cpu().numpy()[:, :p].astype(np.int64)
    pad = codes.cpu().numpy()[:, p:]
    ok_codes = np.array_equal(c_h, codes_ref)
    ok_pad = not pad.any()
    ok_sum = np.array_equal(colsum.cpu().numpy()[:p].astype(np.int64), codes_ref.sum(0))
    C = D.syrk_i8(ctx, codes, p)
    ctx.synchronize()
    Ch = C.cpu().numpy().astype(np.int64)  # [j, i] = C[i + j*n]; symmetric anyway
    ref = codes_ref @ codes_ref.T
    iu = np.triu_indices(n)
    got_u = Ch.T[iu]
    ok_C = np.array_equal(got_u, ref[iu])
    print(f"[int n={n} p={p} m={m}] detect={md} flags={flags[0].item()} codes={ok_codes} pad0={ok_pad} colsum={ok_sum} "
          f"C_upper={ok_C}")
    if not ok_C:
        bad = np.argwhere(np.triu(Ch.T != ref))
        print("   mismatches:", len(bad), "first:", bad[:8].tolist())
        for (i, j) in bad[:4]:
            print(f"   ({i},{j}) got {Ch.T[i, j]} want {ref[i, j]}")
        rows = np.unique(bad[:, 0])
        cols = np.unique(bad[:, 1])
        print("   bad rows range", rows.min(), rows.max(), "n rows", len(rows), "| bad cols range", cols.min(), cols.max(),
              "n cols", len(cols))
    return ok_codes and ok_pad and ok_sum and ok_C


def stage_int_small():
    ctx = grm_b200.default_context(0)
    for (n, p, m) in [(100, 1000, 2), (128, 128, 1), (257, 333, 4), (300, 4097, 2), (1000, 10000, 4), (513, 129, 64)]:
        try:
            check_int(ctx, n, p, m)
        except Exception:
            traceback.print_exc()


def stage_int_c2():
    ctx = grm_b200.default_context(0)
    n, p, m = 5000, 100000, 2
    At = synth.dosage_device(n, p, m, 7, dev)
    torch.cuda.synchronize()
    codes, colsum, flags = D.pack_dosage(ctx, At, m)
    ctx.synchronize()
    ldc = codes.shape[1]
    t_pack, all_pack = ev_time(lambda: D.pack_dosage(ctx, At, m, codes=codes, colsum=colsum, flags=flags), ctx)
    print(f"[C2] pack {t_pack:.3f} ms  -> {n*p*8/t_pack/1e6:.0f} GB/s read ({n*p*9/t_pack/1e6:.0f} GB/s r+w)", all_pack)
    C = torch.zeros((n, n), dtype=torch.int32, device=dev)
    t_syrk, all_syrk = ev_time(lambda: D.syrk_i8(ctx, codes, p, out=C), ctx)
    macs = n * n * p / 2
    print(f"[C2] syrk_i8 {t_syrk:.3f} ms -> {macs/t_syrk/1e9:.1f} TMAC/s ({2*macs/t_syrk/1e9:.0f} TOP/s)", all_syrk)
    G = torch.zeros((n, n), dtype=torch.float64, device=dev)
    t_f, all_f = ev_time(lambda: D.syrk_i8_f64(ctx, codes, p, 0.25, 0.0, 0.0, float(p), out=G), ctx)
    print(f"[C2] syrk_i8_f64 {t_f:.3f} ms -> {macs/t_f/1e9:.1f} TMAC/s", all_f)
    # cross-check against cuBLASLt int8 (library, test-only comparator)
    ctx.synchronize()
    cc = codes[:, :p].contiguous() if p % 8 == 0 else codes
    try:
        ref = torch._int_mm(cc, cc.T.contiguous())
        iu = torch.triu_indices(n, n, device=dev)
        ok = torch.equal(C.T[iu[0], iu[1]], ref[iu[0], iu[1]])
        print("[C2] C upper == torch._int_mm:", ok)
        if not ok:
            diff = (C.T != ref).triu()
            idx = diff.nonzero()
            print("   mismatches", idx.shape[0], idx[:8].tolist())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bt = cc.T.contiguous()
        torch._int_mm(cc, bt)
        a.record()
        for _ in range(3):
            torch._int_mm(cc, bt)
        b.record()
        b.synchronize()
        t = a.elapsed_time(b) / 3
        print(f"[C2] cuBLASLt full n x n x p int8 GEMM {t:.3f} ms -> {n*n*p/t/1e9:.1f} TMAC/s")
    except Exception:
        traceback.print_exc()
    # whole device-resident call
    Gout = torch.empty((n, n), dtype=torch.float64, device=dev)
    for _ in range(2):
        _, info = D.grm_device(ctx, At, None, skip_repair=True, out=Gout)
    print("[C2] grm_device simple (no repair):", info)
    _, info = D.grm_device(ctx, At, 2, skip_repair=True, out=Gout)
    _, info = D.grm_device(ctx, At, 2, skip_repair=True, out=Gout)
    print("[C2] grm_device ploidyaware (no repair):", info)
    _, info = D.grm_device(ctx, At, None, skip_repair=False, out=Gout)
    print("[C2] grm_device simple + repair:", info)


def stage_peaks():
    n = 8192
    a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    torch._int_mm(a, b)
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch._int_mm(a, b)
        e.record()
        e.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"[peak] cuBLASLt int8 8192^3 burst: {2*n**3/best/1e9:.0f} TOP/s ({best:.3f} ms)")
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    iters = 0
    s.record()
    while time.time() - t0 < 3.0:
        for _ in range(50):
            torch._int_mm(a, b)
        iters += 50
        torch.cuda.synchronize()
    e.record()
    e.synchronize()
    print(f"[peak] cuBLASLt int8 8192^3 sustained: {2*n**3*iters/s.elapsed_time(e)/1e9:.0f} TOP/s")
    x = torch.randn((n, n), dtype=torch.float64, device=dev)
    y = torch.randn((n, n), dtype=torch.float64, device=dev)
    torch.matmul(x, y)
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(x, y)
        e.record()
        e.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"[peak] cuBLAS dgemm 8192^3 burst: {2*n**3/best/1e9:.1f} TFLOP/s ({best:.3f} ms)")
    # H2D pinned bandwidth
    h = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
    d = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    e.record()
    e.synchronize()
    print(f"[peak] H2D pinned: {4*(1<<28)/s.elapsed_time(e)/1e6:.1f} GB/s")
    s.record()
    for _ in range(4):
        h.copy_(d, non_blocking=True)
    e.record()
    e.synchronize()
    print(f"[peak] D2H pinned: {4*(1<<28)/s.elapsed_time(e)/1e6:.1f} GB/s")


def stage_f64():
    ctx = grm_b200.default_context(0)
    for (n, p) in [(100, 1000), (257, 333), (1000, 5000)]:
        A = synth.continuous(n, p, seed=3)
        At = to_dev_colmajor(A)
        G = D.syrk_f64(ctx, At, False)
        D.scale_mirror(ctx, G, float(p))
        ctx.synchronize()
        ref = (A @ A.T) / p
        err = np.abs(G.cpu().numpy() - ref).max() / np.abs(ref).max()
        q, stats, flags = D.locus_stats_f64(ctx, At)
        ctx.synchronize()
        qe = np.abs(q.cpu().numpy() - A.mean(0)).max()
        Z = 2.0 * (A - 0.5) - A.mean(0)
        G2 = D.syrk_f64(ctx, At, True, 2.0, q)
        d = 2.0 * stats[0].item()
        D.scale_mirror(ctx, G2, d)
        ctx.synchronize()
        ref2 = (Z @ Z.T) / (2.0 * (A.mean(0) * (1 - A.mean(0))).sum())
        err2 = np.abs(G2.cpu().numpy() - ref2).max() / np.abs(ref2).max()
        sym = torch.equal(G2, G2.T)
        print(f"[f64 n={n} p={p}] simple rel err {err:.2e}; q abs err {qe:.2e}; centred rel err {err2:.2e}; symmetric {sym}")
    n, p = 4096, 32768
    At = synth.continuous_device(n, p, 5, dev)
    G = torch.zeros((n, n), dtype=torch.float64, device=dev)
    t, allt = ev_time(lambda: D.syrk_f64(ctx, At, False, out=G), ctx)
    print(f"[f64 n={n} p={p}] syrk_f64 {t:.2f} ms -> {n*n*p/2/t/1e9:.2f} TMAC/s = {n*n*p/t/1e9:.1f} TFLOP/s (useful)", allt)


def stage_host():
    from grm_b200 import Genomes, grmsimple, grmploidyaware, inflatediagonals
    for kind in ("dosage", "continuous"):
        A = synth.dosage(100, 10000, 2, seed=42) if kind == "dosage" else synth.continuous(100, 10000, seed=42)
        g = Genomes.from_matrix(A)
        for fn, kw in ((grmsimple, {}), (grmploidyaware, {"ploidy": 2})):
            t0 = time.time()
            out = fn(g, **kw)
            dt = time.time() - t0
            G = out.genomic_relationship_matrix
            print(f"[host {kind} {fn.__name__}] {dt*1e3:.1f} ms path={out.info['path']} m={out.info['denominator']} "
                  f"iters={out.info['repair_iters']} eps_total={out.info['eps_total']:.6g} G[0,0]={G[0,0]:.6f} "
                  f"sym={np.array_equal(G, G.T)} det>eps={np.linalg.det(G) > np.finfo(float).eps}")
    x = np.random.default_rng(0).random(10)
    X = np.outer(x, x)
    inflatediagonals(X, verbose=True)
    print("[host] rank-1 10x10 det after:", np.linalg.det(X))


def stage_pack_ab():
    """TMA-staged vs LDG pack kernel, alternating in one process, at C3's entry count."""
    ctx = grm_b200.default_context(0)
    n, p, m = 20000, 200000, 4
    At = synth.dosage_device(n, p, m, 9, dev)
    torch.cuda.synchronize()
    codes, colsum, flags = D.pack_dosage(ctx, At, m)
    res = {"tma": [], "ldg": []}
    for rep in range(4):
        for var in ("tma", "ldg"):
            os.environ["GRM_CUDA_PACK"] = var
            t, _ = ev_time(lambda: D.pack_dosage(ctx, At, m, codes=codes, colsum=colsum, flags=flags), ctx, reps=3)
            res[var].append(t)
    os.environ.pop("GRM_CUDA_PACK", None)
    for var, ts in res.items():
        best = min(ts)
        print(f"[pack n={n} p={p}] {var}: best {best:.3f} ms -> {9*n*p/best/1e6:.0f} GB/s r+w; all {['%.3f' % t for t in ts]}")


STAGES = {"pack_ab": stage_pack_ab, "env": stage_env, "int_small": stage_int_small, "int_c2": stage_int_c2, "peaks": stage_peaks,
          "f64": stage_f64, "host": stage_host}

def stage_f64_ab():
    """FP64 SYRK: burst vs sustained, and the cost of centring / mean-fill on operand load."""
    ctx = grm_b200.default_context(0)
    n, p = 4096, 32768
    At = synth.continuous_device(n, p, 5, dev)
    G = torch.zeros((n, n), dtype=torch.float64, device=dev)
    t, _ = ev_time(lambda: D.syrk_f64(ctx, At, False, out=G), ctx)
    print(f"[f64 burst n={n} p={p}] {t:.2f} ms -> {n*n*p/t/1e9:.1f} TFLOP/s")
    with D.on_ctx_stream(ctx):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(150):
            D.syrk_f64(ctx, At, False, out=G)
        b.record()
        b.synchronize()
    t = a.elapsed_time(b) / 150
    print(f"[f64 sustained 150 launches] {t:.2f} ms -> {n*n*p/t/1e9:.1f} TFLOP/s")
    sh("nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader")
    del At, G
    n, p = 10000, 100000
    At = synth.continuous_device(n, p, 6, dev)
    q, stats, flags = D.locus_stats_f64(ctx, At)
    G = torch.zeros((n, n), dtype=torch.float64, device=dev)
    for name, kw in (("simple", dict(centred=False)), ("centred", dict(centred=True, q=q)),
                     ("centred+fill", dict(centred=True, q=q, fill_missing=True))):
        t, ts = ev_time(lambda: D.syrk_f64(ctx, At, ploidy=2.0, out=G, **kw), ctx, reps=2)
        print(f"[f64 n={n} p={p} {name}] {t:.1f} ms -> {n*n*p/t/1e9:.1f} TFLOP/s useful", ts)


STAGES["f64_ab"] = stage_f64_ab


def stage_repair_ab():
    """inflatediagonals!: legacy cusolverDnDgetrf vs the 64-bit cusolverDnXgetrf (ALG_0 / ALG_1)."""
    import ctypes as C
    ctx = grm_b200.default_context(0)
    for (n, p, m, ploidy) in [(5000, 20000, 2, None), (20000, 20000, 4, 4)]:
        At = synth.dosage_device(n, p, m, 3, dev)
        G0, info = D.grm_device(ctx, At, ploidy, denominator=m, skip_repair=True)
        ctx.synchronize()
        base = None
        for var in ("legacy", "x0", "x1", "legacy"):
            os.environ["GRM_CUDA_GETRF"] = var
            G = G0.clone()
            torch.cuda.synchronize()
            it, tot = C.c_int32(0), C.c_double(0.0)
            t0 = time.perf_counter()
            _lib.check(ctx.lib.grm_cuda_inflate_diagonals(ctx.handle, G.data_ptr(), n, n, 1000, _lib.LOC_DEVICE,
                                                          C.byref(it), C.byref(tot)))
            ctx.synchronize()
            dt = time.perf_counter() - t0
            res = (it.value, tot.value)
            same = base is None or (res == base[0] and torch.equal(G, base[1]))
            if base is None:
                base = (res, G)
            print(f"[repair n={n}] {var}: {dt*1e3:.1f} ms iters={it.value} eps_total={tot.value:.6g} identical_to_legacy={same}")
    os.environ.pop("GRM_CUDA_GETRF", None)


STAGES["repair_ab"] = stage_repair_ab


def stage_syrk_ab():
    """2-CTA SYRK at C3 scale with and without the K-window synchronisation, alternating in one process."""
    ctx = grm_b200.default_context(0)
    n, p = 20000, 500000
    ldc = D.padded_loci(p)
    codes = torch.randint(0, 5, (n, ldc), dtype=torch.int8, device=dev)
    codes[:, p:] = 0
    G = torch.zeros((n, n), dtype=torch.float64, device=dev)
    ref = None
    macs = n * n * p / 2
    for rep in range(3):
        for var in ("1", "0"):
            os.environ["GRM_CUDA_SYRK_SYNC"] = var
            t, ts = ev_time(lambda: D.syrk_i8_f64(ctx, codes, p, 1.0 / 16, 0.0, 0.0, float(p), out=G), ctx, reps=2)
            ctx.synchronize()
            chk = float(G.sum())
            if ref is None:
                ref = chk
            print(f"[syrk C3 sync={var}] {t:.2f} ms -> {2*macs/t/1e9:.0f} TOP/s  {['%.2f' % x for x in ts]} same={chk == ref}", flush=True)
    os.environ.pop("GRM_CUDA_SYRK_SYNC", None)
    n, p = 5000, 100000
    codes = torch.randint(0, 3, (n, D.padded_loci(p)), dtype=torch.int8, device=dev)
    codes[:, p:] = 0
    G = torch.zeros((n, n), dtype=torch.float64, device=dev)
    for var in ("1", "0", "1", "0"):
        os.environ["GRM_CUDA_SYRK_SYNC"] = var
        t, ts = ev_time(lambda: D.syrk_i8_f64(ctx, codes, p, 0.25, 0.0, 0.0, float(p), out=G), ctx, reps=3)
        print(f"[syrk C2 sync={var}] {t:.3f} ms -> {n*n*p/t/1e9:.0f} TOP/s", flush=True)
    os.environ.pop("GRM_CUDA_SYRK_SYNC", None)


STAGES["syrk_ab"] = stage_syrk_ab


def stage_syrk_window():
    """Sweep of the K-window of the pair synchronisation at C3 scale."""
    ctx = grm_b200.default_context(0)
    n, p = 20000, 500000
    codes = torch.randint(0, 5, (n, D.padded_loci(p)), dtype=torch.int8, device=dev)
    codes[:, p:] = 0
    G = torch.zeros((n, n), dtype=torch.float64, device=dev)
    for rep in range(2):
        for w in ("16", "32", "64", "128", "256", "512"):
            os.environ["GRM_CUDA_SYRK_WINDOW"] = w
            t, ts = ev_time(lambda: D.syrk_i8_f64(ctx, codes, p, 1.0 / 16, 0.0, 0.0, float(p), out=G), ctx, reps=2)
            print(f"[syrk C3 window={w}] {t:.2f} ms {['%.2f' % x for x in ts]}", flush=True)
    os.environ.pop("GRM_CUDA_SYRK_WINDOW", None)


STAGES["syrk_window"] = stage_syrk_window


def stage_pipeline_ab():
    """C3 step (device-resident input) with the serial flow and the pipelined flow at several chunk counts."""
    from grm_b200 import synth
    ctx = grm_b200.default_context(0)
    n, p, ploidy = 20000, 500000, 4
    A = torch.empty((p, n), dtype=torch.float64, device=dev)
    for c0 in range(0, p, 2500):
        A[c0:c0 + 2500] = synth.dosage_device(n, 2500, ploidy, 1000 + c0 // 2500, dev, chunk=2500)
    G = torch.empty((n, n), dtype=torch.float64, device=dev)
    ref = None
    for rep in range(2):
        for ch in ("1", "2", "3", "4", "6", "8"):
            os.environ["GRM_CUDA_PIPELINE"] = ch
            infos = []
            t, ts = ev_time(lambda: infos.append(D.grm_device(ctx, A, ploidy, denominator=ploidy, skip_repair=True, out=G)[1]),
                            ctx, reps=3)
            i = infos[-1]
            if ref is None:
                ref = G.clone()
            same = bool(torch.equal(ref, G))
            print(f"[C3 pipeline chunks={ch}] {t:.2f} ms {['%.2f' % x for x in ts]} pack {i['ms_pack']:.2f} stats "
                  f"{i['ms_rowdot']:.2f} syrk {i['ms_syrk']:.2f} launches {i['kernel_launches']} identical={same}", flush=True)
    os.environ.pop("GRM_CUDA_PIPELINE", None)


STAGES["pipeline_ab"] = stage_pipeline_ab


def stage_pack_occ():
    """LDG pack kernel at 2 CTAs/SM (default) and capped to 1 CTA/SM by padding shared memory: is the co-resident
    pack of the pipelined flow limited by occupancy or by contention with the SYRK?"""
    ctx = grm_b200.default_context(0)
    n, p, m = 20000, 200000, 4
    At = synth.dosage_device(n, p, m, 9, dev)
    torch.cuda.synchronize()
    codes, colsum, flags = D.pack_dosage(ctx, At, m)
    os.environ["GRM_CUDA_PACK"] = "ldg"
    for rep in range(2):
        for pad in ("0", "120"):
            os.environ["GRM_CUDA_PACK_PAD"] = pad
            t, ts = ev_time(lambda: D.pack_dosage(ctx, At, m, codes=codes, colsum=colsum, flags=flags), ctx, reps=3)
            print(f"[pack ldg pad={pad} KB] {t:.3f} ms -> {9*n*p/t/1e6:.0f} GB/s r+w {['%.3f' % x for x in ts]}", flush=True)
    os.environ.pop("GRM_CUDA_PACK", None)
    os.environ.pop("GRM_CUDA_PACK_PAD", None)


STAGES["pack_occ"] = stage_pack_occ


def stage_distances():
    """distances(genomes) pair kernel: entries x entries and loci x loci at C3's entry count (device-resident input)."""
    import ctypes as C
    from grm_b200 import _lib
    ctx = grm_b200.default_context(0)
    for (n, p, t, what) in [(20000, 2000, 100, ("euclidean", "correlation", "mad", "rmsd", "chi2")),
                            (20000, 2000, 100, ("euclidean",)), (5000, 4000, 2000, ("euclidean", "correlation", "mad", "rmsd", "chi2")),
                            (1000, 8192, 8192, ("correlation",))]:  # estimateld scale: one chromosome of 8192 loci
        At = synth.continuous_device(n, p, 3, dev)
        idx0 = np.sort(np.random.default_rng(0).choice(p, t, replace=False)).astype(np.int64)
        for dim, r, c in ((_lib.DIM_ENTRIES, n, t), (_lib.DIM_LOCI, t, n)):
            if what == ("correlation",) and dim == _lib.DIM_ENTRIES:
                continue
            mats = {k: torch.empty((r, r), dtype=torch.float64, device=dev) for k in what}
            cnt = torch.empty((r, r), dtype=torch.float64, device=dev)
            ptr = lambda k: C.c_void_p(mats[k].data_ptr()) if k in mats else None
            torch.cuda.synchronize()
            def call():
                _lib.check(ctx.lib.grm_cuda_distances(ctx.handle, C.c_void_p(At.data_ptr()), n, p, n, _lib.LOC_DEVICE,
                                                      idx0.ctypes.data, t, dim, ptr("euclidean"), ptr("correlation"), ptr("mad"),
                                                      ptr("rmsd"), ptr("chi2"), C.c_void_p(cnt.data_ptr()), r, _lib.LOC_DEVICE))
            tms, ts = ev_time(call, ctx, reps=3)
            pe = float(r) * r * c
            print(f"[distances n={n} t={t} dim={'entries' if dim else 'loci'} metrics={len(what)}] {tms:.3f} ms -> "
                  f"{pe / tms / 1e6:.1f} G pair-elements/s ({r}x{r} pairs x {c} positions)", flush=True)
        del At


STAGES["distances"] = stage_distances


def stage_syrk_quad():
    """Clusters of 4 (two CTA pairs, shared column panel multicast) against the CTA-pair kernel: bit-identity on ragged
    shapes (odd and even tile counts, both epilogues), then timing at C2 and C3."""
    os.environ["GRM_CUDA_DEBUG"] = "1"
    MODES = os.environ.get("SYRK_MODES", "2cta,wide").split(",")
    ctx = grm_b200.default_context(0)
    for (n, p, m, ploidy) in [(300, 1000, 2, None), (700, 3001, 4, 4), (1100, 777, 2, None), (2049, 5001, 4, 2), (5000, 20000, 2, 2)]:
        At = synth.dosage_device(n, p, m, 5, dev)
        os.environ["GRM_CUDA_SYRK_I8"] = "2cta"
        G0 = D.grm_device(ctx, At, ploidy, denominator=m, skip_repair=True)[0].clone()
        res = {}
        for mode in MODES[1:]:
            os.environ["GRM_CUDA_SYRK_I8"] = mode
            G1 = D.grm_device(ctx, At, ploidy, denominator=m, skip_repair=True)[0]
            torch.cuda.synchronize()
            res[mode] = bool(torch.equal(G0, G1))
        print(f"[n={n} p={p} ploidy={ploidy}] tiles={((n + 255) // 256) * ((n + 255) // 256 + 1) // 2} bit-identical={res}",
              flush=True)
    # C2: SYRK alone
    n2, p2 = 5000, 100000
    codes = torch.randint(0, 3, (n2, D.padded_loci(p2)), dtype=torch.int8, device=dev)
    codes[:, p2:] = 0
    G2 = torch.zeros((n2, n2), dtype=torch.float64, device=dev)
    for rep in range(2):
        for mode in MODES:
            os.environ["GRM_CUDA_SYRK_I8"] = mode
            t, ts = ev_time(lambda: D.syrk_i8_f64(ctx, codes, p2, 0.25, 0.0, 0.0, float(p2), out=G2), ctx, reps=5)
            print(f"[C2 syrk {mode}] best {t:.3f} ms", flush=True)
    del codes, G2
    # C3: whole step
    n, p, ploidy = 20000, 500000, 4
    A = torch.empty((p, n), dtype=torch.float64, device=dev)
    for c0 in range(0, p, 2500):
        A[c0:c0 + 2500] = synth.dosage_device(n, 2500, ploidy, 1000 + c0 // 2500, dev, chunk=2500)
    G = torch.empty((n, n), dtype=torch.float64, device=dev)
    ref = None
    for rep in range(3):
        for mode in MODES:
            os.environ["GRM_CUDA_SYRK_I8"] = mode
            infos = []
            t, ts = ev_time(lambda: infos.append(D.grm_device(ctx, A, ploidy, denominator=ploidy, skip_repair=True, out=G)[1]),
                            ctx, reps=4)
            if ref is None:
                ref = G.clone()
            print(f"[C3 {mode}] best {t:.2f} ms mean {sum(ts)/len(ts):.2f} syrk {infos[-1]['ms_syrk']:.2f} "
                  f"identical={bool(torch.equal(ref, G))}", flush=True)
    os.environ.pop("GRM_CUDA_SYRK_I8", None)


STAGES["syrk_quad"] = stage_syrk_quad


if __name__ == "__main__":
    names = sys.argv[1:] or list(STAGES)
    for nm in names:
        print(f"===== {nm} =====", flush=True)
        try:
            STAGES[nm]()
        except Exception:
            traceback.print_exc()
        sys.stdout.flush()

