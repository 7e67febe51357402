#!/usr/bin/env python
"""bench.py — GRM hot path on B200: n^2 p / 2 MAC/s (BASELINE.json metric) on synthetic dosage data.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port), rank 0 only

One "step" = one pass of the hot path over the workload: FP64 allele frequencies already resident in HBM ->
column statistics + int8 pack -> locus stats / row dots -> tcgen05 int8 SYRK with the FP64 epilogue -> mirror
(= the raw GRM, src/grm/calc.jl:189-209).  `value` times exactly that.  `e2e` times the reference-facing call
(grm_cuda_ploidyaware through the C ABI) on HOST buffers: pinned H2D of the FP64 matrix, the same device pipeline,
the inflatediagonals! repair (calc.jl:211) and the D2H of the n x n result are all inside the timed region.

Workloads (BASELINE.json configs): c3 = 20,000 x 500,000 tetraploid grmploidyaware (default: the configuration the
north_star target is quoted on; it fits one GPU), c2 = 5,000 x 100,000 diploid grmsimple, c1 = 100 x 10,000.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(n=100, p=10_000, ploidy=2, fn="grmsimple", gen_chunk=2500,
               name="C1 simulategenomes-like 100 x 10,000 diploid grmsimple"),
    "c2": dict(n=5_000, p=100_000, ploidy=2, fn="grmsimple", gen_chunk=2500,
               name="C2 5,000 x 100,000 diploid grmsimple (int8 dosage path)"),
    "c3": dict(n=20_000, p=500_000, ploidy=4, fn="grmploidyaware", gen_chunk=2500,
               name="C3 20,000 x 500,000 tetraploid grmploidyaware (int8 dosage path)"),
}
METRIC = "grm_mac_per_s"
UNIT = "n^2*p/2 MAC/s"


def macs(n, p):
    return n * n * p / 2.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the upper half: samples taken while the GPU was busy
        busy = sm[len(sm) // 2:] if sm else []
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(wl, A_sub, rows, nthreads):
    """Reference algorithm (oracle port, oracle/grm_oracle.c) on a bounded sample: pairs (i < rows, j >= i) over the
    loci of A_sub (n x p_sub, column-major).  Returns (MAC/s, seconds, MACs)."""
    from oracle import oracle as O

    n, ps = A_sub.shape
    t0 = time.perf_counter()
    if wl["fn"] == "grmsimple":
        st, _ = O.c_grmsimple_raw(A_sub, rows_limit=rows, nthreads=nthreads)
    else:
        st, _, _, _ = O.c_grmploidyaware_raw(A_sub, wl["ploidy"], rows_limit=rows, nthreads=nthreads)
    dt = time.perf_counter() - t0
    assert st == 0
    pairs = sum(n - i for i in range(rows))
    return pairs * ps / dt, dt, pairs * ps


def cpu_rows_for(wl, A_sub, target_s, nthreads):
    """Rows of the sample such that one pass takes about target_s seconds (calibrated on a one-row-per-thread probe)."""
    n = wl["n"]
    rows = max(1, min(n, nthreads))
    _, dt, _ = cpu_sample(wl, A_sub, rows, nthreads)
    want = int(rows * target_s / max(dt, 1e-3))
    return max(1, min(n, max(nthreads, want // nthreads * nthreads)))


def cpu_sample_loci(wl):
    return min(wl["p"], max(2000, int(1.0e9 / (8 * wl["n"]))))  # <= 1 GB host matrix


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import grm_b200  # noqa: F401  (import shim for grm_b200.synth)
    from grm_b200 import synth

    nthreads = os.cpu_count() or 1
    p_sub = cpu_sample_loci(wl)
    A = synth.dosage(wl["n"], p_sub, wl["ploidy"], seed=42)
    rows = cpu_rows_for(wl, A, 4.0, nthreads)
    for _ in range(args.warmup):
        cpu_sample(wl, A, max(1, rows // 4), nthreads)
    t_tot, m_tot = 0.0, 0.0
    for _ in range(args.steps):
        _, dt, m = cpu_sample(wl, A, rows, nthreads)
        t_tot += dt
        m_tot += m
    value = m_tot / t_tot
    sample = (f"oracle port of the src/grm/calc.jl pair loop ({wl['fn']}): pairs (i<{rows}, j>=i) of n={wl['n']} entries "
              f"over the first {p_sub} of {wl['p']} loci per step, {nthreads} host threads (Threads.@threads-style "
              f"contiguous chunks); MAC/s of the sample, which the algorithm sustains independently of the sample size")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "n": wl["n"], "p": wl["p"], "ploidy": wl["ploidy"], "function": wl["fn"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def ncu_traffic():
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return {}
    return {}


def run_b200(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist

    import grm_b200
    from grm_b200 import _lib, device as D, dist as GD, grm_call, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch multi-GPU runs with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = grm_b200.Context(local)
    peaks = load_peaks()
    n, p, ploidy = wl["n"], wl["p"], wl["ploidy"]
    m = ploidy  # generator emits c/ploidy; ploidy 2 and 4 are power-of-two denominators
    centred = wl["fn"] == "grmploidyaware"
    k0, k1 = GD.loci_range(p, rank, world)
    pc = k1 - k0

    # ---- synthetic input, generated on the device in fixed 2500-locus chunks (same data for every world size) ----
    gc = wl["gen_chunk"]
    A_dev = torch.empty((pc, n), dtype=torch.float64, device=dev)
    for c0 in range(0, p, gc):
        lo, hi = max(c0, k0), min(c0 + gc, k1)
        if lo >= hi:
            continue
        blk = synth.dosage_device(n, min(gc, p - c0), ploidy, 1000 + c0 // gc, dev, chunk=gc)
        A_dev[lo - k0:hi - k0] = blk[lo - c0:hi - c0]
        del blk
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident step --------------------------------------------------------------------------------
    G_dev = None
    sharded = None
    stage_ms = {}
    if world == 1:
        G_dev = torch.empty((n, n), dtype=torch.float64, device=dev)

        def step():
            _, info = D.grm_device(ctx, A_dev, ploidy if centred else None, denominator=m, skip_repair=True, out=G_dev)
            return info
    else:
        sharded = GD.ShardedGRM(n, p, ploidy if centred else None, m, ctx)

        def step():
            t = {}
            sharded.step(A_dev, timings=t)
            return t

    launches = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # sampled from the warm-up on: short multi-GPU timed regions would otherwise see no sample
    for _ in range(args.warmup):
        step()
    barrier()
    with D.on_ctx_stream(ctx):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        infos = [step() for _ in range(args.steps)]
        e1.record()
        e1.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    ms_step = ms / args.steps
    value = macs(n, p) / (ms_step * 1e-3)
    for key in ("ms_pack", "ms_exchange", "ms_rowdot", "ms_syrk", "ms_finalize"):
        vals = [i[key] for i in infos if key in i]
        if vals:
            stage_ms[key] = sum(vals) / len(vals)
    if world == 1:
        launches = sum(i["kernel_launches"] for i in infos)
    else:
        launches = sharded.launches * args.steps // (args.steps + args.warmup)

    # ---- roofline of the dominant kernel (int8 tcgen05 SYRK) and of the HBM-bound pack kernel --------------------
    T = (n + 255) // 256
    tiles_rank = _lib.load().grm_cuda_tile_count(n, rank, world)
    syrk_ms = stage_ms.get("ms_syrk", float("nan"))
    if world > 1:
        t = torch.tensor([syrk_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        syrk_ms = float(t[0])
    alg_ops = 2.0 * macs(n, p) / world  # algorithmic int8 ops per launch per GPU: n^2 p / 2 MACs dealt over the ranks
    achieved = alg_ops / (syrk_ms * 1e-3) / 1e12
    # sustained peak for a kernel that keeps the tensor pipe busy for most of a long run (power-capped regime),
    # burst peak for short launches (B200_PROFILING.md)
    busy_ms = syrk_ms * (args.steps + args.warmup)
    regime = "sustained" if busy_ms >= 400.0 else "burst"
    int8_peak = 2.0 * (peaks["bf16_sustained"] if regime == "sustained" else peaks["bf16_burst"])
    tr = ncu_traffic() if world == 1 else {}
    wk = args.workload
    roofline = {
        "bound": "tensor", "kernel": ("syrk_i8_wide_kernel" if (world == 1 and (p + 127) // 128 >= 2048 and T * (T + 1) // 2 >= 296)
                                     else "syrk_i8_2cta_kernel") + " (tcgen05.mma cta_group::2 kind::i8, TMA, TMEM)", "achieved": achieved,
        "peak": int8_peak, "unit": "TFLOP/s", "frac": achieved / int8_peak,
        "frac_of_nominal_int8": achieved / 4500.0, "frac_of_burst_derived": achieved / (2.0 * peaks["bf16_burst"]),
        "traffic": tr.get(f"syrk_i8_dram_bytes_per_launch_{wk}"),
        "peak_note": f"int8 dense = 2 x measured {regime} bf16 ({peaks['source']}); unit counts int8 ops (1 MAC = 2 ops); "
                     f"cuBLASLt int8 8192^3 measured on this pool: 3213 burst / 2579 sustained TOP/s (profiles/). "
                     f"A fraction above 1 means the kernel does more int8 work per joule than cuBLAS does bf16 work "
                     f"(the denominator is a power-capped library GEMM), not skipped work: results are bit-exact "
                     f"against cuBLASLt int8 and the CPU oracle (tests/test_gpu_parity.py)",
        "ms_per_launch": syrk_ms, "algorithmic_ops_per_launch": alg_ops, "tiles_owned": int(tiles_rank),
    }
    if world == 1 and infos and infos[0]["kernel_launches"] > 6:
        roofline["launches_per_step"] = 4
        roofline["note"] = ("pipelined flow: the loci are contracted in 4 chunk launches (the pack of the next chunk and "
                            "the integer statistics run underneath on a second stream); ms_per_launch and "
                            "algorithmic_ops_per_launch are the sums over the 4 launches of one step")

    pack_ms = stage_ms.get("ms_pack", float("nan"))
    pack_note = "timed inside the step"
    if world == 1 and infos and infos[0]["kernel_launches"] > 6:
        # pipelined flow: the step's pack runs on the auxiliary stream underneath the SYRK launches, so its span is not
        # a kernel time; time the pack kernel alone (same input, CUDA events on the launching stream)
        codes_b, colsum_b, flags_b = D.pack_dosage(ctx, A_dev, m)
        with D.on_ctx_stream(ctx):
            pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pe0.record()
            for _ in range(3):
                D.pack_dosage(ctx, A_dev, m, codes=codes_b, colsum=colsum_b, flags=flags_b)
            pe1.record()
            pe1.synchronize()
        pack_ms = pe0.elapsed_time(pe1) / 3
        pack_note = ("timed alone, 3 launches after the timed region (inside the pipelined step 3/4 of the pack runs "
                     "underneath the SYRK launches at a lower rate)")
        del codes_b, colsum_b, flags_b
    pack_bytes = 9.0 * n * pc  # 8 B read + 1 B written per cell
    roofline_pack = {
        "bound": "hbm", "kernel": "pack_dosage_kernel", "achieved": pack_bytes / (pack_ms * 1e-3) / 1e9,
        "peak": peaks["hbm"], "unit": "GB/s", "frac": pack_bytes / (pack_ms * 1e-3) / 1e9 / peaks["hbm"],
        "traffic": tr.get(f"pack_dram_bytes_per_launch_{wk}"), "ms_per_launch": pack_ms, "timing": pack_note,
    }

    # ---- end to end through the C ABI with host buffers -----------------------------------------------------------
    e2e = None
    e2e_detail = {}
    A_host = None
    e2e_skip = None
    if not args.no_e2e:
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = None
        need = 8 * n * pc + (8 * n * n if rank == 0 else 0)
        if avail is not None and avail < 1.15 * need * world + (2 << 30):  # every rank allocates its share of the host
            e2e_skip = f"host memory: {avail / 2**30:.0f} GiB available, {need / 2**30:.0f} GiB needed for the host buffers"
    if not args.no_e2e and e2e_skip is None:
        try:
            A_host = torch.empty((pc, n), dtype=torch.float64, pin_memory=True)
            pinned = True
        except Exception:
            A_host = torch.empty((pc, n), dtype=torch.float64)
            pinned = False
        A_host.copy_(A_dev)
        torch.cuda.synchronize()
        esteps = args.e2e_steps or args.steps
        h2d = 8 * n * pc
        if world == 1:
            del A_dev
            torch.cuda.empty_cache()
            A_np = A_host.numpy().T  # (n, p) column-major view, no copy
            G_host = torch.empty((n, n), dtype=torch.float64, pin_memory=pinned)
            G_np = G_host.numpy().T
            pl = ploidy if centred else None
            _, info = grm_call(A_np, pl, ctx=ctx, denominator=0, out=G_np)  # warm-up (allocations, cuSOLVER handle)
            barrier()
            t0 = time.perf_counter()
            for _ in range(esteps):
                _, info = grm_call(A_np, pl, ctx=ctx, denominator=0, out=G_np)
                launches_e2e = info["kernel_launches"]
            barrier()
            dt = (time.perf_counter() - t0) / esteps
            d2h = 8 * n * n
            e2e_detail = {k: info[k] for k in ("ms_h2d", "ms_syrk", "ms_repair", "ms_d2h", "ms_total", "repair_iters",
                                               "eps_total", "denominator", "path")}
            e2e_detail["pinned"] = pinned
            e2e_detail["includes"] = "H2D (chunked, overlapped with pack), detect, pack, stats, rowdot, SYRK, mirror, inflatediagonals!, D2H"
        else:
            G_host = torch.empty((n, n), dtype=torch.float64, pin_memory=pinned) if rank == 0 else None
            A_in = torch.empty((pc, n), dtype=torch.float64, device=dev)
            import ctypes as C

            def e2e_step():
                with D.on_ctx_stream(ctx):
                    A_in.copy_(A_host, non_blocking=True)
                sharded.step(A_in)
                G = sharded.gather()
                # inflatediagonals! with the rounds probed in parallel, one per rank (every rank holds G after gather)
                it, tot = GD.inflatediagonals_parallel(ctx, G, 1000)
                if rank == 0:
                    with D.on_ctx_stream(ctx):
                        G_host.copy_(G, non_blocking=True)
                ctx.synchronize()
                return it, tot
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(esteps):
                it, tot = e2e_step()
            barrier()
            dt = (time.perf_counter() - t0) / esteps
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
            d2h = 8 * n * n if rank == 0 else 0
            e2e_detail = {"repair_iters": it, "eps_total": tot, "pinned": pinned,
                          "includes": "per-rank H2D of its loci slab, pack, integer statistics + int64 all-reduce, SYRK over own loci, "
                                      "int32 reduce-scatter onto the tile owners, finalize, all-reduce gather of G, "
                                      "inflatediagonals! with its rounds probed in parallel (one per rank), D2H on rank 0"}
        e2e = {"value": macs(n, p) / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "steps": esteps, **e2e_detail}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        nthreads = os.cpu_count() or 1
        p_sub = cpu_sample_loci(wl)
        if A_host is not None:
            A_sub = np.asfortranarray(A_host[:p_sub].numpy().T)
        else:
            A_sub = synth.dosage(n, p_sub, ploidy, seed=42)
        rows = cpu_rows_for(wl, A_sub, 12.0, nthreads)
        rate, dt_cpu, m_cpu = cpu_sample(wl, A_sub, rows, nthreads)
        # not the reference's algorithm, reported so the GPU is not compared with the pair loop only: the same
        # contraction as one OpenBLAS dgemm (NumPy) on all host cores, on a 4096-entry x p_sub sample
        blas = None
        try:
            ns = min(n, 4096)
            Xs = np.ascontiguousarray(A_sub[:ns])
            t0 = time.perf_counter()
            _ = Xs @ Xs.T
            dtb = time.perf_counter() - t0
            blas = {"value": ns * ns * p_sub / 2.0 / dtb, "unit": UNIT, "what": f"NumPy/OpenBLAS X@X.T on {ns} x {p_sub} "
                    f"(n^2 p/2 useful MACs of the n^2 p computed), {nthreads} host threads", "seconds": dtb}
        except Exception as ex:  # pragma: no cover
            blas = {"error": repr(ex)}
        # the reference README's own setting is `julia --threads 2`: the same port on 2 threads, smaller sample
        two = None
        if nthreads > 2:
            rows2 = max(2, rows * 2 // nthreads // 2)  # about half the wall time of the all-threads sample
            rate2, dt2, m2 = cpu_sample(wl, A_sub, rows2, 2)
            two = {"value": rate2, "unit": UNIT, "cores": 2, "seconds": dt2,
                   "sample": f"pairs (i<{rows2}, j>=i) over the first {p_sub} loci = {m2:.3g} MACs"}
        cpu = {"value": rate, "unit": UNIT, "cores": nthreads, "kind": "port", "optimised_blas": blas, "two_threads": two,
               "sample": f"oracle port of the src/grm/calc.jl pair loop ({wl['fn']}): pairs (i<{rows}, j>=i) of n={n} over the "
                         f"first {p_sub} of {p} loci = {m_cpu:.3g} MACs in {dt_cpu:.1f} s on {nthreads} threads",
               "seconds": dt_cpu}

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8",
        "data": "synthetic",
        "config": {"workload": wl["name"], "n": n, "p": p, "ploidy": ploidy, "function": wl["fn"],
                   "dosage_denominator": m, "path": "int8 tcgen05 SYRK + FP64 rank-1 epilogue",
                   "step": "FP64 A resident in HBM -> pack/stats -> rowdot -> SYRK -> mirror (raw GRM); the "
                           "inflatediagonals! repair is in e2e, not in value",
                   "l2": "inputs larger than L2 (FP64 input and int8 codes >> 126 MB)",
                   "parallelism": f"loci-split pack + tcgen05 SYRK over own loci + int32 reduce-scatter onto tile owners x{world} ({sharded.mode})" if world > 1
                   else "single GPU", "macs_per_step": macs(n, p)},
        "clocks": clocks, "e2e": e2e if e2e is not None else ({"unavailable": e2e_skip} if e2e_skip else None),
        "gpu_launches": int(launches), "roofline": roofline, "roofline_pack": roofline_pack,
        "cpu_baseline": cpu, "stage_ms": stage_ms,
    }
    if world > 1:
        dist.destroy_process_group()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GRM_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    args.warmup = max(args.warmup, 3)
    result = run_b200(args, wl)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(result))


if __name__ == "__main__":
    main()
