"""Multi-GPU parity check of the distributed path INSIDE libgrm_cuda (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/dist_check.py

Every rank builds the same seeded matrix, computes the single-GPU result on its own GPU (non-distributed call) and then
makes the distributed call (`opts.distributed = 1`, the communicator of csrc/comm.cu; torch.distributed — gloo, CPU —
only broadcasts the 128-byte id).  Bars: int8 path bit-identical for any number of ranks (exact integers), FP64 path
<= 1e-13 relative; identical repair rounds / eps_total; identical statuses on every rank when only ONE rank's loci hold
the offending cell.  Prints DIST_CHECK_OK on every rank.  tests/test_gpu_dist.py runs this under pytest when >= 2 GPUs
are visible."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def relerr(G, ref):
    return np.abs(G - ref).max() / max(np.abs(ref).max(), 1e-300)


def main():
    import ctypes as C
    import faulthandler

    # a stuck collective must show WHERE: every rank dumps its stack and exits instead of hanging the box
    faulthandler.dump_traceback_later(int(os.environ.get("DIST_CHECK_WATCHDOG_S", "240")), exit=True)

    import torch
    import torch.distributed as dist

    import grm_b200
    from grm_b200 import _lib, grm_call, synth
    from grm_b200 import device as D
    from grm_b200 import dist as GD

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = grm_b200.Context(local)
    GD.init_comm(ctx)
    r2, w2, ver = ctx.comm_info()
    assert (r2, w2) == (rank, world) and ver >= 20000, (r2, w2, ver)
    dev = torch.device("cuda", local)
    lib = ctx.lib

    def check_tiles(T, ref, n):
        ti, tj = C.c_int64(), C.c_int64()
        cnt = lib.grm_cuda_tile_count(n, rank, world)
        for t in range(cnt):
            assert lib.grm_cuda_tile_at(n, rank, world, t, C.byref(ti), C.byref(tj)) == 0
            r0, c0 = ti.value * 256, tj.value * 256
            blk = np.zeros((256, 256))
            sub = ref[r0:r0 + 256, c0:c0 + 256]
            if ti.value == tj.value:
                sub = np.triu(sub)  # diagonal tiles hold the upper triangle only
            blk[:sub.shape[0], :sub.shape[1]] = sub
            got = T[t].T  # tile is [col][row]
            if ti.value == tj.value:
                got = np.triu(got)
            yield got, blk

    cases = [
        # (kind, n, p, m/ploidy of generator, ploidy of the call)  — ksplit: p > 4n, gather: p <= 4n
        ("dosage", 600, 5000, 4, 4), ("dosage", 600, 5000, 2, None), ("dosage", 1100, 2000, 2, 2),
        ("dosage", 1100, 2000, 4, None), ("continuous", 500, 3000, 0, 2), ("continuous", 700, 900, 0, None),
        ("dosage", 300, 3, 2, 2), ("dosage", 130, 1, 2, None),  # fewer loci than ranks x 128: short and empty slabs
        # ONE output tile: with two or more ranks some of them own no tile at all (what 8 ranks see at n = 600)
        ("dosage", 200, 2000, 2, 2), ("dosage", 200, 2000, 4, None), ("continuous", 200, 1500, 0, 2), ("dosage", 256, 900, 2, 2),
    ]
    for kind, n, p, gm, ploidy in cases:
        A = synth.dosage(n, p, gm, seed=n + p) if kind == "dosage" else synth.continuous(n, p, seed=n + p)
        exact = kind == "dosage"
        ref_raw, _ = grm_call(A, ploidy, ctx=ctx, skip_repair=True)
        ref, iref = grm_call(A, ploidy, ctx=ctx)
        # (1) whole matrix on every rank, the library reads only its loci range; with the parallel-probe repair
        G, info = grm_call(A, ploidy, ctx=ctx, distributed=True)
        assert info["world"] == world and info["path"] == iref["path"], (info, iref)
        assert info["repair_iters"] == iref["repair_iters"], (kind, n, p, info["repair_iters"], iref["repair_iters"])
        if exact:
            assert np.array_equal(G, ref), (kind, n, p, ploidy, relerr(G, ref))
            assert info["eps_total"] == iref["eps_total"]
        else:
            assert relerr(G, ref) <= 1e-13, (kind, n, p, ploidy, relerr(G, ref))
        assert np.array_equal(G, G.T)
        # (2) the rank's own loci slab as input (what a host that shards its Genomes passes), raw GRM
        k0, k1 = GD.loci_range(p, rank, world)
        slab = np.asfortranarray(A[:, k0:k1]) if k1 > k0 else np.zeros((n, 1), order="F")
        G2, i2 = grm_call(slab, ploidy, ctx=ctx, distributed=True, input_slab=True, p_total=p, skip_repair=True)
        assert (np.array_equal(G2, ref_raw) if exact else relerr(G2, ref_raw) <= 1e-13), (kind, n, p, ploidy)
        # (3) device-resident slab, sharded output (tiles stay on their owners)
        slab_dev = torch.from_numpy(np.ascontiguousarray(A[:, k0:k1].T)).to(dev)
        T, i3 = D.grm_device(ctx, slab_dev, ploidy, skip_repair=True, distributed=True, p_total=p, tiles=True)
        ctx.synchronize()
        Th = T.cpu().numpy()
        for got, blk in check_tiles(Th, ref_raw, n):
            assert (np.array_equal(got, blk) if exact else relerr(got, blk) <= 1e-13 or np.abs(blk).max() == 0), (kind, n, p)
        # (4) the same from host memory
        T2, i4 = grm_call(A, ploidy, ctx=ctx, distributed=True, tiles=True, skip_repair=True)
        cnt = lib.grm_cuda_tile_count(n, rank, world)
        if cnt > 0:  # a rank may own no tile at all
            assert (np.array_equal(T2[:cnt], Th[:cnt]) if exact else relerr(T2[:cnt], Th[:cnt]) <= 1e-13), (kind, n, p)
        if p > 4 * n and exact and world > 1:
            assert i3["exchange_waves"] >= 1
        dist.barrier()

    # exchange waves: any number gives the same bits
    A = synth.dosage(1300, 6000, 4, seed=3)
    ref, _ = grm_call(A, 4, ctx=ctx, skip_repair=True)
    # ... through every exchange flow: 0 = one launch per wave + NCCL (default), 1 = streamed + NCCL, 2 = streamed + copy engines
    for exchange in (0, 1, 2):
        ctx.set_tuning("exchange", exchange)
        for waves in (1, 2, 3, 8):
            ctx.set_tuning("ksplit_waves", waves)
            for pl in (4, None):
                G, info = grm_call(A, pl, ctx=ctx, distributed=True, skip_repair=True)
                want = ref if pl == 4 else grm_call(A, None, ctx=ctx, skip_repair=True)[0]
                assert info["exchange_waves"] == min(waves, lib.grm_cuda_packed_tiles_max(1300, world)), (exchange, waves)
                assert np.array_equal(G, want), (exchange, waves, pl)
    ctx.set_tuning("ksplit_waves", 0)
    ctx.set_tuning("exchange", 0)
    # back-to-back calls reuse the peer buffers: the epoch flags must keep the steps apart
    for _ in range(5):
        G, info = grm_call(A, 4, ctx=ctx, distributed=True, skip_repair=True)
        assert np.array_equal(G, ref)

    # QC mask + mean-fill, distributed: kept loci / scale agree with the single-GPU call
    A = synth.add_missing(synth.dosage(400, 3000, 2, seed=8), 0.002, seed=9)
    for kw in (dict(maf=0.05, max_locus_sparsity=0.0), dict(maf=0.02, max_locus_sparsity=0.05, missing="meanfill")):
        ref, iref = grm_call(A, 2, ctx=ctx, skip_repair=True, **kw)
        G, info = grm_call(A, 2, ctx=ctx, distributed=True, skip_repair=True, **kw)
        assert info["loci_kept"] == iref["loci_kept"] and info["path"] == iref["path"]
        assert (np.array_equal(G, ref) if info["path"] == _lib.PATH_I8 else relerr(G, ref) <= 1e-13), kw

    # consensus: the offending cell lives in the LAST rank's loci only, every rank must take the same branch
    n, p = 2100, 2600 * max(world, 1)
    A = synth.dosage(n, p, 2, seed=13)
    A[11, p - 1] = np.nan
    try:
        grm_call(A, 2, ctx=ctx, distributed=True, skip_repair=True)
        raise AssertionError("missing cell not reported")
    except _lib.GrmCudaError as e:
        assert e.status == _lib.ERR_MISSING
    A[11, p - 1] = 0.3  # beyond every rank's detection sample: int8 m=2 -> retry m=64 -> FP64, together
    G, info = grm_call(A, None, ctx=ctx, distributed=True, skip_repair=True)
    ref, iref = grm_call(A, None, ctx=ctx, skip_repair=True)
    assert info["path"] == iref["path"] == _lib.PATH_F64 and relerr(G, ref) <= 1e-13
    A[11, p - 1] = 5.0 / 64.0
    G, info = grm_call(A, 2, ctx=ctx, distributed=True, skip_repair=True)
    ref, iref = grm_call(A, 2, ctx=ctx, skip_repair=True)
    assert info["denominator"] == iref["denominator"] == 64 and np.array_equal(G, ref)

    # a call without a communicator must fail loudly, not hang
    c2 = grm_b200.Context(local)
    try:
        grm_call(synth.dosage(10, 10, 2), None, ctx=c2, distributed=True)
        raise AssertionError("distributed call without a communicator succeeded")
    except _lib.GrmCudaError as e:
        assert e.status == _lib.ERR_COMM
    c2.close()

    dist.barrier()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()
    print(f"DIST_CHECK_OK rank {rank} of {world}", flush=True)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        # leave at once: a destructor that waits for a collective the peers are still in would keep the launcher from
        # noticing the failure (and from tearing the other ranks down)
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
