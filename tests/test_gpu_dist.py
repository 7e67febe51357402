"""GPU: the distributed path of libgrm_cuda (opts.distributed, csrc/run.cu + comm.cu).

On ONE GPU: a communicator of size 1 walks every collective code path (integer statistics all-reduce, partial-tile
reduce-scatter in waves, code all-gather, tile all-gather + unpack, parallel-probe repair) and must reproduce the plain
call bit for bit.  On >= 2 GPUs: scripts/dist_check.py under torchrun, one rank per GPU (run here with
`gpurun --gpus 2|4|8 -- python -m pytest tests/test_gpu_dist.py -m gpu`), and two contexts on two devices in one
process (per-device kernel attributes)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import grm_b200
from grm_b200 import _lib, grm_call, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch

    return torch.cuda.device_count()


def relerr(G, ref):
    return np.abs(G - ref).max() / np.abs(ref).max()


@pytest.fixture(scope="module")
def ctx1():
    c = grm_b200.Context(0)
    c.comm_init(_lib.comm_unique_id(), 0, 1)
    yield c
    c.comm_destroy()
    c.close()


@pytest.mark.parametrize("kind,n,p,gm,ploidy", [
    ("dosage", 600, 5000, 4, 4), ("dosage", 600, 5000, 2, None),        # loci-split + reduce-scatter waves (p > 4n)
    ("dosage", 1100, 2000, 2, 2), ("dosage", 1100, 2000, 4, None),      # code all-gather + tile-split (p <= 4n)
    ("continuous", 500, 3000, 0, 2), ("continuous", 700, 900, 0, None),  # FP64 loci-split
])
def test_world_of_one_reproduces_the_plain_call(ctx1, kind, n, p, gm, ploidy):
    import ctypes as C

    assert ctx1.comm_info()[:2] == (0, 1)
    A = synth.dosage(n, p, gm, seed=n + p) if kind == "dosage" else synth.continuous(n, p, seed=n + p)
    ref, iref = grm_call(A, ploidy, ctx=ctx1)
    G, info = grm_call(A, ploidy, ctx=ctx1, distributed=True)
    assert info["world"] == 1 and info["path"] == iref["path"]
    assert (info["repair_iters"], info["eps_total"]) == (iref["repair_iters"], iref["eps_total"])
    assert np.array_equal(G, ref) if kind == "dosage" else relerr(G, ref) <= 1e-13
    assert np.array_equal(G, G.T)
    if kind == "dosage" and p > 4 * n:
        for exchange in (0, 1, 2):  # per-wave launches + NCCL (default) / streamed + NCCL / streamed + copy engines
            ctx1.set_tuning("exchange", exchange)
            for waves in (1, 3, 8):
                ctx1.set_tuning("ksplit_waves", waves)
                G2, i2 = grm_call(A, ploidy, ctx=ctx1, distributed=True)
                assert i2["exchange_waves"] == min(waves, ctx1.lib.grm_cuda_packed_tiles_max(n, 1)) and np.array_equal(G2, ref), (exchange, waves)
        ctx1.set_tuning("ksplit_waves", 0)
        ctx1.set_tuning("exchange", 0)
    # sharded output: every upper-triangular tile, packed [col][row]
    raw, _ = grm_call(A, ploidy, ctx=ctx1, skip_repair=True)
    T, it = grm_call(A, ploidy, ctx=ctx1, distributed=True, tiles=True, skip_repair=True)
    ti, tj = C.c_int64(), C.c_int64()
    cnt = ctx1.lib.grm_cuda_tile_count(n, 0, 1)
    assert T.shape[0] == cnt
    for t in range(cnt):
        assert ctx1.lib.grm_cuda_tile_at(n, 0, 1, t, C.byref(ti), C.byref(tj)) == 0
        r0, c0 = ti.value * 256, tj.value * 256
        blk = np.zeros((256, 256))
        sub = raw[r0:r0 + 256, c0:c0 + 256]
        blk[:sub.shape[0], :sub.shape[1]] = sub
        got = T[t].T
        if ti.value == tj.value:
            got, blk = np.triu(got), np.triu(blk)
        assert np.array_equal(got, blk) if kind == "dosage" else relerr(got, blk) <= 1e-13


def test_distributed_needs_a_communicator_and_excludes_legacy_sharding(ctx1):
    A = synth.dosage(20, 50, 2, seed=1)
    c = grm_b200.Context(0)
    try:
        with pytest.raises(_lib.GrmCudaError) as e:
            grm_call(A, None, ctx=c, distributed=True)
        assert e.value.status == _lib.ERR_COMM
    finally:
        c.close()
    with pytest.raises(_lib.GrmCudaError) as e:
        grm_call(A, None, ctx=ctx1, distributed=True, rank=0, world=2)
    assert e.value.status == _lib.ERR_BAD_ARGUMENT


@pytest.mark.skipif("ngpus() < 2", reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_ranks_on_separate_gpus_match_the_single_gpu_result():
    world = ngpus()
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "dist_check.py")]
    env = dict(os.environ, DIST_CHECK_WATCHDOG_S="120")  # a stuck rank dumps its stack and exits instead of hanging the box
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=ROOT, env=env)
    tail = (r.stdout + r.stderr)[-6000:]
    assert r.returncode == 0, tail
    assert r.stdout.count("DIST_CHECK_OK") == world, tail


@pytest.mark.skipif("ngpus() < 2", reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_two_contexts_on_two_devices_in_one_process():
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: a second context on another GPU of the same
    process must raise the limit again (the flags used to be process-wide statics)."""
    A = synth.dosage(700, 3001, 4, seed=5)
    B = synth.continuous(300, 900, seed=6)
    c0, c1 = grm_b200.Context(0), grm_b200.Context(1)
    first = {}
    try:
        for c in (c0, c1, c0):
            for key, (M, pl) in enumerate(((A, 4), (A, None), (B, 2))):
                G, info = grm_call(M, pl, ctx=c)
                assert np.array_equal(G, first.setdefault(key, G))  # device 1 and the second visit of device 0
    finally:
        c0.close()
        c1.close()
