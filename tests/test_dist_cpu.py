"""CPU, world_size 2 and 3 over gloo: the host-side logic of the multi-GPU path — loci ranges, slab layout after the
all-gather, tile ownership, and the exactness of the sum-gather of zero-padded tile sets.  The compute itself (pack,
SYRK) is emulated here with exact integer NumPy arithmetic; the CUDA kernels are covered by tests/test_gpu_parity.py
(sharded tiles on one GPU) and by bench.py --gpus N on real ranks."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import grm_b200
from grm_b200 import _lib, dist as GD, synth

N, P, M = 700, 1000, 4  # 3 x 3 tiles of 256 entries; p not a multiple of the slab granule


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = synth.dosage(N, P, M, seed=77)
        codes = np.rint(A * M).astype(np.int8)
        k0, k1 = GD.loci_range(P, rank, world)
        lds = GD.slab_stride(P, world)
        slab = np.zeros((N, lds), dtype=np.int8)
        slab[:, :k1 - k0] = codes[:, k0:k1]  # what pack_dosage writes for this rank's loci
        colsum = np.zeros((lds,), dtype=np.int32)
        colsum[:k1 - k0] = codes[:, k0:k1].sum(0)
        codes_all, colsum_all = GD.exchange_codes(torch.from_numpy(slab), torch.from_numpy(colsum))
        assert tuple(codes_all.shape) == (world, N, lds)
        # every slab r holds loci_range(r), zero padded
        for r in range(world):
            a, b = GD.loci_range(P, r, world)
            assert np.array_equal(codes_all[r].numpy()[:, :b - a], codes[:, a:b])
            assert not codes_all[r].numpy()[:, b - a:].any()
            assert np.array_equal(colsum_all[r].numpy()[:b - a], codes[:, a:b].astype(np.int64).sum(0))
        # own tiles of the integer Gram matrix from the slab layout (what the SYRK kernel iterates over)
        lib = _lib.load()
        flat = codes_all.numpy().astype(np.int64)  # [world, N, lds]
        G = np.zeros((N, N), dtype=np.float64)
        ti, tj = C.c_int64(), C.c_int64()
        for t in range(lib.grm_cuda_tile_count(N, rank, world)):
            assert lib.grm_cuda_tile_at(N, rank, world, t, C.byref(ti), C.byref(tj)) == 0
            r0, c0 = ti.value * 256, tj.value * 256
            blk = sum(flat[s, r0:r0 + 256] @ flat[s, c0:c0 + 256].T for s in range(world))
            G[r0:r0 + 256, c0:c0 + 256] = blk
        Gt = torch.from_numpy(G)
        GD.gather_tiles_sum(Gt)
        full = codes.astype(np.int64) @ codes.astype(np.int64).T
        iu = np.triu_indices(N)
        assert np.array_equal(Gt.numpy()[iu], full[iu].astype(np.float64))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_loci_split_allgather_and_tile_gather(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


@pytest.mark.parametrize("p,world", [(1000, 1), (1000, 3), (500_000, 8), (7, 8), (128, 2)])
def test_loci_ranges_cover(p, world):
    ranges = [GD.loci_range(p, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == p
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c and a <= b
    lds = GD.slab_stride(p, world)
    assert lds % 128 == 0 and all(b - a <= lds for a, b in ranges)
