"""BASELINE config 5 at full size on 8 GPUs: 100,000 entries x 50,000 loci diploid, output sharded as packed FP64 tiles
(80 GB as a dense matrix; 10 GB of tiles per GPU), grmsimple.  Run under torch.distributed.run, one rank per GPU.
Checks (size-independent properties, exact): every diagonal tile is symmetric; for sampled rows, G_ii and a sampled
off-diagonal G_ij equal the value computed directly from the codes."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import grm_b200
from grm_b200 import device as D, dist as GD, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = grm_b200.Context(local)
n, p, m = int(os.environ.get("C5_N", 100_000)), int(os.environ.get("C5_P", 50_000)), 2
k0, k1 = GD.loci_range(p, rank, world)
A = synth.dosage_device(n, k1 - k0, m, 500 + rank, dev, chunk=1250)
torch.cuda.synchronize()
sh = GD.ShardedGRM(n, p, None, m, ctx, mode="gather", output="tiles")
for _ in range(2):
    sh.step(A)
dist.barrier()
torch.cuda.synchronize()
with D.on_ctx_stream(ctx):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t = {}
    for _ in range(3):
        sh.step(A, timings=t)
    e1.record()
    e1.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
tiles = sh.tiles
ok = True
ti, tj = C.c_int64(), C.c_int64()
codes = sh.codes_all  # [world, n, lds]
checked = 0
for tt in range(0, tiles.shape[0], max(1, tiles.shape[0] // 6)):
    assert ctx.lib.grm_cuda_tile_at(n, rank, world, tt, C.byref(ti), C.byref(tj)) == 0
    r0, c0 = ti.value * 256, tj.value * 256
    rows = codes[:, r0:r0 + 256].long()            # [world, 256, lds]
    cols = codes[:, c0:c0 + 256].long()
    ref = torch.einsum("sik,sjk->ij", rows.double(), cols.double())  # exact: integers far below 2^53
    want = torch.div(torch.div(ref, torch.tensor(float(m * m), dtype=torch.float64, device=dev)),
                     torch.tensor(float(p), dtype=torch.float64, device=dev))
    got = tiles[tt].T[:want.shape[0], :want.shape[1]]
    ok = ok and torch.equal(got, want)
    if ti.value == tj.value:
        ok = ok and torch.equal(tiles[tt], tiles[tt].T)
    checked += 1
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    macs = n * n * p / 2
    print(json.dumps({"config": f"C5 {n} x {p} diploid grmsimple, sharded tile output", "n_gpus": world,
                      "ms_per_step": float(ms[0]), "mac_per_s": macs / (float(ms[0]) * 1e-3),
                      "stage_ms": {k: v / 3 for k, v in t.items()}, "tiles_per_rank": int(tiles.shape[0]),
                      "tile_bytes_per_rank": int(tiles.numel() * 8), "tiles_checked_per_rank": checked,
                      "bit_exact_vs_codes": bool(float(flag[0]) == 1.0)}))
dist.destroy_process_group()
sys.exit(0 if float(flag[0]) == 1.0 else 1)
