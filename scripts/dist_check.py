"""Multi-GPU correctness check (run under torch.distributed.run, one rank per GPU):
the sharded path (loci-split pack -> int8 all-gather -> tile-sharded SYRK -> gather) against the single-GPU call."""
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
ok = True
MODE = os.environ.get("GRM_DIST_MODE", "auto")
for (n, p, m, ploidy) in [(3000, 20000, 4, 4), (1100, 777, 2, None), (2049, 5001, 4, 2)]:
    A = synth.dosage_device(n, p, m, 5, dev)  # same seed on every rank => identical matrix
    k0, k1 = GD.loci_range(p, rank, world)
    sh = GD.ShardedGRM(n, p, ploidy, m, ctx, mode=MODE)
    sh.step(A[k0:k1].contiguous())
    G = sh.gather().clone()
    sh.step(A[k0:k1].contiguous())  # second step after a gather must start from zeros again
    G2 = sh.gather().clone()
    for _ in range(4):  # back-to-back steps: a rank may not overwrite a peer that still finalises the previous step
        sh.step(A[k0:k1].contiguous())
    G3 = sh.gather()
    ref, info = D.grm_device(ctx, A, ploidy, denominator=m, skip_repair=True)
    ctx.synchronize()
    same = torch.equal(G, ref)
    rel = float((G - ref).abs().max() / ref.abs().max())
    again = torch.equal(G, G2) and torch.equal(G, G3)
    sym = torch.equal(G, G.T)
    print(f"[rank {rank}/{world}] n={n} p={p} ploidy={ploidy}: bit-identical={same} rel={rel:.2e} repeat={again} symmetric={sym} mode={sh.mode} exchange={getattr(sh, 'exchange', '-')}",
          flush=True)
    ok = ok and same and again and sym  # integer statistics: bit-identical for any number of ranks
    exch = getattr(sh, "exchange", "-")
    sh.close()
t = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if float(t[0]) == 1.0 else 1)
