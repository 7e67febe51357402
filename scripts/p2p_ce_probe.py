"""Copy-engine peer-to-peer bandwidth on one node (single process, all GPUs): GPU 0 -> every other GPU, `streams` streams,
chunks of `mb` MiB (cudaMemcpyPeerAsync through torch).  Also all GPUs sending at once (the all-to-all pattern of the
K-split exchange).  JSON lines."""
import json
import sys
import time

import torch

n = torch.cuda.device_count()
assert n >= 2
total_mb = 840  # what one rank sends per C3 step
for mb in (16, 64, 128):
    src = [torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{g}") for g in range(n)]
    dst = [[torch.empty(mb << 20, dtype=torch.uint8, device=f"cuda:{d}") for _ in range(n)] for d in range(n)]
    for streams in (1, 2, 4, 7, 14):
        for pattern in ("one_sender", "all_senders"):
            senders = [0] if pattern == "one_sender" else list(range(n))
            ss = {g: [torch.cuda.Stream(device=f"cuda:{g}") for _ in range(streams)] for g in senders}
            reps = max(1, total_mb // (mb * (n - 1)))
            def run():
                for g in senders:
                    k = 0
                    for _ in range(reps):
                        for d in range(n):
                            if d == g:
                                continue
                            with torch.cuda.device(g), torch.cuda.stream(ss[g][k % streams]):
                                dst[d][g].copy_(src[g], non_blocking=True)
                            k += 1
                for g in range(n):
                    torch.cuda.synchronize(g)
            run()
            t = time.perf_counter()
            run()
            dt = time.perf_counter() - t
            sent = reps * (n - 1) * mb / 1024.0  # GiB per sender
            print(json.dumps({"chunk_mib": mb, "streams": streams, "pattern": pattern, "gib_per_sender": round(sent, 3),
                              "ms": round(dt * 1e3, 2), "gbs_per_sender": round(sent * 1.0737 / dt, 1)}), flush=True)
    del src, dst
