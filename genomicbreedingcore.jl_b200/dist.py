"""Multi-GPU GRM: one process per GPU.  The whole distributed pipeline lives INSIDE libgrm_cuda (csrc/run.cu, comm.cu:
NCCL over NVLink / NVSwitch, dlopen'ed by the library) behind the same grm_cuda_simple / grm_cuda_ploidyaware calls —
`opts.distributed = 1` on a context that carries a communicator.  This module is the thin Python caller:

  init_comm(ctx, group)   give the context its communicator (rank 0 creates the 128-byte id, torch.distributed — any
                          backend — only broadcasts it; a Julia host does the same with Distributed.jl / MPI.jl)
  ShardedGRM.step(slab)   ONE C call per rank on the rank's device-resident loci slab

What the library does per call (both decompositions start from loci-split input — rank r ingests only the columns of
its loci range, every input byte crosses PCIe once — and both give bit-identical results for any number of ranks,
because the integer Gram matrix and the integer centring statistics are exact under any partition):

  "ksplit" (p > 4n: BASELINE C2, C3) — every rank contracts ITS loci over ALL tiles (tcgen05 SYRK, raw int32 tiles) in
      waves; the reduce-scatter (int32 sum, exact) of wave w runs on a second stream under the contraction of wave
      w + 1 and leaves each rank with the tiles it owns, finalised in FP64.  Exchange: 2 n^2 bytes per GPU.
  "gather" (otherwise, e.g. C5) — the packed int8 slabs are all-gathered (n p bytes in total) and each rank runs the
      SYRK over its own tiles with all loci (3-D tensor map over [slab][entry][locus]).
  FP64 path — loci-split SYRK into a local matrix, FP64 reduce-scatter of the upper-triangle tiles, / d on the owner.

Output: the rank's packed tiles (OUT_TILES), or the all-gathered dense matrix on every rank followed by
inflatediagonals! with the rounds probed in parallel (rank r evaluates round base + r).

The helpers below that take plain tensors and a process group run under `gloo` on CPU tensors in the tests.
"""
from __future__ import annotations

import os

from typing import Optional, Tuple

from . import _lib


def _torch():
    import torch

    return torch


def loci_range(p: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous loci range of `rank`: equal slabs of ceil(p/world) loci (the last ones may be short or empty)."""
    L = (p + world - 1) // world
    return min(p, rank * L), min(p, (rank + 1) * L)


def slab_stride(p: int, world: int) -> int:
    """Bytes per packed row of one slab: ceil(p/world) rounded up to the 128-byte TMA/K-block granule."""
    L = (p + world - 1) // world
    return (L + 127) // 128 * 128


def exchange_codes(codes_slab, colsum_slab, group=None):
    """all-gather of the packed slabs and their column sums -> ([world, n, lds] int8, [world, lds] int32)."""
    torch = _torch()
    import torch.distributed as dist

    world = dist.get_world_size(group)
    codes_all = torch.empty((world,) + tuple(codes_slab.shape), dtype=codes_slab.dtype, device=codes_slab.device)
    colsum_all = torch.empty((world,) + tuple(colsum_slab.shape), dtype=colsum_slab.dtype, device=colsum_slab.device)
    # output viewed as the dim-0 concatenation of the inputs (the shape gloo insists on; NCCL only counts elements)
    dist.all_gather_into_tensor(codes_all.view((-1,) + tuple(codes_slab.shape[1:])), codes_slab.contiguous(), group=group)
    dist.all_gather_into_tensor(colsum_all.view(-1), colsum_slab.contiguous(), group=group)
    return codes_all, colsum_all


def gather_tiles_sum(G_local, group=None):
    """Every rank's local matrix is zero outside its own tiles; the sum over ranks is the full (upper) matrix."""
    import torch.distributed as dist

    dist.all_reduce(G_local, op=dist.ReduceOp.SUM, group=group)
    return G_local


def init_comm(ctx: _lib.Context, group=None) -> None:
    """Collective: create the library communicator of `ctx` over the ranks of `group`."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [_lib.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(box[0], rank, world)


class ShardedGRM:
    """Per-rank driver of the distributed path on device-resident loci slabs: `step` is ONE library call."""

    def __init__(self, n: int, p: int, ploidy: Optional[int], denominator: int, ctx: _lib.Context, group=None,
                 output: str = "tiles", path: int = _lib.PATH_I8, skip_repair: bool = True):
        """output="tiles": packed [tiles_owned, 256, 256] FP64 tiles of this rank (the GRM stays sharded — the form for
        matrices larger than one device, BASELINE config 5); output="dense": the complete matrix on every rank."""
        torch = _torch()
        self.n, self.p, self.ploidy, self.m = n, p, ploidy, denominator
        self.ctx = ctx
        self.rank, self.world, _ = ctx.comm_info()
        self.k0, self.k1 = loci_range(p, self.rank, self.world)
        self.output, self.path, self.skip_repair = output, path, skip_repair
        dev = torch.device("cuda", ctx.device)
        if output == "tiles":
            cnt = max(ctx.lib.grm_cuda_tile_count(n, self.rank, self.world), 1)
            self.out = torch.zeros((cnt, 256, 256), dtype=torch.float64, device=dev)
        elif output == "dense":
            self.out = torch.zeros((n, n), dtype=torch.float64, device=dev)
        else:
            raise ValueError(f"output must be 'tiles' or 'dense', got {output!r}")
        self.launches = 0
        self.info = None
        self.scale = None

    def step(self, A_slab_pn, timings: Optional[dict] = None):
        """A_slab_pn: (k1-k0, n) contiguous FP64 CUDA tensor = this rank's loci, column-major n x pc."""
        from . import device as D

        out, info = D.grm_device(self.ctx, A_slab_pn, self.ploidy, path=self.path, denominator=self.m,
                                 skip_repair=self.skip_repair, distributed=True, p_total=self.p,
                                 tiles=self.output == "tiles", out=self.out)
        self.info, self.scale = info, info["scale"]
        self.launches += info["kernel_launches"]
        if timings is not None:
            for k in ("ms_pack", "ms_rowdot", "ms_syrk", "ms_exchange", "ms_gather", "ms_finalize", "ms_repair"):
                timings[k] = timings.get(k, 0.0) + float(info[k])
        return out


def repair_probe(ctx, G, rounds: int):
    """(det, posdef, scan_flags, maxabs) of G after `rounds` rounds of inflation; G (device, n x n) is not modified."""
    import ctypes as C

    from . import device as _D  # noqa: F401  (installs Context.synchronize_with_torch)

    n = G.shape[0]
    det, pd, fl, mx = C.c_double(0.0), C.c_int32(-1), C.c_uint32(0), C.c_double(0.0)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_repair_probe(ctx.handle, C.c_void_p(G.data_ptr()), n, n, int(rounds), C.byref(det),
                                             C.byref(pd), C.byref(fl), C.byref(mx)))
    return det.value, pd.value, fl.value, mx.value


def inflatediagonals_parallel(ctx, G, max_iter: int = 1000, group=None, world: Optional[int] = None,
                              rank: Optional[int] = None, exchange=None):
    """inflatediagonals!(G) (src/grm/calc.jl:41-70) with the rounds evaluated in parallel: every rank holds the same G
    and probes round base + rank; the first round whose test passes is applied on every rank.  Identical outcome to
    the sequential loop (same arithmetic per probe).  Returns (iters, eps_total).  `exchange(list) -> list of lists`
    replaces the all-gather in single-process tests."""
    import ctypes as C
    import math

    torch = _torch()
    eps = 2.220446049250313e-16
    if world is None:
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = G.shape[0]
    base, iters, maxabs = 0, None, 0.0
    while iters is None:
        r = base + rank
        mine = [float("nan"), -1.0, 0.0, 0.0]
        if r <= max_iter:
            det, pd, fl, mx = repair_probe(ctx, G, r)
            mine = [det, float(pd), float(fl), mx]
        if exchange is not None:
            allr = exchange(mine)
        else:
            import torch.distributed as dist

            t = torch.tensor(mine, dtype=torch.float64, device=G.device)
            out = torch.empty((world, 4), dtype=torch.float64, device=G.device)
            dist.all_gather_into_tensor(out.view(-1), t, group=group)
            allr = out.cpu().tolist()
        if base == 0:
            det0, pd0, fl0, maxabs = allr[0]
            flags = int(fl0)
            if flags & 1:
                raise _lib.GrmCudaError(_lib.ERR_NOT_SYMMETRIC, "The input matrix is not symmetric.")  # calc.jl:43-45
            if det0 > eps and pd0 == 1.0:
                return 0, 0.0  # calc.jl:49-51
            if flags & 2:
                raise _lib.GrmCudaError(_lib.ERR_NAN_MATRIX, "The input matrix contains NaN values.")
            if flags & 4:
                raise _lib.GrmCudaError(_lib.ERR_INF_MATRIX, "The input matrix contains Inf values.")
        for k, (det, pd, _fl, _mx) in enumerate(allr):
            t_round = base + k
            need = (det < eps) or (pd != 1.0)  # calc.jl:61 (NaN det compares false, then isposdef decides)
            if (not need) or t_round >= max_iter:
                iters = t_round
                break
        base += world
    tot = C.c_double(0.0)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_repair_apply(ctx.handle, C.c_void_p(G.data_ptr()), n, n, int(iters), float(maxabs),
                                             C.byref(tot)))
    from . import device as _D  # noqa: F401  (installs Context.order_torch_after)

    ctx.order_torch_after()
    return iters, tot.value
