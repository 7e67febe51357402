"""Multi-GPU GRM: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

Two decompositions of the int8 path; both start from loci-split input (rank r holds, and end to end uploads over its
own PCIe link, only the FP64 columns of its loci range and packs them on its GPU) and both END with the finished
upper-triangular 256 x 256 tiles sharded over the ranks in contiguous runs of an L2-friendly order (grm_cuda_tile_*):

  "ksplit" (p > 4n: BASELINE C2, C3) — every rank contracts ITS loci over ALL tiles (tcgen05 SYRK, raw int32 tiles),
      then ONE reduce-scatter (int32 sum, exact) leaves each rank with the tiles it owns, which it finalises in FP64.
      Exchange: 2 n^2 bytes per GPU, independent of p.
  "gather" (otherwise, e.g. C5) — the packed int8 slabs are all-gathered (n p bytes in total) and each rank runs the
      SYRK over its own tiles with all loci (3-D tensor map over [slab][entry][locus]).

The integer Gram matrix and the integer centring statistics are exact under any partition, so both give bit-identical
results for any number of ranks.

Decomposition details (north_star):
  * loci are split into `world` contiguous ranges; rank r holds (and, end to end, uploads over its own PCIe link)
    only the FP64 columns of its range and packs them to int8 on its GPU — every input byte crosses PCIe once;
  * the packed int8 slabs (n*p bytes in total — 1/8 of the FP64 input) and the integer column sums are all-gathered:
    this is the one real exchange step of the path;
  * the upper-triangular 256 x 256 output tiles are dealt to the ranks in contiguous runs of an L2-friendly order
    (grm_cuda_tile_*); each rank runs the tcgen05 SYRK over its own tiles with all loci.  The integer Gram matrix
    is exact under any partition, so results do not depend on the number of ranks;
  * the finished tiles stay sharded; `gather()` sums the zero-padded local matrices onto every rank (each element is
    written by exactly one rank, so the sum is exact) when a single-device copy is requested, e.g. for the repair.

The exchange helpers take plain tensors and a process group, so they run under `gloo` on CPU tensors in the tests.
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


class ShardedGRM:
    """Per-rank driver of the sharded int8 path on device-resident loci slabs."""

    def __init__(self, n: int, p: int, ploidy: Optional[int], denominator: int, ctx: _lib.Context, group=None,
                 mode: str = "auto", output: str = "dense", exchange: str = "auto"):
        """output="dense": this rank's tiles inside a zero-initialised local n x n matrix (self.G);
        output="tiles" (gather mode): packed [tiles_owned, 256, 256] FP64 tiles (self.tiles), never a dense matrix —
        the form for GRMs larger than one device (BASELINE config 5: n = 100,000, 10 GB of tiles per GPU on 8)."""
        torch = _torch()
        import torch.distributed as dist

        self.n, self.p, self.ploidy, self.m = n, p, ploidy, denominator
        self.ctx = ctx
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.k0, self.k1 = loci_range(p, self.rank, self.world)
        self.lds = slab_stride(p, self.world)
        self.L = (p + self.world - 1) // self.world
        dev = torch.device("cuda", ctx.device)
        if mode == "auto":
            mode = "ksplit" if (p > 4 * n and self.world > 1) else "gather"
        self.mode = mode
        self.codes_slab = torch.zeros((n, self.lds), dtype=torch.int8, device=dev)
        self.colsum_slab = torch.zeros((self.lds,), dtype=torch.int32, device=dev)
        if mode == "gather":
            self.codes_all = torch.zeros((self.world, n, self.lds), dtype=torch.int8, device=dev)
            self.colsum_all = torch.zeros((self.world, self.lds), dtype=torch.int32, device=dev)
        else:
            # exchange of the K-split partial tiles: "nccl" = packed buffer + reduce-scatter + finalize;
            # "p2p" = the SYRK epilogue stores every tile into its owner's receive buffer over NVLink peer memory
            # (CUDA IPC mappings), then a one-element all-reduce as barrier and a finalize that sums the parts
            if exchange == "auto":
                exchange = os.environ.get("GRM_CUDA_EXCHANGE", "nccl")
            if exchange not in ("nccl", "p2p"):
                raise ValueError(f"exchange must be 'nccl' or 'p2p', got {exchange!r}")
            if exchange == "p2p" and not (1 < self.world <= 8):
                exchange = "nccl"
            self.exchange = exchange
            if denominator * denominator * p >= 2 ** 31:
                raise _lib.GrmCudaError(_lib.ERR_OVERFLOW, f"m^2 * p = {denominator}^2 * {p} overflows int32")
            self.tiles_max = ctx.lib.grm_cuda_packed_tiles_max(n, self.world)
            if exchange == "p2p":
                from . import device as D

                self.recv = D.PeerBuffer(ctx, self.world * self.tiles_max * 65536 * 4)
                mine = torch.tensor(list(self.recv.handle), dtype=torch.uint8, device=dev)
                allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(allh.view(-1), mine, group=group)
                self.recv.open_peers([bytes(row.tolist()) for row in allh.cpu()], self.rank)
                self._token = torch.zeros((1,), dtype=torch.int32, device=dev)
            else:
                self.packed = torch.zeros((self.world * self.tiles_max, 65536), dtype=torch.int32, device=dev)
                self.reduced = torch.zeros((self.tiles_max, 65536), dtype=torch.int32, device=dev)
        self.flags = torch.zeros((16,), dtype=torch.int32, device=dev)
        self.output = output
        if output == "tiles":
            if mode != "gather":
                raise ValueError('output="tiles" is available in "gather" mode')
            self.tiles = torch.zeros((ctx.lib.grm_cuda_tile_count(n, self.rank, self.world), 256, 256),
                                     dtype=torch.float64, device=dev)
            self.G = None
        else:
            self.G = torch.zeros((n, n), dtype=torch.float64, device=dev)
        self.launches = 0
        self._gathered = False

    def step(self, A_slab_pn, timings: Optional[dict] = None):
        """A_slab_pn: (k1-k0, n) contiguous FP64 CUDA tensor = this rank's loci, column-major n x pc.
        Leaves this rank's tiles of the raw GRM in self.G (zeros elsewhere).  Runs on the context stream."""
        torch = _torch()
        import torch.distributed as dist

        from . import device as D

        ctx, n, p, m = self.ctx, self.n, self.p, self.m
        pc = self.k1 - self.k0
        with D.on_ctx_stream(ctx):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timings is not None else None
            if ev:
                ev[0].record()
            self.colsum_slab.zero_()
            self.flags.zero_()
            if self._gathered and self.G is not None:  # other ranks' tiles were summed in by gather()
                self.G.zero_()
                self._gathered = False
            if pc > 0:
                D.pack_dosage(ctx, A_slab_pn, m, codes=self.codes_slab, colsum=self.colsum_slab, flags=self.flags)
                self.launches += 1
            if ev:
                ev[1].record()
            if self.mode == "gather":
                if self.world > 1:
                    dist.all_gather_into_tensor(self.codes_all.view(-1, self.lds), self.codes_slab, group=self.group)
                    dist.all_gather_into_tensor(self.colsum_all.view(-1), self.colsum_slab, group=self.group)
                else:
                    self.codes_all[0].copy_(self.codes_slab)
                    self.colsum_all[0].copy_(self.colsum_slab)
            if ev:
                ev[2].record()
            if self.ploidy is None:
                alpha, beta, gamma, denom, u = 1.0 / (m * m), 0.0, 0.0, float(p), None
            else:
                # integer statistics of the rank's OWN loci, then an exact integer all-reduce (r, T and the two sums
                # are additive over loci): no rank re-reads the other slabs, and any rank count gives the same bits
                if pc > 0:
                    sums, r_, T_ = D.dosage_stats(ctx, self.codes_slab, pc, self.colsum_slab, m=m)
                    self.launches += 2
                else:
                    sums = torch.zeros((2,), dtype=torch.int64, device=self.codes_slab.device)
                    r_ = torch.zeros((n,), dtype=torch.int64, device=self.codes_slab.device)
                    T_ = torch.zeros((n,), dtype=torch.int64, device=self.codes_slab.device)
                if self.world > 1:
                    packed = torch.cat([sums, r_, T_])
                    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
                    sums, r_, T_ = packed[:2].contiguous(), packed[2:2 + n].contiguous(), packed[2 + n:].contiguous()
                u, (alpha, beta, gamma, denom) = D.dosage_centre(ctx, sums, r_, T_, p, m, self.ploidy)
                self.launches += 1
            self.scale = denom
            if ev:
                ev[3].record()
            if self.mode == "gather" and self.output == "tiles":
                D.syrk_i8_f64_tiles(ctx, self.codes_all, min(self.L, p), alpha, beta, gamma, denom, u, rank=self.rank,
                                    world=self.world, out=self.tiles)
                self.launches += 1
                if ev:
                    ev[4].record()
            elif self.mode == "gather":
                D.syrk_i8_f64(ctx, self.codes_all, min(self.L, p), alpha, beta, gamma, denom, u, rank=self.rank,
                              world=self.world, out=self.G)
                self.launches += 1
                if ev:
                    ev[4].record()
            elif self.exchange == "p2p":
                if self.ploidy is None:
                    # no statistics all-reduce in this step: a one-element all-reduce keeps any rank from storing into
                    # a peer that is still finalising the previous step
                    dist.all_reduce(self._token, group=self.group)
                D.syrk_i8_scatter(ctx, self.codes_slab, max(pc, 1), self.rank, self.world, self.recv.peers)
                self.launches += 1
                if ev:
                    ev[4].record()
                    ev5, ev6 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.all_reduce(self._token, group=self.group)  # barrier: every peer's stores have landed
                if ev:
                    ev5.record()
                D.finalize_tiles_parts(ctx, self.recv.ptr, n, alpha, beta, gamma, denom, u, self.G, self.rank,
                                       self.world)
                self.launches += 1
                if ev:
                    ev6.record()
            else:
                D.syrk_i8_packed(ctx, self.codes_slab, max(pc, 1), self.world, out=self.packed)
                self.launches += 1
                if ev:
                    ev[4].record()
                    ev5, ev6 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if self.world > 1:
                    dist.reduce_scatter_tensor(self.reduced.view(-1), self.packed.view(-1), op=dist.ReduceOp.SUM,
                                               group=self.group)
                else:
                    self.reduced.copy_(self.packed[:self.tiles_max])
                if ev:
                    ev5.record()
                D.finalize_tiles(ctx, self.reduced, n, alpha, beta, gamma, denom, u, self.G, self.rank, self.world)
                self.launches += 1
                if ev:
                    ev6.record()
            if ev:
                torch.cuda.current_stream().synchronize()
                for name, i in (("ms_pack", 0), ("ms_exchange", 1), ("ms_rowdot", 2), ("ms_syrk", 3)):
                    timings[name] = timings.get(name, 0.0) + ev[i].elapsed_time(ev[i + 1])
                if self.mode == "ksplit":
                    timings["ms_exchange"] = timings.get("ms_exchange", 0.0) + ev[4].elapsed_time(ev5)
                    timings["ms_finalize"] = timings.get("ms_finalize", 0.0) + ev5.elapsed_time(ev6)
        flags = int(self.flags[0])
        if flags:
            raise _lib.GrmCudaError(_lib.ERR_MISSING if flags & 1 else _lib.ERR_NONFINITE if flags & 2
                                    else _lib.ERR_NOT_DOSAGE, f"input flags {flags} on rank {self.rank}")
        return self.G if self.output == "dense" else self.tiles

    def close(self):
        """Release the peer mappings and the receive buffer of the "p2p" exchange (collective: call on every rank)."""
        recv = getattr(self, "recv", None)
        if recv is not None:
            import torch.distributed as dist

            self.ctx.synchronize()
            if dist.is_initialized() and self.world > 1:
                dist.barrier(group=self.group)  # no peer may still be storing into a buffer that is about to go
            recv.close()
            self.recv = None

    def gather(self):
        """Single-device copy on every rank: sum of the zero-padded tile sets, then mirror."""
        from . import device as D

        with D.on_ctx_stream(self.ctx):
            if self.world > 1:
                gather_tiles_sum(self.G, self.group)
            D.scale_mirror(self.ctx, self.G, 1.0)
        self._gathered = True
        return self.G


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
