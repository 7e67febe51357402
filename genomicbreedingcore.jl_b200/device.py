"""Staged device-side API over torch CUDA tensors (PyTorch only provides device memory and streams here).

Layout convention: an n x p column-major matrix is a contiguous torch tensor of shape (p, n); the n x n column-major
GRM is a contiguous (n, n) tensor whose [j, i] element is G[i, j] (symmetric, so the distinction only matters for
partially filled, sharded outputs).  All work is enqueued on the context's stream; call `ctx.synchronize()` (or use
`on_ctx_stream`) before touching results from torch.
"""
from __future__ import annotations

import ctypes as C
from contextlib import contextmanager

from . import _lib


def _torch():
    import torch

    return torch


@contextmanager
def on_ctx_stream(ctx: _lib.Context):
    """Make the library stream torch's current stream (so torch ops, NCCL collectives and torch.cuda.Event timings
    are ordered with the library's kernels).  The library stream first waits for the work already enqueued on the
    caller's stream, and the caller's stream waits for the library stream on exit — no host synchronisation."""
    torch = _torch()
    dev = torch.device("cuda", ctx.device)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    outer = torch.cuda.current_stream(dev)
    if outer.cuda_stream != ext.cuda_stream:
        ext.wait_stream(outer)
    try:
        with torch.cuda.stream(ext):
            yield ext
    finally:
        # also on an exception: library work may still be queued, the caller's stream must stay ordered after it
        if outer.cuda_stream != ext.cuda_stream:
            outer.wait_stream(ext)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def padded_loci(p: int) -> int:
    return (p + 127) // 128 * 128


def detect_denominator(ctx, A_pn, max_cells: int = 0, ignore_missing: bool = False) -> int:
    p, n = A_pn.shape
    m = C.c_int32(0)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_detect_denominator(ctx.handle, _ptr(A_pn), n, p, n, max_cells,
                                                   1 if ignore_missing else 0, C.byref(m)))
    return m.value


def locus_qc(ctx, A_pn, maf: float, max_locus_sparsity: float, error_on_missing: bool = True, with_stats: bool = False):
    """QC locus mask (filterbysparsity + filterbymaf rules): returns (q, keep uint8, kept count, stats, flags)."""
    torch = _torch()
    p, n = A_pn.shape
    dev = A_pn.device
    q = torch.empty((p,), dtype=torch.float64, device=dev)
    keep = torch.zeros((p,), dtype=torch.uint8, device=dev)
    kept = torch.zeros((1,), dtype=torch.int64, device=dev)
    stats = torch.zeros((3,), dtype=torch.float64, device=dev) if with_stats else None
    flags = torch.zeros((16,), dtype=torch.int32, device=dev)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_locus_qc(ctx.handle, _ptr(A_pn), n, p, n, float(maf), float(max_locus_sparsity),
                                         1 if error_on_missing else 0, _ptr(q), _ptr(keep), _ptr(kept),
                                         _ptr(stats) if with_stats else None, 0, _ptr(flags)))
    ctx.order_torch_after()
    return q, keep, kept, stats, flags


def pack_dosage(ctx, A_pn, m: int, codes=None, colsum=None, flags=None, k0: int = 0, miss=None):
    """Quantise + transpose A (loci k0..k0+pc of the full problem) into codes[:, k0:k0+pc]; returns
    (codes[n, ldc] int8, colsum[ldc] int32, flags[16] uint32-as-int32).  `miss` (int32 [ldc], zeroed) switches to QC
    mode: NaN cells are packed as 0 and counted per locus instead of raising flag 1."""
    torch = _torch()
    pc, n = A_pn.shape
    dev = A_pn.device
    if codes is None:
        codes = torch.zeros((n, padded_loci(k0 + pc)), dtype=torch.int8, device=dev)
    ldc = codes.shape[1]
    if colsum is None:
        colsum = torch.zeros((ldc,), dtype=torch.int32, device=dev)
    if flags is None:
        flags = torch.zeros((16,), dtype=torch.int32, device=dev)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_pack_dosage(ctx.handle, _ptr(A_pn), n, pc, n, m,
                                            C.c_void_p(codes.data_ptr() + k0), ldc,
                                            C.c_void_p(colsum.data_ptr() + 4 * k0), _ptr(flags),
                                            C.c_void_p(miss.data_ptr() + 4 * k0) if miss is not None else None))
    ctx.order_torch_after()
    return codes, colsum, flags


def pack_codes(ctx, src_pn, codes=None, colsum=None, flags=None, k0: int = 0, miss=None):
    """Host-quantised codes (int8 [pc, lds], locus-major like the caller's matrix; lds % 4 == 0, entries beyond n
    ignored via `n = codes.shape[0]` when `codes` is given, else n = lds) -> K-major codes + column sums + flags."""
    torch = _torch()
    pc, lds = src_pn.shape
    dev = src_pn.device
    n = lds if codes is None else codes.shape[0]
    if codes is None:
        codes = torch.zeros((n, padded_loci(k0 + pc)), dtype=torch.int8, device=dev)
    ldc = codes.shape[1]
    if colsum is None:
        colsum = torch.zeros((ldc,), dtype=torch.int32, device=dev)
    if flags is None:
        flags = torch.zeros((16,), dtype=torch.int32, device=dev)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_pack_codes(ctx.handle, _ptr(src_pn), n, pc, lds, C.c_void_p(codes.data_ptr() + k0), ldc,
                                           C.c_void_p(colsum.data_ptr() + 4 * k0), _ptr(flags),
                                           C.c_void_p(miss.data_ptr() + 4 * k0) if miss is not None else None))
    ctx.order_torch_after()
    return codes, colsum, flags


def locus_keep(ctx, colsum, miss, n: int, p: int, m: int, maf: float, max_locus_sparsity: float,
               error_on_missing: bool = True, flags=None):
    """QC mask from the integers of a QC-mode pack -> (keep uint8 [p], kept count tensor, flags); colsum of dropped
    loci is zeroed in place."""
    torch = _torch()
    dev = colsum.device
    keep = torch.zeros((p,), dtype=torch.uint8, device=dev)
    kept = torch.zeros((1,), dtype=torch.int64, device=dev)
    if flags is None:
        flags = torch.zeros((16,), dtype=torch.int32, device=dev)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_locus_keep(ctx.handle, _ptr(colsum), _ptr(miss), n, p, m, float(maf),
                                           float(max_locus_sparsity), 1 if error_on_missing else 0, _ptr(keep), _ptr(kept),
                                           _ptr(flags)))
    ctx.order_torch_after()
    return keep, kept, flags


def mask_codes(ctx, codes, p: int, keep):
    n, ldc = codes.shape
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_mask_codes(ctx.handle, _ptr(codes), n, p, ldc, _ptr(keep)))
    ctx.order_torch_after()
    return codes


def _slabs(codes):
    if codes.dim() == 3:
        return codes.shape[0], codes.shape[1], codes.shape[2]
    return 1, codes.shape[0], codes.shape[1]


def syrk_i8(ctx, codes, p: int, rank: int = 0, world: int = 1, out=None):
    """codes: [n, ldc] or [nslabs, n, ldc] int8; p = valid loci per slab."""
    torch = _torch()
    nslabs, n, ldc = _slabs(codes)
    if out is None:
        out = torch.zeros((n, n), dtype=torch.int32, device=codes.device)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_syrk_i8(ctx.handle, _ptr(codes), n, p, ldc, nslabs, _ptr(out), n, rank, world))
    ctx.order_torch_after()
    return out


def syrk_i8_f64(ctx, codes, p: int, alpha: float, beta: float, gamma: float, denom: float, u=None, rank: int = 0,
                world: int = 1, out=None):
    torch = _torch()
    nslabs, n, ldc = _slabs(codes)
    if out is None:
        out = torch.zeros((n, n), dtype=torch.float64, device=codes.device)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_syrk_i8_f64(ctx.handle, _ptr(codes), n, p, ldc, nslabs, alpha, beta, gamma, denom,
                                            _ptr(u) if u is not None else None, _ptr(out), n, rank, world))
    ctx.order_torch_after()
    return out


def syrk_i8_f64_tiles(ctx, codes, p: int, alpha: float, beta: float, gamma: float, denom: float, u=None, rank: int = 0,
                      world: int = 1, out=None):
    """Sharded-output SYRK: returns [tiles_owned, 256 (col), 256 (row)] FP64 tiles of this rank."""
    torch = _torch()
    nslabs, n, ldc = _slabs(codes)
    count = ctx.lib.grm_cuda_tile_count(n, rank, world)
    if out is None:
        out = torch.empty((count, 256, 256), dtype=torch.float64, device=codes.device)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_syrk_i8_f64_tiles(ctx.handle, _ptr(codes), n, p, ldc, nslabs, alpha, beta, gamma, denom,
                                                  _ptr(u) if u is not None else None, _ptr(out), rank, world))
    ctx.order_torch_after()
    return out


def syrk_i8_packed(ctx, codes, p: int, world: int, out=None):
    """Raw int32 tiles of ALL upper-triangular tiles over this GPU's loci, in reduce-scatter layout
    [world * tiles_max, 65536]."""
    torch = _torch()
    n, ldc = codes.shape
    tmax = ctx.lib.grm_cuda_packed_tiles_max(n, world)
    if out is None:
        out = torch.zeros((world * tmax, 65536), dtype=torch.int32, device=codes.device)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_syrk_i8_packed(ctx.handle, _ptr(codes), n, p, ldc, world, _ptr(out)))
    ctx.order_torch_after()
    return out


def finalize_tiles(ctx, reduced, n: int, alpha: float, beta: float, gamma: float, denom: float, u, G, rank: int,
                   world: int):
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_finalize_tiles(ctx.handle, _ptr(reduced), n, alpha, beta, gamma, denom,
                                               _ptr(u) if u is not None else None, _ptr(G), n, rank, world))
    ctx.order_torch_after()
    return G


def locus_stats_i8(ctx, colsum, p: int, n: int, m: int, ploidy: int):
    torch = _torch()
    dev = colsum.device
    q = torch.empty((p,), dtype=torch.float64, device=dev)
    w = torch.empty((p,), dtype=torch.float64, device=dev)
    stats = torch.zeros((3,), dtype=torch.float64, device=dev)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_locus_stats_i8(ctx.handle, _ptr(colsum), p, n, m, ploidy, _ptr(q), _ptr(w), _ptr(stats)))
    ctx.order_torch_after()
    return q, w, stats


def dosage_stats(ctx, codes, p: int, colsum, sums=None, r=None, T=None, accumulate: bool = False, m: int = 64):
    """Exact integer r_i, T_i and {sum S, sum S^2} of one slab of codes (see grm_cuda_dosage_stats)."""
    torch = _torch()
    n, ldc = codes.shape
    dev = codes.device
    if sums is None:
        sums = torch.zeros((2,), dtype=torch.int64, device=dev)
        r = torch.zeros((n,), dtype=torch.int64, device=dev)
        T = torch.zeros((n,), dtype=torch.int64, device=dev)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_dosage_stats(ctx.handle, _ptr(codes), n, p, ldc, _ptr(colsum), int(m), _ptr(sums),
                                             _ptr(r), _ptr(T), 1 if accumulate else 0))
    ctx.order_torch_after()
    return sums, r, T


def dosage_centre(ctx, sums, r, T, p_total: int, m: int, ploidy: int):
    """-> (u [n] float64 CUDA tensor, (alpha, beta, gamma, denom)) for syrk_i8_f64."""
    torch = _torch()
    n = r.shape[0]
    u = torch.empty((n,), dtype=torch.float64, device=r.device)
    sc = (C.c_double * 4)()
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_dosage_centre(ctx.handle, _ptr(sums), _ptr(r), _ptr(T), n, p_total, m, ploidy, _ptr(u),
                                              sc))
    ctx.order_torch_after()
    return u, tuple(sc)


def rowdot_i8(ctx, codes, p: int, w):
    torch = _torch()
    n, ldc = codes.shape
    u = torch.empty((n,), dtype=torch.float64, device=codes.device)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_rowdot_i8(ctx.handle, _ptr(codes), n, p, ldc, _ptr(w), _ptr(u)))
    ctx.order_torch_after()
    return u


def locus_stats_f64(ctx, A_pn, accumulate: bool = False, q=None, stats=None, flags=None, skip_missing: bool = False):
    torch = _torch()
    p, n = A_pn.shape
    dev = A_pn.device
    if q is None:
        q = torch.empty((p,), dtype=torch.float64, device=dev)
    if stats is None:
        stats = torch.zeros((3,), dtype=torch.float64, device=dev)
    if flags is None:
        flags = torch.zeros((16,), dtype=torch.int32, device=dev)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_locus_stats_f64(ctx.handle, _ptr(A_pn), n, p, n, _ptr(q), _ptr(stats), _ptr(flags),
                                                1 if accumulate else 0, 1 if skip_missing else 0))
    ctx.order_torch_after()
    return q, stats, flags


def centre_f64(ctx, A_pn, centred: bool, ploidy: float = 2.0, q=None, fill_missing: bool = False, keep=None, out=None):
    """SYRK operand Z of a loci chunk: centred per calc.jl:198,205, optionally mean-filled and masked."""
    torch = _torch()
    p, n = A_pn.shape
    if out is None:
        out = torch.empty_like(A_pn)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_centre_f64(ctx.handle, _ptr(A_pn), n, p, n, 1 if centred else 0, float(ploidy),
                                           _ptr(q) if q is not None else None, 1 if fill_missing else 0,
                                           _ptr(keep) if keep is not None else None, _ptr(out), n))
    ctx.order_torch_after()
    return out


def syrk_f64(ctx, A_pn, centred: bool = False, ploidy: float = 2.0, q=None, accumulate: bool = False, rank: int = 0,
             world: int = 1, out=None, fill_missing: bool = False, keep=None):
    """G (+)= Z Z' over the owned upper-triangular tiles; Z = A, or centre_f64(A) when centring / filling / masking."""
    torch = _torch()
    p, n = A_pn.shape
    if out is None:
        out = torch.zeros((n, n), dtype=torch.float64, device=A_pn.device)
    Z = A_pn
    if centred or fill_missing or keep is not None:
        Z = centre_f64(ctx, A_pn, centred, ploidy, q, fill_missing, keep)
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_syrk_f64(ctx.handle, _ptr(Z), n, p, n, 1 if accumulate else 0, _ptr(out), n, rank, world))
    ctx.order_torch_after()
    return out


def scale_mirror(ctx, G, denom: float = 1.0):
    n = G.shape[0]
    ctx.synchronize_with_torch()
    _lib.check(ctx.lib.grm_cuda_scale_mirror(ctx.handle, _ptr(G), n, n, float(denom)))
    ctx.order_torch_after()
    return G


def grm_device(ctx, A_pn, ploidy=None, *, path=_lib.PATH_AUTO, denominator=0, skip_repair=False, max_iter=1000,
               rank=0, world=1, out=None, distributed=False, p_total=None, tiles=False, missing="error", maf=0.0,
               max_locus_sparsity=1.0):
    """One-call device-resident GRM (inputs and output already in HBM): grmsimple if `ploidy` is None, else
    grmploidyaware.  Returns (G, info dict).  distributed=True: the context carries a communicator, A_pn is THIS
    rank's loci slab ((pc, n), may be empty) of p_total loci and every rank makes the same call; tiles=True returns
    the rank's packed [tiles_owned, 256, 256] tiles instead of the dense matrix."""
    torch = _torch()
    p, n = A_pn.shape
    o = _lib.default_options()
    o.path, o.denominator, o.max_iter = path, denominator, max_iter
    o.skip_repair = 1 if skip_repair else 0
    o.input_location = o.output_location = _lib.LOC_DEVICE
    o.rank, o.world = rank, world
    o.missing_policy = {"error": _lib.MISSING_ERROR, "meanfill": _lib.MISSING_MEANFILL}[missing]
    o.maf, o.max_locus_sparsity = float(maf), float(max_locus_sparsity)
    trank, tworld = rank, world
    if distributed:
        o.distributed, o.input_slab = 1, 1
        p = int(p_total)
        trank, tworld, _ = ctx.comm_info()
    if tiles:
        o.output_mode = _lib.OUT_TILES
    if out is None:
        if tiles:
            out = torch.empty((max(ctx.lib.grm_cuda_tile_count(n, trank, tworld), 1), 256, 256), dtype=torch.float64,
                              device=A_pn.device)
        else:
            out = torch.empty((n, n), dtype=torch.float64, device=A_pn.device)
            if world > 1:
                out.zero_()
    info = _lib.new_info()
    ctx.synchronize_with_torch()
    a_ptr = _ptr(A_pn) if A_pn.numel() > 0 else C.c_void_p(out.data_ptr())  # an empty slab is never dereferenced
    if ploidy is None:
        st = ctx.lib.grm_cuda_simple(ctx.handle, a_ptr, n, p, n, _ptr(out), n, C.byref(o), C.byref(info))
    else:
        st = ctx.lib.grm_cuda_ploidyaware(ctx.handle, a_ptr, n, p, n, int(ploidy), _ptr(out), n, C.byref(o),
                                          C.byref(info))
    _lib.check(st)
    ctx.order_torch_after()
    return out, info.as_dict()


def _order_torch_after(self):
    """Make torch's current stream wait (on the device, no host sync) for what the library has enqueued so far, so
    that torch ops on the results are ordered after the kernels that produce them."""
    torch = _torch()
    dev = torch.device("cuda", self.device)
    cur = torch.cuda.current_stream(dev)
    if cur.cuda_stream != self.stream:
        ext = torch.cuda.ExternalStream(self.stream, device=dev)
        ev = torch.cuda.Event()
        ev.record(ext)
        cur.wait_event(ev)


def _synchronize_with_torch(self):
    """Order the library stream after everything torch has enqueued on its current stream for this device."""
    torch = _torch()
    cur = torch.cuda.current_stream(torch.device("cuda", self.device))
    if cur.cuda_stream != self.stream:
        cur.synchronize()


_lib.Context.synchronize_with_torch = _synchronize_with_torch
_lib.Context.order_torch_after = _order_torch_after
