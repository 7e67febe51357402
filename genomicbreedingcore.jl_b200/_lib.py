"""ctypes binding of libgrm_cuda.so — a 1:1 image of include/grm_cuda.h.

This is the same binding a Julia `ccall` shim makes (julia/GRMCuda.jl); Python is used here because
Julia is not available in the build image.  There is no CPU fallback: if the shared library is missing, or the
machine has no sm_100 GPU, the compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgrm_cuda.so")

OK = 0
ERR_BAD_ARGUMENT = 1
ERR_MISSING = 2
ERR_NONFINITE = 3
ERR_NOT_SYMMETRIC = 4
ERR_NAN_MATRIX = 5
ERR_INF_MATRIX = 6
ERR_CUDA = 7
ERR_OOM = 8
ERR_NOT_DOSAGE = 9
ERR_OVERFLOW = 10
ERR_ALL_FILTERED = 11
ERR_NOT_POSDEF = 12
ERR_COMM = 13
DIM_LOCI, DIM_ENTRIES = 0, 1

PATH_AUTO, PATH_I8, PATH_F64 = 0, 1, 2
MISSING_ERROR, MISSING_MEANFILL = 0, 1
LOC_HOST, LOC_DEVICE = 0, 1
OUT_DENSE, OUT_TILES = 0, 1
INGEST_AUTO, INGEST_STREAM, INGEST_HOSTQ = 0, 1, 2


class Options(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("path", C.c_int32),
        ("denominator", C.c_int32),
        ("max_iter", C.c_int32),
        ("skip_repair", C.c_int32),
        ("input_location", C.c_int32),
        ("output_location", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("missing_policy", C.c_int32),
        ("chunk_loci", C.c_int64),
        ("maf", C.c_double),
        ("max_locus_sparsity", C.c_double),
        ("distributed", C.c_int32),
        ("input_slab", C.c_int32),
        ("output_mode", C.c_int32),
        ("reserved0", C.c_int32),
    ]


class Info(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("path", C.c_int32),
        ("denominator", C.c_int32),
        ("repair_iters", C.c_int32),
        ("eps_total", C.c_double),
        ("scale", C.c_double),
        ("tiles", C.c_int64),
        ("ms_h2d", C.c_float),
        ("ms_pack", C.c_float),
        ("ms_rowdot", C.c_float),
        ("ms_syrk", C.c_float),
        ("ms_finalize", C.c_float),
        ("ms_repair", C.c_float),
        ("ms_d2h", C.c_float),
        ("ms_total", C.c_float),
        ("kernel_launches", C.c_int32),
        ("loci_kept", C.c_int32),
        ("ms_exchange", C.c_float),
        ("ms_gather", C.c_float),
        ("ms_host_busy", C.c_float),
        ("ingest", C.c_int32),
        ("host_threads", C.c_int32),
        ("world", C.c_int32),
        ("exchange_waves", C.c_int32),
        ("syrk_window", C.c_int32),
        ("repair_lu", C.c_int32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_ if not name.startswith("reserved")}


_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); every symbol include/grm_cuda.h declares
SIGNATURES = {
    "grm_cuda_version": (_i32, []),
    "grm_cuda_last_error": (C.c_char_p, []),
    "grm_cuda_status_string": (C.c_char_p, [_i32]),
    "grm_cuda_options_default": (None, [C.POINTER(Options)]),
    "grm_cuda_ctx_create": (_i32, [_i32, C.POINTER(_vp)]),
    "grm_cuda_ctx_destroy": (_i32, [_vp]),
    "grm_cuda_ctx_stream": (_vp, [_vp]),
    "grm_cuda_ctx_synchronize": (_i32, [_vp]),
    "grm_cuda_ctx_set_tuning": (_i32, [_vp, C.c_char_p, _i64]),
    "grm_cuda_ctx_get_tuning": (_i32, [_vp, C.c_char_p, C.POINTER(_i64)]),
    "grm_cuda_comm_set_library": (_i32, [C.c_char_p]),
    "grm_cuda_comm_unique_id": (_i32, [_vp]),
    "grm_cuda_comm_init": (_i32, [_vp, _vp, _i32, _i32]),
    "grm_cuda_comm_destroy": (_i32, [_vp]),
    "grm_cuda_comm_info": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "grm_cuda_simple": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, C.POINTER(Options), C.POINTER(Info)]),
    "grm_cuda_ploidyaware": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _i64, C.POINTER(Options), C.POINTER(Info)]),
    "grm_cuda_simple_union": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _vp, _i64, C.POINTER(Options),
                                     C.POINTER(Info)]),
    "grm_cuda_ploidyaware_union": (_i32, [_vp, _vp, _vp, _i32, _i64, _i64, _i64, _i32, _vp, _i64, C.POINTER(Options),
                                          C.POINTER(Info)]),
    "grm_cuda_from_codes": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp, _i64, C.POINTER(Options),
                                   C.POINTER(Info)]),
    "grm_cuda_host_quantise": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i32, _vp, _i64, _i32, C.POINTER(C.c_uint32)]),
    "grm_cuda_host_detect": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _i32, C.POINTER(_i32)]),
    "grm_cuda_host_copy": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _i64, _i32, C.POINTER(C.c_uint32)]),
    "grm_cuda_host_simd": (C.c_char_p, []),
    "grm_cuda_host_tune": (_i32, [C.c_char_p, _i64]),
    "grm_cuda_host_lu_det": (_f64, [_vp, _vp, _i64]),
    "grm_cuda_host_chol_det": (_i32, [_vp, _i64, _f64, C.POINTER(_f64)]),
    "grm_cuda_inflate_diagonals": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, C.POINTER(_i32), C.POINTER(_f64)]),
    "grm_cuda_repair_probe": (_i32, [_vp, _vp, _i64, _i64, _i32, C.POINTER(_f64), C.POINTER(_i32), C.POINTER(C.c_uint32),
                                     C.POINTER(_f64)]),
    "grm_cuda_repair_apply": (_i32, [_vp, _vp, _i64, _i64, _i32, _f64, C.POINTER(_f64)]),
    "grm_cuda_cholesky": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _i64, _i32]),
    "grm_cuda_last_factor": (_i32, [_vp, _i64, _vp, _i64, _i32]),
    "grm_cuda_distances": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64,
                                  _i32]),
    "grm_cuda_detect_denominator": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, C.POINTER(_i32)]),
    "grm_cuda_pack_dosage": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _vp]),
    "grm_cuda_pack_codes": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp]),
    "grm_cuda_locus_keep": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _f64, _f64, _i32, _vp, _vp, _vp]),
    "grm_cuda_mask_codes": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "grm_cuda_locus_qc": (_i32, [_vp, _vp, _i64, _i64, _i64, _f64, _f64, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "grm_cuda_syrk_i8": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i32, _i32]),
    "grm_cuda_syrk_i8_f64": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _f64, _f64, _f64, _f64, _vp, _vp, _i64, _i32, _i32]),
    "grm_cuda_syrk_i8_f64_tiles": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _f64, _f64, _f64, _f64, _vp, _vp, _i32, _i32]),
    "grm_cuda_packed_tiles_max": (_i64, [_i64, _i32]),
    "grm_cuda_syrk_i8_packed": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp]),
    "grm_cuda_finalize_tiles": (_i32, [_vp, _vp, _i64, _f64, _f64, _f64, _f64, _vp, _vp, _i64, _i32, _i32]),
    "grm_cuda_dosage_stats": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _vp, _vp, _vp, _i32]),
    "grm_cuda_dosage_centre": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, C.POINTER(_f64)]),
    "grm_cuda_rowdot_i8": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "grm_cuda_locus_stats_i8": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp]),
    "grm_cuda_locus_stats_f64": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _i32]),
    "grm_cuda_syrk_f64": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _vp, _i64, _i32, _i32]),
    "grm_cuda_centre_f64": (_i32, [_vp, _vp, _i64, _i64, _i64, _i32, _f64, _vp, _i32, _vp, _vp, _i64]),
    "grm_cuda_scale_mirror": (_i32, [_vp, _vp, _i64, _i64, _f64]),
    "grm_cuda_mirror": (_i32, [_vp, _vp, _i64, _i64]),
    "grm_cuda_tile_edge": (_i32, []),
    "grm_cuda_tile_count": (_i64, [_i64, _i32, _i32]),
    "grm_cuda_tile_owner": (_i32, [_i64, _i64, _i64, _i32]),
    "grm_cuda_tile_at": (_i32, [_i64, _i32, _i32, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
}


class GrmCudaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libgrm_cuda status {status}: {message}")
        self.status = status
        self.message = message


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"(nvcc, sm_100a).  There is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != OK:
        msg = load().grm_cuda_last_error().decode("utf-8", "replace")
        raise GrmCudaError(status, msg)


def default_options() -> Options:
    o = Options()
    load().grm_cuda_options_default(C.byref(o))
    return o


def new_info() -> Info:
    i = Info()
    i.struct_size = C.sizeof(Info)
    return i


class Context:
    """Owns a grm_cuda_ctx (device, streams, grow-only workspaces)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _vp()
        check(self.lib.grm_cuda_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    @property
    def stream(self) -> int:
        return int(self.lib.grm_cuda_ctx_stream(self.handle) or 0)

    def synchronize(self) -> None:
        check(self.lib.grm_cuda_ctx_synchronize(self.handle))

    def set_tuning(self, key: str, value: int) -> None:
        """Per-context tuning knob (include/grm_cuda.h lists the keys); the library never reads the environment."""
        check(self.lib.grm_cuda_ctx_set_tuning(self.handle, key.encode(), int(value)))

    def get_tuning(self, key: str) -> int:
        v = _i64(0)
        check(self.lib.grm_cuda_ctx_get_tuning(self.handle, key.encode(), C.byref(v)))
        return v.value

    # ---- communicator (one process per GPU; NCCL lives inside the library) --------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        _name_nccl()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self.lib.grm_cuda_comm_init(self.handle, buf, int(rank), int(world)))

    def comm_info(self):
        r, w, v = _i32(0), _i32(1), _i32(0)
        check(self.lib.grm_cuda_comm_info(self.handle, C.byref(r), C.byref(w), C.byref(v)))
        return r.value, w.value, v.value

    def comm_destroy(self) -> None:
        check(self.lib.grm_cuda_comm_destroy(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.grm_cuda_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_nccl_named = False


def _name_nccl() -> None:
    """Point the library at the NCCL that PyTorch links against (the nvidia-nccl wheel) BEFORE it loads any: the loader
    hands out whichever libnccl.so.2 came first, and a later `import torch` must not be given the older system copy."""
    global _nccl_named
    if _nccl_named:
        return
    _nccl_named = True
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            path = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(path):
                load().grm_cuda_comm_set_library(path.encode())  # refused (harmless) if an NCCL is already loaded
                return
    except Exception:
        pass


def comm_unique_id() -> bytes:
    """128 opaque bytes created on ONE rank; hand them to the others and call Context.comm_init on every rank."""
    _name_nccl()
    buf = (C.c_uint8 * 128)()
    check(load().grm_cuda_comm_unique_id(buf))
    return bytes(buf)


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
