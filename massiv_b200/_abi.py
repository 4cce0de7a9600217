"""ctypes binding of libmassiv_b200.so (include/massiv_b200.h).

The library is the product; this module only marshals.  There is no fallback of any kind: if the
shared library is missing the import of :func:`lib` raises, and every compute entry point raises
:class:`MassivB200Error` when the CUDA device is not usable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MASSIV_B200_LIB", os.path.join(HERE, "libmassiv_b200.so"))  # override: tuning variants only
MAX_RANK = 5
NCCL_ID_BYTES = 128

(OK, ERR_INVALID_DESC, ERR_SIZE_MISMATCH, ERR_INDEX_ZERO, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL,
 ERR_OOM) = range(8)

ELEMS = {"u8": 0, "i8": 1, "u16": 2, "i16": 3, "u32": 4, "i32": 5, "u64": 6, "i64": 7, "f32": 8,
         "f64": 9, "bool32": 10}
KINDS = {"id": 0, "sum": 1, "product": 2, "avg": 3, "max": 4, "min": 5, "conv": 6, "corr": 7,
         "taps": 8, "life": 9, "mag2": 10, "int_midpoint": 11, "int_trapezoid": 12, "int_simpson": 13, "expr": 14}
BORDERS = {"fill": 0, "wrap": 1, "edge": 2, "reflect": 3, "continue": 4}
FLAG_MAG2_CORR = 1
UNTIL_EQUAL, UNTIL_MAX_ABS_DIFF = 0, 1
FLAG_ASSUME_FINITE = 2

# every symbol include/massiv_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "mb_version", "mb_status_string", "mb_ctx_create", "mb_ctx_destroy", "mb_last_error",
    "mb_ctx_set_stream", "mb_sync", "mb_ctx_trim", "mb_wavefront_arrived", "mb_wavefront_region", "mb_wavefront_final", "mb_ctx_set_option", "mb_launch_count", "mb_aux_launch_count", "mb_last_kernel",
    "mb_elem_size", "mb_out_dims", "mb_border_index", "mb_buf_alloc", "mb_buf_free", "mb_upload",
    "mb_download", "mb_host_alloc", "mb_host_free", "mb_run", "mb_run_dev", "mb_run_sep2_dev",
    "mb_run_sep2", "mb_iterate_dev", "mb_iterate", "mb_nccl_unique_id", "mb_shard_init",
    "mb_shard_halo", "mb_shard_plan_make", "mb_iterate_sharded_dev", "mb_iterate_sharded", "mb_shard_transport", "mb_timer_start", "mb_timer_stop", "mb_selftest_div", "mb_compare_dev", "mb_iterate_until_dev", "mb_iterate_until",
]


POSTS = {"negate": 1, "abs": 2, "signum": 3, "square": 4, "add": 5, "sub": 6, "rsub": 7, "mul": 8, "min": 9, "max": 10,
         "div": 11, "rdiv": 12, "recip": 13, "sqrt": 14}
MAX_POST = 8


class PostOp(C.Structure):
    """mb_post_op: one fmap'ed function; the operand c lives in the element type"""
    _fields_ = [("op", C.c_int32), ("reserved", C.c_int32), ("operand", C.c_uint8 * 8)]


def post_array(post, npdt):
    """[(name,) | (name, c), ...] -> ctypes array of PostOp"""
    import numpy as _np
    arr = (PostOp * max(1, len(post)))()
    for i, item in enumerate(post):
        name, c = (item[0], item[1] if len(item) > 1 else 0)
        arr[i].op = POSTS[name]
        raw = _np.array([c]).astype(npdt).tobytes()
        for j, b in enumerate(raw):
            arr[i].operand[j] = b
    return arr


# program steps of an expression stencil (mb_xop)
XOPS = {"term": 1, "const": 2, "add": 3, "sub": 4, "mul": 5, "div": 6, "min": 7, "max": 8, "negate": 9, "abs": 10,
        "signum": 11, "square": 12, "recip": 13, "sqrt": 14}
XOPS_BINARY = ("add", "sub", "mul", "div", "min", "max")
XOPS_UNARY = ("negate", "abs", "signum", "square", "recip", "sqrt")
MAX_TERMS, MAX_PROGRAM, MAX_STACK = 8, 32, 8


class Term(C.Structure):
    """mb_term: one sub-stencil of an expression"""
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("size", C.c_int64 * MAX_RANK), ("center", C.c_int64 * MAX_RANK),
                ("coeffs", C.c_void_p), ("ntaps", C.c_int64), ("tap_offsets", C.c_void_p)]


class XIns(C.Structure):
    """mb_xins: one step of the postfix program"""
    _fields_ = [("op", C.c_int32), ("arg", C.c_int32), ("operand", C.c_uint8 * 8)]


class StencilDesc(C.Structure):
    """mb_stencil_desc"""
    _fields_ = [("kind", C.c_int32), ("elem", C.c_int32), ("rank", C.c_int32), ("border", C.c_int32),
                ("size", C.c_int64 * MAX_RANK), ("center", C.c_int64 * MAX_RANK),
                ("pad_lo", C.c_int64 * MAX_RANK), ("pad_hi", C.c_int64 * MAX_RANK),
                ("stride", C.c_int64 * MAX_RANK), ("fill", C.c_uint8 * 8),
                ("coeffs", C.c_void_p), ("coeffs2", C.c_void_p), ("ntaps", C.c_int64),
                ("tap_offsets", C.c_void_p), ("flags", C.c_uint32), ("n_post", C.c_uint32), ("post", C.c_void_p),
                ("n_terms", C.c_uint32), ("n_program", C.c_uint32), ("terms", C.c_void_p), ("program", C.c_void_p)]


MAX_HALO = 8
GHOST_FILL, GHOST_REMOTE = -1, -2


class ShardPlan(C.Structure):
    """mb_shard_plan"""
    _fields_ = [("first_plane", C.c_int64), ("local_planes", C.c_int64), ("halo_lo", C.c_int64),
                ("halo_hi", C.c_int64), ("recv_lower_from", C.c_int32), ("recv_upper_from", C.c_int32),
                ("send_first_to", C.c_int32), ("send_last_to", C.c_int32),
                ("ghost_lo_src", C.c_int64 * MAX_HALO), ("ghost_hi_src", C.c_int64 * MAX_HALO),
                ("step_halo_lo", C.c_int64), ("step_halo_hi", C.c_int64)]


class MassivB200Error(RuntimeError):
    """Raised for every non-zero mb_status.  `.status` holds the code."""

    def __init__(self, status: int, message: str):
        super().__init__(f"massiv_b200 status {status}: {message}")
        self.status = status


class IndexZeroException(MassivB200Error):
    """MB_ERR_INDEX_ZERO — massiv's IndexZeroException (Core/Index/Internal.hs:1031)."""


class SizeMismatchException(MassivB200Error):
    """MB_ERR_SIZE_MISMATCH — massiv's SizeMismatchException."""


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m massiv_b200.build` "
            "(there is no CPU or PyTorch fallback for the stencil path)")
    L = C.CDLL(LIB_PATH)
    i64p, vp, u8p = C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_uint8)
    dp = C.POINTER(StencilDesc)
    L.mb_version.restype = C.c_int
    L.mb_status_string.restype = C.c_char_p
    L.mb_status_string.argtypes = [C.c_int]
    L.mb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.mb_ctx_destroy.argtypes = [vp]
    L.mb_last_error.restype = C.c_char_p
    L.mb_last_error.argtypes = [vp]
    L.mb_ctx_set_stream.argtypes = [vp, vp]
    L.mb_sync.argtypes = [vp]
    L.mb_ctx_trim.argtypes = [vp]
    L.mb_wavefront_arrived.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.mb_wavefront_arrived.restype = C.c_int64
    L.mb_wavefront_region.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.mb_wavefront_region.restype = None
    L.mb_wavefront_final.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64]
    L.mb_wavefront_final.restype = C.c_int64
    L.mb_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.mb_launch_count.restype = C.c_int64
    L.mb_launch_count.argtypes = [vp]
    L.mb_aux_launch_count.restype = C.c_int64
    L.mb_aux_launch_count.argtypes = [vp]
    L.mb_last_kernel.restype = C.c_char_p
    L.mb_last_kernel.argtypes = [vp]
    L.mb_elem_size.argtypes = [C.c_int]
    L.mb_out_dims.argtypes = [dp, i64p, i64p]
    L.mb_border_index.argtypes = [C.c_int, C.c_int64, C.c_int64, i64p, C.POINTER(C.c_int32)]
    L.mb_buf_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.mb_buf_free.argtypes = [vp, vp]
    L.mb_upload.argtypes = [vp, vp, vp, C.c_size_t]
    L.mb_download.argtypes = [vp, vp, vp, C.c_size_t]
    L.mb_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.mb_host_free.argtypes = [vp, vp]
    L.mb_run.argtypes = [vp, dp, i64p, vp, vp]
    L.mb_run_dev.argtypes = [vp, dp, i64p, vp, vp]
    L.mb_run_sep2_dev.argtypes = [vp, dp, dp, i64p, vp, vp, vp]
    L.mb_run_sep2.argtypes = [vp, dp, dp, i64p, vp, vp]
    L.mb_iterate_dev.argtypes = [vp, dp, i64p, vp, vp, C.c_int64, C.POINTER(vp)]
    L.mb_iterate.argtypes = [vp, dp, i64p, vp, vp, C.c_int64]
    L.mb_nccl_unique_id.argtypes = [u8p]
    L.mb_shard_init.argtypes = [vp, C.c_int, C.c_int, u8p]
    L.mb_shard_halo.argtypes = [dp, i64p, i64p]
    L.mb_shard_plan_make.argtypes = [dp, i64p, C.c_int, C.c_int, C.POINTER(ShardPlan)]
    L.mb_iterate_sharded_dev.argtypes = [vp, dp, i64p, i64p, vp, vp, C.c_int64, C.POINTER(vp)]
    L.mb_iterate_sharded.argtypes = [vp, dp, i64p, i64p, vp, vp, C.c_int64]
    L.mb_shard_transport.restype = C.c_char_p
    L.mb_shard_transport.argtypes = [vp]
    L.mb_selftest_div.argtypes = [vp, C.c_int, C.c_int64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_uint64)]
    L.mb_iterate_until_dev.argtypes = [vp, dp, i64p, vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_double, i64p, C.POINTER(C.c_int32),
                                       C.POINTER(vp)]
    L.mb_iterate_until.argtypes = [vp, dp, i64p, vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_double, i64p, C.POINTER(C.c_int32)]
    L.mb_compare_dev.argtypes = [vp, vp, vp, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.mb_timer_start.argtypes = [vp]
    L.mb_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    _lib = L
    return L


def dims_array(dims):
    a = (C.c_int64 * MAX_RANK)()
    for i, v in enumerate(dims):
        a[i] = int(v)
    return a


def raise_for(status: int, ctx_handle=None):
    if status == OK:
        return
    L = lib()
    msg = L.mb_status_string(status).decode()
    if ctx_handle:
        detail = L.mb_last_error(ctx_handle).decode()
        if detail:
            msg = f"{msg}: {detail}"
    if status == ERR_INDEX_ZERO:
        raise IndexZeroException(status, msg)
    if status == ERR_SIZE_MISMATCH:
        raise SizeMismatchException(status, msg)
    raise MassivB200Error(status, msg)
