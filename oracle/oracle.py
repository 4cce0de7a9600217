"""ctypes front-end of the CPU oracle (oracle/massiv_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, from ``__graft_entry__.smoke()`` and
from ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs — never from ``massiv_b200``.
Nothing in here reads /root/reference at run time.

The functions take plain Python values (no massiv_b200 types) so the oracle stays independent
of the product it checks.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmassiv_oracle.so")
MAX_RANK = 5

KINDS = {"id": 0, "sum": 1, "product": 2, "avg": 3, "max": 4, "min": 5, "conv": 6, "corr": 7,
         "taps": 8, "life": 9, "mag2": 10, "int_midpoint": 11, "int_trapezoid": 12, "int_simpson": 13, "expr": 14}
BORDERS = {"fill": 0, "wrap": 1, "edge": 2, "reflect": 3, "continue": 4}
ELEMS = {"u8": 0, "i8": 1, "u16": 2, "i16": 3, "u32": 4, "i32": 5, "u64": 6, "i64": 7,
         "f32": 8, "f64": 9, "bool32": 10}
_NP2ELEM = {np.dtype(np.uint8): "u8", np.dtype(np.int8): "i8", np.dtype(np.uint16): "u16",
            np.dtype(np.int16): "i16", np.dtype(np.uint32): "u32", np.dtype(np.int32): "i32",
            np.dtype(np.uint64): "u64", np.dtype(np.int64): "i64", np.dtype(np.float32): "f32",
            np.dtype(np.float64): "f64"}
ELEM2NP = {v: k for k, v in _NP2ELEM.items()}
ELEM2NP["bool32"] = np.dtype(np.int32)
FLAG_MAG2_CORR = 1
FLAG_NO_FAST = 1 << 31

OK, ERR_INVALID_DESC, ERR_SIZE_MISMATCH, ERR_INDEX_ZERO, ERR_UNSUPPORTED = range(5)


class OracleError(RuntimeError):
    def __init__(self, status: int):
        super().__init__(f"oracle status {status}")
        self.status = status


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


class _Desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("elem", C.c_int32), ("rank", C.c_int32), ("border", C.c_int32),
                ("size", C.c_int64 * MAX_RANK), ("center", C.c_int64 * MAX_RANK),
                ("pad_lo", C.c_int64 * MAX_RANK), ("pad_hi", C.c_int64 * MAX_RANK),
                ("stride", C.c_int64 * MAX_RANK), ("fill", C.c_uint8 * 8),
                ("coeffs", C.c_void_p), ("coeffs2", C.c_void_p), ("ntaps", C.c_int64),
                ("tap_offsets", C.c_void_p), ("flags", C.c_uint32), ("n_post", C.c_uint32), ("post", C.c_void_p),
                ("n_terms", C.c_uint32), ("n_program", C.c_uint32), ("terms", C.c_void_p), ("program", C.c_void_p)]


class _Term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("size", C.c_int64 * MAX_RANK), ("center", C.c_int64 * MAX_RANK),
                ("coeffs", C.c_void_p), ("ntaps", C.c_int64), ("tap_offsets", C.c_void_p)]


class _XIns(C.Structure):
    _fields_ = [("op", C.c_int32), ("arg", C.c_int32), ("operand", C.c_uint8 * 8)]


XOPS = {"term": 1, "const": 2, "add": 3, "sub": 4, "mul": 5, "div": 6, "min": 7, "max": 8, "negate": 9, "abs": 10,
        "signum": 11, "square": 12, "recip": 13, "sqrt": 14}


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc only; seconds)."""
    srcs = [os.path.join(_HERE, f) for f in
            ("massiv_oracle.c", "massiv_oracle_tmpl.inc", "massiv_oracle_fast.inc", "massiv_oracle.h")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "CC=gcc"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.mo_border_index.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int32)]
        _lib.mo_out_dims.argtypes = [C.POINTER(_Desc), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib.mo_geom.argtypes = [C.POINTER(_Desc), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib.mo_apply.argtypes = [C.POINTER(_Desc), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_int]
        _lib.mo_iterate.argtypes = [C.POINTER(_Desc), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_int]
    return _lib


def expected_center(kind: str, size: Sequence[int], flags: int = 0) -> tuple:
    """stencilCenter of the built-in constructors (TAPS carries its own)."""
    if kind == "conv" or (kind == "mag2" and not flags & FLAG_MAG2_CORR):
        return tuple(s - 1 - s // 2 for s in size)
    if kind == "corr" or kind == "mag2":
        return tuple(s // 2 for s in size)
    if kind == "life":
        return (1, 1)
    return tuple(0 for _ in size)


def geom_size(kind: str, size: Sequence[int]) -> tuple:
    return tuple(max(s, 1) for s in size) if kind == "avg" else tuple(size)


def same_padding(kind: str, size: Sequence[int], center: Optional[Sequence[int]] = None, flags: int = 0):
    """samePadding (Stencil.hs:173-179): lo = centre, hi = size - (centre + 1)."""
    c = tuple(center) if center is not None else expected_center(kind, size, flags)
    g = geom_size(kind, size)
    return c, tuple(max(0, gs - (cc + 1)) for gs, cc in zip(g, c))  # Sz clamps negatives to 0


class _Built:
    """keeps the numpy buffers referenced by the ctypes descriptor alive"""

    def __init__(self, desc, keep):
        self.desc, self.keep = desc, keep


def expr_geom(terms, rank: int):
    """(size, centre) of an expression: unionStencilCenters / unionStencilSizes over its terms
    (Stencil/Internal.hs:83-90); terms are dicts(kind=, size=[, center=, coeffs=, taps=])"""
    cs = [t.get("center") or expected_center(t["kind"], t["size"]) for t in terms]
    gs = [geom_size(t["kind"], t["size"]) for t in terms]
    mc = tuple(max(c[i] for c in cs) for i in range(rank))
    ext = tuple(max(max(g[i] - c[i] for g, c in zip(gs, cs)), 1) for i in range(rank))
    return tuple(a + b for a, b in zip(mc, ext)), mc


def make_desc(kind: str, elem: str, rank: int, *, size=None, center=None, pad_lo=None, pad_hi=None,
              border="fill", fill=0, stride=None, coeffs=None, coeffs2=None, taps=None,
              flags: int = 0, post=None, terms=None, program=None) -> _Built:
    """kind "expr": terms = [dict(kind=, size=[, center=, coeffs=, taps=])], program = [("term", i) | ("const", c) |
    (op,)] in postfix order; size / centre default to the union of the terms'"""
    d = _Desc()
    keep = []
    if kind == "expr":
        usz, uc = expr_geom(terms, rank)
        size = usz if size is None else size
        center = uc if center is None else center
    d.kind, d.elem, d.rank, d.border = KINDS[kind], ELEMS[elem], rank, BORDERS[border]
    size = tuple(int(s) for s in size)
    assert len(size) == rank
    if center is None:
        center = expected_center(kind, size, flags)
    if pad_lo is None or pad_hi is None:  # mapStencil
        pad_lo, pad_hi = same_padding(kind, size, center, flags)
    if stride is None:
        stride = (1,) * rank
    for i in range(rank):
        d.size[i], d.center[i] = size[i], int(center[i])
        d.pad_lo[i], d.pad_hi[i], d.stride[i] = int(pad_lo[i]), int(pad_hi[i]), int(stride[i])
    npdt = ELEM2NP[elem]
    fv = np.array([fill]).astype(npdt) if elem != "bool32" else np.array([1 if fill else 0], np.int32)
    raw = fv.tobytes()
    for i, b in enumerate(raw):
        d.fill[i] = b
    if coeffs is not None:
        k1 = np.ascontiguousarray(np.asarray(coeffs).astype(npdt)).reshape(-1)
        keep.append(k1)
        d.coeffs = k1.ctypes.data
    if coeffs2 is not None:
        k2 = np.ascontiguousarray(np.asarray(coeffs2).astype(npdt)).reshape(-1)
        keep.append(k2)
        d.coeffs2 = k2.ctypes.data
    if taps is not None:
        offs = np.ascontiguousarray(np.asarray(taps, dtype=np.int64).reshape(-1, rank) if len(taps) else
                                    np.zeros((0, rank), np.int64))
        keep.append(offs)
        d.ntaps = offs.shape[0]
        d.tap_offsets = offs.ctypes.data
    d.flags = flags
    if post:
        pa = post_array(post, npdt)
        keep.append(pa)
        d.n_post, d.post = len(post), C.cast(pa, C.c_void_p)
    if kind == "expr":
        ta = (_Term * len(terms))()
        for i, t in enumerate(terms):
            ta[i].kind = KINDS[t["kind"]]
            tsz = tuple(int(v) for v in t["size"])
            tc = t.get("center") or expected_center(t["kind"], tsz)
            for j in range(rank):
                ta[i].size[j], ta[i].center[j] = tsz[j], int(tc[j])
            if t.get("coeffs") is not None:
                kk = np.ascontiguousarray(np.asarray(t["coeffs"]).astype(npdt)).reshape(-1)
                keep.append(kk)
                ta[i].coeffs = kk.ctypes.data
            if t.get("taps") is not None:
                offs = np.ascontiguousarray(np.asarray(t["taps"], dtype=np.int64).reshape(-1, rank))
                keep.append(offs)
                ta[i].ntaps, ta[i].tap_offsets = offs.shape[0], offs.ctypes.data
        xa = (_XIns * len(program))()
        for i, step in enumerate(program):
            xa[i].op = XOPS[step[0]]
            if step[0] == "term":
                xa[i].arg = int(step[1])
            elif step[0] == "const":
                for j, b in enumerate(np.array([step[1]]).astype(npdt).tobytes()):
                    xa[i].operand[j] = b
        keep += [ta, xa]
        d.n_terms, d.n_program = len(terms), len(program)
        d.terms, d.program = C.cast(ta, C.c_void_p), C.cast(xa, C.c_void_p)
    return _Built(d, keep)


def _dims(a: Sequence[int]):
    arr = (C.c_int64 * MAX_RANK)()
    for i, v in enumerate(a):
        arr[i] = int(v)
    return arr


def elem_of(arr: np.ndarray, elem: Optional[str]) -> str:
    return elem if elem is not None else _NP2ELEM[arr.dtype]


def out_dims(kind: str, in_dims: Sequence[int], *, elem="i64", **kw) -> tuple:
    b = make_desc(kind, elem, len(in_dims), **kw)
    od = (C.c_int64 * MAX_RANK)()
    st = lib().mo_out_dims(C.byref(b.desc), _dims(in_dims), od)
    if st:
        raise OracleError(st)
    return tuple(od[i] for i in range(len(in_dims)))


def apply(kind: str, arr: np.ndarray, *, elem: Optional[str] = None, nthreads: int = 1, **kw) -> np.ndarray:
    """compute(WithStride) (applyStencil padding stencil arr) — mapStencil when no padding given."""
    arr = np.ascontiguousarray(arr)
    e = elem_of(arr, elem)
    b = make_desc(kind, e, arr.ndim, **kw)
    od = (C.c_int64 * MAX_RANK)()
    st = lib().mo_out_dims(C.byref(b.desc), _dims(arr.shape), od)
    if st:
        raise OracleError(st)
    out = np.empty(tuple(od[i] for i in range(arr.ndim)), dtype=arr.dtype)
    st = lib().mo_apply(C.byref(b.desc), _dims(arr.shape), arr.ctypes.data, out.ctypes.data, nthreads)
    if st:
        raise OracleError(st)
    return out


def iterate(kind: str, arr: np.ndarray, n_steps: int, *, elem: Optional[str] = None, nthreads: int = 1,
            **kw) -> np.ndarray:
    """iterateUntil (\\n _ _ -> n >= n_steps-1) (\\_ -> mapStencil border stencil)."""
    arr = np.ascontiguousarray(arr)
    e = elem_of(arr, elem)
    b = make_desc(kind, e, arr.ndim, **kw)
    out = np.empty_like(arr)
    tmp = np.empty_like(arr)
    st = lib().mo_iterate(C.byref(b.desc), _dims(arr.shape), arr.ctypes.data, out.ctypes.data,
                          tmp.ctypes.data, n_steps, nthreads)
    if st:
        raise OracleError(st)
    return out


def border_index(border: str, k: int, i: int):
    """handleBorderIndex for one dimension -> (repaired index | None when Fill yields the element)."""
    o, f = C.c_int64(), C.c_int32()
    st = lib().mo_border_index(BORDERS[border], k, i, C.byref(o), C.byref(f))
    if st:
        raise OracleError(st)
    return None if f.value else o.value
