"""Host-side mirror of ``Data.Massiv.Array.Stencil`` for the B200 device path.

Same names, argument order and error behaviour as the reference module
(/root/reference/massiv/src/Data/Massiv/Array/Stencil.hs:14-48), so the parity tests read like
massiv's own.  The difference is what a Stencil *is*: massiv's is an opaque closure
(Stencil/Internal.hs:30-34) which cannot run on a device, so here every constructor returns a
reified descriptor that lowers to ``mb_stencil_desc`` (include/massiv_b200.h).  Arbitrary closure
stencils are not representable — by construction, never by silent CPU fallback.

    arr' = compute(mapStencil(Edge, avgStencil(Sz(3, 3)), arr))          # host -> host (numpy)
    dev  = Device(0); d = dev.fromHost(arr)
    d'   = compute(mapStencil(Fill(0), heat, d))                          # device resident
    d''  = iterateStencil(500, Fill(0), heat, d)                          # iterateUntil counter form

Arrays are numpy arrays (the analogue of Storable ``S`` arrays: dense row-major host memory) or
:class:`DeviceArray` handles.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple, Union

import numpy as np

from . import _abi

Ix = Tuple[int, ...]

_NP2ELEM = {np.dtype(np.uint8): "u8", np.dtype(np.int8): "i8", np.dtype(np.uint16): "u16",
            np.dtype(np.int16): "i16", np.dtype(np.uint32): "u32", np.dtype(np.int32): "i32",
            np.dtype(np.uint64): "u64", np.dtype(np.int64): "i64", np.dtype(np.float32): "f32",
            np.dtype(np.float64): "f64"}
_ELEM2NP = {v: k for k, v in _NP2ELEM.items()}
_ELEM2NP["bool32"] = np.dtype(np.int32)


def Sz(*dims) -> Ix:
    """``Sz`` smart constructor: negatives clamp to 0 (Core/Index/Internal.hs:122-125).
    ``Sz(3)`` is a scalar size that broadcasts to the rank it is used at, like ``Sz 3 :: Sz2``."""
    if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
        dims = tuple(dims[0])
    return tuple(max(0, int(d)) for d in dims)


def _bcast(v, rank: int) -> Ix:
    if isinstance(v, (int, np.integer)):
        return (int(v),) * rank
    v = tuple(int(x) for x in v)
    if len(v) == 1 and rank > 1:
        return v * rank
    if len(v) != rank:
        raise ValueError(f"index {v} does not have rank {rank}")
    return v


# ---- Border (Core/Index.hs:159-195) ------------------------------------------------------------

@dataclass(frozen=True)
class Border:
    kind: str
    value: object = None

    def __repr__(self):
        return f"Fill {self.value}" if self.kind == "fill" else self.kind.capitalize()


def Fill(e) -> Border:
    return Border("fill", e)


Wrap, Edge, Reflect, Continue = Border("wrap"), Border("edge"), Border("reflect"), Border("continue")


def handleBorderIndex(border: Border, sz: Sequence[int], getVal, ix: Sequence[int]):
    """Core/Index.hs:225-253 (host-side; uses the library's own index arithmetic)."""
    L = _abi.lib()
    out = []
    for k, i in zip(sz, ix):
        r, f = C.c_int64(), C.c_int32()
        _abi.raise_for(L.mb_border_index(_abi.BORDERS[border.kind], int(k), int(i), C.byref(r), C.byref(f)))
        if f.value:
            return border.value
        out.append(r.value)
    return getVal(tuple(out))


# ---- Stencil (Stencil/Internal.hs:30-34, reified) -----------------------------------------------

@dataclass(frozen=True)
class Stencil:
    kind: str
    size: Optional[Ix]              # None: scalar size, broadcast to the array's rank on use
    center: Optional[Ix] = None     # None: implied by the constructor
    coeffs: Optional[np.ndarray] = None
    coeffs2: Optional[np.ndarray] = None
    taps: Optional[Tuple[Ix, ...]] = None      # TAPS: effective source offsets, application order
    flags: int = 0
    scalar_size: Optional[int] = None
    along: Optional[Tuple[int, int]] = None    # (Dim, length): size 1 everywhere but along Dim (integral rules)
    post: Tuple[tuple, ...] = ()               # fmap'ed functions applied to the result, in order: (name[, c])
    expr: Optional["_X"] = None                # kind "expr": the expression tree over base stencils (MB_EXPR)

    def _resolved(self, rank: int) -> "Stencil":
        if self.kind == "expr":
            ex = self.expr._resolved(rank)
            return Stencil("expr", ex.size(), ex.center(), flags=self.flags, post=self.post, expr=ex)
        if self.along is not None and self.size is None:
            dim, length = self.along
            if not 1 <= dim <= rank:   # setDim' throws IndexDimensionException
                raise IndexError(f"IndexDimensionException: (Dim {dim}) for an index of rank {rank}")
            size = tuple(length if i == rank - dim else 1 for i in range(rank))   # Dim 1 is the innermost
            return Stencil(self.kind, size, None, self.coeffs, self.coeffs2, self.taps, self.flags, post=self.post)
        if self.size is not None:
            if len(self.size) != rank:
                raise ValueError(f"stencil of rank {len(self.size)} applied to an array of rank {rank}")
            return self
        return Stencil(self.kind, (self.scalar_size,) * rank, None, self.coeffs, self.coeffs2, self.taps, self.flags, post=self.post)

    def assumeFinite(self) -> "Stencil":
        """Device-path extension: promise finite inputs so zero coefficients may be skipped
        (bit-identical for finite data; see MB_FLAG_ASSUME_FINITE)."""
        return Stencil(self.kind, self.size, self.center, self.coeffs, self.coeffs2, self.taps,
                       self.flags | _abi.FLAG_ASSUME_FINITE, self.scalar_size, self.along, self.post)

    # ---- Functor / Num / Fractional / Floating (Stencil/Internal.hs:36-41, 114-146) --------------
    # `fmap f` for the f whose result is exactly specified (include/massiv_b200.h: mb_post); anything
    # else is a closure and stays on massiv's CPU path.
    def fmap(self, name: str, c=None) -> "Stencil":
        """``fmap f stencil`` with f named: "negate", "abs", "signum", "square" (^2), "recip", "sqrt", or with a
        constant c: "add" (+ c), "sub" (subtract c), "rsub" (c -), "mul" (* c), "div" (/ c), "rdiv" (c /), "min", "max"."""
        if name not in _abi.POSTS:
            raise NotImplementedError(f"fmap {name}: not an exactly specified post-map; closures cannot run on the device")
        if self.kind == "expr":
            return Stencil("expr", None, None, flags=self.flags, expr=_node_fmap(self.expr, name, c))
        if self.kind == "life" or len(self.post) >= _abi.MAX_POST:
            raise NotImplementedError("too many post-maps, or a rule stencil")
        item = (name,) if c is None else (name, c)
        return Stencil(self.kind, self.size, self.center, self.coeffs, self.coeffs2, self.taps, self.flags,
                       self.scalar_size, self.along, self.post + (item,))

    def __neg__(self):
        return self.fmap("negate")

    def __abs__(self):
        return self.fmap("abs")

    def signum(self):
        return self.fmap("signum")

    def recip(self):
        return self.fmap("recip")

    def sqrt(self):
        return self.fmap("sqrt")

    # stencil `op` constant  ==  liftA2 op stencil (pure c): same size and centre, lowered to a post-map;
    # stencil `op` stencil   ==  liftA2 op s1 s2 (Stencil/Internal.hs:93-131): an expression stencil (MB_EXPR)
    def __add__(self, c):
        return self.fmap("add", c) if np.isscalar(c) else _lift2("add", self, c)

    def __radd__(self, c):  # IEEE / modular addition commutes
        return self.fmap("add", c) if np.isscalar(c) else _lift2("add", c, self)

    def __sub__(self, c):
        return self.fmap("sub", c) if np.isscalar(c) else _lift2("sub", self, c)

    def __rsub__(self, c):
        return self.fmap("rsub", c) if np.isscalar(c) else _lift2("sub", c, self)

    def __mul__(self, c):
        return self.fmap("mul", c) if np.isscalar(c) else _lift2("mul", self, c)

    def __rmul__(self, c):
        return self.fmap("mul", c) if np.isscalar(c) else _lift2("mul", c, self)

    def __truediv__(self, c):
        return self.fmap("div", c) if np.isscalar(c) else _lift2("div", self, c)

    def __rtruediv__(self, c):
        return self.fmap("rdiv", c) if np.isscalar(c) else _lift2("div", c, self)

    # Num / Floating instances (Stencil/Internal.hs:114-174) — only the Sobel-magnitude shape
    # `sqrt (fmap (^2) a + fmap (^2) b)` lowers to a device kernel (MB_MAG2).
    def __pow__(self, n):
        if n != 2:
            raise NotImplementedError("only `fmap (^2)` lowers to the device")
        return self.fmap("square")


@dataclass(frozen=True)
class _X:
    """A node of a stencil expression: the Applicative / Num / Fractional / Floating instances of Stencil
    (Stencil/Internal.hs:93-174) reified.  op is "term" (a built-in stencil without post-maps), "const" (`pure c`),
    one of _abi.XOPS_BINARY (liftA2) or _abi.XOPS_UNARY (fmap)."""
    op: str
    args: Tuple["_X", ...] = ()
    term: Optional[Stencil] = None
    value: object = None

    def _resolved(self, rank: int) -> "_X":
        if self.op == "term":
            return _X("term", term=self.term._resolved(rank))
        return _X(self.op, tuple(a._resolved(rank) for a in self.args), value=self.value)

    # unionStencilCenters / unionStencilSizes (Stencil/Internal.hs:83-90); `pure` is size 1, centre 0
    def center(self) -> Optional[Ix]:
        if self.op == "term":
            return _center(self.term)
        if self.op == "const":
            return None
        cs = [c for c in (a.center() for a in self.args) if c is not None]
        return tuple(max(v) for v in zip(*cs)) if cs else None

    def _extent(self) -> Optional[Ix]:   # size - centre
        if self.op == "term":
            return tuple(g - c for g, c in zip(_geom_size(self.term), _center(self.term)))
        if self.op == "const":
            return None
        es = [e for e in (a._extent() for a in self.args) if e is not None]
        return tuple(max(max(v), 1) for v in zip(*es)) if es else None

    def size(self) -> Optional[Ix]:
        c, e = self.center(), self._extent()
        return None if c is None else tuple(a + b for a, b in zip(c, e))


def _term_node(st: Stencil) -> _X:
    """a stencil as an expression node: its post-maps become fmap / liftA2-with-pure nodes"""
    if st.kind == "expr":
        return st.expr
    if st.kind in ("life", "mag2") or st.kind.startswith("int_"):
        raise NotImplementedError(f"{st.kind} stencils do not combine into expressions")
    node = _X("term", term=Stencil(st.kind, st.size, st.center, st.coeffs, None, st.taps, 0, st.scalar_size, st.along))
    for item in st.post:
        node = _node_fmap(node, item[0], item[1] if len(item) > 1 else None)
    return node


def _node_fmap(node: _X, name: str, c=None) -> _X:
    if name in _abi.XOPS_UNARY:
        return _X(name, (node,))
    if name in _abi.XOPS_BINARY:            # (`op` c)
        return _X(name, (node, _X("const", value=c)))
    if name == "rsub":                      # (c -)
        return _X("sub", (_X("const", value=c), node))
    if name == "rdiv":                      # (c /)
        return _X("div", (_X("const", value=c), node))
    raise NotImplementedError(f"fmap {name}")


def _as_node(x) -> _X:
    if isinstance(x, Stencil):
        return _term_node(x)
    if isinstance(x, _X):
        return x
    raise TypeError(f"cannot combine a stencil with {type(x).__name__}")


def _lift2(op: str, a, b) -> Stencil:
    """``liftA2 op a b`` of two stencils (Stencil/Internal.hs:104-111): one pass, an MB_EXPR descriptor"""
    if not isinstance(a, Stencil) or not isinstance(b, Stencil):
        return NotImplemented
    flags = (a.flags | b.flags) & _abi.FLAG_ASSUME_FINITE
    if op == "mul" and a is b:   # s * s evaluates s twice in the reference; x * x is (^2) of one evaluation, bit for bit
        return Stencil("expr", None, None, flags=flags, expr=_X("square", (_as_node(a),)))
    return Stencil("expr", None, None, flags=flags, expr=_X(op, (_as_node(a), _as_node(b))))


def minStencils(a: Stencil, b: Stencil) -> Stencil:
    """``liftA2 min a b``"""
    return _lift2("min", a, b)


def maxStencils(a: Stencil, b: Stencil) -> Stencil:
    """``liftA2 max a b``"""
    return _lift2("max", a, b)


def _squared_kernel(node: _X) -> Optional[Stencil]:
    """the FromKernel stencil k when node is ``fmap (^2) k`` or ``k * k``"""
    if node.op == "square" and node.args[0].op == "term":
        t = node.args[0].term
    elif node.op == "mul" and node.args[0].op == "term" and node.args[1].op == "term" and node.args[0].term is node.args[1].term:
        t = node.args[0].term
    else:
        return None
    return t if t.kind in ("conv", "corr") and t.size is not None else None


def sqrt(x: Stencil) -> Stencil:
    """``sqrt`` of the Floating (Stencil ix e) instance: ``fmap sqrt``.  The Sobel-magnitude shape
    ``sqrt (fmap (^2) a + fmap (^2) b)`` of two FromKernel stencils of one size takes the dedicated MB_MAG2 kernels;
    every other expression is evaluated as MB_EXPR."""
    if not isinstance(x, Stencil):
        raise NotImplementedError("sqrt of this object does not lower to the device")
    if x.kind == "expr" and x.expr.op == "add":
        a, b = _squared_kernel(x.expr.args[0]), _squared_kernel(x.expr.args[1])
        if a is not None and b is not None and a.kind == b.kind and a.size == b.size:
            flags = x.flags | (_abi.FLAG_MAG2_CORR if a.kind == "corr" else 0)
            return Stencil("mag2", a.size, None, a.coeffs, b.coeffs, None, flags)
    return x.fmap("sqrt")


def _flatten(node: _X, rank: int):
    """expression tree -> (distinct term stencils, postfix program [(op, arg)]) within the descriptor's limits"""
    terms, prog = [], []

    def key(t: Stencil):
        return (t.kind, t.size, t.center, None if t.coeffs is None else np.asarray(t.coeffs).tobytes(),
                None if t.coeffs is None else str(np.asarray(t.coeffs).dtype), t.taps)

    keys = []

    def walk(n: _X) -> int:   # returns the stack depth the subtree needs
        if n.op == "term":
            k = key(n.term)
            if k not in keys:
                keys.append(k)
                terms.append(n.term)
            prog.append(("term", keys.index(k)))
            return 1
        if n.op == "const":
            prog.append(("const", n.value))
            return 1
        depths = [walk(a) for a in n.args]
        prog.append((n.op, None))
        return max(d + i for i, d in enumerate(depths))

    depth = walk(node)
    if len(terms) > _abi.MAX_TERMS or len(prog) > _abi.MAX_PROGRAM or depth > _abi.MAX_STACK:
        raise NotImplementedError(f"stencil expression too large for one descriptor: {len(terms)} terms, {len(prog)} steps, "
                                  f"stack {depth} (limits {_abi.MAX_TERMS}, {_abi.MAX_PROGRAM}, {_abi.MAX_STACK})")
    return terms, prog


def getStencilSize(st: Stencil) -> Ix:
    return _geom_size(st)


def getStencilCenter(st: Stencil) -> Ix:
    return _center(st)


def _geom_size(st: Stencil) -> Ix:
    if st.kind == "expr":
        return st.size if st.size is not None else st.expr.size()
    if st.kind == "avg":   # liftA2 (/) with `pure n`: unionStencilSizes (Stencil/Internal.hs:83-90)
        return tuple(max(s, 1) for s in st.size)
    return st.size


def _center(st: Stencil) -> Ix:
    if st.center is not None:
        return st.center
    if st.kind == "expr":
        return st.expr.center()
    if st.kind == "conv" or (st.kind == "mag2" and not st.flags & _abi.FLAG_MAG2_CORR):
        return tuple(s - 1 - s // 2 for s in st.size)    # Convolution.hs:71-73
    if st.kind in ("corr", "mag2"):
        return tuple(s // 2 for s in st.size)            # Convolution.hs:114
    if st.kind == "life":
        return (1, 1)
    return tuple(0 for _ in st.size)                     # foldlStencil: centre 0 (Stencil.hs:298-302)


def _fold(kind: str, sz) -> Stencil:
    if isinstance(sz, (int, np.integer)):
        return Stencil(kind, None, scalar_size=max(0, int(sz)))
    sz = Sz(sz)
    if len(sz) == 1:  # `Sz 3` is polymorphic in the rank (Num instance of Sz): resolved on use
        return Stencil(kind, None, scalar_size=sz[0])
    return Stencil(kind, sz)


def idStencil(rank: int = None) -> Stencil:
    """Stencil.hs:269"""
    return Stencil("id", None, scalar_size=1) if rank is None else Stencil("id", (1,) * rank)


def sumStencil(sz) -> Stencil:
    """Stencil.hs:425-427"""
    return _fold("sum", sz)


def productStencil(sz) -> Stencil:
    """Stencil.hs:454-456"""
    return _fold("product", sz)


def avgStencil(sz) -> Stencil:
    """Stencil.hs:479-481: ``sumStencil sz / fromIntegral (totalElem sz)`` — a true division."""
    return _fold("avg", sz)


def maxStencil(sz) -> Stencil:
    """Stencil.hs:368-370 (Bounded e: integral and Bool elements only)"""
    return _fold("max", sz)


def minStencil(sz) -> Stencil:
    """Stencil.hs:401-403"""
    return _fold("min", sz)


def makeConvolutionStencilFromKernel(kArr) -> Stencil:
    """Stencil/Convolution.hs:64-80"""
    k = np.ascontiguousarray(kArr)
    return Stencil("conv", tuple(k.shape), None, k)


def makeCorrelationStencilFromKernel(kArr) -> Stencil:
    """Stencil/Convolution.hs:107-121"""
    k = np.ascontiguousarray(kArr)
    return Stencil("corr", tuple(k.shape), None, k)


def _closure(sz, center, listed, sign: int, inv: bool) -> Stencil:
    sz = Sz(sz)
    rank = len(sz)
    center = _bcast(center, rank)
    listed = [(_bcast(o, rank), k) for o, k in listed]
    rev = list(reversed(listed))  # f a . f b $ 0 applies b first
    taps = tuple(tuple(sign * v for v in o) for o, _ in rev)
    coeffs = np.array([k for _, k in rev])
    cen = tuple(s - 1 - c for s, c in zip(sz, center)) if inv else center
    return Stencil("taps", sz, cen, coeffs, None, taps)


def makeConvolutionStencil(sz, center, taps) -> Stencil:
    """Stencil/Convolution.hs:44-57.  ``taps`` is the closure body ``\\f -> f o1 k1 . f o2 k2 ...``
    given as the list ``[(o1, k1), (o2, k2), ...]`` in source order."""
    return _closure(sz, center, taps, -1, True)


makeUnsafeConvolutionStencil = makeConvolutionStencil  # Stencil/Unsafe.hs:51-64: same arithmetic


def makeCorrelationStencil(sz, center, taps) -> Stencil:
    """Stencil/Convolution.hs:89-100"""
    return _closure(sz, center, taps, +1, False)


makeUnsafeCorrelationStencil = makeCorrelationStencil  # Stencil/Unsafe.hs:71-82


lifeStencil = Stencil("life", (3, 3))
"""massiv-examples/GameOfLife/app/GameOfLife.hs:21-33"""


# ---- Padding (Stencil.hs:152-179) --------------------------------------------------------------

@dataclass(frozen=True)
class Padding:
    paddingFromOrigin: Union[int, Ix]
    paddingFromBottom: Union[int, Ix]
    paddingWithElement: Border


noPadding = Padding(0, 0, Edge)


def samePadding(stencil: Stencil, border: Border) -> Padding:
    if stencil.size is None:
        raise ValueError("samePadding needs a stencil of known rank; pass an explicit Sz")
    c = _center(stencil)
    g = _geom_size(stencil)
    return Padding(Sz(c), Sz(tuple(s - (cc + 1) for s, cc in zip(g, c))), border)


# ---- device, arrays -----------------------------------------------------------------------------

class Device:
    """One B200 + its library context: the analogue of the `Comp` a massiv array carries
    (Core/Common.hs:171-201) — where the load runs is chosen by the array's device."""

    _default = {}
    _lock = threading.Lock()

    def __init__(self, ordinal: int = 0):
        self.ordinal = ordinal
        self._h = C.c_void_p()
        L = _abi.lib()
        st = L.mb_ctx_create(ordinal, C.byref(self._h))
        if st:
            raise _abi.MassivB200Error(st, f"cannot create a context on cuda:{ordinal} "
                                           "(the stencil path has no CPU fallback)")
        self._L = L

    @classmethod
    def default(cls, ordinal: int = 0) -> "Device":
        with cls._lock:
            if ordinal not in cls._default:
                cls._default[ordinal] = Device(ordinal)
            return cls._default[ordinal]

    @property
    def handle(self):
        return self._h

    def check(self, st: int):
        _abi.raise_for(st, self._h)

    def close(self):
        if self._h:
            self._L.mb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def sync(self):
        self.check(self._L.mb_sync(self._h))

    def trim(self):
        """give the context's device scratch (staging arrays, padded pairs) back to the driver"""
        self.check(self._L.mb_ctx_trim(self._h))

    def set_option(self, key: str, value: int):
        self.check(self._L.mb_ctx_set_option(self._h, key.encode(), int(value)))

    def set_stream(self, cuda_stream: int):
        self.check(self._L.mb_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))

    @property
    def launch_count(self) -> int:
        return self._L.mb_launch_count(self._h)

    @property
    def aux_launch_count(self) -> int:
        """of launch_count: conditional follow-ups of the checked kernels, post-map passes, ghost fills"""
        return self._L.mb_aux_launch_count(self._h)

    @property
    def last_kernel(self) -> str:
        return self._L.mb_last_kernel(self._h).decode()

    def alloc(self, shape: Sequence[int], dtype, elem: Optional[str] = None) -> "DeviceArray":
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        p = C.c_void_p()
        self.check(self._L.mb_buf_alloc(self._h, n, C.byref(p)))
        return DeviceArray(self, p.value, tuple(int(s) for s in shape), dtype, elem or _NP2ELEM[dtype], owned=True)

    def fromHost(self, arr: np.ndarray, elem: Optional[str] = None) -> "DeviceArray":
        arr = np.ascontiguousarray(arr)
        d = self.alloc(arr.shape, arr.dtype, elem)
        if arr.nbytes:
            self.check(self._L.mb_upload(self._h, C.c_void_p(d.ptr), C.c_void_p(arr.ctypes.data), arr.nbytes))
        return d

    def wrap(self, ptr: int, shape, dtype, elem: Optional[str] = None) -> "DeviceArray":
        """View device memory owned by someone else (e.g. a torch tensor's data_ptr())."""
        dtype = np.dtype(dtype)
        return DeviceArray(self, int(ptr), tuple(int(s) for s in shape), dtype, elem or _NP2ELEM[dtype], owned=False)

    def timer_start(self):
        self.check(self._L.mb_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.check(self._L.mb_timer_stop(self._h, C.byref(ms)))
        return ms.value


class DeviceArray:
    """A manifest array resident in HBM (dense row-major, like `S`)."""

    def __init__(self, dev: Device, ptr: int, shape: Ix, dtype: np.dtype, elem: str, owned: bool):
        self.dev, self.ptr, self.shape, self.dtype, self.elem, self._owned = dev, ptr, shape, dtype, elem, owned

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def nbytes(self):
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    def toHost(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        if out.nbytes:
            self.dev.check(self.dev._L.mb_download(self.dev.handle, C.c_void_p(out.ctypes.data), C.c_void_p(self.ptr), out.nbytes))
        return out

    def free(self):
        if self._owned and self.ptr and self.dev.handle:
            self.dev._L.mb_buf_free(self.dev.handle, C.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


Array = Union[np.ndarray, DeviceArray]


# ---- DW: the delayed windowed array (Delayed/Windowed.hs:60-65) ---------------------------------

@dataclass
class DW:
    """What mapStencil/applyStencil return: nothing is computed until `compute`."""
    padding: Padding
    stencil: Stencil
    arr: Array
    elem: str
    _keep: list = field(default_factory=list)

    @property
    def rank(self):
        return self.arr.ndim

    def desc(self, stride=None) -> _abi.StencilDesc:
        st = self.stencil
        rank = self.rank
        d = _abi.StencilDesc()
        d.kind, d.elem, d.rank = _abi.KINDS[st.kind], _abi.ELEMS[self.elem], rank
        b = self.padding.paddingWithElement
        d.border = _abi.BORDERS[b.kind]
        cen = _center(st)
        lo = Sz(_bcast(self.padding.paddingFromOrigin, rank))
        hi = Sz(_bcast(self.padding.paddingFromBottom, rank))
        stride = (1,) * rank if stride is None else tuple(max(1, s) for s in _bcast(stride, rank))  # Stride clamps to >= 1
        for i in range(rank):
            d.size[i], d.center[i], d.pad_lo[i], d.pad_hi[i], d.stride[i] = st.size[i], cen[i], lo[i], hi[i], stride[i]
        npdt = _ELEM2NP[self.elem]
        if b.kind == "fill":
            fv = np.array([1 if b.value else 0], np.int32) if self.elem == "bool32" else np.array([b.value]).astype(npdt)
            for i, byte in enumerate(fv.tobytes()):
                d.fill[i] = byte
        keep = []
        if st.coeffs is not None:
            k1 = np.ascontiguousarray(np.asarray(st.coeffs).astype(npdt)).reshape(-1)
            keep.append(k1)
            d.coeffs = k1.ctypes.data
        if st.coeffs2 is not None:
            k2 = np.ascontiguousarray(np.asarray(st.coeffs2).astype(npdt)).reshape(-1)
            keep.append(k2)
            d.coeffs2 = k2.ctypes.data
        if st.taps is not None:
            offs = np.ascontiguousarray(np.asarray(st.taps, dtype=np.int64).reshape(-1, rank))
            keep.append(offs)
            d.ntaps, d.tap_offsets = offs.shape[0], offs.ctypes.data
        d.flags = st.flags
        if st.kind == "expr":
            terms, prog = _flatten(st.expr, rank)
            ta = (_abi.Term * len(terms))()
            for i, t in enumerate(terms):
                ta[i].kind = _abi.KINDS[t.kind]
                tc = _center(t)
                for j in range(rank):
                    ta[i].size[j], ta[i].center[j] = t.size[j], tc[j]
                if t.coeffs is not None:
                    kk = np.ascontiguousarray(np.asarray(t.coeffs).astype(npdt)).reshape(-1)
                    keep.append(kk)
                    ta[i].coeffs = kk.ctypes.data
                if t.taps is not None:
                    offs = np.ascontiguousarray(np.asarray(t.taps, dtype=np.int64).reshape(-1, rank))
                    keep.append(offs)
                    ta[i].ntaps, ta[i].tap_offsets = offs.shape[0], offs.ctypes.data
            pa = (_abi.XIns * len(prog))()
            for i, (op, arg) in enumerate(prog):
                pa[i].op = _abi.XOPS[op]
                if op == "term":
                    pa[i].arg = arg
                elif op == "const":
                    for j, byte in enumerate(np.array([arg]).astype(npdt).tobytes()):
                        pa[i].operand[j] = byte
            keep += [ta, pa]
            d.n_terms, d.n_program = len(terms), len(prog)
            d.terms, d.program = C.cast(ta, C.c_void_p), C.cast(pa, C.c_void_p)
        if st.post:
            pa = _abi.post_array(st.post, npdt)
            keep.append(pa)
            d.n_post, d.post = len(st.post), C.cast(pa, C.c_void_p)
        d._keep = keep  # the descriptor owns the buffers its pointers refer to
        return d

    def fmap(self, name: str, c=None) -> "DW":
        """Functor (Array DW ix) (Delayed/Windowed.hs:78-84): map over the windowed array before it is computed"""
        return DW(self.padding, self.stencil.fmap(name, c), self.arr, self.elem)

    def size(self, stride=None) -> Ix:
        d = self.desc(stride)
        od = (C.c_int64 * _abi.MAX_RANK)()
        _abi.raise_for(_abi.lib().mb_out_dims(C.byref(d), _abi.dims_array(self.arr.shape), od))
        return tuple(od[i] for i in range(self.rank))


def applyStencil(padding: Padding, stencil: Stencil, arr: Array, elem: Optional[str] = None) -> DW:
    """Stencil.hs:187-217"""
    st = stencil._resolved(arr.ndim)
    e = elem or (arr.elem if isinstance(arr, DeviceArray) else _NP2ELEM[np.asarray(arr).dtype])
    if not isinstance(arr, DeviceArray):
        arr = np.ascontiguousarray(arr)
    return DW(padding, st, arr, e)


def mapStencil(border: Border, stencil: Stencil, arr: Array, elem: Optional[str] = None) -> DW:
    """Stencil.hs:76-86: ``applyStencil (samePadding stencil border) stencil``"""
    st = stencil._resolved(arr.ndim)
    return applyStencil(samePadding(st, border), st, arr, elem)


def _device_of(dw: DW, device: Optional[Device]) -> Device:
    if isinstance(dw.arr, DeviceArray):
        return dw.arr.dev
    return device or Device.default()


def computeWithStride(stride, dw: DW, device: Optional[Device] = None) -> Array:
    """Manifest/Internal.hs:248-259 — `compute` when stride is None."""
    dev = _device_of(dw, device)
    d = dw.desc(stride)
    out_shape = dw.size(stride)
    L = dev._L
    if isinstance(dw.arr, DeviceArray):
        out = dev.alloc(out_shape, dw.arr.dtype, dw.elem)
        dev.check(L.mb_run_dev(dev.handle, C.byref(d), _abi.dims_array(dw.arr.shape), C.c_void_p(dw.arr.ptr), C.c_void_p(out.ptr)))
        return out
    out = np.empty(out_shape, dw.arr.dtype)
    dev.check(L.mb_run(dev.handle, C.byref(d), _abi.dims_array(dw.arr.shape), C.c_void_p(dw.arr.ctypes.data),
                       C.c_void_p(out.ctypes.data)))
    return out


def compute(dw: DW, device: Optional[Device] = None) -> Array:
    """Manifest/Internal.hs:65-67"""
    return computeWithStride(None, dw, device)


def computeInto(out: DeviceArray, dw: DW) -> DeviceArray:
    """Array/Mutable.hs:452-467: load into an existing array; sizes must match."""
    if tuple(out.shape) != dw.size():
        raise _abi.SizeMismatchException(_abi.ERR_SIZE_MISMATCH, f"computeInto: {out.shape} vs {dw.size()}")
    dev = out.dev
    d = dw.desc()
    dev.check(dev._L.mb_run_dev(dev.handle, C.byref(d), _abi.dims_array(dw.arr.shape), C.c_void_p(dw.arr.ptr), C.c_void_p(out.ptr)))
    return out


def iterateStencil(n: int, border: Border, stencil: Stencil, arr: Array, device: Optional[Device] = None) -> Array:
    """``iterateUntil (\\k _ _ -> k >= n-1) (\\_ -> mapStencil border stencil) arr``
    (Manifest/Internal.hs:331-397): n >= 1 applications, double-buffered.  `border` may also be a size-preserving
    :class:`Padding` (``applyStencil padding stencil`` as the step)."""
    dw = applyStencil(border, stencil._resolved(arr.ndim), arr) if isinstance(border, Padding) else mapStencil(border, stencil, arr)
    dev = _device_of(dw, device)
    d = dw.desc()
    L = dev._L
    dims = _abi.dims_array(dw.arr.shape)
    if isinstance(arr, DeviceArray):
        a = dev.alloc(arr.shape, arr.dtype, arr.elem)   # iterateUntil never mutates its argument
        b = dev.alloc(arr.shape, arr.dtype, arr.elem)
        if arr.nbytes:
            import ctypes
            # device-to-device copy through the library's own identity stencil keeps this torch-free
            idd = mapStencil(Edge, idStencil(arr.ndim), arr).desc()
            dev.check(L.mb_run_dev(dev.handle, ctypes.byref(idd), dims, C.c_void_p(arr.ptr), C.c_void_p(a.ptr)))
        res = C.c_void_p()
        dev.check(L.mb_iterate_dev(dev.handle, C.byref(d), dims, C.c_void_p(a.ptr), C.c_void_p(b.ptr), n, C.byref(res)))
        dev.sync()
        keep, drop = (a, b) if res.value == a.ptr else (b, a)
        drop.free()
        return keep
    out = np.empty_like(dw.arr)
    dev.check(L.mb_iterate(dev.handle, C.byref(d), dims, C.c_void_p(dw.arr.ctypes.data), C.c_void_p(out.ctypes.data), n))
    return out


def iterateUntil(predicate, border: Border, stencil: Stencil, arr: Array, max_steps: int, check_every: int = 1,
                 device: Optional[Device] = None):
    """``iterateUntil convergence (\\_ -> mapStencil border stencil) arr`` (Manifest/Internal.hs:331-349) with one of the
    predicates the device can evaluate: ``"equal"`` (``\\_ prev cur -> cur == prev``) or ``("max_abs_diff", tol)``
    (``maximum (abs (cur - prev)) < tol``), checked after every `check_every`-th application, at most `max_steps`
    applications.  Returns (array, applications made, converged)."""
    name, tol = (predicate, 0.0) if isinstance(predicate, str) else (predicate[0], float(predicate[1]))
    pred = {"equal": _abi.UNTIL_EQUAL, "max_abs_diff": _abi.UNTIL_MAX_ABS_DIFF}[name]
    dw = applyStencil(border, stencil._resolved(arr.ndim), arr) if isinstance(border, Padding) else mapStencil(border, stencil, arr)
    dev = _device_of(dw, device)
    d = dw.desc()
    L = dev._L
    dims = _abi.dims_array(dw.arr.shape)
    steps, conv = C.c_int64(), C.c_int32()
    if isinstance(arr, DeviceArray):
        a = dev.alloc(arr.shape, arr.dtype, arr.elem)   # iterateUntil never mutates its argument
        b = dev.alloc(arr.shape, arr.dtype, arr.elem)
        if arr.nbytes:
            idd = mapStencil(Edge, idStencil(arr.ndim), arr).desc()
            dev.check(L.mb_run_dev(dev.handle, C.byref(idd), dims, C.c_void_p(arr.ptr), C.c_void_p(a.ptr)))
        res = C.c_void_p()
        dev.check(L.mb_iterate_until_dev(dev.handle, C.byref(d), dims, C.c_void_p(a.ptr), C.c_void_p(b.ptr), max_steps, check_every,
                                         pred, tol, C.byref(steps), C.byref(conv), C.byref(res)))
        keep, drop = (a, b) if res.value == a.ptr else (b, a)
        drop.free()
        return keep, steps.value, bool(conv.value)
    out = np.empty_like(dw.arr)
    dev.check(L.mb_iterate_until(dev.handle, C.byref(d), dims, C.c_void_p(dw.arr.ctypes.data), C.c_void_p(out.ctypes.data), max_steps,
                                 check_every, pred, tol, C.byref(steps), C.byref(conv)))
    return out, steps.value, bool(conv.value)


def computeSeparable(border: Border, first: Stencil, second: Stencil, arr: Array, device: Optional[Device] = None) -> Array:
    """``compute . mapStencil border second . compute . mapStencil border first`` — the two-pass idiom
    for rank-1 factorable kernels, fused on the device when the shapes allow (mb_run_sep2)."""
    d1w, d2w = mapStencil(border, first, arr), mapStencil(border, second, arr)
    dev = _device_of(d1w, device)
    d1, d2 = d1w.desc(), d2w.desc()
    dims = _abi.dims_array(d1w.arr.shape)
    if isinstance(arr, DeviceArray):
        out = dev.alloc(arr.shape, arr.dtype, arr.elem)
        dev.check(dev._L.mb_run_sep2_dev(dev.handle, C.byref(d1), C.byref(d2), dims, C.c_void_p(arr.ptr), C.c_void_p(out.ptr), None))
        return out
    out = np.empty_like(d1w.arr)
    dev.check(dev._L.mb_run_sep2(dev.handle, C.byref(d1), C.byref(d2), dims, C.c_void_p(d1w.arr.ctypes.data), C.c_void_p(out.ctypes.data)))
    return out
