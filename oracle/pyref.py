"""Second, independent restatement of the reference semantics — pure-Python loops, tiny cases only.

TEST INFRASTRUCTURE ONLY (same rule as oracle.py).  It is deliberately written the way the
Haskell is — a Stencil is (size, centre, function of a getter and an index) and ``apply_stencil``
builds every cell through ``border_index`` — so that it shares no code path with
massiv_oracle.c.  tests/test_oracle_golden.py checks both against the reference's golden vectors
and against each other on random small cases.

Reference (paths under /root/reference/massiv/src/Data/Massiv/):
  Stencil         Array/Stencil/Internal.hs:30-34
  applyStencil    Array/Stencil.hs:197-217      mapStencil :76-86     samePadding :173-179
  borderIndex     Core/Common.hs:1024-1026      handleBorderIndex Core/Index.hs:236-253
  foldlStencil    Array/Stencil.hs:298-302      sum/product/max/min/avg :368-481
  FromKernel      Array/Stencil/Convolution.hs:64-80, 107-121    closure forms :44-57, 89-100
  strideSize      Core/Index/Stride.hs:102-105
  integral rules  Array/Numeric/Integral.hs:54-158, 226-287
"""
from __future__ import annotations

import itertools
import warnings
from dataclasses import dataclass
from typing import Callable, Sequence, Tuple

import numpy as np

Ix = Tuple[int, ...]


class IndexZeroException(Exception):
    pass


def _sz(v):  # Sz smart constructor clamps negatives to 0 (Core/Index/Internal.hs:122-125)
    return tuple(max(0, int(x)) for x in v)


def repair_index(k: int, i: int, below, over) -> int:  # Core/Index/Internal.hs:1030-1035
    if k <= 0:
        raise IndexZeroException(k)
    if i < 0:
        return below(k, i)
    if i >= k:
        return over(k, i)
    return i


def handle_border_index(border, sz: Ix, get: Callable[[Ix], object], ix: Ix):
    kind = border[0]
    if kind == "fill":
        if all(0 <= i < k for i, k in zip(ix, sz)):
            return get(ix)
        return border[1]
    if kind == "wrap":
        b = o = lambda k, i: i % k  # Python % is floored like Haskell `mod`
    elif kind == "edge":
        b, o = (lambda k, i: 0), (lambda k, i: k - 1)
    elif kind == "reflect":
        b = o = lambda k, i: (-i - 1) % k
    elif kind == "continue":
        b, o = (lambda k, i: (-i) % k), (lambda k, i: (-i - 2) % k)
    else:
        raise ValueError(border)
    return get(tuple(repair_index(k, i, b, o) for k, i in zip(sz, ix)))


@dataclass
class Stencil:
    size: Ix
    center: Ix
    func: Callable  # (get :: Ix -> e) -> Ix -> a      (absolute source index)


def _row_major(sz: Ix):
    return itertools.product(*[range(s) for s in sz])


def fold_stencil(f, acc0, sz: Ix) -> Stencil:  # foldlStencil: centre 0, row-major, acc = f acc x
    def func(get, ix):
        acc = acc0
        for q in _row_major(sz):
            acc = f(acc, get(tuple(a + b for a, b in zip(ix, q))))
        return acc
    return Stencil(tuple(sz), tuple(0 for _ in sz), func)


def id_stencil(rank: int) -> Stencil:
    return Stencil((1,) * rank, (0,) * rank, lambda get, ix: get(ix))


def sum_stencil(sz, dt) -> Stencil:
    return fold_stencil(lambda a, x: dt.type(a + x), dt.type(0), sz)


def product_stencil(sz, dt) -> Stencil:
    return fold_stencil(lambda a, x: dt.type(a * x), dt.type(1), sz)


def max_stencil(sz, dt, lo=None) -> Stencil:
    lo = np.iinfo(dt).min if lo is None else lo
    return fold_stencil(lambda a, x: x if x > a else a, dt.type(lo), sz)


def min_stencil(sz, dt, hi=None) -> Stencil:
    hi = np.iinfo(dt).max if hi is None else hi
    return fold_stencil(lambda a, x: x if x < a else a, dt.type(hi), sz)


def avg_stencil(sz, dt) -> Stencil:
    # sumStencil sz / fromIntegral (totalElem sz): liftA2 (/) with `pure n` (size 1, centre 0)
    s = sum_stencil(sz, dt)
    n = dt.type(int(np.prod(sz, dtype=np.int64)) if len(sz) else 1)
    size = tuple(max(a, 1) for a in sz)  # unionStencilSizes
    return Stencil(size, s.center, lambda get, ix: dt.type(s.func(get, ix) / n))


def conv_from_kernel(k: np.ndarray) -> Stencil:
    sz = k.shape
    sc = tuple(s // 2 for s in sz)  # quot; sizes >= 1
    center = tuple(s - 1 - c for s, c in zip(sz, sc))
    dt = k.dtype

    def func(get, ix):
        acc = dt.type(0)
        for kix in _row_major(sz):
            src = tuple(i + c - q for i, c, q in zip(ix, sc, kix))
            m = dt.type(get(src) * k[kix])
            acc = dt.type(m + acc)
        return acc
    return Stencil(sz, center, func)


def corr_from_kernel(k: np.ndarray) -> Stencil:
    sz = k.shape
    sc = tuple(s // 2 for s in sz)
    dt = k.dtype

    def func(get, ix):
        acc = dt.type(0)
        for kix in _row_major(sz):
            src = tuple(i - c + q for i, c, q in zip(ix, sc, kix))
            m = dt.type(get(src) * k[kix])
            acc = dt.type(m + acc)
        return acc
    return Stencil(sz, sc, func)


def conv_closure(sz: Ix, center: Ix, taps: Sequence[Tuple[Ix, object]], dt) -> Stencil:
    """makeConvolutionStencil sz centre (\\f -> f o1 k1 . f o2 k2 . ... ): the LAST listed tap is
    applied first; value taken at ix - off; stencilCenter is inverted (Convolution.hs:50-55)."""
    inv = tuple(s - 1 - c for s, c in zip(sz, center))

    def func(get, ix):
        acc = dt.type(0)
        for off, kv in reversed(list(taps)):
            m = dt.type(get(tuple(i - o for i, o in zip(ix, off))) * dt.type(kv))
            acc = dt.type(m + acc)
        return acc
    return Stencil(tuple(sz), inv, func)


def corr_closure(sz: Ix, center: Ix, taps, dt) -> Stencil:
    def func(get, ix):
        acc = dt.type(0)
        for off, kv in reversed(list(taps)):
            m = dt.type(get(tuple(i + o for i, o in zip(ix, off))) * dt.type(kv))
            acc = dt.type(m + acc)
        return acc
    return Stencil(tuple(sz), tuple(center), func)


def lift2(f, s1: Stencil, s2: Stencil) -> Stencil:  # liftA2 (Stencil/Internal.hs:83-111)
    mc = tuple(max(a, b) for a, b in zip(s1.center, s2.center))
    size = tuple(c + max(a - ca, b - cb) for c, a, ca, b, cb in
                 zip(mc, s1.size, s1.center, s2.size, s2.center))
    return Stencil(size, mc, lambda get, ix: f(s1.func(get, ix), s2.func(get, ix)))


def fmap(f, s: Stencil) -> Stencil:
    return Stencil(s.size, s.center, lambda get, ix: f(s.func(get, ix)))


def mag2(s1: Stencil, s2: Stencil, dt) -> Stencil:  # sqrt (fmap (^2) s1 + fmap (^2) s2)
    sq = lambda x: dt.type(x * x)
    return fmap(lambda v: dt.type(np.sqrt(v)), lift2(lambda a, b: dt.type(a + b), fmap(sq, s1), fmap(sq, s2)))


def post_fn(name: str, c, dt):
    """the Prelude function behind one post-map (GHC.Float / GHC.Num / GHC.Classes), in dtype dt"""
    t = dt.type
    isf = dt.kind == "f"

    def wrap(v):  # fixed-width integers wrap
        if isf:
            return t(v)
        bits = 8 * dt.itemsize
        v = int(v) & ((1 << bits) - 1)
        if dt.kind == "i" and v >= 1 << (bits - 1):
            v -= 1 << bits
        return t(v)
    ex = (lambda x: x) if isf else int
    cc = t(c) if c is not None else None
    table = {
        "negate": lambda x: t(-x) if isf else wrap(-int(x)),
        "abs": lambda x: (t(0) if x == 0 else (x if x > 0 else t(-x))) if isf else (x if x >= 0 else wrap(-int(x))),
        "signum": lambda x: (t(1) if x > 0 else (t(-1) if x < 0 else x)) if isf else (t(1) if x > 0 else (wrap(-1) if x < 0 else t(0))),
        "square": lambda x: wrap(ex(x) * ex(x)),
        "add": lambda x: wrap(ex(x) + ex(cc)),
        "sub": lambda x: wrap(ex(x) - ex(cc)),
        "rsub": lambda x: wrap(ex(cc) - ex(x)),
        "mul": lambda x: wrap(ex(x) * ex(cc)),
        "min": lambda x: x if x <= cc else cc,
        "max": lambda x: cc if x <= cc else x,
        "div": lambda x: t(x / cc),
        "rdiv": lambda x: t(cc / x),
        "recip": lambda x: t(t(1) / x),
        "sqrt": lambda x: t(np.sqrt(x)),
    }
    return table[name]


def with_post(st: Stencil, post, dt) -> Stencil:
    """fmap f1 . fmap f2 ... in application order"""
    for item in post:
        st = fmap(post_fn(item[0], item[1] if len(item) > 1 else None, dt), st)
    return st


def life_stencil() -> Stencil:  # GameOfLife.hs:15-36
    u8 = np.uint8

    def rules(c, n):
        if c == 0 and n == 3:
            return u8(1)
        if c == 1 and n in (2, 3):
            return u8(1)
        return u8(0)

    def func(get, ix):
        i, j = ix
        n = u8(0)
        for a, b in ((-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)):
            n = u8((int(n) + int(get((i + a, j + b)))) & 0xFF)
        return rules(get((i, j)), n)
    return Stencil((3, 3), (1, 1), func)


# ---- integral approximation (Array/Numeric/Integral.hs) ------------------------------------------
# Dim 1 is the innermost dimension: setDim' ix (Dim d) touches axis rank - d.

def _along(rank: int, dim: int, v: int, rest: int) -> Ix:
    if not 1 <= dim <= rank:
        raise IndexError(dim)
    return tuple(v if i == rank - dim else rest for i in range(rank))


def midpoint_stencil(dx, dim: int, k: int, rank: int, dt) -> Stencil:  # Integral.hs:54-66
    def func(get, ix):
        acc = dt.type(0)
        for i in range(k):  # loop 0 (< k) (+ 1) 0 (\i -> (+ g i)):  acc + g i
            acc = dt.type(acc + get(tuple(a + b for a, b in zip(ix, _along(rank, dim, i, 0)))))
        return dt.type(dt.type(dx) * acc)
    return Stencil(_along(rank, dim, k, 1), (0,) * rank, func)


def trapezoid_stencil(dx, dim: int, n: int, rank: int, dt) -> Stencil:  # Integral.hs:75-90
    def func(get, ix):
        g = lambda i: get(tuple(a + b for a, b in zip(ix, _along(rank, dim, i, 0))))
        acc = g(0)
        for i in range(1, n):  # loop 1 (< n) (+ 1) (g 0) (\i -> (+ 2 * g i))
            acc = dt.type(acc + dt.type(dt.type(2) * g(i)))
        acc = dt.type(acc + g(n))
        return dt.type(dt.type(dt.type(dx) / dt.type(2)) * acc)
    return Stencil(_along(rank, dim, n + 1, 1), (0,) * rank, func)


def simpsons_stencil(dx, dim: int, n: int, rank: int, dt) -> Stencil:  # Integral.hs:99-118
    if n % 2:
        raise ValueError("Number of sample points for Simpson's rule stencil should be even")

    def func(get, ix):
        g = lambda i: get(tuple(a + b for a, b in zip(ix, _along(rank, dim, i, 0))))

        def sim_acc(i, state):
            prev, acc = state
            fx3 = g(i + 2)
            new = dt.type(dt.type(dt.type(acc + prev) + dt.type(dt.type(4) * g(i + 1))) + fx3)
            return fx3, new
        state = sim_acc(0, (g(0), dt.type(0)))
        i = 2
        while i < n - 1:
            state = sim_acc(i, state)
            i += 2
        return dt.type(dt.type(dt.type(dx) / dt.type(3)) * state[1])
    return Stencil(_along(rank, dim, n + 1, 1), (0,) * rank, func)


def integrate_with(stencil_of, dim: int, n: int, arr: np.ndarray) -> np.ndarray:  # Integral.hs:121-135
    rank = arr.ndim
    st = stencil_of(dim, n)
    return map_stencil(("fill", arr.dtype.type(0)), st, arr, stride=_along(rank, dim, n, 1))


def integral_approx(stencil_of, d, sz: Ix, n: int, arr: np.ndarray) -> np.ndarray:  # Integral.hs:138-158
    dt = arr.dtype
    dx = dt.type(dt.type(d) / dt.type(n))
    for dim in range(1, arr.ndim + 1):
        arr = integrate_with(lambda dm, k: stencil_of(dx, dm, k, arr.ndim, dt), dim, n, arr)
    return arr[tuple(slice(0, s) for s in sz)]


def from_function(f, a, d, sz: Ix, n: int, dt) -> np.ndarray:  # Integral.hs:226-251
    shape = tuple(n * s + 1 for s in sz)
    scale = lambda i: dt.type(dt.type(a) + dt.type(dt.type(dt.type(d) * dt.type(i)) / dt.type(n)))
    out = np.empty(shape, dt)
    for ix in _row_major(shape):
        out[ix] = f(scale, ix)
    return out


def from_function_midpoint(f, a, d, sz: Ix, n: int, dt) -> np.ndarray:  # Integral.hs:253-287
    shape = tuple(n * s for s in sz)
    dx2 = dt.type(dt.type(dt.type(d) / dt.type(n)) / dt.type(2))
    scale = lambda i: dt.type(dt.type(dx2 + dt.type(a)) + dt.type(dt.type(dt.type(d) * dt.type(i)) / dt.type(n)))
    out = np.empty(shape, dt)
    for ix in _row_major(shape):
        out[ix] = f(scale, ix)
    return out


def same_padding(st: Stencil):
    return _sz(st.center), _sz(s - (c + 1) for s, c in zip(st.size, st.center))


def apply_stencil(pad_lo: Ix, pad_hi: Ix, border, st: Stencil, arr: np.ndarray, stride=None,
                  out_dtype=None) -> np.ndarray:
    """computeWithStride stride (applyStencil (Padding lo hi border) st arr)."""
    isz = arr.shape
    offset = tuple(c - p for c, p in zip(st.center, pad_lo))
    shrink = _sz(s - 1 for s in st.size)
    osz = _sz(a + b + n - s for a, b, n, s in zip(_sz(pad_lo), _sz(pad_hi), isz, shrink))
    stride = tuple(max(1, s) for s in (stride or (1,) * arr.ndim))
    ssz = tuple((n - 1) // s + 1 for n, s in zip(osz, stride))  # floored div like Haskell
    out = np.empty(ssz, dtype=out_dtype or arr.dtype)
    get = lambda ix: handle_border_index(border, isz, lambda i: arr[i], ix)
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for j in _row_major(ssz):
            o = tuple(a * s for a, s in zip(j, stride))
            out[j] = st.func(get, tuple(a + b for a, b in zip(o, offset)))
    return out


def map_stencil(border, st: Stencil, arr: np.ndarray, stride=None) -> np.ndarray:
    lo, hi = same_padding(st)
    return apply_stencil(lo, hi, border, st, arr, stride)
