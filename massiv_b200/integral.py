"""Host mirror of ``Data.Massiv.Array.Numeric.Integral`` — the one in-library caller of the stencil
path: integral approximation as 1-D stencils evaluated with a stride.

Same names, argument order and meaning as the reference
(/root/reference/massiv/src/Data/Massiv/Array/Numeric/Integral.hs); the stencils are reified
(MB_INT_MIDPOINT / MB_INT_TRAPEZOID / MB_INT_SIMPSON in include/massiv_b200.h) and run on the
device through ``computeWithStride``.  ``Dim 1`` is the innermost dimension, as in massiv.

Where Haskell infers the element type, these functions take it from the array (``integrateWith``,
``integralApprox``) or from the ``dtype`` argument (``fromFunction*``, the ``*Rule`` drivers).
The function to integrate receives ``scale`` and the index like in the reference, but vectorised:
``ix`` is a tuple of broadcastable integer arrays (one per dimension) and ``scale`` maps an index
array to sample positions in the element type.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from . import stencil as S
from .stencil import Array, Device, DeviceArray, Stencil, Sz


def _dx(dx, dtype=None):
    return np.asarray(dx if dtype is None else np.dtype(dtype).type(dx)).reshape(1)


def midpointStencil(dx, dim: int, k: int) -> Stencil:
    """Integral.hs:54-66: ``dx * (g 0 + g 1 + ... + g (k-1))`` along ``dim`` (size k, centre 0)."""
    return Stencil("int_midpoint", None, None, _dx(dx), along=(int(dim), int(k)))


def trapezoidStencil(dx, dim: int, n: int) -> Stencil:
    """Integral.hs:75-90: ``(dx / 2) * (g 0 + 2 g 1 + ... + 2 g (n-1) + g n)`` (size n + 1)."""
    return Stencil("int_trapezoid", None, None, _dx(dx), along=(int(dim), int(n) + 1))


def simpsonsStencil(dx, dim: int, n: int) -> Stencil:
    """Integral.hs:99-118; ``n`` must be even, otherwise error — as in the reference."""
    if n % 2:
        raise ValueError(f"Number of sample points for Simpson's rule stencil should be even, but received: {n}")
    return Stencil("int_simpson", None, None, _dx(dx), along=(int(dim), int(n) + 1))


def _with_dtype(st: Stencil, dtype) -> Stencil:
    """dx lives in the element type of the array it is applied to (``e`` in the reference)."""
    return Stencil(st.kind, st.size, st.center, np.asarray(st.coeffs).astype(dtype), st.coeffs2, st.taps, st.flags,
                   st.scalar_size, st.along)


def integrateWith(stencil: Callable[[int, int], Stencil], dim: int, n: int, arr: Array,
                  device: Optional[Device] = None) -> Array:
    """Integral.hs:121-135: ``computeWithStride (Stride nsz) $ mapStencil (Fill 0) (stencil dim n) arr``
    with ``nsz = setDim' (pureIndex 1) dim n``."""
    rank = arr.ndim
    if not 1 <= dim <= rank:
        raise IndexError(f"IndexDimensionException: (Dim {dim}) for an index of rank {rank}")
    nsz = tuple(n if i == rank - dim else 1 for i in range(rank))
    dtype = arr.dtype
    st = _with_dtype(stencil(dim, n), dtype)
    return S.computeWithStride(nsz, S.mapStencil(S.Fill(dtype.type(0)), st, arr), device)


def integralApprox(stencil: Callable[..., Stencil], d, sz, n: int, arr: Array,
                   device: Optional[Device] = None) -> Array:
    """Integral.hs:138-158: integrate along every dimension in turn (Dim 1 first), then
    ``extract' zeroIndex sz``.  ``dx = d / fromIntegral n`` in the element type."""
    dtype = arr.dtype
    rank = arr.ndim
    sz = Sz(S._bcast(sz, rank))
    dx = dtype.type(d) / dtype.type(n)
    for dim in range(1, rank + 1):
        arr = integrateWith(lambda dm, k: stencil(dx, dm, k), dim, n, arr, device)
    for have, want in zip(arr.shape, sz):  # extract' throws SizeSubregionException
        if want > have:
            raise IndexError(f"SizeSubregionException: {tuple(arr.shape)} zeroIndex {tuple(sz)}")
    if isinstance(arr, DeviceArray):
        host = arr.toHost()
        return arr.dev.fromHost(np.ascontiguousarray(host[tuple(slice(0, s) for s in sz)]))
    return np.ascontiguousarray(arr[tuple(slice(0, s) for s in sz)])


def fromFunction(comp, f: Callable, a, d, sz, n: int, dtype=np.float64) -> np.ndarray:
    """Integral.hs:226-251: samples at the cell edges, ``n * sz + 1`` per dimension;
    ``scale i = a + d * fromIntegral i / nFrac``.  ``comp`` is accepted for signature parity."""
    del comp
    t = np.dtype(dtype).type
    sz = Sz(sz)
    nf = t(n)

    def scale(i):
        return t(a) + t(d) * np.asarray(i).astype(dtype) / nf

    ix = np.indices(tuple(n * s + 1 for s in sz), sparse=True)
    return np.ascontiguousarray(np.broadcast_to(np.asarray(f(scale, tuple(ix)), dtype=dtype), tuple(n * s + 1 for s in sz)))


def fromFunctionMidpoint(comp, f: Callable, a, d, sz, n: int, dtype=np.float64) -> np.ndarray:
    """Integral.hs:253-287: samples in the middle of the cells, ``n * sz`` per dimension;
    ``scale i = dx2 + a + d * fromIntegral i / nFrac`` with ``dx2 = d / nFrac / 2``."""
    del comp
    t = np.dtype(dtype).type
    sz = Sz(sz)
    nf = t(n)
    dx2 = t(d) / nf / t(2)

    def scale(i):
        return (dx2 + t(a)) + t(d) * np.asarray(i).astype(dtype) / nf

    shape = tuple(n * s for s in sz)
    ix = np.indices(shape, sparse=True)
    return np.ascontiguousarray(np.broadcast_to(np.asarray(f(scale, tuple(ix)), dtype=dtype), shape))


def midpointRule(comp, r, f: Callable, a, d, sz, n: int, dtype=np.float64, device: Optional[Device] = None) -> Array:
    """Integral.hs:160-180.  ``r`` (the intermediate representation) is accepted for parity."""
    del r
    return integralApprox(midpointStencil, d, sz, n, fromFunctionMidpoint(comp, f, a, d, sz, n, dtype), device)


def trapezoidRule(comp, r, f: Callable, a, d, sz, n: int, dtype=np.float64, device: Optional[Device] = None) -> Array:
    """Integral.hs:182-202"""
    del r
    return integralApprox(trapezoidStencil, d, sz, n, fromFunction(comp, f, a, d, sz, n, dtype), device)


def simpsonsRule(comp, r, f: Callable, a, d, sz, n: int, dtype=np.float64, device: Optional[Device] = None) -> Array:
    """Integral.hs:204-225; ``n`` must be even."""
    del r
    return integralApprox(simpsonsStencil, d, sz, n, fromFunction(comp, f, a, d, sz, n, dtype), device)


__all__ = ["midpointStencil", "trapezoidStencil", "simpsonsStencil", "integrateWith", "integralApprox",
           "fromFunction", "fromFunctionMidpoint", "midpointRule", "trapezoidRule", "simpsonsRule"]
