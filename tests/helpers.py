"""Shared helpers for the parity tests: run one case through the product (C ABI via the Python
host mirror) and through the oracle, on identical inputs."""
import numpy as np

import massiv_b200 as M
from massiv_b200 import _abi
from oracle import oracle as O

BORDERS = ["fill", "wrap", "edge", "reflect", "continue"]


def border_obj(border, fill=0):
    return {"fill": M.Fill(fill), "wrap": M.Wrap, "edge": M.Edge, "reflect": M.Reflect, "continue": M.Continue}[border]


def stencil_obj(kind, size, center=None, coeffs=None, coeffs2=None, taps=None, flags=0):
    size = tuple(size)
    if kind == "taps":
        return M.Stencil("taps", size, tuple(center), np.asarray(coeffs), None, tuple(tuple(t) for t in taps), flags)
    co = None if coeffs is None else np.asarray(coeffs).reshape(size)
    co2 = None if coeffs2 is None else np.asarray(coeffs2).reshape(size)
    return M.Stencil(kind, size, None, co, co2, None, flags)


def product_apply(kind, arr, *, size, center=None, pad_lo=None, pad_hi=None, border="fill", fill=0, stride=None,
                  coeffs=None, coeffs2=None, taps=None, flags=0, elem=None, device=None, resident=False):
    """compute(WithStride) (applyStencil ...) through libmassiv_b200.so"""
    st = stencil_obj(kind, size, center, coeffs, coeffs2, taps, flags)
    b = border_obj(border, fill)
    pad = M.samePadding(st, b) if pad_lo is None else M.Padding(tuple(pad_lo), tuple(pad_hi), b)
    src = arr
    if resident:
        dev = device or M.Device.default()
        src = dev.fromHost(arr, elem)
    dw = M.applyStencil(pad, st, src, elem)
    out = M.computeWithStride(stride, dw, device)
    return out.toHost() if resident else out


def oracle_apply(kind, arr, **kw):
    kw.pop("device", None)
    kw.pop("resident", None)
    return O.apply(kind, arr, **kw)


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """bit-exact, except that any NaN matches any NaN (payloads are hardware specific)"""
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype.kind == "f":
        na, nb = np.isnan(a), np.isnan(b)
        if not np.array_equal(na, nb):
            return False
        u = {4: np.uint32, 8: np.uint64}[a.dtype.itemsize]
        return np.array_equal(a.view(u)[~na], b.view(u)[~nb])
    return np.array_equal(a, b)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser over uint64 lanes (SURVEY.md §8d synthetic inputs)"""
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))
