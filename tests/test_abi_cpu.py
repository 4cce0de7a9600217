"""CPU-only checks of the C-ABI library and the host logic: the .so loads and exports every symbol
include/massiv_b200.h declares, the geometry / border arithmetic agrees with the oracle, error
statuses map to the reference's exceptions, and nothing computes without a GPU."""
import ctypes as C
import itertools
import os
import re

import numpy as np
import pytest

import massiv_b200 as M
from massiv_b200 import _abi
from oracle import oracle as O
from helpers import BORDERS, border_obj, stencil_obj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _abi.lib()
    header = open(os.path.join(ROOT, "include", "massiv_b200.h")).read()
    declared = set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", header))
    declared -= {"mb_ctx", "mb_stencil_desc"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in massiv_b200.h but not exported"
    assert set(_abi.SYMBOLS) == declared
    assert L.mb_version() == 100


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(M.MassivB200Error) as e:
        M.Device(0)
    assert e.value.status == _abi.ERR_CUDA
    with pytest.raises(M.MassivB200Error):
        M.compute(M.mapStencil(M.Edge, M.sumStencil(M.Sz(3, 3)), np.zeros((4, 4))))


def test_elem_sizes_follow_storable():
    L = _abi.lib()
    want = {"u8": 1, "i8": 1, "u16": 2, "i16": 2, "u32": 4, "i32": 4, "u64": 8, "i64": 8, "f32": 4, "f64": 8, "bool32": 4}
    for k, v in want.items():
        assert L.mb_elem_size(_abi.ELEMS[k]) == v
    assert L.mb_elem_size(99) == 0


def test_border_index_matches_oracle():
    for b in BORDERS:
        for k in range(1, 9):
            for i in range(-30, 30):
                want = O.border_index(b, k, i)
                got = M.handleBorderIndex(border_obj(b, "FILL"), (k,), lambda ix: ix[0], (i,))
                assert got == ("FILL" if want is None else want), (b, k, i)
    with pytest.raises(M.IndexZeroException):
        M.handleBorderIndex(M.Wrap, (0,), lambda ix: ix, (1,))
    # doctests Core/Index.hs:218-223
    assert M.handleBorderIndex(M.Fill(100), (2, 3), lambda ix: ix, (2, 3)) == 100
    assert M.handleBorderIndex(M.Wrap, (2, 3), lambda ix: ix, (2, 3)) == (0, 0)
    assert M.handleBorderIndex(M.Edge, (2, 3), lambda ix: ix, (2, 3)) == (1, 2)


def test_out_dims_match_oracle():
    rng = np.random.default_rng(0)
    for _ in range(400):
        rank = int(rng.integers(1, 6))
        kind = ["sum", "avg", "conv", "corr", "max", "id"][int(rng.integers(0, 6))]
        dims = tuple(int(rng.integers(0, 9)) for _ in range(rank))
        lo_sz = 1 if kind in ("conv", "corr") else 0
        size = (1,) * rank if kind == "id" else tuple(int(rng.integers(lo_sz, 6)) for _ in range(rank))
        pad_lo = tuple(int(rng.integers(0, 4)) for _ in range(rank))
        pad_hi = tuple(int(rng.integers(0, 4)) for _ in range(rank))
        stride = tuple(int(rng.integers(1, 5)) for _ in range(rank))
        n = int(np.prod(size))
        coeffs = np.ones(n) if kind in ("conv", "corr") else None
        elem = "f64" if kind == "avg" else "i64"
        dt = np.float64 if kind == "avg" else np.int64
        want = O.out_dims(kind, dims, elem=elem, size=size, pad_lo=pad_lo, pad_hi=pad_hi, stride=stride, coeffs=coeffs)
        st = stencil_obj(kind, size, coeffs=coeffs)
        dw = M.applyStencil(M.Padding(pad_lo, pad_hi, M.Edge), st, np.zeros(dims, dt))
        assert dw.size(stride) == want, (kind, dims, size, pad_lo, pad_hi, stride)


def test_map_stencil_keeps_the_size():
    for kind, size in (("sum", (3, 3)), ("avg", (2, 5)), ("conv", (3, 4)), ("corr", (4, 3)), ("life", (3, 3))):
        coeffs = np.ones(size) if kind in ("conv", "corr") else None
        dt = np.uint8 if kind == "life" else np.float64
        dw = M.mapStencil(M.Edge, stencil_obj(kind, size, coeffs=coeffs), np.zeros((7, 9), dt))
        assert dw.size() == (7, 9)


def test_stencil_centres_follow_the_constructors():
    k = np.arange(12.0).reshape(3, 4)
    assert M.getStencilCenter(M.makeConvolutionStencilFromKernel(k)) == (1, 1)   # size-1-(size quot 2)
    assert M.getStencilCenter(M.makeCorrelationStencilFromKernel(k)) == (1, 2)   # size div 2
    assert M.getStencilCenter(M.sumStencil(M.Sz(3, 3))) == (0, 0)
    assert M.getStencilSize(M.avgStencil(M.Sz(0, 3))) == (1, 3)                   # union with `pure`
    s = M.makeConvolutionStencil(M.Sz(4), 1, [((-1,), 1), ((0,), 2), ((1,), 3), ((2,), 4)])
    assert M.getStencilCenter(s) == (2,) and s.taps == ((-2,), (-1,), (0,), (1,))


def test_invalid_descriptors_are_rejected():
    L = _abi.lib()
    od = (C.c_int64 * 5)()

    def status(st, arr, pad=None, stride=None, elem=None):
        dw = M.applyStencil(pad or M.noPadding, st, arr, elem)
        return L.mb_out_dims(C.byref(dw.desc(stride)), _abi.dims_array(arr.shape), od)

    a = np.zeros((4, 4))
    assert status(M.Stencil("max", (3, 3)), a) == _abi.ERR_UNSUPPORTED            # no Bounded Double
    assert status(M.Stencil("avg", (3, 3)), a.astype(np.int64)) == _abi.ERR_UNSUPPORTED  # not Fractional
    assert status(M.Stencil("life", (3, 3)), a) == _abi.ERR_UNSUPPORTED
    assert status(M.Stencil("sum", (3, 3), center=(1, 1)), a) == _abi.ERR_INVALID_DESC  # fold stencils sit at 0
    assert status(M.Stencil("sum", (3, 3)), a.astype(np.int32), elem="bool32") == _abi.ERR_UNSUPPORTED
    assert status(M.Stencil("max", (3, 3)), a.astype(np.int32), elem="bool32") == _abi.OK
    d = M.applyStencil(M.noPadding, M.Stencil("sum", (3, 3)), a).desc()
    d.rank = 9
    assert L.mb_out_dims(C.byref(d), _abi.dims_array((4, 4)), od) == _abi.ERR_INVALID_DESC
    assert L.mb_out_dims(None, _abi.dims_array((4, 4)), od) == _abi.ERR_INVALID_DESC


def test_sobel_magnitude_lowers_to_mag2():
    kx = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32)
    sx, sy = M.makeConvolutionStencilFromKernel(kx), M.makeConvolutionStencilFromKernel(kx.T)
    op = M.sqrt(sx ** 2 + sy ** 2)
    assert op.kind == "mag2" and op.size == (3, 3) and M.getStencilCenter(op) == (1, 1)
    # kernels of different sizes: no MAG2 kernel, but still one pass — an expression stencil with the union geometry
    op2 = M.sqrt(sx ** 2 + M.makeConvolutionStencilFromKernel(np.ones((5, 5), np.float32)) ** 2)._resolved(2)
    assert op2.kind == "expr" and op2.size == (5, 5) and M.getStencilCenter(op2) == (2, 2)
