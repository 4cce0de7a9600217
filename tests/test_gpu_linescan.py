"""The line-window kernels (csrc/k_linescan.cu: windows along one dimension — integral rules and 1-D
folds, strided or not) against the oracle, with the dispatcher forced onto them for small arrays.
Bit-exact."""
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from helpers import bits_equal, border_obj
from oracle import oracle as O

pytestmark = pytest.mark.gpu

BORDERS = ["fill", "wrap", "edge", "reflect", "continue"]


@pytest.fixture(scope="module")
def dev():
    d = M.Device.default()
    d.set_option("tile_min_elems", 0)
    d.set_option("force_generic", 0)
    yield d
    d.set_option("tile_min_elems", 16384)


def run_both(dev, kind, arr, size, border, stride, coeffs=None, fill=0):
    want = O.apply(kind, arr, size=size, border=border, fill=fill, stride=stride, coeffs=coeffs)
    st = M.Stencil(kind, tuple(size), None, None if coeffs is None else np.asarray(coeffs, arr.dtype))
    dw = M.mapStencil(border_obj(border, arr.dtype.type(fill)), st, dev.fromHost(arr))
    got = M.computeWithStride(stride, dw).toHost()
    return got, want


RULES = ["int_midpoint", "int_trapezoid", "int_simpson"]
SHAPES = [  # (shape, axis)
    ((5, 700), 1), ((700, 5), 0), ((3, 130, 4), 1), ((2, 3, 97), 2), ((97, 2, 3), 0), ((1, 1, 260), 2), ((300,), 0),
    ((2, 2, 40, 3), 2), ((2, 2, 3, 41), 3),
]


@pytest.mark.parametrize("kind", RULES)
@pytest.mark.parametrize("shape,axis", SHAPES)
@pytest.mark.parametrize("elem", ["f32", "f64"])
def test_rules(dev, kind, shape, axis, elem):
    dt = np.dtype(O.ELEM2NP[elem])
    rng = np.random.default_rng(zlib.crc32(repr((kind, shape, axis, elem)).encode()))
    arr = rng.standard_normal(shape).astype(dt)
    for n in (2, 6, 16, 40, 100):
        if n >= shape[axis] * 2:
            continue
        k = n if kind == "int_midpoint" else n + 1
        size = tuple(k if i == axis else 1 for i in range(len(shape)))
        dx = dt.type(rng.uniform(0.01, 1.0))
        for border in ("fill", BORDERS[1 + n % 4]):
            for stride in (tuple(n if i == axis else 1 for i in range(len(shape))), None, tuple(3 if i == axis else 1 for i in range(len(shape)))):
                got, want = run_both(dev, kind, arr, size, border, stride, [dx])
                assert dev.last_kernel in ("linescan_rows", "linescan_cols"), dev.last_kernel
                assert dev.last_kernel == ("linescan_rows" if axis == len(shape) - 1 or int(np.prod(shape[axis + 1:])) == 1 else "linescan_cols")
                assert bits_equal(got, want), (n, border, stride)


FOLDS = [("sum", e) for e in ("u8", "i16", "i32", "u64", "f32", "f64")] + [("avg", "f32"), ("avg", "f64")] + \
        [("product", e) for e in ("i8", "u32", "f64")] + [(k, e) for k in ("max", "min") for e in ("u8", "i8", "i16", "u32", "i64")]


@pytest.mark.parametrize("kind,elem", FOLDS)
@pytest.mark.parametrize("shape,axis", SHAPES[:6])
def test_folds(dev, kind, elem, shape, axis):
    dt = np.dtype(O.ELEM2NP[elem])
    rng = np.random.default_rng(zlib.crc32(repr((kind, elem, shape, axis)).encode()))
    if dt.kind == "f":
        arr = (rng.standard_normal(shape) * (1.0 if kind != "product" else 0.2) + (1.0 if kind == "product" else 0.0)).astype(dt)
    else:
        info = np.iinfo(dt)
        arr = rng.integers(info.min, info.max, size=shape, dtype=dt, endpoint=True)
    for k, s in ((9, 9), (33, 7), (70, 1), (2, 5), (12, 12)):
        if k >= shape[axis] * 2:
            continue
        size = tuple(k if i == axis else 1 for i in range(len(shape)))
        stride = tuple(s if i == axis else 1 for i in range(len(shape)))
        for border in ("fill", BORDERS[1 + k % 4]):
            got, want = run_both(dev, kind, arr, size, border, stride, fill=3)
            assert dev.last_kernel.startswith("linescan"), dev.last_kernel
            assert bits_equal(got, want), (k, s, border)


def test_not_taken_when_it_should_not_be(dev):
    arr = np.random.default_rng(1).random((64, 64))
    # padding elsewhere than along the window's dimension: not a pure line problem
    st = M.Stencil("sum", (1, 5))
    dw = M.applyStencil(M.Padding((1, 0), (0, 4), M.Fill(0.0)), st, dev.fromHost(arr))
    got = M.compute(dw).toHost()
    assert not dev.last_kernel.startswith("linescan")
    want = O.apply("sum", arr, size=(1, 5), pad_lo=(1, 0), pad_hi=(0, 4), border="fill", fill=0.0)
    assert bits_equal(got, want)
    # a window in two dimensions
    got = M.computeWithStride((3, 3), M.mapStencil(M.Edge, M.Stencil("sum", (3, 3)), dev.fromHost(arr))).toHost()
    assert not dev.last_kernel.startswith("linescan")
    assert bits_equal(got, O.apply("sum", arr, size=(3, 3), border="edge", stride=(3, 3)))


def test_non_finite_and_signed_zero(dev):
    arr = np.array([[0.0, -0.0, np.inf, -np.inf, np.nan, 1e308, -1e308, 5e-324] * 8] * 4)
    for kind, k in (("int_trapezoid", 5), ("int_simpson", 5), ("int_midpoint", 4), ("sum", 4), ("avg", 8), ("product", 3)):
        coeffs = [0.5] if kind.startswith("int_") else None
        for border in ("fill", "reflect"):
            got, want = run_both(dev, kind, arr, (1, k), border, (1, 4), coeffs, fill=-0.0)
            assert dev.last_kernel == "linescan_rows"
            assert bits_equal(got, want), (kind, border)
            got, want = run_both(dev, kind, np.ascontiguousarray(arr.T), (k, 1), border, (4, 1), coeffs, fill=-0.0)
            assert dev.last_kernel == "linescan_cols"
            assert bits_equal(got, want), (kind, border)
