"""Rank 3-5 arrays whose stencil has extent 1 in the leading dimensions (a 2-D stencil over every slice of
a stack, a 3-D one over every volume of a batch) run the tiled kernels slice by slice; results against
the oracle.  Bit-exact."""
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


CASES = [
    ("conv", (5, 70, 90), (1, 3, 3), "tile2d"), ("avg", (4, 64, 80), (1, 3, 3), "tile2d"), ("sum", (2, 3, 40, 72), (1, 1, 5, 5), "tile2d"),
    ("corr", (3, 2, 2, 33, 70), (1, 1, 1, 3, 3), "tile2d"), ("conv", (3, 20, 36, 40), (1, 3, 3, 3), "vol3d"),
    ("conv", (2, 2, 12, 36, 40), (1, 1, 3, 3, 3), "vol3d"), ("max", (6, 50, 64), (1, 3, 3), "tile2d"), ("mag2", (3, 48, 96), (1, 3, 3), "tile2d"),
]


@pytest.mark.parametrize("kind,shape,size,family", CASES)
@pytest.mark.parametrize("elem", ["f32", "f64"])
def test_stacks(dev, kind, shape, size, family, elem):
    if kind == "max":
        elem = "i32" if elem == "f32" else "i64"
    dt = np.dtype(O.ELEM2NP[elem])
    rng = np.random.default_rng(zlib.crc32(repr((kind, shape, size, elem)).encode()))
    arr = rng.standard_normal(shape).astype(dt) if dt.kind == "f" else rng.integers(-99, 99, size=shape).astype(dt)
    n = int(np.prod(size))
    co = rng.standard_normal(n).astype(dt).tolist() if kind in ("conv", "corr", "mag2") else None
    co2 = rng.standard_normal(n).astype(dt).tolist() if kind == "mag2" else None
    for border in BORDERS:
        want = O.apply(kind, arr, size=size, border=border, fill=1, coeffs=co, coeffs2=co2)
        st = M.Stencil(kind, size, None, None if co is None else np.asarray(co, dt).reshape(size),
                       None if co2 is None else np.asarray(co2, dt).reshape(size))
        got = M.compute(M.mapStencil(border_obj(border, 1), st, dev.fromHost(arr))).toHost()
        assert dev.last_kernel.startswith(family), (dev.last_kernel, border)
        assert bits_equal(got, want), border


def test_iterate_over_a_stack_and_fallbacks(dev):
    rng = np.random.default_rng(3)
    arr = rng.random((4, 64, 96))
    k = rng.standard_normal((1, 3, 3))
    st = M.makeCorrelationStencilFromKernel(k)
    want = O.iterate("corr", arr, 3, size=(1, 3, 3), coeffs=k.reshape(-1).tolist(), border="reflect")
    assert bits_equal(M.iterateStencil(3, M.Reflect, st, dev.fromHost(arr)).toHost(), want)
    assert dev.last_kernel.startswith("tile2d")
    # a window the compile-time-shaped kernels do not have (4x4): the small-window kernel, slice by slice
    got = M.compute(M.mapStencil(M.Edge, M.sumStencil(M.Sz(1, 4, 4)), dev.fromHost(arr))).toHost()
    assert dev.last_kernel == "smallwin" and bits_equal(got, O.apply("sum", arr, size=(1, 4, 4), border="edge"))
    # padding in a leading dimension: not a stack of independent slices
    dw = M.applyStencil(M.Padding((1, 1, 1), (0, 1, 1), M.Fill(0.0)), M.sumStencil(M.Sz(1, 3, 3)), dev.fromHost(arr))
    got = M.compute(dw).toHost()
    assert bits_equal(got, O.apply("sum", arr, size=(1, 3, 3), pad_lo=(1, 1, 1), pad_hi=(0, 1, 1), border="fill", fill=0.0))
    # post-maps ride along
    got = M.compute(M.mapStencil(M.Edge, M.avgStencil(M.Sz(1, 3, 3)).fmap("mul", 2.0), dev.fromHost(arr))).toHost()
    assert bits_equal(got, O.apply("avg", arr, size=(1, 3, 3), border="edge", post=[("mul", 2.0)]))
