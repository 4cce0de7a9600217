"""The strided small-window kernel (csrc/k_direct2d.cu: pooling and strided convolution /
correlation over the two innermost dimensions) against the oracle, with the dispatcher forced onto it
for small arrays.  Includes the reference's own strided vectors (Stencil.hs:346-364,
StencilSpec.hs:298-309).  Bit-exact."""
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


def run_both(dev, kind, arr, size, border, stride, coeffs=None, fill=0, pad=None):
    kw = {} if pad is None else {"pad_lo": pad[0], "pad_hi": pad[1]}
    want = O.apply(kind, arr, size=size, border=border, fill=fill, stride=stride, coeffs=coeffs, **kw)
    co = None if coeffs is None else np.asarray(coeffs, arr.dtype).reshape(size)
    st = M.Stencil(kind, tuple(size), None, co)
    b = border_obj(border, arr.dtype.type(fill))
    padding = M.samePadding(st, b) if pad is None else M.Padding(tuple(pad[0]), tuple(pad[1]), b)
    got = M.computeWithStride(stride, M.applyStencil(padding, st, dev.fromHost(arr))).toHost()
    return got, want


def test_reference_vectors(dev):
    # Stencil.hs:346-364: computeWithStrideAs P (Stride 3) $ mapStencil Edge (maxStencil (Sz 3)) a, a = 10..90 as 9x9
    a = np.arange(10, 91, dtype=np.int64).reshape(9, 9)
    got, want = run_both(dev, "max", a, (3, 3), "edge", (3, 3))
    assert dev.last_kernel == "direct2d"
    assert got.tolist() == [[30, 33, 36], [57, 60, 63], [84, 87, 90]] and bits_equal(got, want)
    # StencilSpec.hs:298-309: convolution with Stride 2 on 3x3 (arr = 1..9) and 5x5 (1..25)
    k = [-1, 0, 1, 0, 1, 0, -1, 0, 1]
    got, _ = run_both(dev, "conv", np.arange(1, 10, dtype=np.int64).reshape(3, 3), (3, 3), "fill", (2, 2), k)
    assert dev.last_kernel == "direct2d" and got.tolist() == [[-4, 8], [2, 14]]
    got, _ = run_both(dev, "conv", np.arange(1, 26, dtype=np.int64).reshape(5, 5), (3, 3), "fill", (2, 2), k)
    assert got.tolist() == [[-6, 1, 14], [-13, 9, 43], [4, 21, 44]]


KINDS = [("max", "i64"), ("max", "u8"), ("min", "i16"), ("min", "u32"), ("sum", "i32"), ("sum", "f32"), ("sum", "f64"),
         ("avg", "f32"), ("avg", "f64"), ("product", "f64"), ("product", "u64"), ("conv", "f32"), ("conv", "f64"), ("conv", "i8"),
         ("corr", "f64"), ("corr", "u16")]
WINDOWS = [((2, 2), (2, 2)), ((3, 3), (3, 3)), ((3, 3), (2, 2)), ((4, 4), (4, 4)), ((2, 3), (1, 2)), ((5, 5), (3, 1)),
           ((1, 4), (3, 4)), ((3, 2), (5, 7)), ((8, 8), (8, 8)), ((3, 3), (1, 3))]


@pytest.mark.parametrize("kind,elem", KINDS)
@pytest.mark.parametrize("size,stride", WINDOWS)
def test_vs_oracle(dev, kind, elem, size, stride):
    dt = np.dtype(O.ELEM2NP[elem])
    seed = zlib.crc32(repr((kind, elem, size, stride)).encode())
    rng = np.random.default_rng(seed)
    for shape in ((37, 53), (64, 96), (3, 29, 31), (2, 2, 17, 40)):
        if dt.kind == "f":
            arr = (rng.standard_normal(shape) * (0.3 if kind == "product" else 1.0) + (1.0 if kind == "product" else 0.0)).astype(dt)
            co = rng.standard_normal(size).astype(dt)
        else:
            info = np.iinfo(dt)
            arr = rng.integers(info.min, info.max, size=shape, dtype=dt, endpoint=True)
            co = rng.integers(-3, 4, size=size).astype(dt)
        full_size = (1,) * (len(shape) - 2) + tuple(size)
        full_stride = (1,) * (len(shape) - 2) + tuple(stride)
        coeffs = co.reshape(-1).tolist() if kind in ("conv", "corr") else None
        for border in ("fill", BORDERS[1 + seed % 4]):
            got, want = run_both(dev, kind, arr, full_size, border, full_stride, coeffs, fill=2)
            if size[0] > 1 and size[1] > 1:
                assert dev.last_kernel == "direct2d", dev.last_kernel
            assert bits_equal(got, want), (shape, border)


def test_explicit_padding_and_non_finite(dev):
    rng = np.random.default_rng(3)
    arr = rng.standard_normal((40, 44))
    arr[3, 5], arr[10, 1], arr[20, 20], arr[0, 0] = np.inf, -np.inf, np.nan, -0.0
    k = rng.standard_normal((3, 3))
    k[1, 1] = 0.0
    for border in BORDERS:
        for pad in (((0, 0), (0, 0)), ((2, 1), (0, 3)), ((4, 4), (4, 4))):
            got, want = run_both(dev, "conv", arr, (3, 3), border, (2, 3), k.reshape(-1).tolist(), fill=-0.0, pad=pad)
            assert dev.last_kernel == "direct2d"
            assert bits_equal(got, want), (border, pad)
            got, want = run_both(dev, "avg", arr, (2, 2), border, (2, 2), fill=1.5, pad=pad)
            assert bits_equal(got, want), (border, pad)
