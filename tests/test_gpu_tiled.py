"""GPU parity of the tiled fast kernels (tile2d / vol3d / life) against the oracle, with the
dispatcher forced onto them even for small arrays (`tile_min_elems = 0`), over shapes chosen to hit
partial tiles, several tile rows/columns, TMA-describable and non-describable pitches, every border
and arbitrary padding.  Bit-exact (0 ulp) for floats and integers."""
import ctypes as C
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from massiv_b200 import _abi
from oracle import oracle as O
from helpers import BORDERS, bits_equal, border_obj, oracle_apply, product_apply

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = M.Device.default()
    d.set_option("tile_min_elems", 0)
    d.set_option("force_generic", 0)
    yield d
    d.set_option("tile_min_elems", 16384)


SHAPES2D = [(33, 130), (64, 128), (5, 7), (97, 260), (32, 512), (70, 131), (1, 300), (257, 64), (40, 1000)]
FLOAT_CASES = [("avg", (3, 3)), ("sum", (3, 3)), ("sum", (5, 5)), ("avg", (7, 7)), ("sum", (2, 2)), ("product", (3, 3)),
               ("avg", (1, 3)), ("sum", (3, 1)), ("conv", (3, 3)), ("corr", (3, 3)), ("conv", (5, 5)), ("corr", (7, 7)),
               ("conv", (7, 7)), ("conv", (1, 7)), ("conv", (7, 1)), ("corr", (1, 5)), ("corr", (5, 1)), ("conv", (1, 3)),
               ("corr", (3, 1)), ("conv", (2, 2)), ("corr", (2, 2)), ("mag2", (3, 3)), ("mag2", (5, 5))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind,size", FLOAT_CASES, ids=[f"{k}{s[0]}x{s[1]}" for k, s in FLOAT_CASES])
def test_tile2d_float_matches_oracle(kind, size, dtype, dev):
    rng = np.random.default_rng(zlib.crc32(f"{kind}{size}{np.dtype(dtype).name}".encode()))
    n = size[0] * size[1]
    for shape in SHAPES2D:
        a = rng.standard_normal(shape).astype(dtype)
        for border in BORDERS:
            kw = dict(size=size, border=border, fill=float(rng.standard_normal()) if rng.random() < 0.5 else 0.0)
            if kind in ("conv", "corr", "mag2"):
                co = rng.standard_normal(n).astype(dtype)
                co[rng.random(n) < 0.2] = 0
                kw["coeffs"] = co.tolist()
            if kind == "mag2":
                kw["coeffs2"] = rng.standard_normal(n).astype(dtype).tolist()
            if rng.random() < 0.3:
                kw["pad_lo"] = (int(rng.integers(0, 4)), int(rng.integers(0, 5)))
                kw["pad_hi"] = (int(rng.integers(0, 4)), int(rng.integers(0, 5)))
            want = oracle_apply(kind, a, **kw)
            got = product_apply(kind, a, resident=True, **kw)
            # mapStencil-shaped calls must take the tiled kernel; with arbitrary padding a Float window may
            # start at a vector offset that is not instantiated and falls back to the generic kernel
            if want.size >= 1 and ("pad_lo" not in kw or np.dtype(dtype).itemsize == 8):
                assert dev.last_kernel == "tile2d", (kind, size, shape, dev.last_kernel)
            assert bits_equal(got, want), (kind, size, shape, kw)


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("kind", ["sum", "max", "min"])
@pytest.mark.parametrize("size", [(3, 3), (2, 2)])
def test_tile2d_int_matches_oracle(kind, size, dtype, dev):
    rng = np.random.default_rng(zlib.crc32(f"{kind}{size}{np.dtype(dtype).name}".encode()))
    info = np.iinfo(dtype)
    for shape in SHAPES2D[:6]:
        a = rng.integers(info.min, info.max, shape, dtype=dtype)
        for border in BORDERS:
            want = oracle_apply(kind, a, size=size, border=border, fill=7)
            got = product_apply(kind, a, size=size, border=border, fill=7, resident=True)
            assert dev.last_kernel == "tile2d"
            assert np.array_equal(got, want), (kind, size, shape, border)


def test_tile2d_non_finite_and_signed_zero(dev):
    a = np.ones((70, 200), np.float32)
    a[3, 4], a[60, 1], a[0, 0], a[69, 199], a[10, 10] = np.inf, np.nan, -np.inf, np.inf, -0.0
    a[20:30, 50:60] = -0.0
    k = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)
    for kind in ("conv", "corr", "sum", "avg"):
        kw = dict(size=(3, 3), border="reflect")
        if kind in ("conv", "corr"):
            kw["coeffs"] = k.reshape(-1).tolist()
        want = oracle_apply(kind, a, **kw)
        got = product_apply(kind, a, resident=True, **kw)
        assert dev.last_kernel in ("tile2d", "tile2d_known")  # this array is the 4-neighbour Laplacian
        assert bits_equal(got, want), kind
    d = a.astype(np.float64)
    d[5, 5] = 5e-324
    d[6, 6] = 1e-308
    assert bits_equal(product_apply("avg", d, size=(3, 3), border="edge", resident=True),
                      oracle_apply("avg", d, size=(3, 3), border="edge"))


def test_tile2d_large_multi_wave(dev):
    """more tiles than resident CTAs, so every CTA loops through its TMA ring several times"""
    rng = np.random.default_rng(1)
    a = rng.random((1100, 2300))
    for border in ("edge", "fill", "wrap"):
        want = O.apply("avg", a, size=(3, 3), border=border, fill=0.0, nthreads=8)
        got = product_apply("avg", a, size=(3, 3), border=border, fill=0.0, resident=True)
        assert dev.last_kernel == "tile2d" and bits_equal(got, want)
    f = rng.random((1500, 4100)).astype(np.float32)   # pitch not a multiple of 16 B ... 4100*4 is; use 4101 below
    kx = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32)
    kw = dict(size=(3, 3), coeffs=kx.reshape(-1).tolist(), coeffs2=kx.T.reshape(-1).tolist(), border="reflect")
    assert bits_equal(product_apply("mag2", f, resident=True, **kw), O.apply("mag2", f, nthreads=8, **kw))
    g = rng.random((700, 4101)).astype(np.float32)    # odd pitch: TMA cannot describe it -> peeled fill everywhere
    assert bits_equal(product_apply("mag2", g, resident=True, **kw), O.apply("mag2", g, nthreads=8, **kw))
    assert dev.last_kernel == "tile2d_known"  # the Sobel pair has compile-time coefficients


@pytest.mark.parametrize("elem,n", [("f64", 9), ("f64", 25), ("f64", 49), ("f64", 6), ("f64", 3), ("f64", 4), ("f64", 1000),
                                    ("f32", 9), ("f32", 25), ("f32", 49), ("f32", 6), ("f32", 10), ("f32", 1024)])
def test_exact_division(elem, n, dev):
    """avgStencil divides by fromIntegral n; the kernels use a 3-operation exact quotient"""
    bad, first = C.c_uint64(), C.c_uint64()
    dev.check(dev._L.mb_selftest_div(dev.handle, _abi.ELEMS[elem], n, 1 << 28, 12345 + n, C.byref(bad), C.byref(first)))
    assert bad.value == 0, f"{bad.value} mismatching quotients, first bits {first.value:#x}"


# ---- Life: bit-packed temporally blocked kernel ---------------------------------------------------

@pytest.mark.parametrize("shape,n", [((64, 64), 40), ((48, 128), 33), ((24, 64), 17), ((96, 192), 50), ((128, 256), 3),
                                     ((16, 64), 70), ((2048, 2048), 64)])
@pytest.mark.parametrize("border", ["wrap", "fill"])
def test_life_bits_matches_oracle(shape, n, border, dev):
    rng = np.random.default_rng(shape[0] * 7 + n)
    a = (rng.random(shape) < 0.37).astype(np.uint8)
    a[rng.random(shape) < 0.01] = 200  # first generation must still be byte-exact
    b = M.Wrap if border == "wrap" else M.Fill(0)
    want = O.iterate("life", a, n, size=(3, 3), border=border, fill=0, nthreads=8)
    got = M.iterateStencil(n, b, M.lifeStencil, dev.fromHost(a)).toHost()
    if shape[1] % 64 == 0 and (shape[1] // 64) & (shape[1] // 64 - 1) == 0:
        assert dev.last_kernel == "life_bits", dev.last_kernel
    assert np.array_equal(got, want)
    dev.set_option("life_bitpack", 0)
    try:
        got2 = M.iterateStencil(n if shape[0] < 1000 else 5, b, M.lifeStencil, dev.fromHost(a)).toHost()
        assert dev.last_kernel == "life_byte"
        if shape[0] < 1000:
            assert np.array_equal(got2, want)
    finally:
        dev.set_option("life_bitpack", 1)


@pytest.mark.parametrize("border", BORDERS)
def test_life_byte_kernel_all_borders(border, dev):
    rng = np.random.default_rng(3)
    for shape in ((33, 64), (7, 12), (100, 260)):
        a = rng.integers(0, 3, shape).astype(np.uint8)
        want = O.apply("life", a, size=(3, 3), border=border, fill=1)
        got = M.compute(M.mapStencil(border_obj(border, 1), M.lifeStencil, dev.fromHost(a))).toHost()
        assert dev.last_kernel == "life_byte"
        assert np.array_equal(got, want), (border, shape)


# ---- rank 3 ---------------------------------------------------------------------------------------------

def _heat(r=0.1, dt=np.float64):
    k = np.zeros((3, 3, 3), dt)
    k[1, 1, 1] = 1 - 6 * r
    for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = r
    return k


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("mode", ["star7", "dense27", "dense27-random", "corr-star7", "corr-random"])
def test_vol3d_matches_oracle(mode, dtype, dev):
    rng = np.random.default_rng(zlib.crc32(f"{mode}{np.dtype(dtype).name}".encode()))
    for shape in ((9, 33, 70), (40, 64, 128), (5, 20, 260), (70, 40, 64), (3, 3, 3), (34, 65, 129)):
        a = rng.standard_normal(shape).astype(dtype)
        k = _heat(dt=dtype) if "random" not in mode else rng.standard_normal((3, 3, 3)).astype(dtype)
        kind = "corr" if mode.startswith("corr") else "conv"
        flags = _abi.FLAG_ASSUME_FINITE if "star7" in mode else 0
        for border in BORDERS:
            fill = 0.0 if rng.random() < 0.6 else 1.5
            kw = dict(size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border=border, fill=fill)
            want = O.apply(kind, a, **kw)
            got = product_apply(kind, a, flags=flags, resident=True, **kw)
            # a star-shaped kernel without the finite promise runs the star kernel optimistically (checked)
            assert dev.last_kernel == ("vol3d_dense27" if "random" in mode else "vol3d_star7"), dev.last_kernel
            assert bits_equal(got, want), (mode, shape, border)


def test_vol3d_non_finite_strict(dev):
    a = np.ones((12, 40, 70))
    a[3, 4, 5], a[8, 30, 60], a[0, 0, 0] = np.inf, np.nan, -np.inf
    k = _heat()
    kw = dict(size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0)
    want = O.apply("conv", a, **kw)
    got = product_apply("conv", a, resident=True, **kw)   # strict: the 20 zero taps spread NaN to the corners
    assert np.isnan(want[2, 3, 4]) and bits_equal(got, want)   # optimistic star pass, check fails, strict redo
    dev.set_option("no_optimistic", 1)
    try:
        got = product_apply("conv", a, resident=True, **kw)
        assert dev.last_kernel == "vol3d_dense27" and bits_equal(got, want)
    finally:
        dev.set_option("no_optimistic", 0)
