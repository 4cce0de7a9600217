"""Compile-time-coefficient kernels for the classic 3x3 arrays (k_tile2d_known.cu): same operation
order and roundings as the generic coefficient path, so results must stay bit-identical to the
oracle — including NaN/Inf flowing through zero coefficients in strict mode."""
import numpy as np
import pytest

import massiv_b200 as M
from massiv_b200 import _abi
from oracle import oracle as O
from helpers import BORDERS, bits_equal, border_obj, product_apply

pytestmark = pytest.mark.gpu

SOBEL_X = [-1, 0, 1, -2, 0, 2, -1, 0, 1]
SOBEL_Y = [-1, -2, -1, 0, 0, 0, 1, 2, 1]
PREWITT_X = [-1, 0, 1, -1, 0, 1, -1, 0, 1]
PREWITT_Y = [-1, -1, -1, 0, 0, 0, 1, 1, 1]
LAPLACE4 = [0, 1, 0, 1, -4, 1, 0, 1, 0]
LAPLACE8 = [1, 1, 1, 1, -8, 1, 1, 1, 1]


@pytest.fixture(scope="module")
def dev():
    d = M.Device.default()
    d.set_option("tile_min_elems", 0)
    d.set_option("force_generic", 0)
    yield d
    d.set_option("tile_min_elems", 16384)


def _inputs(rng, dtype, nonfinite):
    out = []
    for shape in ((33, 130), (64, 128), (97, 260), (5, 9), (200, 1000)):
        a = (rng.standard_normal(shape) * 50).astype(dtype)
        if nonfinite:
            for _ in range(6):
                a[rng.integers(0, shape[0]), rng.integers(0, shape[1])] = rng.choice([np.inf, -np.inf, np.nan])
            a[0, 0] = -0.0
        out.append(a)
    return out


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("finite", [False, True], ids=["strict", "assume_finite"])
@pytest.mark.parametrize("name,k1,k2", [("sobel", SOBEL_X, SOBEL_Y), ("sobel-swapped", SOBEL_Y, SOBEL_X),
                                        ("prewitt", PREWITT_X, PREWITT_Y)])
@pytest.mark.parametrize("corr", [False, True], ids=["conv", "corr"])
def test_known_magnitude(name, k1, k2, corr, finite, dtype, dev):
    rng = np.random.default_rng(len(name) + 10 * corr + 100 * finite)
    flags = (_abi.FLAG_MAG2_CORR if corr else 0) | (_abi.FLAG_ASSUME_FINITE if finite else 0)
    for a in _inputs(rng, dtype, nonfinite=not finite):
        for border in BORDERS:
            kw = dict(size=(3, 3), coeffs=k1, coeffs2=k2, border=border, fill=0.5)
            want = O.apply("mag2", a, flags=O.FLAG_MAG2_CORR if corr else 0, **kw)
            got = product_apply("mag2", a, flags=flags, resident=True, **kw)
            assert dev.last_kernel == "tile2d_known", dev.last_kernel
            assert bits_equal(got, want), (name, corr, finite, a.shape, border)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("finite", [False, True], ids=["strict", "assume_finite"])
@pytest.mark.parametrize("name,k", [("sobelx", SOBEL_X), ("sobely", SOBEL_Y), ("laplace4", LAPLACE4), ("laplace8", LAPLACE8)])
@pytest.mark.parametrize("kind", ["conv", "corr"])
def test_known_linear(name, k, kind, finite, dtype, dev):
    rng = np.random.default_rng(len(name) + 7 * finite)
    flags = _abi.FLAG_ASSUME_FINITE if finite else 0
    for a in _inputs(rng, dtype, nonfinite=not finite):
        for border in BORDERS:
            kw = dict(size=(3, 3), coeffs=k, border=border, fill=-2.0)
            want = O.apply(kind, a, **kw)
            got = product_apply(kind, a, flags=flags, resident=True, **kw)
            assert dev.last_kernel == "tile2d_known", dev.last_kernel
            assert bits_equal(got, want), (name, kind, finite, a.shape, border)


def test_known_equals_generic_coefficient_path(dev):
    """the registry is an optimisation only: switching it off must not change a bit"""
    rng = np.random.default_rng(0)
    a = rng.standard_normal((300, 700)).astype(np.float32)
    kw = dict(size=(3, 3), coeffs=SOBEL_X, coeffs2=SOBEL_Y, border="reflect")
    fast = product_apply("mag2", a, resident=True, **kw)
    assert dev.last_kernel == "tile2d_known"
    dev.set_option("no_known", 1)
    try:
        slow = product_apply("mag2", a, resident=True, **kw)
        assert dev.last_kernel == "tile2d"
    finally:
        dev.set_option("no_known", 0)
    assert bits_equal(fast, slow)
    # a kernel that is almost Sobel is not Sobel
    k = list(SOBEL_X)
    k[4] = 0.25
    got = product_apply("mag2", a, resident=True, size=(3, 3), coeffs=k, coeffs2=SOBEL_Y, border="reflect")
    assert dev.last_kernel == "tile2d"
    assert bits_equal(got, O.apply("mag2", a, size=(3, 3), coeffs=k, coeffs2=SOBEL_Y, border="reflect"))


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_checked_optimistic_form_one_nonfinite_cell_anywhere(dtype, dev):
    """Without the finite promise the known kernels skip their zero taps and inspect their results; the strict launch
    that follows only works when something non-finite showed (k_tile2d_known.cu: known_optimistic_ok).  ONE Inf or NaN
    at every kind of position — corners, edges, interior — must give the strict result for every kernel, both
    directions and all borders, down to the smallest arrays the optimistic form accepts (and below, where the strict
    kernel runs alone)."""
    rng = np.random.default_rng(99)
    lin = [("conv", SOBEL_X), ("corr", SOBEL_X), ("conv", SOBEL_Y), ("corr", SOBEL_Y), ("conv", LAPLACE4), ("corr", PREWITT_X)]
    mag = [(SOBEL_X, SOBEL_Y, False), (SOBEL_Y, SOBEL_X, True), (PREWITT_X, PREWITT_Y, False)]
    for shape in ((3, 3), (3, 6), (7, 3), (40, 70), (2, 9), (9, 1)):
        h, w = shape
        spots = {(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (0, w // 2), (h - 1, w // 2), (h // 2, 0), (h // 2, w - 1), (h // 2, w // 2)}
        for (y, x) in sorted(spots):
            a = (rng.standard_normal(shape) * 3).astype(dtype)
            a[y, x] = rng.choice([np.inf, -np.inf, np.nan])
            for border in BORDERS:
                for kind, k in lin:
                    kw = dict(size=(3, 3), coeffs=k, border=border, fill=1.5)
                    assert bits_equal(product_apply(kind, a, resident=True, **kw), O.apply(kind, a, **kw)), (shape, y, x, border, kind, k)
                for k1, k2, corr in mag:
                    kw = dict(size=(3, 3), coeffs=k1, coeffs2=k2, border=border, fill=1.5)
                    want = O.apply("mag2", a, flags=O.FLAG_MAG2_CORR if corr else 0, **kw)
                    got = product_apply("mag2", a, flags=_abi.FLAG_MAG2_CORR if corr else 0, resident=True, **kw)
                    assert bits_equal(got, want), (shape, y, x, border, "mag", corr)
    # a non-finite Fill element reaches zero taps at the rim: strict kernel alone
    a = rng.standard_normal((20, 30)).astype(dtype)
    kw = dict(size=(3, 3), coeffs=SOBEL_X, coeffs2=SOBEL_Y, border="fill", fill=np.inf)
    assert bits_equal(product_apply("mag2", a, resident=True, **kw), O.apply("mag2", a, **kw))
