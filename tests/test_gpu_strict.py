"""Star-shaped 3x3x3 kernels WITHOUT the finite-input promise: the device runs the 7-tap kernels
optimistically, checks what it read, and redoes the work with the strict 27-tap kernel when it met Inf
or NaN (0 * Inf = NaN through the zero taps).  Results must equal the oracle's for every input."""
import numpy as np
import pytest

import massiv_b200 as M
from helpers import bits_equal
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = M.Device.default()
    d.set_option("tile_min_elems", 0)
    d.set_option("force_generic", 0)
    d.set_option("tb2", 1)
    yield d
    d.set_option("tile_min_elems", 16384)


def heat(dt=np.float64, r=0.1):
    k = np.zeros((3, 3, 3), dt)
    k[1, 1, 1] = 1 - 6 * r
    for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = r
    return k


SPOTS = [(0, 0, 0), (5, 31, 63), (5, 32, 64), (11, 39, 129), (6, 0, 64), (3, 30, 59), (3, 29, 60), (7, 17, 1)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("n_steps", [1, 2, 3, 4])
def test_iterate_strict(dev, dtype, n_steps):
    rng = np.random.default_rng(n_steps)
    k = heat(dtype)
    kw = dict(size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0)
    st = M.makeConvolutionStencilFromKernel(k)   # no assumeFinite()
    shape = (12, 40, 130) if dtype == np.float64 else (12, 40, 132)   # TMA needs 16-byte row pitches
    clean = rng.random(shape).astype(dtype)
    want = O.iterate("conv", clean, n_steps, **kw)
    got = M.iterateStencil(n_steps, M.Fill(0.0), st, dev.fromHost(clean)).toHost()
    assert dev.last_kernel == ("vol3d_tb2" if n_steps % 2 == 0 else "vol3d_star7")
    assert bits_equal(got, want)
    for spot in SPOTS:
        for bad in (np.inf, -np.inf, np.nan):
            a = clean.copy()
            a[spot] = bad
            want = O.iterate("conv", a, n_steps, **kw)
            got = M.iterateStencil(n_steps, M.Fill(0.0), st, dev.fromHost(a)).toHost()
            assert bits_equal(got, want), (spot, bad)
            assert bits_equal(M.iterateStencil(n_steps, M.Fill(0.0), st, a), want), (spot, bad)   # host -> host


def test_overflow_in_the_intermediate(dev):
    # finite input, but the first application overflows to Inf: the second one must see it through its zero taps
    k = heat(np.float64)
    k *= 4.0
    a = np.full((10, 36, 70), 1e307)
    a[5, 20, 33] = 1.7e308
    kw = dict(size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0)
    with np.errstate(all="ignore"):
        want = O.iterate("conv", a, 2, **kw)
    assert not np.isfinite(want).all()
    got = M.iterateStencil(2, M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), dev.fromHost(a)).toHost()
    assert dev.last_kernel == "vol3d_tb2" and bits_equal(got, want)


@pytest.mark.parametrize("border", ["fill", "wrap", "edge", "reflect", "continue"])
def test_single_pass_strict(dev, border):
    rng = np.random.default_rng(7)
    k = heat()
    a = rng.random((9, 40, 72))
    b = {"fill": M.Fill(0.0), "wrap": M.Wrap, "edge": M.Edge, "reflect": M.Reflect, "continue": M.Continue}[border]
    st = M.makeCorrelationStencilFromKernel(k)
    kw = dict(size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border=border, fill=0.0)
    assert bits_equal(M.compute(M.mapStencil(b, st, dev.fromHost(a))).toHost(), O.apply("corr", a, **kw))
    assert dev.last_kernel == "vol3d_star7"
    for spot in [(0, 0, 0), (8, 39, 71), (4, 31, 64), (0, 20, 0)]:
        a2 = a.copy()
        a2[spot] = np.inf
        assert bits_equal(M.compute(M.mapStencil(b, st, dev.fromHost(a2))).toHost(), O.apply("corr", a2, **kw)), spot


def test_cases_that_stay_strict(dev):
    rng = np.random.default_rng(8)
    k = heat()
    a = rng.random((9, 40, 72))
    st = M.makeConvolutionStencilFromKernel(k)
    # a non-finite Fill element reaches the zero taps at the border: the 27-tap kernel from the start
    got = M.compute(M.mapStencil(M.Fill(np.inf), st, dev.fromHost(a))).toHost()
    assert dev.last_kernel == "vol3d_dense27"
    assert bits_equal(got, O.apply("conv", a, size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=np.inf))
    # cells of the source that no output cell sits on (no padding): the check would not see them
    got = M.compute(M.applyStencil(M.noPadding, st, dev.fromHost(a))).toHost()
    assert dev.last_kernel != "vol3d_star7"
    assert bits_equal(got, O.apply("conv", a, size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), pad_lo=(0, 0, 0), pad_hi=(0, 0, 0)))


def test_chunked_and_ranged_launches_see_their_halo_planes():
    """mb_run splits large host->host passes into chunks of output planes: a non-finite value in the plane
    just outside a chunk must still be noticed by that chunk's launch"""
    dev = M.Device.default()
    rng = np.random.default_rng(21)
    k = heat()
    a = rng.random((96, 520, 530))   # ~212 MB in + out: two or more chunks
    kw = dict(size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0)
    st = M.makeConvolutionStencilFromKernel(k)
    for z in (47, 48, 49, 0, 95, 23, 24, 71, 72):
        a2 = a.copy()
        a2[z, 100, 200] = np.inf
        want = O.apply("conv", a2, nthreads=16, **kw)
        assert bits_equal(M.compute(M.mapStencil(M.Fill(0.0), st, a2)), want), z
