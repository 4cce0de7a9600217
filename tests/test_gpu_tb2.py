"""Temporal blocking (two steps of the 3-D star stencil per pass, k_vol3d_tb2.cu) must be bit-identical
to iterating the reference step by step: same per-cell operations, the border rule applied to the
intermediate array as well."""
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from oracle import oracle as O
from helpers import bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = M.Device.default()
    d.set_option("tile_min_elems", 0)
    d.set_option("force_generic", 0)
    d.set_option("tb2", 1)
    yield d
    d.set_option("tile_min_elems", 16384)


def star(rng, dtype, heat=True):
    k = np.zeros((3, 3, 3), dtype)
    if heat:
        k[1, 1, 1] = 0.4
        for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
            k[ix] = 0.1
    else:  # seven different coefficients: any mix-up of the tap order shows
        for ix in ((1, 1, 1), (0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
            k[ix] = rng.standard_normal() * 0.3
    return k


SHAPES = [(9, 33, 72), (40, 64, 128), (5, 20, 260), (70, 40, 64), (3, 3, 4), (34, 65, 132), (2, 31, 60), (130, 61, 124), (37, 45, 136),
          (9, 33, 70), (34, 65, 129)]  # the last two have pitches TMA cannot describe: single-step kernels


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind", ["conv", "corr"])
@pytest.mark.parametrize("n", [2, 3, 4, 7])
def test_tb2_matches_stepwise_oracle(n, kind, dtype, dev):
    rng = np.random.default_rng(zlib.crc32(f"{n}{kind}{np.dtype(dtype).name}".encode()))
    mk = M.makeConvolutionStencilFromKernel if kind == "conv" else M.makeCorrelationStencilFromKernel
    for shape in SHAPES:
        a = rng.standard_normal(shape).astype(dtype)
        k = star(rng, dtype, heat=False)
        want = O.iterate(kind, a, n, size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0, nthreads=8)
        got = M.iterateStencil(n, M.Fill(0.0), mk(k).assumeFinite(), dev.fromHost(a)).toHost()
        # an even count ends on a fused pass, an odd one on a single step
        if (shape[2] * np.dtype(dtype).itemsize) % 16 == 0:
            assert dev.last_kernel == ("vol3d_tb2" if n % 2 == 0 else "vol3d_star7"), (dev.last_kernel, shape, n)
        assert bits_equal(got, want), (shape, n, kind)


def test_tb2_equals_single_step_path(dev):
    rng = np.random.default_rng(1)
    a = rng.random((96, 200, 330))
    st = M.makeConvolutionStencilFromKernel(star(rng, np.float64)).assumeFinite()
    fused = M.iterateStencil(6, M.Fill(0.0), st, dev.fromHost(a)).toHost()
    assert dev.last_kernel == "vol3d_tb2"
    dev.set_option("tb2", 0)
    try:
        single = M.iterateStencil(6, M.Fill(0.0), st, dev.fromHost(a)).toHost()
        assert dev.last_kernel == "vol3d_star7"
    finally:
        dev.set_option("tb2", 1)
    assert bits_equal(fused, single)


def test_tb2_not_used_when_it_must_not_be(dev):
    rng = np.random.default_rng(2)
    a = rng.random((20, 40, 70))
    k = star(rng, np.float64)
    co = k.reshape(-1).tolist()
    # non-zero Fill, other borders: single-step kernels, still exact.  Without the finite promise the two-step
    # kernel runs in its checked form (tests/test_gpu_strict.py) unless that is switched off.
    for border, fill, st, opt in (("fill", 1.5, M.makeConvolutionStencilFromKernel(k).assumeFinite(), 0),
                                  ("edge", 0.0, M.makeConvolutionStencilFromKernel(k).assumeFinite(), 0),
                                  ("fill", 0.0, M.makeConvolutionStencilFromKernel(k), 1)):
        b = M.Fill(fill) if border == "fill" else M.Edge
        dev.set_option("no_optimistic", opt)
        try:
            got = M.iterateStencil(4, b, st, dev.fromHost(a)).toHost()
        finally:
            dev.set_option("no_optimistic", 0)
        assert dev.last_kernel != "vol3d_tb2"
        assert bits_equal(got, O.iterate("conv", a, 4, size=(3, 3, 3), coeffs=co, border=border, fill=fill))
