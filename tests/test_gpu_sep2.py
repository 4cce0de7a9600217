"""Fused separable kernel (k_sep2d.cu) against the reference's own idiom — two `compute`s with the
intermediate rounded to the element type and the border applied to it again — bit for bit."""
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from oracle import oracle as O
from helpers import BORDERS, bits_equal, border_obj

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = M.Device.default()
    d.set_option("tile_min_elems", 0)
    d.set_option("force_generic", 0)
    yield d
    d.set_option("tile_min_elems", 16384)


def two_pass_oracle(kind, a, row, col, border, fill):
    t = O.apply(kind, a, size=(1, len(row)), coeffs=list(row), border=border, fill=fill)
    return O.apply(kind, t, size=(len(col), 1), coeffs=list(col), border=border, fill=fill)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("k", [3, 5, 7])
@pytest.mark.parametrize("kind", ["conv", "corr"])
def test_fused_separable_matches_two_pass_oracle(kind, k, dtype, dev):
    rng = np.random.default_rng(zlib.crc32(f"{kind}{k}{np.dtype(dtype).name}".encode()))
    mk = M.makeConvolutionStencilFromKernel if kind == "conv" else M.makeCorrelationStencilFromKernel
    for shape in ((33, 130), (64, 128), (5, 7), (97, 260), (70, 131), (1, 300), (257, 64), (40, 1000), (3, 3)):
        a = rng.standard_normal(shape).astype(dtype)
        row, col = rng.standard_normal(k).astype(dtype), rng.standard_normal(k).astype(dtype)
        for border in BORDERS:
            fill = 0.0 if rng.random() < 0.5 else 1.25
            want = two_pass_oracle(kind, a, row.tolist(), col.tolist(), border, fill)
            got = M.computeSeparable(border_obj(border, fill), mk(row.reshape(1, k)), mk(col.reshape(k, 1)), dev.fromHost(a)).toHost()
            assert dev.last_kernel == "sep2d", (shape, border, dev.last_kernel)
            assert bits_equal(got, want), (kind, k, shape, border, fill)


def test_fused_separable_non_finite_and_large(dev):
    rng = np.random.default_rng(2)
    a = rng.random((900, 2100)).astype(np.float32)
    a[5, 7], a[100, 2099], a[899, 0] = np.inf, np.nan, -np.inf
    g = np.exp(-0.5 * np.arange(-3, 4) ** 2).astype(np.float32)
    g /= g.sum()
    for border in ("edge", "fill", "wrap"):
        want = two_pass_oracle("conv", a, g.tolist(), g.tolist(), border, 0.0)
        got = M.computeSeparable(border_obj(border, 0.0), M.makeConvolutionStencilFromKernel(g.reshape(1, 7)),
                                 M.makeConvolutionStencilFromKernel(g.reshape(7, 1)), dev.fromHost(a)).toHost()
        assert dev.last_kernel == "sep2d" and bits_equal(got, want), border


def test_unfusable_shapes_fall_back_to_two_passes(dev):
    """a 1x5 followed by a 3x1 (different lengths) is still correct: two launches through scratch"""
    rng = np.random.default_rng(3)
    a = rng.random((80, 200))
    r, c = rng.standard_normal(5), rng.standard_normal(3)
    want = two_pass_oracle("conv", a, r.tolist(), c.tolist(), "reflect", 0.0)
    got = M.computeSeparable(M.Reflect, M.makeConvolutionStencilFromKernel(r.reshape(1, 5)),
                             M.makeConvolutionStencilFromKernel(c.reshape(3, 1)), dev.fromHost(a)).toHost()
    assert dev.last_kernel == "tile2d" and bits_equal(got, want)
