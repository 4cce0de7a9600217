"""Arrays TMA cannot describe (row pitch not a multiple of 16 bytes: odd row lengths) and the `no_tma` switch:
the tiled kernels stage their boxes with cp.async instead.  Results against the oracle, bit-exact."""
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
    d.set_option("no_tma", 0)


CASES2D = [("avg", (3, 3), "f64"), ("avg", (3, 3), "f32"), ("conv", (5, 5), "f32"), ("corr", (3, 3), "f64"), ("sum", (7, 7), "f64"),
           ("max", (3, 3), "i32"), ("mag2", (3, 3), "f32"), ("conv", (1, 5), "f64"), ("product", (2, 2), "f32")]


@pytest.mark.parametrize("kind,size,elem", CASES2D)
@pytest.mark.parametrize("no_tma", [0, 1])
def test_2d_odd_pitches(dev, kind, size, elem, no_tma):
    dev.set_option("no_tma", no_tma)
    dt = np.dtype(O.ELEM2NP[elem])
    rng = np.random.default_rng(zlib.crc32(repr((kind, size, elem)).encode()))
    n = int(np.prod(size))
    co = rng.standard_normal(n).astype(dt).tolist() if kind in ("conv", "corr", "mag2") else None
    co2 = rng.standard_normal(n).astype(dt).tolist() if kind == "mag2" else None
    for shape in ((131, 203), (200, 517), (97, 130), (260, 1001)):
        arr = (rng.standard_normal(shape) * (0.2 if kind == "product" else 1) + (1 if kind == "product" else 0)).astype(dt) \
            if dt.kind == "f" else rng.integers(-999, 999, size=shape).astype(dt)
        for border in BORDERS:
            want = O.apply(kind, arr, size=size, border=border, fill=2, coeffs=co, coeffs2=co2)
            st = M.Stencil(kind, size, None, None if co is None else np.asarray(co, dt).reshape(size),
                           None if co2 is None else np.asarray(co2, dt).reshape(size))
            got = M.compute(M.mapStencil(border_obj(border, 2), st, dev.fromHost(arr))).toHost()
            assert dev.last_kernel.startswith("tile2d"), dev.last_kernel
            assert bits_equal(got, want), (shape, border)
    dev.set_option("no_tma", 0)


@pytest.mark.parametrize("no_tma", [0, 1])
@pytest.mark.parametrize("elem", ["f64", "f32"])
def test_3d_odd_pitches(dev, elem, no_tma):
    dev.set_option("no_tma", no_tma)
    dt = np.dtype(O.ELEM2NP[elem])
    rng = np.random.default_rng(5 + no_tma)
    k = np.zeros((3, 3, 3), dt)
    k[1, 1, 1], k[0, 1, 1], k[2, 1, 1], k[1, 0, 1], k[1, 2, 1], k[1, 1, 0], k[1, 1, 2] = 0.4, .1, .1, .1, .1, .1, .1
    kr = rng.standard_normal((3, 3, 3)).astype(dt)
    for shape in ((20, 37, 71), (13, 66, 133), (40, 33, 65)):
        arr = rng.random(shape).astype(dt)
        for kk in (k, kr):
            co = kk.reshape(-1).tolist()
            for border in BORDERS:
                want = O.apply("conv", arr, size=(3, 3, 3), coeffs=co, border=border, fill=0.0)
                got = M.compute(M.mapStencil(border_obj(border, 0.0), M.makeConvolutionStencilFromKernel(kk), dev.fromHost(arr))).toHost()
                assert dev.last_kernel.startswith("vol3d"), dev.last_kernel
                assert bits_equal(got, want), (shape, border)
        for n in (2, 3, 5):
            want = O.iterate("conv", arr, n, size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0)
            got = M.iterateStencil(n, M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), dev.fromHost(arr)).toHost()
            assert bits_equal(got, want), (shape, n)
            a2 = arr.copy()
            a2[shape[0] // 2, 5, shape[2] - 1] = np.inf
            want = O.iterate("conv", a2, n, size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0)
            got = M.iterateStencil(n, M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), dev.fromHost(a2)).toHost()
            assert bits_equal(got, want), (shape, n, "inf")
    dev.set_option("no_tma", 0)


@pytest.mark.parametrize("no_tma", [0, 1])
def test_two_step_kernel_bulk_rows(dev, no_tma):
    """Double arrays without a tensor map: the two-step kernel's boxes inside the array come row by row through the
    bulk-copy engine, rows that start 8 bytes off a 16-byte boundary sit one element to the right in shared memory
    (k_vol3d_tb2.cu).  Odd and even H (the shifted rows alternate with z when H is odd), boxes on every rim."""
    dev.set_option("no_tma", no_tma)
    rng = np.random.default_rng(77 + no_tma)
    k = np.zeros((3, 3, 3))
    for ix in ((1, 1, 1), (0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = rng.standard_normal() * 0.3
    co = k.reshape(-1).tolist()
    try:
        for shape in ((14, 151, 203), (9, 104, 141), (11, 97, 140)):
            arr = rng.standard_normal(shape)
            for n in (2, 4):
                want = O.iterate("corr", arr, n, size=(3, 3, 3), coeffs=co, border="fill", fill=0.0, nthreads=16)
                got = M.iterateStencil(n, M.Fill(0.0), M.makeCorrelationStencilFromKernel(k), dev.fromHost(arr)).toHost()
                assert dev.last_kernel == "vol3d_tb2", dev.last_kernel
                assert bits_equal(got, want), (shape, n)
            a2 = arr.copy()
            a2[shape[0] // 2, 60, 77] = np.inf
            a2[2, 3, shape[2] - 1] = np.nan
            want = O.iterate("corr", a2, 2, size=(3, 3, 3), coeffs=co, border="fill", fill=0.0, nthreads=16)
            got = M.iterateStencil(2, M.Fill(0.0), M.makeCorrelationStencilFromKernel(k), dev.fromHost(a2)).toHost()
            assert bits_equal(got, want), (shape, "inf")
    finally:
        dev.set_option("no_tma", 0)


@pytest.mark.parametrize("elem", ["f64", "f32"])
def test_iteration_on_padded_rows(dev, elem):
    """An iteration of the two-step kernel's plans on an array TMA cannot describe runs on an internal pair of arrays
    with padded rows (mb_api.cu, iterate_pitch): from the device entry (copied in and out on the device) and from the
    host entry (the upload places the rows).  Same bits as stepping the oracle, with and without Inf / NaN; the
    `no_pitch` switch gives the in-place path."""
    dt = np.dtype(O.ELEM2NP[elem])
    rng = np.random.default_rng(123)
    k = np.zeros((3, 3, 3), dt)
    for ix in ((1, 1, 1), (0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = rng.standard_normal() * 0.3
    co = k.reshape(-1).tolist()
    st = M.makeConvolutionStencilFromKernel(k)
    if True:
        for shape in ((12, 70, 131), (9, 33, 67), (20, 20, 10)):
            arr = rng.standard_normal(shape).astype(dt)
            bad = arr.copy()
            bad[shape[0] // 2, 7, shape[2] - 1] = np.inf
            bad[1, 2, 3] = np.nan
            for n in (8, 9):
                for a in (arr, bad):
                    want = O.iterate("conv", a, n, size=(3, 3, 3), coeffs=co, border="fill", fill=0.0, nthreads=16)
                    got = M.iterateStencil(n, M.Fill(0.0), st, dev.fromHost(a)).toHost()
                    assert bits_equal(got, want), (shape, n, "device entry")
                    got = M.iterateStencil(n, M.Fill(0.0), st, a)
                    assert bits_equal(got, want), (shape, n, "host entry")
            dev.set_option("no_pitch", 1)
            try:
                got = M.iterateStencil(8, M.Fill(0.0), st, arr)
                want = O.iterate("conv", arr, 8, size=(3, 3, 3), coeffs=co, border="fill", fill=0.0, nthreads=16)
                assert bits_equal(got, want)
            finally:
                dev.set_option("no_pitch", 0)
        # iterateUntil on such an array: chunks of applications between the checks run padded, the checks compare dense arrays
        arr = rng.random((10, 40, 67)).astype(dt)
        heat = np.zeros((3, 3, 3), dt)
        heat[1, 1, 1] = 0.4
        for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
            heat[ix] = 0.1
        res, steps, conv = M.iterateUntil(("max_abs_diff", 1e-30), M.Fill(0.0), M.makeConvolutionStencilFromKernel(heat), dev.fromHost(arr),
                                          max_steps=20, check_every=10)
        want = O.iterate("conv", arr, 20, size=(3, 3, 3), coeffs=heat.reshape(-1).tolist(), border="fill", fill=0.0, nthreads=16)
        assert steps == 20 and not conv and bits_equal(res.toHost(), want)


@pytest.mark.parametrize("no_tma", [0, 1])
def test_separable_odd_pitches(dev, no_tma):
    dev.set_option("no_tma", no_tma)
    rng = np.random.default_rng(31)
    for dt, K in ((np.float32, 7), (np.float64, 5), (np.float32, 3)):
        g = rng.standard_normal(K).astype(dt)
        for shape in ((130, 203), (97, 517)):
            img = rng.random(shape).astype(dt)
            for border in BORDERS:
                t = O.apply("conv", img, size=(1, K), coeffs=g.tolist(), border=border, fill=0.5)
                want = O.apply("conv", t, size=(K, 1), coeffs=g.tolist(), border=border, fill=0.5)
                row, col = M.makeConvolutionStencilFromKernel(g.reshape(1, K)), M.makeConvolutionStencilFromKernel(g.reshape(K, 1))
                got = M.computeSeparable(border_obj(border, 0.5), row, col, dev.fromHost(img)).toHost()
                assert dev.last_kernel == "sep2d", dev.last_kernel
                assert bits_equal(got, want), (shape, border)
    dev.set_option("no_tma", 0)
