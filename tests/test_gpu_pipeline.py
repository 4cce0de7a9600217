"""mb_run (host -> host) runs large single passes as a chunked H2D / kernel / D2H pipeline; the result
must not depend on it.  Arrays here are just large enough to be split."""
import numpy as np
import pytest

import massiv_b200 as M
from helpers import bits_equal, border_obj
from oracle import oracle as O

pytestmark = pytest.mark.gpu

BORDERS = ["fill", "wrap", "edge", "reflect", "continue"]


@pytest.mark.parametrize("border", BORDERS)
def test_conv2d_all_borders(border):
    dev = M.Device.default()
    rng = np.random.default_rng(11)
    arr = rng.standard_normal((4100, 4099))
    k = rng.standard_normal((5, 5))
    st = M.makeConvolutionStencilFromKernel(k)
    want = O.apply("conv", arr, size=(5, 5), coeffs=k.reshape(-1).tolist(), border=border, fill=0.5, nthreads=16)
    got = M.compute(M.mapStencil(border_obj(border, 0.5), st, arr))
    assert bits_equal(got, want)
    dev.set_option("pipeline", 0)
    try:
        assert bits_equal(M.compute(M.mapStencil(border_obj(border, 0.5), st, arr)), want)
    finally:
        dev.set_option("pipeline", 1)


def test_other_shapes():
    rng = np.random.default_rng(12)
    # 3-D star under Wrap (first and last chunks read the far end), asymmetric fold window, padding that grows the array
    k = np.zeros((3, 3, 3))
    k[1, 1, 1], k[0, 1, 1], k[2, 1, 1], k[1, 0, 1], k[1, 2, 1], k[1, 1, 0], k[1, 1, 2] = 0.4, .1, .1, .1, .1, .1, .1
    vol = rng.random((140, 300, 310))
    want = O.apply("conv", vol, size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="wrap", nthreads=16)
    assert bits_equal(M.compute(M.mapStencil(M.Wrap, M.makeConvolutionStencilFromKernel(k), vol)), want)
    img = rng.integers(0, 255, size=(9000, 9001), dtype=np.uint8)
    want = O.apply("max", img, size=(7, 3), border="continue", nthreads=16)
    assert bits_equal(M.compute(M.mapStencil(M.Continue, M.maxStencil(M.Sz(7, 3)), img)), want)
    arr = rng.standard_normal((4000, 4000))
    want = O.apply("sum", arr, size=(3, 3), pad_lo=(5, 2), pad_hi=(6, 1), border="reflect", nthreads=16)
    got = M.compute(M.applyStencil(M.Padding((5, 2), (6, 1), M.Reflect), M.sumStencil(M.Sz(3, 3)), arr))
    assert bits_equal(got, want)
    post = M.avgStencil(M.Sz(3, 3)).fmap("mul", 2.0).fmap("abs")
    want = O.apply("avg", arr, size=(3, 3), border="edge", nthreads=16, post=[("mul", 2.0), ("abs",)])
    assert bits_equal(M.compute(M.mapStencil(M.Edge, post, arr)), want)


@pytest.mark.parametrize("border", ["edge", "wrap", "fill"])
def test_separable_two_pass_host_to_host(border):
    """mb_run_sep2 pipelines the fused separable kernel over row chunks the same way; the result must be that of the two
    computes (intermediate rounded to Float) whatever the chunking."""
    dev = M.Device.default()
    rng = np.random.default_rng(13)
    img = rng.random((6200, 6100)).astype(np.float32)
    g = np.array([0.05, 0.1, 0.2, 0.3, 0.2, 0.1, 0.05], np.float32)
    row, col = M.makeConvolutionStencilFromKernel(g.reshape(1, 7)), M.makeConvolutionStencilFromKernel(g.reshape(7, 1))
    t = O.apply("conv", img, size=(1, 7), coeffs=g.tolist(), border=border, fill=0.25, nthreads=16)
    want = O.apply("conv", t, size=(7, 1), coeffs=g.tolist(), border=border, fill=0.25, nthreads=16)
    got = M.computeSeparable(border_obj(border, 0.25), row, col, img)
    assert dev.last_kernel == "sep2d" and bits_equal(got, want)
    dev.set_option("pipeline", 0)
    try:
        assert bits_equal(M.computeSeparable(border_obj(border, 0.25), row, col, img), want)
    finally:
        dev.set_option("pipeline", 1)


@pytest.mark.parametrize("dtname", ["float64", "float32"])
def test_iterate_host_to_host_wavefront(dtname):
    """mb_iterate on the two-step kernel's plans sends the input up in blocks and runs every pass on the part of the
    array whose dependency cone has arrived (mb_api.cu: iterate_wavefront); the result must be that of stepping the
    oracle — even and odd step counts, rows TMA cannot describe (padded on the device), and Inf / NaN in the input
    (the job is then redone strictly)."""
    dev = M.Device.default()
    dt = np.dtype(dtname)
    rng = np.random.default_rng(21)
    k = np.zeros((3, 3, 3), dt)
    for ix in ((1, 1, 1), (0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = rng.standard_normal() * 0.3
    co = k.reshape(-1).tolist()
    st = M.makeCorrelationStencilFromKernel(k)
    dev.set_option("wavefront_min_mb", 0)
    dev.set_option("tile_min_elems", 0)
    try:
        for shape in ((40, 70, 136), (67, 50, 131), (16, 40, 64)):
            arr = rng.standard_normal(shape).astype(dt)
            bad = arr.copy()
            bad[shape[0] - 3, 5, 7] = np.inf
            for n in (4, 7, 12, 30):
                want = O.iterate("corr", arr, n, size=(3, 3, 3), coeffs=co, border="fill", fill=0.0, nthreads=16)
                assert bits_equal(M.iterateStencil(n, M.Fill(0.0), st, arr), want), (shape, n)
                assert dev.last_kernel in ("vol3d_tb2", "vol3d_star7")
            want = O.iterate("corr", bad, 7, size=(3, 3, 3), coeffs=co, border="fill", fill=0.0, nthreads=16)
            assert bits_equal(M.iterateStencil(7, M.Fill(0.0), st, bad), want), (shape, "inf")
            # and with the caller's promise (no flag, no redo)
            want = O.iterate("corr", arr, 6, size=(3, 3, 3), coeffs=co, border="fill", fill=0.0, nthreads=16)
            assert bits_equal(M.iterateStencil(6, M.Fill(0.0), st.assumeFinite(), arr), want), (shape, "promise")
        dev.set_option("pipeline", 0)
        arr = rng.standard_normal((40, 70, 136)).astype(dt)
        want = O.iterate("corr", arr, 5, size=(3, 3, 3), coeffs=co, border="fill", fill=0.0, nthreads=16)
        assert bits_equal(M.iterateStencil(5, M.Fill(0.0), st, arr), want)
    finally:
        dev.set_option("pipeline", 1)
        dev.set_option("wavefront_min_mb", 256)
        dev.set_option("tile_min_elems", 16384)


def test_trim_returns_the_scratch_and_the_next_call_allocates_again():
    import torch
    dev = M.Device.default()
    rng = np.random.default_rng(31)
    arr = rng.standard_normal((3000, 3001))
    st = M.sumStencil(M.Sz(3, 3))
    want = O.apply("sum", arr, size=(3, 3), border="edge", nthreads=16)
    assert bits_equal(M.compute(M.mapStencil(M.Edge, st, arr)), want)   # host -> host: staging arrays in the context
    torch.cuda.synchronize()
    before = torch.cuda.mem_get_info()[0]
    dev.trim()
    after = torch.cuda.mem_get_info()[0]
    assert after - before >= 2 * arr.nbytes - (64 << 20), (before, after)
    assert bits_equal(M.compute(M.mapStencil(M.Edge, st, arr)), want)
