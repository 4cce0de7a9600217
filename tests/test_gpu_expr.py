"""Stencil expressions (MB_EXPR: the Applicative / Num / Fractional / Floating instances of Stencil,
Stencil/Internal.hs:83-174) on the device, against the C oracle on the same descriptor: random expression trees over
every constructor, all borders, element types and ranks 1-3 through the generic kernel, and 2-D shapes large enough
for the tiled small-window kernel; the closure-form Sobel operator of the reference's benchmark."""
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from helpers import bits_equal
from test_expr_cpu import _random_tree, oracle_of_dw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = M.Device.default()
    yield d
    d.set_option("force_generic", 0)
    d.set_option("tile_min_elems", 16384)


@pytest.mark.parametrize("dtname", ["int64", "int32", "int16", "uint16", "int8", "uint8", "float64", "float32"])
@pytest.mark.parametrize("rank", [1, 2, 3])
def test_random_expressions_match_the_oracle(dev, dtname, rank):
    dt = np.dtype(dtname)
    rng = np.random.default_rng(zlib.crc32(f"gpu{dtname}{rank}".encode()))
    done = 0
    for trial in range(30):
        ps, ms = _random_tree(rng, dt, rank)
        shape = tuple(int(rng.integers(1, 40)) for _ in range(rank))
        arr = (rng.standard_normal(shape) if dt.kind == "f" else rng.integers(-9 if dt.kind == "i" else 0, 10, shape)).astype(dt)
        bname = ["fill", "wrap", "edge", "reflect", "continue"][trial % 5]
        border = M.Fill(dt.type(3)) if bname == "fill" else {"wrap": M.Wrap, "edge": M.Edge, "reflect": M.Reflect, "continue": M.Continue}[bname]
        try:
            want, _ = oracle_of_dw(M.mapStencil(border, ms, arr), arr)
        except NotImplementedError:
            continue
        for generic in (1, 0):
            dev.set_option("force_generic", generic)
            dev.set_option("tile_min_elems", 0)
            got = M.compute(M.mapStencil(border, ms, dev.fromHost(arr))).toHost()
            assert bits_equal(got, want), (trial, bname, generic, dev.last_kernel)
        done += 1
    dev.set_option("force_generic", 0)
    assert done >= 20


@pytest.mark.parametrize("dtname", ["float32", "float64", "int32", "uint8", "int16"])
def test_tiled_small_window_kernel_on_larger_2d_arrays(dev, dtname):
    """shapes with partial tiles, odd pitches, windows up to 9x9, host->host and strided evaluation"""
    dt = np.dtype(dtname)
    rng = np.random.default_rng(zlib.crc32(dtname.encode()))
    dev.set_option("force_generic", 0)
    dev.set_option("tile_min_elems", 0)
    for shape in ((70, 300), (131, 517), (33, 1030)):
        arr = (rng.standard_normal(shape) if dt.kind == "f" else rng.integers(-9 if dt.kind == "i" else 0, 10, shape)).astype(dt)
        k9 = (rng.standard_normal((9, 9)) if dt.kind == "f" else rng.integers(0, 3, (9, 9))).astype(dt)
        k35 = (rng.standard_normal((3, 5)) if dt.kind == "f" else rng.integers(0, 3, (3, 5))).astype(dt)
        cases = [M.makeConvolutionStencilFromKernel(k9), M.makeCorrelationStencilFromKernel(k35), M.sumStencil(M.Sz(4, 6)),
                 M.makeConvolutionStencilFromKernel(k35) - M.sumStencil(M.Sz(2, 2)) * 2,
                 abs(M.makeCorrelationStencilFromKernel(k35) - M.makeConvolutionStencilFromKernel(k35.T.copy()))]
        if dt.kind != "f":
            cases.append(M.maxStencils(M.maxStencil(M.Sz(3, 3)), M.sumStencil(M.Sz(1, 2))))
        for i, st in enumerate(cases):
            for border in (M.Fill(dt.type(1)), M.Wrap, M.Reflect):
                want, _ = oracle_of_dw(M.mapStencil(border, st, arr), arr)
                got = M.compute(M.mapStencil(border, st, dev.fromHost(arr))).toHost()
                assert bits_equal(got, want), (shape, i, border, dev.last_kernel)
                assert dev.last_kernel == "smallwin", dev.last_kernel
                got = M.compute(M.mapStencil(border, st, arr))   # host -> host (pipelined for large arrays)
                assert bits_equal(got, want), (shape, i, border, "host")


def test_closure_form_sobel_of_the_reference_benchmark(dev):
    """massiv-bench/src/Data/Massiv/Bench/Sobel.hs:15-44"""
    sx = M.makeConvolutionStencil(M.Sz(3, 3), (1, 1), [((-1, -1), -1), ((0, -1), -2), ((1, -1), -1), ((-1, 1), 1), ((0, 1), 2), ((1, 1), 1)])
    sy = M.makeConvolutionStencil(M.Sz(3, 3), (1, 1), [((-1, -1), -1), ((-1, 0), -2), ((-1, 1), -1), ((1, -1), 1), ((1, 0), 2), ((1, 1), 1)])
    op = M.sqrt(sx * sx + sy * sy)
    rng = np.random.default_rng(4)
    dev.set_option("force_generic", 0)
    dev.set_option("tile_min_elems", 16384)
    for dt, shape in ((np.float64, (1200, 1600)), (np.float32, (1001, 2003))):
        a = rng.random(shape).astype(dt)
        want, _ = oracle_of_dw(M.mapStencil(M.Edge, op, a), a)
        got = M.compute(M.mapStencil(M.Edge, op, dev.fromHost(a))).toHost()
        assert bits_equal(got, want)
        assert dev.last_kernel == "smallwin", dev.last_kernel


@pytest.mark.parametrize("dtname", ["float32", "float64", "uint8", "int64"])
def test_small_window_kernel_rank3(dev, dtname):
    """3-D folds, windows other than 3x3x3, 3-D expressions: a tile is one output plane whose box carries the window's depth"""
    dt = np.dtype(dtname)
    rng = np.random.default_rng(zlib.crc32(("r3" + dtname).encode()))
    dev.set_option("force_generic", 0)
    dev.set_option("tile_min_elems", 0)
    for shape in ((20, 40, 300), (7, 70, 131)):
        arr = (rng.standard_normal(shape) if dt.kind == "f" else rng.integers(-9 if dt.kind == "i" else 0, 10, shape)).astype(dt)
        k234 = (rng.standard_normal((2, 3, 4)) if dt.kind == "f" else rng.integers(0, 3, (2, 3, 4))).astype(dt)
        k511 = (rng.standard_normal((5, 1, 1)) if dt.kind == "f" else rng.integers(0, 3, (5, 1, 1))).astype(dt)
        cases = [M.sumStencil(M.Sz(3, 3, 3)), M.makeCorrelationStencilFromKernel(k234), M.makeConvolutionStencilFromKernel(k511),
                 M.productStencil(M.Sz(2, 2, 2)), M.sumStencil(M.Sz(3, 3, 3)) - M.makeConvolutionStencilFromKernel(k234) * 2,
                 abs(M.makeCorrelationStencilFromKernel(k234) - M.sumStencil(M.Sz(1, 2, 2)))]
        if dt.kind == "f":
            cases.append(M.avgStencil(M.Sz(3, 3, 3)))
        else:
            cases.append(M.maxStencil(M.Sz(3, 2, 3)))
        for i, st in enumerate(cases):
            for border in (M.Fill(dt.type(1)), M.Fill(dt.type(0)), M.Wrap, M.Continue):
                want, _ = oracle_of_dw(M.mapStencil(border, st, arr), arr)
                got = M.compute(M.mapStencil(border, st, dev.fromHost(arr))).toHost()
                assert bits_equal(got, want), (shape, i, border, dev.last_kernel)
                # (a 5-plane box of 8-byte elements does not fit two stages of shared memory: the generic kernel takes it)
                assert dev.last_kernel == "smallwin" or (dt.itemsize == 8 and i == 2), (dev.last_kernel, shape, i)
        # iterated: a 3-D fold step
        want = arr
        for _ in range(3):
            want, _ = oracle_of_dw(M.mapStencil(M.Edge, cases[0] if dt.kind != "f" else cases[-1], want), want)
        got = M.iterateStencil(3, M.Edge, cases[0] if dt.kind != "f" else cases[-1], dev.fromHost(arr)).toHost()
        assert bits_equal(got, want)


@pytest.mark.parametrize("dtname", ["float64", "int64", "float32", "uint8", "int16"])
def test_dense_windows_share_their_loads_between_output_rows(dev, dtname):
    """Dense windows of 8-byte elements, and tall sums / products of any type, take the small-window kernel's dense form
    (k_smallwin.cu: sw_dense): box rows are visited once and every element is applied to all the thread's output rows
    it belongs to — walked top-down for folds / correlations, bottom-up for convolutions (whose taps run in reverse).
    The bits must be those of the tap-list form and of the oracle."""
    dt = np.dtype(dtname)
    rng = np.random.default_rng(zlib.crc32(("dense" + dtname).encode()))
    dev.set_option("force_generic", 0)
    dev.set_option("tile_min_elems", 0)
    shape = (150, 333)
    arr = (rng.standard_normal(shape) if dt.kind == "f" else rng.integers(-9 if dt.kind == "i" else 0, 10, shape)).astype(dt)

    def kern(h, w):
        return (rng.standard_normal((h, w)) if dt.kind == "f" else rng.integers(-2 if dt.kind == "i" else 0, 3, (h, w))).astype(dt)

    cases = [M.sumStencil(M.Sz(9, 9)), M.sumStencil(M.Sz(7, 2)), M.productStencil(M.Sz(8, 3)), M.makeConvolutionStencilFromKernel(kern(9, 4)),
             M.makeCorrelationStencilFromKernel(kern(4, 6)), M.makeConvolutionStencilFromKernel(kern(3, 16)), M.sumStencil(M.Sz(16, 2))]
    if dt.kind == "f":
        cases.append(M.avgStencil(M.Sz(11, 11)))
    else:
        cases += [M.maxStencil(M.Sz(3, 5)), M.minStencil(M.Sz(10, 2))]
    try:
        for i, st in enumerate(cases):
            for border in (M.Edge, M.Fill(dt.type(1)), M.Wrap, M.Reflect):
                want, _ = oracle_of_dw(M.mapStencil(border, st, arr), arr)
                got = M.compute(M.mapStencil(border, st, dev.fromHost(arr))).toHost()
                assert dev.last_kernel == "smallwin", (dev.last_kernel, i)
                assert bits_equal(got, want), (i, border)
    finally:
        dev.set_option("tile_min_elems", 16384)
