"""iterateUntil with a device-side convergence predicate (Array/Manifest/Internal.hs:331-349): the result, the number of
applications and the verdict must equal a step-by-step loop over the oracle with the same predicate in numpy."""
import numpy as np
import pytest

import massiv_b200 as M
from helpers import bits_equal
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def loop(kind, arr, pred, max_steps, check_every, **kw):
    cur, n = arr, 0
    while n < max_steps:
        chunk = min(check_every, max_steps - n)
        for _ in range(chunk):
            prev, cur = cur, O.apply(kind, cur, **kw)
        n += chunk
        if pred(prev, cur):
            return cur, n, True
    return cur, n, False


@pytest.mark.parametrize("check_every", [1, 2, 5])
def test_life_until_nothing_changes(check_every):
    dev = M.Device.default()
    g = np.zeros((64, 128), np.uint8)
    g[10:12, 10:12] = 1                      # a block: still life
    g[30, 40:43] = 1                         # a blinker: period 2, never equal to its predecessor
    g2 = g.copy()
    g2[30, 40:43] = 0
    g2[40, 60], g2[41, 61], g2[42, 59:62] = 1, 1, 1  # a glider that dies against nothing: keeps moving (torus)
    for board, cap in ((g, 12), (np.where(g2, 1, 0).astype(np.uint8), 9), (np.zeros((64, 128), np.uint8) + (np.arange(128) % 7 == 0), 30)):
        want, n, ok = loop("life", board, lambda p, c: np.array_equal(p, c), cap, check_every, size=(3, 3), border="wrap")
        got, steps, conv = M.iterateUntil("equal", M.Wrap, M.lifeStencil, dev.fromHost(board), cap, check_every)
        assert (steps, conv) == (n, ok)
        assert bits_equal(got.toHost(), want)
        got, steps, conv = M.iterateUntil("equal", M.Wrap, M.lifeStencil, board, cap, check_every)   # host -> host
        assert (steps, conv) == (n, ok) and bits_equal(got, want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("check_every", [1, 3, 4])
def test_heat_until_the_largest_change_is_small(dtype, check_every):
    dev = M.Device.default()
    dev.set_option("tile_min_elems", 0)
    rng = np.random.default_rng(1)
    k = np.zeros((3, 3, 3), dtype)
    k[1, 1, 1] = 0.4
    for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = 0.1
    arr = rng.random((20, 36, 72)).astype(dtype)
    tol = 0.02
    kw = dict(size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0)
    pred = lambda p, c: float(np.max(np.abs(c.astype(np.float64) - p.astype(np.float64)))) < tol
    want, n, ok = loop("conv", arr, pred, 40, check_every, **kw)
    got, steps, conv = M.iterateUntil(("max_abs_diff", tol), M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), dev.fromHost(arr), 40, check_every)
    assert (steps, conv) == (n, ok) and 1 < n < 40
    assert bits_equal(got.toHost(), want)
    # a NaN never converges; the cap ends the loop
    arr[3, 4, 5] = np.nan
    got, steps, conv = M.iterateUntil(("max_abs_diff", 1e9), M.Fill(0.0), M.makeConvolutionStencilFromKernel(k), dev.fromHost(arr), 6, check_every)
    assert (steps, conv) == (6, False)
    dev.set_option("tile_min_elems", 16384)


def test_predicate_needs_matching_element_type():
    with pytest.raises(M.MassivB200Error):
        M.iterateUntil(("max_abs_diff", 1.0), M.Wrap, M.lifeStencil, np.zeros((8, 8), np.uint8), 3)
