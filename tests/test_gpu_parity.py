"""GPU parity: libmassiv_b200.so (through its C ABI) against the CPU oracle and the reference's
golden vectors, bit-exact for integers AND floating point (the kernels reproduce the reference's
operation order with unfused IEEE operations, so the stated tolerance is 0 ulp; NaN payloads are
the one documented exception).

Every test here needs a B200 (`-m gpu`)."""
import json
import os
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from massiv_b200 import _abi
from oracle import oracle as O
from helpers import BORDERS, bits_equal, border_obj, oracle_apply, product_apply, splitmix64, stencil_obj

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "massiv_golden.json")))


@pytest.fixture(scope="module")
def dev():
    return M.Device.default()


@pytest.fixture(params=[0, 1], ids=["dispatch", "generic"])
def dev_mode(request, dev):
    """run every parity case twice: through the dispatcher (fast kernels where they apply) and with
    the generic kernel forced"""
    dev.set_option("force_generic", request.param)
    yield dev
    dev.set_option("force_generic", 0)


def _case_kw(c):
    return dict(size=c["size"], center=c["center"], pad_lo=c["pad_lo"], pad_hi=c["pad_hi"], border=c["border"],
                fill=c["fill"], stride=c["stride"], coeffs=c["coeffs"], coeffs2=c["coeffs2"], taps=c["taps"])


@pytest.mark.parametrize("resident", [False, True], ids=["host", "resident"])
@pytest.mark.parametrize("c", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_reference_golden_vectors(c, resident, dev_mode):
    arr = np.array(c["input"], dtype=O.ELEM2NP[c["elem"]]).reshape(c["in_dims"])
    out = product_apply(c["kind"], arr, resident=resident, **_case_kw(c))
    assert list(out.shape) == c["expect_dims"], c["src"]
    assert out.reshape(-1).tolist() == c["expect"], c["src"]


def _rand_arr(rng, dims, dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        return rng.standard_normal(dims).astype(dt)
    if dt.itemsize == 8:
        return rng.integers(-2**62, 2**62, dims).astype(dt)
    info = np.iinfo(dt)
    return rng.integers(info.min, int(info.max) + 1, dims).astype(dt)


def _rand_case(rng, rank, kind, dt, max_dim=9, max_sz=5):
    dt = np.dtype(dt)
    fold = kind in ("sum", "product", "avg", "max", "min")
    dims = tuple(int(rng.integers(0 if rng.random() < 0.08 else 1, max_dim + 1)) for _ in range(rank))
    size = tuple(int(rng.integers(0 if fold and rng.random() < 0.08 else 1, max_sz + 1)) for _ in range(rank))
    kw = dict(size=size, border=BORDERS[int(rng.integers(0, 5))])
    if kw["border"] == "fill":
        kw["fill"] = float(rng.standard_normal()) if dt.kind == "f" else int(rng.integers(0, 100))
    n = int(np.prod(size))
    if kind in ("conv", "corr", "mag2"):
        co = rng.standard_normal(n) if dt.kind == "f" else rng.integers(-5, 6, n)
        co[rng.random(n) < 0.25] = 0  # zero coefficients are NOT skipped by the reference
        kw["coeffs"] = co.astype(dt).tolist() if dt.kind == "f" else (co.astype(np.int64) % 7).tolist()
    if kind == "mag2":
        kw["coeffs2"] = rng.standard_normal(n).astype(dt).tolist()
    if kind == "taps":
        nt = int(rng.integers(0, 7))
        center = tuple(int(rng.integers(0, s)) for s in size)
        kw["center"] = center
        kw["taps"] = [[int(rng.integers(-c, s - c)) for s, c in zip(size, center)] for _ in range(nt)]
        kw["coeffs"] = (rng.standard_normal(nt).astype(dt) if dt.kind == "f" else rng.integers(0, 9, nt)).tolist()
    if rng.random() < 0.4:
        kw["pad_lo"] = tuple(int(rng.integers(0, 4)) for _ in range(rank))
        kw["pad_hi"] = tuple(int(rng.integers(0, 4)) for _ in range(rank))
    if rng.random() < 0.35:
        kw["stride"] = tuple(int(rng.integers(1, 4)) for _ in range(rank))
    return _rand_arr(rng, dims, dt), kw


KIND_TYPES = [("sum", np.int64), ("sum", np.float32), ("sum", np.float64), ("sum", np.uint8), ("sum", np.int16),
              ("product", np.int32), ("product", np.float64), ("product", np.uint64), ("avg", np.float64),
              ("avg", np.float32), ("max", np.int32), ("max", np.uint8), ("min", np.uint16), ("min", np.int64),
              ("max", np.int8), ("conv", np.float32), ("conv", np.float64), ("conv", np.int64), ("conv", np.uint16),
              ("corr", np.float64), ("corr", np.float32), ("corr", np.uint32), ("corr", np.int8), ("mag2", np.float32),
              ("mag2", np.float64), ("taps", np.float32), ("taps", np.int64), ("id", np.float64), ("id", np.uint8)]


@pytest.mark.parametrize("rank", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("kind,dtype", KIND_TYPES, ids=[f"{k}-{np.dtype(d).name}" for k, d in KIND_TYPES])
def test_random_small_cases_match_oracle(rank, kind, dtype, dev_mode):
    """the reference's QuickCheck shape: arrays and stencils of extent 0..9 per dim (SzTiny), all five
    borders, random padding and stride, ranks 1-5 (StencilSpec.hs:107-140, WindowedSpec.hs:14-37)."""
    rng = np.random.default_rng(zlib.crc32(f"{rank}-{kind}-{np.dtype(dtype).name}".encode()))
    trials = {1: 12, 2: 12, 3: 8, 4: 4, 5: 3}[rank]
    md = {1: 40, 2: 14, 3: 9, 4: 6, 5: 5}[rank]
    ms = {1: 7, 2: 5, 3: 4, 4: 3, 5: 2}[rank]
    for _ in range(trials):
        arr, kw = _rand_case(rng, rank, kind, dtype, md, ms)
        if kind == "id":
            kw["size"] = (1,) * rank
        try:
            want = oracle_apply(kind, arr, **kw)
        except O.OracleError as e:
            assert e.status == O.ERR_INDEX_ZERO
            with pytest.raises(M.IndexZeroException):
                product_apply(kind, arr, **kw)
            continue
        got = product_apply(kind, arr, **kw)
        assert bits_equal(got, want), (arr.shape, kw, got, want)


def test_bool32_max_min(dev_mode):
    rng = np.random.default_rng(4)
    a = (rng.integers(0, 3, (9, 11)) * 7).astype(np.int32)   # any non-zero word is True
    for kind in ("max", "min", "id"):
        size = (1, 1) if kind == "id" else (2, 3)
        for b in BORDERS:
            want = oracle_apply(kind, a, elem="bool32", size=size, border=b, fill=1)
            got = product_apply(kind, a, elem="bool32", size=size, border=b, fill=1)
            assert np.array_equal(got, want)
    # an empty window folds to minBound / maxBound = False / True
    assert (product_apply("max", a, elem="bool32", size=(0, 0), pad_lo=(0, 0), pad_hi=(0, 0), border="edge") == 0).all()
    assert (product_apply("min", a, elem="bool32", size=(0, 0), pad_lo=(0, 0), pad_hi=(0, 0), border="edge") == 1).all()


def test_non_finite_inputs_propagate_through_zero_coefficients(dev_mode):
    """zero coefficients are not skipped: 0 * Inf = NaN reaches the output."""
    a = np.ones((8, 9), np.float32)
    a[3, 4] = np.inf
    a[6, 1] = np.nan
    a[0, 0] = -np.inf
    k = np.array([[0, 1, 0], [1, -4, 1], [0, 1, 0]], np.float32)
    for kind in ("conv", "corr"):
        want = oracle_apply(kind, a, size=(3, 3), coeffs=k.reshape(-1).tolist(), border="edge")
        got = product_apply(kind, a, size=(3, 3), coeffs=k.reshape(-1).tolist(), border="edge")
        assert np.isnan(want[2, 3]) and bits_equal(got, want)
    # with the finite promise the zero taps are skipped: differs only where 0 * non-finite occurred
    got = product_apply("conv", a, size=(3, 3), coeffs=k.reshape(-1).tolist(), border="edge", flags=_abi.FLAG_ASSUME_FINITE)
    assert not np.isnan(got[2, 3])
    fin = np.nan_to_num(a, nan=2.0, posinf=3.0, neginf=-3.0)
    want = oracle_apply("conv", fin, size=(3, 3), coeffs=k.reshape(-1).tolist(), border="edge")
    got = product_apply("conv", fin, size=(3, 3), coeffs=k.reshape(-1).tolist(), border="edge", flags=_abi.FLAG_ASSUME_FINITE)
    assert bits_equal(got, want)


def test_signed_zero_and_denormals(dev_mode):
    a = np.array([[-0.0, 0.0, 1e-310, -1e-310], [5e-324, -0.0, -0.0, 2.5e-308]], np.float64)
    for kind, size in (("sum", (1, 2)), ("avg", (2, 2)), ("id", (1, 1)), ("product", (2, 1))):
        for b in ("edge", "wrap"):
            assert bits_equal(product_apply(kind, a, size=size, border=b), oracle_apply(kind, a, size=size, border=b))
    f = a.astype(np.float32) * np.float32(1e-30)
    f[0, 0] = -0.0
    k = [0.5, -0.25, 1.0, 0.0]
    assert bits_equal(product_apply("conv", f, size=(2, 2), coeffs=k, border="reflect"),
                      oracle_apply("conv", f, size=(2, 2), coeffs=k, border="reflect"))


def test_errors_map_to_reference_exceptions(dev):
    with pytest.raises(M.IndexZeroException):
        product_apply("sum", np.zeros((0, 3), np.int64), size=(2, 2), border="edge", pad_lo=(1, 0), pad_hi=(1, 1))
    out = product_apply("sum", np.zeros((0, 3), np.int64), size=(2, 2), border="fill", fill=5, pad_lo=(1, 0), pad_hi=(1, 1))
    assert out.shape == (1, 3) and (out == 20).all()
    assert product_apply("sum", np.zeros((0, 3), np.int64), size=(2, 2), border="edge").shape == (0, 3)
    with pytest.raises(M.SizeMismatchException):  # iterate needs a size-preserving stencil
        d = M.applyStencil(M.noPadding, M.sumStencil(M.Sz(3, 3)), np.zeros((6, 6))).desc()
        import ctypes as C
        a = dev.alloc((6, 6), np.float64)
        b = dev.alloc((6, 6), np.float64)
        res = C.c_void_p()
        dev.check(dev._L.mb_iterate_dev(dev.handle, C.byref(d), _abi.dims_array((6, 6)), C.c_void_p(a.ptr), C.c_void_p(b.ptr), 3, C.byref(res)))


# ---- iteration ------------------------------------------------------------------------------------

@pytest.mark.parametrize("shape", [(5, 5), (37, 53), (64, 64), (100, 260), (1, 7), (3, 1)])
@pytest.mark.parametrize("n", [1, 2, 7, 32, 65])
def test_life_iterate_matches_oracle(shape, n, dev_mode):
    rng = np.random.default_rng(n * 1000 + shape[0])
    a = (rng.random(shape) < 0.4).astype(np.uint8)
    want = O.iterate("life", a, n, size=(3, 3), border="wrap")
    got = M.iterateStencil(n, M.Wrap, M.lifeStencil, a)
    assert np.array_equal(got, want)
    d = M.Device.default().fromHost(a)
    got2 = M.iterateStencil(n, M.Wrap, M.lifeStencil, d).toHost()
    assert np.array_equal(got2, want)


def test_life_cells_other_than_0_1(dev_mode):
    """a centre of 2..255 always dies; neighbours still add modulo 256."""
    rng = np.random.default_rng(12)
    a = rng.integers(0, 256, (33, 47)).astype(np.uint8)
    a[rng.random(a.shape) < 0.5] = 1
    for n in (1, 2, 5):
        assert np.array_equal(M.iterateStencil(n, M.Wrap, M.lifeStencil, a), O.iterate("life", a, n, size=(3, 3), border="wrap"))
    # 255 + 1 + 1 + 1 + 1 = 259 = 3 (mod 256): a dead cell with those neighbours is born
    b = np.zeros((5, 5), np.uint8)
    b[1, 1], b[1, 2], b[1, 3], b[2, 1], b[2, 3] = 255, 1, 1, 1, 1
    assert np.array_equal(M.compute(M.mapStencil(M.Wrap, M.lifeStencil, b)), O.apply("life", b, size=(3, 3), border="wrap"))


def test_life_inf2_seed(dev_mode):
    """the example's inf2 pattern centred by initLife (GameOfLife.hs:38-45, 61-68)."""
    inf2 = np.array([[1, 1, 1, 0, 1], [1, 0, 0, 0, 0], [0, 0, 0, 1, 1], [0, 1, 1, 0, 1], [1, 0, 1, 0, 1]], np.uint8)
    g = np.zeros((96, 128), np.uint8)
    r0, c0 = (96 - 5) // 2, (128 - 5) // 2
    g[r0:r0 + 5, c0:c0 + 5] = inf2
    assert np.array_equal(M.iterateStencil(120, M.Wrap, M.lifeStencil, g), O.iterate("life", g, 120, size=(3, 3), border="wrap"))


def heat_kernel(r=0.1, dt=np.float64):
    k = np.zeros((3, 3, 3), dt)
    k[1, 1, 1] = 1 - 6 * r
    for ix in ((0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2)):
        k[ix] = r
    return k


@pytest.mark.parametrize("shape", [(8, 9, 10), (20, 33, 40), (5, 64, 128), (33, 32, 64), (1, 1, 5)])
@pytest.mark.parametrize("finite", [False, True], ids=["strict", "assume_finite"])
def test_heat3d_iterate_matches_oracle(shape, finite, dev_mode):
    rng = np.random.default_rng(shape[1])
    a = rng.random(shape)
    k = heat_kernel()
    st = M.makeConvolutionStencilFromKernel(k)
    if finite:
        st = st.assumeFinite()
    for n in (1, 4):
        want = O.iterate("conv", a, n, size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0)
        got = M.iterateStencil(n, M.Fill(0.0), st, a)
        assert bits_equal(got, want)


def test_iterate_does_not_mutate_its_argument(dev):
    a = (np.arange(64).reshape(8, 8) % 3 == 0).astype(np.uint8)
    d = dev.fromHost(a)
    M.iterateStencil(3, M.Wrap, M.lifeStencil, d)
    assert np.array_equal(d.toHost(), a)


# ---- separable two-pass ----------------------------------------------------------------------------

@pytest.mark.parametrize("border", BORDERS)
@pytest.mark.parametrize("shape", [(9, 13), (64, 128), (70, 300), (3, 5)])
def test_separable_two_pass_matches_oracle(border, shape, dev_mode):
    rng = np.random.default_rng(shape[1])
    a = rng.random(shape).astype(np.float32)
    g = np.exp(-0.5 * np.arange(-3, 4) ** 2).astype(np.float32)
    g /= g.sum()
    row, col = g.reshape(1, 7), g.reshape(7, 1)
    t = O.apply("conv", a, size=(1, 7), coeffs=g.tolist(), border=border, fill=0.5)
    want = O.apply("conv", t, size=(7, 1), coeffs=g.tolist(), border=border, fill=0.5)
    got = M.computeSeparable(border_obj(border, 0.5), M.makeConvolutionStencilFromKernel(row),
                             M.makeConvolutionStencilFromKernel(col), a)
    assert bits_equal(got, want)
