"""Post-maps (`fmap f` over the results, include/massiv_b200.h: mb_post) through the C ABI against the
oracle, behind every kernel family.  Bit-exact."""
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from helpers import bits_equal, border_obj
from oracle import oracle as O

pytestmark = pytest.mark.gpu

FLOAT_ONLY = {"div", "rdiv", "recip", "sqrt"}
NEEDS_C = {"add", "sub", "rsub", "mul", "min", "max", "div", "rdiv"}


@pytest.fixture(scope="module")
def dev():
    d = M.Device.default()
    d.set_option("tile_min_elems", 0)
    d.set_option("force_generic", 0)
    yield d
    d.set_option("tile_min_elems", 16384)


def special(dt, rng, shape):
    if dt.kind == "f":
        a = rng.standard_normal(shape).astype(dt)
        a.reshape(-1)[:6] = [0.0, -0.0, np.inf, -np.inf, np.nan, np.finfo(dt).tiny / 4]
        return a
    info = np.iinfo(dt)
    a = rng.integers(info.min, info.max, size=shape, dtype=dt, endpoint=True)
    a.reshape(-1)[:3] = [info.min, info.max, 0]
    return a


@pytest.mark.parametrize("elem", ["u8", "i8", "u16", "i16", "u32", "i32", "u64", "i64", "f32", "f64"])
@pytest.mark.parametrize("name", sorted(O.POSTS))
def test_each_op(dev, elem, name):
    dt = np.dtype(O.ELEM2NP[elem])
    rng = np.random.default_rng(zlib.crc32(repr((elem, name)).encode()))
    arr = special(dt, rng, (9, 11))
    c = dt.type(rng.integers(-3, 4) if dt.kind == "i" else (rng.integers(0, 5) if dt.kind == "u" else rng.standard_normal()))
    st = M.idStencil(2).fmap(name, c if name in NEEDS_C else None)
    if name in FLOAT_ONLY and dt.kind != "f":
        with pytest.raises(M.MassivB200Error):
            M.compute(M.mapStencil(M.Edge, st, arr))
        return
    want = O.apply("id", arr, size=(1, 1), border="edge", post=[(name, c)] if name in NEEDS_C else [(name,)])
    assert bits_equal(M.compute(M.mapStencil(M.Edge, st, arr)), want)
    assert bits_equal(M.compute(M.mapStencil(M.Edge, st, dev.fromHost(arr))).toHost(), want)


POST = [("mul", 0.5), ("abs",), ("add", -0.25), ("square",), ("sqrt",), ("min", 0.75), ("rsub", 1.0), ("max", 0.3)]


def chain(st):
    for item in POST:
        st = st.fmap(*item)
    return st


def test_behind_every_family(dev):
    rng = np.random.default_rng(4)
    seen = set()
    k3 = rng.standard_normal((3, 3))
    k333 = np.zeros((3, 3, 3))
    k333[1, 1, 1], k333[0, 1, 1], k333[2, 1, 1], k333[1, 0, 1], k333[1, 2, 1], k333[1, 1, 0], k333[1, 1, 2] = 0.4, .1, .1, .1, .1, .1, .1
    sob = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], float)
    cases = [
        ("conv", rng.standard_normal((70, 90)), dict(size=(3, 3), coeffs=k3.reshape(-1).tolist(), border="reflect"), M.makeConvolutionStencilFromKernel(k3), None),
        ("conv", rng.standard_normal((70, 90)), dict(size=(3, 3), coeffs=sob.reshape(-1).tolist(), border="edge"), M.makeConvolutionStencilFromKernel(sob), None),
        ("avg", rng.standard_normal((64, 80)), dict(size=(3, 3), border="edge"), M.avgStencil(M.Sz(3, 3)), None),
        # 49 taps: the compile-time-shaped kernel + the post-map pass (small windows take the fused small-window kernel)
        ("sum", rng.standard_normal((70, 90)), dict(size=(7, 7), border="reflect"), M.sumStencil(M.Sz(7, 7)), None),
        ("conv", rng.standard_normal((20, 36, 40)), dict(size=(3, 3, 3), coeffs=k333.reshape(-1).tolist(), border="fill", fill=0.0), M.makeConvolutionStencilFromKernel(k333).assumeFinite(), None),
        ("conv", rng.standard_normal((20, 36, 40)), dict(size=(3, 3, 3), coeffs=rng.standard_normal(27).tolist(), border="wrap"), None, None),
        ("sum", rng.standard_normal((40, 640)), dict(size=(1, 16), border="fill", fill=0.0, stride=(1, 16)), M.Stencil("sum", (1, 16)), (1, 16)),
        ("sum", rng.standard_normal((640, 40)), dict(size=(16, 1), border="fill", fill=0.0, stride=(16, 1)), M.Stencil("sum", (16, 1)), (16, 1)),
        ("product", rng.standard_normal((66, 99)), dict(size=(3, 3), border="continue", stride=(3, 3)), M.productStencil(M.Sz(3, 3)), (3, 3)),
        ("sum", rng.standard_normal((5, 6, 7, 8)), dict(size=(2, 1, 2, 2), border="edge"), M.sumStencil(M.Sz(2, 1, 2, 2)), None),
    ]
    for kind, arr, kw, st, stride in cases:
        if st is None:
            st = M.makeConvolutionStencilFromKernel(np.asarray(kw["coeffs"]).reshape(kw["size"]))
        want = O.apply(kind, arr, post=POST, **kw)
        b = border_obj(kw["border"], kw.get("fill", 0))
        got = M.computeWithStride(stride, M.mapStencil(b, chain(st), dev.fromHost(arr))).toHost()
        seen.add(dev.last_kernel)
        assert bits_equal(got, want), (kind, kw["size"], dev.last_kernel)
    assert {"tile2d", "smallwin", "vol3d_star7", "vol3d_dense27", "linescan_rows", "linescan_cols", "direct2d", "generic"} <= seen, seen


def test_iterate_and_separable(dev):
    rng = np.random.default_rng(5)
    k = np.zeros((3, 3, 3))
    k[1, 1, 1], k[0, 1, 1], k[2, 1, 1], k[1, 0, 1], k[1, 2, 1], k[1, 1, 0], k[1, 1, 2] = 0.4, .1, .1, .1, .1, .1, .1
    arr = rng.random((18, 34, 66))
    post = [("mul", 1.01), ("min", 0.9)]
    st = M.makeConvolutionStencilFromKernel(k).assumeFinite().fmap("mul", 1.01).fmap("min", 0.9)
    want = O.iterate("conv", arr, 5, size=(3, 3, 3), coeffs=k.reshape(-1).tolist(), border="fill", fill=0.0, post=post)
    got = M.iterateStencil(5, M.Fill(0.0), st, dev.fromHost(arr)).toHost()
    assert dev.last_kernel == "vol3d_star7"   # the two-step kernel does not take mapped stencils
    assert bits_equal(got, want)
    assert bits_equal(M.iterateStencil(5, M.Fill(0.0), st, arr), want)
    g = np.array([0.1, 0.2, 0.4, 0.2, 0.1], np.float32)
    img = rng.random((50, 70)).astype(np.float32)
    row = M.makeConvolutionStencilFromKernel(g.reshape(1, 5)).fmap("abs")
    col = M.makeConvolutionStencilFromKernel(g.reshape(5, 1)).fmap("rsub", np.float32(1))
    t = O.apply("conv", img, size=(1, 5), coeffs=g.tolist(), border="edge", post=[("abs",)])
    want = O.apply("conv", t, size=(5, 1), coeffs=g.tolist(), border="edge", post=[("rsub", np.float32(1))])
    assert bits_equal(M.computeSeparable(M.Edge, row, col, dev.fromHost(img)).toHost(), want)
