"""Post-maps (`fmap f` over a stencil's results, Stencil/Internal.hs:36-41, 70-83, 114-146): the C oracle
against the independent pure-Python restatement, and the descriptor plumbing.  CPU only."""
import ctypes as C
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from massiv_b200 import _abi
from oracle import oracle as O
from oracle import pyref as P

FLOAT_ONLY = {"div", "rdiv", "recip", "sqrt"}
NEEDS_C = {"add", "sub", "rsub", "mul", "min", "max", "div", "rdiv"}


def test_struct_layout():
    assert C.sizeof(_abi.StencilDesc) == 296 == C.sizeof(O._Desc)
    assert _abi.StencilDesc.n_post.offset == 260 and _abi.StencilDesc.post.offset == 264
    assert (_abi.StencilDesc.n_terms.offset, _abi.StencilDesc.n_program.offset, _abi.StencilDesc.terms.offset,
            _abi.StencilDesc.program.offset) == (272, 276, 280, 288)
    assert C.sizeof(_abi.Term) == 112 == C.sizeof(O._Term) and C.sizeof(_abi.XIns) == 16 == C.sizeof(O._XIns)
    assert C.sizeof(_abi.PostOp) == 16


def special(dt, rng, shape):
    if dt.kind == "f":
        a = rng.standard_normal(shape).astype(dt)
        flat = a.reshape(-1)
        flat[:6] = [0.0, -0.0, np.inf, -np.inf, np.nan, np.finfo(dt).tiny / 4]
        return a
    info = np.iinfo(dt)
    a = rng.integers(info.min, info.max, size=shape, dtype=dt, endpoint=True)
    a.reshape(-1)[:3] = [info.min, info.max, 0]
    return a


@pytest.mark.parametrize("elem", ["u8", "i8", "i16", "u32", "i64", "f32", "f64"])
@pytest.mark.parametrize("name", sorted(O.POSTS))
def test_oracle_equals_pyref(elem, name):
    dt = np.dtype(O.ELEM2NP[elem])
    if name in FLOAT_ONLY and dt.kind != "f":
        with pytest.raises(O.OracleError):
            O.apply("id", np.zeros((3, 3), dt), size=(1, 1), post=[(name, 1)])
        return
    rng = np.random.default_rng(zlib.crc32(repr((elem, name)).encode()))
    arr = special(dt, rng, (5, 7))
    c = dt.type(rng.integers(-3, 4) if dt.kind == "i" else (rng.integers(0, 5) if dt.kind == "u" else rng.standard_normal()))
    post = [(name, c)] if name in NEEDS_C else [(name,)]
    with np.errstate(all="ignore"):
        got = O.apply("id", arr, size=(1, 1), post=post)
        want = P.map_stencil(("edge",), P.with_post(P.id_stencil(2), post, dt), arr)
        ok = np.array_equal(got.view(np.uint8), np.asarray(want, dt).view(np.uint8)) if dt.kind != "f" else \
            np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)].view(np.uint8), want[~np.isnan(want)].view(np.uint8))
    assert ok, (got, want)


def test_chain_order_and_stencil():
    rng = np.random.default_rng(9)
    arr = rng.standard_normal((6, 8))
    k = rng.standard_normal((3, 3))
    post = [("mul", 0.5), ("abs",), ("add", -0.25), ("square",), ("sqrt",), ("min", 0.75), ("rsub", 1.0), ("max", 0.3)]
    got = O.apply("conv", arr, size=(3, 3), coeffs=k.reshape(-1).tolist(), border="reflect", post=post)
    dt = arr.dtype
    want = P.map_stencil(("reflect",), P.with_post(P.conv_from_kernel(k), post, dt), arr)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    with pytest.raises(O.OracleError):
        O.apply("id", arr, size=(1, 1), post=[("abs",)] * 9)


def test_host_mirror_builds_the_chain():
    k = np.ones((3, 3))
    st = 1.0 - abs(M.makeConvolutionStencilFromKernel(k) * 0.5) / 4.0
    assert st.post == (("mul", 0.5), ("abs",), ("div", 4.0), ("rsub", 1.0))
    d = M.mapStencil(M.Edge, st, np.zeros((4, 4))).desc()
    assert d.n_post == 4
    ops = C.cast(d.post, C.POINTER(_abi.PostOp))
    assert [ops[i].op for i in range(4)] == [_abi.POSTS[n] for n in ("mul", "abs", "div", "rsub")]
    assert M.sqrt(M.sumStencil(M.Sz(3, 3))).post == (("sqrt",),)
    sq = M.makeCorrelationStencilFromKernel(k) ** 2
    assert M.mapStencil(M.Edge, sq, np.zeros((4, 4))).stencil.post == (("square",),)
    assert (-M.sumStencil(3)).fmap("max", 0).post == (("negate",), ("max", 0))
    with pytest.raises(NotImplementedError):
        M.sumStencil(3).fmap("exp")
    dw = M.mapStencil(M.Edge, M.avgStencil(M.Sz(3, 3)), np.zeros((4, 4))).fmap("mul", 2.0)
    assert dw.stencil.post == (("mul", 2.0),)
