"""Stencil expressions — the Applicative / Num / Fractional / Floating instances of Stencil
(/root/reference/massiv/src/Data/Massiv/Array/Stencil/Internal.hs:83-174) — on the CPU side: the C oracle against
the closure-style Python restatement (oracle/pyref.py: lift2 / fmap, written like the Haskell), hand-derived Int
vectors, the size / centre union, and the host mirror's lowering to MB_EXPR descriptors (no GPU needed).

No reference test exercises these instances on arrays; the vectors below are derived by hand from the instance
definitions (Int arithmetic is exact), e.g. for ys = [1..5] under Fill 0:
    sumStencil (Sz 3)                  -> window i..i+2          [6, 9, 12, 9, 5]
    makeCorrelationStencilFromKernel [1,2,3] (centre 1)          [8, 14, 20, 26, 14]   (StencilSpec.hs:206)
    their liftA2 (+): centre max(0,1) = 1, size 1 + max(3-0, 3-1) = 4; same padding (1, 2); output cell o is
    evaluated at o + (1 - 1) = o, where BOTH stencils are evaluated:       [14, 23, 32, 35, 19]
"""
import ctypes as C
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from massiv_b200 import _abi
from oracle import oracle as O
from oracle import pyref as P


def test_hand_derived_int_vectors():
    ys = np.arange(1, 6, dtype=np.int64)
    terms = [dict(kind="sum", size=(3,)), dict(kind="corr", size=(3,), coeffs=[1, 2, 3])]
    assert O.expr_geom(terms, 1) == ((4,), (1,))
    got = O.apply("expr", ys, terms=terms, program=[("term", 0), ("term", 1), ("add",)], border="fill", fill=0)
    assert got.tolist() == [14, 23, 32, 35, 19]
    # liftA2 (*) and (-); fromInteger 10 - s == liftA2 (-) (pure 10) s
    got = O.apply("expr", ys, terms=terms, program=[("term", 0), ("term", 1), ("mul",)], border="fill", fill=0)
    assert got.tolist() == [6 * 8, 9 * 14, 12 * 20, 9 * 26, 5 * 14]
    got = O.apply("expr", ys, terms=terms[:1], program=[("const", 10), ("term", 0), ("sub",)], border="fill", fill=0)
    assert got.tolist() == [4, 1, -2, 1, 5]
    # abs (s1 - s2), signum, negate, (^2) on Int; liftA2 max / min
    got = O.apply("expr", ys, terms=terms, program=[("term", 0), ("term", 1), ("sub",), ("abs",)], border="fill", fill=0)
    assert got.tolist() == [2, 5, 8, 17, 9]
    got = O.apply("expr", ys, terms=terms, program=[("term", 0), ("term", 1), ("max",), ("term", 0), ("term", 1), ("min",), ("sub",)],
                  border="fill", fill=0)
    assert got.tolist() == [2, 5, 8, 17, 9]
    # 2-D: the centre of the union moves the first stencil's window: sum (Sz2 2 2) (centre 0) + conv 3x3 identity-like
    a = np.arange(1, 13, dtype=np.int64).reshape(3, 4)
    k = np.zeros((3, 3), np.int64)
    k[1, 1] = 1   # conv with a centred delta == idStencil shifted to centre (1, 1)
    terms = [dict(kind="sum", size=(2, 2)), dict(kind="conv", size=(3, 3), coeffs=k)]
    assert O.expr_geom(terms, 2) == ((3, 3), (1, 1))
    got = O.apply("expr", a, terms=terms, program=[("term", 0), ("term", 1), ("add",)], border="fill", fill=0)
    s22 = O.apply("sum", a, size=(2, 2), border="fill", fill=0)   # same evaluation index (centre 0, pad (0,0),(1,1))
    assert np.array_equal(got, s22 + a)


def _random_tree(rng, dt, rank, depth=0):
    """(pyref stencil, host-mirror stencil) of one random expression"""
    isf = dt.kind == "f"
    t = dt.type

    def wrap(v):
        if isf:
            return t(v)
        bits = 8 * dt.itemsize
        v = int(v) & ((1 << bits) - 1)
        if dt.kind == "i" and v >= 1 << (bits - 1):
            v -= 1 << bits
        return t(v)

    ex = (lambda x: x) if isf else int
    leaf = depth >= 2 or rng.random() < 0.3
    if leaf:
        kind = rng.choice(["sum", "product", "conv", "corr", "taps", "id"] + ([] if isf else ["max", "min"]) + (["avg"] if isf else []))
        size = tuple(int(rng.integers(1, 4)) for _ in range(rank))
        if kind == "id":
            return P.id_stencil(rank), M.idStencil(rank)
        if kind in ("sum", "product", "max", "min", "avg"):
            ps = {"sum": P.sum_stencil, "product": P.product_stencil, "max": P.max_stencil, "min": P.min_stencil, "avg": P.avg_stencil}[kind](size, dt)
            ms = {"sum": M.sumStencil, "product": M.productStencil, "max": M.maxStencil, "min": M.minStencil, "avg": M.avgStencil}[kind](M.Sz(size))
            return ps, ms
        if kind in ("conv", "corr"):
            k = (rng.standard_normal(size) if isf else rng.integers(-3 if dt.kind == 'i' else 0, 4, size)).astype(dt)
            return ((P.conv_from_kernel(k), M.makeConvolutionStencilFromKernel(k)) if kind == "conv"
                    else (P.corr_from_kernel(k), M.makeCorrelationStencilFromKernel(k)))
        center = tuple(int(rng.integers(0, s)) for s in size)
        n = int(rng.integers(1, 4))
        taps = []
        for _ in range(n):
            off = tuple(int(rng.integers(-c, s - c)) for c, s in zip(center, size))
            taps.append((off, (rng.standard_normal() if isf else int(rng.integers(-3 if dt.kind == 'i' else 0, 4)))))
        taps = [(o, dt.type(k)) for o, k in taps]
        return P.corr_closure(size, center, taps, dt), M.makeCorrelationStencil(M.Sz(size), center, taps)
    op = rng.choice(["add", "sub", "mul", "negate", "abs", "square", "addc", "rsubc", "min", "max"] + (["div", "sqrtabs", "recip"] if isf else []))
    pa, ma = _random_tree(rng, dt, rank, depth + 1)
    if op in ("add", "sub", "mul", "div", "min", "max"):
        pb, mb = _random_tree(rng, dt, rank, depth + 1)
        f = {"add": lambda x, y: wrap(ex(x) + ex(y)), "sub": lambda x, y: wrap(ex(x) - ex(y)), "mul": lambda x, y: wrap(ex(x) * ex(y)),
             "div": lambda x, y: t(x / y), "min": lambda x, y: x if x <= y else y, "max": lambda x, y: y if x <= y else x}[op]
        mm = {"add": lambda: ma + mb, "sub": lambda: ma - mb, "mul": lambda: ma * mb, "div": lambda: ma / mb,
              "min": lambda: M.minStencils(ma, mb), "max": lambda: M.maxStencils(ma, mb)}[op]()
        return P.lift2(f, pa, pb), mm
    c = t(rng.standard_normal()) if isf else t(int(rng.integers(-5 if dt.kind == 'i' else 0, 6)))
    if op == "addc":
        return P.fmap(P.post_fn("add", c, dt), pa), ma + c
    if op == "rsubc":
        return P.fmap(P.post_fn("rsub", c, dt), pa), c - ma
    if op == "sqrtabs":
        return P.fmap(P.post_fn("sqrt", None, dt), P.fmap(P.post_fn("abs", None, dt), pa)), M.sqrt(abs(ma))
    if op == "recip":
        return P.fmap(P.post_fn("recip", None, dt), pa), ma.recip()
    return P.fmap(P.post_fn(op, None, dt), pa), {"negate": lambda: -ma, "abs": lambda: abs(ma), "square": lambda: ma ** 2}[op]()


def oracle_of_dw(dw, arr, stride=None):
    """run the host mirror's descriptor through the C oracle (the structures share their layout)"""
    d = dw.desc(stride)
    od = O._Desc.from_buffer_copy(bytes(d))
    dims = O._dims(arr.shape)
    out_dims = (C.c_int64 * O.MAX_RANK)()
    st = O.lib().mo_out_dims(C.byref(od), dims, out_dims)
    assert st == 0, st
    out = np.empty(tuple(out_dims[i] for i in range(arr.ndim)), arr.dtype)
    st = O.lib().mo_apply(C.byref(od), dims, np.ascontiguousarray(arr).ctypes.data, out.ctypes.data, 1)
    assert st == 0, st
    return out, d


@pytest.mark.parametrize("dtname", ["int64", "int16", "uint8", "float64", "float32"])
@pytest.mark.parametrize("rank", [1, 2, 3])
def test_oracle_matches_closure_restatement_on_random_expressions(dtname, rank):
    """lowering (host mirror) + C oracle == closures composed the way the Haskell instances compose them"""
    dt = np.dtype(dtname)
    rng = np.random.default_rng(zlib.crc32(f"{dtname}{rank}".encode()))
    done = 0
    for trial in range(40):
        ps, ms = _random_tree(rng, dt, rank)
        shape = tuple(int(rng.integers(1, 6)) for _ in range(rank))
        arr = (rng.standard_normal(shape) if dt.kind == "f" else rng.integers(-9 if dt.kind == "i" else 0, 10, shape)).astype(dt)
        bname = ["fill", "wrap", "edge", "reflect", "continue"][trial % 5]
        border = M.Fill(dt.type(3)) if bname == "fill" else {"wrap": M.Wrap, "edge": M.Edge, "reflect": M.Reflect, "continue": M.Continue}[bname]
        try:
            dw = M.mapStencil(border, ms, arr)
            got, d = oracle_of_dw(dw, arr)
        except NotImplementedError:
            continue   # expression larger than one descriptor
        rs = ms._resolved(rank)
        assert tuple(M.getStencilSize(rs)) == tuple(ps.size) and tuple(M.getStencilCenter(rs)) == tuple(ps.center)
        want = P.map_stencil(("fill", dt.type(3)) if bname == "fill" else (bname,), ps, arr)
        if dt.kind == "f":
            nan = np.isnan(want)
            assert np.array_equal(np.isnan(got), nan), (trial, bname)
            assert got[~nan].tobytes() == want[~nan].tobytes(), (trial, bname)
        else:
            assert got.tobytes() == want.tobytes(), (trial, bname)
        done += 1
    assert done >= 25


def test_closure_sobel_of_the_reference_benchmark_lowers_to_two_tap_lists():
    """massiv-bench/src/Data/Massiv/Bench/Sobel.hs:15-44: sqrt (sX * sX + sY * sY) with makeConvolutionStencil closures"""
    sx = M.makeConvolutionStencil(M.Sz(3, 3), (1, 1), [((-1, -1), -1), ((0, -1), -2), ((1, -1), -1), ((-1, 1), 1), ((0, 1), 2), ((1, 1), 1)])
    sy = M.makeConvolutionStencil(M.Sz(3, 3), (1, 1), [((-1, -1), -1), ((-1, 0), -2), ((-1, 1), -1), ((1, -1), 1), ((1, 0), 2), ((1, 1), 1)])
    op = M.sqrt(sx * sx + sy * sy)
    rng = np.random.default_rng(3)
    a = rng.random((9, 11))
    dw = M.mapStencil(M.Edge, op, a)
    got, d = oracle_of_dw(dw, a)
    assert (d.kind, d.n_terms, d.n_program) == (_abi.KINDS["expr"], 2, 6)   # T0 SQUARE T1 SQUARE ADD SQRT
    # the same operator from kernels (FromKernel form) differs only in tap order: same values up to rounding, and
    # exactly the closure restatement's
    dt = a.dtype
    px = P.conv_closure((3, 3), (1, 1), [((-1, -1), -1), ((0, -1), -2), ((1, -1), -1), ((-1, 1), 1), ((0, 1), 2), ((1, 1), 1)], dt)
    py = P.conv_closure((3, 3), (1, 1), [((-1, -1), -1), ((-1, 0), -2), ((-1, 1), -1), ((1, -1), 1), ((1, 0), 2), ((1, 1), 1)], dt)
    mul = lambda x, y: dt.type(x * y)
    want = P.map_stencil(("edge",), P.fmap(lambda v: dt.type(np.sqrt(v)), P.lift2(lambda x, y: dt.type(x + y), P.lift2(mul, px, px), P.lift2(mul, py, py))), a)
    assert got.tobytes() == want.tobytes()


def test_descriptor_validation():
    L = _abi.lib()
    a = np.zeros((5, 6), np.int64)
    good = M.mapStencil(M.Edge, M.sumStencil(M.Sz(2, 2)) + M.maxStencil(M.Sz(3, 1)), a).desc()
    od = (C.c_int64 * _abi.MAX_RANK)()
    assert L.mb_out_dims(C.byref(good), _abi.dims_array(a.shape), od) == _abi.OK
    bad = _abi.StencilDesc.from_buffer_copy(bytes(good))
    bad._keep = good._keep
    bad.size[0] += 1                      # not the union of the terms' sizes
    assert L.mb_out_dims(C.byref(bad), _abi.dims_array(a.shape), od) == _abi.ERR_INVALID_DESC
    with pytest.raises(_abi.MassivB200Error):  # (/) on Int: no Fractional instance
        M.mapStencil(M.Edge, M.sumStencil(M.Sz(2, 2)) / M.maxStencil(M.Sz(3, 1)), a).size()
