"""GPU parity for the integral-approximation rules through the C ABI:
the reference's golden numbers (IntegralSpec.hs:14-43, Integral.hs doctests) and randomized
differential tests against the oracle.  Bit-exact (0 ulp)."""
import ctypes
import json
import os
import zlib

import numpy as np
import pytest

import massiv_b200 as M
from massiv_b200 import integral as I
from oracle import oracle as O
from helpers import bits_equal

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "integral_golden.json")))
KIND = {"midpoint": "int_midpoint", "trapezoid": "int_trapezoid", "simpson": "int_simpson"}
STENCIL = {"midpoint": I.midpointStencil, "trapezoid": I.trapezoidStencil, "simpson": I.simpsonsStencil}


def rule_size(rule, n):
    return n if rule == "midpoint" else n + 1


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("c", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_reference_numbers(c, resident):
    arr = np.array(c["samples_bits"], dtype=np.uint32).view(np.float32)
    src = M.Device.default().fromHost(arr) if resident else arr
    got = I.integralApprox(STENCIL[c["rule"]], c["d"], c["sz"], c["n"], src)
    got = got.toHost() if resident else got
    want = np.array([np.float32(s) for s in c["expect"]], np.float32)
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (got, c["expect"])


def test_rule_drivers_match_spec_shape():
    """rule Seq BN (gaussian .) 0 2 (Sz1 1) n  — with this machine's expf the 18 numbers of
    IntegralSpec.hs:20-43 come out as printed there (glibc's expf, which GHC's exp @Float calls)"""
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype, libm.expf.argtypes = ctypes.c_float, [ctypes.c_float]
    g = np.vectorize(lambda x: np.float32(libm.expf(float(np.float32(x) * np.float32(x)))), otypes=[np.float32])
    f = lambda scale, ix: g(scale(ix[0]))
    drivers = {"midpoint": I.midpointRule, "trapezoid": I.trapezoidRule, "simpson": I.simpsonsRule}
    for c in GOLD["cases"]:
        if not c["name"].startswith("IntegralSpec"):
            continue
        got = drivers[c["rule"]]("Seq", "BN", f, c["a"], c["d"], c["sz"], c["n"], dtype=np.float32)
        assert got.shape == (1,) and got[0] == np.float32(c["expect"][0]), (c["name"], got)


CASES = [(rule, rank, dim, n, elem) for rule in KIND for rank in (1, 2, 3, 4) for dim in range(1, rank + 1)
         for n in (2, 6, 16) for elem in ("f32", "f64")]


@pytest.mark.parametrize("rule,rank,dim,n,elem", CASES)
def test_integrate_with_vs_oracle(rule, rank, dim, n, elem):
    seed = zlib.crc32(repr(("gpu", rule, rank, dim, n, elem)).encode())
    rng = np.random.default_rng(seed)
    dt = np.dtype(O.ELEM2NP[elem])
    ax = rank - dim
    shape = tuple(int(rng.integers(1, 5)) * (n if i == ax else 1) + int(rng.integers(0, 2)) + (3 if i != ax else 0) for i in range(rank))
    arr = rng.standard_normal(shape).astype(dt)
    dx = dt.type(rng.uniform(0.01, 2.0))
    size = tuple(rule_size(rule, n) if i == ax else 1 for i in range(rank))
    stride = tuple(n if i == ax else 1 for i in range(rank))
    want = O.apply(KIND[rule], arr, size=size, border="fill", fill=0, stride=stride, coeffs=[dx])
    got = I.integrateWith(lambda dm, k: STENCIL[rule](dx, dm, k), dim, n, arr)
    assert bits_equal(got, want)
    # the same stencil is an ordinary stencil: other borders, no stride, device-resident
    st = STENCIL[rule](dx, dim, n)
    st = M.Stencil(st.kind, st.size, st.center, np.asarray(st.coeffs).astype(dt), None, None, 0, None, st.along)
    dev = M.Device.default()
    for border, bobj in (("edge", M.Edge), ("continue", M.Continue), ("wrap", M.Wrap)):
        want = O.apply(KIND[rule], arr, size=size, border=border, coeffs=[dx])
        got = M.compute(M.mapStencil(bobj, st, dev.fromHost(arr))).toHost()
        assert bits_equal(got, want), border


@pytest.mark.parametrize("rule", list(KIND))
@pytest.mark.parametrize("elem", ["f32", "f64"])
def test_integral_approx_2d_large(rule, elem):
    """the reference bench's shape (massiv-bench/bench/Integral.hs:17-24: Sz2 10 10 cells, n = 100) and larger"""
    dt = np.dtype(O.ELEM2NP[elem])
    for sz, n in (((10, 10), 100), ((37, 19), 8)):
        rng = np.random.default_rng(zlib.crc32(repr((rule, elem, sz, n)).encode()))
        shape = tuple(s * n + (0 if rule == "midpoint" else 1) for s in sz)
        arr = rng.random(shape).astype(dt)
        d = 1.5
        dx = dt.type(d) / dt.type(n)
        want = arr
        for dim in (1, 2):
            ax = 2 - dim
            want = O.apply(KIND[rule], want, size=tuple(rule_size(rule, n) if i == ax else 1 for i in range(2)), border="fill",
                           fill=0, stride=tuple(n if i == ax else 1 for i in range(2)), coeffs=[dx], nthreads=8)
        want = want[:sz[0], :sz[1]]
        got = I.integralApprox(STENCIL[rule], d, sz, n, arr)
        assert bits_equal(got, np.ascontiguousarray(want))


def test_statuses():
    arr = np.zeros((9,), np.float64)
    st = M.Stencil("int_simpson", (4,), None, np.array([0.5]))  # odd n smuggled past the constructor
    with pytest.raises(M.MassivB200Error):
        M.compute(M.mapStencil(M.Fill(0.0), st, arr))
    with pytest.raises(M.MassivB200Error):  # Fractional e
        M.compute(M.mapStencil(M.Fill(0), M.Stencil("int_midpoint", (4,), None, np.array([1])), np.zeros((9,), np.int32)))
    with pytest.raises(IndexError):
        I.integrateWith(lambda dm, k: I.midpointStencil(0.5, dm, k), 3, 2, np.zeros((4, 4)))
