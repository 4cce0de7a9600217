"""Integral approximation rules (Array/Numeric/Integral.hs): the oracle (C) and the independent
pure-Python restatement against the reference's golden numbers (IntegralSpec.hs:14-43, the Integral.hs
doctests — tests/golden/integral_golden.json) and against each other.  CPU only."""
import json
import os
import zlib

import numpy as np
import pytest

from massiv_b200 import integral as I
from oracle import oracle as O
from oracle import pyref as P

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "integral_golden.json")))
KIND = {"midpoint": "int_midpoint", "trapezoid": "int_trapezoid", "simpson": "int_simpson"}
PYREF = {"midpoint": P.midpoint_stencil, "trapezoid": P.trapezoid_stencil, "simpson": P.simpsons_stencil}


def samples(c):
    return np.array(c["samples_bits"], dtype=np.uint32).view(np.float32)


def expect(c):
    return np.array([np.float32(s) for s in c["expect"]], np.float32)


def rule_size(rule, n):
    return n if rule == "midpoint" else n + 1


def oracle_integral_approx(rule, d, sz, n, arr, **extra):
    """integralApprox through the C oracle: Dim 1 (innermost) first, Fill 0, stride n along the dim"""
    dt = arr.dtype
    dx = dt.type(d) / dt.type(n)
    rank = arr.ndim
    for dim in range(1, rank + 1):
        ax = rank - dim
        size = tuple(rule_size(rule, n) if i == ax else 1 for i in range(rank))
        stride = tuple(n if i == ax else 1 for i in range(rank))
        arr = O.apply(KIND[rule], arr, size=size, border="fill", fill=0, stride=stride, coeffs=[dx], **extra)
    return arr[tuple(slice(0, s) for s in sz)]


@pytest.mark.parametrize("c", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_matches_reference_numbers(c):
    got = oracle_integral_approx(c["rule"], c["d"], c["sz"], c["n"], samples(c))
    assert got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), expect(c).view(np.uint32)), (got, c["expect"])


@pytest.mark.parametrize("c", [c for c in GOLD["cases"] if c["n"] <= 32], ids=lambda c: c["name"])
def test_pyref_matches_reference_numbers(c):
    arr = samples(c)
    dt = arr.dtype
    got = P.integral_approx(PYREF[c["rule"]], c["d"], tuple(c["sz"]), c["n"], arr)
    assert np.array_equal(np.asarray(got, dt).view(np.uint32), expect(c).view(np.uint32))


def test_printed_sample_arrays_agree_with_generated_ones():
    # Integral.hs:307-309 prints the n=4 sample array; the IntegralSpec n=4 inputs were generated with expf
    a = next(c for c in GOLD["cases"] if c["name"] == "doctest simpsonsRule n=4")
    b = next(c for c in GOLD["cases"] if c["name"] == "IntegralSpec simpson n=4")
    assert a["samples_bits"] == b["samples_bits"]
    t = next(c for c in GOLD["cases"] if c["name"] == "doctest integralApprox trapezoidStencil")
    assert [np.float32(s).view(np.uint32) for s in t["x_printed"]] == t["x_bits"]


@pytest.mark.parametrize("name", ["fromFunction", "fromFunctionMidpoint"])
def test_from_function_doctests(name):
    g = GOLD["grids"][name]
    f = lambda scale, ix: scale(ix[0]) + scale(ix[1])
    got = getattr(I, name)("Seq", f, g["a"], g["d"], g["sz"], g["n"], dtype=np.float64)
    assert np.array_equal(got, np.array(g["expect"]))
    pf = {"fromFunction": P.from_function, "fromFunctionMidpoint": P.from_function_midpoint}[name]
    ref = pf(lambda scale, ix: scale(ix[0]) + scale(ix[1]), g["a"], g["d"], tuple(g["sz"]), g["n"], np.dtype(np.float64))
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sample_generators_match_pyref(dtype):
    f = lambda scale, ix: scale(ix[0]) * scale(ix[1]) - scale(ix[1])
    dt = np.dtype(dtype)
    for a, d, sz, n in [(-1.25, 0.3, (2, 3), 3), (0.1, 7.0, (3, 1), 5)]:
        assert np.array_equal(I.fromFunction("Par", f, a, d, sz, n, dtype), P.from_function(f, a, d, sz, n, dt))
        assert np.array_equal(I.fromFunctionMidpoint("Par", f, a, d, sz, n, dtype), P.from_function_midpoint(f, a, d, sz, n, dt))


CASES = [(rule, rank, dim, n, elem) for rule in KIND for rank in (1, 2, 3) for dim in range(1, rank + 1)
         for n in (2, 4, 6) for elem in ("f32", "f64")]


@pytest.mark.parametrize("rule,rank,dim,n,elem", CASES)
def test_oracle_equals_pyref_random(rule, rank, dim, n, elem):
    """integrateWith along one dimension, plus the same stencil unstrided and under other borders"""
    seed = zlib.crc32(repr((rule, rank, dim, n, elem)).encode())
    rng = np.random.default_rng(seed)
    dt = np.dtype(O.ELEM2NP[elem])
    shape = tuple(int(rng.integers(1, 4)) * (n if i == rank - dim else 1) + int(rng.integers(0, 2)) for i in range(rank))
    arr = rng.standard_normal(shape).astype(dt)
    dx = dt.type(rng.uniform(0.01, 2.0))
    size = tuple(rule_size(rule, n) if i == rank - dim else 1 for i in range(rank))
    st = PYREF[rule](dx, dim, n, rank, dt)
    assert st.size == size
    for border, stride in [("fill", tuple(n if i == rank - dim else 1 for i in range(rank))), ("reflect", None), ("wrap", (2,) * rank)]:
        want = P.map_stencil((border, dt.type(0)) if border == "fill" else (border,), st, arr, stride)
        for extra in ({}, {"flags": O.FLAG_NO_FAST}, {"nthreads": 3}):
            got = O.apply(KIND[rule], arr, size=size, border=border, fill=0, stride=stride, coeffs=[dx], **extra)
            assert got.shape == want.shape
            assert np.array_equal(got.view(np.uint8), np.asarray(want, dt).view(np.uint8)), (border, stride)


def test_rejections():
    arr = np.zeros((8,), np.float64)
    with pytest.raises(O.OracleError):  # odd n: error in the reference (Integral.hs:106-109)
        O.apply("int_simpson", arr, size=(4,), coeffs=[0.5])
    with pytest.raises(O.OracleError):  # Fractional e
        O.apply("int_midpoint", np.zeros((8,), np.int32), size=(4,), coeffs=[1])
    with pytest.raises(O.OracleError):  # one dimension only
        O.apply("int_trapezoid", np.zeros((8, 8)), size=(3, 3), coeffs=[0.5])
    with pytest.raises(ValueError):
        I.simpsonsStencil(0.5, 1, 3)
