"""Pins the CPU oracle (oracle/massiv_oracle.c) and the pure-Python restatement (oracle/pyref.py)
against the reference's own golden vectors (tests/golden/massiv_golden.json, transcribed from
massiv-test/tests/Test/Massiv/Array/StencilSpec.hs and the Stencil.hs doctests) and against each
other on randomized small cases modelled on the reference's QuickCheck properties
(StencilSpec.hs:34-60, 107-140, 180-182, 248-274; Test/Massiv/Core/Index.hs:254-258, 329-364).

CPU only (no gpu marker).
"""
import itertools
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as P

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "massiv_golden.json")))
BORDERS = ["fill", "wrap", "edge", "reflect", "continue"]


def run_oracle_case(c, **extra):
    arr = np.array(c["input"], dtype=O.ELEM2NP[c["elem"]]).reshape(c["in_dims"])
    return O.apply(c["kind"], arr, size=c["size"], center=c["center"], pad_lo=c["pad_lo"], pad_hi=c["pad_hi"],
                   border=c["border"], fill=c["fill"], stride=c["stride"], coeffs=c["coeffs"],
                   coeffs2=c["coeffs2"], taps=c["taps"], **extra)


def pyref_stencil(kind, dt, size, center=None, coeffs=None, coeffs2=None, taps=None, flags=0):
    dt = np.dtype(dt)
    size = tuple(size)
    if kind == "id":
        return P.id_stencil(len(size))
    if kind == "sum":
        return P.sum_stencil(size, dt)
    if kind == "product":
        return P.product_stencil(size, dt)
    if kind == "avg":
        return P.avg_stencil(size, dt)
    if kind == "max":
        return P.max_stencil(size, dt)
    if kind == "min":
        return P.min_stencil(size, dt)
    if kind == "conv":
        return P.conv_from_kernel(np.array(coeffs).astype(dt).reshape(size))
    if kind == "corr":
        return P.corr_from_kernel(np.array(coeffs).astype(dt).reshape(size))
    if kind == "life":
        return P.life_stencil()
    if kind == "mag2":
        mk = P.corr_from_kernel if flags & O.FLAG_MAG2_CORR else P.conv_from_kernel
        return P.mag2(mk(np.array(coeffs).astype(dt).reshape(size)), mk(np.array(coeffs2).astype(dt).reshape(size)), dt)
    if kind == "taps":
        # already-lowered tap list: application order, effective source offsets
        def func(get, ix):
            acc = dt.type(0)
            for off, kv in zip(taps, coeffs):
                m = dt.type(get(tuple(i + o for i, o in zip(ix, off))) * dt.type(kv))
                acc = dt.type(m + acc)
            return acc
        return P.Stencil(size, tuple(center), func)
    raise ValueError(kind)


def run_pyref(kind, arr, *, size, center=None, pad_lo=None, pad_hi=None, border="fill", fill=0, stride=None,
              coeffs=None, coeffs2=None, taps=None, flags=0):
    st = pyref_stencil(kind, arr.dtype, size, center, coeffs, coeffs2, taps, flags)
    b = ("fill", arr.dtype.type(fill)) if border == "fill" else (border,)
    if pad_lo is None:
        pad_lo, pad_hi = P.same_padding(st)
    return P.apply_stencil(tuple(pad_lo), tuple(pad_hi), b, st, arr, stride)


@pytest.mark.parametrize("c", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_golden_c_oracle(c):
    out = run_oracle_case(c)
    assert list(out.shape) == c["expect_dims"], c["src"]
    assert out.reshape(-1).tolist() == c["expect"], c["src"]
    # the vectorised interior loop nest and the plain tap loop must agree
    out2 = run_oracle_case(c, flags=O.FLAG_NO_FAST)
    assert np.array_equal(out, out2)


@pytest.mark.parametrize("c", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_golden_pyref(c):
    arr = np.array(c["input"], dtype=O.ELEM2NP[c["elem"]]).reshape(c["in_dims"])
    out = run_pyref(c["kind"], arr, size=c["size"], center=c["center"], pad_lo=c["pad_lo"], pad_hi=c["pad_hi"],
                    border=c["border"], fill=c["fill"], stride=c["stride"], coeffs=c["coeffs"],
                    coeffs2=c["coeffs2"], taps=c["taps"])
    assert list(out.shape) == c["expect_dims"], c["src"]
    assert out.reshape(-1).tolist() == c["expect"], c["src"]


@pytest.mark.parametrize("c", GOLD["border_index"], ids=[c["border"] for c in GOLD["border_index"]])
def test_golden_border_index(c):
    got = [O.border_index(c["border"], k, i) for k, i in zip(c["sz"], c["ix"])]
    assert got == c["expect"], c["src"]


def test_foldl_stencil_visits_row_major():
    """Stencil.hs:279-295 doctest: foldlStencil (flip (:)) [] pins the row-major tap order."""
    a = np.arange(11, 23).reshape(3, 4)
    st = P.fold_stencil(lambda acc, x: [int(x)] + acc, [], (3, 2))
    out = P.apply_stencil((0, 0), (0, 0), ("edge",), st, a, out_dtype=object)
    assert out.shape == (1, 3)
    assert out[0, 0] == [20, 19, 16, 15, 12, 11]
    assert out[0, 1] == [21, 20, 17, 16, 13, 12]
    assert out[0, 2] == [22, 21, 18, 17, 14, 13]
    out = P.apply_stencil((1, 0), (0, 0), ("fill", 10), st, a, out_dtype=object)
    assert out.shape == (2, 3)
    assert out[0, 0] == [16, 15, 12, 11, 10, 10]
    assert out[1, 2] == [22, 21, 18, 17, 14, 13]


def test_iterate_until_pascal():
    """Manifest/Internal.hs:301-328 doctest: iterateUntil + mapStencil (Fill 0) pascal."""
    arr = np.zeros((8, 8), np.int64)
    arr[0, 0] = 1

    def pascal_f(get, ix):
        i, j = ix
        cur = get((i, j))
        return cur if cur != 0 else get((i - 1, j - 1)) + get((i - 1, j))
    pascal = P.Stencil((2, 2), (1, 1), pascal_f)
    n = 0
    while True:  # iterateUntil: apply, then test convergence on the new array
        arr = P.map_stencil(("fill", 0), pascal, arr)
        if arr[7, 7] != 0:
            break
        n += 1
    expect = [[1, 0, 0, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0, 0, 0], [1, 2, 1, 0, 0, 0, 0, 0], [1, 3, 3, 1, 0, 0, 0, 0],
              [1, 4, 6, 4, 1, 0, 0, 0], [1, 5, 10, 10, 5, 1, 0, 0], [1, 6, 15, 20, 15, 6, 1, 0],
              [1, 7, 21, 35, 35, 21, 7, 1]]
    assert arr.tolist() == expect
    assert n == 6  # seven applications fill rows 1..7


def test_oracle_iterate_counts_steps():
    """mo_iterate applies the step exactly n times (iterateLoop counter 0..n-1) for odd and even n."""
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2, (9, 11)).astype(np.uint8)
    cur = a
    for n in range(1, 6):
        cur = O.apply("life", cur, size=(3, 3), border="wrap")
        assert np.array_equal(O.iterate("life", a, n, size=(3, 3), border="wrap"), cur)


# ---- handleBorderIndex properties: Test/Massiv/Core/Index.hs:254-258, 329-364 -----------------

def test_border_repair_is_safe_and_periodic():
    rng = np.random.default_rng(7)
    for _ in range(3000):
        k = int(rng.integers(1, 12))
        i = int(rng.integers(-40, 40))
        period = int(rng.integers(1, 5))
        for b in BORDERS[1:]:
            r = O.border_index(b, k, i)
            assert 0 <= r < k
            if 0 <= i < k:
                assert r == i
        if not (0 <= i < k):
            assert O.border_index("fill", k, i) is None
            assert O.border_index("wrap", k, i) == O.border_index("wrap", k, i + period * k)
            assert O.border_index("edge", k, i) == (0 if i < 0 else k - 1)
            sg = -1 if i < 0 else 1
            assert O.border_index("reflect", k, i) == O.border_index("reflect", k, i + 2 * sg * period * k)
            assert O.border_index("continue", k, i) == O.border_index("reflect", k, i - 1 if i < 0 else i + 1)


def test_border_index_zero_exception():
    for b in BORDERS[1:]:
        with pytest.raises(O.OracleError) as e:
            O.border_index(b, 0, 3)
        assert e.value.status == O.ERR_INDEX_ZERO
    assert O.border_index("fill", 0, 3) is None
    with pytest.raises(O.OracleError) as e:
        O.apply("sum", np.zeros((0, 3), np.int64), size=(2, 2), border="edge", pad_lo=(1, 0), pad_hi=(1, 1))
    assert e.value.status == O.ERR_INDEX_ZERO
    # mapStencil over an empty array has no cells to force, so nothing throws (laziness)
    assert O.apply("sum", np.zeros((0, 3), np.int64), size=(2, 2), border="edge").shape == (0, 3)
    # Fill never repairs an index, so an empty source is fine: every cell is the fold of fills
    out = O.apply("sum", np.zeros((0, 3), np.int64), size=(2, 2), border="fill", fill=5,
                  pad_lo=(1, 0), pad_hi=(1, 1))
    assert out.shape == (1, 3) and (out == 20).all()


# ---- randomized differential: C oracle vs pyref ------------------------------------------------

def _rand_case(rng, rank, kind, dt):
    dims = tuple(int(rng.integers(0 if rng.random() < 0.1 else 1, 7)) for _ in range(rank))
    size = tuple(int(rng.integers(0 if kind in ("sum", "product", "avg", "max", "min") and rng.random() < 0.1 else 1, 5))
                 for _ in range(rank))
    if dt.kind == "f":
        arr = rng.standard_normal(dims).astype(dt)
    elif dt.itemsize == 8:
        arr = rng.integers(-2**62, 2**62, dims).astype(dt)
    else:
        info = np.iinfo(dt)
        arr = rng.integers(info.min, int(info.max) + 1, dims).astype(dt)
    kw = dict(size=size, border=BORDERS[int(rng.integers(0, 5))])
    if kw["border"] == "fill":
        kw["fill"] = float(rng.standard_normal()) if dt.kind == "f" else int(rng.integers(-100, 100)) % (2 ** (8 * dt.itemsize - 1))
    n = int(np.prod(size))
    if kind in ("conv", "corr", "mag2"):
        kw["coeffs"] = (rng.standard_normal(n) if dt.kind == "f" else rng.integers(-5, 6, n)).tolist()
    if kind == "mag2":
        kw["coeffs2"] = rng.standard_normal(n).tolist()
    if rng.random() < 0.4:  # applyStencil with arbitrary padding instead of mapStencil
        kw["pad_lo"] = tuple(int(rng.integers(0, 4)) for _ in range(rank))
        kw["pad_hi"] = tuple(int(rng.integers(0, 4)) for _ in range(rank))
    if rng.random() < 0.4:
        kw["stride"] = tuple(int(rng.integers(1, 4)) for _ in range(rank))
    return arr, kw


@pytest.mark.parametrize("rank", [1, 2, 3, 4])
@pytest.mark.parametrize("kind,dtype", [("sum", np.int64), ("sum", np.float32), ("sum", np.uint8), ("product", np.int16),
                                        ("product", np.float64), ("avg", np.float64), ("avg", np.float32),
                                        ("max", np.int32), ("min", np.uint16), ("max", np.int8), ("conv", np.float32),
                                        ("conv", np.int64), ("corr", np.float64), ("corr", np.uint32),
                                        ("mag2", np.float32), ("mag2", np.float64)])
def test_c_oracle_equals_pyref_random(rank, kind, dtype):
    rng = np.random.default_rng(zlib.crc32(f"{rank}-{kind}-{np.dtype(dtype).name}".encode()))
    trials = {1: 25, 2: 20, 3: 8, 4: 3}[rank]
    for _ in range(trials):
        arr, kw = _rand_case(rng, rank, kind, np.dtype(dtype))
        try:
            want = run_pyref(kind, arr, **kw)
        except P.IndexZeroException:
            with pytest.raises(O.OracleError) as e:
                O.apply(kind, arr, **kw)
            assert e.value.status == O.ERR_INDEX_ZERO
            continue
        got = O.apply(kind, arr, **kw)
        assert got.shape == want.shape, (arr.shape, kw)
        # bit-exact, NaN-aware (avg over an empty window is 0/0)
        assert np.array_equal(got.view(np.uint8), want.astype(arr.dtype).view(np.uint8)) or \
            np.array_equal(got, want, equal_nan=True), (arr.shape, kw)
        got_threads = O.apply(kind, arr, nthreads=3, **kw)
        assert np.array_equal(got, got_threads, equal_nan=True)


def test_fast_path_equals_tap_loop():
    rng = np.random.default_rng(11)
    for kind, dt in (("avg", np.float64), ("conv", np.float32), ("mag2", np.float32), ("sum", np.int64), ("corr", np.float64)):
        for rank in (2, 3):
            dims = tuple(int(rng.integers(20, 40)) for _ in range(rank - 1)) + (700,)
            size = (3,) * rank
            arr = rng.standard_normal(dims).astype(dt) if np.dtype(dt).kind == "f" else rng.integers(-9, 9, dims).astype(dt)
            n = 3 ** rank
            kw = dict(size=size, border="reflect")
            if kind in ("conv", "corr", "mag2"):
                kw["coeffs"] = rng.standard_normal(n).tolist() if np.dtype(dt).kind == "f" else rng.integers(-3, 3, n).tolist()
            if kind == "mag2":
                kw["coeffs2"] = rng.standard_normal(n).tolist()
            a = O.apply(kind, arr, nthreads=4, **kw)
            b = O.apply(kind, arr, flags=O.FLAG_NO_FAST, **kw)
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_life_random_vs_pyref():
    rng = np.random.default_rng(3)
    for _ in range(10):
        a = rng.integers(0, 4, (int(rng.integers(1, 8)), int(rng.integers(1, 8)))).astype(np.uint8)  # values > 1 are legal
        want = run_pyref("life", a, size=(3, 3), border="wrap")
        assert np.array_equal(O.apply("life", a, size=(3, 3), border="wrap"), want)


# ---- reference properties restated ------------------------------------------------------------

def test_prop_map_singleton_stencil_is_identity():
    """StencilSpec.hs:34-60 (f = id): a 1-cell stencil maps to the array itself, any border/stride."""
    rng = np.random.default_rng(5)
    for rank in (1, 2, 3, 4):
        dims = tuple(int(rng.integers(1, 6)) for _ in range(rank))
        a = rng.integers(-50, 50, dims).astype(np.int64)
        for b in BORDERS:
            assert np.array_equal(O.apply("id", a, size=(1,) * rank, border=b, fill=9), a)
            assert np.array_equal(O.apply("id", a, size=(1,) * rank, border=b, stride=(1,) * rank), a)


def test_prop_apply_zero_stencil():
    """StencilSpec.hs:44-49: a zero-size stencil under noPadding yields size arr filled with acc0."""
    a = np.arange(12, dtype=np.int64).reshape(3, 4)
    out = O.apply("sum", a, size=(0, 0), pad_lo=(0, 0), pad_hi=(0, 0), border="edge")
    assert out.shape == (3, 4) and (out == 0).all()
    out = O.apply("product", a, size=(0, 2), pad_lo=(0, 0), pad_hi=(0, 0), border="edge")
    assert out.shape == (3, 3) and (out == 1).all()
    out = O.apply("max", a, size=(0, 0), pad_lo=(0, 0), pad_hi=(0, 0), border="edge")
    assert (out == np.iinfo(np.int64).min).all()


def test_prop_avg_padding_1_1_equals_centered_blur():
    """StencilSpec.hs:180-182: mapStencil b avg3x3 == applyStencil (Padding 1 1 b) (avgStencil 3)."""
    rng = np.random.default_rng(9)
    a = rng.integers(-20, 20, (6, 7)).astype(np.float64)  # small ints: sums exact, like Rational
    for b in BORDERS:
        got = O.apply("avg", a, size=(3, 3), pad_lo=(1, 1), pad_hi=(1, 1), border=b, fill=2.0)
        ones = [1.0] * 9
        want = O.apply("conv", a, size=(3, 3), coeffs=ones, border=b, fill=2.0) / 9.0
        assert np.array_equal(got, want)


def test_prop_conv_equals_corr_rotate180():
    """StencilSpec.hs:262-274."""
    rng = np.random.default_rng(13)
    for shape, ksz in (((7,), (3,)), ((9,), (5,)), ((6, 5), (3, 3))):
        a = rng.integers(-9, 9, shape).astype(np.int64)
        k = rng.integers(-4, 4, ksz).astype(np.int64)
        rot = k[tuple(slice(None, None, -1) for _ in ksz)]
        assert np.array_equal(O.apply("conv", a, size=ksz, coeffs=k.reshape(-1).tolist()),
                              O.apply("corr", a, size=ksz, coeffs=rot.reshape(-1).tolist()))


def test_out_dims_formula():
    for in_d, s, lo, hi, st in itertools.product((0, 1, 5), (0, 1, 3, 7), (0, 2), (0, 3), (1, 2, 4)):
        full = max(0, lo + hi + in_d - max(0, s - 1))
        assert O.out_dims("sum", (in_d,), size=(s,), pad_lo=(lo,), pad_hi=(hi,), stride=(st,)) == ((full - 1) // st + 1,)
    # avgStencil unions its size with `pure`'s size 1 (Stencil/Internal.hs:83-90)
    assert O.out_dims("avg", (4,), elem="f64", size=(0,), pad_lo=(0,), pad_hi=(0,)) == (4,)
