"""Golden vectors for the integral-approximation rules.

Expected numbers are TRANSCRIBED from the reference's own tests and doctests:
  massiv-test/tests/Test/Massiv/Array/Numeric/IntegralSpec.hs:14-43   (18 values, Float, e^(x^2) on [0,2])
  massiv/src/Data/Massiv/Array/Numeric/Integral.hs:228-240, 256-267   (fromFunction / fromFunctionMidpoint)
  massiv/src/Data/Massiv/Array/Numeric/Integral.hs:307-317, 345-360   (sample array, Simpson, trapezoid)
The sample arrays fed to the rules are generated here — `exp :: Float -> Float` is C `expf`, taken
from this machine's libm through ctypes — and stored bit-exactly, so the checks do not depend on the
libm of the machine that runs the tests.  Haskell's `show` prints the shortest decimal that reads
back as the same Float, so a result matches when float32(result) == float32(literal).

Run from the repo root:  python tests/golden/make_integral_vectors.py
"""
import ctypes
import json
import os

import numpy as np

f32 = np.float32
libm = ctypes.CDLL("libm.so.6")
libm.expf.restype = ctypes.c_float
libm.expf.argtypes = [ctypes.c_float]


def gaussian(x):  # IntegralSpec.hs:10-11: exp (x ^ (2 :: Int))
    return f32(libm.expf(float(f32(x) * f32(x))))


def edges(a, d, sz, n):  # fromFunction, rank 1
    return [f32(a) + f32(d) * f32(i) / f32(n) for i in range(n * sz + 1)]


def mids(a, d, sz, n):  # fromFunctionMidpoint, rank 1
    dx2 = f32(d) / f32(n) / f32(2)
    return [(dx2 + f32(a)) + f32(d) * f32(i) / f32(n) for i in range(n * sz)]


def bits(xs):
    return [int(np.asarray(x, f32).view(np.uint32)) for x in xs]


SPEC = {  # IntegralSpec.hs:20-43
    "midpoint": ["14.485613", "15.905677", "16.311854", "16.417171", "16.443748", "16.450407"],
    "trapezoid": ["20.644558", "17.565086", "16.735381", "16.523618", "16.470394", "16.457073"],
    "simpson": ["17.353626", "16.538595", "16.458815", "16.453030", "16.452653", "16.452629"],
}
NS = [4, 8, 16, 32, 64, 128]

cases = []
for rule, want in SPEC.items():
    for n, w in zip(NS, want):
        xs = mids(0, 2, 1, n) if rule == "midpoint" else edges(0, 2, 1, n)
        cases.append({"name": f"IntegralSpec {rule} n={n}", "ref": "IntegralSpec.hs:20-43", "rule": rule, "elem": "f32",
                      "a": 0, "d": 2, "sz": [1], "n": n, "samples_bits": bits(gaussian(x) for x in xs), "expect": [w]})

# Integral.hs:307-317: the sample array is printed, so it is pinned independently of expf
cases.append({"name": "doctest simpsonsRule n=4", "ref": "Integral.hs:307-317", "rule": "simpson", "elem": "f32",
              "a": 0, "d": 2, "sz": [1], "n": 4,
              "samples_bits": bits(f32(v) for v in ["1.0", "1.2840254", "2.7182817", "9.487736", "54.59815"]),
              "samples_printed": ["1.0", "1.2840254", "2.7182817", "9.487736", "54.59815"], "expect": ["17.353626"]})
# Integral.hs:345-352: xArrX4 / trapezoid over 4 cells
x4 = edges(-2, 1, 4, 4)
cases.append({"name": "doctest integralApprox trapezoidStencil", "ref": "Integral.hs:340-352", "rule": "trapezoid",
              "elem": "f32", "a": -2, "d": 1, "sz": [4], "n": 4, "samples_bits": bits(gaussian(x) for x in x4),
              "x_printed": ["-2.0", "-1.75", "-1.5", "-1.25", "-1.0", "-0.75", "-0.5", "-0.25", "0.0", "0.25", "0.5", "0.75",
                            "1.0", "1.25", "1.5", "1.75", "2.0"],
              "x_bits": bits(x4), "expect": ["16.074406", "1.4906789", "1.4906789", "16.074408"]})
# Integral.hs:358-360
cases.append({"name": "doctest simpsonsRule n=128 over 4 cells", "ref": "Integral.hs:358-360", "rule": "simpson", "elem": "f32",
              "a": -2, "d": 1, "sz": [4], "n": 128, "samples_bits": bits(gaussian(x) for x in edges(-2, 1, 4, 128)),
              "expect": ["14.989977", "1.4626511", "1.4626517", "14.989977"]})

grids = {
    # Integral.hs:228-240: fromFunction Seq (\ scale (i :. j) -> scale i + scale j :: Double) (-2) 1 (Sz 4) 2
    "fromFunction": {"ref": "Integral.hs:228-240", "a": -2, "d": 1, "sz": [4, 4], "n": 2,
                     "expect": [[-4.0 + 0.5 * (i + j) for j in range(9)] for i in range(9)]},
    # Integral.hs:256-267: fromFunctionMidpoint with the same arguments
    "fromFunctionMidpoint": {"ref": "Integral.hs:256-267", "a": -2, "d": 1, "sz": [4, 4], "n": 2,
                             "expect": [[-3.5 + 0.5 * (i + j) for j in range(8)] for i in range(8)]},
}
# the two printed grids, first and last rows as they appear in the doctests (a transcription check)
assert grids["fromFunction"]["expect"][0] == [-4.0, -3.5, -3.0, -2.5, -2.0, -1.5, -1.0, -0.5, 0.0]
assert grids["fromFunction"]["expect"][8] == [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0]
assert grids["fromFunctionMidpoint"]["expect"][0] == [-3.5, -3.0, -2.5, -2.0, -1.5, -1.0, -0.5, 0.0]
assert grids["fromFunctionMidpoint"]["expect"][7] == [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5]

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "integral_golden.json")
with open(out, "w") as fh:
    json.dump({"source": "lehins/massiv IntegralSpec.hs + Integral.hs doctests (see refs)", "cases": cases, "grids": grids}, fh, indent=1)
print(f"{len(cases)} cases -> {out}")
