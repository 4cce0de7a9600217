#!/usr/bin/env python
"""Writes tests/golden/massiv_golden.json.

Every case below is a LITERAL transcription of a known-answer test or doctest that the reference
(lehins/massiv 1.0.5.0) holds for the stencil path; `src` cites it.  No GHC exists in this image, so
the expected values are the ones printed in the reference's own sources, not regenerated.  The
script needs nothing but the standard library; re-run it after editing to refresh the JSON.

Lowering used for closure stencils:
  makeConvolutionStencil sz c (\\f -> f o1 k1 . f o2 k2 ...)  ->  kind "taps", centre sz-1-c,
      source offsets -o_i, applied LAST-LISTED FIRST   (Convolution.hs:50-55)
  makeCorrelationStencil sz c (...)                           ->  kind "taps", centre c, offsets +o_i,
      last-listed first                                       (Convolution.hs:95-98)
  makeStencil sz c (\\f -> f o)                                ->  kind "taps", one tap, coefficient 1
"""
import json
import os

S = "massiv-test/tests/Test/Massiv/Array/StencilSpec.hs"
D = "massiv/src/Data/Massiv/Array/Stencil.hs"
cases = []


def case(name, src, kind, elem, in_dims, inp, expect_dims, expect, *, size, center=None, pad_lo=None,
         pad_hi=None, border="fill", fill=0, stride=None, coeffs=None, coeffs2=None, taps=None):
    cases.append(dict(name=name, src=src, kind=kind, elem=elem, in_dims=in_dims, input=inp,
                      expect_dims=expect_dims, expect=expect, size=size, center=center, pad_lo=pad_lo,
                      pad_hi=pad_hi, border=border, fill=fill, stride=stride, coeffs=coeffs,
                      coeffs2=coeffs2, taps=taps))


def conv_closure(sz, c, listed):
    """listed = [(offset, k), ...] in source order; returns (centre, taps, coeffs) in application order"""
    inv = [s - 1 - cc for s, cc in zip(sz, c)]
    rev = list(reversed(listed))
    return inv, [[-o for o in off] for off, _ in rev], [k for _, k in rev]


def corr_closure(sz, c, listed):
    rev = list(reversed(listed))
    return list(c), [list(off) for off, _ in rev], [k for _, k in rev]


ys = [1, 2, 3, 4, 5]
xs3, xs4 = [1, 2, 3], [1, 2, 3, 4]
xs3f = [([-1], 1), ([0], 2), ([1], 3)]
xs4f = [([-2], 1), ([-1], 2), ([0], 3), ([1], 4)]
xs4f_ = [([-1], 1), ([0], 2), ([1], 3), ([2], 4)]
ysConvXs3, ysConvXs4 = [4, 10, 16, 22, 22], [10, 20, 30, 34, 31]
ysCorrXs3, ysCorrXs4 = [8, 14, 20, 26, 14], [11, 20, 30, 40, 26]
ysConvXs4_, ysCorrXs4_ = [4, 10, 20, 30, 34], [20, 30, 40, 26, 14]

# makeConvolutionStencilFromKernel — StencilSpec.hs:226-235
case("convFromKernel 1x3 map", S + ":227", "conv", "i64", [5], ys, [5], ysConvXs3, size=[3], coeffs=xs3)
case("convFromKernel 1x4 map", S + ":228", "conv", "i64", [5], ys, [5], ysConvXs4, size=[4], coeffs=xs4)
case("convFromKernel 1x3 apply noPadding", S + ":229-231", "conv", "i64", [5], ys, [3], ysConvXs3[1:4],
     size=[3], coeffs=xs3, pad_lo=[0], pad_hi=[0], border="edge")
case("convFromKernel 1x4 apply noPadding", S + ":232-234", "conv", "i64", [5], ys, [2], ysConvXs4[1:3],
     size=[4], coeffs=xs4, pad_lo=[0], pad_hi=[0], border="edge")
# makeCorrelationStencilFromKernel — StencilSpec.hs:236-238
case("corrFromKernel 1x3 map", S + ":237", "corr", "i64", [5], ys, [5], ysCorrXs3, size=[3], coeffs=xs3)
case("corrFromKernel 1x4 map", S + ":238", "corr", "i64", [5], ys, [5], ysCorrXs4, size=[4], coeffs=xs4)
# makeConvolutionStencil (closure) — StencilSpec.hs:239-242
for nm, sz, c, listed, exp, line in (("1x3", [3], [1], xs3f, ysConvXs3, 240), ("1x4 c2", [4], [2], xs4f, ysConvXs4, 241),
                                     ("1x4 c1", [4], [1], xs4f_, ysConvXs4_, 242)):
    cen, taps, co = conv_closure(sz, c, listed)
    case("makeConvolutionStencil " + nm, f"{S}:{line}", "taps", "i64", [5], ys, [5], exp, size=sz, center=cen,
         taps=taps, coeffs=co)
# makeCorrelationStencil (closure) — StencilSpec.hs:243-246
for nm, sz, c, listed, exp, line in (("1x3", [3], [1], xs3f, ysCorrXs3, 244), ("1x4 c2", [4], [2], xs4f, ysCorrXs4, 245),
                                     ("1x4 c1", [4], [1], xs4f_, ysCorrXs4_, 246)):
    cen, taps, co = corr_closure(sz, c, listed)
    case("makeCorrelationStencil " + nm, f"{S}:{line}", "taps", "i64", [5], ys, [5], exp, size=sz, center=cen,
         taps=taps, coeffs=co)

# Unit tests Ix2 — StencilSpec.hs:280-297: makeStencil (Sz 3) centre (\f -> f ix), Fill 0
arr3 = [1, 2, 3, 4, 5, 6, 7, 8, 9]
for nm, cen, off, exp, line in (
        ("Direction Left", [1, 1], [0, 1], [2, 3, 0, 5, 6, 0, 8, 9, 0], 283),
        ("Direction Right", [1, 1], [0, -1], [0, 1, 2, 0, 4, 5, 0, 7, 8], 285),
        ("Direction Down", [1, 1], [1, 0], [4, 5, 6, 7, 8, 9, 0, 0, 0], 287),
        ("Direction Up", [1, 1], [-1, 0], [0, 0, 0, 1, 2, 3, 4, 5, 6], 289),
        ("Left/Top Corner", [0, 0], [2, 2], [9, 0, 0, 0, 0, 0, 0, 0, 0], 291),
        ("Right/Top Corner", [0, 2], [2, -2], [0, 0, 7, 0, 0, 0, 0, 0, 0], 293),
        ("Right/Bottom Corner", [2, 2], [-2, -2], [0, 0, 0, 0, 0, 0, 0, 0, 1], 295),
        ("Left/Bottom Corner", [2, 0], [-2, 2], [0, 0, 0, 0, 0, 0, 3, 0, 0], 297)):
    case("makeStencil " + nm, f"{S}:{line}", "taps", "i64", [3, 3], arr3, [3, 3], exp, size=[3, 3], center=cen,
         taps=[off], coeffs=[1])

# mapStencil with stride — StencilSpec.hs:298-309
kern = [-1, 0, 1, 0, 1, 0, -1, 0, 1]
case("conv stride 2 small", S + ":302-304", "conv", "i64", [3, 3], arr3, [2, 2], [-4, 8, 2, 14], size=[3, 3],
     coeffs=kern, stride=[2, 2])
case("conv stride 2 larger", S + ":305-309", "conv", "i64", [5, 5], list(range(1, 26)), [3, 3],
     [-6, 1, 14, -13, 9, 43, 4, 21, 44], size=[3, 3], coeffs=kern, stride=[2, 2])

# idStencil + Padding doctests — Stencil.hs:95-149
a23 = [1, 2, 3, 4, 5, 6]
case("idStencil noPadding", D + ":97-101", "id", "i64", [2, 3], a23, [2, 3], a23, size=[1, 1], pad_lo=[0, 0],
     pad_hi=[0, 0], border="edge")
case("idStencil Padding (1,2) (3,4) Fill 0", D + ":102-110", "id", "i64", [2, 3], a23, [6, 9],
     [0] * 9 + [0, 0, 1, 2, 3, 0, 0, 0, 0] + [0, 0, 4, 5, 6, 0, 0, 0, 0] + [0] * 27,
     size=[1, 1], pad_lo=[1, 2], pad_hi=[3, 4], border="fill", fill=0)
pads = dict(size=[1, 1], pad_lo=[2, 3], pad_hi=[2, 3])
case("idStencil Wrap", D + ":114-122", "id", "i64", [2, 3], a23, [6, 9],
     ([1, 2, 3] * 3 + [4, 5, 6] * 3) * 3, border="wrap", **pads)
case("idStencil Edge", D + ":123-131", "id", "i64", [2, 3], a23, [6, 9],
     [1, 1, 1, 1, 2, 3, 3, 3, 3] * 3 + [4, 4, 4, 4, 5, 6, 6, 6, 6] * 3, border="edge", **pads)
r1, r2 = [6, 5, 4, 4, 5, 6, 6, 5, 4], [3, 2, 1, 1, 2, 3, 3, 2, 1]
case("idStencil Reflect", D + ":132-140", "id", "i64", [2, 3], a23, [6, 9], r1 + r2 + r2 + r1 + r1 + r2,
     border="reflect", **pads)
c1, c2 = [1, 3, 2, 1, 2, 3, 2, 1, 3], [4, 6, 5, 4, 5, 6, 5, 4, 6]
case("idStencil Continue", D + ":141-149", "id", "i64", [2, 3], a23, [6, 9], (c1 + c2) * 3,
     border="continue", **pads)

# max / min pooling doctests — Stencil.hs:346-365, 379-398
a99 = list(range(10, 91))
case("maxStencil 3x3 Edge Stride 3", D + ":360-365", "max", "i64", [9, 9], a99, [3, 3],
     [30, 33, 36, 57, 60, 63, 84, 87, 90], size=[3, 3], border="edge", stride=[3, 3])
case("minStencil 3x3 Edge Stride 3", D + ":393-398", "min", "i64", [9, 9], a99, [3, 3],
     [10, 13, 16, 37, 40, 43, 64, 67, 70], size=[3, 3], border="edge", stride=[3, 3])
# sumStencil doctest — Stencil.hs:409-422
case("sumStencil (Sz2 1 2) noPadding", D + ":416-420", "sum", "i64", [2, 5],
     [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024], [2, 4], [6, 12, 24, 48, 192, 384, 768, 1536],
     size=[1, 2], pad_lo=[0, 0], pad_hi=[0, 0], border="edge")
# productStencil doctests — Stencil.hs:433-451
case("productStencil 2 Padding 0 2 (Fill 0)", D + ":440-445", "product", "i64", [2, 2], [1, 2, 3, 4], [3, 3],
     [24, 0, 0, 0, 0, 0, 0, 0, 0], size=[2, 2], pad_lo=[0, 0], pad_hi=[2, 2], border="fill", fill=0)
case("productStencil 2 Padding 0 2 Reflect", D + ":446-451", "product", "i64", [2, 2], [1, 2, 3, 4], [3, 3],
     [24, 64, 24, 144, 256, 144, 24, 64, 24], size=[2, 2], pad_lo=[0, 0], pad_hi=[2, 2], border="reflect")
# avgStencil doctest — Stencil.hs:462-476
case("avgStencil (Sz2 2 3) noPadding", D + ":470-474", "avg", "f64", [3, 4],
     [float(v) for v in range(11, 23)], [2, 2], [14.0, 15.0, 18.0, 19.0], size=[2, 3], pad_lo=[0, 0],
     pad_hi=[0, 0], border="edge")

# handleBorderIndex doctests — Core/Index.hs:218-223
border_cases = [
    dict(src="massiv/src/Data/Massiv/Core/Index.hs:218-219", border="fill", sz=[2, 3], ix=[2, 3], expect=[None, None]),
    dict(src="massiv/src/Data/Massiv/Core/Index.hs:220-221", border="wrap", sz=[2, 3], ix=[2, 3], expect=[0, 0]),
    dict(src="massiv/src/Data/Massiv/Core/Index.hs:222-223", border="edge", sz=[2, 3], ix=[2, 3], expect=[1, 2]),
]
# Border haddock pictures — Core/Index.hs:163-194: array 1 2 3 4 with four cells outside either side
pictures = {
    "fill": [0, 0, 0, 0, 1, 2, 3, 4, 0, 0, 0, 0],
    "wrap": [1, 2, 3, 4, 1, 2, 3, 4, 1, 2, 3, 4],
    "edge": [1, 1, 1, 1, 1, 2, 3, 4, 4, 4, 4, 4],
    "reflect": [4, 3, 2, 1, 1, 2, 3, 4, 4, 3, 2, 1],
    "continue": [1, 4, 3, 2, 1, 2, 3, 4, 3, 2, 1, 4],
}
for b, exp in pictures.items():
    case(f"Border picture {b}", "massiv/src/Data/Massiv/Core/Index.hs:163-194", "id", "i64", [4], [1, 2, 3, 4],
         [12], exp, size=[1], pad_lo=[4], pad_hi=[4], border=b, fill=0)

# Life: blinker oscillates with period 2 (GameOfLife.hs:47-52 pattern; rule :15-19) — known answer
blinker5 = [0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0]
blinker5h = [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
case("life blinker one generation", "massiv-examples/GameOfLife/app/GameOfLife.hs:15-36,47-52", "life", "u8",
     [5, 5], blinker5, [5, 5], blinker5h, size=[3, 3], border="wrap")

out = dict(
    note="Literal transcriptions of the reference's known-answer tests; see transcribe_reference_vectors.py",
    reference="lehins/massiv 1.0.5.0",
    cases=cases,
    border_index=border_cases,
)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "massiv_golden.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(f"wrote {len(cases)} cases + {len(border_cases)} border cases to {path}")
