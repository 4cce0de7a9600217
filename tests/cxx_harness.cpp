// Tests of the C++ host mirror (include/massiv_b200.hpp) written the way the reference's own tests are:
// same constructors, same calls, the reference's expected values transcribed literally
// (massiv-test/tests/Test/Massiv/Array/StencilSpec.hs, the Stencil.hs / Integral.hs doctests,
// IntegralSpec.hs).  Built and run by tests/test_cxx_harness.py:
//     cxx_harness geometry    what needs no device
//     cxx_harness gpu         the golden vectors through compute / computeWithStride / iterateStencil
#include <cstdio>
#include <cstring>
#include <string>

#include "massiv_b200.hpp"

using namespace massiv_b200;

static int failures = 0;
#define CHECK(cond, msg)                                                                    \
  do {                                                                                      \
    if (!(cond)) { std::printf("FAIL %s (%s:%d)\n", msg, __FILE__, __LINE__); ++failures; } \
  } while (0)

template <typename E> static Array<E> iota(const Sz &sz, E first) {
  Array<E> a(sz);
  for (size_t i = 0; i < a.elems.size(); ++i) a.elems[i] = (E)(first + (E)i);
  return a;
}

static int geometry() {
  // Stencil.hs:102-110: applyStencil (Padding (Sz2 1 2) (Sz2 3 4) (Fill 0)) idStencil on a 2x3 array is 6x9
  const Array<int64_t> a = iota<int64_t>(Sz{2, 3}, 1);
  const Padding<int64_t> pad{{1, 2}, {3, 4}, Border<int64_t>::Fill(0)};
  CHECK((outputSize(applyStencil(pad, idStencil<int64_t>(2), a)) == Sz{6, 9}), "padded idStencil size");
  // Stencil.hs:346-364: Stride 3 of maxStencil (Sz 3) on 9x9 is 3x3
  const Array<int64_t> b = iota<int64_t>(Sz{9, 9}, 10);
  const Ix stride{3, 3};
  CHECK((outputSize(mapStencil(Border<int64_t>::Edge(), maxStencil<int64_t>(Sz{3, 3}), b), &stride) == Sz{3, 3}), "strided size");
  // Bounded has no Double instance
  bool threw = false;
  try {
    const Array<double> c(Sz{4, 4});
    outputSize(mapStencil(Border<double>::Edge(), maxStencil<double>(Sz{3, 3}), c));
  } catch (const B200Error &e) { threw = e.status == MB_ERR_UNSUPPORTED; }
  CHECK(threw, "maxStencil on Double is rejected");
  // Integral.hs:106-109
  threw = false;
  try { simpsonsStencil<float>(0.5f, 1, 3, 1); } catch (const std::invalid_argument &) { threw = true; }
  CHECK(threw, "Simpson's rule with an odd number of samples is an error");
  // makeConvolutionStencil inverts the centre (Convolution.hs:50-55); samePadding follows it
  const auto s = makeConvolutionStencil<int64_t>(Sz{3, 3}, Ix{0, 1}, {{Ix{0, 0}, 1}});
  CHECK((s.center == Ix{2, 1}), "inverted centre");
  const auto p = samePadding(s, Border<int64_t>::Edge());
  CHECK((p.paddingFromOrigin == Ix{2, 1} && p.paddingFromBottom == Ix{0, 1}), "samePadding");
  CHECK((Sz{-3, 2}.dims == Ix{0, 2}), "Sz clamps negative extents");
  {  // Stencil/Internal.hs:83-90: liftA2 unions centres and sizes: sum (Sz 3) [centre 0] + corr [1,2,3] [centre 1]
    const Array<int64_t> k(Sz{3}, {1, 2, 3}), ys = iota<int64_t>(Sz{5}, 1);
    const auto e = sumStencil<int64_t>(Sz{3}) + makeCorrelationStencilFromKernel(k);
    CHECK((e.kind == MB_EXPR && e.center == Ix{1} && e.size == Ix{4} && e.terms.size() == 2 && e.program.size() == 3), "union of centres and sizes");
    CHECK((outputSize(mapStencil(Border<int64_t>::Fill(0), e, ys)) == Sz{5}), "mapStencil of an expression keeps the size");
    bool threw = false;  // (/) needs Fractional
    try { outputSize(mapStencil(Border<int64_t>::Fill(0), e / e, ys)); } catch (const B200Error &x) { threw = x.status == MB_ERR_UNSUPPORTED; }
    CHECK(threw, "(/) on Int stencils is rejected");
  }
  return failures;
}

// every Int case of tests/golden/massiv_golden.json (the reference's StencilSpec / doctest vectors)
struct GoldenCase {
  const char *name;
  mb_kind kind;
  Ix in_dims;
  std::vector<int64_t> input;
  Ix expect_dims;
  std::vector<int64_t> expect;
  Ix size;
  int has_center;
  Ix center, pad_lo, pad_hi;
  mb_border border;
  int64_t fill;
  Ix stride;
  std::vector<int64_t> coeffs, taps;
};
static const GoldenCase kGolden[] = {
#include "golden/golden_i64.inc"
};

static void golden(const Device &dev) {
  for (const GoldenCase &c : kGolden) {
    Stencil<int64_t> st;
    st.kind = c.kind; st.size = c.size; st.coeffs = c.coeffs; st.tap_offsets = c.taps;
    if (c.has_center) st.center = c.center;
    else
      for (int64_t n : c.size)  // stencilCenter of the built-in constructors
        st.center.push_back(c.kind == MB_CONV ? n - 1 - n / 2 : (c.kind == MB_CORR ? n / 2 : 0));
    const Array<int64_t> arr(Sz(c.in_dims), c.input);
    const Border<int64_t> b{c.border, c.fill};
    const DW<int64_t> dw = c.pad_lo.empty() ? mapStencil(b, st, arr) : applyStencil(Padding<int64_t>{c.pad_lo, c.pad_hi, b}, st, arr);
    const Array<int64_t> got = c.stride.empty() ? compute(dev, dw) : computeWithStride(dev, c.stride, dw);
    CHECK((got.sz == Sz(c.expect_dims) && got.elems == c.expect), c.name);
  }
}

static int gpu() {
  Device dev(0);
  golden(dev);
  {  // StencilSpec.hs:227, 237: conv / corr of [1,2,3] over [1..5], Fill 0
    const Array<int64_t> ys = iota<int64_t>(Sz{5}, 1), k(Sz{3}, {1, 2, 3});
    const auto conv = compute(dev, mapStencil(Border<int64_t>::Fill(0), makeConvolutionStencilFromKernel(k), ys));
    CHECK((conv.elems == std::vector<int64_t>{4, 10, 16, 22, 22}), "convolution known answer");
    const auto corr = compute(dev, mapStencil(Border<int64_t>::Fill(0), makeCorrelationStencilFromKernel(k), ys));
    CHECK((corr.elems == std::vector<int64_t>{8, 14, 20, 26, 14}), "correlation known answer");
  }
  {  // Stencil.hs:346-364: max pooling
    const Array<int64_t> a = iota<int64_t>(Sz{9, 9}, 10);
    const auto r = computeWithStride(dev, Ix{3, 3}, mapStencil(Border<int64_t>::Edge(), maxStencil<int64_t>(Sz{3, 3}), a));
    CHECK((r.sz == Sz{3, 3} && r.elems == std::vector<int64_t>{30, 33, 36, 57, 60, 63, 84, 87, 90}), "max pooling doctest");
  }
  {  // StencilSpec.hs:298-309: mapStencil with stride
    const Array<int64_t> kernel(Sz{3, 3}, {-1, 0, 1, 0, 1, 0, -1, 0, 1});
    const auto st = makeConvolutionStencilFromKernel(kernel);
    const Array<int64_t> arr = iota<int64_t>(Sz{3, 3}, 1), large = iota<int64_t>(Sz{5, 5}, 1);
    const auto small = computeWithStride(dev, Ix{2, 2}, mapStencil(Border<int64_t>::Fill(0), st, arr));
    CHECK((small.elems == std::vector<int64_t>{-4, 8, 2, 14}), "map stencil with stride on small array");
    const auto big = computeWithStride(dev, Ix{2, 2}, mapStencil(Border<int64_t>::Fill(0), st, large));
    CHECK((big.elems == std::vector<int64_t>{-6, 1, 14, -13, 9, 43, 4, 21, 44}), "map stencil with stride on larger array");
  }
  {  // StencilSpec.hs:311-326: the closure form of Sobel X equals the kernel form, as convolution and as correlation
    const std::vector<std::pair<Ix, int64_t>> sobelX = {{{-1, -1}, -1}, {{-1, 1}, 1}, {{0, -1}, -2}, {{0, 1}, 2}, {{1, -1}, -1}, {{1, 1}, 1}};
    const Array<int64_t> sobelKernelX(Sz{3, 3}, {-1, 0, 1, -2, 0, 2, -1, 0, 1});
    Array<int64_t> img(Sz{17, 23});
    for (size_t i = 0; i < img.elems.size(); ++i) img.elems[i] = (int64_t)((i * 2654435761u) % 1000) - 500;
    const auto b = Border<int64_t>::Reflect();
    CHECK((compute(dev, mapStencil(b, makeCorrelationStencil<int64_t>(Sz{3, 3}, Ix{1, 1}, sobelX), img)) ==
           compute(dev, mapStencil(b, makeCorrelationStencilFromKernel(sobelKernelX), img))), "correlation: closure form == kernel form");
    CHECK((compute(dev, mapStencil(b, makeConvolutionStencil<int64_t>(Sz{3, 3}, Ix{1, 1}, sobelX), img)) ==
           compute(dev, mapStencil(b, makeConvolutionStencilFromKernel(sobelKernelX), img))), "convolution: closure form == kernel form");
  }
  {  // Stencil.hs:470-474: avgStencil (Sz2 2 3) with noPadding on 11..22
    const Array<double> a = iota<double>(Sz{3, 4}, 11.0);
    const auto r = compute(dev, applyStencil(noPadding<double>(2), avgStencil<double>(Sz{2, 3}), a));
    CHECK((r.sz == Sz{2, 2} && r.elems == std::vector<double>{14, 15, 18, 19}), "avgStencil doctest");
  }
  {  // Integral.hs:307-317: Simpson's rule over the printed sample array; IntegralSpec.hs:36
    const Array<float> y(Sz{5}, {1.0f, 1.2840254f, 2.7182817f, 9.487736f, 54.59815f});
    const float dx = 2.0f / 4.0f;
    const auto r = integrateWith(dev, [&](int dim, int n) { return simpsonsStencil<float>(dx, dim, n, 1); }, 1, 4, y);
    CHECK(r.elems.size() == 2 && r.elems[0] == 17.353626f, "simpsonsRule doctest");
    const auto t = integrateWith(dev, [&](int dim, int n) { return trapezoidStencil<float>(dx, dim, n, 1); }, 1, 4, y);
    CHECK(t.elems[0] == 20.644558f, "trapezoidRule (IntegralSpec.hs:28)");
  }
  {  // GameOfLife.hs: a blinker has period 2 on a torus
    Array<uint8_t> g(Sz{64, 64});
    g.elems[10 * 64 + 20] = g.elems[11 * 64 + 20] = g.elems[12 * 64 + 20] = 1;
    const auto one = iterateStencil(dev, 1, Border<uint8_t>::Wrap(), lifeStencil(), g);
    CHECK(one.elems[11 * 64 + 19] == 1 && one.elems[11 * 64 + 20] == 1 && one.elems[11 * 64 + 21] == 1 && one.elems[10 * 64 + 20] == 0,
          "blinker after one generation");
    CHECK((iterateStencil(dev, 100, Border<uint8_t>::Wrap(), lifeStencil(), g) == g), "blinker after 100 generations");
  }
  {  // fmap over a stencil (Stencil/Internal.hs:114-146): abs (sobel) * 2
    const Array<double> kx(Sz{3, 3}, {-1, 0, 1, -2, 0, 2, -1, 0, 1});
    Array<double> img(Sz{40, 50});
    for (size_t i = 0; i < img.elems.size(); ++i) img.elems[i] = (double)((i * 40503u) % 97) / 7.0;
    const auto plain = compute(dev, mapStencil(Border<double>::Edge(), makeConvolutionStencilFromKernel(kx), img));
    const auto mapped = compute(dev, mapStencil(Border<double>::Edge(), abs(makeConvolutionStencilFromKernel(kx)) * 2.0, img));
    bool ok = plain.elems.size() == mapped.elems.size();
    for (size_t i = 0; ok && i < plain.elems.size(); ++i) {
      const double v = plain.elems[i], w = (v == 0 ? 0.0 : (v > 0 ? v : -v)) * 2.0;
      ok = std::memcmp(&w, &mapped.elems[i], 8) == 0;
    }
    CHECK(ok, "fmap abs, (* 2)");
  }
  {  // Num (Stencil ix e) on two stencils (Stencil/Internal.hs:104-131), hand-derived: over [1..5] under Fill 0
     // sumStencil (Sz 3) gives [6,9,12,9,5], the correlation with [1,2,3] gives [8,14,20,26,14] (StencilSpec.hs:206)
    const Array<int64_t> k(Sz{3}, {1, 2, 3}), ys = iota<int64_t>(Sz{5}, 1);
    const auto s1 = sumStencil<int64_t>(Sz{3});
    const auto s2 = makeCorrelationStencilFromKernel(k);
    const auto b = Border<int64_t>::Fill(0);
    CHECK((compute(dev, mapStencil(b, s1 + s2, ys)).elems == std::vector<int64_t>{14, 23, 32, 35, 19}), "s1 + s2");
    CHECK((compute(dev, mapStencil(b, s1 * s2, ys)).elems == std::vector<int64_t>{48, 126, 240, 234, 70}), "s1 * s2");
    CHECK((compute(dev, mapStencil(b, abs(s1 - s2), ys)).elems == std::vector<int64_t>{2, 5, 8, 17, 9}), "abs (s1 - s2)");
    CHECK((compute(dev, mapStencil(b, maxStencils(s1, s2) - minStencils(s1, s2), ys)).elems == std::vector<int64_t>{2, 5, 8, 17, 9}), "liftA2 max - liftA2 min");
    CHECK((compute(dev, mapStencil(b, (s1 + s2) * int64_t(2) - int64_t(1), ys)).elems == std::vector<int64_t>{27, 45, 63, 69, 37}), "fmap over an expression");
  }
  {  // the closure-form Sobel operator of the reference's benchmark (massiv-bench/src/Data/Massiv/Bench/Sobel.hs:15-44)
     // against the same operator assembled from the FromKernel magnitude: equal wherever both are exact (Int-valued data)
    const std::vector<std::pair<Ix, double>> sx = {{{-1, -1}, -1}, {{0, -1}, -2}, {{1, -1}, -1}, {{-1, 1}, 1}, {{0, 1}, 2}, {{1, 1}, 1}};
    const std::vector<std::pair<Ix, double>> sy = {{{-1, -1}, -1}, {{-1, 0}, -2}, {{-1, 1}, -1}, {{1, -1}, 1}, {{1, 0}, 2}, {{1, 1}, 1}};
    const auto cx = makeConvolutionStencil<double>(Sz{3, 3}, Ix{1, 1}, sx), cy = makeConvolutionStencil<double>(Sz{3, 3}, Ix{1, 1}, sy);
    const auto op = sqrt(cx * cx + cy * cy);
    Array<double> img(Sz{120, 200});
    for (size_t i = 0; i < img.elems.size(); ++i) img.elems[i] = (double)((i * 2654435761u) % 256);
    const Array<double> kx(Sz{3, 3}, {-1, 0, 1, -2, 0, 2, -1, 0, 1}), ky(Sz{3, 3}, {-1, -2, -1, 0, 0, 0, 1, 2, 1});
    const auto ref = compute(dev, mapStencil(Border<double>::Edge(), magnitude(makeConvolutionStencilFromKernel(kx), makeConvolutionStencilFromKernel(ky)), img));
    const auto got = compute(dev, mapStencil(Border<double>::Edge(), op, img));
    CHECK((got == ref), "closure-form Sobel operator == kernel-form magnitude on integer-valued data");
  }
  {  // iterateUntil (\_ prev cur -> cur == prev): a block is a still life, found after one generation
    Array<uint8_t> g(Sz{32, 64});
    g.elems[5 * 64 + 5] = g.elems[5 * 64 + 6] = g.elems[6 * 64 + 5] = g.elems[6 * 64 + 6] = 1;
    const auto r = iterateUntil(dev, Until::Equal(), 50, 1, Border<uint8_t>::Wrap(), lifeStencil(), g);
    CHECK((r.converged && r.steps == 1 && r.array == g), "iterateUntil: fixed point of Life");
    g.elems[20 * 64 + 30] = g.elems[20 * 64 + 31] = g.elems[20 * 64 + 32] = 1;  // plus a blinker: never equal
    const auto q = iterateUntil(dev, Until::Equal(), 7, 1, Border<uint8_t>::Wrap(), lifeStencil(), g);
    CHECK((!q.converged && q.steps == 7), "iterateUntil: the step cap ends an oscillator");
  }
  {  // exceptions: IndexZeroException, SizeMismatchException
    bool threw = false;
    try {
      const Array<int64_t> empty(Sz{0, 3});
      compute(dev, applyStencil(Padding<int64_t>{{1, 0}, {1, 1}, Border<int64_t>::Edge()}, sumStencil<int64_t>(Sz{2, 2}), empty));
    } catch (const IndexZeroException &) { threw = true; }
    CHECK(threw, "IndexZeroException");
    threw = false;
    try {
      const Array<uint8_t> g(Sz{8, 8});
      auto shrinking = lifeStencil();
      iterateStencil(dev, 2, Border<uint8_t>::Wrap(), shrinking, Array<uint8_t>(Sz{8, 8}));
      const auto dw = applyStencil(noPadding<uint8_t>(2), shrinking, g);
      mb_stencil_desc d;
      dw.describe(&d, nullptr);
      Array<uint8_t> out(g.sz);
      dev.check(mb_iterate(dev.handle(), &d, g.sz.dims.data(), g.elems.data(), out.elems.data(), 2));
    } catch (const SizeMismatchException &) { threw = true; }
    CHECK(threw, "iterate refuses a shrinking stencil (SizeMismatchException)");
  }
  return failures;
}

int main(int argc, char **argv) {
  const std::string mode = argc > 1 ? argv[1] : "geometry";
  int bad = 0;
  try {
    bad = mode == "gpu" ? geometry() + gpu() : geometry();
  } catch (const std::exception &e) {
    std::printf("FAIL uncaught exception: %s\n", e.what());
    bad = 1;
  }
  if (bad == 0) std::printf("CXX_HARNESS_OK\n");
  return bad == 0 ? 0 : 1;
}
