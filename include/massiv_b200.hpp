// massiv_b200.hpp — C++ host mirror of massiv's stencil interface over the C ABI (massiv_b200.h).
//
// The reference is compiled code (Haskell) whose toolchain is absent from the build image, so next to the
// Haskell shim (haskell/Data/Massiv/Array/Stencil/B200.hs, source only) this header gives the same
// interface in a language that IS compiled and tested here: same names, argument order, meaning and
// error behaviour as Data.Massiv.Array.Stencil / Data.Massiv.Array.Numeric.Integral, so that tests
// written against it read like the reference's own (tests/cxx_harness.cpp).  Header-only; link with
// libmassiv_b200.so.  Nothing here computes: every compute call goes through mb_run / mb_iterate, and
// there is no CPU fallback.
//
//   reference (massiv/src/Data/Massiv/...)                    here
//   Sz, Ix (Core/Index/Internal.hs)                            Sz
//   Border: Fill e | Wrap | Edge | Reflect | Continue         Border<E>::Fill(e), Wrap(), ...
//   Array S ix e (Array/Manifest/Storable.hs)                  Array<E>   (dense row-major, last index fastest)
//   Stencil ix e e (Array/Stencil/Internal.hs:30-34)           Stencil<E> (reified, see mb_kind)
//   idStencil / sumStencil / ... / makeConvolutionStencil...   same names
//   Padding, samePadding, noPadding (Array/Stencil.hs:152-179) same names
//   applyStencil, mapStencil (Array/Stencil.hs:76-86,187-217)  same names -> DW<E>
//   compute, computeWithStride (Array/Manifest/Internal.hs)    same names (take the Device)
//   iterateUntil (\n _ _ -> n >= k-1) (\_ -> mapStencil b st) iterateStencil
//   iterateUntil (\_ p c -> c == p | maximum |c - p| < tol)    iterateUntil (Until::Equal / Until::MaxAbsDiff)
//   Applicative / Num / Fractional / Floating (Stencil ix e)   operators + - * / on two stencils, abs, signum,
//     (Array/Stencil/Internal.hs:83-174)                         recip, sqrt, square, minStencils, maxStencils, pure
//   midpoint/trapezoid/simpsonsStencil, integrateWith          same names (Array/Numeric/Integral.hs)
//   IndexZeroException, SizeMismatchException                  same names
#ifndef MASSIV_B200_HPP
#define MASSIV_B200_HPP

#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "massiv_b200.h"

namespace massiv_b200 {

using Ix = std::vector<int64_t>;

// ---- exceptions (Core/Index/Internal.hs:1031, 1136-1142) ---------------------------------------------
struct B200Error : std::runtime_error {
  int status;
  B200Error(int st, const std::string &msg) : std::runtime_error(msg), status(st) {}
};
struct IndexZeroException : B200Error { using B200Error::B200Error; };
struct SizeMismatchException : B200Error { using B200Error::B200Error; };

// ---- element types -----------------------------------------------------------------------------------
template <typename E> struct Elem;
#define MB_HPP_ELEM(T, CODE) template <> struct Elem<T> { static constexpr int code = CODE; }
MB_HPP_ELEM(uint8_t, MB_U8); MB_HPP_ELEM(int8_t, MB_I8); MB_HPP_ELEM(uint16_t, MB_U16); MB_HPP_ELEM(int16_t, MB_I16);
MB_HPP_ELEM(uint32_t, MB_U32); MB_HPP_ELEM(int32_t, MB_I32); MB_HPP_ELEM(uint64_t, MB_U64); MB_HPP_ELEM(int64_t, MB_I64);
MB_HPP_ELEM(float, MB_F32); MB_HPP_ELEM(double, MB_F64);
#undef MB_HPP_ELEM

// ---- Sz: negative extents clamp to 0 (Core/Index/Internal.hs:122-125) ----------------------------------
struct Sz {
  Ix dims;
  Sz() = default;
  Sz(std::initializer_list<int64_t> d) { for (int64_t v : d) dims.push_back(v < 0 ? 0 : v); }
  explicit Sz(const Ix &d) { for (int64_t v : d) dims.push_back(v < 0 ? 0 : v); }
  int rank() const { return (int)dims.size(); }
  int64_t totalElem() const { int64_t n = 1; for (int64_t v : dims) n *= v; return n; }
  bool operator==(const Sz &o) const { return dims == o.dims; }
};

// ---- Border (Core/Index.hs:159-195) ------------------------------------------------------------------
template <typename E> struct Border {
  mb_border kind;
  E fill;
  static Border Fill(E e) { return {MB_FILL, e}; }
  static Border Wrap() { return {MB_WRAP, E()}; }
  static Border Edge() { return {MB_EDGE, E()}; }
  static Border Reflect() { return {MB_REFLECT, E()}; }
  static Border Continue() { return {MB_CONTINUE, E()}; }
};

// ---- Array S ix e ------------------------------------------------------------------------------------
template <typename E> struct Array {
  Sz sz;
  std::vector<E> elems;
  Array() = default;
  explicit Array(const Sz &s) : sz(s), elems((size_t)s.totalElem()) {}
  Array(const Sz &s, std::vector<E> v) : sz(s), elems(std::move(v)) {
    if ((int64_t)elems.size() != sz.totalElem()) throw SizeMismatchException(MB_ERR_SIZE_MISMATCH, "Array: size and element count differ");
  }
  const Sz &size() const { return sz; }
  bool operator==(const Array &o) const { return sz == o.sz && elems == o.elems; }
};

// ---- Stencil ix e e, reified -------------------------------------------------------------------------
template <typename E> struct Stencil {
  mb_kind kind = MB_ID;
  Ix size, center;
  std::vector<E> coeffs, coeffs2;
  std::vector<int64_t> tap_offsets;  // TAPS: ntaps * rank, application order
  uint32_t flags = 0;
  std::vector<mb_post_op> post;
  // kind MB_EXPR (liftA2 / fmap / pure over built-in stencils, Array/Stencil/Internal.hs:93-174): the sub-stencils
  // and the postfix program; size / center hold the union of Internal.hs:83-90
  std::vector<Stencil> terms;
  std::vector<mb_xins> program;

  // Functor / Num / Fractional / Floating instances (Array/Stencil/Internal.hs:36-41, 114-146) for the
  // functions with an exactly specified result; everything else is a closure and stays on massiv's CPU path
  Stencil fmap(mb_post op, E c = E()) const {
    Stencil s = *this;
    if (kind == MB_EXPR) {  // fmap over an expression: further program steps
      mb_xins k;
      std::memset(&k, 0, sizeof(k));
      k.op = MB_X_CONST;
      std::memcpy(k.operand, &c, sizeof(E));
      mb_xins o;
      std::memset(&o, 0, sizeof(o));
      switch (op) {
        case MB_POST_NEGATE: o.op = MB_X_NEGATE; break;
        case MB_POST_ABS: o.op = MB_X_ABS; break;
        case MB_POST_SIGNUM: o.op = MB_X_SIGNUM; break;
        case MB_POST_SQUARE: o.op = MB_X_SQUARE; break;
        case MB_POST_RECIP: o.op = MB_X_RECIP; break;
        case MB_POST_SQRT: o.op = MB_X_SQRT; break;
        case MB_POST_ADD: s.program.push_back(k); o.op = MB_X_ADD; break;
        case MB_POST_SUB: s.program.push_back(k); o.op = MB_X_SUB; break;
        case MB_POST_MUL: s.program.push_back(k); o.op = MB_X_MUL; break;
        case MB_POST_MIN: s.program.push_back(k); o.op = MB_X_MIN; break;
        case MB_POST_MAX: s.program.push_back(k); o.op = MB_X_MAX; break;
        case MB_POST_DIV: s.program.push_back(k); o.op = MB_X_DIV; break;
        case MB_POST_RSUB: s.program.insert(s.program.begin(), k); o.op = MB_X_SUB; break;
        default: s.program.insert(s.program.begin(), k); o.op = MB_X_DIV; break;
      }
      s.program.push_back(o);
      return s;
    }
    mb_post_op p;
    std::memset(&p, 0, sizeof(p));
    p.op = op;
    std::memcpy(p.operand, &c, sizeof(E));
    s.post.push_back(p);
    return s;
  }
  Stencil operator-() const { return fmap(MB_POST_NEGATE); }
  Stencil operator+(E c) const { return fmap(MB_POST_ADD, c); }
  Stencil operator-(E c) const { return fmap(MB_POST_SUB, c); }
  Stencil operator*(E c) const { return fmap(MB_POST_MUL, c); }
  Stencil operator/(E c) const { return fmap(MB_POST_DIV, c); }
  Stencil assumeFinite() const { Stencil s = *this; s.flags |= MB_FLAG_ASSUME_FINITE; return s; }
};
namespace detail {
inline mb_xins xins(int op, int arg = 0) {
  mb_xins x;
  std::memset(&x, 0, sizeof(x));
  x.op = op; x.arg = arg;
  return x;
}
// a stencil as (terms, program): a built-in one is a single term; its post-maps become program steps
template <typename E> void as_expr(const Stencil<E> &s, std::vector<Stencil<E>> *terms, std::vector<mb_xins> *prog) {
  const int base = (int)terms->size();
  std::vector<mb_xins> body;
  if (s.kind == MB_EXPR) {
    for (const Stencil<E> &t : s.terms) terms->push_back(t);
    for (mb_xins x : s.program) { if (x.op == MB_X_TERM) x.arg += base; body.push_back(x); }
  } else {
    if (s.kind == MB_LIFE || s.kind == MB_MAG2 || s.kind >= MB_INT_MIDPOINT) throw B200Error(MB_ERR_UNSUPPORTED, "this stencil does not combine into expressions");
    Stencil<E> t = s;
    t.post.clear(); t.flags = 0;
    terms->push_back(t);
    body.push_back(xins(MB_X_TERM, base));
  }
  for (const mb_post_op &q : s.post) {
    mb_xins c = xins(MB_X_CONST);
    std::memcpy(c.operand, q.operand, 8);
    switch (q.op) {
      case MB_POST_NEGATE: body.push_back(xins(MB_X_NEGATE)); break;
      case MB_POST_ABS: body.push_back(xins(MB_X_ABS)); break;
      case MB_POST_SIGNUM: body.push_back(xins(MB_X_SIGNUM)); break;
      case MB_POST_SQUARE: body.push_back(xins(MB_X_SQUARE)); break;
      case MB_POST_RECIP: body.push_back(xins(MB_X_RECIP)); break;
      case MB_POST_SQRT: body.push_back(xins(MB_X_SQRT)); break;
      case MB_POST_ADD: body.push_back(c); body.push_back(xins(MB_X_ADD)); break;
      case MB_POST_SUB: body.push_back(c); body.push_back(xins(MB_X_SUB)); break;
      case MB_POST_MUL: body.push_back(c); body.push_back(xins(MB_X_MUL)); break;
      case MB_POST_MIN: body.push_back(c); body.push_back(xins(MB_X_MIN)); break;
      case MB_POST_MAX: body.push_back(c); body.push_back(xins(MB_X_MAX)); break;
      case MB_POST_DIV: body.push_back(c); body.push_back(xins(MB_X_DIV)); break;
      case MB_POST_RSUB: body.insert(body.begin(), c); body.push_back(xins(MB_X_SUB)); break;  // c - x
      default: body.insert(body.begin(), c); body.push_back(xins(MB_X_DIV)); break;             // MB_POST_RDIV: c / x
    }
  }
  prog->insert(prog->end(), body.begin(), body.end());
}
template <typename E> int64_t geom_of(const Stencil<E> &s, size_t i) { return (s.kind == MB_AVG && s.size[i] < 1) ? 1 : s.size[i]; }
// liftA2 f a b: both evaluated at the same index; centre = max of the centres, size = centre + max (size - centre)
template <typename E> Stencil<E> lift2(int op, const Stencil<E> &a, const Stencil<E> &b) {
  if (a.size.size() != b.size.size()) throw B200Error(MB_ERR_INVALID_DESC, "stencils of different ranks");
  Stencil<E> r;
  r.kind = MB_EXPR;
  r.flags = (a.flags | b.flags) & MB_FLAG_ASSUME_FINITE;
  as_expr(a, &r.terms, &r.program);
  as_expr(b, &r.terms, &r.program);
  r.program.push_back(xins(op));
  for (size_t i = 0; i < a.size.size(); ++i) {
    const int64_t c = a.center[i] > b.center[i] ? a.center[i] : b.center[i];
    const int64_t ea = geom_of(a, i) - a.center[i], eb = geom_of(b, i) - b.center[i];
    r.center.push_back(c);
    r.size.push_back(c + (ea > eb ? ea : eb));
  }
  return r;
}
}  // namespace detail
// Num / Fractional instances on two stencils (Array/Stencil/Internal.hs:114-131): liftA2 (+), (-), (*), (/)
template <typename E> Stencil<E> operator+(const Stencil<E> &a, const Stencil<E> &b) { return detail::lift2(MB_X_ADD, a, b); }
template <typename E> Stencil<E> operator-(const Stencil<E> &a, const Stencil<E> &b) { return detail::lift2(MB_X_SUB, a, b); }
template <typename E> Stencil<E> operator*(const Stencil<E> &a, const Stencil<E> &b) { return detail::lift2(MB_X_MUL, a, b); }
template <typename E> Stencil<E> operator/(const Stencil<E> &a, const Stencil<E> &b) { return detail::lift2(MB_X_DIV, a, b); }
template <typename E> Stencil<E> minStencils(const Stencil<E> &a, const Stencil<E> &b) { return detail::lift2(MB_X_MIN, a, b); }  // liftA2 min
template <typename E> Stencil<E> maxStencils(const Stencil<E> &a, const Stencil<E> &b) { return detail::lift2(MB_X_MAX, a, b); }  // liftA2 max
template <typename E> Stencil<E> square(const Stencil<E> &s) { return s.fmap(MB_POST_SQUARE); }  // fmap (^2)
template <typename E> Stencil<E> abs(const Stencil<E> &s) { return s.fmap(MB_POST_ABS); }
template <typename E> Stencil<E> signum(const Stencil<E> &s) { return s.fmap(MB_POST_SIGNUM); }
template <typename E> Stencil<E> recip(const Stencil<E> &s) { return s.fmap(MB_POST_RECIP); }
template <typename E> Stencil<E> sqrt(const Stencil<E> &s) { return s.fmap(MB_POST_SQRT); }

namespace detail {
template <typename E> Stencil<E> fold(mb_kind k, const Sz &sz) {  // foldlStencil: centre 0 (Stencil.hs:298-302)
  Stencil<E> s;
  s.kind = k; s.size = sz.dims; s.center.assign(sz.dims.size(), 0);
  return s;
}
inline int64_t floor_div(int64_t a, int64_t b) { int64_t q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
}  // namespace detail

template <typename E> Stencil<E> idStencil(int rank) { return detail::fold<E>(MB_ID, Sz(Ix((size_t)rank, 1))); }          // Stencil.hs:269
template <typename E> Stencil<E> sumStencil(const Sz &sz) { return detail::fold<E>(MB_SUM, sz); }                          // :425-427
template <typename E> Stencil<E> productStencil(const Sz &sz) { return detail::fold<E>(MB_PRODUCT, sz); }                  // :454-456
template <typename E> Stencil<E> avgStencil(const Sz &sz) { return detail::fold<E>(MB_AVG, sz); }                          // :479-481
template <typename E> Stencil<E> maxStencil(const Sz &sz) { return detail::fold<E>(MB_MAX, sz); }                          // :368-370
template <typename E> Stencil<E> minStencil(const Sz &sz) { return detail::fold<E>(MB_MIN, sz); }                          // :401-403

// Stencil/Convolution.hs:64-80: centre = size - 1 - (size `quot` 2)
template <typename E> Stencil<E> makeConvolutionStencilFromKernel(const Array<E> &k) {
  Stencil<E> s;
  s.kind = MB_CONV; s.size = k.sz.dims; s.coeffs = k.elems;
  for (int64_t n : s.size) s.center.push_back(n - 1 - n / 2);
  return s;
}
// Stencil/Convolution.hs:107-121: centre = size `div` 2
template <typename E> Stencil<E> makeCorrelationStencilFromKernel(const Array<E> &k) {
  Stencil<E> s;
  s.kind = MB_CORR; s.size = k.sz.dims; s.coeffs = k.elems;
  for (int64_t n : s.size) s.center.push_back(detail::floor_div(n, 2));
  return s;
}
// Stencil/Convolution.hs:44-57 / 89-100: the closure body `\f -> f o1 k1 . f o2 k2 ...` as a list in source
// order; the LAST listed tap is applied first; a convolution reads at ix - o and inverts the centre
template <typename E>
Stencil<E> makeConvolutionStencil(const Sz &sz, const Ix &center, const std::vector<std::pair<Ix, E>> &taps) {
  Stencil<E> s;
  s.kind = MB_TAPS; s.size = sz.dims;
  for (size_t i = 0; i < sz.dims.size(); ++i) s.center.push_back(sz.dims[i] - 1 - center[i]);
  for (auto it = taps.rbegin(); it != taps.rend(); ++it) {
    for (int64_t o : it->first) s.tap_offsets.push_back(-o);
    s.coeffs.push_back(it->second);
  }
  return s;
}
template <typename E>
Stencil<E> makeCorrelationStencil(const Sz &sz, const Ix &center, const std::vector<std::pair<Ix, E>> &taps) {
  Stencil<E> s;
  s.kind = MB_TAPS; s.size = sz.dims; s.center = center;
  for (auto it = taps.rbegin(); it != taps.rend(); ++it) {
    for (int64_t o : it->first) s.tap_offsets.push_back(o);
    s.coeffs.push_back(it->second);
  }
  return s;
}
// massiv-examples/GameOfLife/app/GameOfLife.hs:21-33
inline Stencil<uint8_t> lifeStencil() {
  Stencil<uint8_t> s;
  s.kind = MB_LIFE; s.size = {3, 3}; s.center = {1, 1};
  return s;
}
// sqrt (fmap (^2) sx + fmap (^2) sy) of two FromKernel convolution stencils (Bench/Sobel.hs:39-44), one fused pass
template <typename E> Stencil<E> magnitude(const Stencil<E> &a, const Stencil<E> &b) {
  if (a.kind != b.kind || a.size != b.size || (a.kind != MB_CONV && a.kind != MB_CORR))
    throw B200Error(MB_ERR_UNSUPPORTED, "magnitude: two FromKernel stencils of one constructor and size");
  Stencil<E> s = a;
  s.kind = MB_MAG2; s.coeffs2 = b.coeffs;
  if (a.kind == MB_CORR) s.flags |= MB_FLAG_MAG2_CORR;
  return s;
}

// ---- integral approximation (Array/Numeric/Integral.hs:54-135); Dim 1 is the innermost dimension -------
namespace detail {
template <typename E> Stencil<E> rule(mb_kind k, E dx, int dim, int64_t len, int rank) {
  if (dim < 1 || dim > rank) throw std::out_of_range("IndexDimensionException");
  Stencil<E> s;
  s.kind = k; s.size.assign((size_t)rank, 1); s.center.assign((size_t)rank, 0);
  s.size[(size_t)(rank - dim)] = len;
  s.coeffs = {dx};
  return s;
}
}  // namespace detail
template <typename E> Stencil<E> midpointStencil(E dx, int dim, int k, int rank) { return detail::rule<E>(MB_INT_MIDPOINT, dx, dim, k, rank); }
template <typename E> Stencil<E> trapezoidStencil(E dx, int dim, int n, int rank) { return detail::rule<E>(MB_INT_TRAPEZOID, dx, dim, n + 1, rank); }
template <typename E> Stencil<E> simpsonsStencil(E dx, int dim, int n, int rank) {
  if (n % 2) throw std::invalid_argument("Number of sample points for Simpson's rule stencil should be even, but received: " + std::to_string(n));
  return detail::rule<E>(MB_INT_SIMPSON, dx, dim, n + 1, rank);
}

// ---- Padding (Array/Stencil.hs:152-179) --------------------------------------------------------------
template <typename E> struct Padding {
  Ix paddingFromOrigin, paddingFromBottom;
  Border<E> paddingWithElement;
};
template <typename E> Padding<E> noPadding(int rank) { return {Ix((size_t)rank, 0), Ix((size_t)rank, 0), Border<E>::Edge()}; }
template <typename E> Padding<E> samePadding(const Stencil<E> &st, const Border<E> &b) {
  Padding<E> p{{}, {}, b};
  for (size_t i = 0; i < st.size.size(); ++i) {
    const int64_t g = (st.kind == MB_AVG && st.size[i] < 1) ? 1 : st.size[i];  // liftA2 (/) with `pure n`
    p.paddingFromOrigin.push_back(st.center[i] < 0 ? 0 : st.center[i]);
    const int64_t hi = g - (st.center[i] + 1);
    p.paddingFromBottom.push_back(hi < 0 ? 0 : hi);
  }
  return p;
}

// ---- the delayed windowed array ------------------------------------------------------------------------
template <typename E> struct DW {
  Padding<E> padding;
  Stencil<E> stencil;
  const Array<E> *arr;
  mutable std::vector<mb_term> term_store;  // MB_EXPR: the mb_term array the descriptor points at
  // fills a descriptor; the vectors referenced live in this object
  void describe(mb_stencil_desc *d, const Ix *stride) const {
    std::memset(d, 0, sizeof(*d));
    const int rank = arr->sz.rank();
    if ((int)stencil.size.size() != rank) throw B200Error(MB_ERR_INVALID_DESC, "stencil and array ranks differ");
    d->kind = stencil.kind; d->elem = Elem<E>::code; d->rank = rank; d->border = padding.paddingWithElement.kind;
    for (int i = 0; i < rank; ++i) {
      d->size[i] = stencil.size[(size_t)i]; d->center[i] = stencil.center[(size_t)i];
      d->pad_lo[i] = padding.paddingFromOrigin[(size_t)i] < 0 ? 0 : padding.paddingFromOrigin[(size_t)i];
      d->pad_hi[i] = padding.paddingFromBottom[(size_t)i] < 0 ? 0 : padding.paddingFromBottom[(size_t)i];
      d->stride[i] = stride ? ((*stride)[(size_t)i] < 1 ? 1 : (*stride)[(size_t)i]) : 1;  // Stride clamps to >= 1
    }
    std::memcpy(d->fill, &padding.paddingWithElement.fill, sizeof(E));
    d->coeffs = stencil.coeffs.empty() ? nullptr : stencil.coeffs.data();
    d->coeffs2 = stencil.coeffs2.empty() ? nullptr : stencil.coeffs2.data();
    if (stencil.kind == MB_TAPS) {
      d->ntaps = (int64_t)stencil.coeffs.size();
      d->tap_offsets = stencil.tap_offsets.data();
    }
    d->flags = stencil.flags;
    d->n_post = (uint32_t)stencil.post.size();
    d->post = stencil.post.empty() ? nullptr : stencil.post.data();
    if (stencil.kind == MB_EXPR) {
      term_store.assign(stencil.terms.size(), mb_term());
      for (size_t t = 0; t < stencil.terms.size(); ++t) {
        const Stencil<E> &ts = stencil.terms[t];
        mb_term &m = term_store[t];
        std::memset(&m, 0, sizeof(m));
        m.kind = ts.kind;
        for (int i = 0; i < rank; ++i) { m.size[i] = ts.size[(size_t)i]; m.center[i] = ts.center[(size_t)i]; }
        m.coeffs = ts.coeffs.empty() ? nullptr : ts.coeffs.data();
        if (ts.kind == MB_TAPS) { m.ntaps = (int64_t)ts.coeffs.size(); m.tap_offsets = ts.tap_offsets.data(); }
      }
      d->n_terms = (uint32_t)term_store.size();
      d->terms = term_store.data();
      d->n_program = (uint32_t)stencil.program.size();
      d->program = stencil.program.data();
    }
  }
};
template <typename E> DW<E> applyStencil(const Padding<E> &p, const Stencil<E> &st, const Array<E> &arr) { return {p, st, &arr}; }
template <typename E> DW<E> mapStencil(const Border<E> &b, const Stencil<E> &st, const Array<E> &arr) { return {samePadding(st, b), st, &arr}; }

// ---- the device (replaces the scheduler hand-off of withMassivScheduler_, Core/Loop.hs:468-474) --------
class Device {
 public:
  explicit Device(int ordinal = 0) {
    const int st = mb_ctx_create(ordinal, &ctx_);
    if (st) throw B200Error(st, "mb_ctx_create: no usable CUDA device (there is no CPU fallback)");
  }
  ~Device() { if (ctx_) mb_ctx_destroy(ctx_); }
  Device(const Device &) = delete;
  Device &operator=(const Device &) = delete;
  mb_ctx *handle() const { return ctx_; }
  std::string lastKernel() const { return mb_last_kernel(ctx_); }
  // gives the context's device scratch (staging arrays, padded iteration pairs) back to the driver
  void trim() const { check(mb_ctx_trim(ctx_)); }
  void check(int st) const {
    if (!st) return;
    const char *m = mb_last_error(ctx_);
    const std::string msg = std::string(mb_status_string(st)) + (m && *m ? std::string(": ") + m : std::string());
    if (st == MB_ERR_INDEX_ZERO) throw IndexZeroException(st, msg);
    if (st == MB_ERR_SIZE_MISMATCH) throw SizeMismatchException(st, msg);
    throw B200Error(st, msg);
  }
 private:
  mb_ctx *ctx_ = nullptr;
};

namespace detail {
template <typename E> Sz out_size(const DW<E> &dw, const mb_stencil_desc &d) {
  int64_t od[MB_MAX_RANK];
  const int st = mb_out_dims(&d, dw.arr->sz.dims.data(), od);
  if (st) throw B200Error(st, mb_status_string(st));
  return Sz(Ix(od, od + d.rank));
}
}  // namespace detail

// size of the array `compute` would produce (host only: usable without a GPU)
template <typename E> Sz outputSize(const DW<E> &dw, const Ix *stride = nullptr) {
  mb_stencil_desc d;
  dw.describe(&d, stride);
  return detail::out_size(dw, d);
}

// computeWithStride (Array/Manifest/Internal.hs:248-259)
template <typename E> Array<E> computeWithStride(const Device &dev, const Ix &stride, const DW<E> &dw) {
  mb_stencil_desc d;
  dw.describe(&d, &stride);
  Array<E> out(detail::out_size(dw, d));
  dev.check(mb_run(dev.handle(), &d, dw.arr->sz.dims.data(), dw.arr->elems.data(), out.elems.data()));
  return out;
}
// compute (Array/Manifest/Internal.hs:65-67)
template <typename E> Array<E> compute(const Device &dev, const DW<E> &dw) {
  mb_stencil_desc d;
  dw.describe(&d, nullptr);
  Array<E> out(detail::out_size(dw, d));
  dev.check(mb_run(dev.handle(), &d, dw.arr->sz.dims.data(), dw.arr->elems.data(), out.elems.data()));
  return out;
}
// iterateUntil (\n _ _ -> n >= k - 1) (\_ -> mapStencil border stencil) arr   (Array/Manifest/Internal.hs:331-397)
template <typename E> Array<E> iterateStencil(const Device &dev, int64_t k, const Border<E> &b, const Stencil<E> &st, const Array<E> &arr) {
  const DW<E> dw = mapStencil(b, st, arr);
  mb_stencil_desc d;
  dw.describe(&d, nullptr);
  Array<E> out(arr.sz);
  dev.check(mb_iterate(dev.handle(), &d, arr.sz.dims.data(), arr.elems.data(), out.elems.data(), k));
  return out;
}
// iterateUntil convergence (\_ -> mapStencil border stencil) arr with a predicate the device can evaluate
// (Array/Manifest/Internal.hs:331-349): Equal = \_ prev cur -> cur == prev; MaxAbsDiff tol = maximum |cur - prev| < tol;
// checked after every checkEvery-th application, at most maxSteps applications
struct Until {
  mb_until predicate;
  double tol;
  static Until Equal() { return {MB_UNTIL_EQUAL, 0.0}; }
  static Until MaxAbsDiff(double t) { return {MB_UNTIL_MAX_ABS_DIFF, t}; }
};
template <typename E> struct UntilResult {
  Array<E> array;
  int64_t steps;
  bool converged;
};
template <typename E>
UntilResult<E> iterateUntil(const Device &dev, const Until &u, int64_t maxSteps, int64_t checkEvery, const Border<E> &b, const Stencil<E> &st,
                            const Array<E> &arr) {
  const DW<E> dw = mapStencil(b, st, arr);
  mb_stencil_desc d;
  dw.describe(&d, nullptr);
  UntilResult<E> r{Array<E>(arr.sz), 0, false};
  int32_t conv = 0;
  dev.check(mb_iterate_until(dev.handle(), &d, arr.sz.dims.data(), arr.elems.data(), r.array.elems.data(), maxSteps, checkEvery, u.predicate,
                             u.tol, &r.steps, &conv));
  r.converged = conv != 0;
  return r;
}
// integrateWith (Array/Numeric/Integral.hs:121-135): computeWithStride (Stride nsz) $ mapStencil (Fill 0) (stencil dim n) arr
template <typename E, typename F> Array<E> integrateWith(const Device &dev, F stencil, int dim, int n, const Array<E> &arr) {
  const int rank = arr.sz.rank();
  if (dim < 1 || dim > rank) throw std::out_of_range("IndexDimensionException");
  Ix nsz((size_t)rank, 1);
  nsz[(size_t)(rank - dim)] = n;
  return computeWithStride(dev, nsz, mapStencil(Border<E>::Fill(E(0)), stencil(dim, n), arr));
}

}  // namespace massiv_b200

#endif  // MASSIV_B200_HPP
