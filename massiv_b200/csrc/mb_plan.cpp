// Host-side normalisation of a stencil descriptor: validation, geometry, ordered tap list.
// Restates (does not copy) the reference's size arithmetic and tap orders; see the citations.
#include "mb_internal.h"

namespace mb {

static inline int64_t floor_mod(int64_t a, int64_t k) {
  int64_t m = a % k;
  return (m != 0 && ((m < 0) != (k < 0))) ? m + k : m;
}
static inline int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b;
  return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// handleBorderIndex, one dimension (Core/Index.hs:236-253; repairIndex Core/Index/Internal.hs:1030-1035)
int border_index(int border, int64_t k, int64_t i, int64_t *out, int32_t *is_fill) {
  *is_fill = 0;
  if (border == MB_FILL) {
    if (i < 0 || i >= k) { *is_fill = 1; *out = 0; } else { *out = i; }
    return MB_OK;
  }
  if (k <= 0) return MB_ERR_INDEX_ZERO;
  if (i >= 0 && i < k) { *out = i; return MB_OK; }
  switch (border) {
    case MB_WRAP: *out = floor_mod(i, k); break;
    case MB_EDGE: *out = i < 0 ? 0 : k - 1; break;
    case MB_REFLECT: *out = floor_mod(-i - 1, k); break;
    case MB_CONTINUE: *out = i < 0 ? floor_mod(-i, k) : floor_mod(-i - 2, k); break;
    default: return MB_ERR_INVALID_DESC;
  }
  return MB_OK;
}

static bool is_float(int e) { return e == MB_F32 || e == MB_F64; }

static int elem_size(int e) {
  switch (e) {
    case MB_U8: case MB_I8: return 1;
    case MB_U16: case MB_I16: return 2;
    case MB_U32: case MB_I32: case MB_F32: case MB_BOOL32: return 4;
    case MB_U64: case MB_I64: case MB_F64: return 8;
  }
  return 0;
}

// what one built-in constructor needs of its arguments and of the element type (the class constraints of the
// reference: Num / Fractional / Bounded / Floating)
static int validate_kind(int kind, int elem, int rank, const int64_t *size, const void *coeffs, const void *coeffs2,
                         int64_t ntaps, const int64_t *tap_offsets, std::string *err) {
  auto bad = [&](int st, const char *m) { if (err) *err = m; return st; };
  for (int i = 0; i < rank; ++i)
    if (size[i] < 0) return bad(MB_ERR_INVALID_DESC, "negative stencil size");
  switch (kind) {
    case MB_ID:
      for (int i = 0; i < rank; ++i) if (size[i] != 1) return bad(MB_ERR_INVALID_DESC, "idStencil has size 1");
      break;
    case MB_SUM: case MB_PRODUCT:
      if (elem == MB_BOOL32) return bad(MB_ERR_UNSUPPORTED, "Bool is not Num");
      break;
    case MB_AVG:
      if (!is_float(elem)) return bad(MB_ERR_UNSUPPORTED, "avgStencil needs Fractional e");
      break;
    case MB_MAX: case MB_MIN:
      if (is_float(elem)) return bad(MB_ERR_UNSUPPORTED, "max/minStencil need Bounded e (no Float/Double instance)");
      break;
    case MB_CONV: case MB_CORR:
      if (elem == MB_BOOL32) return bad(MB_ERR_UNSUPPORTED, "Bool is not Num");
      if (!coeffs) return bad(MB_ERR_INVALID_DESC, "missing kernel array");
      for (int i = 0; i < rank; ++i) if (size[i] < 1) return bad(MB_ERR_UNSUPPORTED, "empty kernel array");
      break;
    case MB_MAG2:
      if (!is_float(elem)) return bad(MB_ERR_UNSUPPORTED, "sqrt needs Floating e");
      if (!coeffs || !coeffs2) return bad(MB_ERR_INVALID_DESC, "missing kernel array");
      for (int i = 0; i < rank; ++i) if (size[i] < 1) return bad(MB_ERR_UNSUPPORTED, "empty kernel array");
      break;
    case MB_TAPS:
      if (elem == MB_BOOL32) return bad(MB_ERR_UNSUPPORTED, "Bool is not Num");
      if (ntaps < 0) return bad(MB_ERR_INVALID_DESC, "negative tap count");
      if (ntaps > 0 && (!coeffs || !tap_offsets)) return bad(MB_ERR_INVALID_DESC, "missing tap list");
      break;
    case MB_LIFE:
      if (elem != MB_U8 || rank != 2 || size[0] != 3 || size[1] != 3)
        return bad(MB_ERR_UNSUPPORTED, "lifeStencil is a 3x3 Word8 stencil");
      break;
    case MB_INT_MIDPOINT: case MB_INT_TRAPEZOID: case MB_INT_SIMPSON: {
      if (!is_float(elem)) return bad(MB_ERR_UNSUPPORTED, "integral rules need Fractional e");
      if (!coeffs) return bad(MB_ERR_INVALID_DESC, "missing dx");
      int along = 0;
      int64_t len = 1;
      for (int i = 0; i < rank; ++i) {
        if (size[i] < 1) return bad(MB_ERR_UNSUPPORTED, "empty integration stencil");
        if (size[i] != 1) { ++along; len = size[i]; }
      }
      if (along > 1) return bad(MB_ERR_INVALID_DESC, "an integration stencil runs along one dimension");
      // Simpson's rule: `odd n` is an error in the reference; n = 0 would read outside its own window
      if (kind == MB_INT_SIMPSON && (len < 3 || (len - 1) % 2 != 0))
        return bad(MB_ERR_UNSUPPORTED, "Number of sample points for Simpson's rule stencil should be even");
      break;
    }
    default: return bad(MB_ERR_INVALID_DESC, "unknown stencil kind");
  }
  return MB_OK;
}

static bool term_kind_ok(int kind) { return kind >= MB_ID && kind <= MB_TAPS; }

// the postfix program of an MB_EXPR descriptor: operators the element type has, a stack that never underflows
// nor exceeds MB_MAX_STACK, exactly one value left
static int validate_program(const mb_stencil_desc *d, std::string *err) {
  auto bad = [&](int st, const char *m) { if (err) *err = m; return st; };
  if (d->n_terms < 1 || d->n_terms > MB_MAX_TERMS || !d->terms) return bad(MB_ERR_INVALID_DESC, "an expression needs 1..8 terms");
  if (d->n_program < 1 || d->n_program > MB_MAX_PROGRAM || !d->program) return bad(MB_ERR_INVALID_DESC, "an expression needs a program of 1..32 steps");
  if (d->elem == MB_BOOL32) return bad(MB_ERR_UNSUPPORTED, "Bool is not Num");
  int depth = 0;
  for (uint32_t i = 0; i < d->n_program; ++i) {
    const mb_xins &x = d->program[i];
    switch (x.op) {
      case MB_X_TERM:
        if (x.arg < 0 || (uint32_t)x.arg >= d->n_terms) return bad(MB_ERR_INVALID_DESC, "program refers to a term that does not exist");
        ++depth;
        break;
      case MB_X_CONST: ++depth; break;
      case MB_X_DIV:
        if (!is_float(d->elem)) return bad(MB_ERR_UNSUPPORTED, "(/) needs Fractional e");
        // fall through
      case MB_X_ADD: case MB_X_SUB: case MB_X_MUL: case MB_X_MIN: case MB_X_MAX:
        if (depth < 2) return bad(MB_ERR_INVALID_DESC, "program stack underflow");
        --depth;
        break;
      case MB_X_RECIP: case MB_X_SQRT:
        if (!is_float(d->elem)) return bad(MB_ERR_UNSUPPORTED, "recip / sqrt need Fractional / Floating e");
        // fall through
      case MB_X_NEGATE: case MB_X_ABS: case MB_X_SIGNUM: case MB_X_SQUARE:
        if (depth < 1) return bad(MB_ERR_INVALID_DESC, "program stack underflow");
        break;
      default: return bad(MB_ERR_INVALID_DESC, "unknown program step");
    }
    if (depth > MB_MAX_STACK) return bad(MB_ERR_UNSUPPORTED, "program stack deeper than MB_MAX_STACK");
  }
  if (depth != 1) return bad(MB_ERR_INVALID_DESC, "a program must leave exactly one value");
  return MB_OK;
}

static int validate(const mb_stencil_desc *d, std::string *err) {
  auto bad = [&](int st, const char *m) { if (err) *err = m; return st; };
  if (!d) return bad(MB_ERR_INVALID_DESC, "null descriptor");
  if (d->rank < 1 || d->rank > MB_MAX_RANK) return bad(MB_ERR_INVALID_DESC, "rank must be 1..5");
  if (elem_size(d->elem) == 0) return bad(MB_ERR_INVALID_DESC, "unknown element type");
  if (d->border < MB_FILL || d->border > MB_CONTINUE) return bad(MB_ERR_INVALID_DESC, "unknown border");
  for (int i = 0; i < d->rank; ++i) {
    if (d->size[i] < 0) return bad(MB_ERR_INVALID_DESC, "negative stencil size");
    if (d->pad_lo[i] < 0 || d->pad_hi[i] < 0) return bad(MB_ERR_INVALID_DESC, "negative padding (Sz is non-negative)");
    if (d->stride[i] < 1) return bad(MB_ERR_INVALID_DESC, "stride must be >= 1");
  }
  int st;
  if (d->kind == MB_EXPR) {
    if ((st = validate_program(d, err))) return st;
    for (uint32_t t = 0; t < d->n_terms; ++t) {
      const mb_term &tm = d->terms[t];
      if (!term_kind_ok(tm.kind)) return bad(MB_ERR_UNSUPPORTED, "an expression combines fold, convolution, correlation and tap-list stencils");
      if ((st = validate_kind(tm.kind, d->elem, d->rank, tm.size, tm.coeffs, nullptr, tm.ntaps, tm.tap_offsets, err))) return st;
    }
  } else if ((st = validate_kind(d->kind, d->elem, d->rank, d->size, d->coeffs, d->coeffs2, d->ntaps, d->tap_offsets, err))) {
    return st;
  }
  if (d->n_post > MB_MAX_POST) return bad(MB_ERR_UNSUPPORTED, "too many post-maps");
  if (d->n_post > 0 && !d->post) return bad(MB_ERR_INVALID_DESC, "missing post-map list");
  for (uint32_t i = 0; i < d->n_post; ++i) {
    const int op = d->post[i].op;
    if (op < MB_POST_NEGATE || op > MB_POST_SQRT) return bad(MB_ERR_INVALID_DESC, "unknown post-map");
    if (op >= MB_POST_DIV && !is_float(d->elem)) return bad(MB_ERR_UNSUPPORTED, "post-map needs Fractional / Floating e");
    if (d->elem == MB_BOOL32) return bad(MB_ERR_UNSUPPORTED, "Bool is not Num");
  }
  return MB_OK;
}

// size used for the output-size arithmetic and stencilCenter of one built-in constructor, one dimension
static void kind_geom(int kind, uint32_t flags, int64_t s, int64_t given_center, int64_t *g, int64_t *c) {
  const bool mag_corr = kind == MB_MAG2 && (flags & MB_FLAG_MAG2_CORR);
  *g = s; *c = 0;
  switch (kind) {
    case MB_AVG: *g = s > 1 ? s : 1; break;  // liftA2 (/) with `pure n`: Stencil/Internal.hs:83-90
    case MB_CONV: *c = s - 1 - s / 2; break;  // Stencil/Convolution.hs:71-73
    case MB_MAG2: *c = mag_corr ? floor_div(s, 2) : s - 1 - s / 2; break;
    case MB_CORR: *c = floor_div(s, 2); break;  // Stencil/Convolution.hs:114
    case MB_TAPS: *c = given_center; break;
    case MB_LIFE: *g = 3; *c = 1; break;
    default: break;  // fold stencils: centre 0 (Stencil.hs:298-302)
  }
}

int plan_geometry(const mb_stencil_desc *d, const int64_t *in_dims, Plan *p, std::string *err) {
  int st = validate(d, err);
  if (st) return st;
  p->kind = d->kind; p->elem = d->elem; p->rank = d->rank; p->border = d->border; p->flags = d->flags;
  p->esize = elem_size(d->elem);
  std::memcpy(p->fill, d->fill, 8);
  if (d->n_post) p->post.assign(d->post, d->post + d->n_post);
  p->in_elems = 1; p->out_elems = 1;
  for (int i = 0; i < d->rank; ++i) {
    const int64_t s = d->size[i];
    int64_t g = s, c = 0;
    if (d->kind == MB_EXPR) {
      // unionStencilCenters / unionStencilSizes (Stencil/Internal.hs:83-90), folded over the terms; `pure c`
      // (size 1, centre 0) never widens a stencil of positive size
      int64_t mc = 0, ext = 1;
      for (uint32_t t = 0; t < d->n_terms; ++t) {
        int64_t tg, tc;
        kind_geom(d->terms[t].kind, 0, d->terms[t].size[i], d->terms[t].center[i], &tg, &tc);
        if (d->terms[t].center[i] != tc) { if (err) *err = "a term's centre does not match its constructor"; return MB_ERR_INVALID_DESC; }
        if (tc > mc) mc = tc;
        if (tg - tc > ext) ext = tg - tc;
      }
      c = mc; g = mc + ext;
      if (s != g) { if (err) *err = "expression size is not the union of its terms' sizes"; return MB_ERR_INVALID_DESC; }
    } else {
      kind_geom(d->kind, d->flags, s, d->center[i], &g, &c);
    }
    if (d->center[i] != c) { if (err) *err = "stencil centre does not match its constructor"; return MB_ERR_INVALID_DESC; }
    if (in_dims[i] < 0) { if (err) *err = "negative array extent"; return MB_ERR_INVALID_DESC; }
    p->size[i] = s; p->geom[i] = g; p->center[i] = c;
    p->pad_lo[i] = d->pad_lo[i]; p->pad_hi[i] = d->pad_hi[i];
    p->in_dims[i] = in_dims[i]; p->stride[i] = d->stride[i];
    p->shift[i] = c - d->pad_lo[i];                                      // Stencil.hs:200
    const int64_t shrink = g - 1 > 0 ? g - 1 : 0;                        // Stencil.hs:207
    int64_t full = d->pad_lo[i] + d->pad_hi[i] + in_dims[i] - shrink;    // Stencil.hs:208
    if (full < 0) full = 0;
    p->full_dims[i] = full;
    p->out_dims[i] = floor_div(full - 1, d->stride[i]) + 1;              // Stride.hs:102-105
    p->in_elems *= in_dims[i]; p->out_elems *= p->out_dims[i];
  }
  return MB_OK;
}

// The ordered tap list of one built-in constructor, appended to `out` (n taps x rank source offsets relative to the
// evaluated index).  fold: Stencil.hs:298-302; FromKernel: Stencil/Convolution.hs:64-80, 107-121; TAPS: as given.
static int kind_taps(int kind, uint32_t flags, int r, const int64_t *size, const int64_t *center, const int64_t *geom,
                     int64_t ntaps_in, const int64_t *tap_offsets, std::vector<int32_t> *out, int64_t *n_out, std::string *err) {
  int64_t n = 1;
  if (kind == MB_TAPS) n = ntaps_in;
  else if (kind == MB_LIFE) n = 9;
  else for (int i = 0; i < r; ++i) n *= size[i];
  if (n > (int64_t)1 << 24) { if (err) *err = "stencil window too large"; return MB_ERR_UNSUPPORTED; }
  *n_out = n;
  const size_t base = out->size();
  out->resize(base + (size_t)n * r, 0);
  int32_t *tap = out->data() + base;
  const bool corr = kind == MB_CORR || (kind == MB_MAG2 && (flags & MB_FLAG_MAG2_CORR));
  const bool conv = kind == MB_CONV || (kind == MB_MAG2 && !corr);
  if (kind == MB_TAPS) {
    for (int64_t t = 0; t < n; ++t)
      for (int i = 0; i < r; ++i) {
        const int64_t o = tap_offsets[t * r + i];
        if (o < -center[i] || o >= geom[i] - center[i]) {
          if (err) *err = "tap offset outside the stencil box";  // the reference raises at stencil creation
          return MB_ERR_INVALID_DESC;
        }
        tap[t * r + i] = (int32_t)o;
      }
  } else if (kind == MB_LIFE) {
    // neighbours first (summed mod 256), centre last
    static const int lo[9][2] = {{-1,-1},{-1,0},{-1,1},{0,-1},{0,1},{1,-1},{1,0},{1,1},{0,0}};
    for (int t = 0; t < 9; ++t) { tap[t * 2] = lo[t][0]; tap[t * 2 + 1] = lo[t][1]; }
  } else {
    int64_t q[MB_MAX_RANK] = {0};
    for (int64_t t = 0; t < n; ++t) {  // row-major over [0,size): iter, Core/Index.hs:556-573
      for (int i = 0; i < r; ++i) {
        const int64_t s = size[i];
        int64_t o = q[i];                                  // fold: get (ix + q)
        if (conv) o = s / 2 - q[i];                        // ix + sCenter - kIx
        else if (corr) o = q[i] - floor_div(s, 2);         // ix - sCenter + kIx
        tap[t * r + i] = (int32_t)o;
      }
      for (int i = r - 1; i >= 0; --i) { if (++q[i] < size[i]) break; q[i] = 0; }
    }
  }
  return MB_OK;
}

int make_plan(const mb_stencil_desc *d, const int64_t *in_dims, Plan *p, std::string *err) {
  int st = plan_geometry(d, in_dims, p, err);
  if (st) return st;
  const int r = d->rank;
  int64_t n = 0;
  if (d->kind == MB_EXPR) {
    for (uint32_t t = 0; t < d->n_terms; ++t) {
      const mb_term &tm = d->terms[t];
      int64_t tg[MB_MAX_RANK], tc[MB_MAX_RANK], tn = 0, total = 1;
      for (int i = 0; i < r; ++i) {
        if (tm.size[i] < 1) { if (err) *err = "terms of an expression need a positive size"; return MB_ERR_UNSUPPORTED; }
        kind_geom(tm.kind, 0, tm.size[i], tm.center[i], &tg[i], &tc[i]);
        total *= tm.size[i];
      }
      if ((st = kind_taps(tm.kind, 0, r, tm.size, tc, tg, tm.ntaps, tm.tap_offsets, &p->tap_off, &tn, err))) return st;
      Plan::Term pt;
      pt.kind = tm.kind; pt.begin = (int32_t)n; pt.ntaps = (int32_t)tn; pt.avg_n = (int32_t)total;
      p->terms.push_back(pt);
      // coefficients travel with the taps (zeros where a term has none) so that one index serves both
      const bool has_k = tm.kind == MB_CONV || tm.kind == MB_CORR || tm.kind == MB_TAPS;
      const size_t at = p->c1.size();
      p->c1.resize(at + (size_t)tn * p->esize, 0);
      if (has_k && tn > 0) std::memcpy(p->c1.data() + at, tm.coeffs, (size_t)tn * p->esize);
      n += tn;
      if (n > (int64_t)1 << 20) { if (err) *err = "stencil window too large"; return MB_ERR_UNSUPPORTED; }
    }
    p->prog.assign(d->program, d->program + d->n_program);
    p->avg_n = 0;
  } else {
    if ((st = kind_taps(d->kind, d->flags, r, d->size, p->center, p->geom, d->ntaps, d->tap_offsets, &p->tap_off, &n, err))) return st;
    p->avg_n = n;
    const bool corr = d->kind == MB_CORR || (d->kind == MB_MAG2 && (d->flags & MB_FLAG_MAG2_CORR));
    const bool conv = d->kind == MB_CONV || (d->kind == MB_MAG2 && !corr);
    if (d->kind != MB_TAPS && d->kind != MB_LIFE) { p->dense = true; p->descending = conv; }
  }
  p->ntaps = n;
  for (int64_t t = 0; t < n; ++t)
    for (int i = 0; i < r; ++i) {
      const int64_t o = p->tap_off[t * r + i];
      if (t == 0 || o < p->min_off[i]) p->min_off[i] = o;
      if (t == 0 || o > p->max_off[i]) p->max_off[i] = o;
    }
  const bool is_rule = d->kind == MB_INT_MIDPOINT || d->kind == MB_INT_TRAPEZOID || d->kind == MB_INT_SIMPSON;
  if (is_rule) p->c1.assign((const uint8_t *)d->coeffs, (const uint8_t *)d->coeffs + (size_t)p->esize);  // dx
  const bool has_k = d->kind == MB_CONV || d->kind == MB_CORR || d->kind == MB_TAPS || d->kind == MB_MAG2;
  if (has_k && n > 0) {
    p->c1.assign((const uint8_t *)d->coeffs, (const uint8_t *)d->coeffs + (size_t)n * p->esize);
    if (d->kind == MB_MAG2)
      p->c2.assign((const uint8_t *)d->coeffs2, (const uint8_t *)d->coeffs2 + (size_t)n * p->esize);
  }
  // IndexZeroException: a non-Fill border forced against a zero extent (Core/Index/Internal.hs:1031)
  bool zero_in = false;
  for (int i = 0; i < r; ++i) if (in_dims[i] == 0) zero_in = true;
  if (p->out_elems > 0 && n > 0 && zero_in && d->border != MB_FILL) {
    if (err) *err = "IndexZeroException: border resolution on an empty dimension";
    return MB_ERR_INDEX_ZERO;
  }
  return MB_OK;
}

}  // namespace mb
