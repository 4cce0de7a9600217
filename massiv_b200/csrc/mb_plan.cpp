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
  switch (d->kind) {
    case MB_ID:
      for (int i = 0; i < d->rank; ++i) if (d->size[i] != 1) return bad(MB_ERR_INVALID_DESC, "idStencil has size 1");
      break;
    case MB_SUM: case MB_PRODUCT:
      if (d->elem == MB_BOOL32) return bad(MB_ERR_UNSUPPORTED, "Bool is not Num");
      break;
    case MB_AVG:
      if (!is_float(d->elem)) return bad(MB_ERR_UNSUPPORTED, "avgStencil needs Fractional e");
      break;
    case MB_MAX: case MB_MIN:
      if (is_float(d->elem)) return bad(MB_ERR_UNSUPPORTED, "max/minStencil need Bounded e (no Float/Double instance)");
      break;
    case MB_CONV: case MB_CORR:
      if (d->elem == MB_BOOL32) return bad(MB_ERR_UNSUPPORTED, "Bool is not Num");
      if (!d->coeffs) return bad(MB_ERR_INVALID_DESC, "missing kernel array");
      for (int i = 0; i < d->rank; ++i) if (d->size[i] < 1) return bad(MB_ERR_UNSUPPORTED, "empty kernel array");
      break;
    case MB_MAG2:
      if (!is_float(d->elem)) return bad(MB_ERR_UNSUPPORTED, "sqrt needs Floating e");
      if (!d->coeffs || !d->coeffs2) return bad(MB_ERR_INVALID_DESC, "missing kernel array");
      for (int i = 0; i < d->rank; ++i) if (d->size[i] < 1) return bad(MB_ERR_UNSUPPORTED, "empty kernel array");
      break;
    case MB_TAPS:
      if (d->elem == MB_BOOL32) return bad(MB_ERR_UNSUPPORTED, "Bool is not Num");
      if (d->ntaps < 0) return bad(MB_ERR_INVALID_DESC, "negative tap count");
      if (d->ntaps > 0 && (!d->coeffs || !d->tap_offsets)) return bad(MB_ERR_INVALID_DESC, "missing tap list");
      break;
    case MB_LIFE:
      if (d->elem != MB_U8 || d->rank != 2 || d->size[0] != 3 || d->size[1] != 3)
        return bad(MB_ERR_UNSUPPORTED, "lifeStencil is a 3x3 Word8 stencil");
      break;
    case MB_INT_MIDPOINT: case MB_INT_TRAPEZOID: case MB_INT_SIMPSON: {
      if (!is_float(d->elem)) return bad(MB_ERR_UNSUPPORTED, "integral rules need Fractional e");
      if (!d->coeffs) return bad(MB_ERR_INVALID_DESC, "missing dx");
      int along = 0;
      int64_t len = 1;
      for (int i = 0; i < d->rank; ++i) {
        if (d->size[i] < 1) return bad(MB_ERR_UNSUPPORTED, "empty integration stencil");
        if (d->size[i] != 1) { ++along; len = d->size[i]; }
      }
      if (along > 1) return bad(MB_ERR_INVALID_DESC, "an integration stencil runs along one dimension");
      // Simpson's rule: `odd n` is an error in the reference; n = 0 would read outside its own window
      if (d->kind == MB_INT_SIMPSON && (len < 3 || (len - 1) % 2 != 0))
        return bad(MB_ERR_UNSUPPORTED, "Number of sample points for Simpson's rule stencil should be even");
      break;
    }
    default: return bad(MB_ERR_INVALID_DESC, "unknown stencil kind");
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

int plan_geometry(const mb_stencil_desc *d, const int64_t *in_dims, Plan *p, std::string *err) {
  int st = validate(d, err);
  if (st) return st;
  p->kind = d->kind; p->elem = d->elem; p->rank = d->rank; p->border = d->border; p->flags = d->flags;
  p->esize = elem_size(d->elem);
  std::memcpy(p->fill, d->fill, 8);
  if (d->n_post) p->post.assign(d->post, d->post + d->n_post);
  const bool mag_corr = d->kind == MB_MAG2 && (d->flags & MB_FLAG_MAG2_CORR);
  p->in_elems = 1; p->out_elems = 1;
  for (int i = 0; i < d->rank; ++i) {
    const int64_t s = d->size[i];
    int64_t g = s, c = 0;
    switch (d->kind) {
      case MB_AVG: g = s > 1 ? s : 1; break;  // liftA2 (/) with `pure n`: Stencil/Internal.hs:83-90
      case MB_CONV: c = s - 1 - s / 2; break;  // Stencil/Convolution.hs:71-73
      case MB_MAG2: c = mag_corr ? floor_div(s, 2) : s - 1 - s / 2; break;
      case MB_CORR: c = floor_div(s, 2); break;  // Stencil/Convolution.hs:114
      case MB_TAPS: c = d->center[i]; break;
      case MB_LIFE: g = 3; c = 1; break;
      default: break;  // fold stencils: centre 0 (Stencil.hs:298-302)
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

int make_plan(const mb_stencil_desc *d, const int64_t *in_dims, Plan *p, std::string *err) {
  int st = plan_geometry(d, in_dims, p, err);
  if (st) return st;
  const int r = d->rank;
  int64_t n = 1;
  if (d->kind == MB_TAPS) n = d->ntaps;
  else if (d->kind == MB_LIFE) n = 9;
  else for (int i = 0; i < r; ++i) n *= d->size[i];
  if (n > (int64_t)1 << 24) { if (err) *err = "stencil window too large"; return MB_ERR_UNSUPPORTED; }
  p->ntaps = n;
  p->avg_n = n;
  p->tap_off.assign((size_t)n * r, 0);
  const bool corr = d->kind == MB_CORR || (d->kind == MB_MAG2 && (d->flags & MB_FLAG_MAG2_CORR));
  const bool conv = d->kind == MB_CONV || (d->kind == MB_MAG2 && !corr);
  if (d->kind == MB_TAPS) {
    for (int64_t t = 0; t < n; ++t)
      for (int i = 0; i < r; ++i) {
        const int64_t o = d->tap_offsets[t * r + i];
        if (o < -p->center[i] || o >= p->geom[i] - p->center[i]) {
          if (err) *err = "tap offset outside the stencil box";  // the reference raises at stencil creation
          return MB_ERR_INVALID_DESC;
        }
        p->tap_off[t * r + i] = (int32_t)o;
      }
  } else if (d->kind == MB_LIFE) {
    // neighbours first (summed mod 256), centre last
    static const int lo[9][2] = {{-1,-1},{-1,0},{-1,1},{0,-1},{0,1},{1,-1},{1,0},{1,1},{0,0}};
    for (int t = 0; t < 9; ++t) { p->tap_off[t * 2] = lo[t][0]; p->tap_off[t * 2 + 1] = lo[t][1]; }
  } else {
    p->dense = true; p->descending = conv;
    int64_t q[MB_MAX_RANK] = {0};
    for (int64_t t = 0; t < n; ++t) {  // row-major over [0,size): iter, Core/Index.hs:556-573
      for (int i = 0; i < r; ++i) {
        const int64_t s = d->size[i];
        int64_t o = q[i];                                  // fold: get (ix + q)
        if (conv) o = s / 2 - q[i];                        // ix + sCenter - kIx
        else if (corr) o = q[i] - floor_div(s, 2);         // ix - sCenter + kIx
        p->tap_off[t * r + i] = (int32_t)o;
      }
      for (int i = r - 1; i >= 0; --i) { if (++q[i] < d->size[i]) break; q[i] = 0; }
    }
  }
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
