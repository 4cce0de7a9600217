/*
 * massiv_oracle.c — CPU ORACLE (test infrastructure; see massiv_oracle.h for the scope rule and
 * the reference file:line each piece restates).  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fno-fast-math, pthreads).
 */
#include "massiv_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* Haskell `mod` for Int (floored) — modInt, used by handleBorderIndex (Core/Index.hs:243-253) */
static inline int64_t floor_mod(int64_t a, int64_t k) {
  int64_t m = a % k;
  return (m != 0 && ((m < 0) != (k < 0))) ? m + k : m;
}
static inline int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b;
  return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q;
}

/* Core/Index/Internal.hs:1030-1035 repairIndex (Int instance) + Core/Index.hs:236-253.
 * The repair function is applied ONCE; for |offset| > k, Reflect/Continue are the mod formulae,
 * not a true 2k-periodic mirror. */
int mo_border_index(int32_t border, int64_t k, int64_t i, int64_t *out, int32_t *fill) {
  *fill = 0;
  if (border == MO_FILL) {
    if (i < 0 || i >= k) { *fill = 1; *out = 0; } else *out = i;
    return MO_OK;
  }
  if (k <= 0) return MO_ERR_INDEX_ZERO; /* throwIndexZeroException */
  if (i >= 0 && i < k) { *out = i; return MO_OK; }
  switch (border) {
    case MO_WRAP: *out = floor_mod(i, k); break;
    case MO_EDGE: *out = (i < 0) ? 0 : k - 1; break;
    case MO_REFLECT: *out = floor_mod(-i - 1, k); break;
    case MO_CONTINUE: *out = (i < 0) ? floor_mod(-i, k) : floor_mod(-i - 2, k); break;
    default: return MO_ERR_INVALID_DESC;
  }
  return MO_OK;
}

int mo_elem_size(int32_t elem) {
  switch (elem) {
    case MO_U8: case MO_I8: return 1;
    case MO_U16: case MO_I16: return 2;
    case MO_U32: case MO_I32: case MO_F32: case MO_BOOL32: return 4;
    case MO_U64: case MO_I64: case MO_F64: return 8;
  }
  return 0;
}

static int is_float(int32_t e) { return e == MO_F32 || e == MO_F64; }

static int validate(const mo_desc *d) {
  if (d->rank < 1 || d->rank > MO_MAX_RANK) return MO_ERR_INVALID_DESC;
  if (mo_elem_size(d->elem) == 0) return MO_ERR_INVALID_DESC;
  if (d->border < MO_FILL || d->border > MO_CONTINUE) return MO_ERR_INVALID_DESC;
  for (int i = 0; i < d->rank; ++i) {
    if (d->size[i] < 0 || d->pad_lo[i] < 0 || d->pad_hi[i] < 0 || d->stride[i] < 1)
      return MO_ERR_INVALID_DESC;
  }
  switch (d->kind) {
    case MO_ID:
      for (int i = 0; i < d->rank; ++i) if (d->size[i] != 1) return MO_ERR_INVALID_DESC;
      break;
    case MO_SUM: case MO_PRODUCT:
      if (d->elem == MO_BOOL32) return MO_ERR_UNSUPPORTED;
      break;
    case MO_AVG:
      if (!is_float(d->elem)) return MO_ERR_UNSUPPORTED; /* Fractional e */
      break;
    case MO_MAX: case MO_MIN:
      if (is_float(d->elem)) return MO_ERR_UNSUPPORTED; /* Bounded e: no Float/Double instance */
      break;
    case MO_CONV: case MO_CORR:
      if (d->elem == MO_BOOL32 || !d->coeffs) return MO_ERR_UNSUPPORTED;
      for (int i = 0; i < d->rank; ++i) if (d->size[i] < 1) return MO_ERR_UNSUPPORTED;
      break;
    case MO_MAG2:
      if (!is_float(d->elem) || !d->coeffs || !d->coeffs2) return MO_ERR_UNSUPPORTED;
      for (int i = 0; i < d->rank; ++i) if (d->size[i] < 1) return MO_ERR_UNSUPPORTED;
      break;
    case MO_TAPS:
      if (d->elem == MO_BOOL32 || d->ntaps < 0) return MO_ERR_UNSUPPORTED;
      if (d->ntaps > 0 && (!d->coeffs || !d->tap_offsets)) return MO_ERR_INVALID_DESC;
      break;
    case MO_LIFE:
      if (d->elem != MO_U8 || d->rank != 2 || d->size[0] != 3 || d->size[1] != 3)
        return MO_ERR_UNSUPPORTED;
      break;
    case MO_INT_MIDPOINT: case MO_INT_TRAPEZOID: case MO_INT_SIMPSON: {
      /* Integral.hs:62-63: Sz (setDim' (pureIndex 1) dim k) — size 1 everywhere but along `dim` */
      if (!is_float(d->elem)) return MO_ERR_UNSUPPORTED; /* Fractional e */
      if (!d->coeffs) return MO_ERR_INVALID_DESC;
      int along = 0; int64_t len = 1;
      for (int i = 0; i < d->rank; ++i) {
        if (d->size[i] < 1) return MO_ERR_UNSUPPORTED;
        if (d->size[i] != 1) { ++along; len = d->size[i]; }
      }
      if (along > 1) return MO_ERR_INVALID_DESC;
      /* Integral.hs:106-109: odd n is an error; n = 0 would index outside the stencil's own window */
      if (d->kind == MO_INT_SIMPSON && (len < 3 || (len - 1) % 2 != 0)) return MO_ERR_UNSUPPORTED;
      break;
    }
    case MO_EXPR: {
      if (d->elem == MO_BOOL32) return MO_ERR_UNSUPPORTED;
      if (d->n_terms < 1 || d->n_terms > MO_MAX_TERMS || !d->terms) return MO_ERR_INVALID_DESC;
      if (d->n_program < 1 || d->n_program > MO_MAX_PROGRAM || !d->program) return MO_ERR_INVALID_DESC;
      for (uint32_t t = 0; t < d->n_terms; ++t) {
        mo_desc td = *d; /* same element type, rank, border; the term's own constructor arguments */
        td.kind = d->terms[t].kind;
        if (td.kind < MO_ID || td.kind > MO_TAPS) return MO_ERR_UNSUPPORTED;
        td.coeffs = d->terms[t].coeffs; td.coeffs2 = NULL; td.ntaps = d->terms[t].ntaps; td.tap_offsets = d->terms[t].tap_offsets;
        for (int i = 0; i < d->rank; ++i) {
          td.size[i] = d->terms[t].size[i];
          if (td.size[i] < 1) return MO_ERR_UNSUPPORTED;
        }
        int st = validate(&td);
        if (st) return st;
      }
      int depth = 0;
      for (uint32_t i = 0; i < d->n_program; ++i) {
        const int op = d->program[i].op;
        if (op == MO_X_TERM) { if (d->program[i].arg < 0 || (uint32_t)d->program[i].arg >= d->n_terms) return MO_ERR_INVALID_DESC; ++depth; }
        else if (op == MO_X_CONST) ++depth;
        else if (op >= MO_X_ADD && op <= MO_X_MAX) { if (op == MO_X_DIV && !is_float(d->elem)) return MO_ERR_UNSUPPORTED; if (depth < 2) return MO_ERR_INVALID_DESC; --depth; }
        else if (op >= MO_X_NEGATE && op <= MO_X_SQRT) { if (op >= MO_X_RECIP && !is_float(d->elem)) return MO_ERR_UNSUPPORTED; if (depth < 1) return MO_ERR_INVALID_DESC; }
        else return MO_ERR_INVALID_DESC;
        if (depth > MO_MAX_STACK) return MO_ERR_UNSUPPORTED;
      }
      if (depth != 1) return MO_ERR_INVALID_DESC;
      break;
    }
    default: return MO_ERR_INVALID_DESC;
  }
  return MO_OK;
}

/* geometry (size used for the output-size arithmetic) and stencilCenter of one constructor, one dimension */
static void kind_geom(int32_t kind, uint32_t flags, int64_t s, int64_t given, int64_t *g, int64_t *c) {
  *g = s; *c = 0;
  switch (kind) {
    case MO_AVG: *g = s > 1 ? s : 1; break;
    case MO_CONV: *c = s - 1 - s / 2; break;
    case MO_MAG2: *c = (flags & MO_FLAG_MAG2_CORR) ? floor_div(s, 2) : s - 1 - s / 2; break;
    case MO_CORR: *c = floor_div(s, 2); break;
    case MO_TAPS: *c = given; break;
    case MO_LIFE: *g = 3; *c = 1; break;
    default: break;
  }
}

int mo_geom(const mo_desc *d, int64_t *g, int64_t *c) {
  int st = validate(d);
  if (st) return st;
  if (d->kind == MO_EXPR) {
    /* unionStencilCenters: max of the centres; unionStencilSizes: centre + max (size - centre)
     * (Stencil/Internal.hs:83-90), folded over all terms; `pure c` is size 1, centre 0 */
    for (int i = 0; i < d->rank; ++i) {
      int64_t mc = 0, ext = 1;
      for (uint32_t t = 0; t < d->n_terms; ++t) {
        int64_t tg, tc;
        kind_geom(d->terms[t].kind, 0, d->terms[t].size[i], d->terms[t].center[i], &tg, &tc);
        if (d->terms[t].center[i] != tc) return MO_ERR_INVALID_DESC;
        if (tc > mc) mc = tc;
        if (tg - tc > ext) ext = tg - tc;
      }
      c[i] = mc; g[i] = mc + ext;
      if (d->center[i] != c[i] || d->size[i] != g[i]) return MO_ERR_INVALID_DESC;
    }
    return MO_OK;
  }
  for (int i = 0; i < d->rank; ++i) {
    int64_t s = d->size[i];
    switch (d->kind) {
      case MO_AVG: /* liftA2 (/) with `pure n` (size 1, centre 0): Stencil/Internal.hs:83-90 */
        g[i] = s > 1 ? s : 1; c[i] = 0; break;
      case MO_CONV: case MO_MAG2:
        if (d->kind == MO_MAG2 && (d->flags & MO_FLAG_MAG2_CORR)) { g[i] = s; c[i] = floor_div(s, 2); break; }
        g[i] = s; c[i] = s - 1 - s / 2; break; /* Convolution.hs:71-73 (quot; s >= 1) */
      case MO_CORR: g[i] = s; c[i] = floor_div(s, 2); break; /* Convolution.hs:114 */
      case MO_TAPS: g[i] = s; c[i] = d->center[i]; break;
      case MO_LIFE: g[i] = 3; c[i] = 1; break;
      default: g[i] = s; c[i] = 0; break; /* fold stencils are centred at 0: Stencil.hs:298-302 */
    }
  }
  for (int i = 0; i < d->rank; ++i) if (d->center[i] != c[i]) return MO_ERR_INVALID_DESC;
  if (d->kind == MO_TAPS) {
    for (int64_t t = 0; t < d->ntaps; ++t)
      for (int i = 0; i < d->rank; ++i) {
        int64_t o = d->tap_offsets[t * d->rank + i];
        if (o < -c[i] || o >= g[i] - c[i]) return MO_ERR_INVALID_DESC; /* outside stencil box */
      }
  }
  return MO_OK;
}

int mo_out_dims(const mo_desc *d, const int64_t *in_dims, int64_t *out_dims) {
  int64_t g[MO_MAX_RANK], c[MO_MAX_RANK];
  int st = mo_geom(d, g, c);
  if (st) return st;
  for (int i = 0; i < d->rank; ++i) {
    if (in_dims[i] < 0) return MO_ERR_INVALID_DESC;
    int64_t shrink = g[i] - 1 > 0 ? g[i] - 1 : 0;               /* Stencil.hs:207 (Sz clamps) */
    int64_t full = d->pad_lo[i] + d->pad_hi[i] + in_dims[i] - shrink; /* Stencil.hs:208 */
    if (full < 0) full = 0;
    out_dims[i] = floor_div(full - 1, d->stride[i]) + 1;         /* Stride.hs:102-105 */
  }
  return MO_OK;
}

typedef struct {
  int32_t kind, rank, border;
  int64_t in_dims[MO_MAX_RANK], out_dims[MO_MAX_RANK], stride[MO_MAX_RANK];
  int64_t shift[MO_MAX_RANK];   /* centre - pad_lo: source index of output cell 0 (Stencil.hs:200) */
  int64_t ntaps;
  int64_t *tap_off;             /* ntaps * rank, application order */
  int64_t *tap_lin;             /* linearised against in_dims */
  int64_t min_off[MO_MAX_RANK], max_off[MO_MAX_RANK];
  const void *c1, *c2;
  int64_t avg_n;
  uint8_t fill[8];
  int no_fast;                  /* MO_FLAG_NO_FAST: force the plain tap loop everywhere */
} mo_plan;

static int build_plan(const mo_desc *d, const int64_t *in_dims, mo_plan *p) {
  int64_t g[MO_MAX_RANK], c[MO_MAX_RANK];
  int st = mo_geom(d, g, c);
  if (st) return st;
  memset(p, 0, sizeof(*p));
  p->kind = d->kind; p->rank = d->rank; p->border = d->border;
  memcpy(p->fill, d->fill, 8);
  st = mo_out_dims(d, in_dims, p->out_dims);
  if (st) return st;
  const int r = d->rank;
  for (int i = 0; i < r; ++i) {
    p->in_dims[i] = in_dims[i]; p->stride[i] = d->stride[i];
    p->shift[i] = c[i] - d->pad_lo[i];
  }
  p->c1 = d->coeffs; p->c2 = d->coeffs2;
  p->no_fast = (d->flags & MO_FLAG_NO_FAST) != 0;
  /* tap list */
  int64_t n = 1;
  if (d->kind == MO_TAPS) n = d->ntaps;
  else if (d->kind == MO_LIFE) n = 9;
  else for (int i = 0; i < r; ++i) n *= d->size[i];
  p->ntaps = n;
  p->avg_n = n; /* totalElem sz of the ORIGINAL size (0 if any dim is 0 -> 0/0 = NaN) */
  p->tap_off = (int64_t *)calloc((size_t)(n > 0 ? n : 1) * r, sizeof(int64_t));
  p->tap_lin = (int64_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int64_t));
  if (d->kind == MO_TAPS) {
    memcpy(p->tap_off, d->tap_offsets, sizeof(int64_t) * (size_t)n * r);
  } else if (d->kind == MO_LIFE) {
    static const int lo[9][2] = {{-1,-1},{-1,0},{-1,1},{0,-1},{0,1},{1,-1},{1,0},{1,1},{0,0}};
    for (int t = 0; t < 9; ++t) { p->tap_off[t*2] = lo[t][0]; p->tap_off[t*2+1] = lo[t][1]; }
  } else {
    int corr = (d->kind == MO_CORR) || (d->kind == MO_MAG2 && (d->flags & MO_FLAG_MAG2_CORR));
    int conv = (d->kind == MO_CONV) || (d->kind == MO_MAG2 && !corr);
    int64_t q[MO_MAX_RANK] = {0};
    for (int64_t t = 0; t < n; ++t) {  /* row-major walk over [0, size) — iter, Core/Index.hs:556-573 */
      for (int i = 0; i < r; ++i) {
        int64_t s = d->size[i];
        if (conv) p->tap_off[t*r+i] = s / 2 - q[i];                 /* ix + sCenter - kIx  */
        else if (corr) p->tap_off[t*r+i] = q[i] - floor_div(s, 2);  /* ix - sCenter + kIx  */
        else p->tap_off[t*r+i] = q[i];                              /* fold: get (ix + q)  */
      }
      for (int i = r - 1; i >= 0; --i) { if (++q[i] < d->size[i]) break; q[i] = 0; }
    }
  }
  for (int64_t t = 0; t < n; ++t) {
    int64_t lin = 0;
    for (int i = 0; i < r; ++i) {
      int64_t o = p->tap_off[t*r+i];
      lin = lin * in_dims[i] + o;
      if (t == 0 || o < p->min_off[i]) p->min_off[i] = o;
      if (t == 0 || o > p->max_off[i]) p->max_off[i] = o;
    }
    p->tap_lin[t] = lin;
  }
  /* IndexZeroException: a non-Fill border evaluated against a zero extent */
  int64_t total_out = 1; int zero_in = 0;
  for (int i = 0; i < r; ++i) { total_out *= p->out_dims[i]; if (in_dims[i] == 0) zero_in = 1; }
  if (total_out > 0 && n > 0 && zero_in && d->border != MO_FILL) {
    free(p->tap_off); free(p->tap_lin);
    return MO_ERR_INDEX_ZERO;
  }
  return MO_OK;
}

static void free_plan(mo_plan *p) { free(p->tap_off); free(p->tap_lin); }

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(x) CAT(x, SUF)

#define LOADX(x) (x)
#define ISFLT 0
#define ACC uint64_t
#define T uint8_t
#define SUF u8
#define TMIN 0
#define TMAX UINT8_MAX
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
#define T int8_t
#define SUF i8
#define TMIN INT8_MIN
#define TMAX INT8_MAX
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
#define T uint16_t
#define SUF u16
#define TMIN 0
#define TMAX UINT16_MAX
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
#define T int16_t
#define SUF i16
#define TMIN INT16_MIN
#define TMAX INT16_MAX
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
#define T uint32_t
#define SUF u32
#define TMIN 0
#define TMAX UINT32_MAX
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
#define T int32_t
#define SUF i32
#define TMIN INT32_MIN
#define TMAX INT32_MAX
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
#define T uint64_t
#define SUF u64
#define TMIN 0
#define TMAX UINT64_MAX
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
#define T int64_t
#define SUF i64
#define TMIN INT64_MIN
#define TMAX INT64_MAX
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
/* Storable Bool: 4-byte HsBool; peek = (/= 0); poke True = 1.  Bounded: False..True */
#undef LOADX
#define LOADX(x) ((x) != 0)
#define T int32_t
#define SUF b32
#define TMIN 0
#define TMAX 1
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef TMIN
#undef TMAX
#undef LOADX
#define LOADX(x) (x)
#undef ISFLT
#undef ACC
#define ISFLT 1
#define ACC float
#define T float
#define SUF f32
#define TMIN 0
#define TMAX 0
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef ACC
#define ACC double
#define T double
#define SUF f64
#include "massiv_oracle_tmpl.inc"
#undef T
#undef SUF
#undef ACC

/* ---- post-maps: fmap f over the results ------------------------------------------------------- */
#define POST_FLOAT(T, NAME, SQRT)                                                              \
  static T NAME(int op, T x, T c) {                                                            \
    switch (op) {                                                                              \
      case MO_POST_NEGATE: return -x;                 /* negateFloat#: sign flip */            \
      case MO_POST_ABS: return x == 0 ? (T)0 : (x > 0 ? x : -x); /* GHC.Float abs */           \
      case MO_POST_SIGNUM: return x > 0 ? (T)1 : (x < 0 ? (T)-1 : x); /* GHC.Float signum */   \
      case MO_POST_SQUARE: return x * x;              /* x ^ 2 */                              \
      case MO_POST_ADD: return x + c;                                                          \
      case MO_POST_SUB: return x - c;                                                          \
      case MO_POST_RSUB: return c - x;                                                         \
      case MO_POST_MUL: return x * c;                                                          \
      case MO_POST_MIN: return x <= c ? x : c;        /* GHC.Classes: min x y = if x <= y then x else y */ \
      case MO_POST_MAX: return x <= c ? c : x;        /*              max x y = if x <= y then y else x */ \
      case MO_POST_DIV: return x / c;                                                          \
      case MO_POST_RDIV: return c / x;                                                         \
      case MO_POST_RECIP: return (T)1 / x;            /* recip x = 1.0 / x */                  \
      default: return SQRT(x);                                                                 \
    }                                                                                          \
  }
POST_FLOAT(float, post_f32, sqrtf)
POST_FLOAT(double, post_f64, sqrt)

#define POST_INT(T, U, NAME)                                                                   \
  static T NAME(int op, T x, T c) {                                                            \
    switch (op) { /* fixed-width wraparound: computed in the unsigned type */                  \
      case MO_POST_NEGATE: return (T)((U)0 - (U)x);                                            \
      case MO_POST_ABS: return x >= 0 ? x : (T)((U)0 - (U)x); /* abs minBound = minBound */    \
      case MO_POST_SIGNUM: return x > 0 ? (T)1 : (x < 0 ? (T)((U)0 - (U)1) : (T)0);            \
      case MO_POST_SQUARE: return (T)((U)x * (U)x);                                            \
      case MO_POST_ADD: return (T)((U)x + (U)c);                                               \
      case MO_POST_SUB: return (T)((U)x - (U)c);                                               \
      case MO_POST_RSUB: return (T)((U)c - (U)x);                                              \
      case MO_POST_MUL: return (T)((U)x * (U)c);                                               \
      case MO_POST_MIN: return x <= c ? x : c;                                                 \
      default: return x <= c ? c : x;                                                          \
    }                                                                                          \
  }
POST_INT(uint8_t, uint32_t, post_u8)
POST_INT(int8_t, uint32_t, post_i8)
POST_INT(uint16_t, uint32_t, post_u16)
POST_INT(int16_t, uint32_t, post_i16)
POST_INT(uint32_t, uint32_t, post_u32)
POST_INT(int32_t, uint32_t, post_i32)
POST_INT(uint64_t, uint64_t, post_u64)
POST_INT(int64_t, uint64_t, post_i64)

static int apply_post(const mo_desc *d, void *out, int64_t n) {
  if (d->n_post > MO_MAX_POST) return MO_ERR_UNSUPPORTED;
  if (!d->post) return MO_ERR_INVALID_DESC;
  if (d->elem == MO_BOOL32) return MO_ERR_UNSUPPORTED;
  for (uint32_t k = 0; k < d->n_post; ++k) {
    const int op = d->post[k].op;
    if (op < MO_POST_NEGATE || op > MO_POST_SQRT) return MO_ERR_INVALID_DESC;
    if (op >= MO_POST_DIV && !is_float(d->elem)) return MO_ERR_UNSUPPORTED;
#define RUN(T, F) { T c; memcpy(&c, d->post[k].operand, sizeof(T)); T *o = (T *)out; for (int64_t i = 0; i < n; ++i) o[i] = F(op, o[i], c); }
    switch (d->elem) {
      case MO_U8: RUN(uint8_t, post_u8) break;
      case MO_I8: RUN(int8_t, post_i8) break;
      case MO_U16: RUN(uint16_t, post_u16) break;
      case MO_I16: RUN(int16_t, post_i16) break;
      case MO_U32: RUN(uint32_t, post_u32) break;
      case MO_I32: RUN(int32_t, post_i32) break;
      case MO_U64: RUN(uint64_t, post_u64) break;
      case MO_I64: RUN(int64_t, post_i64) break;
      case MO_F32: RUN(float, post_f32) break;
      case MO_F64: RUN(double, post_f64) break;
      default: return MO_ERR_INVALID_DESC;
    }
#undef RUN
  }
  return MO_OK;
}

static void run_plan(int32_t elem, mo_plan *p, const void *in, void *out, int nthreads) {
  switch (elem) {
    case MO_U8: apply_u8(p, (const uint8_t *)in, (uint8_t *)out, nthreads); break;
    case MO_I8: apply_i8(p, (const int8_t *)in, (int8_t *)out, nthreads); break;
    case MO_U16: apply_u16(p, (const uint16_t *)in, (uint16_t *)out, nthreads); break;
    case MO_I16: apply_i16(p, (const int16_t *)in, (int16_t *)out, nthreads); break;
    case MO_U32: apply_u32(p, (const uint32_t *)in, (uint32_t *)out, nthreads); break;
    case MO_I32: apply_i32(p, (const int32_t *)in, (int32_t *)out, nthreads); break;
    case MO_U64: apply_u64(p, (const uint64_t *)in, (uint64_t *)out, nthreads); break;
    case MO_I64: apply_i64(p, (const int64_t *)in, (int64_t *)out, nthreads); break;
    case MO_BOOL32: apply_b32(p, (const int32_t *)in, (int32_t *)out, nthreads); break;
    case MO_F32: apply_f32(p, (const float *)in, (float *)out, nthreads); break;
    default: apply_f64(p, (const double *)in, (double *)out, nthreads); break;
  }
}

/* MO_EXPR (Stencil/Internal.hs:93-174): every term is evaluated over the whole output at the SAME absolute
 * source index — output cell o sits at o + (centre - pad_lo) with the UNION centre (Stencil.hs:200) — into
 * an array of its own; the postfix program then combines the arrays cell by cell in the element type.
 * The reference composes the closures per cell instead; the values are the same (each f_t is a pure function
 * of the index).  liftA2 min / max use GHC.Classes' defaults (min x y = if x <= y then x else y). */
static int apply_expr(const mo_desc *d, const int64_t *in_dims, const void *in, void *out, int nthreads) {
  int64_t g[MO_MAX_RANK], c[MO_MAX_RANK], od[MO_MAX_RANK];
  int st = mo_geom(d, g, c);
  if (st) return st;
  if ((st = mo_out_dims(d, in_dims, od))) return st;
  const int r = d->rank;
  const int es = mo_elem_size(d->elem);
  int64_t n = 1; int zero_in = 0;
  for (int i = 0; i < r; ++i) { n *= od[i]; if (in_dims[i] == 0) zero_in = 1; }
  void *tv[MO_MAX_TERMS] = {0};
  for (uint32_t t = 0; t < d->n_terms && !st; ++t) {
    mo_desc td = *d;
    td.kind = d->terms[t].kind; td.coeffs = d->terms[t].coeffs; td.coeffs2 = NULL;
    td.ntaps = d->terms[t].ntaps; td.tap_offsets = d->terms[t].tap_offsets;
    td.n_post = 0; td.n_terms = 0; td.n_program = 0; td.flags = d->flags & MO_FLAG_NO_FAST;
    for (int i = 0; i < r; ++i) {
      td.size[i] = d->terms[t].size[i]; td.center[i] = d->terms[t].center[i];
      td.pad_lo[i] = 0; td.pad_hi[i] = 0;
    }
    mo_plan p;
    if ((st = build_plan(&td, in_dims, &p))) break;
    for (int i = 0; i < r; ++i) { p.shift[i] = c[i] - d->pad_lo[i]; p.out_dims[i] = od[i]; }
    if (n > 0 && p.ntaps > 0 && zero_in && d->border != MO_FILL) { free_plan(&p); st = MO_ERR_INDEX_ZERO; break; }
    tv[t] = malloc((size_t)(n > 0 ? n : 1) * es);
    if (n > 0) run_plan(d->elem, &p, in, tv[t], nthreads);
    free_plan(&p);
  }
  if (!st) {
    static const int un_map[] = {MO_POST_NEGATE, MO_POST_ABS, MO_POST_SIGNUM, MO_POST_SQUARE, MO_POST_RECIP, MO_POST_SQRT};
    static const int bin_map[] = {MO_POST_ADD, MO_POST_SUB, MO_POST_MUL, MO_POST_DIV, MO_POST_MIN, MO_POST_MAX};
#define RUNX(T, F)                                                                             \
    for (int64_t i = 0; i < n; ++i) {                                                          \
      T stack[MO_MAX_STACK]; int sp = 0;                                                       \
      for (uint32_t k = 0; k < d->n_program; ++k) {                                            \
        const mo_xins *x = &d->program[k];                                                     \
        if (x->op == MO_X_TERM) stack[sp++] = ((const T *)tv[x->arg])[i];                      \
        else if (x->op == MO_X_CONST) { T cv; memcpy(&cv, x->operand, sizeof(T)); stack[sp++] = cv; } \
        else if (x->op <= MO_X_MAX) { T b = stack[--sp]; T a = stack[sp - 1]; stack[sp - 1] = F(bin_map[x->op - MO_X_ADD], a, b); } \
        else stack[sp - 1] = F(un_map[x->op - MO_X_NEGATE], stack[sp - 1], (T)0);              \
      }                                                                                        \
      ((T *)out)[i] = stack[0];                                                                \
    }
    switch (d->elem) {
      case MO_U8: RUNX(uint8_t, post_u8) break;
      case MO_I8: RUNX(int8_t, post_i8) break;
      case MO_U16: RUNX(uint16_t, post_u16) break;
      case MO_I16: RUNX(int16_t, post_i16) break;
      case MO_U32: RUNX(uint32_t, post_u32) break;
      case MO_I32: RUNX(int32_t, post_i32) break;
      case MO_U64: RUNX(uint64_t, post_u64) break;
      case MO_I64: RUNX(int64_t, post_i64) break;
      case MO_F32: RUNX(float, post_f32) break;
      case MO_F64: RUNX(double, post_f64) break;
      default: st = MO_ERR_UNSUPPORTED;
    }
#undef RUNX
  }
  for (uint32_t t = 0; t < d->n_terms; ++t) free(tv[t]);
  if (!st && d->n_post) st = apply_post(d, out, n);
  return st;
}

int mo_apply(const mo_desc *d, const int64_t *in_dims, const void *in, void *out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (d->kind == MO_EXPR) return apply_expr(d, in_dims, in, out, nthreads);
  mo_plan p;
  int st = build_plan(d, in_dims, &p);
  if (st) return st;
  switch (d->elem) {
    case MO_U8: apply_u8(&p, (const uint8_t *)in, (uint8_t *)out, nthreads); break;
    case MO_I8: apply_i8(&p, (const int8_t *)in, (int8_t *)out, nthreads); break;
    case MO_U16: apply_u16(&p, (const uint16_t *)in, (uint16_t *)out, nthreads); break;
    case MO_I16: apply_i16(&p, (const int16_t *)in, (int16_t *)out, nthreads); break;
    case MO_U32: apply_u32(&p, (const uint32_t *)in, (uint32_t *)out, nthreads); break;
    case MO_I32: apply_i32(&p, (const int32_t *)in, (int32_t *)out, nthreads); break;
    case MO_U64: apply_u64(&p, (const uint64_t *)in, (uint64_t *)out, nthreads); break;
    case MO_I64: apply_i64(&p, (const int64_t *)in, (int64_t *)out, nthreads); break;
    case MO_BOOL32: apply_b32(&p, (const int32_t *)in, (int32_t *)out, nthreads); break;
    case MO_F32: apply_f32(&p, (const float *)in, (float *)out, nthreads); break;
    case MO_F64: apply_f64(&p, (const double *)in, (double *)out, nthreads); break;
    default: st = MO_ERR_INVALID_DESC;
  }
  if (!st && d->n_post) {
    int64_t n = 1;
    for (int i = 0; i < d->rank; ++i) n *= p.out_dims[i];
    st = apply_post(d, out, n);
  }
  free_plan(&p);
  return st;
}

/* iterateUntil (Manifest/Internal.hs:331-349) with convergence `n >= n_steps - 1`: the step is
 * applied n_steps times (counter n = 0 .. n_steps-1), double-buffered (iterateLoop :376-397). */
int mo_iterate(const mo_desc *d, const int64_t *dims, const void *in, void *out, void *tmp,
               int64_t n_steps, int nthreads) {
  int64_t od[MO_MAX_RANK];
  int st = mo_out_dims(d, dims, od);
  if (st) return st;
  for (int i = 0; i < d->rank; ++i) if (od[i] != dims[i]) return MO_ERR_SIZE_MISMATCH;
  if (n_steps < 1) return MO_ERR_INVALID_DESC; /* iterateUntil always applies the step at least once */
  const void *cur = in;
  for (int64_t s = 0; s < n_steps; ++s) {
    void *dst = ((n_steps - 1 - s) % 2 == 0) ? out : tmp;
    st = mo_apply(d, dims, cur, dst, nthreads);
    if (st) return st;
    cur = dst;
  }
  return MO_OK;
}
