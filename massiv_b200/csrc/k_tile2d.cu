// Dispatch of the tiled rank-2 kernel (k_tile2d.cuh) over element types, window shapes and ops.
#include "k_tile2d.cuh"

namespace mb {

// the window's offset into the 16-byte aligned staged row; instantiated for 0 (fold stencils, valid
// padding) and for the value mapStencil gives a centred kernel — every offset for 8-byte elements
template <typename T, int OP, bool REV, int KH, int KW>
static int launch_cfg(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  constexpr int V = 16 / (int)sizeof(T);
  const int sx = (int)(dp.p.shift[1] + dp.p.min_off[1]);
  const int xoff = ((sx % V) + V) % V;
  constexpr int NAT = (V - ((KW / 2) % V)) % V;  // centred odd kernel under mapStencil
  if (xoff == 0) return launch_x<T, OP, REV, KH, KW, 0>(c, dp, d_in, d_out, range);
  if constexpr (NAT != 0) {
    if (xoff == NAT) return launch_x<T, OP, REV, KH, KW, NAT>(c, dp, d_in, d_out, range);
  }
  if constexpr (V == 2 && NAT != 1) {
    if (xoff == 1) return launch_x<T, OP, REV, KH, KW, 1>(c, dp, d_in, d_out, range);
  }
  return -1;
}

template <typename T, int KH, int KW>
static int launch_shape(Ctx *c, const DevPlan &dp, const void *i, void *o, const OuterRange *r) {
  const Plan &p = dp.p;
  switch (p.kind) {
    case MB_SUM: case MB_AVG: return launch_cfg<T, OP_SUM, false, KH, KW>(c, dp, i, o, r);
    case MB_PRODUCT: return launch_cfg<T, OP_PRODUCT, false, KH, KW>(c, dp, i, o, r);
    case MB_CONV: return launch_cfg<T, OP_LIN, true, KH, KW>(c, dp, i, o, r);
    case MB_CORR: return launch_cfg<T, OP_LIN, false, KH, KW>(c, dp, i, o, r);
    case MB_MAG2:
      if (p.flags & MB_FLAG_MAG2_CORR) return launch_cfg<T, OP_MAG, false, KH, KW>(c, dp, i, o, r);
      return launch_cfg<T, OP_MAG, true, KH, KW>(c, dp, i, o, r);
  }
  return -1;
}

template <typename T, int KH, int KW>
static int launch_shape_int(Ctx *c, const DevPlan &dp, const void *i, void *o, const OuterRange *r) {
  switch (dp.p.kind) {
    case MB_SUM: return launch_cfg<T, OP_SUM, false, KH, KW>(c, dp, i, o, r);
    case MB_MAX: return launch_cfg<T, OP_MAX, false, KH, KW>(c, dp, i, o, r);
    case MB_MIN: return launch_cfg<T, OP_MIN, false, KH, KW>(c, dp, i, o, r);
  }
  return -1;
}

#define MB_SHAPES(X) X(3, 3) X(5, 5) X(7, 7) X(2, 2) X(1, 3) X(3, 1) X(1, 5) X(5, 1) X(1, 7) X(7, 1)

template <typename T>
static int launch_float(Ctx *c, const DevPlan &dp, const void *i, void *o, const OuterRange *r) {
  const int kh = (int)dp.p.size[0], kw = (int)dp.p.size[1];
#define X(H, W) if (kh == H && kw == W) return launch_shape<T, H, W>(c, dp, i, o, r);
  MB_SHAPES(X)
#undef X
  return -1;
}

template <typename T>
static int launch_int(Ctx *c, const DevPlan &dp, const void *i, void *o, const OuterRange *r) {
  const int kh = (int)dp.p.size[0], kw = (int)dp.p.size[1];
  if (kh == 3 && kw == 3) return launch_shape_int<T, 3, 3>(c, dp, i, o, r);
  if (kh == 2 && kw == 2) return launch_shape_int<T, 2, 2>(c, dp, i, o, r);
  return -1;
}

int launch_tile2d(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p = dp.p;
  if (p.rank != 2 || !p.dense || !p.unit_stride()) return -1;
  if (p.in_dims[0] < 1 || p.in_dims[1] < 1) return -1;
  if (p.in_dims[0] >= (1ll << 30) || p.in_dims[1] >= (1ll << 30) || p.out_dims[0] >= (1ll << 30) ||
      p.out_dims[1] >= (1ll << 30))
    return -1;
  if (p.out_elems < c->tile_min_elems) return -1;  // tiny arrays: one generic launch is cheaper
  {
    const int st = launch_tile2d_known(c, dp, d_in, d_out, range);
    if (st != -1) return st;
  }
  switch (p.elem) {
    case MB_F32: return launch_float<float>(c, dp, d_in, d_out, range);
    case MB_F64: return launch_float<double>(c, dp, d_in, d_out, range);
    case MB_I32: return launch_int<int32_t>(c, dp, d_in, d_out, range);
    case MB_I64: return launch_int<int64_t>(c, dp, d_in, d_out, range);
  }
  return -1;
}

}  // namespace mb
