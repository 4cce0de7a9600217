// Well-known 3x3 kernels with compile-time coefficients on the tiled rank-2 kernel.
//
// A FromKernel stencil multiplies every tap by its coefficient and adds, two rounded operations per
// tap (Array/Stencil/Convolution.hs:74-78) — 36 + 3 operations plus a square root per cell for the
// Sobel magnitude, more than the FP32 pipe can issue in the time HBM needs to move the cell.  When the
// kernel array is one of the classics below, its small integer coefficients are template constants
// and each tap costs what it is worth (tap_known in k_tile2d.cuh): +-1 -> one add, 0 -> one exact
// FMA (or nothing under MB_FLAG_ASSUME_FINITE), anything else -> the reference's multiply and add.
// Same operation order, same roundings, bit-identical results; any other kernel array takes the
// generic coefficient path.
#include "k_tile2d.cuh"

namespace mb {

#define MB_KC(NAME, ...)                                                                     \
  struct NAME {                                                                              \
    __host__ __device__ static constexpr int k1(int t) { constexpr int a[9] = {__VA_ARGS__}; return a[t]; } \
    __host__ __device__ static constexpr int k2(int) { return 0; }                          \
  };
#define MB_KC2(NAME, A1, A2)                                                                 \
  struct NAME {                                                                              \
    __host__ __device__ static constexpr int k1(int t) { constexpr int a[9] = A1; return a[t]; } \
    __host__ __device__ static constexpr int k2(int t) { constexpr int a[9] = A2; return a[t]; } \
  };
#define MB_ARR(...) {__VA_ARGS__}

MB_KC(KSobelX, -1, 0, 1, -2, 0, 2, -1, 0, 1)
MB_KC(KSobelY, -1, -2, -1, 0, 0, 0, 1, 2, 1)
MB_KC(KLaplace4, 0, 1, 0, 1, -4, 1, 0, 1, 0)
MB_KC(KLaplace8, 1, 1, 1, 1, -8, 1, 1, 1, 1)
MB_KC2(KSobelXY, MB_ARR(-1, 0, 1, -2, 0, 2, -1, 0, 1), MB_ARR(-1, -2, -1, 0, 0, 0, 1, 2, 1))
MB_KC2(KSobelYX, MB_ARR(-1, -2, -1, 0, 0, 0, 1, 2, 1), MB_ARR(-1, 0, 1, -2, 0, 2, -1, 0, 1))
MB_KC2(KPrewittXY, MB_ARR(-1, 0, 1, -1, 0, 1, -1, 0, 1), MB_ARR(-1, -1, -1, 0, 0, 0, 1, 1, 1))

template <typename T, typename KC>
static bool matches(const Plan &p, bool two) {
  const T *a = reinterpret_cast<const T *>(p.c1.data());
  const T *b = two ? reinterpret_cast<const T *>(p.c2.data()) : nullptr;
  for (int t = 0; t < 9; ++t) {
    if (a[t] != T(KC::k1(t))) return false;
    if (two && b[t] != T(KC::k2(t))) return false;
  }
  return true;
}

// May zero taps be skipped WITHOUT the caller's promise, with a check of the results (k_tile2d.cuh, FIN = 2)?
// The two forms differ only where a non-finite input sits on a zero tap (0 * Inf = NaN).  Under mapStencil's geometry
// every input cell (y, x) is window position (1 - dy, 1 - dx) (mirrored for a convolution) of output (y + dy, x + dx);
// in an array of at least 3 x 3 cells the outputs with dy in {0, sy}, dx in {0, sx} exist for some pair of signs, so
// the cell is seen at the centre, at one horizontal, one vertical and one diagonal neighbour position.  If in every
// such quadrant one of these four coefficients (of either kernel of a magnitude pair) is non-zero, a non-finite cell
// makes some result non-finite — and that is what the launch looks for.  The whole array in one launch (no row
// ranges), Fill elements finite.
template <typename T, typename KC>
static bool known_optimistic_ok(Ctx *c, const Plan &p, const OuterRange *r) {
  if (c->no_optimistic || r != nullptr || !p.same_shape() || p.in_dims[0] < 3 || p.in_dims[1] < 3) return false;
  if (p.shift[0] + p.min_off[0] != -1 || p.shift[1] + p.min_off[1] != -1) return false;
  if (p.border == MB_FILL) {
    T fv;
    std::memcpy(&fv, p.fill, sizeof(T));
    if (!(fv - fv == T(0))) return false;
  }
  auto nz = [](int ry, int cx) { return KC::k1(ry * 3 + cx) != 0 || KC::k2(ry * 3 + cx) != 0; };
  for (int ry = 0; ry <= 2; ry += 2)
    for (int cx = 0; cx <= 2; cx += 2)
      if (!(nz(1, 1) || nz(1, cx) || nz(ry, 1) || nz(ry, cx))) return false;
  return true;
}

// centred 3x3 under mapStencil only (XOFF = the natural offset); everything else is generic
template <typename T, int OP, typename KC>
static int launch_kc(Ctx *c, const DevPlan &dp, const void *i, void *o, const OuterRange *r, bool rev) {
  constexpr int V = 16 / (int)sizeof(T);
  constexpr int NAT = (V - (1 % V)) % V;
  const int sx = (int)(dp.p.shift[1] + dp.p.min_off[1]);
  if ((((sx % V) + V) % V) != NAT) return -1;
  const bool fin = (dp.p.flags & MB_FLAG_ASSUME_FINITE) != 0;
  if (!fin && known_optimistic_ok<T, KC>(c, dp.p, r) && ensure_flag(c) == MB_OK) {
    // optimistic launch that inspects its results, then the strict one, which only works if the flag went up
    MB_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(unsigned), c->stream));
    int st = rev ? launch_x<T, OP, true, 3, 3, NAT, KC, 2>(c, dp, i, o, r, c->d_flag) : launch_x<T, OP, false, 3, 3, NAT, KC, 2>(c, dp, i, o, r, c->d_flag);
    if (st) return st;
    c->aux_launches++;
    return rev ? launch_x<T, OP, true, 3, 3, NAT, KC, 0>(c, dp, i, o, r, c->d_flag) : launch_x<T, OP, false, 3, 3, NAT, KC, 0>(c, dp, i, o, r, c->d_flag);
  }
  if (rev) return fin ? launch_x<T, OP, true, 3, 3, NAT, KC, 1>(c, dp, i, o, r) : launch_x<T, OP, true, 3, 3, NAT, KC, 0>(c, dp, i, o, r);
  return fin ? launch_x<T, OP, false, 3, 3, NAT, KC, 1>(c, dp, i, o, r) : launch_x<T, OP, false, 3, 3, NAT, KC, 0>(c, dp, i, o, r);
}

template <typename T>
static int known_t(Ctx *c, const DevPlan &dp, const void *i, void *o, const OuterRange *r) {
  const Plan &p = dp.p;
  if (p.kind == MB_MAG2) {
    const bool rev = !(p.flags & MB_FLAG_MAG2_CORR);
    if (matches<T, KSobelXY>(p, true)) return launch_kc<T, OP_MAG, KSobelXY>(c, dp, i, o, r, rev);
    if (matches<T, KSobelYX>(p, true)) return launch_kc<T, OP_MAG, KSobelYX>(c, dp, i, o, r, rev);
    if (matches<T, KPrewittXY>(p, true)) return launch_kc<T, OP_MAG, KPrewittXY>(c, dp, i, o, r, rev);
    return -1;
  }
  const bool rev = p.kind == MB_CONV;
  if (matches<T, KSobelX>(p, false)) return launch_kc<T, OP_LIN, KSobelX>(c, dp, i, o, r, rev);
  if (matches<T, KSobelY>(p, false)) return launch_kc<T, OP_LIN, KSobelY>(c, dp, i, o, r, rev);
  if (matches<T, KLaplace4>(p, false)) return launch_kc<T, OP_LIN, KLaplace4>(c, dp, i, o, r, rev);
  if (matches<T, KLaplace8>(p, false)) return launch_kc<T, OP_LIN, KLaplace8>(c, dp, i, o, r, rev);
  return -1;
}

int launch_tile2d_known(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p = dp.p;
  if (c->no_known) return -1;
  if (p.rank != 2 || p.size[0] != 3 || p.size[1] != 3) return -1;
  if (p.kind != MB_CONV && p.kind != MB_CORR && p.kind != MB_MAG2) return -1;
  switch (p.elem) {
    case MB_F32: return known_t<float>(c, dp, d_in, d_out, range);
    case MB_F64: return known_t<double>(c, dp, d_in, d_out, range);
  }
  return -1;
}

}  // namespace mb
