// Generic stencil kernel: any rank <= 5, any window, any stride/padding, all kinds and element
// types.  One thread per output cell, every tap through the border rule — the direct restatement
// of `warr`'s index function in applyStencil (Array/Stencil.hs:201-205).  It is the correctness
// baseline the tiled kernels are tested against and the path for shapes they do not cover
// (strides, ranks 1/4/5, odd windows, tiny arrays).  It is still a GPU kernel: there is no CPU path.
#include "mb_arith.cuh"
#include "mb_internal.h"
#include "mb_post.cuh"

namespace mb {

struct GParams {
  int rank, border, kind, bool_canon, skip_zero, small;
  int64_t in_dims[MB_MAX_RANK], out_dims[MB_MAX_RANK], stride[MB_MAX_RANK], shift[MB_MAX_RANK];
  int64_t ntaps, cell0, ncells, avg_n;
  int64_t min_off[MB_MAX_RANK], max_off[MB_MAX_RANK];
  const int32_t *taps;
  const int64_t *tap_lin;
  const void *c1, *c2;
  unsigned long long fill_bits;
  XParams X;  // MB_EXPR: terms and program (mb_post.cuh)
};

// a cell whose whole window lies inside the array reads its taps at precomputed linear offsets
struct Cell {
  bool interior;
  int64_t lin;
};

template <typename T>
__device__ __forceinline__ T fetch(const GParams &P, const T *__restrict__ in, const int64_t *ix,
                                   const int32_t *__restrict__ off, T fillv, const Cell &cell, int64_t t) {
  if (cell.interior) {
    T v = in[cell.lin + P.tap_lin[t]];
    if (P.bool_canon) v = (T)(v != T(0));
    return v;
  }
  int64_t lin = 0;
#pragma unroll 1
  for (int d = 0; d < P.rank; ++d) {
    int64_t i = ix[d] + off[d];
    if (!resolve_border(P.border, P.in_dims[d], i)) return fillv;  // Fill: ANY dim out => element
    lin = lin * P.in_dims[d] + i;
  }
  T v = in[lin];
  if (P.bool_canon) v = (T)(v != T(0));  // Storable Bool peek: (/= 0)
  return v;
}

// one built-in constructor evaluated at source index ix: taps [tb, tb + ntaps) of the plan's tap table
template <typename T>
__device__ __forceinline__ T eval_kind(const GParams &P, int kind, int64_t tb, int64_t ntaps, int64_t avg_n, const T *__restrict__ in,
                                       const int64_t *ix, T fillv, const Cell &win) {
  using A = Arith<T>;
  using Acc = typename A::Acc;
  const int r = P.rank;
  const int32_t *taps = P.taps;
  T res;
  switch (kind) {
      case MB_ID: res = fetch<T>(P, in, ix, taps + tb * r, fillv, win, tb); break;
      case MB_SUM: case MB_AVG: {
        Acc acc = A::from_count(0);
        for (int64_t t = 0; t < ntaps; ++t) acc = A::add(acc, A::load(fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t)));
        if (kind == MB_AVG) acc = A::div(acc, A::from_count(avg_n));
        res = A::store(acc);
        break;
      }
      case MB_PRODUCT: {
        Acc acc = A::from_count(1);
        for (int64_t t = 0; t < ntaps; ++t) acc = A::mul(acc, A::load(fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t)));
        res = A::store(acc);
        break;
      }
      case MB_MAX: {
        T acc = P.bool_canon ? T(0) : Bounds<T>::lo();  // Bounded Bool: False..True
        for (int64_t t = 0; t < ntaps; ++t) { T x = fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t); acc = x > acc ? x : acc; }
        res = acc;
        break;
      }
      case MB_MIN: {
        T acc = P.bool_canon ? T(1) : Bounds<T>::hi();
        for (int64_t t = 0; t < ntaps; ++t) { T x = fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t); acc = x < acc ? x : acc; }
        res = acc;
        break;
      }
      case MB_CONV: case MB_CORR: case MB_TAPS: {
        const T *k = (const T *)P.c1 + tb;
        Acc acc = A::from_count(0);
        for (int64_t t = 0; t < ntaps; ++t) {
          const Acc kv = A::load(k[t]);
          if (P.skip_zero && A::is_zero(kv)) continue;
          acc = A::add(A::mul(A::load(fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t)), kv), acc);
        }
        res = A::store(acc);
        break;
      }
      case MB_MAG2: {
        const T *k1 = (const T *)P.c1, *k2 = (const T *)P.c2;
        Acc a = A::from_count(0), b = A::from_count(0);
        for (int64_t t = 0; t < ntaps; ++t) {
          const Acc kv = A::load(k1[t]);
          if (P.skip_zero && A::is_zero(kv)) continue;
          a = A::add(A::mul(A::load(fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t)), kv), a);
        }
        for (int64_t t = 0; t < ntaps; ++t) {
          const Acc kv = A::load(k2[t]);
          if (P.skip_zero && A::is_zero(kv)) continue;
          b = A::add(A::mul(A::load(fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t)), kv), b);
        }
        res = A::store(A::sqrt(A::add(A::mul(a, a), A::mul(b, b))));
        break;
      }
      case MB_INT_MIDPOINT: {  // Integral.hs:64-66: dx * loop 0 (< k) (+ 1) 0 (\i -> (+ g i))
        Acc acc = A::from_count(0);
        for (int64_t t = 0; t < ntaps; ++t) acc = A::add(acc, A::load(fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t)));
        res = A::store(A::mul(A::load(*(const T *)P.c1), acc));
        break;
      }
      case MB_INT_TRAPEZOID: {  // Integral.hs:85-89: (dx / 2) * (loop 1 (< n) .. (g 0) (+ 2 * g i) + g n)
        const int64_t n = ntaps - 1;
        Acc acc = A::load(fetch<T>(P, in, ix, taps + tb * r, fillv, win, tb));
        for (int64_t t = 1; t < n; ++t)
          acc = A::add(acc, A::mul(A::from_count(2), A::load(fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t))));
        acc = A::add(acc, A::load(fetch<T>(P, in, ix, taps + (tb + n) * r, fillv, win, tb + n)));
        res = A::store(A::mul(A::div(A::load(*(const T *)P.c1), A::from_count(2)), acc));
        break;
      }
      case MB_INT_SIMPSON: {  // Integral.hs:112-117: acc + prev + 4 * g (i + 1) + fx3, then dx / 3 * acc
        const int64_t n = ntaps - 1;
        Acc acc = A::from_count(0);
        Acc prev = A::load(fetch<T>(P, in, ix, taps + tb * r, fillv, win, tb));
        for (int64_t i = 0; i < n - 1; i += 2) {
          const Acc fx3 = A::load(fetch<T>(P, in, ix, taps + (tb + (i + 2)) * r, fillv, win, tb + i + 2));
          const Acc mid = A::mul(A::from_count(4), A::load(fetch<T>(P, in, ix, taps + (tb + (i + 1)) * r, fillv, win, tb + i + 1)));
          acc = A::add(A::add(A::add(acc, prev), mid), fx3);
          prev = fx3;
        }
        res = A::store(A::mul(A::div(A::load(*(const T *)P.c1), A::from_count(3)), acc));
        break;
      }
      default: {  // MB_LIFE (T = uint8_t): GameOfLife.hs:15-36
        unsigned n = 0;
        for (int t = 0; t < 8; ++t) n += (unsigned)fetch<T>(P, in, ix, taps + (tb + t) * r, fillv, win, tb + t);
        n &= 0xFFu;
        const unsigned c = (unsigned)fetch<T>(P, in, ix, taps + (tb + 8) * r, fillv, win, tb + 8);
        res = (T)(((c == 0 && n == 3) || (c == 1 && (n == 2 || n == 3))) ? 1 : 0);
        break;
      }
    }
  return res;
}

template <typename T>
__global__ void __launch_bounds__(256) generic_kernel(const GParams P, const T *__restrict__ in, T *__restrict__ out) {
  using A = Arith<T>;
  using Acc = typename A::Acc;
  T fillv;
  {
    unsigned long long b = P.fill_bits;
    memcpy(&fillv, &b, sizeof(T));
  }
  const int r = P.rank;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < P.ncells;
       cell += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lin_out = P.cell0 + cell;
    int64_t ix[MB_MAX_RANK];
    if (P.small) {  // everything fits 32 bits: the 64-bit divisions are most of a cell's cost otherwise
      unsigned rem = (unsigned)lin_out;
#pragma unroll 1
      for (int d = r - 1; d >= 0; --d) {
        const unsigned od = (unsigned)P.out_dims[d];
        const unsigned q = rem / od, j = rem - q * od;
        rem = q;
        ix[d] = (int64_t)j * P.stride[d] + P.shift[d];  // Stencil.hs:200 + Stride.hs:110-120
      }
    } else {
      int64_t rem = lin_out;
#pragma unroll 1
      for (int d = r - 1; d >= 0; --d) {
        const int64_t j = rem % P.out_dims[d];
        rem /= P.out_dims[d];
        ix[d] = j * P.stride[d] + P.shift[d];
      }
    }
    const int32_t *taps = P.taps;
    Cell win;
    win.interior = P.tap_lin != nullptr;
    win.lin = 0;
#pragma unroll 1
    for (int d = 0; d < r; ++d) {
      win.interior = win.interior && ix[d] + P.min_off[d] >= 0 && ix[d] + P.max_off[d] < P.in_dims[d];
      win.lin = win.lin * P.in_dims[d] + ix[d];
    }
    T res;
    if (P.kind == MB_EXPR) {
      // Applicative / Num instances of Stencil (Stencil/Internal.hs:93-174): every term at the same index, then
      // the postfix program over the values
      T tv[MB_MAX_TERMS];
      for (int t = 0; t < P.X.n_terms; ++t)
        tv[t] = eval_kind<T>(P, P.X.term_kind[t], P.X.term_begin[t], P.X.term_ntaps[t], P.X.term_avg_n[t], in, ix, fillv, win);
      res = run_program<T>(P.X, tv);
    } else {
      res = eval_kind<T>(P, P.kind, 0, P.ntaps, P.avg_n, in, ix, fillv, win);
    }
    out[lin_out] = res;
  }
}

template <typename T>
static int launch_t(Ctx *c, const GParams &P, const void *in, void *out) {
  const int block = 256;
  int64_t blocks = (P.ncells + block - 1) / block;
  const int64_t cap = (int64_t)c->num_sms * 32;
  if (blocks > cap) blocks = cap;
  generic_kernel<T><<<(unsigned)blocks, block, 0, c->stream>>>(P, (const T *)in, (T *)out);
  return check_cuda(c, cudaGetLastError(), "generic_kernel launch");
}

int launch_generic(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p = dp.p;
  GParams P;
  std::memset(&P, 0, sizeof(P));
  P.rank = p.rank; P.border = p.border; P.kind = p.kind;
  P.bool_canon = p.elem == MB_BOOL32;
  P.skip_zero = (p.flags & MB_FLAG_ASSUME_FINITE) != 0;
  int64_t plane = 1;
  for (int i = 0; i < p.rank; ++i) {
    P.in_dims[i] = p.in_dims[i]; P.out_dims[i] = p.out_dims[i];
    P.stride[i] = p.stride[i]; P.shift[i] = p.shift[i];
    if (i > 0) plane *= p.out_dims[i];
  }
  P.ntaps = p.ntaps; P.avg_n = p.avg_n;
  P.taps = dp.d_taps; P.c1 = dp.d_c1; P.c2 = dp.d_c2;
  P.tap_lin = dp.d_tap_lin;
  P.small = p.out_elems < ((int64_t)1 << 31);
  for (int i = 0; i < p.rank; ++i) { P.min_off[i] = p.min_off[i]; P.max_off[i] = p.max_off[i]; }
  std::memcpy(&P.fill_bits, p.fill, 8);
  P.X = make_xparams(p);
  int64_t o0 = 0, o1 = p.out_dims[0];
  if (range) { o0 = range->o0; o1 = range->o1; }
  P.cell0 = o0 * plane;
  P.ncells = (o1 - o0) * plane;
  if (P.ncells <= 0) return MB_OK;
  c->launches++;
  c->last_kernel = "generic";
  switch (p.elem) {
    case MB_U8: return launch_t<uint8_t>(c, P, d_in, d_out);
    case MB_I8: return launch_t<int8_t>(c, P, d_in, d_out);
    case MB_U16: return launch_t<uint16_t>(c, P, d_in, d_out);
    case MB_I16: return launch_t<int16_t>(c, P, d_in, d_out);
    case MB_U32: return launch_t<uint32_t>(c, P, d_in, d_out);
    case MB_I32: case MB_BOOL32: return launch_t<int32_t>(c, P, d_in, d_out);
    case MB_U64: return launch_t<uint64_t>(c, P, d_in, d_out);
    case MB_I64: return launch_t<int64_t>(c, P, d_in, d_out);
    case MB_F32: return launch_t<float>(c, P, d_in, d_out);
    case MB_F64: return launch_t<double>(c, P, d_in, d_out);
  }
  return fail(c, MB_ERR_INVALID_DESC, "unknown element type");
}

}  // namespace mb
