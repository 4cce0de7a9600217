// Strided small windows over the two innermost dimensions: `computeWithStride stride (mapStencil b st)`
// for pooling (max/min/sum/avg/product stencils, Array/Stencil.hs:360-481) and strided convolution /
// correlation (StencilSpec.hs:298-309).  With a stride the windows of neighbouring outputs overlap
// little or not at all, so there is nothing to stage: every thread owns one output, the lanes of a
// warp sit on consecutive outputs of a row, and a window row is read with KW loads that the warp
// coalesces into one contiguous stretch.  Index arithmetic is 32-bit and done once per output; outputs
// whose window leaves the array take the per-tap border path.  Taps are visited in the reference's
// order (row-major over the kernel; a convolution walks the source backwards), unfused IEEE operations.
#include "mb_arith.cuh"
#include "mb_internal.h"

namespace mb {

enum { D2_SUM = 0, D2_PRODUCT, D2_MAX, D2_MIN, D2_LIN };

struct D2Params {
  int A;                 // product of the outer dimensions (the window has extent 1 there)
  int H, W, OH, OW;
  int KH, KW, sy, sx;
  int oy, ox;            // source index of the window's first cell for output (0, 0)
  int op, rev, avg, border;
  long long avg_n;
  unsigned long long fill_bits;
  const void *coeffs;    // KH*KW elements, kernel order
};

template <typename T, int KWT>
__global__ void __launch_bounds__(256) direct2d_kernel(const D2Params P, const T *__restrict__ in, T *__restrict__ out) {
  using A = Arith<T>;
  using Acc = typename A::Acc;
  __shared__ T ksm[64];
  const int KW = KWT ? KWT : P.KW, KH = P.KH;
  if (P.op == D2_LIN) {
    for (int t = threadIdx.x; t < KH * KW; t += blockDim.x) ksm[t] = ((const T *)P.coeffs)[t];
    __syncthreads();
  }
  T fillv;
  {
    unsigned long long b = P.fill_bits;
    memcpy(&fillv, &b, sizeof(T));
  }
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy_l = threadIdx.x >> 5;
  if (ox >= P.OW) return;
  const int x0 = ox * P.sx + P.ox;
  const bool in_x = x0 >= 0 && x0 + KW <= P.W;
  for (int a = blockIdx.z; a < P.A; a += gridDim.z) {
    const T *img = in + (size_t)a * P.H * P.W;
    for (int oyb = blockIdx.y * 8; oyb < P.OH; oyb += gridDim.y * 8) {
      const int oy = oyb + oy_l;
      if (oy >= P.OH) continue;
      const int y0 = oy * P.sy + P.oy;
      const bool inside = in_x && y0 >= 0 && y0 + KH <= P.H;
      // tap (r, c) in kernel order; its source cell inside the window is (rr, cc)
      auto visit = [&](auto &&f) {
        for (int r = 0; r < KH; ++r) {
          const int rr = P.rev ? KH - 1 - r : r;
          if (inside) {
            const T *row = img + (size_t)(y0 + rr) * P.W + x0;
            if (KWT) {
              T v[KWT ? KWT : 1];
#pragma unroll
              for (int c = 0; c < KWT; ++c) v[c] = row[P.rev ? KWT - 1 - c : c];
#pragma unroll
              for (int c = 0; c < KWT; ++c) f(r * KWT + c, v[c]);
            } else {
#pragma unroll 4
              for (int c = 0; c < KW; ++c) f(r * KW + c, row[P.rev ? KW - 1 - c : c]);
            }
          } else {
            int y = y0 + rr;
            const bool y_ok = resolve_border32(P.border, P.H, y);
            for (int c = 0; c < KW; ++c) {
              int x = x0 + (P.rev ? KW - 1 - c : c);
              T v = fillv;
              if (y_ok && resolve_border32(P.border, P.W, x)) v = img[(size_t)y * P.W + x];
              f(r * KW + c, v);
            }
          }
        }
      };
      T res;
      switch (P.op) {
        case D2_SUM: {
          Acc acc = A::from_count(0);
          visit([&](int, T x) { acc = A::add(acc, A::load(x)); });
          if (P.avg) acc = A::div(acc, A::from_count(P.avg_n));
          res = A::store(acc);
          break;
        }
        case D2_PRODUCT: {
          Acc acc = A::from_count(1);
          visit([&](int, T x) { acc = A::mul(acc, A::load(x)); });
          res = A::store(acc);
          break;
        }
        case D2_MAX: {
          T acc = Bounds<T>::lo();
          visit([&](int, T x) { acc = x > acc ? x : acc; });
          res = acc;
          break;
        }
        case D2_MIN: {
          T acc = Bounds<T>::hi();
          visit([&](int, T x) { acc = x < acc ? x : acc; });
          res = acc;
          break;
        }
        default: {  // D2_LIN: acc = x * k + acc (Stencil/Convolution.hs:74-78, 115-119)
          Acc acc = A::from_count(0);
          visit([&](int t, T x) { acc = A::add(A::mul(A::load(x), A::load(ksm[t])), acc); });
          res = A::store(acc);
          break;
        }
      }
      out[((size_t)a * P.OH + oy) * P.OW + ox] = res;
    }
  }
}

template <typename T>
static int launch_d2(Ctx *c, const D2Params &P, const void *d_in, void *d_out) {
  dim3 grid((unsigned)((P.OW + 31) / 32), (unsigned)std::min((P.OH + 7) / 8, 16384), (unsigned)std::min(P.A, 64));
  switch (P.KW) {
    case 2: direct2d_kernel<T, 2><<<grid, 256, 0, c->stream>>>(P, (const T *)d_in, (T *)d_out); break;
    case 3: direct2d_kernel<T, 3><<<grid, 256, 0, c->stream>>>(P, (const T *)d_in, (T *)d_out); break;
    case 4: direct2d_kernel<T, 4><<<grid, 256, 0, c->stream>>>(P, (const T *)d_in, (T *)d_out); break;
    default: direct2d_kernel<T, 0><<<grid, 256, 0, c->stream>>>(P, (const T *)d_in, (T *)d_out); break;
  }
  c->launches++;
  c->last_kernel = "direct2d";
  return check_cuda(c, cudaGetLastError(), "direct2d_kernel launch");
}

// -1: not a strided small window over the two innermost dimensions
int launch_direct2d(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p = dp.p;
  if (range || p.rank < 2 || p.elem == MB_BOOL32 || p.out_elems < c->tile_min_elems) return -1;
  D2Params P;
  std::memset(&P, 0, sizeof(P));
  switch (p.kind) {
    case MB_SUM: P.op = D2_SUM; break;
    case MB_AVG: P.op = D2_SUM; P.avg = 1; break;
    case MB_PRODUCT: P.op = D2_PRODUCT; break;
    case MB_MAX: P.op = D2_MAX; break;
    case MB_MIN: P.op = D2_MIN; break;
    case MB_CONV: P.op = D2_LIN; P.rev = 1; break;
    case MB_CORR: P.op = D2_LIN; break;
    default: return -1;
  }
  if (p.flags & MB_FLAG_ASSUME_FINITE) { /* zero taps are not skipped here: identical for finite data */ }
  const int r = p.rank, iy = r - 2, ix = r - 1;
  if (p.unit_stride()) return -1;  // unit stride: the tiled kernels (or the generic one) own it
  int64_t a = 1;
  for (int i = 0; i < iy; ++i) {
    if (p.size[i] != 1 || p.stride[i] != 1 || p.shift[i] != 0 || p.out_dims[i] != p.in_dims[i]) return -1;
    a *= p.in_dims[i];
  }
  if (p.size[iy] < 1 || p.size[ix] < 1 || p.size[iy] * p.size[ix] > 64 || p.size[iy] * p.size[ix] < 1) return -1;
  const int64_t lim = (int64_t)1 << 30;
  if (a >= lim || p.in_dims[iy] >= lim || p.in_dims[ix] >= lim || p.in_dims[iy] < 1 || p.in_dims[ix] < 1) return -1;
  if (p.stride[iy] >= lim || p.stride[ix] >= lim || p.out_dims[iy] * p.stride[iy] >= lim || p.out_dims[ix] * p.stride[ix] >= lim) return -1;
  P.A = (int)a; P.H = (int)p.in_dims[iy]; P.W = (int)p.in_dims[ix];
  P.OH = (int)p.out_dims[iy]; P.OW = (int)p.out_dims[ix];
  P.KH = (int)p.size[iy]; P.KW = (int)p.size[ix];
  P.sy = (int)p.stride[iy]; P.sx = (int)p.stride[ix];
  P.oy = (int)(p.shift[iy] + p.min_off[iy]); P.ox = (int)(p.shift[ix] + p.min_off[ix]);
  P.border = p.border; P.avg_n = p.avg_n;
  std::memcpy(&P.fill_bits, p.fill, 8);
  P.coeffs = dp.d_c1;
  if (P.op == D2_LIN && !P.coeffs) return -1;
  switch (p.elem) {
    case MB_U8: return launch_d2<uint8_t>(c, P, d_in, d_out);
    case MB_I8: return launch_d2<int8_t>(c, P, d_in, d_out);
    case MB_U16: return launch_d2<uint16_t>(c, P, d_in, d_out);
    case MB_I16: return launch_d2<int16_t>(c, P, d_in, d_out);
    case MB_U32: return launch_d2<uint32_t>(c, P, d_in, d_out);
    case MB_I32: return launch_d2<int32_t>(c, P, d_in, d_out);
    case MB_U64: return launch_d2<uint64_t>(c, P, d_in, d_out);
    case MB_I64: return launch_d2<int64_t>(c, P, d_in, d_out);
    case MB_F32: return launch_d2<float>(c, P, d_in, d_out);
    case MB_F64: return launch_d2<double>(c, P, d_in, d_out);
  }
  return -1;
}

}  // namespace mb
