// Line-window kernels: stencils whose window runs along ONE dimension (size 1 everywhere else) —
// the integral-approximation rules (Array/Numeric/Integral.hs:54-118, evaluated with a stride by
// integrateWith :121-135), and the fold stencils with 1-D windows (sum/avg/product/max/min pooling
// along a dimension).  The array is viewed as (A, D, B): D the window's dimension, A the product of
// the dimensions outside it, B the product of those inside it.
//
//   B > 1  "columns": a thread owns one output cell and walks its window down the D axis; the lanes
//          of a warp sit on consecutive B positions, so every load is one coalesced request.
//   B == 1 "rows": the window runs along the contiguous axis.  A warp takes 32 consecutive outputs
//          of one row; their windows are fetched with coalesced loads (consecutive lanes read
//          consecutive elements), parked in a per-warp shared-memory buffer with an odd pitch, and
//          each lane then folds ITS window sequentially out of shared memory without bank
//          conflicts.  Long windows go through the buffer in chunks; the fold state carries over.
// The fold is sequential per output in the reference's order (no tree reduction), so results are
// bit-identical.  The border rule applies along D only (the window has no extent elsewhere).
#include "mb_arith.cuh"
#include "mb_internal.h"
#include "mb_tma.cuh"

namespace mb {

enum { LS_SUM = 0, LS_PRODUCT, LS_MAX, LS_MIN, LS_MIDPOINT, LS_TRAPEZOID, LS_SIMPSON };

struct LSParams {
  int64_t A, D, B, OD;   // view of the input (A, D, B) and output extent along D
  int K;                   // window length
  int64_t s, shift;      // stride along D; source index of window element 0 of output 0
  int border, avg;
  int64_t avg_n;
  unsigned long long fill_bits, dx_bits;
};

// the per-output fold: feed the window's elements in order, then finish
template <typename T, int OP> struct Fold {
  using A = Arith<T>;
  using Acc = typename A::Acc;
  Acc acc, aux, held;
  __device__ __forceinline__ void init() {
    aux = held = A::from_count(0);
    if (OP == LS_PRODUCT) acc = A::from_count(1);
    else if (OP == LS_MAX) acc = A::load(Bounds<T>::lo());
    else if (OP == LS_MIN) acc = A::load(Bounds<T>::hi());
    else acc = A::from_count(0);
  }
  // i = position in the window, n = K - 1
  __device__ __forceinline__ void step(int i, int n, T xv) {
    const Acc x = A::load(xv);
    if (OP == LS_SUM || OP == LS_MIDPOINT) acc = A::add(acc, x);
    else if (OP == LS_PRODUCT) acc = A::mul(acc, x);
    else if (OP == LS_MAX) { if (xv > A::store(acc)) acc = x; }
    else if (OP == LS_MIN) { if (xv < A::store(acc)) acc = x; }
    else if (OP == LS_TRAPEZOID) {  // Integral.hs:85-89: g 0, then + 2 * g i for 0 < i < n, then + g n
      if (i == 0) acc = x;
      else if (i < n) acc = A::add(acc, A::mul(A::from_count(2), x));
      else acc = A::add(acc, x);
    } else {  // LS_SIMPSON, Integral.hs:112-117: acc + prev + 4 * g (i + 1) + fx3, left to right
      if (i == 0) aux = x;                                   // prev = g 0
      else if (i & 1) held = A::mul(A::from_count(4), x);    // 4 * g (i + 1)
      else { acc = A::add(A::add(A::add(acc, aux), held), x); aux = x; }
    }
  }
  // a whole window at once: at(i) yields element i; same operations, no per-element dispatch
  template <typename At> __device__ __forceinline__ void run(At at, int K) {
    const int n = K - 1;
    if (OP == LS_TRAPEZOID) {
      acc = A::load(at(0));
      for (int i = 1; i < n; ++i) acc = A::add(acc, A::mul(A::from_count(2), A::load(at(i))));
      acc = A::add(acc, A::load(at(n)));
    } else if (OP == LS_SIMPSON) {
      Acc prev = A::load(at(0));
      for (int i = 0; i < n - 1; i += 2) {
        const Acc fx3 = A::load(at(i + 2));
        const Acc mid = A::mul(A::from_count(4), A::load(at(i + 1)));
        acc = A::add(A::add(A::add(acc, prev), mid), fx3);
        prev = fx3;
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < K; ++i) step(i, n, at(i));
    }
  }
  __device__ __forceinline__ T finish(const LSParams &P) const {
    Acc r = acc;
    if (OP == LS_SUM && P.avg) r = A::div(r, A::from_count(P.avg_n));
    if (OP == LS_MIDPOINT || OP == LS_TRAPEZOID || OP == LS_SIMPSON) {
      T dxv;
      unsigned long long b = P.dx_bits;
      memcpy(&dxv, &b, sizeof(T));
      Acc dx = A::load(dxv);
      if (OP == LS_TRAPEZOID) dx = A::div(dx, A::from_count(2));
      if (OP == LS_SIMPSON) dx = A::div(dx, A::from_count(3));
      r = A::mul(dx, r);
    }
    return A::store(r);
  }
};

template <typename T> __device__ __forceinline__ T ls_fill(const LSParams &P) {
  T v;
  unsigned long long b = P.fill_bits;
  memcpy(&v, &b, sizeof(T));
  return v;
}

// ---- B > 1 -----------------------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(256) linescan_cols_kernel(const LSParams P, const T *__restrict__ in, T *__restrict__ out) {
  const T fillv = ls_fill<T>(P);
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  const int n = P.K - 1;
  for (int64_t a = blockIdx.z; a < P.A; a += gridDim.z) {
    const T *src = in + a * P.D * P.B + b;
    for (int64_t o = blockIdx.y; o < P.OD; o += gridDim.y) {
      const int64_t y0 = o * P.s + P.shift;
      Fold<T, OP> f;
      f.init();
      if (y0 >= 0 && y0 + P.K <= P.D) {  // the whole window is inside the array
        const T *p = src + y0 * P.B;
#pragma unroll 4
        for (int i = 0; i < P.K; ++i) f.step(i, n, p[(int64_t)i * P.B]);
      } else {
        for (int i = 0; i < P.K; ++i) {
          int64_t y = y0 + i;
          T x = fillv;
          if (resolve_border(P.border, P.D, y)) x = src[y * P.B];
          f.step(i, n, x);
        }
      }
      out[(a * P.OD + o) * P.B + b] = f.finish(P);
    }
  }
}

// ---- B == 1 ----------------------------------------------------------------------------------------
// Two ways to park the windows of a warp's 32 outputs in its buffer:
//   SPAN  (windows tile or touch: K <= s + 1, modest s): the segment's inputs are ONE contiguous span;
//         it is copied as is, with PAD elements inserted after every s so that the lanes' window
//         starts, s + PAD apart, fall on different banks ((s + PAD) odd).  No index arithmetic per
//         element beyond a shift (s a power of two) or nothing at all (s odd).
//   piece (anything else): CH elements of every window per pass, lane j's piece at j * PITCH (odd).
template <typename T, int OP, bool SPAN>
__global__ void __launch_bounds__(256) linescan_rows_kernel(const LSParams P, const T *__restrict__ in, T *__restrict__ out,
                                                              int CH, int PITCH, int64_t nseg, int64_t nitems, int PAD,
                                                              int sshift, int buf_elems, int OPL) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T *buf = reinterpret_cast<T *>(smem_raw) + (size_t)warp * buf_elems;
  const T fillv = ls_fill<T>(P);
  const int n = P.K - 1;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t item = blockIdx.x * (int64_t)(blockDim.x >> 5) + warp; item < nitems; item += wstride) {
    const int64_t row = item / nseg, seg = item - row * nseg;
    const int per_item = 32 * OPL;  // outputs per item: OPL per lane (short windows need more than 32 to amortise an item)
    const int64_t o0 = seg * per_item;
    const int nout = (int)min((int64_t)per_item, (int64_t)(P.OD - o0));
    const T *src = in + row * P.D;
    const int64_t x0 = o0 * P.s + P.shift;  // window element 0 of the segment's first output
    const bool inside = x0 >= 0 && x0 + (int64_t)(nout - 1) * P.s + P.K <= P.D;
    Fold<T, OP> f;
    f.init();
    if constexpr (SPAN) {
      const int si = (int)P.s;
      const int span = (nout - 1) * si + P.K;
      constexpr int U = 4;
      if (sizeof(T) >= 4 && inside) {
        for (int e = lane; e < span; e += 32)
          cp_async_elem<(sizeof(T) >= 4 ? (int)sizeof(T) : 4)>(&buf[e + (PAD ? (sshift >= 0 ? e >> sshift : e / si) : 0)], &src[x0 + e]);
        cp_async_wait_all();
      } else
      for (int e0 = 0; e0 < span; e0 += 32 * U) {
        T v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int e = e0 + 32 * u + lane;
          v[u] = fillv;
          if (e < span) {
            if (inside) v[u] = src[x0 + e];
            else { int64_t x = x0 + e; if (resolve_border(P.border, P.D, x)) v[u] = src[x]; }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int e = e0 + 32 * u + lane;
          if (e < span) buf[e + (PAD ? (sshift >= 0 ? e >> sshift : e / si) : 0)] = v[u];
        }
      }
      __syncwarp();
      for (int oj = lane; oj < nout; oj += 32) {
        const T *w = buf + oj * (si + PAD);
        Fold<T, OP> g;
        g.init();
        g.run([&](int i) { return w[i + (i >= si ? PAD : 0)]; }, P.K);
        out[row * P.OD + o0 + oj] = g.finish(P);
      }
      __syncwarp();
      continue;
    } else {
      for (int c0 = 0; c0 < P.K; c0 += CH) {
        const int L = min(CH, P.K - c0);
        // fetch: piece j = elements [c0, c0 + L) of output j's window; consecutive lanes read consecutive
        // elements of a piece; several loads are issued before the first store (requests in flight)
        constexpr int U = 8;
        int j = lane / L, i = lane - j * L;
        if (sizeof(T) >= 4 && inside) {
          while (j < nout) {  // no register staging: every copy of the lane's share is in flight at once
            cp_async_elem<(sizeof(T) >= 4 ? (int)sizeof(T) : 4)>(&buf[j * PITCH + i], &src[x0 + (int64_t)j * P.s + c0 + i]);
            i += 32;
            while (i >= L) { i -= L; ++j; }
          }
          cp_async_wait_all();
        }
        while (j < nout) {
          int pos[U];
          T v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            pos[u] = j < nout ? j * PITCH + i : -1;
            int64_t x = x0 + (int64_t)j * P.s + c0 + i;
            v[u] = fillv;
            if (j < nout) {
              if (inside) v[u] = src[x];
              else if (resolve_border(P.border, P.D, x)) v[u] = src[x];
            }
            i += 32;
            while (i >= L) { i -= L; ++j; }
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (pos[u] >= 0) buf[pos[u]] = v[u];
        }
        __syncwarp();
        if (lane < nout) {
          const T *w = buf + lane * PITCH;
#pragma unroll 4
          for (int t = 0; t < L; ++t) f.step(c0 + t, n, w[t]);
        }
        __syncwarp();
      }
    }
    if (lane < nout) out[row * P.OD + o0 + lane] = f.finish(P);
  }
}

template <typename T, int OP>
static int launch_ls(Ctx *c, const LSParams &P, const void *d_in, void *d_out) {
  if (P.B > 1) {
    dim3 grid((unsigned)((P.B + 255) / 256), (unsigned)std::min<int64_t>(P.OD, 65535), (unsigned)std::min<int64_t>(P.A, 1024));
    // few columns: fewer threads per block would idle less, but such arrays are tiny anyway
    linescan_cols_kernel<T, OP><<<grid, 256, 0, c->stream>>>(P, (const T *)d_in, (T *)d_out);
    c->last_kernel = "linescan_cols";
  } else {
    const int CH = P.K < 32 ? P.K : 32;
    const int PITCH = CH | 1;
    const bool span = P.K <= P.s + 1 && P.s <= 64;
    int PAD = 0, sshift = -1, buf_elems = 32 * PITCH, OPL = 1;
    if (span) {
      PAD = (P.s % 2 == 0) ? 1 : 0;
      for (int b = 0; b < 7; ++b) if (((int64_t)1 << b) == P.s) sshift = b;
      OPL = (int)std::max<int64_t>(1, std::min<int64_t>(8, 512 / (32 * P.s)));  // ~512 elements per item and warp
      buf_elems = (32 * OPL - 1) * (int)P.s + P.K + 32 * OPL * PAD + 1;
    }
    const int64_t nseg = (P.OD + 32 * OPL - 1) / (32 * OPL), nitems = P.A * nseg;
    const size_t smem = (size_t)8 * buf_elems * sizeof(T);
    int64_t blocks = (nitems + 7) / 8;
    const int64_t cap = (int64_t)c->num_sms * 8;
    if (blocks > cap) blocks = cap;
    auto kern = span ? linescan_rows_kernel<T, OP, true> : linescan_rows_kernel<T, OP, false>;
    static bool attr_set[64][2] = {{false}};
    if (!attr_set[c->device & 63][span]) {
      MB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (31 * 64 + 65 + 33) * (int)sizeof(T)));
      attr_set[c->device & 63][span] = true;
    }
    kern<<<(unsigned)blocks, 256, smem, c->stream>>>(P, (const T *)d_in, (T *)d_out, CH, PITCH, nseg, nitems, PAD, sshift, buf_elems, OPL);
    c->last_kernel = "linescan_rows";
  }
  c->launches++;
  return check_cuda(c, cudaGetLastError(), "linescan kernel launch");
}

template <typename T>
static int launch_ls_fold(Ctx *c, int kind, const LSParams &P, const void *i, void *o) {
  switch (kind) {
    case MB_SUM: case MB_AVG: return launch_ls<T, LS_SUM>(c, P, i, o);
    case MB_PRODUCT: return launch_ls<T, LS_PRODUCT>(c, P, i, o);
    case MB_MAX: return launch_ls<T, LS_MAX>(c, P, i, o);
    case MB_MIN: return launch_ls<T, LS_MIN>(c, P, i, o);
  }
  return -1;
}

template <typename T>
static int launch_ls_float(Ctx *c, int kind, const LSParams &P, const void *i, void *o) {
  switch (kind) {
    case MB_SUM: case MB_AVG: return launch_ls<T, LS_SUM>(c, P, i, o);
    case MB_PRODUCT: return launch_ls<T, LS_PRODUCT>(c, P, i, o);
    case MB_INT_MIDPOINT: return launch_ls<T, LS_MIDPOINT>(c, P, i, o);
    case MB_INT_TRAPEZOID: return launch_ls<T, LS_TRAPEZOID>(c, P, i, o);
    case MB_INT_SIMPSON: return launch_ls<T, LS_SIMPSON>(c, P, i, o);
  }
  return -1;
}

// -1: not a line window (or a shape this family does not take); the caller tries the next family
int launch_linescan(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p = dp.p;
  if (range || p.elem == MB_BOOL32 || p.out_elems < c->tile_min_elems) return -1;
  switch (p.kind) {
    case MB_SUM: case MB_AVG: case MB_PRODUCT: case MB_MAX: case MB_MIN:
    case MB_INT_MIDPOINT: case MB_INT_TRAPEZOID: case MB_INT_SIMPSON: break;
    default: return -1;
  }
  int d = -1;
  for (int i = 0; i < p.rank; ++i)
    if (p.size[i] != 1) { if (d >= 0) return -1; d = i; }
  if (d < 0) return -1;  // a 1-element window: nothing to gain
  if (p.size[d] < 2 || p.size[d] > (1 << 24)) return -1;
  // a line window that is also a plain small 2-D window with unit stride belongs to the tiled kernel
  if (p.unit_stride() && p.rank == 2 && p.size[d] <= 7 && (p.kind == MB_SUM || p.kind == MB_AVG || p.kind == MB_PRODUCT ||
                                                           p.kind == MB_MAX || p.kind == MB_MIN)) return -1;
  LSParams P;
  std::memset(&P, 0, sizeof(P));
  P.A = 1; P.B = 1;
  for (int i = 0; i < p.rank; ++i) {
    if (i == d) continue;
    // no extent outside the window's dimension: the cell grid there must be the array's own
    if (p.stride[i] != 1 || p.shift[i] != 0 || p.out_dims[i] != p.in_dims[i]) return -1;
    if (i < d) P.A *= p.in_dims[i]; else P.B *= p.in_dims[i];
  }
  P.D = p.in_dims[d]; P.OD = p.out_dims[d];
  if (P.D < 1) return -1;
  P.K = (int)p.size[d]; P.s = p.stride[d]; P.shift = p.shift[d];
  P.border = p.border; P.avg = p.kind == MB_AVG; P.avg_n = p.avg_n;
  std::memcpy(&P.fill_bits, p.fill, 8);
  if (!p.c1.empty()) std::memcpy(&P.dx_bits, p.c1.data(), p.c1.size() < 8 ? p.c1.size() : 8);
  switch (p.elem) {
    case MB_U8: return launch_ls_fold<uint8_t>(c, p.kind, P, d_in, d_out);
    case MB_I8: return launch_ls_fold<int8_t>(c, p.kind, P, d_in, d_out);
    case MB_U16: return launch_ls_fold<uint16_t>(c, p.kind, P, d_in, d_out);
    case MB_I16: return launch_ls_fold<int16_t>(c, p.kind, P, d_in, d_out);
    case MB_U32: return launch_ls_fold<uint32_t>(c, p.kind, P, d_in, d_out);
    case MB_I32: return launch_ls_fold<int32_t>(c, p.kind, P, d_in, d_out);
    case MB_U64: return launch_ls_fold<uint64_t>(c, p.kind, P, d_in, d_out);
    case MB_I64: return launch_ls_fold<int64_t>(c, p.kind, P, d_in, d_out);
    case MB_F32: return launch_ls_float<float>(c, p.kind, P, d_in, d_out);
    case MB_F64: return launch_ls_float<double>(c, p.kind, P, d_in, d_out);
  }
  return -1;
}

}  // namespace mb
