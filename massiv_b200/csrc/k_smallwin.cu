// "Any small window" rank-2 kernel: every stencil the descriptor can express over a window of up to 16 x 16 cells,
// for every Storable element type — run-time window shape, run-time tap lists, stencil expressions (MB_EXPR: the
// Applicative / Num / Fractional / Floating instances of Stencil, Array/Stencil/Internal.hs:83-174) and post-maps
// evaluated in ONE pass over HBM.  It takes what the compile-time-shaped kernels (k_tile2d*.cu) decline: closure-form
// convolutions / correlations (MB_TAPS, Stencil/Convolution.hs:44-57, 89-100 — the form the reference's own Sobel
// benchmark uses, massiv-bench/src/Data/Massiv/Bench/Sobel.hs:15-44), windows other than 3x3 / 5x5 / 7x7 / 2x2 /
// 1xk / kx1, Word8 / Word16 / Int16 ... images, sums / products of several stencils, fused fmap.
//
// Same skeleton as k_tile2d.cuh — persistent CTAs, the tile's input box staged by ONE TMA box load into an mbarrier
// ring (cp.async copies for arrays without a tensor map, a peeled path through handleBorderIndex for edge tiles
// under borders other than Fill 0) — but the arithmetic is table driven: every term is a list of (shared-memory
// offset, coefficient) pairs that the CTA keeps in shared memory; a thread owns 4 x 4 outputs whose columns are 32
// apart, so that for one tap a warp reads 32 CONSECUTIVE elements of a box row: no bank conflicts for any element
// size, window offset or alignment.  The bound is the shared-memory pipe: one LDS per tap and output,
// 32 lanes per clock and SM.  Tap order and the one-rounding-per-operation arithmetic are the reference's
// (mb_arith.cuh), so results are bit-identical to the generic kernel's and to massiv's.
#include <cstdlib>

#include "mb_arith.cuh"
#include "mb_internal.h"
#include "mb_post.cuh"
#include "mb_tma.cuh"

namespace mb {

enum { SW_TW = 128, SW_TH = 32, SW_CX = 4, SW_RY = 4, SW_S = 3, SW_MAX_TAPS = 1024, SW_MAX_WIN = 16 };


struct SWProgram {
  int n_terms, n_prog;
  int term_kind[MB_MAX_TERMS], term_begin[MB_MAX_TERMS], term_ntaps[MB_MAX_TERMS], term_avg_n[MB_MAX_TERMS];
  unsigned char op[MB_MAX_PROGRAM + MB_MAX_POST + 8], arg[MB_MAX_PROGRAM + MB_MAX_POST + 8];
  unsigned long long operand[MB_MAX_PROGRAM + MB_MAX_POST + 8];
};

struct SWParams {
  int rank;              // 2, or 3: a tile is 32 x 128 cells of ONE output plane, its box carries the window's depth
  int in_d, in_h, in_w, out_d, out_h, out_w;
  int shift_z, shift_y, shift_x;  // source index of the window's first cell for output (0, 0, 0)
  int o0, o1;            // outermost output range produced: rows (rank 2) or planes (rank 3)
  int tiles_x, tiles_y, ntiles;
  int bd, nstages;       // box depth (1 for rank 2); stages of the ring (3, or 2 when the boxes are large)
  int min_dz;
  int border, use_tma, zero_fill_ok;
  int xoff;              // the window begins xoff elements into the staged row (boxes start on 16-byte boundaries)
  int bw, bh;            // staged box, elements (bw a multiple of the vector length)
  int wused;             // columns [xoff, wused) of a staged row are read
  int stage_bytes;
  int ntaps;             // all terms together
  int bool_canon, skip_zero, trivial;  // trivial: one term, no program beyond it
  // dense (rank 2, trivial plans): the taps are a full kh x kw window walked row-major (1) or in reverse (2): a loaded
  // element then serves all the thread's output rows it belongs to, in the order each of them needs it
  int dense, kh, kw;
  unsigned long long fill_bits;
  const int32_t *taps;   // ntaps x 2 offsets relative to the evaluated index (the plan's table)
  const void *coef;      // ntaps coefficients (zeros where a term has none)
  int min_dy, min_dx;
  SWProgram X;
};

template <typename T> struct SWAcc {
  using A = Arith<T>;
  using Acc = typename A::Acc;
};

// one term over RP rows x 4 columns (32 apart) of the thread's outputs, starting at `base`
template <typename T, bool canon, int RP>
__device__ __forceinline__ void sw_term_rows(int kind, int nt, int avg_n, const int *__restrict__ s_off, const T *__restrict__ s_k,
                                             const T *__restrict__ base, int pitch, bool skip_zero, T (&out)[RP][SW_CX]) {
  using A = Arith<T>;
  using Acc = typename A::Acc;
  Acc acc[RP][SW_CX];
  auto ld = [&](int off, int r, int c) -> T {
    T v = base[off + r * pitch + 32 * c];
    if (canon) v = (T)(v != T(0));  // Storable Bool peek: (/= 0)
    return v;
  };
#define SW_EACH(BODY)                      \
  _Pragma("unroll") for (int r = 0; r < RP; ++r) _Pragma("unroll") for (int c = 0; c < SW_CX; ++c) { BODY; }
  switch (kind) {
    case MB_ID:
      SW_EACH(out[r][c] = ld(s_off[0], r, c))
      return;
    case MB_SUM: case MB_AVG:
      SW_EACH(acc[r][c] = A::from_count(0))
#pragma unroll 1
      for (int t = 0; t < nt; ++t) {
        const int off = s_off[t];
        SW_EACH(acc[r][c] = A::add(acc[r][c], A::load(ld(off, r, c))))
      }
      if (kind == MB_AVG) {
        const Acc n = A::from_count(avg_n);
        SW_EACH(acc[r][c] = A::div(acc[r][c], n))
      }
      break;
    case MB_PRODUCT:
      SW_EACH(acc[r][c] = A::from_count(1))
#pragma unroll 1
      for (int t = 0; t < nt; ++t) {
        const int off = s_off[t];
        SW_EACH(acc[r][c] = A::mul(acc[r][c], A::load(ld(off, r, c))))
      }
      break;
    case MB_MAX: case MB_MIN: {
      const bool mx = kind == MB_MAX;
      T m[RP][SW_CX];
      const T init = mx ? (canon ? T(0) : Bounds<T>::lo()) : (canon ? T(1) : Bounds<T>::hi());
      SW_EACH(m[r][c] = init)
#pragma unroll 1
      for (int t = 0; t < nt; ++t) {
        const int off = s_off[t];
        SW_EACH(const T x = ld(off, r, c); m[r][c] = mx ? (x > m[r][c] ? x : m[r][c]) : (x < m[r][c] ? x : m[r][c]))
      }
      SW_EACH(out[r][c] = m[r][c])
      return;
    }
    default:  // MB_CONV / MB_CORR / MB_TAPS: acc = src * k + acc in the listed order (Convolution.hs:52-55, 74-78, 115-119)
      SW_EACH(acc[r][c] = A::from_count(0))
      if (skip_zero) {
#pragma unroll 1
        for (int t = 0; t < nt; ++t) {
          const Acc kv = A::load(s_k[t]);
          if (A::is_zero(kv)) continue;
          const int off = s_off[t];
          SW_EACH(acc[r][c] = A::add(A::mul(A::load(ld(off, r, c)), kv), acc[r][c]))
        }
      } else {
        // two taps per trip: the second tap's table and data loads are in flight while the first is accumulated
        int t = 0;
#pragma unroll 1
        for (; t + 2 <= nt; t += 2) {
          const Acc k0 = A::load(s_k[t]), k1 = A::load(s_k[t + 1]);
          const int o0 = s_off[t], o1 = s_off[t + 1];
          T x0[RP][SW_CX], x1[RP][SW_CX];
          SW_EACH(x0[r][c] = ld(o0, r, c))
          SW_EACH(x1[r][c] = ld(o1, r, c))
          SW_EACH(acc[r][c] = A::add(A::mul(A::load(x0[r][c]), k0), acc[r][c]))
          SW_EACH(acc[r][c] = A::add(A::mul(A::load(x1[r][c]), k1), acc[r][c]))
        }
        if (t < nt) {
          const Acc kv = A::load(s_k[t]);
          const int off = s_off[t];
          SW_EACH(acc[r][c] = A::add(A::mul(A::load(ld(off, r, c)), kv), acc[r][c]))
        }
      }
      break;
  }
  SW_EACH(out[r][c] = A::store(acc[r][c]))
#undef SW_EACH
}

// A dense kh x kw window over the thread's 4 rows x 4 columns (32 apart): box rows are visited once, in the order the
// window is walked (top-down, or bottom-up for a reversed walk); the element of box row y and window column tx is the
// tap (y - r, tx) of output row r for every r with 0 <= y - r < kh.  Every output still accumulates its taps in the
// reference's order (row-major over the window, or its reverse), so the bits are those of the tap-list form; one
// shared-memory load now feeds up to four accumulations.
template <typename T, bool canon>
__device__ __forceinline__ void sw_dense(int kind, int kh, int kw, bool reversed, int avg_n, const T *__restrict__ s_k,
                                         const T *__restrict__ base, int pitch, bool skip_zero, T (&out)[SW_RY][SW_CX]) {
  using A = Arith<T>;
  using Acc = typename A::Acc;
  Acc acc[SW_RY][SW_CX];
  T m[SW_RY][SW_CX];
  const bool lin = kind == MB_CONV || kind == MB_CORR || kind == MB_TAPS;
  const bool mx = kind == MB_MAX;
#define SW_ALL(BODY)                          \
  _Pragma("unroll") for (int r = 0; r < SW_RY; ++r) _Pragma("unroll") for (int c = 0; c < SW_CX; ++c) { BODY; }
  SW_ALL(acc[r][c] = A::from_count(kind == MB_PRODUCT ? 1 : 0))
  SW_ALL(m[r][c] = mx ? (canon ? T(0) : Bounds<T>::lo()) : (canon ? T(1) : Bounds<T>::hi()))
  const int nrows = kh + SW_RY - 1;
  // (measured and rejected: a separate test-free loop for the box rows that lie inside all four windows — sum 9x9
  // Word8 / Float / Double 93 / 103 / 98 Gcells/s against 108 / 136 / 109 for this form)
#pragma unroll 1
  for (int i = 0; i < nrows; ++i) {
    const int y = reversed ? nrows - 1 - i : i;
    const T *rowp = base + y * pitch;
#pragma unroll 1
    for (int j = 0; j < kw; ++j) {
      const int tx = reversed ? kw - 1 - j : j;
      T x[SW_CX];
#pragma unroll
      for (int c = 0; c < SW_CX; ++c) {
        T v = rowp[tx + 32 * c];
        if (canon) v = (T)(v != T(0));
        x[c] = v;
      }
#pragma unroll
      for (int r = 0; r < SW_RY; ++r) {
        const int ty = y - r;
        if ((unsigned)ty < (unsigned)kh) {  // (uniform)
          if (lin) {
            const Acc kv = A::load(s_k[reversed ? (kh - 1 - ty) * kw + (kw - 1 - tx) : ty * kw + tx]);
            if (!(skip_zero && A::is_zero(kv))) {
#pragma unroll
              for (int c = 0; c < SW_CX; ++c) acc[r][c] = A::add(A::mul(A::load(x[c]), kv), acc[r][c]);
            }
          } else if (kind == MB_PRODUCT) {
#pragma unroll
            for (int c = 0; c < SW_CX; ++c) acc[r][c] = A::mul(acc[r][c], A::load(x[c]));
          } else if (kind == MB_MAX || kind == MB_MIN) {
#pragma unroll
            for (int c = 0; c < SW_CX; ++c) m[r][c] = mx ? (x[c] > m[r][c] ? x[c] : m[r][c]) : (x[c] < m[r][c] ? x[c] : m[r][c]);
          } else {
#pragma unroll
            for (int c = 0; c < SW_CX; ++c) acc[r][c] = A::add(acc[r][c], A::load(x[c]));
          }
        }
      }
    }
  }
  if (kind == MB_MAX || kind == MB_MIN) {
    SW_ALL(out[r][c] = m[r][c])
    return;
  }
  if (kind == MB_AVG) {
    const Acc n = A::from_count(avg_n);
    SW_ALL(acc[r][c] = A::div(acc[r][c], n))
  }
  SW_ALL(out[r][c] = A::store(acc[r][c]))
#undef SW_ALL
}

template <typename T, bool TMA, bool CANON>
__global__ void __launch_bounds__(256, 2)
smallwin_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ SWParams P, const T *__restrict__ in, T *__restrict__ out) {
  constexpr int V = 16 / (int)sizeof(T);
  extern __shared__ __align__(128) unsigned char smem[];
  const int S = P.nstages;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + S * P.stage_bytes);
  int *s_off = reinterpret_cast<int *>(full + 8);
  T *s_k = reinterpret_cast<T *>(s_off + ((P.ntaps + 3) & ~3));
  const int tid = threadIdx.x;
  if (tid == 0) {
    if (TMA) tma_prefetch_desc(&tmap);
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  const int pitch = P.bw, plane = P.bw * P.bh;
  // the tap table: shared-memory offset of every tap relative to a cell's window origin, and its coefficient
  for (int t = tid; t < P.ntaps; t += 256) {
    const int32_t *o = P.taps + P.rank * t;
    const int dz = P.rank == 3 ? o[0] - P.min_dz : 0;
    s_off[t] = dz * plane + (o[P.rank - 2] - P.min_dy) * pitch + (o[P.rank - 1] - P.min_dx) + P.xoff;
    s_k[t] = P.coef ? reinterpret_cast<const T *>(P.coef)[t] : T(0);
  }
  __syncthreads();

  const int my_tiles = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // tiles run x fastest, then y, then (rank 3) the output plane: the planes a box shares with the tile one plane
  // further are re-read out of L2 a few hundred tiles later
  // Within wave i (gridDim.x consecutive tiles) CTA b takes tile (b + i) mod gridDim.x: the tiles on the array's rim,
  // which cost a patch, would otherwise fall to the same few CTAs every wave (148 = 4 mod 8 tiles per row).
  const int full_waves = P.ntiles / (int)gridDim.x;
  auto tile_origin = [&](int i, int &oz, int &oy0, int &ox0) {
    int slot = (int)blockIdx.x;
    if (i < full_waves) {
      slot += i % (int)gridDim.x;
      if (slot >= (int)gridDim.x) slot -= (int)gridDim.x;
    }
    const int t = slot + i * (int)gridDim.x;
    ox0 = (t % P.tiles_x) * SW_TW;
    const int rest = t / P.tiles_x;
    if (P.rank == 3) {
      oy0 = (rest % P.tiles_y) * SW_TH;
      oz = P.o0 + rest / P.tiles_y;
    } else {
      oy0 = P.o0 + rest * SW_TH;
      oz = 0;
    }
  };
  auto inside = [&](int oz, int oy0, int ox0) -> bool {
    const int bz0 = oz + P.shift_z, by0 = oy0 + P.shift_y, bx0 = ox0 + P.shift_x - P.xoff;
    return bz0 >= 0 && by0 >= 0 && bx0 >= 0 && bz0 + P.bd <= P.in_d && by0 + P.bh <= P.in_h && bx0 + P.bw <= P.in_w;
  };
  // every box comes by TMA (cells outside the array arrive as zeros); a box that crosses the array's edge under a
  // border other than Fill 0 has its outside cells patched afterwards
  auto prefetch = [&](int i) {
    if (!TMA || tid != 0 || i >= my_tiles) return;
    int oz, oy0, ox0;
    tile_origin(i, oz, oy0, ox0);
    const int s = i % S;
    mbar_arrive_expect_tx(&full[s], (uint32_t)(P.bw * P.bh * P.bd * (int)sizeof(T)));
    if (P.rank == 3) tma_load_3d(smem + s * P.stage_bytes, &tmap, &full[s], ox0 + P.shift_x - P.xoff, oy0 + P.shift_y, oz + P.shift_z);
    else tma_load_2d(smem + s * P.stage_bytes, &tmap, &full[s], ox0 + P.shift_x - P.xoff, oy0 + P.shift_y);
  };
#pragma unroll 1
  for (int i = 0; i < S - 1; ++i) prefetch(i);
  T fillv;
  {
    unsigned long long b = P.fill_bits;
    memcpy(&fillv, &b, sizeof(T));
  }
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t phase = 0;
  const bool skip = P.skip_zero != 0;

  for (int i = 0; i < my_tiles; ++i) {
    const int s = i % S;
    int oz, oy0, ox0;
    tile_origin(i, oz, oy0, ox0);
    if (TMA && tid == 0 && i + S - 1 < my_tiles) fence_proxy_async();
    prefetch(i + S - 1);
    T *sm = reinterpret_cast<T *>(smem + s * P.stage_bytes);
    const bool all_in = inside(oz, oy0, ox0);
    const bool patch = !(TMA && (all_in || P.zero_fill_ok));
    if (TMA) {
      mbar_wait(&full[s], (phase >> s) & 1u);
      phase ^= 1u << s;
    }
    if (patch) {
      // box cells through the border rule (Core/Index.hs:236-253).  After a TMA fetch only the cells outside the array
      // are written: whole rows whose plane or row index is outside, else the columns left and right of the array.
      const int bz0 = oz + P.shift_z, by0 = oy0 + P.shift_y, bx0 = ox0 + P.shift_x - P.xoff;
      int cl = bx0 < 0 ? -bx0 : 0, cr = P.in_w - bx0;  // box columns [cl, cr) lie inside the array
      if (cl > P.bw) cl = P.bw;
      if (cr > P.bw) cr = P.bw;
      if (cr < cl) cr = cl;
      auto cell = [&](int q, int r, int c, int gz, int gy, bool row_ok) {
        int gx = bx0 + c;
        T v = fillv;
        if (row_ok && resolve_border32(P.border, P.in_w, gx)) v = in[((size_t)(P.rank == 3 ? gz : 0) * P.in_h + gy) * P.in_w + gx];
        sm[q * plane + r * pitch + c] = v;
      };
      // whole rows: all of them without TMA, else those whose plane or row index is outside
      for (int rr = warp; rr < P.bh * P.bd; rr += 8) {
        const int q = rr / P.bh, r = rr - q * P.bh;
        int gz = bz0 + q, gy = by0 + r;
        const bool row_in = (P.rank < 3 || (unsigned)gz < (unsigned)P.in_d) && (unsigned)gy < (unsigned)P.in_h;
        if (TMA && row_in) continue;
        const bool row_ok = row_in || ((P.rank < 3 || resolve_border32(P.border, P.in_d, gz)) && resolve_border32(P.border, P.in_h, gy));
        for (int c = lane; c < P.bw; c += 32) cell(q, r, c, gz, gy, row_ok);
      }
      if (TMA) {
        // the columns left and right of the array in the rows inside it, one cell per thread and trip
        const int l0 = P.xoff < cl ? P.xoff : cl, r1 = P.wused > cr ? P.wused : cr;
        const int nl = cl - l0, nc = nl + (r1 - cr);
        for (int item = tid; item < nc * P.bh * P.bd; item += 256) {
          const int rr = item / nc, k = item - rr * nc;
          const int q = rr / P.bh, r = rr - q * P.bh;
          const int gz = bz0 + q, gy = by0 + r;
          const bool row_in = (P.rank < 3 || (unsigned)gz < (unsigned)P.in_d) && (unsigned)gy < (unsigned)P.in_h;
          if (row_in) cell(q, r, k < nl ? l0 + k : cr + (k - nl), gz, gy, true);
        }
      }
      __syncthreads();
    }

    // ---- compute: 4 x 4 outputs per thread (columns 32 apart), RP rows at a time ----
    // The host has turned the postfix program (Stencil/Internal.hs:104-111) into steps on two fixed value slots
    // (programs that need a deeper stack take the generic kernel): a term is evaluated where the program pushes it,
    // every step updates its slot in place.  8-byte elements go two rows at a time to stay within 128 registers.
    constexpr int RP = sizeof(T) == 8 ? 2 : 4;
    const T *base = sm + (warp * SW_RY) * pitch + lane;
    if (P.dense) {
      T vd[SW_RY][SW_CX];
      sw_dense<T, CANON>(P.X.term_kind[0], P.kh, P.kw, P.dense == 2, P.X.term_avg_n[0], s_k, base + P.xoff, pitch, skip, vd);
#pragma unroll
      for (int r = 0; r < SW_RY; ++r) {
        const int gy = oy0 + warp * SW_RY + r;
        if (gy < P.o1) {
#pragma unroll
          for (int c = 0; c < SW_CX; ++c) {
            const int gx = ox0 + lane + 32 * c;
            if (gx < P.out_w) out[(size_t)gy * P.out_w + gx] = vd[r][c];
          }
        }
      }
    }
#pragma unroll 1
    for (int r0 = P.dense ? SW_RY : 0; r0 < SW_RY; r0 += RP) {
      const T *rbase = base + r0 * pitch;
      T va[RP][SW_CX], vb[RP][SW_CX];
#define SW_EACH(BODY)                      \
  _Pragma("unroll") for (int r = 0; r < RP; ++r) _Pragma("unroll") for (int c = 0; c < SW_CX; ++c) { BODY; }
      if (P.trivial) {
        sw_term_rows<T, CANON, RP>(P.X.term_kind[0], P.X.term_ntaps[0], P.X.term_avg_n[0], s_off, s_k, rbase, pitch, skip, va);
      } else {
        SW_EACH(va[r][c] = vb[r][c] = T(0))
#pragma unroll 1
        for (int k = 0; k < P.X.n_prog; ++k) {
          const int code = P.X.op[k];
          const int op = code & 31, form = (code >> 5) & 3;
          const bool on_b = (code & 0x80) != 0;
          T cv;
          {
            unsigned long long b = P.X.operand[k];
            memcpy(&cv, &b, sizeof(T));
          }
          if (form == 1) {  // slot <- slot op c
            if (on_b) { SW_EACH(vb[r][c] = xop_binary<T>(op, vb[r][c], cv)) } else { SW_EACH(va[r][c] = xop_binary<T>(op, va[r][c], cv)) }
          } else if (form == 2) {  // slot <- c op slot
            if (on_b) { SW_EACH(vb[r][c] = xop_binary<T>(op, cv, vb[r][c])) } else { SW_EACH(va[r][c] = xop_binary<T>(op, cv, va[r][c])) }
          } else if (form == 3) {  // a <- b op a
            SW_EACH(va[r][c] = xop_binary<T>(op, vb[r][c], va[r][c]))
          } else if (op == MB_X_TERM) {
            const int t = P.X.arg[k];
            if (on_b)
              sw_term_rows<T, CANON, RP>(P.X.term_kind[t], P.X.term_ntaps[t], P.X.term_avg_n[t], s_off + P.X.term_begin[t],
                                         s_k + P.X.term_begin[t], rbase, pitch, skip, vb);
            else
              sw_term_rows<T, CANON, RP>(P.X.term_kind[t], P.X.term_ntaps[t], P.X.term_avg_n[t], s_off + P.X.term_begin[t],
                                         s_k + P.X.term_begin[t], rbase, pitch, skip, va);
          } else if (op == MB_X_CONST) {
            if (on_b) { SW_EACH(vb[r][c] = cv) } else { SW_EACH(va[r][c] = cv) }
          } else if (op <= MB_X_MAX) {  // a <- a op b
            SW_EACH(va[r][c] = xop_binary<T>(op, va[r][c], vb[r][c]))
          } else if (on_b) {
            SW_EACH(vb[r][c] = xop_unary<T>(op, vb[r][c]))
          } else {
            SW_EACH(va[r][c] = xop_unary<T>(op, va[r][c]))
          }
        }
      }
#undef SW_EACH
#pragma unroll
      for (int r = 0; r < RP; ++r) {
        const int gy = oy0 + warp * SW_RY + r0 + r;
        if (gy < (P.rank == 3 ? P.out_h : P.o1)) {
#pragma unroll
          for (int c = 0; c < SW_CX; ++c) {
            const int gx = ox0 + lane + 32 * c;
            if (gx < P.out_w) out[((size_t)oz * P.out_h + gy) * P.out_w + gx] = va[r][c];
          }
        }
      }
    }
    if (TMA && patch) fence_proxy_async();  // generic-proxy writes into the stage come before its next TMA refill
    __syncthreads();  // everyone is done with stage s before it is refilled
  }
}

// The plan as terms + a program for the kernel's TWO value slots.  Single constructors are one term, MAG2 is
// sqrt (a*a + b*b) of two, MB_EXPR brings its own postfix program, post-maps are further steps (so every fmap runs in
// the store epilogue).  The postfix program is parsed into its tree and code is generated Sethi-Ullman style: a
// subtree that is a chain of unary / with-constant operators over one term needs one slot; a binary node whose
// operands both need two slots cannot be evaluated here (the generic kernel takes it).
// step code: bit 7 = works on slot B; bits 5-6 = form (0 plain, 1 slot <- slot op c, 2 slot <- c op slot,
// 3 a <- b op a); bits 0-4 = mb_xop.
namespace {
struct XNode {
  int op = 0, arg = 0;
  unsigned long long operand = 0;
  int a = -1, b = -1;  // children
};
struct XGen {
  std::vector<XNode> nodes;
  SWProgram *X;
  int n = 0;
  bool ok = true;
  void emit(int code, int arg, unsigned long long operand) {
    if (n >= (int)sizeof(X->op)) { ok = false; return; }
    X->op[n] = (unsigned char)code; X->arg[n] = (unsigned char)arg; X->operand[n] = operand;
    ++n;
  }
  bool is_const(int i) const { return nodes[i].op == MB_X_CONST; }
  bool is_binary(int i) const { return nodes[i].op >= MB_X_ADD && nodes[i].op <= MB_X_MAX; }
  // can node i be evaluated in ONE slot?
  bool single(int i) const {
    const XNode &x = nodes[i];
    if (x.op == MB_X_TERM || x.op == MB_X_CONST) return true;
    if (is_binary(i)) return (is_const(x.b) && single(x.a)) || (is_const(x.a) && single(x.b));
    return single(x.a);
  }
  void gen(int i, bool slot_b) {
    if (!ok) return;
    const XNode &x = nodes[i];
    const int sb = slot_b ? 0x80 : 0;
    if (x.op == MB_X_TERM || x.op == MB_X_CONST) { emit(x.op | sb, x.arg, x.operand); return; }
    if (!is_binary(i)) { gen(x.a, slot_b); emit(x.op | sb, 0, 0); return; }
    if (is_const(x.b)) { gen(x.a, slot_b); emit(x.op | (1 << 5) | sb, 0, nodes[x.b].operand); return; }
    if (is_const(x.a)) { gen(x.b, slot_b); emit(x.op | (2 << 5) | sb, 0, nodes[x.a].operand); return; }
    if (slot_b) { ok = false; return; }  // two live values and a third wanted
    if (single(x.b)) { gen(x.a, false); gen(x.b, true); emit(x.op, 0, 0); return; }
    if (single(x.a)) { gen(x.b, false); gen(x.a, true); emit(x.op | (3 << 5), 0, 0); return; }
    ok = false;
  }
};
}  // namespace

static bool build_program(const Plan &p, SWProgram *X, std::vector<uint8_t> *coef) {
  std::memset(X, 0, sizeof(*X));
  XGen g;
  g.X = X;
  std::vector<int> stack;
  auto leaf = [&](int op, int arg, const uint8_t *operand) {
    XNode nd;
    nd.op = op; nd.arg = arg;
    if (operand) std::memcpy(&nd.operand, operand, 8);
    g.nodes.push_back(nd);
    stack.push_back((int)g.nodes.size() - 1);
  };
  auto unary = [&](int op) {
    XNode nd;
    nd.op = op; nd.a = stack.back();
    g.nodes.push_back(nd);
    stack.back() = (int)g.nodes.size() - 1;
  };
  auto binary = [&](int op) {
    XNode nd;
    nd.op = op; nd.b = stack.back(); stack.pop_back(); nd.a = stack.back();
    g.nodes.push_back(nd);
    stack.back() = (int)g.nodes.size() - 1;
  };
  if (p.kind == MB_EXPR) {
    X->n_terms = (int)p.terms.size();
    for (int t = 0; t < X->n_terms; ++t) {
      X->term_kind[t] = p.terms[t].kind; X->term_begin[t] = p.terms[t].begin;
      X->term_ntaps[t] = p.terms[t].ntaps; X->term_avg_n[t] = p.terms[t].avg_n;
    }
    for (const mb_xins &x : p.prog) {  // validated by the planner: no underflow, one value left
      if (x.op == MB_X_TERM || x.op == MB_X_CONST) leaf(x.op, x.arg, x.operand);
      else if (x.op <= MB_X_MAX) binary(x.op);
      else unary(x.op);
    }
    *coef = p.c1;
  } else if (p.kind == MB_MAG2) {
    const int nt = (int)p.ntaps;
    const int k = (p.flags & MB_FLAG_MAG2_CORR) ? MB_CORR : MB_CONV;
    X->n_terms = 2;
    X->term_kind[0] = X->term_kind[1] = k;
    X->term_begin[0] = 0; X->term_begin[1] = nt;
    X->term_ntaps[0] = X->term_ntaps[1] = nt;
    coef->assign(p.c1.begin(), p.c1.end());
    coef->insert(coef->end(), p.c2.begin(), p.c2.end());
    // sqrt (fmap (^2) a + fmap (^2) b)  (Stencil/Internal.hs:105-146; Bench/Sobel.hs:39-44)
    leaf(MB_X_TERM, 0, nullptr); unary(MB_X_SQUARE); leaf(MB_X_TERM, 1, nullptr); unary(MB_X_SQUARE);
    binary(MB_X_ADD); unary(MB_X_SQRT);
  } else {
    switch (p.kind) {
      case MB_ID: case MB_SUM: case MB_PRODUCT: case MB_AVG: case MB_MAX: case MB_MIN: case MB_CONV: case MB_CORR: case MB_TAPS: break;
      default: return false;
    }
    X->n_terms = 1;
    X->term_kind[0] = p.kind; X->term_begin[0] = 0; X->term_ntaps[0] = (int)p.ntaps; X->term_avg_n[0] = (int)p.avg_n;
    leaf(MB_X_TERM, 0, nullptr);
    *coef = p.c1;
  }
  for (const mb_post_op &q : p.post) {  // fmap f: the same functions as operators with a constant
    switch (q.op) {
      case MB_POST_NEGATE: unary(MB_X_NEGATE); break;
      case MB_POST_ABS: unary(MB_X_ABS); break;
      case MB_POST_SIGNUM: unary(MB_X_SIGNUM); break;
      case MB_POST_SQUARE: unary(MB_X_SQUARE); break;
      case MB_POST_RECIP: unary(MB_X_RECIP); break;
      case MB_POST_SQRT: unary(MB_X_SQRT); break;
      case MB_POST_ADD: leaf(MB_X_CONST, 0, q.operand); binary(MB_X_ADD); break;
      case MB_POST_SUB: leaf(MB_X_CONST, 0, q.operand); binary(MB_X_SUB); break;
      case MB_POST_MUL: leaf(MB_X_CONST, 0, q.operand); binary(MB_X_MUL); break;
      case MB_POST_MIN: leaf(MB_X_CONST, 0, q.operand); binary(MB_X_MIN); break;
      case MB_POST_MAX: leaf(MB_X_CONST, 0, q.operand); binary(MB_X_MAX); break;
      case MB_POST_DIV: leaf(MB_X_CONST, 0, q.operand); binary(MB_X_DIV); break;
      case MB_POST_RSUB: case MB_POST_RDIV: {  // c - x, c / x: the constant is the FIRST operand
        const int x = stack.back();
        stack.pop_back();
        leaf(MB_X_CONST, 0, q.operand);
        stack.push_back(x);
        binary(q.op == MB_POST_RSUB ? MB_X_SUB : MB_X_DIV);
        break;
      }
      default: return false;
    }
  }
  if (stack.size() != 1) return false;
  g.gen(stack.back(), false);
  X->n_prog = g.n;
  return g.ok;
}

constexpr size_t SW_SMEM_MAX = 227 * 1024;  // the opt-in limit of a CTA on sm_100

template <typename T, bool CANON = false>
static int launch_sw_t(Ctx *c, const DevPlan &dp, const SWProgram &X, const void *d_coef, const void *d_taps2, const void *d_in, void *d_out,
                       const OuterRange *range) {
  constexpr int V = 16 / (int)sizeof(T);
  const Plan &p = dp.p;
  SWParams P;
  std::memset(&P, 0, sizeof(P));
  const int r = p.rank, iy = r - 2, ix = r - 1;
  P.rank = r;
  P.in_d = r == 3 ? (int)p.in_dims[0] : 1; P.in_h = (int)p.in_dims[iy]; P.in_w = (int)p.in_dims[ix];
  P.out_d = r == 3 ? (int)p.out_dims[0] : 1; P.out_h = (int)p.out_dims[iy]; P.out_w = (int)p.out_dims[ix];
  P.min_dz = r == 3 ? (int)p.min_off[0] : 0; P.min_dy = (int)p.min_off[iy]; P.min_dx = (int)p.min_off[ix];
  P.shift_z = r == 3 ? (int)(p.shift[0] + p.min_off[0]) : 0;
  P.shift_y = (int)(p.shift[iy] + p.min_off[iy]); P.shift_x = (int)(p.shift[ix] + p.min_off[ix]);
  P.o0 = range ? (int)range->o0 : 0;
  P.o1 = range ? (int)range->o1 : (int)p.out_dims[0];
  if (P.o1 <= P.o0) return MB_OK;
  P.tiles_x = (P.out_w + SW_TW - 1) / SW_TW;
  P.tiles_y = (P.out_h + SW_TH - 1) / SW_TH;
  P.ntiles = r == 3 ? P.tiles_x * P.tiles_y * (P.o1 - P.o0) : P.tiles_x * ((P.o1 - P.o0 + SW_TH - 1) / SW_TH);
  P.border = p.border;
  std::memcpy(&P.fill_bits, p.fill, 8);
  bool fill_zero = p.border == MB_FILL;
  for (int i = 0; i < p.esize; ++i) if (p.fill[i] != 0) fill_zero = false;
  P.zero_fill_ok = fill_zero;
  const int wd = r == 3 ? (int)(p.max_off[0] - p.min_off[0] + 1) : 1;
  const int wh = (int)(p.max_off[iy] - p.min_off[iy] + 1), ww = (int)(p.max_off[ix] - p.min_off[ix] + 1);
  P.xoff = ((P.shift_x % V) + V) % V;
  P.bw = ((P.xoff + SW_TW + ww - 1 + V - 1) / V) * V;
  P.bh = SW_TH + wh - 1;
  P.wused = P.xoff + SW_TW + ww - 1;
  P.bd = wd;
  P.stage_bytes = ((P.bw * P.bh * P.bd * (int)sizeof(T) + 127) / 128) * 128;
  P.ntaps = (int)p.ntaps * (p.kind == MB_MAG2 ? 2 : 1);
  P.bool_canon = p.elem == MB_BOOL32;
  P.skip_zero = (p.flags & MB_FLAG_ASSUME_FINITE) != 0;
  P.trivial = X.n_terms == 1 && X.n_prog == 1;
  // a dense window walked in (reversed) row-major order?  rank 2, one term, no program, at least three rows (below
  // that nothing is shared between the thread's output rows)
  P.dense = 0;
  // Where it pays (scripts/probe_dense.py, 16384^2, Gcells/s dense vs tap list): 8-byte elements always (the tap-list
  // form works on two rows at a time there: sum 9x9 Double 109 vs 52, conv 9x9 56 vs 48, sum 3x11 169 vs 120); sums and
  // products over seven or more rows (sum 9x9 Word8 / Float: see DESIGN.md); not the 4-byte multiply-add kernels
  // (conv 9x9 Float 52 vs 93: the per-row coefficient fetch costs more than the shared loads save) nor short windows.
  const bool dense_pays = sizeof(T) == 8 || ((p.kind == MB_SUM || p.kind == MB_AVG || p.kind == MB_PRODUCT) && wh >= 7);
  if (r == 2 && P.trivial && p.kind != MB_MAG2 && p.kind != MB_ID && wh >= 3 && (int64_t)p.ntaps == (int64_t)wh * ww &&
      (dense_pays || std::getenv("MASSIV_B200_SW_DENSE")) && !std::getenv("MASSIV_B200_SW_NO_DENSE")) {
    bool asc = true, desc = true;
    for (int t = 0; t < (int)p.ntaps; ++t) {
      const int dy = (int)(p.tap_off[2 * t] - p.min_off[0]), dx = (int)(p.tap_off[2 * t + 1] - p.min_off[1]);
      const int ty = t / ww, tx = t % ww;
      if (dy != ty || dx != tx) asc = false;
      if (dy != wh - 1 - ty || dx != ww - 1 - tx) desc = false;
    }
    P.dense = asc ? 1 : (desc ? 2 : 0);
    P.kh = wh; P.kw = ww;
  }
  P.taps = (const int32_t *)d_taps2;
  P.coef = d_coef;
  P.X = X;
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  const uint32_t box2[2] = {(uint32_t)P.bh, (uint32_t)P.bw}, box3[3] = {(uint32_t)P.bd, (uint32_t)P.bh, (uint32_t)P.bw};
  P.use_tma = (!c->no_tma && P.bw <= 256 && P.bh <= 256 && P.bd <= 256 &&
               make_tensor_map(&tmap, p.elem, p.esize, r, d_in, p.in_dims, r == 3 ? box3 : box2)) ? 1 : 0;
  // three stages when they fit, two for large boxes (deep windows, 8-byte elements)
  const size_t tables = 64 + (size_t)((P.ntaps + 3) & ~3) * 4 + (size_t)P.ntaps * sizeof(T) + 16;
  P.nstages = SW_S;
  if ((size_t)P.nstages * P.stage_bytes + tables > 200 * 1024) P.nstages = 2;  // (three stages leave room for L1)
  if ((size_t)P.nstages * P.stage_bytes + tables > SW_SMEM_MAX) return -1;
  auto kern = P.use_tma ? smallwin_kernel<T, true, CANON> : smallwin_kernel<T, false, CANON>;
  static bool configured[64][2] = {{false}};  // (per instantiation: T, CANON)
  if (!configured[c->device & 63][P.use_tma]) {
    MB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SW_SMEM_MAX));
    configured[c->device & 63][P.use_tma] = true;
  }
  // residency before ring depth: two stages when that lets another CTA onto the SM (rank-3 boxes carry the window's
  // depth: avg 3x3x3 Float 1024^3 runs two CTAs of two stages, not one of three)
  int occ = 0;
  MB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, (size_t)P.nstages * P.stage_bytes + tables));
  if (P.nstages == 3) {
    int occ2 = 0;
    MB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, 256, (size_t)2 * P.stage_bytes + tables));
    if (occ2 > occ) { occ = occ2; P.nstages = 2; }
  }
  if (const char *e = std::getenv("MASSIV_B200_SW_STAGES")) {  // measurement knob
    P.nstages = std::atoi(e) == 2 ? 2 : P.nstages;
    MB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, (size_t)P.nstages * P.stage_bytes + tables));
  }
  const size_t smem_bytes = (size_t)P.nstages * P.stage_bytes + tables;
  if (occ < 1) occ = 1;
  int grid = c->num_sms * occ;
  if (grid > P.ntiles) grid = P.ntiles;
  kern<<<grid, 256, smem_bytes, c->stream>>>(tmap, P, (const T *)d_in, (T *)d_out);
  c->launches++;
  c->last_kernel = "smallwin";
  c->post_fused = true;  // the post-maps ran as program steps
  return check_cuda(c, cudaGetLastError(), "smallwin_kernel launch");
}

// -1: not a rank-2 / rank-3, unit-stride, small-window plan
int launch_smallwin(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p = dp.p;
  if ((p.rank != 2 && p.rank != 3) || !p.unit_stride() || p.ntaps < 1) return -1;
  for (int i = 0; i < p.rank; ++i) {
    if (p.in_dims[i] < 1 || p.in_dims[i] >= (1ll << 30) || p.out_dims[i] >= (1ll << 30)) return -1;
    if (p.max_off[i] - p.min_off[i] + 1 > SW_MAX_WIN) return -1;
  }
  if (p.out_elems < c->tile_min_elems || p.out_elems >= (1ll << 40)) return -1;
  if (p.ntaps * (p.kind == MB_MAG2 ? 2 : 1) > SW_MAX_TAPS) return -1;
  if (p.post.size() + p.prog.size() + 8 > sizeof(SWProgram::op)) return -1;
  SWProgram X;
  std::vector<uint8_t> coef;
  if (!build_program(p, &X, &coef)) return -1;
  // the plan's device tables as they are, except for MAG2: its two kernels share ONE tap list in the plan, the two
  // terms here each want their own — a doubled copy is made on first use and lives with the cached plan
  const void *d_coef = dp.d_c1;
  const void *d_taps = dp.d_taps;
  if (p.kind == MB_MAG2) {
    const size_t tb = (size_t)p.ntaps * p.rank * sizeof(int32_t), cb = coef.size();
    if (!dp.d_sw) {
      std::vector<uint8_t> host(2 * tb + cb);
      std::memcpy(host.data(), p.tap_off.data(), tb);
      std::memcpy(host.data() + tb, p.tap_off.data(), tb);
      std::memcpy(host.data() + 2 * tb, coef.data(), cb);
      MB_CUDA(c, cudaMalloc(&dp.d_sw, host.size()));
      MB_CUDA(c, cudaMemcpy(dp.d_sw, host.data(), host.size(), cudaMemcpyHostToDevice));
    }
    d_taps = dp.d_sw;
    d_coef = (const char *)dp.d_sw + 2 * tb;
  }
  switch (p.elem) {
    case MB_U8: return launch_sw_t<uint8_t>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_I8: return launch_sw_t<int8_t>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_U16: return launch_sw_t<uint16_t>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_I16: return launch_sw_t<int16_t>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_U32: return launch_sw_t<uint32_t>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_I32: return launch_sw_t<int32_t>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_BOOL32: return launch_sw_t<int32_t, true>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_U64: return launch_sw_t<uint64_t>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_I64: return launch_sw_t<int64_t>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_F32: return launch_sw_t<float>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
    case MB_F64: return launch_sw_t<double>(c, dp, X, d_coef, d_taps, d_in, d_out, range);
  }
  return -1;
}

}  // namespace mb
