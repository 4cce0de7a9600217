// Tiled rank-2 stencil kernel — the hot loop of the windowed load for 2-D arrays
// (reference: loadWithIx2 / unrollAndJam, Array/Delayed/Windowed.hs:282-305, 476-512).
//
// Design (B200):
//   * persistent CTAs (grid = SMs x resident CTAs) walk the output tiles round-robin;
//   * the input box of a tile (tile + stencil halo) is staged in shared memory by ONE TMA
//     instruction (cp.async.bulk.tensor.2d) into a 4-deep ring guarded by mbarriers, so the loads of
//     the next three tiles are in flight while a tile is computed; out-of-range box cells are
//     zero-filled by the TMA unit, which IS the `Fill 0` border;
//   * tiles whose box crosses the array edge under any other border (and arrays TMA cannot
//     describe) are filled by a peeled path that resolves every cell through handleBorderIndex;
//     the compute loop itself is identical and branch-free for every tile;
//   * each thread produces V x R outputs (V = 16 B of elements) from a register window that slides
//     down the tile: every shared-memory row is read once per thread with 128-bit loads, rows are
//     reused across the R vertically adjacent outputs (the unroll-and-jam of the reference);
//   * outputs leave as 128-bit coalesced stores.
// Arithmetic follows the reference's tap order with unfused IEEE operations (mb_arith.cuh).
#pragma once
#include <type_traits>

#include "mb_arith.cuh"
#include "mb_internal.h"
#include "mb_tma.cuh"

namespace mb {

enum { OP_SUM = 0, OP_PRODUCT, OP_MAX, OP_MIN, OP_LIN, OP_MAG };

struct T2Params {
  int in_h, in_w, out_h, out_w;
  int shift_y, shift_x;  // source index of the window's first cell for output (0,0)
  int y0, y1;            // output rows produced
  int tiles_x, ntiles;
  int border, use_tma, zero_fill_ok, vec_store, avg;
  unsigned long long fill_bits;
  DivCount dc;
};

template <typename T, int N> struct Coef {
  T c1[N];
  T c2[N];
};

// XOFF: TMA needs the first byte of every box row 16-byte aligned in global memory, so the box starts
// at the source column rounded DOWN to a whole vector and the window begins XOFF (< V) elements into
// the staged row.  XOFF is a template parameter so that the register window is indexed statically.
template <typename T, int OP, bool REV, int KH, int KW, int XOFF, bool TMA = true> struct T2Cfg {
  static constexpr int V = 16 / (int)sizeof(T);
#ifndef MB_T2_WX64
#define MB_T2_WX64 2
#endif
#ifndef MB_T2_R64
#define MB_T2_R64 3
#endif
#ifndef MB_T2_R32
#define MB_T2_R32 4
#endif
  // Double kernels with windows up to three rows are DRAM-bound: two warps side by side make the box rows 1 KiB long
  // (longer contiguous DRAM bursts) and three rows per thread keep the tile at 128 x 12 cells.  Measured at 16384^2
  // (scripts/probe_tile2d.py; 1 warp x 4 rows -> 2 x 4 -> 2 x 3): avg3x3 335 -> 346 -> 348 Gcells/s, conv3x3 Reflect
  // 314 -> 341 -> 364, conv3x3 Fill 0 328 -> 346 -> 386; the 5x5 / 7x7 kernels (two rows per thread) and the Int64
  // folds lose 3-10 % and keep the one-warp rows.
  static constexpr bool WIDE = sizeof(T) == 8 && KH <= 3 && std::is_floating_point<T>::value;
  static constexpr int WX = WIDE ? MB_T2_WX64 : 1;
  static constexpr int TX = 32 * WX, TY = 8 / WX;
  static constexpr int TW = TX * V;
  // (without a tensor map the cp.async-staged wide tile does better with four rows: avg3x3 16383^2 286 against 274)
  static constexpr int R = (KH >= 5 && sizeof(T) == 8) ? 2 : (WIDE ? (TMA ? MB_T2_R64 : 4) : (sizeof(T) == 4 && KH <= 3 ? MB_T2_R32 : 4));
  static constexpr int TH = TY * R;
  static constexpr int HALO = ((XOFF + KW - 1 + V - 1) / V) * V; // halo rounded to whole vectors
  static constexpr int NV = 1 + HALO / V;
  static constexpr int BW = TW + HALO, BH = TH + KH - 1;
  static constexpr int STAGE = ((BW * BH * (int)sizeof(T) + 127) / 128) * 128;
  // ring depth: 8-byte elements are DRAM-bound and gain from residency (two stages: six CTAs per SM instead of three —
  // avg3x3 Double 16384^2: S = 2: 338 Gcells/s, 3: 311, 4: 308); the 4-byte kernels are issue-bound and want the deeper
  // prefetch (Sobel Float 32768^2: S = 4: 585, 3: 545, 2: 553)
#ifndef MB_T2_S
#define MB_T2_S (sizeof(T) == 8 ? 2 : 4)
#endif
  static constexpr int S = MB_T2_S;
  static constexpr int SMEM = S * STAGE + 64;
  // Arrays without a tensor map (row pitch not a multiple of 16 bytes): boxes go through REGISTERS — a warp per box row,
  // one element per lane and trip (coalesced whatever the row's alignment), the next tile's loads in flight while this
  // tile is computed, two stages.  LDG + STS cost 1.8 + 1 LSU cycles per warp and trip, LDGSTS (cp.async) 8
  // (B300_MICROARCH.md, LDG / LDGSTS tables), and rows that are not 16-byte aligned need element-sized copies.
  static constexpr int RS_ROWS = (BH + 7) / 8, RS_COLS = (BW + 31) / 32;
#ifndef MB_T2_REGSTAGE
#define MB_T2_REGSTAGE 1
#endif
  // 4-byte elements only: Sobel Float 32767^2 420 Gcells/s against 379 with cp.async; with 8-byte elements every row can
  // move in 16- or 8-byte cp.async copies and those win (avg3x3 Double 16383^2: 275 against 261)
  static constexpr bool REGSTAGE = MB_T2_REGSTAGE && sizeof(T) == 4 && RS_ROWS * RS_COLS <= 36;
  static constexpr int SMEM_RS = 2 * STAGE + 64;
};

template <typename T, int NVV> __device__ __forceinline__ void load_row(T (&dst)[NVV], const T *src) {
  constexpr int V = 16 / (int)sizeof(T);
#pragma unroll
  for (int i = 0; i < NVV / V; ++i) {
    const uint4 q = *reinterpret_cast<const uint4 *>(src + i * V);
    T tmp[V];
    memcpy(tmp, &q, 16);
#pragma unroll
    for (int v = 0; v < V; ++v) dst[i * V + v] = tmp[v];
  }
}

// Compile-time integer coefficients for well-known kernels (KC != void): a tap with coefficient 1 or -1
// is one add (x * 1 is x exactly), a zero coefficient is one exact fused multiply-add fma(x, 0, acc) —
// x*0 is exactly +-0 or NaN, so the fused and the unfused forms round identically — or nothing at all
// when the caller promised finite inputs (FIN = 1) or the launch checks its results (FIN = 2, below); everything else is
// the reference's multiply then add.
// FIN = 2, the checked optimistic form (no promise needed): zero taps are skipped, every stored result is inspected,
// and a non-finite one raises `flag`; the strict launch (FIN = 0) that follows on the stream returns at once while the
// flag is clear.  Sound because a non-finite input that sits on a zero tap of one output sits on a non-zero tap of a
// neighbouring output (the host checks the coefficient pattern and the geometry: known_optimistic_ok), so whenever the
// two forms could differ somewhere, some result is non-finite.
template <typename T, int FIN> __device__ __forceinline__ T tap_known(T x, T acc, int kv) {
  using A = Arith<T>;
  if (kv == 0) {
    if (FIN) return acc;
    if constexpr (sizeof(T) == 4) return __fmaf_rn(x, T(0), acc); else return __fma_rn(x, T(0), acc);
  }
  if (kv == 1) return A::add(x, acc);
  if (kv == -1) return A::add(-x, acc);
  return A::add(A::mul(x, T(kv)), acc);
}

// TMA: boxes come through the tensor map; false: the array has none (row pitch not a multiple of 16 bytes)
// and boxes inside the array are staged with cp.async — a separate instantiation, so that the TMA
// variant's register allocation is not disturbed (the 7x7 Float kernel sits at the 128-register limit).
// (register staging holds a box in registers: small kernels are kept at 80 registers so that three CTAs share an SM)
template <typename T, int OP, bool REV, int KH, int KW, int XOFF, typename KC = void, int FIN = 0, bool TMA = true>
#ifndef MB_T2_MINB8
#define MB_T2_MINB8 2  // (experiment: minimum CTAs per SM of the small 8-byte kernels)
#endif
__global__ void __launch_bounds__(256, (!TMA && T2Cfg<T, OP, REV, KH, KW, XOFF, TMA>::REGSTAGE && KH * KW <= 9) ? 3
                                       : (sizeof(T) == 8 && KH * KW <= 9) ? MB_T2_MINB8 : 2)
tile2d_kernel(const __grid_constant__ CUtensorMap tmap, const T2Params P, const Coef<T, KH * KW> K,
              const T *__restrict__ in, T *__restrict__ out, unsigned *flag) {
  // FIN = 0 with a flag: the strict redo of a checked optimistic launch — nothing to do unless that one met Inf / NaN
  if (FIN == 0 && flag != nullptr && *reinterpret_cast<const volatile unsigned *>(flag) == 0u) return;
  unsigned bad = 0;  // FIN = 2: running maximum of the results' exponent-and-up bits
  using C = T2Cfg<T, OP, REV, KH, KW, XOFF, TMA>;
  using A = Arith<T>;
  using Acc = typename A::Acc;
  constexpr int V = C::V, R = C::R, NVV = C::NV * V;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr bool RS = C::REGSTAGE && !TMA;  // register staging, two stages
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (RS ? 2 : C::S) * C::STAGE);
  const int tid = threadIdx.x;
  if (tid == 0) {
    if (TMA) tma_prefetch_desc(&tmap);
#pragma unroll
    for (int s = 0; s < C::S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int my_tiles = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_origin = [&](int i, int &oy0, int &ox0) {
    const int t = (int)blockIdx.x + i * (int)gridDim.x;
    oy0 = P.y0 + (t / P.tiles_x) * C::TH;
    ox0 = (t % P.tiles_x) * C::TW;
  };
  // may tile i be fetched by TMA?  yes when out-of-range cells are never used or zero is the border
  auto tma_ok = [&](int oy0, int ox0) -> bool {
    if (!TMA) return false;
    if (P.zero_fill_ok) return true;
    const int by0 = oy0 + P.shift_y, bx0 = ox0 + P.shift_x;
    const int th = min(C::TH, P.y1 - oy0), tw = min(C::TW, P.out_w - ox0);
    return by0 >= 0 && bx0 >= 0 && by0 + th + KH - 1 <= P.in_h && bx0 + tw + KW - 1 <= P.in_w;
  };
  auto issue = [&](int i) {  // thread 0 only
    int oy0, ox0;
    tile_origin(i, oy0, ox0);
    if (!tma_ok(oy0, ox0)) return;
    const int s = i % C::S;
    mbar_arrive_expect_tx(&full[s], (uint32_t)(C::BW * C::BH * sizeof(T)));
    tma_load_2d(smem + s * C::STAGE, &tmap, &full[s], ox0 + P.shift_x - XOFF, oy0 + P.shift_y);
  };
  // Arrays TMA cannot describe (row pitch not a multiple of 16 bytes): boxes that lie inside the array are
  // staged with cp.async by all threads, one commit group per tile, S - 1 tiles ahead like the TMA ring.
  auto box_inside = [&](int oy0, int ox0) -> bool {
    const int by0 = oy0 + P.shift_y, bx0 = ox0 + P.shift_x - XOFF;
    return by0 >= 0 && bx0 >= 0 && by0 + C::BH <= P.in_h && bx0 + C::BW <= P.in_w;
  };
  auto prefetch = [&](int i) {  // every thread
    if (TMA) {
      if (tid == 0 && i < my_tiles) issue(i);
      return;
    }
    if (C::REGSTAGE && !TMA) return;
    if (sizeof(T) >= 4 && i < my_tiles) {
      int oy0, ox0;
      tile_origin(i, oy0, ox0);
      if (box_inside(oy0, ox0)) {
        // a warp per box row, the widest copies the row's alignment allows (mb_tma.cuh)
        stage_box<T, C::BH, C::BW, 8>(reinterpret_cast<T *>(smem + (i % C::S) * C::STAGE), in, in, P.in_h, P.in_w, oy0 + P.shift_y,
                                      ox0 + P.shift_x - XOFF, tid);
      }
    }
    cp_async_commit();
  };
  // register staging: tile i's box is held in stg from one trip to the next
  T stg[RS ? C::RS_ROWS : 1][RS ? C::RS_COLS : 1];
  auto rs_load = [&](int i) {
    if (!RS || i >= my_tiles) return;
    int oy0, ox0;
    tile_origin(i, oy0, ox0);
    if (!box_inside(oy0, ox0)) return;
    const T *src = in + (size_t)(oy0 + P.shift_y + (tid >> 5)) * P.in_w + (ox0 + P.shift_x - XOFF) + (tid & 31);
#pragma unroll
    for (int k = 0; k < C::RS_ROWS; ++k) {
      if (k * 8 + 8 <= C::BH || (tid >> 5) + k * 8 < C::BH) {
#pragma unroll
        for (int j = 0; j < C::RS_COLS; ++j)
          if (j * 32 + 32 <= C::BW || (tid & 31) + j * 32 < C::BW) stg[k][j] = __ldg(src + (size_t)k * 8 * P.in_w + j * 32);
      }
    }
  };
  auto rs_store = [&](T *sm) {
    if (!RS) return;
    T *dst = sm + (tid >> 5) * C::BW + (tid & 31);
#pragma unroll
    for (int k = 0; k < C::RS_ROWS; ++k) {
      if (k * 8 + 8 <= C::BH || (tid >> 5) + k * 8 < C::BH) {
#pragma unroll
        for (int j = 0; j < C::RS_COLS; ++j)
          if (j * 32 + 32 <= C::BW || (tid & 31) + j * 32 < C::BW) dst[k * 8 * C::BW + j * 32] = stg[k][j];
      }
    }
  };
  if (RS) rs_load(0);
#pragma unroll 1
  for (int i = 0; i < C::S - 1; ++i) prefetch(i);
  T fillv;
  {
    unsigned long long b = P.fill_bits;
    memcpy(&fillv, &b, sizeof(T));
  }
  const int lx = (tid % C::TX) * V, ly = (tid / C::TX) * R;
  uint32_t phase = 0;

  for (int i = 0; i < my_tiles; ++i) {
    const int s = RS ? (i & 1) : i % C::S;
    int oy0, ox0;
    tile_origin(i, oy0, ox0);
    if (TMA && tid == 0 && i + C::S - 1 < my_tiles) fence_proxy_async();
    prefetch(i + C::S - 1);
    T *sm = reinterpret_cast<T *>(smem + s * C::STAGE);
    const bool by_tma = tma_ok(oy0, ox0);
    if (by_tma) {
      mbar_wait(&full[s], (phase >> s) & 1u);
      phase ^= 1u << s;
    } else if (RS && box_inside(oy0, ox0)) {
      rs_store(sm);
      __syncthreads();
      rs_load(i + 1);
    } else if (!TMA && sizeof(T) >= 4 && box_inside(oy0, ox0)) {
      cp_async_wait_group<C::S - 1>();  // this tile's group has landed (S - 1 younger ones may be in flight)
      __syncthreads();
    } else {
      if (!TMA) cp_async_wait_group<C::S - 1>();  // keeps the group count in step
      // peeled edge path: every box cell through the border rule (Core/Index.hs:236-253)
      const int by0 = oy0 + P.shift_y, bx0 = ox0 + P.shift_x - XOFF;
      for (int e = tid; e < C::BW * C::BH; e += 256) {
        const int r = e / C::BW, c = e - r * C::BW;
        int gy = by0 + r, gx = bx0 + c;
        T v = fillv;
        if (resolve_border32(P.border, P.in_h, gy) && resolve_border32(P.border, P.in_w, gx))
          v = in[(size_t)gy * P.in_w + gx];
        sm[e] = v;
      }
      __syncthreads();
      rs_load(i + 1);
    }

    // ---- compute: V x R outputs per thread from a sliding register window ----
    T win[KH][NVV];
    const T *base = sm + ly * C::BW + lx;
#pragma unroll
    for (int j = 0; j < KH - 1; ++j) load_row<T, NVV>(win[j], base + j * C::BW);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      load_row<T, NVV>(win[(r + KH - 1) % KH], base + (r + KH - 1) * C::BW);
      T res[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if constexpr (OP == OP_SUM) {  // foldlStencil: acc = acc + x, row-major (Stencil.hs:298-302)
          Acc acc = A::from_count(0);
#pragma unroll
          for (int ty = 0; ty < KH; ++ty)
#pragma unroll
            for (int tx = 0; tx < KW; ++tx) acc = A::add(acc, A::load(win[(r + ty) % KH][XOFF + v + tx]));
          T o = A::store(acc);
          if (P.avg) o = div_count(o, P.dc);
          res[v] = o;
        } else if constexpr (OP == OP_PRODUCT) {
          Acc acc = A::from_count(1);
#pragma unroll
          for (int ty = 0; ty < KH; ++ty)
#pragma unroll
            for (int tx = 0; tx < KW; ++tx) acc = A::mul(acc, A::load(win[(r + ty) % KH][XOFF + v + tx]));
          res[v] = A::store(acc);
        } else if constexpr (OP == OP_MAX || OP == OP_MIN) {
          T acc = OP == OP_MAX ? Bounds<T>::lo() : Bounds<T>::hi();
#pragma unroll
          for (int ty = 0; ty < KH; ++ty)
#pragma unroll
            for (int tx = 0; tx < KW; ++tx) {
              const T x = win[(r + ty) % KH][XOFF + v + tx];
              acc = OP == OP_MAX ? (x > acc ? x : acc) : (x < acc ? x : acc);
            }
          res[v] = acc;
        } else {  // OP_LIN / OP_MAG: acc = src * k + acc in kernel order (Convolution.hs:74-78, 115-119)
          Acc a = A::from_count(0), b = A::from_count(0);
#pragma unroll
          for (int ty = 0; ty < KH; ++ty)
#pragma unroll
            for (int tx = 0; tx < KW; ++tx) {
              const int wy = REV ? KH - 1 - ty : ty, wx = REV ? KW - 1 - tx : tx;
              const Acc x = A::load(win[(r + wy) % KH][XOFF + v + wx]);
              if constexpr (std::is_void<KC>::value) a = A::add(A::mul(x, A::load(K.c1[ty * KW + tx])), a);
              else a = tap_known<T, FIN>(x, a, KC::k1(ty * KW + tx));
            }
          if constexpr (OP == OP_MAG) {
#pragma unroll
            for (int ty = 0; ty < KH; ++ty)
#pragma unroll
              for (int tx = 0; tx < KW; ++tx) {
                const int wy = REV ? KH - 1 - ty : ty, wx = REV ? KW - 1 - tx : tx;
                const Acc x = A::load(win[(r + wy) % KH][XOFF + v + wx]);
                if constexpr (std::is_void<KC>::value) b = A::add(A::mul(x, A::load(K.c2[ty * KW + tx])), b);
                else b = tap_known<T, FIN>(x, b, KC::k2(ty * KW + tx));
              }
            a = A::sqrt(A::add(A::mul(a, a), A::mul(b, b)));
          }
          res[v] = A::store(a);
        }
      }
      const int gy = oy0 + ly + r, gx = ox0 + lx;
      if (gy < P.y1 && gx < P.out_w) {
        if (FIN == 2) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            unsigned hi;
            if constexpr (sizeof(T) == 8) hi = (unsigned)__double2hiint((double)res[v]); else hi = __float_as_uint((float)res[v]);
            if (gx + v < P.out_w) bad = max(bad, hi & 0x7fffffffu);
          }
        }
        T *dst = out + (size_t)gy * P.out_w + gx;
        if (P.vec_store && gx + V <= P.out_w) {
          uint4 q;
          memcpy(&q, res, 16);
          *reinterpret_cast<uint4 *>(dst) = q;
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (gx + v < P.out_w) dst[v] = res[v];
        }
      }
    }
    if (RS) continue;  // two stages taken in turn, a barrier after every fill: no thread can still read the stage filled next
    if (!by_tma) fence_proxy_async();
    __syncthreads();  // everyone is done with stage s before it is refilled
  }
  if (FIN == 2 && bad >= (sizeof(T) == 8 ? 0x7ff00000u : 0x7f800000u)) atomicOr(flag, 1u);
}

// ---- host side ------------------------------------------------------------------------------------

// one launch with the tile geometry of the TMA (tmap != nullptr) or the non-TMA instantiation
template <typename T, int OP, bool REV, int KH, int KW, int XOFF, typename KC, int FIN, bool TMA>
static int launch_cfg(Ctx *c, const DevPlan &dp, const CUtensorMap &tmap, const void *d_in, void *d_out, const OuterRange *range,
                      unsigned *flag) {
  using C = T2Cfg<T, OP, REV, KH, KW, XOFF, TMA>;
  const Plan &p = dp.p;
  T2Params P;
  std::memset(&P, 0, sizeof(P));
  P.in_h = (int)p.in_dims[0]; P.in_w = (int)p.in_dims[1];
  P.out_h = (int)p.out_dims[0]; P.out_w = (int)p.out_dims[1];
  P.shift_y = (int)(p.shift[0] + p.min_off[0]); P.shift_x = (int)(p.shift[1] + p.min_off[1]);
  P.y0 = range ? (int)range->o0 : 0;
  P.y1 = range ? (int)range->o1 : P.out_h;
  if (P.y1 <= P.y0) return MB_OK;
  P.tiles_x = (P.out_w + C::TW - 1) / C::TW;
  const int tiles_y = (P.y1 - P.y0 + C::TH - 1) / C::TH;
  P.ntiles = P.tiles_x * tiles_y;
  P.border = p.border;
  std::memcpy(&P.fill_bits, p.fill, 8);
  bool fill_zero = p.border == MB_FILL;
  for (int i = 0; i < p.esize; ++i) if (p.fill[i] != 0) fill_zero = false;
  P.zero_fill_ok = fill_zero;
  P.vec_store = ((uintptr_t)d_out % 16 == 0) && ((size_t)P.out_w * sizeof(T)) % 16 == 0;
  P.avg = p.kind == MB_AVG;
  P.dc.n_d = (double)p.avg_n; P.dc.n_f = (float)p.avg_n;
  P.dc.fast = p.avg_n >= 1 && p.avg_n <= 1024;
  if (P.dc.fast) { P.dc.y_d = 1.0 / P.dc.n_d; P.dc.y_f = 1.0f / P.dc.n_f; }
  Coef<T, KH * KW> K;
  std::memset(&K, 0, sizeof(K));
  if (!p.c1.empty()) std::memcpy(K.c1, p.c1.data(), sizeof(T) * KH * KW);
  if (!p.c2.empty()) std::memcpy(K.c2, p.c2.data(), sizeof(T) * KH * KW);
  P.use_tma = TMA ? 1 : 0;
  auto kern = tile2d_kernel<T, OP, REV, KH, KW, XOFF, KC, FIN, TMA>;
  static int occ_dev[64] = {0};  // per device: the smem opt-in is a per-device function attribute
  int &occ = occ_dev[c->device & 63];
  const int smem_bytes = (!TMA && C::REGSTAGE) ? C::SMEM_RS : C::SMEM;
  if (occ == 0) {
    MB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    MB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem_bytes));
    if (occ < 1) occ = 1;
  }
  int grid = c->num_sms * occ;
  if (grid > P.ntiles) grid = P.ntiles;
  kern<<<grid, 256, smem_bytes, c->stream>>>(tmap, P, K, (const T *)d_in, (T *)d_out, flag);
  c->launches++;
  c->last_kernel = std::is_void<KC>::value ? "tile2d" : "tile2d_known";
  return check_cuda(c, cudaGetLastError(), "tile2d_kernel launch");
}

template <typename T, int OP, bool REV, int KH, int KW, int XOFF, typename KC = void, int FIN = 0>
static int launch_x(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range, unsigned *flag = nullptr) {
  using CT = T2Cfg<T, OP, REV, KH, KW, XOFF, true>;
  const Plan &p = dp.p;
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  const uint32_t box[2] = {(uint32_t)CT::BH, (uint32_t)CT::BW};
  if (!c->no_tma && make_tensor_map(&tmap, p.elem, p.esize, 2, d_in, p.in_dims, box))
    return launch_cfg<T, OP, REV, KH, KW, XOFF, KC, FIN, true>(c, dp, tmap, d_in, d_out, range, flag);
  return launch_cfg<T, OP, REV, KH, KW, XOFF, KC, FIN, false>(c, dp, tmap, d_in, d_out, range, flag);
}

// well-known 3x3 kernels with compile-time coefficients (k_tile2d_known.cu); -1 when none matches
int launch_tile2d_known(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range);

}  // namespace mb
