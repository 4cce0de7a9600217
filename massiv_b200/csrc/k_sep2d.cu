// Fused separable application:  compute (mapStencil b (K x 1)) . compute (mapStencil b (1 x K))
// in ONE pass over HBM.  The reference has no fused form — the idiom is two `compute`s with the
// intermediate array rounded to the element type and the border rule applied to it again (SURVEY.md
// Appendix A4); this kernel reproduces exactly that: the row pass writes its result, rounded to T,
// into a shared-memory tile (including the K-1 halo rows the column pass needs, each resolved through
// the border rule on the INTERMEDIATE's row index), the column pass reads it back.  One read and one
// write of every element: 2 x sizeof(e) bytes per cell instead of 4 x.
//
// Same machinery as k_tile2d.cu: persistent CTAs, one TMA box load per tile into an mbarrier ring,
// peeled border fill for edge tiles, 128-bit shared loads, 128-bit coalesced stores.
#include "mb_arith.cuh"
#include "mb_internal.h"
#include "mb_tma.cuh"

namespace mb {

struct S2Params {
  int h, w;              // array extent (input == intermediate == output)
  int shift_y, shift_x;  // source row / column of the first window cell for output (0,0)
  int y0, y1;            // output rows produced
  int tiles_x, ntiles;
  int border, use_tma, zero_fill_ok, vec_store;
  unsigned long long fill_bits;
};

template <typename T, int K> struct SepCoef {
  T row[K];  // 1 x K kernel (first pass)
  T col[K];  // K x 1 kernel (second pass)
};

template <typename T, int K, int XOFF> struct S2Cfg {
  static constexpr int V = 16 / (int)sizeof(T);
  static constexpr int TW = 32 * V, TY = 8, R = 4, TH = TY * R;
  static constexpr int HALO = ((XOFF + K - 1 + V - 1) / V) * V;
  static constexpr int NV = 1 + HALO / V;
  static constexpr int BW = TW + HALO, BH = TH + K - 1;
  static constexpr int STAGE = ((BW * BH * (int)sizeof(T) + 127) / 128) * 128;
  static constexpr int TBUF = ((TW * BH * (int)sizeof(T) + 127) / 128) * 128;
// ring depth: two stages make three CTAs fit an SM (62 KB each) instead of two — 24 warps hide the barrier and
// shared-memory waits between the two passes far better than a third tile in flight does:
// Gauss 7x7 Float 32768^2: S = 2: 593-597 Gcells/s, S = 3: 490-498, S = 4: 499 (same box)
#ifndef MB_S2_S
#define MB_S2_S 2
#endif
  static constexpr int S = MB_S2_S;
  static constexpr int SMEM = S * STAGE + TBUF + 64;
  // arrays without a tensor map, 4-byte elements: boxes go through registers (k_tile2d.cuh); the box is only read by
  // the row pass, so ONE stage does
  static constexpr int RS_ROWS = (BH + 7) / 8, RS_COLS = (BW + 31) / 32;
#ifndef MB_S2_REGSTAGE
#define MB_S2_REGSTAGE 1
#endif
  static constexpr bool REGSTAGE = MB_S2_REGSTAGE && sizeof(T) == 4 && RS_ROWS * RS_COLS <= 36;
  static constexpr int SMEM_RS = STAGE + TBUF + 64;
};

template <typename T, bool REV, int K, int XOFF, bool TMA>
__global__ void __launch_bounds__(256, (!TMA && S2Cfg<T, K, XOFF>::REGSTAGE) ? 3 : 2)
sep2d_kernel(const __grid_constant__ CUtensorMap tmap, const S2Params P, const SepCoef<T, K> C2,
             const T *__restrict__ in, T *__restrict__ out) {
  using C = S2Cfg<T, K, XOFF>;
  using A = Arith<T>;
  constexpr int V = C::V, R = C::R, NVV = C::NV * V;
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr bool RS = C::REGSTAGE && !TMA;
  constexpr int NS = RS ? 1 : C::S;
  T *tbuf = reinterpret_cast<T *>(smem + NS * C::STAGE);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + NS * C::STAGE + C::TBUF);
  const int tid = threadIdx.x;
  if (tid == 0) {
    if (TMA) tma_prefetch_desc(&tmap);
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int my_tiles = (P.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_origin = [&](int i, int &oy0, int &ox0) {
    const int t = (int)blockIdx.x + i * (int)gridDim.x;
    oy0 = P.y0 + (t / P.tiles_x) * C::TH;
    ox0 = (t % P.tiles_x) * C::TW;
  };
  auto tma_ok = [&](int oy0, int ox0) -> bool {
    if (!TMA) return false;
    if (P.zero_fill_ok) return true;
    const int by0 = oy0 + P.shift_y, bx0 = ox0 + P.shift_x;
    const int th = min(C::TH, P.y1 - oy0), tw = min(C::TW, P.w - ox0);
    return by0 >= 0 && bx0 >= 0 && by0 + th + K - 1 <= P.h && bx0 + tw + K - 1 <= P.w;
  };
  auto issue = [&](int i) {
    int oy0, ox0;
    tile_origin(i, oy0, ox0);
    if (!tma_ok(oy0, ox0)) return;
    const int s = i % C::S;
    mbar_arrive_expect_tx(&full[s], (uint32_t)(C::BW * C::BH * sizeof(T)));
    tma_load_2d(smem + s * C::STAGE, &tmap, &full[s], ox0 + P.shift_x - XOFF, oy0 + P.shift_y);
  };
  // arrays TMA cannot describe: boxes inside the array are staged with cp.async by all threads, one commit
  // group per tile, S - 1 tiles ahead like the TMA ring (see k_tile2d.cuh)
  auto box_inside = [&](int oy0, int ox0) -> bool {
    const int by0 = oy0 + P.shift_y, bx0 = ox0 + P.shift_x - XOFF;
    return by0 >= 0 && bx0 >= 0 && by0 + C::BH <= P.h && bx0 + C::BW <= P.w;
  };
  auto prefetch = [&](int i) {
    if (TMA) {
      if (tid == 0 && i < my_tiles) issue(i);
      return;
    }
    if (RS) return;
    if (i < my_tiles) {
      int oy0, ox0;
      tile_origin(i, oy0, ox0);
      if (box_inside(oy0, ox0)) {
        // a warp per box row, the widest copies the row's alignment allows (mb_tma.cuh)
        stage_box<T, C::BH, C::BW, 8>(reinterpret_cast<T *>(smem + (i % C::S) * C::STAGE), in, in, P.h, P.w, oy0 + P.shift_y,
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
    const T *src = in + (size_t)(oy0 + P.shift_y + (tid >> 5)) * P.w + (ox0 + P.shift_x - XOFF) + (tid & 31);
#pragma unroll
    for (int k = 0; k < C::RS_ROWS; ++k) {
      if (k * 8 + 8 <= C::BH || (tid >> 5) + k * 8 < C::BH) {
#pragma unroll
        for (int j = 0; j < C::RS_COLS; ++j)
          if (j * 32 + 32 <= C::BW || (tid & 31) + j * 32 < C::BW) stg[k][j] = __ldg(src + (size_t)k * 8 * P.w + j * 32);
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
  const int lane = tid % 32, warp = tid / 32;
  const int lx = lane * V, ly = warp * R;
  uint32_t phase = 0;

  for (int i = 0; i < my_tiles; ++i) {
    const int s = RS ? 0 : i % C::S;
    int oy0, ox0;
    tile_origin(i, oy0, ox0);
    if (TMA && tid == 0 && i + C::S - 1 < my_tiles) fence_proxy_async();
    prefetch(i + C::S - 1);
    T *sm = reinterpret_cast<T *>(smem + s * C::STAGE);
    const bool by_tma = tma_ok(oy0, ox0);
    const int by0 = oy0 + P.shift_y, bx0 = ox0 + P.shift_x - XOFF;
    if (by_tma) {
      mbar_wait(&full[s], (phase >> s) & 1u);
      phase ^= 1u << s;
    } else if (RS && box_inside(oy0, ox0)) {
      rs_store(sm);
      __syncthreads();
      rs_load(i + 1);
    } else if (!TMA && box_inside(oy0, ox0)) {
      cp_async_wait_group<C::S - 1>();
      __syncthreads();
    } else {
      if (!TMA) cp_async_wait_group<C::S - 1>();  // keeps the group count in step
      for (int e = tid; e < C::BW * C::BH; e += 256) {
        const int r = e / C::BW, c = e - r * C::BW;
        int gy = by0 + r, gx = bx0 + c;
        T v = fillv;
        if (resolve_border32(P.border, P.h, gy) && resolve_border32(P.border, P.w, gx)) v = in[(size_t)gy * P.w + gx];
        sm[e] = v;
      }
      __syncthreads();
      rs_load(i + 1);
    }

    // ---- pass 1: 1 x K along the rows of the box, result rounded to T into tbuf ----
    // (fully unrolled: the trip count is fixed, the row's shared-memory address becomes an immediate offset; the
    // Fill test only exists for tiles that touch the array's first or last rows)
    const bool fill_rows = P.border == MB_FILL && (by0 < 0 || by0 + C::BH > P.h);
#pragma unroll
    for (int j = 0; j < (C::BH + C::TY - 1) / C::TY; ++j) {
      const int rr = warp + j * C::TY;
      if (rr >= C::BH) break;
      T res[V];
      const int gy = by0 + rr;
      if (fill_rows && (gy < 0 || gy >= P.h)) {
        // the intermediate array is padded with the Fill element itself, not with a filtered fill row
#pragma unroll
        for (int v = 0; v < V; ++v) res[v] = fillv;
      } else {
        T x[NVV];
        const T *src = sm + rr * C::BW + lx;
#pragma unroll
        for (int q = 0; q < NVV / V; ++q) {
          const uint4 u = *reinterpret_cast<const uint4 *>(src + q * V);
          T tmp[V];
          memcpy(tmp, &u, 16);
#pragma unroll
          for (int v = 0; v < V; ++v) x[q * V + v] = tmp[v];
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          T acc = A::from_count(0);
#pragma unroll
          for (int tx = 0; tx < K; ++tx) {
            const int wx = REV ? K - 1 - tx : tx;
            acc = A::add(A::mul(x[XOFF + v + wx], C2.row[tx]), acc);
          }
          res[v] = acc;
        }
      }
      uint4 u;
      memcpy(&u, res, 16);
      *reinterpret_cast<uint4 *>(tbuf + rr * C::TW + lx) = u;
    }
    __syncthreads();

    // ---- pass 2: K x 1 down the columns of tbuf, sliding register window ----
    T win[K][V];
    const T *tb = tbuf + ly * C::TW + lx;
    auto ldrow = [&](T (&dst)[V], const T *p) {
      const uint4 u = *reinterpret_cast<const uint4 *>(p);
      T tmp[V];
      memcpy(tmp, &u, 16);
#pragma unroll
      for (int v = 0; v < V; ++v) dst[v] = tmp[v];
    };
#pragma unroll
    for (int j = 0; j < K - 1; ++j) ldrow(win[j], tb + j * C::TW);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ldrow(win[(r + K - 1) % K], tb + (r + K - 1) * C::TW);
      T res[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        T acc = A::from_count(0);
#pragma unroll
        for (int ty = 0; ty < K; ++ty) {
          const int wy = REV ? K - 1 - ty : ty;
          acc = A::add(A::mul(win[(r + wy) % K][v], C2.col[ty]), acc);
        }
        res[v] = acc;
      }
      const int gy = oy0 + ly + r, gx = ox0 + lx;
      if (gy < P.y1 && gx < P.w) {
        T *dst = out + (size_t)gy * P.w + gx;
        if (P.vec_store && gx + V <= P.w) {
          uint4 u;
          memcpy(&u, res, 16);
          *reinterpret_cast<uint4 *>(dst) = u;
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (gx + v < P.w) dst[v] = res[v];
        }
      }
    }
    // (register staging: the stage is free after the barrier between the passes, tbuf after the one that follows
    // the next fill)
    if (RS) continue;
    if (!by_tma) fence_proxy_async();
    __syncthreads();  // stage s and tbuf are free again
  }
}

template <typename T, bool REV, int K, int XOFF>
static int launch_s2(Ctx *c, const Plan &p1, const Plan &p2, const void *d_in, void *d_out, const OuterRange *range) {
  using C = S2Cfg<T, K, XOFF>;
  S2Params P;
  std::memset(&P, 0, sizeof(P));
  P.h = (int)p1.in_dims[0]; P.w = (int)p1.in_dims[1];
  P.shift_y = (int)(p2.shift[0] + p2.min_off[0]);
  P.shift_x = (int)(p1.shift[1] + p1.min_off[1]);
  P.y0 = range ? (int)range->o0 : 0;
  P.y1 = range ? (int)range->o1 : P.h;
  if (P.y1 <= P.y0) return MB_OK;
  P.tiles_x = (P.w + C::TW - 1) / C::TW;
  P.ntiles = P.tiles_x * ((P.y1 - P.y0 + C::TH - 1) / C::TH);
  P.border = p1.border;
  std::memcpy(&P.fill_bits, p1.fill, 8);
  bool fill_zero = p1.border == MB_FILL;
  for (int i = 0; i < p1.esize; ++i) if (p1.fill[i] != 0) fill_zero = false;
  SepCoef<T, K> CO;
  std::memcpy(CO.row, p1.c1.data(), sizeof(T) * K);
  std::memcpy(CO.col, p2.c1.data(), sizeof(T) * K);
  // TMA zero fill stands in for `Fill 0` only if filtering zeros gives zero: finite row coefficients
  for (int i = 0; i < K; ++i) if (!(CO.row[i] - CO.row[i] == T(0))) fill_zero = false;
  P.zero_fill_ok = fill_zero;
  P.vec_store = ((uintptr_t)d_out % 16 == 0) && ((size_t)P.w * sizeof(T)) % 16 == 0;
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  const uint32_t box[2] = {(uint32_t)C::BH, (uint32_t)C::BW};
  P.use_tma = (!c->no_tma && make_tensor_map(&tmap, p1.elem, p1.esize, 2, d_in, p1.in_dims, box)) ? 1 : 0;
  auto kern_tma = sep2d_kernel<T, REV, K, XOFF, true>;
  auto kern_cp = sep2d_kernel<T, REV, K, XOFF, false>;
  auto kern = P.use_tma ? kern_tma : kern_cp;
  static int occ_dev[64][2] = {{0}};
  int &occ = occ_dev[c->device & 63][P.use_tma];
  const int smem_bytes = (!P.use_tma && C::REGSTAGE) ? C::SMEM_RS : C::SMEM;
  if (occ == 0) {
    MB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    MB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem_bytes));
    if (occ < 1) occ = 1;
  }
  int grid = c->num_sms * occ;
  if (grid > P.ntiles) grid = P.ntiles;
  kern<<<grid, 256, smem_bytes, c->stream>>>(tmap, P, CO, (const T *)d_in, (T *)d_out);
  c->launches++;
  c->last_kernel = "sep2d";
  return check_cuda(c, cudaGetLastError(), "sep2d_kernel launch");
}

template <typename T, bool REV, int K>
static int launch_s2_x(Ctx *c, const Plan &p1, const Plan &p2, const void *i, void *o, const OuterRange *r) {
  constexpr int V = 16 / (int)sizeof(T);
  const int sx = (int)(p1.shift[1] + p1.min_off[1]);
  const int xoff = ((sx % V) + V) % V;
  constexpr int NAT = (V - ((K / 2) % V)) % V;
  if (xoff == 0) return launch_s2<T, REV, K, 0>(c, p1, p2, i, o, r);
  if constexpr (NAT != 0) {
    if (xoff == NAT) return launch_s2<T, REV, K, NAT>(c, p1, p2, i, o, r);
  }
  return -1;
}

template <typename T>
static int launch_s2_t(Ctx *c, const Plan &p1, const Plan &p2, const void *i, void *o, const OuterRange *r) {
  const bool rev = p1.kind == MB_CONV;
  const int k = (int)p1.size[1];
#define MB_S2(KK)                                                                  \
  if (k == KK) return rev ? launch_s2_x<T, true, KK>(c, p1, p2, i, o, r) : launch_s2_x<T, false, KK>(c, p1, p2, i, o, r);
  MB_S2(3) MB_S2(5) MB_S2(7)
#undef MB_S2
  return -1;
}

int run_sep2_fused(Ctx *c, const DevPlan &first, const DevPlan &second, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p1 = first.p, &p2 = second.p;
  if (p1.rank != 2 || p2.rank != 2 || !p1.unit_stride() || !p2.unit_stride()) return -1;
  if ((p1.kind != MB_CONV && p1.kind != MB_CORR) || p2.kind != p1.kind) return -1;
  if (p1.size[0] != 1 || p2.size[1] != 1 || p1.size[1] != p2.size[0]) return -1;
  if (!p1.same_shape() || !p2.same_shape()) return -1;
  if (p1.in_dims[0] < 1 || p1.in_dims[1] < 1 || p1.in_dims[0] >= (1ll << 30) || p1.in_dims[1] >= (1ll << 30)) return -1;
  if (p1.out_elems < c->tile_min_elems) return -1;
  if (p1.shift[0] + p1.min_off[0] != 0 || p2.shift[1] + p2.min_off[1] != 0) return -1;
  switch (p1.elem) {
    case MB_F32: return launch_s2_t<float>(c, p1, p2, d_in, d_out, range);
    case MB_F64: return launch_s2_t<double>(c, p1, p2, d_in, d_out, range);
  }
  return -1;
}

}  // namespace mb
