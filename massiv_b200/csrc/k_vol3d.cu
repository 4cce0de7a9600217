// Rank-3 3x3x3 convolution / correlation stencils — the iterated heat / Laplacian step
// (reference: loadWithIxN, Array/Delayed/Windowed.hs:432-474: one job per outer page, each page a
// rank-2 windowed load; FromKernel arithmetic, Array/Stencil/Convolution.hs:64-80, 107-121).
//
// Design (B200): 2.5-D streaming.  A persistent CTA owns a (TY x TX) column of the grid and marches
// along the outermost dimension.  Every source plane tile (with its one-cell halo) is brought into
// shared memory exactly once by a 3-D TMA box load into an S-deep mbarrier ring that runs several
// planes ahead of the arithmetic; TMA's out-of-range zero fill is the `Fill 0` border in all three
// dimensions.  Planes z-1 and z live in registers (the window slides along z), so one plane is read
// from HBM once per pass and the kernel moves the algorithmic 2 x sizeof(e) bytes per cell.
//   STAR7   : the seven face/centre coefficients are the only non-zero ones and the caller promised
//             finite inputs (MB_FLAG_ASSUME_FINITE): 7 multiply + 7 add per cell, in kernel order.
//   DENSE27 : any 3x3x3 kernel, all 27 taps (zero coefficients included — strict NaN/Inf semantics).
// Tiles whose box leaves the array under a non-zero border are filled through the border rule by a
// peeled path; the arithmetic loop is the same for every tile.
#include "mb_arith.cuh"
#include "mb_internal.h"
#include "mb_tma.cuh"

namespace mb {

enum { V3_STAR7 = 0, V3_DENSE27 = 1 };

struct V3Params {
  int in_d, in_h, in_w, out_d, out_h, out_w;
  int in_pitch, out_pitch;  // row pitches in elements (in_w / out_w unless that array is padded: Ctx::in_pitch, out_pitch)
  int sz, sy, sx;  // source index of the window centre for output (0,0,0)
  int z0, z1;      // output planes produced
  int tiles_x, tiles_y, zc, nitems;
  int border, use_tma, zero_fill_ok, vec_store;
  unsigned long long fill_bits;
};

template <typename T> struct Coef27 {
  T k[27];
};

template <typename T, int MODE> struct V3Cfg {
  static constexpr int V = 16 / (int)sizeof(T);
  static constexpr int TX = 32 * V;
  static constexpr int R = MODE == V3_STAR7 ? 4 : 2;
  static constexpr int TY = 8 * R;
  static constexpr int BW = TX + 2 * V;  // left/right halo padded to a whole vector: rows stay 16 B aligned
  static constexpr int BH = TY + 2;
  static constexpr int STAGE = ((BW * BH * (int)sizeof(T) + 127) / 128) * 128;
#ifndef MB_V3_S
#define MB_V3_S 5
#endif
#ifndef MB_V3_MINB
#define MB_V3_MINB 2
#endif
  static constexpr int S = MODE == V3_STAR7 ? MB_V3_S : 6;
  static constexpr int SMEM = S * STAGE + 64;
};

template <typename T> __device__ __forceinline__ void lds_vec(T *dst, const T *src) {
  constexpr int V = 16 / (int)sizeof(T);
  const uint4 q = *reinterpret_cast<const uint4 *>(src);
  T tmp[V];
  memcpy(tmp, &q, 16);
#pragma unroll
  for (int v = 0; v < V; ++v) dst[v] = tmp[v];
}

// DETECT (STAR7 only): the kernel skips the 20 zero taps WITHOUT the caller's promise and reports through
// `flag` whether an Inf or NaN was involved — the only inputs for which skipping a zero tap changes the
// result (0 * Inf = NaN).  DETECT 2 inspects every source value that passes through registers.  DETECT 1
// (centre coefficient non-zero) inspects the RESULTS instead, half the work: a non-finite source cell
// makes the result at its own position non-finite through the centre tap (Inf * k, NaN * k; nothing
// added to it brings it back).  The dense kernel that follows on the stream takes `flag` as its condition
// and returns at once when it is clear, so the strict result costs a star pass for finite data.
template <typename T, int MODE, bool REV, int DETECT, bool TMA>
__global__ void __launch_bounds__(256, MODE == V3_STAR7 ? MB_V3_MINB : 1)
vol3d_kernel(const __grid_constant__ CUtensorMap tmap, const V3Params P, const Coef27<T> K,
             const T *__restrict__ in, T *__restrict__ out, unsigned *flag) {
  if (MODE == V3_DENSE27 && flag != nullptr && *reinterpret_cast<volatile unsigned *>(flag) == 0u) return;
  unsigned bad = 0;
  using C = V3Cfg<T, MODE>;
  using A = Arith<T>;
  constexpr int V = C::V, R = C::R, S = C::S;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + S * C::STAGE);
  const int tid = threadIdx.x;
  if (tid == 0) {
    if (TMA) tma_prefetch_desc(&tmap);
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int my_items = (P.nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto item_geom = [&](int it, int &oz0, int &nz, int &oy0, int &ox0) {
    const int t = (int)blockIdx.x + it * (int)gridDim.x;
    const int tx = t % P.tiles_x, rest = t / P.tiles_x;
    const int ty = rest % P.tiles_y, zi = rest / P.tiles_y;
    oz0 = P.z0 + zi * P.zc;
    nz = min(P.zc, P.z1 - oz0);
    oy0 = ty * C::TY;
    ox0 = tx * C::TX;
  };
  // may the plane box be fetched by TMA (out-of-range cells unused, or zero IS the border)?
  auto tma_ok = [&](int zsrc, int oy0, int ox0) -> bool {
    if (!TMA) return false;
    if (P.zero_fill_ok) return true;
    const int y0 = oy0 + P.sy - 1, x0 = ox0 + P.sx - 1;
    const int th = min(C::TY, P.out_h - oy0), tw = min(C::TX, P.out_w - ox0);
    return zsrc >= 0 && zsrc < P.in_d && y0 >= 0 && x0 >= 0 && y0 + th + 2 <= P.in_h && x0 + tw + 2 <= P.in_w;
  };

  // without TMA: may the plane box be staged with cp.async (cells outside the array written as zeros)?
  auto cp_ok = [&](int zsrc, int oy0, int ox0) -> bool {
    if (TMA) return false;
    if (P.zero_fill_ok) return true;
    const int y0 = oy0 + P.sy - 1, x0 = ox0 + P.sx - V;
    return zsrc >= 0 && zsrc < P.in_d && y0 >= 0 && x0 >= 0 && y0 + C::BH <= P.in_h && x0 + C::BW <= P.in_w;
  };

  // ---- producer state (thread 0; every thread without TMA): loads are numbered globally, stage = number % S ----
  int p_it = 0, p_pl = 0, issued = 0, released = 0;
  int total_loads = 0;
  for (int it = 0; it < my_items; ++it) {
    int a, nz, b, c;
    item_geom(it, a, nz, b, c);
    total_loads += nz + 2;
  }
  auto issue_next = [&]() {  // thread 0 only
    int oz0, nz, oy0, ox0;
    item_geom(p_it, oz0, nz, oy0, ox0);
    const int zsrc = oz0 + P.sz - 1 + p_pl;
    const int s = issued % S;
    if (tma_ok(zsrc, oy0, ox0)) {
      mbar_arrive_expect_tx(&full[s], (uint32_t)(C::BW * C::BH * sizeof(T)));
      tma_load_3d(smem + s * C::STAGE, &tmap, &full[s], ox0 + P.sx - V, oy0 + P.sy - 1, zsrc);
    }
    if (!TMA) {  // one commit group per load, empty when the box goes through the border rule at acquire time
      if (cp_ok(zsrc, oy0, ox0)) {
        // a warp per box row, the widest copies the row's alignment allows; zeros outside the array (mb_tma.cuh)
        const T *pl = (zsrc >= 0 && zsrc < P.in_d) ? in + (size_t)zsrc * P.in_h * P.in_w : nullptr;  // (dense arrays only)
        stage_box<T, C::BH, C::BW, 8>(reinterpret_cast<T *>(smem + s * C::STAGE), pl, in, P.in_h, P.in_w, oy0 + P.sy - 1, ox0 + P.sx - V, tid);
      }
      cp_async_commit();
    }
    ++issued;
    if (++p_pl == nz + 2) { p_pl = 0; ++p_it; }
  };
  if (tid == 0 || !TMA) {
    while (issued < total_loads && issued < released + S) issue_next();
  }

  T fillv;
  {
    unsigned long long b = P.fill_bits;
    memcpy(&fillv, &b, sizeof(T));
  }
  uint32_t phase = 0;
  // consumer side of one load: wait for the TMA, or fill the stage through the border rule
  auto acquire = [&](int L, int zsrc, int oy0, int ox0) {
    const int s = L % S;
    if (tma_ok(zsrc, oy0, ox0)) {
      mbar_wait(&full[s], (phase >> s) & 1u);
      phase ^= 1u << s;
    } else if (cp_ok(zsrc, oy0, ox0)) {
      switch (issued - 1 - L) {  // younger groups that may still be in flight
        case 0: cp_async_wait_group<0>(); break;
        case 1: cp_async_wait_group<1>(); break;
        case 2: cp_async_wait_group<2>(); break;
        case 3: cp_async_wait_group<3>(); break;
        case 4: cp_async_wait_group<4>(); break;
        default: cp_async_wait_group<5>(); break;
      }
      __syncthreads();
    } else {
      T *sm = reinterpret_cast<T *>(smem + s * C::STAGE);
      const int y0 = oy0 + P.sy - 1, x0 = ox0 + P.sx - V;
      int gz = zsrc;
      const bool zin = resolve_border32(P.border, P.in_d, gz);
      for (int e = tid; e < C::BW * C::BH; e += 256) {
        const int r = e / C::BW, c = e - r * C::BW;
        int gy = y0 + r, gx = x0 + c;
        T v = fillv;
        if (zin && resolve_border32(P.border, P.in_h, gy) && resolve_border32(P.border, P.in_w, gx))
          v = in[((size_t)gz * P.in_h + gy) * P.in_pitch + gx];
        sm[e] = v;
      }
      fence_proxy_async();
      __syncthreads();
    }
  };

  const int lx = (tid % 32) * V, ly = (tid / 32) * R;
  int Lc = 0;
  for (int it = 0; it < my_items; ++it) {
    int oz0, nz, oy0, ox0;
    item_geom(it, oz0, nz, oy0, ox0);
    const int zs0 = oz0 + P.sz - 1;  // source plane of load Lc
    acquire(Lc, zs0, oy0, ox0);
    acquire(Lc + 1, zs0 + 1, oy0, ox0);

    if constexpr (MODE == V3_STAR7) {
      T zm[R][V], zc[R][V];
      {
        const T *s0 = reinterpret_cast<const T *>(smem + (Lc % S) * C::STAGE) + (ly + 1) * C::BW + V + lx;
        const T *s1 = reinterpret_cast<const T *>(smem + ((Lc + 1) % S) * C::STAGE) + (ly + 1) * C::BW + V + lx;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          lds_vec<T>(zm[r], s0 + r * C::BW);
          lds_vec<T>(zc[r], s1 + r * C::BW);
          if (DETECT) {  // the plane below the item's first result plane is nobody's result in this launch
#pragma unroll
            for (int v = 0; v < V; ++v) { track_magnitude(bad, zm[r][v]); if (DETECT == 2) track_magnitude(bad, zc[r][v]); }
          }
        }
      }
#pragma unroll 1
      for (int j = 0; j < nz; ++j) {
        acquire(Lc + j + 2, zs0 + j + 2, oy0, ox0);
        const T *sm = reinterpret_cast<const T *>(smem + ((Lc + j + 1) % S) * C::STAGE) + (ly + 1) * C::BW + V + lx;
        const T *sn = reinterpret_cast<const T *>(smem + ((Lc + j + 2) % S) * C::STAGE) + (ly + 1) * C::BW + V + lx;
        T zp[R][V], ytop[V], ybot[V], xl[R], xr[R];
        lds_vec<T>(ytop, sm - C::BW);
        lds_vec<T>(ybot, sm + R * C::BW);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          lds_vec<T>(zp[r], sn + r * C::BW);
          xl[r] = sm[r * C::BW - 1];
          xr[r] = sm[r * C::BW + V];
          if (DETECT == 2 || (DETECT == 1 && j == nz - 1)) {  // ... and neither is the plane above its last one
#pragma unroll
            for (int v = 0; v < V; ++v) track_magnitude(bad, zp[r][v]);
          }
        }
        const int gz = oz0 + j;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          T res[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const T up = r > 0 ? zc[r - 1][v] : ytop[v];        // y - 1
            const T dn = r + 1 < R ? zc[r + 1][v] : ybot[v];    // y + 1
            const T lf = v > 0 ? zc[r][v - 1] : xl[r];          // x - 1
            const T rt = v + 1 < V ? zc[r][v + 1] : xr[r];      // x + 1
            // kernel index order 4,10,12,13,14,16,22 (row-major over the 3x3x3 kernel); source offset of
            // kernel index (p,q,r) is (1-p,1-q,1-r) for a convolution, (p-1,q-1,r-1) for a correlation
            T acc = A::from_count(0);
            acc = A::add(A::mul(REV ? zp[r][v] : zm[r][v], K.k[4]), acc);
            acc = A::add(A::mul(REV ? dn : up, K.k[10]), acc);
            acc = A::add(A::mul(REV ? rt : lf, K.k[12]), acc);
            acc = A::add(A::mul(zc[r][v], K.k[13]), acc);
            acc = A::add(A::mul(REV ? lf : rt, K.k[14]), acc);
            acc = A::add(A::mul(REV ? up : dn, K.k[16]), acc);
            acc = A::add(A::mul(REV ? zm[r][v] : zp[r][v], K.k[22]), acc);
            res[v] = acc;
            if (DETECT == 1) track_magnitude(bad, acc);
          }
          const int gy = oy0 + ly + r, gx = ox0 + lx;
          if (gy < P.out_h && gx < P.out_w) {
            T *dst = out + ((size_t)gz * P.out_h + gy) * P.out_pitch + gx;
            if (P.vec_store && gx + V <= P.out_w) {
              uint4 q;
              memcpy(&q, res, 16);
#ifdef MB_V3_STCS
              __stcs(reinterpret_cast<uint4 *>(dst), q);
#else
              *reinterpret_cast<uint4 *>(dst) = q;
#endif
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v)
                if (gx + v < P.out_w) dst[v] = res[v];
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int v = 0; v < V; ++v) { zm[r][v] = zc[r][v]; zc[r][v] = zp[r][v]; }
        __syncthreads();  // plane Lc+j+1 (the middle one) is no longer read by anyone
        if (tid == 0 || !TMA) {
          released = (j == nz - 1) ? Lc + nz + 2 : Lc + j + 2;
          while (issued < total_loads && issued < released + S) issue_next();
        }
      }
    } else {
      // DENSE27: full (R+2) x (V+2) windows of planes z-1, z, z+1 in registers
      T w[3][R + 2][V + 2];
      auto load_win = [&](T (&dst)[R + 2][V + 2], int L) {
        const T *sp = reinterpret_cast<const T *>(smem + (L % S) * C::STAGE) + ly * C::BW + V + lx;
#pragma unroll
        for (int r = 0; r < R + 2; ++r) {
          T mid[V];
          lds_vec<T>(mid, sp + r * C::BW);
          dst[r][0] = sp[r * C::BW - 1];
#pragma unroll
          for (int v = 0; v < V; ++v) dst[r][1 + v] = mid[v];
          dst[r][V + 1] = sp[r * C::BW + V];
        }
      };
      load_win(w[0], Lc);
      load_win(w[1], Lc + 1);
      __syncthreads();
      if (tid == 0 || !TMA) {
        released = Lc + 2;
        while (issued < total_loads && issued < released + S) issue_next();
      }
#pragma unroll 1
      for (int j = 0; j < nz; ++j) {
        acquire(Lc + j + 2, zs0 + j + 2, oy0, ox0);
        load_win(w[2], Lc + j + 2);
        const int gz = oz0 + j;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          T res[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            T acc = A::from_count(0);
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
              for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                  const int dz = REV ? 1 - p : p - 1, dy = REV ? 1 - q : q - 1, dx = REV ? 1 - t : t - 1;
                  acc = A::add(A::mul(w[1 + dz][r + 1 + dy][v + 1 + dx], K.k[p * 9 + q * 3 + t]), acc);
                }
            res[v] = acc;
          }
          const int gy = oy0 + ly + r, gx = ox0 + lx;
          if (gy < P.out_h && gx < P.out_w) {
            T *dst = out + ((size_t)gz * P.out_h + gy) * P.out_pitch + gx;
            if (P.vec_store && gx + V <= P.out_w) {
              uint4 q;
              memcpy(&q, res, 16);
#ifdef MB_V3_STCS
              __stcs(reinterpret_cast<uint4 *>(dst), q);
#else
              *reinterpret_cast<uint4 *>(dst) = q;
#endif
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v)
                if (gx + v < P.out_w) dst[v] = res[v];
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R + 2; ++r)
#pragma unroll
          for (int v = 0; v < V + 2; ++v) { w[0][r][v] = w[1][r][v]; w[1][r][v] = w[2][r][v]; }
        __syncthreads();  // the newest plane now lives in registers: its stage is free
        if (tid == 0 || !TMA) {
          released = Lc + j + 3;
          while (issued < total_loads && issued < released + S) issue_next();
        }
      }
    }
    Lc += nz + 2;
  }
  if (DETECT && magnitude_nonfinite<T>(bad)) atomicOr(flag, 1u);
}

template <typename T, int MODE, bool REV, int DETECT = 0>
static int launch_v3(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range, unsigned *flag = nullptr) {
  using C = V3Cfg<T, MODE>;
  const Plan &p = dp.p;
  V3Params P;
  std::memset(&P, 0, sizeof(P));
  P.in_d = (int)p.in_dims[0]; P.in_h = (int)p.in_dims[1]; P.in_w = (int)p.in_dims[2];
  P.out_d = (int)p.out_dims[0]; P.out_h = (int)p.out_dims[1]; P.out_w = (int)p.out_dims[2];
  P.sz = (int)p.shift[0]; P.sy = (int)p.shift[1]; P.sx = (int)p.shift[2];
  P.z0 = range ? (int)range->o0 : 0;
  P.z1 = range ? (int)range->o1 : P.out_d;
  if (P.z1 <= P.z0) return MB_OK;
  P.tiles_x = (P.out_w + C::TX - 1) / C::TX;
  P.tiles_y = (P.out_h + C::TY - 1) / C::TY;
  P.border = p.border;
  std::memcpy(&P.fill_bits, p.fill, 8);
  bool fill_zero = p.border == MB_FILL;
  for (int i = 0; i < p.esize; ++i) if (p.fill[i] != 0) fill_zero = false;
  P.zero_fill_ok = fill_zero;
  P.in_pitch = c->in_pitch ? (int)c->in_pitch : P.in_w;
  P.out_pitch = c->out_pitch ? (int)c->out_pitch : P.out_w;
  P.vec_store = ((uintptr_t)d_out % 16 == 0) && ((size_t)P.out_pitch * sizeof(T)) % 16 == 0;
  Coef27<T> K;
  std::memcpy(K.k, p.c1.data(), sizeof(T) * 27);
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  const uint32_t box[3] = {1u, (uint32_t)C::BH, (uint32_t)C::BW};
  P.use_tma = (!c->no_tma && make_tensor_map(&tmap, p.elem, p.esize, 3, d_in, p.in_dims, box, c->in_pitch)) ? 1 : 0;
  if (c->in_pitch && !P.use_tma) return fail(c, MB_ERR_UNSUPPORTED, "internal: padded arrays need a tensor map");
  auto kern_tma = vol3d_kernel<T, MODE, REV, DETECT, true>;
  auto kern_cp = vol3d_kernel<T, MODE, REV, DETECT, false>;
  auto kern = P.use_tma ? kern_tma : kern_cp;
  static int occ_dev[64][2] = {{0}};
  int &occ = occ_dev[c->device & 63][P.use_tma];
  if (occ == 0) {
    MB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    MB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, C::SMEM));
    if (occ < 1) occ = 1;
  }
  // z-chunks balance the persistent CTAs
  const int resident = c->num_sms * occ;
  const int tiles = P.tiles_x * P.tiles_y;
  const int nzt = P.z1 - P.z0;
  int chunks = 1;
  {
    // the chunk count with the smallest estimated makespan: whole waves of resident CTAs, each item
    // paying its halo planes and the pipeline fill
    const int max_chunks = nzt / 16 > 1 ? (nzt / 16 < 256 ? nzt / 16 : 256) : 1;
    long best = -1;
    for (int cnum = 1; cnum <= max_chunks; ++cnum) {
      const int zc = (nzt + cnum - 1) / cnum;
      const int real = (nzt + zc - 1) / zc;
      const long waves = ((long)tiles * real + resident - 1) / resident;
      const long cost = waves * (zc + 2 + 3);
      if (best < 0 || cost < best) { best = cost; chunks = real; P.zc = zc; }
    }
  }
  P.nitems = tiles * chunks;
  int grid = resident < P.nitems ? resident : P.nitems;
  kern<<<grid, 256, C::SMEM, c->stream>>>(tmap, P, K, (const T *)d_in, (T *)d_out, flag);
  c->launches++;
  c->last_kernel = MODE == V3_STAR7 ? "vol3d_star7" : "vol3d_dense27";
  return check_cuda(c, cudaGetLastError(), "vol3d_kernel launch");
}

template <typename T>
static bool is_star(const Plan &p) {
  const T *k = reinterpret_cast<const T *>(p.c1.data());
  for (int t = 0; t < 27; ++t) {
    const bool star = t == 4 || t == 10 || t == 12 || t == 13 || t == 14 || t == 16 || t == 22;
    if (!star && k[t] != T(0)) return false;
  }
  return true;
}

template <typename T>
static int launch_v3_t(Ctx *c, const DevPlan &dp, const void *i, void *o, const OuterRange *r) {
  const Plan &p = dp.p;
  // TMA wants every box row to start on a 16-byte boundary: the box begins one vector left of the
  // tile, which is aligned exactly when the window centre of output column 0 is (mapStencil: 0)
  if (p.shift[2] % (16 / (int)sizeof(T)) != 0) return -1;
  const bool rev = p.kind == MB_CONV;
  const bool starshape = is_star<T>(p);
  if (starshape && (p.flags & MB_FLAG_ASSUME_FINITE))
    return rev ? launch_v3<T, V3_STAR7, true>(c, dp, i, o, r) : launch_v3<T, V3_STAR7, false>(c, dp, i, o, r);
  // No promise: run the star kernel optimistically and let it report non-finite inputs; the strict
  // 27-tap kernel follows on the stream and only does work when it has to.  Every source cell must pass
  // through some thread's registers for the report to be complete: in-plane the cell grid has to be the
  // array's own (along dim 0 every plane that is read is also loaded as a centre plane).
  const bool inplane_same = p.shift[1] == 0 && p.shift[2] == 0 && p.out_dims[1] == p.in_dims[1] && p.out_dims[2] == p.in_dims[2];
  bool fill_finite = true;  // a Fill element of Inf / NaN reaches the zero taps at the border
  if (p.border == MB_FILL) {
    T fv;
    std::memcpy(&fv, p.fill, sizeof(T));
    fill_finite = fv - fv == T(0);
  }
  if (starshape && inplane_same && fill_finite && !c->no_optimistic && ensure_flag(c) == MB_OK) {
    MB_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(unsigned), c->stream));
    const T centre = reinterpret_cast<const T *>(p.c1.data())[13];
    int st;
    if (centre != T(0))  // (a NaN centre makes every result NaN: the check fires, the redo agrees)
      st = rev ? launch_v3<T, V3_STAR7, true, 1>(c, dp, i, o, r, c->d_flag) : launch_v3<T, V3_STAR7, false, 1>(c, dp, i, o, r, c->d_flag);
    else
      st = rev ? launch_v3<T, V3_STAR7, true, 2>(c, dp, i, o, r, c->d_flag) : launch_v3<T, V3_STAR7, false, 2>(c, dp, i, o, r, c->d_flag);
    if (st) return st;
    st = rev ? launch_v3<T, V3_DENSE27, true>(c, dp, i, o, r, c->d_flag) : launch_v3<T, V3_DENSE27, false>(c, dp, i, o, r, c->d_flag);
    c->aux_launches++;
    c->last_kernel = "vol3d_star7";
    return st;
  }
  return rev ? launch_v3<T, V3_DENSE27, true>(c, dp, i, o, r) : launch_v3<T, V3_DENSE27, false>(c, dp, i, o, r);
}

// the strict kernel, conditional on *cond (used by the optimistic two-step path)
template <typename T>
static int dense_cond_t(Ctx *c, const DevPlan &dp, const void *i, void *o, const unsigned *cond) {
  if (dp.p.shift[2] % (16 / (int)sizeof(T)) != 0) return -1;
  return dp.p.kind == MB_CONV ? launch_v3<T, V3_DENSE27, true>(c, dp, i, o, nullptr, const_cast<unsigned *>(cond))
                              : launch_v3<T, V3_DENSE27, false>(c, dp, i, o, nullptr, const_cast<unsigned *>(cond));
}
int launch_vol3d_dense_cond(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const unsigned *cond) {
  if (cond) c->aux_launches++;
  if (dp.p.elem == MB_F32) return dense_cond_t<float>(c, dp, d_in, d_out, cond);
  if (dp.p.elem == MB_F64) return dense_cond_t<double>(c, dp, d_in, d_out, cond);
  return -1;
}

__global__ void cond_copy_kernel(uint4 *dst, const uint4 *src, size_t n16, unsigned char *dst_tail, const unsigned char *src_tail,
                                 int tail, const unsigned *cond) {
  if (*reinterpret_cast<const volatile unsigned *>(cond) == 0u) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}
int launch_cond_copy(Ctx *c, void *dst, const void *src, size_t bytes, const unsigned *cond) {
  const size_t n16 = bytes / 16;
  cond_copy_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>((uint4 *)dst, (const uint4 *)src, n16, (unsigned char *)dst + n16 * 16,
                                                         (const unsigned char *)src + n16 * 16, (int)(bytes - n16 * 16), cond);
  c->launches++;
  c->aux_launches++;
  return check_cuda(c, cudaGetLastError(), "cond_copy_kernel launch");
}

__global__ void cond_fill_kernel(unsigned char *dst, size_t n_elems, int esize, unsigned long long bits, const unsigned *cond) {
  if (*reinterpret_cast<const volatile unsigned *>(cond) == 0u) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_elems; i += (size_t)gridDim.x * blockDim.x)
    for (int b = 0; b < esize; ++b) dst[i * esize + b] = (unsigned char)(bits >> (8 * b));
}
int launch_cond_fill(Ctx *c, void *dst, size_t n_elems, int esize, unsigned long long bits, const unsigned *cond) {
  cond_fill_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>((unsigned char *)dst, n_elems, esize, bits, cond);
  c->launches++;
  c->aux_launches++;
  return check_cuda(c, cudaGetLastError(), "cond_fill_kernel launch");
}

int ensure_flag(Ctx *c) {
  if (c->d_flag) return MB_OK;
  if (cudaMalloc(&c->d_flag, 256) != cudaSuccess) { cudaGetLastError(); return MB_ERR_CUDA; }
  return MB_OK;
}

int launch_vol3d(Ctx *c, const DevPlan &dp, const void *d_in, void *d_out, const OuterRange *range) {
  const Plan &p = dp.p;
  if (p.rank != 3 || !p.unit_stride() || (p.kind != MB_CONV && p.kind != MB_CORR)) return -1;
  for (int i = 0; i < 3; ++i) {
    if (p.size[i] != 3 || p.in_dims[i] < 1 || p.in_dims[i] >= (1ll << 30) || p.out_dims[i] >= (1ll << 30)) return -1;
  }
  if (p.out_elems < c->tile_min_elems) return -1;
  switch (p.elem) {
    case MB_F32: return launch_v3_t<float>(c, dp, d_in, d_out, range);
    case MB_F64: return launch_v3_t<double>(c, dp, d_in, d_out, range);
  }
  return -1;
}

}  // namespace mb
