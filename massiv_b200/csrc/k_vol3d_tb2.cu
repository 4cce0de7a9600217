// Temporal blocking for the iterated 3-D 7-point stencil: TWO applications of the stencil per pass
// over HBM (iterateUntil's step applied twice, Array/Manifest/Internal.hs:331-397), so every element
// is read once and written once per two steps — half the traffic of k_vol3d.cu.
//
// A CTA streams a (TYB x TXB) column along the outermost dimension.  For each new source plane it
//   stage 1: computes the INTERMEDIATE array (one application) on the whole column tile and parks it in
//            shared memory (cells outside the global array are the Fill element 0, not stencil values:
//            the border rule applies to the intermediate array exactly as it does in the reference),
//   stage 2: computes the RESULT (second application) one plane behind, on the tile minus a one-cell
//            rim (a vector wide in x), from the parked intermediate planes.
// Tiles overlap by that rim (overlapped tiling): the rim's stage-1 work and source loads are done
// twice (1.14x arithmetic, 1.28x reads per two steps) in exchange for no inter-CTA communication.
// Per-cell arithmetic is exactly that of two separate passes — same seven multiply-adds in kernel
// order per application — so the result is bit-identical to iterating the single-step kernel.
// Only for `Fill 0` (TMA out-of-range zero fill is the border at both levels), star-shaped finite
// kernels (MB_FLAG_ASSUME_FINITE); everything else iterates the single-step kernels.
#include "mb_arith.cuh"
#include "mb_internal.h"
#include "mb_tma.cuh"

#ifndef MB_TB2_R
#define MB_TB2_R 4  // rows per thread (tile height MB_TB2_NW * R)
#endif
#ifndef MB_TB2_NW
#define MB_TB2_NW 12  // warps per CTA: a warp owns R rows of the tile.  12 x 4 = 48-row tiles, one CTA of 384 threads per SM
                      // with 160 registers each: measured best of 8 / 10 / 12 / 14 / 16 warps x 2..6 rows
#endif
#ifndef MB_TB2_MINB
#define MB_TB2_MINB (MB_TB2_NW <= 8 ? 2 : 1)
#endif
#ifndef MB_TB2_S
#define MB_TB2_S 4
#endif

namespace mb {

struct TB2Params {
  int in_d, in_h, in_w;  // source buffer (for sharded slabs: own planes + 2 ghost planes each side)
  int out_d;             // result planes
  int sz2;               // source-frame plane of result plane 0 (0 single GPU, 2 for a slab)
  int zoff, dglob;       // source-frame plane p is global plane p + zoff of a dglob-plane array
  int z0, z1;            // result planes produced
  int tiles_x, tiles_y, zc, nitems;
  int vec_store;
  int pitch;             // row pitch in elements of the OUTPUT array (in_w unless it is padded: Ctx::out_pitch; a padded
                         // input is described by the tensor map)
  int use_tma;           // 0: the array cannot be described to TMA; boxes are staged with cp.async (zero fill by hand)
};

template <typename T> struct Coef27b {
  T k[27];
};

template <typename T, int R_> struct TB2Cfg {
  static constexpr int V = 16 / (int)sizeof(T);
  static constexpr int R = R_;
  static constexpr int NT = 32 * MB_TB2_NW;            // threads per CTA
#ifndef MB_TB2_WX
#define MB_TB2_WX 1  // warps side by side in x (experiment: longer box rows)
#endif
  static constexpr int WX = sizeof(T) == 8 ? MB_TB2_WX : 1;  // (a Float box twice as wide would exceed TMA's 256-element box)
  static constexpr int TXB = 32 * WX * V, TYB = (MB_TB2_NW / WX) * R;  // stage-1 tile
  static constexpr int OX = TXB - 2 * V, OY = TYB - 2;  // result tile
  static constexpr int BW = TXB + 2 * V, BH = TYB + 2;  // staged box (one-cell halo, vector-padded in x)
  static constexpr int STAGE = ((BW * BH * (int)sizeof(T) + 127) / 128) * 128;
  static constexpr int S = MB_TB2_S;
  static constexpr int SMEM = (S + 2) * STAGE + 128;    // S source stages + 2 intermediate planes + barriers + the producer's books
  // Arrays without a tensor map, 8-byte elements: every box row comes through the bulk-copy engine from the 16-byte
  // boundary at or one element before its first cell, so a row may sit ONE ELEMENT to the right in its (one vector
  // wider) shared-memory row; the readers add that element to the row's address.  For 8-byte elements a shifted row
  // costs the same shared-memory wavefronts (two LDS.64 instead of one LDS.128).
#ifndef MB_TB2_BULK
#define MB_TB2_BULK 1
#endif
  static constexpr bool BULK = MB_TB2_BULK && sizeof(T) == 8;
  static constexpr int SP_NT = BULK ? BW + V : BW;
  static constexpr int STAGE_NT = ((SP_NT * BH * (int)sizeof(T) + 127) / 128) * 128;
  static constexpr int SMEM_NT = S * STAGE_NT + 2 * STAGE + 128;
};

template <typename T> __device__ __forceinline__ void lds_v(T *dst, const T *src) {
  constexpr int V = 16 / (int)sizeof(T);
  const uint4 q = *reinterpret_cast<const uint4 *>(src);
  T tmp[V];
  memcpy(tmp, &q, 16);
#pragma unroll
  for (int v = 0; v < V; ++v) dst[v] = tmp[v];
}

// The in-plane neighbours one application needs beyond the thread's own R x V cells of the middle plane: the row
// above / below (a vector each) and the cell left / right of every row.
template <typename T, int R, int V> struct Nbrs {
  T ytop[V], ybot[V], xl[R], xr[R];
};

// experiment switches (each measured on the GPU before it became a default)
#ifndef MB_TB2_BOOKS_SMEM
#define MB_TB2_BOOKS_SMEM 0  // 1: the TMA producer's bookkeeping lives in shared memory instead of registers
#endif
#ifndef MB_TB2_INSIDE
#define MB_TB2_INSIDE 1      // 1: tiles that lie inside the array skip the masking of the intermediate values (uniform branch)
#endif
#ifndef MB_TB2_TRACKALL
#define MB_TB2_TRACKALL 1    // 1: the finite check inspects every result (rim included), stores are predicated, no per-row branches
#endif
#ifndef MB_TB2_L2HINT
#define MB_TB2_L2HINT 0  // 1 / 2: source boxes are fetched with an L2 evict_last / evict_first hint (the rims are read again by the neighbour tiles)
#endif
#ifndef MB_TB2_STCS
#define MB_TB2_STCS 0    // 1: results leave with streaming stores (st.global.cs: first to be evicted from L2)
#endif
#ifndef MB_TB2_HOIST
#define MB_TB2_HOIST 0  // 1: the in-plane neighbour loads of stage 1 are issued before the wait for the new source plane
#endif
#ifndef MB_TB2_SHFL
#define MB_TB2_SHFL 0  // 1: x neighbours come from the adjacent lanes' registers (warp shuffle) instead of shared memory
#endif

// `mid` points at the thread's first cell of the middle plane in shared memory (row pitch BW); pc are the same
// cells in registers.  Lane stride is one 16-byte vector, so the scalar xl / xr loads are two-way bank
// conflicted; with MB_TB2_SHFL they travel by shuffle and only lanes 0 / 31 touch shared memory.
template <typename T, int R, int V, int BW>
__device__ __forceinline__ void load_nbrs(const T *mid, const T (&pc)[R][V], Nbrs<T, R, V> &nb) {
  lds_v<T>(nb.ytop, mid - BW);
  lds_v<T>(nb.ybot, mid + R * BW);
#if MB_TB2_SHFL
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    T l = __shfl_up_sync(0xffffffffu, pc[r][V - 1], 1);
    T rr = __shfl_down_sync(0xffffffffu, pc[r][0], 1);
    if (lane == 0) l = mid[r * BW - 1];
    if (lane == 31) rr = mid[r * BW + V];
    nb.xl[r] = l;
    nb.xr[r] = rr;
  }
#else
#pragma unroll
  for (int r = 0; r < R; ++r) {
    nb.xl[r] = mid[r * BW - 1];
    nb.xr[r] = mid[r * BW + V];
  }
#endif
}

// one application on the thread's R x V cells: planes below / at / above in registers (pm, pc, pp).
// Kernel index order 4,10,12,13,14,16,22; see k_vol3d.cu.
template <typename T, bool REV, int R, int V>
__device__ __forceinline__ void star7(const Nbrs<T, R, V> &nb, const T (&pm)[R][V], const T (&pc)[R][V], const T (&pp)[R][V],
                                      const Coef27b<T> &K, T (&res)[R][V]) {
  using A = Arith<T>;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const T up = r > 0 ? pc[r - 1][v] : nb.ytop[v];
      const T dn = r + 1 < R ? pc[r + 1][v] : nb.ybot[v];
      const T lf = v > 0 ? pc[r][v - 1] : nb.xl[r];
      const T rt = v + 1 < V ? pc[r][v + 1] : nb.xr[r];
      T acc = A::from_count(0);
      acc = A::add(A::mul(REV ? pp[r][v] : pm[r][v], K.k[4]), acc);
      acc = A::add(A::mul(REV ? dn : up, K.k[10]), acc);
      acc = A::add(A::mul(REV ? rt : lf, K.k[12]), acc);
      acc = A::add(A::mul(pc[r][v], K.k[13]), acc);
      acc = A::add(A::mul(REV ? lf : rt, K.k[14]), acc);
      acc = A::add(A::mul(REV ? up : dn, K.k[16]), acc);
      acc = A::add(A::mul(REV ? pm[r][v] : pp[r][v], K.k[22]), acc);
      res[r][v] = acc;
    }
}

// Inf / NaN detection on the integer pipe without masking the sign: a positive non-finite value has the largest
// high word among the positives (signed maximum), a negative one the largest among all as unsigned.
#ifndef MB_TB2_TRACK2
#define MB_TB2_TRACK2 1
#endif
struct Finite2 {
  int smax = 0;
  unsigned umax = 0;
};
template <typename T, int N> __device__ __forceinline__ void track_row(Finite2 &f, const T (&x)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    int hi;
    if constexpr (sizeof(T) == 8) hi = __double2hiint(x[i]); else hi = __float_as_int(x[i]);
#if MB_TB2_TRACK2
    f.smax = max(f.smax, hi);
    f.umax = max(f.umax, (unsigned)hi);
#else
    f.umax = max(f.umax, (unsigned)hi & 0x7fffffffu);
#endif
  }
}
template <typename T> __device__ __forceinline__ bool nonfinite_seen(const Finite2 &f) {
#if MB_TB2_TRACK2
  return sizeof(T) == 8 ? (f.smax >= 0x7ff00000 || f.umax >= 0xfff00000u) : (f.smax >= 0x7f800000 || f.umax >= 0xff800000u);
#else
  return f.umax >= (sizeof(T) == 8 ? 0x7ff00000u : 0x7f800000u);
#endif
}

// DETECT: no finite-input promise.  The kernel reports through `flag` whether an Inf or NaN was involved
// — in the source or in the intermediate array (a finite input can overflow); the caller then redoes the
// two steps with the strict 27-tap kernel (conditional launches that return at once when the flag is
// clear).  DETECT 2 inspects every source and intermediate value.  DETECT 1 (centre coefficient non-zero)
// inspects the RESULTS only: a non-finite value at a cell makes the next application's value at the same
// cell non-finite through the centre tap, so it always surfaces in the result tile that owns the cell.
template <typename T, bool REV, int R_, int DETECT, bool TMA>
__global__ void __launch_bounds__(32 * MB_TB2_NW, MB_TB2_MINB)
vol3d_tb2_kernel(const __grid_constant__ CUtensorMap tmap, const TB2Params P, const Coef27b<T> K, const T *__restrict__ in,
                 T *__restrict__ out, unsigned *flag) {
  Finite2 bad;
  using C = TB2Cfg<T, R_>;
  constexpr int V = C::V, R = C::R, S = C::S, BW = C::BW;
  constexpr bool BULK = !TMA && C::BULK;
  constexpr int SP = TMA ? C::BW : C::SP_NT;         // row pitch of a staged source plane
  constexpr int SSTAGE = TMA ? C::STAGE : C::STAGE_NT;  // bytes per source stage
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *const ubase = smem + S * SSTAGE;  // two intermediate planes (row pitch BW), STAGE bytes apart
  uint64_t *full = reinterpret_cast<uint64_t *>(ubase + 2 * C::STAGE);
  uint64_t *ufree = full + S;  // "every warp has read intermediate buffer b": it may be overwritten
  const int tid = threadIdx.x;
  if (tid == 0) {
    if (TMA) tma_prefetch_desc(&tmap);
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_init(&ufree[0], MB_TB2_NW);
    mbar_init(&ufree[1], MB_TB2_NW);
    fence_mbar_init();
  }
  // The rim of the two intermediate buffers (box cells outside the stage-1 tile) is never written by a step;
  // the result cells that read it are themselves rim cells and never stored, but they ARE inspected by the
  // finite check, so the rim holds zeros rather than whatever the last kernel left in shared memory.
  for (int e = tid; e < 2 * C::STAGE / (int)sizeof(T); e += C::NT) reinterpret_cast<T *>(ubase)[e] = T(0);
  __syncthreads();

  const int my_items = (P.nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto item_geom = [&](int it, int &oz0, int &nz, int &oy0, int &ox0) {
    const int t = (int)blockIdx.x + it * (int)gridDim.x;
    const int tx = t % P.tiles_x, rest = t / P.tiles_x;
    const int ty = rest % P.tiles_y, zi = rest / P.tiles_y;
    oz0 = P.z0 + zi * P.zc;
    nz = min(P.zc, P.z1 - oz0);
    oy0 = ty * C::OY;
    ox0 = tx * C::OX;
  };
  // producer (thread 0): source planes numbered over this CTA's items; an item of nz result planes
  // consumes nz + 4 source planes (two above, two below).  The next load's coordinates are kept
  // incrementally so that the steady state costs a handful of instructions.
  // With TMA only thread 0 produces: its books live in shared memory so that they do not occupy registers in
  // every thread of a kernel that sits at the 128-register limit.  (Without TMA every thread copies and keeps
  // them in registers.)
  struct Books { int it, left, x, y, z, issued, total, code; };
  Books *const pb = reinterpret_cast<Books *>(ufree + 2);
  constexpr bool BOOKS_SMEM = TMA && MB_TB2_BOOKS_SMEM;
  Books rb = {0, 0, 0, 0, 0, 0, 0, -1};
  if (tid == 0 || !TMA) {
    for (int it = 0; it < my_items; ++it) {
      int a, nz, b, c;
      item_geom(it, a, nz, b, c);
      rb.total += nz + 4;
    }
    if (BOOKS_SMEM) *pb = rb;
  }
  // May the box of plane z at (y, x) come through the bulk-copy engine?  It must lie inside the array, and not end with
  // the array's last row (a shifted row reads one element past its end).  The in-plane part of the answer and the
  // parity of the box's first element are taken once per item (xy_code: -1 no, else bit 0 = the parity of plane 0's
  // box, bit 1 = the box ends with the plane's last row, bit 2 = it starts with the plane's first element);
  // row r of plane z's box is shifted when (parity + z * (H * W & 1) + r * (W & 1)) is odd.
  const unsigned base_el = (unsigned)(reinterpret_cast<uintptr_t>(in) / sizeof(T));
  const int hw1 = (P.in_h & P.in_w) & 1, sb = P.in_w & 1;
  auto xy_code = [&](int y, int x) -> int {
    if (!BULK || y < 0 || y + C::BH > P.in_h || x < 0 || x + C::BW > P.in_w) return -1;
    return (int)((base_el + (unsigned)y * (unsigned)P.in_w + (unsigned)x) & 1u) | (y + C::BH == P.in_h ? 2 : 0) | ((y | x) == 0 ? 4 : 0);
  };
  auto bulk_plane = [&](int code, int z, int &sa) -> bool {
    sa = 0;
    if (!BULK || code < 0 || z < 0 || z >= P.in_d || ((code & 2) && z == P.in_d - 1)) return false;
    const int a = (code + (z & hw1)) & 1;
    if (a && (code & 4) && z == 0) return false;  // (a shifted first row would read the element before the array)
    sa = a;
    return true;
  };
  auto issue_upto = [&](int limit) {  // limit = planes released so far + S
    Books q = BOOKS_SMEM ? *pb : rb;
    while (q.issued < q.total && q.issued < limit) {
      if (q.left == 0) {
        int oz0, nz, oy0, ox0;
        item_geom(q.it++, oz0, nz, oy0, ox0);
        q.left = nz + 4;
        q.x = ox0 - 2 * V; q.y = oy0 - 2; q.z = oz0 + P.sz2 - 2;
        q.code = xy_code(q.y, q.x);
      }
      const int s = q.issued % S;
      if (TMA) {
        mbar_arrive_expect_tx(&full[s], (uint32_t)(C::BW * C::BH * sizeof(T)));
#if MB_TB2_L2HINT
        tma_load_3d_hint(smem + s * SSTAGE, &tmap, &full[s], q.x, q.y, q.z, l2_policy(MB_TB2_L2HINT));
#else
        tma_load_3d(smem + s * SSTAGE, &tmap, &full[s], q.x, q.y, q.z);
#endif
      } else {
        int sa;
        if (BULK && bulk_plane(q.code, q.z, sa)) {
          // a box inside the array: one bulk copy per row, (shifted rows are 16 bytes
          // longer: the element before and the element after travel too)
          if (tid == 0) {
            const int nsh = sb ? (C::BH + sa) / 2 : sa * C::BH;  // rows r < BH with (sa + sb * r) odd
            mbar_arrive_expect_tx(&full[s], (uint32_t)(C::BW * C::BH * sizeof(T)) + 16u * (uint32_t)nsh);
          }
          // (the copy instruction takes its operands from uniform registers: a warp issues its lanes' copies one
          // after the other, so the rows are dealt out over all warps, a few lanes each)
          const int row = (tid >> 5) + MB_TB2_NW * (tid & 31);
          if ((tid & 31) < (C::BH + MB_TB2_NW - 1) / MB_TB2_NW && row < C::BH) {
            fence_proxy_async();  // the stage's last readers (generic proxy) come before the engine's writes
            const int sh = (sa + sb * row) & 1;
            const T *src = in + ((size_t)q.z * P.in_h + (q.y + row)) * P.in_w + q.x - sh;
            bulk_load_1d(smem + s * SSTAGE + (size_t)row * SP * sizeof(T), src, (uint32_t)(C::BW * sizeof(T)) + 16u * (uint32_t)sh, &full[s]);
          }
        } else {
          // the whole CTA copies the box, a warp per row with the widest copies the row's alignment allows; cells
          // outside the array become zeros, as with TMA (mb_tma.cuh).
          const T *pl = (q.z >= 0 && q.z < P.in_d) ? in + (size_t)q.z * P.in_h * P.in_w : nullptr;
          stage_box<T, C::BH, C::BW, MB_TB2_NW, SP>(reinterpret_cast<T *>(smem + s * SSTAGE), pl, in, P.in_h, P.in_w, q.y, q.x, tid);
        }
        cp_async_commit();  // one commit group per plane (empty for a bulk plane: the counts stay in step)
      }
      ++q.issued; ++q.z; --q.left;
    }
    if (BOOKS_SMEM) *pb = q; else rb = q;
  };
  if (tid == 0 || !TMA) issue_upto(S);

  uint32_t bphase = 0;  // (bulk path) parity of every stage's barrier: only some planes complete a phase
  // returns the plane's shift code: bit 0 = box row 0 is shifted, bit 1 = the rows alternate
  auto acquire = [&](int L, int z, int code) -> int {
    const int s = L % S;
    int sa;
    if (TMA) {
      mbar_wait(&full[s], (uint32_t)(L / S) & 1u);  // load L is the (L / S)-th use of its stage
      return 0;
    } else if (BULK && bulk_plane(code, z, sa)) {
      mbar_wait(&full[s], (bphase >> s) & 1u);
      bphase ^= 1u << s;
      return sa | (sb << 1);
    } else {
      // plane L's group must have landed; issued - 1 - L younger groups may still be in flight
      const int younger = rb.issued - 1 - L;
      if (younger <= 0) cp_async_wait_group<0>();
      else if (younger == 1) cp_async_wait_group<1>();
      else if (younger == 2) cp_async_wait_group<2>();
      else if (younger == 3) cp_async_wait_group<3>();
      else if (younger == 4) cp_async_wait_group<4>();
      else cp_async_wait_group<5>();
      __syncthreads();  // everybody's share of the plane is visible
      return 0;
    }
  };
  const int lx = (tid % (32 * C::WX)) * V, ly = (tid / (32 * C::WX)) * R;
  const int coff = (ly + 1) * BW + V + lx;   // the thread's first cell inside an intermediate plane
  const int scoff = (ly + 1) * SP + V + lx;  // ... inside a staged source box
  // (bulk path) the thread's column in box rows ly + j of a staged plane with shift code sc: rows with even j start at
  // pe + j * SP, rows with odd j at po + j * SP
  auto row_ptrs = [&](int L, int sc, const T *&pe, const T *&po) {
    const T *p = reinterpret_cast<const T *>(smem + (L % S) * SSTAGE) + ly * SP + V + lx;
    const int a = (sc + (sc >> 1) * ly) & 1;
    pe = p + a;
    po = p + (a ^ (sc >> 1));
  };
  auto centres = [&](T (&dst)[R][V], int L, int sc) {
    if constexpr (BULK) {
      const T *pe, *po;
      row_ptrs(L, sc, pe, po);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const T *p = (((r + 1) & 1) ? po : pe) + (r + 1) * SP;
#pragma unroll
        for (int v = 0; v < V; ++v) dst[r][v] = p[v];
      }
    } else {
      const T *sp = reinterpret_cast<const T *>(smem + (L % S) * SSTAGE) + scoff;
#pragma unroll
      for (int r = 0; r < R; ++r) lds_v<T>(dst[r], sp + r * SP);
    }
  };
  // the in-plane neighbours of the thread's cells in a staged source plane
  auto src_nbrs = [&](int L, int sc, const T (&pc)[R][V], Nbrs<T, R, V> &nb) {
    if constexpr (BULK) {
      const T *pe, *po;
      row_ptrs(L, sc, pe, po);
      const T *pb2 = (((R + 1) & 1) ? po : pe) + (R + 1) * SP;
#pragma unroll
      for (int v = 0; v < V; ++v) { nb.ytop[v] = pe[v]; nb.ybot[v] = pb2[v]; }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const T *p = (((r + 1) & 1) ? po : pe) + (r + 1) * SP;
        nb.xl[r] = p[-1];
        nb.xr[r] = p[V];
      }
    } else {
      load_nbrs<T, R, V, SP>(reinterpret_cast<const T *>(smem + (L % S) * SSTAGE) + scoff, pc, nb);
    }
  };

  int Lc = 0;
  uint32_t g = 0;  // steps taken by this CTA over all its items: intermediate planes alternate buffers with it
  for (int it = 0; it < my_items; ++it) {
    int oz0, nz, oy0, ox0;
    item_geom(it, oz0, nz, oy0, ox0);
    const int yB0 = oy0 - 1, xB0 = ox0 - V;  // origin of the stage-1 tile
    const int cs = oz0 + P.sz2;              // source-frame plane of the first result plane
    const int bz = oz0 + P.sz2 - 2, icode = xy_code(oy0 - 2, ox0 - 2 * V);  // the item's first source box
    const int sc0 = acquire(Lc, bz, icode);
    int sc_mid = acquire(Lc + 1, bz + 1, icode);  // shift code of the middle source plane of the next step
    // the three source planes and the three intermediate planes a step touches live in registers and
    // rotate by NAME (the loop is unrolled three times), not by copying
    T Z0[R][V], Z1[R][V], Z2[R][V], U0[R][V], U1[R][V], U2[R][V];
    centres(Z0, Lc, sc0);
    centres(Z1, Lc + 1, sc_mid);
    if (DETECT) {  // the two planes below the item's first result plane are nobody's result in this launch
#pragma unroll
      for (int r = 0; r < R; ++r) { track_row(bad, Z0[r]); track_row(bad, Z1[r]); }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < V; ++v) U0[r][v] = U1[r][v] = T(0);
    // which of the thread's cells exist in the global array (the intermediate is Fill 0 elsewhere),
    // and which of its rows are result cells (the tile minus its rim, inside the array)
    uint32_t rows = 0;
    const int gx = xB0 + lx, gy0 = yB0 + ly;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int by = ly + r, gy = yB0 + by;
      if (by >= 1 && by < C::TYB - 1 && lx >= V && lx < C::TXB - V && gy < P.in_h && gx < P.in_w) rows |= 1u << r;
    }
    // CTA-uniform: the whole stage-1 tile lies inside the array in y and x (true for all but the edge tiles):
    // the intermediate values then need no masking on planes inside the array
    const bool tile_inside = yB0 >= 0 && yB0 + C::TYB <= P.in_h && xB0 >= 0 && xB0 + C::TXB <= P.in_w;
    const bool whole = P.vec_store && gx + V <= P.in_w;
    const long long plane = (long long)P.in_h * P.pitch;
    long long ooff = ((long long)oz0 * P.in_h + (yB0 + ly)) * P.pitch + gx;  // row 0 of result plane oz0 (used where rows says so)

    // one step: source planes zm (below), zc (middle) in, zp loaded; intermediate planes um, uc in, up produced
    auto step = [&](int k, const T (&zm)[R][V], const T (&zc)[R][V], T (&zp)[R][V], const T (&um)[R][V], const T (&uc)[R][V],
                    T (&up)[R][V]) {
      const int ip = cs - 1 + k;  // intermediate plane produced in this step (source frame)
      // the middle source plane has been resident since the last step: its in-plane neighbours are fetched
      // before the wait for the new plane, so their latency overlaps it
      Nbrs<T, R, V> nb;
#if MB_TB2_HOIST
      src_nbrs(Lc + k + 1, sc_mid, zc, nb);
#endif
      const int sc_new = acquire(Lc + k + 2, bz + k + 2, icode);
      centres(zp, Lc + k + 2, sc_new);
#if !MB_TB2_HOIST
      src_nbrs(Lc + k + 1, sc_mid, zc, nb);
#endif
      sc_mid = sc_new;
      if (DETECT == 2 || (DETECT == 1 && k >= nz)) {  // ... nor are the two planes above its last one
#pragma unroll
        for (int r = 0; r < R; ++r) track_row(bad, zp[r]);
      }
      // ---- stage 1: first application, planes ip-1, ip, ip+1 of the source ----
      star7<T, REV, R, V>(nb, zm, zc, zp, K, up);
      const bool in_z = ip + P.zoff >= 0 && ip + P.zoff < P.dglob;
      T *ub = reinterpret_cast<T *>(ubase + (g & 1u) * C::STAGE) + coff;
      if (!MB_TB2_INSIDE || !(tile_inside && in_z)) {  // uniform: edge tiles and the planes beyond the array's ends only
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (!(in_z && (unsigned)(gy0 + r) < (unsigned)P.in_h && (unsigned)(gx + v) < (unsigned)P.in_w)) up[r][v] = T(0);
      }
      if (DETECT == 2 || (DETECT == 1 && (k == 0 || k == nz + 1))) {  // intermediate planes outside the result range
#pragma unroll
        for (int r = 0; r < R; ++r) track_row(bad, up[r]);
      }
      // the buffer was last read by step g-1's second application
      if (g > 0) mbar_wait(&ufree[g & 1u], ((g - 1) >> 1) & 1u);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        uint4 q;
        memcpy(&q, up[r], 16);
        *reinterpret_cast<uint4 *>(ub + r * BW) = q;
      }
      __syncthreads();  // intermediate plane ip is complete; source plane ip (the middle one) is no longer read
      if (tid == 0 || !TMA) issue_upto(((k == nz + 1) ? Lc + nz + 4 : Lc + k + 2) + S);
      // ---- stage 2: second application one plane behind, intermediate planes ip-2, ip-1, ip ----
      const T *umid = reinterpret_cast<const T *>(ubase + ((g + 1u) & 1u) * C::STAGE) + coff;
      if (k >= 2) {
        load_nbrs<T, R, V, BW>(umid, uc, nb);
        __syncwarp();  // this warp's reads of the plane are done: one arrival per warp
        if ((tid & 31) == 0) mbar_arrive(&ufree[(g + 1u) & 1u]);
        T res[R][V];
        star7<T, REV, R, V>(nb, um, uc, up, K, res);
#if MB_TB2_TRACKALL
        if (DETECT == 1) {
          // every cell's result is inspected, rim cells included (they are computed from finite-or-real data:
          // the buffers' rim was zeroed), so the check is branch-free
#pragma unroll
          for (int r = 0; r < R; ++r) track_row(bad, res[r]);
        }
        T *dst = out + ooff;
        if (whole) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            uint4 q;
            memcpy(&q, res[r], 16);
#if MB_TB2_STCS
            if ((rows >> r) & 1u) __stcs(reinterpret_cast<uint4 *>(dst + (size_t)r * P.pitch), q);
#else
            if ((rows >> r) & 1u) *reinterpret_cast<uint4 *>(dst + (size_t)r * P.pitch) = q;
#endif
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int v = 0; v < V; ++v)
              if (((rows >> r) & 1u) && gx + v < P.in_w) dst[(size_t)r * P.pitch + v] = res[r][v];
        }
#else
        T *dst = out + ooff;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((rows >> r) & 1u) {
            if (DETECT == 1) track_row(bad, res[r]);
            if (whole) {
              uint4 q;
              memcpy(&q, res[r], 16);
              *reinterpret_cast<uint4 *>(dst + (size_t)r * P.pitch) = q;
            } else {
#pragma unroll
              for (int v = 0; v < V; ++v)
                if (gx + v < P.in_w) dst[(size_t)r * P.pitch + v] = res[r][v];
            }
          }
        }
#endif
        ooff += plane;
      } else {
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&ufree[(g + 1u) & 1u]);
      }
      ++g;
    };
    const int nsteps = nz + 2;
#pragma unroll 1
    for (int k = 0;;) {
      step(k, Z0, Z1, Z2, U0, U1, U2);
      if (++k >= nsteps) break;
      step(k, Z1, Z2, Z0, U1, U2, U0);
      if (++k >= nsteps) break;
      step(k, Z2, Z0, Z1, U2, U0, U1);
      if (++k >= nsteps) break;
    }
    Lc += nz + 4;
  }
  if (DETECT && nonfinite_seen<T>(bad)) atomicOr(flag, 1u);
}

template <typename T, bool REV, int R_, int DETECT>
static int launch_tb2_r(Ctx *c, const Plan &p, const void *d_in, void *d_out, const OuterRange *range, int sz2, int64_t zoff,
                        int64_t dglob, unsigned *flag) {
  using C = TB2Cfg<T, R_>;
  TB2Params P;
  std::memset(&P, 0, sizeof(P));
  P.in_d = (int)p.in_dims[0]; P.in_h = (int)p.in_dims[1]; P.in_w = (int)p.in_dims[2];
  P.out_d = P.in_d - 2 * sz2;
  P.sz2 = sz2; P.zoff = (int)zoff; P.dglob = (int)dglob;
  P.z0 = range ? (int)range->o0 : 0;
  P.z1 = range ? (int)range->o1 : P.out_d;
  if (P.z1 <= P.z0) return MB_OK;
  P.tiles_x = (P.in_w + C::OX - 1) / C::OX;
  P.tiles_y = (P.in_h + C::OY - 1) / C::OY;
  P.pitch = c->out_pitch ? (int)c->out_pitch : P.in_w;
  P.vec_store = ((uintptr_t)d_out % 16 == 0) && ((size_t)P.pitch * sizeof(T)) % 16 == 0;
  Coef27b<T> K;
  std::memcpy(K.k, p.c1.data(), sizeof(T) * 27);
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  const uint32_t box[3] = {1u, (uint32_t)C::BH, (uint32_t)C::BW};
  P.use_tma = (!c->no_tma && make_tensor_map(&tmap, p.elem, p.esize, 3, d_in, p.in_dims, box, c->in_pitch)) ? 1 : 0;
  if (c->in_pitch && !P.use_tma) return fail(c, MB_ERR_UNSUPPORTED, "internal: padded arrays need a tensor map");
  auto kern_tma = vol3d_tb2_kernel<T, REV, R_, DETECT, true>;
  auto kern_cp = vol3d_tb2_kernel<T, REV, R_, DETECT, false>;
  auto kern = P.use_tma ? kern_tma : kern_cp;
  const int smem_bytes = P.use_tma ? C::SMEM : C::SMEM_NT;
  static int occ_dev[64][2] = {{0}};
  int &occ = occ_dev[c->device & 63][P.use_tma];
  if (occ == 0) {
    MB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    MB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, smem_bytes));
    if (occ < 1) occ = 1;
  }
  const int resident = c->num_sms * occ;
  const int tiles = P.tiles_x * P.tiles_y;
  const int nzt = P.z1 - P.z0;
  // Split the outer dimension into chunks so that the items fill whole waves of the resident CTAs:
  // every chunk costs four extra source planes (plus the pipeline fill), every partial wave costs a
  // whole item.  Pick the chunk count with the smallest estimated makespan.
  int chunks = 1;
  {
    const int max_chunks = nzt / 32 > 1 ? (nzt / 32 < 256 ? nzt / 32 : 256) : 1;
    long best = -1;
    for (int cnum = 1; cnum <= max_chunks; ++cnum) {
      const int zc = (nzt + cnum - 1) / cnum;
      const int real = (nzt + zc - 1) / zc;
      const long waves = ((long)tiles * real + resident - 1) / resident;
      const long cost = waves * (zc + 4 + 3);
      if (best < 0 || cost < best) { best = cost; chunks = real; P.zc = zc; }
    }
  }
  P.nitems = tiles * chunks;
  const int grid = resident < P.nitems ? resident : P.nitems;
  kern<<<grid, C::NT, smem_bytes, c->stream>>>(tmap, P, K, (const T *)d_in, (T *)d_out, flag);
  c->launches++;
  c->last_kernel = "vol3d_tb2";
  return check_cuda(c, cudaGetLastError(), "vol3d_tb2_kernel launch");
}

template <typename T>
static bool star_only(const Plan &p) {
  const T *k = reinterpret_cast<const T *>(p.c1.data());
  for (int t = 0; t < 27; ++t) {
    const bool star = t == 4 || t == 10 || t == 12 || t == 13 || t == 14 || t == 16 || t == 22;
    if (!star && k[t] != T(0)) return false;
    if (!(k[t] - k[t] == T(0))) return false;  // finite coefficients: Fill 0 stays 0 after one application
  }
  return true;
}

// descriptor-only eligibility (decides how many ghost planes a sharded slab carries)
bool tb2_desc_eligible(const mb_stencil_desc *d) {
  if (!d || d->rank != 3 || (d->kind != MB_CONV && d->kind != MB_CORR) || !d->coeffs || d->n_post) return false;
  if (d->border != MB_FILL) return false;  // (with or without the finite promise: the checked form covers the rest)
  if (d->elem != MB_F32 && d->elem != MB_F64) return false;
  const int es = d->elem == MB_F32 ? 4 : 8;
  for (int i = 0; i < es; ++i) if (d->fill[i] != 0) return false;
  for (int i = 0; i < 3; ++i) if (d->size[i] != 3 || d->stride[i] != 1) return false;
  for (int t = 0; t < 27; ++t) {
    const bool star = t == 4 || t == 10 || t == 12 || t == 13 || t == 14 || t == 16 || t == 22;
    const double k = d->elem == MB_F32 ? (double)((const float *)d->coeffs)[t] : ((const double *)d->coeffs)[t];
    if ((!star && k != 0.0) || !(k - k == 0.0)) return false;
  }
  return true;
}

// May two steps of this single-step plan be fused?  (the plan describes ONE application)
// 1: yes, under the caller's finite promise; 2: yes, optimistically — the kernel checks what the promise
// would have covered and the caller redoes the pair strictly when the check fails; 0: no.
int vol3d_tb2_mode(Ctx *c, const Plan &p) {
  if (!c->tb2 || c->force_generic || !p.post.empty()) return 0;
  if (p.rank != 3 || !p.unit_stride() || (p.kind != MB_CONV && p.kind != MB_CORR)) return 0;
  if (p.border != MB_FILL) return 0;
  for (int i = 0; i < p.esize; ++i) if (p.fill[i] != 0) return 0;
  for (int i = 0; i < 3; ++i)
    if (p.size[i] != 3 || p.in_dims[i] < 1 || p.in_dims[i] >= (1ll << 30)) return 0;
  if (p.shift[0] != 0 || p.shift[1] != 0 || p.shift[2] != 0) return 0;
  for (int i = 0; i < 3; ++i) if (p.out_dims[i] != p.in_dims[i]) return 0;
  if (p.out_elems < c->tile_min_elems) return 0;
  const bool star = p.elem == MB_F32 ? star_only<float>(p) : (p.elem == MB_F64 ? star_only<double>(p) : false);
  if (!star) return 0;
  if (p.flags & MB_FLAG_ASSUME_FINITE) return 1;
  return c->no_optimistic ? 0 : 2;
}

// the form the sharded loop (slabs with ghost planes) uses: 0 / 1 (promise) / 2 (checked), as above
int vol3d_tb2_slab_mode(Ctx *c, const Plan &p) {
  Plan q = p;
  q.flags |= MB_FLAG_ASSUME_FINITE;
  if (!vol3d_tb2_applicable(c, q)) return 0;
  if (p.flags & MB_FLAG_ASSUME_FINITE) return 1;
  return c->no_optimistic ? 0 : 2;
}

bool vol3d_tb2_applicable(Ctx *c, const Plan &p) {
  if (!(p.flags & MB_FLAG_ASSUME_FINITE)) return false;
  if (!c->tb2 || c->force_generic || !p.post.empty()) return false;
  if (p.rank != 3 || !p.unit_stride() || (p.kind != MB_CONV && p.kind != MB_CORR)) return false;
  if (p.border != MB_FILL) return false;
  for (int i = 0; i < p.esize; ++i) if (p.fill[i] != 0) return false;
  for (int i = 0; i < 3; ++i)
    if (p.size[i] != 3 || p.in_dims[i] < 1 || p.in_dims[i] >= (1ll << 30)) return false;
  if (p.shift[1] != 0 || p.shift[2] != 0 || p.out_dims[1] != p.in_dims[1] || p.out_dims[2] != p.in_dims[2]) return false;
  if (p.out_elems < c->tile_min_elems) return false;
  if (p.elem == MB_F32) return star_only<float>(p);
  if (p.elem == MB_F64) return star_only<double>(p);
  return false;
}

// Two applications in one pass.  `p` is the single-step plan whose in_dims describe the SOURCE buffer
// handed in here; sz2 = 0: both applications are size-preserving along dim 0 (single GPU);
// sz2 = 2: the buffer carries two ghost planes each side and in_dims[0] - 4 result planes come out.
int launch_vol3d_tb2(Ctx *c, const Plan &p, const void *d_in, void *d_out, const OuterRange *range, int sz2, int64_t zoff,
                     int64_t dglob, unsigned *detect) {
  const bool rev = p.kind == MB_CONV;
#define MB_TB2_GO(T, REV, DET) launch_tb2_r<T, REV, MB_TB2_R, DET>(c, p, d_in, d_out, range, sz2, zoff, dglob, detect)
#ifdef MB_TB2_QUICK  // compile-time experiments: only the Double convolution instantiations
  if (p.elem != MB_F64 || !rev) return -1;
  return detect ? MB_TB2_GO(double, true, 1) : MB_TB2_GO(double, true, 0);
#else
  if (p.elem == MB_F64) {
    const bool centre = reinterpret_cast<const double *>(p.c1.data())[13] != 0.0;
    if (detect && centre) return rev ? MB_TB2_GO(double, true, 1) : MB_TB2_GO(double, false, 1);
    if (detect) return rev ? MB_TB2_GO(double, true, 2) : MB_TB2_GO(double, false, 2);
    return rev ? MB_TB2_GO(double, true, 0) : MB_TB2_GO(double, false, 0);
  }
  const bool centre = reinterpret_cast<const float *>(p.c1.data())[13] != 0.0f;
  if (detect && centre) return rev ? MB_TB2_GO(float, true, 1) : MB_TB2_GO(float, false, 1);
  if (detect) return rev ? MB_TB2_GO(float, true, 2) : MB_TB2_GO(float, false, 2);
  return rev ? MB_TB2_GO(float, true, 0) : MB_TB2_GO(float, false, 0);
#endif
#undef MB_TB2_GO
}

}  // namespace mb
