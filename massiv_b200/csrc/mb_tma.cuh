// TMA (cp.async.bulk.tensor) + mbarrier plumbing: PTX wrappers for the device side and the
// tensor-map encoder for the host side.  The driver entry point is resolved at run time through
// the CUDA runtime, so libmassiv_b200.so does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/massiv_b200.h"

namespace mb {

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

inline bool tma_dtype(int elem, CUtensorMapDataType *dt) {
  switch (elem) {
    case MB_U8: case MB_I8: *dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; return true;
    case MB_U16: case MB_I16: *dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; return true;
    case MB_U32: case MB_I32: case MB_BOOL32: *dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; return true;
    case MB_U64: case MB_I64: *dt = CU_TENSOR_MAP_DATA_TYPE_UINT64; return true;
    case MB_F32: *dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; return true;
    case MB_F64: *dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64; return true;
  }
  return false;
}

// dims/box are given OUTERMOST FIRST (the library's convention); TMA wants innermost first.
// Returns false when the array cannot be described (alignment / size limits): the caller then
// fills tiles with ordinary loads instead.
// row_pitch (elements, 0 = dense): the distance between rows of the innermost dimension when the array is padded.
inline bool make_tensor_map(CUtensorMap *m, int elem, int esize, int rank, const void *base, const int64_t *dims,
                            const uint32_t *box, int64_t row_pitch = 0) {
  EncodeTiledFn enc = get_encode_tiled();
  CUtensorMapDataType dt;
  if (!enc || !tma_dtype(elem, &dt)) return false;
  if (((uintptr_t)base & 15) != 0) return false;
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bdim[5], estr[5];
  uint64_t pitch = (uint64_t)esize;
  for (int i = 0; i < rank; ++i) {
    const int o = rank - 1 - i;  // outermost-first index of TMA dim i
    if (dims[o] <= 0 || dims[o] > (int64_t)1 << 31) return false;
    gdim[i] = (cuuint64_t)dims[o];
    bdim[i] = box[o];
    estr[i] = 1;
    if (box[o] == 0 || box[o] > 256) return false;
    if (i > 0) {
      if (pitch % 16 != 0 || pitch >= ((uint64_t)1 << 40)) return false;
      gstr[i - 1] = pitch;
    }
    pitch *= (uint64_t)((i == 0 && row_pitch > 0) ? row_pitch : dims[o]);
  }
  if (((uint64_t)box[rank - 1] * esize) % 16 != 0) return false;
  // L2 promotion: experiment knob MASSIV_B200_L2PROMO = 0 (none) / 64 / 128 / 256
  static int promo = -1;
  if (promo < 0) {
    const char *e = getenv("MASSIV_B200_L2PROMO");
    promo = e ? atoi(e) : 64;
  }
  const CUtensorMapL2promotion l2p = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                     : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                    : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  CUresult r = enc(m, dt, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// ---- device ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's generic-proxy shared-memory accesses before later async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): 1 = evict_last (data other CTAs will read again), 2 = evict_first
__device__ __forceinline__ uint64_t l2_policy(int kind) {
  uint64_t pol;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "l"(pol)
      : "memory");
}
// one contiguous run global -> shared through the bulk-copy engine (UBLKCP): both addresses 16-byte aligned, bytes a
// multiple of 16, completion counted on the mbarrier like a tensor load
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// ---- cp.async (LDGSTS): the staging path for arrays TMA cannot describe (row pitch not a multiple of
// 16 bytes, e.g. an odd number of Floats per row) ----
// one element (4 or 8 bytes) global -> shared, no register staging; src_bytes = 0 writes zeros instead
template <int BYTES> __device__ __forceinline__ void cp_async_elem(void *dst_smem, const void *src, int src_bytes = BYTES) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(smem_u32(dst_smem)), "l"(src), "n"(BYTES), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- staging WITHOUT a tensor map ----
// A tensor map needs row pitches in whole 16-byte units; massiv arrays have no pitch rule, so an array with an
// odd number of Doubles (or a number of Floats that is not a multiple of 4) per row has none.  Its boxes are staged
// with cp.async (LDGSTS) by the whole CTA, a warp per box row, each row with the widest copy its alignment allows:
// the destination row starts on a 16-byte boundary, so a row whose first source byte does too moves in 16-byte
// units, one that is 8-byte aligned in 8-byte units, the rest element by element.  A copy costs the same issue
// and shared-memory slots whatever its size (the one-element version this replaces was LSU-bound), so this is
// 1.3x (Double, odd width) to 1.5x (Float, odd width) fewer copy instructions than element copies.
// Measured and rejected (Sobel Float 32767^2, tensor-map path 594 Gcells/s, element copies 364): one cp.async.bulk
// per box row on the TMA ring followed by a realignment pass in shared memory, 377; 16-byte copies from the
// rounded-down address + realignment pass, 393 (the pass and its barrier cost more than the wider copies save).
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
template <typename T> __device__ __forceinline__ int lead_elems(const T *p) {
  return (int)((reinterpret_cast<uintptr_t>(p) & 15) / sizeof(T));
}
// One BH x BW box (BW a multiple of the vector length, dst 16-byte aligned, rows BW apart) whose first cell is
// (y0, x0) of a plane of H rows x W columns at `plane` (nullptr: the whole plane lies outside the array).
// zero_outside: cells outside the array are written as zeros (src-size 0) — what TMA's out-of-range fill does;
// otherwise the caller guarantees that the box lies inside.  `any` is a valid address for the zero-size copies.
// The caller commits the group.
template <typename T, int BH, int BW, int NWARPS, int PITCH = BW>
__device__ __forceinline__ void stage_box(T *dst, const T *plane, const T *any, int H, int W, int y0, int x0, int tid) {
  constexpr int V = 16 / (int)sizeof(T);
  const int lane = tid & 31;
  for (int r = tid >> 5; r < BH; r += NWARPS) {
    T *drow = dst + r * PITCH;
    const int gy = y0 + r;
    const bool row_ok = plane != nullptr && gy >= 0 && gy < H;
    if (row_ok && x0 >= 0 && x0 + BW <= W) {
      const T *srow = plane + (size_t)gy * W + x0;
      const int lead_bytes = (int)(reinterpret_cast<uintptr_t>(srow) & 15);
      if (lead_bytes == 0) {
        for (int j = lane; j < BW / V; j += 32) cp_async16(drow + j * V, srow + j * V);
      } else if (sizeof(T) == 8 || (lead_bytes & 7) == 0) {
        constexpr int E8 = 8 / (int)sizeof(T) > 0 ? 8 / (int)sizeof(T) : 1;  // elements per 8-byte copy
        for (int j = lane; j < BW / E8; j += 32) cp_async_elem<8>(drow + j * E8, srow + j * E8);
      } else {
        for (int c = lane; c < BW; c += 32) cp_async_elem<(sizeof(T) >= 4 ? (int)sizeof(T) : 4)>(drow + c, srow + c);
      }
    } else {
      for (int c = lane; c < BW; c += 32) {
        const int gx = x0 + c;
        const bool ok = row_ok && gx >= 0 && gx < W;
        const T *src = ok ? plane + (size_t)gy * W + gx : any;
        cp_async_elem<(sizeof(T) >= 4 ? (int)sizeof(T) : 4)>(drow + c, src, ok ? (int)sizeof(T) : 0);
      }
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace mb
