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
inline bool make_tensor_map(CUtensorMap *m, int elem, int esize, int rank, const void *base, const int64_t *dims,
                            const uint32_t *box) {
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
    pitch *= (uint64_t)dims[o];
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
// ---- cp.async (LDGSTS): the staging path for arrays TMA cannot describe (row pitch not a multiple of
// 16 bytes, e.g. an odd number of Floats per row) ----
// one element (4 or 8 bytes) global -> shared, no register staging; src_bytes = 0 writes zeros instead
template <int BYTES> __device__ __forceinline__ void cp_async_elem(void *dst_smem, const void *src, int src_bytes = BYTES) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(smem_u32(dst_smem)), "l"(src), "n"(BYTES), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace mb
