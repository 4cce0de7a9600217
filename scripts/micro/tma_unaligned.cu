// Does a TMA tile load accept a box whose first element is NOT 16-byte aligned in global memory (inner coordinate not a
// multiple of 16 / sizeof(T))?  And row-phase maps: an array of odd width W read through maps whose base is rounded
// down to 16 bytes and whose rows are V apart.     nvcc -arch=sm_100a -o build/tma_unaligned scripts/micro/tma_unaligned.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../massiv_b200/csrc/mb_tma.cuh"
using namespace mb;

__global__ void k(const __grid_constant__ CUtensorMap m, int x, int y, float *out) {
  __shared__ __align__(128) float sm[8 * 16];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&bar, 8 * 16 * 4);
    tma_load_2d(sm, &m, &bar, x, y);
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) out[i] = sm[i];
}

int main() {
  const int H = 37, W = 67;  // odd width: rows are 268 bytes apart
  std::vector<float> h(H * W + 64);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, 128 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  std::vector<float> r(128);
  int bad_total = 0;
  {  // 1. an aligned array (W = 64 view of the same memory), unaligned inner coordinate
    CUtensorMap m;
    const int64_t dims[2] = {32, 64};
    const uint32_t box[2] = {8, 16};
    if (!make_tensor_map(&m, MB_F32, 4, 2, d, dims, box)) { printf("map failed\n"); return 1; }
    for (int x : {0, 1, 2, 3, 5, 49, 61, -3}) {
      cudaMemset(o, 0xff, 512);
      k<<<1, 64>>>(m, x, 3, o);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("x=%d: %s\n", x, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(r.data(), o, 512, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int rr = 0; rr < 8; ++rr)
        for (int c = 0; c < 16; ++c) {
          const int gx = x + c, gy = 3 + rr;
          const float want = (gx >= 0 && gx < 64) ? (float)(gy * 64 + gx) : 0.f;
          bad += r[rr * 16 + c] != want;
        }
      printf("aligned array, x=%d: %d wrong cells\n", x, bad);
      bad_total += bad;
    }
  }
  {  // 2. row-phase maps of the odd-width array: phase p holds rows p, p + 4, ...; box of ONE row
    for (int p = 0; p < 4; ++p) {
      const uintptr_t addr = (uintptr_t)d + (size_t)p * W * 4, base = addr & ~(uintptr_t)15;
      const int lead = (int)((addr - base) / 4);
      EncodeTiledFn enc = get_encode_tiled();
      CUtensorMap m;
      cuuint64_t gdim[2] = {(cuuint64_t)(W + lead), (cuuint64_t)((H - p + 3) / 4)}, gstr[1] = {(cuuint64_t)16 * W};
      cuuint32_t bdim[2] = {16, 8}, estr[2] = {1, 1};
      CUresult cr = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) { printf("phase map %d failed %d\n", p, (int)cr); return 1; }
      for (int x : {0, 7, 60}) {
        cudaMemset(o, 0xff, 512);
        k<<<1, 64>>>(m, x + lead, 1, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("phase %d x=%d: %s\n", p, x, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(r.data(), o, 512, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < 8; ++rr)
          for (int c = 0; c < 16; ++c) {
            const int gx = x + c, gy = p + 4 * (1 + rr);
            const float want = (gx < W && gy < H) ? (float)(gy * W + gx) : 0.f;
            bad += r[rr * 16 + c] != want;
          }
        printf("phase %d (lead %d), x=%d: %d wrong cells\n", p, lead, x, bad);
        bad_total += bad;
      }
    }
  }
  printf(bad_total ? "FAIL\n" : "ALL OK\n");
  return bad_total != 0;
}
