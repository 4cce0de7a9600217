// Device self-tests exported through the C ABI (used by tests/, never by the compute path).
#include "mb_arith.cuh"
#include "mb_internal.h"

namespace mb {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Compares div_count (the three-operation exact quotient used by avgStencil) with the IEEE divide
// over `count` pseudo-random bit patterns: uniformly random bits (all exponents, NaN/Inf, denormals),
// plus patterns steered towards the dangerous regions (results near the subnormal boundary and
// mantissas of the form n*(2k+1) +- 1 that bring x/n close to a rounding boundary).
template <typename T, typename U>
__global__ void div_selftest_kernel(DivCount dc, unsigned long long count, unsigned long long seed,
                                    unsigned long long *mismatches, unsigned long long *first_bad) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count; i += stride) {
    unsigned long long h = mix64(seed + i);
    U bits = (U)h;
    const int mode = (int)(i % 4);
    constexpr int MB = sizeof(T) == 8 ? 52 : 23;
    if (mode == 1) {  // quotient around the smallest normal number
      const U mant = bits & (((U)1 << MB) - 1);
      const U e = (U)(1 + (h >> 60));  // biased exponent 1..16
      bits = (bits & ((U)1 << (sizeof(T) * 8 - 1))) | (e << MB) | mant;
    } else if (mode == 2) {  // x = n * (odd 54-bit-ish value) +- small: near-midpoint quotients
      const unsigned long long k = (h >> 12) | 1ull;
      const T base = (T)((double)k) * (sizeof(T) == 8 ? (T)dc.n_d : (T)dc.n_f);
      U b2;
      memcpy(&b2, &base, sizeof(T));
      b2 += (U)((h >> 3) % 5) - 2;
      bits = b2;
    }
    T x;
    memcpy(&x, &bits, sizeof(T));
    T want, got;
    if constexpr (sizeof(T) == 8) { want = __ddiv_rn(x, dc.n_d); } else { want = __fdiv_rn(x, dc.n_f); }
    got = div_count(x, dc);
    U wb, gb;
    memcpy(&wb, &want, sizeof(T));
    memcpy(&gb, &got, sizeof(T));
    const bool both_nan = (want != want) && (got != got);
    if (wb != gb && !both_nan) {
      ++bad;
      atomicCAS(first_bad, 0ull, (unsigned long long)bits | (sizeof(T) == 4 ? (1ull << 63) : 0ull));
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}

// bitwise comparison of two device buffers (tests / bench verification): counts differing bytes' words and
// reports the first differing byte offset
__global__ void compare_kernel(const unsigned char *a, const unsigned char *b, unsigned long long bytes, bool words,
                               unsigned long long *count, unsigned long long *first) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long bad = 0, where = ~0ull;
  if (words) {
    const unsigned long long n = bytes / 8;
    const unsigned long long *x = reinterpret_cast<const unsigned long long *>(a), *y = reinterpret_cast<const unsigned long long *>(b);
    for (unsigned long long i = tid; i < n; i += stride)
      if (x[i] != y[i]) { ++bad; if (where == ~0ull) where = i * 8; }
    for (unsigned long long i = n * 8 + tid; i < bytes; i += stride)
      if (a[i] != b[i]) { ++bad; if (where == ~0ull) where = i; }
  } else {
    for (unsigned long long i = tid; i < bytes; i += stride)
      if (a[i] != b[i]) { ++bad; if (where == ~0ull) where = i; }
  }
  if (bad) { atomicAdd(count, bad); atomicMin(first, where); }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_compare_dev(mb_ctx *h, const void *dev_a, const void *dev_b, size_t bytes, uint64_t *mismatches,
                              uint64_t *first_offset) {
  if (!h || !mismatches) return MB_ERR_INVALID_DESC;
  Ctx *c = &h->c;
  std::lock_guard<std::mutex> lk(c->mu);
  MB_CUDA(c, cudaSetDevice(c->device));
  unsigned long long *d = nullptr;
  MB_CUDA(c, cudaMalloc(&d, 16));
  const unsigned long long init[2] = {0ull, ~0ull};
  MB_CUDA(c, cudaMemcpyAsync(d, init, 16, cudaMemcpyHostToDevice, c->stream));
  if (bytes) {
    const bool words = ((uintptr_t)dev_a % 8 == 0) && ((uintptr_t)dev_b % 8 == 0);
    compare_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>((const unsigned char *)dev_a, (const unsigned char *)dev_b,
                                                          (unsigned long long)bytes, words, d, d + 1);
    c->launches++;
    c->aux_launches++;
  }
  unsigned long long hres[2] = {0, 0};
  MB_CUDA(c, cudaMemcpyAsync(hres, d, 16, cudaMemcpyDeviceToHost, c->stream));
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  MB_CUDA(c, cudaFree(d));
  *mismatches = hres[0];
  if (first_offset) *first_offset = hres[1];
  return MB_OK;
}

extern "C" int mb_selftest_div(mb_ctx *h, int elem, int64_t n, uint64_t count, uint64_t seed, uint64_t *mismatches,
                               uint64_t *first_bad_bits) {
  if (!h || !mismatches) return MB_ERR_INVALID_DESC;
  Ctx *c = &h->c;
  std::lock_guard<std::mutex> lk(c->mu);
  MB_CUDA(c, cudaSetDevice(c->device));
  if (elem != MB_F32 && elem != MB_F64) return fail(c, MB_ERR_UNSUPPORTED, "division self-test needs a floating type");
  DivCount dc;
  dc.n_d = (double)n; dc.n_f = (float)n;
  dc.fast = n >= 1 && n <= 1024;
  dc.y_d = 1.0 / dc.n_d; dc.y_f = 1.0f / dc.n_f;
  unsigned long long *d = nullptr;
  MB_CUDA(c, cudaMalloc(&d, 16));
  MB_CUDA(c, cudaMemsetAsync(d, 0, 16, c->stream));
  if (elem == MB_F64)
    div_selftest_kernel<double, unsigned long long><<<c->num_sms * 8, 256, 0, c->stream>>>(dc, count, seed, d, d + 1);
  else
    div_selftest_kernel<float, unsigned int><<<c->num_sms * 8, 256, 0, c->stream>>>(dc, count, seed, d, d + 1);
  c->launches++;
  unsigned long long hres[2] = {0, 0};
  MB_CUDA(c, cudaMemcpyAsync(hres, d, 16, cudaMemcpyDeviceToHost, c->stream));
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  MB_CUDA(c, cudaFree(d));
  *mismatches = hres[0];
  if (first_bad_bits) *first_bad_bits = hres[1];
  return MB_OK;
}
