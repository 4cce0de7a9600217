// Convergence predicates for iterateUntil (Array/Manifest/Internal.hs:331-349): the reference hands the last two
// arrays to a closure; the device path evaluates the two predicates people write — "nothing changed" and "the largest
// change is below a tolerance" — as one reduction over both buffers, and only the verdict travels to the host.
#include "mb_arith.cuh"
#include "mb_internal.h"

namespace mb {

// result[0]: number of cells with !(cur == prev) (Eq e: NaN /= NaN, -0 == 0)
// result[1]: bit pattern of max |cur - prev| as a non-negative Double (NaN sorts above +Inf: never below a tolerance)
template <typename T>
__global__ void __launch_bounds__(256) converge_kernel(const T *__restrict__ cur, const T *__restrict__ prev, size_t n,
                                                       unsigned long long *result) {
  unsigned long long unequal = 0, maxbits = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const T a = cur[i], b = prev[i];
    if (!(a == b)) ++unequal;
    if constexpr (Arith<T>::is_float) {
      const double d = fabs((double)a - (double)b);  // exact for Float; one rounding for Double, as `abs (a - b)` has
      const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
      maxbits = bits > maxbits ? bits : maxbits;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    unequal += __shfl_down_sync(0xffffffffu, unequal, o);
    const unsigned long long other = __shfl_down_sync(0xffffffffu, maxbits, o);
    maxbits = other > maxbits ? other : maxbits;
  }
  if ((threadIdx.x & 31) == 0) {
    if (unequal) atomicAdd(&result[0], unequal);
    if (maxbits) atomicMax(&result[1], maxbits);
  }
}

template <typename T>
static int launch_conv_t(Ctx *c, const void *cur, const void *prev, size_t n, unsigned long long *d_res) {
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)c->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  converge_kernel<T><<<(unsigned)blocks, 256, 0, c->stream>>>((const T *)cur, (const T *)prev, n, d_res);
  c->launches++;
  c->aux_launches++;
  return check_cuda(c, cudaGetLastError(), "converge_kernel launch");
}

// blocks until the verdict is known: *unequal cells that differ, *max_abs_diff the largest |cur - prev| (0 for integers)
int converge_check(Ctx *c, int elem, const void *cur, const void *prev, size_t n, unsigned long long *unequal, double *max_abs_diff) {
  if (!c->d_conv) MB_CUDA(c, cudaMalloc(&c->d_conv, 16));
  MB_CUDA(c, cudaMemsetAsync(c->d_conv, 0, 16, c->stream));
  int st = MB_OK;
  switch (elem) {
    case MB_U8: st = launch_conv_t<uint8_t>(c, cur, prev, n, c->d_conv); break;
    case MB_I8: st = launch_conv_t<int8_t>(c, cur, prev, n, c->d_conv); break;
    case MB_U16: st = launch_conv_t<uint16_t>(c, cur, prev, n, c->d_conv); break;
    case MB_I16: st = launch_conv_t<int16_t>(c, cur, prev, n, c->d_conv); break;
    case MB_U32: st = launch_conv_t<uint32_t>(c, cur, prev, n, c->d_conv); break;
    case MB_I32: case MB_BOOL32: st = launch_conv_t<int32_t>(c, cur, prev, n, c->d_conv); break;
    case MB_U64: st = launch_conv_t<uint64_t>(c, cur, prev, n, c->d_conv); break;
    case MB_I64: st = launch_conv_t<int64_t>(c, cur, prev, n, c->d_conv); break;
    case MB_F32: st = launch_conv_t<float>(c, cur, prev, n, c->d_conv); break;
    case MB_F64: st = launch_conv_t<double>(c, cur, prev, n, c->d_conv); break;
    default: return fail(c, MB_ERR_INVALID_DESC, "unknown element type");
  }
  if (st) return st;
  unsigned long long h[2] = {0, 0};
  MB_CUDA(c, cudaMemcpyAsync(h, c->d_conv, 16, cudaMemcpyDeviceToHost, c->stream));
  MB_CUDA(c, cudaStreamSynchronize(c->stream));
  *unequal = h[0];
  long long bits = (long long)h[1];
  std::memcpy(max_abs_diff, &bits, 8);
  return MB_OK;
}

}  // namespace mb
