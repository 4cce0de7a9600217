// Post-maps: `fmap f` over a stencil's results (Functor (Stencil ix e) / Functor (Array DW ix),
// Stencil/Internal.hs:36-41, Delayed/Windowed.hs:78-84) for the f listed in mb_post — each one an
// exactly specified IEEE-754 / two's-complement operation, so the mapped result stays bit-identical.
// First version: an in-place elementwise pass over what the stencil kernel wrote (one extra read and
// write of the OUTPUT per launch); fusing it into the kernels' store epilogues is the follow-up.
#include "mb_arith.cuh"
#include "mb_internal.h"

namespace mb {

struct PostParams {
  int n;
  int op[MB_MAX_POST];
  unsigned long long operand[MB_MAX_POST];
};

template <typename T> __device__ __forceinline__ T post_one(int op, T x, T c) {
  using A = Arith<T>;
  if constexpr (A::is_float) {
    switch (op) {
      case MB_POST_NEGATE: return -x;                                    // negateFloat#: sign flip, also for 0 and NaN
      case MB_POST_ABS: return x == T(0) ? T(0) : (x > T(0) ? x : -x);   // GHC.Float: abs (-0) = 0
      case MB_POST_SIGNUM: return x > T(0) ? T(1) : (x < T(0) ? T(-1) : x);
      case MB_POST_SQUARE: return A::mul(x, x);
      case MB_POST_ADD: return A::add(x, c);
      case MB_POST_SUB: return x - c;
      case MB_POST_RSUB: return c - x;
      case MB_POST_MUL: return A::mul(x, c);
      case MB_POST_MIN: return x <= c ? x : c;
      case MB_POST_MAX: return x <= c ? c : x;
      case MB_POST_DIV: return A::div(x, c);
      case MB_POST_RDIV: return A::div(c, x);
      case MB_POST_RECIP: return A::div(T(1), x);
      default: return A::sqrt(x);
    }
  } else {
    using U = typename A::Acc;  // unsigned arithmetic wraps like Haskell's fixed-width integers
    switch (op) {
      case MB_POST_NEGATE: return (T)(U(0) - (U)x);
      case MB_POST_ABS: return x >= T(0) ? x : (T)(U(0) - (U)x);
      case MB_POST_SIGNUM: return x > T(0) ? T(1) : (x < T(0) ? (T)(U(0) - U(1)) : T(0));
      case MB_POST_SQUARE: return (T)((U)x * (U)x);
      case MB_POST_ADD: return (T)((U)x + (U)c);
      case MB_POST_SUB: return (T)((U)x - (U)c);
      case MB_POST_RSUB: return (T)((U)c - (U)x);
      case MB_POST_MUL: return (T)((U)x * (U)c);
      case MB_POST_MIN: return x <= c ? x : c;
      default: return x <= c ? c : x;  // MB_POST_MAX (the Fractional ones are rejected at validation)
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) post_kernel(const PostParams P, T *__restrict__ data, size_t n) {
  T c[MB_MAX_POST];
#pragma unroll
  for (int i = 0; i < MB_MAX_POST; ++i) {
    unsigned long long b = P.operand[i];
    memcpy(&c[i], &b, sizeof(T));
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    T x = data[i];
#pragma unroll
    for (int k = 0; k < MB_MAX_POST; ++k)
      if (k < P.n) x = post_one<T>(P.op[k], x, c[k]);
    data[i] = x;
  }
}

template <typename T>
static int launch_post_t(Ctx *c, const PostParams &P, void *data, size_t n) {
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)c->num_sms * 16;
  if (blocks > cap) blocks = cap;
  post_kernel<T><<<(unsigned)blocks, 256, 0, c->stream>>>(P, (T *)data, n);
  c->launches++;
  c->aux_launches++;
  return check_cuda(c, cudaGetLastError(), "post_kernel launch");
}

int launch_post(Ctx *c, const Plan &p, void *d_out, const OuterRange *range) {
  if (p.post.empty()) return MB_OK;
  PostParams P;
  std::memset(&P, 0, sizeof(P));
  P.n = (int)p.post.size();
  for (int i = 0; i < P.n; ++i) {
    P.op[i] = p.post[i].op;
    std::memcpy(&P.operand[i], p.post[i].operand, 8);
  }
  int64_t plane = 1;
  for (int i = 1; i < p.rank; ++i) plane *= p.out_dims[i];
  int64_t o0 = 0, o1 = p.out_dims[0];
  if (range) { o0 = range->o0; o1 = range->o1; }
  if (o1 <= o0 || plane == 0) return MB_OK;
  char *base = (char *)d_out + (size_t)o0 * plane * p.esize;
  const size_t n = (size_t)(o1 - o0) * plane;
  switch (p.elem) {
    case MB_U8: return launch_post_t<uint8_t>(c, P, base, n);
    case MB_I8: return launch_post_t<int8_t>(c, P, base, n);
    case MB_U16: return launch_post_t<uint16_t>(c, P, base, n);
    case MB_I16: return launch_post_t<int16_t>(c, P, base, n);
    case MB_U32: return launch_post_t<uint32_t>(c, P, base, n);
    case MB_I32: return launch_post_t<int32_t>(c, P, base, n);
    case MB_U64: return launch_post_t<uint64_t>(c, P, base, n);
    case MB_I64: return launch_post_t<int64_t>(c, P, base, n);
    case MB_F32: return launch_post_t<float>(c, P, base, n);
    case MB_F64: return launch_post_t<double>(c, P, base, n);
  }
  return fail(c, MB_ERR_UNSUPPORTED, "post-maps on this element type");
}

}  // namespace mb
