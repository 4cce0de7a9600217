// Post-maps (`fmap f` over a stencil's results, include/massiv_b200.h: mb_post): the per-element
// functions, shared by the stand-alone pass (k_post.cu) and the kernels that apply them in their store
// epilogue.
#pragma once
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

// out of line on purpose: the tiled kernels call it only when a descriptor carries post-maps, and their
// unmapped fast path should not pay for its registers
template <typename T> __device__ __noinline__ T post_apply(T x, const PostParams &P) {
  for (int k = 0; k < P.n; ++k) {
    T c;
    unsigned long long b = P.operand[k];
    memcpy(&c, &b, sizeof(T));
    x = post_one<T>(P.op[k], x, c);
  }
  return x;
}

// inlined form for kernels with a single store site
template <typename T> __device__ __forceinline__ T post_apply_inline(T x, const int n, const int *op, const unsigned long long *operand) {
#pragma unroll 1
  for (int k = 0; k < n; ++k) {
    T c;
    unsigned long long b = operand[k];
    memcpy(&c, &b, sizeof(T));
    x = post_one<T>(op[k], x, c);
  }
  return x;
}

inline PostParams make_post_params(const Plan &p) {
  PostParams P;
  std::memset(&P, 0, sizeof(P));
  P.n = (int)p.post.size();
  for (int i = 0; i < P.n; ++i) {
    P.op[i] = p.post[i].op;
    std::memcpy(&P.operand[i], p.post[i].operand, 8);
  }
  return P;
}

}  // namespace mb
