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

// ---- expression stencils (MB_EXPR): the terms' tap ranges and the postfix program, passed by value ----
struct XParams {
  int n_terms, n_prog;
  int term_kind[MB_MAX_TERMS], term_begin[MB_MAX_TERMS], term_ntaps[MB_MAX_TERMS], term_avg_n[MB_MAX_TERMS];
  unsigned char op[MB_MAX_PROGRAM], arg[MB_MAX_PROGRAM];
  unsigned long long operand[MB_MAX_PROGRAM];
};

template <typename T> __device__ __forceinline__ T xop_binary(int op, T a, T b) {
  switch (op) {  // liftA2 (+|-|*|/|min|max): the post-map forms with the second operand as the constant
    case MB_X_ADD: return post_one<T>(MB_POST_ADD, a, b);
    case MB_X_SUB: return post_one<T>(MB_POST_SUB, a, b);
    case MB_X_MUL: return post_one<T>(MB_POST_MUL, a, b);
    case MB_X_DIV: return post_one<T>(MB_POST_DIV, a, b);
    case MB_X_MIN: return post_one<T>(MB_POST_MIN, a, b);
    default: return post_one<T>(MB_POST_MAX, a, b);
  }
}
template <typename T> __device__ __forceinline__ T xop_unary(int op, T a) {
  switch (op) {
    case MB_X_NEGATE: return post_one<T>(MB_POST_NEGATE, a, T(0));
    case MB_X_ABS: return post_one<T>(MB_POST_ABS, a, T(0));
    case MB_X_SIGNUM: return post_one<T>(MB_POST_SIGNUM, a, T(0));
    case MB_X_SQUARE: return post_one<T>(MB_POST_SQUARE, a, T(0));
    case MB_X_RECIP: return post_one<T>(MB_POST_RECIP, a, T(0));
    default: return post_one<T>(MB_POST_SQRT, a, T(0));
  }
}
// the program over one cell's term values (validated on the host: no underflow, depth <= MB_MAX_STACK)
template <typename T> __device__ __forceinline__ T run_program(const XParams &X, const T *tv) {
  T stack[MB_MAX_STACK];
  int sp = 0;
  for (int k = 0; k < X.n_prog; ++k) {
    const int op = X.op[k];
    if (op == MB_X_TERM) {
      stack[sp++] = tv[X.arg[k]];
    } else if (op == MB_X_CONST) {
      T c;
      unsigned long long b = X.operand[k];
      memcpy(&c, &b, sizeof(T));
      stack[sp++] = c;
    } else if (op <= MB_X_MAX) {
      const T b = stack[--sp];
      stack[sp - 1] = xop_binary<T>(op, stack[sp - 1], b);
    } else {
      stack[sp - 1] = xop_unary<T>(op, stack[sp - 1]);
    }
  }
  return stack[0];
}

inline XParams make_xparams(const Plan &p) {
  XParams X;
  std::memset(&X, 0, sizeof(X));
  X.n_terms = (int)p.terms.size();
  X.n_prog = (int)p.prog.size();
  for (int t = 0; t < X.n_terms; ++t) {
    X.term_kind[t] = p.terms[t].kind; X.term_begin[t] = p.terms[t].begin;
    X.term_ntaps[t] = p.terms[t].ntaps; X.term_avg_n[t] = p.terms[t].avg_n;
  }
  for (int k = 0; k < X.n_prog; ++k) {
    X.op[k] = (unsigned char)p.prog[k].op; X.arg[k] = (unsigned char)p.prog[k].arg;
    std::memcpy(&X.operand[k], p.prog[k].operand, 8);
  }
  return X;
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
