// Element arithmetic with the reference's rounding model, shared by all kernels.
//
// Floating point: GHC on x86-64 compiles `x * k + acc` to separate IEEE multiply and add
// (SURVEY.md §8c "FP model"), so every operation here is an explicit round-to-nearest intrinsic
// (__fmul_rn / __fadd_rn / ...), which nvcc never contracts into FMA.  Division and sqrt are the
// IEEE-correct __fdiv_rn / __fsqrt_rn forms.  Integers wrap modulo 2^n like Haskell's Int/Word
// types: arithmetic is done in an unsigned accumulator and truncated on store.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/massiv_b200.h"

namespace mb {

template <typename T> struct Arith;

template <> struct Arith<float> {
  using Acc = float;
  static constexpr bool is_float = true;
  __device__ __forceinline__ static Acc add(Acc a, Acc b) { return __fadd_rn(a, b); }
  __device__ __forceinline__ static Acc mul(Acc a, Acc b) { return __fmul_rn(a, b); }
  __device__ __forceinline__ static Acc div(Acc a, Acc b) { return __fdiv_rn(a, b); }
  __device__ __forceinline__ static Acc sqrt(Acc a) { return __fsqrt_rn(a); }
  __device__ __forceinline__ static Acc from_count(int64_t n) { return (float)n; }
  __device__ __forceinline__ static Acc load(float v) { return v; }
  __device__ __forceinline__ static float store(Acc v) { return v; }
  __device__ __forceinline__ static bool is_zero(Acc k) { return k == 0.0f; }
};

template <> struct Arith<double> {
  using Acc = double;
  static constexpr bool is_float = true;
  __device__ __forceinline__ static Acc add(Acc a, Acc b) { return __dadd_rn(a, b); }
  __device__ __forceinline__ static Acc mul(Acc a, Acc b) { return __dmul_rn(a, b); }
  __device__ __forceinline__ static Acc div(Acc a, Acc b) { return __ddiv_rn(a, b); }
  __device__ __forceinline__ static Acc sqrt(Acc a) { return __dsqrt_rn(a); }
  __device__ __forceinline__ static Acc from_count(int64_t n) { return (double)n; }
  __device__ __forceinline__ static Acc load(double v) { return v; }
  __device__ __forceinline__ static double store(Acc v) { return v; }
  __device__ __forceinline__ static bool is_zero(Acc k) { return k == 0.0; }
};

template <typename T, typename A> struct IntArith {
  using Acc = A;
  static constexpr bool is_float = false;
  __device__ __forceinline__ static Acc add(Acc a, Acc b) { return a + b; }
  __device__ __forceinline__ static Acc mul(Acc a, Acc b) { return a * b; }
  __device__ __forceinline__ static Acc div(Acc a, Acc) { return a; }   // unreachable: AVG is Fractional
  __device__ __forceinline__ static Acc sqrt(Acc a) { return a; }       // unreachable: MAG2 is Floating
  __device__ __forceinline__ static Acc from_count(int64_t n) { return (Acc)n; }
  __device__ __forceinline__ static Acc load(T v) { return (Acc)v; }    // sign-extends, then wraps
  __device__ __forceinline__ static T store(Acc v) { return (T)v; }
  __device__ __forceinline__ static bool is_zero(Acc k) { return k == 0; }
};
template <> struct Arith<uint8_t> : IntArith<uint8_t, uint32_t> {};
template <> struct Arith<int8_t> : IntArith<int8_t, uint32_t> {};
template <> struct Arith<uint16_t> : IntArith<uint16_t, uint32_t> {};
template <> struct Arith<int16_t> : IntArith<int16_t, uint32_t> {};
template <> struct Arith<uint32_t> : IntArith<uint32_t, uint32_t> {};
template <> struct Arith<int32_t> : IntArith<int32_t, uint32_t> {};
template <> struct Arith<uint64_t> : IntArith<uint64_t, uint64_t> {};
template <> struct Arith<int64_t> : IntArith<int64_t, uint64_t> {};

// Running maximum of |x|'s high word: two integer instructions per value; Inf or NaN was seen iff the
// maximum reaches the all-ones exponent
__device__ __forceinline__ void track_magnitude(unsigned &m, float x) { m = max(m, __float_as_uint(x) & 0x7fffffffu); }
__device__ __forceinline__ void track_magnitude(unsigned &m, double x) { m = max(m, (unsigned)__double2hiint(x) & 0x7fffffffu); }
template <typename T> __device__ __forceinline__ bool magnitude_nonfinite(unsigned m) {
  return m >= (sizeof(T) == 4 ? 0x7f800000u : 0x7ff00000u);
}

// Inf or NaN?  (exponent field all ones) — integer pipe only, for the optimistic star kernels
__device__ __forceinline__ unsigned nonfinite_bits(float x) { return (__float_as_uint(x) & 0x7f800000u) == 0x7f800000u; }
__device__ __forceinline__ unsigned nonfinite_bits(double x) {
  return ((unsigned)__double2hiint(x) & 0x7ff00000u) == 0x7ff00000u;
}

// Correctly rounded x / n for a small positive integer n (avgStencil's `/ fromIntegral n`,
// Array/Stencil.hs:479-481) in three operations instead of the ~15-instruction IEEE divide:
//     q = RN(x * y), y = RN(1/n);   r = x - n*q  (exact in one FMA);   q' = RN(q + r*y)
// Why q' == RN(x/n) for every x (n <= 1024): |q - x/n| < 2.0001 ulp(q)*n/n so r is a multiple of
// ulp(q) below 2^12 ulp(q) and the FMA computes it exactly; q + r*y differs from x/n by less than
// 2^-52 ulp(q), while x/n is at least ulp(q)/(4n) away from any rounding boundary (n*(2k+1) has more
// than 53 significant bits, so x/n is never a midpoint).  The FMAs are internal to a sequence whose
// RESULT equals the IEEE quotient; they do not fuse any two reference operations.  r is formed as
// -fma(n, q, -x) so that x = -0 yields -0.  Guards fall back to the IEEE divide: non-finite
// quotients (Inf would make r NaN) and subnormal-range quotients, where an even n can hit an exact
// tie that the correction term would break.  tests/test_gpu_parity.py::test_exact_division checks
// it against __ddiv_rn / __fdiv_rn on 2^28 random and adversarial bit patterns per n.
struct DivCount {
  double n_d, y_d;
  float n_f, y_f;
  int fast;
};
__device__ __forceinline__ double div_count(double x, const DivCount &d) {
  if (!d.fast) return __ddiv_rn(x, d.n_d);
  const double q = __dmul_rn(x, d.y_d);
  const double r = -__fma_rn(d.n_d, q, -x);
  const double q2 = __fma_rn(r, d.y_d, q);
  const double aq = fabs(q);
  if (!(aq < __longlong_as_double(0x7ff0000000000000LL)) || (aq < 8.9e-308 && x != 0.0)) return __ddiv_rn(x, d.n_d);
  return q2;
}
__device__ __forceinline__ float div_count(float x, const DivCount &d) {
  if (!d.fast) return __fdiv_rn(x, d.n_f);
  const float q = __fmul_rn(x, d.y_f);
  const float r = -__fmaf_rn(d.n_f, q, -x);
  const float q2 = __fmaf_rn(r, d.y_f, q);
  const float aq = fabsf(q);
  if (!(aq < __int_as_float(0x7f800000)) || (aq < 4.8e-38f && x != 0.0f)) return __fdiv_rn(x, d.n_f);
  return q2;
}
template <typename T> __device__ __forceinline__ T div_count(T x, const DivCount &) { return x; }  // integers: unreachable

template <typename T> struct Bounds {  // minBound / maxBound
  __device__ __forceinline__ static T lo() { return T(0); }
  __device__ __forceinline__ static T hi() { return T(0); }
};
#define MB_BOUNDS(T, LO, HI)                                   \
  template <> struct Bounds<T> {                               \
    __device__ __forceinline__ static T lo() { return LO; }    \
    __device__ __forceinline__ static T hi() { return HI; }    \
  };
MB_BOUNDS(uint8_t, 0, UINT8_MAX)
MB_BOUNDS(int8_t, INT8_MIN, INT8_MAX)
MB_BOUNDS(uint16_t, 0, UINT16_MAX)
MB_BOUNDS(int16_t, INT16_MIN, INT16_MAX)
MB_BOUNDS(uint32_t, 0u, UINT32_MAX)
MB_BOUNDS(int32_t, INT32_MIN, INT32_MAX)
MB_BOUNDS(uint64_t, 0ull, UINT64_MAX)
MB_BOUNDS(int64_t, INT64_MIN, INT64_MAX)
#undef MB_BOUNDS

// handleBorderIndex for one dimension (Core/Index.hs:236-253).  Returns false when a Fill border
// yields the fill element; otherwise `i` is replaced by the repaired index.  k > 0 for non-Fill
// borders is guaranteed by the host (MB_ERR_INDEX_ZERO otherwise).
__device__ __forceinline__ int64_t floor_mod_dev(int64_t a, int64_t k) {
  int64_t m = a % k;
  return (m < 0) ? m + k : m;
}
__device__ __forceinline__ bool resolve_border(int border, int64_t k, int64_t &i) {
  if (i >= 0 && i < k) return true;
  switch (border) {
    case MB_FILL: return false;
    case MB_WRAP: i = floor_mod_dev(i, k); break;
    case MB_EDGE: i = i < 0 ? 0 : k - 1; break;
    case MB_REFLECT: i = floor_mod_dev(-i - 1, k); break;
    default: i = i < 0 ? floor_mod_dev(-i, k) : floor_mod_dev(-i - 2, k); break;  // MB_CONTINUE
  }
  return true;
}
// 32-bit variant for the tile kernels (extents < 2^31)
__device__ __forceinline__ int floor_mod_dev32(int a, int k) {
  int m = a % k;
  return (m < 0) ? m + k : m;
}
__device__ __forceinline__ bool resolve_border32(int border, int k, int &i) {
  if (i >= 0 && i < k) return true;
  switch (border) {
    case MB_FILL: return false;
    case MB_WRAP: i = floor_mod_dev32(i, k); break;
    case MB_EDGE: i = i < 0 ? 0 : k - 1; break;
    case MB_REFLECT: i = floor_mod_dev32(-i - 1, k); break;
    default: i = i < 0 ? floor_mod_dev32(-i, k) : floor_mod_dev32(-i - 2, k); break;
  }
  return true;
}

}  // namespace mb
