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
