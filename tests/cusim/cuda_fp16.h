// cusim: <cuda_fp16.h> on top of the compiler's _Float16 (IEEE binary16, round to nearest even) — TEST INFRASTRUCTURE.
#pragma once
struct __half {
  _Float16 v;
  __half() = default;
  explicit __half(float f) : v((_Float16)f) {}
};
struct alignas(4) __half2 {
  __half x, y;
};
inline __half __float2half_rn(float f) { return __half(f); }
inline __half __float2half(float f) { return __half(f); }
inline float __half2float(__half h) { return (float)h.v; }
inline __half2 __floats2half2_rn(float a, float b) { __half2 r; r.x = __half(a); r.y = __half(b); return r; }
#include "cuda_runtime.h"
inline float2 __half22float2(__half2 h) { return make_float2((float)h.x.v, (float)h.y.v); }
