// Packed fp32x2 arithmetic helpers.  On sm_100a these map to FADD2 / FMUL2 / FFMA2 (one issue slot
// for two fp32 lanes); on the host (unit tests of the generated FFT code) they are emulated.
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define AA_HD __host__ __device__ __forceinline__
#else
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#define AA_HD static inline
#define __device__
#define __forceinline__ inline
#endif

AA_HD float2 padd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
AA_HD float2 psub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fadd2_rn(a, make_float2(-b.x, -b.y));  // FADD2 with a negated operand
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
AA_HD float2 pmul(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
  return __fmul2_rn(a, b);
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
// a * c + b with a scalar (broadcast) multiplier
AA_HD float2 pfma(float2 a, float c, float2 b) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(a, make_float2(c, c), b);
#else
  return make_float2(fmaf(a.x, c, b.x), fmaf(a.y, c, b.y));
#endif
}
AA_HD float2 pfma2(float2 a, float2 c, float2 b) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(a, c, b);
#else
  return make_float2(fmaf(a.x, c.x, b.x), fmaf(a.y, c.y, b.y));
#endif
}
// a * c with a scalar (broadcast) multiplier
AA_HD float2 pmuls(float2 a, float c) {
#if defined(__CUDA_ARCH__)
  return __fmul2_rn(a, make_float2(c, c));
#else
  return make_float2(a.x * c, a.y * c);
#endif
}
