// Host shims for the arithmetic headers (mas_math.cuh, kspace_ops.cuh): when one of them is compiled as plain
// C++ by tests/hostcheck/ (never by the product build) the CUDA qualifiers vanish and the round-to-nearest
// intrinsics become the plain operators -- identical results as long as the host compiler does not contract
// into FMAs (-ffp-contract=off).  Under nvcc this header is empty.
#pragma once
#if !defined(__CUDACC__)
#include <math.h>
#include <stddef.h>
#define __device__
#define __forceinline__ inline
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  const unsigned long long old = *p;
  *p = old + v;
  return old;
}
struct float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float atomicAdd(float* p, float v) {   // the host check runs one particle at a time
  const float old = *p;
  *p = old + v;
  return old;
}
#endif
