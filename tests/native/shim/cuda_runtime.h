/* TEST INFRASTRUCTURE: stand-in for <cuda_runtime.h> so that the product's device headers (mc_device.cuh) compile with g++
 * for the host replica tests.  Only what those headers use. */
#pragma once
#include <cmath>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline)) inline
#define __constant__ static const
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
