/* Test infrastructure: lets g++ compile a one-thread-per-robot CUDA kernel source as plain host C++ so that its LOGIC can be
 * compared with the compiled reference in the build container (no GPU there).  Never part of the product library. */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#undef __device__
#undef __global__
#undef __forceinline__
#undef __launch_bounds__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
inline void __syncthreads() {}
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
const uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
const dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);
using std::fmin; using std::fmax; using std::fabs; using std::sqrt; using std::sin; using std::cos; using std::atan2; using std::acos;
using std::asin; using std::fmod; using std::floor; using std::ceil;
