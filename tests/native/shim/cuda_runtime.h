// TEST INFRASTRUCTURE (never part of the product): a stand-in for <cuda_runtime.h> that lets g++ compile the search
// kernels of erweitern_55_b200/csrc/e55_gpu_selfplay.cuh for the HOST, one "thread" at a time, so that they can run
// under AddressSanitizer / UndefinedBehaviorSanitizer (compute-sanitizer is closed on the GPU pool) and be fuzzed with a
// pseudo-network on a machine without a GPU.  tests/native/gs_host.cpp is the driver.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __host__
#define __device__
#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__

struct gs_dim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local gs_dim3 blockIdx, threadIdx;
inline thread_local gs_dim3 blockDim{1, 1, 1}, gridDim{1, 1, 1};

struct uint4 { unsigned x, y, z, w; };

inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p += v; return o; }
inline void __threadfence() {}
inline void __syncthreads() {}
// IEEE single-precision operations in the order written (build with -ffp-contract=off, no -ffast-math)
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __expf(float x) { return std::exp(x); }
inline float __shfl_xor_sync(unsigned, float v, int) { return v; }   // (only in kernels the host driver never runs)
