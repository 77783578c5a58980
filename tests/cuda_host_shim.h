// cuda_host_shim.h — TEST-ONLY stand-ins for the CUDA device intrinsics used by the search headers of slam3d_b200/csrc
// (common.cuh, nn_search.cuh, knn_walk.cuh), so that g++ can compile them for tests/hostsearch.cpp: one host thread then
// plays one device thread.  Build with -ffp-contract=off (no FMA contraction, like -fmad=false on the device); the float
// operations below are then the same IEEE round-to-nearest operations as the __f*_rn intrinsics.
#pragma once

#include <cuda_runtime.h>  // vector types (float4, uint4, ...) and the empty host definitions of __device__ / __forceinline__

#include <cmath>
#include <cstdint>
#include <cstring>

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
using std::isfinite;

// one host thread = thread 0 of a one-thread block
static const struct { unsigned x, y, z; } threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
