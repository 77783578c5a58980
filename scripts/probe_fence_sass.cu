#include <cuda_runtime.h>
__global__ void k_threadfence(int* a, unsigned* t) { a[threadIdx.x] = 1; __threadfence(); if (threadIdx.x == 0) atomicAdd(t, 1u); }
__global__ void k_release(int* a, unsigned* t) { a[threadIdx.x] = 1; unsigned old; asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(t) : "memory"); a[64 + threadIdx.x] = old; }
__global__ void k_acquire(int* a, unsigned* t) { unsigned old; asm volatile("atom.add.acquire.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(t) : "memory"); a[threadIdx.x] = a[old & 63]; }
__global__ void k_strelease(int* a, unsigned* t) { a[threadIdx.x] = 1; asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(t), "r"(5u) : "memory"); }
__global__ void k_ldacquire(int* a, unsigned* t) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(t) : "memory"); a[threadIdx.x] = a[v & 63]; }
__global__ void k_fence_acqrel(int* a, unsigned* t) { a[threadIdx.x] = 1; asm volatile("fence.acq_rel.gpu;" ::: "memory"); if (threadIdx.x == 0) atomicAdd(t, 1u); }
