// csrc/vbz_device.h - what every kernel file includes first.
//
// The kernels are written for sm_100a and compiled by nvcc.  The same source can also be compiled
// as plain C++ with -DVBZ_CPUSIM against tests/cpusim/cpusim.h, a fiber-based lock-step emulator
// used ONLY by the test-suite to debug kernels in the GPU-less build container (it is not part of
// the shipped library, which has no CPU path).  Kernel code therefore sticks to:
//   * dynamic shared memory only (VBZ_DYN_SMEM), no static __shared__;
//   * VBZ_LAUNCH(...) instead of <<<>>>;
//   * VBZ_SPIN_YIELD() inside every spin-wait;
//   * the helpers below for cache-hinted / volatile accesses.
#pragma once
#include <stdint.h>

#ifdef VBZ_CPUSIM
#include "cpusim.h"
#define VBZ_DYN_SMEM(name) unsigned char* name = (unsigned char*)cpusim::g_cur->cta->smem
#define VBZ_SPIN_YIELD() cpusim::yield()
#define VBZ_SPIN_YIELD_SHARED() cpusim::yield()
#define VBZ_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cpusim::launch(cpusim::Dim3(grid), cpusim::Dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })
typedef void* vbz_stream_t;

template <class T> static inline T vbz_ld_stream(const T* p) { return *p; }
template <class T> static inline void vbz_st_stream(T* p, T v) { *p = v; }
static inline uint64_t vbz_ld_volatile_u64(const uint64_t* p) { return *(const volatile uint64_t*)p; }
static inline void vbz_st_volatile_u64(uint64_t* p, uint64_t v) { *(volatile uint64_t*)p = v; }
static inline uint32_t vbz_ld_volatile_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }

#else  // ---------------------------------------------------------------- real CUDA
#include <cuda_runtime.h>
#define VBZ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define VBZ_SPIN_YIELD() __nanosleep(20)
#define VBZ_SPIN_YIELD_SHARED() asm volatile("" ::: "memory")
#define VBZ_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
typedef cudaStream_t vbz_stream_t;

// streaming (evict-first) global accesses for data touched exactly once
template <class T> __device__ __forceinline__ T vbz_ld_stream(const T* p) { return __ldcs(p); }
template <class T> __device__ __forceinline__ void vbz_st_stream(T* p, T v) { __stcs(p, v); }
__device__ __forceinline__ uint64_t vbz_ld_volatile_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void vbz_st_volatile_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t vbz_ld_volatile_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif
