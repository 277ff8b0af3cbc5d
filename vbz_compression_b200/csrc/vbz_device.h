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
#include <string.h>
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
static inline void vbz_prefetch_l2(const void*) {}
static inline uint32_t vbz_ld_acquire_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }
static inline void vbz_st_release_u32(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; }
static inline void vbz_slot_post(uint64_t* slot, uint64_t v) { *(volatile uint64_t*)slot = v; }
static inline uint64_t vbz_slot_read(const uint64_t* slot) { return *(const volatile uint64_t*)slot; }

// packed 16-bit SIMD-in-a-word arithmetic (VIADD.16x2 on sm_100a)
static inline uint32_t vbz_add16x2(uint32_t a, uint32_t b) { return ((a + b) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16); }
static inline uint32_t vbz_sub16x2(uint32_t a, uint32_t b) { return ((a - b) & 0xFFFFu) | (((a >> 16) - (b >> 16)) << 16); }
// full PRMT (selector bit 3 of a nibble replicates the sign of the selected byte; __byte_perm ignores that bit)
static inline uint32_t vbz_prmt(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t ab = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(ab >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) { s = s > 32 ? 32 : s; return s == 32 ? lo : (unsigned)(((((uint64_t)hi << 32) | lo) << s) >> 32); }
// TMA 1-D bulk copies and their mbarrier: the emulator copies synchronously.  The 64-bit barrier word keeps
// the number of completed phases (low half) and the bytes still expected in the current one (high half).
static inline void vbz_mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
static inline void vbz_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    if (bytes) *bar += (uint64_t)bytes << 32;
    else *bar += 1;
}
static inline void vbz_bulk_load(void* s, const void* g, uint32_t n, uint64_t* bar)
{
    memcpy(s, g, n);
    *bar -= (uint64_t)n << 32;
    if ((*bar >> 32) == 0) *bar += 1;
}
static inline void vbz_mbar_wait(uint64_t* bar, uint32_t parity)
{
    while ((*(volatile uint64_t*)bar >> 32) != 0 || ((uint32_t)*(volatile uint64_t*)bar & 1u) == parity) cpusim::yield();
}
static inline void vbz_fence_async_smem() {}
// 8-byte asynchronous global -> shared copies (cp.async): the emulator copies at once
static inline void vbz_cp_async8(void* s, const void* g) { memcpy(s, g, 8); }
static inline void vbz_cp_async_commit() {}
static inline void vbz_cp_async_wait_all() {}
static inline void vbz_bulk_store(void* g, const void* s, uint32_t n) { memcpy(g, s, n); }
static inline void vbz_bulk_store_commit() {}
static inline void vbz_bulk_store_wait_read() {}

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
// hand-off of data through global memory between CTAs: everything written before the release store is visible to whoever sees its value
__device__ __forceinline__ uint32_t vbz_ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void vbz_st_release_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// pull the line that holds p towards the SM ahead of its use (no register, no wait)
__device__ __forceinline__ void vbz_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// hand-off of one 64-bit word (value + ready bit) between warps of a CTA through shared memory: release store, acquire load
// (strong operations at CTA scope: defined under the PTX memory model without a barrier in between)
__device__ __forceinline__ void vbz_slot_post(uint64_t* slot, uint64_t v)
{
    asm volatile("st.release.cta.shared.u64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t vbz_slot_read(const uint64_t* slot)
{
    uint64_t v;
    asm volatile("ld.acquire.cta.shared.u64 %0, [%1];" : "=l"(v) : "r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
    return v;
}
// packed 16-bit SIMD-in-a-word arithmetic (VIADD.16x2 on sm_100a)
__device__ __forceinline__ uint32_t vbz_add16x2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t vbz_sub16x2(uint32_t a, uint32_t b) { return __vsub2(a, b); }
// full PRMT (selector bit 3 of a nibble replicates the sign of the selected byte; __byte_perm ignores that bit)
__device__ __forceinline__ uint32_t vbz_prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}

// ---- TMA 1-D bulk copies (cp.async.bulk) and the mbarrier that tracks a load ------------------------
// Addresses and sizes must be multiples of 16 bytes.
__device__ __forceinline__ uint32_t vbz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vbz_mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vbz_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void vbz_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vbz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vbz_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vbz_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(vbz_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void vbz_mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "VBZ_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra VBZ_DONE;\n\t"
        "bra VBZ_WAIT;\n\t"
        "VBZ_DONE:\n\t"
        "}" ::"r"(vbz_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 8-byte asynchronous global -> shared copy (LDGSTS): no register holds the value while it is in flight
__device__ __forceinline__ void vbz_cp_async8(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(vbz_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void vbz_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void vbz_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// make this thread's shared-memory writes visible to the async (TMA) proxy
__device__ __forceinline__ void vbz_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void vbz_bulk_store(void* gdst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(vbz_smem_u32(smem_src)), "r"(bytes) : "memory");
}
// the issuing thread: close the group / wait until every group's shared-memory source has been read
__device__ __forceinline__ void vbz_bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void vbz_bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
#endif
