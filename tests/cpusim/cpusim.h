// tests/cpusim/cpusim.h - TEST INFRASTRUCTURE ONLY.
//
// A small lock-step emulator that runs the product's CUDA kernel SOURCE (csrc/*.cuh, compiled as
// plain C++ with -DVBZ_CPUSIM) on the CPU so kernels can be debugged in the build container,
// which has no GPU.  One user-level fiber per CUDA thread; fibers of up to `concurrent_ctas` CTAs
// are interleaved round-robin and switch at every barrier / warp collective / spin-wait, so
// decoupled look-back between CTAs is exercised too.  It is never linked into libvbz_b200.so and
// is not a fallback: the shipped library has no CPU path.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>

namespace cpusim {

struct Dim3 { unsigned x = 1, y = 1, z = 1; Dim3() {} Dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

struct Warp;
struct Cta;
struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    Dim3 tid, bid;
    Cta* cta = nullptr;
    Warp* warp = nullptr;
    unsigned lane = 0;
    bool done = false;
};
struct Warp {
    unsigned nlanes = 0;
    unsigned arrived = 0, gen = 0;
    uint64_t slot[32];
    unsigned ballot = 0;
};
struct Cta {
    unsigned nthreads = 0, arrived = 0, gen = 0;
    char* smem = nullptr;
    Dim3 bid;
};

extern Fiber* g_cur;
extern Dim3 g_block_dim, g_grid_dim;
extern unsigned g_concurrent_ctas;  // how many CTAs are interleaved (default 3)
extern unsigned long long g_switches;

void yield();
void launch(Dim3 grid, Dim3 block, size_t smem_bytes, const std::function<void()>& body);

inline void warp_barrier()
{
    Warp* w = g_cur->warp;
    unsigned my = w->gen;
    if (++w->arrived == w->nlanes) { w->arrived = 0; w->gen++; }
    else while (w->gen == my) yield();
}
inline void cta_barrier()
{
    Cta* c = g_cur->cta;
    unsigned my = c->gen;
    if (++c->arrived == c->nthreads) { c->arrived = 0; c->gen++; }
    else while (c->gen == my) yield();
}
// exchange: every lane deposits, then reads lane `src`
template <class T>
inline T warp_exchange(T v, unsigned src)
{
    static_assert(sizeof(T) <= 8, "");
    Warp* w = g_cur->warp;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w->slot[g_cur->lane] = raw;
    warp_barrier();
    uint64_t got = w->slot[src < w->nlanes ? src : g_cur->lane];
    warp_barrier();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
inline unsigned warp_ballot(bool pred)
{
    Warp* w = g_cur->warp;
    w->slot[g_cur->lane] = pred ? 1 : 0;
    warp_barrier();
    unsigned m = 0;
    for (unsigned i = 0; i < w->nlanes; ++i) m |= (unsigned)(w->slot[i] & 1) << i;
    warp_barrier();
    return m;
}

}  // namespace cpusim

// ---- the CUDA surface the kernels use ---------------------------------------------------
#define __global__
#define __device__ inline
#define __host__
#define __forceinline__
#define __noinline__
#define __launch_bounds__(...)
#define __constant__ static const
#ifndef __restrict__
#define __restrict__ __restrict
#endif

#define threadIdx (cpusim::g_cur->tid)
#define blockIdx (cpusim::g_cur->bid)
#define blockDim (cpusim::g_block_dim)
#define gridDim (cpusim::g_grid_dim)

struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

static inline void __syncthreads() { cpusim::cta_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cpusim::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) { cpusim::yield(); }

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    unsigned lane = cpusim::g_cur->lane, base = lane & ~(unsigned)(width - 1);
    return cpusim::warp_exchange(v, base + ((unsigned)src & (unsigned)(width - 1)));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32)
{
    unsigned lane = cpusim::g_cur->lane;
    return cpusim::warp_exchange(v, lane >= d ? lane - d : lane);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32)
{
    unsigned lane = cpusim::g_cur->lane;
    return cpusim::warp_exchange(v, lane + d < 32 ? lane + d : lane);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32)
{
    return cpusim::warp_exchange(v, cpusim::g_cur->lane ^ (unsigned)m);
}
static inline unsigned __ballot_sync(unsigned, int p) { return cpusim::warp_ballot(p != 0); }
static inline int __any_sync(unsigned, int p) { return cpusim::warp_ballot(p != 0) != 0; }
static inline int __all_sync(unsigned, int p) { return cpusim::warp_ballot(p == 0) == 0; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
    unsigned s = 0;
    for (unsigned i = 0; i < cpusim::g_cur->warp->nlanes; ++i) s += cpusim::warp_exchange(v, i);
    return s;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
    unsigned s = 0;
    for (unsigned i = 0; i < cpusim::g_cur->warp->nlanes; ++i) s |= cpusim::warp_exchange(v, i);
    return s;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
    unsigned s = 0;
    for (unsigned i = 0; i < cpusim::g_cur->warp->nlanes; ++i) { const unsigned x = cpusim::warp_exchange(v, i); if (x > s) s = x; }
    return s;
}
static inline unsigned __match_any_sync(unsigned, int v)
{
    unsigned m = 0;
    for (unsigned i = 0; i < cpusim::g_cur->warp->nlanes; ++i) if (cpusim::warp_exchange((unsigned)v, i) == (unsigned)v) m |= 1u << i;
    return m;
}
static inline unsigned __activemask() { return cpusim::g_cur->warp->nlanes == 32 ? 0xffffffffu : ((1u << cpusim::g_cur->warp->nlanes) - 1); }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(v);
}
// like the CUDA intrinsic, only 3 bits per selector nibble are honoured (no sign replication)
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    uint64_t ab = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned byte = (unsigned)(ab >> (8 * (sel & 7))) & 0xFF;
        r |= byte << (8 * i);
    }
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { s = s > 32 ? 32 : s; return s == 32 ? hi : (unsigned)((((uint64_t)hi << 32) | lo) >> s); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

template <class T, class U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> static inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> static inline T atomicCAS(T* p, U c, V v) { T o = *p; if (o == (T)c) *p = (T)v; return o; }
