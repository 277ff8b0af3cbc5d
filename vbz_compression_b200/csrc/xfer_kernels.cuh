// csrc/xfer_kernels.cuh - moving a wave of reads between the caller's page-locked HOST buffers and the wave's
// device buffers with ONE kernel launch per direction (the SMs drive the PCIe link through the unified address
// space) instead of one DMA per read.
//
// Why: measured on the B200 box (tools/probe/pcie_probe.cu), every cudaMemcpyAsync costs the copy engine 4 us on top of
// its transfer time and 8-14 us while the link is busy in the other direction - a batch of 430 KB reads moved
// read by read reaches 27 GB/s of the link's 55 GB/s.  A kernel has no per-read cost: a warp owns a 16 KiB chunk of
// a read, keeps 4 KiB of 128-bit loads in flight and writes full 512-byte rows.
//
// Both ends may have any byte alignment (the reference API takes plain pointers, ref vbz/vbz.h:76-130): the copy is
// aligned to the DESTINATION (head and tail bytes go out one by one), and a misaligned source is read with aligned
// 128-bit loads whose neighbours are merged by byte shifts - every source granule crosses the link once (the next
// granule comes from the neighbouring lane by shuffle, not from a second load).  Aligned loads never leave the 16-byte
// granules that hold the read's own bytes, so nothing outside the caller's pages is touched.
//
// No sample arithmetic happens here - bytes are only moved.
#pragma once
#include "vbz_device.h"
#include "svb_types.h"

namespace vbz {

// bytes [m, m+16) of the 32-byte pair (a, b); m in 0..15
__device__ __forceinline__ uint4 xfer_merge(const uint4& a, const uint4& b, uint32_t m)
{
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const uint32_t bits = (m & 3u) * 8u;
    uint4 r;
    switch (m >> 2) {
    case 0: r.x = __funnelshift_r(w[0], w[1], bits); r.y = __funnelshift_r(w[1], w[2], bits); r.z = __funnelshift_r(w[2], w[3], bits); r.w = __funnelshift_r(w[3], w[4], bits); break;
    case 1: r.x = __funnelshift_r(w[1], w[2], bits); r.y = __funnelshift_r(w[2], w[3], bits); r.z = __funnelshift_r(w[3], w[4], bits); r.w = __funnelshift_r(w[4], w[5], bits); break;
    case 2: r.x = __funnelshift_r(w[2], w[3], bits); r.y = __funnelshift_r(w[3], w[4], bits); r.z = __funnelshift_r(w[4], w[5], bits); r.w = __funnelshift_r(w[5], w[6], bits); break;
    default: r.x = __funnelshift_r(w[3], w[4], bits); r.y = __funnelshift_r(w[4], w[5], bits); r.z = __funnelshift_r(w[5], w[6], bits); r.w = __funnelshift_r(w[6], w[7], bits); break;
    }
    return r;
}

__device__ __forceinline__ uint4 xfer_shfl(const uint4& v, int lane)
{
    uint4 r;
    r.x = __shfl_sync(0xFFFFFFFFu, v.x, lane); r.y = __shfl_sync(0xFFFFFFFFu, v.y, lane);
    r.z = __shfl_sync(0xFFFFFFFFu, v.z, lane); r.w = __shfl_sync(0xFFFFFFFFu, v.w, lane);
    return r;
}

__global__ void __launch_bounds__(kXferThreads) xfer_kernel(const XferParams p)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < p.n_chunks; c += warps) {
        // the read this chunk belongs to: last one whose first_chunk <= c
        uint32_t lo = 0, hi = p.n_reads;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.reads[mid].first_chunk <= c) lo = mid; else hi = mid;
        }
        const XferRead r = p.reads[lo];
        uint32_t n = r.n;
        if (p.sizes) {
            const uint32_t s = p.sizes[r.idx];
            n = (s >= 0xFFFFFFF9u || s > r.n) ? 0u : s;
        }
        if (n == 0) continue;
        const uint32_t k = c - r.first_chunk;
        uint8_t* dst = reinterpret_cast<uint8_t*>(r.dst);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(r.src);
        uint32_t head = (16u - (uint32_t)(r.dst & 15u)) & 15u;
        if (head > n) head = n;
        const uint32_t body = (n - head) >> 4;  // 128-bit words of the destination
        const uint32_t tail = (n - head) & 15u;
        if (k == 0) {
            if (lane < head) dst[lane] = src[lane];
            if (lane < tail) dst[head + body * 16u + lane] = src[head + body * 16u + lane];
        }
        const uint32_t j0 = k * kXferChunkVecs;
        if (j0 >= body) continue;
        const uint32_t j1 = min(body, j0 + kXferChunkVecs);
        uint4* out = reinterpret_cast<uint4*>(dst + head);
        const uint64_t sp = r.src + head;
        const uint32_t m = (uint32_t)(sp & 15u);
        const uint4* in = reinterpret_cast<const uint4*>(sp - m);
        if (m == 0) {
            for (uint32_t j = j0; j < j1; j += 32u * kXferRows) {
                uint4 v[kXferRows];
#pragma unroll
                for (uint32_t u = 0; u < kXferRows; ++u) {
                    const uint32_t q = j + u * 32u + lane;
                    if (q < j1) v[u] = in[q];
                }
#pragma unroll
                for (uint32_t u = 0; u < kXferRows; ++u) {
                    const uint32_t q = j + u * 32u + lane;
                    if (q < j1) out[q] = v[u];
                }
            }
        } else {
            for (uint32_t j = j0; j < j1; j += 32u * kXferRows) {
                uint4 v[kXferRows + 1];
#pragma unroll
                for (uint32_t u = 0; u < kXferRows; ++u) {
                    const uint32_t q = j + u * 32u + lane;
                    v[u] = make_uint4(0, 0, 0, 0);
                    if (q <= j1) v[u] = in[q];  // granule j1 holds the last bytes of word j1-1
                }
                v[kXferRows] = make_uint4(0, 0, 0, 0);
                if (lane == 0 && j + kXferRows * 32u <= j1) v[kXferRows] = in[j + kXferRows * 32u];
#pragma unroll
                for (uint32_t u = 0; u < kXferRows; ++u) {
                    const uint32_t q = j + u * 32u + lane;
                    uint4 nx = xfer_shfl(v[u], (int)((lane + 1u) & 31u));
                    const uint4 first_of_next = xfer_shfl(v[u + 1], 0);
                    if (lane == 31u) nx = first_of_next;
                    if (q < j1) out[q] = xfer_merge(v[u], nx, m);
                }
            }
        }
    }
}

// ---- small control-plane moves, by kernel -----------------------------------------------------------------------------
// Descriptor tables go up, per-read results come down and look-back state is cleared by these two kernels rather than
// by cudaMemcpyAsync / cudaMemsetAsync: a copy-engine operation queues behind every DMA already handed to that engine,
// so a 5 KB descriptor upload would wait for the 30 MB wave in front of it and stall the codec (measured: the kernels
// of a wave took 0.9 ms instead of 0.05 ms).  Both pointers 16-byte aligned (pinned staging and device scratch are);
// sizes in 32-bit words.
__global__ void __launch_bounds__(256) words_copy_kernel(uint32_t* dst, const uint32_t* src, uint64_t n_words)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n4 = n_words >> 2;
    for (uint64_t k = i; k < n4; k += stride) reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(src)[k];
    for (uint64_t k = (n4 << 2) + i; k < n_words; k += stride) dst[k] = src[k];
}
__global__ void __launch_bounds__(256) words_zero_kernel(uint32_t* dst, uint64_t n_words)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n4 = n_words >> 2;
    for (uint64_t k = i; k < n4; k += stride) reinterpret_cast<uint4*>(dst)[k] = make_uint4(0, 0, 0, 0);
    for (uint64_t k = (n4 << 2) + i; k < n_words; k += stride) dst[k] = 0;
}

}  // namespace vbz
