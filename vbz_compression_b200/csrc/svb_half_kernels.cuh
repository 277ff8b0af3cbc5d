// csrc/svb_half_kernels.cuh - stage 1 of VBZ for integer_size 1 with vbz_version 1: the "half" packing.
//
// Replaces StreamVByteWorkerV1<int8_t, zz>::compress / ::decompress and the streamvbyte_*_half helpers
// (ref vbz/v1/vbz_streamvbyte_impl.h:20-301, dispatch vbz/v1/vbz_streamvbyte.cpp:37-65, 82-113):
//   value -> 2-bit code 0/1/2/3 = 0/4/8/16 data BITS (code 3 for every value >= 256; only its low 16 bits are
//   kept), control bytes laid out as in v0, data = a nibble stream, low nibble of a byte first, a value's
//   nibbles least significant first; a trailing half-filled byte counts (impl.h:150-153).
//   pre-transform as the generic v0 worker: int8 widened to int32, 32-bit delta + zigzag or sign extension.
//
// Parallelism: one WARP per read, reads in parallel across the grid.  A nibble stream has no byte-aligned cut
// points, so a read is walked front to back in chunks of 512 values (a lane owns 16 consecutive values = 4
// control bytes = one 128-bit load); the offsets inside a chunk come from a shuffle scan over nibble counts,
// the odd nibble of a chunk is carried into the next one in a register.  int8 data is not the signal path
// (SURVEY.md 8f #4 ranks "at speed" after everything else); this kernel is the parity row, not a tuned one.
#pragma once
#include "svb_kernels.cuh"

namespace vbz {

constexpr int kHalfWarps = 8;                    // warps (reads) per CTA
constexpr uint32_t kHalfChunk = 512;             // values per warp iteration
constexpr uint32_t kHalfStage = 1024 + 16 + 32;  // bytes of staging per warp: 2048 nibbles + the carried one, + copy slack

// value -> code (impl.h:85-103)
__device__ __forceinline__ uint32_t half_code(uint32_t u) { return u == 0 ? 0u : (u < 16u ? 1u : (u < 256u ? 2u : 3u)); }
// nibbles the 16 codes of a 32-bit control word ask for: code c -> (1 << c) >> 1 = b0 + 2 b1 + (b0 & b1)
__device__ __forceinline__ uint32_t half_word_nibbles(uint32_t w)
{
    return __popc(w & 0x55555555u) + 2 * __popc(w & 0xAAAAAAAAu) + __popc(w & (w >> 1) & 0x55555555u);
}

template <bool ZZ>
__global__ void __launch_bounds__(kHalfWarps * 32) svb_half_encode_kernel(SvbParams p)
{
    VBZ_DYN_SMEM(smem);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kHalfWarps + wid;
    if (r >= p.n_reads) return;
    const SvbRead rd = p.reads[r];
    if (rd.rejected) {
        if (lane == 0) p.results[r] = rd.preset;
        return;
    }
    uint8_t* stage = smem + wid * kHalfStage;
    uint32_t* stage32 = reinterpret_cast<uint32_t*>(stage);
    const uint32_t n = rd.n_values, nkeys = (n + 3) / 4;
    const uint8_t* src = p.src + rd.src_off;
    uint8_t* out = p.dst + rd.dst_off;
    uint8_t* keys = out + p.hdr_bytes;
    uint8_t* data = keys + nkeys;
    if (lane == 0 && p.hdr_bytes) *reinterpret_cast<uint32_t*>(out) = rd.src_bytes;
    uint64_t written = 0;      // data bytes already in global memory
    uint32_t carry = 0;        // 1: a nibble is waiting for its other half
    uint32_t carry_nib = 0;
    int32_t prev_chunk = 0;    // last sample of the previous chunk (zigzag)
    for (uint32_t c0 = 0; c0 < n; c0 += kHalfChunk) {
        const uint32_t v0 = c0 + lane * 16;
        const uint32_t cnt = v0 >= n ? 0u : (n - v0 < 16u ? n - v0 : 16u);
        uint32_t w[4] = {0, 0, 0, 0};
        if (cnt == 16) {
            const uint4 q = vbz_ld_stream(reinterpret_cast<const uint4*>(src + v0));
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        } else {
            for (uint32_t b = 0; b < cnt; ++b) w[b >> 2] |= (uint32_t)src[v0 + b] << (8 * (b & 3));
        }
        int32_t prev = 0;
        if (ZZ) {
            const int32_t last = (int32_t)(int8_t)(w[3] >> 24);  // only meaningful in full lanes, which all lanes before a non-empty one are
            prev = __shfl_up_sync(0xffffffffu, last, 1);
            if (lane == 0) prev = prev_chunk;
            prev_chunk = __shfl_sync(0xffffffffu, last, 31);
        }
        // zero the staging area (the carried nibble goes back in afterwards)
        for (uint32_t i = lane; i < kHalfStage / 4; i += 32) stage32[i] = 0;
        __syncwarp();
        uint32_t u[16], key = 0, nibs = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int32_t x = (int32_t)(int8_t)(w[k >> 2] >> (8 * (k & 3)));
            uint32_t uz;
            if (ZZ) {
                const uint32_t d = (uint32_t)x - (uint32_t)prev;  // streamvbyte_zigzag.c:13-18 on the widened values
                uz = (d + d) ^ (uint32_t)((int32_t)d >> 31);
                prev = x;
            } else {
                uz = (uint32_t)x;
            }
            const uint32_t code = half_code(uz);
            u[k] = uz & 0xFFFFu;
            if ((uint32_t)k < cnt) {
                key |= code << (2 * k);
                nibs += (1u << code) >> 1;
            }
        }
        // control bytes: 4 per lane, at a 4-byte aligned address (slots are 16-byte aligned, header 0 or 4 bytes)
        if (cnt == 16) *reinterpret_cast<uint32_t*>(keys + (v0 >> 2)) = key;
        else for (uint32_t b = 0; b < (cnt + 3) / 4; ++b) keys[(v0 >> 2) + b] = (uint8_t)(key >> (8 * b));
        // nibble offsets inside the chunk
        const uint32_t incl = warp_scan_u32(nibs);
        const uint32_t total = carry + __shfl_sync(0xffffffffu, incl, 31);
        uint32_t pos = carry + incl - nibs;
        if (lane == 0 && carry) stage32[0] = carry_nib;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if ((uint32_t)k >= cnt) break;
            const uint32_t code = (key >> (2 * k)) & 3u;
            if (code == 0) continue;
            const uint32_t bits = 2u << code;  // 4, 8, 16
            const uint32_t v = u[k] & ((1u << bits) - 1u);
            const uint32_t word = pos >> 3, sh = 4u * (pos & 7u);
            atomicOr(stage32 + word, v << sh);
            if (sh + bits > 32u) atomicOr(stage32 + word + 1, v >> (32u - sh));
            pos += bits >> 2;
        }
        __syncwarp();
        // whole bytes leave; an odd nibble waits for the next chunk
        const uint32_t nbytes = total >> 1;
        warp_copy_out_shifted(data + written, stage, nbytes);
        carry = total & 1u;
        carry_nib = carry ? (uint32_t)(stage[nbytes] & 0xFu) : 0u;
        written += nbytes;
        __syncwarp();
    }
    if (lane == 0) {
        if (carry) data[written] = (uint8_t)carry_nib;  // impl.h:150-153: the half-filled byte counts
        p.results[r] = (uint32_t)(p.hdr_bytes + nkeys + written + carry);
    }
}

template <bool ZZ>
__global__ void __launch_bounds__(kHalfWarps * 32) svb_half_decode_kernel(SvbParams p)
{
    VBZ_DYN_SMEM(smem);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kHalfWarps + wid;
    if (r >= p.n_reads) return;
    const SvbRead rd = p.reads[r];
    const uint32_t n = rd.n_values;
    const uint32_t src_bytes = p.src_sizes ? p.src_sizes[r] : rd.src_bytes;
    const uint32_t nkeys = (n + 3) / 4;
    uint32_t early = 0;
    bool stop = false;
    if (p.src_sizes && src_bytes >= kErrOutOfMemory) { early = src_bytes; stop = true; }  // the zstd stage failed first
    else if (rd.rejected) { early = rd.preset; stop = true; }
    else if (n == 0 || src_bytes == 0) { early = (n == 0 && src_bytes == 0) ? 0u : (uint32_t)kErrStream; stop = true; }  // impl.h:186-189
    else if (nkeys > src_bytes) { early = kErrStream; stop = true; }                                                    // impl.h:194-197
    if (stop) {
        if (lane == 0) p.results[r] = early;
        return;
    }
    const uint8_t* keys = p.src + rd.src_off;  // 4-byte aligned
    const uint8_t* data = keys + nkeys;
    const uint32_t ndata = src_bytes - nkeys;
    // ---- validate (impl.h:183-216): the control bytes must ask for exactly the nibbles that are there -------
    {
        uint64_t sum = 0;
        for (uint32_t b = lane * 4; b < nkeys; b += 128) {
            uint32_t word;
            if (b + 4 <= nkeys) word = *reinterpret_cast<const uint32_t*>(keys + b);
            else {
                word = 0;
                for (uint32_t i = b; i < nkeys; ++i) word |= (uint32_t)keys[i] << (8 * (i - b));
            }
            const uint32_t vleft = n - 4 * b;
            if (vleft < 16) word &= (1u << (2 * vleft)) - 1u;
            sum += half_word_nibbles(word);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if ((sum + 1) / 2 != (uint64_t)ndata) {
            if (lane == 0) p.results[r] = kErrStream;
            return;
        }
    }
    uint8_t* stage_base = smem + wid * kHalfStage;
    uint8_t* out = p.dst + rd.dst_off;
    uint64_t nib_base = 0;  // nibbles consumed by earlier chunks
    uint32_t run_chunk = 0; // running sample value (zigzag)
    for (uint32_t c0 = 0; c0 < n; c0 += kHalfChunk) {
        const uint32_t v0 = c0 + lane * 16;
        const uint32_t cnt = v0 >= n ? 0u : (n - v0 < 16u ? n - v0 : 16u);
        uint32_t key = 0;
        if (cnt == 16) key = *reinterpret_cast<const uint32_t*>(keys + (v0 >> 2));
        else if (cnt) {
            for (uint32_t b = 0; b < (cnt + 3) / 4; ++b) key |= (uint32_t)keys[(v0 >> 2) + b] << (8 * b);
            key &= (1u << (2 * cnt)) - 1u;
        }
        const uint32_t nibs = half_word_nibbles(key);
        const uint32_t incl = warp_scan_u32(nibs);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        // stage the bytes that hold this chunk's nibbles
        const uint64_t first_byte = nib_base >> 1;
        const uint32_t nbytes = (uint32_t)(((nib_base + total + 1) >> 1) - first_byte);
        const uint8_t* g = data + first_byte;
        uint8_t* stage = stage_base + ((uintptr_t)g & 15u);
        warp_copy_in(stage, g, nbytes);
        if (lane < 4) stage[nbytes + lane] = 0;
        __syncwarp();
        uint32_t pos = (uint32_t)(nib_base & 1u) + incl - nibs;
        uint32_t val[16], run = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t v = 0;
            if ((uint32_t)k < cnt) {
                const uint32_t code = (key >> (2 * k)) & 3u;
                if (code) {
                    const uint32_t bits = 2u << code, b = pos >> 1;
                    const uint32_t raw = (uint32_t)stage[b] | ((uint32_t)stage[b + 1] << 8) | ((uint32_t)stage[b + 2] << 16);
                    v = (raw >> (4u * (pos & 1u))) & ((1u << bits) - 1u);
                    pos += bits >> 2;
                }
            }
            if (ZZ) {
                run += (v >> 1) ^ (0u - (v & 1u));  // streamvbyte_zigzag.c:20-38; truncated to the sample width on store
                val[k] = run;
            } else {
                val[k] = v;
            }
        }
        uint32_t base = 0;
        if (ZZ) {
            const uint32_t si = warp_scan_u32(run);
            base = run_chunk + si - run;
            run_chunk += __shfl_sync(0xffffffffu, si, 31);
        }
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; ++k) w[k >> 2] |= ((val[k] + base) & 0xFFu) << (8 * (k & 3));
        if (cnt == 16) vbz_st_stream(reinterpret_cast<uint4*>(out + v0), make_uint4(w[0], w[1], w[2], w[3]));
        else for (uint32_t b = 0; b < cnt; ++b) out[v0 + b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
        nib_base += total;
        __syncwarp();
    }
    if (lane == 0) p.results[r] = n;
}

}  // namespace vbz
