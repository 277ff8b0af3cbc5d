// csrc/svb_kernels.cuh - stage 1 of VBZ on the GPU: fused (delta + zigzag +) StreamVByte
// pack / unpack, one launch per batch of reads.
//
// Replaces, for every read of a batch at once:
//   encode  StreamVByteWorkerV0<int16,true>::compress   (ref vbz/v0/vbz_streamvbyte_impl_sse3.h:406-466)
//           StreamVByteWorkerV0<T,zz>::compress         (ref vbz/v0/vbz_streamvbyte_impl.h:18-41,
//                                                        streamvbyte_encode.c:14-62, streamvbyte_zigzag.c:13-18)
//   decode  StreamVByteWorkerV0<int16,true>::decompress (ref vbz/v0/vbz_streamvbyte_impl_sse3.h:468-580)
//           StreamVByteWorkerV0<T,zz>::decompress       (ref vbz/v0/vbz_streamvbyte_impl.h:43-81,
//                                                        streamvbyte_decode.c:18-136, streamvbyte_zigzag.c:20-38)
//
// Layout.  A batch is one byte buffer in HBM plus a descriptor per read (SvbRead): where the read
// starts, how many values it has, where its output slot starts.  Reads are cut into TILES of
// 256 threads x 2 groups x 16 raw bytes = 8 KiB of samples; one CTA per tile, tiles numbered
// read-major so that blockIdx order is the order of the data.  Within a tile every thread owns two
// "groups" of 16 raw bytes (8 int16 / 4 int32 / 16 int8 values = 2 / 1 / 4 control bytes), chosen
// strided (group = j*256 + tid) so that every 128-bit global load and store of a warp is one
// fully-coalesced 512-byte access.
//
// The variable-length part: the position of a value's data bytes is the sum of the byte lengths of
// all earlier values of the READ.  Inside the tile that is a warp-shuffle scan + one cross-warp
// step; across the tiles of a read it is a single-pass decoupled look-back (one 64-bit status word
// per tile: 2 flag bits + 62-bit value) - no second pass over the samples.  Packed bytes are staged
// in shared memory at the same 16-byte phase as their global destination, so the copy-out is
// aligned 128-bit stores whatever byte offset the tile's data starts at.  Decode runs two chained
// look-backs: data-byte offsets (known from the control bytes alone, published early) and the
// running sample value (the delta prefix sum).
//
// HBM traffic per int16 sample: 2 B in + 0.25 B keys + ~1.01 B data out (encode) and the reverse
// (decode) = the algorithmic minimum; status words add 8 B per 8 KiB tile (0.1 %).
#pragma once
#include "vbz_device.h"
#include "svb_types.h"

namespace vbz {

// ---- CTA-wide exclusive scan of two independent counters per thread (shared barriers) -----------
// scratch: 2*(NW+1) words.  Two __syncthreads.
template <int NT>
__device__ __forceinline__ void cta_scan_pair(uint32_t a, uint32_t b, uint32_t* scratch, uint32_t& ea, uint32_t& eb,
                                              uint32_t& ta, uint32_t& tb)
{
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t ia = a, ib = b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t xa = __shfl_up_sync(0xffffffffu, ia, d);
        uint32_t xb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) { ia += xa; ib += xb; }
    }
    if (lane == 31) { scratch[wid] = ia; scratch[NW + 1 + wid] = ib; }
    __syncthreads();
    if (wid == 0) {
        uint32_t wa = lane < NW ? scratch[lane] : 0, wb = lane < NW ? scratch[NW + 1 + lane] : 0;
        uint32_t sa = wa, sb = wb;
#pragma unroll
        for (int d = 1; d < NW; d <<= 1) {
            uint32_t xa = __shfl_up_sync(0xffffffffu, sa, d);
            uint32_t xb = __shfl_up_sync(0xffffffffu, sb, d);
            if (lane >= d) { sa += xa; sb += xb; }
        }
        if (lane < NW) { scratch[lane] = sa - wa; scratch[NW + 1 + lane] = sb - wb; }
        if (lane == NW - 1) { scratch[NW] = sa; scratch[2 * NW + 1] = sb; }
    }
    __syncthreads();
    ea = ia - a + scratch[wid];
    eb = ib - b + scratch[NW + 1 + wid];
    ta = scratch[NW];
    tb = scratch[2 * NW + 1];
}

// ---- decoupled look-back over the earlier tiles of the same read (call with all lanes of one warp)
__device__ __forceinline__ uint64_t lookback_exclusive(const uint64_t* status, uint32_t tile, uint32_t first_tile)
{
    const int lane = threadIdx.x & 31;
    uint64_t excl = 0;
    long long j = (long long)tile - 1;
    for (;;) {
        const long long idx = j - lane;
        const bool valid = idx >= (long long)first_tile;
        uint64_t s;
        unsigned inc, upto;
        for (;;) {
            s = valid ? vbz_ld_volatile_u64(status + idx) : kStInclusive;
            const unsigned st = (unsigned)(s >> 62);
            const unsigned inv = __ballot_sync(0xffffffffu, st == 0);
            inc = __ballot_sync(0xffffffffu, st == 2);
            upto = inc ? ((2u << (__ffs((int)inc) - 1)) - 1u) : 0xffffffffu;
            if ((inv & upto) == 0) break;
            VBZ_SPIN_YIELD();
        }
        uint64_t v = ((upto >> lane) & 1u) ? (s & kStValueMask) : 0ull;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        excl += v;
        if (inc) break;
        j -= 32;
    }
    return excl;
}

// ---- shared <-> global copies where both sides sit at the same 16-byte phase ------------------------
template <int NT>
__device__ __forceinline__ void copy_out_aligned(uint8_t* g, const uint8_t* s, uint32_t nbytes)
{
    const uint32_t tid = threadIdx.x;
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > nbytes) head = nbytes;
    if (tid < head) g[tid] = s[tid];
    const uint32_t body = (nbytes - head) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(s + head);
    uint4* g4 = reinterpret_cast<uint4*>(g + head);
    for (uint32_t i = tid; i < body; i += NT) vbz_st_stream(g4 + i, s4[i]);
    const uint32_t done = head + (body << 4);
    if (tid < nbytes - done) g[done + tid] = s[done + tid];
}
template <int NT>
__device__ __forceinline__ void copy_in_aligned(uint8_t* s, const uint8_t* g, uint32_t nbytes)
{
    const uint32_t tid = threadIdx.x;
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > nbytes) head = nbytes;
    if (tid < head) s[tid] = g[tid];
    const uint32_t body = (nbytes - head) >> 4;
    uint4* s4 = reinterpret_cast<uint4*>(s + head);
    const uint4* g4 = reinterpret_cast<const uint4*>(g + head);
    for (uint32_t i = tid; i < body; i += NT) s4[i] = vbz_ld_stream(g4 + i);
    const uint32_t done = head + (body << 4);
    if (tid < nbytes - done) s[done + tid] = g[done + tid];
}

template <int ISZ>
__device__ __forceinline__ int32_t svb_extract(const uint32_t (&w)[4], int k)
{
    if (ISZ == 4) return (int32_t)w[k];
    if (ISZ == 2) return (int32_t)(int16_t)(w[k >> 1] >> (16 * (k & 1)));
    return (int32_t)(int8_t)(w[k >> 2] >> (8 * (k & 3)));
}
template <int ISZ>
__device__ __forceinline__ int32_t svb_load_elem(const uint8_t* p, uint64_t i)
{
    if (ISZ == 4) return *reinterpret_cast<const int32_t*>(p + 4 * i);
    if (ISZ == 2) return *reinterpret_cast<const int16_t*>(p + 2 * i);
    return *reinterpret_cast<const int8_t*>(p + i);
}

// shared memory carve-up (bytes)
constexpr int kSvbScratchBytes = 3 * 2 * (kSvbThreads / 32 + 1) * 4;  // three scan scratch areas
template <int ISZ, int MODE>
struct SvbSmem {
    static constexpr int kTileValues = kSvbTileBytes / ISZ;
    static constexpr int kMaxValueBytes = MODE == kModeI16X86 ? 2 : 4;
    static constexpr int kKeyStage = kTileValues / 4 + 32;
    static constexpr int kDataStage = kTileValues * kMaxValueBytes + 32;
    static constexpr int kScratchOff = 0;
    static constexpr int kBcastOff = (kSvbScratchBytes + 15) & ~15;  // 16 bytes of broadcast slots
    static constexpr int kKeyOff = kBcastOff + 16;
    static constexpr int kDataOff = (kKeyOff + kKeyStage + 15) & ~15;
    static constexpr int kEncodeBytes = kDataOff + kDataStage;
    // decode: a (possibly corrupt) stream may carry 4-byte codes whatever the mode; keys area unused
    static constexpr int kDecodeBytes = kDataOff + kTileValues * 4 + 32;
};

// =====================================================================================================
// encode
// =====================================================================================================
template <int ISZ, bool ZZ, int MODE>
__global__ void __launch_bounds__(kSvbThreads) svb_encode_kernel(SvbParams p)
{
    constexpr int NT = kSvbThreads, G = kSvbGroups, VPG = 16 / ISZ, KB = VPG / 4;
    constexpr uint32_t TV = kSvbTileBytes / ISZ;
    using L = SvbSmem<ISZ, MODE>;
    VBZ_DYN_SMEM(smem);
    uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + L::kScratchOff);
    uint64_t* bcast = reinterpret_cast<uint64_t*>(smem + L::kBcastOff);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t t = blockIdx.x;
    const uint32_t r = p.tile_read[t];
    const SvbRead rd = p.reads[r];
    const uint32_t ti = t - rd.first_tile;
    if (rd.rejected) {
        if (ti == 0 && tid == 0) p.results[r] = rd.preset;
        return;
    }
    const uint32_t n = rd.n_values;
    const uint32_t v0 = ti * TV;
    const uint32_t nv = n - v0 < TV ? n - v0 : TV;
    const uint8_t* src = p.src + rd.src_off;

    // ---- load 16 raw bytes per group, pre-transform, classify -------------------------------------
    uint32_t u[G][VPG];
    uint32_t key[G], len[G], cnt[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t g = j * NT + tid, e0 = g * VPG;
        cnt[j] = e0 >= nv ? 0u : (nv - e0 < (uint32_t)VPG ? nv - e0 : (uint32_t)VPG);
        uint32_t w[4] = {0, 0, 0, 0};
        const uint8_t* gp = src + ((uint64_t)v0 + e0) * ISZ;
        if (cnt[j] == (uint32_t)VPG) {
            const uint4 q = vbz_ld_stream(reinterpret_cast<const uint4*>(gp));
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        } else {
            for (uint32_t b = 0; b < cnt[j] * ISZ; ++b) w[b >> 2] |= (uint32_t)gp[b] << (8 * (b & 3));
        }
        int32_t x[VPG];
#pragma unroll
        for (int k = 0; k < VPG; ++k) x[k] = svb_extract<ISZ>(w, k);
        int32_t prev = 0;
        if (ZZ) {
            prev = __shfl_up_sync(0xffffffffu, x[VPG - 1], 1);
            if (lane == 0) prev = (cnt[j] && (uint64_t)v0 + e0 > 0) ? svb_load_elem<ISZ>(src, (uint64_t)v0 + e0 - 1) : 0;
        }
        key[j] = 0;
        len[j] = 0;
#pragma unroll
        for (int k = 0; k < VPG; ++k) {
            uint32_t uz;
            if (!ZZ) {
                uz = (uint32_t)x[k];
            } else if (MODE == kModeI16X86) {
                const int32_t d = (int32_t)(int16_t)(x[k] - prev);            // 16-bit wrapping delta
                uz = (uint32_t)((d << 1) ^ (d >> 15)) & 0xFFFFu;              // 16-bit zigzag, zero-extended
            } else {
                const uint32_t d = (uint32_t)x[k] - (uint32_t)prev;           // 32-bit wrapping delta
                uz = (d + d) ^ (uint32_t)((int32_t)d >> 31);
            }
            prev = x[k];
            u[j][k] = uz;
            const uint32_t code = MODE == kModeI16X86 ? (uint32_t)(uz > 0xFFu)
                                                      : (uint32_t)(uz > 0xFFu) + (uz > 0xFFFFu) + (uz > 0xFFFFFFu);
            if ((uint32_t)k < cnt[j]) {
                key[j] |= code << (2 * k);
                len[j] += code + 1;
            }
        }
    }

    // ---- where do this tile's data bytes go? -------------------------------------------------------
    uint32_t off0, off1, tot0, tot1;
    cta_scan_pair<NT>(len[0], len[1], scratch, off0, off1, tot0, tot1);
    off1 += tot0;
    const uint32_t tile_total = tot0 + tot1;
    if (tid < 32) {
        uint64_t excl = 0;
        if (ti == 0) {
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStInclusive | tile_total);
        } else {
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStAggregate | tile_total);
            excl = lookback_exclusive(p.status_a, t, rd.first_tile);
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStInclusive | (excl + tile_total));
        }
        if (lane == 0) bcast[0] = excl;
    }
    __syncthreads();
    const uint64_t tile_excl = bcast[0];

    const uint32_t nkeys = (n + 3) / 4;
    uint8_t* out = p.dst + rd.dst_off;
    uint8_t* kdst = out + p.hdr_bytes + (v0 >> 2);
    uint8_t* ddst = out + p.hdr_bytes + nkeys + tile_excl;
    uint8_t* kst = smem + L::kKeyOff + ((uintptr_t)kdst & 15);
    uint8_t* dstg = smem + L::kDataOff + ((uintptr_t)ddst & 15);

    // ---- stage control bytes and data bytes in shared memory ----------------------------------------
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t g = j * NT + tid;
#pragma unroll
        for (int b = 0; b < KB; ++b) kst[g * KB + b] = (uint8_t)(key[j] >> (8 * b));
        uint32_t o = j == 0 ? off0 : off1;
#pragma unroll
        for (int k = 0; k < VPG; ++k) {
            if ((uint32_t)k < cnt[j]) {
                const uint32_t v = u[j][k];
                dstg[o++] = (uint8_t)v;
                if (v > 0xFFu) dstg[o++] = (uint8_t)(v >> 8);
                if (MODE != kModeI16X86) {
                    if (v > 0xFFFFu) dstg[o++] = (uint8_t)(v >> 16);
                    if (v > 0xFFFFFFu) dstg[o++] = (uint8_t)(v >> 24);
                }
            }
        }
    }
    __syncthreads();

    // ---- coalesced copy-out ---------------------------------------------------------------------------
    copy_out_aligned<NT>(kdst, kst, (nv + 3) >> 2);
    copy_out_aligned<NT>(ddst, dstg, tile_total);
    if (tid == 0) {
        if (ti == 0 && p.hdr_bytes) *reinterpret_cast<uint32_t*>(out) = rd.src_bytes;
        if (v0 + nv == n) p.results[r] = (uint32_t)(p.hdr_bytes + nkeys + tile_excl + tile_total);
    }
}

// =====================================================================================================
// decode
// =====================================================================================================
template <int ISZ, bool ZZ, int MODE>
__global__ void __launch_bounds__(kSvbThreads) svb_decode_kernel(SvbParams p)
{
    constexpr int NT = kSvbThreads, G = kSvbGroups, VPG = 16 / ISZ, KB = VPG / 4;
    constexpr uint32_t TV = kSvbTileBytes / ISZ;
    constexpr int NW = NT / 32;
    using L = SvbSmem<ISZ, MODE>;
    VBZ_DYN_SMEM(smem);
    uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + L::kScratchOff);
    uint32_t* scratch2 = scratch + 2 * (NW + 1);
    uint64_t* bcast = reinterpret_cast<uint64_t*>(smem + L::kBcastOff);

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t t = blockIdx.x;
    const uint32_t r = p.tile_read[t];
    const SvbRead rd = p.reads[r];
    const uint32_t ti = t - rd.first_tile;
    const uint32_t n = rd.n_values;
    const uint32_t src_bytes = p.src_sizes ? p.src_sizes[r] : rd.src_bytes;
    const uint32_t nkeys = (n + 3) / 4;
    {   // whole-read rejections: every tile of the read takes the same branch
        uint32_t early = 0;
        bool stop = false;
        if (rd.rejected) { early = rd.preset; stop = true; }
        else if (p.src_sizes && src_bytes >= kErrOutOfMemory) { early = src_bytes; stop = true; }
        else if (n == 0) {  // x86 int16 path returns 0 before looking at the input (impl_sse3.h:473-476);
                            // streamvbyte_validate_stream wants both counts to be zero (streamvbyte_decode.c:91-92)
            early = (MODE == kModeI16X86 || src_bytes == 0) ? 0u : (uint32_t)kErrStream;
            stop = true;
        } else if (nkeys > src_bytes) {  // impl_sse3.h:479-482 vs streamvbyte_decode.c:100-101
            early = MODE == kModeI16X86 ? (uint32_t)kErrInputSize : (uint32_t)kErrStream;
            stop = true;
        }
        if (stop) {
            if (ti == 0 && tid == 0) p.results[r] = early;
            return;
        }
    }
    const uint32_t ndata = src_bytes - nkeys;
    const uint32_t v0 = ti * TV;
    const uint32_t nv = n - v0 < TV ? n - v0 : TV;
    const uint8_t* keys = p.src + rd.src_off + (v0 >> 2);
    const uint8_t* data = p.src + rd.src_off + nkeys;

    // ---- control bytes -> per-group byte counts ------------------------------------------------------------
    uint32_t key[G], len[G], cnt[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t g = j * NT + tid, e0 = g * VPG;
        cnt[j] = e0 >= nv ? 0u : (nv - e0 < (uint32_t)VPG ? nv - e0 : (uint32_t)VPG);
        uint32_t kbits = 0;
        if (cnt[j] == (uint32_t)VPG) {
            if (KB == 4) kbits = *reinterpret_cast<const uint32_t*>(keys + g * 4);
            else if (KB == 2) kbits = *reinterpret_cast<const uint16_t*>(keys + g * 2);
            else kbits = keys[g];
        } else if (cnt[j]) {
            for (uint32_t b = 0; b < (cnt[j] + 3) / 4; ++b) kbits |= (uint32_t)keys[g * KB + b] << (8 * b);
            kbits &= (1u << (2 * cnt[j])) - 1u;  // bits of values past the end are ignored by the reference
        }
        key[j] = kbits;
        len[j] = cnt[j] + __popc(kbits & 0x55555555u) + 2 * __popc(kbits & 0xAAAAAAAAu);
    }
    uint32_t off0, off1, tot0, tot1;
    cta_scan_pair<NT>(len[0], len[1], scratch, off0, off1, tot0, tot1);
    off1 += tot0;
    const uint32_t tile_total = tot0 + tot1;
    if (tid < 32) {
        uint64_t excl = 0;
        if (ti == 0) {
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStInclusive | tile_total);
        } else {
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStAggregate | tile_total);
            excl = lookback_exclusive(p.status_a, t, rd.first_tile);
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStInclusive | (excl + tile_total));
        }
        if (lane == 0) bcast[0] = excl;
    }
    __syncthreads();
    const uint64_t tile_excl = bcast[0];

    // ---- stage this tile's data bytes (clipped to the read's data region; the rest reads as zero) -------
    const uint8_t* gsrc = data + tile_excl;
    uint8_t* stg = smem + L::kDataOff + ((uintptr_t)gsrc & 15);
    const uint32_t avail = tile_excl >= ndata ? 0u : (ndata - tile_excl < tile_total ? (uint32_t)(ndata - tile_excl) : tile_total);
    copy_in_aligned<NT>(stg, gsrc, avail);
    for (uint32_t i = avail + tid; i < tile_total; i += NT) stg[i] = 0;
    __syncthreads();

    // ---- unpack, un-zigzag, prefix-sum ---------------------------------------------------------------------
    uint32_t val[G][VPG];
    uint32_t gsum[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t e0 = (j * NT + tid) * VPG;
        uint32_t o = j == 0 ? off0 : off1;
        // x86 build: a group of 8 goes through the SIMD body (16-bit un-zigzag of the low 16 bits) while it is
        // one of the count/8 full groups and at least 32 data bytes remain; otherwise the scalar tail
        // (32-bit un-zigzag, truncated on store) - impl_sse3.h:494-500,505-520,542-572.
        const bool simd16 = MODE == kModeI16X86 && ((v0 + e0) >> 3) < (n >> 3) && tile_excl + o + 32 <= (uint64_t)ndata;
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < VPG; ++k) {
            uint32_t v = 0;
            if ((uint32_t)k < cnt[j]) {
                const uint32_t code = (key[j] >> (2 * k)) & 3u;
                v = stg[o];
                if (code >= 1) v |= (uint32_t)stg[o + 1] << 8;
                if (code >= 2) v |= (uint32_t)stg[o + 2] << 16;
                if (code == 3) v |= (uint32_t)stg[o + 3] << 24;
                o += code + 1;
            }
            if (ZZ) {
                if (simd16) v &= 0xFFFFu;
                run += (v >> 1) ^ (0u - (v & 1u));  // sums are taken mod 2^32 and truncated to the sample width on store
                val[j][k] = run;
            } else {
                val[j][k] = v;
            }
        }
        gsum[j] = run;
    }
    uint32_t carry0 = 0, carry1 = 0;
    if (ZZ) {
        uint32_t e0s, e1s, t0s, t1s;
        cta_scan_pair<NT>(gsum[0], gsum[1], scratch2, e0s, e1s, t0s, t1s);
        e1s += t0s;
        const uint32_t tile_sum = t0s + t1s;
        if (tid < 32) {
            uint64_t excl = 0;
            if (ti == 0) {
                if (lane == 0) vbz_st_volatile_u64(p.status_b + t, kStInclusive | tile_sum);
            } else {
                if (lane == 0) vbz_st_volatile_u64(p.status_b + t, kStAggregate | tile_sum);
                excl = lookback_exclusive(p.status_b, t, rd.first_tile) & 0xFFFFFFFFull;
                if (lane == 0) vbz_st_volatile_u64(p.status_b + t, kStInclusive | ((excl + tile_sum) & 0xFFFFFFFFull));
            }
            if (lane == 0) bcast[1] = excl;
        }
        __syncthreads();
        const uint32_t base = (uint32_t)bcast[1];
        carry0 = base + e0s;
        carry1 = base + e1s;
    }

    // ---- store samples ----------------------------------------------------------------------------------------
    uint8_t* out = p.dst + rd.dst_off;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t e0 = (j * NT + tid) * VPG;
        const uint32_t c = j == 0 ? carry0 : carry1;
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < VPG; ++k) {
            const uint32_t x = val[j][k] + c;
            if (ISZ == 4) w[k] = x;
            else if (ISZ == 2) w[k >> 1] |= (x & 0xFFFFu) << (16 * (k & 1));
            else w[k >> 2] |= (x & 0xFFu) << (8 * (k & 3));
        }
        uint8_t* gp = out + ((uint64_t)v0 + e0) * ISZ;
        if (cnt[j] == (uint32_t)VPG) {
            vbz_st_stream(reinterpret_cast<uint4*>(gp), make_uint4(w[0], w[1], w[2], w[3]));
        } else {
            for (uint32_t b = 0; b < cnt[j] * ISZ; ++b) gp[b] = (uint8_t)(w[b >> 2] >> (8 * (b & 3)));
        }
    }
    if (tid == 0 && v0 + nv == n) {
        // all data bytes must be consumed exactly: impl_sse3.h:574-577,638-642 / streamvbyte_decode.c:135
        p.results[r] = (tile_excl + tile_total == (uint64_t)ndata) ? n * (uint32_t)ISZ : (uint32_t)kErrStream;
    }
}

// ---- tile table: tile -> read (one warp per read) ------------------------------------------------------------
__global__ void svb_fill_tiles_kernel(const SvbRead* reads, uint32_t n_reads, uint32_t tile_values, uint32_t* tile_read)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_reads) return;
    const SvbRead rd = reads[warp];
    uint32_t nt = (uint32_t)(((uint64_t)rd.n_values + tile_values - 1) / tile_values);
    if (nt == 0) nt = 1;
    for (uint32_t i = lane; i < nt; i += 32) tile_read[rd.first_tile + i] = warp;
}

}  // namespace vbz
