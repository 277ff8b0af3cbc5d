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
// starts, how many values it has, where its output slot starts.  Reads are cut into TILES of 8 KiB
// of samples, one CTA (8 warps) per tile, tiles numbered read-major so that blockIdx order is the
// order of the data.  Each warp owns a 1 KiB WARP TILE; a lane owns two "groups" of 16 raw bytes
// (8 int16 / 4 int32 / 16 int8 values = 2 / 1 / 4 control bytes), strided (group = j*32 + lane) so
// that every 128-bit global load / store of a warp is one fully-coalesced 512-byte access.
//
// The variable-length part: the position of a value's data bytes is the sum of the byte lengths of
// all earlier values of the READ.  Three levels, no CTA-wide barrier after the prologue:
//   lane   -> shuffle scan inside the warp;
//   warp   -> each warp posts its total in a shared-memory slot (value + ready bit in one word) and
//             sums the slots of the warps before it, spinning only on those;
//   tile   -> the last warp posts the CTA total as the tile's AGGREGATE in a 64-bit global status
//             word (2 flag bits + 62-bit value) and runs a decoupled look-back, 32 predecessors at a
//             time, over the earlier tiles of the same read - single pass, no second read of the data.
// The dependency is pushed to the end of the tile's work: encode stages its packed bytes in shared
// memory at tile-local offsets BEFORE the global offset is known and only the final copy-out
// (which re-aligns with funnel shifts to aligned 128-bit stores) waits for it; decode gets its
// data-byte offsets from a tiny pre-pass over the control bytes alone (svb_key_offsets_kernel, 1/8
// of the bytes) and its running sample value (the delta prefix sum) from the look-back, needed
// only for the final add + store.
//
// HBM traffic per int16 sample: 2 B in + 0.25 B keys + ~1.01 B data out (encode) and the reverse
// (decode, + 0.25 B for the control-byte pre-pass) = the algorithmic minimum + 8 %.
#pragma once
#include "vbz_device.h"
#include "svb_types.h"

namespace vbz {

enum : uint64_t { kSlotReady = 1ull << 63 };

// ---- decoupled look-back over the earlier tiles of the same read (all 32 lanes of a warp) ----------
// `pf`: this lane's status word of the first window (status[tile - 1 - lane]) if the caller loaded it earlier
// (have_pf); any valid status, however old, gives the right answer because aggregates only turn into inclusives.
__device__ __forceinline__ uint64_t lookback_exclusive_pf(const uint64_t* status, uint32_t tile, uint32_t first_tile, uint64_t pf, bool have_pf)
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
            s = valid ? (have_pf ? pf : vbz_ld_volatile_u64(status + idx)) : kStInclusive;
            have_pf = false;
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
// The same walk with the first kLookWindows windows (32 tiles each) already sitting in shared memory: win[w * 32 + lane] holds
// status[tile - 1 - 32 w - lane], or 0 where that index lies before first_tile or the word was still empty when it was fetched.
// Inside one read nobody can post an INCLUSIVE prefix before the tile in front of it has resolved, so the walk of tile number
// ti of its read covers ti tiles: fetched one window after the other that was up to four dependent global round trips on
// the critical path of every tile of a long read; fetched together, ahead of time, it is none (and one for a window that
// was not complete yet).
constexpr int kLookWindows = 4;
__device__ __forceinline__ uint64_t lookback_exclusive_win(const uint64_t* status, uint32_t tile, uint32_t first_tile, const uint64_t* win)
{
    const int lane = threadIdx.x & 31;
    uint64_t excl = 0;
    long long j = (long long)tile - 1;
    for (int w = 0;; ++w) {
        const long long idx = j - lane;
        const bool valid = idx >= (long long)first_tile;
        bool have = w < kLookWindows;
        uint64_t s = 0;
        unsigned inc, upto;
        for (;;) {
            s = valid ? (have ? win[w * 32 + lane] : vbz_ld_volatile_u64(status + idx)) : kStInclusive;
            have = false;
            const unsigned st = (unsigned)(s >> 62);
            const unsigned inv = __ballot_sync(0xffffffffu, st == 0);
            inc = __ballot_sync(0xffffffffu, st == 2);
            upto = inc ? ((2u << (__ffs((int)inc) - 1)) - 1u) : 0xffffffffu;
            if ((inv & upto) == 0) break;
            VBZ_SPIN_YIELD();
        }
        // at most one INCLUSIVE value (< 2^62) and 31 tile aggregates: two hardware reductions instead of a 64-bit shuffle tree
        const uint64_t v = ((upto >> lane) & 1u) ? (s & kStValueMask) : 0ull;
        const uint64_t lo = __reduce_add_sync(0xffffffffu, (uint32_t)(v & 0xFFFFFFu)) , mid = __reduce_add_sync(0xffffffffu, (uint32_t)((v >> 24) & 0xFFFFFFu)),
                       hi = __reduce_add_sync(0xffffffffu, (uint32_t)(v >> 48));
        excl += lo + (mid << 24) + (hi << 48);
        if (inc) break;
        j -= 32;
    }
    return excl;
}
__device__ __forceinline__ uint64_t lookback_exclusive(const uint64_t* status, uint32_t tile, uint32_t first_tile)
{
    return lookback_exclusive_pf(status, tile, first_tile, 0ull, false);
}

// Level "warp -> tile -> read" of the offset computation (see the header comment).  Every warp of
// the CTA calls it once per chain with its own total; returns the exclusive prefix of this warp
// within the tile (*warp_excl) and posts, through `cta_slot`, the exclusive prefix of the tile
// within the read (call chain_wait() to read it when it is finally needed).
// slots: kSvbWarps words of shared memory, zeroed in the prologue.  mask trims the running value.
__device__ __forceinline__ void chain_post(uint64_t* slots, uint64_t* cta_slot, uint64_t* status, uint32_t tile,
                                           uint32_t first_tile, uint64_t warp_total, uint64_t mask, uint64_t* warp_excl)
{
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) vbz_slot_post(slots + wid, kSlotReady | (warp_total & mask));
    uint64_t v = 0;
    if (wid > 0) {
        for (;;) {
            v = lane < wid ? vbz_slot_read(slots + lane) : kSlotReady;
            if (__all_sync(0xffffffffu, (v & kSlotReady) != 0)) break;
            VBZ_SPIN_YIELD_SHARED();
        }
        v &= ~kSlotReady;
#pragma unroll
        for (int d = 4; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);  // only lanes < 8 carry values
        v = __shfl_sync(0xffffffffu, v, 0) & mask;
    }
    *warp_excl = v;
    if (wid == kSvbWarps - 1) {  // the last warp knows the tile total first
        const uint64_t aggregate = (v + warp_total) & mask;
        uint64_t excl = 0;
        if (tile == first_tile) {
            if (lane == 0) vbz_st_volatile_u64(status + tile, kStInclusive | aggregate);
        } else {
            if (lane == 0) vbz_st_volatile_u64(status + tile, kStAggregate | aggregate);
            excl = lookback_exclusive(status, tile, first_tile) & mask;
            if (lane == 0) vbz_st_volatile_u64(status + tile, kStInclusive | ((excl + aggregate) & mask));
        }
        if (lane == 0) vbz_slot_post(cta_slot, kSlotReady | excl);
    }
}
__device__ __forceinline__ uint64_t chain_wait(const uint64_t* cta_slot)
{
    uint64_t v;
    for (;;) {
        v = vbz_slot_read(cta_slot);
        if (v & kSlotReady) break;
        VBZ_SPIN_YIELD_SHARED();
    }
    return v & ~kSlotReady;
}

// ---- warp-wide shared -> global copy, same 16-byte phase on both sides -----------------------------------
__device__ __forceinline__ void warp_copy_out(uint8_t* g, const uint8_t* s, uint32_t nbytes)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > nbytes) head = nbytes;
    if (lane < head) g[lane] = s[lane];
    const uint32_t body = (nbytes - head) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(s + head);
    uint4* g4 = reinterpret_cast<uint4*>(g + head);
    for (uint32_t i = lane; i < body; i += 32) vbz_st_stream(g4 + i, s4[i]);
    const uint32_t done = head + (body << 4);
    if (lane < nbytes - done) g[done + lane] = s[done + lane];
}
// ---- warp-wide shared -> global copy, arbitrary phase on both sides: aligned 128-bit global stores fed
// by 32-bit shared loads re-aligned with funnel shifts (the shared buffer must be readable 8 bytes past
// the end of the range) -------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_copy_out_shifted(uint8_t* g, const uint8_t* s, uint32_t nbytes)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > nbytes) head = nbytes;
    if (lane < head) g[lane] = s[lane];
    const uint32_t body = (nbytes - head) >> 4;
    const uint8_t* sb = s + head;
    const uint32_t sh = 8u * (uint32_t)((uintptr_t)sb & 3u);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(sb - ((uintptr_t)sb & 3u));
    uint4* g4 = reinterpret_cast<uint4*>(g + head);
    for (uint32_t i = lane; i < body; i += 32) {
        const uint32_t* q = sw + 4 * i;
        const uint32_t a = q[0], b = q[1], c = q[2], d = q[3], e = q[4];
        vbz_st_stream(g4 + i, make_uint4(__funnelshift_r(a, b, sh), __funnelshift_r(b, c, sh), __funnelshift_r(c, d, sh), __funnelshift_r(d, e, sh)));
    }
    const uint32_t done = head + (body << 4);
    if (lane < nbytes - done) g[done + lane] = s[done + lane];
}
__device__ __forceinline__ void warp_copy_in(uint8_t* s, const uint8_t* g, uint32_t nbytes)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > nbytes) head = nbytes;
    if (lane < head) s[lane] = g[lane];
    const uint32_t body = (nbytes - head) >> 4;
    uint4* s4 = reinterpret_cast<uint4*>(s + head);
    const uint4* g4 = reinterpret_cast<const uint4*>(g + head);
    for (uint32_t i = lane; i < body; i += 32) s4[i] = vbz_ld_stream(g4 + i);
    const uint32_t done = head + (body << 4);
    if (lane < nbytes - done) s[done + lane] = g[done + lane];
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

// inclusive warp scan (also used on two 16-bit counters packed in one word, each total < 65536)
__device__ __forceinline__ uint32_t warp_scan_u32(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += x;
    }
    return v;
}

// shared memory carve-up (bytes)
constexpr int kSvbCtrlBytes = (2 * kSvbWarps + 2) * 8;  // warp slots of both chains + the two tile slots
template <int ISZ, int MODE>
struct SvbSmem {
    static constexpr int kWarpValues = kSvbWarpTileBytes / ISZ;
    static constexpr int kTileValues = kSvbTileBytes / ISZ;
    static constexpr int kMaxValueBytes = MODE == kModeI16X86 ? 2 : 4;
    static constexpr int kCtrlOff = 0;
    // encode: per-warp control-byte staging, then one tile-wide data staging area at tile-local offsets
    static constexpr int kKeyStageWarp = (kWarpValues / 4 + 16 + 15) & ~15;
    static constexpr int kKeyOff = (kSvbCtrlBytes + 15) & ~15;
    static constexpr int kEncDataOff = kKeyOff + kKeyStageWarp * kSvbWarps;
    static constexpr int kEncodeBytes = kEncDataOff + kTileValues * kMaxValueBytes + 32;
    // decode: per-warp data staging; a (possibly corrupt) stream may carry 4-byte codes whatever the mode
    static constexpr int kDecStageWarp = (kWarpValues * 4 + 32 + 15) & ~15;
    static constexpr int kDecDataOff = (kSvbCtrlBytes + 15) & ~15;
    static constexpr int kDecodeBytes = kDecDataOff + kDecStageWarp * kSvbWarps;
};

// =====================================================================================================
// encode
// =====================================================================================================
template <int ISZ, bool ZZ, int MODE>
__global__ void __launch_bounds__(kSvbThreads) svb_encode_kernel(SvbParams p)
{
    constexpr int G = kSvbGroups, VPG = 16 / ISZ, KB = VPG / 4;
    constexpr uint32_t TV = kSvbTileBytes / ISZ, WV = kSvbWarpTileBytes / ISZ;
    using L = SvbSmem<ISZ, MODE>;
    VBZ_DYN_SMEM(smem);
    uint64_t* ctrl = reinterpret_cast<uint64_t*>(smem + L::kCtrlOff);

    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x;
    const uint32_t r = p.tile_read[t];
    const SvbRead rd = p.reads[r];
    const uint32_t ti = t - rd.first_tile;
    if (rd.rejected) {
        if (ti == 0 && threadIdx.x == 0) p.results[r] = rd.preset;
        return;
    }
    if (threadIdx.x < 2 * kSvbWarps + 2) ctrl[threadIdx.x] = 0;
    __syncthreads();  // the only CTA-wide barrier: nothing is in flight yet

    const uint32_t n = rd.n_values;
    const uint32_t v0 = ti * TV + wid * WV;                                 // first value of this warp tile
    const uint32_t nv = v0 >= n ? 0u : (n - v0 < WV ? n - v0 : WV);         // values in this warp tile
    const uint8_t* src = p.src + rd.src_off;

    // ---- load 16 raw bytes per group, pre-transform, classify -------------------------------------
    uint32_t u[G][VPG];
    uint32_t key[G], len[G], cnt[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t g = j * 32 + lane, e0 = g * VPG;
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

    // ---- offsets: lane -> warp -> tile (shared slots) -> read (look-back, by the last warp) --------------
    static_assert(G == 2, "the packed scan below assumes two groups per lane");
    const uint32_t incl = warp_scan_u32(len[0] | (len[1] << 16));
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t tot0 = tot & 0xFFFFu, warp_total = tot0 + (tot >> 16);
    uint64_t warp_excl;
    chain_post(ctrl, ctrl + 2 * kSvbWarps, p.status_a, t, rd.first_tile, warp_total, kStValueMask, &warp_excl);
    const uint32_t off0 = (uint32_t)warp_excl + (incl & 0xFFFFu) - len[0];
    const uint32_t off1 = (uint32_t)warp_excl + tot0 + (incl >> 16) - len[1];

    // ---- stage control bytes (final phase known) and data bytes (tile-local offsets) ---------------------
    const uint32_t nkeys = (n + 3) / 4;
    uint8_t* out = p.dst + rd.dst_off;
    uint8_t* kdst = out + p.hdr_bytes + (v0 >> 2);
    uint8_t* kst = smem + L::kKeyOff + wid * L::kKeyStageWarp + ((uintptr_t)kdst & 15);
    uint8_t* dstg = smem + L::kEncDataOff;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t g = j * 32 + lane;
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
    __syncwarp();
    warp_copy_out(kdst, kst, (nv + 3) >> 2);

    // ---- the only use of the global offset: copy this warp's bytes out -------------------------------------
    const uint64_t tile_excl = chain_wait(ctrl + 2 * kSvbWarps);
    warp_copy_out_shifted(out + p.hdr_bytes + nkeys + tile_excl + warp_excl, dstg + warp_excl, warp_total);
    if (lane == 0) {
        if (ti == 0 && wid == 0 && p.hdr_bytes) *reinterpret_cast<uint32_t*>(out) = rd.src_bytes;
        // the warp holding the read's last value (or warp 0 of an empty read) reports the size
        if ((nv && v0 + nv == n) || (n == 0 && wid == 0)) p.results[r] = (uint32_t)(p.hdr_bytes + nkeys + tile_excl + warp_excl + warp_total);
    }
}

// =====================================================================================================
// decode, pre-pass: data-byte offset of every warp tile from the control bytes alone (one warp per read)
// =====================================================================================================
template <int ISZ, int MODE>
__global__ void __launch_bounds__(kSvbThreads) svb_key_offsets_kernel(SvbParams p)
{
    constexpr uint32_t WV = kSvbWarpTileBytes / ISZ;  // values per warp tile
    constexpr uint32_t KPW = WV / 4;                  // control bytes per warp tile (64 / 128 / 256)
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= p.n_reads) return;
    const SvbRead rd = p.reads[r];
    if (rd.rejected) return;
    const uint32_t n = rd.n_values;
    const uint32_t src_bytes = p.src_sizes ? p.src_sizes[r] : rd.src_bytes;
    const uint32_t nkeys = (n + 3) / 4;
    if (n == 0 || src_bytes >= kErrOutOfMemory || nkeys > src_bytes) return;  // svb_decode_kernel reports these
    const uint8_t* keys = p.src + rd.src_off;
    uint32_t* out = p.wtile_off + (uint64_t)rd.first_tile * kSvbWarps;
    const uint32_t n_wt = (n + WV - 1) / WV;
    uint32_t running = 0;
    constexpr int U = 4;  // warp tiles per iteration: independent loads in flight, interleaved reductions
    for (uint32_t wt0 = 0; wt0 < n_wt; wt0 += U) {
        uint32_t extra[U];  // sum of the 2-bit codes of each warp tile
#pragma unroll
        for (int q = 0; q < U; ++q) {
            extra[q] = 0;
            const uint32_t kb0 = (wt0 + q) * KPW;
            for (uint32_t wq = lane; wq < KPW / 4; wq += 32) {
                const uint32_t b0 = kb0 + 4 * wq;  // first control byte of this word; covers values 4*b0 .. 4*b0+15
                if (b0 >= nkeys) continue;
                uint32_t word;
                if (b0 + 4 <= nkeys) word = *reinterpret_cast<const uint32_t*>(keys + b0);
                else {
                    word = 0;
                    for (uint32_t b = b0; b < nkeys; ++b) word |= (uint32_t)keys[b] << (8 * (b - b0));
                }
                const uint32_t vleft = n - 4 * b0;  // values covered from this word on (>= 1)
                if (vleft < 16) word &= (1u << (2 * vleft)) - 1u;
                extra[q] += __popc(word & 0x55555555u) + 2 * __popc(word & 0xAAAAAAAAu);
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
#pragma unroll
            for (int q = 0; q < U; ++q) extra[q] += __shfl_xor_sync(0xffffffffu, extra[q], d);
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const uint32_t wt = wt0 + q;
            if (wt < n_wt) {
                if (lane == 0) out[wt] = running;
                running += (n - wt * WV < WV ? n - wt * WV : WV) + extra[q];
            }
        }
    }
}

// =====================================================================================================
// decode
// =====================================================================================================
template <int ISZ, bool ZZ, int MODE>
__global__ void __launch_bounds__(kSvbThreads) svb_decode_kernel(SvbParams p)
{
    constexpr int G = kSvbGroups, VPG = 16 / ISZ, KB = VPG / 4;
    constexpr uint32_t TV = kSvbTileBytes / ISZ, WV = kSvbWarpTileBytes / ISZ;
    using L = SvbSmem<ISZ, MODE>;
    VBZ_DYN_SMEM(smem);
    uint64_t* ctrl = reinterpret_cast<uint64_t*>(smem + L::kCtrlOff);

    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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
        if (p.src_sizes && src_bytes >= kErrOutOfMemory) { early = src_bytes; stop = true; }  // the zstd stage failed first
        else if (rd.rejected) { early = rd.preset; stop = true; }
        else if (n == 0) {  // x86 int16 path returns 0 before looking at the input (impl_sse3.h:473-476);
                            // streamvbyte_validate_stream wants both counts to be zero (streamvbyte_decode.c:91-92)
            early = (MODE == kModeI16X86 || src_bytes == 0) ? 0u : (uint32_t)kErrStream;
            stop = true;
        } else if (nkeys > src_bytes) {  // impl_sse3.h:479-482 vs streamvbyte_decode.c:100-101
            early = MODE == kModeI16X86 ? (uint32_t)kErrInputSize : (uint32_t)kErrStream;
            stop = true;
        }
        if (stop) {
            if (ti == 0 && threadIdx.x == 0) p.results[r] = early;
            return;
        }
    }
    if (ZZ) {
        if (threadIdx.x < 2 * kSvbWarps + 2) ctrl[threadIdx.x] = 0;
        __syncthreads();  // the only CTA-wide barrier: nothing is in flight yet
    }
    const uint32_t ndata = src_bytes - nkeys;
    const uint32_t v0 = ti * TV + wid * WV;
    const uint32_t nv = v0 >= n ? 0u : (n - v0 < WV ? n - v0 : WV);
    const uint8_t* keys = p.src + rd.src_off + (v0 >> 2);
    const uint8_t* data = p.src + rd.src_off + nkeys;
    const uint32_t tile_excl = nv ? p.wtile_off[(uint64_t)t * kSvbWarps + wid] : 0u;  // data bytes before this warp tile

    // ---- control bytes -> per-group byte counts ------------------------------------------------------------
    uint32_t key[G], len[G], cnt[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t g = j * 32 + lane, e0 = g * VPG;
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
    static_assert(G == 2, "the packed scan below assumes two groups per lane");
    const uint32_t incl = warp_scan_u32(len[0] | (len[1] << 16));
    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t tot0 = tot & 0xFFFFu, warp_total = tot0 + (tot >> 16);
    const uint32_t off0 = (incl & 0xFFFFu) - len[0];
    const uint32_t off1 = tot0 + (incl >> 16) - len[1];

    // ---- stage this warp tile's data bytes (clipped to the read's data region; the rest reads as zero) -------
    const uint8_t* gsrc = data + tile_excl;
    uint8_t* stg = smem + L::kDecDataOff + wid * L::kDecStageWarp + ((uintptr_t)gsrc & 15);
    const uint32_t avail = tile_excl >= ndata ? 0u : (ndata - tile_excl < warp_total ? ndata - tile_excl : warp_total);
    warp_copy_in(stg, gsrc, avail);
    for (uint32_t i = avail + lane; i < warp_total; i += 32) stg[i] = 0;
    __syncwarp();

    // ---- unpack, un-zigzag, prefix-sum ---------------------------------------------------------------------
    uint32_t val[G][VPG];
    uint32_t gsum[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t e0 = (j * 32 + lane) * VPG;
        uint32_t o = j == 0 ? off0 : off1;
        // x86 build: a group of 8 goes through the SIMD body (16-bit un-zigzag of the low 16 bits) while it is
        // one of the count/8 full groups and at least 32 data bytes remain; otherwise the scalar tail
        // (32-bit un-zigzag, truncated on store) - impl_sse3.h:494-500,505-520,542-572.
        const bool simd16 = MODE == kModeI16X86 && ((v0 + e0) >> 3) < (n >> 3) && (uint64_t)tile_excl + o + 32 <= (uint64_t)ndata;
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
        const uint32_t i0 = warp_scan_u32(gsum[0]), i1 = warp_scan_u32(gsum[1]);
        const uint32_t t0s = __shfl_sync(0xffffffffu, i0, 31), t1s = __shfl_sync(0xffffffffu, i1, 31);
        uint64_t warp_excl;
        chain_post(ctrl + kSvbWarps, ctrl + 2 * kSvbWarps + 1, p.status_b, t, rd.first_tile, (uint64_t)(t0s + t1s), 0xFFFFFFFFull, &warp_excl);
        const uint32_t base = (uint32_t)chain_wait(ctrl + 2 * kSvbWarps + 1) + (uint32_t)warp_excl;
        carry0 = base + i0 - gsum[0];
        carry1 = base + t0s + i1 - gsum[1];
    }

    // ---- store samples ----------------------------------------------------------------------------------------
    uint8_t* out = p.dst + rd.dst_off;
#pragma unroll
    for (int j = 0; j < G; ++j) {
        const uint32_t e0 = (j * 32 + lane) * VPG;
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
    if (lane == 0 && nv && v0 + nv == n) {
        // all data bytes must be consumed exactly: impl_sse3.h:574-577,638-642 / streamvbyte_decode.c:135
        p.results[r] = ((uint64_t)tile_excl + warp_total == (uint64_t)ndata) ? n * (uint32_t)ISZ : (uint32_t)kErrStream;
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
