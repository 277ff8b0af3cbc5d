// csrc/svb16_kernels.cuh - the signal path of VBZ stage 1: int16 + delta + zigzag + StreamVByte with the
// arithmetic of the x86 reference build, written for instruction count.
//
// Replaces StreamVByteWorkerV0<int16_t, true>::compress / ::decompress
// (ref vbz/v0/vbz_streamvbyte_impl_sse3.h:406-466 and :468-580) for every read of a batch at once.
// The other option combinations keep the generic kernels of csrc/svb_kernels.cuh.
//
// Why a second set of kernels: the generic ones spend ~80 thread-instructions per sample and are issue
// bound at 13 % of the HBM roofline.  Here a lane handles GROUPS of 8 samples (one 128-bit word) and all
// per-sample arithmetic is done two samples at a time in packed 16-bit lanes (VIADD.16x2 / PRMT / LOP3):
//   encode  delta        pw = PRMT(prev, x)            ; d = x -16x2 pw
//           zigzag       z = (d +16x2 d) ^ PRMT.sign(d)
//           codes        PRMT gathers the 4 high bytes of 4 samples, a carry trick flags the non-zero
//                        ones, one IMAD moves the 4 flags to the control-byte bit positions
//           pack         a 16-entry shared-memory table of PRMT selectors compacts 4 samples (4..8
//                        bytes) at a time; funnel shifts splice the two halves and place the group at
//                        its byte offset in the shared-memory staging tile (word stores + edge bytes)
//   decode  the inverse: 5-word window at the group's byte offset, selector table, packed un-zigzag,
//           packed in-register prefix sum, VIADD.16x2 warp scan.
// Geometry: a CTA (8 warps) owns a TILE of 8192 samples (16 KiB raw); a warp owns 1024 of them as
// 4 rows of 32 groups (group = row * 32 + lane) so that every global 128-bit access of a warp is one
// contiguous 512-byte segment.  Tiles are numbered read-major; blockIdx order is data order.
// Offsets: lane -> warp by shuffle scan, warp -> tile through shared memory, tile -> read by a
// decoupled look-back over 64-bit status words (csrc/svb_kernels.cuh: lookback_exclusive).
// Data movement: encode reads samples with coalesced streaming 128-bit loads and writes BOTH output
// streams (control bytes, data bytes) from shared-memory staging with TMA bulk stores
// (cp.async.bulk.global.shared::cta), staged at the same 16-byte phase as the destination; decode pulls
// its variable-length slice of data bytes and its control bytes into shared memory with TMA bulk loads
// (cp.async.bulk.shared::cluster.global + mbarrier) issued before anything else, and stores samples with
// coalesced streaming 128-bit stores.
//
// HBM traffic per sample: encode 2 B in + 0.25 B + ~1.01 B out; decode the reverse + 0.25 B (the control
// bytes are read twice: once by svb16_tile_sums_kernel, which turns them into per-tile data offsets).
#pragma once
#include "svb_kernels.cuh"

namespace vbz {

constexpr int kS16Threads = 256;
constexpr int kS16Warps = 8;
constexpr int kS16Rows = 4;                                         // groups per lane
constexpr uint32_t kS16WarpValues = 32 * kS16Rows * 8;              // 1024
constexpr uint32_t kS16TileValues = kS16Warps * kS16WarpValues;     // 8192
constexpr uint32_t kS16TileKeys = kS16TileValues / 4;               // 2048
constexpr uint32_t kS16LutWords = 96;                               // indexed by the sparse 4-code key (bits 0,2,4,6)

// shared memory carve-up (bytes)
struct S16Smem {
    static constexpr int kLut = 0;                                  // encode: 96 x u32, decode: 96 x uint2
    static constexpr int kWtot = kLut + kS16LutWords * 8;           // 8 warp totals (chain A) + 8 (chain B)
    static constexpr int kSlots = kWtot + 16 * 4;                   // two u64 look-back results, one mbarrier, flags
    static constexpr int kKeys = kSlots + 48;                       // control bytes (+16 phase)
    static constexpr int kData = kKeys + kS16TileKeys + 16;         // data bytes (+16 phase, +32 window over-read)
    static constexpr int kBytes = kData + kS16TileValues * 2 + 64;
};
static_assert(S16Smem::kKeys % 16 == 0 && S16Smem::kData % 16 == 0, "staging areas must be 16-byte aligned");

VBZ_HD uint32_t s16_tiles_of(uint32_t n_values)
{
    const uint32_t t = (uint32_t)(((uint64_t)n_values + kS16TileValues - 1) / kS16TileValues);
    return t ? t : 1;
}

// ---- selector tables ------------------------------------------------------------------------------------
// k = codes of 4 samples at bits 0, 2, 4, 6 (a control byte restricted to codes 0/1).
// encode: 8 PRMT selector nibbles that compact (lo0 hi0 lo1 hi1 lo2 hi2 lo3 hi3) to the kept bytes; unused
// output bytes select a high byte that is known to be zero.
__device__ __forceinline__ uint32_t s16_make_enc_sel(uint32_t k)
{
    uint32_t sel = 0, pos = 0, zero_src = 0;
    for (uint32_t m = 0; m < 4; ++m)
        if (!((k >> (2 * m)) & 1u)) { zero_src = 2 * m + 1; break; }
    for (uint32_t m = 0; m < 4; ++m) {
        sel |= (2 * m) << (4 * pos++);
        if ((k >> (2 * m)) & 1u) sel |= (2 * m + 1) << (4 * pos++);
    }
    while (pos < 8) sel |= zero_src << (4 * pos++);
    return sel;
}
// decode: selector nibbles that expand 4..8 packed bytes (two window words, the second masked to the
// bytes that belong to these 4 samples so that byte 7 reads as zero whenever a high byte is missing)
__device__ __forceinline__ uint2 s16_make_dec_sel(uint32_t k)
{
    uint32_t sel = 0, pos = 0;
    for (uint32_t m = 0; m < 4; ++m) {
        sel |= pos++ << (8 * m);
        sel |= (((k >> (2 * m)) & 1u) ? pos++ : 7u) << (8 * m + 4);
    }
    const uint32_t extra = pos - 4;  // bytes taken from the second window word
    return make_uint2(sel, extra == 4 ? 0xFFFFFFFFu : ((1u << (8 * extra)) - 1u));
}

// flags (bit 7 of each byte) of the non-zero bytes of w
__device__ __forceinline__ uint32_t s16_nonzero_bytes(uint32_t w) { return (((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u; }
// bits 7/15/23/31 -> bits 0/2/4/6
__device__ __forceinline__ uint32_t s16_flags_to_codes(uint32_t f) { return ((f >> 7) * 0x01041040u) >> 24; }

// Place `len` bytes (little-endian in r0..r3, zero above) at byte offset P of the staging area: whole words
// where the group owns them, single bytes at the two edges it shares with its neighbours.
__device__ __forceinline__ void s16_place(uint8_t* stage, uint32_t P, uint32_t len, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3)
{
    const uint32_t a = P & 3u, sh = a * 8u, end = a + len;
    const uint32_t s0 = r0 << sh;
    const uint32_t s1 = __funnelshift_l(r0, r1, sh), s2 = __funnelshift_l(r1, r2, sh), s3 = __funnelshift_l(r2, r3, sh);
    const uint32_t s4 = __funnelshift_l(r3, 0u, sh);
    uint8_t* b = stage + (P & ~3u);
    uint32_t* w = reinterpret_cast<uint32_t*>(b);
    if (a == 0 && end >= 4) w[0] = s0;
    if (end >= 8) w[1] = s1;
    if (end >= 12) w[2] = s2;
    if (end >= 16) w[3] = s3;
    if (a != 0) {
        if (a <= 1 && 1 < end) b[1] = (uint8_t)(s0 >> 8);
        if (a <= 2 && 2 < end) b[2] = (uint8_t)(s0 >> 16);
        if (3 < end) b[3] = (uint8_t)(s0 >> 24);
    }
    const uint32_t re = end & 3u, qe = end >> 2;
    if (re != 0 && (qe != 0 || a == 0)) {
        const uint32_t st = qe == 0 ? s0 : qe == 1 ? s1 : qe == 2 ? s2 : qe == 3 ? s3 : s4;
        uint8_t* t = b + 4 * qe;
        t[0] = (uint8_t)st;
        if (re > 1) t[1] = (uint8_t)(st >> 8);
        if (re > 2) t[2] = (uint8_t)(st >> 16);
    }
}

// Copy `n` staged bytes to global memory: TMA bulk store for the 16-byte aligned body, byte stores for the
// (< 16 byte) head and tail.  stage + (g & 15) is where byte 0 sits, so both sides share their phase.
// Called by all threads of the CTA after the staging area has been fenced and barriered.
__device__ __forceinline__ void s16_copy_out(uint8_t* g, const uint8_t* stage_at_g, uint32_t n, uint32_t tid, uint32_t issuer)
{
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > n) head = n;
    const uint32_t body = (n - head) & ~15u, tail = n - head - body;
    if (tid == issuer && body) vbz_bulk_store(g + head, stage_at_g + head, body);
    const uint32_t i = tid - issuer - 1;  // the threads after the issuer take the edges
    if (i < head) g[i] = stage_at_g[i];
    else if (i - 16u < tail) g[head + body + (i - 16u)] = stage_at_g[head + body + (i - 16u)];
}

// =====================================================================================================
// encode
// =====================================================================================================
__global__ void __launch_bounds__(kS16Threads, 4) svb16_encode_kernel(SvbParams p)
{
    VBZ_DYN_SMEM(smem);
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(smem + S16Smem::kLut);
    uint32_t* s_wtot = reinterpret_cast<uint32_t*>(smem + S16Smem::kWtot);
    uint64_t* s_excl = reinterpret_cast<uint64_t*>(smem + S16Smem::kSlots);
    uint8_t* s_keys = smem + S16Smem::kKeys;
    uint8_t* s_data = smem + S16Smem::kData;

    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t t = blockIdx.x;
    const uint32_t r = p.tile_read[t];
    const SvbRead rd = p.reads[r];
    const uint32_t ti = t - rd.first_tile;
    if (rd.rejected) {
        if (ti == 0 && tid == 0) p.results[r] = rd.preset;
        return;
    }
    const uint32_t n = rd.n_values;
    uint8_t* out = p.dst + rd.dst_off;
    const uint32_t nkeys = (n + 3) / 4;
    if (n == 0) {  // an empty read owns one empty tile
        if (tid == 0) {
            if (p.hdr_bytes) *reinterpret_cast<uint32_t*>(out) = rd.src_bytes;
            p.results[r] = p.hdr_bytes;
        }
        return;
    }
    if (tid < 86 && !(tid & 0xAAu)) s_lut[tid] = s16_make_enc_sel(tid);

    // ---- load: 4 rows of 32 groups per warp, 16 bytes per lane and row -----------------------------------
    const uint8_t* src = p.src + rd.src_off;
    const uint32_t tv0 = ti * kS16TileValues, wv0 = tv0 + wid * kS16WarpValues;
    uint4 q[kS16Rows];
    uint32_t cnt[kS16Rows];
#pragma unroll
    for (int j = 0; j < kS16Rows; ++j) {
        const uint32_t s0 = wv0 + j * 256 + lane * 8;
        cnt[j] = s0 >= n ? 0u : (n - s0 < 8u ? n - s0 : 8u);
        // the slot of a read is padded to 16 bytes, so the last (partial) group can be loaded whole
        q[j] = cnt[j] ? vbz_ld_stream(reinterpret_cast<const uint4*>(src + 2ull * s0)) : make_uint4(0, 0, 0, 0);
    }
    uint32_t carry_w = 0;  // lane 0: the sample before this warp's first one, in the high half
    if (lane == 0 && wv0 > 0 && wv0 < n) carry_w = (uint32_t)(*reinterpret_cast<const uint16_t*>(src + 2ull * (wv0 - 1))) << 16;

    // ---- delta, zigzag, codes --------------------------------------------------------------------------------
    uint32_t z[kS16Rows][4], key[kS16Rows], len[kS16Rows];
#pragma unroll
    for (int j = 0; j < kS16Rows; ++j) {
        const uint32_t send = (lane == 31 && j > 0) ? q[j > 0 ? j - 1 : 0].w : q[j].w;
        const uint32_t recv = __shfl_sync(0xffffffffu, send, (lane + 31) & 31);
        const uint32_t prevw = (lane == 0 && j == 0) ? carry_w : recv;
        const uint32_t x[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
        const uint32_t pw[4] = {__byte_perm(prevw, x[0], 0x5432), __byte_perm(x[0], x[1], 0x5432), __byte_perm(x[1], x[2], 0x5432),
                                __byte_perm(x[2], x[3], 0x5432)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t d = vbz_sub16x2(x[i], pw[i]);                                  // 16-bit wrapping delta (impl_sse3.h:434-436)
            z[j][i] = vbz_add16x2(d, d) ^ vbz_prmt(d, 0u, 0xBB99);                     // (d << 1) ^ (d >> 15)  (impl_sse3.h:438-440)
        }
        if (cnt[j] < 8) {  // the read's last, partial group: samples past the end contribute nothing
#pragma unroll
            for (int i = 0; i < 4; ++i) z[j][i] &= cnt[j] > 2u * i + 1 ? 0xFFFFFFFFu : (cnt[j] > 2u * i ? 0x0000FFFFu : 0u);
        }
        const uint32_t k0 = s16_flags_to_codes(s16_nonzero_bytes(__byte_perm(z[j][0], z[j][1], 0x7531)));
        const uint32_t k1 = s16_flags_to_codes(s16_nonzero_bytes(__byte_perm(z[j][2], z[j][3], 0x7531)));
        key[j] = k0 | (k1 << 8);
        len[j] = cnt[j] + __popc(key[j]);
    }

    // ---- offsets: lane -> warp (packed shuffle scan) -> tile (shared) -> read (look-back) -----------------------
    const uint32_t ia = warp_scan_u32(len[0] | (len[1] << 16)), ib = warp_scan_u32(len[2] | (len[3] << 16));
    const uint32_t ta = __shfl_sync(0xffffffffu, ia, 31), tb = __shfl_sync(0xffffffffu, ib, 31);
    const uint32_t t0 = ta & 0xFFFFu, t1 = ta >> 16, t2 = tb & 0xFFFFu, t3 = tb >> 16;
    const uint32_t warp_total = t0 + t1 + t2 + t3;
    uint32_t off[kS16Rows];
    off[0] = (ia & 0xFFFFu) - len[0];
    off[1] = t0 + (ia >> 16) - len[1];
    off[2] = t0 + t1 + (ib & 0xFFFFu) - len[2];
    off[3] = t0 + t1 + t2 + (ib >> 16) - len[3];
    if (lane == 0) s_wtot[wid] = warp_total;
    __syncthreads();
    uint32_t warp_excl = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kS16Warps; ++w) {
        const uint32_t v = s_wtot[w];
        if ((uint32_t)w < wid) warp_excl += v;
        tile_total += v;
    }
    if (wid == 0) {
        uint64_t excl = 0;
        if (ti == 0) {
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStInclusive | tile_total);
        } else {
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStAggregate | tile_total);
            excl = lookback_exclusive(p.status_a, t, rd.first_tile);
            if (lane == 0) vbz_st_volatile_u64(p.status_a + t, kStInclusive | ((excl + tile_total) & kStValueMask));
        }
        if (lane == 0) *s_excl = excl;
    }
    __syncthreads();
    const uint64_t tile_excl = *s_excl;

    // ---- stage both streams at the 16-byte phase of their destination ---------------------------------------------
    uint8_t* kg = out + p.hdr_bytes + (tv0 >> 2);
    uint8_t* dg = out + p.hdr_bytes + nkeys + tile_excl;
    const uint32_t dphase = (uint32_t)((uintptr_t)dg & 15u);
    uint8_t* kstage = s_keys + ((uintptr_t)kg & 15u);
    uint8_t* dstage = s_data + dphase;
#pragma unroll
    for (int j = 0; j < kS16Rows; ++j) {
        if (cnt[j]) {
            *reinterpret_cast<uint16_t*>(kstage + wid * 256 + j * 64 + lane * 2) = (uint16_t)key[j];  // kg is even: hdr 0 or 4, 16-byte slots
            const uint32_t k0 = key[j] & 0x55u, k1 = key[j] >> 8;
            const uint32_t sel0 = s_lut[k0], sel1 = s_lut[k1];
            const uint32_t c0 = __byte_perm(z[j][0], z[j][1], sel0), c1 = __byte_perm(z[j][0], z[j][1], sel0 >> 16);
            const uint32_t c2 = __byte_perm(z[j][2], z[j][3], sel1), c3 = __byte_perm(z[j][2], z[j][3], sel1 >> 16);
            const uint32_t sh = 8u * __popc(k0);  // the second half starts at byte 4 + popc(k0)
            s16_place(s_data, dphase + warp_excl + off[j], len[j], c0, c1 | __funnelshift_lc(0u, c2, sh), __funnelshift_lc(c2, c3, sh),
                      __funnelshift_lc(c3, 0u, sh));
        }
    }
    vbz_fence_async_smem();
    __syncthreads();

    // ---- out: TMA bulk stores for the aligned bodies, byte stores for the edges -------------------------------------
    const uint32_t tile_values = n - tv0 < kS16TileValues ? n - tv0 : kS16TileValues;
    s16_copy_out(kg, kstage, (tile_values + 3) >> 2, tid, 0);
    s16_copy_out(dg, dstage, tile_total, tid, 64);
    if (tid == 0 || tid == 64) vbz_bulk_store_commit_wait();
    if (tid == 128) {
        if (ti == 0 && p.hdr_bytes) *reinterpret_cast<uint32_t*>(out) = rd.src_bytes;
        if (tv0 + tile_values == n) p.results[r] = (uint32_t)(p.hdr_bytes + nkeys + tile_excl + tile_total);
    }
}

// =====================================================================================================
// decode, pre-pass 1: sum of the data-byte lengths of every tile, from the control bytes alone.
// One warp per tile (2 KiB of control bytes = 4 x 128-bit loads per lane).
// =====================================================================================================
__global__ void __launch_bounds__(kS16Threads) svb16_tile_sums_kernel(SvbParams p)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= p.n_tiles) return;
    const uint32_t r = p.tile_read[t];
    const SvbRead rd = p.reads[r];
    const uint32_t n = rd.n_values;
    const uint32_t src_bytes = p.src_sizes ? p.src_sizes[r] : rd.src_bytes;
    const uint32_t nkeys = (n + 3) / 4;
    uint32_t sum = 0;
    if (!rd.rejected && n != 0 && src_bytes < kErrOutOfMemory && nkeys <= src_bytes) {
        const uint32_t ti = t - rd.first_tile;
        const uint8_t* keys = p.src + rd.src_off;  // 4-byte aligned (16-byte slots, optional 4-byte header)
        const uint32_t kb0 = ti * kS16TileKeys;
        const uint32_t kb1 = nkeys - kb0 < kS16TileKeys ? nkeys : kb0 + kS16TileKeys;
        for (uint32_t b = kb0 + lane * 4; b < kb1; b += 128) {
            uint32_t word;
            if (b + 4 <= kb1) word = *reinterpret_cast<const uint32_t*>(keys + b);
            else {
                word = 0;
                for (uint32_t i = b; i < kb1; ++i) word |= (uint32_t)keys[i] << (8 * (i - b));
            }
            const uint32_t vleft = n - 4 * b;  // values from this word on (>= 1)
            if (vleft < 16) word &= (1u << (2 * vleft)) - 1u;
            sum += __popc(word & 0x55555555u) + 2 * __popc(word & 0xAAAAAAAAu);
        }
        const uint32_t tv0 = ti * kS16TileValues;
        if (lane == 0) sum += n - tv0 < kS16TileValues ? n - tv0 : kS16TileValues;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) p.wtile_off[(uint64_t)p.n_tiles + t] = sum;
}
// pre-pass 2: exclusive scan of the tile sums inside each read (one warp per read)
__global__ void __launch_bounds__(kS16Threads) svb16_tile_offsets_kernel(SvbParams p)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= p.n_reads) return;
    const SvbRead rd = p.reads[r];
    const uint32_t nt = s16_tiles_of(rd.n_values);
    const uint32_t* sums = p.wtile_off + p.n_tiles + rd.first_tile;
    uint32_t* offs = p.wtile_off + rd.first_tile;
    uint32_t running = 0;
    for (uint32_t i0 = 0; i0 < nt; i0 += 32) {
        const uint32_t v = i0 + lane < nt ? sums[i0 + lane] : 0u;
        const uint32_t incl = warp_scan_u32(v);
        if (i0 + lane < nt) offs[i0 + lane] = running + incl - v;
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// =====================================================================================================
// decode
// =====================================================================================================
__global__ void __launch_bounds__(kS16Threads, 4) svb16_decode_kernel(SvbParams p)
{
    VBZ_DYN_SMEM(smem);
    uint2* s_lut = reinterpret_cast<uint2*>(smem + S16Smem::kLut);
    uint32_t* s_wtot = reinterpret_cast<uint32_t*>(smem + S16Smem::kWtot);
    uint64_t* s_excl = reinterpret_cast<uint64_t*>(smem + S16Smem::kSlots);
    uint64_t* s_bar = s_excl + 2;
    uint32_t* s_flag = reinterpret_cast<uint32_t*>(s_excl + 3);
    uint8_t* s_keys = smem + S16Smem::kKeys;
    uint8_t* s_data = smem + S16Smem::kData;

    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
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
        else if (n == 0) { early = 0; stop = true; }                              // impl_sse3.h:473-476
        else if (nkeys > src_bytes) { early = kErrInputSize; stop = true; }       // impl_sse3.h:479-482
        if (stop) {
            if (ti == 0 && tid == 0) p.results[r] = early;
            return;
        }
    }
    const uint32_t ndata = src_bytes - nkeys;
    const uint32_t tv0 = ti * kS16TileValues;
    const uint32_t tile_values = n - tv0 < kS16TileValues ? n - tv0 : kS16TileValues;
    const uint32_t tile_keys = (tile_values + 3) >> 2;
    const uint32_t tile_off = p.wtile_off[t];                       // data bytes of this read before the tile
    const uint32_t tile_len = p.wtile_off[(uint64_t)p.n_tiles + t];  // data bytes the tile's control bytes ask for
    const uint8_t* kg = p.src + rd.src_off + (tv0 >> 2);
    const uint8_t* dbase = p.src + rd.src_off + nkeys;
    const uint8_t* dg = dbase + tile_off;
    // the packed path needs the tile's data bytes in shared memory: all present, at most 2 per sample
    const bool short_data = (uint64_t)tile_off + tile_len > (uint64_t)ndata || tile_len > 2u * tile_values;
    const uint32_t kphase = (uint32_t)((uintptr_t)kg & 15u), dphase = (uint32_t)((uintptr_t)dg & 15u);

    // ---- TMA: control bytes and this tile's slice of the data bytes -> shared memory ---------------------------
    if (tid == 0) {
        vbz_mbar_init(s_bar, 1);
        *s_flag = 0;
        const uint32_t kbytes = (kphase + tile_keys + 15u) & ~15u;
        const uint32_t dbytes = short_data ? 0u : (dphase + tile_len + 15u) & ~15u;
        vbz_mbar_expect_tx(s_bar, kbytes + dbytes);
        vbz_bulk_load(s_keys, kg - kphase, kbytes, s_bar);
        if (dbytes) vbz_bulk_load(s_data, dg - dphase, dbytes, s_bar);
    }
    if (tid >= 32 && tid < 32 + 86 && !((tid - 32) & 0xAAu)) s_lut[tid - 32] = s16_make_dec_sel(tid - 32);
    __syncthreads();  // mbarrier initialised, table ready
    vbz_mbar_wait(s_bar, 0);

    // ---- control bytes -> lengths -> offsets ------------------------------------------------------------------
    const uint32_t wv0 = tv0 + wid * kS16WarpValues;
    uint32_t key[kS16Rows], len[kS16Rows], cnt[kS16Rows];
    uint32_t wide = 0;
#pragma unroll
    for (int j = 0; j < kS16Rows; ++j) {
        const uint32_t s0 = wv0 + j * 256 + lane * 8;
        cnt[j] = s0 >= n ? 0u : (n - s0 < 8u ? n - s0 : 8u);
        uint32_t kb = cnt[j] ? *reinterpret_cast<const uint16_t*>(s_keys + kphase + wid * 256 + j * 64 + lane * 2) : 0u;
        if (cnt[j] < 8) kb &= (1u << (2 * cnt[j])) - 1u;  // bits of values past the end are ignored by the reference
        key[j] = kb;
        wide |= kb & 0xAAAAu;
        len[j] = cnt[j] + __popc(kb & 0x5555u) + 2 * __popc(kb & 0xAAAAu);
    }
    const uint32_t ia = warp_scan_u32(len[0] | (len[1] << 16)), ib = warp_scan_u32(len[2] | (len[3] << 16));
    const uint32_t ta = __shfl_sync(0xffffffffu, ia, 31), tb = __shfl_sync(0xffffffffu, ib, 31);
    const uint32_t t0 = ta & 0xFFFFu, t1 = ta >> 16, t2 = tb & 0xFFFFu, t3 = tb >> 16;
    uint32_t off[kS16Rows];
    off[0] = (ia & 0xFFFFu) - len[0];
    off[1] = t0 + (ia >> 16) - len[1];
    off[2] = t0 + t1 + (ib & 0xFFFFu) - len[2];
    off[3] = t0 + t1 + t2 + (ib >> 16) - len[3];
    if (lane == 0) s_wtot[wid] = t0 + t1 + t2 + t3;
    if (__any_sync(0xffffffffu, wide != 0) && lane == 0) *s_flag = 1;  // 3- or 4-byte codes: not an x86-int16 stream
    __syncthreads();
    uint32_t warp_excl = 0;
#pragma unroll
    for (int w = 0; w < kS16Warps; ++w)
        if ((uint32_t)w < wid) warp_excl += s_wtot[w];
    const bool slow = short_data || *s_flag != 0;

    // ---- unpack + un-zigzag + prefix sum inside the group (two samples per 32-bit word) ---------------------------
    uint32_t y[kS16Rows][4];
    if (!slow) {
#pragma unroll
        for (int j = 0; j < kS16Rows; ++j) {
            const uint32_t P = dphase + warp_excl + off[j];
            const uint32_t* w = reinterpret_cast<const uint32_t*>(s_data + (P & ~3u));
            const uint32_t sh = 8u * (P & 3u);
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
            const uint32_t r0 = __funnelshift_r(w0, w1, sh), r1 = __funnelshift_r(w1, w2, sh), r2 = __funnelshift_r(w2, w3, sh),
                           r3 = __funnelshift_r(w3, w4, sh);
            const uint32_t k0 = key[j] & 0x55u, k1 = key[j] >> 8;
            const uint2 e0 = s_lut[k0], e1 = s_lut[k1];
            const uint32_t s2 = 8u * __popc(k0);  // the second half starts at byte 4 + popc(k0)
            const uint32_t a1 = r1 & e0.y;
            const uint32_t b0 = __funnelshift_rc(r1, r2, s2), b1 = __funnelshift_rc(r2, r3, s2) & e1.y;
            const uint32_t zz[4] = {__byte_perm(r0, a1, e0.x), __byte_perm(r0, a1, e0.x >> 16), __byte_perm(b0, b1, e1.x), __byte_perm(b0, b1, e1.x >> 16)};
            uint32_t run = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // (z >> 1) ^ -(z & 1) in both halves (impl_sse3.h:516-520), then x_lo, x_lo + x_hi, plus the running sum
                const uint32_t d = ((zz[i] >> 1) & 0x7FFF7FFFu) ^ vbz_prmt(zz[i] << 7, 0u, 0xAA88);
                const uint32_t pre = vbz_add16x2(d, d << 16);
                y[j][i] = vbz_add16x2(pre, run);
                run = __byte_perm(y[j][i], 0u, 0x3232);
            }
        }
    } else {
        // exact scalar semantics for tiles the packed path cannot express: codes 2/3 in the stream (the x86
        // build then decodes 8-value groups 16 bits wide while 32 data bytes remain and 32 bits wide in its
        // scalar tail, impl_sse3.h:494-572) or a stream that ends early (bytes past the end read as zero)
#pragma unroll
        for (int j = 0; j < kS16Rows; ++j) {
            const uint32_t s0 = wv0 + j * 256 + lane * 8;
            uint64_t o = (uint64_t)tile_off + warp_excl + off[j];
            const bool simd16 = (s0 >> 3) < (n >> 3) && o + 32 <= (uint64_t)ndata;
            uint32_t run = 0, prevv = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t v = 0;
                if ((uint32_t)k < cnt[j]) {
                    const uint32_t code = (key[j] >> (2 * k)) & 3u;
                    for (uint32_t b = 0; b <= code; ++b, ++o) v |= (o < ndata ? (uint32_t)dbase[o] : 0u) << (8 * b);
                }
                if (simd16) v &= 0xFFFFu;
                run += (v >> 1) ^ (0u - (v & 1u));
                if (k & 1) y[j][k >> 1] = (prevv & 0xFFFFu) | (run << 16);
                prevv = run;
            }
        }
    }

    // ---- running sample value: lane -> warp (packed 16-bit scan) -> tile -> read (look-back) -----------------------
    const uint32_t la = __byte_perm(y[0][3], y[1][3], 0x7632), lb = __byte_perm(y[2][3], y[3][3], 0x7632);  // row totals, two per word
    uint32_t sa = la, sb = lb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t xa = __shfl_up_sync(0xffffffffu, sa, d), xb = __shfl_up_sync(0xffffffffu, sb, d);
        if (lane >= (uint32_t)d) { sa = vbz_add16x2(sa, xa); sb = vbz_add16x2(sb, xb); }
    }
    const uint32_t wa = __shfl_sync(0xffffffffu, sa, 31), wb = __shfl_sync(0xffffffffu, sb, 31);
    const uint32_t r0t = wa & 0xFFFFu, r1t = wa >> 16, r2t = wb & 0xFFFFu, r3t = wb >> 16;
    const uint32_t warp_sum = (r0t + r1t + r2t + r3t) & 0xFFFFu;
    if (lane == 0) s_wtot[8 + wid] = warp_sum;
    __syncthreads();
    uint32_t wsum_excl = 0, tile_sum = 0;
#pragma unroll
    for (int w = 0; w < kS16Warps; ++w) {
        const uint32_t v = s_wtot[8 + w];
        if ((uint32_t)w < wid) wsum_excl += v;
        tile_sum += v;
    }
    tile_sum &= 0xFFFFu;
    if (wid == 0) {
        uint64_t excl = 0;
        if (ti == 0) {
            if (lane == 0) vbz_st_volatile_u64(p.status_b + t, kStInclusive | tile_sum);
        } else {
            if (lane == 0) vbz_st_volatile_u64(p.status_b + t, kStAggregate | tile_sum);
            excl = lookback_exclusive(p.status_b, t, rd.first_tile) & 0xFFFFu;
            if (lane == 0) vbz_st_volatile_u64(p.status_b + t, kStInclusive | ((excl + tile_sum) & 0xFFFFu));
        }
        if (lane == 0) s_excl[1] = excl;
    }
    __syncthreads();
    const uint32_t base = (uint32_t)s_excl[1] + wsum_excl;
    // exclusive carry of each row: everything before the warp, the rows before it, the lanes before it
    const uint32_t ea = vbz_sub16x2(sa, la), eb = vbz_sub16x2(sb, lb);
    uint32_t carry[kS16Rows];
    carry[0] = base + (ea & 0xFFFFu);
    carry[1] = base + r0t + (ea >> 16);
    carry[2] = base + r0t + r1t + (eb & 0xFFFFu);
    carry[3] = base + r0t + r1t + r2t + (eb >> 16);

    // ---- store samples -------------------------------------------------------------------------------------------
    uint8_t* outp = p.dst + rd.dst_off;
#pragma unroll
    for (int j = 0; j < kS16Rows; ++j) {
        const uint32_t s0 = wv0 + j * 256 + lane * 8;
        const uint32_t c = __byte_perm(carry[j], 0u, 0x1010);
        const uint4 v = make_uint4(vbz_add16x2(y[j][0], c), vbz_add16x2(y[j][1], c), vbz_add16x2(y[j][2], c), vbz_add16x2(y[j][3], c));
        uint8_t* gp = outp + 2ull * s0;
        if (cnt[j] == 8) vbz_st_stream(reinterpret_cast<uint4*>(gp), v);
        else if (cnt[j]) {
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
            for (uint32_t b = 0; b < cnt[j] * 2; ++b) gp[b] = (uint8_t)(wv[b >> 2] >> (8 * (b & 3)));
        }
    }
    if (tid == 0 && tv0 + tile_values == n) {
        // all data bytes must be consumed exactly: impl_sse3.h:574-577, 638-642
        p.results[r] = ((uint64_t)tile_off + tile_len == (uint64_t)ndata) ? n * 2u : (uint32_t)kErrStream;
    }
}

}  // namespace vbz
