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
// Geometry: a TILE is 8192 samples (16 KiB raw) of one read; tiles are numbered read-major.  A warp owns
// 1024 samples of the tile as 4 rows of 32 groups (group = row * 32 + lane): conflict-free 128-bit
// shared-memory reads, contiguous 512-byte global stores.
// Scheduling: PERSISTENT CTAs (8 warps; as many as are co-resident on the 148 SMs, launched cooperatively so that
// they ARE co-resident) walk the tiles round-robin (tile = blockIdx + k * gridDim): the tiles being worked on at
// any time are consecutive, which keeps the look-back between them short.
// Input: the TMA engine fills the input stage of tile k+1 (cp.async.bulk.shared::cluster.global + mbarrier, one
// instruction for 16 KiB) as soon as tile k's rows are in registers, so HBM latency is off the critical path and
// the compute warps issue no global load instruction at all.
// Offsets: lane -> warp by shuffle scan, warp -> tile through shared memory, tile -> read by a decoupled
// look-back over 64-bit status words (csrc/svb_kernels.cuh: lookback_exclusive).  The look-back is DEFERRED by
// one tile: a CTA posts the aggregate of tile k, and only then resolves tile k-1, whose predecessors posted theirs a
// whole tile time ago - the wait for a predecessor, which bounded the first version of these kernels, is gone.
// To make that possible encode packs tile k into shared memory at tile-local offsets (double buffered) and the
// whole CTA copies tile k-1 out once its offset is known (aligned 128-bit streaming stores fed by funnel-shifted
// 128-bit shared loads); the control bytes, whose position is known up front, leave through a TMA bulk store
// (cp.async.bulk.global.shared::cta).  Decode parks the samples of tile k - complete except for the carry of the
// tile - in a thread-private shared-memory slot and adds the carry when it stores tile k-1's samples with coalesced
// 128-bit streaming stores.
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
constexpr uint32_t kS16TileBytes = kS16TileValues * 2;              // 16384
constexpr uint32_t kS16LutWords = 96;                               // indexed by the sparse 4-code key (bits 0,2,4,6)

struct S16Ctl {           // control block in shared memory
    uint32_t wtot[8];     // per-warp data-byte totals of the current tile
    uint32_t wsum[8];     // decode: per-warp sample sums of the current tile
    uint32_t base[2][8];  // decode: sample value before each warp, tile-relative, for the two tiles in flight
    uint32_t flag;        // decode: the current tile holds 3/4-byte codes
    uint32_t carry;       // decode: look-back result (sample value before the pending tile)
    uint64_t excl;        // encode: look-back result (data bytes before the pending tile)
    uint64_t bar[3];      // mbarriers of the TMA loads
};
// shared memory carve-up (bytes)
struct S16EncSmem {
    static constexpr int kLut = 0;                                   // 96 x u32
    static constexpr int kCtl = 384;
    static constexpr int kWin = 576;                                 // look-back windows of the pending tile: kLookWindows x 32 x u64
    static constexpr int kIn = kWin + kLookWindows * 256;            // 16 B left margin + tile + 16 B slack
    static constexpr int kInBytes = 16 + kS16TileBytes + 16;
    static constexpr int kKeys = kIn + kInBytes;                     // control bytes (+16 phase)
    static constexpr int kData = kKeys + kS16TileKeys + 16;          // 2 x data bytes at tile-local offsets (+64 slack)
    static constexpr int kDataStride = kS16TileBytes + 64;
    static constexpr int kBytes = kData + 2 * kDataStride;
};
struct S16DecSmem {
    static constexpr int kLut = 0;                                   // 96 x uint2
    static constexpr int kCtl = 768;
    static constexpr int kWin = 960;                                 // look-back windows of the pending tile: kLookWindows x 32 x u64
    static constexpr int kKeys = kWin + kLookWindows * 256;          // 2 x control bytes (+16 phase)
    static constexpr int kKeysStride = kS16TileKeys + 16;
    static constexpr int kData = kKeys + 2 * kKeysStride;            // data-byte slice (+16 phase, +48 slack)
    static constexpr int kOut = kData + kS16TileBytes + 64;          // 2 x parked samples, thread-private slots
    static constexpr int kBytes = kOut + 2 * kS16TileBytes;
};
static_assert(sizeof(S16Ctl) <= 192 && S16EncSmem::kKeys % 16 == 0 && S16EncSmem::kData % 16 == 0 && S16EncSmem::kDataStride % 16 == 0 &&
                  S16DecSmem::kData % 16 == 0 && S16DecSmem::kOut % 16 == 0,
              "smem layout");

// a tile whose offset (encode) / carry (decode) is resolved one tile later; CTA-uniform, lives in registers
struct S16Pending {
    uint32_t valid = 0, t = 0, r = 0, first_tile = 0, n = 0, src_bytes = 0, total = 0, tile_off = 0, tile_len = 0, ndata = 0;
    uint64_t dst_off = 0;
};

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
// where the group owns them, single bytes at the two edges it shares with its neighbours.  Any len (ragged tiles).
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

// One warp copies `n` staged bytes to global memory: a TMA bulk store for the 16-byte aligned body (lane 0),
// byte stores for the (< 16 byte) head (lanes 1..15) and tail (lanes 16..30).  stage_at_g sits at the same
// 16-byte phase as g.  The staging area must have been fenced (fence.proxy.async) and barriered.
__device__ __forceinline__ void s16_copy_out_warp(uint8_t* g, const uint8_t* stage_at_g, uint32_t n, uint32_t lane)
{
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > n) head = n;
    const uint32_t body = (n - head) & ~15u, tail = n - head - body;
    if (lane == 0) {
        if (body) {
            vbz_bulk_store(g + head, stage_at_g + head, body);
            vbz_bulk_store_commit();
        }
    } else if (lane < 16) {
        if (lane - 1 < head) g[lane - 1] = stage_at_g[lane - 1];
    } else if (lane - 16 < tail) {
        g[head + body + (lane - 16)] = stage_at_g[head + body + (lane - 16)];
    }
}
// The whole CTA copies `n` bytes from a 16-byte aligned staging area to an arbitrarily aligned global address:
// aligned 128-bit streaming stores, each fed by two aligned 128-bit shared loads re-aligned with funnel shifts
// (the staging area must be readable 32 bytes past the end).
__device__ __forceinline__ void s16_copy_out_shifted(uint8_t* g, const uint8_t* stage, uint32_t n, uint32_t tid)
{
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > n) head = n;
    const uint32_t nchunks = (n - head) >> 4, done = head + (nchunks << 4);
    if (tid < head) g[tid] = stage[tid];
    else if (tid >= 32 && tid - 32 < n - done) g[done + tid - 32] = stage[done + tid - 32];
    const uint4* s4 = reinterpret_cast<const uint4*>(stage);  // chunk c starts at staging byte head + 16 c
    uint4* g4 = reinterpret_cast<uint4*>(g + head);
    const uint32_t wsel = head >> 2, sh = 8u * (head & 3u);   // CTA-uniform: which words of the 32-byte window, which byte phase
    for (uint32_t c = tid; c < nchunks; c += kS16Threads) {
        const uint4 A = s4[c], B = s4[c + 1];
        uint32_t w0, w1, w2, w3, w4;
        if (wsel == 0) { w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; }
        else if (wsel == 1) { w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; }
        else if (wsel == 2) { w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; }
        else { w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; }
        vbz_st_stream(g4 + c, make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh)));
    }
}

// ---- the packed arithmetic of one group of 8 samples -----------------------------------------------------------
// x[0..3]: 8 samples; prevw: the sample before them in the high half.  -> zigzag words z, control bits key.
__device__ __forceinline__ void s16_transform_group(const uint4 q, uint32_t prevw, uint32_t (&z)[4])
{
    const uint32_t x[4] = {q.x, q.y, q.z, q.w};
    const uint32_t pw[4] = {__byte_perm(prevw, x[0], 0x5432), __byte_perm(x[0], x[1], 0x5432), __byte_perm(x[1], x[2], 0x5432),
                            __byte_perm(x[2], x[3], 0x5432)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t d = vbz_sub16x2(x[i], pw[i]);                 // 16-bit wrapping delta (impl_sse3.h:434-436)
        z[i] = vbz_add16x2(d, d) ^ vbz_prmt(d, 0u, 0xBB99);          // (d << 1) ^ (d >> 15)  (impl_sse3.h:438-440)
    }
}
__device__ __forceinline__ uint32_t s16_group_key(const uint32_t (&z)[4])
{
    const uint32_t k0 = s16_flags_to_codes(s16_nonzero_bytes(__byte_perm(z[0], z[1], 0x7531)));
    const uint32_t k1 = s16_flags_to_codes(s16_nonzero_bytes(__byte_perm(z[2], z[3], 0x7531)));
    return k0 | (k1 << 8);
}
// compact the group's 8..16 data bytes into r[0..3] (zero above)
__device__ __forceinline__ void s16_compact_group(const uint32_t (&z)[4], uint32_t key, const uint32_t* s_lut, uint32_t (&r)[4])
{
    const uint32_t k0 = key & 0x55u, k1 = key >> 8;
    const uint32_t sel0 = s_lut[k0], sel1 = s_lut[k1];
    const uint32_t c0 = __byte_perm(z[0], z[1], sel0), c1 = __byte_perm(z[0], z[1], sel0 >> 16);
    const uint32_t c2 = __byte_perm(z[2], z[3], sel1), c3 = __byte_perm(z[2], z[3], sel1 >> 16);
    const uint32_t sh = 8u * __popc(k0);  // the second half starts at byte 4 + popc(k0)
    r[0] = c0;
    r[1] = c1 | __funnelshift_lc(0u, c2, sh);
    r[2] = __funnelshift_lc(c2, c3, sh);
    r[3] = __funnelshift_lc(c3, 0u, sh);
}

// =====================================================================================================
// encode
// =====================================================================================================
// The TMA load of a tile's samples (plus the 16 bytes before them, which hold the previous sample), as a
// descriptor: svb16_enc_prepare_kernel writes one per tile so that the issuing thread needs a single load.
__device__ __forceinline__ uint4 s16_enc_describe_load(const uint8_t* src, const SvbRead& rd, uint32_t tile)
{
    const uint32_t n = rd.n_values;
    if (rd.rejected || n == 0) return make_uint4(0, 0, 0, 0);
    const uint32_t ti = tile - rd.first_tile;
    const uint64_t b0 = (uint64_t)ti * kS16TileBytes, total = 2ull * n;
    uint32_t bytes = total - b0 < kS16TileBytes ? (uint32_t)(total - b0) : kS16TileBytes;
    bytes = (bytes + 15u) & ~15u;  // the slot of a read is padded to 16 bytes
    const uint32_t margin = ti ? 16u : 0u;
    const uint64_t g = (uint64_t)(uintptr_t)(src + rd.src_off + b0 - margin);
    return make_uint4((uint32_t)g, (uint32_t)(g >> 32), bytes + margin, 16u - margin);
}
__device__ __forceinline__ uint4 s16_fetch_load(const SvbLoad* d) { return *reinterpret_cast<const uint4*>(d); }
__device__ __forceinline__ void s16_store_load(SvbLoad* d, const uint4 v) { *reinterpret_cast<uint4*>(d) = v; }
__device__ __forceinline__ void s16_issue(const uint4 d, uint8_t* stage, uint64_t* bar)
{
    if (d.z == 0) return;
    vbz_mbar_expect_tx(bar, d.z);
    vbz_bulk_load(stage + d.w, reinterpret_cast<const void*>((uintptr_t)(((uint64_t)d.y << 32) | d.x)), d.z, bar);
}
// tile -> read table and the encode-side load descriptors (one warp per read)
__global__ void svb16_enc_prepare_kernel(SvbParams p)
{
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= p.n_reads) return;
    const SvbRead rd = p.reads[r];
    const uint32_t nt = s16_tiles_of(rd.n_values);
    for (uint32_t i = lane; i < nt; i += 32) s16_store_load(p.issue_a + rd.first_tile + i, s16_enc_describe_load(p.src, rd, rd.first_tile + i));
}

// the part of a tile that waits for its offset: copy the packed data bytes out, report the size
__device__ __forceinline__ void s16_encode_finish(const SvbParams& p, const S16Pending& pend, uint64_t excl, const uint8_t* stage, uint32_t tid)
{
    const uint32_t nkeys = (pend.n + 3) / 4;
    uint8_t* out = p.dst + pend.dst_off;
    s16_copy_out_shifted(out + p.hdr_bytes + nkeys + excl, stage, pend.total, tid);
    const uint32_t ti = pend.t - pend.first_tile;
    if (tid == 96 && (uint64_t)ti * kS16TileValues + kS16TileValues >= (uint64_t)pend.n)  // the read's last tile reports
        p.results[pend.r] = (uint32_t)(p.hdr_bytes + nkeys + excl + pend.total);
}
// warp 0: resolve the pending tile's offset (its predecessors posted their aggregates a tile time ago)
__device__ __forceinline__ uint64_t s16_prefetch_status(const uint64_t* status, const S16Pending& pend, uint32_t lane)
{
    const long long idx = (long long)pend.t - 1 - lane;
    return (pend.valid && idx >= (long long)pend.first_tile) ? vbz_ld_volatile_u64(status + idx) : 0ull;
}
// warp 0: start fetching every status word the pending tile's look-back can need (back to the first tile of its read, at most
// kLookWindows x 32 tiles) into shared memory, asynchronously; s16_resolve_win consumes them a phase later
__device__ __forceinline__ void s16_prefetch_windows(const uint64_t* status, const S16Pending& pend, uint32_t lane, uint64_t* win)
{
#pragma unroll
    for (int w = 0; w < kLookWindows; ++w) {
        const long long idx = (long long)pend.t - 1 - 32 * w - lane;
        if (pend.valid && idx >= (long long)pend.first_tile) vbz_cp_async8(win + w * 32 + lane, status + idx);
        else win[w * 32 + lane] = 0;
    }
    vbz_cp_async_commit();
}
__device__ __forceinline__ uint64_t s16_resolve_win(uint64_t* status, const S16Pending& pend, uint64_t mask, uint32_t lane, const uint64_t* win)
{
    vbz_cp_async_wait_all();
    __syncwarp();
    if (pend.t == pend.first_tile) return 0;  // posted as inclusive right away
    const uint64_t excl = lookback_exclusive_win(status, pend.t, pend.first_tile, win) & mask;
    if (lane == 0) vbz_st_volatile_u64(status + pend.t, kStInclusive | ((excl + pend.total) & mask));
    return excl;
}
__device__ __forceinline__ uint64_t s16_resolve(uint64_t* status, const S16Pending& pend, uint64_t mask, uint32_t lane, uint64_t pf = 0, bool have_pf = false)
{
    if (pend.t == pend.first_tile) return 0;  // posted as inclusive right away
    const uint64_t excl = lookback_exclusive_pf(status, pend.t, pend.first_tile, pf, have_pf) & mask;
    if (lane == 0) vbz_st_volatile_u64(status + pend.t, kStInclusive | ((excl + pend.total) & mask));
    return excl;
}

template <bool FULL>
__device__ __forceinline__ void s16_encode_tile(const SvbParams& p, const SvbRead& rd, uint32_t r, uint32_t t, uint32_t ti, unsigned char* smem,
                                                 S16Pending& pend, uint32_t sbuf)
{
    const uint32_t* s_lut = reinterpret_cast<const uint32_t*>(smem + S16EncSmem::kLut);
    S16Ctl* ctl = reinterpret_cast<S16Ctl*>(smem + S16EncSmem::kCtl);
    const uint8_t* in = smem + S16EncSmem::kIn + 16;
    uint8_t* s_keys = smem + S16EncSmem::kKeys;
    uint8_t* s_data = smem + S16EncSmem::kData + sbuf * S16EncSmem::kDataStride;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t n = rd.n_values, tv0 = ti * kS16TileValues;
    const uint32_t tile_values = FULL ? kS16TileValues : n - tv0;
    const uint32_t local0 = wid * kS16WarpValues + lane * 8;  // this lane's first value in row 0, tile-relative
    // warp 2 fetches the next tile's load descriptor now and issues the load once this tile's rows are in registers
    uint4 next_load = make_uint4(0, 0, 0, 0);
    if (wid == 2 && t + gridDim.x < p.n_tiles) next_load = s16_fetch_load(p.issue_a + (t + gridDim.x));
    // warp 0 starts fetching the status words of the pending tile's whole look-back; they are consumed after barrier (1)
    uint64_t* win = reinterpret_cast<uint64_t*>(smem + S16EncSmem::kWin);
    if (wid == 0) s16_prefetch_windows(p.status_a, pend, lane, win);

    // ---- rows: shared -> registers, delta, zigzag, codes -------------------------------------------------------
    uint32_t z[kS16Rows][4], key[kS16Rows], len[kS16Rows], cnt[kS16Rows];
    {
        uint32_t prev_row_w = 0;
        // lane 0 of row 0 continues from the sample before the warp's first one (zero at the start of a read)
        const uint32_t carry_w = (wid == 0 && ti == 0) ? 0u : (uint32_t)(*reinterpret_cast<const uint16_t*>(in + 2 * wid * kS16WarpValues - 2)) << 16;
#pragma unroll
        for (int j = 0; j < kS16Rows; ++j) {
            const uint32_t l0 = local0 + j * 256;
            cnt[j] = FULL ? 8u : (l0 >= tile_values ? 0u : (tile_values - l0 < 8u ? tile_values - l0 : 8u));
            const uint4 q = *reinterpret_cast<const uint4*>(in + 2 * l0);
            const uint32_t send = lane == 31 ? prev_row_w : q.w;  // lane 31 hands its previous row's last word to lane 0
            const uint32_t recv = __shfl_sync(0xffffffffu, send, (lane + 31) & 31);
            prev_row_w = q.w;
            s16_transform_group(q, (lane == 0 && j == 0) ? carry_w : recv, z[j]);
            if (!FULL && cnt[j] < 8) {  // the read's last, partial group: samples past the end contribute nothing
#pragma unroll
                for (int i = 0; i < 4; ++i) z[j][i] &= cnt[j] > 2u * i + 1 ? 0xFFFFFFFFu : (cnt[j] > 2u * i ? 0x0000FFFFu : 0u);
            }
            key[j] = s16_group_key(z[j]);
            len[j] = cnt[j] + __popc(key[j]);
        }
    }

    // ---- offsets inside the tile: lane -> warp (packed shuffle scan) -> tile (shared) --------------------------------
    const uint32_t ia = warp_scan_u32(len[0] | (len[1] << 16)), ib = warp_scan_u32(len[2] | (len[3] << 16));
    const uint32_t ta = __shfl_sync(0xffffffffu, ia, 31), tb = __shfl_sync(0xffffffffu, ib, 31);
    const uint32_t t0 = ta & 0xFFFFu, t1 = ta >> 16, t2 = tb & 0xFFFFu, t3 = tb >> 16;
    uint32_t off[kS16Rows];
    off[0] = (ia & 0xFFFFu) - len[0];
    off[1] = t0 + (ia >> 16) - len[1];
    off[2] = t0 + t1 + (ib & 0xFFFFu) - len[2];
    off[3] = t0 + t1 + t2 + (ib >> 16) - len[3];
    if (lane == 0) {
        ctl->wtot[wid] = t0 + t1 + t2 + t3;
        if (wid == 0) vbz_bulk_store_wait_read();  // the previous tile's control bytes must have left their staging area
    }
    __syncthreads();  // (1) warp totals posted, input stage consumed
    if (tid == 64) s16_issue(next_load, smem + S16EncSmem::kIn, &ctl->bar[0]);
    const uint32_t wt = ctl->wtot[lane & 7];
    const uint32_t warp_excl = __reduce_add_sync(0xffffffffu, lane < wid ? wt : 0u);
    const uint32_t tile_total = __reduce_add_sync(0xffffffffu, lane < kS16Warps ? wt : 0u);
    if (wid == 0) {
        // post this tile's aggregate FIRST, then resolve the tile before it
        if (lane == 0) vbz_st_volatile_u64(p.status_a + t, (ti == 0 ? kStInclusive : kStAggregate) | tile_total);
        if (pend.valid) {
            const uint64_t excl = s16_resolve_win(p.status_a, pend, kStValueMask, lane, win);
            if (lane == 0) ctl->excl = excl;
        }
    }

    // ---- pack: control bytes at the phase of their destination, data bytes at tile-local offsets --------------------
    uint8_t* out = p.dst + rd.dst_off;
    uint8_t* kg = out + p.hdr_bytes + (tv0 >> 2);
    uint8_t* kstage = s_keys + ((uintptr_t)kg & 15u);  // kg is even: header of 0 or 4 bytes, 16-byte aligned slots
    uint32_t tail_prev = 0;
#pragma unroll
    for (int j = 0; j < kS16Rows; ++j) {
        if (!FULL && cnt[j] == 0) continue;
        *reinterpret_cast<uint16_t*>(kstage + wid * 256 + j * 64 + lane * 2) = (uint16_t)key[j];
        uint32_t rr[4];
        s16_compact_group(z[j], key[j], s_lut, rr);
        const uint32_t P = warp_excl + off[j];
        if (!FULL) {
            s16_place(s_data, P, len[j], rr[0], rr[1], rr[2], rr[3]);
        } else {
            // every group has 8..16 bytes: the word that holds its last bytes also holds the first bytes of the next
            // group, which receives them by shuffle and stores the merged word - whole-word stores only
            const uint32_t a = P & 3u, sh = a * 8u, end = a + len[j], qe = end >> 2;
            const uint32_t s0 = rr[0] << sh;
            const uint32_t s1 = __funnelshift_l(rr[0], rr[1], sh), s2 = __funnelshift_l(rr[1], rr[2], sh), s3 = __funnelshift_l(rr[2], rr[3], sh);
            const uint32_t s4 = __funnelshift_l(rr[3], 0u, sh);
            const uint32_t tailw = qe == 2 ? s2 : (qe == 3 ? s3 : s4);  // zero if the group ends on a word boundary
            const uint32_t recv = __shfl_sync(0xffffffffu, lane == 31 ? tail_prev : tailw, (lane + 31) & 31);
            tail_prev = tailw;
            uint32_t* w = reinterpret_cast<uint32_t*>(s_data + (P & ~3u));
            if (j == 0 && lane == 0) {  // first group of the warp: the bytes before it belong to another warp
                if (a == 0) w[0] = s0;
                else {
                    uint8_t* b = reinterpret_cast<uint8_t*>(w);
                    if (a <= 1) b[1] = (uint8_t)(s0 >> 8);
                    if (a <= 2) b[2] = (uint8_t)(s0 >> 16);
                    b[3] = (uint8_t)(s0 >> 24);
                }
            } else {
                w[0] = s0 | recv;
            }
            w[1] = s1;
            if (end >= 12) w[2] = s2;
            if (end >= 16) w[3] = s3;
            if (j == kS16Rows - 1 && lane == 31) {  // last group of the warp: nobody merges its trailing bytes
                uint8_t* b = reinterpret_cast<uint8_t*>(w + qe);
                const uint32_t re = end & 3u;
                if (re > 0) b[0] = (uint8_t)tailw;
                if (re > 1) b[1] = (uint8_t)(tailw >> 8);
                if (re > 2) b[2] = (uint8_t)(tailw >> 16);
            }
        }
    }
    vbz_fence_async_smem();
    __syncthreads();  // (2) both staging areas complete, the pending tile's offset known

    // ---- out: this tile's control bytes (TMA bulk store), the PENDING tile's data bytes (whole CTA) ------------------
    if (wid == 0) s16_copy_out_warp(kg, kstage, (tile_values + 3) >> 2, lane);
    if (tid == 96 && ti == 0 && p.hdr_bytes) *reinterpret_cast<uint32_t*>(out) = rd.src_bytes;
    if (pend.valid) s16_encode_finish(p, pend, ctl->excl, smem + S16EncSmem::kData + (sbuf ^ 1u) * S16EncSmem::kDataStride, tid);
    pend.valid = 1; pend.t = t; pend.r = r; pend.first_tile = rd.first_tile; pend.n = n; pend.src_bytes = rd.src_bytes;
    pend.total = tile_total; pend.dst_off = rd.dst_off;
}

__global__ void __launch_bounds__(kS16Threads, 4) svb16_encode_kernel(SvbParams p)
{
    VBZ_DYN_SMEM(smem);
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(smem + S16EncSmem::kLut);
    S16Ctl* ctl = reinterpret_cast<S16Ctl*>(smem + S16EncSmem::kCtl);
    const uint32_t tid = threadIdx.x;
    if (tid < 86 && !(tid & 0xAAu)) s_lut[tid] = s16_make_enc_sel(tid);
    if (tid == 0) {
        vbz_mbar_init(&ctl->bar[0], 1);
        s16_issue(s16_fetch_load(p.issue_a + (blockIdx.x)), smem + S16EncSmem::kIn, &ctl->bar[0]);
    }
    __syncthreads();
    S16Pending pend;
    uint32_t phase = 0, sbuf = 0;
    for (uint32_t cur = blockIdx.x; cur < p.n_tiles; cur += gridDim.x) {
        const uint32_t r = p.tile_read[cur];
        const SvbRead rd = p.reads[r];
        const uint32_t ti = cur - rd.first_tile;
        if (rd.rejected || rd.n_values == 0) {  // decided on the host, or an empty read (one empty tile): nothing was loaded
            if (tid == 0) {
                if (rd.rejected) p.results[r] = rd.preset;
                else {
                    if (p.hdr_bytes) *reinterpret_cast<uint32_t*>(p.dst + rd.dst_off) = rd.src_bytes;
                    p.results[r] = p.hdr_bytes;
                }
            }
            __syncthreads();  // nobody is still reading the input stage
            if (tid == 0 && cur + gridDim.x < p.n_tiles) s16_issue(s16_fetch_load(p.issue_a + (cur + gridDim.x)), smem + S16EncSmem::kIn, &ctl->bar[0]);
            continue;
        }
        vbz_mbar_wait(&ctl->bar[0], phase);
        phase ^= 1u;
        if ((uint64_t)ti * kS16TileValues + kS16TileValues <= (uint64_t)rd.n_values) s16_encode_tile<true>(p, rd, r, cur, ti, smem, pend, sbuf);
        else s16_encode_tile<false>(p, rd, r, cur, ti, smem, pend, sbuf);
        sbuf ^= 1u;
    }
    if (pend.valid) {  // flush the last tile
        __syncthreads();
        if (tid < 32) {
            const uint64_t excl = s16_resolve(p.status_a, pend, kStValueMask, tid);
            if (tid == 0) ctl->excl = excl;
        }
        __syncthreads();
        s16_encode_finish(p, pend, ctl->excl, smem + S16EncSmem::kData + (sbuf ^ 1u) * S16EncSmem::kDataStride, tid);
    }
    if (tid == 0) vbz_bulk_store_wait_read();  // shared memory must outlive the last bulk store
}

struct S16DecTile {
    uint32_t n, nkeys, ndata, ti, tv0, tile_values, tile_off, tile_len, kphase, dphase;
    const uint8_t* kg;     // the tile's control bytes
    const uint8_t* dbase;  // the read's data bytes
    bool active, short_data;
    uint32_t early;        // result of an inactive read
};
// what every tile of a read agrees on, and this tile's slices (the same function tells the TMA issuer what to
// load and the consumers what to expect)
__device__ __forceinline__ S16DecTile s16_dec_describe(const SvbParams& p, const SvbRead* rd, uint32_t r, uint32_t t)
{
    S16DecTile d;
    d.n = rd->n_values;
    const uint32_t src_bytes = p.src_sizes ? p.src_sizes[r] : rd->src_bytes;
    d.nkeys = (d.n + 3) / 4;
    d.ti = t - rd->first_tile;
    d.active = true;
    d.early = 0;
    if (p.src_sizes && src_bytes >= kErrOutOfMemory) { d.early = src_bytes; d.active = false; }  // the zstd stage failed first
    else if (rd->rejected) { d.early = rd->preset; d.active = false; }
    else if (d.n == 0) { d.early = 0; d.active = false; }                               // impl_sse3.h:473-476
    else if (d.nkeys > src_bytes) { d.early = kErrInputSize; d.active = false; }         // impl_sse3.h:479-482
    if (!d.active) return d;
    d.ndata = src_bytes - d.nkeys;
    d.tv0 = d.ti * kS16TileValues;
    d.tile_values = d.n - d.tv0 < kS16TileValues ? d.n - d.tv0 : kS16TileValues;
    d.tile_off = p.wtile_off[t];                        // data bytes of this read before the tile
    d.tile_len = p.wtile_off[(uint64_t)p.n_tiles + t];  // data bytes the tile's control bytes ask for
    d.kg = p.src + rd->src_off + (d.tv0 >> 2);
    d.dbase = p.src + rd->src_off + d.nkeys;
    // the packed path needs the tile's data bytes in shared memory: all present, at most 2 per sample
    d.short_data = (uint64_t)d.tile_off + d.tile_len > (uint64_t)d.ndata || d.tile_len > 2u * d.tile_values;
    d.kphase = (uint32_t)((uintptr_t)d.kg & 15u);
    d.dphase = (uint32_t)((uintptr_t)(d.dbase + d.tile_off) & 15u);
    return d;
}
// =====================================================================================================
// decode, pre-pass 1: sum of the data-byte lengths of every tile, from the control bytes alone.
// One warp per tile (2 KiB of control bytes).
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
        // 2 KiB of control bytes per tile = 16 words per lane: all sixteen loads are issued before the first one is used
        uint32_t word[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t b = kb0 + (lane + 32u * i) * 4u;
            word[i] = 0;
            if (b + 4 <= kb1) word[i] = *reinterpret_cast<const uint32_t*>(keys + b);
            else if (b < kb1)
                for (uint32_t k = b; k < kb1; ++k) word[i] |= (uint32_t)keys[k] << (8 * (k - b));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t b = kb0 + (lane + 32u * i) * 4u;
            if (b >= kb1) continue;
            uint32_t w = word[i];
            const uint32_t vleft = n - 4 * b;  // values from this word on (>= 1)
            if (vleft < 16) w &= (1u << (2 * vleft)) - 1u;
            sum += __popc(w & 0x55555555u) + 2 * __popc(w & 0xAAAAAAAAu);
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
    __syncwarp();
    // the decode-side TMA load descriptors of every tile: control bytes (issue_a), data-byte slice (issue_b)
    for (uint32_t i = lane; i < nt; i += 32) {
        const uint32_t t = rd.first_tile + i;
        const S16DecTile d = s16_dec_describe(p, &rd, r, t);
        uint4 ka = make_uint4(0, 0, 0, 0), da = make_uint4(0, 0, 0, 0);
        if (d.active) {
            const uint64_t kgp = (uint64_t)(uintptr_t)(d.kg - d.kphase);
            ka = make_uint4((uint32_t)kgp, (uint32_t)(kgp >> 32), (d.kphase + ((d.tile_values + 3) >> 2) + 15u) & ~15u, 0u);
            if (!d.short_data) {
                const uint64_t dgp = (uint64_t)(uintptr_t)(d.dbase + d.tile_off - d.dphase);
                da = make_uint4((uint32_t)dgp, (uint32_t)(dgp >> 32), (d.dphase + d.tile_len + 15u) & ~15u, 0u);
            }
        }
        s16_store_load(p.issue_a + t, ka);
        s16_store_load(p.issue_b + t, da);
    }
}

// =====================================================================================================
// decode
// =====================================================================================================
// the part of a tile that waits for its carry: add it to the parked samples and store them
__device__ __forceinline__ void s16_decode_finish(const SvbParams& p, const S16Pending& pend, uint32_t carry, const uint8_t* park, uint32_t tid)
{
    const uint32_t lane = tid & 31, wid = tid >> 5;
    const uint32_t ti = pend.t - pend.first_tile, tv0 = ti * kS16TileValues;
    const uint32_t tile_values = pend.n - tv0 < kS16TileValues ? pend.n - tv0 : kS16TileValues;
    const uint32_t c = __byte_perm(carry, 0u, 0x1010);
    const uint32_t local0 = wid * kS16WarpValues + lane * 8;
    uint8_t* outp = p.dst + pend.dst_off + 2ull * tv0;
#pragma unroll
    for (int j = 0; j < kS16Rows; ++j) {
        const uint32_t l0 = local0 + j * 256;
        if (l0 >= tile_values) continue;
        const uint32_t cnt = tile_values - l0 < 8u ? tile_values - l0 : 8u;
        uint4 v = *reinterpret_cast<const uint4*>(park + 2 * l0);
        v = make_uint4(vbz_add16x2(v.x, c), vbz_add16x2(v.y, c), vbz_add16x2(v.z, c), vbz_add16x2(v.w, c));
        uint8_t* gp = outp + 2 * l0;
        if (cnt == 8) vbz_st_stream(reinterpret_cast<uint4*>(gp), v);
        else {
            uint16_t* g16 = reinterpret_cast<uint16_t*>(gp);
            g16[0] = (uint16_t)v.x;
            if (cnt > 1) g16[1] = (uint16_t)(v.x >> 16);
            if (cnt > 2) g16[2] = (uint16_t)v.y;
            if (cnt > 3) g16[3] = (uint16_t)(v.y >> 16);
            if (cnt > 4) g16[4] = (uint16_t)v.z;
            if (cnt > 5) g16[5] = (uint16_t)(v.z >> 16);
            if (cnt > 6) g16[6] = (uint16_t)v.w;
        }
    }
    if (tid == 0 && tv0 + tile_values == pend.n) {
        // all data bytes must be consumed exactly: impl_sse3.h:574-577, 638-642
        p.results[pend.r] = ((uint64_t)pend.tile_off + pend.tile_len == (uint64_t)pend.ndata) ? pend.n * 2u : (uint32_t)kErrStream;
    }
}

// Exact scalar semantics for tiles the packed path cannot express: codes 2/3 in the stream (the x86 build then
// decodes 8-value groups 16 bits wide while 32 data bytes remain and 32 bits wide in its scalar tail,
// impl_sse3.h:494-572) or a stream that ends early (bytes past the end read as zero).  One lane's 4 groups;
// the in-group running sums go to slot[row * 512 ..] (two samples per word, as the packed path leaves them).
__device__ __noinline__ void s16_decode_rows_scalar(const uint8_t* dbase, uint32_t ndata, uint32_t n, uint32_t s_first, uint64_t o_base, uint32_t k0,
                                                    uint32_t k1, uint32_t k2, uint32_t k3, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t o0,
                                                    uint32_t o1, uint32_t o2, uint32_t o3, uint8_t* slot)
{
#pragma unroll 1
    for (int j = 0; j < kS16Rows; ++j) {
        const uint32_t key = j == 0 ? k0 : j == 1 ? k1 : j == 2 ? k2 : k3;
        const uint32_t cnt = j == 0 ? c0 : j == 1 ? c1 : j == 2 ? c2 : c3;
        const uint32_t s0 = s_first + j * 256;
        uint64_t o = o_base + (j == 0 ? o0 : j == 1 ? o1 : j == 2 ? o2 : o3);
        const bool simd16 = (s0 >> 3) < (n >> 3) && o + 32 <= (uint64_t)ndata;
        uint32_t run = 0, prevv = 0;
        uint32_t* out = reinterpret_cast<uint32_t*>(slot + j * 512);
#pragma unroll 1
        for (uint32_t k = 0; k < 8; ++k) {
            uint32_t v = 0;
            if (k < cnt) {
                const uint32_t code = (key >> (2 * k)) & 3u;
                for (uint32_t b = 0; b <= code; ++b, ++o) v |= (o < ndata ? (uint32_t)dbase[o] : 0u) << (8 * b);
            }
            if (simd16) v &= 0xFFFFu;
            run += (v >> 1) ^ (0u - (v & 1u));
            if (k & 1) out[k >> 1] = (prevv & 0xFFFFu) | (run << 16);
            prevv = run;
        }
    }
}

template <bool FULL>
__device__ __forceinline__ void s16_decode_tile(const SvbParams& p, const SvbRead& rd, const S16DecTile& d, uint32_t r, uint32_t t, unsigned char* smem,
                                                 S16Pending& pend, uint32_t sbuf, uint32_t data_phase)
{
    const uint2* s_lut = reinterpret_cast<const uint2*>(smem + S16DecSmem::kLut);
    S16Ctl* ctl = reinterpret_cast<S16Ctl*>(smem + S16DecSmem::kCtl);
    const uint8_t* s_keys = smem + S16DecSmem::kKeys + sbuf * S16DecSmem::kKeysStride;
    const uint8_t* s_data = smem + S16DecSmem::kData;
    uint8_t* park = smem + S16DecSmem::kOut + sbuf * kS16TileBytes;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t n = d.n, ndata = d.ndata;
    const uint32_t local0 = wid * kS16WarpValues + lane * 8;
    // warp 2 fetches the next tile's load descriptors now; the loads are issued as the stages become free
    uint4 next_keys = make_uint4(0, 0, 0, 0), next_data = make_uint4(0, 0, 0, 0);
    if (wid == 2 && t + gridDim.x < p.n_tiles) { next_keys = s16_fetch_load(p.issue_a + (t + gridDim.x)); next_data = s16_fetch_load(p.issue_b + (t + gridDim.x)); }
    // warp 0 starts fetching the status words of the pending tile's whole look-back; they are consumed after barrier (2)
    uint64_t* win = reinterpret_cast<uint64_t*>(smem + S16DecSmem::kWin);
    if (wid == 0) s16_prefetch_windows(p.status_b, pend, lane, win);

    // ---- control bytes -> lengths -> offsets ------------------------------------------------------------------
    uint32_t key[kS16Rows], len[kS16Rows], cnt[kS16Rows];
    uint32_t wide = 0;
#pragma unroll
    for (int j = 0; j < kS16Rows; ++j) {
        const uint32_t l0 = local0 + j * 256;
        cnt[j] = FULL ? 8u : (l0 >= d.tile_values ? 0u : (d.tile_values - l0 < 8u ? d.tile_values - l0 : 8u));
        uint32_t kb = *reinterpret_cast<const uint16_t*>(s_keys + d.kphase + wid * 256 + j * 64 + lane * 2);
        if (!FULL && cnt[j] < 8) kb &= (1u << (2 * cnt[j])) - 1u;  // bits of values past the end are ignored by the reference
        key[j] = kb;
        wide |= kb;
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
    const bool any_wide = __any_sync(0xffffffffu, (wide & 0xAAAAu) != 0);  // 3- or 4-byte codes: not an x86-int16 stream
    if (lane == 0) {
        ctl->wtot[wid] = t0 + t1 + t2 + t3;
        if (any_wide) ctl->flag = 1;
    }
    __syncthreads();  // (1) warp totals, flag; this tile's control-byte stage is consumed
    if (tid == 64) s16_issue(next_keys, smem + S16DecSmem::kKeys + (sbuf ^ 1u) * S16DecSmem::kKeysStride, &ctl->bar[sbuf ^ 1u]);
    const bool slow = d.short_data || ctl->flag != 0;
    const uint32_t warp_excl = __reduce_add_sync(0xffffffffu, lane < wid ? ctl->wtot[lane & 7] : 0u);

    // ---- unpack + un-zigzag + prefix sum inside the group (two samples per 32-bit word) ---------------------------
    uint32_t y[kS16Rows][4];
    if (!slow) {
        vbz_mbar_wait(&ctl->bar[2], data_phase);
#pragma unroll
        for (int j = 0; j < kS16Rows; ++j) {
            const uint32_t P = d.dphase + warp_excl + off[j];
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
                const uint32_t dd = ((zz[i] >> 1) & 0x7FFF7FFFu) ^ vbz_prmt(zz[i] << 7, 0u, 0xAA88);
                y[j][i] = vbz_add16x2(vbz_add16x2(dd, dd << 16), run);
                run = __byte_perm(y[j][i], 0u, 0x3232);
            }
        }
    } else {
        if (!d.short_data) vbz_mbar_wait(&ctl->bar[2], data_phase);  // consume the phase of the load that was issued
        // rare: through a call, via the thread's slot of the park area, to keep it out of the packed path's registers
        s16_decode_rows_scalar(d.dbase, ndata, n, d.tv0 + local0, (uint64_t)d.tile_off + warp_excl, key[0], key[1], key[2], key[3], cnt[0], cnt[1],
                               cnt[2], cnt[3], off[0], off[1], off[2], off[3], park + 2 * local0);
#pragma unroll
        for (int j = 0; j < kS16Rows; ++j) {
            const uint4 v = *reinterpret_cast<const uint4*>(park + 2 * (local0 + j * 256));
            y[j][0] = v.x; y[j][1] = v.y; y[j][2] = v.z; y[j][3] = v.w;
        }
    }

    // ---- running sample value inside the warp (packed 16-bit scan); park the samples ---------------------------------
    const uint32_t la = __byte_perm(y[0][3], y[1][3], 0x7632), lb = __byte_perm(y[2][3], y[3][3], 0x7632);  // row totals, two per word
    uint32_t sa = la, sb = lb;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const uint32_t xa = __shfl_up_sync(0xffffffffu, sa, dd), xb = __shfl_up_sync(0xffffffffu, sb, dd);
        if (lane >= (uint32_t)dd) { sa = vbz_add16x2(sa, xa); sb = vbz_add16x2(sb, xb); }
    }
    const uint32_t wa = __shfl_sync(0xffffffffu, sa, 31), wb = __shfl_sync(0xffffffffu, sb, 31);
    const uint32_t r0t = wa & 0xFFFFu, r1t = wa >> 16, r2t = wb & 0xFFFFu, r3t = wb >> 16;
    if (lane == 0) ctl->wsum[wid] = (r0t + r1t + r2t + r3t) & 0xFFFFu;
    {   // exclusive carry of each row inside the warp: the rows before it, the lanes before it
        const uint32_t ea = vbz_sub16x2(sa, la), eb = vbz_sub16x2(sb, lb);
        const uint32_t carry[kS16Rows] = {ea & 0xFFFFu, r0t + (ea >> 16), r0t + r1t + (eb & 0xFFFFu), r0t + r1t + r2t + (eb >> 16)};
#pragma unroll
        for (int j = 0; j < kS16Rows; ++j) {
            const uint32_t c = __byte_perm(carry[j], 0u, 0x1010);
            *reinterpret_cast<uint4*>(park + 2 * (local0 + j * 256)) =
                make_uint4(vbz_add16x2(y[j][0], c), vbz_add16x2(y[j][1], c), vbz_add16x2(y[j][2], c), vbz_add16x2(y[j][3], c));
        }
    }
    __syncthreads();  // (2) warp sums; the data-byte stage is consumed
    if (tid == 64) s16_issue(next_data, smem + S16DecSmem::kData, &ctl->bar[2]);
    if (wid == 0) {
        const uint32_t v = lane < kS16Warps ? ctl->wsum[lane] : 0u;
        const uint32_t incl = warp_scan_u32(v);
        const uint32_t tile_sum = __shfl_sync(0xffffffffu, incl, 31) & 0xFFFFu;
        if (lane < kS16Warps) ctl->base[sbuf][lane] = incl - v;  // sample value before the warp's first sample, tile-relative
        // post this tile's aggregate FIRST, then resolve the tile before it
        if (lane == 0) vbz_st_volatile_u64(p.status_b + t, (d.ti == 0 ? kStInclusive : kStAggregate) | tile_sum);
        if (pend.valid) {
            const uint32_t carry = (uint32_t)s16_resolve_win(p.status_b, pend, 0xFFFFull, lane, win);
            if (lane == 0) ctl->carry = carry;
        }
        if (lane == 0) ctl->flag = 0;
        pend.total = tile_sum;  // (warp 0's copy is the one s16_resolve uses)
    }
    __syncthreads();  // (3) the pending tile's carry is known
    if (pend.valid) s16_decode_finish(p, pend, ctl->carry + ctl->base[sbuf ^ 1u][wid], smem + S16DecSmem::kOut + (sbuf ^ 1u) * kS16TileBytes, tid);
    const uint32_t keep_total = pend.total;
    pend.valid = 1; pend.t = t; pend.r = r; pend.first_tile = rd.first_tile; pend.n = n; pend.src_bytes = rd.src_bytes;
    pend.total = keep_total; pend.dst_off = rd.dst_off; pend.tile_off = d.tile_off; pend.tile_len = d.tile_len; pend.ndata = ndata;
}

__global__ void __launch_bounds__(kS16Threads, 4) svb16_decode_kernel(SvbParams p)
{
    VBZ_DYN_SMEM(smem);
    uint2* s_lut = reinterpret_cast<uint2*>(smem + S16DecSmem::kLut);
    S16Ctl* ctl = reinterpret_cast<S16Ctl*>(smem + S16DecSmem::kCtl);
    const uint32_t tid = threadIdx.x;
    if (tid < 86 && !(tid & 0xAAu)) s_lut[tid] = s16_make_dec_sel(tid);
    if (tid == 0) {
        vbz_mbar_init(&ctl->bar[0], 1);
        vbz_mbar_init(&ctl->bar[1], 1);
        vbz_mbar_init(&ctl->bar[2], 1);
        ctl->flag = 0;
        s16_issue(s16_fetch_load(p.issue_a + (blockIdx.x)), smem + S16DecSmem::kKeys, &ctl->bar[0]);
        s16_issue(s16_fetch_load(p.issue_b + (blockIdx.x)), smem + S16DecSmem::kData, &ctl->bar[2]);
    }
    __syncthreads();
    S16Pending pend;
    uint32_t sbuf = 0, key_phases = 0, data_phase = 0;
    for (uint32_t cur = blockIdx.x; cur < p.n_tiles; cur += gridDim.x) {
        const uint32_t r = p.tile_read[cur];
        const SvbRead rd = p.reads[r];
        const S16DecTile d = s16_dec_describe(p, &rd, r, cur);
        if (!d.active) {  // whole-read rejections (every tile of the read takes this branch): nothing was loaded
            if (tid == 0 && d.ti == 0) p.results[r] = d.early;
            __syncthreads();  // nobody is still reading the stages
            if (tid == 0 && cur + gridDim.x < p.n_tiles) {
                s16_issue(s16_fetch_load(p.issue_a + (cur + gridDim.x)), smem + S16DecSmem::kKeys + sbuf * S16DecSmem::kKeysStride, &ctl->bar[sbuf]);
                s16_issue(s16_fetch_load(p.issue_b + (cur + gridDim.x)), smem + S16DecSmem::kData, &ctl->bar[2]);
            }
            continue;
        }
        vbz_mbar_wait(&ctl->bar[sbuf], (key_phases >> sbuf) & 1u);
        key_phases ^= 1u << sbuf;
        if (d.tile_values == kS16TileValues) s16_decode_tile<true>(p, rd, d, r, cur, smem, pend, sbuf, data_phase);
        else s16_decode_tile<false>(p, rd, d, r, cur, smem, pend, sbuf, data_phase);
        if (!d.short_data) data_phase ^= 1u;
        sbuf ^= 1u;
    }
    if (pend.valid) {  // flush the last tile
        __syncthreads();
        if (tid < 32) {
            const uint32_t carry = (uint32_t)s16_resolve(p.status_b, pend, 0xFFFFull, tid);
            if (tid == 0) ctl->carry = carry;
        }
        __syncthreads();
        s16_decode_finish(p, pend, ctl->carry + ctl->base[sbuf ^ 1u][tid >> 5], smem + S16DecSmem::kOut + (sbuf ^ 1u) * kS16TileBytes, tid);
    }
}

}  // namespace vbz
