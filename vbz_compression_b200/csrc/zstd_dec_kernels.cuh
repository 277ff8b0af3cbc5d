// csrc/zstd_dec_kernels.cuh - stage 2 of VBZ decode on the GPU: Zstandard frame decoder (RFC 8878).
//
// Replaces the reference's calls into libzstd on the decode path: ZSTD_getFrameContentSize + ZSTD_decompress
// (ref vbz/vbz.cpp:236-240, 263-272; libzstd itself is an un-vendored dependency, pinned zstd/1.4.8 by
// CMakeLists.txt:93).  Written from the format specification; accepts everything ZSTD_compress emits at any
// level: Raw / RLE / Compressed blocks, Raw / RLE / Huffman / Treeless literals with 1 or 4 streams, direct and
// FSE-compressed Huffman weights, Predefined / RLE / FSE / Repeat sequence tables, repeat offsets, windowed and
// single-segment frames, concatenated and skippable frames, the optional XXH64 content checksum.
//
// Parallelism: one WARP per frame (one frame = one read), frames taken from an atomic counter by persistent
// warps - a batch holds 10^4..10^6 frames, far more than the 148 SMs have warp slots.  Inside a frame:
//   * the four Huffman literal streams of a block are decoded by four lanes at once, table in shared memory;
//     a block without sequences (all but the first block of a signal frame) decodes straight into its final
//     position in the output;
//   * sequences are decoded 32 at a time by one lane into shared memory and then executed by the whole warp:
//     literal runs and matches are warp-wide copies (a match with offset < 32 is a periodic fill, computed
//     without waiting on the bytes it produces);
//   * Raw / RLE blocks, table builds are warp-wide loops.
// Per-warp tables: Huffman 2048 x u16, three FSE tables 512/512/256 x u32 -> 10.5 KiB of shared memory.
#pragma once
#include "vbz_device.h"
#include "svb_types.h"
#ifdef VBZ_CPUSIM
#include <stdio.h>
#include <stdlib.h>
#endif

namespace vbz {


struct ZWarp {  // per-warp shared state
    uint16_t huf[2048];   // sym | nbBits << 8
    uint32_t ll[512], ml[512], of[256];  // sym | nbBits << 8 | base << 16
    uint32_t seq_ll[32], seq_ml[32], seq_of[32];
    uint32_t wtab[64];    // FSE table of the Huffman weights (accuracy log <= 6)
    uint8_t weights[256];
    int16_t norm[256];
    uint16_t next[256];
};

// ---- small readers over global memory ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t z_le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ uint32_t z_le24(const uint8_t* p) { return z_le16(p) | ((uint32_t)p[2] << 16); }
__device__ __forceinline__ uint32_t z_le32(const uint8_t* p) { return z_le24(p) | ((uint32_t)p[3] << 24); }
__device__ __forceinline__ uint64_t z_le64(const uint8_t* p) { return (uint64_t)z_le32(p) | ((uint64_t)z_le32(p + 4) << 32); }
__device__ __forceinline__ int z_highbit(uint32_t v) { return 31 - __clz((int)v); }

// backward bit reader: the stream is a little-endian big integer whose highest set bit is a padding marker;
// bits are consumed from the top.  `pos` = number of unread bits; reading below bit 0 yields zeros.
struct ZBack {
    const uint8_t* p;
    long long pos;
};
__device__ __forceinline__ bool zb_init(ZBack& b, const uint8_t* p, uint32_t n)
{
    if (n == 0 || p[n - 1] == 0) return false;
    b.p = p;
    b.pos = (long long)(n - 1) * 8 + z_highbit(p[n - 1]);
    return true;
}
// the 32 bits ending just below bit `pos` (bit 31 of the result = bit pos-1); bits below 0 read as zero
__device__ __forceinline__ uint32_t zb_window(const ZBack& b)
{
    const long long lo = b.pos - 32;  // bit index of result bit 0
    if (lo >= 0) {
        const uint8_t* q = b.p + (lo >> 3);
        const uint32_t sh = (uint32_t)(lo & 7);
        const uint64_t v = (uint64_t)z_le32(q) | ((uint64_t)(sh ? q[4] : 0) << 32);
        return (uint32_t)(v >> sh);
    }
    if (b.pos <= 0) return 0;
    // fewer than 32 bits left: assemble from the first bytes
    uint64_t v = 0;
    const uint32_t nbytes = (uint32_t)((b.pos + 7) >> 3);
    for (uint32_t i = 0; i < nbytes && i < 5; ++i) v |= (uint64_t)b.p[i] << (8 * i);
    v &= (b.pos >= 64) ? ~0ull : ((1ull << b.pos) - 1ull);
    return (uint32_t)(v << (uint32_t)(32 - b.pos));
}
__device__ __forceinline__ uint32_t zb_peek(const ZBack& b, uint32_t nb) { return nb ? zb_window(b) >> (32 - nb) : 0u; }
__device__ __forceinline__ uint32_t zb_read(ZBack& b, uint32_t nb)
{
    const uint32_t v = zb_peek(b, nb);
    b.pos -= nb;
    return v;
}

// forward, LSB-first bit reader (FSE table descriptions)
struct ZFwd {
    const uint8_t* p;
    uint32_t n, bit;
};
__device__ __forceinline__ uint32_t zf_peek(const ZFwd& f, uint32_t nb)
{
    uint64_t v = 0;
    const uint32_t byte = f.bit >> 3;
    for (uint32_t i = 0; i < 4; ++i)
        if (byte + i < f.n) v |= (uint64_t)f.p[byte + i] << (8 * i);
    return (uint32_t)(v >> (f.bit & 7)) & ((1u << nb) - 1u);
}

// ---- FSE (one lane) -----------------------------------------------------------------------------------------
// table description -> normalised counts.  Returns bytes consumed or 0 on error.
__device__ uint32_t z_fse_read_ncount(const uint8_t* p, uint32_t n, int max_al, int max_sym, int16_t* norm, int* nsym, int* al_out)
{
    ZFwd b = {p, n, 0};
    if (n == 0) return 0;
    const int al = (int)zf_peek(b, 4) + 5;
    b.bit += 4;
    if (al > max_al) return 0;
    int remaining = (1 << al) + 1, thr = 1 << al, nb = al + 1, sym = 0, prev0 = 0;
    while (remaining > 1 && sym <= max_sym) {
        if (prev0) {
            for (;;) {
                const uint32_t r = zf_peek(b, 2);
                b.bit += 2;
                for (uint32_t k = 0; k < r; ++k) {
                    if (sym > max_sym) return 0;
                    norm[sym++] = 0;
                }
                if (r != 3) break;
            }
            if (sym > max_sym) return 0;
        }
        const int max = (2 * thr - 1) - remaining;
        int c;
        const uint32_t low = zf_peek(b, (uint32_t)(nb - 1));
        if ((int)low < max) { c = (int)low; b.bit += (uint32_t)(nb - 1); }
        else {
            c = (int)zf_peek(b, (uint32_t)nb);
            if (c >= thr) c -= max;
            b.bit += (uint32_t)nb;
        }
        c -= 1;
        remaining -= c < 0 ? -c : c;
        norm[sym++] = (int16_t)c;
        prev0 = (c == 0);
        while (remaining < thr) { nb--; thr >>= 1; }
    }
    if (remaining != 1) return 0;
    const uint32_t used = (b.bit + 7) / 8;
    if (used > n) return 0;
    *nsym = sym;
    *al_out = al;
    return used;
}
// normalised counts -> decoding table (cells: sym | nb << 8 | base << 16; cell u lives at cell[u * STRIDE])
template <int STRIDE = 1>
__device__ bool z_fse_build(uint32_t* cell, uint16_t* next, const int16_t* norm, int nsym, int al)
{
    const int T = 1 << al;
    int high = T - 1;
    for (int s = 0; s < nsym; ++s) {
        if (norm[s] == -1) { cell[(high--) * STRIDE] = (uint32_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    const int step = (T >> 1) + (T >> 3) + 3;
    int pos = 0;
    for (int s = 0; s < nsym; ++s)
        for (int i = 0; i < norm[s]; ++i) {
            cell[pos * STRIDE] = (uint32_t)s;
            do { pos = (pos + step) & (T - 1); } while (pos > high);
        }
    if (pos != 0) return false;
    for (int u = 0; u < T; ++u) {
        const uint32_t s = cell[u * STRIDE] & 0xFFu;
        const uint32_t x = next[s]++;
        const uint32_t nb = (uint32_t)(al - z_highbit(x));
        cell[u * STRIDE] = s | (nb << 8) | ((((x << nb) - (uint32_t)T) & 0xFFFFu) << 16);
    }
    return true;
}

// ---- Huffman ------------------------------------------------------------------------------------------------
// tree description -> weights[0..nw) (one lane).  Returns bytes consumed or 0 on error.
__device__ uint32_t z_huf_read_weights(uint8_t* w, uint32_t* wtab, int16_t* norm, uint16_t* next, const uint8_t* p, uint32_t n, int* nw_out)
{
    int nw = 0;
    uint32_t used;
    if (n < 1) return 0;
    const int hb = p[0];
    if (hb >= 128) {
        nw = hb - 127;
        used = 1 + (uint32_t)(nw + 1) / 2;
        if (used > n) return 0;
        for (int i = 0; i < nw; ++i) w[i] = (i & 1) ? (p[1 + i / 2] & 0xF) : (p[1 + i / 2] >> 4);
    } else {
        used = 1 + (uint32_t)hb;
        if (used > n || hb == 0) return 0;
        int nsym, al;
        const uint32_t hdr = z_fse_read_ncount(p + 1, (uint32_t)hb, 6, 255, norm, &nsym, &al);
        if (!hdr) return 0;
        uint32_t* cell = wtab;
        if (!z_fse_build(cell, next, norm, nsym, al)) return 0;
        ZBack b;
        if (!zb_init(b, p + 1 + hdr, (uint32_t)hb - hdr)) return 0;
        uint32_t s1 = zb_read(b, (uint32_t)al), s2 = zb_read(b, (uint32_t)al);
        for (;;) {  // two interleaved states; stop one symbol after the stream overruns
            if (nw > 253) return 0;
            w[nw++] = (uint8_t)cell[s1];
            s1 = (cell[s1] >> 16) + zb_read(b, (cell[s1] >> 8) & 0xFF);
            if (b.pos < 0) { w[nw++] = (uint8_t)cell[s2]; break; }
            if (nw > 253) return 0;
            w[nw++] = (uint8_t)cell[s2];
            s2 = (cell[s2] >> 16) + zb_read(b, (cell[s2] >> 8) & 0xFF);
            if (b.pos < 0) { w[nw++] = (uint8_t)cell[s1]; break; }
        }
    }
    *nw_out = nw;
    return used;
}
// weights -> decode table huf (one lane).  Returns max_bits or 0 on error.
__device__ int z_huf_build(uint16_t* huf, uint8_t* w, int nw)
{
    uint32_t total = 0, rank[16];
    for (int i = 0; i < 16; ++i) rank[i] = 0;
    for (int i = 0; i < nw; ++i) {
        if (w[i] > 12) return 0;
        rank[w[i]]++;
        total += (1u << w[i]) >> 1;
    }
    if (total == 0) return 0;
    const int max_bits = z_highbit(total) + 1;
    if (max_bits > 11) return 0;  // HUF_TABLELOG_MAX of the format (larger tables are rejected by libzstd too)
    const uint32_t rest = (1u << max_bits) - total;
    if ((rest & (rest - 1)) != 0) return 0;
    w[nw] = (uint8_t)(z_highbit(rest) + 1);
    rank[w[nw]]++;
    nw++;
    if (rank[1] < 2 || (rank[1] & 1)) return 0;
    // cells: weight 1 (longest codes) first, symbols ascending within a weight
    uint32_t start[16], acc = 0;
    for (int r = 1; r <= max_bits; ++r) { start[r] = acc; acc += rank[r] << (r - 1); }
    for (int s = 0; s < nw; ++s) {
        const int r = w[s];
        if (!r) continue;
        const uint32_t len = (1u << r) >> 1;
        const uint32_t e = (uint32_t)s | ((uint32_t)(max_bits + 1 - r) << 8);
        uint16_t* c = huf + start[r];  // a run of `len` cells starts at a multiple of len: wide stores are aligned
        if (len >= 4) {
            const uint2 e4 = make_uint2(e | (e << 16), e | (e << 16));
            for (uint32_t k = 0; k < len; k += 4) *reinterpret_cast<uint2*>(c + k) = e4;
        } else {
            for (uint32_t k = 0; k < len; ++k) c[k] = (uint16_t)e;
        }
        start[r] += len;
    }
    return max_bits;
}
// one Huffman stream (one lane): `count` symbols to out.  Returns true if the stream ends exactly.
// The stream is read backwards through ALIGNED 32-bit words (the word holding the last byte first) into a 64-bit
// window whose top bit is the next unread bit; one refill test per two symbols (2 x 11 bits <= 32).  Words below
// `p & ~3` are never touched (a damaged stream reads zeros there and fails the final test), bytes of the lowest word
// that precede p can only be looked at by the padding of a final table look-up, which does not depend on them.
// Output leaves in 16-byte stores once `out` is aligned.
__device__ bool z_huf_decode_stream(const uint16_t* huf, int max_bits, const uint8_t* p, uint32_t n, uint8_t* out, uint32_t count)
{
    if (n == 0 || p[n - 1] == 0) return false;
    const uint32_t* lim = reinterpret_cast<const uint32_t*>((uintptr_t)p & ~(uintptr_t)3);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>((uintptr_t)(p + n - 1) & ~(uintptr_t)3);
    const uint32_t v0 = 8u * (uint32_t)((uintptr_t)(p + n - 1) & 3u) + (uint32_t)z_highbit(p[n - 1]);  // unread bits in the first word
    // `avail` = bits of the stream proper
    const long long avail = (long long)(n - 1) * 8 + z_highbit(p[n - 1]);
    uint64_t buf = v0 ? ((uint64_t)*wp << (64 - v0)) : 0ull;
    int nbits = (int)v0;
    --wp;
    const uint32_t* const wp0 = wp;
    uint32_t nxt = wp >= lim ? *wp : 0u;
    const uint32_t sh = 32u - (uint32_t)max_bits;
    uint32_t i = 0;
#define Z_HUF_REFILL()                                              \
    if (nbits < 32) {                                               \
        buf |= (uint64_t)nxt << (32 - nbits);                       \
        nbits += 32;                                                \
        --wp;                                                       \
        nxt = wp >= lim ? *wp : 0u;                                 \
    }
#define Z_HUF_SYM(dst_expr)                                         \
    {                                                               \
        const uint32_t e = huf[(uint32_t)(buf >> 32) >> sh];        \
        buf <<= (e >> 8);                                           \
        nbits -= (int)(e >> 8);                                     \
        dst_expr;                                                   \
    }
    // head: up to the first 16-byte boundary of out
    uint32_t head = (16u - (uint32_t)((uintptr_t)out & 15u)) & 15u;
    if (head > count) head = count;
    for (; i < head; ++i) {
        Z_HUF_REFILL();
        Z_HUF_SYM(out[i] = (uint8_t)e);
    }
    for (; i + 16 <= count; i += 16) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t acc = 0;
            Z_HUF_REFILL();
            Z_HUF_SYM(acc = e & 0xFFu);
            Z_HUF_SYM(acc |= (e & 0xFFu) << 8);
            Z_HUF_REFILL();
            Z_HUF_SYM(acc |= (e & 0xFFu) << 16);
            Z_HUF_SYM(acc |= e << 24);
            w[q] = acc;
        }
        *reinterpret_cast<uint4*>(out + i) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (; i < count; ++i) {
        Z_HUF_REFILL();
        Z_HUF_SYM(out[i] = (uint8_t)e);
    }
#undef Z_HUF_REFILL
#undef Z_HUF_SYM
    // bits handed to the window (32 per refill) minus bits still in it = bits used; the stream must end exactly
    return (long long)v0 + 32ll * (long long)(wp0 - wp) - nbits == avail;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += x;
    }
    return v;
}

// ---- warp-wide copies -------------------------------------------------------------------------------------------
__device__ __forceinline__ void z_warp_copy(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane)
{
    for (uint32_t k = lane; k < n; k += 32) dst[k] = src[k];
}
__device__ __forceinline__ void z_warp_fill(uint8_t* dst, uint8_t v, uint32_t n, uint32_t lane)
{
    for (uint32_t k = lane; k < n; k += 32) dst[k] = v;
}
// out[op .. op + mlen) = the mlen bytes starting `offset` back, overlap allowed
__device__ __forceinline__ void z_warp_match(uint8_t* out, uint64_t op, uint32_t offset, uint32_t mlen, uint32_t lane)
{
    const uint8_t* base = out + op - offset;
    if (offset >= 32) {
        for (uint32_t k0 = 0; k0 < mlen; k0 += 32) {
            const uint32_t k = k0 + lane;
            if (k < mlen) out[op + k] = base[k];
            __syncwarp();
        }
    } else {  // periodic fill: every source byte precedes op
        for (uint32_t k = lane; k < mlen; k += 32) out[op + k] = base[k % offset];
    }
}

__constant__ int16_t kZ_LL_DEF[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
__constant__ int16_t kZ_ML_DEF[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                      1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
__constant__ int16_t kZ_OF_DEF[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
__constant__ uint32_t kZ_LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
__constant__ uint8_t kZ_LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__constant__ uint32_t kZ_ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                                        30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
__constant__ uint8_t kZ_ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                       0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

// XXH64 with seed 0 (one lane): only the optional frame checksum needs it
__device__ __forceinline__ uint64_t z_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ uint64_t z_xxh64(const uint8_t* p, uint64_t n)
{
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P3 = 1609587929392839161ULL, P4 = 9650029242287828579ULL,
                   P5 = 2870177450012600261ULL;
    const uint8_t* e = p + n;
    uint64_t h;
    if (n >= 32) {
        uint64_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;
        do {
            v1 = z_rotl64(v1 + z_le64(p) * P2, 31) * P1; p += 8;
            v2 = z_rotl64(v2 + z_le64(p) * P2, 31) * P1; p += 8;
            v3 = z_rotl64(v3 + z_le64(p) * P2, 31) * P1; p += 8;
            v4 = z_rotl64(v4 + z_le64(p) * P2, 31) * P1; p += 8;
        } while (p + 32 <= e);
        h = z_rotl64(v1, 1) + z_rotl64(v2, 7) + z_rotl64(v3, 12) + z_rotl64(v4, 18);
        h = (h ^ (z_rotl64(v1 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (z_rotl64(v2 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (z_rotl64(v3 * P2, 31) * P1)) * P1 + P4;
        h = (h ^ (z_rotl64(v4 * P2, 31) * P1)) * P1 + P4;
    } else h = P5;
    h += n;
    while (p + 8 <= e) { h ^= z_rotl64(z_le64(p) * P2, 31) * P1; h = z_rotl64(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= e) { h ^= (uint64_t)z_le32(p) * P1; h = z_rotl64(h, 23) * P2 + P3; p += 4; }
    while (p < e) { h ^= (*p++) * P5; h = z_rotl64(h, 11) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// per-frame decoder state that lives in registers (warp-uniform)
struct ZFrame {
    int huf_bits;      // 0 = no Huffman table yet
    int al_ll, al_of, al_ml;
    bool seq_valid;
    uint32_t rep0, rep1, rep2;
};

// One of the three sequence tables, by mode (one lane).  Returns bytes consumed, 0xFFFFFFFF on error.
__device__ uint32_t z_seq_table(ZWarp* zw, uint32_t* cell, int* al, int mode, const uint8_t* p, uint32_t n, int max_al, int max_sym,
                                       const int16_t* def, int def_n, int def_al, bool have_prev)
{
    if (mode == 0) {
        for (int i = 0; i < def_n; ++i) zw->norm[i] = def[i];
        if (!z_fse_build(cell, zw->next, zw->norm, def_n, def_al)) return 0xFFFFFFFFu;
        *al = def_al;
        return 0;
    }
    if (mode == 1) {
        if (n < 1 || p[0] > max_sym) return 0xFFFFFFFFu;
        cell[0] = p[0];
        *al = 0;
        return 1;
    }
    if (mode == 3) return have_prev ? 0u : 0xFFFFFFFFu;
    int nsym, a;
    const uint32_t used = z_fse_read_ncount(p, n, max_al, max_sym, zw->norm, &nsym, &a);
    if (!used) return 0xFFFFFFFFu;
    if (!z_fse_build(cell, zw->next, zw->norm, nsym, a)) return 0xFFFFFFFFu;
    *al = a;
    return used;
}

// literals section of a compressed block -> lit_dst[0 .. *lit_n).  Whole warp.  Returns bytes consumed, 0 on error.
__device__ uint32_t z_decode_literals(ZWarp* zw, ZFrame& f, const uint8_t* p, uint32_t n, uint8_t* lit_dst, uint32_t lit_cap, uint32_t* lit_n,
                                             uint32_t lane)
{
    if (n < 1) return 0;
    const int type = p[0] & 3, sf = (p[0] >> 2) & 3;
    uint32_t hdr, regen, comp = 0;
    int streams = 4;
    if (type < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = p[0] >> 3; }
        else if (sf == 1) { if (n < 2) return 0; hdr = 2; regen = z_le16(p) >> 4; }
        else { if (n < 3) return 0; hdr = 3; regen = z_le24(p) >> 4; }
        if (regen > kZBlockMax || regen > lit_cap) return 0;
        if (type == 0) {
            if (hdr + regen > n) return 0;
            z_warp_copy(lit_dst, p + hdr, regen, lane);
            *lit_n = regen;
            return hdr + regen;
        }
        if (hdr + 1 > n) return 0;
        z_warp_fill(lit_dst, p[hdr], regen, lane);
        *lit_n = regen;
        return hdr + 1;
    }
    if (n < 3) return 0;
    if (sf <= 1) { hdr = 3; const uint32_t v = z_le24(p); regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF; if (sf == 0) streams = 1; }
    else if (sf == 2) { if (n < 4) return 0; hdr = 4; const uint32_t v = z_le32(p); regen = (v >> 4) & 0x3FFF; comp = (v >> 18) & 0x3FFF; }
    else { if (n < 5) return 0; hdr = 5; const uint64_t v = (uint64_t)z_le32(p) | ((uint64_t)p[4] << 32); regen = (uint32_t)((v >> 4) & 0x3FFFF); comp = (uint32_t)((v >> 22) & 0x3FFFF); }
    if (regen > kZBlockMax || regen > lit_cap || hdr + comp > n) return 0;
    if (streams == 4 && regen < 6) return 0;  // libzstd >= 1.5.4: MIN_LITERALS_FOR_4_STREAMS
    const uint8_t* q = p + hdr;
    uint32_t left = comp;
    if (type == 2) {
        int bits = 0;
        uint32_t t = 0;
        if (lane == 0) {
            int nw = 0;
            t = z_huf_read_weights(zw->weights, zw->wtab, zw->norm, zw->next, q, left, &nw);
            if (t) bits = z_huf_build(zw->huf, zw->weights, nw);
            if (!bits) t = 0;
        }
        t = __shfl_sync(0xffffffffu, t, 0);
        bits = __shfl_sync(0xffffffffu, bits, 0);
        if (!t) return 0;
        f.huf_bits = bits;
        q += t;
        left -= t;
    } else if (!f.huf_bits) return 0;
    __syncwarp();
    bool ok = true;
    if (streams == 1) {
        if (lane == 0) ok = z_huf_decode_stream(zw->huf, f.huf_bits, q, left, lit_dst, regen);
    } else {
        if (left < 10) return 0;
        const uint32_t s1 = z_le16(q), s2 = z_le16(q + 2), s3 = z_le16(q + 4);
        if (6 + s1 + s2 + s3 > left) return 0;
        const uint32_t s4 = left - 6 - s1 - s2 - s3;
        const uint32_t seg = (regen + 3) / 4;
        if (3 * seg > regen) return 0;
        const uint8_t* d = q + 6;
        if (lane < 4) {
            const uint32_t so = lane == 0 ? 0u : lane == 1 ? s1 : lane == 2 ? s1 + s2 : s1 + s2 + s3;
            const uint32_t sn = lane == 0 ? s1 : lane == 1 ? s2 : lane == 2 ? s3 : s4;
            ok = z_huf_decode_stream(zw->huf, f.huf_bits, d + so, sn, lit_dst + lane * seg, lane == 3 ? regen - 3 * seg : seg);
        }
    }
    if (!__all_sync(0xffffffffu, ok)) return 0;
    *lit_n = regen;
    return hdr + comp;
}

// one compressed block: literals + sequences -> out[*op ..].  Whole warp.
__device__ bool z_decode_compressed_block(ZWarp* zw, ZFrame& f, const uint8_t* p, uint32_t n, uint8_t* out, uint64_t cap, uint64_t* op_io,
                                                 uint8_t* lit_scratch, uint32_t lane)
{
    uint64_t op = *op_io;
    // does the block carry sequences?  Peek past the literals section: cheaper to decode the literals header twice
    // than to stage every literal.  Literals go straight to the output when there are none.
    uint32_t lit_n = 0;
    bool direct = false;
    {
        // size of the literals section from its header alone
        if (n < 1) return false;
        const int type = p[0] & 3, sf = (p[0] >> 2) & 3;
        uint32_t sec = 0;
        bool known = true;
        if (type < 2) {
            uint32_t hdr, regen;
            if (sf == 0 || sf == 2) { hdr = 1; regen = p[0] >> 3; }
            else if (sf == 1) { if (n < 2) return false; hdr = 2; regen = z_le16(p) >> 4; }
            else { if (n < 3) return false; hdr = 3; regen = z_le24(p) >> 4; }
            sec = type == 0 ? hdr + regen : hdr + 1;
        } else {
            if (n < 3) return false;
            if (sf <= 1) sec = 3 + ((z_le24(p) >> 14) & 0x3FF);
            else if (sf == 2) { if (n < 4) return false; sec = 4 + ((z_le32(p) >> 18) & 0x3FFF); }
            else { if (n < 5) return false; sec = 5 + (uint32_t)((((uint64_t)z_le32(p) | ((uint64_t)p[4] << 32)) >> 22) & 0x3FFFF); }
        }
        if (known && sec < n && p[sec] == 0 && sec + 1 == n) direct = true;
    }
    uint8_t* lit = direct ? out + op : lit_scratch;
    const uint64_t room = cap - op;
    const uint32_t lit_cap = direct ? (uint32_t)(room < kZBlockMax ? room : kZBlockMax) : kZBlockMax;
    const uint32_t used = z_decode_literals(zw, f, p, n, lit, lit_cap, &lit_n, lane);
    if (!used) return false;
    p += used;
    n -= used;
    if (n < 1) return false;
    uint32_t nseq = p[0];
    uint32_t h = 1;
    if (nseq == 0) {
        if (n != 1) return false;
        if (!direct) {  // (cannot happen: direct was derived from the same bytes)
            if (op + lit_n > cap) return false;
            z_warp_copy(out + op, lit, lit_n, lane);
        }
        *op_io = op + lit_n;
        __syncwarp();
        return true;
    }
    if (direct) return false;
    if (nseq > 0x7F) {
        if (nseq == 0xFF) { if (n < 3) return false; nseq = z_le16(p + 1) + 0x7F00; h = 3; }
        else { if (n < 2) return false; nseq = ((nseq - 0x80) << 8) + p[1]; h = 2; }
    }
    if (h + 1 > n) return false;
    const int modes = p[h++];
    if (modes & 3) return false;  // reserved bits (checked by libzstd 1.5.x)
    // tables (one lane), then broadcast
    uint32_t ok = 1;
    int al_ll = f.al_ll, al_of = f.al_of, al_ml = f.al_ml;
    if (lane == 0) {
        uint32_t r;
        r = z_seq_table(zw, zw->ll, &al_ll, modes >> 6, p + h, n - h, 9, 35, kZ_LL_DEF, 36, 6, f.seq_valid);
        if (r == 0xFFFFFFFFu) ok = 0; else h += r;
        if (ok) { r = z_seq_table(zw, zw->of, &al_of, (modes >> 4) & 3, p + h, n - h, 8, 31, kZ_OF_DEF, 29, 5, f.seq_valid); if (r == 0xFFFFFFFFu) ok = 0; else h += r; }
        if (ok) { r = z_seq_table(zw, zw->ml, &al_ml, (modes >> 2) & 3, p + h, n - h, 9, 52, kZ_ML_DEF, 53, 6, f.seq_valid); if (r == 0xFFFFFFFFu) ok = 0; else h += r; }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) return false;
    h = __shfl_sync(0xffffffffu, h, 0);
    f.al_ll = __shfl_sync(0xffffffffu, al_ll, 0);
    f.al_of = __shfl_sync(0xffffffffu, al_of, 0);
    f.al_ml = __shfl_sync(0xffffffffu, al_ml, 0);
    f.seq_valid = true;
    __syncwarp();

    // sequences: lane 0 decodes up to 32 into shared memory, the warp executes them
    ZBack b;
    b.p = p + h;
    b.pos = 0;
    uint32_t sl = 0, so = 0, sm = 0;
    if (lane == 0) {
        if (!zb_init(b, p + h, n - h)) ok = 0;
        else {
            sl = zb_read(b, (uint32_t)f.al_ll);
            so = zb_read(b, (uint32_t)f.al_of);
            sm = zb_read(b, (uint32_t)f.al_ml);
            if (b.pos < 0) ok = 0;
        }
    }
    if (!__shfl_sync(0xffffffffu, ok, 0)) return false;
    uint32_t lp = 0;
    uint32_t rep0 = f.rep0, rep1 = f.rep1, rep2 = f.rep2;
    for (uint32_t i0 = 0; i0 < nseq; i0 += 32) {
        const uint32_t cnt = nseq - i0 < 32 ? nseq - i0 : 32;
        if (lane == 0) {
            for (uint32_t k = 0; k < cnt && ok; ++k) {
                const uint32_t cl = zw->ll[sl], co = zw->of[so], cm = zw->ml[sm];
                const uint32_t llc = cl & 0xFF, ofc = co & 0xFF, mlc = cm & 0xFF;
                if (llc > 35 || mlc > 52 || ofc > 31) { ok = 0; break; }
                const uint32_t ofv = (ofc ? (1u << ofc) : 1u) + (ofc ? zb_read(b, ofc) : 0u);
                const uint32_t mlen = kZ_ML_BASE[mlc] + zb_read(b, kZ_ML_BITS[mlc]);
                const uint32_t llen = kZ_LL_BASE[llc] + zb_read(b, kZ_LL_BITS[llc]);
                uint32_t offset;
                if (ofv > 3) {
                    offset = ofv - 3;
                    rep2 = rep1; rep1 = rep0; rep0 = offset;
                } else {
                    const uint32_t kk = ofv + (llen == 0 ? 1u : 0u);
                    if (kk == 1) offset = rep0;
                    else {
                        offset = (kk == 4) ? rep0 - 1 : (kk == 2 ? rep1 : rep2);
                        if (offset == 0) offset = 1;  // libzstd: "0 is not valid; input is corrupted; force offset to 1"
                        if (kk != 2) rep2 = rep1;
                        rep1 = rep0;
                        rep0 = offset;
                    }
                }
                if (i0 + k + 1 < nseq) {
                    sl = (cl >> 16) + zb_read(b, (cl >> 8) & 0xFF);
                    sm = (cm >> 16) + zb_read(b, (cm >> 8) & 0xFF);
                    so = (co >> 16) + zb_read(b, (co >> 8) & 0xFF);
                }
                if (b.pos < 0) { ok = 0; break; }
                zw->seq_ll[k] = llen;
                zw->seq_ml[k] = mlen;
                zw->seq_of[k] = offset;
            }
        }
        __syncwarp();
        if (!__shfl_sync(0xffffffffu, ok, 0)) return false;
        for (uint32_t k = 0; k < cnt; ++k) {
            const uint32_t llen = zw->seq_ll[k], mlen = zw->seq_ml[k], offset = zw->seq_of[k];
            if ((uint64_t)lp + llen > lit_n) return false;
            if (op + llen + mlen > cap) return false;
            z_warp_copy(out + op, lit + lp, llen, lane);
            lp += llen;
            op += llen;
            if ((uint64_t)offset > op) return false;
            __syncwarp();
            z_warp_match(out, op, offset, mlen, lane);
            op += mlen;
            __syncwarp();
        }
        __syncwarp();
    }
    if (lane == 0 && b.pos != 0) ok = 0;
    if (!__shfl_sync(0xffffffffu, ok, 0)) return false;
    f.rep0 = __shfl_sync(0xffffffffu, rep0, 0);
    f.rep1 = __shfl_sync(0xffffffffu, rep1, 0);
    f.rep2 = __shfl_sync(0xffffffffu, rep2, 0);
    const uint32_t rest = lit_n - lp;
    if (op + rest > cap) return false;
    z_warp_copy(out + op, lit + lp, rest, lane);
    op += rest;
    __syncwarp();
    *op_io = op;
    return true;
}

// frame header.  Returns header length, 0 on error / truncated input.
__device__ uint32_t z_frame_header(const uint8_t* p, uint64_t n, uint64_t* fcs, bool* has_fcs, bool* checksum)
{
    if (n < 5 || z_le32(p) != 0xFD2FB528u) return 0;
    const int fhd = p[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    if (fhd & 0x08) return 0;  // reserved bit
    uint32_t h = 5;
    if (!single) {
        if (h + 1 > n) return 0;
        const int wl = 10 + (p[h] >> 3);
        if (wl > 31) return 0;
        h++;
    }
    const uint32_t did_bytes = did == 3 ? 4u : (uint32_t)did;
    if (h + did_bytes > n) return 0;
    uint32_t dict_id = 0;
    for (uint32_t i = 0; i < did_bytes; ++i) dict_id |= (uint32_t)p[h + i] << (8 * i);
    h += did_bytes;
    if (dict_id != 0) return 0;  // no dictionary is ever supplied on this path
    const uint32_t fb = fcs_flag == 0 ? (single ? 1u : 0u) : (fcs_flag == 1 ? 2u : (fcs_flag == 2 ? 4u : 8u));
    if (h + fb > n) return 0;
    *has_fcs = fb != 0;
    *fcs = 0;
    if (fb == 1) *fcs = p[h];
    else if (fb == 2) *fcs = z_le16(p + h) + 256;
    else if (fb == 4) *fcs = z_le32(p + h);
    else if (fb == 8) *fcs = z_le64(p + h);
    *checksum = ((fhd >> 2) & 1) != 0;
    return h + fb;
}

// ZSTD_decompress for one read: every frame in src, back to back.  Whole warp.  Returns content bytes or kErrZstd.
__device__ uint32_t z_decompress(ZWarp* zw, const uint8_t* p, uint64_t n, uint8_t* out, uint64_t cap, uint8_t* lit_scratch, uint32_t lane)
{
    uint64_t total = 0;
    while (n > 0) {
        if (n >= 8 && (z_le32(p) & 0xFFFFFFF0u) == 0x184D2A50u) {  // skippable frame
            const uint64_t sk = (uint64_t)z_le32(p + 4) + 8;
            if (sk > n) return kErrZstd;
            p += sk;
            n -= sk;
            continue;
        }
        uint64_t fcs;
        bool has, ck;
        const uint32_t h = z_frame_header(p, n, &fcs, &has, &ck);
        if (!h) return kErrZstd;
        p += h;
        n -= h;
        ZFrame f;
        f.huf_bits = 0;
        f.al_ll = f.al_of = f.al_ml = 0;
        f.seq_valid = false;
        f.rep0 = 1; f.rep1 = 4; f.rep2 = 8;
        uint8_t* fout = out + total;
        const uint64_t fcap = cap - total;
        uint64_t op = 0;
        for (;;) {
            if (n < 3) return kErrZstd;
            const uint32_t bh = z_le24(p);
            const int last = bh & 1, type = (bh >> 1) & 3;
            const uint32_t bs = bh >> 3;
            p += 3;
            n -= 3;
            if (type == 0) {
                if (bs > n || op + bs > fcap) return kErrZstd;
                z_warp_copy(fout + op, p, bs, lane);
                op += bs; p += bs; n -= bs;
            } else if (type == 1) {
                if (n < 1 || op + bs > fcap) return kErrZstd;
                z_warp_fill(fout + op, p[0], bs, lane);
                op += bs; p += 1; n -= 1;
            } else if (type == 2) {
                if (bs > n || bs >= kZBlockMax) return kErrZstd;
                if (!z_decode_compressed_block(zw, f, p, bs, fout, fcap, &op, lit_scratch, lane)) return kErrZstd;
                p += bs; n -= bs;
            } else return kErrZstd;
            __syncwarp();
            if (last) break;
        }
        if (has && op != fcs) return kErrZstd;
        if (ck) {
            if (n < 4) return kErrZstd;
            uint32_t good = 1;
            if (lane == 0) good = (uint32_t)z_xxh64(fout, op) == z_le32(p);
            if (!__shfl_sync(0xffffffffu, good, 0)) return kErrZstd;
            p += 4; n -= 4;
        }
        total += op;
    }
    if (total > 0xFFFFFFF0ull) return kErrZstd;
    return (uint32_t)total;
}

// The general kernel: whole frames, one warp each - the frames the walker did not list as units (sequences, checksums,
// several frames in one buffer, anything malformed: this code decides their result).
__global__ void __launch_bounds__(kZdecWarps * 32) zstd_decode_kernel(ZstdDecParams p)
{
    VBZ_DYN_SMEM(smem);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ZWarp* zw = reinterpret_cast<ZWarp*>(smem) + wid;
    uint8_t* lit_scratch = p.lit_scratch + (uint64_t)(blockIdx.x * kZdecWarps + wid) * kZLitStride;
    const uint32_t n_general = p.counter[2];
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(p.counter, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_general) break;
        const uint32_t r = p.general[k];
        const ZstdRead rd = p.reads[r];
        const uint32_t res = z_decompress(zw, p.src + rd.src_off, rd.src_bytes, p.dst + rd.dst_off, rd.dst_cap, lit_scratch, lane);
        __syncwarp();
        if (lane == 0) p.results[r] = res;
    }
}

// ---- block-parallel path ----------------------------------------------------------------------------------------
// Signal frames are literals: VBZ's own writer emits nothing else, and whatever wrote the frame, a block without
// sequences regenerates exactly its literals.  For such frames every block's position in the output follows from the
// headers alone, so the blocks - and the four Huffman streams inside each - can all be decoded at the same time.
//   zstd_walk_kernel   one THREAD per frame hops over the block headers (twice: count, then list) and turns the frame
//                      into units; a frame it cannot fully vouch for goes to the general kernel's list instead
//   zstd_units_kernel  persistent warps take 8 units at a time: 4 lanes per unit build its Huffman table in shared
//                      memory and decode one stream each; Raw / RLE pieces are warp-wide copies and fills

// header of a literals section.  Returns false if it does not fit in n bytes.
__device__ __forceinline__ bool z_lit_header(const uint8_t* p, uint32_t n, int* type, uint32_t* hdr, uint32_t* regen, uint32_t* comp, int* streams)
{
    if (n < 1) return false;
    const int t = p[0] & 3, sf = (p[0] >> 2) & 3;
    *type = t;
    *streams = 4;
    *comp = 0;
    if (t < 2) {
        if (sf == 0 || sf == 2) { *hdr = 1; *regen = p[0] >> 3; }
        else if (sf == 1) { if (n < 2) return false; *hdr = 2; *regen = z_le16(p) >> 4; }
        else { if (n < 3) return false; *hdr = 3; *regen = z_le24(p) >> 4; }
        return true;
    }
    if (n < 3) return false;
    if (sf <= 1) { *hdr = 3; const uint32_t v = z_le24(p); *regen = (v >> 4) & 0x3FF; *comp = (v >> 14) & 0x3FF; if (sf == 0) *streams = 1; }
    else if (sf == 2) { if (n < 4) return false; *hdr = 4; const uint32_t v = z_le32(p); *regen = (v >> 4) & 0x3FFF; *comp = (v >> 18) & 0x3FFF; }
    else { if (n < 5) return false; *hdr = 5; const uint64_t v = (uint64_t)z_le32(p) | ((uint64_t)p[4] << 32); *regen = (uint32_t)((v >> 4) & 0x3FFFF); *comp = (uint32_t)((v >> 22) & 0x3FFFF); }
    return true;
}

// number of sequences of a block, from the first bytes of its sequences section (RFC 8878 3.1.1.3.2.1).  *h = bytes used; 0 on truncation
__device__ __forceinline__ bool z_seq_count(const uint8_t* q, uint32_t left, uint32_t* nseq, uint32_t* h)
{
    if (left < 1) return false;
    const uint32_t b0 = q[0];
    if (b0 < 128) { *nseq = b0; *h = 1; return true; }
    if (b0 < 255) { if (left < 2) return false; *nseq = ((b0 - 128) << 8) + q[1]; *h = 2; return true; }
    if (left < 3) return false;
    *nseq = z_le16(q + 1) + 0x7F00; *h = 3;
    return true;
}

// One pass over a frame.  emit == nullptr: count.  Returns the number of units, or 0xFFFFFFFF if the frame is not for
// this path.  *total = content bytes (only meaningful when no block carries sequences); *n_seqs = sequences of all Seq units
// (their records are numbered from seq_base on, block after block).
// A CHAIN = a Seq unit whose block brings all three sequence tables itself (or is the frame's first) and the units up to the next such
// unit: what one warp of zstd_seq_kernel decodes - the chains of a frame need nothing from each other but the repeat offsets, which
// are settled when the last of them is done.  The list lives behind the executor's list (ZstdDecParams::seq_list + unit_cap).
struct ZstdSeqChain { uint32_t frame, first, count, nseq; };  // slot in seq_frames, first unit (relative to the frame's), units, their sequences
struct ZWalk { uint32_t total, groups, seqs; bool deferred; uint32_t chains; };
__device__ uint32_t z_walk_frame(const uint8_t* src, const ZstdRead& rd, uint32_t r, ZstdUnit* emit, ZstdGroup* gemit, uint32_t unit_base,
                                 uint32_t seq_base, ZWalk* out, ZstdSeqChain* cemit = nullptr, uint32_t frame_slot = 0)
{
    const uint32_t kNo = 0xFFFFFFFFu;
    const uint8_t* p = src + rd.src_off;
    uint64_t n = rd.src_bytes;
    uint64_t fcs;
    bool has, ck;
    const uint32_t h = z_frame_header(p, n, &fcs, &has, &ck);
    if (!h || ck) return kNo;
    p += h;
    n -= h;
    const uint64_t cap = rd.dst_cap;
    uint64_t op = 0;
    uint32_t units = 0, groups = 0, seqs = 0, chains = 0, chain_first = 0, chain_seqs = 0;
    bool deferred = false;         // a block with sequences came by: positions from here on are known after zstd_seq_kernel only
    const uint8_t* tab = nullptr;  // the last literals section that carried a tree
    // a GROUP = up to 16 consecutive Huffman units decoded with the same table (the one that carries the tree and the
    // treeless ones after it), or a single Raw / RLE piece.  Seq units belong to no group.
    uint32_t g_first = 0, g_count = 0;
    for (;;) {
        if (n < 3) return kNo;
        const uint32_t bh = z_le24(p);
        const int last = bh & 1, type = (bh >> 1) & 3;
        const uint32_t bs = bh >> 3;
        p += 3;
        n -= 3;
        ZstdUnit u;
        u.dst_off = rd.dst_off + op;
        u.read = r;
        u.tab_back = 0;
        u.kind = kZUnitNop;
        u.n = 0;
        u.regen = 0;
        u.nseq = 0;
        u.seq_first = 0;
        u.pad = 0;
        u.src_off = (uint64_t)(p - src);
        uint32_t adv;
        int seq_lt = 0;  // Seq units: type of the literals section
        if (type == 0) {
            if (bs > n || (!deferred && op + bs > cap)) return kNo;
            u.kind = kZUnitCopy; u.n = bs; u.regen = bs; adv = bs;
        } else if (type == 1) {
            if (n < 1 || bs > kZBlockMax || (!deferred && op + bs > cap)) return kNo;
            u.kind = kZUnitFill; u.n = bs; u.regen = bs; adv = 1;
        } else if (type == 2) {
            if (bs > n || bs >= kZBlockMax) return kNo;
            int lt, streams;
            uint32_t hdr, regen, comp;
            if (!z_lit_header(p, bs, &lt, &hdr, &regen, &comp, &streams)) return kNo;
            const uint32_t sec = lt == 0 ? hdr + regen : lt == 1 ? hdr + 1 : hdr + comp;
            if (sec >= bs || regen > kZBlockMax) return kNo;
            if (lt == 2) tab = p;
            else if (lt == 3 && !tab) return kNo;
            if (p[sec] == 0) {  // "0 sequences": the block regenerates exactly its literals
                if ((uint64_t)sec + 1 != bs || (!deferred && op + regen > cap)) return kNo;
                if (lt == 0) { u.kind = kZUnitCopy; u.src_off += hdr; u.n = regen; }
                else if (lt == 1) { u.kind = kZUnitFill; u.src_off += hdr; u.n = regen; }
                else { u.kind = kZUnitHuf; u.n = sec; u.tab_back = (uint32_t)(p - tab); }
                u.regen = regen;
            } else {
                uint32_t nseq, sh;
                if (!z_seq_count(p + sec, bs - sec, &nseq, &sh)) return kNo;
                u.kind = kZUnitSeq; u.n = bs; u.tab_back = lt == 3 ? (uint32_t)(p - tab) : 0u;
                seq_lt = lt;
                u.nseq = nseq; u.seq_first = seq_base + seqs;
                seqs += nseq;
                // the compression modes byte: no Repeat_Mode among the three = the block stands alone (a damaged one: the sequence kernel says so)
                const uint32_t modes = sec + sh < bs ? p[sec + sh] : 0u;
                const bool alone = ((modes >> 6) & 3u) != 3u && ((modes >> 4) & 3u) != 3u && ((modes >> 2) & 3u) != 3u;
                if (!deferred || alone) {
                    if (deferred) { if (cemit) cemit[chains] = ZstdSeqChain{frame_slot, chain_first, units - chain_first, chain_seqs}; ++chains; }
                    chain_first = units;
                    chain_seqs = 0;
                }
                chain_seqs += nseq;
                deferred = true;
            }
            adv = bs;
        } else return kNo;
        op += u.regen;
        if (u.kind == kZUnitSeq || u.n != 0) {
            // a Seq unit's literals are decoded by the units kernel like any other literals section (to the END of the block's place in the
            // output, csrc: zstd_seqexec_kernel): with Huffman literals it joins the groups, with Raw / RLE ones it stands alone
            const bool huf = u.kind == kZUnitHuf || (u.kind == kZUnitSeq && seq_lt >= 2);
            const bool extend = huf && u.tab_back != 0 && g_count != 0 && g_count < 16;  // (an open group is always a Huffman one)
            if (!extend) {
                if (g_count) { if (gemit) gemit[groups] = ZstdGroup{unit_base + g_first, g_count}; ++groups; }
                g_first = units;
                g_count = 0;
            }
            if (emit) emit[units] = u;
            ++units;
            ++g_count;
            if (!huf) {  // Raw / RLE pieces stand alone
                if (gemit) gemit[groups] = ZstdGroup{unit_base + g_first, g_count};
                ++groups;
                g_count = 0;
            }
        }
        p += adv;
        n -= adv;
        if (last) break;
    }
    if (n != 0 || op > 0xFFFFFFF0ull) return kNo;
    if (!deferred && has && op != fcs) return kNo;
    if (g_count) { if (gemit) gemit[groups] = ZstdGroup{unit_base + g_first, g_count}; ++groups; }
    if (deferred) { if (cemit) cemit[chains] = ZstdSeqChain{frame_slot, chain_first, units - chain_first, chain_seqs}; ++chains; }
    out->total = (uint32_t)op;
    out->groups = groups;
    out->seqs = seqs;
    out->deferred = deferred;
    out->chains = chains;
    return units;
}

__global__ void zstd_walk_kernel(ZstdDecParams p)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_reads) return;
    const ZstdRead rd = p.reads[r];
    if (rd.rejected) { p.results[r] = rd.preset; return; }
    ZWalk w = {0, 0, 0, false, 0};
    const uint32_t cnt = z_walk_frame(p.src, rd, r, nullptr, nullptr, 0, 0, &w);
    bool mine = cnt != 0xFFFFFFFFu;
    if (mine && cnt != 0) {
        const uint32_t base = atomicAdd(p.counter + 1, cnt);
        const uint32_t gbase = w.groups ? atomicAdd(p.counter + 4, w.groups) : 0u;
        const uint32_t sbase = w.seqs ? atomicAdd(p.counter + 5, w.seqs) : 0u;
        const uint32_t gcnt = w.groups;
        const uint32_t cbase = w.chains ? atomicAdd(p.counter + 10, w.chains) : 0u;
        if (base >= p.unit_cap || cnt > p.unit_cap - base || (gcnt && (gbase >= p.unit_cap || gcnt > p.unit_cap - gbase)) ||
            (w.seqs && (sbase >= p.seq_cap || w.seqs > p.seq_cap - sbase)) || (w.chains && (cbase >= p.unit_cap || w.chains > p.unit_cap - cbase))) {
            // out of slots: the units kernel stops at unit_cap; what was allotted here stays empty - and must say so: the executor walks from a
            // Seq unit to the units behind it for as long as they name the same read, and what an earlier batch left in these slots may
            for (uint32_t k = base; k < p.unit_cap && k - base < cnt; ++k) { p.units[k].kind = kZUnitNop; p.units[k].read = 0xFFFFFFFFu; }
            for (uint32_t k = gbase; gcnt && k < p.unit_cap && k - gbase < gcnt; ++k) p.groups[k] = ZstdGroup{0, 0};
            for (uint32_t k = cbase; w.chains && k < p.unit_cap && k - cbase < w.chains; ++k)
                reinterpret_cast<ZstdSeqChain*>(p.seq_list + p.unit_cap)[k] = ZstdSeqChain{0xFFFFFFFFu, 0, 0, 0};
            mine = false;
        } else {
            const uint32_t slot = w.deferred ? atomicAdd(p.counter + 6, 1u) : 0u;
            if (w.deferred) p.seq_frames[slot] = ZstdSeqFrame{r, base, cnt, w.chains};  // (pad: chains still to be decoded; top bit: one of them failed)
            z_walk_frame(p.src, rd, r, p.units + base, p.groups + gbase, base, sbase, &w, reinterpret_cast<ZstdSeqChain*>(p.seq_list + p.unit_cap) + cbase, slot);
        }
    }
    if (mine) { if (!w.deferred) p.results[r] = w.total; }
    else p.general[atomicAdd(p.counter + 2, 1u)] = r;
#ifdef VBZ_CPUSIM
    if (getenv("VBZ_B200_TRACE_WALK")) fprintf(stderr, "walk: read %u -> %s (%u units, %u groups, %u sequences, %u bytes)\n", r, mine ? (w.deferred ? "units + sequence kernels" : "units") : "general", mine ? cnt : 0u, mine ? w.groups : 0u, mine ? w.seqs : 0u, w.total);
#endif
}

// ---- the units kernel: one CTA of 64 threads per group -------------------------------------------------------------------
// The table is built once per group, by all threads; then every unit of the group gets 4 lanes, one per Huffman stream:
// with the writer of csrc/zstd_enc_kernels.cuh (16 blocks of 8 KiB per table) all 64 lanes decode 2 KiB each.  Frames of
// other writers (libzstd: one tree per 128 KiB block, or a few treeless blocks after it) use 4 lanes per block.
constexpr int kZUnitThreads = 64;
struct ZGroupSmem {
    uint16_t huf[2048];    // sym | nbBits << 8.  While the weights are read: the tree description (160 bytes at +2048) and the FSE
                           // scratch (table 64 x u32, norm 256 x i16, next 256 x u16 at +0)
    uint8_t weights[256];
    uint16_t pos[256];     // first cell of each symbol
    uint32_t rank[16], start[16];
    uint32_t v[8];         // [0] group index, [1] weights, [2] tree bytes, [3] max_bits (0 = bad table), [4] sum of 2^(w-1), [5] error
};

// weights[0..nw) -> decode table, all threads of the CTA.  Returns max_bits or 0 on error (same checks as z_huf_build).
__device__ __forceinline__ int z_huf_build_cta(ZGroupSmem* sm, int nw, uint32_t tid)
{
    if (tid < 16) sm->rank[tid] = 0;
    if (tid == 0) { sm->v[4] = 0; sm->v[5] = 0; }
    __syncthreads();
    for (int i = (int)tid; i < nw; i += kZUnitThreads) {
        const uint32_t w = sm->weights[i];
        if (w > 12) { sm->v[5] = 1; continue; }
        atomicAdd(&sm->rank[w], 1u);
        atomicAdd(&sm->v[4], (1u << w) >> 1);
    }
    __syncthreads();
    if (tid == 0) {
        int max_bits = 0;
        const uint32_t total = sm->v[4];
        if (!sm->v[5] && total != 0 && nw > 0 && nw < 256) {
            max_bits = z_highbit(total) + 1;
            const uint32_t rest = (1u << max_bits) - total;
            if (max_bits > 11 || (rest & (rest - 1)) != 0) max_bits = 0;  // HUF_TABLELOG_MAX of the format
            else {
                const uint32_t wl = (uint32_t)z_highbit(rest) + 1;
                sm->weights[nw] = (uint8_t)wl;
                sm->rank[wl]++;
                if (sm->rank[1] < 2 || (sm->rank[1] & 1)) max_bits = 0;
                uint32_t acc = 0;
                for (int r = 1; r <= 12; ++r) { sm->start[r] = acc; acc += sm->rank[r] << (r - 1); }
            }
        }
        sm->v[3] = (uint32_t)max_bits;
    }
    __syncthreads();
    const int max_bits = (int)sm->v[3];
    if (!max_bits) return 0;
    nw += 1;
    // cells: weight 1 (longest codes) first, symbols ascending within a weight.  Thread r walks the symbols of weight r.
    if (tid >= 1 && tid <= 12 && sm->rank[tid]) {
        uint32_t cur = sm->start[tid];
        const uint32_t len = (1u << tid) >> 1;
        for (int sy = 0; sy < nw; ++sy)
            if (sm->weights[sy] == tid) { sm->pos[sy] = (uint16_t)cur; cur += len; }
    }
    __syncthreads();
    for (int sy = (int)tid; sy < nw; sy += kZUnitThreads) {
        const uint32_t r = sm->weights[sy];
        if (!r) continue;
        const uint32_t len = (1u << r) >> 1;
        const uint32_t e = (uint32_t)sy | ((uint32_t)(max_bits + 1 - (int)r) << 8);
        uint16_t* c = sm->huf + sm->pos[sy];  // a run of `len` cells starts at a multiple of len: wide stores are aligned
        const uint32_t e2 = e | (e << 16);
        if (len >= 8) {
            const uint4 e8 = make_uint4(e2, e2, e2, e2);
            for (uint32_t k = 0; k < len; k += 8) *reinterpret_cast<uint4*>(c + k) = e8;
        } else if (len == 4) *reinterpret_cast<uint2*>(c) = make_uint2(e2, e2);
        else if (len == 2) *reinterpret_cast<uint32_t*>(c) = e2;
        else c[0] = (uint16_t)e;
    }
    __syncthreads();
    return max_bits;
}

__global__ void __launch_bounds__(kZUnitThreads) zstd_units_kernel(ZstdDecParams p)
{
    VBZ_DYN_SMEM(smem);
    ZGroupSmem* sm = reinterpret_cast<ZGroupSmem*>(smem);
    const uint32_t tid = threadIdx.x, j = tid & 3;
    uint32_t n_groups = p.counter[4];
    if (n_groups > p.unit_cap) n_groups = p.unit_cap;
    for (;;) {
        __syncthreads();  // the previous group's table and scalars are no longer in use
        if (tid == 0) sm->v[0] = atomicAdd(p.counter + 3, 1u);
        __syncthreads();
        const uint32_t g = sm->v[0];
        if (g >= n_groups) break;
        const ZstdGroup grp = p.groups[g];
        if (grp.count == 0) continue;
        const ZstdUnit u0 = p.units[grp.first];
        if (u0.kind == kZUnitCopy || u0.kind == kZUnitFill) {  // Raw / RLE piece: the whole CTA
            uint8_t* d = p.dst + u0.dst_off;
            const uint8_t* sp = p.src + u0.src_off;
            const uint32_t cn = u0.n;
            if (u0.kind == kZUnitFill) {
                const uint8_t v = sp[0];
                for (uint32_t k = tid; k < cn; k += kZUnitThreads) d[k] = v;
            } else if ((((uintptr_t)d ^ (uintptr_t)sp) & 3u) == 0 && cn >= 64) {  // same word phase: bytes to the boundary, words, bytes
                const uint32_t head = (4u - (uint32_t)((uintptr_t)d & 3u)) & 3u;
                const uint32_t words = (cn - head) >> 2, done = head + 4 * words;
                if (tid < head) d[tid] = sp[tid];
                else if (tid >= 8 && tid - 8 < cn - done) d[done + tid - 8] = sp[done + tid - 8];
                const uint32_t* s4 = reinterpret_cast<const uint32_t*>(sp + head);
                uint32_t* d4 = reinterpret_cast<uint32_t*>(d + head);
                for (uint32_t k = tid; k < words; k += kZUnitThreads) d4[k] = s4[k];
            } else {
                for (uint32_t k = tid; k < cn; k += kZUnitThreads) d[k] = sp[k];
            }
            continue;
        }
        if (u0.kind == kZUnitSeq) {  // Raw / RLE literals of a block with sequences: to the end of the block's place (the Huffman ones: below)
            int lt0 = 0, st0 = 0;
            uint32_t hdr0 = 0, lit0 = 0, comp0 = 0;
            if (z_lit_header(p.src + u0.src_off, u0.n, &lt0, &hdr0, &lit0, &comp0, &st0) && lt0 < 2 && lit0 <= u0.regen) {
                uint8_t* d = p.dst + u0.dst_off + (u0.regen - lit0);
                const uint8_t* sp = p.src + u0.src_off + hdr0;
                if (lt0 == 1) { const uint8_t v = sp[0]; for (uint32_t k = tid; k < lit0; k += kZUnitThreads) d[k] = v; }
                else for (uint32_t k = tid; k < lit0; k += kZUnitThreads) d[k] = sp[k];
                continue;
            }
        }
        if (u0.kind != kZUnitHuf && u0.kind != kZUnitSeq) continue;
        // ---- the group's table: the tree sits in the literals section tab_back bytes before the first unit's ----------------------
        int bits = 0;
        {
            const uint8_t* tp = p.src + u0.src_off - u0.tab_back;
            int lt2 = 0, st2 = 0;
            uint32_t hdr2 = 0, regen2 = 0, comp2 = 0;
            const bool hok = z_lit_header(tp, u0.tab_back ? 5u : u0.n, &lt2, &hdr2, &regen2, &comp2, &st2) && lt2 == 2 && (u0.tab_back != 0 || hdr2 + comp2 <= u0.n);
            // the description is at most 1 + 127 bytes (FSE) or 1 + 64 (direct): stage it in shared memory for the serial reader
            uint8_t* tree = reinterpret_cast<uint8_t*>(sm->huf) + 2048;
            const uint32_t tn = hok ? (comp2 < 160u ? comp2 : 160u) : 0u;
            for (uint32_t k = tid; k < tn; k += kZUnitThreads) tree[k] = tp[hdr2 + k];
            __syncthreads();
            if (tid == 0) {
                int nw = 0;
                uint32_t t = 0;
                if (hok) {
                    uint8_t* scratch = reinterpret_cast<uint8_t*>(sm->huf);
                    t = z_huf_read_weights(sm->weights, reinterpret_cast<uint32_t*>(scratch), reinterpret_cast<int16_t*>(scratch + 256),
                                           reinterpret_cast<uint16_t*>(scratch + 768), tree, tn, &nw);
                }
                sm->v[1] = (uint32_t)nw;
                sm->v[2] = t;
            }
            __syncthreads();
            if (sm->v[2]) bits = z_huf_build_cta(sm, (int)sm->v[1], tid);
        }
        const uint32_t t = sm->v[2];
        // ---- the units, 16 at a time: 4 lanes each -----------------------------------------------------------------------------
        for (uint32_t k0 = 0; k0 < grp.count; k0 += 16) {
            const uint32_t ui = k0 + (tid >> 2);
            if (ui >= grp.count) continue;
            const ZstdUnit un = p.units[grp.first + ui];
            const uint8_t* lp = p.src + un.src_off;
            int lt = 0, streams = 4;
            uint32_t hdr = 0, regen = 0, comp = 0;
            if (un.kind == kZUnitNop) continue;
            bool ok = bits != 0 && (un.kind == kZUnitHuf || un.kind == kZUnitSeq) && z_lit_header(lp, un.n, &lt, &hdr, &regen, &comp, &streams) && lt >= 2 &&
                      regen <= kZBlockMax && hdr + comp <= un.n && !(streams == 4 && regen < 6) && (un.kind == kZUnitHuf || regen <= un.regen);
            if (ok) {
                const uint8_t* q = lp + hdr;
                uint32_t left = comp;
                if (lt == 2) {  // this unit carries the tree (t <= comp: read_weights never reads past what it was given)
                    if (t > left) ok = false;
                    else { q += t; left -= t; }
                }
                uint8_t* dst = p.dst + un.dst_off + (un.kind == kZUnitSeq && ok ? un.regen - regen : 0u);  // Seq: the literals wait at the end of the block's place
                if (!ok) {
                } else if (streams == 1) {
                    if (j == 0) ok = z_huf_decode_stream(sm->huf, bits, q, left, dst, regen);
                } else if (left < 10) ok = false;
                else {
                    const uint32_t s1 = z_le16(q), s2 = z_le16(q + 2), s3 = z_le16(q + 4);
                    const uint32_t seg = (regen + 3) / 4;
                    if (6 + s1 + s2 + s3 > left || 3 * seg > regen) ok = false;
                    else {
                        const uint32_t s4 = left - 6 - s1 - s2 - s3;
                        const uint32_t so = j == 0 ? 0u : j == 1 ? s1 : j == 2 ? s1 + s2 : s1 + s2 + s3;
                        const uint32_t sn = j == 0 ? s1 : j == 1 ? s2 : j == 2 ? s3 : s4;
                        ok = z_huf_decode_stream(sm->huf, bits, q + 6 + so, sn, dst + j * seg, j == 3 ? regen - 3 * seg : seg);
                    }
                }
            }
            if (!ok) p.results[un.read] = kErrZstd;
        }
    }
}

// ---- frames with sequences: the two sequence kernels ------------------------------------------------------------------------
// A block with sequences regenerates lit_n + sum(match lengths) bytes - known only once its sequences have been decoded, and
// the FSE states make that a serial walk.  So the work is split by what it depends on:
//   zstd_seq_kernel      a WARP per CHAIN of a frame's Seq units (a block that brings its own sequence tables and the blocks that repeat
//                        them; the walker lists the chains): lanes 0 - 2 build the tables, a lane per block decodes the sequences into
//                        RECORDS {offset, ll, ml}.  The warp that finishes a frame's last chain resolves the repeat-offset codes,
//                        derives every unit's position from the block sizes (dst_off of ALL the frame's units is written here) and
//                        lists the executor's work.  Nothing is copied yet.
//   zstd_units_kernel    then decodes every block WITHOUT sequences of every frame, as before - including the blocks of a
//                        libzstd-written signal frame that follow its one or two blocks with sequences - and the LITERALS of the
//                        blocks with sequences, right-aligned in the block's place in the output.
//   zstd_seqexec_kernel  a warp per executor chain (a Seq unit whose matches stay inside its block, and the units behind it whose matches
//                        do not): the records expand the block in place, 32 per round.  Matches may reach back into earlier blocks:
//                        those are complete (the units kernel ran before; earlier chains of the frame report done, release/acquire).
// A frame the sequence kernel cannot vouch for (damaged, more records than the pool holds) goes to the general kernel.
// entries of ZstdDecParams::seq_list: a unit index, whether the unit is the first Seq unit of its frame, and - set by the executor - done
constexpr uint32_t kZListIndexMask = 0x3FFFFFFFu, kZListFirst = 0x40000000u, kZListDone = 0x80000000u;
struct ZSeqWarp {  // zstd_seq_kernel: per-warp shared state
    // packed cells: x = next-state base | state bits << 16 | extra bits << 24, y = value base (a sequence is three look-ups, nothing else)
    uint2 ll[512], ml[512], of[256];
    int16_t norm[3][64];                 // scratch of the three table builders (lanes 0 - 2)
    uint16_t next[3][64];
};
// cells as z_fse_build leaves them (sym | nb << 8 | base << 16 in x) -> packed cells; which: 0 literal lengths, 1 offsets, 2 match lengths
__device__ __forceinline__ void z_seq_pack(uint2* cell, int al, int which, uint32_t lane)
{
    const uint32_t T = 1u << al;  // (RLE: al = 0, one cell)
    for (uint32_t u = lane; u < T; u += 32) {
        const uint32_t x = cell[u].x, sym = x & 0xFFu;
        uint32_t extra, base;
        if (which == 0) { extra = sym < 36 ? kZ_LL_BITS[sym] : 0u; base = sym < 36 ? kZ_LL_BASE[sym] : 0u; }
        else if (which == 1) { extra = sym & 31u; base = 1u << (sym & 31u); }
        else { extra = sym < 53 ? kZ_ML_BITS[sym] : 0u; base = sym < 53 ? kZ_ML_BASE[sym] : 0u; }
        cell[u] = make_uint2((x >> 16) | (((x >> 8) & 0xFFu) << 16) | (extra << 24), base);
    }
}

// One of the three sequence tables, by mode (one lane); norm / next: scratch for 64 symbols.  The cells go to every other word of `cell`
// (the x words of the packed table, see z_seq_pack).  Returns bytes consumed, 0xFFFFFFFF on error.
__device__ uint32_t z_seq_table2(int16_t* norm, uint16_t* next, uint32_t* cell, int* al, int mode, const uint8_t* p, uint32_t n, int max_al, int max_sym,
                                 const int16_t* def, int def_n, int def_al, bool have_prev)
{
    if (mode == 0) {
        for (int i = 0; i < def_n; ++i) norm[i] = def[i];
        if (!z_fse_build<2>(cell, next, norm, def_n, def_al)) return 0xFFFFFFFFu;
        *al = def_al;
        return 0;
    }
    if (mode == 1) {
        if (n < 1 || p[0] > max_sym) return 0xFFFFFFFFu;
        cell[0] = p[0];
        *al = 0;
        return 1;
    }
    if (mode == 3) return have_prev ? 0u : 0xFFFFFFFFu;
    int nsym, a;
    const uint32_t used = z_fse_read_ncount(p, n, max_al, max_sym, norm, &nsym, &a);
    if (!used) return 0xFFFFFFFFu;
    if (!z_fse_build<2>(cell, next, norm, nsym, a)) return 0xFFFFFFFFu;
    *al = a;
    return used;
}

// what a lane knows about "its" Seq unit before the sequences are decoded
struct ZSeqHead {
    const uint8_t* sec;  // the sequences section behind its count: the modes byte first
    uint32_t n;          // bytes from there to the end of the block
    uint32_t nseq, lit_n;
    int modes;           // -1: the block is damaged
};
__device__ __forceinline__ ZSeqHead z_seq_head(const uint8_t* bp, uint32_t bs, uint32_t nseq_expected)
{
    ZSeqHead h = {nullptr, 0, 0, 0, -1};
    int lt, streams;
    uint32_t hdr, lit_n, comp;
    if (!z_lit_header(bp, bs, &lt, &hdr, &lit_n, &comp, &streams)) return h;
    const uint32_t sec = lt == 0 ? hdr + lit_n : lt == 1 ? hdr + 1 : hdr + comp;
    if (sec >= bs || lit_n > kZBlockMax) return h;
    uint32_t nseq, used;
    if (!z_seq_count(bp + sec, bs - sec, &nseq, &used) || nseq != nseq_expected || nseq == 0) return h;
    if (sec + used + 1 > bs) return h;
    const int modes = bp[sec + used];
    if (modes & 3) return h;  // reserved bits (checked by libzstd 1.5.x)
    h.sec = bp + sec + used;
    h.n = bs - sec - used;
    h.nseq = nseq;
    h.lit_n = lit_n;
    h.modes = modes;
    return h;
}

// The sequences of one block -> records, one lane.  The bitstream is p[0 .. n).  With `track` the lane knows the repeat offsets the block
// starts with (rep[]) and resolves the repeat codes as it goes; without, an offset value of 1 .. 3 is left in the record as 0x80000000 | value
// (*raw says whether there was any) for z_resolve_repeats.  Returns the bytes the block regenerates, 0xFFFFFFFF if it is damaged.
// *back: a match may reach back beyond the start of the block.
// A sequence is a serial step (the three states of sequence i + 1 come out of the bits behind those of sequence i), so the step is
// kept short: `bp` counts the unread bits, the 64 bits below it come from three aligned words, and the packed cells say how many bits
// each of the six fields takes - no refill tests, no second look-up between the cell and the value.
__device__ uint32_t z_lane_sequences(const ZSeqWarp* zw, int al_ll, int al_of, int al_ml, const uint8_t* p, uint32_t n, uint32_t nseq, uint32_t lit_n,
                                     ZstdSeq* out, bool track, uint32_t* rep, bool* raw, bool* back)
{
    const uint32_t kBad = 0xFFFFFFFFu;
    if (n == 0 || p[n - 1] == 0 || n > (1u << 24)) return kBad;
    const uint32_t phase = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(p - phase);  // stream bit k = bit 8 * phase + k of this array of words
    const uint32_t lead = 8u * phase, wlast = (phase + n - 1) >> 2;
    int32_t left = (int32_t)((n - 1) * 8) + z_highbit(p[n - 1]);  // unread bits (n <= 2^24; negative: the reader ran past the start)
    // n bits below bit position lead + left (n <= 32); bits before the stream read as zero
    auto slow_read = [&](uint32_t nb) -> uint32_t {
        if (nb == 0) return 0u;
        left -= nb;
        uint64_t v = 0;
        const int32_t from = left;  // stream bit index of the field's lowest bit (may be negative)
        const int32_t lo = from < 0 ? 0 : from;
        const int32_t hi = from + (int32_t)nb;  // exclusive
        if (hi > 0) {
            const uint32_t a = lead + (uint32_t)lo, wi = a >> 5, sh = a & 31u;
            const uint64_t two = (uint64_t)words[wi] | ((uint64_t)(((a + (uint32_t)(hi - lo) - 1) >> 5) > wi ? words[wi + 1] : 0u) << 32);
            v = (two >> sh) & ((1ull << (uint32_t)(hi - lo)) - 1ull);
            v <<= (uint32_t)(lo - from);
        }
        return (uint32_t)v;
    };
    uint32_t sl = slow_read((uint32_t)al_ll), so = slow_read((uint32_t)al_of), sm = slow_read((uint32_t)al_ml);
    if (left < 0) return kBad;
    uint32_t rep0 = rep[0], rep1 = rep[1], rep2 = rep[2];
    uint32_t lits = 0, regen = 0;
    bool any_raw = false, reach = false;
    // the window: hi:lo = the 64 stream bits below bit position `wtop` (top bit first), of which the top `wtop - left` are spent already.
    // It is reloaded from three aligned words whenever fewer than 64 - spent bits could be needed; between reloads fields are cut out
    // of the pair at their offsets - six independent extractions, no chain through a shifting window.
    uint32_t lo = 0, hi = 0;
    int32_t wtop = -1;  // (no window yet)
    for (uint32_t i = 0; i < nseq; ++i) {
        const uint2 cl = zw->ll[sl], co = zw->of[so], cm = zw->ml[sm];
        const uint32_t ofb = co.x >> 24, mlb = cm.x >> 24, llb = cl.x >> 24;
        const bool more = i + 1 < nseq;
        const uint32_t nls = more ? (cl.x >> 16) & 0xFFu : 0u, nms = more ? (cm.x >> 16) & 0xFFu : 0u, nos = more ? (co.x >> 16) & 0xFFu : 0u;
        const uint32_t c1 = ofb, c2 = c1 + mlb, c3 = c2 + llb, c4 = c3 + nls, c5 = c4 + nms, need = c5 + nos;
        uint32_t ofx, mlx, llx, xl, xm, xo;
        if (left >= 96 && need <= 64) {
            uint32_t spent = (uint32_t)(wtop - left);
            if (wtop < 0 || spent + need > 64) {
                const uint32_t a = lead + (uint32_t)left - 64u, wi = a >> 5, sh = a & 31u;
                const uint32_t w0 = words[wi], w1 = words[wi + 1], w2 = words[wi + 2 < wlast ? wi + 2 : wlast];  // (with sh = 0 the third word is not part of the 64 bits)
                lo = __funnelshift_r(w0, w1, sh);
                hi = __funnelshift_r(w1, w2, sh);
                wtop = left;
                spent = 0;
            }
            // nb bits that start t bits below the top of hi:lo (t + nb <= 64, nb <= 31)
#define Z_CUT(dst, t, nb) { const uint32_t t_ = spent + (t); const uint32_t x_ = t_ < 32 ? __funnelshift_l(lo, hi, t_) : lo << (t_ - 32); dst = (x_ >> 1) >> (31u - (nb)); }
            Z_CUT(ofx, 0u, ofb) Z_CUT(mlx, c1, mlb) Z_CUT(llx, c2, llb) Z_CUT(xl, c3, nls) Z_CUT(xm, c4, nms) Z_CUT(xo, c5, nos)
#undef Z_CUT
            left -= need;
        } else {
            ofx = slow_read(ofb); mlx = slow_read(mlb); llx = slow_read(llb);
            xl = slow_read(nls); xm = slow_read(nms); xo = slow_read(nos);
            if (left < 0) return kBad;
            wtop = -1;
        }
        const uint32_t ofv = co.y + ofx, mlen = cm.y + mlx, llen = cl.y + llx;
        sl = (cl.x & 0xFFFFu) + xl;
        sm = (cm.x & 0xFFFFu) + xm;
        so = (co.x & 0xFFFFu) + xo;
        uint32_t offset;
        if (ofv > 3) {
            offset = ofv - 3;
            rep2 = rep1; rep1 = rep0; rep0 = offset;
        } else if (!track) {
            offset = 0x80000000u | ofv;
            any_raw = true;
        } else {
            const uint32_t kk = ofv + (llen == 0 ? 1u : 0u);
            if (kk == 1) offset = rep0;
            else {
                offset = (kk == 4) ? rep0 - 1 : (kk == 2 ? rep1 : rep2);
                if (offset == 0) offset = 1;  // libzstd: "0 is not valid; input is corrupted; force offset to 1"
                if (kk != 2) rep2 = rep1;
                rep1 = rep0;
                rep0 = offset;
            }
        }
        lits += llen;
        if (lits > lit_n || regen + llen + mlen > 2u * kZBlockMax) return kBad;
        if (offset > regen + llen) reach = true;  // (an unresolved code counts: its value has the top bit set)
        regen += llen + mlen;
        out[i] = ZstdSeq{offset, llen, mlen, 0};
    }
    if (left != 0) return kBad;
    regen += lit_n - lits;  // the literals after the last sequence
    if (regen > kZBlockMax) return kBad;
    rep[0] = rep0; rep[1] = rep1; rep[2] = rep2;
    *raw = any_raw;
    *back = reach;
    return regen;
}
__device__ __forceinline__ uint32_t z_ld_cg(const uint32_t* p)
{
#ifdef VBZ_CPUSIM
    return *p;
#else
    return __ldcg(p);
#endif
}
// The repeat codes the lanes left in the records of a block, resolved in order from the offsets the block starts with - by the whole
// warp: 32 records per round are loaded side by side (ld.cg: they may have been written by a warp of another SM), then walked in
// order with their fields broadcast by shuffles, so no load sits in the serial chain.  rep[] is the same in all lanes.
__device__ void z_resolve_repeats(ZstdSeq* sq, uint32_t nseq, uint32_t* rep, uint32_t lane)
{
    uint32_t rep0 = rep[0], rep1 = rep[1], rep2 = rep[2];
    for (uint32_t base = 0; base < nseq; base += 32) {
        const uint32_t i = base + lane, cnt = nseq - base < 32u ? nseq - base : 32u;
        const uint32_t off = i < nseq ? z_ld_cg(&sq[i].offset) : 0u;
        const uint32_t ll = i < nseq ? z_ld_cg(&sq[i].ll) : 1u;
        if (!__ballot_sync(0xffffffffu, (off & 0x80000000u) != 0)) {  // explicit offsets only: the last three are what is left behind
            const uint32_t a = __shfl_sync(0xffffffffu, off, (int)cnt - 1), b = __shfl_sync(0xffffffffu, off, cnt >= 2 ? (int)cnt - 2 : 0),
                           c = __shfl_sync(0xffffffffu, off, cnt >= 3 ? (int)cnt - 3 : 0);
            if (cnt >= 3) { rep2 = c; rep1 = b; rep0 = a; }
            else if (cnt == 2) { rep2 = rep0; rep1 = b; rep0 = a; }
            else { rep2 = rep1; rep1 = rep0; rep0 = a; }
            continue;
        }
        uint32_t mine = off;
        for (uint32_t j = 0; j < cnt; ++j) {
            uint32_t offset = __shfl_sync(0xffffffffu, off, (int)j);
            if (!(offset & 0x80000000u)) { rep2 = rep1; rep1 = rep0; rep0 = offset; continue; }
            const uint32_t kk = (offset & 3u) + (__shfl_sync(0xffffffffu, ll, (int)j) == 0 ? 1u : 0u);
            if (kk == 1) offset = rep0;
            else {
                offset = (kk == 4) ? rep0 - 1 : (kk == 2 ? rep1 : rep2);
                if (offset == 0) offset = 1;  // libzstd: "0 is not valid; input is corrupted; force offset to 1"
                if (kk != 2) rep2 = rep1;
                rep1 = rep0;
                rep0 = offset;
            }
            if (lane == j) mine = offset;
        }
        if (i < nseq && mine != off) sq[i].offset = mine;
    }
    rep[0] = rep0; rep[1] = rep1; rep[2] = rep2;
}

// zstd_seq_kernel: a WARP per CHAIN of a frame's Seq units (zstd_walk_kernel lists them).  The chain's Seq units are taken in SETS: a
// unit whose block brings tables (or names the predefined / RLE ones) and the units after it that repeat all three (Repeat_Mode x 3:
// what the writer of csrc/zstd_match_kernels.cuh gives every block of a group but the first).  Lanes 0 - 2 build the set's tables,
// then every unit of the set is decoded by a lane of its own - sequences are serial work per block, not per frame.  libzstd's sets are
// single blocks, so its frames are decoded a warp per block (the blocks that repeat a table ride with the one before them).
// No lane knows the repeat offsets its block starts with: repeat codes stay in the records, and the warp that finishes a frame's LAST
// chain walks the frame's units once - resolves the codes in order, gives every unit its place in the output and lists the executor's
// chains.  (What other warps wrote is read with ld.cg there: their lines may sit stale in this SM's L1.)
constexpr int kZSeqWarps = 4;
#ifdef VBZ_CPUSIM
#define Z_SEQ_TRACE(...) do { if (lane == 0 && getenv("VBZ_B200_TRACE_SEQ")) fprintf(stderr, __VA_ARGS__); } while (0)
#else
#define Z_SEQ_TRACE(...) do { } while (0)
#endif
__global__ void __launch_bounds__(kZSeqWarps * 32) zstd_seq_kernel(ZstdDecParams p)
{
    VBZ_DYN_SMEM(smem);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ZSeqWarp* zw = reinterpret_cast<ZSeqWarp*>(smem) + wid;
    const ZstdSeqChain* chains = reinterpret_cast<const ZstdSeqChain*>(p.seq_list + p.unit_cap);
    const uint32_t n_chains = p.counter[10] < p.unit_cap ? p.counter[10] : p.unit_cap;
    const uint32_t kBad = 0xFFFFFFFFu;
    // the list is walked twice: the long chains first (a block of control bytes written by libzstd holds 19 k sequences, serial work for one
    // lane: started last, it would run alone at the end), then the rest
    constexpr uint32_t kLong = 2048;
    for (uint32_t round = 0;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(p.counter + (round == 0 ? 7 : 11), 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_chains) {
            if (round == 0) { round = 1; continue; }
            break;
        }
        const ZstdSeqChain ch = chains[k];
        if (ch.frame == 0xFFFFFFFFu) continue;  // (slots of a frame that found no room for its units)
        if ((ch.nseq >= kLong) != (round == 0)) continue;
        const ZstdSeqFrame sf = p.seq_frames[ch.frame];
        int al[3] = {0, 0, 0};  // literal lengths, offsets, match lengths
        bool tables_valid = false;
        bool ok = true;
        const uint32_t ch_end = ch.first + ch.count;
        uint32_t ui = ch.first;
        while (ui < ch_end && ok) {
            // ---- the next 32 units: which are Seq, what their blocks say ---------------------------------------------------------------
            const uint32_t mine = ui + lane;
            ZstdUnit u = {};
            u.kind = kZUnitNop;
            if (mine < ch_end) u = p.units[sf.unit_first + mine];
            const bool is_seq = u.kind == kZUnitSeq;
            const uint32_t seqmask = __ballot_sync(0xffffffffu, is_seq);
            if (!seqmask) { ui += 32; continue; }
            const uint32_t head_lane = (uint32_t)__ffs((int)seqmask) - 1;
            ZSeqHead h = {nullptr, 0, 0, 0, -1};
            if (is_seq) h = z_seq_head(p.src + u.src_off, u.n, u.nseq);
            if (__ballot_sync(0xffffffffu, is_seq && h.modes < 0)) { Z_SEQ_TRACE("seq: chain %u of read %u: a block's sequences section is damaged\n", k, sf.read); ok = false; break; }
            // the set: the head and the Seq units behind it (among these 32) that repeat all three tables
            const uint32_t stop = __ballot_sync(0xffffffffu, is_seq && lane > head_lane && h.modes != 0xFC);
            const uint32_t upto = stop ? (uint32_t)__ffs((int)stop) - 1 : 32u;
            const uint32_t members = seqmask & (upto < 32 ? (1u << upto) - 1u : 0xFFFFFFFFu);
            const bool member = (members >> lane) & 1u;
            // ---- the head's tables: lanes 0, 1, 2 build one each ---------------------------------------------------------------------------
            const int hmodes = __shfl_sync(0xffffffffu, h.modes, head_lane);
            const uint64_t hsec = __shfl_sync(0xffffffffu, (uint64_t)(uintptr_t)h.sec, head_lane);
            const uint32_t hn = __shfl_sync(0xffffffffu, h.n, head_lane);
            const uint8_t* tp = reinterpret_cast<const uint8_t*>((uintptr_t)hsec) + 1;  // behind the modes byte
            uint32_t tn = hn - 1;
            uint32_t used_all = 0;
            {
                // the descriptions follow each other: a lane must know the sizes of those before its own.  Sizes first (lane 0 reads the
                // counts of all three - cheap next to the table builds), then the builds side by side.
                uint32_t off3[3] = {0, 0, 0}, bad = 0;
                if (lane == 0) {
                    uint32_t at = 0;
                    for (int t = 0; t < 3 && !bad; ++t) {
                        const int mode = (hmodes >> (6 - 2 * t)) & 3;
                        off3[t] = at;
                        if (mode == 1) { if (at + 1 > tn) bad = 1; else at += 1; }
                        else if (mode == 2) {
                            int nsym, a;
                            const uint32_t used = at <= tn ? z_fse_read_ncount(tp + at, tn - at, t == 1 ? 8 : 9, t == 0 ? 35 : (t == 1 ? 31 : 52), zw->norm[0], &nsym, &a) : 0u;
                            if (!used) bad = 1; else at += used;
                        } else if (mode == 3 && !tables_valid) bad = 1;
                    }
                    used_all = at;
                }
                bad = __shfl_sync(0xffffffffu, bad, 0);
                if (bad) { Z_SEQ_TRACE("seq: chain %u of read %u: table descriptions\n", k, sf.read); ok = false; break; }
                used_all = __shfl_sync(0xffffffffu, used_all, 0);
                const uint32_t o0 = __shfl_sync(0xffffffffu, off3[0], 0), o1 = __shfl_sync(0xffffffffu, off3[1], 0), o2 = __shfl_sync(0xffffffffu, off3[2], 0);
                uint32_t tbad = 0;
                int my_al = 0;
                if (lane < 3) {
                    const int mode = (hmodes >> (6 - 2 * (int)lane)) & 3;
                    const uint32_t at = lane == 0 ? o0 : (lane == 1 ? o1 : o2);
                    uint32_t* cell = reinterpret_cast<uint32_t*>(lane == 0 ? zw->ll : (lane == 1 ? zw->of : zw->ml));  // (the x words)
                    my_al = al[lane];
                    const uint32_t r = z_seq_table2(zw->norm[lane], zw->next[lane], cell, &my_al, mode, tp + at, tn - at, lane == 1 ? 8 : 9,
                                                    lane == 0 ? 35 : (lane == 1 ? 31 : 52), lane == 0 ? kZ_LL_DEF : (lane == 1 ? kZ_OF_DEF : kZ_ML_DEF),
                                                    lane == 0 ? 36 : (lane == 1 ? 29 : 53), lane == 1 ? 5 : 6, tables_valid);
                    if (r == kBad) tbad = 1;
                }
                if (__ballot_sync(0xffffffffu, tbad != 0)) { Z_SEQ_TRACE("seq: chain %u of read %u: table build\n", k, sf.read); ok = false; break; }
                al[0] = __shfl_sync(0xffffffffu, my_al, 0);
                al[1] = __shfl_sync(0xffffffffu, my_al, 1);
                al[2] = __shfl_sync(0xffffffffu, my_al, 2);
                tables_valid = true;
                __syncwarp();
                // the tables that were (re)built: cells -> packed cells, all lanes
                if (((hmodes >> 6) & 3) != 3) z_seq_pack(zw->ll, al[0], 0, lane);
                if (((hmodes >> 4) & 3) != 3) z_seq_pack(zw->of, al[1], 1, lane);
                if (((hmodes >> 2) & 3) != 3) z_seq_pack(zw->ml, al[2], 2, lane);
            }
            __syncwarp();
            // ---- every member decodes its block; repeat codes stay in the records ------------------------------------------------------------------
            uint32_t regen = 0, my_rep[3] = {1, 4, 8};
            bool raw = false, back = false;
            if (member) {
                const bool is_head = lane == head_lane;
                const uint8_t* bits = is_head ? h.sec + 1 + used_all : h.sec + 1;
                const uint32_t nbits = is_head ? (h.n >= 1 + used_all ? h.n - 1 - used_all : 0u) : h.n - 1;
                regen = (u.seq_first >= p.seq_cap || u.nseq > p.seq_cap - u.seq_first)
                            ? kBad
                            : z_lane_sequences(zw, al[0], al[1], al[2], bits, nbits, h.nseq, h.lit_n, p.seqs + u.seq_first, false, my_rep, &raw, &back);
            }
            if (__ballot_sync(0xffffffffu, member && regen == kBad)) { Z_SEQ_TRACE("seq: chain %u of read %u: a block's sequences (or no room for its records)\n", k, sf.read); ok = false; break; }
            if (member) {
                ZstdUnit* up = p.units + sf.unit_first + mine;
                up->regen = regen;
                up->pad = (back || raw ? 1u : 0u) | (raw ? 2u : 0u);  // 1: a match may reach back beyond the block (the executor waits for the units before it); 2: repeat codes to resolve
            }
            __syncwarp();
            ui += upto;  // (upto >= head_lane + 1: the units before the head are not Seq)
        }
        // ---- report; the warp that finishes the frame's last chain goes on --------------------------------------------------------------------
        __threadfence();
        __syncwarp();
        uint32_t old = 0;
        if (lane == 0) {
            if (!ok) atomicOr(&p.seq_frames[ch.frame].pad, 0x80000000u);
            old = atomicAdd(&p.seq_frames[ch.frame].pad, 0xFFFFFFFFu);  // (minus one)
        }
        old = __shfl_sync(0xffffffffu, old, 0);
        if ((old & 0x7FFFFFFFu) != 1u) continue;
        __threadfence();
        ok = !(old & 0x80000000u);
        Z_SEQ_TRACE("seq: read %u: last chain done, %s\n", sf.read, ok ? "all good" : "a chain failed");
        const ZstdRead rd = p.reads[sf.read];
        uint64_t fcs = 0;
        bool has = false, ck = false;
        z_frame_header(p.src + rd.src_off, rd.src_bytes, &fcs, &has, &ck);  // (the walker accepted it)
        // ---- the repeat offsets, unit after unit (all lanes alike): a block without codes leaves its last three offsets behind, one with codes is walked ----
        if (ok) {
            uint32_t rep[3] = {1, 4, 8};
            for (uint32_t i = 0; i < sf.unit_count; ++i) {
                const ZstdUnit* up = p.units + sf.unit_first + i;
                if (up->kind != kZUnitSeq) continue;  // (kind, nseq, seq_first: the walker's, read-only here)
                ZstdSeq* sq = p.seqs + up->seq_first;
                const uint32_t nseq = up->nseq;
                if (z_ld_cg(&up->pad) & 2u) z_resolve_repeats(sq, nseq, rep, lane);
                else {
                    const uint32_t r0 = z_ld_cg(&sq[nseq - 1].offset);
                    if (nseq >= 3) { rep[2] = z_ld_cg(&sq[nseq - 3].offset); rep[1] = z_ld_cg(&sq[nseq - 2].offset); rep[0] = r0; }
                    else if (nseq == 2) { rep[2] = rep[0]; rep[1] = z_ld_cg(&sq[0].offset); rep[0] = r0; }
                    else { rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = r0; }
                }
            }
            __threadfence();
        }
        __syncwarp();
        // ---- positions: every unit of the frame, 32 at a time ----------------------------------------------------------------------------
        uint64_t op = 0;
        for (uint32_t u0 = 0; u0 < sf.unit_count && ok; u0 += 32) {
            const uint32_t mine = u0 + lane;
            const uint32_t regen = mine < sf.unit_count ? z_ld_cg(&p.units[sf.unit_first + mine].regen) : 0u;
            const uint32_t incl = warp_incl_scan(regen, lane);
            if (mine < sf.unit_count) p.units[sf.unit_first + mine].dst_off = rd.dst_off + op + incl - regen;
            op += __shfl_sync(0xffffffffu, incl, 31);
            if (op > rd.dst_cap) ok = false;
        }
        if (ok && has && op != fcs) ok = false;
        if (ok && (uint64_t)sf.unit_first + sf.unit_count >= kZListIndexMask) ok = false;
        Z_SEQ_TRACE("seq: read %u: %llu bytes, content size %llu -> %s\n", sf.read, (unsigned long long)op, (unsigned long long)fcs, ok ? "executor" : "general kernel");
        if (ok) {
            if (lane == 0) p.results[sf.read] = (uint32_t)op;
            // the executor's work list: the frame's CHAINS in block order, side by side.  A chain = a Seq unit that needs nothing from the
            // Seq units before it (or the frame's first) and the Seq units behind it that do (pad & 1): one warp runs it from end to end
            uint32_t heads = 0;
            for (uint32_t u0 = 0; u0 < sf.unit_count; u0 += 32) {
                const uint32_t mine = u0 + lane;
                const bool is_seq = mine < sf.unit_count && p.units[sf.unit_first + mine].kind == kZUnitSeq;
                const uint32_t m = __ballot_sync(0xffffffffu, is_seq);
                const bool head = is_seq && ((heads == 0 && !(m & ((1u << lane) - 1u))) || !(z_ld_cg(&p.units[sf.unit_first + mine].pad) & 1u));
                heads += (uint32_t)__popc((int)__ballot_sync(0xffffffffu, head));
            }
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(p.counter + 9, heads);
            base = __shfl_sync(0xffffffffu, base, 0);
            uint32_t done = 0;
            for (uint32_t u0 = 0; u0 < sf.unit_count; u0 += 32) {
                const uint32_t mine = u0 + lane;
                const bool is_seq = mine < sf.unit_count && p.units[sf.unit_first + mine].kind == kZUnitSeq;
                const uint32_t m = __ballot_sync(0xffffffffu, is_seq);
                const bool head = is_seq && ((done == 0 && !(m & ((1u << lane) - 1u))) || !(z_ld_cg(&p.units[sf.unit_first + mine].pad) & 1u));
                const uint32_t hm = __ballot_sync(0xffffffffu, head);
                const uint32_t rank = done + (uint32_t)__popc((int)(hm & ((1u << lane) - 1u)));
                if (head) p.seq_list[base + rank] = (sf.unit_first + mine) | (rank == 0 ? kZListFirst : 0u);
                done += (uint32_t)__popc((int)hm);
            }
        } else {
            // not for this path after all: its units do nothing, the general kernel decides the frame's result
            for (uint32_t ui2 = lane; ui2 < sf.unit_count; ui2 += 32) p.units[sf.unit_first + ui2].kind = kZUnitNop;
            if (lane == 0) {
                p.seq_frames[ch.frame].unit_count = 0;
                p.general[atomicAdd(p.counter + 2, 1u)] = sf.read;
            }
        }
        __syncwarp();
    }
}

// zstd_seqexec_kernel: a WARP per chain of Seq units, taken from the list in order.  The unit's literals already sit at the END of the block's place in
// the output (the units kernel put them there); the records are executed from the front: a literal run moves down to where it belongs,
// a match copies what lies `offset` back.  What a sequence writes ends at or below where the literals of the next one wait, so nothing
// is overwritten before it is read - the block expands in place, and the literals behind the last sequence are where they belong already.
// The list holds CHAINS: a unit that needs nothing from the frame's Seq units before it, and the units behind it with a match that may
// reach back beyond their block - those run behind it in the same warp.  Before the first such unit the warp waits until the frame's
// earlier chains are done (the entries of the list before its own; everything else was written by the units kernel).  Chains are taken in
// list order: what a chain waits for is running or done.
constexpr int kZExecWarps = 4;
// wait until every Seq unit of the frame before list entry k is done: the entries from k - 1 down to the one that carries the First mark,
// 32 per look (lane 0 holds the nearest)
__device__ __forceinline__ void z_exec_wait(const uint32_t* list, uint32_t k, uint32_t lane)
{
    long long hi = (long long)k;
    for (;;) {
        const long long idx = hi - 1 - (long long)lane;
        bool found;
        for (;;) {
            const uint32_t e = idx >= 0 ? vbz_ld_acquire_u32(list + idx) : (kZListDone | kZListFirst);
            const uint32_t firsts = __ballot_sync(0xffffffffu, (e & kZListFirst) != 0);
            const uint32_t upto = firsts ? (2u << ((uint32_t)__ffs((int)firsts) - 1)) - 1u : 0xFFFFFFFFu;  // the lanes that hold entries of this frame
            const uint32_t open = __ballot_sync(0xffffffffu, !(e & kZListDone)) & upto;
            found = firsts != 0;
            if (!open) break;
            VBZ_SPIN_YIELD();
        }
        if (found) break;
        hi -= 32;
    }
    __syncwarp();
}
// A short match copied by its own lane.  Byte after byte every load would wait for the store before it (the compiler must assume they
// meet), and that wait was half of the executor's time on libzstd's frames.  So: a source eight or more bytes behind is copied in chunks
// whose loads are all issued before their stores; a nearer one (the match repeats the `offset` bytes before it) is loaded once into a
// register and written out of it.
__device__ __forceinline__ void z_lane_match(uint8_t* out, uint64_t dst, uint32_t offset, uint32_t n)
{
    const uint8_t* from = out + dst - offset;
    uint8_t* to = out + dst;
    if (offset >= 8) {
        uint32_t q = 0;
        for (; q + 8 <= n; q += 8) {
            uint8_t b[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) b[k] = from[q + k];
#pragma unroll
            for (int k = 0; k < 8; ++k) to[q + k] = b[k];
        }
        const uint32_t r = n - q;
        uint8_t b[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) b[k] = (uint32_t)k < r ? from[q + k] : (uint8_t)0;
#pragma unroll
        for (int k = 0; k < 7; ++k)
            if ((uint32_t)k < r) to[q + k] = b[k];
        return;
    }
    uint64_t pat = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k)
        if ((uint32_t)k < offset) pat |= (uint64_t)from[k] << (8 * k);
    for (uint32_t q = 0, ph = 0; q < n; ++q) {
        to[q] = (uint8_t)(pat >> (8u * ph));
        ph = ph + 1 == offset ? 0u : ph + 1;
    }
}
// the records of one Seq unit (all lanes); false: the block is damaged
__device__ bool z_exec_unit(const ZstdDecParams& p, const ZstdUnit& u, uint64_t frame_off, uint32_t lane)
{
    uint8_t* out = p.dst;
    const uint8_t* bp = p.src + u.src_off;
    int lt = 0, streams = 4;
    uint32_t hdr = 0, lit_n = 0, comp = 0;
    if (!z_lit_header(bp, u.n, &lt, &hdr, &lit_n, &comp, &streams) || lit_n > u.regen) return false;
    const uint8_t* lit = out + u.dst_off + (u.regen - lit_n);  // literal i of the block waits at lit[i]
    uint64_t op = u.dst_off;
    uint32_t lp = 0;
    const ZstdSeq* sq = p.seqs + u.seq_first;
    for (uint32_t i0 = 0; i0 < u.nseq; i0 += 32) {
        const uint32_t i = i0 + lane;
        const ZstdSeq rec = i < u.nseq ? sq[i] : ZstdSeq{1, 0, 0, 0};
        const uint32_t span = rec.ll + rec.ml;
        const uint32_t incl = warp_incl_scan(span, lane), lincl = warp_incl_scan(rec.ll, lane);
        const uint64_t my_out = op + incl - span;        // where this sequence's literals go
        const uint64_t mdst = my_out + rec.ll, msrc = mdst - rec.offset;
        // an offset beyond the frame's first byte: damaged (checked here: the sequence kernel decodes blocks side by side and knows no positions)
        if (__ballot_sync(0xffffffffu, rec.ml != 0 && (uint64_t)rec.offset > mdst - frame_off)) return false;
        // ---- literal runs, in literal order, 32 bytes per step: literal o of the round belongs to the first sequence whose inclusive
        // count exceeds o.  Reads come before the writes of a step, and a step writes below what later steps read.
        const uint32_t total = __shfl_sync(0xffffffffu, lincl, 31);
        for (uint32_t o0 = 0; o0 < total; o0 += 32) {
            const uint32_t o = o0 + lane;
            uint32_t j = 0;
#pragma unroll
            for (uint32_t step = 16; step != 0; step >>= 1) {
                const uint32_t v = __shfl_sync(0xffffffffu, lincl, j + step - 1);
                if (v <= o) j += step;
            }
            j &= 31u;
            const uint32_t ej = __shfl_sync(0xffffffffu, lincl - rec.ll, j);
            const uint32_t oj_lo = __shfl_sync(0xffffffffu, (uint32_t)my_out, j), oj_hi = __shfl_sync(0xffffffffu, (uint32_t)(my_out >> 32), j);
            uint8_t v = 0;
            if (o < total) v = lit[lp + o];
            __syncwarp();
            if (o < total) out[(((uint64_t)oj_hi << 32) | oj_lo) + (o - ej)] = v;
        }
        __syncwarp();
        // ---- matches: short and independent of this round's other sequences -> the lane itself
        const bool dep = msrc < my_out && msrc + rec.ml > op;
        // (a run - offset 1 - of any reasonable length too: libzstd turns the zero runs among control bytes into matches of hundreds of bytes, and
        // a lane that fills 16 bytes per store is done long before the warp has made its round trip through memory for it)
        const bool self = rec.ml != 0 && !dep && (rec.ml <= 48 || (rec.offset == 1 && rec.ml <= 4096));
        if (self) {
            if (rec.offset == 1) {
                const uint32_t v = out[mdst - 1];
                uint8_t* a = out + mdst;
                uint8_t* const e = a + rec.ml;
                while (a < e && ((uintptr_t)a & 15u)) *a++ = (uint8_t)v;
                const uint32_t w = v * 0x01010101u;
                for (; a + 16 <= e; a += 16) *reinterpret_cast<uint4*>(a) = make_uint4(w, w, w, w);
                while (a < e) *a++ = (uint8_t)v;
            } else {
                z_lane_match(out, mdst, rec.offset, rec.ml);
            }
        }
        // ---- the others (libzstd's control bytes: eight or nine of a round's 32, nearly all short and fed by bytes this round writes): what is
        // final now are the literals and the matches above.  A waiting match is READY when its source meets no destination of a waiting match
        // before it; the ready ones run side by side (the short ones in their lanes, the long ones by the warp), then the test is made
        // again - a round trip through memory per level of dependence instead of one per match.  (A later destination never reaches down
        // into an earlier source, so nothing is overwritten before it is read.)
        uint32_t waiting = __ballot_sync(0xffffffffu, rec.ml != 0 && !self);
        const uint32_t d_lo = (uint32_t)(mdst - op), d_hi = d_lo + rec.ml;  // relative to the round's first byte
        const uint32_t s_hi = msrc + rec.ml > op ? (uint32_t)(msrc + rec.ml - op) : 0u, s_lo = msrc > op ? (uint32_t)(msrc - op) : 0u;
        __syncwarp();
        while (waiting) {
            bool blocked = false;
            for (uint32_t scan = waiting; scan; scan &= scan - 1) {
                const uint32_t f = (uint32_t)__ffs((int)scan) - 1;
                const uint32_t f_lo = __shfl_sync(0xffffffffu, d_lo, f), f_hi = __shfl_sync(0xffffffffu, d_hi, f);
                if (lane > f && s_lo < f_hi && s_hi > f_lo) blocked = true;
            }
            const bool go = ((waiting >> lane) & 1u) && !blocked;
            const uint32_t ready = __ballot_sync(0xffffffffu, go);
            if (go && rec.ml <= 48) z_lane_match(out, mdst, rec.offset, rec.ml);
            uint32_t longs = __ballot_sync(0xffffffffu, go && rec.ml > 48);
            __syncwarp();
            while (longs) {
                const uint32_t fl = (uint32_t)__ffs((int)longs) - 1;
                longs &= longs - 1;
                const uint32_t n = __shfl_sync(0xffffffffu, rec.ml, fl), fo = __shfl_sync(0xffffffffu, rec.offset, fl);
                const uint32_t fd_lo = __shfl_sync(0xffffffffu, (uint32_t)mdst, fl), fd_hi = __shfl_sync(0xffffffffu, (uint32_t)(mdst >> 32), fl);
                z_warp_match(out, ((uint64_t)fd_hi << 32) | fd_lo, fo, n, lane);
                __syncwarp();
            }
            waiting &= ~ready;
        }
        op += __shfl_sync(0xffffffffu, incl, 31);
        lp += total;
    }
    return lp <= lit_n && op + (lit_n - lp) == u.dst_off + u.regen;  // (the sequence kernel's sums say so)
}

__global__ void __launch_bounds__(kZExecWarps * 32) zstd_seqexec_kernel(ZstdDecParams p)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_list = p.counter[9];
    uint32_t n_units = p.counter[1];
    if (n_units > p.unit_cap) n_units = p.unit_cap;
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(p.counter + 8, 1u);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_list) break;
        const uint32_t entry = p.seq_list[k];
        uint32_t idx = entry & kZListIndexMask;
        ZstdUnit u = p.units[idx];
        const uint32_t read = u.read;
        const uint64_t frame_off = p.reads[read].dst_off;
        bool ok = true, waited = (entry & kZListFirst) != 0;
        // the chain: this unit, then the Seq units of the frame behind it that need what lies before them
        for (;;) {
            if (!z_exec_unit(p, u, frame_off, lane)) { ok = false; break; }
            __syncwarp();
            bool more = false;
            while (++idx < n_units) {
                u = p.units[idx];
                if (u.read != read) break;
                if (u.kind != kZUnitSeq) continue;
                more = (u.pad & 1u) != 0;
                break;
            }
            if (!more) break;
            if (!waited) { z_exec_wait(p.seq_list, k, lane); waited = true; }  // the chains of the frame before this one
        }
        if (!ok && lane == 0) p.results[read] = kErrZstd;
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            vbz_st_release_u32(p.seq_list + k, entry | kZListDone);
        }
    }
}

}  // namespace vbz
