// csrc/zstd_enc_kernels.cuh - stage 2 of VBZ encode on the GPU: a Zstandard frame writer (RFC 8878).
//
// Replaces the reference's call ZSTD_compress(dst, cap, stage1, n, level) (ref vbz/vbz.cpp:194-204; libzstd is an
// un-vendored dependency, pinned zstd/1.4.8 by CMakeLists.txt:93).  The bytes libzstd would write are not
// reproducible across its versions, so the contract is the FORMAT: every frame written here carries
// its content size (vbz.cpp:236-240 demands it), is a single frame, and decodes with stock libzstd and with the
// reference's vbz_decompress to the same bytes - the parity tests check exactly that.
//
// What is emitted.  On nanopore signal zstd level 1 finds almost no matches: 98 % of its output is Huffman-coded
// literals.  This writer therefore entropy-codes and does not search.  The content is cut
// into GROUPS of 128 KiB that share one Huffman code (own length-limited code, <= 11 bits, weights FSE-compressed
// or direct), and every group into SUB-BLOCKS of 8 KiB, each its own zstd block: a Compressed block holding one
// Huffman literals section (4 streams; 1 below 1 KiB) and no sequences - the first one of the group carries the
// tree, the others are "treeless" (they reuse it, RFC 8878 3.1.1.3.1.1) - or a Raw block when that is smaller.  A
// group of one repeated byte is a single RLE block, an incompressible one a single Raw block.  Small blocks with a
// shared table are what lets the DECODER run 64 Huffman streams per table at once (csrc/zstd_dec_kernels.cuh);
// the table is also built once per 128 KiB here.  The compression level only has to be non-zero.
//
// Parallelism: groups are numbered frame-major; two kernels.
//   zstd_enc_code_kernel    a WARP per CODE SET - eight consecutive groups of a region share one code (every group still carries the
//                           tree in its first Huffman block, so groups stay independent of each other in the frame) - no CTA-wide
//                           barrier: histogram with 128-bit loads and
//                           shared-memory atomics (zero bytes counted in a register), rank sort of the 256 counts by the
//                           warp, then lane 0 runs the inherently serial part - two-queue Huffman construction, length
//                           limiting, FSE coding of the weights - while the SM's other 31 warps work on their own groups.
//                           (In the first version this serial part sat in the middle of the 256-thread encode kernel and
//                           60 % of that kernel's warp samples were the other 255 threads waiting at the barrier behind it.)
//                           The code, the tree description and the kind of the group go to a 1.2 KB record in global memory.
//   zstd_encode_kernel      a CTA (8 warps) per group:
//     sizes     a warp per stream, 8 streams after each other: a lane owns a contiguous run of <= 64 symbols, which
//               it loads with aligned 128-bit loads and shifts into place in registers; code lengths are summed, a
//               shuffle scan gives the lane's bit offset (kept for the next pass) and the stream's size
//     position  one thread sizes the sub-blocks; the group's size is posted, a decoupled look-back over the
//               frame's earlier groups (csrc/svb_kernels.cuh) gives its position
//     pack      two sub-blocks at a time (4 warps = 4 streams each): codes of two symbols are merged and OR-ed
//               into a zeroed shared-memory staging area, which then leaves with aligned 128-bit stores
#pragma once
#include "zstd_dec_kernels.cuh"
#include "svb_kernels.cuh"
#include "zstd_match_kernels.cuh"

namespace vbz {

constexpr int kZencThreads = 256;
constexpr int kZencWarps = kZencThreads / 32;
constexpr uint32_t kZSubs = kZEncBlock / kZEncSub;   // sub-blocks per group (16)
constexpr uint32_t kZStreams = kZSubs * 4;           // Huffman streams per group (64)

// the state of one group's code construction (zstd_enc_code_kernel: one per warp)
struct ZCode {
    uint32_t cnt[256];        // the histogram
    uint32_t code[256];       // code | len << 16
    uint8_t order[256];       // symbols by ascending count
    uint8_t len[256];
    uint16_t parent[512];
    uint32_t nodew[256];      // weights of the internal nodes
    uint32_t wcell[64];       // FSE table of the weights
    int16_t norm[16];
    uint16_t next[16];
    uint8_t cellidx[13][64];  // per weight value: its cells in table order
    uint8_t weights[256];
    uint8_t tree[160];        // the Huffman tree description
    uint32_t max_bits;
    uint32_t pad[3];
};
// what zstd_enc_code_kernel leaves for zstd_encode_kernel, one per group (ZstdEncParams::codes)
struct ZGroupCode {
    uint32_t code[256];  // code | len << 16; 0 for symbols that do not occur
    uint8_t tree[160];
    uint32_t kind, tree_bytes, nsym, pad;
};
static_assert(sizeof(ZGroupCode) == kZGroupCodeBytes, "host and device disagree on the size of a group's code record");
struct ZEncSmem {
    uint16_t lane_off[kZStreams][32];  // bit offset of every lane's run inside its stream
    uint32_t code[256];       // code | len << 16
    uint8_t len[256];
    uint8_t tree[160];        // the Huffman tree description
    uint32_t v[16];           // scalars shared by the CTA (see the kZv* indices)
    uint64_t excl;
    uint32_t sbits[kZStreams];     // bits of every stream of the group
    uint32_t sub_off[kZSubs + 1];  // where each sub-block starts in the group's output
    uint32_t sub_lit[kZSubs];      // S groups: size of each block's literals section
    uint32_t seq_off[kZSubs], seq_len[kZSubs];  // S groups: where each block's sequences section sits in the group's literal slot, its bytes
    uint8_t sub_kind[kZSubs];      // kZKindRaw / kZKindHuf
    uint8_t sub_tree[kZSubs];      // != 0: this sub-block carries the tree
    alignas(16) uint8_t stage[2][kZEncSub + 96];  // (sized header, frame header,) block header, block payload; slack for the copy
};
enum { kZvNsym = 0, kZvMaxBits, kZvTreeBytes, kZvKind, kZvTotal, kZvPre };
enum { kZKindRaw = 0, kZKindRle = 1, kZKindHuf = 2 };

// little-endian bit writer that grows upward from bit 0 of a byte buffer (one thread)
struct ZBitW {
    uint8_t* p;
    uint64_t acc;
    uint32_t nbits, nbytes;
};
__device__ __forceinline__ void zw_init(ZBitW& w, uint8_t* p) { w.p = p; w.acc = 0; w.nbits = 0; w.nbytes = 0; }
__device__ __forceinline__ void zw_put(ZBitW& w, uint32_t v, uint32_t nb)
{
    w.acc |= (uint64_t)v << w.nbits;
    w.nbits += nb;
    while (w.nbits >= 8) { w.p[w.nbytes++] = (uint8_t)w.acc; w.acc >>= 8; w.nbits -= 8; }
}
__device__ __forceinline__ uint32_t zw_close(ZBitW& w, bool marker)  // -> bytes written
{
    if (marker) zw_put(w, 1, 1);
    if (w.nbits) { w.p[w.nbytes++] = (uint8_t)w.acc; w.acc = 0; w.nbits = 0; }
    return w.nbytes;
}

// ---- Huffman code of one block (one thread; cnt[], order[] are ready) -------------------------------------------
// -> len[], code[], max_bits.  Returns false if no valid code could be built (the block is then stored raw).
// All lanes call.  The two-queue merge and the depth chain are serial (lane 0: ~2 n dependent steps); everything that is a map or a
// reduction over the symbols - leaf lengths, Kraft sum, the census of lengths, the canonical codes - is spread over the warp.
__device__ bool zenc_build_code(ZCode* s, int nsym, uint32_t lane)
{
    const int n = nsym;
    uint32_t* scratch = s->nodew + 224;  // 32 words at the end of nodew: census of lengths [0..12], first cell per weight [16..28] (n <= 256 nodes use [0, n - 1))
    if (lane == 0) {
        // two-queue construction over the leaves sorted by ascending count: leaves 0..nsym-1, internal nodes nsym..2nsym-2
        int li = 0, ii = 0, made = 0;
        for (made = 0; made < n - 1; ++made) {
            uint32_t w = 0;
            for (int k = 0; k < 2; ++k) {
                const bool leaf_ok = li < n, int_ok = ii < made;
                bool take_leaf;
                if (leaf_ok && int_ok) take_leaf = s->cnt[s->order[li]] <= s->nodew[ii];
                else take_leaf = leaf_ok;
                if (take_leaf) { w += s->cnt[s->order[li]]; s->parent[li] = (uint16_t)(n + made); ++li; }
                else { w += s->nodew[ii]; s->parent[n + ii] = (uint16_t)(n + made); ++ii; }
            }
            s->nodew[made] = w;
        }
        // depths: the root is node 2n-2; a parent always has a larger index than its children
        // (the depth of an internal node is kept in nodew, whose weights are no longer needed)
        s->nodew[n - 2] = 0;
        for (int k = n - 3; k >= 0; --k) s->nodew[k] = s->nodew[s->parent[n + k] - n] + 1;
    }
    __syncwarp();
    // leaf lengths (limit 11) and the Kraft sum in units of 2^-11
    int kraft = 0;
    for (int i = (int)lane; i < n; i += 32) {
        int d = (int)s->nodew[s->parent[i] - n] + 1;
        if (d > 11) d = 11;
        s->len[s->order[i]] = (uint8_t)d;
        kraft += 1 << (11 - d);
    }
    kraft = (int)__reduce_add_sync(0xffffffffu, (uint32_t)kraft);
    __syncwarp();
    if (kraft != 2048) {
        // length limit 11: repair the Kraft sum.  Too large -> lengthen the rarest symbols that can still grow; too small -> shorten, each
        // time the most frequent symbol whose gain fits (lane 0: a few steps, each depends on the one before)
        uint32_t ok = 1;
        if (lane == 0) {
            int guard = 0;
            // order[] is ascending by count and a symbol only ever grows to 11, so "the rarest symbol that can still grow" is a
            // pointer that moves one way: everything before it has reached 11 bits
            for (int pick = 0; kraft > 2048 && ok;) {
                while (pick < n && s->len[s->order[pick]] >= 11) ++pick;
                if (pick >= n) { ok = 0; break; }
                const int sym = s->order[pick];
                kraft -= 1 << (10 - s->len[sym]);
                s->len[sym]++;
            }
            while (kraft < 2048 && ok) {
                const int slack = 2048 - kraft;
                int pick = -1;
                for (int i = n - 1; i >= 0; --i) {  // most frequent first
                    const int L = s->len[s->order[i]];
                    if (L > 1 && (1 << (11 - L)) <= slack) { pick = i; break; }
                }
                if (pick < 0 || ++guard > 4096) { ok = 0; break; }
                const int sym = s->order[pick];
                kraft += 1 << (11 - s->len[sym]);
                s->len[sym]--;
            }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) return false;
        __syncwarp();
    }
    // the longest code, the census of lengths
    uint32_t longest = 0;
    for (int i = (int)lane; i < n; i += 32) { const uint32_t L = s->len[s->order[i]]; if (L > longest) longest = L; }
    const int max_len = (int)__reduce_max_sync(0xffffffffu, longest);
    scratch[lane] = 0;
    __syncwarp();
    for (int i = (int)lane; i < n; i += 32) atomicAdd(&scratch[max_len + 1 - s->len[s->order[i]]], 1u);
    __syncwarp();
    // canonical code in the decoder's table order: longest codes (weight 1) first, symbols ascending within a weight
    uint32_t ok = 1;
    if (lane == 0) {
        uint32_t acc = 0;
        for (int w = 1; w <= max_len; ++w) { scratch[16 + w] = acc; acc += scratch[w] << (w - 1); }
        if (acc != (1u << max_len)) ok = 0;
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) return false;
    __syncwarp();
    // a row of 32 consecutive symbols at a time: a symbol's place among those of its weight = what earlier rows used up (kept in
    // scratch[16 + w]) + the lanes before it in the row with the same weight
    for (int row = 0; row < 8; ++row) {
        const int sym = row * 32 + (int)lane;
        const int L = s->cnt[sym] ? s->len[sym] : 0;
        const int w = L ? max_len + 1 - L : 0;
        const uint32_t peers = __match_any_sync(0xffffffffu, w);
        if (!L) { s->code[sym] = 0; s->len[sym] = 0; }
        else {
            const uint32_t idx = (uint32_t)__popc((int)(peers & ((1u << lane) - 1u)));
            const uint32_t first = scratch[16 + w];
            s->code[sym] = ((first >> (w - 1)) + idx) | ((uint32_t)L << 16);
        }
        __syncwarp();
        if (L && lane == (uint32_t)__ffs((int)peers) - 1) scratch[16 + w] += (uint32_t)__popc((int)peers) << (w - 1);
        __syncwarp();
    }
    if (lane == 0) s->max_bits = (uint32_t)max_len;
    __syncwarp();
    return true;
}

// ---- tree description (one thread) -> s->tree, returns its size in bytes (0 = cannot be described) -------------------
// All lanes call: the weights and their census are spread over the warp; the description proper - a few dozen values, then a chain of FSE
// states - is lane 0's.  Returns (in every lane) the size in bytes, 0 = cannot be described.
__device__ uint32_t zenc_describe_tree_lane0(ZCode* s, int nw, const uint32_t* census);
__device__ uint32_t zenc_describe_tree(ZCode* s, uint32_t lane)
{
    const int max_bits = (int)s->max_bits;
    uint32_t top = 0;
    for (int k = 0; k < 8; ++k) { const uint32_t sym = (uint32_t)k * 32 + lane; if (s->len[sym]) top = sym; }
    int last = (int)__reduce_max_sync(0xffffffffu, top);
    const int nw = last;  // weights of symbols 0..last-1; the last one is implied
    uint32_t* census = s->nodew + 224;  // (the code is built: nodew is free)
    census[lane] = 0;
    __syncwarp();
    for (int i = (int)lane; i < nw; i += 32) {
        const uint8_t w = s->len[i] ? (uint8_t)(max_bits + 1 - s->len[i]) : 0;
        s->weights[i] = w;
        atomicAdd(&census[w], 1u);
    }
    __syncwarp();
    uint32_t bytes = 0;
    if (lane == 0) bytes = zenc_describe_tree_lane0(s, nw, census);
    bytes = __shfl_sync(0xffffffffu, bytes, 0);
    __syncwarp();
    return bytes;
}
__device__ uint32_t zenc_describe_tree_lane0(ZCode* s, int nw, const uint32_t* census)
{
    uint32_t direct = 0;
    if (nw <= 128) direct = 1 + (uint32_t)(nw + 1) / 2;
    uint32_t fse_bytes = 0;
    // An alphabet of up to 128 symbols (control bytes, literals of S groups) sends its weights as they are, 4 bits each: some 30 bytes more
    // per tree than FSE-compressed ones (+0.03 % on signal), but neither this lane nor the decoder's has to walk a chain of FSE states for
    // it - and half of all trees are of this kind (encode 88.8 -> 86.4 ms, decode 66.9 -> 65.1 ms per 100,000 reads).
    if (nw >= 2 && !(direct && zm_tune("VBZ_TUNE_DIRECTW", 1u))) {
        // FSE-compress the weights: accuracy log 6, every present value gets at least one cell
        int hist[13];
        for (int w = 0; w < 13; ++w) hist[w] = (int)census[w];
        int maxw = 12;
        while (maxw > 0 && hist[maxw] == 0) --maxw;
        int distinct = 0;
        for (int w = 0; w <= maxw; ++w) distinct += hist[w] != 0;
        int norm[13], sum = 0, big = 0;
        for (int w = 0; w <= maxw; ++w) {
            norm[w] = hist[w] ? (hist[w] * 64) / nw : 0;
            if (hist[w] && norm[w] == 0) norm[w] = 1;
            sum += norm[w];
            if (hist[w] > hist[big]) big = w;
        }
        if (distinct == 1) {  // a lone value would own the whole table and its states would read no bits: lend one cell
            const int other = big == 0 ? 1 : big - 1;
            norm[other] = 1;
            if (other > maxw) maxw = other;
            sum += 1;
        }
        // bring the sum to 64 on the most frequent value; if that is impossible spread the correction
        int diff = 64 - sum;
        if (norm[big] + diff >= 1) norm[big] += diff;
        else {
            while (diff < 0) {
                int pick = -1;
                for (int w = 0; w <= maxw; ++w)
                    if (norm[w] > 1 && (pick < 0 || norm[w] > norm[pick])) pick = w;
                if (pick < 0) return direct ? direct : 0;
                norm[pick]--;
                diff++;
            }
        }
        const int nsymw = maxw + 1;
        for (int w = 0; w < nsymw; ++w) s->norm[w] = (int16_t)norm[w];
        // table description (forward bit stream)
        ZBitW bw;
        zw_init(bw, s->tree + 1);
        zw_put(bw, 6 - 5, 4);
        {
            int remaining = 64 + 1, thr = 64, nb = 7, sym = 0;
            while (remaining > 1 && sym < nsymw) {
                const int c = norm[sym];
                const int max = (2 * thr - 1) - remaining;
                int v = c + 1;
                if (v < max) zw_put(bw, (uint32_t)v, (uint32_t)(nb - 1));
                else {
                    if (v >= thr) v += max;
                    zw_put(bw, (uint32_t)v, (uint32_t)nb);
                }
                remaining -= c;
                ++sym;
                if (c == 0) {  // how many more zero-probability values follow, in 2-bit chunks (3 = "three and go on")
                    int zeros = 0;
                    while (sym + zeros < nsymw && norm[sym + zeros] == 0) ++zeros;
                    sym += zeros;
                    while (zeros >= 3) { zw_put(bw, 3, 2); zeros -= 3; }
                    zw_put(bw, (uint32_t)zeros, 2);
                }
                while (remaining < thr) { nb--; thr >>= 1; }
            }
            if (remaining != 1) return direct ? direct : 0;
        }
        const uint32_t desc_bytes = zw_close(bw, false);
        // decoding table, and per value the list of its cells in table order
        if (!z_fse_build(s->wcell, s->next, s->norm, nsymw, 6)) return direct ? direct : 0;
        int filled[13];
        for (int w = 0; w < 13; ++w) filled[w] = 0;
        for (int u = 0; u < 64; ++u) {
            const int w = (int)(s->wcell[u] & 0xFF);
            s->cellidx[w][filled[w]++] = (uint8_t)u;
        }
        // choose the cells from the last weight back to the first; the states alternate (even index: state 1)
        ZBitW sw;
        zw_init(sw, s->tree + 1 + desc_bytes);
        int cell_a = s->cellidx[s->weights[nw - 1]][0];  // cell of weight i + 2 while walking down
        int cell_b = s->cellidx[s->weights[nw - 2]][0];  // cell of weight i + 1
        if (((s->wcell[cell_b] >> 8) & 0xFF) == 0) return direct ? direct : 0;  // the decoder must run dry on this state
        for (int i = nw - 3; i >= 0; --i) {
            const int w = s->weights[i];
            const int c = norm[w];
            const uint32_t X = (uint32_t)cell_a + 64u;  // the state that follows this symbol in its chain is cell_a
            int k = z_highbit(X) - z_highbit((uint32_t)c);
            if ((int)(X >> k) < c) --k;
            const int x = (int)(X >> k);
            zw_put(sw, X & ((1u << k) - 1u), (uint32_t)k);
            const int cell_i = s->cellidx[w][x - c];
            cell_a = cell_b;
            cell_b = cell_i;
        }
        // now cell_b is the cell of weight 0 (state 1), cell_a of weight 1 (state 2); state 1 is read first
        zw_put(sw, (uint32_t)cell_a, 6);
        zw_put(sw, (uint32_t)cell_b, 6);
        const uint32_t stream_bytes = zw_close(sw, true);
        fse_bytes = 1 + desc_bytes + stream_bytes;
        if (desc_bytes + stream_bytes >= 128) fse_bytes = 0;
        else s->tree[0] = (uint8_t)(desc_bytes + stream_bytes);
    }
    if (fse_bytes && (!direct || fse_bytes <= direct)) return fse_bytes;
    if (!direct) return 0;
    s->tree[0] = (uint8_t)(127 + nw);
    for (int i = 0; i < nw; i += 2) s->tree[1 + i / 2] = (uint8_t)((s->weights[i] << 4) | (i + 1 < nw ? s->weights[i + 1] : 0));
    return direct;
}

// ---- a lane's run of symbols ------------------------------------------------------------------------------------
// bytes [p, p + n), n <= 64, any alignment -> r[0..15], little-endian (bytes past n are unspecified).  Aligned
// 128-bit loads; the 16-byte granules that hold the first and the last byte are read whole.
__device__ __forceinline__ void zenc_load_run(const uint8_t* p, uint32_t n, uint32_t (&r)[16])
{
    const uint4* q = reinterpret_cast<const uint4*>((uintptr_t)p & ~(uintptr_t)15);
    const uint32_t sh = (uint32_t)((uintptr_t)p & 15u);
    const uint32_t chunks = n ? (sh + n + 15) >> 4 : 0;  // 0..5
    uint32_t w[20];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        uint4 t = make_uint4(0, 0, 0, 0);
        if ((uint32_t)k < chunks) t = q[k];
        w[4 * k] = t.x; w[4 * k + 1] = t.y; w[4 * k + 2] = t.z; w[4 * k + 3] = t.w;
    }
    const uint32_t wo = sh >> 2, bs = 8u * (sh & 3u);
    if (wo == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(w[i], w[i + 1], bs);
    } else if (wo == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(w[i + 1], w[i + 2], bs);
    } else if (wo == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(w[i + 2], w[i + 3], bs);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(w[i + 3], w[i + 4], bs);
    }
}
// the same in two steps, so that the loads of the next run can be in flight while this one is processed
struct ZRun { uint4 g[5]; uint32_t sh; };
__device__ __forceinline__ void zenc_fetch_run(const uint8_t* p, uint32_t n, ZRun& f)
{
    const uint4* q = reinterpret_cast<const uint4*>((uintptr_t)p & ~(uintptr_t)15);
    f.sh = (uint32_t)((uintptr_t)p & 15u);
    const uint32_t chunks = n ? (f.sh + n + 15) >> 4 : 0;  // 0..5
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        f.g[k] = make_uint4(0, 0, 0, 0);
        if ((uint32_t)k < chunks) f.g[k] = q[k];
    }
}
__device__ __forceinline__ void zenc_align_run(const ZRun& f, uint32_t (&r)[16])
{
    uint32_t w[20];
#pragma unroll
    for (int k = 0; k < 5; ++k) { w[4 * k] = f.g[k].x; w[4 * k + 1] = f.g[k].y; w[4 * k + 2] = f.g[k].z; w[4 * k + 3] = f.g[k].w; }
    const uint32_t wo = f.sh >> 2, bs = 8u * (f.sh & 3u);
    if (wo == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(w[i], w[i + 1], bs);
    } else if (wo == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(w[i + 1], w[i + 2], bs);
    } else if (wo == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(w[i + 2], w[i + 3], bs);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __funnelshift_r(w[i + 3], w[i + 4], bs);
    }
}
// Lane l of the stream's warp owns the symbols [count - (l+1) C, count - l C), C = ceil(count / 32) <= 64: the
// stream is written from its last symbol up, so lane 0 writes the lowest bits.
__device__ __forceinline__ void zenc_lane_run(uint32_t count, uint32_t lane, uint32_t* lo, uint32_t* n)
{
    const uint32_t C = (count + 31) / 32;
    const uint32_t hi = count > lane * C ? count - lane * C : 0;
    *lo = hi > C ? hi - C : 0;
    *n = hi - *lo;
}
__device__ __forceinline__ uint32_t zenc_run_bits(const uint8_t* len, const uint32_t (&r)[16], uint32_t n)
{
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i)
        if ((uint32_t)i < n) bits += len[(r[i >> 2] >> (8 * (i & 3))) & 0xFFu];
    return bits;
}
// the run's codes, last symbol first, OR-ed into the zeroed staging words from bit `bit0` on
__device__ __forceinline__ void zenc_run_write(const uint32_t* code, const uint32_t (&r)[16], uint32_t n, uint32_t bit0, uint32_t* words)
{
    uint32_t wi = bit0 >> 5, nb = bit0 & 31u;
    uint64_t acc = 0;
#pragma unroll
    for (int k = 31; k >= 0; --k) {
        if ((uint32_t)(2 * k) < n) {  // symbols 2k+1 (if it exists) and 2k, in that order
            const uint32_t c0 = code[(r[k >> 1] >> (16 * (k & 1))) & 0xFFu];
            uint32_t v = c0 & 0xFFFFu, l = c0 >> 16;
            if ((uint32_t)(2 * k + 1) < n) {
                const uint32_t c1 = code[(r[k >> 1] >> (16 * (k & 1) + 8)) & 0xFFu];
                v = (c1 & 0xFFFFu) | (v << (c1 >> 16));
                l += c1 >> 16;
            }
            acc |= (uint64_t)v << nb;
            nb += l;
            if (nb >= 32) {
                atomicOr(words + wi, (uint32_t)acc);
                ++wi;
                acc >>= 32;
                nb -= 32;
            }
        }
    }
    if (nb) atomicOr(words + wi, (uint32_t)acc);
}

// `nthreads` threads (t = 0 .. nthreads-1) copy n bytes from a 16-byte aligned staging area to an arbitrarily aligned
// global address: aligned 128-bit stores, each fed by two aligned 128-bit shared loads re-aligned with funnel shifts
// (the staging area must be readable 32 bytes past the end).
__device__ __forceinline__ void zenc_copy_out(uint8_t* g, const uint8_t* stage, uint32_t n, uint32_t t, uint32_t nthreads)
{
    uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
    if (head > n) head = n;
    const uint32_t nchunks = (n - head) >> 4, done = head + (nchunks << 4);
    if (t < head) g[t] = stage[t];
    else if (t >= 32 && t - 32 < n - done) g[done + t - 32] = stage[done + t - 32];
    const uint4* s4 = reinterpret_cast<const uint4*>(stage);
    uint4* g4 = reinterpret_cast<uint4*>(g + head);
    const uint32_t wsel = head >> 2, sh = 8u * (head & 3u);
    for (uint32_t c = t; c < nchunks; c += nthreads) {
        const uint4 A = s4[c], B = s4[c + 1];
        uint32_t w0, w1, w2, w3, w4;
        if (wsel == 0) { w0 = A.x; w1 = A.y; w2 = A.z; w3 = A.w; w4 = B.x; }
        else if (wsel == 1) { w0 = A.y; w1 = A.z; w2 = A.w; w3 = B.x; w4 = B.y; }
        else if (wsel == 2) { w0 = A.z; w1 = A.w; w2 = B.x; w3 = B.y; w4 = B.z; }
        else { w0 = A.w; w1 = B.x; w2 = B.y; w3 = B.z; w4 = B.w; }
        g4[c] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
    }
}

__global__ void zstd_enc_fill_blocks_kernel(const ZstdEncRead* reads, uint32_t n_reads, uint32_t* block_read)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_reads) return;
    const ZstdEncRead rd = reads[warp];
    for (uint32_t i = lane; i < rd.n_blocks; i += 32) block_read[rd.first_block + i] = warp;
}

// geometry of sub-block q of a group of B bytes
struct ZSub { uint32_t start, bytes, streams, seg; };
__device__ __forceinline__ ZSub zenc_sub(uint32_t B, uint32_t q, uint32_t slice = kZEncSub)
{
    ZSub s;
    s.start = q * slice;
    s.bytes = B - s.start < slice ? B - s.start : slice;
    s.streams = s.bytes < 1024 ? 1u : 4u;
    s.seg = s.streams == 1 ? s.bytes : (s.bytes + 3) / 4;
    return s;
}

// the lanes of a warp fetch their runs of stream sidx (= 4 * sub-block + stream) of a group of B bytes; n = 0 if there is none
__device__ __forceinline__ void zenc_fetch_stream(const uint8_t* src, uint32_t B, uint32_t sidx, uint32_t nstreams, uint32_t lane, ZRun& f, uint32_t* n_out, uint32_t slice)
{
    uint32_t n = 0;
    f.sh = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) f.g[k] = make_uint4(0, 0, 0, 0);
    if (sidx < nstreams) {
        const ZSub sb = zenc_sub(B, sidx >> 2, slice);
        const uint32_t j = sidx & 3u;
        if (j < sb.streams && sb.bytes >= 32) {
            const uint32_t cnt = (sb.streams == 4 && j == 3) ? sb.bytes - 3 * sb.seg : sb.seg;
            uint32_t lo;
            zenc_lane_run(cnt, lane, &lo, &n);
            zenc_fetch_run(src + sb.start + j * sb.seg + lo, n, f);
        }
    }
    *n_out = n;
}

// ---- kernel 1: histogram + Huffman code of every group, a warp per group -------------------------------------------------
constexpr int kZcodeWarps = 1;
__global__ void __launch_bounds__(kZcodeWarps * 32) zstd_enc_code_kernel(ZstdEncParams p)
{
    VBZ_DYN_SMEM(smem);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ZCode* s = reinterpret_cast<ZCode*>(smem) + wid;
    const uint32_t b = blockIdx.x * kZcodeWarps + wid;
    if (b >= p.n_blocks) return;
    const uint32_t r = p.block_read[b];
    const ZstdEncRead rd = p.reads[r];
    ZGroupLoc g;
    if (!zenc_locate(p, b, rd, r, g) || g.code_first != 0) return;  // the first group of a code set builds the code for all of them
    const uint32_t B = g.code_bytes;
    for (uint32_t i = lane; i < 256; i += 32) s->cnt[i] = 0;
    __syncwarp();
    // ---- histogram of the set's LITERALS.  L groups (no sequences): their bytes, with 128-bit loads over the aligned body (four in
    // flight per lane) and byte loads at the two ends.  S groups: the match kernel counted their literals (csrc/zstd_match_kernels.cuh).
    // Large code sets are SAMPLED: one 512-byte row in four (from 256 KiB) or in eight (from 512 KiB) is counted (the shared-memory atomic unit bounds this
    // kernel), the counts are scaled and every byte value gets a floor of one, so that a value the sample missed still has a code.
    // A sample that saw a single value is checked against the whole set: constant content stays one RLE block per group.
    const bool sampled = B >= (256u << 10);
    const uint32_t rate = B >= (512u << 10) ? 8u : 4u;  // one row in `rate`
    const ZGroupMeta* meta = reinterpret_cast<const ZGroupMeta*>(p.meta) + b;
    const uint32_t set_groups = (B + g.gsize - 1) / g.gsize;
    bool any_s = false, any_l = false;
    {
        uint32_t* h = s->cnt;
        uint32_t zeros = 0;
        for (uint32_t j = 0; j < set_groups; ++j) {
            if (meta[j].mode == kZModeS) { any_s = true; continue; }
            any_l = true;
            const uint8_t* src = g.src + j * g.gsize;
            const uint32_t gB = B - j * g.gsize < g.gsize ? B - j * g.gsize : g.gsize;
            uint32_t head = (16u - (uint32_t)((uintptr_t)src & 15u)) & 15u;
            if (head > gB) head = gB;
            const uint32_t nch = (gB - head) >> 4, done = head + (nch << 4);
            if (lane < head) atomicAdd(h + src[lane], 1u);
            if (lane < gB - done) atomicAdd(h + src[done + lane], 1u);
            const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
            const uint32_t spacing = sampled ? 32u * rate : 32u;  // chunks between the rows a lane has in flight
            for (uint32_t c0 = 0; c0 < nch; c0 += 4 * spacing) {
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t c = c0 + spacing * u + lane;
                    q[u] = make_uint4(0, 0, 0, 0);
                    if (c < nch) q[u] = s4[c];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t c = c0 + spacing * u + lane;
                    if (c >= nch) continue;
                    const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const uint32_t v = (w[k] >> (8 * jj)) & 0xFFu;
                            if (v) atomicAdd(h + v, 1u);
                            else ++zeros;
                        }
                    }
                }
            }
        }
        if (zeros) atomicAdd(h, zeros);
        __syncwarp();
        if (sampled && any_l) {
            uint32_t seen = 0, which = 0;
            for (uint32_t i = lane; i < 256; i += 32)
                if (s->cnt[i]) { ++seen; which = i; }
            seen = __reduce_add_sync(0xffffffffu, seen);
            bool constant = false;
            if (seen == 1 && !any_s) {  // one value in the sample: is it the only one in the whole set?
                which = __reduce_add_sync(0xffffffffu, which);  // (one lane holds it, the others 0)
                const uint32_t rep = which * 0x01010101u;
                uint32_t diff = 0;
                const uint8_t* src = g.src;
                uint32_t head = (16u - (uint32_t)((uintptr_t)src & 15u)) & 15u;
                if (head > B) head = B;
                const uint32_t nch = (B - head) >> 4, done = head + (nch << 4);
                const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
                for (uint32_t c = lane; c < nch; c += 32) {
                    const uint4 t = s4[c];
                    diff |= (t.x ^ rep) | (t.y ^ rep) | (t.z ^ rep) | (t.w ^ rep);
                }
                if (lane < head) diff |= src[lane] ^ which;
                if (lane < B - done) diff |= src[done + lane] ^ which;
                constant = !__any_sync(0xffffffffu, diff != 0);
            }
            if (!constant)
                for (uint32_t i = lane; i < 256; i += 32) s->cnt[i] = s->cnt[i] * rate + 1u;
        }
        __syncwarp();
        if (any_s)
            for (uint32_t j = 0; j < set_groups; ++j)
                if (meta[j].mode == kZModeS)
                    for (uint32_t i = lane; i < 256; i += 32) s->cnt[i] += p.hist[(uint64_t)(b + j) * 256 + i];
    }
    __syncwarp();
    // ---- rank sort of the symbols that occur, ascending by (count, symbol): a lane ranks 8 symbols --------------------------
    uint32_t nsym = 0;
    {
        uint32_t c[8], rank[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = s->cnt[lane * 8 + k]; rank[k] = 0; }
        for (uint32_t j = 0; j < 256; ++j) {
            const uint32_t cj = s->cnt[j];
            if (cj == 0) continue;
#pragma unroll
            for (int k = 0; k < 8; ++k) rank[k] += (cj < c[k] || (cj == c[k] && j < lane * 8 + k)) ? 1u : 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (c[k]) { s->order[rank[k]] = (uint8_t)(lane * 8 + k); ++nsym; }
    }
    nsym = __reduce_add_sync(0xffffffffu, nsym);
    __syncwarp();
    // ---- the code and the kind of group (lane 0) -------------------------------------------------------------------------
    uint32_t kind = kZKindRaw, tree_bytes = 0;
    if (B > 0 && nsym == 1 && !any_s) kind = kZKindRle;
    else if (nsym >= 2 && B >= 32 && zenc_build_code(s, (int)nsym, lane)) {
        tree_bytes = zenc_describe_tree(s, lane);
        if (tree_bytes) kind = kZKindHuf;
    }
    __syncwarp();
    ZGroupCode* out = reinterpret_cast<ZGroupCode*>(p.codes) + b;
    if (kind == kZKindHuf) {
        for (uint32_t i = lane; i < 256; i += 32) out->code[i] = s->code[i];
        for (uint32_t i = lane; i < 160 / 4; i += 32) reinterpret_cast<uint32_t*>(out->tree)[i] = reinterpret_cast<const uint32_t*>(s->tree)[i];
    }
    if (lane == 0) { out->kind = kind; out->tree_bytes = tree_bytes; out->nsym = nsym; }
}

// `nthreads` threads copy n bytes between two arbitrarily aligned global addresses (the sequences sections: a few KB at most)
__device__ __forceinline__ void zenc_copy_global(uint8_t* d, const uint8_t* sp, uint32_t n, uint32_t t, uint32_t nthreads)
{
    uint32_t head = (4u - (uint32_t)((uintptr_t)d & 3u)) & 3u;
    if (head > n) head = n;
    const uint32_t words = (n - head) >> 2, done = head + 4 * words;
    if (t < head) d[t] = sp[t];
    else if (t >= 32 && t - 32 < n - done) d[done + t - 32] = sp[done + t - 32];
    uint32_t* d4 = reinterpret_cast<uint32_t*>(d + head);
    const uint8_t* s0 = sp + head;
    for (uint32_t k = t; k < words; k += nthreads) d4[k] = zm_load32(s0 + 4 * k);
}

// ---- kernel 2: sizes, position, pack - a CTA per group --------------------------------------------------------------------
// An S group (csrc/zstd_match_kernels.cuh) runs through the same three steps with its compacted LITERALS as the bytes to code:
// block q holds literals [8 KiB q, 8 KiB (q + 1)) of the group, as a Huffman or Raw literals section, followed by the block's
// sequences section, which zstd_enc_seq_kernel left behind the literals in the literal scratch.
__global__ void __launch_bounds__(kZencThreads) zstd_encode_kernel(ZstdEncParams p)
{
    VBZ_DYN_SMEM(smem);
    ZEncSmem* s = reinterpret_cast<ZEncSmem*>(smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t b = blockIdx.x;
    const uint32_t r = p.block_read[b];
    const ZstdEncRead rd = p.reads[r];
    ZGroupLoc g;
    if (!zenc_locate(p, b, rd, r, g)) {
        if (g.gi == 0 && tid == 0) {
            if (rd.rejected) p.results[r] = rd.preset;
            else if (p.src_sizes && g.content >= kErrOutOfMemory) p.results[r] = g.content;
        }
        return;
    }
    const uint32_t gi = g.gi, content = g.content;
    const bool last = g.last;
    const ZGroupMeta* meta = reinterpret_cast<const ZGroupMeta*>(p.meta) + b;
    const bool smode = meta->mode == kZModeS;
    // the bytes to entropy-code: the group itself, or the literals its matches left over
    const uint8_t* slot = p.lits + (uint64_t)(g.src - p.src);
    const uint8_t* src = smode ? slot : g.src;
    const uint32_t B = smode ? meta->lits : g.B;
    const uint32_t slice = smode ? 1u << zm_slice_shift(g.keys) : kZEncSub;  // literals per block
    const uint32_t nsub = smode ? meta->nblk : (B ? (B + kZEncSub - 1) / kZEncSub : 1u);
    // the group's code, built by zstd_enc_code_kernel
    const ZGroupCode* gc = reinterpret_cast<const ZGroupCode*>(p.codes) + (b - g.code_first);
    const bool huf_ok = gc->kind == kZKindHuf;
    uint32_t kind = smode ? (uint32_t)kZKindHuf : gc->kind;  // (an S group always takes the block-by-block path; its literals may still be stored raw)
    const uint32_t tree_bytes = gc->tree_bytes;
    if (huf_ok) {
        const uint32_t c = gc->code[tid];
        s->code[tid] = c;
        s->len[tid] = (uint8_t)(c >> 16);
        if (tid < 160 / 4) reinterpret_cast<uint32_t*>(s->tree)[tid] = reinterpret_cast<const uint32_t*>(gc->tree)[tid];
    }
    if (smode && tid < kZSubs) { s->seq_off[tid] = meta->seq_off[tid]; s->seq_len[tid] = meta->seq_len[tid]; }
    // bytes before the first block header: the sized header and the frame header, in front of the frame's first group
    uint32_t pre = 0;
    uint8_t prebuf[16];
    if (gi == 0) {
        if (p.hdr_bytes) { prebuf[0] = (uint8_t)rd.hdr_value; prebuf[1] = (uint8_t)(rd.hdr_value >> 8); prebuf[2] = (uint8_t)(rd.hdr_value >> 16); prebuf[3] = (uint8_t)(rd.hdr_value >> 24); pre = 4; }
        const uint32_t fcs_bytes = content < 256 ? 1u : (content < 65536 + 256 ? 2u : 4u);
        prebuf[pre] = 0x28; prebuf[pre + 1] = 0xB5; prebuf[pre + 2] = 0x2F; prebuf[pre + 3] = 0xFD;
        prebuf[pre + 4] = (uint8_t)(0x20 | (fcs_bytes == 1 ? 0x00 : fcs_bytes == 2 ? 0x40 : 0x80));  // single segment, content size present
        const uint32_t f = fcs_bytes == 2 ? content - 256 : content;
        for (uint32_t k = 0; k < fcs_bytes; ++k) prebuf[pre + 5 + k] = (uint8_t)(f >> (8 * k));
        pre += 5 + fcs_bytes;
    }
    __syncthreads();
    uint16_t* lane_off = &s->lane_off[0][0];  // [stream][lane]

    // ---- sizes: every stream of the group, a warp per stream ----------------------------------------------------------------
    if (kind == kZKindHuf && huf_ok) {
        ZRun f;
        uint32_t fn;
        zenc_fetch_stream(src, B, wid, nsub * 4, lane, f, &fn, slice);
        for (uint32_t sidx = wid; sidx < nsub * 4; sidx += kZencWarps) {
            uint32_t run[16];
            const uint32_t n = fn;
            zenc_align_run(f, run);
            zenc_fetch_stream(src, B, sidx + kZencWarps, nsub * 4, lane, f, &fn, slice);  // in flight while this stream is summed
            const uint32_t bits = zenc_run_bits(s->len, run, n);
            const uint32_t incl = warp_scan_u32(bits);
            lane_off[sidx * 32 + lane] = (uint16_t)(incl - bits);
            if (lane == 31) s->sbits[sidx] = incl;
        }
    }
    __syncthreads();
    // ---- sub-block kinds and sizes (one thread) -----------------------------------------------------------------------------
    if (tid == 0) {
        uint32_t off = pre;
        if (kind == kZKindHuf) {
            bool tree_out = false, any = false;
            for (uint32_t q = 0; q < nsub; ++q) {
                const ZSub sb = zenc_sub(B, q, slice);
                // the literals stored as they are: a Raw block (L), or a Raw literals section with a 1 / 2 / 3-byte header (S)
                const uint32_t raw_lit = smode ? (sb.bytes < 32 ? 1u : sb.bytes < 4096 ? 2u : 3u) + sb.bytes : sb.bytes;
                const uint32_t seq_sec = smode && s->seq_len[q] ? s->seq_len[q] : 1u;  // 1 = "0 sequences"
                uint32_t payload = smode ? raw_lit + seq_sec : sb.bytes;
                s->sub_kind[q] = kZKindRaw;
                s->sub_tree[q] = 0;
                s->sub_lit[q] = raw_lit;
                if (huf_ok && sb.bytes >= 32) {
                    uint32_t comp = (tree_out ? 0u : tree_bytes) + (sb.streams == 4 ? 6u : 0u);
                    for (uint32_t j = 0; j < sb.streams; ++j) comp += (s->sbits[q * 4 + j] >> 3) + 1;  // + the marker bit
                    const uint32_t lit_hdr = sb.streams == 1 ? 3u : 4u;  // 4 streams: 1024 <= regenerated size <= 8192 and comp below it
                    const uint32_t huf = lit_hdr + comp + seq_sec;
                    if (huf < payload && comp < (sb.streams == 1 ? 1024u : 16384u)) {
                        payload = huf;
                        s->sub_lit[q] = lit_hdr + comp;
                        s->sub_kind[q] = kZKindHuf;
                        s->sub_tree[q] = tree_out ? 0 : 1;
                        tree_out = true;
                        any = true;
                    }
                }
                s->sub_off[q] = off;
                off += 3 + payload;
            }
            s->sub_off[nsub] = off;
            if (!any && !smode) { kind = kZKindRaw; off = pre; }  // nothing gained anywhere: one Raw block for the group
        }
        if (kind == kZKindRle) off = pre + 3 + 1;
        else if (kind == kZKindRaw) off = pre + 3 + B;
        s->v[kZvKind] = kind;
        s->v[kZvTotal] = off;
    }
    __syncthreads();
    kind = s->v[kZvKind];
    const uint32_t total = s->v[kZvTotal];

    // ---- position inside the frame: look-back over the frame's earlier groups -------------------------------------------------
    if (wid == 0) {
        uint64_t excl = 0;
        if (gi == 0) {
            if (lane == 0) vbz_st_volatile_u64(p.status + b, kStInclusive | total);
        } else {
            if (lane == 0) vbz_st_volatile_u64(p.status + b, kStAggregate | total);
            excl = lookback_exclusive(p.status, b, rd.first_block);
            if (lane == 0) vbz_st_volatile_u64(p.status + b, kStInclusive | ((excl + total) & kStValueMask));
        }
        if (lane == 0) s->excl = excl;
    }
    __syncthreads();
    const uint64_t excl = s->excl;
    const uint64_t end = excl + total;
    const bool fits = end <= rd.dst_cap;
    if (last && tid == 0) p.results[r] = fits ? (uint32_t)end : kErrZstd + p.err_add;  // ZSTD_compress: dstSize_tooSmall (vbz.cpp:201-204)
    uint8_t* dst = p.dst + rd.dst_off + excl;

    if (kind != kZKindHuf) {
        if (!fits) return;
        const uint32_t bh = (last ? 1u : 0u) | ((kind == kZKindRle ? 1u : 0u) << 1) | (B << 3);
        if (tid == 0) {
            for (uint32_t k = 0; k < pre; ++k) dst[k] = prebuf[k];
            dst[pre] = (uint8_t)bh; dst[pre + 1] = (uint8_t)(bh >> 8); dst[pre + 2] = (uint8_t)(bh >> 16);
            if (kind == kZKindRle) dst[pre + 3] = src[0];
        }
        if (kind == kZKindRaw)
            for (uint32_t i = tid; i < B; i += kZencThreads) dst[pre + 3 + i] = src[i];
        return;
    }

    // ---- pack: two sub-blocks at a time, four warps (= four streams) each ---------------------------------------------------------
    const uint32_t half = wid >> 2, j = wid & 3u, ht = tid & 127u;
    uint8_t* stage = s->stage[half];
    for (uint32_t q0 = 0; q0 < nsub; q0 += 2) {
        const uint32_t q = q0 + half;
        const bool live = q < nsub;
        const ZSub sb = zenc_sub(B, live ? q : 0, slice);
        const uint32_t skind = live ? s->sub_kind[q] : kZKindRaw;
        const uint32_t front = (live && q == 0) ? pre : 0u;  // the frame's headers travel with its first sub-block
        for (uint32_t i = ht; i < (kZEncSub + 96) / 16; i += 128) reinterpret_cast<uint4*>(stage)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        uint8_t* blk = stage + front;
        uint32_t sbytes[4] = {0, 0, 0, 0};
        uint32_t lit_hdr = 0, tb = 0;
        if (live) {
            const uint32_t out_bytes = s->sub_off[q + 1] - s->sub_off[q];
            const uint32_t payload = out_bytes - 3;
            if (ht == 0) {
                for (uint32_t k = 0; k < front; ++k) stage[k] = prebuf[k];
                const bool blast = last && q + 1 == nsub;
                const uint32_t btype = (smode || skind == kZKindHuf) ? 2u : 0u;  // Compressed / Raw
                const uint32_t bh = (blast ? 1u : 0u) | (btype << 1) | (payload << 3);
                blk[0] = (uint8_t)bh; blk[1] = (uint8_t)(bh >> 8); blk[2] = (uint8_t)(bh >> 16);
            }
            if (skind == kZKindHuf) {
                tb = s->sub_tree[q] ? tree_bytes : 0u;
                lit_hdr = sb.streams == 1 ? 3u : 4u;
                uint32_t comp = tb + (sb.streams == 4 ? 6u : 0u);
                for (uint32_t k = 0; k < sb.streams; ++k) { sbytes[k] = (s->sbits[q * 4 + k] >> 3) + 1; comp += sbytes[k]; }
                uint8_t* lit = blk + 3;
                if (ht == 0) {
                    const uint32_t type = s->sub_tree[q] ? 2u : 3u;  // Compressed / Treeless
                    uint32_t v = type | ((sb.streams == 1 ? 0u : 2u) << 2) | (sb.bytes << 4);
                    v |= sb.streams == 1 ? (comp << 14) : (comp << 18);
                    for (uint32_t k = 0; k < lit_hdr; ++k) lit[k] = (uint8_t)(v >> (8 * k));
                }
                for (uint32_t k = ht; k < tb; k += 128) lit[lit_hdr + k] = s->tree[k];
                if (sb.streams == 4 && ht >= 32 && ht < 35) {
                    const uint32_t k = ht - 32;
                    lit[lit_hdr + tb + 2 * k] = (uint8_t)sbytes[k];
                    lit[lit_hdr + tb + 2 * k + 1] = (uint8_t)(sbytes[k] >> 8);
                }
            } else {
                uint32_t rh = 0;  // S: Raw literals section header (RFC 8878 3.1.1.3.1.1), L: the block is the bytes
                if (smode) {
                    rh = sb.bytes < 32 ? 1u : sb.bytes < 4096 ? 2u : 3u;
                    if (ht == 0) {
                        const uint32_t v = rh == 1 ? (sb.bytes << 3) : (rh == 2 ? ((sb.bytes << 4) | (1u << 2)) : ((sb.bytes << 4) | (3u << 2)));
                        for (uint32_t k = 0; k < rh; ++k) blk[3 + k] = (uint8_t)(v >> (8 * k));
                    }
                }
                const uint8_t* from = src + sb.start;
                for (uint32_t i = ht; i < sb.bytes; i += 128) blk[3 + rh + i] = from[i];
            }
        }
        __syncthreads();  // byte-wise header writes land before the word-wise stream writes next to them
        if (live && skind == kZKindHuf && j < sb.streams) {
            uint8_t* body = blk + 3 + lit_hdr + tb + (sb.streams == 4 ? 6u : 0u);
            uint32_t so = 0;
            for (uint32_t k = 0; k < j; ++k) so += sbytes[k];
            uint8_t* sp = body + so;
            // each stream starts at a byte position; the bit writer works on aligned words, so shift by the byte phase
            const uint32_t phase = (uint32_t)((uintptr_t)sp & 3u);
            uint32_t* words = reinterpret_cast<uint32_t*>(sp - phase);
            const uint32_t cnt = (sb.streams == 4 && j == 3) ? sb.bytes - 3 * sb.seg : sb.seg;
            uint32_t lo, n, run[16];
            zenc_lane_run(cnt, lane, &lo, &n);
            zenc_load_run(src + sb.start + j * sb.seg + lo, n, run);
            if (n) zenc_run_write(s->code, run, n, (uint32_t)lane_off[(q * 4 + j) * 32 + lane] + 8 * phase, words);
            if (lane == 0) {  // the marker bit above the last bit
                const uint32_t mb = s->sbits[q * 4 + j] + 8 * phase;
                atomicOr(words + (mb >> 5), 1u << (mb & 31u));
            }
        }
        __syncthreads();
        if (live) {
            const uint32_t o0 = q == 0 ? 0u : s->sub_off[q], o1 = s->sub_off[q + 1];  // sub-block 0 starts at 0 with the frame's headers
            if (excl + o1 <= rd.dst_cap) {
                const uint32_t seq_n = smode ? s->seq_len[q] : 0u;
                // staged: headers, the literals section and - when there are no sequences - the zero byte that says so
                zenc_copy_out(dst + o0, stage, o1 - o0 - seq_n, ht, 128);
                if (seq_n) zenc_copy_global(dst + o1 - seq_n, slot + s->seq_off[q], seq_n, ht, 128);
            }
        }
        __syncthreads();
    }
}

}  // namespace vbz
