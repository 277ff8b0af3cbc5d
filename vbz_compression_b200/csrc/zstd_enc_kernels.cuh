// csrc/zstd_enc_kernels.cuh - stage 2 of VBZ encode on the GPU: a Zstandard frame writer (RFC 8878).
//
// Replaces the reference's call ZSTD_compress(dst, cap, stage1, n, level) (ref vbz/vbz.cpp:194-204; libzstd is an
// un-vendored dependency, pinned zstd/1.4.8 by CMakeLists.txt:93).  The bytes libzstd would write are not
// reproducible across its versions, so the contract is the FORMAT: every frame written here carries
// its content size (vbz.cpp:236-240 demands it), is a single frame, and decodes with stock libzstd and with the
// reference's vbz_decompress to the same bytes - the parity tests check exactly that.
//
// What is emitted.  On nanopore signal zstd level 1 finds almost no matches: 98 % of its output is Huffman-coded
// literals.  This writer therefore entropy-codes and does not search: the content is cut
// into blocks of 32 KiB; every block is, whichever is smallest, an RLE block, a Compressed block holding one
// Huffman-coded literals section (own length-limited code, <= 11 bits, weights FSE-compressed or direct, 1 or 4
// streams) and no sequences, or a Raw block.  The compression level only has to be non-zero.
//
// Parallelism: one CTA (4 warps) per block, blocks numbered frame-major.
//   histogram   per-warp shared-memory histograms, merged
//   code        rank sort of the 256 counts by all threads; the two-queue Huffman construction, the length
//               limiting and the (small) FSE coding of the weights are serial work of one thread
//   streams     a warp per stream; each lane owns a contiguous run of symbols, sums its code lengths, a shuffle
//               scan gives its bit offset, and it writes its bits into the shared-memory staging area (plain
//               stores for the words it owns, atomicOr for the two it shares with its neighbours)
//   out         the size of the block is posted, a decoupled look-back over the frame's earlier blocks
//               (csrc/svb_kernels.cuh) gives its position, the CTA copies the staging area out
#pragma once
#include "zstd_dec_kernels.cuh"
#include "svb_kernels.cuh"

namespace vbz {

constexpr int kZencThreads = 128;
constexpr int kZencWarps = kZencThreads / 32;

struct ZEncSmem {
    uint32_t hist[kZencWarps][256];
    uint32_t cnt[256];
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
    uint32_t v[16];           // scalars shared by the CTA (see the kZv* indices)
    uint64_t excl;
    alignas(16) uint8_t stage[kZEncBlock + 64];  // sized header, frame header, block header, block payload
};
enum { kZvNsym = 0, kZvMaxBits, kZvTreeBytes, kZvKind, kZvStreams, kZvBits0, kZvBits1, kZvBits2, kZvBits3, kZvTotal };
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
__device__ bool zenc_build_code(ZEncSmem* s, int nsym)
{
    // two-queue construction over the leaves sorted by ascending count: leaves 0..nsym-1, internal nodes nsym..2nsym-2
    const int n = nsym;
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
    // (the depth of an internal node is kept in the low byte of nodew, which is no longer needed)
    s->nodew[n - 2] = 0;
    for (int k = n - 3; k >= 0; --k) s->nodew[k] = s->nodew[s->parent[n + k] - n] + 1;
    int max_len = 0;
    for (int i = 0; i < n; ++i) {
        int d = (int)s->nodew[s->parent[i] - n] + 1;
        if (d > 11) d = 11;
        s->len[s->order[i]] = (uint8_t)d;
        if (d > max_len) max_len = d;
    }
    // length limit 11: repair the Kraft sum (in units of 2^-11).  Too large -> lengthen the rarest symbols that can
    // still grow; too small -> shorten, each time the most frequent symbol whose gain fits.
    int kraft = 0;
    for (int i = 0; i < n; ++i) kraft += 1 << (11 - s->len[s->order[i]]);
    int guard = 0;
    // order[] is ascending by count and a symbol only ever grows to 11, so "the rarest symbol that can still grow" is a
    // pointer that moves one way: everything before it has reached 11 bits
    for (int pick = 0; kraft > 2048;) {
        while (pick < n && s->len[s->order[pick]] >= 11) ++pick;
        if (pick >= n) return false;
        const int sym = s->order[pick];
        kraft -= 1 << (10 - s->len[sym]);
        s->len[sym]++;
    }
    while (kraft < 2048) {
        const int slack = 2048 - kraft;
        int pick = -1;
        for (int i = n - 1; i >= 0; --i) {  // most frequent first
            const int L = s->len[s->order[i]];
            if (L > 1 && (1 << (11 - L)) <= slack) { pick = i; break; }
        }
        if (pick < 0 || ++guard > 4096) return false;
        const int sym = s->order[pick];
        kraft += 1 << (11 - s->len[sym]);
        s->len[sym]--;
    }
    max_len = 0;
    for (int i = 0; i < n; ++i)
        if (s->len[s->order[i]] > max_len) max_len = s->len[s->order[i]];
    // canonical code in the decoder's table order: longest codes (weight 1) first, symbols ascending within a weight
    uint32_t rank[13], start[13];
    for (int w = 0; w < 13; ++w) rank[w] = 0;
    for (int i = 0; i < n; ++i) rank[max_len + 1 - s->len[s->order[i]]]++;
    uint32_t acc = 0;
    for (int w = 1; w <= max_len; ++w) { start[w] = acc; acc += rank[w] << (w - 1); }
    if (acc != (1u << max_len)) return false;
    for (int sym = 0; sym < 256; ++sym) {
        const int L = s->cnt[sym] ? s->len[sym] : 0;
        if (!L) { s->code[sym] = 0; s->len[sym] = 0; continue; }
        const int w = max_len + 1 - L;
        s->code[sym] = (start[w] >> (w - 1)) | ((uint32_t)L << 16);
        start[w] += 1u << (w - 1);
    }
    s->v[kZvMaxBits] = (uint32_t)max_len;
    return true;
}

// ---- tree description (one thread) -> s->tree, returns its size in bytes (0 = cannot be described) -------------------
__device__ uint32_t zenc_describe_tree(ZEncSmem* s)
{
    const int max_bits = (int)s->v[kZvMaxBits];
    int last = 255;
    while (last > 0 && s->len[last] == 0) --last;
    const int nw = last;  // weights of symbols 0..last-1; the last one is implied
    for (int i = 0; i < nw; ++i) s->weights[i] = s->len[i] ? (uint8_t)(max_bits + 1 - s->len[i]) : 0;
    uint32_t direct = 0;
    if (nw <= 128) direct = 1 + (uint32_t)(nw + 1) / 2;
    uint32_t fse_bytes = 0;
    if (nw >= 2) {
        // FSE-compress the weights: accuracy log 6, every present value gets at least one cell
        int hist[13];
        for (int w = 0; w < 13; ++w) hist[w] = 0;
        for (int i = 0; i < nw; ++i) hist[s->weights[i]]++;
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

// ---- one Huffman stream (one warp): symbols src[0..count) -> bits at stage (zeroed), returns the bit count -----------
// Lane l owns the symbols [count - (l+1)*C, count - l*C) (the stream is written from the last symbol up).
__device__ __forceinline__ uint32_t zenc_stream_bits(const uint32_t* code, const uint8_t* src, uint32_t count, uint32_t lane, uint32_t* chunk_out)
{
    const uint32_t C = (count + 31) / 32;
    uint32_t bits = 0;
    const uint32_t hi = count > lane * C ? count - lane * C : 0, lo = hi > C ? hi - C : 0;
    for (uint32_t i = lo; i < hi; ++i) bits += code[src[i]] >> 16;
    *chunk_out = C;
    return bits;
}
__device__ __forceinline__ void zenc_stream_write(const uint32_t* code, const uint8_t* src, uint32_t count, uint32_t lane, uint32_t C, uint32_t bit0,
                                                  uint8_t* stage)
{
    // stage is 4-byte aligned; this lane's bits start at bit0
    uint32_t* words = reinterpret_cast<uint32_t*>(stage);
    const uint32_t hi = count > lane * C ? count - lane * C : 0, lo = hi > C ? hi - C : 0;
    if (hi == lo) return;
    uint32_t wi = bit0 >> 5;
    uint32_t nb = bit0 & 31u;
    uint64_t acc = 0;
    bool first = true;
    for (uint32_t i = hi; i-- > lo;) {
        const uint32_t c = code[src[i]];
        acc |= (uint64_t)(c & 0xFFFFu) << nb;
        nb += c >> 16;
        if (nb >= 32) {
            if (first) { atomicOr(words + wi, (uint32_t)acc); first = false; }
            else words[wi] = (uint32_t)acc;
            ++wi;
            acc >>= 32;
            nb -= 32;
        }
    }
    if (nb) atomicOr(words + wi, (uint32_t)acc);
}

__global__ void zstd_enc_fill_blocks_kernel(const ZstdEncRead* reads, uint32_t n_reads, uint32_t* block_read)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_reads) return;
    const ZstdEncRead rd = reads[warp];
    for (uint32_t i = lane; i < rd.n_blocks; i += 32) block_read[rd.first_block + i] = warp;
}

__global__ void __launch_bounds__(kZencThreads) zstd_encode_kernel(ZstdEncParams p)
{
    VBZ_DYN_SMEM(smem);
    ZEncSmem* s = reinterpret_cast<ZEncSmem*>(smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t b = blockIdx.x;
    const uint32_t r = p.block_read[b];
    const ZstdEncRead rd = p.reads[r];
    const uint32_t bi = b - rd.first_block;
    if (rd.rejected) {
        if (bi == 0 && tid == 0) p.results[r] = rd.preset;
        return;
    }
    const uint32_t content = p.src_sizes ? p.src_sizes[r] : rd.src_bytes;
    if (p.src_sizes && content >= kErrOutOfMemory) {  // stage 1 reported an error for this read
        if (bi == 0 && tid == 0) p.results[r] = content;
        return;
    }
    // blocks of the first region (stage 1's control bytes: their statistics differ completely from the data bytes'), then the rest
    const uint32_t split = rd.split < content ? rd.split : 0u;
    const uint32_t n_a = (split + kZEncBlock - 1) / kZEncBlock, n_b = (content - split + kZEncBlock - 1) / kZEncBlock;
    const uint32_t nblk = n_a + n_b ? n_a + n_b : 1u;
    if (bi >= nblk) return;  // the table was sized for the largest content stage 1 could have produced
    const uint32_t start = bi < n_a ? bi * kZEncBlock : split + (bi - n_a) * kZEncBlock;
    const uint32_t region_end = bi < n_a ? split : content;
    const uint32_t B = region_end - start < kZEncBlock ? region_end - start : kZEncBlock;
    const bool last = bi + 1 == nblk;
    const uint8_t* src = p.src + rd.src_off + start;

    // ---- zero the staging area and the histograms; headers that do not depend on the payload ---------------------------
    for (uint32_t i = tid; i < (kZEncBlock + 64) / 16; i += kZencThreads) reinterpret_cast<uint4*>(s->stage)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < kZencWarps * 256; i += kZencThreads) (&s->hist[0][0])[i] = 0;
    __syncthreads();
    uint32_t pre = 0;  // bytes before the block header
    if (bi == 0) {
        if (p.hdr_bytes) pre = 4;
        const uint32_t fcs_bytes = content < 256 ? 1u : (content < 65536 + 256 ? 2u : 4u);
        if (tid == 0) {
            uint8_t* h = s->stage;
            if (p.hdr_bytes) { h[0] = (uint8_t)rd.hdr_value; h[1] = (uint8_t)(rd.hdr_value >> 8); h[2] = (uint8_t)(rd.hdr_value >> 16); h[3] = (uint8_t)(rd.hdr_value >> 24); }
            h += pre;
            h[0] = 0x28; h[1] = 0xB5; h[2] = 0x2F; h[3] = 0xFD;
            h[4] = (uint8_t)(0x20 | (fcs_bytes == 1 ? 0x00 : fcs_bytes == 2 ? 0x40 : 0x80));  // single segment, content size present
            const uint32_t f = fcs_bytes == 2 ? content - 256 : content;
            for (uint32_t k = 0; k < fcs_bytes; ++k) h[5 + k] = (uint8_t)(f >> (8 * k));
        }
        pre += 5 + fcs_bytes;
    }
    uint8_t* blk = s->stage + pre;  // block header, then the payload

    // ---- histogram ----------------------------------------------------------------------------------------------------
    for (uint32_t i = tid; i < B; i += kZencThreads) atomicAdd(&s->hist[wid][src[i]], 1u);
    __syncthreads();
    for (uint32_t i = tid; i < 256; i += kZencThreads) s->cnt[i] = s->hist[0][i] + s->hist[1][i] + s->hist[2][i] + s->hist[3][i];
    __syncthreads();
    // rank sort of the symbols that occur, ascending by (count, symbol)
    uint32_t nsym = 0;
    for (uint32_t sym = tid; sym < 256; sym += kZencThreads) {
        const uint32_t c = s->cnt[sym];
        if (c) {
            uint32_t rank = 0;
            for (uint32_t j = 0; j < 256; ++j) {
                const uint32_t cj = s->cnt[j];
                rank += (cj != 0 && (cj < c || (cj == c && j < sym))) ? 1u : 0u;
            }
            s->order[rank] = (uint8_t)sym;
            nsym++;
        }
    }
    nsym = __reduce_add_sync(0xffffffffu, nsym);
    if (lane == 0) s->v[12 + wid] = nsym;
    __syncthreads();
    nsym = s->v[12] + s->v[13] + s->v[14] + s->v[15];
    __syncthreads();

    // ---- the code and the kind of block (one thread) -----------------------------------------------------------------------
    if (tid == 0) {
        uint32_t kind = kZKindRaw, tree_bytes = 0;
        if (B > 0 && nsym == 1) kind = kZKindRle;
        else if (nsym >= 2 && B >= 32 && zenc_build_code(s, (int)nsym)) {
            tree_bytes = zenc_describe_tree(s);
            if (tree_bytes) kind = kZKindHuf;
        }
        s->v[kZvKind] = kind;
        s->v[kZvTreeBytes] = tree_bytes;
        s->v[kZvNsym] = nsym;
    }
    __syncthreads();
    uint32_t kind = s->v[kZvKind];
    const uint32_t tree_bytes = s->v[kZvTreeBytes];

    // ---- stream sizes: a warp per stream ---------------------------------------------------------------------------------------
    const uint32_t streams = B < 1024 ? 1u : 4u;
    const uint32_t seg = (B + 3) / 4;
    uint32_t s_off = 0, s_cnt = 0, chunk = 0, my_bits = 0, lane_bit0 = 0;
    if (kind == kZKindHuf) {
        if (streams == 1) { s_off = 0; s_cnt = wid == 0 ? B : 0; }
        else { s_off = wid * seg; s_cnt = wid == 3 ? B - 3 * seg : seg; }
        my_bits = zenc_stream_bits(s->code, src + s_off, s_cnt, lane, &chunk);
        const uint32_t incl = warp_scan_u32(my_bits);
        lane_bit0 = incl - my_bits;
        if (lane == 31) s->v[kZvBits0 + wid] = incl;
    }
    __syncthreads();
    uint32_t sbytes[4] = {0, 0, 0, 0}, comp = 0, lit_hdr = 0;
    if (kind == kZKindHuf) {
        for (uint32_t k = 0; k < streams; ++k) sbytes[k] = (s->v[kZvBits0 + k] >> 3) + 1;  // + the marker bit
        comp = tree_bytes + (streams == 4 ? 6 : 0) + sbytes[0] + sbytes[1] + sbytes[2] + sbytes[3];
        lit_hdr = (streams == 1 || (B < 1024 && comp < 1024)) ? 3u : ((B < 16384 && comp < 16384) ? 4u : 5u);
        if (streams == 1 && comp >= 1024) kind = kZKindRaw;
        if (lit_hdr + comp + 1 >= B) kind = kZKindRaw;  // not smaller than the bytes themselves
    }

    // ---- payload ----------------------------------------------------------------------------------------------------------------------
    const uint32_t payload = kind == kZKindHuf ? lit_hdr + comp + 1 : (kind == kZKindRle ? 1u : B);  // Huffman: literals section + "0 sequences"
    if (tid == 0) {
        const uint32_t bsize = kind == kZKindRle ? B : payload;
        const uint32_t bh = (last ? 1u : 0u) | ((kind == kZKindHuf ? 2u : kind == kZKindRle ? 1u : 0u) << 1) | (bsize << 3);
        blk[0] = (uint8_t)bh; blk[1] = (uint8_t)(bh >> 8); blk[2] = (uint8_t)(bh >> 16);
    }
    if (kind == kZKindHuf) {
        uint8_t* lit = blk + 3;
        if (tid == 0) {
            const uint32_t sf = lit_hdr == 3 ? (streams == 1 ? 0u : 1u) : (lit_hdr == 4 ? 2u : 3u);
            uint64_t v = 2u | (sf << 2);
            if (lit_hdr == 3) v |= ((uint64_t)B << 4) | ((uint64_t)comp << 14);
            else if (lit_hdr == 4) v |= ((uint64_t)B << 4) | ((uint64_t)comp << 18);
            else v |= ((uint64_t)B << 4) | ((uint64_t)comp << 22);
            for (uint32_t k = 0; k < lit_hdr; ++k) lit[k] = (uint8_t)(v >> (8 * k));
        }
        for (uint32_t k = tid; k < tree_bytes; k += kZencThreads) lit[lit_hdr + k] = s->tree[k];
        uint8_t* body = lit + lit_hdr + tree_bytes;
        if (streams == 4) {
            if (tid < 3) { body[2 * tid] = (uint8_t)sbytes[tid]; body[2 * tid + 1] = (uint8_t)(sbytes[tid] >> 8); }
            body += 6;
        }
        __syncthreads();  // byte-wise header writes land before the word-wise stream writes next to them
        // each stream starts at a byte position; the bit writer works on aligned words, so shift by the byte phase
        if (wid < streams) {
            uint32_t so = 0;
            for (uint32_t k = 0; k < wid; ++k) so += sbytes[k];
            uint8_t* sp = body + so;
            const uint32_t phase = (uint32_t)((uintptr_t)sp & 3u);
            zenc_stream_write(s->code, src + s_off, s_cnt, lane, chunk, lane_bit0 + 8 * phase, sp - phase);
            if (lane == 0) {  // the marker bit above the last bit
                const uint32_t mb = s->v[kZvBits0 + wid] + 8 * phase;
                atomicOr(reinterpret_cast<uint32_t*>(sp - phase) + (mb >> 5), 1u << (mb & 31u));
            }
        }
    } else if (kind == kZKindRle) {
        if (tid == 0) blk[3] = src[0];
    } else {
        for (uint32_t i = tid; i < B; i += kZencThreads) blk[3 + i] = src[i];
    }
    const uint32_t total = pre + 3 + payload;

    // ---- position inside the frame (look-back over the frame's earlier blocks), copy out ----------------------------------------------------
    if (wid == 0) {
        uint64_t excl = 0;
        if (bi == 0) {
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
    if (end <= rd.dst_cap) {
        uint8_t* dst = p.dst + rd.dst_off + excl;
        for (uint32_t i = tid; i < total; i += kZencThreads) dst[i] = s->stage[i];
    }
    if (last && tid == 0) p.results[r] = end <= rd.dst_cap ? (uint32_t)end : kErrZstd + p.err_add;  // ZSTD_compress: dstSize_tooSmall (vbz.cpp:201-204)
}

}  // namespace vbz
