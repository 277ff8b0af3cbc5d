// csrc/zstd_match_kernels.cuh - the sequence half of the zstd frame writer: match finding, and the sequences sections (RFC 8878 3.1.1.3.2).
//
// Replaces what ZSTD_compress(level) does beyond entropy-coding literals (ref call site vbz/vbz.cpp:194-204): block splitting,
// hash-table match finding, FSE table construction for literal lengths / match lengths / offsets and the interleaved sequences
// bitstream.  On nanopore signal the matches that matter are the runs of zero control bytes at the head of every stage-1 stream
// (96.9 % of them are 0x00) - Huffman cannot go below one bit per byte there; on repetitive input (the
// reference's own benchmark tiles one trace, vbz/perf/test_data_generator.h:29-74; numpy.arange in pyvbz's README) matches are everything.
//
// Units of work.  The content of a frame is cut into GROUPS of <= 128 KiB (csrc/zstd_enc_kernels.cuh).  A group is written either
//   L  as up to 16 literal-only blocks of 8 KiB (no sequences: what the block-parallel decoder likes best), or
//   S  as ceil(literals / 8 KiB) blocks that each hold <= 8 KiB of the group's LITERALS (the bytes no match covers, compacted) plus
//      the sequences that end inside that stretch of literals - so the literal machinery of the encode kernel (4 Huffman streams per
//      block, lanes own runs of <= 64 symbols) serves both kinds unchanged.
// zstd_enc_match_kernel   a WARP per group: greedy parse, 32 positions per step.  Every lane looks at "its" position: a repeat of the
//                         previous byte (4 bytes equal to the 4 bytes one back: a run) or a candidate from a hash table of 4-byte
//                         6-byte strings (2^10 tagged entries per warp in shared memory, seeded with the 8 KiB before the group), verified against
//                         the data.  The first lane with a match wins, the warp extends it 128 bytes per step, the parse continues
//                         behind it.  The longer the stretch without a match, the more steps are skipped between examined ones (level 1: up
//                         to 3 in 4; higher levels later and less), as libzstd's fast strategy accelerates over incompressible data.  Sequences go
//                         to a record array as {offset, literal run, match length}; a group becomes S only if its matches pay for
//                         their sequences (12 bytes of match per sequence - which also bounds every scratch area), and only then are
//                         its literals compacted and counted (per-group histogram for the Huffman code).
// zstd_enc_seq_kernel     a warp per S group: sequence codes and their histograms, FSE tables (normalised counts, the table description,
//                         the encoder's state tables; RLE mode for a lone code, the predefined tables for a handful of sequences),
//                         then lane 0 writes the sequences section of every block of the group: the first one carries the tables,
//                         the others repeat them (Repeat_Mode).  Offsets are always explicit (offset + 3): no repeat-offset history
//                         has to be tracked, blocks stay independent of what precedes them.
// The sections land behind the group's literals in the literal scratch; zstd_encode_kernel copies them in behind each literals section.
#pragma once
#include "zstd_dec_kernels.cuh"
#include "svb_kernels.cuh"

namespace vbz {

constexpr uint32_t kZCodeSet = 8;  // consecutive groups of a region that share one Huffman code (built once, over up to 1 MiB)

// where group b of the batch lives: false = nothing to encode (rejected read, upstream error, or a group beyond the content)
struct ZGroupLoc {
    const uint8_t* src;
    uint32_t r, gi, ngrp, content, B;
    uint32_t code_first;  // groups back to the first group of this one's CODE SET (kZCodeSet consecutive groups of a region share one code)
    uint32_t code_bytes;  // content bytes of the code set, from its first group on
    uint32_t region_off;  // content bytes of the group's region (control bytes | data bytes) before the group
    bool keys;            // the group lies in the control-byte region of a stage-1 stream
    uint32_t gsize;       // nominal size of the groups of this region
    bool last;
};
__device__ __forceinline__ bool zenc_locate(const ZstdEncParams& p, uint32_t b, const ZstdEncRead& rd, uint32_t r, ZGroupLoc& g)
{
    g.r = r;
    g.gi = b - rd.first_block;
    if (rd.rejected) return false;
    g.content = p.src_sizes ? p.src_sizes[r] : rd.src_bytes;
    if (p.src_sizes && g.content >= kErrOutOfMemory) return false;  // stage 1 reported an error for this read
    // groups of the first region (stage 1's control bytes: their statistics differ completely from the data bytes'; groups of 32 KiB there:
    // nearly all of them is matches, so a group is sequences - serial work for one warp in the match finder, the sequence writer and
    // the decoder's sequence reader alike), then the rest in groups of 128 KiB
    const uint32_t split = rd.split < g.content ? rd.split : 0u;
    const uint32_t n_a = (split + kZEncBlockKeys - 1) / kZEncBlockKeys, n_b = (g.content - split + kZEncBlock - 1) / kZEncBlock;
    g.ngrp = n_a + n_b ? n_a + n_b : 1u;
    if (g.gi >= g.ngrp) return false;  // the table was sized for the largest content stage 1 could have produced
    g.keys = g.gi < n_a;
    g.gsize = g.keys ? kZEncBlockKeys : kZEncBlock;
    const uint32_t start = g.keys ? g.gi * kZEncBlockKeys : split + (g.gi - n_a) * kZEncBlock;
    const uint32_t region_end = g.keys ? split : g.content;
    g.B = region_end - start < g.gsize ? region_end - start : g.gsize;
    g.last = g.gi + 1 == g.ngrp;
    g.src = p.src + rd.src_off + start;
    const uint32_t in_region = g.keys ? g.gi : g.gi - n_a;
    g.code_first = in_region % kZCodeSet;
    g.region_off = in_region * g.gsize;
    const uint32_t set_start = start - g.code_first * g.gsize;
    g.code_bytes = region_end - set_start < kZCodeSet * g.gsize ? region_end - set_start : kZCodeSet * g.gsize;
    return true;
}



constexpr uint32_t kZHashLog = 10;
constexpr uint32_t kZTagBits = 14;          // a table entry = (window coordinate + 1) << 14 | 14 more bits of the hash: a lane only goes to memory to verify a candidate whose tag agrees
constexpr uint32_t kZPreWindow = 65536;      // bytes before the group that seed the hash table (a period of the reference benchmark's tiled trace is 16 KB of data bytes)
constexpr uint32_t kZPreDense = 4096;        // ... the nearest of them one position in four, the rest one in sixteen
constexpr uint32_t kZMinRun = 12;            // shortest run (offset 1) worth a sequence
constexpr uint32_t kZMinMatch = 6;           // shortest match at any other offset
constexpr uint32_t kZMinKeyMatch = 32;       // ... among the control bytes of a stage-1 stream
constexpr int kZMatchWarps = 1;  // (a warp per CTA: a group of control bytes is dense steps from end to end, a group of noise a few dozen strided ones -
                                 // sharing a CTA, the short ones would hold their slots until the long one is done)
#ifdef VBZ_CPUSIM
static inline uint32_t zm_tune(const char* name, uint32_t dflt) { const char* e = getenv(name); return e ? (uint32_t)atoi(e) : dflt; }  // experiments on the emulator
#else
#define zm_tune(name, dflt) (dflt)
#endif

enum : uint32_t { kZModeL = 0, kZModeS = 1 };
// what the match kernel (and then the sequence kernel) leaves per group
struct ZGroupMeta {
    uint32_t mode;          // kZModeL / kZModeS
    uint32_t nseq;          // S: sequences of the group
    uint32_t lits;          // S: literals of the group (compacted in the literal scratch, at the group's content offset)
    uint32_t nblk;          // S: blocks = max(1, ceil(lits / slice)), slice = zm_slice(control bytes?)
    uint32_t seq_off[16];   // S: the sequences section of block q = seq_len[q] bytes at seq_off[q] of the group's literal slot (0 bytes: no sequences)
    uint32_t seq_len[16];
};
static_assert(sizeof(ZGroupMeta) == kZGroupMetaBytes, "ZGroupMeta layout");
// literals per block of an S group: 2 KiB among control bytes (a group there is sequences more than anything else, and a block's sequences are
// serial work for one lane in zstd_enc_seq_kernel and in the decoder alike: 16 blocks = 16 lanes), 8 KiB among data bytes
__device__ __forceinline__ uint32_t zm_slice_shift(bool keys) { return keys ? 12u : 13u; }

__device__ __forceinline__ uint64_t zm_pack(uint32_t off, uint32_t ll, uint32_t ml) { return (uint64_t)off | ((uint64_t)ll << 18) | ((uint64_t)ml << 36); }
__device__ __forceinline__ uint32_t zm_off(uint64_t r) { return (uint32_t)r & 0x3FFFFu; }
__device__ __forceinline__ uint32_t zm_ll(uint64_t r) { return (uint32_t)(r >> 18) & 0x3FFFFu; }
__device__ __forceinline__ uint32_t zm_ml(uint64_t r) { return (uint32_t)(r >> 36) & 0x3FFFFu; }

// 4 bytes at any address, little-endian, through aligned 32-bit loads
__device__ __forceinline__ uint32_t zm_load32(const uint8_t* p)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>((uintptr_t)p & ~(uintptr_t)3);
    const uint32_t sh = 8u * (uint32_t)((uintptr_t)p & 3u);
    const uint32_t a = w[0];
    return sh ? __funnelshift_r(a, w[1], sh) : a;
}
// hash of 6 bytes (w = the first four, w2 = the next two): what libzstd's fast strategy hashes at level 1 for inputs of this size -
// four bytes repeat by chance all the time in a narrow alphabet, and the greedy parse would take those accidents
// -> index in the top kZHashLog bits, tag in the kZTagBits below them
__device__ __forceinline__ uint32_t zm_hash(uint32_t w, uint32_t w2)
{
    const uint64_t k = ((uint64_t)w | ((uint64_t)(w2 & 0xFFFFu) << 32)) << 16;
    return (uint32_t)((k * 227718039650203ull) >> (64 - kZHashLog - kZTagBits));
}
__device__ __forceinline__ uint32_t zm_index(uint32_t h) { return h >> kZTagBits; }
__device__ __forceinline__ uint32_t zm_entry(uint32_t h, uint32_t wpos) { return ((wpos + 1) << kZTagBits) | (h & ((1u << kZTagBits) - 1u)); }

// first record of group b in the record scratch, and how many it may hold
__device__ __forceinline__ uint64_t zm_rec_base(uint64_t content_off, uint32_t b) { return content_off / 16 + b; }
__device__ __forceinline__ uint32_t zm_rec_cap(uint32_t B) { return B / 16; }

struct ZMatchSmem { uint32_t table[1u << kZHashLog]; };  // (its first KiB is the literal histogram once the parse is over)

// 8 bytes at any address, little-endian, through three aligned 32-bit loads (bytes up to p + 11 - (p & 3) are touched)
__device__ __forceinline__ void zm_load64(const uint8_t* p, uint32_t& lo, uint32_t& hi)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>((uintptr_t)p & ~(uintptr_t)3);
    const uint32_t sh = 8u * (uint32_t)((uintptr_t)p & 3u);
    const uint32_t x0 = w[0], x1 = w[1], x2 = w[2];
    lo = __funnelshift_r(x0, x1, sh);
    hi = __funnelshift_r(x1, x2, sh);
}
__device__ __forceinline__ uint32_t zm_rot8(uint32_t w) { return (w >> 8) | (w << 24); }
// bytes of w that are not zero
__device__ __forceinline__ uint32_t zm_nonzero_bytes(uint32_t w)
{
    const uint32_t t = (w | (w >> 4));
    const uint32_t u = (t | (t >> 2));
    return (uint32_t)__popc((u | (u >> 1)) & 0x01010101u);
}

// what the parse carries from step to step (the same in every lane)
struct ZParse {
    uint32_t pos, anchor, nseq, matched, rep_off;
};

// one sequence: `ml` bytes at `mp` repeat what lies `off` back.  With `back`, the match first grows towards the anchor as far as the bytes
// before it agree too (a strided step lands in the middle of a repeat; libzstd's fast strategy "catches up" the same way).  The
// credit rule: every sequence must be paid for by 12 bytes of match - which also bounds the record and section scratch areas.
__device__ __forceinline__ bool zm_emit(ZParse& z, uint64_t* recs, uint32_t rec_cap, const uint8_t* src, uint32_t pre, uint32_t mp, uint32_t off,
                                        uint32_t ml, bool back, uint32_t need, uint32_t lane)
{
    if (back) {
        for (;;) {
            const uint32_t room = mp - z.anchor;                 // bytes between the anchor and the match
            uint32_t lim = mp + pre - off;                       // ... and how far back the source may go (window coordinate 0)
            if (lim > room) lim = room;
            if (lim == 0) break;
            const uint8_t* win = src - pre;
            const bool eq = lane < lim && win[mp + pre - 1 - lane] == win[mp + pre - 1 - lane - off];
            const uint32_t ne = ~__ballot_sync(0xffffffffu, eq);
            const uint32_t k = ne ? (uint32_t)__ffs((int)ne) - 1 : 32u;
            mp -= k;
            ml += k;
            if (k < 32) break;
        }
    }
    if (ml < need || z.nseq >= rec_cap || !(ml >= 12 || 12ull * (z.nseq + 1) <= (uint64_t)z.matched + ml)) return false;
    if (lane == 0) recs[z.nseq] = zm_pack(off, mp - z.anchor, ml);
    ++z.nseq;
    z.matched += ml;
    z.anchor = mp + ml;
    return true;
}

// how far the bytes from mp + m on repeat what lies `o` back: 4 bytes per lane, 128 per round (m bytes are known to agree)
__device__ __forceinline__ uint32_t zm_extend(const uint8_t* src, uint32_t B, uint32_t mp, uint32_t o, uint32_t m, uint32_t lane)
{
    for (;;) {
        const uint32_t at = mp + m + 4 * lane;
        uint32_t eq = 0;
        if (at < B) {
            const uint32_t x = zm_load32(src + at) ^ zm_load32(src + at - o);
            eq = x ? ((uint32_t)__ffs((int)x) - 1) >> 3 : 4u;
            if (eq > B - at) eq = B - at;
        }
        const uint32_t stop = __ballot_sync(0xffffffffu, eq < 4);
        if (stop) {
            const uint32_t fl = (uint32_t)__ffs((int)stop) - 1;
            return m + 4 * fl + __shfl_sync(0xffffffffu, eq, fl);
        }
        m += 128;
    }
}

// ---- runs among control bytes: 128 bytes per step ---------------------------------------------------------------------------------------
// Nearly every sequence of a signal read is a run of zero control bytes, and there is one for every few dozen bytes: the parse of a
// control-byte group is therefore a scan for runs that never walks over breaks or runs one at a time.
//   zm_runs_line   one aligned 128-byte line: a lane holds 4 bytes and the 4 bits "equals its predecessor"; a run of kZMinRun (>= 11) bytes
//                  covers two whole lanes at least, so runs are stretches of FULL lanes (one ballot) plus the ones at the top of the lane
//                  before and at the bottom of the lane after.  The lane at the head of a stretch owns the run; owners of finished runs
//                  write their records side by side (rank by popc).  The run that reaches the end of the line is carried into the next.
//   zm_runs_bytes  any other range, 32 positions per step (byte loads): heads and tails of spans, and around a match found through the table.
struct ZRuns { uint32_t carry; };  // bytes [pos - carry, pos) each equal their predecessor: the run that is open at pos

__device__ __forceinline__ void zm_runs_line(ZParse& z, ZRuns& rs, uint64_t* recs, uint32_t rec_cap, const uint8_t* src, uint32_t pre, uint32_t P,
                                             uint32_t min_run, uint32_t lane, uint32_t W, uint32_t before)
{
    // W: the lane's four bytes; before: the byte ahead of the line (0x100: none)
    uint32_t pb = __shfl_up_sync(0xffffffffu, W, 1) >> 24;
    if (lane == 0) pb = before < 0x100u ? before : (~W & 0xFFu);
    const uint32_t x = W ^ ((W << 8) | pb);
    const uint32_t t = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);  // 0x80 where a byte of x is zero
    const uint32_t n = (((t >> 7) * 0x00204081u) >> 21) & 0xFu;                  // bit k: byte 4 * lane + k equals its predecessor
    const bool full = n == 0xFu;
    const uint32_t F = __ballot_sync(0xffffffffu, full);
    const uint32_t low1 = (uint32_t)__ffs((int)~n) - 1;          // ones at the bottom of the lane (0 .. 4)
    const uint32_t top1 = (uint32_t)__clz((int)~(n << 28));      // ones at its top (0 .. 4)
    // owners: the first lane of a stretch of full lanes; lane 0 also when the carried run ends inside it
    const uint32_t below = __shfl_up_sync(0xffffffffu, top1, 1);
    const bool first = full && (lane == 0 || !((F >> (lane - 1)) & 1u));
    const uint32_t notf = ~(F >> lane);                                         // (zeros come in at the top: the stretch ends at lane 32 at the latest)
    const uint32_t nfull = !full ? 0u : (notf ? (uint32_t)__ffs((int)notf) - 1 : 32u);  // (a line of nothing but full lanes: 32)
    const uint32_t after = lane + nfull;                                           // the lane behind the stretch
    const uint32_t tail = __shfl_sync(0xffffffffu, low1, after & 31u);
    const uint32_t head = lane == 0 ? rs.carry : below;
    const bool owner = first || (lane == 0 && rs.carry != 0);
    const bool closed = after < 32;
    const uint32_t len = head + 4 * nfull + (closed ? (full ? tail : low1) : 0u);
    const uint32_t start = P + 4 * lane - head;
    const uint32_t Q = __ballot_sync(0xffffffffu, owner && closed && len >= min_run);
    if (Q && z.nseq + (uint32_t)__popc((int)Q) <= rec_cap) {
        const uint32_t lower = Q & ((1u << lane) - 1u);
        const uint32_t prev_lane = lower ? 31u - (uint32_t)__clz((int)lower) : 0u;
        const uint32_t prev_end = __shfl_sync(0xffffffffu, start + len, prev_lane);
        const bool mine = (Q >> lane) & 1u;
        if (mine) recs[z.nseq + (uint32_t)__popc((int)lower)] = zm_pack(1u, start - (lower ? prev_end : z.anchor), len);
        z.nseq += (uint32_t)__popc((int)Q);
        z.matched += __reduce_add_sync(0xffffffffu, mine ? len : 0u);
        z.anchor = __shfl_sync(0xffffffffu, start + len, 31u - (uint32_t)__clz((int)Q));
    }
    // what is open at the end of the line: the stretch that reaches lane 31, or the ones at the top of lane 31
    const uint32_t top_stretch = (uint32_t)__clz((int)~F);  // full lanes at the top
    const uint32_t own = 32u - top_stretch;                 // its first lane (32: none)
    const uint32_t open_len = __shfl_sync(0xffffffffu, head + 4 * nfull, own & 31u);
    const uint32_t top31 = __shfl_sync(0xffffffffu, top1, 31);
    rs.carry = top_stretch ? open_len : top31;
}

// [from, to): 32 positions per step.  `ends_run`: the byte at `to` is known to differ from its predecessor (the open run stops there).
__device__ __forceinline__ void zm_runs_bytes(ZParse& z, ZRuns& rs, uint64_t* recs, uint32_t rec_cap, const uint8_t* src, uint32_t pre, uint32_t from,
                                              uint32_t to, uint32_t min_run, uint32_t lane)
{
    while (from < to) {
        const uint32_t nlive = to - from < 32 ? to - from : 32u;
        const uint32_t live_mask = nlive < 32 ? (1u << nlive) - 1u : 0xFFFFFFFFu;
        const uint32_t q = from + lane;
        const bool same = lane < nlive && (q + pre) != 0 && src[q] == src[(int)q - 1];
        const uint32_t eq1 = __ballot_sync(0xffffffffu, same) & live_mask;
        const uint32_t carry = rs.carry;
        const uint32_t lowc = carry >= 32 ? 0xFFFFFFFFu : (carry ? ~(0xFFFFFFFFu >> carry) : 0u);
        const uint64_t X = ((uint64_t)eq1 << 32) | lowc;
        uint64_t longs = X;
        {
            uint32_t len = 1;
            while (2 * len <= min_run) { longs &= longs >> len; len *= 2; }
            if (min_run > len) longs &= longs >> (min_run - len);
        }
        const uint64_t ends = X & ~(X >> 1) & ~(1ull << (31 + nlive));
        uint32_t run_ends = (uint32_t)((ends & (longs << (min_run - 1))) >> 31);
        while (run_ends) {
            const uint32_t i = (uint32_t)__ffs((int)run_ends) - 1;
            run_ends &= run_ends - 1;
            const uint64_t zeros = ~X & ((1ull << (31 + i)) - 1ull);
            const uint32_t start = zeros ? from + (63u - (uint32_t)__clzll((long long)zeros)) + 1u - 32u : from - carry;
            zm_emit(z, recs, rec_cap, src, pre, start, 1u, from + i - start, false, min_run, lane);
        }
        const uint32_t brk = ~eq1 & live_mask;
        rs.carry = brk ? nlive - 1u - (31u - (uint32_t)__clz((int)brk)) : carry + nlive;
        from += nlive;
    }
}

// [from, to) by whatever fits: bytes up to the first line boundary, whole lines, bytes
__device__ __forceinline__ void zm_runs(ZParse& z, ZRuns& rs, uint64_t* recs, uint32_t rec_cap, const uint8_t* src, uint32_t pre, uint32_t from, uint32_t to,
                                        uint32_t min_run, uint32_t lane)
{
    const uint32_t mis = (uint32_t)((uintptr_t)(src + from) & 127u);
    if (mis && to - from >= 128 + (128 - mis)) {
        zm_runs_bytes(z, rs, recs, rec_cap, src, pre, from, from + (128 - mis), min_run, lane);
        from += 128 - mis;
    }
    if (((uintptr_t)(src + from) & 127u) == 0 && to - from >= 128 && !zm_tune("VBZ_TUNE_NOLINES", 0u)) {
        // the next line is on its way while this one is looked at
        uint32_t W = reinterpret_cast<const uint32_t*>(src + from)[lane];
        uint32_t before = (from + pre) ? src[(int)from - 1] : 0x100u;
        for (; to - from >= 128; from += 128) {
            const uint32_t Wn = to - from >= 256 ? reinterpret_cast<const uint32_t*>(src + from + 128)[lane] : 0u;
            zm_runs_line(z, rs, recs, rec_cap, src, pre, from, min_run, lane, W, before);
            before = __shfl_sync(0xffffffffu, W, 31) >> 24;
            W = Wn;
        }
    }
    zm_runs_bytes(z, rs, recs, rec_cap, src, pre, from, to, min_run, lane);
}

// ---- kernel: greedy parse of every group ---------------------------------------------------------------------------------------------
// A step looks at 32 positions, one per lane, `s` bytes apart.  s = 1 + (bytes since the last match >> skip_shift): over data without
// matches the parse accelerates exactly as libzstd's fast strategy does (ZSTD_compressBlock_fast: step = 1 + (ip - anchor >> 8)), so a
// 128 KiB group of noise costs a few dozen steps, and the table learns the positions the steps looked at.
//   dense steps (s == 1)  bit i of `eq1` says "the byte at pos + i equals the one before it"; a RUN is a stretch of ones, and a run of
//                         kZMinRun bytes or more is a match at offset 1 - found with no table, no memory access beyond the step's own
//                         loads, carried from step to step while it lasts.  The other positions (the BREAKS, where a run ends) ask the
//                         hash table; a verified candidate is extended by the whole warp, forwards and backwards.
//   strided steps         any candidate (a table hit, the repeated offset, five equal bytes) sends the parse back to dense steps at that
//                         position; the match then grows backwards over what the stride jumped.
__global__ void __launch_bounds__(kZMatchWarps * 32) zstd_enc_match_kernel(ZstdEncParams p)
{
    VBZ_DYN_SMEM(smem);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ZMatchSmem* sm = reinterpret_cast<ZMatchSmem*>(smem) + wid;
    const uint32_t b = blockIdx.x * kZMatchWarps + wid;
    if (b >= p.n_blocks) return;
    const uint32_t r = p.block_read[b];
    const ZstdEncRead rd = p.reads[r];
    ZGroupLoc g;
    ZGroupMeta* meta = reinterpret_cast<ZGroupMeta*>(p.meta) + b;
    if (!zenc_locate(p, b, rd, r, g)) return;
    if (lane == 0) { meta->mode = kZModeL; meta->nseq = 0; meta->lits = g.B; meta->nblk = 0; }
    const uint32_t B = g.B;
    if (B < 64) return;
    const uint8_t* src = g.src;
    const uint64_t content_off = (uint64_t)(src - p.src);
    uint64_t* recs = p.seqs + zm_rec_base(content_off, b);
    const uint32_t rec_cap = zm_rec_cap(B);
    // how far back a match may reach: the group's region (control bytes | data bytes), at most kZPreWindow before the group
    const uint32_t pre = g.region_off < kZPreWindow ? g.region_off : kZPreWindow;
    const uint8_t* win = src - pre;  // window coordinate w <-> byte win[w]; the group starts at w = pre
    const uint32_t skip_shift = zm_tune("VBZ_TUNE_SKIPSHIFT", p.level >= 3 ? 10u : (p.level == 2 ? 9u : 8u));
    const uint32_t min_run = zm_tune("VBZ_TUNE_MINRUN", kZMinRun);
    // what a match at a distance must bring.  Control bytes are nearly all zero and cost a bit each as literals - and while every
    // sequence of a group has offset 1 the offsets cost nothing (RLE table): only a long match pays there; among data bytes (6 to 7
    // bits each) a short one does
    const uint32_t need = g.keys ? zm_tune("VBZ_TUNE_KEYMATCH", kZMinKeyMatch) : zm_tune("VBZ_TUNE_MINMATCH", kZMinMatch);
    const uint32_t key_nz = zm_tune("VBZ_TUNE_KEYNZ", 2u);
    const uint32_t key_stride = zm_tune("VBZ_TUNE_KEYSTRIDE", 32u);
    const uint32_t end = B - 11;  // positions q < end are looked at: the loads of a position reach up to q + 10

    for (uint32_t i = lane; i < (1u << kZHashLog); i += 32) sm->table[i] = 0;  // entry = (window coordinate + 1, tag), 0 = empty
    __syncwarp();
    // seed with the pre-window: one position in sixteen, one in four over the nearest 4 KiB (a repeat that matters is long: the dense steps
    // at the head of the group look at every phase of it, one hit is enough).  Later positions replace earlier ones.
    {
        const uint32_t far_end = pre > kZPreDense ? (pre - kZPreDense) & ~15u : 0u;
        for (uint32_t w0 = 0; w0 < far_end; w0 += 512) {
            const uint32_t w = w0 + 16 * lane;
            if (w < far_end) {
                uint32_t lo, hi;
                zm_load64(win + w, lo, hi);
                const uint32_t h = zm_hash(lo, hi);
                atomicMax(&sm->table[zm_index(h)], zm_entry(h, w));  // (lanes of a step may share an entry: the latest position stays, whatever the order)
            }
        }
        __syncwarp();
        for (uint32_t w0 = far_end; w0 + 12 <= pre; w0 += 128) {
            const uint32_t w = w0 + 4 * lane;
            if (w + 12 <= pre) {
                uint32_t lo, hi;
                zm_load64(win + w, lo, hi);
                const uint32_t h = zm_hash(lo, hi);
                atomicMax(&sm->table[zm_index(h)], zm_entry(h, w));
            }
        }
        __syncwarp();
    }

    ZParse z = {0, 0, 0, 0, 0};
    if (g.keys) {
        // ---- control bytes: runs by zm_runs over every byte; the table is asked at 32 positions per SPAN, `s` bytes apart, s growing with
        // the distance from the last match found through the table (runs do not count: there is one every few dozen bytes)
        ZRuns rs = {0};
        uint32_t hanchor = 0;
        static_assert(kZMinRun >= 11, "zm_runs_line: a run must cover two whole lanes");
        const uint32_t kmin_run = min_run < 11 ? 11u : min_run;
        while (z.pos < B) {
            uint32_t st = 1 + ((z.pos - hanchor) >> skip_shift);
            if (st > key_stride) st = key_stride;
            uint32_t span_end = z.pos + 32 * st;
            span_end += (128u - (uint32_t)((uintptr_t)(src + span_end) & 127u)) & 127u;  // spans end on line boundaries
            if (span_end > B) span_end = B;
            const uint32_t q = z.pos + lane * st;
            const bool live = q < end && q < span_end;
            uint32_t w = 0, w2 = 0, prevb = 0x100;
            if (live) {
                uint32_t lo, hi;
                if (q + pre != 0) {
                    zm_load64(src + q - 1, lo, hi);
                    prevb = lo & 0xFFu;
                    w = __funnelshift_r(lo, hi, 8);
                    w2 = (hi >> 8) & 0xFFFFu;
                } else {
                    zm_load64(src + q, lo, hi);
                    w = lo;
                    w2 = hi & 0xFFFFu;
                }
            }
            // a position inside a run belongs to the run; a match must be long to pay for its offset (while every sequence of a group
            // has offset 1 the offsets cost nothing: RLE table) - four times as long where it starts with "one code byte and zeros"
            const bool hl = live && (w & 0xFFu) != prevb && w != zm_rot8(w);
            const uint32_t far = zm_nonzero_bytes(w) >= key_nz ? need : 4u * need;
            const uint32_t h = zm_hash(w, w2);
            const uint32_t at = q + pre;
            uint32_t dist = 0;
            if (hl) {
                const uint32_t e = sm->table[zm_index(h)], ep = e >> kZTagBits;
                if (ep != 0 && ((e ^ h) & ((1u << kZTagBits) - 1u)) == 0 && ep - 1 < at) {
                    uint32_t lo, hi;
                    zm_load64(win + (ep - 1), lo, hi);
                    if (lo == w && (hi & 0xFFFFu) == w2) dist = at - (ep - 1);
                }
                if (dist == 0 && z.rep_off > 1 && at >= z.rep_off && zm_load32(src + q - z.rep_off) == w) dist = z.rep_off;
                // ... look at its far end and its middle before the warp spends a round on it
                if (dist != 0 && !(q + far <= B && zm_load32(src + q + far - 4) == zm_load32(src + q + far - 4 - dist) &&
                                   zm_load32(src + q + far / 2) == zm_load32(src + q + far / 2 - dist)))
                    dist = 0;
            }
            uint32_t others = __ballot_sync(0xffffffffu, dist != 0);
            uint32_t cur = z.pos;
            bool moved = false;
            while (others) {
                const uint32_t f = (uint32_t)__ffs((int)others) - 1;
                others &= others - 1;
                const uint32_t mp = z.pos + f * st;
                zm_runs(z, rs, recs, rec_cap, src, pre, cur, mp, kmin_run, lane);
                if (rs.carry >= kmin_run) zm_emit(z, recs, rec_cap, src, pre, mp - rs.carry, 1u, rs.carry, false, kmin_run, lane);  // (the byte at mp differs from its predecessor)
                rs.carry = 0;
                cur = mp;
                const uint32_t o = __shfl_sync(0xffffffffu, dist, f);
                const uint32_t ml = zm_extend(src, B, mp, o, 4, lane);
                if (zm_emit(z, recs, rec_cap, src, pre, mp, o, ml, true, __shfl_sync(0xffffffffu, far, f), lane)) {
                    z.rep_off = o;
                    hanchor = z.anchor;
                    moved = true;
                    break;
                }
            }
            __syncwarp();  // (the step's look-ups are done)
            if (hl && (!moved || q < z.anchor)) atomicMax(&sm->table[zm_index(h)], zm_entry(h, at));
            __syncwarp();
            if (moved) { z.pos = z.anchor; continue; }
            zm_runs(z, rs, recs, rec_cap, src, pre, cur, span_end, kmin_run, lane);
            z.pos = span_end;
        }
        if (rs.carry >= kmin_run) zm_emit(z, recs, rec_cap, src, pre, B - rs.carry, 1u, rs.carry, false, kmin_run, lane);
    }
    uint32_t carry = 0;        // dense steps: the bytes [pos - carry, pos) each equal their predecessor (a run that is still open)
    bool dense_next = false;   // the next step is dense whatever the distance from the anchor
    uint32_t caught_at = 0xFFFFFFFFu;  // where a strided step sent the parse back to dense ones: a run that starts there may have begun earlier
    while (!g.keys && z.pos < end) {
        uint32_t s = 1;
        if (!dense_next && carry == 0) {
            s += (z.pos - z.anchor) >> skip_shift;
            if (s > 255) s = 255;
        }
        dense_next = false;
        const uint32_t q = z.pos + lane * s;
        const bool live = q < end;
        if (s == 1 && lane < 2 && z.pos + 256 + 128 * lane < B) vbz_prefetch_l2(src + z.pos + 256 + 128 * lane);
        uint32_t w = 0, w2 = 0, prevb = 0x100;
        if (live) {
            uint32_t lo, hi;
            if (q + pre != 0) {
                zm_load64(src + q - 1, lo, hi);
                prevb = lo & 0xFFu;
                w = __funnelshift_r(lo, hi, 8);
                w2 = (hi >> 8) & 0xFFFFu;
            } else {  // nothing before the region's first byte
                zm_load64(src + q, lo, hi);
                w = lo;
                w2 = hi & 0xFFFFu;
            }
        }
        const bool all4 = w == zm_rot8(w);
        const bool same = live && (w & 0xFFu) == prevb;
        // lanes that ask the table: not inside a run (the run logic owns those positions)
        const bool hl = live && (s > 1 || (!same && !all4));
        // a match among control bytes must be long to pay for its offset - and four times as long where it starts with "one code byte
        // and zeros", which repeats all the time there
        const uint32_t far = !g.keys ? 0u : (zm_nonzero_bytes(w) >= key_nz ? need : 4u * need);
        const uint32_t h = zm_hash(w, w2);
        const uint32_t at = q + pre;
        uint32_t dist = 0;
        if (hl) {
            const uint32_t e = sm->table[zm_index(h)], ep = e >> kZTagBits;  // ep = window coordinate + 1, 0 = empty
            if (ep != 0 && ((e ^ h) & ((1u << kZTagBits) - 1u)) == 0 && ep - 1 < at) {
                uint32_t lo, hi;
                zm_load64(win + (ep - 1), lo, hi);
                if (lo == w && (hi & 0xFFFFu) == w2) dist = at - (ep - 1);
            }
            // the offset of the last match found through the table, tried again: once a period has been found the parse stays on it
            if (dist == 0 && z.rep_off > 1 && at >= z.rep_off && zm_load32(src + q - z.rep_off) == w) dist = z.rep_off;
            // ... look at its far end before the warp spends a round on it
            if (dist != 0 && far != 0 && !(q + far <= B && zm_load32(src + q + far - 4) == zm_load32(src + q + far - 4 - dist))) dist = 0;
        }
        const uint32_t others = __ballot_sync(0xffffffffu, dist != 0);
        if (s > 1) {
            // ---- strided step: any sign of a repeat -> dense steps from there -------------------------------------------------------------
            const uint32_t cand = others | __ballot_sync(0xffffffffu, same && all4);
            const uint32_t f = cand ? (uint32_t)__ffs((int)cand) - 1 : 32u;
            __syncwarp();  // (the step's look-ups are done)
            if (hl && lane <= f) atomicMax(&sm->table[zm_index(h)], zm_entry(h, at));
            __syncwarp();
            if (cand) { z.pos += f * s; dense_next = true; caught_at = z.pos; }
            else z.pos += 32 * s;
            continue;
        }
        // ---- dense step: the runs that end in it, the candidates at the breaks, in order --------------------------------------------------
        // X = the step's 32 "equals its predecessor" bits in the high word, below them the run that was open when the step began (as many
        // ones as it is long, 32 at most).  Bit j of `longs` says "ones at j .. j + min_run - 1" (erosion by doubling), an END is a one
        // below a zero: a run worth a sequence is an end whose bit min_run - 1 down is set in `longs` - found without walking the breaks.
        const uint32_t nlive = end - z.pos < 32 ? end - z.pos : 32u;
        const uint32_t live_mask = nlive < 32 ? (1u << nlive) - 1u : 0xFFFFFFFFu;
        const uint32_t eq1 = __ballot_sync(0xffffffffu, same) & live_mask;
        const uint32_t lowc = carry >= 32 ? 0xFFFFFFFFu : (carry ? ~(0xFFFFFFFFu >> carry) : 0u);
        const uint64_t X = ((uint64_t)eq1 << 32) | lowc;
        uint64_t longs = X;
        {
            uint32_t len = 1;
            while (2 * len <= min_run) { longs &= longs >> len; len *= 2; }
            if (min_run > len) longs &= longs >> (min_run - len);
        }
        // ends at window index -1 (the open run stops at the step's first position) .. nlive - 2 (a run that reaches the last live position
        // stays open); bit i of run_ends: a run worth a sequence ends just before window index i
        const uint64_t ends = X & ~(X >> 1) & ~(1ull << (31 + nlive));
        const uint32_t run_ends = (uint32_t)((ends & (longs << (min_run - 1))) >> 31);
        uint32_t events = run_ends | others;
        bool moved = false;
        while (events) {
            const uint32_t i = (uint32_t)__ffs((int)events) - 1;
            events &= events - 1;
            if ((run_ends >> i) & 1u) {
                // the run's first position: above the highest zero below its end
                const uint64_t zeros = ~X & ((1ull << (31 + i)) - 1ull);
                const uint32_t start = zeros ? z.pos + (63u - (uint32_t)__clzll((long long)zeros)) + 1u - 32u : z.pos - carry;
                zm_emit(z, recs, rec_cap, src, pre, start, 1u, z.pos + i - start, start == caught_at, min_run, lane);
            }
            if ((others >> i) & 1u) {
                const uint32_t o = __shfl_sync(0xffffffffu, dist, i), mp = z.pos + i;
                const uint32_t ml = zm_extend(src, B, mp, o, 4, lane);
                const uint32_t need_f = g.keys ? __shfl_sync(0xffffffffu, far, i) : need;
                if (zm_emit(z, recs, rec_cap, src, pre, mp, o, ml, true, need_f, lane)) {
                    z.rep_off = o;
                    moved = true;
                    break;
                }
            }
        }
        // the table learns the positions the parse has passed - not those behind a match that ends inside the step, which the next
        // step looks at again (their entries would point at themselves)
        __syncwarp();  // (the step's look-ups are done)
        if (hl && (!moved || q < z.anchor)) atomicMax(&sm->table[zm_index(h)], zm_entry(h, at));
        __syncwarp();
        if (moved) { z.pos = z.anchor; carry = 0; }
        else {
            const uint32_t brk = ~eq1 & live_mask;  // the run that is open at the end of the step: the ones above the last break
            carry = brk ? nlive - 1u - (31u - (uint32_t)__clz((int)brk)) : carry + nlive;
            z.pos += nlive;
        }
    }
    if (!g.keys && carry) {  // the run that is open at the end: to the group's last byte
        const uint32_t start = z.pos - carry;
        const uint32_t ml = zm_extend(src, B, start, 1u, carry, lane);
        if (ml >= min_run) zm_emit(z, recs, rec_cap, src, pre, start, 1u, ml, start == caught_at, min_run, lane);
    }
    const uint32_t nseq = z.nseq, matched = z.matched;
    __syncwarp();
    // ---- L or S? ----------------------------------------------------------------------------------------------------------------------
    const uint32_t floor_bytes = B / 32 > 64 ? B / 32 : 64;
#ifdef VBZ_CPUSIM
    if (lane == 0 && getenv("VBZ_B200_TRACE_MATCH")) fprintf(stderr, "match: group %u B %u nseq %u matched %u rep_off %u\n", b, B, nseq, matched, z.rep_off);
#endif
    // (12 nseq <= matched holds by construction; the sections - at most 10 bytes per sequence and 256 of headers and tables - must fit behind the literals)
    // and with 96 more (16 blocks, 3 + 3 bytes of headers each) an S group is never larger than the bytes it stands for
    if (nseq == 0 || matched < floor_bytes || 10ull * nseq + 256 + 96 + 16 * 8 + 16 > matched) return;  // stays L
    // ---- S: compact the literals into the literal scratch, count them --------------------------------------------------------------------
    uint8_t* lits = p.lits + content_off;
    uint32_t* hist = sm->table;  // (the parse is over)
    for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    uint32_t from = 0, to = 0;  // source position / literal count before the records of this round
    uint32_t zeros = 0;         // literals that are 0x00 (most of them among control bytes: counted by ballot, not by 32 atomics on one word)
    uint64_t rec_next = lane < nseq ? recs[lane] : 0ull;  // (the next round's records are on their way while this round's literals move)
    for (uint32_t i0 = 0; i0 < nseq; i0 += 32) {
        const uint64_t rec = rec_next;
        rec_next = i0 + 32 + lane < nseq ? recs[i0 + 32 + lane] : 0ull;
        const uint32_t ll = zm_ll(rec), ml = zm_ml(rec);
        const uint32_t span = warp_scan_u32(ll + ml), lsum = warp_scan_u32(ll);
        const uint32_t my_from = from + span - (ll + ml);
        const uint32_t total = __shfl_sync(0xffffffffu, lsum, 31);
        // a gather: literal o of the round belongs to the first record whose inclusive count exceeds o (binary search over the lanes);
        // two steps of 32 literals at a time, both loads in flight before the first store
        for (uint32_t o0 = 0; o0 < total; o0 += 64) {
            uint32_t v[2], at[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t o = o0 + 32 * h + lane;
                uint32_t j = 0;
#pragma unroll
                for (uint32_t step = 16; step != 0; step >>= 1) {
                    const uint32_t c = __shfl_sync(0xffffffffu, lsum, j + step - 1);
                    if (c <= o) j += step;
                }
                j &= 31u;
                const uint32_t fj = __shfl_sync(0xffffffffu, my_from, j), ej = __shfl_sync(0xffffffffu, lsum - ll, j);
                at[h] = o;
                v[h] = o < total ? (uint32_t)__ldg(src + fj + (o - ej)) : 0x100u;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (at[h] < total) {
                    lits[to + at[h]] = (uint8_t)v[h];
                    if (v[h]) atomicAdd(&hist[v[h]], 1u);
                }
                zeros += (uint32_t)__popc((int)__ballot_sync(0xffffffffu, v[h] == 0));
            }
        }
        from += __shfl_sync(0xffffffffu, span, 31);
        to += total;
    }
    for (uint32_t k0 = from; k0 < B; k0 += 32) {  // the literals after the last match
        const uint32_t k = k0 + lane;
        uint32_t v = 0x100;
        if (k < B) {
            v = src[k];
            lits[to + k - from] = (uint8_t)v;
            if (v) atomicAdd(&hist[v], 1u);
        }
        zeros += (uint32_t)__popc((int)__ballot_sync(0xffffffffu, v == 0));
    }
    __syncwarp();
    if (lane == 0) hist[0] = zeros;
    to += B - from;
    __syncwarp();
    uint32_t* gh = p.hist + (uint64_t)b * 256;
    for (uint32_t i = lane; i < 256; i += 32) gh[i] = hist[i];
    if (lane == 0) { meta->mode = kZModeS; meta->nseq = nseq; meta->lits = to; meta->nblk = to ? (to + (1u << zm_slice_shift(g.keys)) - 1) >> zm_slice_shift(g.keys) : 1u; }
}

// ---- sequence codes (RFC 8878 3.1.1.3.2.1.1) ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t zm_ll_code(uint32_t ll)
{
    if (ll < 16) return ll;
    if (ll < 64) {
        if (ll < 24) return 16 + ((ll - 16) >> 1);
        if (ll < 32) return 20 + ((ll - 24) >> 2);
        if (ll < 48) return 22 + ((ll - 32) >> 3);
        return 24;
    }
    return (uint32_t)z_highbit(ll) + 19;  // 64 -> 25, 128 -> 26, ...
}
__device__ __forceinline__ uint32_t zm_ml_code(uint32_t ml)  // ml >= 3
{
    const uint32_t v = ml - 3;
    if (v < 32) return v;
    if (v < 128) {
        if (v < 40) return 32 + ((v - 32) >> 1);
        if (v < 48) return 36 + ((v - 40) >> 2);
        if (v < 64) return 38 + ((v - 48) >> 3);
        if (v < 96) return 40 + ((v - 64) >> 4);
        return 42;
    }
    return (uint32_t)z_highbit(v) + 36;  // 128 -> 43, 256 -> 44, ...
}

// one lane's FSE encoder for one of the three symbol types
struct ZFseEnc {
    uint16_t state[256];  // next state by rank of the symbol's occurrences in table order (accuracy log <= 8)
    int32_t dnb[56];      // deltaNbBits
    int16_t dfs[56];      // deltaFindState
    int16_t norm[56];
    uint8_t cell[256];    // symbol of every cell
    uint32_t al, mode, nsym, rle_sym;
};
struct ZSeqEncSmem {
    ZFseEnc t[3];         // 0 = literal lengths, 1 = offsets, 2 = match lengths
    uint32_t hist[3][56];
    uint32_t blk_cnt[16];
    uint8_t desc[3][64];  // table descriptions
    uint32_t desc_n[3];
};

// counts -> normalised counts that sum to 2^al, every present symbol >= 1.  false if impossible.
__device__ bool zm_normalize(const uint32_t* cnt, int nsym, uint32_t total, int al, int16_t* norm)
{
    const int T = 1 << al;
    int sum = 0, big = 0, present = 0;
    for (int s = 0; s < nsym; ++s) {
        int v = 0;
        if (cnt[s]) {
            v = (int)(((uint64_t)cnt[s] * (uint64_t)T) / total);
            if (v == 0) v = 1;
            ++present;
        }
        norm[s] = (int16_t)v;
        sum += v;
        if (cnt[s] > cnt[big]) big = s;
    }
    if (present > T) return false;
    int diff = T - sum;
    if (diff > 0 || norm[big] + diff >= 1) { norm[big] = (int16_t)(norm[big] + diff); return true; }
    while (diff < 0) {  // take the excess from the largest counts, one at a time
        int pick = -1;
        for (int s = 0; s < nsym; ++s)
            if (norm[s] > 1 && (pick < 0 || norm[s] > norm[pick])) pick = s;
        if (pick < 0) return false;
        norm[pick]--;
        diff++;
    }
    return true;
}
// normalised counts -> table description (forward bit stream, RFC 8878 4.1.1).  Returns its size in bytes, 0 if it does not fit `cap`.
__device__ uint32_t zm_describe(const int16_t* norm, int nsym, int al, uint8_t* out, uint32_t cap)
{
    uint64_t acc = 0;
    uint32_t nbits = 0, nbytes = 0;
    auto put = [&](uint32_t v, uint32_t nb) {
        acc |= (uint64_t)v << nbits;
        nbits += nb;
        while (nbits >= 8) { if (nbytes < cap) out[nbytes] = (uint8_t)acc; ++nbytes; acc >>= 8; nbits -= 8; }
    };
    put((uint32_t)(al - 5), 4);
    int remaining = (1 << al) + 1, thr = 1 << al, nb = al + 1, sym = 0;
    while (remaining > 1 && sym < nsym) {
        const int c = norm[sym];
        const int max = (2 * thr - 1) - remaining;
        int v = c + 1;  // (-1, "less than one", would be written as 0: this writer gives every present symbol at least one cell)
        if (v < max) put((uint32_t)v, (uint32_t)(nb - 1));
        else {
            if (v >= thr) v += max;
            put((uint32_t)v, (uint32_t)nb);
        }
        remaining -= c < 0 ? -c : c;
        ++sym;
        if (c == 0) {
            int zeros = 0;
            while (sym + zeros < nsym && norm[sym + zeros] == 0) ++zeros;
            sym += zeros;
            while (zeros >= 3) { put(3, 2); zeros -= 3; }
            put((uint32_t)zeros, 2);
        }
        while (remaining < thr) { nb--; thr >>= 1; }
    }
    if (remaining != 1) return 0;
    if (nbits) { if (nbytes < cap) out[nbytes] = (uint8_t)acc; ++nbytes; }
    return nbytes <= cap ? nbytes : 0;
}
// normalised counts (-1 = "less than one", as in the predefined tables) -> the encoder's tables
__device__ void zm_build_enc(ZFseEnc* t, int nsym, int al)
{
    const int T = 1 << al;
    int cumul[57];
    int high = T - 1;
    cumul[0] = 0;
    for (int s = 0; s < nsym; ++s) {
        const int n = t->norm[s];
        cumul[s + 1] = cumul[s] + (n == -1 ? 1 : n);
        if (n == -1) t->cell[high--] = (uint8_t)s;
    }
    const int step = (T >> 1) + (T >> 3) + 3;
    int pos = 0;
    for (int s = 0; s < nsym; ++s)
        for (int i = 0; i < t->norm[s]; ++i) {
            t->cell[pos] = (uint8_t)s;
            do { pos = (pos + step) & (T - 1); } while (pos > high);
        }
    for (int u = 0; u < T; ++u) { const int s = t->cell[u]; t->state[cumul[s]++] = (uint16_t)(T + u); }
    int total = 0;
    for (int s = 0; s < nsym; ++s) {
        const int n = t->norm[s];
        if (n == 0) { t->dnb[s] = ((al + 1) << 16) - T; t->dfs[s] = 0; }
        else if (n == -1 || n == 1) { t->dnb[s] = (al << 16) - T; t->dfs[s] = (int16_t)(total - 1); total += 1; }
        else {
            const int max_bits_out = al - z_highbit((uint32_t)(n - 1));
            t->dnb[s] = (max_bits_out << 16) - (n << max_bits_out);
            t->dfs[s] = (int16_t)(total - n);
            total += n;
        }
    }
    t->al = (uint32_t)al;
}

// forward bit writer into global memory (one lane): fewer than 32 bits wait in `acc` between calls, four bytes leave at a time
struct ZGBitW { uint8_t* p; uint64_t acc; uint32_t nbits, nbytes; };
__device__ __forceinline__ void zg_put(ZGBitW& w, uint32_t v, uint32_t nb)  // nb <= 32
{
    w.acc |= (uint64_t)v << w.nbits;
    w.nbits += nb;
    if (w.nbits >= 32) {
        const uint32_t x = (uint32_t)w.acc;
        uint8_t* o = w.p + w.nbytes;
        o[0] = (uint8_t)x; o[1] = (uint8_t)(x >> 8); o[2] = (uint8_t)(x >> 16); o[3] = (uint8_t)(x >> 24);
        w.nbytes += 4;
        w.acc >>= 32;
        w.nbits -= 32;
    }
}
// what is left, the last partial byte included
__device__ __forceinline__ void zg_flush(ZGBitW& w)
{
    while (w.nbits > 0) { w.p[w.nbytes++] = (uint8_t)w.acc; w.acc >>= 8; w.nbits = w.nbits > 8 ? w.nbits - 8 : 0; }
}
__device__ __forceinline__ void zm_fse_init(const ZFseEnc* t, uint32_t sym, uint32_t* st)
{
    if (t->mode == 1) { *st = 0; return; }  // RLE: no state
    const int32_t d = t->dnb[sym];
    const uint32_t nb = (uint32_t)(d + (1 << 15)) >> 16;
    const uint32_t value = (nb << 16) - (uint32_t)d;
    *st = t->state[(value >> nb) + t->dfs[sym]];
}
__device__ __forceinline__ void zm_fse_encode(const ZFseEnc* t, uint32_t sym, uint32_t* st, ZGBitW& w)
{
    if (t->mode == 1) return;
    const uint32_t nb = (uint32_t)((int32_t)*st + t->dnb[sym]) >> 16;
    zg_put(w, *st & ((1u << nb) - 1u), nb);
    *st = t->state[(*st >> nb) + t->dfs[sym]];
}
__device__ __forceinline__ void zm_fse_flush(const ZFseEnc* t, uint32_t st, ZGBitW& w)
{
    if (t->mode == 1) return;
    zg_put(w, st & ((1u << t->al) - 1u), t->al);
}

// ---- kernel: the sequences sections of every S group, a warp per group ------------------------------------------------------------------
constexpr int kZSeqEncWarps = 2;
__global__ void __launch_bounds__(kZSeqEncWarps * 32) zstd_enc_seq_kernel(ZstdEncParams p)
{
    VBZ_DYN_SMEM(smem);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ZSeqEncSmem* sm = reinterpret_cast<ZSeqEncSmem*>(smem) + wid;
    const uint32_t b = blockIdx.x * kZSeqEncWarps + wid;
    if (b >= p.n_blocks) return;
    ZGroupMeta* meta = reinterpret_cast<ZGroupMeta*>(p.meta) + b;
    const uint32_t r = p.block_read[b];
    const ZstdEncRead rd = p.reads[r];
    ZGroupLoc g;
    if (!zenc_locate(p, b, rd, r, g) || meta->mode != kZModeS) return;
    const uint32_t nseq = meta->nseq, lits = meta->lits, nblk = meta->nblk;
    const uint64_t content_off = (uint64_t)(g.src - p.src);
    uint64_t* recs = p.seqs + zm_rec_base(content_off, b);
    uint8_t* slot = p.lits + content_off;

    for (uint32_t i = lane; i < 3 * 56; i += 32) (&sm->hist[0][0])[i] = 0;
    if (lane < 16) sm->blk_cnt[lane] = 0;
    __syncwarp();
    const uint32_t shift = zm_slice_shift(g.keys);
    // ---- pass 1: which block a sequence belongs to, its literal run inside that block, the code histograms -----------------------------
    uint32_t lit_before = 0;
    for (uint32_t i0 = 0; i0 < nseq; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint64_t rec = i < nseq ? recs[i] : 0ull;
        const uint32_t ll = zm_ll(rec), ml = zm_ml(rec), off = zm_off(rec);
        const uint32_t lsum = warp_scan_u32(ll);
        if (i < nseq) {
            const uint32_t c = lit_before + lsum - ll;  // literals of the group before this sequence's run
            // the block that holds the last literal in front of the match (none: block 0).  A run that ends on a slice boundary stays with the
            // block it fills: no block starts with a match, so only a genuine long offset makes a block wait for the one before it in the decoder
            uint32_t blk = c + ll ? (c + ll - 1) >> shift : 0u;
            if (blk >= nblk) blk = nblk - 1;
            const uint32_t start = blk << shift;
            const uint32_t ll_in = c + ll - (c > start ? c : start);
            recs[i] = zm_pack(off, ll_in, ml);
            atomicAdd(&sm->hist[0][zm_ll_code(ll_in)], 1u);
            atomicAdd(&sm->hist[1][(uint32_t)z_highbit(off + 3)], 1u);
            atomicAdd(&sm->hist[2][zm_ml_code(ml)], 1u);
            atomicAdd(&sm->blk_cnt[blk], 1u);
        }
        lit_before += __shfl_sync(0xffffffffu, lsum, 31);
    }
    __syncwarp();
    // ---- the three tables: lanes 0, 1, 2 ---------------------------------------------------------------------------------------------------
    if (lane < 3) {
        ZFseEnc* t = &sm->t[lane];
        const int max_sym = lane == 0 ? 36 : (lane == 1 ? 32 : 53);
        const uint32_t* h = sm->hist[lane];
        int nsym = 0, present = 0, only = 0;
        for (int s = 0; s < max_sym; ++s)
            if (h[s]) { nsym = s + 1; ++present; only = s; }
        t->nsym = (uint32_t)nsym;
        sm->desc_n[lane] = 0;
        bool done = false;
        if (present == 1) {  // RLE_Mode: one byte, no bits per sequence
            t->mode = 1; t->rle_sym = (uint32_t)only; t->al = 0;
            sm->desc[lane][0] = (uint8_t)only; sm->desc_n[lane] = 1;
            done = true;
        }
        if (!done && nseq >= 48) {  // FSE_Compressed_Mode
            const int max_al = lane == 1 ? 7 : 8;
            int al = z_highbit(nseq) - 2;
            al = al < 5 ? 5 : (al > max_al ? max_al : al);
            while ((1 << al) < present) ++al;
            if (al <= max_al && zm_normalize(h, nsym, nseq, al, t->norm)) {
                const uint32_t n = zm_describe(t->norm, nsym, al, sm->desc[lane], 64);
                if (n) { sm->desc_n[lane] = n; t->mode = 2; zm_build_enc(t, nsym, al); done = true; }
            }
        }
        if (!done) {  // Predefined_Mode
            const int16_t* def = lane == 0 ? kZ_LL_DEF : (lane == 1 ? kZ_OF_DEF : kZ_ML_DEF);
            const int def_n = lane == 0 ? 36 : (lane == 1 ? 29 : 53);
            for (int s = 0; s < def_n; ++s) t->norm[s] = def[s];
            t->mode = 0;
            zm_build_enc(t, def_n, lane == 1 ? 5 : 6);
        }
    }
    __syncwarp();
    // offsets beyond the predefined table's 29 codes cannot occur: the window is 136 KiB (code <= 17)
    // ---- the sections: a lane per block.  Every block's section has its own place behind the literals, sized for the worst case
    // (10 bytes per sequence: 26 bits of states, 50 of extra bits at most; headers and table descriptions in the first one) --------------
    {
        const ZFseEnc* tl = &sm->t[0];
        const ZFseEnc* to = &sm->t[1];
        const ZFseEnc* tm = &sm->t[2];
        const uint32_t n = lane < nblk ? sm->blk_cnt[lane] : 0u;
        const uint32_t with = __ballot_sync(0xffffffffu, n != 0);
        const bool tables = with && lane == (uint32_t)__ffs((int)with) - 1;  // the first block with sequences carries the tables, the others repeat them
        const uint32_t cap = n ? 8u + 10u * n + (tables ? 200u : 0u) : 0u;
        const uint32_t cap_incl = warp_scan_u32(cap), n_incl = warp_scan_u32(n);
        const uint32_t at = ((lits + 3u) & ~3u) + cap_incl - cap;
        const uint32_t first_rec = n_incl - n;
        uint32_t bytes = 0;
        if (n) {
            uint8_t* o = slot + at;
            uint32_t hb = 0;
            if (n < 128) o[hb++] = (uint8_t)n;
            else if (n < 0x7F00) { o[hb++] = (uint8_t)((n >> 8) + 128); o[hb++] = (uint8_t)n; }
            else { o[hb++] = 255; o[hb++] = (uint8_t)(n - 0x7F00); o[hb++] = (uint8_t)((n - 0x7F00) >> 8); }
            if (tables) {
                o[hb++] = (uint8_t)((tl->mode << 6) | (to->mode << 4) | (tm->mode << 2));
                for (int k = 0; k < 3; ++k)
                    for (uint32_t j = 0; j < sm->desc_n[k]; ++j) o[hb++] = sm->desc[k][j];
            } else {
                o[hb++] = (uint8_t)((3u << 6) | (3u << 4) | (3u << 2));  // Repeat_Mode x 3
            }
            ZGBitW w = {o + hb, 0, 0, 0};
            // written last sequence first: the decoder reads the stream from its end
            uint64_t rec = recs[first_rec + n - 1];
            uint32_t llv = zm_ll(rec), mlv = zm_ml(rec), ofv = zm_off(rec) + 3;
            uint32_t llc = zm_ll_code(llv), mlc = zm_ml_code(mlv), ofc = (uint32_t)z_highbit(ofv);
            uint32_t sl, so, s_m;
            zm_fse_init(tm, mlc, &s_m);
            zm_fse_init(to, ofc, &so);
            zm_fse_init(tl, llc, &sl);
            zg_put(w, llv - kZ_LL_BASE[llc], kZ_LL_BITS[llc]);
            zg_put(w, mlv - kZ_ML_BASE[mlc], kZ_ML_BITS[mlc]);
            zg_put(w, ofv - (1u << ofc), ofc);
            uint64_t rec_ahead = n >= 2 ? recs[first_rec + n - 2] : 0ull;  // (one record ahead of the one being written)
            for (uint32_t k = n - 1; k-- > 0;) {
                rec = rec_ahead;
                if (k) rec_ahead = recs[first_rec + k - 1];
                llv = zm_ll(rec); mlv = zm_ml(rec); ofv = zm_off(rec) + 3;
                llc = zm_ll_code(llv); mlc = zm_ml_code(mlv); ofc = (uint32_t)z_highbit(ofv);
                zm_fse_encode(to, ofc, &so, w);
                zm_fse_encode(tm, mlc, &s_m, w);
                zm_fse_encode(tl, llc, &sl, w);
                zg_put(w, llv - kZ_LL_BASE[llc], kZ_LL_BITS[llc]);
                zg_put(w, mlv - kZ_ML_BASE[mlc], kZ_ML_BITS[mlc]);
                zg_put(w, ofv - (1u << ofc), ofc);
            }
            zm_fse_flush(tm, s_m, w);
            zm_fse_flush(to, so, w);
            zm_fse_flush(tl, sl, w);
            zg_put(w, 1, 1);  // the marker above the last bit
            zg_flush(w);
            bytes = hb + w.nbytes;
        }
        if (lane < 16) { meta->seq_off[lane] = at; meta->seq_len[lane] = bytes; }
    }
}

}  // namespace vbz
