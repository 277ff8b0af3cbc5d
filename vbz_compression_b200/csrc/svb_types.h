// csrc/svb_types.h - descriptors and geometry shared by the host layer and the stage-1 kernels
// (csrc/svb_kernels.cuh).  Plain structs, no device code.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VBZ_HD __host__ __device__ __forceinline__
#else
#define VBZ_HD inline
#endif

namespace vbz {

enum : uint32_t {
    kErrZstd = 0xFFFFFFFFu,
    kErrInputSize = 0xFFFFFFFEu,
    kErrIntegerSize = 0xFFFFFFFDu,
    kErrDestinationSize = 0xFFFFFFFCu,
    kErrStream = 0xFFFFFFFBu,
    kErrVersion = 0xFFFFFFFAu,
    kErrOutOfMemory = 0xFFFFFFF9u,
};

// MODE of the pre-transform / value semantics
enum : int {
    kModeI16X86 = 0,   // int16 + zigzag exactly as the x86 reference build computes it (16-bit delta)
    kModeGeneric = 1,  // widen to 32 bits (sign extension), 32-bit delta + zigzag or none
};

struct SvbRead {
    uint64_t src_off;      // byte offset of this read's input from SvbParams::src (multiple of 16)
    uint64_t dst_off;      // byte offset of this read's output slot from SvbParams::dst (multiple of 16)
    uint32_t src_bytes;    // encode: raw bytes; decode: stage-1 bytes (unless SvbParams::src_sizes is set)
    uint32_t n_values;     // number of integers in the read
    uint32_t first_tile;   // index of the read's first tile
    uint32_t rejected;     // != 0: the read was decided on the host; copy `preset` to results[], do nothing else
    uint32_t preset;
    uint32_t reserved;
};

// one TMA bulk load: global address, byte count (0 = nothing to load), offset inside the shared-memory stage
struct alignas(16) SvbLoad { uint32_t addr_lo, addr_hi, bytes, stage_off; };

struct SvbParams {
    const uint8_t* src;
    uint8_t* dst;
    const SvbRead* reads;
    const uint32_t* tile_read;  // tile -> read index
    uint64_t* status_a;         // look-back chain A (data bytes), one word per tile, zeroed before launch
    uint64_t* status_b;         // look-back chain B (sample value), decode only
    uint32_t* results;          // per read: bytes produced, or a VBZ error code
    const uint32_t* src_sizes;  // decode: optional per-read stage-1 size computed on the device (zstd stage);
                                // a value >= kErrOutOfMemory is an upstream error and is passed through
    uint32_t* wtile_off;        // decode: data-byte offsets from the control-byte pre-pass (8 * n_tiles words).  Generic
                                // kernels: one per warp tile; int16 path: [0, n_tiles) tile offsets, [n_tiles, 2 n_tiles) tile sums
    SvbLoad* issue_a;             // int16 path: per-tile TMA descriptor {address lo, hi, bytes (0 = nothing to load), stage offset}
    SvbLoad* issue_b;             //   encode: samples (issue_a); decode: control bytes (issue_a), data-byte slice (issue_b)
    uint32_t* ticket;           // int16 path: tile ticket counter of the persistent CTAs, zeroed before launch
    uint32_t hdr_bytes;         // encode: 4 = put the "sized" header (ref vbz/vbz.cpp:316-317) before the payload
    uint32_t n_reads;
    uint32_t n_tiles;
    uint32_t debug;             // experiments only (VBZ_B200_DEBUG); 0 in production
    uint32_t ctas_per_sm;       // int16 path: cap on the persistent CTAs per SM (0 = as many as fit).  The host-wave pipeline leaves
                                // a quarter of every SM to the transfer kernels that run next to the codec (csrc/xfer_kernels.cuh)
};

// ---- stage 2 (zstd) ----------------------------------------------------------------------------------------
struct ZstdRead {
    uint64_t src_off;    // the frame(s): bytes [src_off, src_off + src_bytes) of ZstdDecParams::src
    uint64_t dst_off;    // where the content goes, offset into ZstdDecParams::dst
    uint32_t src_bytes;
    uint32_t dst_cap;    // capacity of the destination slot
    uint32_t rejected;   // != 0: decided on the host; copy `preset` to results[]
    uint32_t preset;
};
// One block (or piece of a block) of a frame.  The walker (zstd_walk_kernel) lists them; frames whose blocks carry no sequences
// are executed by zstd_units_kernel alone, in parallel over blocks and Huffman streams.  A block WITH sequences becomes a Seq unit:
// zstd_seq_kernel decodes its sequences into records (and only then knows how many bytes the block regenerates, i.e. where the
// blocks after it go), zstd_seqexec_kernel decodes its literals and executes the records (csrc/zstd_dec_kernels.cuh).
enum : uint32_t { kZUnitNop = 0, kZUnitCopy = 1, kZUnitFill = 2, kZUnitHuf = 3, kZUnitSeq = 4 };
struct ZstdUnit {
    uint64_t src_off;    // Copy: the bytes; Fill: the byte; Huf / Seq: the literals section (its header first), offsets into ZstdDecParams::src
    uint64_t dst_off;    // offset into ZstdDecParams::dst (Seq frames: filled in by zstd_seq_kernel)
    uint32_t n;          // Copy / Fill: bytes to produce; Huf: size of the literals section; Seq: size of the block
    uint32_t tab_back;   // Huf / Seq, treeless literals: distance back from src_off to the literals section that carries the tree
    uint32_t read;       // whose result becomes kErrZstd if the unit fails
    uint32_t kind;
    uint32_t regen;      // content bytes the unit produces (Seq: filled in by zstd_seq_kernel)
    uint32_t nseq;       // Seq: number of sequences
    uint32_t seq_first;  // Seq: index of its first record in ZstdDecParams::seqs
    uint32_t pad;
};
struct ZstdGroup { uint32_t first, count; };  // units [first, first + count): decoded together by one CTA of zstd_units_kernel
struct alignas(16) ZstdSeq { uint32_t offset, ll, ml, pad; };  // one decoded sequence: literal run, then a match `offset` back
struct ZstdSeqFrame { uint32_t read, unit_first, unit_count, pad; };  // a frame with Seq units: its units, in block order
struct ZstdDecParams {
    const uint8_t* src;
    uint8_t* dst;
    const ZstdRead* reads;
    uint32_t* results;   // per read: content bytes, or kErrZstd
    uint8_t* lit_scratch;  // 128 KiB + 32 per warp of the general kernel's / the sequence executor's grid
    uint32_t* counter;   // 16 words, zeroed before launch: [0] work counter of the general kernel, [1] units listed,
                         // [2] frames left to the general kernel, [3] work counter of the units kernel, [4] groups listed,
                         // [5] sequence records allotted, [6] frames with Seq units, [7] / [8] work counters of the two sequence kernels,
                         // [9] entries of seq_list, [10] chains listed for the sequence kernel, [11] its second work counter
    ZstdUnit* units;     // unit_cap entries
    ZstdGroup* groups;   // unit_cap entries
    uint32_t* general;   // n_reads entries: the frames the walker left to the general kernel
    ZstdSeq* seqs;       // seq_cap records
    ZstdSeqFrame* seq_frames;  // n_reads entries
    uint32_t* seq_list;  // unit_cap entries: the Seq units of the frames the sequence kernel vouches for, in block order (the executor's work list);
                         // behind them unit_cap ZstdSeqChain records (csrc/zstd_dec_kernels.cuh): the sequence kernel's own work list
    uint32_t unit_cap;
    uint32_t n_reads;
    uint32_t seq_cap;
    uint32_t pad;
};
// unit slots the host reserves per frame.  This library's writer: a block per 8 KiB of content, per 2 KiB of literals in the S groups of the
// control-byte region (a fifth of a stage-1 stream) - content / 5 KiB at most; libzstd: a block per 128 KiB.  A frame of a batch that runs
// the pool dry (other writers may cut blocks of any size) is decoded by the general kernel.
VBZ_HD uint32_t zstd_units_of(uint32_t dst_cap) { return dst_cap / 4096u + 8u; }

struct ZstdEncRead {
    uint64_t src_off;     // the content: bytes [src_off, src_off + size) of ZstdEncParams::src
    uint64_t dst_off;     // the output slot, offset into ZstdEncParams::dst
    uint32_t src_bytes;   // content size, unless ZstdEncParams::src_sizes says otherwise
    uint32_t dst_cap;     // capacity of the output slot (sized header included)
    uint32_t first_block; // index of the frame's first block
    uint32_t n_blocks;    // blocks reserved for the frame (enough for the largest content stage 1 could produce)
    uint32_t hdr_value;   // value of the 4-byte "sized" header: the original byte count (ref vbz/vbz.cpp:316-317)
    uint32_t rejected;    // != 0: decided on the host; copy `preset` to results[]
    uint32_t preset;
    uint32_t split;       // blocks never straddle this content offset (the control-byte / data-byte boundary of stage 1); 0 = none
};
struct ZstdEncParams {
    const uint8_t* src;
    uint8_t* dst;
    const ZstdEncRead* reads;
    const uint32_t* block_read;  // block -> read
    const uint32_t* src_sizes;   // optional per-read content size written by stage 1; >= kErrOutOfMemory is passed through
    uint8_t* codes;              // n_blocks records of kZGroupCodeBytes: the Huffman code of every group (written by zstd_enc_code_kernel)
    uint64_t* status;            // look-back chain over the blocks of a frame, zeroed before launch
    uint32_t* results;           // per read: bytes written (headers included) or an error code
    uint32_t hdr_bytes;          // 4 = put the "sized" header before the frame
    uint32_t err_add;            // what the sized entry point adds to error codes (ref vbz/vbz.cpp:329)
    uint32_t n_reads, n_blocks;
    uint32_t level;              // zstd_compression_level of the call (1, 2, 3 ... select the match finder's effort)
    uint32_t pad;
    // scratch of the match finder (csrc/zstd_match_kernels.cuh)
    uint8_t* lits;               // as large as the content buffer: the literals of an S group sit at the group's content offset, its sequences sections behind them
    uint64_t* seqs;              // sequence records: content extent / 16 + n_blocks + 1
    uint8_t* meta;               // n_blocks records of kZGroupMetaBytes
    uint32_t* hist;              // n_blocks x 256: literal histogram of every S group
};
constexpr uint32_t kZEncBlock = 128u * 1024u;  // content bytes per GROUP of the encoder: blocks that share one Huffman table
constexpr uint32_t kZEncBlockKeys = 64u * 1024u;  // ... in the control-byte region of a stage-1 stream (csrc/zstd_match_kernels.cuh: zenc_locate)
constexpr uint32_t kZEncSub = 8u * 1024u;     // content bytes per block written by the encoder
constexpr uint32_t kZGroupMetaBytes = 144;    // sizeof(ZGroupMeta), csrc/zstd_match_kernels.cuh
constexpr uint32_t kZGroupCodeBytes = 1200;   // sizeof(ZGroupCode), csrc/zstd_enc_kernels.cuh
constexpr uint32_t kZBlockMax = 128u * 1024u;
constexpr int kZdecWarps = 4;
constexpr uint32_t kZLitStride = kZBlockMax + 32;  // literal scratch per decoding warp

constexpr int kSvbThreads = 256;
constexpr int kSvbWarps = kSvbThreads / 32;
constexpr int kSvbGroups = 2;                                    // groups of 16 raw bytes per lane
constexpr int kSvbWarpTileBytes = 32 * kSvbGroups * 16;          // raw bytes per warp tile (1 KiB)
constexpr int kSvbTileBytes = kSvbWarps * kSvbWarpTileBytes;     // raw bytes per (CTA) tile (8 KiB)
constexpr uint64_t kStAggregate = 1ull << 62, kStInclusive = 2ull << 62, kStValueMask = (1ull << 62) - 1;

// values per tile: the int16 x86-semantics path (csrc/svb16_kernels.cuh) uses 16 KiB tiles, the generic kernels 8 KiB
VBZ_HD uint32_t svb_tile_values(int isz, int mode) { return (isz == 2 && mode == kModeI16X86) ? 8192u : (uint32_t)(kSvbTileBytes / isz); }
VBZ_HD uint32_t svb_tiles_of(uint32_t n_values, uint32_t tile_values)
{
    uint32_t t = (uint32_t)(((uint64_t)n_values + tile_values - 1) / tile_values);
    return t ? t : 1;  // an empty read still owns one (empty) tile: it writes the header and the result
}

// ---- host <-> device transfers of a wave by kernel (csrc/xfer_kernels.cuh) ------------------------------------------
// One read of a transfer.  `n` is the byte count, unless the launch carries a device array of sizes (downloads
// after compression: the size is the kernel-produced result): then bytes = sizes[idx] when that is a valid size
// no larger than n, else nothing is copied (the host reports the error).
struct XferRead {
    uint64_t src;
    uint64_t dst;
    uint32_t n;
    uint32_t idx;
    uint32_t first_chunk;
    uint32_t pad;
};
struct XferParams {
    const XferRead* reads;
    const uint32_t* sizes;  // nullptr: XferRead::n is exact
    uint32_t n_reads;
    uint32_t n_chunks;
};

constexpr uint32_t kXferThreads = 128;   // a CTA fits next to the codec's persistent CTAs (<= 16K registers)
constexpr uint32_t kXferChunkVecs = 1024;  // 128-bit words of the destination per chunk (16 KiB)
constexpr uint32_t kXferRows = 8;          // rows of 32 x 16 bytes a warp keeps in flight

// chunks a read of at most n bytes can need, whatever its destination alignment
VBZ_HD uint32_t xfer_chunks_of(uint32_t n) { return n / (16u * kXferChunkVecs) + 1; }

}  // namespace vbz
