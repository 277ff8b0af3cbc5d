// csrc/host/batch.cpp - the C++ host layer: turns a batch of reads into descriptor tables, stages
// buffers, and drives the CUDA kernels through the C-ABI shim (csrc/vbzk.h).
//
// Per-read semantics follow the reference call for call:
//   compress    vbz_compress / vbz_compress_sized       (ref vbz/vbz.cpp:116-208, 302-330)
//   decompress  vbz_decompress / vbz_decompress_sized   (ref vbz/vbz.cpp:210-300, 332-367)
// including the order in which option / size errors are detected.  All sample arithmetic happens in
// the kernels; the host only validates sizes, lays buffers out and moves bytes.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/vbz_b200.h"
#include "../svb_types.h"
#include "../vbzk.h"
#include "context.h"
#include "copy_pool.h"
#include "options.h"

namespace vbz {

// ---- errors -----------------------------------------------------------------------------------------
static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }
const char* get_last_error() { return t_last_error.c_str(); }
bool cuda_ok(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    set_last_error(std::string(what) + ": " + cudaGetErrorString(e));
    return false;
}

bool DevBuf::ensure(size_t n)
{
    if (n <= cap) return true;
    release();
    size_t want = std::max(n, cap + cap / 2);
    want = (want + 255) & ~(size_t)255;
    if (!cuda_ok(cudaMalloc(&p, want), "cudaMalloc")) { p = nullptr; return false; }
    cap = want;
    return true;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
bool PinBuf::ensure(size_t n)
{
    if (n <= cap) return true;
    release();
    size_t want = std::max(n, cap + cap / 2);
    want = (want + 255) & ~(size_t)255;
    if (!cuda_ok(cudaMallocHost(&p, want), "cudaMallocHost")) { p = nullptr; return false; }
    cap = want;
    return true;
}
void PinBuf::release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }

bool StageRing::init()
{
    for (cudaEvent_t& e : ev)
        if (!cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) return false;
    return true;
}
void* StageRing::take(size_t n)
{
    // every block has the size of the largest request so far: a block that grows (cudaFreeHost + cudaMallocHost, which
    // synchronise the device) in the middle of a run cost one bench step half a second
    if (n > high_water) {
        high_water = n + n / 4 + 4096;
        for (int i = 0; i < kEntries; ++i) {
            if (!cuda_ok(cudaEventSynchronize(ev[i]), "cudaEventSynchronize")) return nullptr;
            if (!buf[i].ensure(high_water)) return nullptr;
        }
    }
    cur = (cur + 1) % kEntries;
    if (!cuda_ok(cudaEventSynchronize(ev[cur]), "cudaEventSynchronize")) return nullptr;
    return buf[cur].p;
}
bool StageRing::mark(cudaStream_t s) { return cuda_ok(cudaEventRecord(ev[cur], s), "cudaEventRecord"); }
void StageRing::release()
{
    for (PinBuf& b : buf) b.release();
    for (cudaEvent_t& e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
}

static bool zero_device(vbz_b200_ctx* ctx, void* d, size_t bytes, cudaStream_t stream)
{
    const int e = vbzk_zero_words(d, bytes, stream);
    if (e) { cuda_ok((cudaError_t)e, "reset device state"); return false; }
    ctx->launches += 1;
    return true;
}
// copies n bytes of descriptors to the device through the pinned ring, without waiting for the stream
static bool upload_descriptors(vbz_b200_ctx* ctx, void* d_to, const void* from, size_t n, cudaStream_t stream)
{
    void* h = ctx->stage.take(n + 4);
    if (!h) return false;
    memcpy(h, from, n);
    const int e = vbzk_copy_words(d_to, h, (n + 3) & ~(size_t)3, stream);  // by kernel: never queues behind a wave's DMA
    if (e) { cuda_ok((cudaError_t)e, "upload descriptors"); return false; }
    ctx->launches += 1;
    return ctx->stage.mark(stream);
}

static inline uint64_t align16(uint64_t v) { return (v + 15) & ~(uint64_t)15; }

// One read of a device-resident batch as the host sees it.
struct ReadJob {
    uint64_t src_off = 0, dst_off = 0;
    uint32_t src_size = 0;   // bytes at src_off (decode: including the sized header, if any)
    uint32_t dst_size = 0;   // compress: capacity of the slot; decompress: exact decompressed size
    bool rejected = false;   // decided on the host; `code` is the result
    uint32_t code = 0;
    uint32_t post_code = 0;  // decompress with a zstd stage: what the read returns once its frame has decoded (0 = go on to stage 1)
    uint64_t mid_cap = 0;    // decompress with a zstd stage: content size of the frame if the host could read it (0 = use the bound)
};

// ---- stage 1 on the device -----------------------------------------------------------------------------
// Enqueues: descriptor upload, tile table, look-back reset, the kernel.  reads[i].first_tile is set here.
static int enqueue_stage1(vbz_b200_ctx* ctx, bool decode, const Stage1Kind& kind, const uint8_t* d_src, uint8_t* d_dst,
                          std::vector<SvbRead>& reads, uint32_t hdr_bytes, uint32_t* d_results, const uint32_t* d_src_sizes,
                          cudaStream_t stream)
{
    const uint32_t n = (uint32_t)reads.size();
    if (n == 0) return 0;
    if (kind.half) {  // v1 nibble packing of int8: one warp per read, no tiles, no look-back (csrc/svb_half_kernels.cuh)
        if (!ctx->d_reads.ensure((size_t)n * sizeof(SvbRead))) return -1;
        if (!upload_descriptors(ctx, ctx->d_reads.p, reads.data(), (size_t)n * sizeof(SvbRead), stream)) return -1;
        SvbParams p;
        memset(&p, 0, sizeof(p));
        p.src = d_src;
        p.dst = d_dst;
        p.reads = ctx->d_reads.as<SvbRead>();
        p.results = d_results;
        p.src_sizes = d_src_sizes;
        p.hdr_bytes = hdr_bytes;
        p.n_reads = n;
        const VbzkSvbParams* pp = reinterpret_cast<const VbzkSvbParams*>(&p);
        const int e = decode ? vbzk_svb_half_decode(kind.zz, pp, stream) : vbzk_svb_half_encode(kind.zz, pp, stream);
        if (e) { cuda_ok((cudaError_t)e, decode ? "svb_half_decode launch" : "svb_half_encode launch"); return -1; }
        ctx->launches += 1;
        return 0;
    }
    uint64_t tiles = 0;
    const uint32_t tile_values = vbzk_svb_tile_values(kind.isz, kind.mode);
    for (uint32_t i = 0; i < n; ++i) {
        reads[i].first_tile = (uint32_t)tiles;
        tiles += svb_tiles_of(reads[i].n_values, tile_values);
    }
    if (tiles > 0x7FFFFFFFull) { set_last_error("batch too large: more than 2^31 tiles"); return -1; }
    const uint32_t n_tiles = (uint32_t)tiles;
    const size_t status_words = (size_t)n_tiles * (decode ? 2 : 1);
    if (!ctx->d_reads.ensure((size_t)n * sizeof(SvbRead)) || !ctx->d_tiles.ensure((size_t)n_tiles * 4) ||
        !ctx->d_status.ensure((status_words + 1) * 8) || (decode && !ctx->d_wtoff.ensure((size_t)n_tiles * kSvbWarps * 4)) ||
        (kind.mode == kModeI16X86 && !ctx->d_issue.ensure((size_t)n_tiles * 32)))
        return -1;
    if (!upload_descriptors(ctx, ctx->d_reads.p, reads.data(), (size_t)n * sizeof(SvbRead), stream)) return -1;
    if (!zero_device(ctx, ctx->d_status.p, (status_words + 1) * 8, stream)) return -1;

    int e = vbzk_svb_fill_tiles(ctx->d_reads.as<VbzkSvbRead>(), n, tile_values, ctx->d_tiles.as<uint32_t>(), stream);
    if (e) { cuda_ok((cudaError_t)e, "svb_fill_tiles launch"); return -1; }
    SvbParams p;
    p.src = d_src;
    p.dst = d_dst;
    p.reads = ctx->d_reads.as<SvbRead>();
    p.tile_read = ctx->d_tiles.as<uint32_t>();
    p.status_a = ctx->d_status.as<uint64_t>();
    p.status_b = ctx->d_status.as<uint64_t>() + n_tiles;
    p.issue_a = reinterpret_cast<SvbLoad*>(ctx->d_issue.p);
    p.issue_b = p.issue_a ? reinterpret_cast<SvbLoad*>(ctx->d_issue.p) + n_tiles : nullptr;
    p.ticket = reinterpret_cast<uint32_t*>(ctx->d_status.as<uint64_t>() + status_words);
    p.results = d_results;
    p.src_sizes = d_src_sizes;
    p.hdr_bytes = hdr_bytes;
    p.wtile_off = decode ? ctx->d_wtoff.as<uint32_t>() : nullptr;
    p.n_reads = n;
    p.n_tiles = n_tiles;
    { const char* dbg = getenv("VBZ_B200_DEBUG"); p.debug = dbg ? (uint32_t)atoi(dbg) : 0u; }
    p.ctas_per_sm = ctx->codec_ctas_per_sm;
    const VbzkSvbParams* pp = reinterpret_cast<const VbzkSvbParams*>(&p);
    if (decode) {
        e = vbzk_svb_key_offsets(kind.isz, kind.mode, pp, stream);
        if (e) { cuda_ok((cudaError_t)e, "svb_key_offsets launch"); return -1; }
        ctx->launches += 1;
    }
    e = decode ? vbzk_svb_decode(kind.isz, kind.zz, kind.mode, pp, n_tiles, stream)
               : vbzk_svb_encode(kind.isz, kind.zz, kind.mode, pp, n_tiles, stream);
    if (e) { cuda_ok((cudaError_t)e, decode ? "svb_decode launch" : "svb_encode launch"); return -1; }
    ctx->launches += kind.mode == kModeI16X86 ? 3 : 2;  // tile table, (int16: load descriptors / second pre-pass kernel,) the codec kernel
    return 0;
}

// ---- stage 2 (zstd), decode: frames -> content --------------------------------------------------------------------
// `phase`: 0 = everything on `stream`; 1 = the front half only (descriptors, walker, sequence kernel; *keep receives the launch parameters);
// 2 = the back half with the parameters phase 1 left in *keep (units kernel, executor, general kernel: what writes d_dst)
static int enqueue_zstd_decode(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const std::vector<ZstdRead>& reads, uint32_t* d_results,
                               cudaStream_t stream, int phase = 0, ZstdDecParams* keep = nullptr)
{
    const uint32_t n = (uint32_t)reads.size();
    if (n == 0) return 0;
    if (phase == 2) {
        if (!ctx->d_zlit.ensure(vbzk_zstd_decode_scratch_bytes())) return -1;
        keep->lit_scratch = ctx->d_zlit.as<uint8_t>();
        const int e = vbzk_zstd_decode_back(reinterpret_cast<const VbzkZstdDecParams*>(keep), stream);
        if (e) { cuda_ok((cudaError_t)e, "zstd_decode launch"); return -1; }
        ctx->launches += 3;
        ctx->last_zdecode_stream = stream;
        return 0;
    }
    const size_t bytes = (size_t)n * sizeof(ZstdRead);
    uint64_t unit_cap = 0;
    for (uint32_t i = 0; i < n; ++i)
        if (!reads[i].rejected) unit_cap += zstd_units_of(reads[i].dst_cap);
    unit_cap = std::min<uint64_t>(std::max<uint64_t>(unit_cap, 8), 0x7FFFFFF0u);
    // records for the sequences of frames with sequences: one per 32 content bytes on average (a signal frame needs one per ~150; a frame
    // of a batch that runs the pool dry is decoded by the general kernel instead)
    uint64_t seq_cap = 0;
    for (uint32_t i = 0; i < n; ++i)
        if (!reads[i].rejected) seq_cap += reads[i].dst_cap / 32u + 64u;
    seq_cap = std::min<uint64_t>(std::max<uint64_t>(seq_cap, 1024), 0x7FFFFFF0u);
    if (!ctx->d_zseqs.ensure((size_t)seq_cap * sizeof(ZstdSeq)) || !ctx->d_zseqframes.ensure((size_t)n * sizeof(ZstdSeqFrame))) return -1;
    if (!ctx->d_zreads.ensure(bytes) || !ctx->d_zcounter.ensure(256) || (phase == 0 && !ctx->d_zlit.ensure(vbzk_zstd_decode_scratch_bytes())) ||
        !ctx->d_zunits.ensure((size_t)unit_cap * sizeof(ZstdUnit)) || !ctx->d_zgroups.ensure((size_t)unit_cap * sizeof(ZstdGroup)) ||
        !ctx->d_zgeneral.ensure((size_t)n * 4) || !ctx->d_zseqlist.ensure((size_t)unit_cap * 20))  // the executor's list, and behind it the sequence kernel's chains (4 words each)
        return -1;
    if (!upload_descriptors(ctx, ctx->d_zreads.p, reads.data(), bytes, stream) ||
        !zero_device(ctx, ctx->d_zcounter.p, 256, stream))
        return -1;
    ZstdDecParams p;
    p.src = d_src;
    p.dst = d_dst;
    p.reads = ctx->d_zreads.as<ZstdRead>();
    p.results = d_results;
    p.lit_scratch = ctx->d_zlit.as<uint8_t>();
    p.counter = ctx->d_zcounter.as<uint32_t>();
    p.units = ctx->d_zunits.as<ZstdUnit>();
    p.groups = ctx->d_zgroups.as<ZstdGroup>();
    p.general = ctx->d_zgeneral.as<uint32_t>();
    p.seqs = ctx->d_zseqs.as<ZstdSeq>();
    p.seq_frames = ctx->d_zseqframes.as<ZstdSeqFrame>();
    p.seq_list = ctx->d_zseqlist.as<uint32_t>();
    p.unit_cap = (uint32_t)unit_cap;
    p.n_reads = n;
    p.seq_cap = (uint32_t)seq_cap;
    p.pad = 0;
    if (phase == 1) {
        *keep = p;
        const int e = vbzk_zstd_decode_front(reinterpret_cast<const VbzkZstdDecParams*>(&p), stream);
        if (e) { cuda_ok((cudaError_t)e, "zstd_decode launch"); return -1; }
        ctx->launches += 2;
        return 0;
    }
    const int e = vbzk_zstd_decode(reinterpret_cast<const VbzkZstdDecParams*>(&p), stream);
    if (e) { cuda_ok((cudaError_t)e, "zstd_decode launch"); return -1; }
    ctx->launches += 5;  // walker, sequence kernel, units kernel, sequence executor, general kernel
    ctx->last_zdecode_stream = stream;
    return 0;
}

// ---- stage 2 (zstd), encode: content -> frames ----------------------------------------------------------------------
// reads[i].first_block is set here.
static int enqueue_zstd_encode(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, std::vector<ZstdEncRead>& reads, const uint32_t* d_src_sizes,
                               uint32_t* d_results, bool sized, unsigned level, uint64_t src_extent, cudaStream_t stream)
{
    const uint32_t n = (uint32_t)reads.size();
    if (n == 0) return 0;
    uint64_t blocks = 0;
    for (uint32_t i = 0; i < n; ++i) {
        reads[i].first_block = (uint32_t)blocks;
        blocks += reads[i].n_blocks;
    }
    if (blocks > 0x7FFFFFFFull) { set_last_error("batch too large: more than 2^31 zstd blocks"); return -1; }
    const uint32_t n_blocks = (uint32_t)blocks;
    const size_t bytes = (size_t)n * sizeof(ZstdEncRead);
    if (!ctx->d_zreads.ensure(bytes) || !ctx->d_zblocks.ensure((size_t)n_blocks * 4) || !ctx->d_zstatus.ensure((size_t)n_blocks * 8) ||
        !ctx->d_zcodes.ensure((size_t)n_blocks * kZGroupCodeBytes) ||
        // scratch of the match finder: literals of the S groups (at their content offsets), sequence records, per-group meta + histogram
        !ctx->d_zelits.ensure((size_t)src_extent + 64) || !ctx->d_zeseqs.ensure(((size_t)src_extent / 16 + n_blocks + 2) * 8) ||
        !ctx->d_zemeta.ensure((size_t)n_blocks * kZGroupMetaBytes) || !ctx->d_zehist.ensure((size_t)n_blocks * 1024))
        return -1;
    if (!upload_descriptors(ctx, ctx->d_zreads.p, reads.data(), bytes, stream) ||
        !zero_device(ctx, ctx->d_zstatus.p, (size_t)n_blocks * 8, stream))
        return -1;
    int e = vbzk_zstd_enc_fill_blocks(ctx->d_zreads.as<VbzkZstdEncRead>(), n, ctx->d_zblocks.as<uint32_t>(), stream);
    if (e) { cuda_ok((cudaError_t)e, "zstd_enc_fill_blocks launch"); return -1; }
    ZstdEncParams p;
    p.src = d_src;
    p.dst = d_dst;
    p.reads = ctx->d_zreads.as<ZstdEncRead>();
    p.block_read = ctx->d_zblocks.as<uint32_t>();
    p.src_sizes = d_src_sizes;
    p.codes = ctx->d_zcodes.as<uint8_t>();
    p.status = ctx->d_zstatus.as<uint64_t>();
    p.results = d_results;
    p.hdr_bytes = sized ? 4u : 0u;
    p.err_add = sized ? 4u : 0u;
    p.n_reads = n;
    p.n_blocks = n_blocks;
    p.level = level;
    p.pad = 0;
    p.lits = ctx->d_zelits.as<uint8_t>();
    p.seqs = ctx->d_zeseqs.as<uint64_t>();
    p.meta = ctx->d_zemeta.as<uint8_t>();
    p.hist = ctx->d_zehist.as<uint32_t>();
    e = vbzk_zstd_encode(reinterpret_cast<const VbzkZstdEncParams*>(&p), stream);
    if (e) { cuda_ok((cudaError_t)e, "zstd_encode launch"); return -1; }
    ctx->launches += 5;  // block table, matches, sequences sections, codes, frames
    return 0;
}

// ---- compress / decompress, device-resident ------------------------------------------------------------------
// The two-stage paths keep the stage-1 bytes of the reads in flight in an intermediate buffer (d_mid).  A batch is cut into
// WAVES of consecutive reads whose intermediate slots fit a budget (VBZ_B200_MID_MIB, default 8 GiB): the buffer stays
// bounded whatever the batch size, so a batch is limited by its own input and output buffers only.  Waves run one after
// the other on the caller's stream; each is large enough (GiBs) that its dozen launches do not show.
static uint64_t mid_budget()
{
    const char* e = getenv("VBZ_B200_MID_MIB");  // (read per call: tests change it)
    const long m = e ? atol(e) : 0;
    return (uint64_t)(m > 0 ? m : 8192) << 20;
}
// format bound of a stage-1 stream of n values: control bytes + 4 data bytes per value (ref streamvbyte.h:33-39 without the padding).
// A valid stream written by ANY build of the reference fits (the non-SSSE3 builds use 3-byte codes for int16 deltas beyond
// 16 bits, so the x86 encoder's keys + 2N is not a decode-side bound).
static inline uint64_t stage1_format_bound(uint32_t n_values) { return ((uint64_t)n_values + 3) / 4 + 4ull * n_values; }

static void stage_mark(vbz_b200_ctx* ctx, cudaStream_t stream)
{
    if (!ctx->stage_timing) return;
    if (ctx->t_used == ctx->t_marks.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        ctx->t_marks.push_back(e);
    }
    cudaEventRecord(ctx->t_marks[ctx->t_used++], stream);
}

static int compress_wave(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const ReadJob* jobs, uint32_t n, uint32_t* d_results,
                         const Stage1Kind& kind, bool sized, cudaStream_t stream)
{
    const bool zstd = kind.zstd_level != 0;
    const uint32_t blk = vbzk_zstd_enc_block_bytes(), kblk = vbzk_zstd_enc_key_block_bytes();
    std::vector<uint64_t> mid_off(zstd ? n : 0);
    uint64_t mid_total = 0;
    if (kind.isz != 0) {
        // stage 1: straight into the destination (with the sized header) when there is no zstd stage, else into the
        // intermediate buffer (ref vbz/vbz.cpp:163-186)
        std::vector<SvbRead> reads(n);
        for (uint32_t i = 0; i < n; ++i) {
            const ReadJob& j = jobs[i];
            SvbRead& r = reads[i];
            memset(&r, 0, sizeof(r));
            r.src_off = j.src_off;
            r.dst_off = j.dst_off;
            r.src_bytes = j.src_size;
            r.rejected = j.rejected;
            r.preset = j.code;
            if (!j.rejected) {
                if ((j.src_off | j.dst_off) & 15) { set_last_error("device batch offsets must be multiples of 16"); return -1; }
                r.n_values = j.src_size / (unsigned)kind.isz;
                if (zstd) {
                    mid_off[i] = mid_total;
                    r.dst_off = mid_total;
                    mid_total += align16(stage1_tight_bound(r.n_values, kind)) + 16;
                }
            }
        }
        if (!zstd) return enqueue_stage1(ctx, false, kind, d_src, d_dst, reads, sized ? 4u : 0u, d_results, nullptr, stream);
        if (!ctx->d_mid.ensure(mid_total + 16) || !ctx->d_sizes.ensure((size_t)n * 4)) return -1;
        if (enqueue_stage1(ctx, false, kind, d_src, ctx->d_mid.as<uint8_t>(), reads, 0, ctx->d_sizes.as<uint32_t>(), nullptr, stream) != 0) return -1;
        stage_mark(ctx, stream);
    }
    // stage 2: one zstd frame per read (ref vbz/vbz.cpp:194-204)
    std::vector<ZstdEncRead> zr(n);
    for (uint32_t i = 0; i < n; ++i) {
        const ReadJob& j = jobs[i];
        ZstdEncRead& z = zr[i];
        memset(&z, 0, sizeof(z));
        z.rejected = j.rejected;
        z.preset = j.code;
        z.n_blocks = 1;
        if (j.rejected) continue;
        z.dst_off = j.dst_off;
        z.dst_cap = j.dst_size;
        z.hdr_value = j.src_size;
        uint64_t largest = j.src_size;
        if (kind.isz != 0) {
            const uint32_t nv = j.src_size / (unsigned)kind.isz;
            z.src_off = mid_off[i];
            z.split = (uint32_t)(((uint64_t)nv + 3) / 4);  // control bytes | data bytes
            if (z.split < 512) z.split = 0;                 // too small to be worth a block (and a table) of its own
            largest = stage1_tight_bound(nv, kind);
        } else {
            z.src_off = j.src_off;
            z.src_bytes = j.src_size;
        }
        z.n_blocks = (uint32_t)std::max<uint64_t>(1, (z.split + kblk - 1) / kblk + (largest - z.split + blk - 1) / blk);
    }
    uint64_t extent = mid_total + 16;
    if (kind.isz == 0) {
        extent = 0;
        for (uint32_t i = 0; i < n; ++i)
            if (!jobs[i].rejected) extent = std::max<uint64_t>(extent, jobs[i].src_off + jobs[i].src_size);
    }
    return enqueue_zstd_encode(ctx, kind.isz != 0 ? ctx->d_mid.as<uint8_t>() : d_src, d_dst, zr, kind.isz != 0 ? ctx->d_sizes.as<uint32_t>() : nullptr,
                               d_results, sized, kind.zstd_level, extent, stream);
}

static int compress_device_impl(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const std::vector<ReadJob>& jobs,
                                uint32_t* d_results, const Stage1Kind& kind, bool sized, cudaStream_t stream)
{
    const uint32_t n = (uint32_t)jobs.size();
    if (kind.zstd_level == 0 && kind.isz == 0) { set_last_error("plain copies (integer_size 0, level 0) are not a device operation"); return -1; }
    if (kind.zstd_level == 0 || kind.isz == 0) return compress_wave(ctx, d_src, d_dst, jobs.data(), n, d_results, kind, sized, stream);
    const uint64_t budget = mid_budget();
    for (uint32_t a = 0; a < n;) {
        uint64_t mid = 0;
        uint32_t b = a;
        while (b < n) {
            const uint64_t m = jobs[b].rejected ? 0 : align16(stage1_tight_bound(jobs[b].src_size / (unsigned)kind.isz, kind)) + 16;
            if (b > a && mid + m > budget) break;
            mid += m;
            ++b;
        }
        if (a) stage_mark(ctx, stream);  // wave boundary: the previous wave's zstd stage ends here
        if (compress_wave(ctx, d_src, d_dst, jobs.data() + a, b - a, d_results + a, kind, sized, stream) != 0) return -1;
        a = b;
    }
    return 0;
}

// intermediate capacity of one read on the decode side
static inline uint64_t decode_mid_cap(const ReadJob& j, const Stage1Kind& kind)
{
    const uint64_t bound = stage1_format_bound(j.dst_size / (unsigned)kind.isz) + 1;  // + 1: a frame one byte too long must not pass for a valid one
    uint64_t cap = j.mid_cap ? std::min<uint64_t>(j.mid_cap, bound) : bound;         // device-resident frames: the host cannot read the header
    return std::min<uint64_t>(cap, 0xFFFFFFF0u);
}

// what the front half of a two-stage wave leaves for its back half (decoder look-ahead)
struct DecodeAhead { ZstdDecParams p; std::vector<uint64_t> mid_off; uint32_t* sizes = nullptr; };

// `phase` / `ahead`: as enqueue_zstd_decode - 1 = the zstd stage's front half only, 2 = its back half and stage 1 (two-stage calls only)
static int decompress_wave(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const ReadJob* jobs, uint32_t n, uint32_t* d_results,
                           const Stage1Kind& kind, bool sized, cudaStream_t stream, int phase = 0, DecodeAhead* ahead = nullptr)
{
    const uint32_t hdr = sized ? 4u : 0u;
    const uint8_t* s1_src = d_src;
    const uint32_t* s1_sizes = nullptr;
    std::vector<uint64_t> mid_off;
    if (phase == 2) {
        std::vector<ZstdRead> none(n);  // (only its size is looked at)
        if (enqueue_zstd_decode(ctx, d_src, nullptr, none, nullptr, stream, 2, &ahead->p) != 0) return -1;
        stage_mark(ctx, stream);
        mid_off.swap(ahead->mid_off);
        s1_src = ctx->d_mid.as<uint8_t>();
        s1_sizes = ahead->sizes;
    } else if (kind.zstd_level != 0) {
        // stage 2 first (ref vbz/vbz.cpp:234-273): each frame decodes into its slot of the intermediate buffer - or
        // straight into the destination when there is no stage 1
        std::vector<ZstdRead> zr(n);
        mid_off.resize(n);
        uint64_t mid_total = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const ReadJob& j = jobs[i];
            ZstdRead& z = zr[i];
            memset(&z, 0, sizeof(z));
            z.rejected = j.rejected;
            z.preset = j.code;
            if (j.rejected) continue;
            z.src_off = j.src_off + hdr;
            z.src_bytes = j.src_size - hdr;
            if (kind.isz == 0) {
                z.dst_off = j.dst_off;
                z.dst_cap = j.dst_size;
            } else {
                const uint64_t cap = decode_mid_cap(j, kind);
                mid_off[i] = mid_total;
                z.dst_off = mid_total;
                z.dst_cap = (uint32_t)cap;
                mid_total += align16(cap) + 16;
            }
        }
        uint32_t* zres = d_results;
        uint8_t* zdst = d_dst;
        if (kind.isz != 0) {
            if (!ctx->d_mid.ensure(mid_total + 16) || !ctx->d_sizes.ensure((size_t)n * 4)) return -1;
            zres = ctx->d_sizes.as<uint32_t>();
            zdst = ctx->d_mid.as<uint8_t>();
        }
        if (phase == 1) {
            if (enqueue_zstd_decode(ctx, d_src, zdst, zr, zres, stream, 1, &ahead->p) != 0) return -1;
            ahead->mid_off.swap(mid_off);
            ahead->sizes = zres;
            return 0;
        }
        if (enqueue_zstd_decode(ctx, d_src, zdst, zr, zres, stream) != 0) return -1;
        if (kind.isz == 0) return 0;
        stage_mark(ctx, stream);
        s1_src = ctx->d_mid.as<uint8_t>();
        s1_sizes = ctx->d_sizes.as<uint32_t>();
    }
    std::vector<SvbRead> reads(n);
    for (uint32_t i = 0; i < n; ++i) {
        const ReadJob& j = jobs[i];
        SvbRead& r = reads[i];
        memset(&r, 0, sizeof(r));
        r.dst_off = j.dst_off;
        r.rejected = j.rejected || j.post_code != 0;
        r.preset = j.rejected ? j.code : j.post_code;
        if (kind.zstd_level != 0) {
            r.src_off = j.rejected ? 0 : mid_off[i];
            r.src_bytes = 0;  // taken from the zstd stage's result
        } else {
            r.src_off = j.src_off + hdr;
            r.src_bytes = j.src_size - (j.rejected ? 0 : hdr);
        }
        if (!r.rejected) {
            if ((j.src_off & 3) || (j.dst_off & 15)) { set_last_error("device batch offsets must be multiples of 16"); return -1; }
            r.n_values = j.dst_size / (unsigned)kind.isz;
        }
    }
    return enqueue_stage1(ctx, true, kind, s1_src, d_dst, reads, 0, d_results, s1_sizes, stream);
}

static int decompress_device_impl(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const std::vector<ReadJob>& jobs,
                                  uint32_t* d_results, const Stage1Kind& kind, bool sized, cudaStream_t stream)
{
    const uint32_t n = (uint32_t)jobs.size();
    if (kind.zstd_level == 0 && kind.isz == 0) { set_last_error("plain copies (integer_size 0, level 0) are not a device operation"); return -1; }
    if (kind.zstd_level == 0 || kind.isz == 0) return decompress_wave(ctx, d_src, d_dst, jobs.data(), n, d_results, kind, sized, stream);
    const uint64_t budget = mid_budget();
    std::vector<std::pair<uint32_t, uint32_t>> waves;
    uint64_t mid_max = 0;
    for (uint32_t a = 0; a < n;) {
        uint64_t mid = 0;
        uint32_t b = a;
        while (b < n) {
            const uint64_t m = jobs[b].rejected ? 0 : align16(decode_mid_cap(jobs[b], kind)) + 16;
            if (b > a && mid + m > budget) break;
            mid += m;
            ++b;
        }
        waves.push_back({a, b});
        mid_max = std::max(mid_max, mid);
        a = b;
    }
    // Look-ahead: the walker and the sequence kernel of wave w + 1 (latency-bound: a thread per frame hops over block headers, a frame's
    // sequences are serial work per block) run on the context's second stream while the kernels that write wave w's output keep the SMs busy.
    // They write units, records and per-frame sizes only, so those exist twice (ZDecSet); the intermediate buffer is one, sized up front.
    const char* la = getenv("VBZ_B200_DECODE_LOOKAHEAD");  // "0": one stream, wave after wave
    const bool lookahead_allowed = !(la && atoi(la) == 0);
    const bool ahead_ok = lookahead_allowed && waves.size() > 1 && ctx->scratch_in_use == 0 && ctx->zdec_in_use == 0 && stream != ctx->kstream[1] && ctx->ev_fork;
    if (!ahead_ok) {
        for (size_t w = 0; w < waves.size(); ++w) {
            if (w) stage_mark(ctx, stream);
            if (decompress_wave(ctx, d_src, d_dst, jobs.data() + waves[w].first, waves[w].second - waves[w].first, d_results + waves[w].first, kind, sized, stream) != 0) return -1;
        }
        return 0;
    }
    if (!ctx->d_mid.ensure(mid_max + 32)) return -1;  // (never reallocated while a wave is in flight)
    struct Restore { vbz_b200_ctx* c; ~Restore() { c->use_zdec(0); } } restore{ctx};
    cudaStream_t s2 = ctx->kstream[1];
    DecodeAhead st[2];
    int rc = 0;
    auto front = [&](size_t w) {
        ctx->use_zdec((int)(w & 1));
        if (w >= 2) cudaStreamWaitEvent(s2, ctx->ev_back[w & 1], 0);  // the set's previous user (wave w - 2) is done with it
        const int e = decompress_wave(ctx, d_src, d_dst, jobs.data() + waves[w].first, waves[w].second - waves[w].first, d_results + waves[w].first, kind, sized, s2, 1, &st[w & 1]);
        // the units kernel of this wave leaves a seventh of its persistent grid out: room for the look-ahead of the wave behind it to start
        // at once instead of behind it (measured: 100 % -> 414 GB/s over the step, 87 % -> 421-423, 75 % -> 416)
        st[w & 1].p.pad = 86;
        cudaEventRecord(ctx->ev_front[w & 1], s2);
        return e;
    };
    cudaEventRecord(ctx->ev_fork, stream);  // the second stream starts behind whatever the caller has queued before this call
    cudaStreamWaitEvent(s2, ctx->ev_fork, 0);
    rc = front(0);
    for (size_t w = 0; w < waves.size() && rc == 0; ++w) {
        if (w + 1 < waves.size()) rc = front(w + 1);
        if (rc != 0) break;
        if (w) stage_mark(ctx, stream);
        ctx->use_zdec((int)(w & 1));
        cudaStreamWaitEvent(stream, ctx->ev_front[w & 1], 0);
        rc = decompress_wave(ctx, d_src, d_dst, jobs.data() + waves[w].first, waves[w].second - waves[w].first, d_results + waves[w].first, kind, sized, stream, 2, &st[w & 1]);
        cudaEventRecord(ctx->ev_back[w & 1], stream);
    }
    if (rc != 0) { cudaStreamSynchronize(s2); cudaStreamSynchronize(stream); }
    return rc;
}

}  // namespace vbz

using namespace vbz;

// Every entry point runs on the context's device and leaves the caller's current device as it found it (a host process may
// drive several GPUs, e.g. torch with cuda:1 current while the H5Z filter uses device 0).
namespace {
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        ok = prev == device || cuda_ok(cudaSetDevice(device), "cudaSetDevice");
        if (prev == device) prev = -1;  // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

// ======================================================================================================
// C ABI (include/vbz_b200.h)
// ======================================================================================================
vbz_b200_ctx::~vbz_b200_ctx()
{
    release_all();
    for (vbz::ScratchSet& a : alt) a.release_all();
    h_results.release(); stage.release();
    for (WaveSlot& w : slots) {
        w.d_in.release(); w.d_out.release(); w.d_results.release(); w.d_pack.release(); w.d_xin.release(); w.d_xout.release();
        w.h_in.release(); w.h_out.release(); w.h_results.release();
        if (w.ev_in) cudaEventDestroy(w.ev_in);
        if (w.ev_k) cudaEventDestroy(w.ev_k);
        if (w.ev_out) cudaEventDestroy(w.ev_out);
        if (w.ev_outs) cudaEventDestroy(w.ev_outs);
    }
    for (cudaEvent_t e : t_marks) cudaEventDestroy(e);
    for (cudaEvent_t e : {ev_fork, ev_front[0], ev_front[1], ev_back[0], ev_back[1]})
        if (e) cudaEventDestroy(e);
    for (vbz::DevBuf* b : {&zdec_alt.d_zreads, &zdec_alt.d_zcounter, &zdec_alt.d_zunits, &zdec_alt.d_zgroups, &zdec_alt.d_zgeneral, &zdec_alt.d_zseqs,
                           &zdec_alt.d_zseqframes, &zdec_alt.d_zseqlist, &zdec_alt.d_sizes})
        b->release();
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    if (s_d2hs) cudaStreamDestroy(s_d2hs);
    for (int k = 1; k < kKernelStreams; ++k)
        if (kstream[k]) cudaStreamDestroy(kstream[k]);
    if (stream) cudaStreamDestroy(stream);
}

extern "C" int vbz_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" vbz_b200_ctx* vbz_b200_create(int device)
{
    int n = vbz_b200_device_count();
    if (device < 0 || device >= n) {
        set_last_error("no such CUDA device (" + std::to_string(device) + " of " + std::to_string(n) + "); libvbz_b200 has no CPU fallback");
        return nullptr;
    }
    DeviceGuard guard(device);
    if (!guard.ok) return nullptr;
    vbz_b200_ctx* c = new vbz_b200_ctx();
    c->device = device;
    bool ok = cuda_ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate") &&
              cuda_ok(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking), "cudaStreamCreate") &&
              cuda_ok(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking), "cudaStreamCreate") &&
              cuda_ok(cudaStreamCreateWithFlags(&c->s_d2hs, cudaStreamNonBlocking), "cudaStreamCreate") &&
              c->stage.init();
    c->kstream[0] = c->stream;
    ok = ok && cuda_ok(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming), "cudaEventCreate");
    for (int k = 0; ok && k < 2; ++k)
        ok = cuda_ok(cudaEventCreateWithFlags(&c->ev_front[k], cudaEventDisableTiming), "cudaEventCreate") &&
             cuda_ok(cudaEventCreateWithFlags(&c->ev_back[k], cudaEventDisableTiming), "cudaEventCreate");
    {
        // kstream[1] carries the decoder's look-ahead (the walker and the sequence kernel of the next wave): highest priority, so that its
        // CTAs are placed before the persistent grids of the wave that is being written fill the SMs
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        for (int k = 1; ok && k < vbz_b200_ctx::kKernelStreams; ++k)
            ok = cuda_ok(cudaStreamCreateWithPriority(&c->kstream[k], cudaStreamNonBlocking, k == 1 ? hi_prio : lo_prio), "cudaStreamCreate");
    }
    for (vbz_b200_ctx::WaveSlot& w : c->slots)
        ok = ok && cuda_ok(cudaEventCreateWithFlags(&w.ev_in, cudaEventDisableTiming), "cudaEventCreate") &&
             cuda_ok(cudaEventCreateWithFlags(&w.ev_k, cudaEventDisableTiming), "cudaEventCreate") &&
             cuda_ok(cudaEventCreateWithFlags(&w.ev_out, cudaEventDisableTiming), "cudaEventCreate") &&
             cuda_ok(cudaEventCreateWithFlags(&w.ev_outs, cudaEventDisableTiming), "cudaEventCreate");
    if (!ok) {
        delete c;
        return nullptr;
    }
    return c;
}

extern "C" void vbz_b200_destroy(vbz_b200_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    for (cudaStream_t ks : ctx->kstream)
        if (ks) cudaStreamSynchronize(ks);
    delete ctx;
}

extern "C" const char* vbz_b200_last_error(void) { return get_last_error(); }

extern "C" uint64_t vbz_b200_compress_bound(vbz_size_t source_size, const CompressionOptions* options, int sized)
{
    const Stage1Kind kind = classify_options(options);
    const uint64_t hdr = sized ? 4 : 0;
    if (kind.bad_isz || (kind.isz != 0 && (kind.bad_version || source_size % (unsigned)kind.isz != 0))) return hdr;
    uint64_t m = source_size;
    if (kind.isz != 0) m = stage1_tight_bound(source_size / (unsigned)kind.isz, kind);
    if (kind.zstd_level != 0) m = zstd_compress_bound(m);
    return align16(m + hdr);
}
extern "C" uint64_t vbz_b200_kernel_launches(vbz_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* vbz_b200_stream(vbz_b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int vbz_b200_synchronize(vbz_b200_ctx* ctx)
{
    if (!ctx) return -1;
    return cuda_ok(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize") ? 0 : -1;
}

extern "C" int vbz_b200_compress_device(vbz_b200_ctx* ctx, const void* d_src, const uint64_t* src_off, const vbz_size_t* src_size,
                                        void* d_dst, const uint64_t* dst_off, const vbz_size_t* dst_cap, vbz_size_t* d_results,
                                        uint32_t n, const CompressionOptions* options, int sized, void* stream)
{
    if (!ctx || !options) { set_last_error("null context or options"); return -1; }
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return -1;
    const Stage1Kind kind = classify_options(options);
    std::vector<ReadJob> jobs(n);
    for (uint32_t i = 0; i < n; ++i) {
        ReadJob& j = jobs[i];
        j.src_off = src_off[i]; j.dst_off = dst_off[i]; j.src_size = src_size[i]; j.dst_size = dst_cap[i];
        j.rejected = compress_precheck(kind, src_size[i], dst_cap[i], sized != 0, &j.code, true);
    }
    if (kind.bad_isz || (kind.isz != 0 && kind.bad_version)) {
        set_last_error("invalid options: every read of the batch would be rejected");
        return -1;
    }
    ctx->t_used = 0;
    ctx->t_decode = false;
    ctx->t_two_stage = kind.zstd_level != 0 && kind.isz != 0;
    stage_mark(ctx, (cudaStream_t)stream);
    const int rc = compress_device_impl(ctx, (const uint8_t*)d_src, (uint8_t*)d_dst, jobs, d_results, kind, sized != 0, (cudaStream_t)stream);
    stage_mark(ctx, (cudaStream_t)stream);
    return rc;
}

extern "C" int vbz_b200_decompress_device(vbz_b200_ctx* ctx, const void* d_src, const uint64_t* src_off, const vbz_size_t* src_size,
                                          void* d_dst, const uint64_t* dst_off, const vbz_size_t* dst_size, vbz_size_t* d_results,
                                          uint32_t n, const CompressionOptions* options, int sized, void* stream)
{
    if (!ctx || !options) { set_last_error("null context or options"); return -1; }
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return -1;
    const Stage1Kind kind = classify_options(options);
    if (kind.bad_isz || (kind.isz != 0 && kind.bad_version && kind.zstd_level == 0)) {
        set_last_error("invalid options: every read of the batch would be rejected");
        return -1;
    }
    const uint32_t hdr = sized ? 4u : 0u;
    std::vector<ReadJob> jobs(n);
    for (uint32_t i = 0; i < n; ++i) {
        ReadJob& j = jobs[i];
        j.src_off = src_off[i]; j.dst_off = dst_off[i]; j.src_size = src_size[i]; j.dst_size = dst_size[i];
        if (src_size[i] < hdr) { j.rejected = true; j.code = kErrInputSize; }
        else j.rejected = decompress_precheck(kind, src_size[i] - hdr, dst_size[i], &j.code);
        if (!j.rejected && kind.zstd_level != 0 && kind.isz != 0) {  // checked by the reference after the zstd stage (vbz.cpp:282-299)
            if (kind.bad_version) j.post_code = kErrVersion;
            else if (dst_size[i] % (unsigned)kind.isz != 0) j.post_code = kErrDestinationSize;
        }
    }
    ctx->t_used = 0;
    ctx->t_decode = true;
    ctx->t_two_stage = kind.zstd_level != 0 && kind.isz != 0;
    stage_mark(ctx, (cudaStream_t)stream);
    const int rc = decompress_device_impl(ctx, (const uint8_t*)d_src, (uint8_t*)d_dst, jobs, d_results, kind, sized != 0, (cudaStream_t)stream);
    stage_mark(ctx, (cudaStream_t)stream);
    return rc;
}

// where the frames of the last zstd-stage decode (the last wave of the last call) went: out[0] = units listed, out[1] = frames decoded by
// the general (warp per frame) kernel, out[2] = frames with sequences decoded block-parallel, out[3] = sequence records used
extern "C" int vbz_b200_decode_stats(vbz_b200_ctx* ctx, uint32_t out[4])
{
    if (!ctx || !out || !ctx->d_zcounter.p) { set_last_error("no zstd-stage decode has run on this context"); return -1; }
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return -1;
    uint32_t c[16] = {};
    if (!cuda_ok(cudaStreamSynchronize(ctx->last_zdecode_stream), "cudaStreamSynchronize") ||
        !cuda_ok(cudaMemcpy(c, ctx->d_zcounter.p, sizeof(c), cudaMemcpyDeviceToHost), "cudaMemcpy"))
        return -1;
    out[0] = c[1]; out[1] = c[2]; out[2] = c[6]; out[3] = c[5];
    return 0;
}

extern "C" int vbz_b200_set_stage_timing(vbz_b200_ctx* ctx, int on)
{
    if (!ctx) return -1;
    ctx->stage_timing = on != 0;
    ctx->t_used = 0;
    return 0;
}
// The marks of the last *_device call: begin, then per wave (first stage end, second stage end).  A single-stage call has two marks.
extern "C" int vbz_b200_stage_times(vbz_b200_ctx* ctx, float* stage1_ms, float* zstd_ms)
{
    if (!ctx || !stage1_ms || !zstd_ms || ctx->t_used < 2) { set_last_error("stage timing is not enabled"); return -1; }
    if (!cuda_ok(cudaEventSynchronize(ctx->t_marks[ctx->t_used - 1]), "cudaEventSynchronize")) return -1;
    float first = 0.f, second = 0.f;
    for (size_t k = 0; k + 1 < ctx->t_used; ++k) {
        float ms = 0.f;
        if (!cuda_ok(cudaEventElapsedTime(&ms, ctx->t_marks[k], ctx->t_marks[k + 1]), "cudaEventElapsedTime")) return -1;
        if (!ctx->t_two_stage || (k & 1) == 0) first += ms; else second += ms;
    }
    if (!ctx->t_two_stage) { *stage1_ms = first; *zstd_ms = 0.f; return 0; }  // a single stage ran (reported as stage 1)
    if (ctx->t_decode) { *zstd_ms = first; *stage1_ms = second; }  // decode runs the zstd stage first
    else { *stage1_ms = first; *zstd_ms = second; }
    return 0;
}

// ---- host-pointer batches -------------------------------------------------------------------------------------
// A batch of host reads runs as WAVES (64 .. 768 MiB of input + output) through kWaveSlots slots (device in/out buffers, pinned staging,
// events).  The host enqueues the upload, the kernels and the download of wave w, then turns to wave w-D
// (D = kWaveSlots - 2), whose kernels have long finished, reads its per-read results, and finally hands wave w-D-1 to
// the caller.  The upload queue therefore always holds D waves and the PCIe link (both directions) and the SMs work at
// the same time; nothing on the host waits for a stream it has just fed.
//
// Moving bytes.  Measured on the B200 box (tools/probe/pcie_probe.cu, xfer_probe.cu): one large DMA runs at 55 GB/s,
// but every cudaMemcpyAsync costs the copy engine ~4 us on top of its transfer time (8-14 us while the link is busy
// in the other direction), so one copy per read (430 KB on average) delivers 36 GB/s, and 27 GB/s with both
// directions busy.  Hence no per-read DMA anywhere.  Per read and per direction:
//   direct  the caller's buffer is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory): ONE transfer
//           kernel per wave and direction (csrc/xfer_kernels.cuh) reads / writes the caller's buffers in place over
//           the link, any alignment, exact sizes (a compressed read's size is taken from the device-side result, so
//           the download does not wait for the host).  Uploads whose reads sit back to back in host memory are
//           fetched run by run with large DMAs instead (the copy engine coexists best with a downloading kernel:
//           48 GB/s each way) and spread to their 16-byte aligned slots by the same kernel, device to device;
//   staged  pageable buffers: host threads (copy_pool.h) gather the wave's staged reads into pinned staging, which
//           crosses the link as ONE DMA per wave.  Staged reads sit first in the wave's device buffers.
namespace vbz {

struct HostJob {
    ReadJob j;          // offsets are relative to the wave's buffers
    uint32_t read;      // index into the caller's arrays
    uint32_t in_slot;   // bytes reserved in the wave's input / output buffers (multiples of 16)
    uint32_t out_slot;
    bool in_direct = false, out_direct = false;
    uint64_t src_dev = 0, dst_dev = 0;  // device-visible addresses of page-locked caller buffers
};
struct HostBatch {
    bool compress, sized;
    Stage1Kind kind;
    const void* const* src;
    const vbz_size_t* src_size;
    void* const* dst;
    const vbz_size_t* dst_cap;
    vbz_size_t* results;
    std::vector<HostJob> jobs;
};
struct Wave {
    uint32_t k0, k1;
    uint64_t in_bytes, out_bytes;    // whole device buffers of the wave
    uint64_t in_staged, out_staged;  // their leading parts that travel through pinned staging
    uint64_t in_direct_bytes;        // bytes of the reads uploaded in place
    uint32_t in_runs;                // host-contiguous runs among them
    uint32_t n_in_direct, n_out_direct;
};

// page-locked and visible to the device?  -> its device-side address
static bool host_pinned(const void* p, uint64_t* dev)
{
#ifdef VBZ_CPUSIM
    *dev = (uint64_t)(uintptr_t)p;
    return getenv("VBZ_CPUSIM_PINNED") != nullptr;  // the emulator's "device" shares the address space
#else
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;
    *dev = (uint64_t)(uintptr_t)a.devicePointer;
    return true;
#endif
}

// Wave sizes.  Large waves amortise the per-wave costs (a dozen launches, the fixed part of the zstd kernels - a libzstd-written
// frame's sequences are a serial walk of several milliseconds whatever the wave holds), but the first wave's upload and the
// last wave's download overlap with nothing, so waves grow geometrically from 128 MiB to the cap (an eighth of the batch, at
// most 768 MiB) and shrink again towards the end of the batch, down to 64 MiB.  Measured on reference-written chunks, 4 GiB
// calls, host to host: 27.7 GB/s with 16 MiB .. 256 MiB waves, 33.7 with these, 35.9 with 768 MiB throughout.
// VBZ_B200_WAVE_MIB fixes the size (tests); VBZ_B200_WAVE_CAP_MIB / _DIV / _FIRST_MIB move the three numbers.
struct WaveSizer {
    size_t fixed = 0, cap, prev = 0;
    uint64_t first = 128ull << 20;
    explicit WaveSizer(uint64_t total)
    {
        if (const char* e = getenv("VBZ_B200_WAVE_MIB")) { const long v = atol(e); if (v > 0) fixed = (size_t)v << 20; }
        uint64_t top = 768, div = 8;
        if (const char* e = getenv("VBZ_B200_WAVE_CAP_MIB")) { const long v = atol(e); if (v > 0) top = (uint64_t)v; }
        if (const char* e = getenv("VBZ_B200_WAVE_DIV")) { const long v = atol(e); if (v > 0) div = (uint64_t)v; }
        if (const char* e = getenv("VBZ_B200_WAVE_FIRST_MIB")) { const long v = atol(e); if (v > 0) first = (uint64_t)v << 20; }
        cap = (size_t)std::min<uint64_t>(std::max<uint64_t>(total / div, first), top << 20);
    }
    size_t next(uint64_t remaining)
    {
        if (fixed) return fixed;
        uint64_t b = prev ? 2 * (uint64_t)prev : first;
        b = std::min<uint64_t>(std::min<uint64_t>(b, cap), remaining / 2);
        prev = (size_t)std::max<uint64_t>(b, std::min<uint64_t>(first, 64ull << 20));
        return prev;
    }
};
static uint32_t env_u32(const char* name, uint32_t dflt)
{
    if (const char* e = getenv(name)) { const long v = atol(e); if (v > 0) return (uint32_t)v; }
    return dflt;
}

struct HostTrace {  // VBZ_B200_TRACE=1: where the host thread of a batch spends its time (stderr)
    bool on = getenv("VBZ_B200_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double acc[8] = {};
    void lap(int k)
    {
        if (!on) return;
        const auto t = std::chrono::steady_clock::now();
        acc[k] += std::chrono::duration<double, std::milli>(t - t0).count();
        t0 = t;
    }
    // device-side timeline of the waves: events around upload / kernels / download
    std::vector<cudaEvent_t> ev;
    void mark(cudaStream_t s)
    {
#ifndef VBZ_CPUSIM
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.push_back(e);
#else
        (void)s;
#endif
    }
    void timeline()
    {
#ifndef VBZ_CPUSIM
        if (!on || ev.size() < 6) return;
        const size_t waves = ev.size() / 6;
        double sum[3] = {}, span = 0;
        float ms = 0;
        for (size_t w = 0; w < waves; ++w)
            for (int k = 0; k < 3; ++k) { cudaEventElapsedTime(&ms, ev[w * 6 + 2 * k], ev[w * 6 + 2 * k + 1]); sum[k] += ms; }
        cudaEventElapsedTime(&ms, ev[0], ev[waves * 6 - 1]);
        span = ms;
        fprintf(stderr, "[vbz_b200] device timeline: %zu waves in %.2f ms; per wave upload %.3f  kernels %.3f  download %.3f ms\n", waves, span,
                sum[0] / waves, sum[1] / waves, sum[2] / waves);
        for (size_t w = 0; w < waves && w < 8; ++w) {
            float t[6];
            for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], ev[0], ev[w * 6 + k]);
            fprintf(stderr, "[vbz_b200]   wave %zu: upload %.3f-%.3f  kernels %.3f-%.3f  download %.3f-%.3f\n", w, t[0], t[1], t[2], t[3], t[4], t[5]);
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
#endif
    }
    void report(uint32_t reads, uint32_t waves, uint32_t dma_in, uint32_t dma_out, uint32_t xfers)
    {
        if (on)
            fprintf(stderr, "[vbz_b200] %u reads, %u waves, %u+%u DMAs, %u transfer kernels: plan %.2f  gather %.2f  upload-issue %.2f  launch %.2f  wait-kernels %.2f  staged-d2h-issue %.2f  wait-download %.2f  scatter %.2f ms\n",
                    reads, waves, dma_in, dma_out, xfers, acc[0], acc[6], acc[1], acc[2], acc[3], acc[4], acc[5], acc[7]);
    }
};

// uploads the records and launches one transfer kernel
static bool enqueue_xfer(vbz_b200_ctx* ctx, DevBuf& d_records, std::vector<XferRead>& recs, const uint32_t* d_sizes, uint32_t ctas, cudaStream_t stream)
{
    if (recs.empty()) return true;
    uint32_t chunks = 0;
    for (XferRead& r : recs) { r.first_chunk = chunks; chunks += xfer_chunks_of(r.n); }
    const size_t bytes = recs.size() * sizeof(XferRead);
    if (!d_records.ensure(bytes) || !upload_descriptors(ctx, d_records.p, recs.data(), bytes, stream)) return false;
    XferParams p;
    p.reads = d_records.as<XferRead>();
    p.sizes = d_sizes;
    p.n_reads = (uint32_t)recs.size();
    p.n_chunks = chunks;
    const int e = vbzk_xfer(reinterpret_cast<const VbzkXferParams*>(&p), ctas, stream);
    if (e) { cuda_ok((cudaError_t)e, "xfer launch"); return false; }
    ctx->launches += 1;
    return true;
}

static int run_host_batch(vbz_b200_ctx* ctx, HostBatch& b)
{
    const uint32_t m = (uint32_t)b.jobs.size();
    if (m == 0) return 0;
    struct NoStageTiming {  // the stage marks belong to the *_device calls
        vbz_b200_ctx* c; bool keep;
        explicit NoStageTiming(vbz_b200_ctx* ctx) : c(ctx), keep(ctx->stage_timing) { c->stage_timing = false; }
        ~NoStageTiming() { c->stage_timing = keep; }
    } no_timing(ctx);
    HostTrace tr;
    // ---- plan the waves -------------------------------------------------------------------------------------------
    uint64_t total = 0;
    for (const HostJob& h : b.jobs) total += (uint64_t)h.in_slot + h.out_slot;
    WaveSizer sizer(total);
    uint64_t planned = 0;
    const char* force = getenv("VBZ_B200_HOST_PATH");  // "staged": never touch caller buffers from the device (experiments and tests)
    const bool never_direct = force && force[0] == 's';
    const char* up_mode = getenv("VBZ_B200_UPLOAD");   // "kernel" / "dma": how page-locked inputs are fetched (default: by run length)
    const uint32_t up_ctas = env_u32("VBZ_B200_XFER_UP_CTAS", 32), spread_ctas = env_u32("VBZ_B200_XFER_SPREAD_CTAS", 148),
                   down_ctas = env_u32("VBZ_B200_XFER_DOWN_CTAS", 16);
    std::vector<Wave> waves;
    uint64_t max_in = 0, max_out = 0, max_in_staged = 0, max_out_staged = 0, max_in_direct = 0;
    uint32_t max_jobs = 0;
    for (uint32_t k = 0; k < m;) {
        Wave w = {k, k, 0, 0, 0, 0, 0, 0, 0, 0};
        const size_t budget = sizer.next(total - planned);
        uint64_t sum = 0;
        while (w.k1 < m && (w.k1 == w.k0 || sum + b.jobs[w.k1].in_slot + b.jobs[w.k1].out_slot <= budget)) {
            sum += (uint64_t)b.jobs[w.k1].in_slot + b.jobs[w.k1].out_slot;
            ++w.k1;
        }
        planned += sum;
        uint64_t prev_end = 0;
        for (uint32_t q = w.k0; q < w.k1; ++q) {
            HostJob& h = b.jobs[q];
            h.in_direct = !never_direct && h.j.src_size != 0 && host_pinned(b.src[h.read], &h.src_dev);
            h.out_direct = !never_direct && host_pinned(b.dst[h.read], &h.dst_dev);
            if (h.in_direct) {
                if (w.n_in_direct == 0 || (uint64_t)(uintptr_t)b.src[h.read] != prev_end) ++w.in_runs;
                prev_end = (uint64_t)(uintptr_t)b.src[h.read] + h.j.src_size;
                w.in_direct_bytes += h.j.src_size;
                ++w.n_in_direct;
            }
            w.n_out_direct += h.out_direct;
        }
        for (int pass = 0; pass < 2; ++pass) {  // staged reads first, then the direct ones
            for (uint32_t q = w.k0; q < w.k1; ++q) {
                HostJob& h = b.jobs[q];
                if ((int)h.in_direct == pass) { h.j.src_off = w.in_bytes; w.in_bytes += h.in_slot; }
                if ((int)h.out_direct == pass) { h.j.dst_off = w.out_bytes; w.out_bytes += h.out_slot; }
            }
            if (pass == 0) { w.in_staged = w.in_bytes; w.out_staged = w.out_bytes; }
        }
        max_in = std::max(max_in, w.in_bytes);
        max_out = std::max(max_out, w.out_bytes);
        max_in_staged = std::max(max_in_staged, w.in_staged);
        max_out_staged = std::max(max_out_staged, w.out_staged);
        max_in_direct = std::max(max_in_direct, w.in_direct_bytes);
        max_jobs = std::max(max_jobs, w.k1 - w.k0);
        waves.push_back(w);
        k = w.k1;
    }
    const uint32_t W = (uint32_t)waves.size();
    const int S = W > 1 ? vbz_b200_ctx::kWaveSlots : 1;
    const uint32_t D = W > 1 ? (uint32_t)S - 2 : 1;  // waves between enqueueing the kernels and looking at their results
    for (int s = 0; s < S; ++s) {
        vbz_b200_ctx::WaveSlot& sl = ctx->slots[s];
        if (!sl.d_in.ensure(max_in + 16) || !sl.d_out.ensure(max_out + 16) || !sl.d_results.ensure((size_t)max_jobs * 4) ||
            !sl.h_results.ensure((size_t)max_jobs * 4))
            return -1;
        if ((max_in_staged && !sl.h_in.ensure(max_in_staged + 16)) || (max_out_staged && !sl.h_out.ensure(max_out_staged + 16))) return -1;
    }
    cudaStream_t sk = ctx->stream, sh = ctx->s_h2d, sd = ctx->s_d2h, sds = ctx->s_d2hs;
    // the kernels of wave w run on stream w mod n with scratch set w mod n (VBZ_B200_STREAMS=1: everything on one)
    const uint32_t n_streams = W > 2 ? std::min<uint32_t>(env_u32("VBZ_B200_STREAMS", vbz_b200_ctx::kKernelStreams), vbz_b200_ctx::kKernelStreams) : 1;
    struct ScratchTurn {
        vbz_b200_ctx* c;
        explicit ScratchTurn(vbz_b200_ctx* ctx) : c(ctx) {}
        ~ScratchTurn() { c->use_scratch(0); }
    } turn(ctx);
    struct CtaCap {  // the codec shares the SMs with the transfer kernels for the duration of the batch
        vbz_b200_ctx* c;
        CtaCap(vbz_b200_ctx* ctx, uint32_t v) : c(ctx) { c->codec_ctas_per_sm = v; }
        ~CtaCap() { c->codec_ctas_per_sm = 0; }
    } cap(ctx, W > 1 ? env_u32("VBZ_B200_CODEC_CTAS_PER_SM", 3) : 0);
    tr.lap(0);
    std::vector<CopyTask> copies;
    std::vector<ReadJob> wave_jobs;
    std::vector<XferRead> recs;
    uint32_t dma_in = 0, dma_out = 0, xfers = 0;
    int rc = 0;

    // one iteration: upload + launch + download wave w, collect the results of wave w-D, deliver wave w-D-1
    for (uint32_t it = 0; it < W + D + 1 && rc == 0; ++it) {
        if (it < W) {
            const Wave& w = waves[it];
            vbz_b200_ctx::WaveSlot& sl = ctx->slots[it % S];
            uint8_t* d_in = sl.d_in.as<uint8_t>();
            uint8_t* d_out = sl.d_out.as<uint8_t>();
            if (n_streams > 1) { ctx->use_scratch((int)(it % n_streams)); sk = ctx->kstream[it % n_streams]; }
            tr.mark(sh);
            if (w.in_staged) {
                copies.clear();
                for (uint32_t q = w.k0; q < w.k1; ++q) {
                    const HostJob& h = b.jobs[q];
                    if (!h.in_direct && h.j.src_size) copies.push_back({sl.h_in.as<uint8_t>() + h.j.src_off, b.src[h.read], h.j.src_size});
                }
                CopyPool::instance().run(copies);
                tr.lap(6);
                if (!cuda_ok(cudaMemcpyAsync(d_in, sl.h_in.p, w.in_staged, cudaMemcpyHostToDevice, sh), "H2D input")) { rc = -1; break; }
                ++dma_in;
            }
            if (w.n_in_direct) {
                // long host-contiguous runs: a DMA per run into d_pack, then device-to-device into the slots; else straight from the host
                bool by_dma = w.in_direct_bytes / w.in_runs >= (2u << 20);
                if (up_mode) by_dma = up_mode[0] == 'd';
                if (by_dma && !sl.d_pack.ensure(max_in_direct + 16)) { rc = -1; break; }
                uint8_t* d_pack = sl.d_pack.as<uint8_t>();
                recs.clear();
                uint64_t run_src = 0, run_len = 0, run_pack = 0, packed = 0;
                for (uint32_t q = w.k0; q < w.k1 && rc == 0; ++q) {
                    const HostJob& h = b.jobs[q];
                    if (!h.in_direct) continue;
                    uint64_t from = h.src_dev;
                    if (by_dma) {
                        const uint64_t host_at = (uint64_t)(uintptr_t)b.src[h.read];
                        if (run_len && host_at != run_src + run_len) {
                            if (!cuda_ok(cudaMemcpyAsync(d_pack + run_pack, (const void*)(uintptr_t)run_src, run_len, cudaMemcpyHostToDevice, sh), "H2D input")) rc = -1;
                            ++dma_in;
                            run_len = 0;
                        }
                        if (run_len == 0) { run_src = host_at; run_pack = packed; }
                        from = (uint64_t)(uintptr_t)(d_pack + packed);
                        run_len += h.j.src_size;
                        packed += h.j.src_size;
                    }
                    recs.push_back({from, (uint64_t)(uintptr_t)(d_in + h.j.src_off), h.j.src_size, 0, 0, 0});
                }
                if (rc == 0 && by_dma && run_len) {
                    if (!cuda_ok(cudaMemcpyAsync(d_pack + run_pack, (const void*)(uintptr_t)run_src, run_len, cudaMemcpyHostToDevice, sh), "H2D input")) rc = -1;
                    ++dma_in;
                }
                if (rc == 0 && !enqueue_xfer(ctx, sl.d_xin, recs, nullptr, by_dma ? spread_ctas : up_ctas, sh)) rc = -1;
                ++xfers;
            }
            if (rc != 0) break;
            tr.mark(sh);
            tr.lap(1);
            wave_jobs.resize(w.k1 - w.k0);
            for (uint32_t q = w.k0; q < w.k1; ++q) wave_jobs[q - w.k0] = b.jobs[q].j;
            // the kernels wait for the upload, and for the downloads that last used this slot's output buffer
            if (!cuda_ok(cudaEventRecord(sl.ev_in, sh), "cudaEventRecord") || !cuda_ok(cudaStreamWaitEvent(sk, sl.ev_in, 0), "cudaStreamWaitEvent") ||
                (it >= (uint32_t)S && (!cuda_ok(cudaStreamWaitEvent(sk, sl.ev_out, 0), "cudaStreamWaitEvent") ||
                                       !cuda_ok(cudaStreamWaitEvent(sk, sl.ev_outs, 0), "cudaStreamWaitEvent")))) { rc = -1; break; }
            tr.mark(sk);
            const int e = b.compress ? compress_device_impl(ctx, d_in, d_out, wave_jobs, sl.d_results.as<uint32_t>(), b.kind, b.sized, sk)
                                     : decompress_device_impl(ctx, d_in, d_out, wave_jobs, sl.d_results.as<uint32_t>(), b.kind, b.sized, sk);
            if (e != 0) { rc = -1; break; }
            if (const int ce = vbzk_copy_words(sl.h_results.p, sl.d_results.p, (size_t)(w.k1 - w.k0) * 4, sk)) { cuda_ok((cudaError_t)ce, "results to host"); rc = -1; break; }
            ctx->launches += 1;
            if (!cuda_ok(cudaEventRecord(sl.ev_k, sk), "cudaEventRecord")) { rc = -1; break; }
            tr.mark(sk);
            // download in place: the transfer kernel takes each read's size from the device-side result
            if (!cuda_ok(cudaStreamWaitEvent(sd, sl.ev_k, 0), "cudaStreamWaitEvent")) { rc = -1; break; }
            tr.mark(sd);
            if (w.n_out_direct) {
                recs.clear();
                for (uint32_t q = w.k0; q < w.k1; ++q) {
                    const HostJob& h = b.jobs[q];
                    if (!h.out_direct) continue;
                    const uint32_t most = b.compress ? std::min<uint32_t>(h.out_slot, b.dst_cap[h.read]) : h.j.dst_size;
                    recs.push_back({(uint64_t)(uintptr_t)(d_out + h.j.dst_off), h.dst_dev, most, q - w.k0, 0, 0});
                }
                if (!enqueue_xfer(ctx, sl.d_xout, recs, sl.d_results.as<uint32_t>(), down_ctas, sd)) { rc = -1; break; }
                ++xfers;
            }
            if (!cuda_ok(cudaEventRecord(sl.ev_out, sd), "cudaEventRecord")) { rc = -1; break; }
            tr.mark(sd);
            tr.lap(2);
        }
        if (it >= D && it - D < W) {
            const Wave& w = waves[it - D];
            vbz_b200_ctx::WaveSlot& sl = ctx->slots[(it - D) % S];
            if (!cuda_ok(cudaEventSynchronize(sl.ev_k), "batch kernels")) { rc = -1; break; }
            tr.lap(3);
            const uint32_t* dres = sl.h_results.as<uint32_t>();
            uint64_t staged_end = 0;  // the staged part of the output buffer, up to the last byte a read produced
            for (uint32_t q = w.k0; q < w.k1; ++q) {
                const HostJob& h = b.jobs[q];
                uint32_t r = dres[q - w.k0];
                if (!is_error(r) && r > b.dst_cap[h.read]) r = b.sized ? kErrDestinationSize + 4 : kErrDestinationSize;  // (compress with a tight device slot)
                b.results[h.read] = r;
                if (!is_error(r) && r != 0 && !h.out_direct) staged_end = std::max<uint64_t>(staged_end, h.j.dst_off + r);
            }
            if (staged_end) {
                if (!cuda_ok(cudaMemcpyAsync(sl.h_out.p, sl.d_out.p, staged_end, cudaMemcpyDeviceToHost, sds), "D2H output")) { rc = -1; break; }
                ++dma_out;
            }
            if (!cuda_ok(cudaEventRecord(sl.ev_outs, sds), "cudaEventRecord")) { rc = -1; break; }
            tr.lap(4);
        }
        if (it >= D + 1) {
            const Wave& w = waves[it - D - 1];
            vbz_b200_ctx::WaveSlot& sl = ctx->slots[(it - D - 1) % S];
            if (!cuda_ok(cudaEventSynchronize(sl.ev_out), "batch download") || !cuda_ok(cudaEventSynchronize(sl.ev_outs), "batch download")) { rc = -1; break; }
            tr.lap(5);
            if (w.out_staged) {
                copies.clear();
                for (uint32_t q = w.k0; q < w.k1; ++q) {
                    const HostJob& h = b.jobs[q];
                    const uint32_t r = b.results[h.read];
                    if (!h.out_direct && !is_error(r) && r > 0) copies.push_back({b.dst[h.read], sl.h_out.as<uint8_t>() + h.j.dst_off, r});
                }
                CopyPool::instance().run(copies);
                tr.lap(7);
            }
        }
    }
    tr.report(m, W, dma_in, dma_out, xfers);
    if (rc == 0) tr.timeline();
    if (rc != 0) {  // leave nothing in flight that still points at the caller's buffers
        cudaStreamSynchronize(sh); for (cudaStream_t ks : ctx->kstream) cudaStreamSynchronize(ks); cudaStreamSynchronize(sd); cudaStreamSynchronize(sds);
    }
    return rc;
}

}  // namespace vbz

extern "C" int vbz_b200_compress_batch(vbz_b200_ctx* ctx, const void* const* src, const vbz_size_t* src_size, void* const* dst,
                                       const vbz_size_t* dst_cap, vbz_size_t* results, uint32_t n, const CompressionOptions* o, int sized_i)
{
    if (!ctx || !o) { set_last_error("null context or options"); return -1; }
    if (n == 0) return 0;
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return -1;
    const bool sized = sized_i != 0;
    const uint32_t hdr = sized ? 4u : 0u;
    HostBatch b;
    b.compress = true; b.sized = sized; b.kind = classify_options(o);
    b.src = src; b.src_size = src_size; b.dst = dst; b.dst_cap = dst_cap; b.results = results;
    const Stage1Kind& kind = b.kind;

    // per-read checks in the reference's order; plain copies are done right here
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t code = 0;
        if (compress_precheck(kind, src_size[i], dst_cap[i], sized, &code)) { results[i] = code; continue; }
        if (kind.passthrough) {  // level 0 and integer_size 0: a copy (ref vbz/vbz.cpp:130-133)
            if (sized) memcpy(dst[i], &src_size[i], 4);
            if (src_size[i]) memcpy((uint8_t*)dst[i] + hdr, src[i], src_size[i]);
            results[i] = src_size[i] + hdr;
            continue;
        }
        HostJob h;
        h.read = i;
        h.j.src_size = src_size[i];
        h.in_slot = (uint32_t)std::min<uint64_t>(align16(src_size[i]), 0xFFFFFFF0u);
        uint64_t produced = kind.isz != 0 ? stage1_tight_bound(src_size[i] / (unsigned)kind.isz, kind) : src_size[i];
        if (kind.zstd_level != 0) produced = zstd_enc_worst_case(produced);
        const uint64_t slot = align16(hdr + produced);
        if (slot > 0xFFFFFFF0ull) { results[i] = sized ? kErrOutOfMemory + 4 : kErrOutOfMemory; continue; }
        // with a zstd stage the caller's capacity decides between a frame and VBZ_ZSTD_ERROR (vbz.cpp:194-204)
        h.j.dst_size = (uint32_t)(kind.zstd_level != 0 ? std::min<uint64_t>(slot, dst_cap[i]) : slot);
        h.out_slot = (uint32_t)slot;
        b.jobs.push_back(h);
    }
    if (b.jobs.empty()) return 0;
    return run_host_batch(ctx, b);
}

extern "C" int vbz_b200_decompress_batch(vbz_b200_ctx* ctx, const void* const* src, const vbz_size_t* src_size, void* const* dst,
                                         const vbz_size_t* dst_size, vbz_size_t* results, uint32_t n, const CompressionOptions* o, int sized_i)
{
    if (!ctx || !o) { set_last_error("null context or options"); return -1; }
    if (n == 0) return 0;
    DeviceGuard guard(ctx->device);
    if (!guard.ok) return -1;
    const bool sized = sized_i != 0;
    const uint32_t hdr = sized ? 4u : 0u;
    HostBatch b;
    b.compress = false; b.sized = sized; b.kind = classify_options(o);
    b.src = src; b.src_size = src_size; b.dst = dst; b.dst_cap = dst_size; b.results = results;
    const Stage1Kind& kind = b.kind;

    for (uint32_t i = 0; i < n; ++i) {
        if (kind.bad_isz) { results[i] = kErrIntegerSize; continue; }
        uint32_t exact = dst_size[i];
        if (sized) {  // ref vbz/vbz.cpp:345-366: header present?  capacity large enough?
            if (src_size[i] < 4) { results[i] = kErrInputSize; continue; }
            uint32_t orig;
            memcpy(&orig, src[i], 4);
            if (dst_size[i] < orig) { results[i] = kErrDestinationSize; continue; }
            exact = orig;
        }
        uint32_t code = 0;
        if (decompress_precheck(kind, src_size[i] - hdr, exact, &code)) { results[i] = code; continue; }
        if (kind.passthrough) {
            if (src_size[i] - hdr) memcpy(dst[i], (const uint8_t*)src[i] + hdr, src_size[i] - hdr);
            results[i] = src_size[i] - hdr;
            continue;
        }
        HostJob h;
        if (kind.zstd_level != 0) {  // ref vbz/vbz.cpp:234-262: the frame must announce its content size
            const uint64_t fcs = zstd_frame_content_size((const uint8_t*)src[i] + hdr, src_size[i] - hdr);
            if (fcs == kZstdContentSizeUnknown || fcs == kZstdContentSizeError) { results[i] = kErrZstd; continue; }
            if (kind.isz == 0) {
                if (fcs > exact) { results[i] = kErrDestinationSize; continue; }
            } else {
                // the reference mallocs the content size here (vbz.cpp:252).  The slot a read gets on the device is capped by what stage 1
                // can consume for its `exact` output bytes (decode_mid_cap): a header that announces more cannot starve the batch of
                // memory - the frame then fails to decode into its slot and this read alone reports the error
                if (fcs > (1ull << 31)) { results[i] = kErrOutOfMemory; continue; }
                h.j.mid_cap = fcs ? fcs : 1;
                if (kind.bad_version) h.j.post_code = kErrVersion;
                else if (exact % (unsigned)kind.isz != 0) h.j.post_code = kErrDestinationSize;
            }
        }
        if (src_size[i] > 0xFFFFFFE0u || exact > 0xFFFFFFE0u) { results[i] = kErrOutOfMemory; continue; }
        h.read = i;
        h.j.src_size = src_size[i];
        h.j.dst_size = exact;
        h.in_slot = (uint32_t)align16(src_size[i]);
        h.out_slot = (uint32_t)align16(exact);
        b.jobs.push_back(h);
    }
    if (b.jobs.empty()) return 0;
    return run_host_batch(ctx, b);
}
