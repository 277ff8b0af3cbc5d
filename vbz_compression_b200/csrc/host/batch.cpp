// csrc/host/batch.cpp - the C++ host layer: turns a batch of reads into descriptor tables, stages
// buffers, and drives the CUDA kernels through the C-ABI shim (csrc/vbzk.h).
//
// Per-read semantics follow the reference call for call:
//   compress    vbz_compress / vbz_compress_sized       (ref vbz/vbz.cpp:116-208, 302-330)
//   decompress  vbz_decompress / vbz_decompress_sized   (ref vbz/vbz.cpp:210-300, 332-367)
// including the order in which option / size errors are detected.  All sample arithmetic happens in
// the kernels; the host only validates sizes, lays buffers out and moves bytes.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/vbz_b200.h"
#include "../svb_types.h"
#include "../vbzk.h"
#include "context.h"
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
    // the pinned staging may still be feeding the previous batch's upload
    if (!cuda_ok(cudaEventSynchronize(ctx->upload_done), "cudaEventSynchronize")) return -1;
    if (!ctx->h_reads.ensure((size_t)n * sizeof(SvbRead))) return -1;
    memcpy(ctx->h_reads.p, reads.data(), (size_t)n * sizeof(SvbRead));
    if (!cuda_ok(cudaMemcpyAsync(ctx->d_reads.p, ctx->h_reads.p, (size_t)n * sizeof(SvbRead), cudaMemcpyHostToDevice, stream), "upload descriptors"))
        return -1;
    if (!cuda_ok(cudaEventRecord(ctx->upload_done, stream), "cudaEventRecord")) return -1;
    if (!cuda_ok(cudaMemsetAsync(ctx->d_status.p, 0, (status_words + 1) * 8, stream), "reset look-back state")) return -1;

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
static int enqueue_zstd_decode(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const std::vector<ZstdRead>& reads, uint32_t* d_results,
                               cudaStream_t stream)
{
    const uint32_t n = (uint32_t)reads.size();
    if (n == 0) return 0;
    const size_t bytes = (size_t)n * sizeof(ZstdRead);
    if (!ctx->d_zreads.ensure(bytes) || !ctx->d_zcounter.ensure(256) || !ctx->d_zlit.ensure(vbzk_zstd_decode_scratch_bytes())) return -1;
    if (!cuda_ok(cudaEventSynchronize(ctx->upload_done), "cudaEventSynchronize")) return -1;
    if (!ctx->h_zreads.ensure(bytes)) return -1;
    memcpy(ctx->h_zreads.p, reads.data(), bytes);
    if (!cuda_ok(cudaMemcpyAsync(ctx->d_zreads.p, ctx->h_zreads.p, bytes, cudaMemcpyHostToDevice, stream), "upload frame descriptors") ||
        !cuda_ok(cudaEventRecord(ctx->upload_done, stream), "cudaEventRecord") ||
        !cuda_ok(cudaMemsetAsync(ctx->d_zcounter.p, 0, 256, stream), "reset work counter"))
        return -1;
    ZstdDecParams p;
    p.src = d_src;
    p.dst = d_dst;
    p.reads = ctx->d_zreads.as<ZstdRead>();
    p.results = d_results;
    p.lit_scratch = ctx->d_zlit.as<uint8_t>();
    p.counter = ctx->d_zcounter.as<uint32_t>();
    p.n_reads = n;
    const int e = vbzk_zstd_decode(reinterpret_cast<const VbzkZstdDecParams*>(&p), stream);
    if (e) { cuda_ok((cudaError_t)e, "zstd_decode launch"); return -1; }
    ctx->launches += 1;
    return 0;
}

// ---- stage 2 (zstd), encode: content -> frames ----------------------------------------------------------------------
// reads[i].first_block is set here.
static int enqueue_zstd_encode(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, std::vector<ZstdEncRead>& reads, const uint32_t* d_src_sizes,
                               uint32_t* d_results, bool sized, cudaStream_t stream)
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
    if (!ctx->d_zreads.ensure(bytes) || !ctx->d_zblocks.ensure((size_t)n_blocks * 4) || !ctx->d_zstatus.ensure((size_t)n_blocks * 8)) return -1;
    if (!cuda_ok(cudaEventSynchronize(ctx->upload_done), "cudaEventSynchronize")) return -1;
    if (!ctx->h_zreads.ensure(bytes)) return -1;
    memcpy(ctx->h_zreads.p, reads.data(), bytes);
    if (!cuda_ok(cudaMemcpyAsync(ctx->d_zreads.p, ctx->h_zreads.p, bytes, cudaMemcpyHostToDevice, stream), "upload frame descriptors") ||
        !cuda_ok(cudaEventRecord(ctx->upload_done, stream), "cudaEventRecord") ||
        !cuda_ok(cudaMemsetAsync(ctx->d_zstatus.p, 0, (size_t)n_blocks * 8, stream), "reset look-back state"))
        return -1;
    int e = vbzk_zstd_enc_fill_blocks(ctx->d_zreads.as<VbzkZstdEncRead>(), n, ctx->d_zblocks.as<uint32_t>(), stream);
    if (e) { cuda_ok((cudaError_t)e, "zstd_enc_fill_blocks launch"); return -1; }
    ZstdEncParams p;
    p.src = d_src;
    p.dst = d_dst;
    p.reads = ctx->d_zreads.as<ZstdEncRead>();
    p.block_read = ctx->d_zblocks.as<uint32_t>();
    p.src_sizes = d_src_sizes;
    p.status = ctx->d_zstatus.as<uint64_t>();
    p.results = d_results;
    p.hdr_bytes = sized ? 4u : 0u;
    p.err_add = sized ? 4u : 0u;
    p.n_reads = n;
    p.n_blocks = n_blocks;
    e = vbzk_zstd_encode(reinterpret_cast<const VbzkZstdEncParams*>(&p), stream);
    if (e) { cuda_ok((cudaError_t)e, "zstd_encode launch"); return -1; }
    ctx->launches += 2;
    return 0;
}

static bool device_path_available(const Stage1Kind& kind, bool decode = false)
{
    (void)decode;
    if (kind.half) {
        set_last_error("this option combination is not implemented on the device yet");
        return false;
    }
    return true;
}

// ---- compress / decompress, device-resident ------------------------------------------------------------------
static int compress_device_impl(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const std::vector<ReadJob>& jobs,
                                uint32_t* d_results, const Stage1Kind& kind, bool sized, cudaStream_t stream)
{
    if (!device_path_available(kind)) return -1;
    const uint32_t n = (uint32_t)jobs.size();
    const bool zstd = kind.zstd_level != 0;
    if (!zstd && kind.isz == 0) { set_last_error("plain copies (integer_size 0, level 0) are not a device operation"); return -1; }
    const uint32_t blk = vbzk_zstd_enc_block_bytes();
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
        z.n_blocks = (uint32_t)std::max<uint64_t>(1, (z.split + blk - 1) / blk + (largest - z.split + blk - 1) / blk);
    }
    return enqueue_zstd_encode(ctx, kind.isz != 0 ? ctx->d_mid.as<uint8_t>() : d_src, d_dst, zr, kind.isz != 0 ? ctx->d_sizes.as<uint32_t>() : nullptr,
                               d_results, sized, stream);
}

static int decompress_device_impl(vbz_b200_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, const std::vector<ReadJob>& jobs,
                                  uint32_t* d_results, const Stage1Kind& kind, bool sized, cudaStream_t stream)
{
    if (!device_path_available(kind, true)) return -1;
    if (kind.zstd_level == 0 && kind.isz == 0) { set_last_error("plain copies (integer_size 0, level 0) are not a device operation"); return -1; }
    const uint32_t hdr = sized ? 4u : 0u;
    const uint32_t n = (uint32_t)jobs.size();
    const uint8_t* s1_src = d_src;
    const uint32_t* s1_sizes = nullptr;
    std::vector<uint64_t> mid_off;
    if (kind.zstd_level != 0) {
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
                uint64_t cap = j.mid_cap;
                if (cap == 0) cap = stage1_tight_bound(j.dst_size / (unsigned)kind.isz, kind) + 16;  // device-resident frames: the host cannot read the header
                cap = std::min<uint64_t>(cap, 0xFFFFFFF0u);
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
        if (enqueue_zstd_decode(ctx, d_src, zdst, zr, zres, stream) != 0) return -1;
        if (kind.isz == 0) return 0;
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

}  // namespace vbz

using namespace vbz;

// ======================================================================================================
// C ABI (include/vbz_b200.h)
// ======================================================================================================
vbz_b200_ctx::~vbz_b200_ctx()
{
    d_reads.release(); d_tiles.release(); d_status.release(); d_wtoff.release(); d_issue.release(); d_results.release(); d_sizes.release();
    h_reads.release(); h_results.release();
    d_in.release(); d_out.release(); d_mid.release();
    h_in.release(); h_out.release();
    if (upload_done) cudaEventDestroy(upload_done);
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
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
    vbz_b200_ctx* c = new vbz_b200_ctx();
    c->device = device;
    if (!cuda_ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !cuda_ok(cudaEventCreateWithFlags(&c->upload_done, cudaEventDisableTiming), "cudaEventCreate")) {
        delete c;
        return nullptr;
    }
    return c;
}

extern "C" void vbz_b200_destroy(vbz_b200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
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
    if (!cuda_ok(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
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
    return compress_device_impl(ctx, (const uint8_t*)d_src, (uint8_t*)d_dst, jobs, d_results, kind, sized != 0,
                                (cudaStream_t)stream);
}

extern "C" int vbz_b200_decompress_device(vbz_b200_ctx* ctx, const void* d_src, const uint64_t* src_off, const vbz_size_t* src_size,
                                          void* d_dst, const uint64_t* dst_off, const vbz_size_t* dst_size, vbz_size_t* d_results,
                                          uint32_t n, const CompressionOptions* options, int sized, void* stream)
{
    if (!ctx || !options) { set_last_error("null context or options"); return -1; }
    if (!cuda_ok(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
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
    return decompress_device_impl(ctx, (const uint8_t*)d_src, (uint8_t*)d_dst, jobs, d_results, kind, sized != 0,
                                  (cudaStream_t)stream);
}

// ---- host-pointer batches -------------------------------------------------------------------------------------
extern "C" int vbz_b200_compress_batch(vbz_b200_ctx* ctx, const void* const* src, const vbz_size_t* src_size, void* const* dst,
                                       const vbz_size_t* dst_cap, vbz_size_t* results, uint32_t n, const CompressionOptions* o, int sized_i)
{
    if (!ctx || !o) { set_last_error("null context or options"); return -1; }
    if (n == 0) return 0;
    if (!cuda_ok(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
    const bool sized = sized_i != 0;
    const uint32_t hdr = sized ? 4u : 0u;
    const Stage1Kind kind = classify_options(o);

    // pass 1: per-read checks in the reference's order; plain copies are done right here
    std::vector<ReadJob> jobs;
    std::vector<uint32_t> job_read;
    uint64_t in_total = 0, out_total = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t code = 0;
        if (compress_precheck(kind, src_size[i], dst_cap[i], sized, &code)) { results[i] = code; continue; }
        if (kind.passthrough) {  // level 0 and integer_size 0: a copy (ref vbz/vbz.cpp:130-133)
            if (sized) memcpy(dst[i], &src_size[i], 4);
            if (src_size[i]) memcpy((uint8_t*)dst[i] + hdr, src[i], src_size[i]);
            results[i] = src_size[i] + hdr;
            continue;
        }
        ReadJob j;
        j.src_off = in_total;
        j.dst_off = out_total;
        j.src_size = src_size[i];
        in_total += align16(src_size[i]);
        uint64_t produced = kind.isz != 0 ? stage1_tight_bound(src_size[i] / (unsigned)kind.isz, kind) : src_size[i];
        if (kind.zstd_level != 0) produced = zstd_enc_worst_case(produced);
        const uint64_t slot = align16(hdr + produced);
        // with a zstd stage the caller's capacity decides between a frame and VBZ_ZSTD_ERROR (vbz.cpp:194-204)
        j.dst_size = (uint32_t)std::min<uint64_t>(kind.zstd_level != 0 ? std::min<uint64_t>(slot, dst_cap[i]) : slot, 0xFFFFFFF0u);
        out_total += slot;
        jobs.push_back(j);
        job_read.push_back(i);
    }
    if (jobs.empty()) return 0;
    if (!device_path_available(kind)) return -1;
    const uint32_t m = (uint32_t)jobs.size();

    // pass 2: stage inputs, run, fetch results
    if (!ctx->h_in.ensure(in_total + 16) || !ctx->d_in.ensure(in_total + 16) || !ctx->d_out.ensure(out_total + 16) ||
        !ctx->d_results.ensure((size_t)m * 4) || !ctx->h_results.ensure((size_t)m * 4))
        return -1;
    for (uint32_t k = 0; k < m; ++k)
        if (jobs[k].src_size) memcpy(ctx->h_in.as<uint8_t>() + jobs[k].src_off, src[job_read[k]], jobs[k].src_size);
    cudaStream_t s = ctx->stream;
    if (!cuda_ok(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in_total, cudaMemcpyHostToDevice, s), "H2D input")) return -1;
    if (compress_device_impl(ctx, ctx->d_in.as<uint8_t>(), ctx->d_out.as<uint8_t>(), jobs, ctx->d_results.as<uint32_t>(), kind, sized, s) != 0)
        return -1;
    if (!cuda_ok(cudaMemcpyAsync(ctx->h_results.p, ctx->d_results.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s), "D2H results") ||
        !cuda_ok(cudaStreamSynchronize(s), "compress batch"))
        return -1;
    // pass 3: bring back exactly the bytes each read produced
    const uint32_t* dres = ctx->h_results.as<uint32_t>();
    if (!ctx->h_out.ensure(out_total + 16)) return -1;
    for (uint32_t k = 0; k < m; ++k) {
        results[job_read[k]] = dres[k];
        if (!is_error(dres[k]) && dres[k] > 0 &&
            !cuda_ok(cudaMemcpyAsync(ctx->h_out.as<uint8_t>() + jobs[k].dst_off, ctx->d_out.as<uint8_t>() + jobs[k].dst_off, dres[k],
                                     cudaMemcpyDeviceToHost, s), "D2H output"))
            return -1;
    }
    if (!cuda_ok(cudaStreamSynchronize(s), "compress batch (copy back)")) return -1;
    for (uint32_t k = 0; k < m; ++k) {
        const uint32_t i = job_read[k];
        if (!is_error(dres[k]) && dres[k] > 0) {
            if (dres[k] > dst_cap[i]) { results[i] = sized ? kErrDestinationSize + 4 : kErrDestinationSize; continue; }
            memcpy(dst[i], ctx->h_out.as<uint8_t>() + jobs[k].dst_off, dres[k]);
        }
    }
    return 0;
}

extern "C" int vbz_b200_decompress_batch(vbz_b200_ctx* ctx, const void* const* src, const vbz_size_t* src_size, void* const* dst,
                                         const vbz_size_t* dst_size, vbz_size_t* results, uint32_t n, const CompressionOptions* o, int sized_i)
{
    if (!ctx || !o) { set_last_error("null context or options"); return -1; }
    if (n == 0) return 0;
    if (!cuda_ok(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
    const bool sized = sized_i != 0;
    const uint32_t hdr = sized ? 4u : 0u;
    const Stage1Kind kind = classify_options(o);

    std::vector<ReadJob> jobs;
    std::vector<uint32_t> job_read;
    uint64_t in_total = 0, out_total = 0;
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
        ReadJob j;
        if (kind.zstd_level != 0) {  // ref vbz/vbz.cpp:234-262: the frame must announce its content size
            const uint64_t fcs = zstd_frame_content_size((const uint8_t*)src[i] + hdr, src_size[i] - hdr);
            if (fcs == kZstdContentSizeUnknown || fcs == kZstdContentSizeError) { results[i] = kErrZstd; continue; }
            if (kind.isz == 0) {
                if (fcs > exact) { results[i] = kErrDestinationSize; continue; }
            } else {
                if (fcs > (1ull << 31)) { results[i] = kErrOutOfMemory; continue; }  // the reference mallocs the content size here
                j.mid_cap = fcs ? fcs : 1;
                if (kind.bad_version) j.post_code = kErrVersion;
                else if (exact % (unsigned)kind.isz != 0) j.post_code = kErrDestinationSize;
            }
        }
        j.src_off = in_total;
        j.dst_off = out_total;
        j.src_size = src_size[i];
        j.dst_size = exact;
        in_total += align16(src_size[i]);
        out_total += align16(exact);
        jobs.push_back(j);
        job_read.push_back(i);
    }
    if (jobs.empty()) return 0;
    if (!device_path_available(kind, true)) return -1;
    const uint32_t m = (uint32_t)jobs.size();

    if (!ctx->h_in.ensure(in_total + 16) || !ctx->d_in.ensure(in_total + 16) || !ctx->d_out.ensure(out_total + 16) ||
        !ctx->h_out.ensure(out_total + 16) || !ctx->d_results.ensure((size_t)m * 4) || !ctx->h_results.ensure((size_t)m * 4))
        return -1;
    for (uint32_t k = 0; k < m; ++k)
        if (jobs[k].src_size) memcpy(ctx->h_in.as<uint8_t>() + jobs[k].src_off, src[job_read[k]], jobs[k].src_size);
    cudaStream_t s = ctx->stream;
    if (!cuda_ok(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in_total, cudaMemcpyHostToDevice, s), "H2D input")) return -1;
    if (decompress_device_impl(ctx, ctx->d_in.as<uint8_t>(), ctx->d_out.as<uint8_t>(), jobs, ctx->d_results.as<uint32_t>(), kind, sized, s) != 0)
        return -1;
    if (!cuda_ok(cudaMemcpyAsync(ctx->h_results.p, ctx->d_results.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s), "D2H results") ||
        !cuda_ok(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out_total, cudaMemcpyDeviceToHost, s), "D2H output") ||
        !cuda_ok(cudaStreamSynchronize(s), "decompress batch"))
        return -1;
    const uint32_t* dres = ctx->h_results.as<uint32_t>();
    for (uint32_t k = 0; k < m; ++k) {
        const uint32_t i = job_read[k];
        results[i] = dres[k];
        if (!is_error(dres[k]) && dres[k] > 0) memcpy(dst[i], ctx->h_out.as<uint8_t>() + jobs[k].dst_off, dres[k]);
    }
    return 0;
}
