// csrc/host/c_api.cpp - the reference's eight C entry points (include/vbz.h) as thin wrappers over a
// batch of one read (csrc/host/batch.cpp).  Callable concurrently from many host threads: contexts
// (stream + scratch) are pooled and handed out one per call.
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../../include/vbz_b200.h"
#include "context.h"
#include "options.h"

namespace {

struct Pool {
    std::mutex mu;
    std::vector<vbz_b200_ctx*> free_list;
    ~Pool()
    {
        // contexts are intentionally leaked at process exit: the CUDA runtime may already be gone
    }
};
Pool& pool()
{
    static Pool* p = new Pool();
    return *p;
}

int default_device()
{
    const char* e = getenv("VBZ_B200_DEVICE");
    return e ? atoi(e) : 0;
}

// VBZ_B200_QUIET=1: a library should not write to its host's stderr unasked; the text stays available through vbz_b200_last_error()
bool quiet()
{
    const char* e = getenv("VBZ_B200_QUIET");
    return e && *e && *e != '0';
}

struct Lease {
    vbz_b200_ctx* ctx = nullptr;
    Lease()
    {
        {
            std::lock_guard<std::mutex> g(pool().mu);
            if (!pool().free_list.empty()) { ctx = pool().free_list.back(); pool().free_list.pop_back(); }
        }
        if (!ctx) ctx = vbz_b200_create(default_device());
        if (!ctx) {
            static bool warned = false;
            if (!warned && !quiet()) {
                warned = true;
                fprintf(stderr, "libvbz_b200: cannot run on a GPU (%s). This library has no CPU implementation.\n", vbz_b200_last_error());
            }
        }
    }
    ~Lease()
    {
        if (ctx) { std::lock_guard<std::mutex> g(pool().mu); pool().free_list.push_back(ctx); }
    }
};

vbz_size_t run_one(bool compress, bool sized, const void* src, vbz_size_t n, void* dst, vbz_size_t cap, const CompressionOptions* o)
{
    vbz::Stage1Kind kind = vbz::classify_options(o);
    if (kind.bad_isz) return VBZ_INTEGER_SIZE_ERROR;  // needs no device
    Lease lease;
    if (!lease.ctx) return VBZ_OUT_OF_MEMORY_ERROR;
    vbz_size_t result = 0;
    const void* srcs[1] = {src};
    void* dsts[1] = {dst};
    int rc = compress ? vbz_b200_compress_batch(lease.ctx, srcs, &n, dsts, &cap, &result, 1, o, sized)
                      : vbz_b200_decompress_batch(lease.ctx, srcs, &n, dsts, &cap, &result, 1, o, sized);
    if (rc != 0) {
        if (!quiet()) fprintf(stderr, "libvbz_b200: %s\n", vbz_b200_last_error());
        return VBZ_OUT_OF_MEMORY_ERROR;
    }
    return result;
}

}  // namespace

extern "C" {

bool vbz_is_error(vbz_size_t v) { return v >= VBZ_FIRST_ERROR; }

char const* vbz_error_string(vbz_size_t v)
{
    switch (v) {
    case VBZ_ZSTD_ERROR: return "VBZ_ZSTD_ERROR";
    case VBZ_INPUT_SIZE_ERROR: return "VBZ_INPUT_SIZE_ERROR";
    case VBZ_INTEGER_SIZE_ERROR: return "VBZ_INTEGER_SIZE_ERROR";
    case VBZ_DESTINATION_SIZE_ERROR: return "VBZ_DESTINATION_SIZE_ERROR";
    case VBZ_STREAMVBYTE_STREAM_ERROR: return "VBZ_STREAMVBYTE_STREAM_ERROR";
    case VBZ_VERSION_ERROR: return "VBZ_VERSION_ERROR";
    case VBZ_OUT_OF_MEMORY_ERROR: return "VBZ_OUT_OF_MEMORY_ERROR";
    default: return "VBZ_UNKNOWN_ERROR";
    }
}

vbz_size_t vbz_max_compressed_size(vbz_size_t source_size, CompressionOptions const* options)
{
    return vbz::max_compressed_size(source_size, options);
}

vbz_size_t vbz_compress(void const* source, vbz_size_t source_size, void* destination, vbz_size_t destination_capacity,
                        CompressionOptions const* options)
{
    return run_one(true, false, source, source_size, destination, destination_capacity, options);
}

vbz_size_t vbz_decompress(void const* source, vbz_size_t source_size, void* destination, vbz_size_t destination_size,
                          CompressionOptions const* options)
{
    return run_one(false, false, source, source_size, destination, destination_size, options);
}

vbz_size_t vbz_compress_sized(void const* source, vbz_size_t source_size, void* destination, vbz_size_t destination_capacity,
                              CompressionOptions const* options)
{
    return run_one(true, true, source, source_size, destination, destination_capacity, options);
}

vbz_size_t vbz_decompress_sized(void const* source, vbz_size_t source_size, void* destination, vbz_size_t destination_capacity,
                                CompressionOptions const* options)
{
    return run_one(false, true, source, source_size, destination, destination_capacity, options);
}

vbz_size_t vbz_decompressed_size(void const* source, vbz_size_t source_size, CompressionOptions const* options)
{
    const vbz::Stage1Kind kind = vbz::classify_options(options);
    if (kind.bad_isz) return VBZ_INTEGER_SIZE_ERROR;
    if (source_size < 4) return VBZ_INPUT_SIZE_ERROR;
    vbz_size_t v;
    __builtin_memcpy(&v, source, 4);
    return v;
}

}  // extern "C"
