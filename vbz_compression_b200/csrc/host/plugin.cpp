// csrc/host/plugin.cpp - HDF5 filter 32020 (include/vbz_plugin.h).
// Behaviour follows reference vbz_plugin/vbz_plugin.cpp:97-229 call for call: cd_values ->
// CompressionOptions, reverse = vbz_decompressed_size + vbz_decompress_sized, forward =
// vbz_max_compressed_size + vbz_compress_sized, buffers swapped with malloc/free, 0 on failure with
// the caller's pointers untouched.
#include <cstdio>
#include <cstdlib>
#include <limits>

#include "../../../include/vbz_b200.h"
#include "../../../include/vbz_plugin.h"

// cd_values -> CompressionOptions exactly as the filter callback reads them (vbz_plugin.cpp:109-134)
static bool options_from_cd_values(size_t cd_nelmts, const unsigned int cd_values[], CompressionOptions* options)
{
    if (cd_nelmts < 3) return false;  // vbz_plugin.cpp:109-112
    options->vbz_version = cd_values[FILTER_VBZ_VERSION_OPTION];
    options->integer_size = cd_values[FILTER_VBZ_INTEGER_SIZE_OPTION];
    options->perform_delta_zig_zag = cd_values[FILTER_VBZ_USE_DELTA_ZIG_ZAG_COMPRESSION] != 0;
    options->zstd_compression_level = cd_nelmts > FILTER_VBZ_ZSTD_COMPRESSION_LEVEL_OPTION ? cd_values[FILTER_VBZ_ZSTD_COMPRESSION_LEVEL_OPTION] : 1;
    return true;
}

// Many chunks of one dataset per call: what libhdf5 hands the filter one chunk at a time (vbz_plugin.cpp:97-229), as a batch.
// Chunk buffers are `[u32 size][payload]` (vbz.cpp:302-367); buffers belong to the caller (no malloc/free swap here).
extern "C" int vbz_b200_h5z_filter_batch(vbz_b200_ctx* ctx, unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[], uint32_t n_chunks,
                                         const void* const* src, const vbz_size_t* src_size, void* const* dst, const vbz_size_t* dst_cap,
                                         vbz_size_t* results)
{
    CompressionOptions options;
    if (!options_from_cd_values(cd_nelmts, cd_values, &options)) return -1;
    if (flags & VBZ_H5Z_FLAG_REVERSE) return vbz_b200_decompress_batch(ctx, src, src_size, dst, dst_cap, results, n_chunks, &options, 1);
    return vbz_b200_compress_batch(ctx, src, src_size, dst, dst_cap, results, n_chunks, &options, 1);
}

extern "C" size_t vbz_b200_h5z_filter(unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[], size_t /*nbytes*/,
                                      size_t* buf_size, void** buf)
{
    CompressionOptions options;
    if (!options_from_cd_values(cd_nelmts, cd_values, &options)) return 0;

    void* out = nullptr;
    vbz_size_t out_alloc = 0, used = 0;
    if (*buf_size > std::numeric_limits<vbz_size_t>::max()) {
        fprintf(stderr, "vbz_filter: Chunk size too large.\n");
        return 0;
    }
    if (flags & VBZ_H5Z_FLAG_REVERSE) {
        const vbz_size_t expect = vbz_decompressed_size(*buf, (vbz_size_t)*buf_size, &options);
        if (vbz_is_error(expect)) { fprintf(stderr, "vbz_filter: size error\n"); return 0; }
        out = malloc(expect ? expect : 1);
        if (!out) { fprintf(stderr, "vbz_filter: out of memory\n"); return 0; }
        used = vbz_decompress_sized(*buf, (vbz_size_t)*buf_size, out, expect, &options);
        if (vbz_is_error(used)) { fprintf(stderr, "vbz_filter: compression error\n"); free(out); return 0; }
        if (used != expect) { fprintf(stderr, "vbz_filter: decompressed size error\n"); free(out); return 0; }
        // the reference leaves outbuf_size at 0 on this branch (vbz_plugin.cpp:106,227); libhdf5 only
        // needs *buf_size >= the returned byte count to be safe, so report the real allocation instead
        out_alloc = expect;
    } else {
        if (options.integer_size == 0 || *buf_size % options.integer_size != 0) {  // :194-199 (0 divides by zero there)
            fprintf(stderr, "vbz_filter: Invalid integer_size specified\n");
            return 0;
        }
        out_alloc = vbz_max_compressed_size((vbz_size_t)*buf_size, &options);
        if (vbz_is_error(out_alloc)) { fprintf(stderr, "vbz_filter: compression error\n"); return 0; }
        out = malloc(out_alloc);
        if (!out) { fprintf(stderr, "vbz_filter: out of memory\n"); return 0; }
        used = vbz_compress_sized(*buf, (vbz_size_t)*buf_size, out, out_alloc, &options);
        if (vbz_is_error(used)) { fprintf(stderr, "vbz_filter: compression error\n"); free(out); return 0; }
    }
    free(*buf);
    *buf = out;
    *buf_size = out_alloc;
    return used;
}

static const vbz_h5z_class2_t vbz_filter_struct = {1, FILTER_VBZ_ID, 1, 1, "vbz", nullptr, nullptr, vbz_b200_h5z_filter};

extern "C" const void* vbz_plugin_info(void) { return &vbz_filter_struct; }
extern "C" int H5PLget_plugin_type(void) { return 0; /* H5PL_TYPE_FILTER */ }
extern "C" const void* H5PLget_plugin_info(void) { return vbz_plugin_info(); }
