// csrc/host/options.h - option validation and the size / error rules of the reference C API,
// restated per call (no sample arithmetic here).
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include "../../../include/vbz.h"
#include "../svb_types.h"

namespace vbz {

struct Stage1Kind {
    int isz = 0;              // 1, 2, 4 (0 = stage 1 disabled)
    int zz = 0;
    int mode = kModeGeneric;  // kModeI16X86 for int16 + zigzag (what the x86 reference build computes)
    bool half = false;        // vbz_version 1 and integer_size 1: nibble packing (ref vbz/v1/vbz_streamvbyte.cpp:37-65)
    bool passthrough = false; // integer_size 0 and level 0: plain copy (ref vbz/vbz.cpp:130-133)
    unsigned zstd_level = 0;
    bool bad_isz = false;     // ref vbz/vbz.cpp:44-50
    bool bad_version = false; // ref vbz/vbz.cpp:147-150, 286-289 (only looked at when integer_size != 0)
};

// int16+zigzag follows the x86 (SSSE3) reference build bit for bit by default.  Setting
// VBZ_B200_INT16_GENERIC=1 selects the generic template's 32-bit delta instead (what ARM builds of
// the reference compute) - the two only differ when |delta| > 32767.
inline bool int16_generic_semantics()
{
    static const bool v = [] { const char* e = getenv("VBZ_B200_INT16_GENERIC"); return e && e[0] == '1'; }();
    return v;
}

inline Stage1Kind classify_options(const CompressionOptions* o)
{
    Stage1Kind k;
    const unsigned isz = o->integer_size;
    k.bad_isz = !(isz == 0 || isz == 1 || isz == 2 || isz == 4);
    k.isz = k.bad_isz ? 0 : (int)isz;
    k.zz = o->perform_delta_zig_zag ? 1 : 0;
    k.zstd_level = o->zstd_compression_level;
    k.bad_version = isz != 0 && o->vbz_version != 0 && o->vbz_version != 1;
    k.passthrough = !k.bad_isz && isz == 0 && k.zstd_level == 0;
    k.half = isz == 1 && o->vbz_version == 1;
    k.mode = (isz == 2 && k.zz && !int16_generic_semantics()) ? kModeI16X86 : kModeGeneric;
    return k;
}

// streamvbyte_max_compressedbytes (ref third_party/streamvbyte/include/streamvbyte.h:33-39) through the
// 32-bit truncation of vbz_max_streamvbyte_compressed_size_v0/v1 (ref vbz/v0/vbz_streamvbyte.cpp:7-18)
inline uint32_t max_streamvbyte_size(unsigned isz, uint32_t source_size)
{
    if (source_size % isz != 0) return kErrInputSize;
    const uint32_t n = source_size / isz;
    return (uint32_t)((uint64_t)((uint32_t)(n + 3u) / 4u) + (uint64_t)n * 4 + 16);  // `length + 3` wraps in 32 bits there, and so does the result
}
// ZSTD_COMPRESSBOUND (zstd.h), the value ZSTD_compressBound returns for sizes in 32-bit range
inline uint64_t zstd_compress_bound(uint64_t n)
{
    return n + (n >> 8) + (n < (128u << 10) ? (((128u << 10) - n) >> 11) : 0);
}
inline bool is_error(uint32_t v) { return v >= kErrOutOfMemory; }
// Largest frame the device encoder writes for `n` content bytes: frame header, then per 8 KiB block (two regions,
// each with a partial last block) a 3-byte header and at most the block itself (csrc/zstd_enc_kernels.cuh).
// Never above ZSTD_compressBound(n).
inline uint64_t zstd_enc_worst_case(uint64_t n)
{
    const uint64_t blocks = n / kZEncSub + n / kZEncBlockKeys + 4;  // every group may end with a partial block; two regions
    return 9 + 3 * blocks + n;
}

// ref vbz/vbz.cpp:79-114
inline uint32_t max_compressed_size(uint32_t source_size, const CompressionOptions* o)
{
    const Stage1Kind k = classify_options(o);
    if (k.bad_isz) return kErrIntegerSize;
    uint32_t m = source_size;
    if (k.isz != 0) {
        if (k.bad_version) return kErrVersion;
        m = max_streamvbyte_size((unsigned)k.isz, m);
        if (is_error(m)) return m;
    }
    if (k.zstd_level != 0) m = (uint32_t)zstd_compress_bound(m);
    return m + 4;
}

// What vbz_compress[_sized] returns before touching any sample (ref vbz/vbz.cpp:116-175, 302-330), in
// the reference's order.  Returns true when the read is rejected; *code is then the value the caller
// sees (the sized form adds 4 to whatever vbz_compress returned, errors included - vbz.cpp:329).
// Largest stage-1 output the kernels can produce for a read (tight, unlike the API's K + 4N + 16): the
// x86 int16 path never emits more than 2 data bytes per value.
inline uint64_t stage1_tight_bound(uint32_t n_values, const Stage1Kind& k)
{
    const uint64_t keys = ((uint64_t)n_values + 3) / 4;
    return keys + (uint64_t)n_values * (k.mode == kModeI16X86 ? 2 : 4);
}

// `device_contract`: the device-resident API only asks for vbz_b200_compress_bound() bytes per slot.
inline bool compress_precheck(const Stage1Kind& k, uint32_t src_size, uint32_t dst_cap, bool sized, uint32_t* code,
                              bool device_contract = false)
{
    if (k.bad_isz) { *code = kErrIntegerSize; return true; }
    uint32_t cap = dst_cap;
    if (sized) {
        if (dst_cap < 4) { *code = kErrDestinationSize; return true; }  // the reference terminates here (gsl contract)
        cap = dst_cap - 4;
    }
    uint32_t e = 0;
    if (k.passthrough) {
        if (src_size > cap) e = kErrDestinationSize;
    } else if (k.isz != 0) {
        if (k.bad_version) e = kErrVersion;
        else {
            const uint32_t m = max_streamvbyte_size((unsigned)k.isz, src_size);
            if (is_error(m)) e = m;
            else if (k.zstd_level == 0 && (device_contract ? stage1_tight_bound(src_size / (unsigned)k.isz, k) : (uint64_t)m) > cap) e = kErrDestinationSize;
        }
    }
    if (!e) return false;
    *code = sized ? e + 4 : e;
    return true;
}

// ZSTD_getFrameContentSize of the first frame in [p, p + n) (ref call site vbz/vbz.cpp:236): the content size, or
// kZstdContentSizeUnknown / kZstdContentSizeError (both count as errors at vbz.cpp:237).  Header parsing only.
constexpr uint64_t kZstdContentSizeUnknown = ~0ull, kZstdContentSizeError = ~0ull - 1;
inline uint64_t zstd_frame_content_size(const uint8_t* p, uint64_t n)
{
    auto le32 = [&](uint64_t o) { return (uint32_t)p[o] | ((uint32_t)p[o + 1] << 8) | ((uint32_t)p[o + 2] << 16) | ((uint32_t)p[o + 3] << 24); };
    if (n >= 8 && (le32(0) & 0xFFFFFFF0u) == 0x184D2A50u) return 0;  // skippable frame
    if (n < 5 || le32(0) != 0xFD2FB528u) return kZstdContentSizeError;
    const int fhd = p[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    if (fhd & 0x08) return kZstdContentSizeError;
    uint64_t h = 5;
    if (!single) {
        if (h + 1 > n || 10 + (p[h] >> 3) > 31) return kZstdContentSizeError;
        h++;
    }
    h += did == 3 ? 4 : did;
    const uint64_t fb = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (h + fb > n) return kZstdContentSizeError;
    if (fb == 0) return kZstdContentSizeUnknown;
    uint64_t v = 0;
    for (uint64_t i = 0; i < fb; ++i) v |= (uint64_t)p[h + i] << (8 * i);
    return fb == 2 ? v + 256 : v;
}

// What vbz_decompress returns before touching the payload when there is no zstd stage
// (ref vbz/vbz.cpp:210-299, vbz/v0/vbz_streamvbyte.cpp:73-76).  payload_size excludes the sized header.
inline bool decompress_precheck(const Stage1Kind& k, uint32_t payload_size, uint32_t dst_size, uint32_t* code)
{
    if (k.bad_isz) { *code = kErrIntegerSize; return true; }
    if (k.passthrough) {
        if (payload_size > dst_size) { *code = kErrDestinationSize; return true; }
        return false;
    }
    if (k.zstd_level != 0) return false;  // the zstd stage reports first (vbz.cpp:234-273)
    if (k.isz == 0) return false;
    if (k.bad_version) { *code = kErrVersion; return true; }
    if (dst_size % (unsigned)k.isz != 0) { *code = kErrDestinationSize; return true; }
    return false;
}

}  // namespace vbz
