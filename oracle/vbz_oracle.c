/* oracle/vbz_oracle.c - TEST INFRASTRUCTURE ONLY.  See vbz_oracle.h.
 *
 * Scalar restatement of the reference's stage 1 (delta + zigzag + StreamVByte)
 * and of the vbz C API that frames it.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference).  Nothing here is
 * copied from the reference: the reference's SIMD/LUT code is restated as the
 * arithmetic it performs.
 */
#include "vbz_oracle.h"
#include <stdlib.h>
#include <string.h>

static int g_int16_x86 = 1;
static int g_use_restated_zstd = 1;
void vbzo_set_int16_x86_semantics(int on) { g_int16_x86 = on; }
void vbzo_set_use_restated_zstd_decoder(int on) { g_use_restated_zstd = on; }

/* ---- error helpers : vbz/vbz.cpp:61-77 ------------------------------------ */
int vbzo_is_error(vbzo_size_t v) { return v >= VBZO_OUT_OF_MEMORY_ERROR; }
const char* vbzo_error_string(vbzo_size_t v)
{
    switch (v) {
    case VBZO_ZSTD_ERROR: return "VBZ_ZSTD_ERROR";
    case VBZO_INPUT_SIZE_ERROR: return "VBZ_INPUT_SIZE_ERROR";
    case VBZO_INTEGER_SIZE_ERROR: return "VBZ_INTEGER_SIZE_ERROR";
    case VBZO_DESTINATION_SIZE_ERROR: return "VBZ_DESTINATION_SIZE_ERROR";
    case VBZO_STREAMVBYTE_STREAM_ERROR: return "VBZ_STREAMVBYTE_STREAM_ERROR";
    case VBZO_VERSION_ERROR: return "VBZ_VERSION_ERROR";
    case VBZO_OUT_OF_MEMORY_ERROR: return "VBZ_OUT_OF_MEMORY_ERROR";
    default: return "VBZ_UNKNOWN_ERROR";
    }
}

/* ---- element access: the reference treats every dtype as SIGNED of its width
 * (vbz/v0/vbz_streamvbyte.cpp:35-62) and widens by sign extension
 * (vbz/v0/vbz_streamvbyte_impl.h:83-91). ------------------------------------ */
static int32_t load_elem(const uint8_t* p, int isz, size_t i)
{
    if (isz == 1) return (int8_t)p[i];
    if (isz == 2) { int16_t v; memcpy(&v, p + 2 * i, 2); return v; }
    { int32_t v; memcpy(&v, p + 4 * i, 4); return v; }
}
static void store_elem(uint8_t* p, int isz, size_t i, uint32_t v)
{
    if (isz == 1) p[i] = (uint8_t)v;
    else if (isz == 2) { uint16_t t = (uint16_t)v; memcpy(p + 2 * i, &t, 2); }
    else memcpy(p + 4 * i, &v, 4);
}

/* third_party/streamvbyte/src/streamvbyte_zigzag.c:3-6 */
static uint32_t zigzag32(uint32_t v) { return (v + v) ^ (uint32_t)((int32_t)v >> 31); }
/* third_party/streamvbyte/src/streamvbyte_zigzag.c:20-23 */
static uint32_t unzigzag32(uint32_t v) { return (v >> 1) ^ (0u - (v & 1u)); }

/* third_party/streamvbyte/include/streamvbyte.h:33-39, vbz/v0/vbz_streamvbyte.cpp:7-18 */
vbzo_size_t vbzo_max_streamvbyte_compressed_size(unsigned integer_size, vbzo_size_t source_size)
{
    if (source_size % integer_size != 0) return VBZO_INPUT_SIZE_ERROR;
    uint32_t n = source_size / integer_size;
    uint64_t cb = (uint32_t)(n + 3u) / 4u, db = (uint64_t)n * 4; /* `length + 3` is uint32 arithmetic in streamvbyte.h:35 */
    return (vbzo_size_t)(cb + db + 16);
}

/* The pre-transform: value i -> the unsigned 32-bit word StreamVByte packs.
 *   zigzag, int16, x86 build : 16-bit delta, 16-bit zigzag, zero-extended
 *       (vbz/v0/vbz_streamvbyte_impl_sse3.h:425-447 SIMD body, :352-365 scalar tail - same arithmetic)
 *   zigzag, otherwise        : 32-bit delta of the sign-extended values, 32-bit zigzag
 *       (vbz/v0/vbz_streamvbyte_impl.h:32-34, streamvbyte_zigzag.c:13-18)
 *   no zigzag                : the sign-extended value itself (vbz/v0/vbz_streamvbyte_impl.h:24-31) */
static uint32_t pre_transform(const uint8_t* src, int isz, int zz, size_t i, int32_t* prev)
{
    int32_t x = load_elem(src, isz, i);
    if (!zz) return (uint32_t)x;
    if (isz == 2 && g_int16_x86) {
        int16_t d = (int16_t)((uint16_t)x - (uint16_t)*prev);
        *prev = x;
        return (uint16_t)(((uint16_t)d << 1) ^ (uint16_t)(d >> 15));
    }
    uint32_t d = (uint32_t)x - (uint32_t)*prev;
    *prev = x;
    return zigzag32(d);
}

/* StreamVByte v0 packing: third_party/streamvbyte/src/streamvbyte_encode.c:14-62,113-137
 * (the SSE4.1 / SSSE3 variants emit the same bytes). */
static vbzo_size_t encode_v0(const uint8_t* src, uint32_t n, uint8_t* dst, int isz, int zz)
{
    if (n == 0) return 0; /* impl_sse3.h:409-413; streamvbyte_encode.c:45-46 writes nothing either */
    uint32_t klen = (n + 3) / 4;
    uint8_t* kp = dst;
    uint8_t* dp = dst + klen;
    int32_t prev = 0;
    uint8_t key = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t v = pre_transform(src, isz, zz, i, &prev);
        uint32_t code = (v > 0xFFu) + (v > 0xFFFFu) + (v > 0xFFFFFFu);
        for (uint32_t b = 0; b <= code; ++b) *dp++ = (uint8_t)(v >> (8 * b));
        key |= (uint8_t)(code << (2 * (i & 3)));
        if ((i & 3) == 3 || i + 1 == n) { *kp++ = key; key = 0; }
    }
    return (vbzo_size_t)(dp - dst);
}

/* v1 "half" packing, integer_size 1 only: vbz/v1/vbz_streamvbyte_impl.h:85-157,
 * codes 0/1/2/3 = 0/4/8/16 data bits, nibble stream low-nibble first, values >= 2^16
 * keep their low 16 bits. */
static vbzo_size_t encode_v1_half(const uint8_t* src, uint32_t n, uint8_t* dst, int zz)
{
    uint32_t klen = (n + 3) / 4;
    uint8_t* kp = dst;
    uint8_t* dp = dst + klen;
    if (n == 0) return (vbzo_size_t)(dp - dst); /* :131-132 - returns dataPtr, i.e. 0 bytes */
    uint64_t nib = 0; /* nibbles written so far */
    int32_t prev = 0;
    uint8_t key = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t v = pre_transform(src, 1, zz, i, &prev);
        uint32_t code, nn;
        if (v == 0) { code = 0; nn = 0; }
        else if (v < 16u) { code = 1; nn = 1; }
        else if (v < 256u) { code = 2; nn = 2; }
        else { code = 3; nn = 4; }
        for (uint32_t k = 0; k < nn; ++k) {
            uint8_t q = (uint8_t)((v >> (4 * k)) & 0xF);
            if ((nib & 1) == 0) dp[nib >> 1] = q; else dp[nib >> 1] |= (uint8_t)(q << 4);
            ++nib;
        }
        key |= (uint8_t)(code << (2 * (i & 3)));
        if ((i & 3) == 3 || i + 1 == n) { *kp++ = key; key = 0; }
    }
    return (vbzo_size_t)(klen + (nib + 1) / 2);
}

/* vbz/v0/vbz_streamvbyte.cpp:20-63 and vbz/v1/vbz_streamvbyte.cpp:22-65 */
vbzo_size_t vbzo_stage1_encode(const void* src, vbzo_size_t src_size, void* dst, vbzo_size_t dst_cap,
                               int isz, int zz, int version)
{
    (void)dst_cap; /* the reference never checks it here */
    if (isz != 1 && isz != 2 && isz != 4) return VBZO_INTEGER_SIZE_ERROR;
    if (src_size % (unsigned)isz != 0) return VBZO_INPUT_SIZE_ERROR;
    uint32_t n = src_size / (unsigned)isz;
    if (version == 1 && isz == 1) return encode_v1_half((const uint8_t*)src, n, (uint8_t*)dst, zz);
    return encode_v0((const uint8_t*)src, n, (uint8_t*)dst, isz, zz);
}

/* third_party/streamvbyte/src/streamvbyte_decode.c:89-136 (v0) and
 * vbz/v1/vbz_streamvbyte_impl.h:183-216 (half): does the key stream describe exactly the
 * bytes that follow it? */
static int validate_stream(const uint8_t* in, size_t in_count, uint32_t out_count, int half)
{
    if (in_count == 0 || out_count == 0) return in_count == out_count;
    uint32_t klen = out_count / 4 + ((out_count & 3) ? 1 : 0);
    if (klen > in_count) return 0;
    uint64_t enc = 0;
    for (uint32_t c = 0; c < out_count; ++c) {
        uint32_t code = (in[c >> 2] >> (2 * (c & 3))) & 3;
        enc += half ? ((1u << code) >> 1) : (code + 1);
    }
    if (half) enc = (enc + 1) / 2;
    return enc == in_count - klen;
}

/* generic worker: vbz/v0/vbz_streamvbyte_impl.h:43-81 */
static vbzo_size_t decode_v0_generic(const uint8_t* in, uint32_t in_size, uint8_t* out, uint32_t n, int isz, int zz)
{
    if (!validate_stream(in, in_size, n, 0)) return VBZO_STREAMVBYTE_STREAM_ERROR;
    uint32_t klen = (n + 3) / 4;
    const uint8_t* dp = in + klen;
    uint32_t prev = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t code = (in[i >> 2] >> (2 * (i & 3))) & 3, v = 0;
        for (uint32_t b = 0; b <= code; ++b) v |= (uint32_t)(*dp++) << (8 * b);
        if (zz) { prev += unzigzag32(v); v = prev; } /* streamvbyte_zigzag.c:32-38 */
        store_elem(out, isz, i, v);                   /* truncating cast, impl.h:94-100 */
    }
    return n * (uint32_t)isz;
}

/* x86 int16+zigzag worker: vbz/v0/vbz_streamvbyte_impl_sse3.h:468-580.
 * Groups of 8 values are taken by the SIMD body while at least 32 data bytes remain
 * (:494-500); that body keeps only the low 16 bits of each value and un-zigzags in 16 bits
 * (:505-520).  Everything after the first group that fails the test goes through the scalar
 * tail, which un-zigzags in 32 bits and truncates on store (:542-572, :367-378). */
static vbzo_size_t decode_v0_int16_x86(const uint8_t* in, uint32_t in_size, uint8_t* out, uint32_t n)
{
    if (n == 0) return 0; /* :473-476 - before any input check */
    uint32_t klen = (n + 3) / 4;
    if (in_size < klen) return VBZO_INPUT_SIZE_ERROR; /* :479-482 */
    const uint8_t* data = in + klen;
    uint32_t dlen = in_size - klen, pos = 0, i = 0;
    uint16_t prev = 0;
    for (uint32_t g = 0; g < n / 8; ++g) {
        if (dlen - pos < 32) break;
        for (uint32_t k = 0; k < 8; ++k, ++i) {
            uint32_t code = (in[i >> 2] >> (2 * (i & 3))) & 3, v = 0;
            for (uint32_t b = 0; b <= code; ++b) v |= (uint32_t)data[pos++] << (8 * b);
            uint16_t u = (uint16_t)v;
            uint16_t d = (uint16_t)((u >> 1) ^ (uint16_t)(0u - (u & 1u)));
            prev = (uint16_t)(prev + d);
            store_elem(out, 2, i, prev);
        }
    }
    for (; i < n; ++i) {
        uint32_t code = (in[i >> 2] >> (2 * (i & 3))) & 3, v = 0;
        if (dlen - pos < code + 1) return VBZO_STREAMVBYTE_STREAM_ERROR; /* :638-642 */
        for (uint32_t b = 0; b <= code; ++b) v |= (uint32_t)data[pos++] << (8 * b);
        prev = (uint16_t)(unzigzag32(v) + (uint32_t)(int32_t)(int16_t)prev);
        store_elem(out, 2, i, prev);
    }
    if (pos != dlen) return VBZO_STREAMVBYTE_STREAM_ERROR; /* :574-577 */
    return n * 2;
}

/* vbz/v1/vbz_streamvbyte_impl.h:20-83,170-181,246-279 */
static vbzo_size_t decode_v1_half(const uint8_t* in, uint32_t in_size, uint8_t* out, uint32_t n, int zz)
{
    if (!validate_stream(in, in_size, n, 1)) return VBZO_STREAMVBYTE_STREAM_ERROR;
    uint32_t klen = (n + 3) / 4;
    const uint8_t* dp = in + klen;
    uint64_t nib = 0;
    uint32_t prev = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t code = (in[i >> 2] >> (2 * (i & 3))) & 3, nn = (1u << code) >> 1, v = 0;
        for (uint32_t k = 0; k < nn; ++k, ++nib) v |= (uint32_t)((dp[nib >> 1] >> (4 * (nib & 1))) & 0xF) << (4 * k);
        if (zz) { prev += unzigzag32(v); v = prev; }
        store_elem(out, 1, i, v);
    }
    return n;
}

/* vbz/v0/vbz_streamvbyte.cpp:65-108 and vbz/v1/vbz_streamvbyte.cpp:67-113 */
vbzo_size_t vbzo_stage1_decode(const void* src, vbzo_size_t src_size, void* dst, vbzo_size_t dst_size,
                               int isz, int zz, int version)
{
    if (isz != 1 && isz != 2 && isz != 4) return VBZO_INTEGER_SIZE_ERROR;
    if (dst_size % (unsigned)isz != 0) return VBZO_DESTINATION_SIZE_ERROR;
    uint32_t n = dst_size / (unsigned)isz;
    if (version == 1 && isz == 1) return decode_v1_half((const uint8_t*)src, src_size, (uint8_t*)dst, n, zz);
    if (isz == 2 && zz && g_int16_x86) return decode_v0_int16_x86((const uint8_t*)src, src_size, (uint8_t*)dst, n);
    return decode_v0_generic((const uint8_t*)src, src_size, (uint8_t*)dst, n, isz, zz);
}

/* ---- C API ------------------------------------------------------------------ */
static int valid_isz(const VbzoOptions* o) /* vbz/vbz.cpp:44-50 */
{
    return o->integer_size == 0 || o->integer_size == 1 || o->integer_size == 2 || o->integer_size == 4;
}

/* vbz/vbz.cpp:79-114 */
vbzo_size_t vbzo_max_compressed_size(vbzo_size_t source_size, const VbzoOptions* o)
{
    if (!valid_isz(o)) return VBZO_INTEGER_SIZE_ERROR;
    vbzo_size_t m = source_size;
    if (o->integer_size != 0) {
        if (o->vbz_version != 0 && o->vbz_version != 1) return VBZO_VERSION_ERROR;
        m = vbzo_max_streamvbyte_compressed_size(o->integer_size, m);
        if (vbzo_is_error(m)) return m;
    }
    if (o->zstd_compression_level != 0) m = (vbzo_size_t)vbzo_zstd_compress_bound(m);
    return m + 4;
}

/* vbz/vbz.cpp:116-208 */
vbzo_size_t vbzo_compress(const void* src, vbzo_size_t n, void* dst, vbzo_size_t cap, const VbzoOptions* o)
{
    if (!valid_isz(o)) return VBZO_INTEGER_SIZE_ERROR;
    if (o->zstd_compression_level == 0 && o->integer_size == 0) {
        if (n > cap) return VBZO_DESTINATION_SIZE_ERROR;
        memcpy(dst, src, n);
        return n;
    }
    void* tmp = NULL;
    const void* cur = src;
    vbzo_size_t cur_n = n;
    if (o->integer_size != 0) {
        if (o->vbz_version != 0 && o->vbz_version != 1) return VBZO_VERSION_ERROR;
        vbzo_size_t mx = vbzo_max_streamvbyte_compressed_size(o->integer_size, n);
        if (vbzo_is_error(mx)) return mx;
        void* s1 = dst;
        if (o->zstd_compression_level != 0) {
            tmp = malloc(mx ? mx : 1);
            if (!tmp) return VBZO_OUT_OF_MEMORY_ERROR;
            s1 = tmp;
        } else if (mx > cap) {
            return VBZO_DESTINATION_SIZE_ERROR;
        }
        cur_n = vbzo_stage1_encode(src, n, s1, mx, (int)o->integer_size, o->perform_delta_zig_zag, (int)o->vbz_version);
        cur = s1;
    }
    if (o->zstd_compression_level == 0) { free(tmp); return cur_n; }
    size_t z = vbzo_libzstd_compress(dst, cap, cur, cur_n, (int)o->zstd_compression_level);
    free(tmp);
    if (z == (size_t)-1) return VBZO_ZSTD_ERROR;
    return (vbzo_size_t)z;
}

/* vbz/vbz.cpp:210-300 */
vbzo_size_t vbzo_decompress(const void* src, vbzo_size_t n, void* dst, vbzo_size_t dst_size, const VbzoOptions* o)
{
    if (!valid_isz(o)) return VBZO_INTEGER_SIZE_ERROR;
    if (o->zstd_compression_level == 0 && o->integer_size == 0) {
        if (n > dst_size) return VBZO_DESTINATION_SIZE_ERROR;
        memcpy(dst, src, n);
        return n;
    }
    void* tmp = NULL;
    const void* cur = src;
    vbzo_size_t cur_n = n;
    if (o->zstd_compression_level != 0) {
        unsigned long long fcs = vbzo_zstd_frame_content_size(src, n);
        if (fcs >= VBZO_ZSTD_CONTENTSIZE_ERROR) return VBZO_ZSTD_ERROR; /* ZSTD_isError() of either sentinel, :236-240 */
        void* zdst = dst;
        size_t zcap = dst_size;
        if (o->integer_size != 0) {
            if (fcs > ((size_t)1 << 40)) return VBZO_OUT_OF_MEMORY_ERROR; /* malloc of a wild size fails, :252-255 */
            tmp = malloc(fcs ? (size_t)fcs : 1);
            if (!tmp) return VBZO_OUT_OF_MEMORY_ERROR;
            zdst = tmp;
            zcap = (vbzo_size_t)fcs; /* the span is built from a truncated size, :256 */
        } else if (fcs > dst_size) {
            return VBZO_DESTINATION_SIZE_ERROR;
        }
        size_t z = g_use_restated_zstd ? vbzo_zstd_decompress(zdst, zcap, src, n) : vbzo_libzstd_decompress(zdst, zcap, src, n);
        if (z == (size_t)-1) { free(tmp); return VBZO_ZSTD_ERROR; }
        cur = zdst;
        cur_n = (vbzo_size_t)z;
    }
    if (o->integer_size == 0) { free(tmp); return cur_n; }
    if (o->vbz_version != 0 && o->vbz_version != 1) { free(tmp); return VBZO_VERSION_ERROR; }
    vbzo_size_t r = vbzo_stage1_decode(cur, cur_n, dst, dst_size, (int)o->integer_size, o->perform_delta_zig_zag, (int)o->vbz_version);
    free(tmp);
    return r;
}

/* vbz/vbz.cpp:302-330.  Note the reference adds 4 to whatever vbz_compress returned, error
 * codes included - restated as is.  A capacity below 4 makes the reference
 * terminate (gsl contract, :316); the port reports DESTINATION_SIZE instead. */
vbzo_size_t vbzo_compress_sized(const void* src, vbzo_size_t n, void* dst, vbzo_size_t cap, const VbzoOptions* o)
{
    if (!valid_isz(o)) return VBZO_INTEGER_SIZE_ERROR;
    if (cap < 4) return VBZO_DESTINATION_SIZE_ERROR;
    memcpy(dst, &n, 4);
    return vbzo_compress(src, n, (uint8_t*)dst + 4, cap - 4, o) + 4;
}

/* vbz/vbz.cpp:332-367 */
vbzo_size_t vbzo_decompress_sized(const void* src, vbzo_size_t n, void* dst, vbzo_size_t cap, const VbzoOptions* o)
{
    if (!valid_isz(o)) return VBZO_INTEGER_SIZE_ERROR;
    if (n < 4) return VBZO_INPUT_SIZE_ERROR;
    uint32_t orig;
    memcpy(&orig, src, 4);
    if (cap < orig) return VBZO_DESTINATION_SIZE_ERROR;
    return vbzo_decompress((const uint8_t*)src + 4, n - 4, dst, orig, o);
}

/* vbz/vbz.cpp:369-386 */
vbzo_size_t vbzo_decompressed_size(const void* src, vbzo_size_t n, const VbzoOptions* o)
{
    if (!valid_isz(o)) return VBZO_INTEGER_SIZE_ERROR;
    if (n < 4) return VBZO_INPUT_SIZE_ERROR;
    uint32_t orig;
    memcpy(&orig, src, 4);
    return orig;
}
