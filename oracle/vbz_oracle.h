/* oracle/vbz_oracle.h - TEST INFRASTRUCTURE ONLY (parity checker; never linked
 * into, imported by or called from the product library libvbz_b200.so).
 *
 * Plain-C, scalar restatement of the reference's signal-compression path:
 *   stage 1  delta + zigzag + StreamVByte   (reference vbz/v0, vbz/v1, third_party/streamvbyte)
 *   stage 2  zstd frame                     (reference calls libzstd: vbz/vbz.cpp:194,236,263)
 *   framing  4-byte "sized" header          (reference vbz/vbz.cpp:302-386)
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this port against
 *   - every byte-exact KAT of the reference's own tests (tests/golden/kat.json),
 *   - the reference's stored fast5 fixtures (tests/golden/fast5_chunks.npz),
 *   - the compiled reference itself (oracle/_ref/libvbz_ref.so) on random inputs
 *     and on the reference's fuzz corpus.
 *
 * zstd: the reference's stage 2 is the third-party libzstd (conan pin zstd/1.4.8,
 * CMakeLists.txt:93; any >=1.3.1 accepted, CMakeLists.txt:156), whose source is not
 * under /root/reference.  Decoding is fully determined by the published format
 * (RFC 8878) and is restated here in zstd_oracle.c.  Encoding is not byte-pinned by
 * the reference - vbzo_compress() therefore produces its stage-2
 * bytes by calling the image's runtime libzstd.so.1 through dlopen, exactly as the
 * reference would.
 */
#ifndef VBZ_ORACLE_H
#define VBZ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef uint32_t vbzo_size_t;
#define VBZO_ZSTD_ERROR ((vbzo_size_t)-1)
#define VBZO_INPUT_SIZE_ERROR ((vbzo_size_t)-2)
#define VBZO_INTEGER_SIZE_ERROR ((vbzo_size_t)-3)
#define VBZO_DESTINATION_SIZE_ERROR ((vbzo_size_t)-4)
#define VBZO_STREAMVBYTE_STREAM_ERROR ((vbzo_size_t)-5)
#define VBZO_VERSION_ERROR ((vbzo_size_t)-6)
#define VBZO_OUT_OF_MEMORY_ERROR ((vbzo_size_t)-7)

/* Same layout as the reference's CompressionOptions (vbz/vbz.h:29-53): 16 bytes. */
typedef struct {
    _Bool perform_delta_zig_zag;
    unsigned int integer_size;
    unsigned int zstd_compression_level;
    unsigned int vbz_version;
} VbzoOptions;

/* 1 (default): int16+zigzag follows the x86/SSSE3 specialisation (16-bit delta, the
 * SIMD/scalar split of the decoder) - vbz/v0/vbz_streamvbyte_impl_sse3.h:403-659.
 * 0: follows the generic template (32-bit delta) - vbz/v0/vbz_streamvbyte_impl.h:15-101. */
void vbzo_set_int16_x86_semantics(int on);

/* stage 1 only (vbz/v0/vbz_streamvbyte.cpp:7-108, vbz/v1/vbz_streamvbyte.cpp:9-113) */
vbzo_size_t vbzo_max_streamvbyte_compressed_size(unsigned integer_size, vbzo_size_t source_size);
vbzo_size_t vbzo_stage1_encode(const void* src, vbzo_size_t src_size, void* dst, vbzo_size_t dst_cap,
                               int integer_size, int zigzag, int version);
vbzo_size_t vbzo_stage1_decode(const void* src, vbzo_size_t src_size, void* dst, vbzo_size_t dst_size,
                               int integer_size, int zigzag, int version);

/* the C API (vbz/vbz.cpp:61-386) */
int vbzo_is_error(vbzo_size_t v);
const char* vbzo_error_string(vbzo_size_t v);
vbzo_size_t vbzo_max_compressed_size(vbzo_size_t source_size, const VbzoOptions* o);
vbzo_size_t vbzo_compress(const void* src, vbzo_size_t n, void* dst, vbzo_size_t cap, const VbzoOptions* o);
vbzo_size_t vbzo_decompress(const void* src, vbzo_size_t n, void* dst, vbzo_size_t dst_size, const VbzoOptions* o);
vbzo_size_t vbzo_compress_sized(const void* src, vbzo_size_t n, void* dst, vbzo_size_t cap, const VbzoOptions* o);
vbzo_size_t vbzo_decompress_sized(const void* src, vbzo_size_t n, void* dst, vbzo_size_t cap, const VbzoOptions* o);
vbzo_size_t vbzo_decompressed_size(const void* src, vbzo_size_t n, const VbzoOptions* o);

/* stage 2 (zstd_oracle.c) */
#define VBZO_ZSTD_CONTENTSIZE_UNKNOWN ((unsigned long long)-1)
#define VBZO_ZSTD_CONTENTSIZE_ERROR ((unsigned long long)-2)
size_t vbzo_zstd_compress_bound(size_t n);                              /* ZSTD_COMPRESSBOUND macro, zstd.h */
unsigned long long vbzo_zstd_frame_content_size(const void* src, size_t n); /* ZSTD_getFrameContentSize */
/* RFC 8878 decoder restatement.  Returns bytes written or (size_t)-1 on any format error. */
size_t vbzo_zstd_decompress(void* dst, size_t cap, const void* src, size_t n);
/* 1 = route vbzo_decompress's stage 2 through the restated decoder (default), 0 = through libzstd */
void vbzo_set_use_restated_zstd_decoder(int on);
/* libzstd through dlopen (encode side; (size_t)-1 on failure) */
size_t vbzo_libzstd_compress(void* dst, size_t cap, const void* src, size_t n, int level);
size_t vbzo_libzstd_decompress(void* dst, size_t cap, const void* src, size_t n);

#ifdef __cplusplus
}
#endif
#endif
