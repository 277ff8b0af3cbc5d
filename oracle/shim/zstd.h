/* Minimal zstd.h for building the reference against the image's runtime
 * libzstd.so.1 (1.5.5): the image ships no zstd development header.  Only the
 * five functions the reference calls (vbz/vbz.cpp:109,194,201,236,263) are
 * declared.  TEST INFRASTRUCTURE ONLY - never included by the product. */
#ifndef VBZ_B200_ORACLE_ZSTD_SHIM_H
#define VBZ_B200_ORACLE_ZSTD_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
unsigned long long ZSTD_getFrameContentSize(const void* src, size_t srcSize);
#ifdef __cplusplus
}
#endif
#endif
