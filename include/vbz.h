/* include/vbz.h - the reference's C API, re-declared for libvbz_b200.so.
 *
 * Every symbol here replaces the reference symbol of the same name and signature
 * (reference vbz/vbz.h:13-141, implemented in vbz/vbz.cpp:61-386), so a program or
 * binding linked against libvbz.so can be re-linked against libvbz_b200.so unchanged.
 * All buffers are HOST pointers; the library stages them through the GPU.  There is no
 * CPU implementation behind these calls: without a usable CUDA device every data call
 * fails loudly - a diagnostic on stderr and VBZ_OUT_OF_MEMORY_ERROR (the reference's
 * only "resource" error code) as the result ("no CPU fallback").
 */
#ifndef VBZ_B200_VBZ_H
#define VBZ_B200_VBZ_H

#include <stdbool.h>
#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#ifndef VBZ_EXPORT
#define VBZ_EXPORT __attribute__((visibility("default")))
#endif

#define VBZ_DEFAULT_VERSION 0

typedef uint32_t vbz_size_t; /* reference vbz/vbz.h:13 */

/* error codes: reference vbz/vbz.h:15-22 */
#define VBZ_ZSTD_ERROR ((vbz_size_t)-1)
#define VBZ_INPUT_SIZE_ERROR ((vbz_size_t)-2)
#define VBZ_INTEGER_SIZE_ERROR ((vbz_size_t)-3)
#define VBZ_DESTINATION_SIZE_ERROR ((vbz_size_t)-4)
#define VBZ_STREAMVBYTE_STREAM_ERROR ((vbz_size_t)-5)
#define VBZ_VERSION_ERROR ((vbz_size_t)-6)
#define VBZ_OUT_OF_MEMORY_ERROR ((vbz_size_t)-7)
#define VBZ_FIRST_ERROR VBZ_OUT_OF_MEMORY_ERROR
/* deprecated aliases: reference vbz/vbz.h:24-27 */
#define VBZ_STREAMVBYTE_INPUT_SIZE_ERROR VBZ_INPUT_SIZE_ERROR
#define VBZ_STREAMVBYTE_INTEGER_SIZE_ERROR VBZ_INTEGER_SIZE_ERROR
#define VBZ_STREAMVBYTE_DESTINATION_SIZE_ERROR VBZ_DESTINATION_SIZE_ERROR

/* reference vbz/vbz.h:29-53.  Layout is ABI: sizeof == 16, bool at 0, unsigned at 4/8/12. */
struct CompressionOptions
{
    bool perform_delta_zig_zag;          /* delta + zigzag before packing */
    unsigned int integer_size;           /* 0 (no packing), 1, 2 or 4 bytes per sample */
    unsigned int zstd_compression_level; /* 0 = no zstd stage */
    unsigned int vbz_version;            /* 0 or 1 (1 differs only for integer_size 1) */
};
#if !defined(__cplusplus)
typedef struct CompressionOptions CompressionOptions;
#endif

/* reference vbz/vbz.h:56, vbz/vbz.cpp:61-64 */
VBZ_EXPORT bool vbz_is_error(vbz_size_t result_value);
/* reference vbz/vbz.h:59, vbz/vbz.cpp:66-77 */
VBZ_EXPORT char const* vbz_error_string(vbz_size_t error_value);
/* reference vbz/vbz.h:65-67, vbz/vbz.cpp:79-114 */
VBZ_EXPORT vbz_size_t vbz_max_compressed_size(vbz_size_t source_size, struct CompressionOptions const* options);
/* reference vbz/vbz.h:76-81, vbz/vbz.cpp:116-208 */
VBZ_EXPORT vbz_size_t vbz_compress(void const* source, vbz_size_t source_size, void* destination,
                                   vbz_size_t destination_capacity, struct CompressionOptions const* options);
/* reference vbz/vbz.h:94-99, vbz/vbz.cpp:210-300 */
VBZ_EXPORT vbz_size_t vbz_decompress(void const* source, vbz_size_t source_size, void* destination,
                                     vbz_size_t destination_size, struct CompressionOptions const* options);
/* reference vbz/vbz.h:109-114, vbz/vbz.cpp:302-330 */
VBZ_EXPORT vbz_size_t vbz_compress_sized(void const* source, vbz_size_t source_size, void* destination,
                                         vbz_size_t destination_capacity, struct CompressionOptions const* options);
/* reference vbz/vbz.h:125-130, vbz/vbz.cpp:332-367 */
VBZ_EXPORT vbz_size_t vbz_decompress_sized(void const* source, vbz_size_t source_size, void* destination,
                                           vbz_size_t destination_capacity, struct CompressionOptions const* options);
/* reference vbz/vbz.h:138-141, vbz/vbz.cpp:369-386 */
VBZ_EXPORT vbz_size_t vbz_decompressed_size(void const* source, vbz_size_t source_size,
                                            struct CompressionOptions const* options);

#if defined(__cplusplus)
}
#endif
#endif
