/* include/vbz_b200.h - batch extension of the VBZ C API (not present in the reference).
 *
 * The reference compresses one read per call on the calling thread (vbz/vbz.h:76-130); a GPU
 * needs many reads per launch.  These entry points take a whole batch of reads and run the same
 * per-read semantics as vbz_compress[_sized] / vbz_decompress[_sized] (vbz/vbz.cpp:116-367) for
 * each of them: results[i] is what the reference call would have returned for read i (a byte count
 * or one of the VBZ_*_ERROR codes).  A bad read never poisons the batch.
 *
 * The single-read functions of include/vbz.h and the H5Z filter of include/vbz_plugin.h are thin
 * wrappers over a batch of one.
 *
 * Two flavours:
 *   *_batch   HOST pointers (one per read).  The library stages through pinned memory, copies to the
 *             device, runs the kernels, copies back.  Synchronous.
 *   *_device  DEVICE pointers: one input buffer and one output buffer in HBM plus per-read offsets
 *             (host arrays).  Work is enqueued on `stream` and the call returns without
 *             synchronising; results[] is a DEVICE array.  Offsets must be multiples of 16.
 *
 * Return value of every function: 0 = the batch ran (inspect results[]), < 0 = the batch could not
 * run at all (no CUDA device, out of device memory, invalid arguments); vbz_b200_last_error() then
 * describes why.  There is no CPU fallback.
 */
#ifndef VBZ_B200_H
#define VBZ_B200_H

#include <stddef.h>
#include "vbz.h"

#if defined(__cplusplus)
extern "C" {
#endif

typedef struct vbz_b200_ctx vbz_b200_ctx;

/* number of usable CUDA devices (0 if none / driver missing) */
VBZ_EXPORT int vbz_b200_device_count(void);
/* A context owns one device's stream, scratch HBM and pinned staging.  Not thread-safe: use one
 * per thread (the single-read API keeps a pool internally).  NULL on failure. */
VBZ_EXPORT vbz_b200_ctx* vbz_b200_create(int device);
VBZ_EXPORT void vbz_b200_destroy(vbz_b200_ctx* ctx);
/* thread-local description of the last failure */
VBZ_EXPORT const char* vbz_b200_last_error(void);

/* ---- host-pointer batches ------------------------------------------------------------------- */
/* per read i: results[i] = vbz_compress[_sized](src[i], src_size[i], dst[i], dst_cap[i], options) */
VBZ_EXPORT int vbz_b200_compress_batch(vbz_b200_ctx* ctx, const void* const* src, const vbz_size_t* src_size,
                                       void* const* dst, const vbz_size_t* dst_cap, vbz_size_t* results,
                                       uint32_t n_reads, const struct CompressionOptions* options, int sized);
/* per read i: results[i] = vbz_decompress[_sized](src[i], src_size[i], dst[i], dst_size[i], options)
 * (dst_size = exact size for the unsized form, capacity for the sized form) */
VBZ_EXPORT int vbz_b200_decompress_batch(vbz_b200_ctx* ctx, const void* const* src, const vbz_size_t* src_size,
                                         void* const* dst, const vbz_size_t* dst_size, vbz_size_t* results,
                                         uint32_t n_reads, const struct CompressionOptions* options, int sized);

/* Many HDF5 chunks of one dataset per call: the batch form of the H5Z filter 32020 callback (include/vbz_plugin.h, reference
 * vbz_plugin/vbz_plugin.cpp:97-229).  flags & 0x0100 (H5Z_FLAG_REVERSE) decodes; cd_values = {version, integer_size, zigzag
 * [, zstd level = 1]} as stored in the dataset's filter pipeline message.  Chunk i is src[i] (`[u32 size][payload]`, src_size[i]
 * bytes) -> dst[i] (capacity dst_cap[i]); results[i] = what vbz_decompress_sized / vbz_compress_sized returns for it.  Returns -1
 * when cd_nelmts < 3 (the filter callback returns 0 for every chunk then). */
VBZ_EXPORT int vbz_b200_h5z_filter_batch(vbz_b200_ctx* ctx, unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[],
                                         uint32_t n_chunks, const void* const* src, const vbz_size_t* src_size, void* const* dst,
                                         const vbz_size_t* dst_cap, vbz_size_t* results);

/* ---- device-resident batches ---------------------------------------------------------------- */
/* Size of the output slot the device-resident API needs for a read of source_size bytes: a tight
 * bound on what the kernels can emit, rounded up to 16 (the reference's vbz_max_compressed_size,
 * vbz/vbz.cpp:79-114, over-asks by up to 2x for int16 and would not let a 100k-read batch fit in HBM). */
VBZ_EXPORT uint64_t vbz_b200_compress_bound(vbz_size_t source_size, const struct CompressionOptions* options, int sized);
/* Read i occupies d_src[src_off[i] .. +src_size[i]); its output goes to d_dst[dst_off[i] ..
 * +dst_cap[i]).  dst_cap[i] must be >= vbz_b200_compress_bound(src_size[i], options, sized).  d_results[i] (device) receives the per-read
 * result.  `stream` is a cudaStream_t used as is (NULL = the CUDA legacy default stream, as in a kernel
 * launch); vbz_b200_stream() returns the context's own non-blocking stream. */
VBZ_EXPORT int vbz_b200_compress_device(vbz_b200_ctx* ctx, const void* d_src, const uint64_t* src_off,
                                        const vbz_size_t* src_size, void* d_dst, const uint64_t* dst_off,
                                        const vbz_size_t* dst_cap, vbz_size_t* d_results, uint32_t n_reads,
                                        const struct CompressionOptions* options, int sized, void* stream);
/* dst_size[i] is the exact decompressed size of read i (the caller allocated it, so it knows it).  With sized != 0 each read starts with
 * the 4-byte size header of vbz_compress_sized; it is SKIPPED here (the host cannot read device memory without a round trip): dst_size[i]
 * must be the value it holds, or the read fails with a stream error.  The host-pointer batches and vbz_decompress_sized read the header
 * and apply the reference's rules (vbz/vbz.cpp:350-366). */
VBZ_EXPORT int vbz_b200_decompress_device(vbz_b200_ctx* ctx, const void* d_src, const uint64_t* src_off,
                                          const vbz_size_t* src_size, void* d_dst, const uint64_t* dst_off,
                                          const vbz_size_t* dst_size, vbz_size_t* d_results, uint32_t n_reads,
                                          const struct CompressionOptions* options, int sized, void* stream);
/* the context's own (non-blocking) stream, as a cudaStream_t */
VBZ_EXPORT void* vbz_b200_stream(vbz_b200_ctx* ctx);
/* wait for everything enqueued on the context's own stream */
VBZ_EXPORT int vbz_b200_synchronize(vbz_b200_ctx* ctx);
/* number of kernels this context has launched so far (bench.py reports it as gpu_launches) */
VBZ_EXPORT uint64_t vbz_b200_kernel_launches(vbz_b200_ctx* ctx);

/* Diagnostics: where the frames of the last zstd-stage decode of this context went (waits for it).  out[0] = blocks listed as units,
 * out[1] = frames decoded by the general warp-per-frame kernel (checksums, several frames in one buffer, damaged input, pool overflow),
 * out[2] = frames with sequences decoded block-parallel (sequence kernels), out[3] = sequence records used.  A batch that runs as
 * several waves reports its last wave. */
VBZ_EXPORT int vbz_b200_decode_stats(vbz_b200_ctx* ctx, uint32_t out[4]);

/* Optional: time the two stages of each *_device call with CUDA events on the call's stream (bench.py's roofline lines).
 * vbz_b200_stage_times waits for the last call and returns the milliseconds its stage-1 kernels (delta + zigzag + StreamVByte,
 * ref vbz/v0, vbz/v1) and its zstd-stage kernels (ref vbz/vbz.cpp:194-204, 236-272) took, launch overheads included. */
VBZ_EXPORT int vbz_b200_set_stage_timing(vbz_b200_ctx* ctx, int on);
VBZ_EXPORT int vbz_b200_stage_times(vbz_b200_ctx* ctx, float* stage1_ms, float* zstd_ms);

#if defined(__cplusplus)
}
#endif
#endif
