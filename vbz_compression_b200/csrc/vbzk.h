// csrc/vbzk.h - the thin C-ABI shim between the C++ host layer and the CUDA kernels.
// Plain pointers, sizes and a stream handle; no C++ types cross it.  Every function only enqueues
// work on `stream` and returns 0, or a non-zero cudaError_t from the launch.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct VbzkSvbRead;   /* == vbz::SvbRead   (csrc/svb_kernels.cuh) */
struct VbzkSvbParams; /* == vbz::SvbParams (csrc/svb_kernels.cuh) */

/* geometry helpers (host side needs them to size tables) */
uint32_t vbzk_svb_tile_values(int integer_size, int mode);
uint32_t vbzk_svb_smem_bytes(int integer_size, int mode, int decode);

/* tile -> read table */
int vbzk_svb_fill_tiles(const struct VbzkSvbRead* d_reads, uint32_t n_reads, uint32_t tile_values, uint32_t* d_tile_read, void* stream);
/* stage 1, v0 packing: integer_size 1/2/4, zigzag 0/1, mode 0 = int16 x86 semantics, 1 = generic */
int vbzk_svb_encode(int integer_size, int zigzag, int mode, const struct VbzkSvbParams* params, uint32_t n_tiles, void* stream);
/* decode pre-pass: fills SvbParams::wtile_off from the control bytes (must precede vbzk_svb_decode) */
int vbzk_svb_key_offsets(int integer_size, int mode, const struct VbzkSvbParams* params, void* stream);
int vbzk_svb_decode(int integer_size, int zigzag, int mode, const struct VbzkSvbParams* params, uint32_t n_tiles, void* stream);
/* stage 1, v1 "half" packing (integer_size 1, vbz_version 1; csrc/svb_half_kernels.cuh): one warp per read, no tile tables -
 * only src, dst, reads, results, src_sizes, hdr_bytes and n_reads of the params are used */
int vbzk_svb_half_encode(int zigzag, const struct VbzkSvbParams* params, void* stream);
int vbzk_svb_half_decode(int zigzag, const struct VbzkSvbParams* params, void* stream);

/* stage 2, decode: one zstd frame per read (csrc/zstd_dec_kernels.cuh).  The launch uses vbzk_zstd_decode_grid() CTAs
 * whatever the batch size; lit_scratch must hold vbzk_zstd_decode_scratch_bytes() bytes and *counter must be 0. */
struct VbzkZstdDecParams; /* == vbz::ZstdDecParams (csrc/svb_types.h) */
uint32_t vbzk_zstd_decode_grid(void);
uint64_t vbzk_zstd_decode_scratch_bytes(void);
int vbzk_zstd_decode(const struct VbzkZstdDecParams* params, void* stream);       /* = front, then back */
int vbzk_zstd_decode_front(const struct VbzkZstdDecParams* params, void* stream); /* walker + sequence kernel: nothing of the output is written */
int vbzk_zstd_decode_back(const struct VbzkZstdDecParams* params, void* stream);  /* units kernel, executor, general kernel */

/* stage 2, encode: one zstd frame per read (csrc/zstd_enc_kernels.cuh); one CTA per 32 KiB block */
struct VbzkZstdEncRead;   /* == vbz::ZstdEncRead */
struct VbzkZstdEncParams; /* == vbz::ZstdEncParams */
uint32_t vbzk_zstd_enc_block_bytes(void);
uint32_t vbzk_zstd_enc_key_block_bytes(void); /* group size inside the control-byte region (ZstdEncRead::split) */
int vbzk_zstd_enc_fill_blocks(const struct VbzkZstdEncRead* d_reads, uint32_t n_reads, uint32_t* d_block_read, void* stream);
int vbzk_zstd_encode(const struct VbzkZstdEncParams* params, void* stream);

/* host <-> device transfer of a wave of reads by ONE kernel (csrc/xfer_kernels.cuh): either side may be page-locked host
 * memory (its device-visible address) with any byte alignment */
struct VbzkXferParams; /* == vbz::XferParams */
uint32_t vbzk_xfer_chunks_of(uint32_t max_bytes);
int vbzk_xfer(const struct VbzkXferParams* params, uint32_t ctas, void* stream);
/* control-plane moves by kernel instead of the copy engine (see csrc/xfer_kernels.cuh): 16-byte aligned pointers, bytes a multiple
 * of 4; src / dst may be page-locked host memory */
int vbzk_copy_words(void* dst, const void* src, uint64_t bytes, void* stream);
int vbzk_zero_words(void* dst, uint64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
