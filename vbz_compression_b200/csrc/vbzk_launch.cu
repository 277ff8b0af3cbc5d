// csrc/vbzk_launch.cu - kernel instantiations + the C-ABI launch shim (csrc/vbzk.h).
#include "vbzk.h"
#include "svb_kernels.cuh"
#include "svb16_kernels.cuh"
#include "svb_half_kernels.cuh"
#include "zstd_dec_kernels.cuh"
#include "zstd_enc_kernels.cuh"
#include "xfer_kernels.cuh"

using namespace vbz;

#ifdef VBZ_CPUSIM
#define VBZK_SET_SMEM(kernel, bytes) 0
#define VBZK_LAST_ERROR() 0
#else
#define VBZK_SET_SMEM(kernel, bytes) \
    ((bytes) > 48 * 1024 ? (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)) : 0)
#define VBZK_LAST_ERROR() ((int)cudaGetLastError())
#endif

extern "C" uint32_t vbzk_svb_tile_values(int isz, int mode) { return svb_tile_values(isz, mode); }

template <int ISZ, int MODE>
static uint32_t smem_of(int decode) { return decode ? SvbSmem<ISZ, MODE>::kDecodeBytes : SvbSmem<ISZ, MODE>::kEncodeBytes; }

extern "C" uint32_t vbzk_svb_smem_bytes(int isz, int mode, int decode)
{
    if (isz == 2 && mode == kModeI16X86) return decode ? S16DecSmem::kBytes : S16EncSmem::kBytes;
    if (isz == 2) return smem_of<2, kModeGeneric>(decode);
    if (isz == 4) return smem_of<4, kModeGeneric>(decode);
    return smem_of<1, kModeGeneric>(decode);
}

extern "C" int vbzk_svb_fill_tiles(const VbzkSvbRead* d_reads, uint32_t n_reads, uint32_t tile_values, uint32_t* d_tile_read, void* stream)
{
    if (n_reads == 0) return 0;
    const uint32_t warps_per_cta = 8;
    const uint32_t grid = (n_reads + warps_per_cta - 1) / warps_per_cta;
    VBZ_LAUNCH(svb_fill_tiles_kernel, grid, warps_per_cta * 32, 0, (vbz_stream_t)stream,
               reinterpret_cast<const SvbRead*>(d_reads), n_reads, tile_values, d_tile_read);
    return VBZK_LAST_ERROR();
}

template <int ISZ, bool ZZ, int MODE>
static int launch_encode(const SvbParams& p, uint32_t n_tiles, void* stream)
{
    const uint32_t smem = SvbSmem<ISZ, MODE>::kEncodeBytes;
    if (int e = VBZK_SET_SMEM((svb_encode_kernel<ISZ, ZZ, MODE>), smem)) return e;
    VBZ_LAUNCH((svb_encode_kernel<ISZ, ZZ, MODE>), n_tiles, kSvbThreads, smem, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}
template <int ISZ, bool ZZ, int MODE>
static int launch_decode(const SvbParams& p, uint32_t n_tiles, void* stream)
{
    const uint32_t smem = SvbSmem<ISZ, MODE>::kDecodeBytes;
    if (int e = VBZK_SET_SMEM((svb_decode_kernel<ISZ, ZZ, MODE>), smem)) return e;
    VBZ_LAUNCH((svb_decode_kernel<ISZ, ZZ, MODE>), n_tiles, kSvbThreads, smem, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}

template <int ISZ, int MODE>
static int launch_key_offsets(const SvbParams& p, void* stream)
{
    const uint32_t grid = (p.n_reads + kSvbWarps - 1) / kSvbWarps;
    VBZ_LAUNCH((svb_key_offsets_kernel<ISZ, MODE>), grid, kSvbThreads, 0, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}
// ---- int16 x86-semantics path (csrc/svb16_kernels.cuh): persistent CTAs, as many as are co-resident -----------
template <class K>
static uint32_t persistent_grid(K kernel, int threads, uint32_t smem, uint32_t n_tiles, uint32_t ctas_per_sm)
{
#ifdef VBZ_CPUSIM
    (void)kernel; (void)threads; (void)smem; (void)ctas_per_sm;
    return n_tiles < 3 ? n_tiles : 3;
#else
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || sms <= 0 || per_sm <= 0)
        return n_tiles < 148 ? n_tiles : 148;
    if (ctas_per_sm && (uint32_t)per_sm > ctas_per_sm) per_sm = (int)ctas_per_sm;
    const uint32_t g = (uint32_t)sms * (uint32_t)per_sm;
    return n_tiles < g ? n_tiles : g;
#endif
}
// Cooperative launch: the persistent CTAs wait for each other's tiles (look-back), so all of them must be resident.
template <class K>
static int launch_persistent(K kernel, uint32_t smem, const SvbParams& p, uint32_t n_tiles, void* stream)
{
    if (int e = VBZK_SET_SMEM(kernel, smem)) return e;
    const uint32_t grid = persistent_grid(kernel, kS16Threads, smem, n_tiles, p.ctas_per_sm);
#ifdef VBZ_CPUSIM
    VBZ_LAUNCH(kernel, grid, kS16Threads, smem, (vbz_stream_t)stream, p);
    return 0;
#else
    SvbParams copy = p;
    void* args[] = {&copy};
    return (int)cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(kS16Threads), args, smem, (cudaStream_t)stream);
#endif
}
static int launch_s16_encode(const SvbParams& p, uint32_t n_tiles, void* stream)
{
    VBZ_LAUNCH(svb16_enc_prepare_kernel, (p.n_reads + kS16Warps - 1) / kS16Warps, kS16Threads, 0, (vbz_stream_t)stream, p);
    if (int e = VBZK_LAST_ERROR()) return e;
    return launch_persistent(svb16_encode_kernel, S16EncSmem::kBytes, p, n_tiles, stream);
}
static int launch_s16_decode(const SvbParams& p, uint32_t n_tiles, void* stream)
{
    return launch_persistent(svb16_decode_kernel, S16DecSmem::kBytes, p, n_tiles, stream);
}
static int launch_s16_offsets(const SvbParams& p, void* stream)
{
    VBZ_LAUNCH(svb16_tile_sums_kernel, (p.n_tiles + kS16Warps - 1) / kS16Warps, kS16Threads, 0, (vbz_stream_t)stream, p);
    if (int e = VBZK_LAST_ERROR()) return e;
    VBZ_LAUNCH(svb16_tile_offsets_kernel, (p.n_reads + kS16Warps - 1) / kS16Warps, kS16Threads, 0, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}

extern "C" int vbzk_svb_key_offsets(int isz, int mode, const VbzkSvbParams* params, void* stream)
{
    const SvbParams& p = *reinterpret_cast<const SvbParams*>(params);
    if (p.n_reads == 0) return 0;
    if (isz == 2 && mode == kModeI16X86) return launch_s16_offsets(p, stream);
    if (isz == 2) return launch_key_offsets<2, kModeGeneric>(p, stream);
    if (isz == 4) return launch_key_offsets<4, kModeGeneric>(p, stream);
    if (isz == 1) return launch_key_offsets<1, kModeGeneric>(p, stream);
    return -1;
}

#define VBZK_DISPATCH(fn)                                                                            \
    const SvbParams& p = *reinterpret_cast<const SvbParams*>(params);                                \
    if (n_tiles == 0) return 0;                                                                      \
    if (isz == 2) return zz ? fn<2, true, kModeGeneric>(p, n_tiles, stream) : fn<2, false, kModeGeneric>(p, n_tiles, stream); \
    if (isz == 4) return zz ? fn<4, true, kModeGeneric>(p, n_tiles, stream) : fn<4, false, kModeGeneric>(p, n_tiles, stream); \
    if (isz == 1) return zz ? fn<1, true, kModeGeneric>(p, n_tiles, stream) : fn<1, false, kModeGeneric>(p, n_tiles, stream); \
    return -1;

extern "C" int vbzk_svb_encode(int isz, int zz, int mode, const VbzkSvbParams* params, uint32_t n_tiles, void* stream)
{
    if (isz == 2 && zz && mode == kModeI16X86 && n_tiles) return launch_s16_encode(*reinterpret_cast<const SvbParams*>(params), n_tiles, stream);
    VBZK_DISPATCH(launch_encode)
}
extern "C" int vbzk_svb_decode(int isz, int zz, int mode, const VbzkSvbParams* params, uint32_t n_tiles, void* stream)
{
    if (isz == 2 && zz && mode == kModeI16X86 && n_tiles) return launch_s16_decode(*reinterpret_cast<const SvbParams*>(params), n_tiles, stream);
    VBZK_DISPATCH(launch_decode)
}

// ---- v1 "half" packing (int8): one warp per read -----------------------------------------------------------------
extern "C" int vbzk_svb_half_encode(int zz, const VbzkSvbParams* params, void* stream)
{
    const SvbParams& p = *reinterpret_cast<const SvbParams*>(params);
    if (p.n_reads == 0) return 0;
    const uint32_t grid = (p.n_reads + kHalfWarps - 1) / kHalfWarps, smem = kHalfWarps * kHalfStage;
    if (zz) VBZ_LAUNCH(svb_half_encode_kernel<true>, grid, kHalfWarps * 32, smem, (vbz_stream_t)stream, p);
    else VBZ_LAUNCH(svb_half_encode_kernel<false>, grid, kHalfWarps * 32, smem, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}
extern "C" int vbzk_svb_half_decode(int zz, const VbzkSvbParams* params, void* stream)
{
    const SvbParams& p = *reinterpret_cast<const SvbParams*>(params);
    if (p.n_reads == 0) return 0;
    const uint32_t grid = (p.n_reads + kHalfWarps - 1) / kHalfWarps, smem = kHalfWarps * kHalfStage;
    if (zz) VBZ_LAUNCH(svb_half_decode_kernel<true>, grid, kHalfWarps * 32, smem, (vbz_stream_t)stream, p);
    else VBZ_LAUNCH(svb_half_decode_kernel<false>, grid, kHalfWarps * 32, smem, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}

// ---- stage 2 (zstd) ---------------------------------------------------------------------------------------------
extern "C" uint32_t vbzk_zstd_decode_grid(void)
{
    static const uint32_t g = persistent_grid(zstd_decode_kernel, kZdecWarps * 32, (uint32_t)(sizeof(ZWarp) * kZdecWarps), 0xFFFFFFFFu, 0);
    return g;
}
extern "C" uint64_t vbzk_zstd_decode_scratch_bytes(void) { return (uint64_t)vbzk_zstd_decode_grid() * kZdecWarps * kZLitStride; }
// The decode of a batch in two halves (csrc/host/batch.cpp runs the first half of the next wave beside the second half of this one):
//   front  walker + sequence kernel: reads the frames, writes units / groups / records / positions / per-frame results - nothing of the output
//   back   units kernel + executor + general kernel: writes the output
extern "C" int vbzk_zstd_decode_front(const VbzkZstdDecParams* params, void* stream)
{
    const ZstdDecParams& p = *reinterpret_cast<const ZstdDecParams*>(params);
    if (p.n_reads == 0) return 0;
    // 1. frames -> units (or the general kernel's list)
    VBZ_LAUNCH(zstd_walk_kernel, (p.n_reads + 127) / 128, 128, 0, (vbz_stream_t)stream, p);
    if (int e = VBZK_LAST_ERROR()) return e;
    // 2. frames with sequences: decode the sequences into records, which gives every unit of those frames its position
    const uint32_t ssmem = (uint32_t)(sizeof(ZSeqWarp) * kZSeqWarps);
    if (int e = VBZK_SET_SMEM(zstd_seq_kernel, ssmem)) return e;
    static const uint32_t sgrid_max = persistent_grid(zstd_seq_kernel, kZSeqWarps * 32, ssmem, 0xFFFFFFFFu, 0);
    const uint32_t sneed = p.n_reads < (1u << 24) ? (p.n_reads * 16u + kZSeqWarps - 1) / kZSeqWarps : 0xFFFFFFFFu;  // a warp per chain: a frame brings a handful (one per block with its own tables)
    VBZ_LAUNCH(zstd_seq_kernel, sneed < sgrid_max ? sneed : sgrid_max, kZSeqWarps * 32, ssmem, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}
extern "C" int vbzk_zstd_decode_back(const VbzkZstdDecParams* params, void* stream)
{
    const ZstdDecParams& p = *reinterpret_cast<const ZstdDecParams*>(params);
    if (p.n_reads == 0) return 0;
    // 3. the groups of units, a CTA of 64 threads each from a work counter; the grid is sized for the most there can be
    const uint32_t usmem = (uint32_t)sizeof(ZGroupSmem);
    if (int e = VBZK_SET_SMEM(zstd_units_kernel, usmem)) return e;
    static const uint32_t ugrid_max = persistent_grid(zstd_units_kernel, kZUnitThreads, usmem, 0xFFFFFFFFu, 0);
    // (ZstdDecParams::pad = per cent of the persistent grid to use, 0 = all: with the next wave's walker and sequence kernel waiting on a second
    // stream, a grid that fills every SM to the last thread keeps them out until it retires - csrc/host/batch.cpp: decompress_device_impl)
    const uint32_t ucap = p.pad > 0 && p.pad < 100 ? (uint32_t)((uint64_t)ugrid_max * p.pad / 100) : ugrid_max;
    uint32_t ugrid = p.unit_cap < ucap ? p.unit_cap : ucap;
    if (ugrid) VBZ_LAUNCH(zstd_units_kernel, ugrid, kZUnitThreads, usmem, (vbz_stream_t)stream, p);
    if (int e = VBZK_LAST_ERROR()) return e;
    // 4. frames with sequences: the records of their Seq units, a warp per chain (the literals are in place, and so is everything the units kernel wrote)
    {
        static const uint32_t xgrid_max = persistent_grid(zstd_seqexec_kernel, kZExecWarps * 32, 0, 0xFFFFFFFFu, 0);
        const uint64_t xneed = ((uint64_t)p.unit_cap + kZExecWarps - 1) / kZExecWarps;
        VBZ_LAUNCH(zstd_seqexec_kernel, (uint32_t)(xneed < xgrid_max ? xneed : xgrid_max), kZExecWarps * 32, 0, (vbz_stream_t)stream, p);
    }
    if (int e = VBZK_LAST_ERROR()) return e;
    // 5. whatever the walker and the sequence kernel left over, a warp per frame
    const uint32_t smem = (uint32_t)(sizeof(ZWarp) * kZdecWarps);
    if (int e = VBZK_SET_SMEM(zstd_decode_kernel, smem)) return e;
    uint32_t grid = vbzk_zstd_decode_grid();
    const uint32_t need = (p.n_reads + kZdecWarps - 1) / kZdecWarps;
    if (need < grid) grid = need;
    VBZ_LAUNCH(zstd_decode_kernel, grid, kZdecWarps * 32, smem, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}
extern "C" int vbzk_zstd_decode(const VbzkZstdDecParams* params, void* stream)
{
    if (int e = vbzk_zstd_decode_front(params, stream)) return e;
    return vbzk_zstd_decode_back(params, stream);
}

extern "C" uint32_t vbzk_zstd_enc_block_bytes(void) { return kZEncBlock; }
extern "C" uint32_t vbzk_zstd_enc_key_block_bytes(void) { return kZEncBlockKeys; }
extern "C" int vbzk_zstd_enc_fill_blocks(const VbzkZstdEncRead* d_reads, uint32_t n_reads, uint32_t* d_block_read, void* stream)
{
    if (n_reads == 0) return 0;
    VBZ_LAUNCH(zstd_enc_fill_blocks_kernel, (n_reads + 7) / 8, 256, 0, (vbz_stream_t)stream, reinterpret_cast<const ZstdEncRead*>(d_reads), n_reads, d_block_read);
    return VBZK_LAST_ERROR();
}
extern "C" int vbzk_zstd_encode(const VbzkZstdEncParams* params, void* stream)
{
    const ZstdEncParams& p = *reinterpret_cast<const ZstdEncParams*>(params);
    if (p.n_blocks == 0) return 0;
    // matches first: they decide which bytes are literals (csrc/zstd_match_kernels.cuh), then the sequences sections of the S groups
    const uint32_t msmem = (uint32_t)(sizeof(ZMatchSmem) * kZMatchWarps);
    if (int e = VBZK_SET_SMEM(zstd_enc_match_kernel, msmem)) return e;
    VBZ_LAUNCH(zstd_enc_match_kernel, (p.n_blocks + kZMatchWarps - 1) / kZMatchWarps, kZMatchWarps * 32, msmem, (vbz_stream_t)stream, p);
    if (int e = VBZK_LAST_ERROR()) return e;
    const uint32_t qsmem = (uint32_t)(sizeof(ZSeqEncSmem) * kZSeqEncWarps);
    if (int e = VBZK_SET_SMEM(zstd_enc_seq_kernel, qsmem)) return e;
    VBZ_LAUNCH(zstd_enc_seq_kernel, (p.n_blocks + kZSeqEncWarps - 1) / kZSeqEncWarps, kZSeqEncWarps * 32, qsmem, (vbz_stream_t)stream, p);
    if (int e = VBZK_LAST_ERROR()) return e;
    const uint32_t csmem = (uint32_t)(sizeof(ZCode) * kZcodeWarps);
    if (int e = VBZK_SET_SMEM(zstd_enc_code_kernel, csmem)) return e;
    VBZ_LAUNCH(zstd_enc_code_kernel, (p.n_blocks + kZcodeWarps - 1) / kZcodeWarps, kZcodeWarps * 32, csmem, (vbz_stream_t)stream, p);
    if (int e = VBZK_LAST_ERROR()) return e;
    const uint32_t smem = (uint32_t)sizeof(ZEncSmem);
    if (int e = VBZK_SET_SMEM(zstd_encode_kernel, smem)) return e;
    VBZ_LAUNCH(zstd_encode_kernel, p.n_blocks, kZencThreads, smem, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}

// ---- host <-> device transfers of a wave by kernel (csrc/xfer_kernels.cuh) ----------------------------------------------
extern "C" uint32_t vbzk_xfer_chunks_of(uint32_t n) { return xfer_chunks_of(n); }
extern "C" int vbzk_xfer(const VbzkXferParams* params, uint32_t ctas, void* stream)
{
    const XferParams& p = *reinterpret_cast<const XferParams*>(params);
    if (p.n_reads == 0 || p.n_chunks == 0) return 0;
    const uint32_t need = (p.n_chunks + kXferThreads / 32 - 1) / (kXferThreads / 32);
    VBZ_LAUNCH(xfer_kernel, need < ctas ? need : ctas, kXferThreads, 0, (vbz_stream_t)stream, p);
    return VBZK_LAST_ERROR();
}
static uint32_t words_grid(uint64_t n_words)
{
    const uint64_t g = (n_words / 4 + 255) / 256;
    return (uint32_t)(g < 1 ? 1 : (g > 1184 ? 1184 : g));
}
extern "C" int vbzk_copy_words(void* dst, const void* src, uint64_t bytes, void* stream)
{
    if (bytes == 0) return 0;
    VBZ_LAUNCH(words_copy_kernel, words_grid(bytes / 4), 256, 0, (vbz_stream_t)stream, (uint32_t*)dst, (const uint32_t*)src, bytes / 4);
    return VBZK_LAST_ERROR();
}
extern "C" int vbzk_zero_words(void* dst, uint64_t bytes, void* stream)
{
    if (bytes == 0) return 0;
    VBZ_LAUNCH(words_zero_kernel, words_grid(bytes / 4), 256, 0, (vbz_stream_t)stream, (uint32_t*)dst, bytes / 4);
    return VBZK_LAST_ERROR();
}
