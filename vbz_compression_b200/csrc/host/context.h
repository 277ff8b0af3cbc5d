// csrc/host/context.h - per-device state of the host layer: a stream, grow-only scratch in HBM,
// pinned staging.  One context is used by one thread at a time.
#pragma once
#include <utility>
#include <initializer_list>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "runtime.h"

namespace vbz {

void set_last_error(const std::string& msg);
const char* get_last_error();
// returns false and records "<what>: <cuda error string>" when e != cudaSuccess
bool cuda_ok(cudaError_t e, const char* what);

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n);  // grow-only; contents are NOT preserved
    void release();
    template <class T> T* as() const { return static_cast<T*>(p); }
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n);
    void release();
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// Pinned staging for descriptor uploads.  An upload reads its block asynchronously, so a block may only be rewritten
// once the copy that used it has completed: kEntries blocks are handed out round robin, each guarded by the event of
// its last upload - with several waves in flight the event of 16 uploads ago has long fired and take() never blocks.
struct StageRing {
    static constexpr int kEntries = 16;
    PinBuf buf[kEntries];
    cudaEvent_t ev[kEntries] = {};
    int cur = 0;
    size_t high_water = 0;
    bool init();
    void* take(size_t n);          // nullptr on failure (last error set)
    bool mark(cudaStream_t s);     // call after enqueueing the copy that reads the block returned by take()
    void release();
};

// Everything the kernels of one wave (or one *_device call) scribble on, besides the caller's buffers.
struct ScratchSet {
    // descriptors and look-back state (stage 1)
    DevBuf d_reads, d_tiles, d_status, d_wtoff, d_issue, d_results, d_sizes;
    // stage 2 (zstd): descriptors, work counters, units / groups / sequence records of the decoder, literal scratch of the general kernel
    DevBuf d_zreads, d_zcounter, d_zlit, d_zblocks, d_zstatus, d_zcodes, d_zunits, d_zgroups, d_zgeneral, d_zseqs, d_zseqframes, d_zseqlist;
    DevBuf d_zelits, d_zeseqs, d_zemeta, d_zehist;  // the encoder's match finder (csrc/zstd_match_kernels.cuh)
    // intermediate stage-1 bytes of the two-stage paths
    DevBuf d_mid;
    void release_all()
    {
        for (DevBuf* b : {&d_reads, &d_tiles, &d_status, &d_wtoff, &d_issue, &d_results, &d_sizes, &d_zreads, &d_zcounter, &d_zlit, &d_zblocks, &d_zstatus, &d_zcodes,
                          &d_zunits, &d_zgroups, &d_zgeneral, &d_zseqs, &d_zseqframes, &d_zseqlist, &d_zelits, &d_zeseqs, &d_zemeta, &d_zehist, &d_mid})
            b->release();
    }
};

}  // namespace vbz

struct vbz_b200_ctx : vbz::ScratchSet {
    int device = 0;
    cudaStream_t stream = nullptr;
    vbz::StageRing stage;  // pinned descriptor staging
    uint64_t launches = 0;
    // optional CUDA-event timing of the stages of the last *_device call (vbz_b200_set_stage_timing)
    bool stage_timing = false, t_decode = false, t_two_stage = false;
    std::vector<cudaEvent_t> t_marks;  // stage boundaries of the last call, in stream order (begin, then two per wave)
    size_t t_used = 0;
    cudaStream_t last_zdecode_stream = nullptr;
    uint32_t codec_ctas_per_sm = 0;  // != 0 while a host batch runs: room for the transfer kernels (SvbParams::ctas_per_sm)
    vbz::PinBuf h_results;
    // more sets of scratch buffers and more kernel streams: the waves of a host batch take them in turn, so that the latency-bound
    // head and tail of one wave's kernel chain (a frame's sequences are serial work for one lane) overlap with the waves next to it
    // (csrc/host/batch.cpp: run_host_batch).  The base class holds the set in use; use_scratch(k) swaps set k in.
    static constexpr int kKernelStreams = 4;
    vbz::ScratchSet alt[kKernelStreams - 1];
    cudaStream_t kstream[kKernelStreams] = {};  // [0] == stream
    int scratch_in_use = 0;
    // the decoder's look-ahead (csrc/host/batch.cpp: decompress_device_impl): the walker and the sequence kernel of the NEXT wave of a
    // device-resident call run on kstream[1] beside the kernels that write this wave's output, so what they write exists twice
    struct ZDecSet { vbz::DevBuf d_zreads, d_zcounter, d_zunits, d_zgroups, d_zgeneral, d_zseqs, d_zseqframes, d_zseqlist, d_sizes; } zdec_alt;
    int zdec_in_use = 0;
    cudaEvent_t ev_front[2] = {}, ev_back[2] = {}, ev_fork = nullptr;
    void use_zdec(int k)
    {
        if (k == zdec_in_use) return;
        std::swap(d_zreads, zdec_alt.d_zreads); std::swap(d_zcounter, zdec_alt.d_zcounter); std::swap(d_zunits, zdec_alt.d_zunits);
        std::swap(d_zgroups, zdec_alt.d_zgroups); std::swap(d_zgeneral, zdec_alt.d_zgeneral); std::swap(d_zseqs, zdec_alt.d_zseqs);
        std::swap(d_zseqframes, zdec_alt.d_zseqframes); std::swap(d_zseqlist, zdec_alt.d_zseqlist); std::swap(d_sizes, zdec_alt.d_sizes);
        zdec_in_use = k;
    }
    void use_scratch(int k)
    {
        if (k == scratch_in_use) return;
        if (scratch_in_use != 0) std::swap(static_cast<vbz::ScratchSet&>(*this), alt[scratch_in_use - 1]);
        if (k != 0) std::swap(static_cast<vbz::ScratchSet&>(*this), alt[k - 1]);
        scratch_in_use = k;
    }

    // host-pointer API: a batch runs as WAVES through kWaveSlots slots so that the upload of wave w+1, the kernels of
    // wave w and the download of wave w-1 overlap (csrc/host/batch.cpp: run_host_batch)
    static constexpr int kWaveSlots = 7;
    struct WaveSlot {
        vbz::DevBuf d_in, d_out, d_results;
        vbz::DevBuf d_pack;         // host-contiguous runs of reads land here by DMA before they are spread to their slots
        vbz::DevBuf d_xin, d_xout;  // transfer records of the wave (csrc/xfer_kernels.cuh)
        vbz::PinBuf h_in, h_out, h_results;
        // upload done; kernels + results done; download by kernel done; download of the staged part done
        cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_out = nullptr, ev_outs = nullptr;
    };
    WaveSlot slots[kWaveSlots];
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_d2hs = nullptr;

    ~vbz_b200_ctx();
};
