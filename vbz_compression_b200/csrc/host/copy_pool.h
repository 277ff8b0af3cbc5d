// csrc/host/copy_pool.h - a few persistent host threads that move bytes between the caller's (pageable)
// buffers and the pinned staging of a wave.  One memcpy stream tops out near 10 GB/s; the PCIe link takes
// five times that, so the staging copies of a batch are cut into pieces and spread over the pool.
// Sample bytes are only MOVED here - every transform runs on the device.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__) && !defined(VBZ_CPUSIM)
#include <emmintrin.h>
#define VBZ_STREAM_COPY 1
#endif

namespace vbz {

struct CopyTask { void* dst; const void* src; size_t n; };

// memcpy for pieces that are written once and not read again by this core: non-temporal 16-byte stores skip the
// read-for-ownership of the destination lines (a third of a plain copy's memory traffic), which is what bounds the staged
// path (the copy threads together move the batch at ~40 GB/s on the B200 host).
static inline void stream_copy(void* dst, const void* src, size_t n)
{
#ifdef VBZ_STREAM_COPY
    if (n >= (64u << 10)) {
        uint8_t* d = static_cast<uint8_t*>(dst);
        const uint8_t* s = static_cast<const uint8_t*>(src);
        const size_t head = (16 - ((uintptr_t)d & 15)) & 15;
        if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
        const size_t blocks = n / 64;
        for (size_t i = 0; i < blocks; ++i) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16)),
                          c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
            s += 64;
            d += 64;
        }
        _mm_sfence();
        memcpy(d, s, n - blocks * 64);
        return;
    }
#endif
    memcpy(dst, src, n);
}

class CopyPool {
public:
    static CopyPool& instance()
    {
        static CopyPool* p = new CopyPool();  // leaked on purpose: worker threads must not be joined from a static destructor
        return *p;
    }
    // runs every task; returns when all of them are done.  Safe to call from several threads (calls are serialised).
    void run(std::vector<CopyTask>& tasks)
    {
        size_t total = 0;
        for (const CopyTask& t : tasks) total += t.n;
        if (workers_.empty() || total < (4u << 20)) {  // small batches: not worth a hand-off
            for (const CopyTask& t : tasks) if (t.n) stream_copy(t.dst, t.src, t.n);
            return;
        }
        // cut large copies so that the pieces balance
        std::vector<CopyTask> pieces;
        const size_t kPiece = 2u << 20;
        for (const CopyTask& t : tasks)
            for (size_t o = 0; o < t.n; o += kPiece)
                pieces.push_back({(uint8_t*)t.dst + o, (const uint8_t*)t.src + o, t.n - o < kPiece ? t.n - o : kPiece});
        std::unique_lock<std::mutex> call(call_mu_);
        {
            std::lock_guard<std::mutex> g(mu_);
            tasks_ = &pieces;
            next_.store(0);
            pending_ = (unsigned)workers_.size();
            ++generation_;
        }
        cv_.notify_all();
        work(pieces);  // the caller helps
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [&] { return pending_ == 0; });
        tasks_ = nullptr;
    }
    unsigned threads() const { return (unsigned)workers_.size() + 1; }

private:
    CopyPool()
    {
        unsigned n = std::thread::hardware_concurrency();
        n = n >= 16 ? 12 : (n >= 4 ? n / 2 : 1);  // measured on the 16-core B200 host: 4 threads 17, 8 threads 23, 16 threads 25 GB/s end to end
        if (const char* e = getenv("VBZ_B200_COPY_THREADS")) n = (unsigned)atoi(e);
        if (n < 1) n = 1;
        if (n > 32) n = 32;
        for (unsigned i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
        for (std::thread& t : workers_) t.detach();
    }
    void work(std::vector<CopyTask>& pieces)
    {
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= pieces.size()) break;
            const CopyTask& t = pieces[i];
            stream_copy(t.dst, t.src, t.n);
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            std::vector<CopyTask>* mine;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return generation_ != seen; });
                seen = generation_;
                mine = tasks_;
            }
            if (mine) work(*mine);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<CopyTask>* tasks_ = nullptr;
    std::atomic<size_t> next_{0};
    unsigned pending_ = 0;
    uint64_t generation_ = 0;
};

}  // namespace vbz
