// oracle/cpu_bench.cpp - TEST/BENCH INFRASTRUCTURE ONLY (the CPU-baseline leg of bench.py).
//
// Times the reference's own C API (vbz_compress_sized / vbz_decompress_sized, vbz/vbz.h:111-131)
// on the host cores: one std::thread per core, reads pulled from an atomic counter, exactly one
// API call per read - the way MinKNOW / ont_fast5_api drive the reference (one HDF5 chunk per call,
// vbz_plugin/vbz_plugin.cpp:97-228).  The library under test is dlopen'ed (default
// oracle/_ref/libvbz_ref.so = the unmodified reference compiled by oracle/Makefile).
//
//   cpu_bench <lib.so> <raw.bin> <lens.i64> <zz> <isz> <lvl> <ver> <threads> <reps>
//
// raw.bin = the reads back to back; lens.i64 = int64 byte count per read.
// Prints one JSON object: best-of-reps encode / decode GB/s (raw bytes / wall time) and the ratio.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

struct Options { bool zz; unsigned isz, lvl, ver; };
typedef uint32_t (*xfn)(const void*, uint32_t, void*, uint32_t, const Options*);
typedef uint32_t (*mfn)(uint32_t, const Options*);

static std::vector<char> slurp(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> v((size_t)n);
    if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) { perror("read"); exit(2); }
    fclose(f);
    return v;
}

template <class F>
static double run_parallel(int threads, size_t n_items, F&& body)
{
    std::atomic<size_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n_items;) body(i); });
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv)
{
    if (argc < 10) { fprintf(stderr, "usage: %s lib raw lens zz isz lvl ver threads reps\n", argv[0]); return 2; }
    void* lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    xfn comp = (xfn)dlsym(lib, "vbz_compress_sized");
    xfn decomp = (xfn)dlsym(lib, "vbz_decompress_sized");
    mfn maxsz = (mfn)dlsym(lib, "vbz_max_compressed_size");
    if (!comp || !decomp || !maxsz) { fprintf(stderr, "missing symbols\n"); return 2; }
    std::vector<char> raw = slurp(argv[2]);
    std::vector<char> lens_b = slurp(argv[3]);
    size_t n = lens_b.size() / 8;
    const int64_t* lens = (const int64_t*)lens_b.data();
    Options o{atoi(argv[4]) != 0, (unsigned)atoi(argv[5]), (unsigned)atoi(argv[6]), (unsigned)atoi(argv[7])};
    int threads = atoi(argv[8]);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    int reps = atoi(argv[9]);

    std::vector<size_t> off(n + 1, 0), coff(n + 1, 0);
    for (size_t i = 0; i < n; ++i) {
        off[i + 1] = off[i] + (size_t)lens[i];
        coff[i + 1] = coff[i] + maxsz((uint32_t)lens[i], &o);
    }
    if (off[n] != raw.size()) { fprintf(stderr, "lens do not sum to raw size\n"); return 2; }
    std::vector<char> cbuf(coff[n]), back(raw.size());
    memset(cbuf.data(), 1, cbuf.size());  // pre-touch
    memset(back.data(), 1, back.size());
    std::vector<uint32_t> csz(n, 0);
    std::atomic<int> errors{0};
    double best_e = 1e30, best_d = 1e30;
    for (int r = 0; r < reps; ++r) {
        double te = run_parallel(threads, n, [&](size_t i) {
            uint32_t c = comp(raw.data() + off[i], (uint32_t)lens[i], cbuf.data() + coff[i], (uint32_t)(coff[i + 1] - coff[i]), &o);
            if (c >= 0xFFFFFFF9u) errors++;
            csz[i] = c;
        });
        double td = run_parallel(threads, n, [&](size_t i) {
            uint32_t d = decomp(cbuf.data() + coff[i], csz[i], back.data() + off[i], (uint32_t)lens[i], &o);
            if (d != (uint32_t)lens[i]) errors++;
        });
        best_e = std::min(best_e, te);
        best_d = std::min(best_d, td);
    }
    if (memcmp(back.data(), raw.data(), raw.size()) != 0) errors++;
    size_t ctotal = 0;
    for (size_t i = 0; i < n; ++i) ctotal += csz[i];
    printf("{\"reads\": %zu, \"raw_bytes\": %zu, \"compressed_bytes\": %zu, \"ratio\": %.6f, \"threads\": %d, \"reps\": %d, "
           "\"encode_s\": %.6f, \"decode_s\": %.6f, \"encode_gbs\": %.4f, \"decode_gbs\": %.4f, \"errors\": %d}\n",
           n, raw.size(), ctotal, raw.size() ? (double)ctotal / (double)raw.size() : 0.0, threads, reps, best_e, best_d,
           raw.size() / best_e / 1e9, raw.size() / best_d / 1e9, errors.load());
    return errors.load() ? 1 : 0;
}
