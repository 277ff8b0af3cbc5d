// oracle/cpu_bench.cpp - TEST/BENCH INFRASTRUCTURE ONLY (the CPU-baseline leg of bench.py).
//
// Times the reference's own C API (vbz_compress_sized / vbz_decompress_sized, vbz/vbz.h:111-131)
// on the host cores: one std::thread per core, reads pulled from an atomic counter, exactly one
// API call per read - the way MinKNOW / ont_fast5_api drive the reference (one HDF5 chunk per call,
// vbz_plugin/vbz_plugin.cpp:97-228).  The library under test is dlopen'ed (default
// oracle/_ref/libvbz_ref.so = the unmodified reference compiled by oracle/Makefile).
//
//   cpu_bench <lib.so> <raw.bin> <lens.i64> <zz> <isz> <lvl> <ver> <threads> <reps> [dump_prefix]
//   cpu_bench <lib.so> --decode <chunks.bin> <csizes.u32> <raw.bin> <lens.i64> <zz> <isz> <lvl> <ver> <threads> <reps>
//
// raw.bin = the reads back to back; lens.i64 = int64 byte count per read.
// Prints one JSON object: best-of-reps encode / decode GB/s (raw bytes / wall time) and the ratio.
// dump_prefix: also write the chunks the reference produced (<prefix>.bin, back to back, and <prefix>.sizes.u32) - the
// "reference-written H5Z chunk buffers" of bench.py --workload h5z-decode.
// --decode: only decode the given chunks ([u32 size][payload] each, back to back; whoever wrote them) with the library
// and compare with raw.bin: the cross-decode check of GPU-written frames, and the CPU arm of the decode-only workload.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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

static int decode_only(void* lib, int argc, char** argv);

int main(int argc, char** argv)
{
    if (argc >= 3 && strcmp(argv[2], "--decode") == 0) {
        if (argc < 13) { fprintf(stderr, "usage: %s lib --decode chunks csizes raw lens zz isz lvl ver threads reps\n", argv[0]); return 2; }
        void* l = dlopen(argv[1], RTLD_NOW);
        if (!l) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
        return decode_only(l, argc, argv);
    }
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
    if (argc >= 11 && !errors.load()) {  // the chunks as the reference wrote them
        std::string pre = argv[10];
        FILE* fb = fopen((pre + ".bin").c_str(), "wb");
        FILE* fs = fopen((pre + ".sizes.u32").c_str(), "wb");
        if (!fb || !fs) { perror("dump"); return 2; }
        for (size_t i = 0; i < n; ++i) fwrite(cbuf.data() + coff[i], 1, csz[i], fb);
        fwrite(csz.data(), 4, n, fs);
        fclose(fb);
        fclose(fs);
    }
    printf("{\"reads\": %zu, \"raw_bytes\": %zu, \"compressed_bytes\": %zu, \"ratio\": %.6f, \"threads\": %d, \"reps\": %d, "
           "\"encode_s\": %.6f, \"decode_s\": %.6f, \"encode_gbs\": %.4f, \"decode_gbs\": %.4f, \"errors\": %d}\n",
           n, raw.size(), ctotal, raw.size() ? (double)ctotal / (double)raw.size() : 0.0, threads, reps, best_e, best_d,
           raw.size() / best_e / 1e9, raw.size() / best_d / 1e9, errors.load());
    return errors.load() ? 1 : 0;
}

static int decode_only(void* lib, int argc, char** argv)
{
    (void)argc;
    xfn decomp = (xfn)dlsym(lib, "vbz_decompress_sized");
    if (!decomp) { fprintf(stderr, "missing symbols\n"); return 2; }
    std::vector<char> chunks = slurp(argv[3]);
    std::vector<char> cs_b = slurp(argv[4]);
    std::vector<char> raw = slurp(argv[5]);
    std::vector<char> lens_b = slurp(argv[6]);
    const size_t n = lens_b.size() / 8;
    if (cs_b.size() / 4 != n) { fprintf(stderr, "csizes and lens disagree\n"); return 2; }
    const int64_t* lens = (const int64_t*)lens_b.data();
    const uint32_t* csz = (const uint32_t*)cs_b.data();
    Options o{atoi(argv[7]) != 0, (unsigned)atoi(argv[8]), (unsigned)atoi(argv[9]), (unsigned)atoi(argv[10])};
    int threads = atoi(argv[11]);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    const int reps = atoi(argv[12]);
    std::vector<size_t> off(n + 1, 0), coff(n + 1, 0);
    for (size_t i = 0; i < n; ++i) { off[i + 1] = off[i] + (size_t)lens[i]; coff[i + 1] = coff[i] + csz[i]; }
    if (off[n] != raw.size() || coff[n] != chunks.size()) { fprintf(stderr, "sizes do not sum to the file sizes\n"); return 2; }
    std::vector<char> back(raw.size());
    memset(back.data(), 1, back.size());
    std::atomic<int> errors{0};
    double best_d = 1e30;
    for (int r = 0; r < reps; ++r) {
        double td = run_parallel(threads, n, [&](size_t i) {
            uint32_t d = decomp(chunks.data() + coff[i], csz[i], back.data() + off[i], (uint32_t)lens[i], &o);
            if (d != (uint32_t)lens[i]) errors++;
        });
        best_d = std::min(best_d, td);
    }
    if (memcmp(back.data(), raw.data(), raw.size()) != 0) errors++;
    printf("{\"reads\": %zu, \"raw_bytes\": %zu, \"compressed_bytes\": %zu, \"ratio\": %.6f, \"threads\": %d, \"reps\": %d, "
           "\"decode_s\": %.6f, \"decode_gbs\": %.4f, \"errors\": %d}\n",
           n, raw.size(), chunks.size(), raw.size() ? (double)chunks.size() / (double)raw.size() : 0.0, threads, reps, best_d,
           raw.size() / best_d / 1e9, errors.load());
    return errors.load() ? 1 : 0;
}
