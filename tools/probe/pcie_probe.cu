// PCIe probe: how fast do per-read cudaMemcpyAsync calls move a batch, next to one big copy?
#include <cuda_runtime.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> u(std::log(1e4), std::log(1e6));
    std::vector<size_t> sz, off;
    size_t tot = 0;
    while (tot < (2ull << 30)) { size_t s = (size_t)std::exp(u(rng)) * 2; off.push_back(tot); sz.push_back(s); tot += (s + 15) & ~15ull; }
    printf("%zu reads, %.2f GiB\n", sz.size(), tot / 1073741824.0);
    char *h, *d, *h2, *d2;
    CK(cudaMallocHost(&h, tot)); CK(cudaMalloc(&d, tot)); CK(cudaMallocHost(&h2, tot)); CK(cudaMalloc(&d2, tot));
    cudaStream_t s[4]; for (auto& x : s) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    for (int rep = 0; rep < 2; ++rep) {
        double t0 = now(); CK(cudaMemcpyAsync(d, h, tot, cudaMemcpyHostToDevice, s[0])); CK(cudaStreamSynchronize(s[0])); double t1 = now();
        printf("one copy h2d: %.1f ms %.1f GB/s\n", (t1 - t0) * 1e3, tot / (t1 - t0) / 1e9);
        t0 = now(); for (size_t i = 0; i < sz.size(); ++i) CK(cudaMemcpyAsync(d + off[i], h + off[i], sz[i], cudaMemcpyHostToDevice, s[0]));
        double ti = now(); CK(cudaStreamSynchronize(s[0])); t1 = now();
        printf("per-read h2d, 1 stream: issue %.1f ms total %.1f ms %.1f GB/s\n", (ti - t0) * 1e3, (t1 - t0) * 1e3, tot / (t1 - t0) / 1e9);
        t0 = now(); for (size_t i = 0; i < sz.size(); ++i) CK(cudaMemcpyAsync(d + off[i], h + off[i], sz[i], cudaMemcpyHostToDevice, s[i & 1]));
        CK(cudaStreamSynchronize(s[0])); CK(cudaStreamSynchronize(s[1])); t1 = now();
        printf("per-read h2d, 2 streams: total %.1f ms %.1f GB/s\n", (t1 - t0) * 1e3, tot / (t1 - t0) / 1e9);
        t0 = now(); for (size_t i = 0; i < sz.size(); ++i) CK(cudaMemcpyAsync(h2 + off[i], d2 + off[i], sz[i], cudaMemcpyDeviceToHost, s[2]));
        CK(cudaStreamSynchronize(s[2])); t1 = now();
        printf("per-read d2h, 1 stream: total %.1f ms %.1f GB/s\n", (t1 - t0) * 1e3, tot / (t1 - t0) / 1e9);
        t0 = now();
        for (size_t i = 0; i < sz.size(); ++i) {
            CK(cudaMemcpyAsync(d + off[i], h + off[i], sz[i], cudaMemcpyHostToDevice, s[0]));
            CK(cudaMemcpyAsync(h2 + off[i], d2 + off[i], sz[i] * 5 / 8, cudaMemcpyDeviceToHost, s[2]));
        }
        CK(cudaStreamSynchronize(s[0])); CK(cudaStreamSynchronize(s[2])); t1 = now();
        printf("per-read h2d + 0.64x d2h at once: total %.1f ms %.1f GB/s up\n", (t1 - t0) * 1e3, tot / (t1 - t0) / 1e9);
        t0 = now();
        {
            size_t nexto = 0;
            for (size_t i = 0; i < sz.size(); ++i) {
                while (nexto <= off[i] && nexto < tot) { CK(cudaMemcpyAsync(d + nexto, h + nexto, std::min<size_t>(32u << 20, tot - nexto), cudaMemcpyHostToDevice, s[0])); nexto += 32u << 20; }
                CK(cudaMemcpyAsync(h2 + off[i], d2 + off[i], sz[i] * 5 / 8, cudaMemcpyDeviceToHost, s[2]));
            }
        }
        double tq = now();
        CK(cudaStreamSynchronize(s[0])); double tu = now(); CK(cudaStreamSynchronize(s[2])); t1 = now();
        printf("32 MiB h2d pieces + per-read 0.64x d2h at once: issue %.1f  h2d done %.1f  all done %.1f ms\n", (tq - t0) * 1e3, (tu - t0) * 1e3, (t1 - t0) * 1e3);
        t0 = now();
        {
            std::thread th([&] { for (size_t i = 0; i < sz.size(); ++i) CK(cudaMemcpyAsync(h2 + off[i], d2 + off[i], sz[i] * 5 / 8, cudaMemcpyDeviceToHost, s[2])); CK(cudaStreamSynchronize(s[2])); });
            for (size_t i = 0; i < sz.size(); ++i) CK(cudaMemcpyAsync(d + off[i], h + off[i], sz[i], cudaMemcpyHostToDevice, s[0]));
            CK(cudaStreamSynchronize(s[0]));
            th.join();
        }
        t1 = now();
        printf("per-read h2d + 0.64x d2h, one issuing thread each: total %.1f ms\n", (t1 - t0) * 1e3);
        // waves: 32 MiB contiguous pieces
        t0 = now(); for (size_t o = 0; o < tot; o += 32u << 20) CK(cudaMemcpyAsync(d + o, h + o, std::min<size_t>(32u << 20, tot - o), cudaMemcpyHostToDevice, s[0]));
        CK(cudaStreamSynchronize(s[0])); t1 = now();
        printf("32 MiB pieces h2d: total %.1f ms %.1f GB/s\n", (t1 - t0) * 1e3, tot / (t1 - t0) / 1e9);
    }
    return 0;
}
