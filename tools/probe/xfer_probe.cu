// Probe: PCIe throughput of the kernel-driven transfer (csrc/xfer_kernels.cuh) for a batch of reads, per grid size,
// one direction and both at once; checks the bytes.
#include <cuda_runtime.h>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "../../vbz_compression_b200/csrc/xfer_kernels.cuh"
using namespace vbz;
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> u(std::log(1e4), std::log(1e6));
    std::vector<size_t> sz, hoff, doff;
    size_t htot = 0, dtot = 0;
    while (htot < (1ull << 30)) { size_t s = (size_t)std::exp(u(rng)) * 2; hoff.push_back(htot + 2); doff.push_back(dtot); sz.push_back(s); htot += s + 2; dtot += (s + 15) & ~15ull; }
    const uint32_t n = (uint32_t)sz.size();
    printf("%u reads, %.2f GiB\n", n, htot / 1073741824.0);
    uint8_t *h, *d, *h2, *d2;
    CK(cudaMallocHost(&h, htot + 64)); CK(cudaMalloc(&d, dtot + 64)); CK(cudaMallocHost(&h2, htot + 64)); CK(cudaMalloc(&d2, dtot + 64));
    for (size_t i = 0; i < htot; i += 1) h[i] = (uint8_t)(i * 2654435761u >> 13);
    memset(h2, 0, htot);
    std::vector<XferRead> up(n), down(n);
    uint32_t chunks = 0;
    for (uint32_t i = 0; i < n; ++i) {
        up[i] = {(uint64_t)(h + hoff[i]), (uint64_t)(d + doff[i]), (uint32_t)sz[i], i, chunks, 0};
        down[i] = {(uint64_t)(d + doff[i]), (uint64_t)(h2 + hoff[i]), (uint32_t)sz[i], i, chunks, 0};
        chunks += xfer_chunks_of((uint32_t)sz[i]);
    }
    XferRead *dup, *ddown;
    CK(cudaMalloc(&dup, n * sizeof(XferRead))); CK(cudaMalloc(&ddown, n * sizeof(XferRead)));
    CK(cudaMemcpy(dup, up.data(), n * sizeof(XferRead), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ddown, down.data(), n * sizeof(XferRead), cudaMemcpyHostToDevice));
    cudaStream_t s0, s1; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    XferParams pu = {dup, nullptr, n, chunks}, pd = {ddown, nullptr, n, chunks};
    for (uint32_t ctas : {8u, 16u, 32u, 64u, 148u, 296u}) {
        double t0 = now(); xfer_kernel<<<ctas, kXferThreads, 0, s0>>>(pu); CK(cudaStreamSynchronize(s0)); double t1 = now();
        double tu = t1 - t0;
        t0 = now(); xfer_kernel<<<ctas, kXferThreads, 0, s1>>>(pd); CK(cudaStreamSynchronize(s1)); t1 = now();
        double td = t1 - t0;
        t0 = now(); xfer_kernel<<<ctas, kXferThreads, 0, s0>>>(pu); xfer_kernel<<<ctas, kXferThreads, 0, s1>>>(pd);
        CK(cudaStreamSynchronize(s0)); CK(cudaStreamSynchronize(s1)); t1 = now();
        printf("ctas %3u: up %.1f GB/s  down %.1f GB/s  both at once %.1f GB/s each\n", ctas, htot / tu / 1e9, htot / td / 1e9, htot / (t1 - t0) / 1e9);
    }
    for (uint32_t cu : {32u, 64u, 128u, 256u})
        for (uint32_t cd : {2u, 4u, 8u, 16u}) {
            double t0 = now(); xfer_kernel<<<cu, kXferThreads, 0, s0>>>(pu); xfer_kernel<<<cd, kXferThreads, 0, s1>>>(pd);
            CK(cudaStreamSynchronize(s0)); double tu = now(); CK(cudaStreamSynchronize(s1)); double t1 = now();
            printf("up %3u ctas, down %2u ctas at once: up done %.1f ms, all done %.1f ms (%.1f GB/s each)\n", cu, cd, (tu - t0) * 1e3, (t1 - t0) * 1e3, htot / (t1 - t0) / 1e9);
        }
    {
        uint8_t* hc; CK(cudaMallocHost(&hc, dtot));
        double t0 = now();
        for (size_t o = 0; o < dtot; o += 32u << 20) CK(cudaMemcpyAsync(d2 + o, hc + o, std::min<size_t>(32u << 20, dtot - o), cudaMemcpyHostToDevice, s0));
        xfer_kernel<<<16, kXferThreads, 0, s1>>>(pd);
        CK(cudaStreamSynchronize(s0)); double tu = now(); CK(cudaStreamSynchronize(s1)); double t1 = now();
        printf("DMA up (32 MiB pieces) + kernel down (16 ctas): up done %.1f ms, all done %.1f ms\n", (tu - t0) * 1e3, (t1 - t0) * 1e3);
        t0 = now();
        xfer_kernel<<<64, kXferThreads, 0, s0>>>(pu);
        for (size_t o = 0; o < dtot; o += 32u << 20) CK(cudaMemcpyAsync(hc + o, d2 + o, std::min<size_t>(32u << 20, dtot - o), cudaMemcpyDeviceToHost, s1));
        CK(cudaStreamSynchronize(s0)); tu = now(); CK(cudaStreamSynchronize(s1)); t1 = now();
        printf("kernel up (64 ctas) + DMA down (32 MiB pieces): up done %.1f ms, all done %.1f ms\n", (tu - t0) * 1e3, (t1 - t0) * 1e3);
        t0 = now();
        for (size_t o = 0; o < dtot; o += 32u << 20) {
            CK(cudaMemcpyAsync(d2 + o, hc + o, std::min<size_t>(32u << 20, dtot - o), cudaMemcpyHostToDevice, s0));
            CK(cudaMemcpyAsync(h + o, d + o, std::min<size_t>(32u << 20, dtot - o), cudaMemcpyDeviceToHost, s1));
        }
        CK(cudaStreamSynchronize(s0)); tu = now(); CK(cudaStreamSynchronize(s1)); t1 = now();
        printf("DMA up + DMA down (32 MiB pieces): up done %.1f ms, all done %.1f ms\n", (tu - t0) * 1e3, (t1 - t0) * 1e3);
    }
    size_t bad = 0;
    for (uint32_t i = 0; i < n; ++i) if (memcmp(h + hoff[i], h2 + hoff[i], sz[i])) ++bad;
    for (uint32_t i = 0; i + 1 < n; ++i) if (h2[hoff[i] + sz[i]] != 0 || h2[hoff[i] + sz[i] + 1] != 0) ++bad;
    printf("round trip host -> device -> host: %zu reads differ\n", bad);
    return bad != 0;
}
