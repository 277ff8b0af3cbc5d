// Probe: how fast can one warp per 128 KiB histogram stage-1-like bytes (skewed: zigzag deltas, mean ~18) in shared memory?
// Variants: (A) one histogram per warp, atomicAdd per non-zero byte (what zstd_enc_code_kernel does); (B) four copies per warp;
// (C) __match_any_sync: one add per distinct value per byte position; (D) per-lane private 8-bit counters (no atomics), flushed
// every 15 chunks.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr uint32_t kGroup = 128u << 10;

template <int MODE>
__global__ void __launch_bounds__(32) hist_kernel(const uint8_t* src, uint32_t n_groups, uint32_t* out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* h = reinterpret_cast<uint32_t*>(smem);
    const uint32_t lane = threadIdx.x, g = blockIdx.x;
    if (g >= n_groups) return;
    const int copies = MODE == 1 ? 4 : 1;
    for (uint32_t i = lane; i < 256u * copies; i += 32) h[i] = 0;
    uint8_t* priv = smem + 4096;  // MODE 3: [256][32] bytes
    if (MODE == 3) for (uint32_t i = lane; i < 256 * 32 / 4; i += 32) reinterpret_cast<uint32_t*>(priv)[i] = 0;
    __syncwarp();
    const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)g * kGroup);
    uint32_t* hc = h + (MODE == 1 ? (lane & 3u) * 256u : 0u);
    uint32_t zeros = 0, since_flush = 0;
    for (uint32_t c0 = 0; c0 < kGroup / 16; c0 += 128) {
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = s4[c0 + 32 * u + lane];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t v = (w[k] >> (8 * j)) & 0xFFu;
                    if (MODE == 0 || MODE == 1) { if (v) atomicAdd(hc + v, 1u); else ++zeros; }
                    else if (MODE == 2) {
                        const uint32_t m = __match_any_sync(0xffffffffu, v);
                        if ((m & ((1u << lane) - 1u)) == 0) h[v] += __popc(m);  // the lowest lane of each value adds
                        __syncwarp();
                    } else {
                        priv[v * 32 + lane]++;
                    }
                }
            if (MODE == 3 && ++since_flush == 15) {
                __syncwarp();
                for (uint32_t i = lane; i < 256; i += 32) {
                    uint32_t t = 0;
                    const uint32_t* row = reinterpret_cast<const uint32_t*>(priv + i * 32);
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const uint32_t x = row[k]; t += (x & 0xFF) + ((x >> 8) & 0xFF) + ((x >> 16) & 0xFF) + (x >> 24); }
                    h[i] += t;
                }
                __syncwarp();
                for (uint32_t i = lane; i < 256 * 32 / 4; i += 32) reinterpret_cast<uint32_t*>(priv)[i] = 0;
                __syncwarp();
                since_flush = 0;
            }
        }
    }
    if (zeros) atomicAdd(hc, zeros);
    __syncwarp();
    if (MODE == 3) {
        for (uint32_t i = lane; i < 256; i += 32) {
            uint32_t t = 0;
            for (int k = 0; k < 32; ++k) t += priv[i * 32 + k];
            h[i] += t;
        }
        __syncwarp();
    }
    for (uint32_t i = lane; i < 256; i += 32) {
        uint32_t t = h[i];
        if (MODE == 1) t += h[256 + i] + h[512 + i] + h[768 + i];
        out[(size_t)g * 256 + i] = t;
    }
}

int main()
{
    const uint32_t n_groups = 16384;  // 2 GiB
    const size_t bytes = (size_t)n_groups * kGroup;
    std::vector<uint8_t> host(bytes);
    std::mt19937 rng(3);
    std::geometric_distribution<int> geo(1.0 / 19.0);
    for (size_t i = 0; i < bytes; ++i) { int v = geo(rng); host[i] = (uint8_t)(v > 255 ? 255 : v); }
    uint8_t* d; uint32_t *o, *oref;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&o, (size_t)n_groups * 1024)); CK(cudaMalloc(&oref, (size_t)n_groups * 1024));
    CK(cudaMemcpy(d, host.data(), bytes, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<uint32_t> ref((size_t)n_groups * 256), got((size_t)n_groups * 256);
    auto run = [&](int mode, const char* name) {
        uint32_t* dst = mode == 0 ? oref : o;
        const size_t smem = mode == 3 ? 4096 + 8192 : (mode == 1 ? 4096 : 1024);
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            if (mode == 0) hist_kernel<0><<<n_groups, 32, smem>>>(d, n_groups, dst);
            if (mode == 1) hist_kernel<1><<<n_groups, 32, smem>>>(d, n_groups, dst);
            if (mode == 2) hist_kernel<2><<<n_groups, 32, smem>>>(d, n_groups, dst);
            if (mode == 3) hist_kernel<3><<<n_groups, 32, smem>>>(d, n_groups, dst);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(mode == 0 ? ref.data() : got.data(), dst, ref.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        if (mode != 0) for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
        printf("%-44s %.2f ms  %.0f GB/s  (%zu wrong bins)\n", name, ms, bytes / ms / 1e6, bad);
    };
    run(0, "A one histogram per warp, atomics");
    run(1, "B four copies per warp, atomics");
    run(2, "C match_any, one add per distinct value");
    run(3, "D per-lane 8-bit counters, no atomics");
    return 0;
}
