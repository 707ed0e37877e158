// tools/membench.cu — microbenchmark behind DESIGN.md §"memory access granularity":
// bandwidth of warp-wide gathers of `chunk`-byte contiguous granules at pseudo-random offsets inside
// a region of `region` bytes (read, and read-modify-write), on one B200.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int VEC>   // VEC = uint32 words per lane per granule (1 -> 128 B, 4 -> 512 B per warp access)
__global__ void gather(uint32_t* base, size_t granules, int per_warp, int rmw, uint32_t* sink) {
    const int lane = threadIdx.x & 31;
    size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint64_t s = warp * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (int i = 0; i < per_warp; i += 4) {
        uint4 v[4]; size_t g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            g[u] = (size_t)((s >> 20) & (granules - 1));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t* p = base + g[u] * (32 * VEC) + lane * VEC;
            if (VEC == 4) v[u] = __ldcg(reinterpret_cast<uint4*>(p));
            else v[u].x = __ldcg(p);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            acc += v[u].x;
            if (rmw) {
                uint32_t* p = base + g[u] * (32 * VEC) + lane * VEC;
                if (VEC == 4) { v[u].x += 1; __stcg(reinterpret_cast<uint4*>(p), v[u]); }
                else __stcg(p, v[u].x + 1);
            }
        }
    }
    if (acc == 0xdeadbeef) *sink = acc;
}

int main(int argc, char** argv) {
    size_t max_region = (size_t)32 << 30;
    uint32_t* buf; uint32_t* sink;
    cudaMalloc(&buf, max_region); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, max_region);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rmw = 0; rmw < 2; ++rmw)
    for (int vec : {1, 4})
    for (size_t region : {(size_t)128 << 20, (size_t)1 << 30, (size_t)8 << 30, (size_t)32 << 30}) {
        size_t chunk = (size_t)128 * vec, granules = region / chunk;
        int blocks = 148 * 8, threads = 256, per_warp = 2048;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (vec == 1) gather<1><<<blocks, threads>>>(buf, granules, per_warp, rmw, sink);
            else gather<4><<<blocks, threads>>>(buf, granules, per_warp, rmw, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double bytes = (double)blocks * (threads / 32) * per_warp * chunk * (rmw ? 2 : 1);
        printf("%s chunk=%zuB region=%6.1fGB  %.1f GB/s\n", rmw ? "rmw " : "read", chunk, region / 1073741824.0, bytes / ms / 1e6);
    }
    return 0;
}
