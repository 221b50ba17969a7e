// store_stride_probe.cu -- do the epilogue's stores pay for address translation?  Every warp of every SM writes tiles of
// 32 rows x 16 B to `planes` planes (the [plane][row][16 B] activation layout of the engine), `stride` bytes apart, walking rows
// tile by tile like the persistent kernels do.  Reports cycles per 16-byte store instruction and GB/s for a small and a large
// plane stride (same bytes, same instruction stream; only the number of 2 MB pages touched per tile differs).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_stride_probe store_stride_probe.cu
//   ./store_stride_probe <planes> <stride_bytes> [tiles_per_warp]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 1) probe(uint8_t* base, int planes, size_t stride, int tiles, unsigned long long* cyc)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    const uint4 v = make_uint4(lane, warp, blockIdx.x, 7);
    const unsigned long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
        const size_t row = ((size_t)t * nw + gw) * 32 + lane;   // consecutive warps take consecutive 32-row tiles
        for (int p = 0; p < planes; ++p) *reinterpret_cast<uint4*>(base + (size_t)p * stride + row * 16) = v;
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv)
{
    const int planes = argc > 1 ? atoi(argv[1]) : 32;
    const size_t stride = argc > 2 ? strtoull(argv[2], nullptr, 10) : (32ull << 20);
    const int tiles = argc > 3 ? atoi(argv[3]) : 400;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t rows = (size_t)tiles * sms * 8 * 32;
    if (rows * 16 > stride) { printf("stride too small for %zu rows\n", rows); return 1; }
    uint8_t* d;
    unsigned long long* c;
    cudaMalloc(&d, (size_t)planes * stride);
    cudaMalloc(&c, sms * sizeof(unsigned long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        probe<<<sms, 256>>>(d, planes, stride, tiles, c);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    unsigned long long h[256];
    cudaMemcpy(h, c, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    const double bytes = (double)rows * 16 * planes;
    printf("planes %3d stride %10zu: %7.1f cycles per store instruction per warp, %7.1f GB/s (%.3f ms)\n", planes, stride,
           (double)h[0] / ((double)tiles * planes), bytes / ms / 1e6, ms);
    return 0;
}
