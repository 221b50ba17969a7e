// mma_rate_probe.cu -- what paces back-to-back tcgen05.mma on sm_100a when both operands come from shared memory
// (SWIZZLE_NONE, K-major core matrices, the layout of dense_gemm*.cuh)?  Every SM (or CTA pair) issues `reps` rounds of a fixed
// MMA pattern on resident operands and reports cycles per MMA and the chip-wide MAC rate.
//   pattern 0: one MMA per round, M x N x 16                                   (rate vs N)
//   pattern 1: the split-precision triple of the engine: (a_hi, w_hi) (a_lo, w_hi) (a_hi, w_lo), each M x N x 16
//   pattern 2: the same products with w_hi | w_lo side by side: (a_hi, [w_hi | w_lo]) as ONE M x 2N x 16 MMA + (a_lo, w_hi) M x N x 16
//              (cta_group::1 only: in a pair the halves of N come from different CTAs)
//   pattern 3: operand form 1 of round 2: one kind::f16 MMA (fp16, K = 16) + one kind::f8f6f4 MMA (e4m3, K = 32) per product
//   pattern 4: the e4m3 K = 32 MMA alone          pattern 5: the fp16 K = 16 MMA alone
//   pattern 6 / 7: pattern 1 with the accumulator alternating between two column ranges every round / every 8 rounds (does a switch
//              of accumulator drain the pipe?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../hifimeth_b200/csrc -o mma_rate_probe mma_rate_probe.cu ; run on a B200:
//   ./mma_rate_probe <pair 0|1> <N> <pattern> [reps]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "umma.cuh"

using namespace hm;

template <bool kPair>
__global__ void __launch_bounds__(128, 1) probe(int n, int pattern, int reps, unsigned long long* cycles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // operands: 8 A tiles of 128 rows x 16 K (4 KB each) then 8 B tiles of up to 512 rows x 16 K (16 KB each); contents irrelevant
    for (uint32_t i = threadIdx.x; i < (8 * 4096 + 8 * 16384) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        umma::mbar_init(&bar, 1);
        umma::fence_barrier_init();
    }
    if (warp == 0) {
        if (kPair) umma::tmem_alloc2(&s_tmem, 512);
        else umma::tmem_alloc(&s_tmem, 512);
    }
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    if (kPair) umma::cluster_sync();
    umma::tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t rank = kPair ? umma::cluster_ctarank() : 0u;
    if (warp == 1 && rank == 0) {
        const uint32_t nb = kPair ? (uint32_t)n / 2 : (uint32_t)n;  // B rows held by this CTA
        const uint32_t desc_hi = (uint32_t)(umma::make_desc(0, 0, 128) >> 32);
        const uint32_t a_base = (uint32_t)umma::make_desc(umma::smem_u32(smem), 128 * 16, 128);
        const uint32_t b_base = (uint32_t)umma::make_desc(umma::smem_u32(smem + 8 * 4096), nb * 16, 128);
        const uint32_t b2_base = (uint32_t)umma::make_desc(umma::smem_u32(smem + 8 * 4096), 2 * nb * 16, 128);  // N' = 2N tile
        const uint32_t idesc = kPair ? umma::make_idesc_bf16_m256((uint32_t)n) : umma::make_idesc_bf16_m128((uint32_t)n);
        const uint32_t idesc2 = kPair ? umma::make_idesc_bf16_m256((uint32_t)(2 * n)) : umma::make_idesc_bf16_m128((uint32_t)(2 * n));
        const uint32_t a_step = 4096 >> 4, b_step = 16384 >> 4;
        unsigned long long t0 = 0, t1 = 0;
        if (lane == 0) {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const uint32_t a0 = a_base + (uint32_t)(r & 3) * 2 * a_step, a1 = a0 + a_step;
                const uint32_t b0 = b_base + (uint32_t)(r & 3) * 2 * b_step, b1 = b0 + b_step;
                auto mma = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t id) {
                    if (kPair) umma::mma2_bf16_w(d, a, b, desc_hi, id, 1);
                    else umma::mma_bf16_w(d, a, b, desc_hi, id, 1);
                };
                if (pattern == 0) {
                    mma(tmem, a0, b0, idesc);
                } else if (pattern == 1) {
                    mma(tmem, a0, b0, idesc);
                    mma(tmem, a1, b0, idesc);
                    mma(tmem, a0, b1, idesc);
                } else if (pattern == 6) {  // the triple, accumulator alternating between two column ranges every round
                    const uint32_t d = tmem + ((uint32_t)r & 1u) * 256u;
                    mma(d, a0, b0, idesc);
                    mma(d, a1, b0, idesc);
                    mma(d, a0, b1, idesc);
                } else if (pattern == 7) {  // the triple, accumulator switched every 8 rounds, first MMA after a switch overwrites
                    const uint32_t d = tmem + (((uint32_t)r >> 3) & 1u) * 256u;
                    if ((r & 7) == 0) { if (kPair) umma::mma2_bf16_w(d, a0, b0, desc_hi, idesc, 0); else umma::mma_bf16_w(d, a0, b0, desc_hi, idesc, 0); }
                    else mma(d, a0, b0, idesc);
                    mma(d, a1, b0, idesc);
                    mma(d, a0, b1, idesc);
                } else if (pattern == 2) {
                    mma(tmem, a0, b2_base + (uint32_t)(r & 3) * 2 * b_step, idesc2);
                    mma(tmem, a1, b0, idesc);
                } else {
                    const uint32_t idf = kPair ? umma::make_idesc_f16_m256((uint32_t)n) : umma::make_idesc_f16_m128((uint32_t)n);
                    if (pattern == 3 || pattern == 5) mma(tmem, a0, b0, idf);
                    if (pattern == 3 || pattern == 4) {
                        if (kPair) umma::mma2_f8_w(tmem, a1, b1, desc_hi, idf, 1);
                        else umma::mma_f8_w(tmem, a1, b1, desc_hi, idf, 1);
                    }
                }
            }
            if (kPair) umma::mma2_commit_mc(&bar);
            else umma::mma_commit(&bar);
        }
        __syncwarp();
        umma::mbar_wait(&bar, 0);
        t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    }
    umma::tc_fence_before();
    __syncthreads();
    if (kPair) umma::cluster_sync();
    umma::tc_fence_after();
    if (warp == 0) {
        if (kPair) umma::tmem_dealloc2(tmem, 512);
        else umma::tmem_dealloc(tmem, 512);
    }
}

int main(int argc, char** argv)
{
    const int pair = argc > 1 ? atoi(argv[1]) : 1, n = argc > 2 ? atoi(argv[2]) : 128, pattern = argc > 3 ? atoi(argv[3]) : 1;
    const int reps = argc > 4 ? atoi(argv[4]) : 20000;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = pair ? (sms & ~1) : sms;
    unsigned long long* d_cyc;
    cudaMalloc(&d_cyc, grid * sizeof(unsigned long long));
    cudaMemset(d_cyc, 0, grid * sizeof(unsigned long long));
    const size_t smem = 8 * 4096 + 8 * 16384 + 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int it = 0; it < 3; ++it) {  // warm-up twice, time the third
        cudaEventRecord(e0);
        if (pair) {
            cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, probe<true>, n, pattern, reps, d_cyc);
        } else {
            cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            probe<false><<<grid, 128, smem>>>(n, pattern, reps, d_cyc);
        }
        cudaEventRecord(e1);
        cudaError_t st = cudaDeviceSynchronize();
        if (st != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(st)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    unsigned long long h[256] = {};
    cudaMemcpy(h, d_cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    const int per_round = pattern == 0 ? 1 : (pattern == 1 || pattern == 6 || pattern == 7) ? 3 : (pattern == 2 || pattern == 3) ? 2 : 1;
    const double macs_round = (double)(pair ? 256 : 128) * n * 16 * (pattern == 0 ? 1 : pattern == 4 ? 2 : pattern == 5 ? 1 : 3);
    const int issuers = pair ? grid / 2 : grid;
    printf("pair %d N %3d pattern %d: %8.1f cycles/round (%6.1f per MMA issued), %7.1f TFLOP/s chip-wide (%.3f ms)\n", pair, n, pattern,
           (double)h[0] / reps, (double)h[0] / reps / per_round, 2.0 * macs_round * reps * issuers / (ms * 1e-3) / 1e12, ms);
    return 0;
}
