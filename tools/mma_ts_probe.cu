// mma_ts_probe.cu -- tcgen05.mma with the A operand in TENSOR MEMORY (the ".ts" form: cute's SM100_MMA_F16BF16_TS / _2x1SM_TS) on
// sm_100a: is the layout what we think it is, and what paces it?  A kernel that keeps its running activation maps in TMEM instead
// of shared memory (site_chain.cuh, round 2) depends on both answers.
//   layout   row m of the 128-row A tile = TMEM lane m; K runs along the columns, two bf16 per 32-bit column (element 2c in the low
//            half of column c): a K = 16 tile is 8 columns, written by tcgen05.st.32x32b.x8 (thread = lane = row).
//   check    D_ts = A(tmem) * B(smem)^T against D_ss = A(smem) * B(smem)^T and against the exact integer result, every element
//   rate     back-to-back TS MMAs (the split-precision triple: two A tiles x two B tiles), cycles per MMA vs N
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../hifimeth_b200/csrc -o mma_ts_probe mma_ts_probe.cu
// Run:   ./mma_ts_probe <pair 0|1> <N> [reps]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "umma.cuh"

using namespace hm;

__device__ __forceinline__ int a_val(int m, int k) { return ((m * 7 + k * 3) % 13) - 6; }
__device__ __forceinline__ int b_val(int n, int k) { return ((n * 5 + k) % 11) - 5; }
__device__ __forceinline__ uint16_t bf16_bits(int v) { return __bfloat16_as_ushort(__float2bfloat16_rn((float)v)); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <bool kPair>
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc)
{
    if (kPair) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc), "r"(0u)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
            "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc), "r"(0u)
            : "memory");
    }
}

template <bool kPair>
__global__ void __launch_bounds__(128, 1) probe(int n, int reps, unsigned long long* cycles, unsigned int* bad)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = kPair ? umma::cluster_ctarank() : 0u;
    const int nb = kPair ? n / 2 : n;  // B rows held by this CTA
    uint8_t* sA = smem;                 // [2 groups][128 rows][16 B]
    uint8_t* sB = smem + 4096;          // [2 groups][nb rows][16 B]
    const int m = (int)threadIdx.x, gm = (int)rank * 128 + m;
    for (int k = 0; k < 16; ++k) reinterpret_cast<uint16_t*>(sA + (k >> 3) * 2048 + m * 16)[k & 7] = bf16_bits(a_val(gm, k));
    for (int r = m; r < nb; r += 128)
        for (int k = 0; k < 16; ++k) reinterpret_cast<uint16_t*>(sB + (k >> 3) * nb * 16 + r * 16)[k & 7] = bf16_bits(b_val((int)rank * nb + r, k));
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
    umma::tc_fence_after();
    const uint32_t tmem = s_tmem;
    // A into TMEM: columns 480..487 (and a second copy at 488..495 for the rate loop)
    {
        uint32_t v[8];
        for (int c = 0; c < 8; ++c) v[c] = (uint32_t)bf16_bits(a_val(gm, 2 * c)) | ((uint32_t)bf16_bits(a_val(gm, 2 * c + 1)) << 16);
        const uint32_t t = tmem + ((warp * 32u) << 16);
        tmem_st8(t + 480, v);
        tmem_st8(t + 488, v);
        tmem_st_wait();
    }
    umma::tc_fence_before();
    __syncthreads();
    if (kPair) umma::cluster_sync();
    umma::tc_fence_after();
    const uint32_t idesc = kPair ? umma::make_idesc_bf16_m256((uint32_t)n) : umma::make_idesc_bf16_m128((uint32_t)n);
    const uint64_t desc_a = umma::make_desc(umma::smem_u32(sA), 2048, 128);
    const uint64_t desc_b = umma::make_desc(umma::smem_u32(sB), (uint32_t)nb * 16, 128);
    if (warp == 1 && rank == 0) {
        if (lane == 0) {
            if (kPair) umma::mma2_bf16_w(tmem, (uint32_t)desc_a, (uint32_t)desc_b, (uint32_t)(desc_a >> 32), idesc, 0);
            else umma::mma_bf16(tmem, desc_a, desc_b, idesc, 0);
            mma_ts<kPair>(tmem + 256, tmem + 480, desc_b, idesc, 0);
            if (kPair) umma::mma2_commit_mc(&bar);
            else umma::mma_commit(&bar);
        }
        __syncwarp();
    }
    umma::mbar_wait(&bar, 0);
    umma::tc_fence_after();
    // every thread checks its row of both results
    unsigned int wrong_ss = 0, wrong_ts = 0;
    for (int c0 = 0; c0 < n; c0 += 16) {
        uint32_t ss[16], ts[16];
        const uint32_t t = tmem + ((warp * 32u) << 16);
        umma::tmem_ld16(t + (uint32_t)c0, ss);
        umma::tmem_ld16(t + 256 + (uint32_t)c0, ts);
        umma::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) {
            int want = 0;
            for (int k = 0; k < 16; ++k) want += a_val(gm, k) * b_val(c0 + j, k);
            wrong_ss += __uint_as_float(ss[j]) != (float)want;
            wrong_ts += __uint_as_float(ts[j]) != (float)want;
        }
    }
    if (wrong_ss) atomicAdd(&bad[0], wrong_ss);
    if (wrong_ts) atomicAdd(&bad[1], wrong_ts);
    umma::tc_fence_before();
    __syncthreads();
    if (kPair) umma::cluster_sync();
    umma::tc_fence_after();
    // ---- rate: the split-precision triple with A in TMEM -------------------------------------------------------------------
    if (warp == 1 && rank == 0) {
        unsigned long long t0 = 0, t1 = 0;
        if (lane == 0) {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                mma_ts<kPair>(tmem, tmem + 480, desc_b, idesc, 1);
                mma_ts<kPair>(tmem, tmem + 488, desc_b, idesc, 1);
                mma_ts<kPair>(tmem, tmem + 480, desc_b, idesc, 1);
            }
            if (kPair) umma::mma2_commit_mc(&bar);
            else umma::mma_commit(&bar);
        }
        __syncwarp();
        umma::mbar_wait(&bar, 1);
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
    const int pair = argc > 1 ? atoi(argv[1]) : 1, n = argc > 2 ? atoi(argv[2]) : 128;
    const int reps = argc > 3 ? atoi(argv[3]) : 20000;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = pair ? (sms & ~1) : sms;
    unsigned long long* d_cyc;
    unsigned int* d_bad;
    cudaMalloc(&d_cyc, grid * sizeof(unsigned long long));
    cudaMalloc(&d_bad, 2 * sizeof(unsigned int));
    cudaMemset(d_cyc, 0, grid * sizeof(unsigned long long));
    cudaMemset(d_bad, 0, 2 * sizeof(unsigned int));
    const size_t smem = 4096 + 256 * 32 + 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int it = 0; it < 2; ++it) {
        cudaMemset(d_bad, 0, 2 * sizeof(unsigned int));
        cudaEventRecord(e0);
        if (pair) {
            cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, probe<true>, n, reps, d_cyc, d_bad);
        } else {
            cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            probe<false><<<grid, 128, smem>>>(n, reps, d_cyc, d_bad);
        }
        cudaEventRecord(e1);
        cudaError_t st = cudaDeviceSynchronize();
        if (st != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(st)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    unsigned long long h[256] = {};
    unsigned int bad[2] = {};
    cudaMemcpy(h, d_cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
    const double macs_round = (double)(pair ? 256 : 128) * n * 16 * 3;
    const int issuers = pair ? grid / 2 : grid;
    printf("pair %d N %3d: wrong elements SS %u, TS %u (of %d per CTA) | TS triple %8.1f cycles (%6.1f per MMA), %7.1f TFLOP/s chip-wide\n", pair, n,
           bad[0], bad[1], 128 * n, (double)h[0] / reps, (double)h[0] / reps / 3, 2.0 * macs_round * reps * issuers / (ms * 1e-3) / 1e12);
    return 0;
}
