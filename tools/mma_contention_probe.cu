// mma_contention_probe.cu -- inside a step the tensor kernels run their MMAs at 70 - 92 cycles each (clock64 stamps of
// dense_gemm2_kernel, dense_fused12_kernel and site_chain_kernel), against 64 - 65 in mma_rate_probe where nothing else runs.
// Which neighbour slows them?  One warp per SM issues the engine's split-precision triple (M = 128, N = 128, K = 16, both operands
// in shared memory, SWIZZLE_NONE) back to back while other warps of the same CTA generate one kind of traffic each:
//   bit 1  bulk copies global -> shared memory (the producers' ring fill), as many bytes per cycle as one warp can keep in flight
//   bit 2  accumulator reads + global stores (the map epilogue: tcgen05.ld 32 columns, 8 x STG.128 per lane), 4 warps
//   bit 4  accumulator reads + shared-memory stores (the fused kernel's epilogue-1: 8 x STS.128 per lane), 4 warps
//   bit 8  accumulator reads only (tcgen05.ld + wait), 4 warps
//   bit 16 no neighbour, but every MMA's operands recomputed from loop-carried values in vector registers (R2UR in front of each)
// Prints cycles per MMA for the chosen mix and the bytes per cycle each neighbour moved.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../hifimeth_b200/csrc -o mma_contention_probe mma_contention_probe.cu
// Run:   ./mma_contention_probe <mask> [reps]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "umma.cuh"

using namespace hm;

constexpr uint32_t kOperandBytes = 8 * 4096 + 8 * 4096;  // 8 A tiles + 8 B tiles (N = 128: 4 KB each)
constexpr uint32_t kFillBytes = 8 * 8192;                // ring the bulk copies land in
constexpr uint32_t kStsBytes = 32768;                    // area the shared-memory stores go to

__global__ void __launch_bounds__(320, 1) probe(int mask, int reps, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                unsigned long long* __restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, fill_bar[8];
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_stop;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (kOperandBytes + kFillBytes + kStsBytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        umma::mbar_init(&bar, 1);
        for (int i = 0; i < 8; ++i) umma::mbar_init(&fill_bar[i], 1);
        umma::fence_barrier_init();
        s_stop = 0;
    }
    if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = s_tmem;
    unsigned long long moved = 0;
    if (warp == 1) {
        // ---- MMA issuer ------------------------------------------------------------------------------------------------------
        const uint32_t desc_hi = (uint32_t)(umma::make_desc(0, 0, 128) >> 32);
        const uint32_t a_base = (uint32_t)umma::make_desc(umma::smem_u32(smem), 128 * 16, 128);
        const uint32_t b_base = (uint32_t)umma::make_desc(umma::smem_u32(smem + 8 * 4096), 128 * 16, 128);
        const uint32_t idesc = umma::make_idesc_bf16_m128(128);
        unsigned long long t0 = 0, t1 = 0;
        if (lane == 0) {
            t0 = clock64();
            if (mask & 16) {
                // the operands of every MMA are recomputed from loop-carried values (as the engine's kernels do: slot, stage and term
                // offsets): one integer op + one register-to-uniform move per operand and MMA in front of the UTCHMMA
                volatile uint32_t* jit = reinterpret_cast<volatile uint32_t*>(smem + kOperandBytes);  // reads 0: defeats hoisting
                for (int r = 0; r < reps; ++r) {
                    const uint32_t z = jit[r & 7];
                    const uint32_t a0 = a_base + (uint32_t)(r & 3) * 2 * (4096 >> 4) + z, a1 = a0 + (4096 >> 4) + z;
                    const uint32_t b0 = b_base + (uint32_t)(r & 3) * 2 * (4096 >> 4) + z, b1 = b0 + (4096 >> 4) + z;
                    const uint32_t d = tmem + (((uint32_t)r >> 5) & 1u) * 128u + z;
                    umma::mma_bf16_w(d, a0, b0, desc_hi, idesc + z, 1);
                    umma::mma_bf16_w(d + z, a1, b0 + z, desc_hi, idesc + z, 1);
                    umma::mma_bf16_w(d + 2 * z, a0 + z, b1, desc_hi, idesc + z, 1);
                }
            } else
            for (int r = 0; r < reps; ++r) {
                const uint32_t a0 = a_base + (uint32_t)(r & 3) * 2 * (4096 >> 4), a1 = a0 + (4096 >> 4);
                const uint32_t b0 = b_base + (uint32_t)(r & 3) * 2 * (4096 >> 4), b1 = b0 + (4096 >> 4);
                const uint32_t d = tmem + (((uint32_t)r >> 5) & 1u) * 128u;  // accumulator switches every 32 triples, like a tile
                umma::mma_bf16_w(d, a0, b0, desc_hi, idesc, 1);
                umma::mma_bf16_w(d, a1, b0, desc_hi, idesc, 1);
                umma::mma_bf16_w(d, a0, b1, desc_hi, idesc, 1);
            }
            umma::mma_commit(&bar);
        }
        __syncwarp();
        umma::mbar_wait(&bar, 0);
        t1 = clock64();
        if (lane == 0) {
            out[4 * blockIdx.x] = t1 - t0;
            s_stop = 1;
        }
    } else if (warp == 2 && (mask & 1)) {
        // ---- bulk copies into a ring of 8 x 8 KB, four 2 KB copies per slot issued by four lanes ------------------------------------
        uint32_t n = 0;
        while (!s_stop) {
            const uint32_t slot = n & 7u;
            if (n >= 8) umma::mbar_wait(&fill_bar[slot], ((n >> 3) - 1u) & 1u);
            if (lane == 0) umma::mbar_arrive_expect_tx(&fill_bar[slot], 8192);
            __syncwarp();
            if (lane < 4) umma::bulk_g2s(smem + kOperandBytes + slot * 8192 + lane * 2048, src + ((size_t)blockIdx.x * 64 + (n & 63u)) * 8192 + lane * 2048, 2048, &fill_bar[slot]);
            __syncwarp();
            ++n;
            moved += 8192;
        }
        for (uint32_t k = n > 8 ? n - 8 : 0; k < n; ++k) umma::mbar_wait(&fill_bar[k & 7u], (k >> 3) & 1u);  // nothing in flight at exit
        if (lane == 0) out[4 * blockIdx.x + 1] = moved;
    } else if (warp >= 4 && warp < 8 && (mask & (2 | 4 | 8))) {
        // ---- accumulator reads (+ global or shared stores) ---------------------------------------------------------------------------
        const uint32_t t_addr = tmem + (((warp & 3u) * 32u) << 16);
        uint8_t* g = dst + ((size_t)blockIdx.x * 128 + (warp & 3u) * 32 + lane) * 16;
        uint8_t* s = smem + kOperandBytes + kFillBytes + ((warp & 3u) * 32 + lane) * 16;
        uint32_t n = 0;
        while (!s_stop) {
            uint32_t v[32];
            umma::tmem_ld32(t_addr + (n & 3u) * 32u, v);
            umma::tmem_ld_wait();
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 q = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (mask & 2) *reinterpret_cast<uint4*>(g + (size_t)(j + 8 * (n & 7u)) * (148ull * 128 * 16)) = q;
                if (mask & 4) *reinterpret_cast<uint4*>(s + (size_t)j * 2048) = q;
            }
            if (mask & 4) umma::fence_proxy_async();
            ++n;
            moved += 512;
        }
        if (lane == 0 && warp == 4) out[4 * blockIdx.x + 2] = moved * 4;
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv)
{
    const int mask = argc > 1 ? atoi(argv[1]) : 0, reps = argc > 2 ? atoi(argv[2]) : 20000;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint8_t *src, *dst;
    unsigned long long* out;
    cudaMalloc(&src, (size_t)sms * 64 * 8192);
    cudaMalloc(&dst, (size_t)64 * 148 * 128 * 16);
    cudaMalloc(&out, (size_t)sms * 4 * sizeof(unsigned long long));
    cudaMemset(src, 0, (size_t)sms * 64 * 8192);
    cudaMemset(out, 0, (size_t)sms * 4 * sizeof(unsigned long long));
    const size_t smem = kOperandBytes + kFillBytes + kStsBytes + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < 2; ++it) {
        probe<<<sms, 320, smem>>>(mask, reps, src, dst, out);
        cudaError_t st = cudaDeviceSynchronize();
        if (st != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(st)); return 1; }
    }
    unsigned long long h[4 * 148] = {};
    cudaMemcpy(h, out, (size_t)sms * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0];
    printf("mask %2d: %6.1f cycles per MMA | bulk copies %5.1f B/cycle, accumulator reads %5.1f B/cycle per SM%s%s%s%s\n", mask, cyc / reps / 3, h[1] / cyc, h[2] / cyc,
           (mask & 1) ? " [+ring fill]" : "", (mask & 2) ? " [+ld, global stores]" : "", (mask & 4) ? " [+ld, shared stores]" : "", (mask & 8) ? " [+ld only]" : "");
    return 0;
}
