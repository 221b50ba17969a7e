// bulk_copy_probe.cu -- how fast can one CTA per SM stream global -> shared memory on sm_100a?
//   mode 0: cp.async.bulk (1-D, TMA engine) of `chunk` bytes, `depth` copies in flight per CTA, one issuing lane
//   mode 1: cp.async 16 B per thread (LDGSTS), `nthreads` threads, groups of `chunk` bytes, `depth` groups in flight
//   mode 2: mode 0 but `split` lanes issue the copies (each its own mbarrier slot set)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_probe bulk_copy_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("{.reg .b64 s; mbarrier.arrive.expect_tx.shared::cta.b64 s, [%0], %1;}" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}

// Each CTA streams `per_cta` bytes starting at src + blockIdx.x * per_cta (pattern 0), or interleaved chunk-by-chunk
// across CTAs (pattern 1: chunk i of CTA b at (i * gridDim + b) * chunk), or `nstreams` separate streams per CTA (pattern 2).
__global__ void __launch_bounds__(256) probe(const uint8_t* src, size_t per_cta, uint32_t chunk, int depth, int mode, int pattern, int nstreams, size_t stream_stride, unsigned long long* cycles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* buf = smem + 1024;
    const size_t nchunks = per_cta / chunk;
    if (threadIdx.x == 0) { for (int i = 0; i < (depth > 64 ? 64 : depth); ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    auto addr = [&](size_t i) -> const uint8_t* {
        if (pattern == 0) return src + (size_t)blockIdx.x * per_cta + i * chunk;
        if (pattern == 1) return src + (i * gridDim.x + blockIdx.x) * (size_t)chunk;
        // pattern 2: nstreams planes; consecutive chunks cycle over planes; within a plane CTAs are adjacent
        size_t pl = i % nstreams, k = i / nstreams;
        return src + pl * stream_stride + (k * gridDim.x + blockIdx.x) * (size_t)chunk;
    };
    long long t0 = clock64();
    if (mode == 0) {
        if (threadIdx.x == 0) {
            for (size_t i = 0; i < nchunks + depth; ++i) {
                int slot = i % depth;
                if (i >= (size_t)depth) mbar_wait(&bars[slot], ((i / depth) - 1) & 1);
                if (i < nchunks) { mbar_expect(&bars[slot], chunk); bulk(buf + (size_t)slot * chunk, addr(i), chunk, &bars[slot]); }
            }
        }
    } else if (mode == 2 || mode == 3) {
        // several issuers: mode 2 = `nstreams` lanes of warp 0 in lockstep, mode 3 = `nstreams` warps (lane 0 of each)
        const int W = nstreams;
        const int me = mode == 2 ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
        const bool on = mode == 2 ? (threadIdx.x < (unsigned)W) : ((threadIdx.x & 31) == 0 && me < W);
        if (on) {
            const int dper = depth / W;  // slots per issuer
            size_t n_me = nchunks / W;
            for (size_t k = 0; k < n_me + dper; ++k) {
                int slot = me * dper + (int)(k % dper);
                if (k >= (size_t)dper) mbar_wait(&bars[slot], ((k / dper) - 1) & 1);
                if (k < n_me) { mbar_expect(&bars[slot], chunk); bulk(buf + (size_t)slot * chunk, src + (size_t)blockIdx.x * per_cta + (k * W + me) * (size_t)chunk, chunk, &bars[slot]); }
            }
        }
    } else if (mode == 4) {
        // the dense kernel's producer: W = nstreams_w warps (blockDim/32), each stage = L lanes x chunk, stage t by warp t % W
        const int W = blockDim.x >> 5, L = 4, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int dper = depth / W;
        const size_t nst = nchunks / L, n_me = nst / W;
        for (size_t k = 0; k < n_me + dper; ++k) {
            int slot = w * dper + (int)(k % dper);
            if (k >= (size_t)dper) mbar_wait(&bars[slot], ((k / dper) - 1) & 1);
            if (k < n_me) {
                size_t stage_id = k * W + w;
                if (lane == 0) mbar_expect(&bars[slot], chunk * L);
                __syncwarp();
                if (lane < L) bulk(buf + ((size_t)slot * L + lane) * chunk, addr(stage_id * L + lane), chunk, &bars[slot]);
            }
            __syncwarp();
        }
    } else {
        // LDGSTS: all threads copy 16 B pieces of chunk i; commit groups; wait depth-1 behind
        const int nt = blockDim.x;
        for (size_t i = 0; i < nchunks; ++i) {
            int slot = i % depth;
            const uint8_t* s = addr(i);
            for (uint32_t o = threadIdx.x * 16; o < chunk; o += nt * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(buf + (size_t)slot * chunk + o)), "l"(s + o) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else if (depth == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
            else if (depth == 8) asm volatile("cp.async.wait_group 7;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

__global__ void issue_latency(const uint8_t* src, long long* out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* buf = smem + 1024;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t[10];
        for (int i = 0; i < 8; ++i) mbar_expect(&bars[i], 2048);
        t[0] = clock64();
        for (int i = 0; i < 8; ++i) { bulk(buf + i * 2048, src + (size_t)i * (1 << 20) + blockIdx.x * 4096, 2048, &bars[i]); t[i + 1] = clock64(); }
        for (int i = 0; i < 8; ++i) { mbar_wait(&bars[i], 0); out[16 + i] = clock64() - t[0]; }
        for (int i = 0; i < 9; ++i) out[i] = t[i] - t[0];
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("{.reg .b64 s; mbarrier.arrive.shared::cta.b64 s, [%0];}" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait_test(uint64_t* b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
// two warps ping-pong over two mbarriers; nwaiters lanes of each warp wait (all arrive count = 1 by lane 0)
__global__ void pingpong(int iters, int use_test, int extra_pollers, long long* out)
{
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    long long t0 = clock64();
    if (warp == 0) {
        for (int i = 0; i < iters; ++i) {
            if (lane == 0) mbar_arrive(&bars[0]);
            if (use_test) mbar_wait_test(&bars[1], i & 1); else mbar_wait(&bars[1], i & 1);
            __syncwarp();
        }
    } else if (warp == 1) {
        for (int i = 0; i < iters; ++i) {
            if (use_test) mbar_wait_test(&bars[0], i & 1); else mbar_wait(&bars[0], i & 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[1]);
        }
    } else if (warp - 2 < extra_pollers) {
        mbar_wait(&bars[2], 0);  // idle pollers: suspended on a barrier that completes at the end
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; mbar_arrive(&bars[2]); }
}

int main()
{
    {
        long long* o; cudaMallocManaged(&o, 64);
        for (int test = 0; test < 2; ++test)
            for (int pol = 0; pol <= 8; pol += 8) {
                pingpong<<<1, 32 * (2 + pol)>>>(1000, test, pol, o); cudaDeviceSynchronize();
                printf("pingpong %s pollers %d: %.1f cycles per round trip\n", test ? "test_wait" : "try_wait", pol, o[0] / 1000.0);
            }
    }
    {
        uint8_t* s0; cudaMalloc(&s0, 64 << 20); cudaMemset(s0, 0, 64 << 20);
        long long* o; cudaMallocManaged(&o, 32 * 8);
        issue_latency<<<1, 32, 32768>>>(s0, o); cudaDeviceSynchronize();
        printf("issue times: "); for (int i = 0; i < 9; ++i) printf("%lld ", o[i]); printf("\ncompletion times: "); for (int i = 0; i < 8; ++i) printf("%lld ", o[16 + i]); printf("\n");
    }
    const size_t total = (size_t)1 << 31;  // 2 GiB source
    uint8_t* src; cudaMalloc(&src, total + (1 << 20)); cudaMemset(src, 1, total);
    unsigned long long* cyc; cudaMallocManaged(&cyc, 148 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int mode, pattern, nstreams; uint32_t chunk; int depth; int threads; } cfgs[] = {
        {14, 2, 32, 2048, 8, 128}, {14, 2, 32, 2048, 8, 128}, {4, 0, 1, 2048, 8, 128}, {4, 1, 1, 2048, 8, 128}, {4, 2, 32, 2048, 8, 128}, {4, 2, 32, 2048, 16, 128}, {4, 2, 32, 2048, 4, 32}, {4, 2, 32, 2048, 8, 32}, {4, 2, 4, 2048, 8, 128}, {4, 2, 32, 4096, 8, 128},
        {2, 0, 4, 2048, 32, 32}, {2, 0, 16, 2048, 64, 32}, {2, 0, 32, 2048, 64, 32}, {3, 0, 4, 2048, 32, 128}, {3, 0, 8, 2048, 64, 256}, {3, 0, 4, 8192, 16, 128}, {2, 0, 4, 8192, 16, 32},
        {0, 0, 1, 2048, 3, 32}, {0, 0, 1, 2048, 8, 32}, {0, 0, 1, 2048, 32, 32}, {0, 0, 1, 8192, 3, 32}, {0, 0, 1, 8192, 8, 32}, {0, 0, 1, 8192, 16, 32},
        {0, 0, 1, 32768, 4, 32}, {0, 1, 1, 2048, 8, 32}, {0, 1, 1, 2048, 32, 32}, {0, 1, 1, 8192, 8, 32}, {0, 2, 32, 2048, 8, 32}, {0, 2, 32, 2048, 32, 32},
        {0, 2, 4, 2048, 32, 32}, {0, 0, 1, 512, 32, 32}, {0, 0, 1, 1024, 64, 32},
        {1, 0, 1, 8192, 4, 128}, {1, 0, 1, 8192, 8, 128}, {1, 0, 1, 8192, 8, 256}, {1, 1, 1, 8192, 8, 256}, {1, 2, 32, 2048, 8, 128}, {1, 0, 1, 32768, 4, 256},
    };
    for (auto& c : cfgs) {
        size_t per_cta = ((size_t)8 << 20), stride = (size_t)64 << 20; int reps = 2;
        if (c.mode == 14) { c.mode = 4; per_cta = (size_t)1 << 20; stride = 4850176; reps = 200; }
        size_t smem = 1024 + (size_t)c.chunk * c.depth * (c.mode == 4 ? 4 : 1);
        if (smem > 200 * 1024) continue;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < reps; ++rep) {
            cudaEventRecord(e0);
            probe<<<148, c.threads, smem>>>(src, per_cta, c.chunk, c.depth, c.mode, c.pattern, c.nstreams, stride, cyc);
            cudaEventRecord(e1);
            cudaError_t st = cudaDeviceSynchronize();
            if (st != cudaSuccess) { printf("error %s\n", cudaGetErrorString(st)); return 1; }
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gbs = 148.0 * per_cta / ms / 1e6;
        printf("mode %d pattern %d streams %2d chunk %6u depth %2d threads %3d : %7.3f ms  %7.1f GB/s  %5.1f B/cyc/SM (cta0 %llu cyc)\n", c.mode, c.pattern, c.nstreams,
               c.chunk, c.depth, c.threads, ms, gbs, (double)per_cta / cyc[0], cyc[0]);
    }
    return 0;
}
