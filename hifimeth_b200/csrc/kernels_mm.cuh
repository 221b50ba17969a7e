// kernels_mm.cuh -- SURVEY.md s8f row N1: the MM:Z skip counts of build_one_mod_bam on the device.
//
// Reference (src/corelib/build_mod_bam.cpp:134-168): MM = "C+m" then, for every forward-strand call, ",<number of C's in the
// forward-strand sequence between the previous call + 1 and this call>"; ";G-m" and the same with G's for the reverse-strand
// calls; ";".  The host loop walks every base of every read (get_bam_fwd_strand_base per position); here one thread per call
// counts the few bases since the previous call in the already resident forward-strand codes and the decimal text is laid
// out with a scan, so the host only copies bytes.  HBM-bound integer/byte work: ~1 B/base read, ~7 B/call written.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hm {

constexpr int kMmBlock = 1024;

__device__ __forceinline__ uint32_t dec_digits(uint32_t v)
{
    return v < 10u ? 1u : v < 100u ? 2u : v < 1000u ? 3u : v < 10000u ? 4u : v < 100000u ? 5u : v < 1000000u ? 6u : v < 10000000u ? 7u
         : v < 100000000u ? 8u : v < 1000000000u ? 9u : 10u;
}

// Pass 1: one thread per call (hm_call_batch order).  delta[k], and per block the number of text bytes (1 + digits each).
__global__ void __launch_bounds__(kMmBlock)
mm_delta_kernel(const uint8_t* __restrict__ bcode, const uint32_t* __restrict__ base_off, const uint32_t* __restrict__ call_off,
                const uint32_t* __restrict__ n_fwd, const int32_t* __restrict__ qoff, uint32_t n_reads, uint32_t n_calls,
                uint32_t* __restrict__ delta, uint32_t* __restrict__ block_sum)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_r0;
    const uint32_t k = blockIdx.x * kMmBlock + threadIdx.x;
    uint32_t len = 0;
    if (threadIdx.x == 0) {
        // read of the block's first call: last r with call_off[r] <= k (one binary search per block; calls are ordered by
        // read, so every other thread walks forward from there -- a read has thousands of calls, a block 1 024)
        uint32_t lo = 0, hi = n_reads;
        const uint32_t k0 = min(k, n_calls ? n_calls - 1 : 0u);
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (call_off[mid] <= k0) lo = mid; else hi = mid;
        }
        s_r0 = lo;
    }
    __syncthreads();
    if (k < n_calls) {
        uint32_t r = s_r0;
        while (r + 1 < n_reads && call_off[r + 1] <= k) ++r;
        const uint32_t a = call_off[r], nf = n_fwd[r], idx = k - a;
        const bool rev = idx >= nf;
        const uint8_t target = rev ? 2 : 1;  // G : C in forward-strand codes
        const int32_t q = qoff[k];
        const int32_t from = (idx == 0 || idx == nf) ? 0 : qoff[k - 1] + 1;
        const uint8_t* s = bcode + base_off[r];
        uint32_t d = 0;
        for (int32_t p = from; p < q; ++p) d += (s[p] == target);
        delta[k] = d;
        len = 1u + dec_digits(d);
    }
    uint32_t v = len;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = s_warp[threadIdx.x];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) block_sum[blockIdx.x] = v;
    }
}

// Pass 2: exclusive prefix over the block sums (one block); block_off[n_blocks] = total text bytes.
__global__ void __launch_bounds__(1024)
mm_scan_blocks_kernel(const uint32_t* __restrict__ block_sum, uint32_t n_blocks, uint32_t* __restrict__ block_off)
{
    __shared__ uint32_t s_warp[32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t per = (n_blocks + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, n_blocks), hi = min(lo + per, n_blocks);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += block_sum[i];
    // inclusive scan of the 1 024 partials: shuffles inside a warp, one shared round over the 32 warp totals
    uint32_t inc = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if ((int)lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + inc - sum;
    for (uint32_t i = lo; i < hi; ++i) {
        block_off[i] = run;
        run += block_sum[i];
    }
    if (threadIdx.x == 1023) block_off[n_blocks] = run;
}

// Pass 3: text offset of every call (text_off[n_calls] = total) and the characters ",<delta>".
__global__ void __launch_bounds__(kMmBlock)
mm_write_kernel(const uint32_t* __restrict__ delta, const uint32_t* __restrict__ block_off, uint32_t n_calls, uint32_t text_cap,
                uint32_t* __restrict__ text_off, uint8_t* __restrict__ text)
{
    __shared__ uint32_t s_warp[32];
    const uint32_t k = blockIdx.x * kMmBlock + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t d = 0, len = 0;
    if (k < n_calls) {
        d = delta[k];
        len = 1u + dec_digits(d);
    }
    uint32_t incl = len;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_warp[lane], w = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if ((int)lane >= o) w += t;
        }
        s_warp[lane] = w - v;
    }
    __syncthreads();
    const uint32_t off = block_off[blockIdx.x] + s_warp[warp] + incl - len;
    if (k < n_calls) {
        text_off[k] = off;
        if (off + len <= text_cap) {
            text[off] = ',';
            uint32_t v = d;
            for (uint32_t i = len - 1; i >= 1; --i) {
                text[off + i] = (uint8_t)('0' + v % 10u);
                v /= 10u;
            }
        }
    }
    if (k == 0) text_off[n_calls] = block_off[gridDim.x];
}

// Pass 4: per read, where its text starts and how much of it belongs to the forward-strand calls.
__global__ void mm_read_offsets_kernel(const uint32_t* __restrict__ text_off, const uint32_t* __restrict__ call_off,
                                       const uint32_t* __restrict__ n_fwd, uint32_t n_reads, uint32_t* __restrict__ mm_off,
                                       uint32_t* __restrict__ mm_fwd_len)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    const uint32_t t0 = text_off[call_off[r]];
    mm_off[r] = t0;
    if (r < n_reads) mm_fwd_len[r] = text_off[call_off[r] + n_fwd[r]] - t0;
}

}  // namespace hm
