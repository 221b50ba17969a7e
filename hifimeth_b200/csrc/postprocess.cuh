// postprocess.cuh -- logits -> probability -> ML byte, shared by both CNN paths.
// Reference: s_logits_to_methy_probs, src/app/hifimeth/mod_batch.cpp:46-64 (max-subtracted 2-way softmax,
// `int v = 255 * p1` truncation, clamp to 255).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hm {

__device__ __forceinline__ float softmax_p1(float v0, float v1)
{
    float m = fmaxf(v0, v1);
    float e0 = expf(v0 - m), e1 = expf(v1 - m);
    return __fdiv_rn(e1, e0 + e1);
}

__device__ __forceinline__ uint8_t prob_to_ml(float p1)
{
    int v = (int)(255.0f * p1);  // truncation, mod_batch.cpp:59
    return (uint8_t)(v > 255 ? 255 : v);
}

// the per-context 256-bin histograms of the ML bytes that `hifimeth pileup` builds while it re-reads
// mod.bam (src/app/hifimeth/pileup.cpp:237-272; records with flag 0x900 -- secondary / supplementary -- are left out), taken
// here from the bytes while they are still on the device.  site lists are laid out [CpG | CHG | CHH fwd | CHH rev]; totals[k]
// holds the four region sizes.  hist = [3][256] u32, zeroed by the caller.  Integer work, one read of 9 B/site.
static __global__ void __launch_bounds__(256)
ml_hist_kernel(const uint32_t* __restrict__ site_read, const uint32_t* __restrict__ site_out, const uint8_t* __restrict__ ml,
               const uint16_t* __restrict__ flag, const uint32_t* __restrict__ totals, uint32_t n_calls, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_h[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += 256) s_h[i] = 0;
    __syncthreads();
    const uint32_t t0 = totals[0], t01 = t0 + totals[1];
    for (uint32_t k = blockIdx.x * 256u + threadIdx.x; k < n_calls; k += gridDim.x * 256u) {
        if (flag[site_read[k]] & 0x900u) continue;
        const uint32_t c = k < t0 ? 0u : k < t01 ? 1u : 2u;
        atomicAdd(&s_h[c * 256u + ml[site_out[k]]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 256; i += 256)
        if (s_h[i]) atomicAdd(&hist[i], s_h[i]);
}

}  // namespace hm
