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

}  // namespace hm
