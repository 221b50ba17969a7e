// kernels_cnn_fp32.cuh -- fp32 CUDA-core evaluation of the DNAModNet graph (hm_cnn_mode HM_CNN_FP32_SIMT).
//
// This is the on-device cross-check for the tensor-core path: same graph, plain fp32 FMAs, one kernel per
// layer, channels-last activations [site][pos][C] in global memory.  Graph: training/model_cnn.py:76-85 as
// exported in models/*.onnx: bn0 -> 8 x (conv1d stride 2 pad 1 + bias + ReLU) -> flatten(channel-major)
// -> fc1 + ReLU -> fc2; post-process s_logits_to_methy_probs (src/app/hifimeth/mod_batch.cpp:46-64).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "postprocess.cuh"

namespace hm {

// Block = COUT threads, TP output positions of one site.  wt = [KW][CIN][COUT], bias [COUT].
// BN0: apply y = x*scale[c] + shift[c] to in-window columns while staging (bn0 is applied before conv1's
// zero padding, so the pad columns stay exactly 0).
template <int CIN, int COUT, int KW, int TP, bool BN0>
__global__ void __launch_bounds__(COUT)
conv_s2_fp32_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wt,
                    const float* __restrict__ bias, int lin, int lout, const float* __restrict__ bn_scale,
                    const float* __restrict__ bn_shift)
{
    constexpr int ROWS = 2 * TP + KW - 2;
    __shared__ __align__(16) float s_in[ROWS * CIN];
    const int site = blockIdx.y;
    const int t0 = blockIdx.x * TP;
    const float* src = in + (size_t)site * lin * CIN;
    for (int idx = threadIdx.x; idx < ROWS * CIN; idx += COUT) {
        int row = idx / CIN, c = idx % CIN;
        int pos = 2 * t0 - 1 + row;
        float v = 0.f;
        if (pos >= 0 && pos < lin) {
            v = src[(size_t)pos * CIN + c];
            if (BN0) v = fmaf(v, bn_scale[c], bn_shift[c]);
        }
        s_in[idx] = v;
    }
    __syncthreads();
    const int co = threadIdx.x;
    float acc[TP];
    const float b = bias[co];
    #pragma unroll
    for (int t = 0; t < TP; ++t) acc[t] = b;
    for (int j = 0; j < KW; ++j) {
        for (int c = 0; c < CIN; c += 4) {
            const float* wp = wt + ((size_t)j * CIN + c) * COUT + co;
            float w0 = wp[0], w1 = wp[COUT], w2 = wp[2 * COUT], w3 = wp[3 * COUT];
            #pragma unroll
            for (int t = 0; t < TP; ++t) {
                float4 x = *reinterpret_cast<const float4*>(&s_in[(2 * t + j) * CIN + c]);
                acc[t] = fmaf(x.x, w0, acc[t]);
                acc[t] = fmaf(x.y, w1, acc[t]);
                acc[t] = fmaf(x.z, w2, acc[t]);
                acc[t] = fmaf(x.w, w3, acc[t]);
            }
        }
    }
    float* dst = out + (size_t)site * lout * COUT;
    #pragma unroll
    for (int t = 0; t < TP; ++t)
        if (t0 + t < lout) dst[(size_t)(t0 + t) * COUT + co] = fmaxf(acc[t], 0.f);
}

// fc1 + ReLU + fc2 + softmax + quantise.  a8 = [site][2][64] channels-last; flatten index = c*2 + t.
// w1t = [128][256] (in, out), w2 = [2][256].  Block = 256 threads, 8 sites.
__global__ void __launch_bounds__(256)
fc_head_fp32_kernel(const float* __restrict__ a8, const float* __restrict__ w1t, const float* __restrict__ b1,
                    const float* __restrict__ w2, const float* __restrict__ b2, const uint32_t* __restrict__ site_out,
                    uint32_t first, uint32_t count, float* __restrict__ logits, uint8_t* __restrict__ ml)
{
    constexpr int SPB = 8;
    __shared__ float s_a[SPB][128];
    __shared__ float s_h[SPB][256];
    const uint32_t s0 = blockIdx.x * SPB;
    for (int idx = threadIdx.x; idx < SPB * 128; idx += 256) {
        int s = idx >> 7, k = idx & 127;   // k = c*2 + t
        int c = k >> 1, t = k & 1;
        s_a[s][k] = (s0 + s < count) ? a8[((size_t)(s0 + s) * 2 + t) * 64 + c] : 0.f;
    }
    __syncthreads();
    const int n = threadIdx.x;
    float acc[SPB];
    #pragma unroll
    for (int s = 0; s < SPB; ++s) acc[s] = b1[n];
    for (int k = 0; k < 128; ++k) {
        float w = w1t[k * 256 + n];
        #pragma unroll
        for (int s = 0; s < SPB; ++s) acc[s] = fmaf(s_a[s][k], w, acc[s]);
    }
    #pragma unroll
    for (int s = 0; s < SPB; ++s) s_h[s][n] = fmaxf(acc[s], 0.f);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (s0 + warp >= count) return;
    float v0 = 0.f, v1 = 0.f;
    for (int k = lane; k < 256; k += 32) {
        float h = s_h[warp][k];
        v0 = fmaf(h, w2[k], v0);
        v1 = fmaf(h, w2[256 + k], v1);
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, off);
        v1 += __shfl_xor_sync(0xffffffffu, v1, off);
    }
    if (lane == 0) {
        v0 += b2[0];
        v1 += b2[1];
        uint32_t o = site_out[first + s0 + warp];
        logits[2 * (size_t)o] = v0;
        logits[2 * (size_t)o + 1] = v1;
        ml[o] = prob_to_ml(softmax_p1(v0, v1));
    }
}

}  // namespace hm
