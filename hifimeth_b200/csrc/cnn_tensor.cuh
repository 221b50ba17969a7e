// cnn_tensor.cuh -- interface of the tensor-core (tcgen05, bf16 split precision) CNN path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/hm_engine.h"
#include "onnx_weights.h"

namespace hm {

struct TensorModel;      // packed bf16 hi/lo weights of one context, device resident
struct TensorWorkspaceImpl;

struct TensorWorkspace {
    TensorWorkspaceImpl* impl = nullptr;
};

struct TensorInputs {
    const uint8_t* bcode;
    const ushort4* kinf;
    const uint32_t* base_off;
    const uint32_t* site_read;
    const uint32_t* site_pos;
    const uint32_t* site_out;
    float* logits;
    uint8_t* ml;
};

struct TensorModelHandle {
    TensorModel* p = nullptr;
};

const char* tensor_last_error();
int tensor_model_build(TensorModelHandle& m, const CnnModel& host);
void tensor_model_free(TensorModelHandle& m);
int tensor_workspace_alloc(TensorWorkspace& w, uint32_t max_bases, uint32_t max_reads);
void tensor_workspace_free(TensorWorkspace& w);
// Runs the CNN of one context over site-list entries [first, first+count); writes logits / ml at site_out.
int tensor_cnn_run(const TensorModelHandle& m, TensorWorkspace& w, const TensorInputs& in, uint32_t first, uint32_t count,
                   cudaStream_t stream, uint32_t* launches, hm_timing* timing);

}  // namespace hm
