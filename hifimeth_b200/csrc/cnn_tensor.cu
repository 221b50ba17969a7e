// cnn_tensor.cu -- tensor-core CNN path (placeholder until the tcgen05 kernels land in this file).
#include "cnn_tensor.cuh"

#include <string>

namespace hm {
namespace { std::string g_err = "tensor-core CNN path not built yet"; }
const char* tensor_last_error() { return g_err.c_str(); }
int tensor_model_build(TensorModelHandle&, const CnnModel&) { return -1; }
void tensor_model_free(TensorModelHandle&) {}
int tensor_workspace_alloc(TensorWorkspace&, uint32_t, uint32_t) { return -1; }
void tensor_workspace_free(TensorWorkspace&) {}
int tensor_cnn_run(const TensorModelHandle&, TensorWorkspace&, const TensorInputs&, uint32_t, uint32_t, cudaStream_t, uint32_t*, hm_timing*) { return -1; }
}  // namespace hm
