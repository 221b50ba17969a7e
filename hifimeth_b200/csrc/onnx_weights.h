// onnx_weights.h -- loads models/{CpG,CHG,CHH}.onnx into plain host arrays.
//
// Replaces ov::Core::read_model + the input-shape checks of ModModels::s_load_one_model
// (reference: src/app/hifimeth/mod_main.cpp:32-67).  Both ONNX dialects that ship are handled
// (SURVEY.md appendix A): opset 17 with named initializers and Gemm(transB=1) (CpG, CHG), and opset 11
// with every weight in a Constant node and MatMul+Add with pre-transposed FC weights (CHH).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace hm {

struct ConvLayer {
    int cout = 0, cin = 0, k = 0;
    std::vector<float> w;  // [cout][cin][k] as stored in the file
    std::vector<float> b;  // [cout]
};

struct CnnModel {
    int kmer = 0, features = 0;            // from the graph input shape [batch, kmer, features]
    float bn_eps = 1e-5f;
    std::vector<float> bn_w, bn_b, bn_mean, bn_var;  // [features]
    std::vector<ConvLayer> convs;          // 8 layers, stride 2, pad 1
    std::vector<float> fc1_w, fc1_b;       // [256][128] (out, in), [256]
    std::vector<float> fc2_w, fc2_b;       // [2][256], [2]
    int fc1_out = 0, fc1_in = 0, fc2_out = 0;
};

// Returns true on success; on failure err holds a one-line reason.
bool load_onnx_model(const std::string& path, CnnModel& out, std::string& err);

// models/{CpG,CHG,CHH}.pt -- the TorchScript exports the reference's app-gpu binary loads (torch::jit::load,
// src/app-gpu/hifimeth-gpu/5mc_call_gpu.cpp:48; written by training/make-torch-script.py:28-30).  A .pt file is a ZIP archive
// whose 24 frozen constants are STORED (uncompressed) members `<name>/constants/0..23`, little-endian f32, in graph order
// (SURVEY.md appendix A): bn0 weight, bias, mean, var; 8 x (conv W [cout][cin][k], b); fc1 [in][out] (transposed), b; fc2
// [in][out], b.  Shapes follow from the member sizes and the fixed channel plan; conv1's kernel size comes from the file.
// No libtorch, no pickle: the central directory is enough.
bool load_pt_model(const std::string& path, CnnModel& out, std::string& err);

// By extension: ".pt" -> load_pt_model, anything else -> load_onnx_model.
bool load_model_file(const std::string& path, CnnModel& out, std::string& err);

}  // namespace hm
