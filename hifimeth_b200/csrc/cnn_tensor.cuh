// cnn_tensor.cuh -- interface of the tensor-core (tcgen05, bf16 split precision) CNN path.
//
// The reference evaluates the network once per site on a 401-wide window (src/app/hifimeth/mod_batch.cpp:66-75).
// This path evaluates it as a *dilated dense plan* over strand positions (see cnn_tensor.cu), every layer being one
// launch of dense_gemm_kernel (dense_gemm.cuh); a site's logits are row (o - 201) of the final map.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/hm_engine.h"
#include "onnx_weights.h"

namespace hm {

struct TensorModel;          // plan + packed bf16 hi/lo weights of one context, device resident
struct TensorWorkspaceImpl;  // activation maps of one sub-batch, logits of one batch, track tables

struct TensorWorkspace {
    TensorWorkspaceImpl* impl = nullptr;
};
struct TensorModelHandle {
    TensorModel* p = nullptr;
};

// One submitted read batch as the CNN stage sees it.
struct TensorBatch {
    // device: outputs of decode_kernel / scan_write_kernel
    const uint8_t* d_bcode;
    const ushort4* d_kinf;
    const uint32_t* d_base_off;
    const uint32_t* d_site_read;
    const uint32_t* d_site_pos;
    const uint32_t* d_site_out;
    // host (pinned staging of the slot)
    const uint32_t* h_base_off;
    const uint8_t* h_valid;
    const uint32_t* h_read_pref;  // [n_reads + 1][4] sites of each class in reads before r (exclusive prefix)
    uint32_t n_reads;
    // class regions of the site list: CpG | CHG | CHH fwd | CHH rev
    uint32_t class_count[4];
    // results, indexed by site_out
    float* d_logits;
    uint8_t* d_ml;
};

const char* tensor_last_error();
int tensor_model_build(TensorModelHandle& m, const CnnModel& host);
void tensor_model_free(TensorModelHandle& m);
// max_rows: dense rows (128-row tiles) of one sub-batch; 0 = default
int tensor_workspace_alloc(TensorWorkspace& w, uint32_t max_bases, uint32_t max_reads, uint32_t max_rows);
void tensor_workspace_free(TensorWorkspace& w);
// Runs the CNN of every enabled context over one batch: features -> dense plan -> per-site logits + ML byte.
int tensor_batch_run(const TensorModelHandle* models /*[3]*/, uint32_t ctx_mask, TensorWorkspace& w, const TensorBatch& b,
                     cudaStream_t stream, int sm_count, uint32_t* launches, hm_timing* timing);

// Device time (ms) of the dense plan of the last batch run on this workspace; call after the stream is synchronised.
float tensor_last_dense_ms(TensorWorkspace& w);

// Debug readback of what the PRODUCT path computes on (hm_debug_dump_xmap / hm_debug_dump_acts; valid after tensor_batch_run,
// stream synchronised).  A site is (read, rev, o = strand offset); compact_row = its row in the compact maps of its context.
//   xwindow:   out [n][401][8] f32 = hi + lo of the X-map rows the site's window covers (raw features, before bn0).
//   site_acts: out [n][n_l][C_l] f32 = the per-site output of conv layer `layer` (1..8) as assembled from the maps the plan left
//              behind: Y_l (dense) for interior positions, F_l / G_l (compact) for the first / last, T7 / T8 for layers 7, 8;
//              NaN where the product path never holds the value in HBM (interior of layer 1 when conv1 + conv2 are fused,
//              except the rows scattered for F2 / G2).  The last run must have been this context alone, in one sub-batch.
void tensor_debug_spill(TensorWorkspace& w, bool on);  // debug reruns: the chain kernel also stores its shared-memory maps to HBM
int tensor_debug_xwindow(TensorWorkspace& w, uint32_t n, const uint32_t* read, const uint8_t* rev, const int32_t* o, float* out, cudaStream_t stream);
int tensor_debug_site_acts(const TensorModelHandle& model, int ctx, TensorWorkspace& w, uint32_t n, const uint32_t* read, const uint8_t* rev,
                           const int32_t* o, const uint32_t* compact_row, int layer, float* out, int* n_out, int* channels, cudaStream_t stream);

// Unit-test hook behind hm_debug_dense_op (include/hm_engine.h).
int tensor_debug_dense_op(int device, uint32_t rows, uint32_t rows_alloc, int cin, int cout, int n_src, const float* const* src,
                          int n_terms, const int32_t* term_src, const int32_t* term_shift, const float* weights, const float* bias,
                          int conv1_taps, const float* w2, const float* b2, const uint32_t* gather_rows, uint32_t gather_mask, float* out);

// Device time (ms) of the kernel launched by the last tensor_debug_dense_op.
float tensor_debug_last_op_ms();

}  // namespace hm
