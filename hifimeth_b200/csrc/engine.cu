// engine.cu -- the C ABI of include/hm_engine.h: engine/slot lifetime, the per-batch stage graph
// (H2D -> decode -> scan -> CNN -> D2H on the slot's stream), validation hooks and microbenchmarks.
//
// Reference control flow being replaced: s_worker_thread, src/app/hifimeth/mod_main.cpp:145-262.
#include <algorithm>
#include <cstdarg>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hm_engine.h"
#include "cnn_tensor.cuh"
#include "kernels_cnn_fp32.cuh"
#include "kernels_front.cuh"
#include "kernels_mm.cuh"
#include "onnx_weights.h"

namespace {

std::mutex g_err_mu;
std::string g_create_error;

struct Fp32Model {
    float *bn_scale = nullptr, *bn_shift = nullptr;
    float* conv_wt[8] = {};
    float* conv_b[8] = {};
    float *fc1_wt = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr;
};

struct Geometry {
    int kw[8], cin[8], cout[8], lin[8], lout[8];
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    // pinned staging (inputs)
    hm_read_batch host{};
    uint32_t *h_chunk_read = nullptr, *h_chunk_pos = nullptr, *h_read_first_chunk = nullptr;
    // device inputs
    uint8_t *d_seq4 = nullptr, *d_fi = nullptr, *d_fp = nullptr, *d_ri = nullptr, *d_rp = nullptr, *d_valid = nullptr;
    uint32_t *d_base_off = nullptr, *d_seq_off = nullptr;
    uint16_t* d_flag = nullptr;
    uint32_t *d_chunk_read = nullptr, *d_chunk_pos = nullptr, *d_read_first_chunk = nullptr;
    // device intermediates
    uint8_t* d_bcode = nullptr;
    ushort4* d_kinf = nullptr;
    hm::ClassCount *d_chunk_cnt = nullptr, *d_pref = nullptr;
    uint64_t* d_cls = nullptr;  // site classes of every position, 4 bits each (decode_kernel -> scan_write_kernel)
    uint32_t* d_totals = nullptr;
    uint32_t *d_site_read = nullptr, *d_site_pos = nullptr, *d_site_out = nullptr;
    // device outputs
    uint32_t *d_call_off = nullptr, *d_n_fwd = nullptr, *d_read_pref = nullptr, *h_read_pref = nullptr;
    int32_t* d_qoff = nullptr;
    uint8_t *d_ml = nullptr, *d_call_ctx = nullptr;
    float* d_logits = nullptr;
    // pinned outputs
    uint32_t *h_call_off = nullptr, *h_n_fwd = nullptr, *h_totals = nullptr;
    int32_t* h_qoff = nullptr;
    uint8_t* h_ml = nullptr;
    // MM skip-count text (row N1; HM_SUBMIT_MM_TEXT)
    uint32_t *d_mm_delta = nullptr, *d_mm_bsum = nullptr, *d_mm_boff = nullptr, *d_mm_toff = nullptr, *d_mm_off = nullptr, *d_mm_fwd_len = nullptr;
    uint8_t* d_mm_text = nullptr;
    uint32_t *h_mm_off = nullptr, *h_mm_fwd_len = nullptr, *h_mm_total = nullptr;
    uint8_t* h_mm_text = nullptr;
    size_t mm_text_cap = 0;
    bool mm_valid = false;
    uint32_t *d_ml_hist = nullptr, *h_ml_hist = nullptr;  // [3][256]
    bool hist_valid = false;
    hm::ReadKinStats *d_read_stats = nullptr, *h_read_stats = nullptr;  // [max_reads]
    bool stats_valid = false;
    // fp32 CNN workspace (site chunk)
    float* d_feat = nullptr;
    float* d_act[8] = {};
    // tensor CNN workspace
    hm::TensorWorkspace tws;
    // state
    uint32_t n_reads = 0, n_bases = 0, n_chunks = 0, n_calls = 0, max_chunks = 0;
    uint32_t totals[5] = {};
    bool resident = false, submitted = false, collected = false;
    hm_timing timing{};
};

}  // namespace

struct hm_engine {
    hm_config cfg{};
    std::string model_dir;
    std::string err;
    int n_slots = 0;
    uint32_t ctx_mask = 7;
    Geometry geo[3];
    hm::CnnModel host_model[3];
    Fp32Model fp32[3];
    hm::TensorModelHandle tensor[3];
    bool have_model[3] = {false, false, false};
    std::vector<Slot> slots;
    uint32_t site_chunk = 4096;  // fp32 path: sites per CNN pass
    int sm_count = 148;
};

namespace {

int fail(hm_engine* e, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (e) e->err = buf;
    else {
        std::lock_guard<std::mutex> g(g_err_mu);
        g_create_error = buf;
    }
    return code;
}

#define HM_CUDA(e, stage, call)                                                                      \
    do {                                                                                             \
        cudaError_t _st = (call);                                                                    \
        if (_st != cudaSuccess) return fail((e), HM_ERR_CUDA, "CUDA error in %s: %s", (stage), cudaGetErrorString(_st)); \
    } while (0)

template <typename T>
cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }
template <typename T>
cudaError_t hmalloc(T** p, size_t n) { return cudaMallocHost((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }

int upload(hm_engine* e, float** dst, const std::vector<float>& src)
{
    HM_CUDA(e, "weight upload", dmalloc(dst, src.size()));
    HM_CUDA(e, "weight upload", cudaMemcpy(*dst, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice));
    return HM_OK;
}

int build_fp32_model(hm_engine* e, int c)
{
    const hm::CnnModel& m = e->host_model[c];
    Fp32Model& d = e->fp32[c];
    std::vector<float> scale(m.features), shift(m.features);
    for (int i = 0; i < m.features; ++i) {
        float inv = 1.0f / sqrtf(m.bn_var[i] + m.bn_eps);
        scale[i] = m.bn_w[i] * inv;
        shift[i] = m.bn_b[i] - m.bn_mean[i] * scale[i];
    }
    int rc;
    if ((rc = upload(e, &d.bn_scale, scale))) return rc;
    if ((rc = upload(e, &d.bn_shift, shift))) return rc;
    for (int l = 0; l < 8; ++l) {
        const hm::ConvLayer& cv = m.convs[l];
        std::vector<float> wt((size_t)cv.k * cv.cin * cv.cout);
        for (int o = 0; o < cv.cout; ++o)
            for (int i = 0; i < cv.cin; ++i)
                for (int j = 0; j < cv.k; ++j) wt[((size_t)j * cv.cin + i) * cv.cout + o] = cv.w[((size_t)o * cv.cin + i) * cv.k + j];
        if ((rc = upload(e, &d.conv_wt[l], wt))) return rc;
        if ((rc = upload(e, &d.conv_b[l], cv.b))) return rc;
    }
    std::vector<float> w1t((size_t)m.fc1_in * m.fc1_out);
    for (int o = 0; o < m.fc1_out; ++o)
        for (int i = 0; i < m.fc1_in; ++i) w1t[(size_t)i * m.fc1_out + o] = m.fc1_w[(size_t)o * m.fc1_in + i];
    if ((rc = upload(e, &d.fc1_wt, w1t))) return rc;
    if ((rc = upload(e, &d.fc1_b, m.fc1_b))) return rc;
    if ((rc = upload(e, &d.fc2_w, m.fc2_w))) return rc;
    if ((rc = upload(e, &d.fc2_b, m.fc2_b))) return rc;
    return HM_OK;
}

int check_geometry(hm_engine* e, int c)
{
    const hm::CnnModel& m = e->host_model[c];
    static const int want_cin[8] = {8, 128, 128, 128, 96, 96, 96, 64};
    static const int want_cout[8] = {128, 128, 128, 96, 96, 96, 64, 64};
    if (m.kmer != HM_KMER || m.features != HM_FEATURES_PER_BASE)
        return fail(e, HM_ERR_MODEL, "model %d: input geometry [%d,%d], engine is built for [401,8]", c, m.kmer, m.features);
    Geometry& g = e->geo[c];
    int len = m.kmer;
    for (int l = 0; l < 8; ++l) {
        const hm::ConvLayer& cv = m.convs[l];
        if (cv.cin != want_cin[l] || cv.cout != want_cout[l] || (l > 0 && cv.k != 3) || (l == 0 && cv.k != 11 && cv.k != 13))
            return fail(e, HM_ERR_MODEL, "model %d: conv%d is [%d,%d,%d], unsupported", c, l + 1, cv.cout, cv.cin, cv.k);
        g.kw[l] = cv.k; g.cin[l] = cv.cin; g.cout[l] = cv.cout; g.lin[l] = len;
        len = (len + 2 - cv.k) / 2 + 1;
        g.lout[l] = len;
    }
    if (len != 2 || m.fc1_in != 128 || m.fc1_out != 256)
        return fail(e, HM_ERR_MODEL, "model %d: head geometry unsupported (L8=%d fc1 %dx%d)", c, len, m.fc1_out, m.fc1_in);
    return HM_OK;
}

void free_slot(Slot& s)
{
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFreeHost(s.host.base_off); cudaFreeHost(s.host.seq_off); cudaFreeHost(s.host.seq4); cudaFreeHost(s.host.flag);
    cudaFreeHost(s.host.valid); cudaFreeHost(s.host.fi); cudaFreeHost(s.host.fp); cudaFreeHost(s.host.ri); cudaFreeHost(s.host.rp);
    cudaFreeHost(s.h_chunk_read); cudaFreeHost(s.h_chunk_pos); cudaFreeHost(s.h_read_first_chunk);
    cudaFreeHost(s.h_read_pref); cudaFree(s.d_read_pref);
    cudaFree(s.d_mm_delta); cudaFree(s.d_mm_bsum); cudaFree(s.d_mm_boff); cudaFree(s.d_mm_toff); cudaFree(s.d_mm_off); cudaFree(s.d_mm_fwd_len);
    cudaFree(s.d_ml_hist); cudaFreeHost(s.h_ml_hist);
    cudaFree(s.d_read_stats); cudaFreeHost(s.h_read_stats);
    cudaFree(s.d_mm_text); cudaFreeHost(s.h_mm_off); cudaFreeHost(s.h_mm_fwd_len); cudaFreeHost(s.h_mm_total); cudaFreeHost(s.h_mm_text);
    cudaFreeHost(s.h_call_off); cudaFreeHost(s.h_n_fwd); cudaFreeHost(s.h_totals); cudaFreeHost(s.h_qoff); cudaFreeHost(s.h_ml);
    void* dev[] = {s.d_seq4, s.d_fi, s.d_fp, s.d_ri, s.d_rp, s.d_valid, s.d_base_off, s.d_seq_off, s.d_flag, s.d_chunk_read,
                   s.d_chunk_pos, s.d_read_first_chunk, s.d_bcode, s.d_kinf, s.d_chunk_cnt, s.d_cls, s.d_pref, s.d_totals, s.d_site_read,
                   s.d_site_pos, s.d_site_out, s.d_call_off, s.d_n_fwd, s.d_qoff, s.d_ml, s.d_call_ctx, s.d_logits, s.d_feat};
    for (void* p : dev) cudaFree(p);
    for (float* p : s.d_act) cudaFree(p);
    hm::tensor_workspace_free(s.tws);
    for (auto& ev : s.ev) if (ev) cudaEventDestroy(ev);
    if (s.stream) cudaStreamDestroy(s.stream);
}

int alloc_slot(hm_engine* e, Slot& s)
{
    const size_t R = e->cfg.max_reads, B = e->cfg.max_bases;
    const size_t seq_cap = B / 2 + R + 16;
    s.max_chunks = (uint32_t)(B / hm::kChunk + R + 1);
    const char* st = "slot allocation";
    const auto ta = std::chrono::steady_clock::now();
    HM_CUDA(e, st, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    for (auto& ev : s.ev) HM_CUDA(e, st, cudaEventCreate(&ev));
    s.host.max_reads = (uint32_t)R;
    s.host.max_bases = (uint32_t)B;
    HM_CUDA(e, st, hmalloc(&s.host.base_off, R + 1));
    HM_CUDA(e, st, hmalloc(&s.host.seq_off, R + 1));
    HM_CUDA(e, st, hmalloc(&s.host.seq4, seq_cap));
    HM_CUDA(e, st, hmalloc(&s.host.flag, R));
    HM_CUDA(e, st, hmalloc(&s.host.valid, R));
    HM_CUDA(e, st, hmalloc(&s.host.fi, B));
    HM_CUDA(e, st, hmalloc(&s.host.fp, B));
    HM_CUDA(e, st, hmalloc(&s.host.ri, B));
    HM_CUDA(e, st, hmalloc(&s.host.rp, B));
    HM_CUDA(e, st, hmalloc(&s.h_chunk_read, s.max_chunks));
    HM_CUDA(e, st, hmalloc(&s.h_chunk_pos, s.max_chunks));
    HM_CUDA(e, st, hmalloc(&s.h_read_first_chunk, R + 1));
    HM_CUDA(e, st, hmalloc(&s.h_call_off, R + 1));
    HM_CUDA(e, st, hmalloc(&s.h_n_fwd, R));
    HM_CUDA(e, st, hmalloc(&s.h_totals, 8));
    HM_CUDA(e, st, hmalloc(&s.h_read_pref, 4 * (R + 1)));
    HM_CUDA(e, st, dmalloc(&s.d_read_pref, 4 * (R + 1)));
    HM_CUDA(e, st, hmalloc(&s.h_qoff, B));
    HM_CUDA(e, st, hmalloc(&s.h_ml, B));
    // decode_kernel stages with 16-byte loads that may over-read by < 16 bytes: 64 bytes of slack on every input array
    HM_CUDA(e, st, dmalloc(&s.d_seq4, seq_cap + 64));
    HM_CUDA(e, st, dmalloc(&s.d_fi, B + 64));
    HM_CUDA(e, st, dmalloc(&s.d_fp, B + 64));
    HM_CUDA(e, st, dmalloc(&s.d_ri, B + 64));
    HM_CUDA(e, st, dmalloc(&s.d_rp, B + 64));
    HM_CUDA(e, st, dmalloc(&s.d_valid, R));
    HM_CUDA(e, st, dmalloc(&s.d_base_off, R + 1));
    HM_CUDA(e, st, dmalloc(&s.d_seq_off, R + 1));
    HM_CUDA(e, st, dmalloc(&s.d_flag, R));
    HM_CUDA(e, st, dmalloc(&s.d_chunk_read, s.max_chunks));
    HM_CUDA(e, st, dmalloc(&s.d_chunk_pos, s.max_chunks));
    HM_CUDA(e, st, dmalloc(&s.d_read_first_chunk, R + 1));
    HM_CUDA(e, st, dmalloc(&s.d_bcode, B));
    HM_CUDA(e, st, dmalloc(&s.d_kinf, B));
    HM_CUDA(e, st, dmalloc(&s.d_chunk_cnt, s.max_chunks));
    HM_CUDA(e, st, dmalloc(&s.d_cls, (size_t)s.max_chunks * hm::kFrontThreads));
    HM_CUDA(e, st, dmalloc(&s.d_pref, s.max_chunks + 1));
    HM_CUDA(e, st, dmalloc(&s.d_totals, 8));
    HM_CUDA(e, st, dmalloc(&s.d_site_read, B));
    HM_CUDA(e, st, dmalloc(&s.d_site_pos, B));
    HM_CUDA(e, st, dmalloc(&s.d_site_out, B));
    HM_CUDA(e, st, dmalloc(&s.d_call_off, R + 1));
    HM_CUDA(e, st, dmalloc(&s.d_n_fwd, R));
    HM_CUDA(e, st, dmalloc(&s.d_qoff, B));
    HM_CUDA(e, st, dmalloc(&s.d_ml, B));
    HM_CUDA(e, st, dmalloc(&s.d_call_ctx, B));
    HM_CUDA(e, st, dmalloc(&s.d_logits, 2 * B));
    s.mm_text_cap = 3 * B + 64;  // ",0" is 2 bytes per call; a skip count of d digits needs d skipped bases
    HM_CUDA(e, st, dmalloc(&s.d_mm_delta, B));
    HM_CUDA(e, st, dmalloc(&s.d_mm_bsum, B / hm::kMmBlock + 2));
    HM_CUDA(e, st, dmalloc(&s.d_mm_boff, B / hm::kMmBlock + 2));
    HM_CUDA(e, st, dmalloc(&s.d_mm_toff, B + 1));
    HM_CUDA(e, st, dmalloc(&s.d_mm_off, R + 1));
    HM_CUDA(e, st, dmalloc(&s.d_mm_fwd_len, R));
    HM_CUDA(e, st, dmalloc(&s.d_mm_text, s.mm_text_cap));
    HM_CUDA(e, st, hmalloc(&s.h_mm_off, R + 1));
    HM_CUDA(e, st, hmalloc(&s.h_mm_fwd_len, R));
    HM_CUDA(e, st, hmalloc(&s.h_mm_total, 1));
    HM_CUDA(e, st, dmalloc(&s.d_read_stats, R));
    HM_CUDA(e, st, hmalloc(&s.h_read_stats, R));
    HM_CUDA(e, st, dmalloc(&s.d_ml_hist, 3 * 256));
    HM_CUDA(e, st, hmalloc(&s.h_ml_hist, 3 * 256));
    HM_CUDA(e, st, hmalloc(&s.h_mm_text, s.mm_text_cap));
    if (e->cfg.cnn_mode == HM_CNN_FP32_SIMT) {
        const size_t S = e->site_chunk;
        HM_CUDA(e, st, dmalloc(&s.d_feat, S * HM_KMER * HM_FEATURES_PER_BASE));
        int lmax[8], cmax[8];
        for (int l = 0; l < 8; ++l) { lmax[l] = 0; cmax[l] = 0; }
        for (int c = 0; c < 3; ++c) {
            if (!e->have_model[c]) continue;
            for (int l = 0; l < 8; ++l) { lmax[l] = std::max(lmax[l], e->geo[c].lout[l]); cmax[l] = std::max(cmax[l], e->geo[c].cout[l]); }
        }
        for (int l = 0; l < 8; ++l) HM_CUDA(e, st, dmalloc(&s.d_act[l], S * (size_t)lmax[l] * cmax[l]));
    } else {
        const auto tb = std::chrono::steady_clock::now();
        int rc = hm::tensor_workspace_alloc(s.tws, e->cfg.max_bases, e->cfg.max_reads, 0);
        if (rc) return fail(e, HM_ERR_CUDA, "CUDA error in tensor workspace allocation: %s", hm::tensor_last_error());
        if (getenv("HM_VERBOSE"))
            fprintf(stderr, "[hm_engine_create]   slot: staging + lists %.3f s, tensor workspace %.3f s\n", std::chrono::duration<double>(tb - ta).count(),
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count());
    }
    return HM_OK;
}

// ---- fp32 CNN over one class region of the site list -----------------------------------------------------
template <int CIN, int COUT, int KW, bool BN0>
void launch_conv_fp32(const float* in, float* out, const float* wt, const float* b, int lin, int lout, uint32_t nsites,
                      const float* sc, const float* sh, cudaStream_t st)
{
    constexpr int TP = 16;
    dim3 grid((lout + TP - 1) / TP, nsites);
    hm::conv_s2_fp32_kernel<CIN, COUT, KW, TP, BN0><<<grid, COUT, 0, st>>>(in, out, wt, b, lin, lout, sc, sh);
}

int run_cnn_fp32(hm_engine* e, Slot& s, int ctx, uint32_t first, uint32_t count, uint32_t& launches)
{
    const Fp32Model& m = e->fp32[ctx];
    const Geometry& g = e->geo[ctx];
    for (uint32_t off = 0; off < count; off += e->site_chunk) {
        const uint32_t n = std::min(e->site_chunk, count - off);
        const uint32_t f = first + off;
        hm::gather_features_kernel<<<n, 128, 0, s.stream>>>(s.d_bcode, s.d_kinf, s.d_base_off, s.d_site_read, s.d_site_pos, f, n, s.d_feat);
        if (g.kw[0] == 11)
            launch_conv_fp32<8, 128, 11, true>(s.d_feat, s.d_act[0], m.conv_wt[0], m.conv_b[0], g.lin[0], g.lout[0], n, m.bn_scale, m.bn_shift, s.stream);
        else
            launch_conv_fp32<8, 128, 13, true>(s.d_feat, s.d_act[0], m.conv_wt[0], m.conv_b[0], g.lin[0], g.lout[0], n, m.bn_scale, m.bn_shift, s.stream);
        launch_conv_fp32<128, 128, 3, false>(s.d_act[0], s.d_act[1], m.conv_wt[1], m.conv_b[1], g.lin[1], g.lout[1], n, nullptr, nullptr, s.stream);
        launch_conv_fp32<128, 128, 3, false>(s.d_act[1], s.d_act[2], m.conv_wt[2], m.conv_b[2], g.lin[2], g.lout[2], n, nullptr, nullptr, s.stream);
        launch_conv_fp32<128, 96, 3, false>(s.d_act[2], s.d_act[3], m.conv_wt[3], m.conv_b[3], g.lin[3], g.lout[3], n, nullptr, nullptr, s.stream);
        launch_conv_fp32<96, 96, 3, false>(s.d_act[3], s.d_act[4], m.conv_wt[4], m.conv_b[4], g.lin[4], g.lout[4], n, nullptr, nullptr, s.stream);
        launch_conv_fp32<96, 96, 3, false>(s.d_act[4], s.d_act[5], m.conv_wt[5], m.conv_b[5], g.lin[5], g.lout[5], n, nullptr, nullptr, s.stream);
        launch_conv_fp32<96, 64, 3, false>(s.d_act[5], s.d_act[6], m.conv_wt[6], m.conv_b[6], g.lin[6], g.lout[6], n, nullptr, nullptr, s.stream);
        launch_conv_fp32<64, 64, 3, false>(s.d_act[6], s.d_act[7], m.conv_wt[7], m.conv_b[7], g.lin[7], g.lout[7], n, nullptr, nullptr, s.stream);
        hm::fc_head_fp32_kernel<<<(n + 7) / 8, 256, 0, s.stream>>>(s.d_act[7], m.fc1_wt, m.fc1_b, m.fc2_w, m.fc2_b, s.d_site_out, f, n, s.d_logits, s.d_ml);
        launches += 10;
        HM_CUDA(e, "fp32 CNN", cudaGetLastError());
    }
    return HM_OK;
}

void build_chunk_table(Slot& s, uint32_t n_reads)
{
    uint32_t nc = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        s.h_read_first_chunk[r] = nc;
        uint32_t L = s.host.base_off[r + 1] - s.host.base_off[r];
        for (uint32_t p = 0; p < L; p += hm::kChunk) {
            s.h_chunk_read[nc] = r;
            s.h_chunk_pos[nc] = p;
            ++nc;
        }
    }
    s.h_read_first_chunk[n_reads] = nc;
    s.n_chunks = nc;
}

int stage_front(hm_engine* e, Slot& s, uint32_t& launches)
{
    const uint32_t nc = s.n_chunks;
    if (nc) {
        hm::decode_kernel<<<nc, hm::kFrontThreads, 0, s.stream>>>(s.d_seq4, s.d_fi, s.d_fp, s.d_ri, s.d_rp, s.d_base_off, s.d_seq_off, s.d_flag,
                                                                 s.d_valid, s.d_chunk_read, s.d_chunk_pos, e->ctx_mask, s.d_bcode, s.d_kinf,
                                                                 s.d_chunk_cnt, s.d_cls);
        ++launches;
    }
    HM_CUDA(e, "decode", cudaGetLastError());
    HM_CUDA(e, "decode", cudaEventRecord(s.ev[2], s.stream));
    hm::scan_offsets_kernel<<<1, 1024, 0, s.stream>>>(s.d_chunk_cnt, nc, s.d_read_first_chunk, s.n_reads, s.d_pref, s.d_totals, s.d_call_off,
                                                      s.d_n_fwd, s.d_read_pref);
    ++launches;
    if (nc) {
        hm::scan_write_kernel<<<nc, hm::kFrontThreads, 0, s.stream>>>(s.d_cls, s.d_chunk_read, s.d_chunk_pos, s.d_read_first_chunk, s.d_pref, nc,
                                                                     s.d_qoff, s.d_call_ctx, s.d_site_read, s.d_site_pos, s.d_site_out);
        ++launches;
    }
    HM_CUDA(e, "scan", cudaGetLastError());
    HM_CUDA(e, "scan", cudaMemcpyAsync(s.h_totals, s.d_totals, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
    HM_CUDA(e, "scan", cudaMemcpyAsync(s.h_read_pref, s.d_read_pref, 4 * (size_t)(s.n_reads + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
    HM_CUDA(e, "scan", cudaEventRecord(s.ev[3], s.stream));
    return HM_OK;
}

int stage_cnn(hm_engine* e, Slot& s, uint32_t& launches)
{
    // class regions of the site list: [CpG | CHG | CHH fwd | CHH rev]
    const uint32_t t0 = s.totals[0], t1 = s.totals[1], t23 = s.totals[2] + s.totals[3];
    const uint32_t first[3] = {0, t0, t0 + t1};
    const uint32_t cnt[3] = {t0, t1, t23};
    if (e->cfg.cnn_mode != HM_CNN_FP32_SIMT) {
        hm::TensorBatch tb{};
        tb.d_bcode = s.d_bcode; tb.d_kinf = s.d_kinf; tb.d_base_off = s.d_base_off;
        tb.d_site_read = s.d_site_read; tb.d_site_pos = s.d_site_pos; tb.d_site_out = s.d_site_out;
        tb.h_base_off = s.host.base_off; tb.h_valid = s.host.valid; tb.h_read_pref = s.h_read_pref; tb.n_reads = s.n_reads;
        for (int k = 0; k < 4; ++k) tb.class_count[k] = s.totals[k];
        tb.d_logits = s.d_logits; tb.d_ml = s.d_ml;
        if (hm::tensor_batch_run(e->tensor, e->ctx_mask, s.tws, tb, s.stream, e->sm_count, &launches, &s.timing))
            return fail(e, HM_ERR_CUDA, "CUDA error in tensor CNN: %s", hm::tensor_last_error());
        return HM_OK;
    }
    for (int c = 0; c < 3; ++c) {
        if (!cnt[c]) continue;
        int rc = run_cnn_fp32(e, s, c, first[c], cnt[c], launches);
        if (rc) return rc;
    }
    return HM_OK;
}

}  // namespace

extern "C" {

const char* hm_version(void) { return "hifimeth-b200 0.1.0 (sm_100a)"; }

const char* hm_last_error(const hm_engine* e)
{
    if (e) return e->err.c_str();
    // a copy private to the calling thread: several device workers may fail in hm_engine_create at once, and a pointer into the
    // shared string could dangle as soon as the lock is dropped
    thread_local std::string copy;
    std::lock_guard<std::mutex> g(g_err_mu);
    copy = g_create_error;
    return copy.c_str();
}

void hm_engine_destroy(hm_engine* e)
{
    if (!e) return;
    const bool verbose = getenv("HM_VERBOSE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    cudaSetDevice(e->cfg.device);
    for (auto& s : e->slots) free_slot(s);
    if (verbose)
        fprintf(stderr, "[hm_engine_destroy] device %d: slots freed after %.3f s\n", e->cfg.device,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    for (int c = 0; c < 3; ++c) {
        Fp32Model& d = e->fp32[c];
        cudaFree(d.bn_scale); cudaFree(d.bn_shift);
        for (int l = 0; l < 8; ++l) { cudaFree(d.conv_wt[l]); cudaFree(d.conv_b[l]); }
        cudaFree(d.fc1_wt); cudaFree(d.fc1_b); cudaFree(d.fc2_w); cudaFree(d.fc2_b);
        hm::tensor_model_free(e->tensor[c]);
    }
    delete e;
}

int hm_engine_create(const hm_config* cfg, hm_engine** out)
{
    if (!cfg || !out || !cfg->model_dir) return fail(nullptr, HM_ERR_ARG, "hm_engine_create: null argument");
    *out = nullptr;
    if (cfg->n_slots < 1 || cfg->n_slots > 4 || cfg->max_reads == 0 || cfg->max_bases == 0 || cfg->max_bases >= 0x7fffffffu)
        return fail(nullptr, HM_ERR_ARG, "hm_engine_create: bad capacities (slots %d, reads %u, bases %u)", cfg->n_slots, cfg->max_reads, cfg->max_bases);
    const auto t_enter = std::chrono::steady_clock::now();
    int ndev = 0;
    cudaError_t st = cudaGetDeviceCount(&ndev);
    const auto t_count = std::chrono::steady_clock::now();
    if (st != cudaSuccess || ndev == 0)
        return fail(nullptr, HM_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", st == cudaSuccess ? "device count 0" : cudaGetErrorString(st));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, HM_ERR_ARG, "device %d out of range (0..%d)", cfg->device, ndev - 1);
    cudaDeviceProp prop;
    if ((st = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return fail(nullptr, HM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(st));
    if (prop.major != 10) return fail(nullptr, HM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    const auto t_prop = std::chrono::steady_clock::now();
    if ((st = cudaSetDevice(cfg->device)) != cudaSuccess) return fail(nullptr, HM_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(st));

    // HM_VERBOSE=1: wall-clock of the creation stages on stderr (context, models + plan lowering, slot allocation)
    const bool verbose = getenv("HM_VERBOSE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = now();
    cudaFree(nullptr);  // forces context creation here so that it is timed on its own
    const auto t1 = now();
    hm_engine* e = new hm_engine();
    e->cfg = *cfg;
    e->model_dir = cfg->model_dir;
    e->cfg.model_dir = e->model_dir.c_str();
    e->ctx_mask = (cfg->ctx_mask & 7) ? (uint32_t)(cfg->ctx_mask & 7) : 7u;
    e->n_slots = cfg->n_slots;
    e->sm_count = prop.multiProcessorCount;
    // The CPU reference loads <dir>/{CpG,CHG,CHH}.onnx (mod_main.cpp:76,85,94), its app-gpu binary the .pt exports of the same
    // networks (5mc_call_gpu.cpp:48).  .onnx is used when present; HM_MODEL_FORMAT=pt (or a directory holding only .pt files)
    // selects the TorchScript archives.  Note that the shipped CHG.pt is a different checkpoint from CHG.onnx.
    static const char* names[3] = {"CpG", "CHG", "CHH"};
    const char* fmt_env = getenv("HM_MODEL_FORMAT");
    int rc = HM_OK;
    for (int c = 0; c < 3 && rc == HM_OK; ++c) {
        if (!(e->ctx_mask & (1u << c))) continue;
        std::string err;
        std::string path = e->model_dir + "/" + names[c] + ".onnx";
        bool use_pt = fmt_env && std::string(fmt_env) == "pt";
        if (!use_pt && !fmt_env) {
            FILE* probe = fopen(path.c_str(), "rb");
            if (probe) fclose(probe);
            else {
                FILE* p2 = fopen((e->model_dir + "/" + names[c] + ".pt").c_str(), "rb");
                if (p2) { fclose(p2); use_pt = true; }
            }
        }
        if (use_pt) path = e->model_dir + "/" + names[c] + ".pt";
        if (!hm::load_model_file(path, e->host_model[c], err)) { rc = fail(e, HM_ERR_MODEL, "%s", err.c_str()); break; }
        if ((rc = check_geometry(e, c))) break;
        e->have_model[c] = true;
        if (cfg->cnn_mode == HM_CNN_FP32_SIMT) rc = build_fp32_model(e, c);
        else if (hm::tensor_model_build(e->tensor[c], e->host_model[c])) rc = fail(e, HM_ERR_CUDA, "CUDA error in tensor model build: %s", hm::tensor_last_error());
    }
    const auto t2 = now();
    if (rc == HM_OK) {
        e->slots.resize(e->n_slots);
        for (auto& s : e->slots)
            if ((rc = alloc_slot(e, s))) break;
    }
    if (verbose)
        fprintf(stderr, "[hm_engine_create] device %d: runtime start (cudaGetDeviceCount) %.3f s, cudaGetDeviceProperties %.3f s, cudaSetDevice %.3f s, "
                        "context (cudaFree(0)) %.3f s, models + plan %.3f s, %d slot(s) %.3f s\n",
                cfg->device, secs(t_enter, t_count), secs(t_count, t_prop), secs(t_prop, t0), secs(t0, t1), secs(t1, t2), e->n_slots, secs(t2, now()));
    if (rc != HM_OK) {
        fail(nullptr, rc, "%s", e->err.c_str());
        hm_engine_destroy(e);
        return rc;
    }
    *out = e;
    return HM_OK;
}

int hm_model_weights(const char* path, float* out, size_t cap, size_t* n_floats, int32_t* conv1_k)
{
    if (!path || !n_floats) return fail(nullptr, HM_ERR_ARG, "hm_model_weights: null argument");
    hm::CnnModel m;
    std::string err;
    if (!hm::load_model_file(path, m, err)) return fail(nullptr, HM_ERR_MODEL, "%s", err.c_str());
    std::vector<float> v;
    auto add = [&](const std::vector<float>& a) { v.insert(v.end(), a.begin(), a.end()); };
    add(m.bn_w); add(m.bn_b); add(m.bn_mean); add(m.bn_var);
    for (const hm::ConvLayer& cv : m.convs) { add(cv.w); add(cv.b); }
    add(m.fc1_w); add(m.fc1_b); add(m.fc2_w); add(m.fc2_b);
    *n_floats = v.size();
    if (conv1_k) *conv1_k = m.convs.empty() ? 0 : m.convs[0].k;
    if (out) {
        if (cap < v.size()) return fail(nullptr, HM_ERR_ARG, "hm_model_weights: %zu floats needed", v.size());
        memcpy(out, v.data(), v.size() * sizeof(float));
    }
    return HM_OK;
}

int hm_batch_acquire(hm_engine* e, int slot, hm_read_batch* out)
{
    if (!e || !out || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_batch_acquire: bad slot %d", slot);
    Slot& s = e->slots[slot];
    cudaSetDevice(e->cfg.device);
    if (s.submitted) HM_CUDA(e, "acquire", cudaStreamSynchronize(s.stream));
    s.submitted = s.collected = s.resident = false;
    *out = s.host;
    return HM_OK;
}

int hm_batch_submit(hm_engine* e, int slot, uint32_t n_reads, uint32_t flags)
{
    if (!e || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_batch_submit: bad slot %d", slot);
    Slot& s = e->slots[slot];
    cudaSetDevice(e->cfg.device);
    const bool skip_h2d = (flags & HM_SUBMIT_SKIP_H2D) != 0;
    if (skip_h2d && (!s.resident || n_reads != s.n_reads)) return fail(e, HM_ERR_STATE, "hm_batch_submit: SKIP_H2D but slot %d holds no resident batch of %u reads", slot, n_reads);
    if (n_reads > e->cfg.max_reads) return fail(e, HM_ERR_ARG, "hm_batch_submit: %u reads exceed capacity %u", n_reads, e->cfg.max_reads);
    if (s.submitted && !s.collected) HM_CUDA(e, "submit", cudaStreamSynchronize(s.stream));
    s.timing = hm_timing{};
    uint32_t launches = 0;
    cudaStream_t st = s.stream;
    HM_CUDA(e, "submit", cudaEventRecord(s.ev[0], st));
    if (!skip_h2d) {
        if (n_reads && s.host.base_off[0] != 0) return fail(e, HM_ERR_ARG, "hm_batch_submit: base_off[0] must be 0");
        if (n_reads && s.host.seq_off[0] != 0) return fail(e, HM_ERR_ARG, "hm_batch_submit: seq_off[0] must be 0");
        const uint32_t nb = n_reads ? s.host.base_off[n_reads] : 0;
        const uint32_t nsb = n_reads ? s.host.seq_off[n_reads] : 0;
        if (nb > e->cfg.max_bases) return fail(e, HM_ERR_ARG, "hm_batch_submit: %u bases exceed capacity %u", nb, e->cfg.max_bases);
        // the staging and device SEQ buffers hold max_bases / 2 + max_reads + 16 bytes (alloc_slot); callers may fill the SoA directly
        const size_t seq_cap = (size_t)e->cfg.max_bases / 2 + e->cfg.max_reads + 16;
        if (nsb > seq_cap) return fail(e, HM_ERR_ARG, "hm_batch_submit: %u packed SEQ bytes exceed capacity %zu", nsb, seq_cap);
        for (uint32_t r = 0; r < n_reads; ++r) {
            uint32_t L = s.host.base_off[r + 1] - s.host.base_off[r];
            if (s.host.base_off[r + 1] < s.host.base_off[r] || s.host.seq_off[r + 1] < s.host.seq_off[r] ||
                s.host.seq_off[r + 1] - s.host.seq_off[r] < (L + 1) / 2)
                return fail(e, HM_ERR_ARG, "hm_batch_submit: inconsistent offsets at read %u", r);
            if ((int32_t)L < e->cfg.min_read_len) s.host.valid[r] = 0;  // -l, mod_main.cpp:189-192
        }
        s.n_reads = n_reads;
        s.n_bases = nb;
        build_chunk_table(s, n_reads);
        const char* stg = "H2D";
        uint64_t bytes = 0;
        auto cp = [&](void* d, const void* h, size_t n) { bytes += n; return n ? cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st) : cudaSuccess; };
        HM_CUDA(e, stg, cp(s.d_base_off, s.host.base_off, (n_reads + 1) * sizeof(uint32_t)));
        HM_CUDA(e, stg, cp(s.d_seq_off, s.host.seq_off, (n_reads + 1) * sizeof(uint32_t)));
        HM_CUDA(e, stg, cp(s.d_seq4, s.host.seq4, nsb));
        HM_CUDA(e, stg, cp(s.d_flag, s.host.flag, n_reads * sizeof(uint16_t)));
        HM_CUDA(e, stg, cp(s.d_valid, s.host.valid, n_reads));
        HM_CUDA(e, stg, cp(s.d_fi, s.host.fi, nb));
        HM_CUDA(e, stg, cp(s.d_fp, s.host.fp, nb));
        HM_CUDA(e, stg, cp(s.d_ri, s.host.ri, nb));
        HM_CUDA(e, stg, cp(s.d_rp, s.host.rp, nb));
        HM_CUDA(e, stg, cp(s.d_chunk_read, s.h_chunk_read, s.n_chunks * sizeof(uint32_t)));
        HM_CUDA(e, stg, cp(s.d_chunk_pos, s.h_chunk_pos, s.n_chunks * sizeof(uint32_t)));
        HM_CUDA(e, stg, cp(s.d_read_first_chunk, s.h_read_first_chunk, (n_reads + 1) * sizeof(uint32_t)));
        s.timing.h2d_bytes = bytes;
        s.resident = true;
    }
    HM_CUDA(e, "submit", cudaEventRecord(s.ev[1], st));
    int rc = stage_front(e, s, launches);
    if (rc) return rc;
    // The CNN launch geometry depends on the site counts: one short host sync on the scan totals.
    HM_CUDA(e, "scan", cudaEventSynchronize(s.ev[3]));
    memcpy(s.totals, s.h_totals, sizeof(s.totals));
    s.n_calls = s.totals[4];
    if (s.n_calls > e->cfg.max_bases) return fail(e, HM_ERR_STATE, "internal: %u calls exceed capacity", s.n_calls);
    const bool want_mm = (flags & HM_SUBMIT_MM_TEXT) != 0;
    s.mm_valid = false;
    if (want_mm) {
        // row N1: MM skip counts + their decimal text from the resident forward-strand codes (independent of the CNN)
        *s.h_mm_total = 0;
        if (s.n_calls) {
            const uint32_t nb = (s.n_calls + hm::kMmBlock - 1) / hm::kMmBlock;
            hm::mm_delta_kernel<<<nb, hm::kMmBlock, 0, st>>>(s.d_bcode, s.d_base_off, s.d_call_off, s.d_n_fwd, s.d_qoff, s.n_reads, s.n_calls,
                                                            s.d_mm_delta, s.d_mm_bsum);
            hm::mm_scan_blocks_kernel<<<1, 1024, 0, st>>>(s.d_mm_bsum, nb, s.d_mm_boff);
            hm::mm_write_kernel<<<nb, hm::kMmBlock, 0, st>>>(s.d_mm_delta, s.d_mm_boff, s.n_calls, (uint32_t)s.mm_text_cap, s.d_mm_toff, s.d_mm_text);
            hm::mm_read_offsets_kernel<<<(s.n_reads + 256) / 256, 256, 0, st>>>(s.d_mm_toff, s.d_call_off, s.d_n_fwd, s.n_reads, s.d_mm_off, s.d_mm_fwd_len);
            launches += 4;
            HM_CUDA(e, "MM text", cudaGetLastError());
            HM_CUDA(e, "MM text", cudaMemcpyAsync(s.h_mm_total, s.d_mm_toff + s.n_calls, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        } else {
            HM_CUDA(e, "MM text", cudaMemsetAsync(s.d_mm_off, 0, (s.n_reads + 1) * sizeof(uint32_t), st));
            HM_CUDA(e, "MM text", cudaMemsetAsync(s.d_mm_fwd_len, 0, std::max<size_t>(s.n_reads, 1) * sizeof(uint32_t), st));
        }
        HM_CUDA(e, "MM text", cudaEventRecord(s.ev[6], st));
    }
    s.stats_valid = false;
    const bool want_stats = (flags & HM_SUBMIT_READ_STATS) != 0;
    if (want_stats && s.n_reads) {
        // row A9 (diagnostics): sum / max of the decoded frames per read
        HM_CUDA(e, "read stats", cudaMemsetAsync(s.d_read_stats, 0, s.n_reads * sizeof(hm::ReadKinStats), st));
        if (s.n_chunks) {
            hm::read_stats_kernel<<<s.n_chunks, hm::kFrontThreads, 0, st>>>(s.d_kinf, s.d_base_off, s.d_chunk_read, s.d_chunk_pos, s.d_read_stats);
            ++launches;
        }
        HM_CUDA(e, "read stats", cudaGetLastError());
    }
    if ((rc = stage_cnn(e, s, launches))) return rc;
    s.hist_valid = false;
    const bool want_hist = (flags & HM_SUBMIT_ML_HIST) != 0;
    if (want_hist) {
        // row N3: per-context histograms of the ML bytes while they are on the device
        HM_CUDA(e, "ML histogram", cudaMemsetAsync(s.d_ml_hist, 0, 3 * 256 * sizeof(uint32_t), st));
        if (s.n_calls) {
            const uint32_t nb = std::min<uint32_t>((s.n_calls + 255u) / 256u, (uint32_t)e->sm_count * 8u);
            hm::ml_hist_kernel<<<nb, 256, 0, st>>>(s.d_site_read, s.d_site_out, s.d_ml, s.d_flag, s.d_totals, s.n_calls, s.d_ml_hist);
            ++launches;
            HM_CUDA(e, "ML histogram", cudaGetLastError());
        }
    }
    HM_CUDA(e, "CNN", cudaEventRecord(s.ev[4], st));
    if (!(flags & HM_SUBMIT_SKIP_D2H)) {
        const char* stg = "D2H";
        uint64_t bytes = 0;
        auto cp = [&](void* h, const void* d, size_t n) { bytes += n; return n ? cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st) : cudaSuccess; };
        HM_CUDA(e, stg, cp(s.h_call_off, s.d_call_off, (s.n_reads + 1) * sizeof(uint32_t)));
        HM_CUDA(e, stg, cp(s.h_n_fwd, s.d_n_fwd, s.n_reads * sizeof(uint32_t)));
        HM_CUDA(e, stg, cp(s.h_qoff, s.d_qoff, (size_t)s.n_calls * sizeof(int32_t)));
        HM_CUDA(e, stg, cp(s.h_ml, s.d_ml, s.n_calls));
        if (want_mm) {
            HM_CUDA(e, stg, cudaEventSynchronize(s.ev[6]));  // total text bytes; those kernels ran ahead of the CNN
            if (*s.h_mm_total > s.mm_text_cap) return fail(e, HM_ERR_STATE, "MM text of %u bytes exceeds the staging capacity", *s.h_mm_total);
            HM_CUDA(e, stg, cp(s.h_mm_text, s.d_mm_text, *s.h_mm_total));
            HM_CUDA(e, stg, cp(s.h_mm_off, s.d_mm_off, (s.n_reads + 1) * sizeof(uint32_t)));
            HM_CUDA(e, stg, cp(s.h_mm_fwd_len, s.d_mm_fwd_len, s.n_reads * sizeof(uint32_t)));
            s.mm_valid = true;
        }
        if (want_hist) {
            HM_CUDA(e, stg, cp(s.h_ml_hist, s.d_ml_hist, 3 * 256 * sizeof(uint32_t)));
            s.hist_valid = true;
        }
        if (want_stats) {
            HM_CUDA(e, stg, cp(s.h_read_stats, s.d_read_stats, s.n_reads * sizeof(hm::ReadKinStats)));
            s.stats_valid = true;
        }
        s.timing.d2h_bytes = bytes + 5 * sizeof(uint32_t);
    }
    HM_CUDA(e, "submit", cudaEventRecord(s.ev[5], st));
    s.timing.kernel_launches = launches;
    s.submitted = true;
    s.collected = false;
    return HM_OK;
}

int hm_batch_collect(hm_engine* e, int slot, hm_call_batch* out)
{
    if (!e || !out || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_batch_collect: bad slot %d", slot);
    Slot& s = e->slots[slot];
    if (!s.submitted) return fail(e, HM_ERR_STATE, "hm_batch_collect: slot %d has no submitted batch", slot);
    cudaSetDevice(e->cfg.device);
    HM_CUDA(e, "collect", cudaEventSynchronize(s.ev[5]));
    HM_CUDA(e, "collect", cudaStreamSynchronize(s.stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, s.ev[0], s.ev[1]); s.timing.h2d_ms = ms;
    cudaEventElapsedTime(&ms, s.ev[1], s.ev[2]); s.timing.decode_ms = ms;
    cudaEventElapsedTime(&ms, s.ev[2], s.ev[3]); s.timing.scan_ms = ms;
    cudaEventElapsedTime(&ms, s.ev[3], s.ev[4]); s.timing.cnn_ms = ms;
    cudaEventElapsedTime(&ms, s.ev[4], s.ev[5]); s.timing.d2h_ms = ms;
    cudaEventElapsedTime(&ms, s.ev[0], s.ev[5]); s.timing.total_ms = ms;
    if (e->cfg.cnn_mode != HM_CNN_FP32_SIMT) s.timing.top_kernel_ms = hm::tensor_last_dense_ms(s.tws);
    out->n_reads = s.n_reads;
    out->n_calls = s.n_calls;
    out->call_off = s.h_call_off;
    out->n_fwd = s.h_n_fwd;
    out->qoff = s.h_qoff;
    out->ml = s.h_ml;
    out->n_sites[0] = s.totals[0];
    out->n_sites[1] = s.totals[1];
    out->n_sites[2] = (uint64_t)s.totals[2] + s.totals[3];
    out->mm_text = s.mm_valid ? s.h_mm_text : nullptr;
    out->mm_off = s.mm_valid ? s.h_mm_off : nullptr;
    out->mm_fwd_len = s.mm_valid ? s.h_mm_fwd_len : nullptr;
    out->ml_hist = s.hist_valid ? s.h_ml_hist : nullptr;
    static_assert(sizeof(hm::ReadKinStats) == sizeof(hm_read_stats), "hm_read_stats layout");
    out->read_stats = s.stats_valid ? reinterpret_cast<const hm_read_stats*>(s.h_read_stats) : nullptr;
    s.collected = true;
    return HM_OK;
}

int hm_batch_timing(hm_engine* e, int slot, hm_timing* out)
{
    if (!e || !out || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_batch_timing: bad slot %d", slot);
    if (!e->slots[slot].collected) return fail(e, HM_ERR_STATE, "hm_batch_timing: collect slot %d first", slot);
    *out = e->slots[slot].timing;
    return HM_OK;
}

// ---- validation hooks ----------------------------------------------------------------------------------------

int hm_debug_dump_decode(hm_engine* e, int slot, uint16_t* fi, uint16_t* fp, uint16_t* ri, uint16_t* rp, uint8_t* fwd_qs, uint8_t* rev_qs)
{
    if (!e || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_debug_dump_decode: bad slot");
    Slot& s = e->slots[slot];
    if (!s.collected) return fail(e, HM_ERR_STATE, "hm_debug_dump_decode: collect slot %d first", slot);
    if (!e->cfg.keep_debug) return fail(e, HM_ERR_STATE, "hm_debug_dump_decode: the engine was created without hm_config.keep_debug");
    cudaSetDevice(e->cfg.device);
    const size_t nb = s.n_bases;
    uint16_t* d16 = nullptr;
    uint8_t* d8 = nullptr;
    HM_CUDA(e, "debug decode", dmalloc(&d16, 4 * nb));
    HM_CUDA(e, "debug decode", dmalloc(&d8, 2 * nb));
    if (s.n_chunks)
        hm::decode_unpack_kernel<<<s.n_chunks, 256, 0, s.stream>>>(s.d_bcode, s.d_kinf, s.d_base_off, s.d_chunk_read, s.d_chunk_pos,
                                                                   d16, d16 + nb, d16 + 2 * nb, d16 + 3 * nb, d8, d8 + nb);
    HM_CUDA(e, "debug decode", cudaGetLastError());
    HM_CUDA(e, "debug decode", cudaStreamSynchronize(s.stream));
    uint16_t* dst16[4] = {fi, fp, ri, rp};
    for (int k = 0; k < 4; ++k)
        if (dst16[k]) HM_CUDA(e, "debug decode", cudaMemcpy(dst16[k], d16 + k * nb, nb * 2, cudaMemcpyDeviceToHost));
    if (fwd_qs) HM_CUDA(e, "debug decode", cudaMemcpy(fwd_qs, d8, nb, cudaMemcpyDeviceToHost));
    if (rev_qs) HM_CUDA(e, "debug decode", cudaMemcpy(rev_qs, d8 + nb, nb, cudaMemcpyDeviceToHost));
    cudaFree(d16);
    cudaFree(d8);
    return HM_OK;
}

int hm_debug_dump_ctx(hm_engine* e, int slot, uint8_t* ctx)
{
    if (!e || !ctx || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_debug_dump_ctx: bad argument");
    Slot& s = e->slots[slot];
    if (!s.collected) return fail(e, HM_ERR_STATE, "hm_debug_dump_ctx: collect slot %d first", slot);
    if (!e->cfg.keep_debug) return fail(e, HM_ERR_STATE, "hm_debug_dump_ctx: the engine was created without hm_config.keep_debug");
    cudaSetDevice(e->cfg.device);
    HM_CUDA(e, "debug ctx", cudaMemcpy(ctx, s.d_call_ctx, s.n_calls, cudaMemcpyDeviceToHost));
    return HM_OK;
}

int hm_debug_dump_features(hm_engine* e, int slot, uint32_t first, uint32_t count, float* out)
{
    if (!e || !out || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_debug_dump_features: bad argument");
    Slot& s = e->slots[slot];
    if (!s.collected) return fail(e, HM_ERR_STATE, "hm_debug_dump_features: collect slot %d first", slot);
    if (!e->cfg.keep_debug) return fail(e, HM_ERR_STATE, "hm_debug_dump_features: the engine was created without hm_config.keep_debug");
    if ((uint64_t)first + count > s.n_calls) return fail(e, HM_ERR_ARG, "hm_debug_dump_features: range beyond %u calls", s.n_calls);
    if (!count) return HM_OK;
    cudaSetDevice(e->cfg.device);
    // site list is in class order; find the list slot of each requested call through site_out
    std::vector<uint32_t> site_out(s.n_calls), inv(s.n_calls);
    HM_CUDA(e, "debug features", cudaMemcpy(site_out.data(), s.d_site_out, (size_t)s.n_calls * 4, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < s.n_calls; ++i) inv[site_out[i]] = i;
    std::vector<uint32_t> h_read(s.n_calls), h_pos(s.n_calls);
    HM_CUDA(e, "debug features", cudaMemcpy(h_read.data(), s.d_site_read, (size_t)s.n_calls * 4, cudaMemcpyDeviceToHost));
    HM_CUDA(e, "debug features", cudaMemcpy(h_pos.data(), s.d_site_pos, (size_t)s.n_calls * 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> sel_read(count), sel_pos(count);
    for (uint32_t k = 0; k < count; ++k) { sel_read[k] = h_read[inv[first + k]]; sel_pos[k] = h_pos[inv[first + k]]; }
    uint32_t *d_r = nullptr, *d_p = nullptr;
    float* d_f = nullptr;
    HM_CUDA(e, "debug features", dmalloc(&d_r, count));
    HM_CUDA(e, "debug features", dmalloc(&d_p, count));
    HM_CUDA(e, "debug features", dmalloc(&d_f, (size_t)count * HM_KMER * HM_FEATURES_PER_BASE));
    HM_CUDA(e, "debug features", cudaMemcpy(d_r, sel_read.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    HM_CUDA(e, "debug features", cudaMemcpy(d_p, sel_pos.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    hm::gather_features_kernel<<<count, 128, 0, s.stream>>>(s.d_bcode, s.d_kinf, s.d_base_off, d_r, d_p, 0, count, d_f);
    HM_CUDA(e, "debug features", cudaGetLastError());
    HM_CUDA(e, "debug features", cudaStreamSynchronize(s.stream));
    HM_CUDA(e, "debug features", cudaMemcpy(out, d_f, (size_t)count * HM_KMER * HM_FEATURES_PER_BASE * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_r); cudaFree(d_p); cudaFree(d_f);
    return HM_OK;
}

int hm_debug_dump_logits(hm_engine* e, int slot, float* out)
{
    if (!e || !out || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_debug_dump_logits: bad argument");
    Slot& s = e->slots[slot];
    if (!s.collected) return fail(e, HM_ERR_STATE, "hm_debug_dump_logits: collect slot %d first", slot);
    if (!e->cfg.keep_debug) return fail(e, HM_ERR_STATE, "hm_debug_dump_logits: the engine was created without hm_config.keep_debug");
    cudaSetDevice(e->cfg.device);
    HM_CUDA(e, "debug logits", cudaMemcpy(out, s.d_logits, (size_t)s.n_calls * 2 * sizeof(float), cudaMemcpyDeviceToHost));
    return HM_OK;
}

namespace {
// Site-list view of calls [first, first + count) of a collected batch (hm_call_batch order): read, strand, strand offset o,
// context and the row the site owns in the compact maps of its context (one sub-batch, one group).
struct DebugSites {
    std::vector<uint32_t> read, compact_row;
    std::vector<uint8_t> rev, ctx;
    std::vector<int32_t> o;
};
int debug_sites(hm_engine* e, Slot& s, uint32_t first, uint32_t count, DebugSites& d)
{
    std::vector<uint32_t> site_out(s.n_calls), inv(s.n_calls), h_read(s.n_calls), h_pos(s.n_calls);
    HM_CUDA(e, "debug sites", cudaMemcpy(site_out.data(), s.d_site_out, (size_t)s.n_calls * 4, cudaMemcpyDeviceToHost));
    HM_CUDA(e, "debug sites", cudaMemcpy(h_read.data(), s.d_site_read, (size_t)s.n_calls * 4, cudaMemcpyDeviceToHost));
    HM_CUDA(e, "debug sites", cudaMemcpy(h_pos.data(), s.d_site_pos, (size_t)s.n_calls * 4, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < s.n_calls; ++i) inv[site_out[i]] = i;
    const uint32_t f1 = s.totals[0], f2 = f1 + s.totals[1], f3 = f2 + s.totals[2];
    d.read.resize(count); d.compact_row.resize(count); d.rev.resize(count); d.ctx.resize(count); d.o.resize(count);
    for (uint32_t i = 0; i < count; ++i) {
        const uint32_t k = inv[first + i];
        const uint32_t r = h_read[k], sp = h_pos[k];
        const int L = (int)(s.host.base_off[r + 1] - s.host.base_off[r]), p = (int)(sp & 0x7fffffffu);
        d.read[i] = r;
        d.rev[i] = (uint8_t)(sp >> 31);
        d.o[i] = d.rev[i] ? L - 1 - p : p;
        d.ctx[i] = (uint8_t)(k < f1 ? 0 : k < f2 ? 1 : 2);
        d.compact_row[i] = k < f1 ? k : k < f2 ? k - f1 : k < f3 ? k - f2 : s.totals[2] + (k - f3);
    }
    return HM_OK;
}
}  // namespace

int hm_debug_dump_xmap(hm_engine* e, int slot, uint32_t first, uint32_t count, float* out)
{
    if (!e || !out || slot < 0 || slot >= e->n_slots) return fail(e, HM_ERR_ARG, "hm_debug_dump_xmap: bad argument");
    Slot& s = e->slots[slot];
    if (!s.collected) return fail(e, HM_ERR_STATE, "hm_debug_dump_xmap: collect slot %d first", slot);
    if (!e->cfg.keep_debug) return fail(e, HM_ERR_STATE, "hm_debug_dump_xmap: the engine was created without hm_config.keep_debug");
    if (e->cfg.cnn_mode != HM_CNN_TENSOR) return fail(e, HM_ERR_STATE, "hm_debug_dump_xmap: only the tensor path keeps an X map");
    if ((uint64_t)first + count > s.n_calls) return fail(e, HM_ERR_ARG, "hm_debug_dump_xmap: range beyond %u calls", s.n_calls);
    if (!count) return HM_OK;
    cudaSetDevice(e->cfg.device);
    DebugSites d;
    int rc = debug_sites(e, s, first, count, d);
    if (rc) return rc;
    if (hm::tensor_debug_xwindow(s.tws, count, d.read.data(), d.rev.data(), d.o.data(), out, s.stream))
        return fail(e, HM_ERR_CUDA, "%s", hm::tensor_last_error());
    return HM_OK;
}

int hm_debug_dump_acts(hm_engine* e, int slot, int ctx, int layer, uint32_t first, uint32_t count, float* out, size_t out_floats,
                       int32_t* n_pos, int32_t* channels)
{
    if (!e || !out || !n_pos || !channels || slot < 0 || slot >= e->n_slots || ctx < 0 || ctx > 2)
        return fail(e, HM_ERR_ARG, "hm_debug_dump_acts: bad argument");
    Slot& s = e->slots[slot];
    if (!s.collected) return fail(e, HM_ERR_STATE, "hm_debug_dump_acts: collect slot %d first", slot);
    if (!e->cfg.keep_debug) return fail(e, HM_ERR_STATE, "hm_debug_dump_acts: the engine was created without hm_config.keep_debug");
    if (e->cfg.cnn_mode != HM_CNN_TENSOR || !e->have_model[ctx]) return fail(e, HM_ERR_STATE, "hm_debug_dump_acts: tensor path with context %d enabled only", ctx);
    if ((uint64_t)first + count > s.n_calls || !count) return fail(e, HM_ERR_ARG, "hm_debug_dump_acts: bad range");
    cudaSetDevice(e->cfg.device);
    DebugSites d;
    int rc = debug_sites(e, s, first, count, d);
    if (rc) return rc;
    // calls of other contexts in the range are reported as NaN rows
    std::vector<uint32_t> pick;
    for (uint32_t i = 0; i < count; ++i)
        if (d.ctx[i] == ctx) pick.push_back(i);
    DebugSites q;
    for (uint32_t i : pick) { q.read.push_back(d.read[i]); q.rev.push_back(d.rev[i]); q.o.push_back(d.o[i]); q.compact_row.push_back(d.compact_row[i]); }
    // re-run this context alone so that every map holds ITS values (a batch run leaves the last context's behind)
    hm_timing keep = s.timing;
    uint32_t launches = 0;
    hm::TensorBatch tb{};
    tb.d_bcode = s.d_bcode; tb.d_kinf = s.d_kinf; tb.d_base_off = s.d_base_off;
    tb.d_site_read = s.d_site_read; tb.d_site_pos = s.d_site_pos; tb.d_site_out = s.d_site_out;
    tb.h_base_off = s.host.base_off; tb.h_valid = s.host.valid; tb.h_read_pref = s.h_read_pref; tb.n_reads = s.n_reads;
    for (int k = 0; k < 4; ++k) tb.class_count[k] = s.totals[k];
    tb.d_logits = s.d_logits; tb.d_ml = s.d_ml;
    hm::tensor_debug_spill(s.tws, true);  // F2.. T8 live in shared memory in the product path: this rerun also stores them
    const int run = hm::tensor_batch_run(e->tensor, 1u << ctx, s.tws, tb, s.stream, e->sm_count, &launches, &s.timing);
    hm::tensor_debug_spill(s.tws, false);
    s.timing = keep;
    if (run) return fail(e, HM_ERR_CUDA, "CUDA error in tensor CNN: %s", hm::tensor_last_error());
    HM_CUDA(e, "debug activations", cudaStreamSynchronize(s.stream));
    int nl = 0, C = 0;
    std::vector<float> tmp((size_t)std::max<size_t>(pick.size(), 1) * 197 * 128);
    if (hm::tensor_debug_site_acts(e->tensor[ctx], ctx, s.tws, (uint32_t)pick.size(), q.read.data(), q.rev.data(), q.o.data(), q.compact_row.data(),
                                   layer, tmp.data(), &nl, &C, s.stream))
        return fail(e, HM_ERR_STATE, "%s", hm::tensor_last_error());
    *n_pos = nl;
    *channels = C;
    const size_t per = (size_t)nl * C;
    if ((size_t)count * per > out_floats) return fail(e, HM_ERR_ARG, "hm_debug_dump_acts: output needs %zu floats", (size_t)count * per);
    for (size_t i = 0; i < (size_t)count * per; ++i) out[i] = NAN;
    for (size_t j = 0; j < pick.size(); ++j) memcpy(out + pick[j] * per, tmp.data() + j * per, per * sizeof(float));
    return HM_OK;
}

int hm_debug_dense_op(int device, uint32_t rows, uint32_t rows_alloc, int cin, int cout, int n_src, const float* const* src, int n_terms,
                      const int32_t* term_src, const int32_t* term_shift, const float* weights, const float* bias, int conv1_taps,
                      const float* w2, const float* b2, const uint32_t* gather_rows, uint32_t gather_mask, float* out)
{
    if (!src || !term_src || !term_shift || !weights || !bias || !out) return fail(nullptr, HM_ERR_ARG, "hm_debug_dense_op: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(nullptr, HM_ERR_CUDA, "hm_debug_dense_op: no such CUDA device");
    if (hm::tensor_debug_dense_op(device, rows, rows_alloc, cin, cout, n_src, src, n_terms, term_src, term_shift, weights, bias, conv1_taps, w2, b2, gather_rows, gather_mask, out))
        return fail(nullptr, HM_ERR_CUDA, "%s", hm::tensor_last_error());
    return HM_OK;
}

float hm_debug_last_op_ms(void) { return hm::tensor_debug_last_op_ms(); }

int hm_microbench(hm_engine* e, int slot, const char* name, uint32_t n_sites, int iters, float* ms_per_launch, double* algo_bytes, double* algo_flops)
{
    if (!e || !name || !ms_per_launch || slot < 0 || slot >= e->n_slots || iters < 1) return fail(e, HM_ERR_ARG, "hm_microbench: bad argument");
    Slot& s = e->slots[slot];
    if (!s.collected) return fail(e, HM_ERR_STATE, "hm_microbench: submit + collect a batch on slot %d first", slot);
    cudaSetDevice(e->cfg.device);
    cudaStream_t st = s.stream;
    const std::string k = name;
    double bytes = 0, flops = 0;
    uint32_t launches = 0;
    cudaEvent_t a = s.ev[6], b = s.ev[7];
    float* d_tmp = nullptr;
    const uint32_t ns = std::min<uint32_t>(n_sites ? n_sites : s.n_calls, s.n_calls);
    if (k == "gather") HM_CUDA(e, "microbench", dmalloc(&d_tmp, (size_t)std::max<uint32_t>(ns, 1) * HM_KMER * HM_FEATURES_PER_BASE));
    for (int it = -1; it < iters; ++it) {  // it = -1: warm-up
        if (it == 0) HM_CUDA(e, "microbench", cudaEventRecord(a, st));
        if (k == "decode") {
            // decode_kernel also classifies the positions it has in shared memory (0.5 B/base of class nibbles, not counted)
            if (s.n_chunks)
                hm::decode_kernel<<<s.n_chunks, hm::kFrontThreads, 0, st>>>(s.d_seq4, s.d_fi, s.d_fp, s.d_ri, s.d_rp, s.d_base_off, s.d_seq_off, s.d_flag,
                                                                         s.d_valid, s.d_chunk_read, s.d_chunk_pos, e->ctx_mask, s.d_bcode, s.d_kinf,
                                                                         s.d_chunk_cnt, s.d_cls);
            bytes = 12.0 * s.n_bases;  // 4 code planes in, 4 x u16 frames out
        } else if (k == "scan") {
            // the part of the site scan that is its own launches: prefix over chunks + site lists from the class nibbles
            hm::scan_offsets_kernel<<<1, 1024, 0, st>>>(s.d_chunk_cnt, s.n_chunks, s.d_read_first_chunk, s.n_reads, s.d_pref, s.d_totals, s.d_call_off,
                                                        s.d_n_fwd, s.d_read_pref);
            if (s.n_chunks)
                hm::scan_write_kernel<<<s.n_chunks, hm::kFrontThreads, 0, st>>>(s.d_cls, s.d_chunk_read, s.d_chunk_pos, s.d_read_first_chunk, s.d_pref,
                                                                             s.n_chunks, s.d_qoff, s.d_call_ctx, s.d_site_read, s.d_site_pos, s.d_site_out);
            bytes = 0.5 * s.n_bases + 5.0 * s.n_calls;  // SURVEY s8d; the lists the engine really writes are 17 B/site
        } else if (k == "stats") {
            if (s.n_chunks) hm::read_stats_kernel<<<s.n_chunks, hm::kFrontThreads, 0, st>>>(s.d_kinf, s.d_base_off, s.d_chunk_read, s.d_chunk_pos, s.d_read_stats);
            bytes = 8.0 * s.n_bases;  // the four decoded frames of every base read once
        } else if (k == "gather") {
            if (ns) hm::gather_features_kernel<<<ns, 128, 0, st>>>(s.d_bcode, s.d_kinf, s.d_base_off, s.d_site_read, s.d_site_pos, 0, ns, d_tmp);
            bytes = 12832.0 * ns;
        } else if (k == "mm") {
            if (s.n_calls) {
                const uint32_t nb = (s.n_calls + hm::kMmBlock - 1) / hm::kMmBlock;
                hm::mm_delta_kernel<<<nb, hm::kMmBlock, 0, st>>>(s.d_bcode, s.d_base_off, s.d_call_off, s.d_n_fwd, s.d_qoff, s.n_reads, s.n_calls,
                                                                s.d_mm_delta, s.d_mm_bsum);
                hm::mm_scan_blocks_kernel<<<1, 1024, 0, st>>>(s.d_mm_bsum, nb, s.d_mm_boff);
                hm::mm_write_kernel<<<nb, hm::kMmBlock, 0, st>>>(s.d_mm_delta, s.d_mm_boff, s.n_calls, (uint32_t)s.mm_text_cap, s.d_mm_toff, s.d_mm_text);
                hm::mm_read_offsets_kernel<<<(s.n_reads + 256) / 256, 256, 0, st>>>(s.d_mm_toff, s.d_call_off, s.d_n_fwd, s.n_reads, s.d_mm_off, s.d_mm_fwd_len);
            }
            // forward-strand codes read once (1 B/base), qoff in (4 B/call), delta out + in (8), text offset (4), ~2.3 B of text
            bytes = 1.0 * s.n_bases + 18.3 * s.n_calls;
        } else if (k == "cnn") {
            hm_timing keep = s.timing;
            int rc = stage_cnn(e, s, launches);
            s.timing = keep;
            if (rc) return rc;
            flops = 22297600.0 * (s.totals[0] + s.totals[1]) + 22881280.0 * ((double)s.totals[2] + s.totals[3]);
        } else {
            cudaFree(d_tmp);
            return fail(e, HM_ERR_ARG, "hm_microbench: unknown kernel family '%s'", name);
        }
        HM_CUDA(e, "microbench", cudaGetLastError());
    }
    HM_CUDA(e, "microbench", cudaEventRecord(b, st));
    HM_CUDA(e, "microbench", cudaEventSynchronize(b));
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    *ms_per_launch = ms / iters;
    if (algo_bytes) *algo_bytes = bytes;
    if (algo_flops) *algo_flops = flops;
    cudaFree(d_tmp);
    return HM_OK;
}

}  // extern "C"
