// cnn_tensor.cu -- tensor-core CNN path: the dilated dense plan, its packed weights, and the batch orchestration.
//
// Network (training/model_cnn.py:31-85 as exported in models/*.onnx): bn0 -> 8 x (conv1d stride 2, pad 1, +bias, ReLU)
// -> flatten -> fc1 + ReLU -> fc2.  The reference runs it once per site on a [401,8] window
// (src/app/hifimeth/mod_batch.cpp:66-75).  Windows of neighbouring sites overlap almost completely, so here every
// layer is evaluated ONCE per strand position as a dilated convolution over the whole read ("track"):
//
//     Y_l[i] = relu(b_l + sum_j W_l[j] . Y_{l-1}[i + j * 2^(l-1)])          l = 1..6,  Y_0 = X (raw features)
//
// Output v of layer l of the site whose window starts at row s + 1 (s = o - 201, o = the site's strand offset) is
// Y_l[s + v * 2^l - (2^l - 2)], except the first and last output of each layer, which see the zero padding.  Those
// depend on s alone and are dense maps as well (F_l, G_l); layers 7, 8 and the FC head are evaluated per row s
// (T7_v, T8_w, head).  bn0 is folded into conv1; the pad columns bn0 never touches are corrected in F_1 / G_1.
// tests/dense_emulator.py restates this plan in numpy and checks it against the per-site oracle (max |dlogit| 4e-6).
//
// Every op of the plan is one launch of dense_gemm_kernel.  Tracks of a batch are processed in sub-batches that fit
// the activation workspace; the final logits of all rows of a batch are kept (8 B/row) and a lookup kernel turns
// them into per-site logits + ML bytes in hm_call_batch order.
#include "cnn_tensor.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "dense_gemm.cuh"
#include "dense_gemm2.cuh"
#include "dense_fused12.cuh"
#include "postprocess.cuh"
#include "site_chain.cuh"

namespace hm {
namespace {

thread_local std::string g_err = "";  // engines of different devices run on different host threads
int tfail(const std::string& s) { g_err = s; return -1; }
#define TCUDA(stage, call)                                                                   \
    do {                                                                                     \
        cudaError_t _st = (call);                                                            \
        if (_st != cudaSuccess) return tfail(std::string(stage) + ": " + cudaGetErrorString(_st)); \
    } while (0)

constexpr int kHaloL = 208;   // rows before strand position 0 (>= 201, multiple of 16)
constexpr int kHaloR = 200;   // zero-feature rows after the last strand position
constexpr int kSlackRows = 1024;  // rows past the sub-batch every plane keeps allocated (shifted reads of the last tile)
constexpr size_t kSmemMax = 232448;
constexpr size_t kSmemAux = 3456;  // barriers + bias / fc2 staging of dense_gemm_kernel

// Default: dense Y chain + F/G/tail ops evaluated at site rows only ("compact").  HM_DENSE_ALL=1 evaluates every op on
// every row (the first version of this path; kept for A/B measurements).
bool compact_mode()
{
    static const bool v = getenv("HM_DENSE_ALL") == nullptr;
    return v;
}

// X and Y1..Y6 are dense (one row per strand position of the sub-batch); F, G, T7, T8 and the scatter buffers S hold one row
// per site of the running context (compact), except under HM_DENSE_ALL where F/G/T are dense too.
constexpr int kMaxScatterMaps = 20;
enum MapId { MAP_X = 0, MAP_Y = 1, MAP_F = 7, MAP_G = 13, MAP_T7 = 19, MAP_T8 = 23, MAP_S = 25, N_MAPS = 25 + kMaxScatterMaps };
const int kLayerCout[6] = {128, 128, 128, 96, 96, 96};

int map_channels(int id)
{
    if (id == MAP_X) return 8;
    if (id < MAP_F) return kLayerCout[id - MAP_Y];
    if (id < MAP_G) return kLayerCout[id - MAP_F];
    if (id < MAP_T7) return kLayerCout[id - MAP_G];
    if (id >= MAP_S) return 128;  // scatter buffers are sized for the widest map
    return 64;
}

// ---- host plan ---------------------------------------------------------------------------------------------------
struct HostTerm {
    int src = 0, shift = 0;
    bool gather = false;   // compact op: rows are picked through the site-row index (source is a dense map)
    std::vector<float> w;  // [cin][cout]; conv1 form: [taps][8][cout]
};
struct HostOp {
    int out = -1;
    int cin = 0, cout = 0;
    std::vector<float> bias;
    std::vector<HostTerm> terms;
    int conv1_taps = 0;  // > 0: conv1 form (one term, weights [taps][8][cout], rows shift .. shift + taps - 1)
    bool head = false;
    bool compact = false;  // evaluated at site rows only (one output row per site of the run)
    std::vector<std::pair<int, int>> scatter;  // dense op: (row shift, compact buffer) copies its epilogue also writes
    std::vector<float> w2, b2;
};

uint16_t f2bf(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float bf2f(uint16_t h)
{
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

bool build_plan(const CnnModel& m, bool compact, std::vector<HostOp>& ops, std::string& err)
{
    if (m.features != 8 || m.kmer != 401 || m.convs.size() != 8) { err = "unsupported model geometry"; return false; }
    const int k1 = m.convs[0].k;
    int lens[9];
    lens[0] = 401;
    for (int l = 0; l < 8; ++l) lens[l + 1] = (lens[l] + 2 - m.convs[l].k) / 2 + 1;
    if (lens[8] != 2 || 2 * (lens[1] - 1) + k1 - 1 != 402) { err = "unsupported conv geometry"; return false; }
    for (int l = 0; l < 6; ++l)
        if (m.convs[l].cout != kLayerCout[l]) { err = "unsupported channel counts"; return false; }
    if (m.convs[6].cout != 64 || m.convs[7].cout != 64 || m.fc1_in != 128 || m.fc1_out != 256) { err = "unsupported head"; return false; }

    double scale[8], shift[8];
    for (int c = 0; c < 8; ++c) {
        scale[c] = (double)m.bn_w[c] / std::sqrt((double)m.bn_var[c] + (double)m.bn_eps);
        shift[c] = (double)m.bn_b[c] - (double)m.bn_mean[c] * scale[c];
    }
    // ---- layer 1 (conv1 form; bn0 folded) ----------------------------------------------------------------------
    const ConvLayer& c1 = m.convs[0];
    auto w1 = [&](int o, int c, int j) { return (double)c1.w[((size_t)o * 8 + c) * k1 + j]; };
    std::vector<double> b1f(128);
    for (int o = 0; o < 128; ++o) {
        double s = c1.b[o];
        for (int c = 0; c < 8; ++c)
            for (int j = 0; j < k1; ++j) s += w1(o, c, j) * shift[c];
        b1f[o] = s;
    }
    auto conv1_op = [&](int out, int base_shift, int skip_tap) {
        HostOp op;
        op.out = out; op.cin = 8; op.cout = 128; op.conv1_taps = k1;
        op.compact = compact && out != MAP_Y;
        op.bias.resize(128);
        HostTerm t;
        t.src = MAP_X; t.shift = base_shift; t.gather = op.compact;
        t.w.assign((size_t)k1 * 8 * 128, 0.f);
        for (int o = 0; o < 128; ++o) {
            double b = b1f[o];
            for (int c = 0; c < 8; ++c) {
                for (int j = 0; j < k1; ++j)
                    if (j != skip_tap) t.w[((size_t)j * 8 + c) * 128 + o] = (float)(w1(o, c, j) * scale[c]);
                // the skipped tap is a conv zero-pad column: bn0 never saw it, so its folded shift goes away too
                if (skip_tap >= 0) b -= w1(o, c, skip_tap) * shift[c];
            }
            op.bias[o] = (float)b;
        }
        op.terms.push_back(std::move(t));
        ops.push_back(std::move(op));
    };
    conv1_op(MAP_Y + 0, 0, -1);
    conv1_op(MAP_F + 0, 0, 0);
    conv1_op(MAP_G + 0, 2 * (lens[1] - 1), k1 - 1);

    // A compact op that reads dense map `mp` at (site row + sh) gets that operand from a compact scatter buffer which the
    // producing dense op fills from its epilogue.
    int n_scatter_maps = 0;
    auto scatter_buffer = [&](int mp, int sh) {
        for (HostOp& pr : ops) {
            if (pr.out != mp || pr.compact) continue;
            for (auto& sc : pr.scatter)
                if (sc.first == sh) return sc.second;
            if ((int)pr.scatter.size() == kMaxScatter || n_scatter_maps == kMaxScatterMaps) return -1;
            pr.scatter.push_back({sh, MAP_S + n_scatter_maps});
            return MAP_S + n_scatter_maps++;
        }
        return -1;
    };
    auto off = [](int l) { return -((1 << l) - 2); };
    // site-level element q of layer l's output -> (map, shift) relative to row s; false = zero pad
    auto src = [&](int l, int q, int& map, int& sh) {
        const int n = lens[l];
        if (q < 0 || q >= n) return false;
        if (l <= 6) {
            if (q == 0) { map = MAP_F + l - 1; sh = 0; }
            else if (q == n - 1) { map = MAP_G + l - 1; sh = 0; }
            else { map = MAP_Y + l - 1; sh = off(l) + q * (1 << l); }
        } else { map = MAP_T7 + q; sh = 0; }
        return true;
    };
    for (int l = 2; l <= 8; ++l) {
        const ConvLayer& cv = m.convs[l - 1];
        if (cv.k != 3) { err = "unsupported kernel size"; return false; }
        auto tap = [&](int j) {
            std::vector<float> w((size_t)cv.cin * cv.cout);
            for (int o = 0; o < cv.cout; ++o)
                for (int c = 0; c < cv.cin; ++c) w[(size_t)c * cv.cout + o] = cv.w[((size_t)o * cv.cin + c) * 3 + j];
            return w;
        };
        auto new_op = [&](int out) {
            HostOp op;
            op.out = out; op.cin = cv.cin; op.cout = cv.cout; op.bias = cv.b;
            return op;
        };
        std::vector<std::pair<int, int>> outs;  // (map, site-level output index)
        if (l <= 6) {
            HostOp op = new_op(MAP_Y + l - 1);
            for (int j = 0; j < 3; ++j) {
                HostTerm t; t.src = MAP_Y + l - 2; t.shift = j * (1 << (l - 1)); t.w = tap(j);
                op.terms.push_back(std::move(t));
            }
            ops.push_back(std::move(op));
            outs = {{MAP_F + l - 1, 0}, {MAP_G + l - 1, lens[l] - 1}};
        } else {
            for (int v = 0; v < lens[l]; ++v) outs.push_back({(l == 7 ? MAP_T7 : MAP_T8) + v, v});
        }
        for (auto& ov : outs) {
            HostOp op = new_op(ov.first);
            op.compact = compact;
            for (int j = 0; j < 3; ++j) {
                int mp, sh;
                if (!src(l - 1, 2 * ov.second - 1 + j, mp, sh)) continue;
                if (sh < 0) { err = "negative shift in plan"; return false; }
                HostTerm t; t.src = mp; t.shift = sh; t.w = tap(j);
                if (compact && mp >= MAP_Y && mp < MAP_F) {  // a dense Y map read at site rows: through a scatter buffer
                    t.src = scatter_buffer(mp, sh);
                    t.shift = 0;
                    if (t.src < 0) { err = "too many scatter buffers in plan"; return false; }
                }
                op.terms.push_back(std::move(t));
            }
            ops.push_back(std::move(op));
        }
    }
    // ---- head: flatten index = c * 2 + t; fc1 + ReLU on the tensor cores, fc2 in the epilogue ---------------------
    HostOp h;
    h.head = true; h.cin = 64; h.cout = 256; h.bias = m.fc1_b; h.w2 = m.fc2_w; h.b2 = m.fc2_b;
    h.compact = compact;
    for (int t = 0; t < 2; ++t) {
        HostTerm tm; tm.src = MAP_T8 + t; tm.shift = 0;
        tm.w.resize((size_t)64 * 256);
        for (int o = 0; o < 256; ++o)
            for (int c = 0; c < 64; ++c) tm.w[(size_t)c * 256 + o] = m.fc1_w[(size_t)o * 128 + c * 2 + t];
        h.terms.push_back(std::move(tm));
    }
    ops.push_back(std::move(h));
    return true;
}

// ---- device op: DenseOp template + the map ids its pointers are patched from ---------------------------------------
struct DevOp {
    DenseOp p{};
    int seg_map[kMaxSegs] = {0, 0, 0};
    int out_map = -1;
    bool compact = false;
    int sc_map[kMaxScatter] = {0, 0, 0, 0, 0};
    bool two_cta = false;  // launched as dense_gemm2_kernel (CTA pairs, each holding half of every weight tile)
    bool conv1 = false;    // conv1 form (reads the X map)
    size_t smem = 0;
    size_t w_off = 0, bias_off = 0, w2_off = 0, b2_off = 0;  // offsets into the model blob
    double macs_per_row = 0;  // executed MACs per output row, one precision pass
};

// Lowers output channels [n0, n0 + n) of a plan op to kernel parameters + packed weights.  Returns false with
// err = "fit" when the slice's resident weights leave no room for a 2-deep ring (the caller then splits it).
bool lower_op(const HostOp& h, int n0, int n, DevOp& d, std::vector<uint8_t>& blob, std::string& err, bool two_cta = false,
              bool conv1_pair = false)
{
    DenseOp& p = d.p;
    p = DenseOp{};
    const int nfull = h.cout;
    if (n % 16 || n > 256 || h.terms.empty() || (int)h.terms.size() > kMaxTerms) { err = "op shape not supported"; return false; }
    p.n = n;
    p.out_groups = (uint32_t)nfull / 8;
    p.out_g0 = (uint32_t)n0 / 8;
    p.tmem_cols = 32;
    while (p.tmem_cols < (uint32_t)(2 * n)) p.tmem_cols <<= 1;
    p.n_terms = (int)h.terms.size();
    p.mode = h.head ? 1 : 0;
    d.out_map = h.out;
    uint32_t stage = 0;
    d.compact = h.compact;
    d.two_cta = two_cta;
    d.conv1 = h.conv1_taps > 0;
    p.n_scatter = (int)h.scatter.size();
    for (int k = 0; k < p.n_scatter; ++k) { p.sc_shift[k] = h.scatter[k].first; d.sc_map[k] = h.scatter[k].second; }
    if (p.n_scatter && (n0 != 0 || n != h.cout)) { err = "a scattering op cannot be split over output channels"; return false; }
    for (const HostTerm& t : h.terms) if (two_cta && t.gather) { err = "op not eligible for the CTA-pair form"; return false; }
    if (two_cta && (h.head || (h.conv1_taps > 0 && !conv1_pair) || n0 != 0 || n != h.cout || n % 32)) { err = "op not eligible for the CTA-pair form"; return false; }
    p.gather_taps = 0;
    p.gather_rows = nullptr;
    if (h.conv1_taps > 0) {
        p.n_segs = 1; p.n_stages = 1; p.ksteps = (h.conv1_taps + 1) / 2;
        DenseSeg& sg = p.seg[0];
        sg.row_off = h.terms[0].shift; sg.groups = 1; sg.smem_off = 0;
        d.seg_map[0] = h.terms[0].src;
        if (h.terms[0].gather) {
            // explicit im2col in shared memory: one 128-row plane per tap, gathered through the site-row index
            const uint32_t T = 2u * p.ksteps, pl = kTileRows * 16u;
            sg.gather = 1; sg.nrows = kTileRows;
            p.gather_taps = (int)T; p.planes_per_seg = (int)(2 * T); p.a_q_off = 2 * pl;
            p.term[0] = DenseTerm{0u, T * pl, pl};
            stage = 2 * T * pl;
        } else {
            // rows are consecutive positions: K-chunk c of row r is row r + c of the same plane (LBO = 16 bytes)
            sg.gather = 0; sg.nrows = kTileRows + 16;
            p.planes_per_seg = 2; p.a_q_off = 32;
            p.term[0] = DenseTerm{0u, sg.nrows * 16u, 16u};
            stage = 2 * sg.nrows * 16u;
        }
        d.macs_per_row = (double)p.ksteps * 16 * n;
    } else {
        if (h.cin % 16) { err = "cin must be a multiple of 16"; return false; }
        p.n_stages = h.cin / 16; p.ksteps = 1; p.a_q_off = 0; p.planes_per_seg = 4; p.n_segs = 0;
        int term_seg[kMaxTerms] = {0, 0, 0};
        for (size_t k = 0; k < h.terms.size(); ++k) {
            int s = -1;
            if (!h.terms[k].gather)
                for (int i = 0; i < p.n_segs; ++i)
                    if (!p.seg[i].gather && d.seg_map[i] == h.terms[k].src) s = i;
            if (s < 0) {
                if (p.n_segs == kMaxSegs) { err = "too many segments"; return false; }
                s = p.n_segs++;
                d.seg_map[s] = h.terms[k].src;
                p.seg[s].gather = h.terms[k].gather ? 1u : 0u;
                p.seg[s].row_off = h.terms[k].shift;
                p.seg[s].nrows = (uint32_t)h.terms[k].shift;  // holds the max shift until finalised
                p.seg[s].groups = (uint32_t)h.cin / 8;
            }
            p.seg[s].row_off = std::min(p.seg[s].row_off, h.terms[k].shift);
            p.seg[s].nrows = std::max<uint32_t>(p.seg[s].nrows, (uint32_t)h.terms[k].shift);
            term_seg[k] = s;
        }
        for (int s = 0; s < p.n_segs; ++s) {
            p.seg[s].nrows = kTileRows + (p.seg[s].nrows - (uint32_t)p.seg[s].row_off);
            p.seg[s].smem_off = stage;
            stage += 4 * p.seg[s].nrows * 16u;
        }
        for (size_t k = 0; k < h.terms.size(); ++k) {
            const int s = term_seg[k];
            const uint32_t pl = p.seg[s].nrows * 16u;
            p.term[k] = DenseTerm{p.seg[s].smem_off + (uint32_t)(h.terms[k].shift - p.seg[s].row_off) * 16u, 2 * pl, pl};
        }
        d.macs_per_row = (double)h.cin * n * h.terms.size();
    }
    p.stage_bytes = (stage + 127u) & ~127u;
    // ---- weight image: tiles [stage][kstep][term][hl], each [2][nw][8] bf16; the CTA-pair form stores one image per rank
    // holding that rank's nw = n / 2 output channels ------------------------------------------------------------------
    const int n_img = two_cta ? 2 : 1, nw = n / n_img;
    const uint32_t tile_elems = 2u * nw * 8u;
    const size_t n_tiles = (size_t)p.n_stages * p.ksteps * p.n_terms * 2;
    std::vector<uint16_t> img(n_img * n_tiles * tile_elems, 0);
    for (int r = 0; r < n_img; ++r)
    for (int st = 0; st < p.n_stages; ++st)
        for (int q = 0; q < p.ksteps; ++q)
            for (int k = 0; k < p.n_terms; ++k) {
                uint16_t* hi = &img[(r * n_tiles + ((size_t)(st * p.ksteps + q) * p.n_terms + k) * 2) * tile_elems];
                uint16_t* lo = hi + tile_elems;
                for (int c = 0; c < 2; ++c)
                    for (int o = 0; o < nw; ++o)
                        for (int e = 0; e < 8; ++e) {
                            float w = 0.f;
                            const int col = n0 + r * nw + o;
                            if (h.conv1_taps > 0) {
                                int tap = 2 * q + c;
                                if (tap < h.conv1_taps) w = h.terms[0].w[((size_t)tap * 8 + e) * nfull + col];
                            } else {
                                w = h.terms[k].w[(size_t)(16 * st + 8 * c + e) * nfull + col];
                            }
                            uint16_t wh = f2bf(w);
                            hi[((size_t)c * nw + o) * 8 + e] = wh;
                            lo[((size_t)c * nw + o) * 8 + e] = f2bf(w - bf2f(wh));
                        }
            }
    p.w_bytes = (uint32_t)(img.size() * 2 / n_img);  // per CTA
    const size_t w_al = (p.w_bytes + 127u) & ~127u;
    if (w_al + 2 * (size_t)p.stage_bytes + kSmemAux > kSmemMax) { err = "fit"; return false; }
    p.ring = (int)std::min<size_t>(kProducerWarps, (kSmemMax - kSmemAux - w_al) / p.stage_bytes);  // producer warp s owns slot s
    d.smem = dense_smem_bytes(p);
    auto append = [&](const void* src, size_t bytes) {
        size_t o = (blob.size() + 255) & ~(size_t)255;
        blob.resize(o + bytes);
        memcpy(blob.data() + o, src, bytes);
        return o;
    };
    d.w_off = append(img.data(), img.size() * 2);
    d.bias_off = append(h.bias.data() + n0, (size_t)n * 4);
    if (h.head) {
        d.w2_off = append(h.w2.data(), h.w2.size() * 4);
        d.b2_off = append(h.b2.data(), h.b2.size() * 4);
    }
    return true;
}

void bind_blob(DevOp& d, const uint8_t* blob)
{
    d.p.w_img = blob + d.w_off;
    d.p.bias = reinterpret_cast<const float*>(blob + d.bias_off);
    if (d.p.mode == 1) {
        d.p.w2 = reinterpret_cast<const float*>(blob + d.w2_off);
        d.p.b2 = reinterpret_cast<const float*>(blob + d.b2_off);
    }
}

bool pdl_enabled()
{
    static const bool v = getenv("HM_NO_PDL") == nullptr;
    return v;
}

bool pair_compact()
{
    static const bool v = getenv("HM_NO_PAIR_COMPACT") == nullptr;
    return v;
}

// conv1 + conv2 in one kernel (dense_fused12.cuh); HM_NO_FUSE12=1 runs them as two launches (A/B measurements)
bool fuse12_enabled()
{
    static const bool v = getenv("HM_NO_FUSE12") == nullptr;
    return v;
}

// the compact chain (F2.. head) of a site tile in one kernel (site_chain.cuh), running maps in tensor memory; HM_NO_CHAIN=1 launches
// the compact ops one by one instead (A/B measurements: 97.7 / 99.2 ms per step op by op against 93.0 / 94.4 ms on the same box).
bool chain_enabled()
{
    static const bool v = getenv("HM_NO_CHAIN") == nullptr;
    return v;
}

bool two_cta_enabled()
{
    static const bool v = getenv("HM_NO_2CTA") == nullptr;
    return v;
}

float g_debug_op_ms = 0.f;
// The dynamic shared memory limit of a kernel is a per-DEVICE attribute: one process may drive several GPUs (the `call` driver's
// --devices), so it is set once per device, under a lock (engines are created from one thread per device).
int ensure_kernel_attr()
{
    static std::mutex m;
    static bool done[64] = {};
    int dev = 0;
    TCUDA("dense kernel attribute", cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(m);
    if (dev >= 0 && dev < 64 && done[dev]) return 0;
    TCUDA("dense kernel attribute", cudaFuncSetAttribute(dense_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    TCUDA("dense kernel attribute", cudaFuncSetAttribute(dense_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    TCUDA("dense kernel attribute", cudaFuncSetAttribute(dense_fused12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    TCUDA("dense kernel attribute", cudaFuncSetAttribute(site_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return 0;
}

// ---- small kernels around the plan ------------------------------------------------------------------------------------

// X map of one sub-batch: one block per 128-row tile.  Row of track position i holds the 8 features of
// s_extract_kmer_features (src/app/hifimeth/eval_kmer_features.cpp:42-64) for that strand position, split hi/lo.
__global__ void __launch_bounds__(128)
track_features_kernel(const uint8_t* __restrict__ bcode, const ushort4* __restrict__ kinf, const uint32_t* __restrict__ base_off,
                      const uint32_t* __restrict__ tile_read, const uint32_t* __restrict__ tile_first, uint32_t gtile0,
                      uint8_t* __restrict__ x_hi, uint8_t* __restrict__ x_lo)
{
    const uint32_t gt = gtile0 + blockIdx.x;
    const uint32_t rs = tile_read[gt];
    const uint32_t r = rs & 0x7fffffffu;
    const bool rev = (rs >> 31) != 0;
    const uint32_t B = base_off[r];
    const int L = (int)(base_off[r + 1] - B);
    const int i = (int)((gt - tile_first[gt]) * kTileRows + threadIdx.x) - kHaloL;
    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
    if (i >= 0 && i < L) {
        const int p = rev ? L - 1 - i : i;
        uint32_t c = bcode[B + p];
        if (rev && c < 4) c = 3u - c;
        const ushort4 k = kinf[B + p];
        const float f0 = __fdiv_rn((float)k.x, 952.0f), f1 = __fdiv_rn((float)k.y, 952.0f);
        const float f2 = __fdiv_rn((float)k.z, 952.0f), f3 = __fdiv_rn((float)k.w, 952.0f);
        const float own_i = rev ? f2 : f0, own_p = rev ? f3 : f1, opp_i = rev ? f0 : f2, opp_p = rev ? f1 : f3;
        const uint32_t one = 0x3f80u;  // bf16 1.0
        hi.x = (c == 0 ? one : 0u) | ((c == 1 ? one : 0u) << 16);
        hi.y = (c == 2 ? one : 0u) | ((c == 3 ? one : 0u) << 16);
        __nv_bfloat16 h0 = __float2bfloat16_rn(own_i), h1 = __float2bfloat16_rn(own_p);
        __nv_bfloat16 h2 = __float2bfloat16_rn(opp_i), h3 = __float2bfloat16_rn(opp_p);
        hi.z = pack_bf16x2(h0, h1);
        hi.w = pack_bf16x2(h2, h3);
        lo.z = pack_bf16x2(__float2bfloat16_rn(own_i - __bfloat162float(h0)), __float2bfloat16_rn(own_p - __bfloat162float(h1)));
        lo.w = pack_bf16x2(__float2bfloat16_rn(opp_i - __bfloat162float(h2)), __float2bfloat16_rn(opp_p - __bfloat162float(h3)));
    }
    const size_t row = (size_t)blockIdx.x * kTileRows + threadIdx.x;
    reinterpret_cast<uint4*>(x_hi)[row] = hi;
    reinterpret_cast<uint4*>(x_lo)[row] = lo;
}

// Per-site lookup of one class region: row (o - 201) of the site's track -> logits, probability, ML byte
// (s_logits_to_methy_probs, src/app/hifimeth/mod_batch.cpp:46-64).
__global__ void __launch_bounds__(256)
site_lookup_kernel(const float2* __restrict__ logit_rows, const uint32_t* __restrict__ track_row_fwd,
                   const uint32_t* __restrict__ track_row_rev, const uint32_t* __restrict__ base_off,
                   const uint32_t* __restrict__ site_read, const uint32_t* __restrict__ site_pos,
                   const uint32_t* __restrict__ site_out, uint32_t first, uint32_t count, float* __restrict__ logits,
                   uint8_t* __restrict__ ml)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint32_t r = site_read[first + k];
    const uint32_t sp = site_pos[first + k];
    const bool rev = (sp >> 31) != 0;
    const int p = (int)(sp & 0x7fffffffu);
    const int L = (int)(base_off[r + 1] - base_off[r]);
    const int o = rev ? L - 1 - p : p;
    const size_t row = (size_t)(rev ? track_row_rev[r] : track_row_fwd[r]) + (size_t)(kHaloL + o - 201);
    const float2 lg = logit_rows[row];
    const uint32_t out = site_out[first + k];
    logits[2 * (size_t)out] = lg.x;
    logits[2 * (size_t)out + 1] = lg.y;
    ml[out] = prob_to_ml(softmax_p1(lg.x, lg.y));
}

// Compact runs: compact row m of a run is site (first_a + m) for m < n_a, else site (first_b + m - n_a) of the class-ordered
// site list.  site_rows_kernel writes the dense row (local to the sub-batch) that holds s = o - 201 of every compact row.
__device__ __forceinline__ uint32_t run_site(uint32_t m, uint32_t first_a, uint32_t n_a, uint32_t first_b)
{
    return m < n_a ? first_a + m : first_b + (m - n_a);
}

// Compact runs may span several sub-batches: this sub-batch's sites are compact rows m_off .. m_off + n.  rows[] holds the row of
// the batch-wide X map (gathered conv1-form ops), site_of_row[] is indexed by the row local to the sub-batch's dense maps.
__global__ void __launch_bounds__(256)
site_rows_kernel(const uint32_t* __restrict__ track_row_fwd, const uint32_t* __restrict__ track_row_rev,
                 const uint32_t* __restrict__ base_off, const uint32_t* __restrict__ site_read, const uint32_t* __restrict__ site_pos,
                 uint32_t first_a, uint32_t n_a, uint32_t first_b, uint32_t n, uint32_t n_pad, uint32_t row_base, uint32_t m_off,
                 const uint32_t* __restrict__ site_out, uint32_t* __restrict__ rows, int32_t* __restrict__ site_of_row,
                 uint32_t* __restrict__ out_idx)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_pad) return;
    uint32_t row = 0;  // padding rows read row 0 (any valid row) and are never looked at
    out_idx[m_off + m] = m < n ? site_out[run_site(m, first_a, n_a, first_b)] : 0xffffffffu;  // where the chain kernel's head writes the call
    if (m < n) {
        const uint32_t k = run_site(m, first_a, n_a, first_b);
        const uint32_t r = site_read[k];
        const uint32_t sp = site_pos[k];
        const bool rev = (sp >> 31) != 0;
        const int p = (int)(sp & 0x7fffffffu);
        const int L = (int)(base_off[r + 1] - base_off[r]);
        const int o = rev ? L - 1 - p : p;
        row = (rev ? track_row_rev[r] : track_row_fwd[r]) + (uint32_t)(kHaloL + o - 201);
        site_of_row[row - row_base] = (int32_t)(m_off + m);  // inverse map for the dense layers' scatter (cleared to -1 before this launch)
    }
    rows[m_off + m] = row;
}

// logits of the compact rows -> per-site logits + ML byte in hm_call_batch order
// (s_logits_to_methy_probs, src/app/hifimeth/mod_batch.cpp:46-64).
__global__ void __launch_bounds__(256)
site_finish_kernel(const float2* __restrict__ clogit, const uint32_t* __restrict__ site_out, uint32_t first_a, uint32_t n_a,
                   uint32_t first_b, uint32_t n, float* __restrict__ logits, uint8_t* __restrict__ ml)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const float2 lg = clogit[m];
    const uint32_t out = site_out[run_site(m, first_a, n_a, first_b)];
    logits[2 * (size_t)out] = lg.x;
    logits[2 * (size_t)out + 1] = lg.y;
    ml[out] = prob_to_ml(softmax_p1(lg.x, lg.y));
}

}  // namespace

// ---- model ---------------------------------------------------------------------------------------------------------------
struct TensorModel {
    std::vector<DevOp> ops;
    // conv1 + conv2 fused (dense_fused12_kernel): launched in place of op i_y1, op i_y2 is skipped
    bool fused12 = false;
    DevOp f12_c1;
    int i_y1 = -1, i_y2 = -1;
    uint8_t* d_blob = nullptr;
    double macs_per_row = 0;  // executed, one precision pass, all ops
    int lens[9] = {};         // per-site output length of every layer (lens[0] = 401)
    // the compact chain as one kernel (site_chain_kernel): every compact op except the conv1-form ones, in chain order
    struct ChainItem {
        DevOp d;                      // pair lowering: weights / bias offsets in the blob, n, macs
        int term_map[kMaxTerms] = {0, 0, 0};  // source map of each term
        bool resident[kMaxTerms] = {false, false, false};  // the term reads a packed map in TMEM (written by an earlier chain op)
        int out_map = -1;
        ChainOp c{};                  // template: pointers are patched per launch
    };
    bool chain = false;
    std::vector<ChainItem> chain_ops;
    size_t chain_w2_off = 0, chain_b2_off = 0;
};

const char* tensor_last_error() { return g_err.c_str(); }

int tensor_model_build(TensorModelHandle& m, const CnnModel& host)
{
    std::vector<HostOp> plan;
    std::string err;
    if (!build_plan(host, compact_mode(), plan, err)) return tfail("dense plan: " + err);
    TensorModel* t = new TensorModel();
    t->lens[0] = host.kmer;
    for (int l = 0; l < 8; ++l) t->lens[l + 1] = (t->lens[l] + 2 - host.convs[l].k) / 2 + 1;
    std::vector<uint8_t> blob;
    for (const HostOp& h : plan) {
        // an op whose resident weights leave no room for the activation ring is split over output channels
        // dense layers with > 64 KiB of weights (conv2 .. conv6) run as CTA pairs: half the weights per SM, deep ring
        if (two_cta_enabled() && (!h.compact || pair_compact()) && !h.head && h.conv1_taps == 0 && h.cout % 32 == 0 &&
            (size_t)h.cin * h.cout * h.terms.size() * 4 > ((h.compact ? 32u : 64u) << 10)) {
            DevOp d;
            std::vector<uint8_t> trial = blob;
            if (lower_op(h, 0, h.cout, d, trial, err, true)) {
                blob.swap(trial);
                t->macs_per_row += d.macs_per_row;
                t->ops.push_back(d);
                continue;
            }
        }
        int parts = 1;
        for (; parts <= 4; parts *= 2) {
            if (h.cout % (16 * parts) || (h.head && parts > 1)) { parts = 8; break; }
            std::vector<DevOp> ds(parts);
            std::vector<uint8_t> trial = blob;
            bool ok = true;
            for (int i = 0; i < parts && ok; ++i) ok = lower_op(h, i * (h.cout / parts), h.cout / parts, ds[i], trial, err);
            if (ok) {
                blob.swap(trial);
                for (DevOp& d : ds) { t->macs_per_row += d.macs_per_row; t->ops.push_back(d); }
                break;
            }
            if (err != "fit") { parts = 8; break; }
        }
        if (parts > 4) { delete t; return tfail("dense plan lowering: " + (err == "fit" ? std::string("op does not fit shared memory") : err)); }
    }
    if (compact_mode() && fuse12_enabled() && two_cta_enabled()) {
        for (size_t i = 0; i < t->ops.size(); ++i) {
            const DevOp& d = t->ops[i];
            if (!d.compact && d.out_map == MAP_Y && d.p.n == 128 && d.p.out_g0 == 0) t->i_y1 = (int)i;
            if (!d.compact && d.out_map == MAP_Y + 1 && d.two_cta && d.p.n == 128 && d.p.n_stages == 8 && d.p.n_terms == 3) t->i_y2 = (int)i;
        }
        const HostOp* h1 = nullptr;
        for (const HostOp& h : plan)
            if (!h.compact && h.out == MAP_Y && h.conv1_taps > 0) h1 = &h;
        if (t->i_y1 >= 0 && t->i_y2 >= 0 && h1 && lower_op(*h1, 0, 128, t->f12_c1, blob, err, true, true)) {
            Fused12Op probe{};
            probe.c1 = t->f12_c1.p;
            probe.c2 = t->ops[t->i_y2].p;
            t->fused12 = fused12_smem_bytes(probe) <= kSmemMax;
        }
    }
    if (compact_mode() && chain_enabled() && two_cta_enabled()) {
        // ---- chain order: F2 G2 .. F6 G6 | T7_1 T7_2 T7_0 T7_3 | T8_0 T8_1 | head ------------------------------------------------
        // F and G alternate so that the MMAs of one run under the epilogue of the other; T7_1 / T7_2 read scatter copies only and
        // go first.  TMEM columns (site_chain.cuh): packed F map [0,128) | packed G map [128,256) | F accumulator [256,384) |
        // G accumulator [384,512).  Tail: T7 accumulators T7_1 256, T7_2 320 (over the drained F accumulator), T7_0 384, T7_3 448
        // (over the drained G accumulator); packed T7 outputs over what F6 / G6 (96 channels) leave free or dead; T8 accumulators
        // 256 / 320, packed T8 outputs [384,512); head accumulator [0,256).
        std::vector<const HostOp*> order;
        auto add = [&](int out, bool head) {
            for (const HostOp& h : plan)
                if (h.compact && ((head && h.head) || (!head && !h.head && h.out == out))) { order.push_back(&h); return true; }
            return false;
        };
        bool ok = true;
        // F1 / G1 (conv1 form, gathered from the X map) open the chain: their outputs never leave the chip either
        for (int l = 1; l <= 6 && ok; ++l) ok = add(MAP_F + l - 1, false) && add(MAP_G + l - 1, false);
        for (int v : {1, 0, 3, 2}) ok = ok && add(MAP_T7 + v, false);
        ok = ok && add(MAP_T8 + 0, false) && add(MAP_T8 + 1, false) && add(-1, true);
        size_t n_compact = 0;
        for (const HostOp& h : plan) n_compact += h.compact ? 1 : 0;
        ok = ok && order.size() == n_compact && order.size() <= (size_t)kChainMaxOps;
        std::vector<TensorModel::ChainItem> items;
        auto index_of = [&](int out_map) {
            for (size_t j = 0; j < items.size(); ++j)
                if (items[j].out_map == out_map && !items[j].c.head) return (int)j;
            return -1;
        };
        size_t steps = 0;
        for (size_t i = 0; i < order.size() && ok; ++i) {
            const HostOp& h = *order[i];
            TensorModel::ChainItem it;
            HostOp hw = h;       // the pair lowering packs the weights; the head is lowered as a plain op (fc2 lives in the epilogue)
            hw.head = false;
            hw.scatter.clear();
            std::vector<uint8_t> tmp;
            const bool c1 = h.conv1_taps > 0;
            if (c1) hw.terms[0].gather = false;  // the pair lowering only has to pack the weights
            if ((!c1 && h.cin % 32) || h.cout % 32 || !lower_op(hw, 0, h.cout, it.d, tmp, err, true, c1)) { ok = false; break; }
            // ---- chain image: per rank [stage pair][term][stage][hl] tiles, so that a ring step is one contiguous block -------------
            // conv1 form: a "stage" is a K-step of two taps; their number is padded to an even one with zero tiles
            const uint32_t n_terms = (uint32_t)h.terms.size(), tile = (uint32_t)h.cout * 16u, src_rank_bytes = it.d.p.w_bytes;
            const uint32_t n_st = c1 ? (uint32_t)it.d.p.ksteps : (uint32_t)h.cin / 16, n_st_pad = (n_st + 1u) & ~1u;
            const uint32_t rank_bytes = n_st_pad * n_terms * 2u * tile;
            if (src_rank_bytes != n_st * n_terms * 2u * tile || (c1 && n_terms != 1)) { ok = false; break; }
            {
                size_t o = (blob.size() + 255) & ~(size_t)255;
                blob.resize(o + 2 * (size_t)rank_bytes, 0);
                const size_t src_off = it.d.w_off;  // where lower_op put the [stage][term][hl] image of rank 0
                it.d.w_off = o;
                uint8_t* dst = blob.data() + o;
                for (uint32_t r = 0; r < 2; ++r)
                    for (uint32_t S = 0; S < n_st_pad / 2; ++S)
                        for (uint32_t k = 0; k < n_terms; ++k)
                            for (uint32_t st = 0; st < 2; ++st)
                                for (uint32_t hl = 0; hl < 2; ++hl) {
                                    if (2 * S + st < n_st)
                                        memcpy(dst, tmp.data() + src_off + (size_t)r * src_rank_bytes + ((size_t)((2 * S + st) * n_terms + k) * 2 + hl) * tile, tile);
                                    dst += tile;
                                }
                o = (blob.size() + 255) & ~(size_t)255;
                blob.resize(o + (size_t)h.cout * 4);
                memcpy(blob.data() + o, h.bias.data(), (size_t)h.cout * 4);
                it.d.bias_off = o;
            }
            const int cin_eff = (int)(16u * n_st_pad);  // conv1 form: 8 features x padded taps
            steps += (size_t)(cin_eff / 32) * n_terms;
            ChainOp& c = it.c;
            c.n = h.cout; c.cin = cin_eff; c.n_terms = (int)n_terms; c.head = h.head ? 1 : 0;
            c.gather = c1 ? 1 : 0;
            c.gather_shift = c1 ? h.terms[0].shift : 0;
            c.w_rank_bytes = rank_bytes;
            c.wait_op = (int)i;
            c.mma_wait = 0;
            it.out_map = h.out;
            // a ring slot holds the step's weight tiles (cout x 64 bytes for this CTA) in front of the 16 KiB slab of a streamed term
            if ((uint32_t)h.cout * 64u > (h.head ? kChainSlotBytes : kChainWBytes)) { ok = false; break; }
            int extra_wait[2] = {-1, -1};
            const uint32_t nh2 = (uint32_t)h.cout / 2;  // columns of the packed hi half (and of the lo half)
            if (h.head) { c.acc_col = 0; if (h.cout != 256) { ok = false; break; } }
            else if (h.out >= MAP_F && h.out < MAP_G) { c.acc_col = 256; c.out_hi_col = 0; c.out_lo_col = nh2; }
            else if (h.out >= MAP_G && h.out < MAP_T7) { c.acc_col = 384; c.out_hi_col = 128; c.out_lo_col = 128 + nh2; }
            else if (h.out >= MAP_T7 && h.out < MAP_T8) {
                if (h.cout != 64 || h.cin != 96) { ok = false; break; }
                const int v = h.out - MAP_T7, f6 = index_of(MAP_F + 5), g6 = index_of(MAP_G + 5);
                // T7_0 is the only reader of F6 and T7_3 of G6: their packed outputs go over their own inputs; T7_1 goes to the 2 x 32
                // columns the 96-channel maps leave free, T7_2 (issued last) to the 2 x 32 columns left in the dead F6 / G6
                static const uint32_t acc[4] = {384, 256, 320, 448}, ohi[4] = {0, 96, 64, 128}, olo[4] = {32, 224, 192, 160};
                c.acc_col = acc[v]; c.out_hi_col = ohi[v]; c.out_lo_col = olo[v];
                extra_wait[0] = (v == 1 || v == 2) ? f6 : g6;  // the accumulator columns were F6's / G6's accumulator
                if (f6 < 0 || g6 < 0) { ok = false; break; }
            } else if (h.out >= MAP_T8 && h.out < MAP_S) {
                if (h.cout != 64 || h.cin != 64) { ok = false; break; }
                const int w = h.out - MAP_T8;
                c.acc_col = w ? 320 : 256; c.out_hi_col = w ? 448 : 384; c.out_lo_col = w ? 480 : 416;
            } else { ok = false; break; }
            int nw = 0;
            auto add_wait = [&](int j) {
                if (j < 0) return true;
                for (int q = 0; q < nw; ++q) if (((c.mma_wait >> (8 * q)) & 0xffu) == (uint32_t)j + 1u) return true;
                if (nw == kChainMaxWait) return false;
                c.mma_wait |= ((uint32_t)j + 1u) << (8 * nw++);
                return true;
            };
            for (size_t k = 0; k < h.terms.size() && ok; ++k) {
                it.term_map[k] = h.terms[k].src;
                if (c1) { ok = h.terms[k].src == MAP_X && h.terms[k].gather; continue; }  // gathered from the batch-wide X map
                if (h.terms[k].shift != 0 || h.terms[k].gather) { ok = false; break; }
                const int sm = h.terms[k].src;
                const bool streamed_kind = sm >= MAP_S;  // scatter copies of the dense maps
                const int dep = index_of(sm);
                it.resident[k] = dep >= 0;
                if ((dep < 0) != streamed_kind || (h.head && streamed_kind)) { ok = false; break; }
                if (dep >= 0) {
                    if (items[dep].c.n != h.cin) { ok = false; break; }
                    c.term[k].a_hi_col = items[dep].c.out_hi_col;
                    c.term[k].a_lo_col = items[dep].c.out_lo_col;
                    ok = add_wait(dep);
                }
            }
            ok = ok && add_wait(extra_wait[0]) && add_wait(extra_wait[1]);
            if (!ok) break;
            items.push_back(it);
        }
        if (ok) {
            // every T7 output goes to columns that only EARLIER MMAs read (order T7_1 T7_0 T7_3 T7_2): no epilogue waits for a later op
            ok = steps < (size_t)kChainMaxSteps && steps >= (size_t)kChainSlots;
        }
        if (ok) {
            const HostOp& hh = *order.back();
            size_t o = (blob.size() + 255) & ~(size_t)255;
            blob.resize(o + hh.w2.size() * 4);
            memcpy(blob.data() + o, hh.w2.data(), hh.w2.size() * 4);
            t->chain_w2_off = o;
            o = (blob.size() + 255) & ~(size_t)255;
            blob.resize(o + hh.b2.size() * 4);
            memcpy(blob.data() + o, hh.b2.data(), hh.b2.size() * 4);
            t->chain_b2_off = o;
            t->chain_ops = std::move(items);
            t->chain = true;
        }
    }
    cudaError_t st = cudaMalloc((void**)&t->d_blob, blob.size());
    if (st == cudaSuccess) st = cudaMemcpy(t->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (st != cudaSuccess) { cudaFree(t->d_blob); delete t; return tfail(std::string("weight upload: ") + cudaGetErrorString(st)); }
    for (DevOp& d : t->ops) bind_blob(d, t->d_blob);
    if (t->fused12) bind_blob(t->f12_c1, t->d_blob);
    for (auto& it : t->chain_ops) {
        it.c.w_img = t->d_blob + it.d.w_off;
        it.c.bias = reinterpret_cast<const float*>(t->d_blob + it.d.bias_off);
    }
    m.p = t;
    return ensure_kernel_attr();
}

void tensor_model_free(TensorModelHandle& m)
{
    if (!m.p) return;
    cudaFree(m.p->d_blob);
    delete m.p;
    m.p = nullptr;
}

// ---- workspace ---------------------------------------------------------------------------------------------------------------
struct TensorWorkspaceImpl {
    uint32_t rows_cap = 0;            // dense rows of one sub-batch
    uint32_t compact_cap = 0;         // sites of one context in one sub-batch
    unsigned long long plane_stride = 0, cplane_stride = 0;  // dense maps / compact maps
    int32_t* d_site_of_row = nullptr; // [rows_cap + slack] compact row of the site whose s-row this is, or -1
    uint8_t* d_maps = nullptr;
    uint8_t* map[N_MAPS] = {};
    size_t logit_rows_cap = 0;
    float* d_logit[3] = {};
    uint32_t tiles_cap = 0, reads_cap = 0;
    uint32_t *h_tile_read = nullptr, *h_tile_first = nullptr, *d_tile_read = nullptr, *d_tile_first = nullptr;
    uint32_t *h_track_row = nullptr, *d_track_row = nullptr;  // [2][reads_cap]
    uint8_t* d_xg = nullptr;          // compact mode: X of the WHOLE batch, planes {hi, lo} of [total rows + slack][8] bf16
    unsigned long long xg_stride = 0;
    const uint8_t* x_cur = nullptr;   // X rows of the sub-batch being launched (points into d_xg) and their plane stride
    unsigned long long x_stride_cur = 0;
    uint32_t* d_site_rows = nullptr;  // [compact_cap] compact row -> X row (global)
    uint32_t* d_out_idx = nullptr;    // [compact_cap + 2 tiles] compact row -> site index in hm_call_batch order (0xffffffff: padding)
    float* d_clogit = nullptr;        // [rows_cap][2] logits of the compact rows
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // shape of the last tensor_batch_run (debug hooks): sub-batches, compact groups per context, rows of the X map
    uint32_t last_subs = 0, last_groups[3] = {0, 0, 0}, last_x_rows = 0;
    double macs = 0;  // executed MACs of the running batch, one precision pass (launch_op / launch_fused12 add to it)
    bool spill_chain = false;  // debug reruns: the chain kernel also stores its maps to HBM (hm_debug_dump_acts reads them)
    bool no_chain = false;
};

int tensor_workspace_alloc(TensorWorkspace& w, uint32_t max_bases, uint32_t max_reads, uint32_t max_rows)
{
    TensorWorkspaceImpl* s = new TensorWorkspaceImpl();
    w.impl = s;
    const size_t total_rows = 2 * (size_t)max_bases + (size_t)1070 * max_reads + 256;  // every track: L + 408 rounded up to 128, x2
    size_t cap = max_rows;
    if (!cap) {
        const char* env = getenv("HM_DENSE_ROWS");
        cap = env ? (size_t)atoll(env) : ((size_t)1 << 21);  // 2 Mi rows: ~20 GB of maps; fewer, larger launches
    }
    cap = std::min(cap, total_rows);
    cap = ((cap + kTileRows - 1) / kTileRows) * kTileRows;
    s->rows_cap = (uint32_t)cap;
    s->compact_cap = compact_mode() ? (uint32_t)(((cap / 2 + kTileRows - 1) / kTileRows) * kTileRows) : (uint32_t)cap;
    s->plane_stride = (unsigned long long)(cap + kSlackRows) * 16ull;
    s->cplane_stride = (unsigned long long)(s->compact_cap + kSlackRows) * 16ull;
    auto is_dense_map = [](int i) { return i < MAP_F || !compact_mode(); };
    size_t bytes = 0;
    for (int i = 0; i < N_MAPS; ++i) {
        if (i >= MAP_S && !compact_mode()) continue;
        bytes += 2 * (size_t)map_channels(i) / 8 * (is_dense_map(i) ? s->plane_stride : s->cplane_stride);
    }
    TCUDA("activation workspace", cudaMalloc((void**)&s->d_maps, bytes));
    TCUDA("activation workspace", cudaMemset(s->d_maps, 0, bytes));
    size_t at = 0;
    for (int i = 0; i < N_MAPS; ++i) {
        if (i >= MAP_S && !compact_mode()) continue;
        s->map[i] = s->d_maps + at;
        at += 2 * (size_t)map_channels(i) / 8 * (is_dense_map(i) ? s->plane_stride : s->cplane_stride);
    }
    TCUDA("site rows", cudaMalloc((void**)&s->d_site_of_row, (cap + kSlackRows) * sizeof(int32_t)));
    if (compact_mode()) {
        s->xg_stride = (unsigned long long)(total_rows + kSlackRows) * 16ull;
        TCUDA("X map", cudaMalloc((void**)&s->d_xg, 2 * s->xg_stride));
        TCUDA("X map", cudaMemset(s->d_xg, 0, 2 * s->xg_stride));
    }
    s->logit_rows_cap = total_rows;
    if (!compact_mode())
        for (int c = 0; c < 3; ++c) TCUDA("logit rows", cudaMalloc((void**)&s->d_logit[c], total_rows * 2 * sizeof(float)));
    TCUDA("site rows", cudaMalloc((void**)&s->d_site_rows, (cap + kTileRows) * sizeof(uint32_t)));
    TCUDA("site rows", cudaMalloc((void**)&s->d_out_idx, (cap + 2 * kTileRows) * sizeof(uint32_t)));
    TCUDA("site rows", cudaMalloc((void**)&s->d_clogit, (cap + kTileRows) * 2 * sizeof(float)));
    s->tiles_cap = (uint32_t)(total_rows / kTileRows + 1);
    s->reads_cap = max_reads;
    TCUDA("track tables", cudaMallocHost((void**)&s->h_tile_read, (size_t)s->tiles_cap * 4));
    TCUDA("track tables", cudaMallocHost((void**)&s->h_tile_first, (size_t)s->tiles_cap * 4));
    TCUDA("track tables", cudaMalloc((void**)&s->d_tile_read, (size_t)s->tiles_cap * 4));
    TCUDA("track tables", cudaMalloc((void**)&s->d_tile_first, (size_t)s->tiles_cap * 4));
    TCUDA("track tables", cudaMallocHost((void**)&s->h_track_row, (size_t)2 * std::max(max_reads, 1u) * 4));
    TCUDA("track tables", cudaMalloc((void**)&s->d_track_row, (size_t)2 * std::max(max_reads, 1u) * 4));
    TCUDA("events", cudaEventCreate(&s->ev0));
    TCUDA("events", cudaEventCreate(&s->ev1));
    return 0;
}

void tensor_workspace_free(TensorWorkspace& w)
{
    TensorWorkspaceImpl* s = w.impl;
    if (!s) return;
    cudaFree(s->d_maps);
    for (float* p : s->d_logit) cudaFree(p);
    cudaFree(s->d_site_rows); cudaFree(s->d_out_idx); cudaFree(s->d_clogit); cudaFree(s->d_site_of_row); cudaFree(s->d_xg);
    cudaFreeHost(s->h_tile_read); cudaFreeHost(s->h_tile_first); cudaFree(s->d_tile_read); cudaFree(s->d_tile_first);
    cudaFreeHost(s->h_track_row); cudaFree(s->d_track_row);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    delete s;
    w.impl = nullptr;
}

namespace {

// Kernel parameters of an op with the workspace pointers of this sub-batch filled in.
DenseOp patch_op(const DevOp& d, const TensorWorkspaceImpl& s, uint32_t n_tiles, float* logit_out)
{
    DenseOp p = d.p;
    auto stride_of = [&](int map) { return (map < MAP_F || !compact_mode()) ? s.plane_stride : s.cplane_stride; };
    for (int i = 0; i < p.n_segs; ++i) {
        p.seg[i].src = s.map[d.seg_map[i]];
        p.seg[i].plane_stride = stride_of(d.seg_map[i]);
        if (d.seg_map[i] == MAP_X && s.x_cur) {
            // compact mode keeps X of the whole batch: dense ops read the running sub-batch's rows, gathered ops index it globally
            p.seg[i].src = p.seg[i].gather ? s.d_xg : s.x_cur;
            p.seg[i].plane_stride = s.x_stride_cur;
        }
    }
    p.n_tiles = n_tiles;
    p.gather_rows = s.d_site_rows;
    if (p.mode == 0) {
        p.out = s.map[d.out_map];
        p.out_plane_stride = stride_of(d.out_map);
    } else {
        p.logits = logit_out;
    }
    for (int k = 0; k < p.n_scatter; ++k) p.sc_out[k] = s.map[d.sc_map[k]];
    p.sc_plane_stride = s.cplane_stride;
    p.site_of_row = s.d_site_of_row;
    return p;
}

long long* g_f12_dbg = nullptr;
long long* g_chain_dbg = nullptr;
long long* g_g2_dbg = nullptr;
int g_g2_which = -1;  // HM_G2_STAMPS=<k>: stamp the k-th dense_gemm2 launch of every batch (0 = first)
int g_g2_count = 0;

// conv1 + conv2 of `rows` dense rows in one launch (dense_fused12_kernel): tiles of 124 output rows, CTA pairs.
int launch_fused12(const DevOp& d1, const DevOp& d2, TensorWorkspaceImpl& s, uint32_t rows, int sm_count, cudaStream_t stream)
{
    Fused12Op f{};
    // HM_F12_STAMPS=1: CTA 0 of every launch stamps its first tiles; the last launch's stamps are printed after the batch
    static const bool stamps = getenv("HM_F12_STAMPS") != nullptr;
    if (stamps) {
        if (!g_f12_dbg) cudaMalloc((void**)&g_f12_dbg, 48 * 16 * sizeof(long long));
        cudaMemsetAsync(g_f12_dbg, 0, 48 * 16 * sizeof(long long), stream);
        f.dbg = g_f12_dbg;
    }
    f.n_tiles = (rows + kF12OutRows - 1) / kF12OutRows;
    s.macs += (double)((f.n_tiles + 1) / 2 * 2) * kTileRows * (d1.macs_per_row + d2.macs_per_row);  // 128 rows computed per 124 kept
    f.c1 = patch_op(d1, s, f.n_tiles, nullptr);
    f.c2 = patch_op(d2, s, f.n_tiles, nullptr);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(std::min<uint32_t>(((f.n_tiles + 1) / 2) * 2, (uint32_t)sm_count & ~1u));
    cfg.blockDim = dim3(kF12Threads);
    cfg.dynamicSmemBytes = fused12_smem_bytes(f);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, dense_fused12_kernel, f) == cudaSuccess ? 0 : -1;
}

int launch_op(const DevOp& d, TensorWorkspaceImpl& s, uint32_t n_tiles, float* logit_out, int sm_count, cudaStream_t stream)
{
    DenseOp p = patch_op(d, s, n_tiles, logit_out);
    s.macs += (double)(d.two_cta ? (n_tiles + 1) / 2 * 2 : n_tiles) * kTileRows * d.macs_per_row;
    uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)sm_count);
    if (d.two_cta) grid = std::min<uint32_t>(((n_tiles + 1) / 2) * 2, (uint32_t)sm_count & ~1u);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kDenseThreads);
    cfg.dynamicSmemBytes = d.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    if (d.two_cta) {
        // HM_G2_STAMPS=<k>: pair 0 of the k-th CTA-pair launch of the batch stamps tiles 16 .. 47 (needs the stamps build)
        static const char* g2 = getenv("HM_G2_STAMPS");
        if (g2) {
            if (!g_g2_dbg) cudaMalloc((void**)&g_g2_dbg, 512 * sizeof(long long));
            if (g_g2_count++ == atoi(g2)) {
                cudaMemsetAsync(g_g2_dbg, 0, 512 * sizeof(long long), stream);
                p.dbg = g_g2_dbg;
                g_g2_which = p.n;
            }
        }
        return cudaLaunchKernelEx(&cfg, dense_gemm2_kernel, p) == cudaSuccess ? 0 : -1;
    }
    return cudaLaunchKernelEx(&cfg, dense_gemm_kernel, p) == cudaSuccess ? 0 : -1;
}

// The compact chain of n_tiles site tiles in one launch (site_chain_kernel).  spill: also store every map to its HBM buffer (debug).
int launch_chain(const TensorModel& tm, TensorWorkspaceImpl& s, uint32_t n_tiles, float* logits, uint8_t* ml, bool spill, int sm_count, cudaStream_t stream)
{
    ChainProgram p{};
    p.n_ops = (int)tm.chain_ops.size();
    p.n_tiles = n_tiles;
    p.plane_stride = s.cplane_stride;
    p.w2 = reinterpret_cast<const float*>(tm.d_blob + tm.chain_w2_off);
    p.b2 = reinterpret_cast<const float*>(tm.d_blob + tm.chain_b2_off);
    p.logits = logits;
    p.ml = ml;
    p.out_idx = s.d_out_idx;
    p.site_rows = s.d_site_rows;
    p.x_lo_off = s.xg_stride;
    // an odd tile count: the peer CTA of the last pair works on one tile of padding rows
    if (n_tiles & 1u) {
        cudaMemsetAsync(s.d_out_idx + (size_t)n_tiles * kTileRows, 0xff, kTileRows * sizeof(uint32_t), stream);
        cudaMemsetAsync(s.d_site_rows + (size_t)n_tiles * kTileRows, 0, kTileRows * sizeof(uint32_t), stream);  // they gather X row 0
    }
    // HM_CHAIN_STAMPS=1: pair 0 stamps the third tile round of every launch; the last launch's stamps are printed after the batch
    static const bool stamps = getenv("HM_CHAIN_STAMPS") != nullptr;
    if (stamps) {
        if (!g_chain_dbg) cudaMalloc((void**)&g_chain_dbg, 512 * sizeof(long long));
        cudaMemsetAsync(g_chain_dbg, 0, 512 * sizeof(long long), stream);
        p.dbg = g_chain_dbg;
    }
    for (int i = 0; i < p.n_ops; ++i) {
        const TensorModel::ChainItem& it = tm.chain_ops[i];
        p.op[i] = it.c;
        for (int k = 0; k < it.c.n_terms; ++k)
            p.op[i].term[k].src = it.resident[k] ? nullptr : it.c.gather ? s.d_xg : s.map[it.term_map[k]];
        p.op[i].spill = (spill && !it.c.head) ? s.map[it.out_map] : nullptr;
        s.macs += (double)((n_tiles + 1) / 2 * 2) * kTileRows * it.d.macs_per_row;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(std::min<uint32_t>(((n_tiles + 1) / 2) * 2, (uint32_t)sm_count & ~1u));
    cfg.blockDim = dim3(kDenseThreads);
    cfg.dynamicSmemBytes = chain_smem_bytes();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, site_chain_kernel, p) == cudaSuccess ? 0 : -1;
}

struct SubBatch {
    uint32_t gtile0, fwd_tiles, tiles;
    uint32_t r0, r1;  // reads [r0, r1)
};

}  // namespace

int tensor_batch_run(const TensorModelHandle* models, uint32_t ctx_mask, TensorWorkspace& w, const TensorBatch& b, cudaStream_t stream,
                     int sm_count, uint32_t* launches, hm_timing* timing)
{
    TensorWorkspaceImpl* s = w.impl;
    if (!s) return tfail("tensor workspace not allocated");
    if (b.n_reads > s->reads_cap) return tfail("batch exceeds the workspace read capacity");
    const bool want_rev = (ctx_mask & 4u) != 0;
    // ---- tracks -> sub-batches: [forward tracks | reverse tracks] per sub-batch ---------------------------------------------
    std::vector<SubBatch> subs;
    uint32_t* trk_f = s->h_track_row;
    uint32_t* trk_r = s->h_track_row + s->reads_cap;
    uint32_t gtile = 0;
    uint32_t r = 0;
    while (r < b.n_reads) {
        uint32_t r1 = r, rows = 0;
        while (r1 < b.n_reads) {
            if (b.h_valid[r1]) {
                const uint32_t L = b.h_base_off[r1 + 1] - b.h_base_off[r1];
                const uint32_t tr = ((L + kHaloL + kHaloR + kTileRows - 1) / kTileRows) * kTileRows;
                const uint32_t need = tr * (want_rev ? 2u : 1u);
                if (need > s->rows_cap) return tfail("a read is longer than the dense workspace (raise HM_DENSE_ROWS)");
                if (rows + need > s->rows_cap) break;
                if (compact_mode() && b.h_read_pref && r1 > r) {
                    // compact buffers hold one row per site of a context: keep every context's sites within compact_cap
                    const uint32_t* p0 = b.h_read_pref + 4 * (size_t)r;
                    const uint32_t* p1 = b.h_read_pref + 4 * (size_t)(r1 + 1);
                    const uint32_t worst = std::max(std::max(p1[0] - p0[0], p1[1] - p0[1]), (p1[2] - p0[2]) + (p1[3] - p0[3]));
                    if (worst > s->compact_cap) break;
                }
                rows += need;
            }
            ++r1;
        }
        if (rows) {
            SubBatch sb{gtile, 0, 0, r, r1};
            for (int strand = 0; strand < (want_rev ? 2 : 1); ++strand) {
                for (uint32_t q = r; q < r1; ++q) {
                    if (!b.h_valid[q]) continue;
                    const uint32_t L = b.h_base_off[q + 1] - b.h_base_off[q];
                    const uint32_t nt = (L + kHaloL + kHaloR + kTileRows - 1) / kTileRows;
                    if (gtile + nt > s->tiles_cap) return tfail("internal: tile table overflow");
                    (strand ? trk_r : trk_f)[q] = gtile * kTileRows;
                    for (uint32_t t = 0; t < nt; ++t) {
                        s->h_tile_read[gtile + t] = q | ((uint32_t)strand << 31);
                        s->h_tile_first[gtile + t] = gtile;
                    }
                    gtile += nt;
                }
                if (strand == 0) sb.fwd_tiles = gtile - sb.gtile0;
            }
            sb.tiles = gtile - sb.gtile0;
            subs.push_back(sb);
        }
        r = r1;
    }
    s->last_subs = (uint32_t)subs.size();
    s->last_x_rows = gtile * kTileRows;
    for (uint32_t& g : s->last_groups) g = 0;
    if (gtile) {
        TCUDA("track tables", cudaMemcpyAsync(s->d_tile_read, s->h_tile_read, (size_t)gtile * 4, cudaMemcpyHostToDevice, stream));
        TCUDA("track tables", cudaMemcpyAsync(s->d_tile_first, s->h_tile_first, (size_t)gtile * 4, cudaMemcpyHostToDevice, stream));
        TCUDA("track tables", cudaMemcpyAsync(s->d_track_row, s->h_track_row, (size_t)2 * s->reads_cap * 4, cudaMemcpyHostToDevice, stream));
    }
    // ---- the plan, sub-batch by sub-batch -------------------------------------------------------------------------------------
    TCUDA("dense plan", cudaEventRecord(s->ev0, stream));
    uint32_t dense_launches = 0;
    s->macs = 0;
    // HM_OP_TIMES=1: per-op device time (events between launches; serialises nothing, but PDL overlap is attributed to the
    // earlier op), summed over the sub-batches of this batch and printed to stderr.  Analysis aid, off by default.
    static const bool prof = getenv("HM_OP_TIMES") != nullptr;
    static std::vector<cudaEvent_t> pe(1 << 16);
    static std::vector<int> pk(1 << 16);
    size_t np = 0;
    const bool compact = compact_mode();
    const uint32_t first[4] = {0, b.class_count[0], b.class_count[0] + b.class_count[1], b.class_count[0] + b.class_count[1] + b.class_count[2]};
    auto stamp = [&](int key) {
        if (!prof || np >= pe.size()) return;
        cudaEventCreate(&pe[np]);
        cudaEventRecord(pe[np], stream);
        pk[np++] = key;
    };
    if (!compact) {
        // every op on every row (HM_DENSE_ALL, the first version of this path): sub-batch by sub-batch
        s->x_cur = nullptr;
        for (const SubBatch& sb : subs) {
            track_features_kernel<<<sb.tiles, 128, 0, stream>>>(b.d_bcode, b.d_kinf, b.d_base_off, s->d_tile_read, s->d_tile_first, sb.gtile0,
                                                              s->map[MAP_X], s->map[MAP_X] + s->plane_stride);
            ++*launches;
            for (int c = 0; c < 3; ++c) {
                const uint32_t n_sites_c = b.class_count[c] + (c == 2 ? b.class_count[3] : 0u);
                if (!(ctx_mask & (1u << c)) || !n_sites_c) continue;
                if (!models[c].p) return tfail("model of an enabled context is missing");
                const uint32_t nt = (c == 2) ? sb.tiles : sb.fwd_tiles;
                if (!nt) continue;
                float* lg = s->d_logit[c] + (size_t)sb.gtile0 * kTileRows * 2;
                for (const DevOp& d : models[c].p->ops) {
                    launch_op(d, *s, nt, lg, sm_count, stream);
                    ++dense_launches;
                }
            }
            TCUDA("dense plan", cudaGetLastError());
        }
    } else if (gtile) {
        // X of the whole batch once (32 B per row); then, context by context, GROUPS of consecutive sub-batches: the dense chain
        // Y1 .. Y6 runs per sub-batch and scatters the rows sites need into the compact buffers, the compact chain (F, G, tail)
        // runs once per group over all its sites -- as many sub-batches as the compact buffers hold, so that the sparse contexts
        // (CpG, CHG: ~60 k sites per sub-batch) get launches of ~1 M rows instead of 15 short ones per op.
        track_features_kernel<<<gtile, 128, 0, stream>>>(b.d_bcode, b.d_kinf, b.d_base_off, s->d_tile_read, s->d_tile_first, 0, s->d_xg,
                                                        s->d_xg + s->xg_stride);
        ++*launches;
        s->x_stride_cur = s->xg_stride;
        struct Seg { uint32_t first_a, n_a, first_b, n, m_off; };
        for (int c = 0; c < 3; ++c) {
            const uint32_t n_sites_c = b.class_count[c] + (c == 2 ? b.class_count[3] : 0u);
            if (!(ctx_mask & (1u << c)) || !n_sites_c) continue;
            if (!models[c].p) return tfail("model of an enabled context is missing");
            const TensorModel& tm = *models[c].p;
            size_t g0 = 0;
            while (g0 < subs.size()) {
                // ---- dense chains of the group's sub-batches ---------------------------------------------------------------
                std::vector<Seg> segs;
                uint32_t m_off = 0;
                size_t g1 = g0;
                for (; g1 < subs.size(); ++g1) {
                    const SubBatch& sb = subs[g1];
                    const uint32_t nt = (c == 2) ? sb.tiles : sb.fwd_tiles;
                    // sites of this context inside the sub-batch: class regions are ordered by read, so each is one range
                    const uint32_t* p0 = b.h_read_pref + 4 * (size_t)sb.r0;
                    const uint32_t* p1 = b.h_read_pref + 4 * (size_t)sb.r1;
                    const uint32_t first_a = first[c] + p0[c], n_a = p1[c] - p0[c];
                    const uint32_t first_b = c == 2 ? first[3] + p0[3] : 0u, n_b = c == 2 ? p1[3] - p0[3] : 0u;
                    const uint32_t n = n_a + n_b;
                    if (!nt || !n) continue;
                    if (n > s->compact_cap) return tfail("internal: more sites than compact rows in a sub-batch");
                    if (m_off + n > s->compact_cap) break;  // the group is full: run its compact chain first
                    TCUDA("site rows", cudaMemsetAsync(s->d_site_of_row, 0xff, ((size_t)nt * kTileRows + kSlackRows) * sizeof(int32_t), stream));
                    const uint32_t n_pad = ((m_off + n + kTileRows - 1) / kTileRows) * kTileRows - m_off;
                    site_rows_kernel<<<(n_pad + 255) / 256, 256, 0, stream>>>(s->d_track_row, s->d_track_row + s->reads_cap, b.d_base_off,
                                                                              b.d_site_read, b.d_site_pos, first_a, n_a, first_b, n, n_pad,
                                                                              sb.gtile0 * kTileRows, m_off, b.d_site_out, s->d_site_rows,
                                                                              s->d_site_of_row, s->d_out_idx);
                    ++*launches;
                    s->x_cur = s->d_xg + (size_t)sb.gtile0 * kTileRows * 16;
                    int op_i = 0;
                    for (const DevOp& d : tm.ops) {
                        if (!d.compact && !(tm.fused12 && op_i == tm.i_y2)) {
                            stamp(c * 64 + op_i);
                            if (tm.fused12 && op_i == tm.i_y1) launch_fused12(tm.f12_c1, tm.ops[tm.i_y2], *s, nt * kTileRows, sm_count, stream);
                            else launch_op(d, *s, nt, nullptr, sm_count, stream);
                            ++dense_launches;
                        }
                        ++op_i;
                    }
                    segs.push_back(Seg{first_a, n_a, first_b, n, m_off});
                    m_off += n;
                }
                // ---- the compact chain over all sites of the group, then logits -> ML bytes per sub-batch segment --------------
                if (m_off) {
                    const uint32_t n_tiles = (m_off + kTileRows - 1) / kTileRows;
                    int op_i = 0;
                    const bool chain = tm.chain && !s->no_chain;
                    for (const DevOp& d : tm.ops) {
                        if (d.compact && !chain) {  // the chain kernel runs every compact op, F1 / G1 included
                            stamp(c * 64 + op_i);
                            launch_op(d, *s, n_tiles, s->d_clogit, sm_count, stream);
                            ++dense_launches;
                        }
                        ++op_i;
                    }
                    if (chain) {
                        stamp(c * 64 + 63);
                        launch_chain(tm, *s, n_tiles, b.d_logits, b.d_ml, s->spill_chain, sm_count, stream);  // writes logits + ML bytes itself
                        ++dense_launches;
                    }
                    stamp(-1);
                    for (const Seg& sg : segs) {
                        if (chain) break;
                        site_finish_kernel<<<(sg.n + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const float2*>(s->d_clogit) + sg.m_off, b.d_site_out,
                                                                                   sg.first_a, sg.n_a, sg.first_b, sg.n, b.d_logits, b.d_ml);
                        ++*launches;
                    }
                }
                TCUDA("dense plan", cudaGetLastError());
                ++s->last_groups[c];
                g0 = g1 > g0 ? g1 : g0 + 1;
            }
        }
    }
    TCUDA("dense plan", cudaEventRecord(s->ev1, stream));
    *launches += dense_launches;
    if (prof && np) {
        cudaStreamSynchronize(stream);
        double acc[3][64] = {};
        for (size_t i = 0; i + 1 < np; ++i) {
            if (pk[i] < 0) continue;
            float ms = 0;
            cudaEventElapsedTime(&ms, pe[i], pe[i + 1]);
            acc[pk[i] / 64][pk[i] % 64] += ms;
        }
        for (size_t i = 0; i < np; ++i) cudaEventDestroy(pe[i]);
        for (int c = 0; c < 3; ++c) {
            if (!models[c].p) continue;
            double tot = 0;
            fprintf(stderr, "ctx %d op ms:", c);
            for (size_t k = 0; k < models[c].p->ops.size(); ++k) {
                fprintf(stderr, " %s%.2f", models[c].p->ops[k].compact ? "c" : "D", acc[c][k]);
                tot += acc[c][k];
            }
            if (acc[c][63] > 0) fprintf(stderr, "  chain %.2f", acc[c][63]);
            fprintf(stderr, "  | total %.2f\n", tot + acc[c][63]);
        }
    }
    // ---- per-site lookup (dense-all mode only; compact runs finish per sub-batch) ----------------------------------------------
    for (int k = 0; k < 4 && !compact; ++k) {
        if (!b.class_count[k]) continue;
        const int c = k < 3 ? k : 2;
        site_lookup_kernel<<<(b.class_count[k] + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const float2*>(s->d_logit[c]), s->d_track_row,
                                                                               s->d_track_row + s->reads_cap, b.d_base_off, b.d_site_read,
                                                                               b.d_site_pos, b.d_site_out, first[k], b.class_count[k], b.d_logits,
                                                                               b.d_ml);
        ++*launches;
    }
    TCUDA("site lookup", cudaGetLastError());
    if (g_f12_dbg) {  // HM_F12_STAMPS: where the fused kernel's CTA 0 spent its first tiles (cycles relative to the first stamp)
        cudaStreamSynchronize(stream);
        std::vector<long long> h(48 * 16);
        cudaMemcpy(h.data(), g_f12_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        fprintf(stderr, "fused12 stamps (cycles since tile 0): tile: mma1-issue-start mma1-issued t_empty | full[0..7] seen | mma2-issued || epi1: a1_full chunk-a chunk-b | epi2: t_full\n");
        for (int t = 8; t < 24; ++t) {
            fprintf(stderr, "%2d:", t);
            for (int k = 0; k < 16; ++k) fprintf(stderr, " %7lld%s", h[16 * t + k] ? h[16 * t + k] - t0 : -1, (k == 2 || k == 10 || k == 11 || k == 14) ? " |" : "");
            fprintf(stderr, "\n");
        }
    }
    if (g_g2_dbg) {  // HM_G2_STAMPS: where pair 0 of one dense_gemm2 launch spent tiles 16 .. 47
        cudaStreamSynchronize(stream);
        g_g2_count = 0;
        std::vector<long long> h(512);
        cudaMemcpy(h.data(), g_g2_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        fprintf(stderr, "dense_gemm2 stamps (N = %d), per tile: MMA warp [wait acc buffer | wait ring stages | tile period] epilogue warp [wait acc | store | period] "
                        "producer warp 0 [wait slot]\n", g_g2_which);
        for (int t = 1; t < 32 && h[16 * t]; ++t) {
            const long long* d = h.data() + 16 * t;
            const long long* q = d - 16;
            fprintf(stderr, " %2d: mma [%5lld | %5lld | %5lld]  epi [%5lld | %5lld | %5lld]  prod [%5lld]\n", 16 + t, d[1] - d[0], d[2], d[0] - q[0], d[5] - d[4],
                    d[6] - d[5], d[6] - q[6], d[8]);
        }
    }
    if (g_chain_dbg) {  // HM_CHAIN_STAMPS: ring timeline of pair 0 of the last chain launch (cycles; each SM has its own clock)
        cudaStreamSynchronize(stream);
        std::vector<long long> h(512);
        cudaMemcpy(h.data(), g_chain_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        fprintf(stderr, "chain MMA warp of pair 0, steps of tile round 2: [cycles waiting for the slot] cycles until the next wait ...\n");
        for (int i = 0; i < 124 && h[2 * i]; ++i)
            fprintf(stderr, " [%lld] %lld", h[2 * i + 1] - h[2 * i], i < 123 && h[2 * i + 2] ? h[2 * i + 2] - h[2 * i + 1] : -1);
        fprintf(stderr, "\nchain epilogue warp 0 of pair 0, ops of tile round 2: wait for the MMAs / work (cycles), start relative to op 0\n");
        for (int i = 0; i < kChainMaxOps && h[256 + 3 * i]; ++i)
            fprintf(stderr, " op%d @%lld: %lld / %lld\n", i, h[256 + 3 * i] - h[256], h[256 + 3 * i + 1] - h[256 + 3 * i], h[256 + 3 * i + 2] - h[256 + 3 * i + 1]);
        fprintf(stderr, "chain MMA warp, op prologues of tile round 2: start (relative to op 0) / cycles waiting for epilogues\n");
        for (int i = 0; i < kChainMaxOps && h[320 + 2 * i]; ++i) fprintf(stderr, " op%d @%lld: %lld\n", i, h[320 + 2 * i] - h[320], h[321 + 2 * i] - h[320 + 2 * i]);
    }
    if (timing) {
        timing->top_kernel_launches = dense_launches;
        timing->executed_flops = s->macs * 2.0 * 3.0;
    }
    return 0;
}

// Device time of the dense plan of the last batch (ms); valid after the stream has been synchronised.
float tensor_last_dense_ms(TensorWorkspace& w)
{
    float ms = 0.f;
    if (w.impl && cudaEventElapsedTime(&ms, w.impl->ev0, w.impl->ev1) != cudaSuccess) ms = 0.f;
    return ms;
}

float tensor_debug_last_op_ms() { return g_debug_op_ms; }

// ---- debug readback of the maps the PRODUCT path works on (hm_debug_dump_xmap / hm_debug_dump_acts) ------------------------------
namespace {

// out[r][g*8 .. g*8+8) = hi + lo of row rows[r] of a plane-layout map; a negative row gives NaN ("the product path never
// materialises this value").
__global__ void __launch_bounds__(256)
debug_rows_kernel(const uint8_t* __restrict__ map, unsigned long long plane_stride, uint32_t groups, const long long* __restrict__ rows,
                  uint32_t n, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * groups) return;
    const uint32_t r = i / groups, g = i % groups;
    const long long row = rows[r];
    float* o = out + ((size_t)r * groups + g) * 8;
    if (row < 0) {
        for (int e = 0; e < 8; ++e) o[e] = __int_as_float(0x7fc00000);
        return;
    }
    const uint4 hi = *reinterpret_cast<const uint4*>(map + (unsigned long long)g * plane_stride + (unsigned long long)row * 16ull);
    const uint4 lo = *reinterpret_cast<const uint4*>(map + (unsigned long long)(groups + g) * plane_stride + (unsigned long long)row * 16ull);
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
    for (int j = 0; j < 4; ++j) {
        o[2 * j] = __uint_as_float(h[j] << 16) + __uint_as_float(l[j] << 16);
        o[2 * j + 1] = __uint_as_float(h[j] & 0xffff0000u) + __uint_as_float(l[j] & 0xffff0000u);
    }
}

int debug_read_rows(const uint8_t* map, unsigned long long plane_stride, uint32_t groups, const std::vector<long long>& rows, float* out,
                    cudaStream_t stream)
{
    if (rows.empty()) return 0;
    long long* d_rows = nullptr;
    float* d_out = nullptr;
    const size_t n = rows.size();
    TCUDA("debug rows", cudaMalloc((void**)&d_rows, n * sizeof(long long)));
    TCUDA("debug rows", cudaMalloc((void**)&d_out, n * groups * 8 * sizeof(float)));
    TCUDA("debug rows", cudaMemcpyAsync(d_rows, rows.data(), n * sizeof(long long), cudaMemcpyHostToDevice, stream));
    debug_rows_kernel<<<(uint32_t)((n * groups + 255) / 256), 256, 0, stream>>>(map, plane_stride, groups, d_rows, (uint32_t)n, d_out);
    TCUDA("debug rows", cudaGetLastError());
    TCUDA("debug rows", cudaMemcpyAsync(out, d_out, n * groups * 8 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    TCUDA("debug rows", cudaStreamSynchronize(stream));
    cudaFree(d_rows);
    cudaFree(d_out);
    return 0;
}

// X-map row of the s-row (strand offset o - 201) of a site; the site's window is rows s + 1 .. s + 401.
long long site_s_row(const TensorWorkspaceImpl& s, uint32_t read, bool rev, int o)
{
    const uint32_t trk = (rev ? s.h_track_row + s.reads_cap : s.h_track_row)[read];
    return (long long)trk + kHaloL + o - 201;
}

}  // namespace

// Debug reruns (hm_debug_dump_acts): the chain kernel also stores the maps it normally keeps in shared memory.
void tensor_debug_spill(TensorWorkspace& w, bool on)
{
    if (w.impl) w.impl->spill_chain = on;
}

int tensor_debug_xwindow(TensorWorkspace& w, uint32_t n, const uint32_t* read, const uint8_t* rev, const int32_t* o, float* out, cudaStream_t stream)
{
    TensorWorkspaceImpl* s = w.impl;
    if (!s || !s->d_xg || !s->last_x_rows) return tfail("debug X window: no X map is resident (tensor path, compact mode, after a batch)");
    std::vector<long long> rows((size_t)n * 401);
    for (uint32_t i = 0; i < n; ++i) {
        if (read[i] >= s->reads_cap) return tfail("debug X window: read index out of range");
        const long long sr = site_s_row(*s, read[i], rev[i] != 0, o[i]);
        for (int j = 0; j < 401; ++j) {
            const long long r = sr + 1 + j;
            if (r < 0 || r >= (long long)s->last_x_rows + kSlackRows) return tfail("debug X window: row outside the X map");
            rows[(size_t)i * 401 + j] = r;
        }
    }
    return debug_read_rows(s->d_xg, s->xg_stride, 1, rows, out, stream);
}

int tensor_debug_site_acts(const TensorModelHandle& model, int ctx, TensorWorkspace& w, uint32_t n, const uint32_t* read, const uint8_t* rev,
                           const int32_t* o, const uint32_t* compact_row, int layer, float* out, int* n_out, int* channels, cudaStream_t stream)
{
    TensorWorkspaceImpl* s = w.impl;
    const TensorModel* tm = model.p;
    if (!s || !tm) return tfail("debug activations: no tensor model / workspace");
    if (!compact_mode()) return tfail("debug activations: not available under HM_DENSE_ALL");
    if (layer < 1 || layer > 8) return tfail("debug activations: layer must be 1..8");
    const int nl = tm->lens[layer];
    const int C = layer <= 6 ? kLayerCout[layer - 1] : 64;
    *n_out = nl;
    *channels = C;
    if (!n) return 0;
    if (s->last_subs != 1 || s->last_groups[ctx] != 1)
        return tfail("debug activations: the batch must fit one sub-batch and one compact group (use a smaller batch)");
    const uint32_t groups = (uint32_t)C / 8;
    // Y1 exists in HBM only as the compact copies the fused kernel scatters for F2 / G2
    const DevOp* y1_prod = nullptr;
    if (layer == 1 && tm->fused12) y1_prod = &tm->f12_c1;
    for (int v = 0; v < nl; ++v) {
        const uint8_t* map = nullptr;
        unsigned long long stride = 0;
        std::vector<long long> rows(n);
        bool compact = false;
        int shift = 0;
        if (layer >= 7) { map = s->map[(layer == 7 ? MAP_T7 : MAP_T8) + v]; compact = true; }
        else if (v == 0) { map = s->map[MAP_F + layer - 1]; compact = true; }
        else if (v == nl - 1) { map = s->map[MAP_G + layer - 1]; compact = true; }
        else {
            shift = -((1 << layer) - 2) + v * (1 << layer);
            if (y1_prod) {
                for (int k = 0; k < y1_prod->p.n_scatter; ++k)
                    if (y1_prod->p.sc_shift[k] == shift) { map = s->map[y1_prod->sc_map[k]]; compact = true; }
            } else {
                map = s->map[MAP_Y + layer - 1];
            }
        }
        stride = compact ? s->cplane_stride : s->plane_stride;
        for (uint32_t i = 0; i < n; ++i) {
            if (!map) { rows[i] = -1; continue; }
            if (compact) { rows[i] = compact_row[i]; continue; }
            const long long r = site_s_row(*s, read[i], rev[i] != 0, o[i]) + shift;
            if (r < 0 || r >= (long long)s->rows_cap + kSlackRows) return tfail("debug activations: row outside the dense map");
            rows[i] = r;
        }
        std::vector<float> tmp((size_t)n * C);
        if (!map) {
            for (float& x : tmp) x = NAN;
        } else if (debug_read_rows(map, stride, groups, rows, tmp.data(), stream)) return -1;
        for (uint32_t i = 0; i < n; ++i) memcpy(out + ((size_t)i * nl + v) * C, tmp.data() + (size_t)i * C, (size_t)C * sizeof(float));
    }
    return 0;
}

// ---- unit-test hook: one op on caller-provided fp32 maps ------------------------------------------------------------------------
int tensor_debug_dense_op(int device, uint32_t rows, uint32_t rows_alloc, int cin, int cout, int n_src, const float* const* src, int n_terms,
                          const int32_t* term_src, const int32_t* term_shift, const float* weights, const float* bias, int conv1_taps,
                          const float* w2, const float* b2, const uint32_t* gather_rows, uint32_t gather_mask, float* out)
{
    if (!gather_rows) gather_mask = 0;
    if (rows == 0 || rows % kTileRows || (!gather_mask && rows_alloc < rows) || n_src < 1 || n_src > kMaxSegs || n_terms < 1 || n_terms > kMaxTerms)
        return tfail("hm_debug_dense_op: bad shape");
    TCUDA("debug op", cudaSetDevice(device));
    if (ensure_kernel_attr()) return -1;
    HostOp h;
    h.out = 0; h.cin = cin; h.cout = cout; h.conv1_taps = conv1_taps; h.head = (w2 != nullptr);
    h.bias.assign(bias, bias + cout);
    if (h.head) { h.w2.assign(w2, w2 + 2 * (size_t)cout); h.b2.assign(b2, b2 + 2); }
    const size_t wsz = conv1_taps > 0 ? (size_t)conv1_taps * 8 * cout : (size_t)cin * cout;
    for (int k = 0; k < n_terms; ++k) {
        HostTerm t;
        if (term_src[k] < 0 || term_src[k] >= n_src || term_shift[k] < 0) return tfail("hm_debug_dense_op: bad term");
        t.src = term_src[k]; t.shift = term_shift[k];
        t.gather = ((gather_mask >> k) & 1u) != 0;
        t.w.assign(weights + k * wsz, weights + (k + 1) * wsz);
        h.terms.push_back(std::move(t));
    }
    DevOp d;
    std::vector<uint8_t> blob;
    std::string err;
    const bool two = getenv("HM_DENSE_2CTA") != nullptr && !gather_mask && !h.head && conv1_taps == 0;
    if (!lower_op(h, 0, cout, d, blob, err, two)) return tfail("hm_debug_dense_op: " + (err == "fit" ? std::string("op does not fit shared memory") : err));
    if (two && rows_alloc < rows + kTileRows + 512) return tfail("hm_debug_dense_op: the CTA-pair form needs one spare tile of rows");
    h.compact = gather_mask != 0;
    uint32_t max_g = 0;
    for (uint32_t r = 0; gather_mask && r < rows; ++r) max_g = std::max(max_g, gather_rows[r]);
    for (int i = 0; i < d.p.n_segs; ++i) {
        const size_t last = d.p.seg[i].gather ? (size_t)max_g + d.p.seg[i].row_off + (d.p.gather_taps > 0 ? d.p.gather_taps : 1)
                                              : (size_t)rows + d.p.seg[i].row_off + d.p.seg[i].nrows - kTileRows;
        if (last > rows_alloc) return tfail("hm_debug_dense_op: shifts run past rows_alloc");
    }
    const int groups = cin / 8;
    const unsigned long long ps = (unsigned long long)rows_alloc * 16ull;
    uint8_t *d_blob = nullptr, *d_in = nullptr, *d_out = nullptr;
    float* d_logit = nullptr;
    uint32_t* d_rows = nullptr;
    const size_t in_bytes = (size_t)n_src * 2 * groups * ps;
    std::vector<uint16_t> img(in_bytes / 2, 0);
    for (int sidx = 0; sidx < n_src; ++sidx)
        for (uint32_t r = 0; r < rows_alloc; ++r)
            for (int c = 0; c < cin; ++c) {
                const float v = src[sidx][(size_t)r * cin + c];
                const uint16_t hi = f2bf(v);
                const size_t base = (size_t)sidx * 2 * groups * (ps / 2);
                img[base + ((size_t)(c / 8) * rows_alloc + r) * 8 + c % 8] = hi;
                img[base + ((size_t)(groups + c / 8) * rows_alloc + r) * 8 + c % 8] = f2bf(v - bf2f(hi));
            }
    const int ogroups = cout / 8;
    const size_t orows = (size_t)rows + kTileRows;  // the CTA-pair form may write one tile past `rows`
    const unsigned long long ops_ = (unsigned long long)orows * 16ull;
    const size_t out_bytes = h.head ? 0 : (size_t)2 * ogroups * ops_;
    const size_t out_alloc = out_bytes + 4096;
    TCUDA("debug op", cudaMalloc((void**)&d_blob, blob.size()));
    TCUDA("debug op", cudaMalloc((void**)&d_in, in_bytes));
    TCUDA("debug op", cudaMalloc((void**)&d_out, out_alloc));
    TCUDA("debug op", cudaMalloc((void**)&d_logit, (size_t)rows * 2 * sizeof(float)));
    TCUDA("debug op", cudaMemcpy(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    TCUDA("debug op", cudaMemcpy(d_in, img.data(), in_bytes, cudaMemcpyHostToDevice));
    bind_blob(d, d_blob);
    DenseOp p = d.p;
    for (int i = 0; i < p.n_segs; ++i) {
        p.seg[i].src = d_in + (size_t)d.seg_map[i] * 2 * groups * ps;
        p.seg[i].plane_stride = ps;
    }
    if (gather_mask) {
        TCUDA("debug op", cudaMalloc((void**)&d_rows, (size_t)rows * 4));
        TCUDA("debug op", cudaMemcpy(d_rows, gather_rows, (size_t)rows * 4, cudaMemcpyHostToDevice));
    }
    p.gather_rows = d_rows;
    p.n_tiles = rows / kTileRows;
    p.out = d_out;
    p.out_plane_stride = ops_;
    p.logits = d_logit;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (const char* v = getenv("HM_DENSE_VARIANT")) p.variant = (uint32_t)atoi(v);
    long long* d_dbg = nullptr;
    if (p.variant & 32u) { cudaMalloc((void**)&d_dbg, 1024 * 8); cudaMemset(d_dbg, 0, 1024 * 8); p.dbg = d_dbg; }
    if (const char* v = getenv("HM_DENSE_RING")) p.ring = std::min(p.ring, std::max(2, atoi(v)));
    const int reps = getenv("HM_DENSE_REPS") ? atoi(getenv("HM_DENSE_REPS")) : 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (p.variant) cudaFuncSetAttribute(dense_gemm_kernel_dbg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    for (int i = 0; i < reps; ++i) {
        if (i == std::max(0, reps - 10)) cudaEventRecord(e0);
        if (d.two_cta) dense_gemm2_kernel<<<std::min<uint32_t>(((p.n_tiles + 1) / 2) * 2, (uint32_t)sms & ~1u), kDenseThreads, d.smem>>>(p);
        else if (p.variant) dense_gemm_kernel_dbg<<<std::min<uint32_t>(p.n_tiles, (uint32_t)sms), kDenseThreads, d.smem>>>(p);
        else dense_gemm_kernel<<<std::min<uint32_t>(p.n_tiles, (uint32_t)sms), kDenseThreads, d.smem>>>(p);
    }
    cudaEventRecord(e1);
    TCUDA("debug op launch", cudaGetLastError());
    TCUDA("debug op run", cudaDeviceSynchronize());
    cudaEventElapsedTime(&g_debug_op_ms, e0, e1);
    g_debug_op_ms /= (float)std::min(reps, 10);
    if (d_dbg) {
        std::vector<long long> h(1024);
        cudaMemcpy(h.data(), d_dbg, 1024 * 8, cudaMemcpyDeviceToHost);
        fprintf(stderr, "stage: issue  full-seen  stage-done  (cycles from first issue)\n");
        for (int i = 16; i < 48; ++i) fprintf(stderr, "%3d: %8lld %8lld %8lld  latency %6lld\n", i, h[i] - h[0], h[256 + i] - h[0], h[512 + i] - h[0], h[256 + i] - h[i]);
        for (int i = 0; i < 8; ++i) fprintf(stderr, "tile %d: t_empty wait %lld -> %lld\n", i, h[768 + 2 * i] - h[0], h[768 + 2 * i + 1] - h[0]);
        cudaFree(d_dbg);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (h.head) {
        TCUDA("debug op", cudaMemcpy(out, d_logit, (size_t)rows * 2 * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        std::vector<uint16_t> o(out_bytes / 2);
        TCUDA("debug op", cudaMemcpy(o.data(), d_out, out_bytes, cudaMemcpyDeviceToHost));
        for (uint32_t r = 0; r < rows; ++r)
            for (int c = 0; c < cout; ++c) {
                const float hi = bf2f(o[((size_t)(c / 8) * orows + r) * 8 + c % 8]);
                const float lo = bf2f(o[((size_t)(ogroups + c / 8) * orows + r) * 8 + c % 8]);
                out[(size_t)r * cout + c] = hi + lo;
            }
    }
    cudaFree(d_blob); cudaFree(d_in); cudaFree(d_out); cudaFree(d_logit); cudaFree(d_rows);
    return 0;
}

}  // namespace hm
