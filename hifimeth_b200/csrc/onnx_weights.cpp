// onnx_weights.cpp -- minimal protobuf wire-format walk over an ONNX ModelProto (no onnx / protobuf
// dependency).  Field numbers used (onnx.proto3): ModelProto.graph=7; GraphProto.node=1, initializer=5,
// input=11; NodeProto.input=1, output=2, op_type=4, attribute=5; AttributeProto.name=1, f=2, i=3, t=5,
// ints=8; TensorProto.dims=1, data_type=2, float_data=4, name=8, raw_data=9; ValueInfoProto.name=1, type=2;
// TypeProto.tensor_type=1; TypeProto.Tensor.shape=2; TensorShapeProto.dim=1; Dimension.dim_value=1.
#include "onnx_weights.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace hm {
namespace {

struct Span {
    const uint8_t* p = nullptr;
    size_t n = 0;
};

struct Field {
    uint32_t no = 0;
    uint32_t wt = 0;
    uint64_t v = 0;  // varint / fixed value
    Span s;          // length-delimited payload
};

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    explicit Reader(Span s) : p(s.p), end(s.p + s.n) {}
    uint64_t varint() {
        uint64_t out = 0;
        int shift = 0;
        while (p < end) {
            uint8_t b = *p++;
            out |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return out;
            shift += 7;
            if (shift > 63) break;
        }
        ok = false;
        return 0;
    }
    bool next(Field& f) {
        if (!ok || p >= end) return false;
        uint64_t key = varint();
        if (!ok) return false;
        f.no = (uint32_t)(key >> 3);
        f.wt = (uint32_t)(key & 7);
        f.s = Span{};
        switch (f.wt) {
        case 0: f.v = varint(); break;
        case 1:
            if (end - p < 8) { ok = false; return false; }
            memcpy(&f.v, p, 8); p += 8; break;
        case 2: {
            uint64_t len = varint();
            if (!ok || (uint64_t)(end - p) < len) { ok = false; return false; }
            f.s = Span{p, (size_t)len};
            p += len;
            break;
        }
        case 5: {
            if (end - p < 4) { ok = false; return false; }
            uint32_t t; memcpy(&t, p, 4); f.v = t; p += 4; break;
        }
        default: ok = false; return false;
        }
        return ok;
    }
};

struct Tensor {
    std::vector<int64_t> dims;
    std::vector<float> data;
    bool is_f32 = false;
};

std::string str(Span s) { return std::string((const char*)s.p, s.n); }

bool parse_tensor(Span s, std::string& name, Tensor& t)
{
    Reader r(s);
    Field f;
    Span raw, fdata;
    int dtype = 0;
    while (r.next(f)) {
        if (f.no == 1) {
            if (f.wt == 0) t.dims.push_back((int64_t)f.v);
            else { Reader rr(f.s); while (rr.p < rr.end && rr.ok) t.dims.push_back((int64_t)rr.varint()); }
        } else if (f.no == 2) dtype = (int)f.v;
        else if (f.no == 8) name = str(f.s);
        else if (f.no == 9) raw = f.s;
        else if (f.no == 4 && f.wt == 2) fdata = f.s;
    }
    if (!r.ok) return false;
    t.is_f32 = (dtype == 1);
    if (!t.is_f32) return true;
    Span src = raw.n ? raw : fdata;
    t.data.resize(src.n / 4);
    memcpy(t.data.data(), src.p, t.data.size() * 4);
    size_t want = 1;
    for (auto d : t.dims) want *= (size_t)d;
    return want == t.data.size();
}

struct Node {
    std::string op;
    std::vector<std::string> in, out;
    std::map<std::string, int64_t> iattr;
    std::map<std::string, float> fattr;
    std::map<std::string, std::vector<int64_t>> ints;
    bool has_t = false;
    Tensor t;
};

bool parse_node(Span s, Node& n)
{
    Reader r(s);
    Field f;
    while (r.next(f)) {
        if (f.no == 1) n.in.push_back(str(f.s));
        else if (f.no == 2) n.out.push_back(str(f.s));
        else if (f.no == 4) n.op = str(f.s);
        else if (f.no == 5) {
            Reader ra(f.s);
            Field a;
            std::string an;
            bool has_i = false, has_f = false;
            int64_t iv = 0;
            float fv = 0;
            std::vector<int64_t> ints;
            Span ts;
            while (ra.next(a)) {
                if (a.no == 1) an = str(a.s);
                else if (a.no == 3) { iv = (int64_t)a.v; has_i = true; }
                else if (a.no == 2) { uint32_t u = (uint32_t)a.v; memcpy(&fv, &u, 4); has_f = true; }
                else if (a.no == 5) ts = a.s;
                else if (a.no == 8) {
                    if (a.wt == 0) ints.push_back((int64_t)a.v);
                    else { Reader rr(a.s); while (rr.p < rr.end && rr.ok) ints.push_back((int64_t)rr.varint()); }
                }
            }
            if (!ra.ok) return false;
            if (has_i) n.iattr[an] = iv;
            if (has_f) n.fattr[an] = fv;
            if (!ints.empty()) n.ints[an] = ints;
            if (ts.n) {
                std::string tn;
                if (!parse_tensor(ts, tn, n.t)) return false;
                n.has_t = true;
            }
        }
    }
    return r.ok;
}

// dims of the first graph input that is not an initializer; 0 for a symbolic dimension
bool parse_input_dims(Span vi, std::string& name, std::vector<int64_t>& dims)
{
    Reader r(vi);
    Field f;
    while (r.next(f)) {
        if (f.no == 1) name = str(f.s);
        else if (f.no == 2) {
            Reader rt(f.s);
            Field ft;
            while (rt.next(ft)) {
                if (ft.no != 1) continue;  // tensor_type
                Reader rtt(ft.s);
                Field fs;
                while (rtt.next(fs)) {
                    if (fs.no != 2) continue;  // shape
                    Reader rs(fs.s);
                    Field fd;
                    while (rs.next(fd)) {
                        if (fd.no != 1) continue;  // dim
                        Reader rd(fd.s);
                        Field fv;
                        int64_t v = 0;
                        while (rd.next(fv)) if (fv.no == 1) v = (int64_t)fv.v;
                        dims.push_back(v);
                    }
                }
            }
        }
    }
    return r.ok;
}

}  // namespace

bool load_onnx_model(const std::string& path, CnnModel& m, std::string& err)
{
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) { err = "cannot open model file " + path; return false; }
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t got;
    while ((got = fread(tmp, 1, sizeof(tmp), fp)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(fp);

    Reader rm(Span{buf.data(), buf.size()});
    Field f;
    Span graph;
    while (rm.next(f)) if (f.no == 7 && f.wt == 2) graph = f.s;
    if (!rm.ok || !graph.n) { err = "not an ONNX model (no graph): " + path; return false; }

    std::map<std::string, Tensor> consts;
    std::vector<Node> nodes;
    std::vector<std::pair<std::string, std::vector<int64_t>>> inputs;
    Reader rg(graph);
    while (rg.next(f)) {
        if (f.no == 5) {
            std::string name;
            Tensor t;
            if (!parse_tensor(f.s, name, t)) { err = "corrupt initializer in " + path; return false; }
            if (t.is_f32) consts[name] = std::move(t);
        } else if (f.no == 1) {
            Node n;
            if (!parse_node(f.s, n)) { err = "corrupt node in " + path; return false; }
            nodes.push_back(std::move(n));
        } else if (f.no == 11) {
            std::string name;
            std::vector<int64_t> dims;
            if (!parse_input_dims(f.s, name, dims)) { err = "corrupt graph input in " + path; return false; }
            inputs.emplace_back(name, dims);
        }
    }
    if (!rg.ok) { err = "corrupt graph in " + path; return false; }
    for (auto& n : nodes)
        if (n.op == "Constant" && n.has_t && n.t.is_f32 && !n.out.empty()) consts[n.out[0]] = n.t;

    // Input geometry: rank 3 with static kmer / feature dims (mod_main.cpp:41-57).
    bool found = false;
    for (auto& in : inputs) {
        if (consts.count(in.first)) continue;
        if (in.second.size() != 3) { err = "model input rank must be 3 (Batch, Kmer, Features): " + path; return false; }
        if (in.second[1] <= 0 || in.second[2] <= 0) { err = "model Kmer or Feature dimension is dynamic: " + path; return false; }
        m.kmer = (int)in.second[1];
        m.features = (int)in.second[2];
        found = true;
        break;
    }
    if (!found) { err = "model has no data input: " + path; return false; }

    auto get = [&](const std::string& name) -> const Tensor* {
        auto it = consts.find(name);
        return it == consts.end() ? nullptr : &it->second;
    };
    bool have_bn = false;
    int n_fc = 0;
    bool pending_matmul_bias = false;
    for (auto& n : nodes) {
        if (n.op == "BatchNormalization") {
            if (n.in.size() < 5) { err = "BatchNormalization arity"; return false; }
            const Tensor *w = get(n.in[1]), *b = get(n.in[2]), *mu = get(n.in[3]), *var = get(n.in[4]);
            if (!w || !b || !mu || !var) { err = "BatchNormalization weights missing"; return false; }
            m.bn_w = w->data; m.bn_b = b->data; m.bn_mean = mu->data; m.bn_var = var->data;
            if (n.fattr.count("epsilon")) m.bn_eps = n.fattr["epsilon"];
            have_bn = true;
        } else if (n.op == "Conv") {
            const Tensor* w = n.in.size() > 1 ? get(n.in[1]) : nullptr;
            const Tensor* b = n.in.size() > 2 ? get(n.in[2]) : nullptr;
            if (!w || !b || w->dims.size() != 3) { err = "Conv weights missing"; return false; }
            auto st = n.ints.find("strides");
            auto pd = n.ints.find("pads");
            if (st == n.ints.end() || st->second.size() != 1 || st->second[0] != 2 || pd == n.ints.end() ||
                pd->second.size() != 2 || pd->second[0] != 1 || pd->second[1] != 1) {
                err = "Conv must be stride 2, pads [1,1]";
                return false;
            }
            ConvLayer c;
            c.cout = (int)w->dims[0]; c.cin = (int)w->dims[1]; c.k = (int)w->dims[2];
            c.w = w->data; c.b = b->data;
            m.convs.push_back(std::move(c));
        } else if (n.op == "Gemm") {
            const Tensor* w = get(n.in[1]);
            const Tensor* b = n.in.size() > 2 ? get(n.in[2]) : nullptr;
            if (!w || !b || w->dims.size() != 2) { err = "Gemm weights missing"; return false; }
            bool transB = n.iattr.count("transB") && n.iattr["transB"];
            int d0 = (int)w->dims[0], d1 = (int)w->dims[1];
            std::vector<float> wm;  // [out][in]
            int out_f, in_f;
            if (transB) { out_f = d0; in_f = d1; wm = w->data; }
            else {
                out_f = d1; in_f = d0; wm.resize(w->data.size());
                for (int i = 0; i < in_f; ++i) for (int o = 0; o < out_f; ++o) wm[(size_t)o * in_f + i] = w->data[(size_t)i * out_f + o];
            }
            if (n_fc == 0) { m.fc1_w = wm; m.fc1_b = b->data; m.fc1_out = out_f; m.fc1_in = in_f; }
            else { m.fc2_w = wm; m.fc2_b = b->data; m.fc2_out = out_f; }
            ++n_fc;
        } else if (n.op == "MatMul") {
            const Tensor* w = get(n.in[1]);
            if (!w || w->dims.size() != 2) { err = "MatMul weights missing"; return false; }
            int in_f = (int)w->dims[0], out_f = (int)w->dims[1];
            std::vector<float> wm(w->data.size());
            for (int i = 0; i < in_f; ++i) for (int o = 0; o < out_f; ++o) wm[(size_t)o * in_f + i] = w->data[(size_t)i * out_f + o];
            if (n_fc == 0) { m.fc1_w = wm; m.fc1_out = out_f; m.fc1_in = in_f; }
            else { m.fc2_w = wm; m.fc2_out = out_f; }
            pending_matmul_bias = true;
        } else if (n.op == "Add" && pending_matmul_bias) {
            const Tensor* b = get(n.in[1]);
            if (!b) b = get(n.in[0]);
            if (!b) { err = "FC bias missing"; return false; }
            if (n_fc == 0) m.fc1_b = b->data; else m.fc2_b = b->data;
            ++n_fc;
            pending_matmul_bias = false;
        }
    }
    if (!have_bn || m.convs.size() != 8 || n_fc != 2) { err = "unexpected graph (need bn0, 8 convs, 2 FCs): " + path; return false; }
    if ((int)m.bn_w.size() != m.features || m.convs[0].cin != m.features) { err = "feature dimension mismatch: " + path; return false; }
    // Conv chain geometry must end in fc1_in values (flatten of [C8][L8]).
    int len = m.kmer;
    for (auto& c : m.convs) len = (len + 2 - c.k) / 2 + 1;
    if (m.convs.back().cout * len != m.fc1_in || m.fc2_out != 2 || (int)m.fc2_w.size() != 2 * m.fc1_out) {
        err = "conv/fc geometry mismatch: " + path;
        return false;
    }
    return true;
}


// ---- TorchScript archive (.pt) ------------------------------------------------------------------------------------------------------
namespace {
uint16_t z16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint32_t z32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
}  // namespace

bool load_pt_model(const std::string& path, CnnModel& m, std::string& err)
{
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) { err = "cannot open model file " + path; return false; }
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t got;
    while ((got = fread(tmp, 1, sizeof(tmp), fp)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(fp);
    const size_t n = buf.size();
    // end-of-central-directory record: signature 0x06054b50 within the last 64 KiB + 22 bytes
    size_t eocd = (size_t)-1;
    for (size_t back = 22; back <= n && back <= 22 + 65535; ++back)
        if (z32(&buf[n - back]) == 0x06054b50u) { eocd = n - back; break; }
    if (eocd == (size_t)-1) { err = "not a TorchScript archive (no ZIP directory): " + path; return false; }
    uint64_t n_entries = z16(&buf[eocd + 10]), cd_size = z32(&buf[eocd + 12]), cd_off = z32(&buf[eocd + 16]);
    if (cd_off == 0xffffffffu || n_entries == 0xffffu) {  // ZIP64: locator 20 bytes before the EOCD -> ZIP64 EOCD record
        if (eocd < 20 || z32(&buf[eocd - 20]) != 0x07064b50u) { err = "corrupt ZIP64 directory in " + path; return false; }
        uint64_t rec = 0;
        for (int i = 7; i >= 0; --i) rec = (rec << 8) | buf[eocd - 20 + 8 + i];
        if (rec + 56 > n || z32(&buf[rec]) != 0x06064b50u) { err = "corrupt ZIP64 directory in " + path; return false; }
        auto z64 = [&](size_t o) { uint64_t v = 0; for (int i = 7; i >= 0; --i) v = (v << 8) | buf[o + i]; return v; };
        n_entries = z64(rec + 32); cd_size = z64(rec + 40); cd_off = z64(rec + 48);
    }
    if (cd_off + cd_size > n) { err = "corrupt ZIP directory in " + path; return false; }
    std::vector<float> c[24];
    bool have[24] = {};
    size_t p = cd_off;
    for (uint64_t e = 0; e < n_entries; ++e) {
        if (p + 46 > n || z32(&buf[p]) != 0x02014b50u) { err = "corrupt ZIP directory entry in " + path; return false; }
        const uint16_t method = z16(&buf[p + 10]), nl = z16(&buf[p + 28]), xl = z16(&buf[p + 30]), cl = z16(&buf[p + 32]);
        uint64_t csize = z32(&buf[p + 20]), usize = z32(&buf[p + 24]), lho = z32(&buf[p + 42]);
        if (p + 46 + nl + xl + cl > n) { err = "corrupt ZIP directory entry in " + path; return false; }
        const std::string name(reinterpret_cast<const char*>(&buf[p + 46]), nl);
        p += 46 + (size_t)nl + xl + cl;
        const size_t k = name.rfind("/constants/");
        if (k == std::string::npos) continue;
        const std::string idx = name.substr(k + 11);
        if (idx.empty() || idx.size() > 2 || idx.find_first_not_of("0123456789") != std::string::npos) continue;
        const int i = atoi(idx.c_str());
        if (i < 0 || i > 23) continue;
        if (method != 0 || csize != usize || usize % 4 || csize == 0xffffffffu || lho == 0xffffffffu) { err = "constant " + idx + " of " + path + " is not a stored f32 member"; return false; }
        if (lho + 30 > n || z32(&buf[lho]) != 0x04034b50u) { err = "corrupt ZIP member header in " + path; return false; }
        const uint64_t data = lho + 30 + z16(&buf[lho + 26]) + z16(&buf[lho + 28]);
        if (data + usize > n) { err = "truncated ZIP member in " + path; return false; }
        c[i].resize(usize / 4);
        memcpy(c[i].data(), &buf[data], usize);
        have[i] = true;
    }
    for (int i = 0; i < 24; ++i)
        if (!have[i]) { err = "TorchScript archive " + path + " lacks constant " + std::to_string(i) + " (expected the 24 frozen tensors of model_cnn)"; return false; }
    static const int cin[8] = {8, 128, 128, 128, 96, 96, 96, 64}, cout[8] = {128, 128, 128, 96, 96, 96, 64, 64};
    m = CnnModel{};
    m.kmer = 401; m.features = 8;  // the exported module has no shape record; training/make-torch-script.py traces [B, 401, 8]
    if (c[0].size() != 8 || c[1].size() != 8 || c[2].size() != 8 || c[3].size() != 8) { err = "bn0 tensors of " + path + " are not [8]"; return false; }
    m.bn_w = c[0]; m.bn_b = c[1]; m.bn_mean = c[2]; m.bn_var = c[3];
    m.convs.resize(8);
    for (int l = 0; l < 8; ++l) {
        ConvLayer& cv = m.convs[l];
        const size_t per = (size_t)cout[l] * cin[l];
        if (c[4 + 2 * l].size() % per || (int)c[5 + 2 * l].size() != cout[l]) { err = "conv" + std::to_string(l + 1) + " of " + path + " does not fit the channel plan"; return false; }
        cv.cout = cout[l]; cv.cin = cin[l]; cv.k = (int)(c[4 + 2 * l].size() / per);
        cv.w = c[4 + 2 * l]; cv.b = c[5 + 2 * l];
    }
    if (c[20].size() != 128 * 256 || c[21].size() != 256 || c[22].size() != 256 * 2 || c[23].size() != 2) { err = "FC tensors of " + path + " have unexpected sizes"; return false; }
    m.fc1_in = 128; m.fc1_out = 256; m.fc2_out = 2;
    m.fc1_w.resize(256 * 128);
    for (int o = 0; o < 256; ++o)
        for (int i = 0; i < 128; ++i) m.fc1_w[(size_t)o * 128 + i] = c[20][(size_t)i * 256 + o];   // stored [in][out]
    m.fc1_b = c[21];
    m.fc2_w.resize(2 * 256);
    for (int o = 0; o < 2; ++o)
        for (int i = 0; i < 256; ++i) m.fc2_w[(size_t)o * 256 + i] = c[22][(size_t)i * 2 + o];
    m.fc2_b = c[23];
    return true;
}

bool load_model_file(const std::string& path, CnnModel& out, std::string& err)
{
    if (path.size() > 3 && path.compare(path.size() - 3, 3, ".pt") == 0) return load_pt_model(path, out, err);
    return load_onnx_model(path, out, err);
}

}  // namespace hm
