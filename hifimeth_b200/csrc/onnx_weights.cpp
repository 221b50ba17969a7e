// onnx_weights.cpp -- minimal protobuf wire-format walk over an ONNX ModelProto (no onnx / protobuf
// dependency).  Field numbers used (onnx.proto3): ModelProto.graph=7; GraphProto.node=1, initializer=5,
// input=11; NodeProto.input=1, output=2, op_type=4, attribute=5; AttributeProto.name=1, f=2, i=3, t=5,
// ints=8; TensorProto.dims=1, data_type=2, float_data=4, name=8, raw_data=9; ValueInfoProto.name=1, type=2;
// TypeProto.tensor_type=1; TypeProto.Tensor.shape=2; TensorShapeProto.dim=1; Dimension.dim_value=1.
#include "onnx_weights.h"

#include <cstdio>
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

}  // namespace hm
