// call_main.cpp -- `hifimeth call [OPTIONS] BAM MOD-BAM` on the B200 engine: the driver shape of the reference's mod_main()
// (src/app/hifimeth/mod_main.cpp:303-412) with its worker pool replaced by the C ABI of include/hm_engine.h.
//
//   reference                                              here
//   SAM_Batch + sam_read1 under a mutex (sam_batch.hpp)     BamReader: block-parallel BGZF inflate, records copied to a batch arena
//   -t worker threads: features + OpenVINO infer()          hm_pack_record -> hm_batch_submit (MM text on device) -> hm_batch_collect
//   build_one_mod_bam per read (build_mod_bam.cpp:125-248)  hm_build_mod_record_mm per read on -t host threads (bytes only)
//   pdqsort by read id + sam_write1 (mod_main.cpp:353-362)  batches are emitted in input order; BamWriter: parallel BGZF deflate
//
// Two staging slots: while the GPU works on batch k the host emits batch k-1 and packs batch k+1.
// Options follow src/app/hifimeth/mod_options.cpp:61-181: -m -l -s -b -k -c -t -v -h; -s is accepted and ignored (the site
// batch is an OpenVINO notion).  Extensions: --device N, --max-bases N, --level N.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include "../../include/hm_engine.h"
#include "bgzf_bam.h"

namespace {

struct Options {
    std::string model_dir, in_path, out_path;
    int min_read_len = 1000, site_batch = 32, reads_per_batch = 10000, keep_kinetics = 0, threads = 0, ctx_mask = 7;
    int device = 0, level = 6;
    long long max_bases = 64ll << 20;
};

std::string default_model_dir()
{
    if (const char* e = getenv("HM_MODEL_DIR")) return e;
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&default_model_dir), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.find_last_of('/');
        const std::string dir = k == std::string::npos ? "." : p.substr(0, k);
        return dir + "/../models";  // <repo>/hifimeth_b200/libhm_engine.so -> <repo>/models
    }
    return "models";
}

void usage(const char* prog, const char* cmd)
{
    fprintf(stderr, "USAGE:\n  %s %s [OPTIONS] BAM MOD-BAM\n\n", prog, cmd);
    fprintf(stderr, "DESCRIPTION:\n  Compute single molecular cytosine methylation states in BAM file reads (B200 engine)\n\n");
    fprintf(stderr, "OPTIONAL ARGUMENTS:\n  -v  Print version info and exit\n  -h  Print this help info and exit\n"
                    "  -m <Model directory>   Default: %s\n  -l <Integer>  Minimum read length considered for 5mC calling. Default = 1000\n"
                    "  -s <Integer>  Accepted for compatibility (site batch of the CPU path); ignored\n"
                    "  -b <Integer>  Number of reads in one batch. Default = 10000\n"
                    "  -k  Keep kinetic values (fi, ri, fp, rp) in modified BAM output\n"
                    "  -c <string>  5mC contexts to detect; comma separated. Default = cpg,chg,chh\n"
                    "  -t <Integer>  Number of CPU threads used for BAM inflate/deflate and record assembly\n"
                    "  --device <Integer>  CUDA device. Default = 0\n  --max-bases <Integer>  Bases per batch. Default = 67108864\n"
                    "  --level <Integer>  BGZF compression level. Default = 6\n",
            default_model_dir().c_str());
}

bool parse_ctx(const std::string& s, int& mask)
{
    mask = 0;
    size_t a = 0;
    while (a <= s.size()) {
        size_t b = s.find(',', a);
        if (b == std::string::npos) b = s.size();
        std::string t = s.substr(a, b - a);
        std::transform(t.begin(), t.end(), t.begin(), [](unsigned char c) { return (char)tolower(c); });
        if (t == "cpg") mask |= HM_CTX_CPG;
        else if (t == "chg") mask |= HM_CTX_CHG;
        else if (t == "chh") mask |= HM_CTX_CHH;
        else if (!t.empty()) return false;
        a = b + 1;
    }
    return mask != 0;
}

// 0 = ok, 1 = exit success (help / version), -1 = usage error
int parse(int argc, char** argv, Options& o)
{
    std::vector<std::string> pos;
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](long long& dst) {
            if (i + 1 >= argc) return false;
            char* end = nullptr;
            dst = strtoll(argv[++i], &end, 10);
            return end && *end == 0;
        };
        long long v = 0;
        if (a == "-h") { usage(argv[0], argv[1]); return 1; }
        if (a == "-v") { fprintf(stdout, "1.1.0 (%s)\n", hm_version()); return 1; }
        if (a == "-k") { o.keep_kinetics = 1; continue; }
        if (a == "-m") { if (i + 1 >= argc) return -1; o.model_dir = argv[++i]; continue; }
        if (a == "-c") { if (i + 1 >= argc || !parse_ctx(argv[++i], o.ctx_mask)) return -1; continue; }
        if (a == "-l") { if (!val(v) || v < 0) return -1; o.min_read_len = (int)v; continue; }
        if (a == "-s") { if (!val(v) || v < 1) return -1; o.site_batch = (int)v; continue; }
        if (a == "-b") { if (!val(v) || v < 1) return -1; o.reads_per_batch = (int)v; continue; }
        if (a == "-t") { if (!val(v) || v < 1) return -1; o.threads = (int)v; continue; }
        if (a == "--device") { if (!val(v) || v < 0) return -1; o.device = (int)v; continue; }
        if (a == "--level") { if (!val(v) || v < 0 || v > 9) return -1; o.level = (int)v; continue; }
        if (a == "--max-bases") { if (!val(v) || v < 1024 || v >= 0x7fffffffll) return -1; o.max_bases = v; continue; }
        if (a.size() > 1 && a[0] == '-') { fprintf(stderr, "unrecognised option '%s'\n", a.c_str()); return -1; }
        pos.push_back(a);
    }
    if (pos.size() != 2) return -1;
    o.in_path = pos[0];
    o.out_path = pos[1];
    if (o.model_dir.empty()) o.model_dir = default_model_dir();
    if (o.threads <= 0) o.threads = (int)std::max(1u, std::thread::hardware_concurrency());
    return 0;
}

struct Entry {
    size_t off, len;   // record body in the batch arena
    int32_t gpu_read;  // index in the submitted batch, or -1: emitted with tags stripped only
};

struct Batch {
    std::vector<uint8_t> arena;
    std::vector<Entry> entries;
    uint32_t n_gpu = 0;
    bool submitted = false;
};

}  // namespace

extern "C" int hm_call_main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s call [OPTIONS] BAM MOD-BAM\n", argc ? argv[0] : "hifimeth-b200"); return EXIT_FAILURE; }
    Options opt;
    const int pr = parse(argc, argv, opt);
    if (pr == 1) return EXIT_SUCCESS;
    if (pr < 0) { usage(argv[0], argv[1]); return EXIT_FAILURE; }
    const auto t_start = std::chrono::steady_clock::now();
    std::string err;
    hm::BamReader in;
    hm::BamHeader hdr;
    if (!in.open(opt.in_path.c_str(), opt.threads, hdr, err)) { fprintf(stderr, "[hifimeth-b200] %s: %s\n", opt.in_path.c_str(), err.c_str()); return EXIT_FAILURE; }
    // @PG line: add_cmd_to_sam_hdr, src/app/hifimeth/mod_main.cpp:101-117
    if (!hdr.text.empty() && hdr.text.back() != '\n') hdr.text += '\n';
    hdr.text += "@PG\tID:hifimeth\tPN:hifimeth\tVN:1.1.0\tCL:";
    for (int i = 0; i < argc; ++i) { if (i) hdr.text += ' '; hdr.text += argv[i]; }
    hdr.text += '\n';
    hm::BamWriter out;
    if (!out.open(opt.out_path.c_str(), opt.threads, opt.level, hdr, err)) { fprintf(stderr, "[hifimeth-b200] %s: %s\n", opt.out_path.c_str(), err.c_str()); return EXIT_FAILURE; }

    hm_config cfg{};
    cfg.model_dir = opt.model_dir.c_str();
    cfg.ctx_mask = opt.ctx_mask;
    cfg.min_read_len = opt.min_read_len;
    cfg.device = opt.device;
    cfg.n_slots = 2;
    cfg.max_reads = (uint32_t)opt.reads_per_batch;
    cfg.max_bases = (uint32_t)opt.max_bases;
    cfg.cnn_mode = HM_CNN_TENSOR;
    hm_engine* eng = nullptr;
    if (hm_engine_create(&cfg, &eng) != HM_OK) { fprintf(stderr, "[hifimeth-b200] %s\n", hm_last_error(nullptr)); return EXIT_FAILURE; }

    Batch batches[2];
    uint64_t n_reads = 0, n_bases = 0, n_sites[3] = {0, 0, 0};
    bool eof = false, failed = false;
    const uint8_t* pending_body = nullptr;  // a record that did not fit the previous batch
    size_t pending_len = 0;
    std::vector<uint8_t> pending_copy;

    auto emit = [&](int slot) -> bool {
        Batch& b = batches[slot];
        hm_call_batch calls{};
        if (b.submitted && hm_batch_collect(eng, slot, &calls) != HM_OK) { fprintf(stderr, "[hifimeth-b200] %s\n", hm_last_error(eng)); return false; }
        if (b.submitted)
            for (int c = 0; c < 3; ++c) n_sites[c] += calls.n_sites[c];
        // output arena: one bounded region per record, filled in parallel, written in order
        std::vector<size_t> ooff(b.entries.size() + 1, 0), olen(b.entries.size(), 0);
        for (size_t i = 0; i < b.entries.size(); ++i) {
            const Entry& e = b.entries[i];
            uint32_t nc = 0, text = 0;
            if (e.gpu_read >= 0) {
                nc = calls.call_off[e.gpu_read + 1] - calls.call_off[e.gpu_read];
                text = calls.mm_off[e.gpu_read + 1] - calls.mm_off[e.gpu_read];
            }
            ooff[i + 1] = ooff[i] + e.len + 64 + nc + text;
        }
        std::vector<uint8_t> obuf(ooff.back());
        std::vector<int> rcs(b.entries.size(), 0);
        hm::parallel_for(b.entries.size(), opt.threads, [&](size_t i) {
            const Entry& e = b.entries[i];
            const uint8_t* body = b.arena.data() + e.off;
            if (e.gpu_read < 0) {
                rcs[i] = hm_build_mod_record_mm(body, e.len, opt.keep_kinetics, nullptr, 0, nullptr, 0, nullptr, 0, 0, obuf.data() + ooff[i], &olen[i]);
                if (rcs[i] == HM_ERR_FORMAT) {  // not parseable as a record: pass the bytes through untouched
                    memcpy(obuf.data() + ooff[i], body, e.len);
                    olen[i] = e.len;
                    rcs[i] = 0;
                }
                return;
            }
            const uint32_t r = (uint32_t)e.gpu_read, a = calls.call_off[r], nc = calls.call_off[r + 1] - a, nf = calls.n_fwd[r];
            const uint8_t* mm = calls.mm_text + calls.mm_off[r];
            const uint32_t fl = calls.mm_fwd_len[r], rl = calls.mm_off[r + 1] - calls.mm_off[r] - fl;
            rcs[i] = hm_build_mod_record_mm(body, e.len, opt.keep_kinetics, mm, fl, mm + fl, rl, calls.ml + a, nf, nc - nf, obuf.data() + ooff[i], &olen[i]);
        });
        for (size_t i = 0; i < b.entries.size(); ++i) {
            if (rcs[i] != 0) { fprintf(stderr, "[hifimeth-b200] cannot assemble output record %zu of a batch (%d)\n", i, rcs[i]); return false; }
            if (!out.write_record(obuf.data() + ooff[i], olen[i], err)) { fprintf(stderr, "[hifimeth-b200] %s\n", err.c_str()); return false; }
        }
        b.entries.clear();
        b.arena.clear();
        b.submitted = false;
        b.n_gpu = 0;
        return true;
    };

    int cur = 0, prev = -1;
    while (!failed) {
        Batch& b = batches[cur];
        hm_read_batch rb{};
        if (hm_batch_acquire(eng, cur, &rb) != HM_OK) { fprintf(stderr, "[hifimeth-b200] %s\n", hm_last_error(eng)); failed = true; break; }
        uint32_t n = 0;
        while (!eof && b.entries.size() < (size_t)opt.reads_per_batch) {
            const uint8_t* body;
            size_t len;
            if (pending_body) { body = pending_body; len = pending_len; pending_body = nullptr; }
            else if (!in.next(body, len, err)) {
                if (!err.empty()) { fprintf(stderr, "[hifimeth-b200] %s: %s\n", opt.in_path.c_str(), err.c_str()); failed = true; }
                eof = true;
                break;
            }
            const int rc = hm_pack_record(&rb, &n, body, len, opt.min_read_len);
            if (rc == HM_ERR_ARG && n > 0) {  // batch full (bases): this record opens the next batch
                pending_copy.assign(body, body + len);
                pending_body = pending_copy.data();
                pending_len = len;
                break;
            }
            Entry e{b.arena.size(), len, rc == HM_OK ? (int32_t)(n - 1) : -1};  // too long for any batch / malformed: pass through
            b.arena.insert(b.arena.end(), body, body + len);
            b.entries.push_back(e);
            if (len >= 32) { uint32_t l; memcpy(&l, body + 16, 4); n_bases += l; }
            ++n_reads;
        }
        if (failed) break;
        b.n_gpu = n;
        if (!b.entries.empty()) {
            if (hm_batch_submit(eng, cur, n, HM_SUBMIT_MM_TEXT) != HM_OK) { fprintf(stderr, "[hifimeth-b200] %s\n", hm_last_error(eng)); failed = true; break; }
            b.submitted = true;
        }
        if (prev >= 0 && !emit(prev)) { failed = true; break; }
        prev = b.entries.empty() ? -1 : cur;
        cur ^= 1;
        if (eof && !pending_body && prev < 0) break;
        if (eof && !pending_body) {
            if (!emit(prev)) failed = true;
            break;
        }
    }
    hm_engine_destroy(eng);
    if (!failed && !out.close(err)) { fprintf(stderr, "[hifimeth-b200] %s\n", err.c_str()); failed = true; }
    if (failed) return EXIT_FAILURE;
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    const uint64_t sites = n_sites[0] + n_sites[1] + n_sites[2];
    fprintf(stderr, "[hifimeth-b200] %llu reads, %llu bases, CpG %llu, CHG %llu, CHH %llu samples in %.2f s (%.3g sites/s, %.3g reads/s)\n",
            (unsigned long long)n_reads, (unsigned long long)n_bases, (unsigned long long)n_sites[0], (unsigned long long)n_sites[1],
            (unsigned long long)n_sites[2], secs, sites / secs, n_reads / secs);
    return EXIT_SUCCESS;
}

// Round trip of the BAM codec alone (tests): read every record of `in_path`, write it unchanged to `out_path`.
extern "C" int hm_bam_copy(const char* in_path, const char* out_path, int threads, int level)
{
    std::string err;
    hm::BamReader in;
    hm::BamHeader hdr;
    if (!in.open(in_path, threads, hdr, err)) return HM_ERR_FORMAT;
    hm::BamWriter out;
    if (!out.open(out_path, threads, level, hdr, err)) return HM_ERR_ARG;
    const uint8_t* body;
    size_t len;
    long long n = 0;
    while (in.next(body, len, err)) {
        if (!out.write_record(body, len, err)) return HM_ERR_ARG;
        ++n;
    }
    if (!err.empty()) return HM_ERR_FORMAT;
    if (!out.close(err)) return HM_ERR_ARG;
    return (int)std::min<long long>(n, 0x7fffffff);
}
