// call_main.cpp -- `hifimeth call [OPTIONS] BAM MOD-BAM` on the B200 engine: the driver shape of the reference's mod_main()
// (src/app/hifimeth/mod_main.cpp:303-412) with its worker pool replaced by the C ABI of include/hm_engine.h.
//
//   reference                                              here
//   SAM_Batch + sam_read1 under a mutex (sam_batch.hpp)     BamReader: read-ahead block-parallel BGZF inflate, zero-copy record framing
//   -t worker threads: features + OpenVINO infer()          hm_pack_record -> hm_batch_submit (MM text on device) -> hm_batch_collect
//   build_one_mod_bam per read (build_mod_bam.cpp:125-248)  hm_build_mod_record_mm per read on -t host threads (bytes only)
//   pdqsort by read id + sam_write1 (mod_main.cpp:353-362)  batches are emitted in input order; BamWriter: parallel BGZF deflate
//
// Threads: one reader (inflate + batch cutting) -> bounded queue of raw batches -> one worker per entry of --devices (each owns an
// engine with two staging slots: while its GPU works on batch k it assembles the records of batch k-1 and packs batch k+1) ->
// ordered outbox -> one writer (deflate).  Reads are independent, so this host work queue is the whole multi-GPU story of the
// path: no collective, weights replicated per engine, batches carry a sequence number and the writer restores
// input order (the reference sorts by read id, mod_main.cpp:353-354).
// Options follow src/app/hifimeth/mod_options.cpp:61-181: -m -l -s -b -k -c -t -v -h; -s is accepted and ignored (the site
// batch is an OpenVINO notion).  Extensions: --devices LIST, --max-bases N, --level N, --hist FILE.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/hm_engine.h"
#include "bgzf_bam.h"
#include "fast_deflate.h"

namespace {

struct Options {
    std::string model_dir, in_path, out_path, hist_path;
    int min_read_len = 1000, site_batch = 32, reads_per_batch = 10000, keep_kinetics = 0, threads = 0, ctx_mask = 7;
    int level = 1;   // BGZF level of the output.  Level 1 = the engine's own run-length + Huffman block coder (fast_deflate.h); 2 .. 9 and 0
                     // = zlib.  htslib's default (what the reference writes) is 6: on synthetic HiFi records that is 1.9 % LARGER than
                     // level 1 here and takes ~8 x the deflate time -- with 8 host cores per GPU the queue ran at 55 % of the device
                     // rate at zlib level 6
    std::vector<int> devices;   // one worker (engine) per entry; an ordinal may repeat
    long long max_bases = 0;    // 0 = chosen from the input size
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
                    "  --devices <list>  CUDA devices, comma separated; batches are dealt to one worker per entry. Default = 0\n"
                    "  --max-bases <Integer>  Bases per batch. Default: from the input size, 2 Mi .. 24 Mi\n"
                    "  --level <Integer>  BGZF compression level of the output: 1 = run-length + Huffman (built in, fastest), 0 and 2-9 = zlib (6 = htslib's default). Default = 1\n"
                    "  --hist <File>  Write the per-context histograms of the ML bytes (the input of pileup's threshold rule) as TSV\n",
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

bool parse_devices(const std::string& s, std::vector<int>& out)
{
    out.clear();
    size_t a = 0;
    while (a <= s.size()) {
        size_t b = s.find(',', a);
        if (b == std::string::npos) b = s.size();
        const std::string t = s.substr(a, b - a);
        if (t.empty() || t.find_first_not_of("0123456789") != std::string::npos || t.size() > 3) return false;
        out.push_back(atoi(t.c_str()));
        a = b + 1;
    }
    return !out.empty() && out.size() <= 64;
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
        if (a == "--device") { if (!val(v) || v < 0) return -1; o.devices.assign(1, (int)v); continue; }
        if (a == "--devices") { if (i + 1 >= argc || !parse_devices(argv[++i], o.devices)) return -1; continue; }
        if (a == "--hist") { if (i + 1 >= argc) return -1; o.hist_path = argv[++i]; continue; }
        if (a == "--level") { if (!val(v) || v < 0 || v > 9) return -1; o.level = (int)v; continue; }
        if (a == "--max-bases") { if (!val(v) || v < 1024 || v >= 0x7fffffffll) return -1; o.max_bases = v; continue; }
        if (a.size() > 1 && a[0] == '-') { fprintf(stderr, "unrecognised option '%s'\n", a.c_str()); return -1; }
        pos.push_back(a);
    }
    if (pos.size() != 2) return -1;
    o.in_path = pos[0];
    o.out_path = pos[1];
    if (o.model_dir.empty()) o.model_dir = default_model_dir();
    if (o.devices.empty()) o.devices.assign(1, 0);
    if (o.threads <= 0) o.threads = (int)std::max(1u, std::thread::hardware_concurrency());
    return 0;
}

struct Entry {
    const uint8_t* body;  // inside one of the batch's slabs
    size_t len;
};

// A run of consecutive input records; `seq` is its position in the input, which the writer restores.  Record bodies are not
// copied: they point into the inflated slabs, which the batch keeps alive.
struct RawBatch {
    uint64_t seq = 0;
    std::vector<std::shared_ptr<hm::Slab>> slabs;
    std::vector<Entry> entries;
    uint64_t bases = 0;
};

// The records of one batch after the engine: assembled output bodies, ready to be written in order.
struct OutBatch {
    uint64_t seq = 0;
    hm::Bytes obuf;  // the batch's records as they go into the BAM stream: block_size + body, one after the other (no zero fill: bgzf_bam.h)
    std::vector<size_t> ooff, olen;  // assembly only: bounded slot and final length of every record
};

// Bounded multi-producer / multi-consumer queue; close() wakes everybody (pop then drains what is left).
template <class T>
class BoundedQueue {
public:
    explicit BoundedQueue(size_t cap) : cap_(cap) {}
    bool push(T&& v)
    {
        std::unique_lock<std::mutex> lk(m_);
        not_full_.wait(lk, [&] { return q_.size() < cap_ || closed_; });
        if (closed_) return false;
        q_.push_back(std::move(v));
        not_empty_.notify_one();
        return true;
    }
    bool pop(T& v)
    {
        std::unique_lock<std::mutex> lk(m_);
        not_empty_.wait(lk, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        v = std::move(q_.front());
        q_.pop_front();
        not_full_.notify_one();
        return true;
    }
    void close()
    {
        std::lock_guard<std::mutex> lk(m_);
        closed_ = true;
        not_full_.notify_all();
        not_empty_.notify_all();
    }
    void abort()  // close and drop what is queued
    {
        std::lock_guard<std::mutex> lk(m_);
        closed_ = true;
        q_.clear();
        not_full_.notify_all();
        not_empty_.notify_all();
    }

private:
    std::mutex m_;
    std::condition_variable not_full_, not_empty_;
    std::deque<T> q_;
    size_t cap_;
    bool closed_ = false;
};

// Finished batches arrive in any order (one producer per GPU); the writer takes them in input order.  The box is BOUNDED: a
// producer blocks in put() once `cap` batches are waiting, so that GPU workers cannot run ahead of a slow writer (zlib deflate is
// the limiter on many GPUs) and pile assembled batches up in memory.  The batch the writer waits for is always admitted -- every
// other batch in the box is younger than it, so the cap cannot deadlock.
class OrderedOutbox {
public:
    explicit OrderedOutbox(size_t cap) : cap_(std::max<size_t>(cap, 1)) {}
    // false once finish() was called (the run failed): the batch is dropped.
    bool put(OutBatch&& b)
    {
        std::unique_lock<std::mutex> lk(m_);
        const uint64_t k = b.seq;
        space_.wait(lk, [&] { return done_ || k == next_ || ready_.size() < cap_; });
        if (done_) return false;
        ready_.emplace(k, std::move(b));
        peak_ = std::max(peak_, ready_.size());
        cv_.notify_all();
        return true;
    }
    // Blocks until batch `seq` is there; false once finish() was called and it never will be.
    bool take(uint64_t seq, OutBatch& b)
    {
        std::unique_lock<std::mutex> lk(m_);
        next_ = seq;
        space_.notify_all();  // a producer holding exactly this batch may now enter even when the box is full
        cv_.wait(lk, [&] { return ready_.count(seq) || done_; });
        auto it = ready_.find(seq);
        if (it == ready_.end()) return false;
        b = std::move(it->second);
        ready_.erase(it);
        next_ = seq + 1;
        space_.notify_all();
        return true;
    }
    void finish()
    {
        std::lock_guard<std::mutex> lk(m_);
        done_ = true;
        cv_.notify_all();
        space_.notify_all();
    }
    size_t peak() { std::lock_guard<std::mutex> lk(m_); return peak_; }

private:
    std::mutex m_;
    std::condition_variable cv_, space_;
    std::map<uint64_t, OutBatch> ready_;
    uint64_t next_ = 0;  // the batch the writer takes next
    size_t cap_, peak_ = 0;
    bool done_ = false;
};

struct Shared {
    Options opt;
    BoundedQueue<RawBatch> raw;
    OrderedOutbox outbox;
    std::atomic<bool> failed{false};
    std::mutex err_m;
    std::string err;
    std::atomic<uint64_t> n_sites[3];
    std::atomic<uint64_t> n_reads{0}, n_bases{0}, n_batches{0};
    std::mutex hist_m;
    uint64_t ml_hist[3][256] = {};  // row N3: ML histograms per context over the whole run
    // phase clocks, seconds summed over threads (report only)
    std::atomic<uint64_t> us_read{0}, us_create{0}, us_pack{0}, us_submit{0}, us_collect{0}, us_assemble{0}, us_write{0};
    // timeline, microseconds since start: last engine ready, input exhausted, last batch collected, engines destroyed
    std::chrono::steady_clock::time_point t_start = std::chrono::steady_clock::now();
    std::atomic<uint64_t> at_ready{0}, at_input_done{0}, at_last_collect{0}, at_destroyed{0};
    uint64_t since_start() const { return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_start).count(); }
    std::atomic<int> producers{0};  // GPU workers that may still hand batches to the writer
    explicit Shared(const Options& o, size_t depth) : opt(o), raw(depth), outbox(depth) { for (auto& v : n_sites) v = 0; }
    // a worker has handed over its last batch (its engine is torn down AFTER this): the last one lets the writer finish
    void producer_done() { if (--producers == 0) outbox.finish(); }
    void fail(const std::string& msg)
    {
        {
            std::lock_guard<std::mutex> lk(err_m);
            if (err.empty()) err = msg;
        }
        failed = true;
        raw.abort();
        outbox.finish();
    }
};

struct Stopwatch {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    uint64_t lap_us()
    {
        const auto t1 = std::chrono::steady_clock::now();
        const uint64_t us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
        t0 = t1;
        return us;
    }
};

// Reader: inflates the input (block-parallel) and cuts it into batches by read count and by bases.
void reader_body(Shared& S, hm::BamReader& in)
{
    std::string err;
    RawBatch cur;
    uint64_t seq = 0;
    Stopwatch sw;
    auto flush = [&]() {
        if (cur.entries.empty()) return true;
        cur.seq = seq++;
        S.n_batches = seq;
        S.us_read += sw.lap_us();
        const bool ok = S.raw.push(std::move(cur));
        sw.lap_us();  // time blocked on a full queue is not reading time
        cur = RawBatch{};
        return ok;
    };
    const uint8_t* body;
    size_t len;
    while (!S.failed) {
        if (!in.next(body, len, err)) {
            if (!err.empty()) { S.fail(S.opt.in_path + ": " + err); return; }
            break;
        }
        uint32_t l = 0;
        memcpy(&l, body + 16, 4);  // l_seq (BamReader guarantees >= 32 bytes)
        if (l > 0x7fffffffu) l = 0;
        if (!cur.entries.empty() && (cur.entries.size() >= (size_t)S.opt.reads_per_batch || cur.bases + l > (uint64_t)S.opt.max_bases))
            if (!flush()) return;
        cur.entries.push_back(Entry{body, len});
        if (cur.slabs.empty() || cur.slabs.back() != in.slab()) cur.slabs.push_back(in.slab());
        if (l <= (uint64_t)S.opt.max_bases) cur.bases += l;  // longer records are passed through, they take no staging space
        S.n_reads++;
        S.n_bases += l;
    }
    flush();
    S.at_input_done = S.since_start();
    S.raw.close();
}

// One worker per entry of --devices: owns an engine with two staging slots.  While the GPU works on batch k the worker
// assembles the records of batch k-1 and packs batch k+1.
void gpu_worker_body(Shared& S, int device, int threads)
{
    const Options& opt = S.opt;
    struct Done {  // exactly one producer_done() per worker, whatever path leaves this function
        Shared& S;
        bool fired = false;
        void fire() { if (!fired) { fired = true; S.producer_done(); } }
        ~Done() { fire(); }
    } done{S};
    Stopwatch sw;
    hm_config cfg{};
    cfg.model_dir = opt.model_dir.c_str();
    cfg.ctx_mask = opt.ctx_mask;
    cfg.min_read_len = opt.min_read_len;
    cfg.device = device;
    cfg.n_slots = 2;
    cfg.max_reads = (uint32_t)opt.reads_per_batch;
    cfg.max_bases = (uint32_t)opt.max_bases;
    cfg.cnn_mode = HM_CNN_TENSOR;
    hm_engine* eng = nullptr;
    if (hm_engine_create(&cfg, &eng) != HM_OK) { S.fail(std::string("device ") + std::to_string(device) + ": " + hm_last_error(nullptr)); return; }
    struct EngineGuard {  // also destroys the engine when an exception unwinds this thread
        hm_engine*& e;
        ~EngineGuard() { if (e) hm_engine_destroy(e); e = nullptr; }
    } guard{eng};
    S.us_create += sw.lap_us();
    S.at_ready = S.since_start();

    struct InFlight {
        RawBatch raw;
        std::vector<int32_t> read_index;
        bool live = false;
    } fl[2];

    auto finish = [&](int slot) -> bool {
        InFlight& f = fl[slot];
        Stopwatch w;
        hm_call_batch calls{};
        if (hm_batch_collect(eng, slot, &calls) != HM_OK) { S.fail(hm_last_error(eng)); return false; }
        S.us_collect += w.lap_us();
        S.at_last_collect = S.since_start();
        for (int c = 0; c < 3; ++c) S.n_sites[c] += calls.n_sites[c];
        if (calls.ml_hist) {
            std::lock_guard<std::mutex> lk(S.hist_m);
            for (int c = 0; c < 3; ++c)
                for (int b = 0; b < 256; ++b) S.ml_hist[c][b] += calls.ml_hist[c * 256 + b];
        }
        const size_t n = f.raw.entries.size();
        OutBatch ob;
        ob.seq = f.raw.seq;
        ob.ooff.assign(n + 1, 0);
        ob.olen.assign(n, 0);
        for (size_t i = 0; i < n; ++i) {
            uint32_t nc = 0, text = 0;
            const int32_t r = f.read_index[i];
            if (r >= 0) {
                nc = calls.call_off[r + 1] - calls.call_off[r];
                text = calls.mm_off[r + 1] - calls.mm_off[r];
            }
            ob.ooff[i + 1] = ob.ooff[i] + 4 + f.raw.entries[i].len + 64 + nc + text;  // 4: the record's block_size field
        }
        ob.obuf.resize(ob.ooff.back());
        std::atomic<int> bad{-1};
        hm::parallel_for(n, threads, [&](size_t i) {
            const Entry& e = f.raw.entries[i];
            const uint8_t* body = e.body;
            uint8_t* dst = ob.obuf.data() + ob.ooff[i] + 4;
            const int32_t ri = f.read_index[i];
            int rc;
            if (ri < 0) {
                rc = hm_build_mod_record_mm(body, e.len, opt.keep_kinetics, nullptr, 0, nullptr, 0, nullptr, 0, 0, dst, &ob.olen[i]);
                if (rc == HM_ERR_FORMAT) {  // not parseable as a record: pass the bytes through untouched
                    memcpy(dst, body, e.len);
                    ob.olen[i] = e.len;
                    rc = 0;
                }
            } else {
                const uint32_t r = (uint32_t)ri, a = calls.call_off[r], nc = calls.call_off[r + 1] - a, nf = calls.n_fwd[r];
                const uint8_t* mm = calls.mm_text + calls.mm_off[r];
                const uint32_t fl_ = calls.mm_fwd_len[r], rl = calls.mm_off[r + 1] - calls.mm_off[r] - fl_;
                rc = hm_build_mod_record_mm(body, e.len, opt.keep_kinetics, mm, fl_, mm + fl_, rl, calls.ml + a, nf, nc - nf, dst, &ob.olen[i]);
            }
            if (rc != 0) bad = (int)i;
        });
        if (bad >= 0) { S.fail("cannot assemble output record " + std::to_string(bad.load()) + " of batch " + std::to_string(ob.seq)); return false; }
        // Close the gaps between the bounded slots, here on the worker: the one writer thread used to copy every record into the
        // BGZF writer's buffer (and that buffer once more into deflate chunks) -- two serial passes over the whole output, 5 GB at
        // 8 GPUs.  The batch now reaches the writer as one finished piece of the BAM stream (BamWriter::write_chunk).
        {
            size_t w_off = 0;
            uint8_t* base = ob.obuf.data();
            for (size_t i = 0; i < n; ++i) {
                const size_t len = ob.olen[i];
                const uint32_t l32 = (uint32_t)len;
                if (w_off != ob.ooff[i]) memmove(base + w_off + 4, base + ob.ooff[i] + 4, len);
                memcpy(base + w_off, &l32, 4);  // little-endian host (x86-64 / aarch64)
                w_off += 4 + len;
            }
            ob.obuf.resize(w_off);
            ob.ooff.clear();
            ob.olen.clear();
        }
        f.raw = RawBatch{};
        f.live = false;
        S.us_assemble += w.lap_us();
        return S.outbox.put(std::move(ob));
    };

    int cur = 0;
    bool ok = true;
    while (ok && !S.failed) {
        RawBatch rb;
        sw.lap_us();
        if (!S.raw.pop(rb)) break;
        sw.lap_us();
        InFlight& f = fl[cur];
        hm_read_batch hb{};
        if (hm_batch_acquire(eng, cur, &hb) != HM_OK) { S.fail(hm_last_error(eng)); ok = false; break; }
        const size_t n = rb.entries.size();
        std::vector<const uint8_t*> bodies(n);
        std::vector<size_t> lens(n);
        for (size_t i = 0; i < n; ++i) { bodies[i] = rb.entries[i].body; lens[i] = rb.entries[i].len; }
        f.read_index.assign(n, -1);
        uint32_t n_packed = 0;
        if (hm_pack_records(&hb, (uint32_t)n, bodies.data(), lens.data(), opt.min_read_len, threads, f.read_index.data(), &n_packed) != HM_OK) {
            S.fail("internal: a batch does not fit its staging slot");
            ok = false;
            break;
        }
        S.us_pack += sw.lap_us();
        if (hm_batch_submit(eng, cur, n_packed, HM_SUBMIT_MM_TEXT | HM_SUBMIT_ML_HIST) != HM_OK) { S.fail(hm_last_error(eng)); ok = false; break; }
        S.us_submit += sw.lap_us();
        f.raw = std::move(rb);
        f.live = true;
        if (fl[cur ^ 1].live && !finish(cur ^ 1)) { ok = false; break; }
        cur ^= 1;
    }
    // drain, older batch first: slot `cur` was finished inside the loop unless it ended early, slot cur ^ 1 holds the newest
    for (int k = 0; k < 2 && ok && !S.failed; ++k)
        if (fl[cur ^ k].live) ok = finish(cur ^ k);
    done.fire();  // the writer may close the output while this engine's pinned and device memory is being released
    hm_engine_destroy(eng);
    eng = nullptr;
    S.at_destroyed = S.since_start();
}

void writer_body(Shared& S, hm::BamWriter& out)
{
    std::string err;
    Stopwatch sw;
    for (uint64_t seq = 0;; ++seq) {
        OutBatch b;
        if (!S.outbox.take(seq, b)) break;
        sw.lap_us();
        if (!out.write_chunk(std::move(b.obuf), err)) { S.fail(S.opt.out_path + ": " + err); return; }
        S.us_write += sw.lap_us();
    }
}

// Thread entry points: an exception in any thread (bad_alloc from a buffer resize, a task of parallel_for) fails the run through
// Shared::fail -- which also wakes every other thread -- instead of reaching std::terminate.
template <class F>
void guarded(Shared& S, const char* role, F&& body)
{
    try {
        body();
    } catch (const std::exception& e) {
        S.fail(std::string(role) + ": " + e.what());
    } catch (...) {
        S.fail(std::string(role) + ": unknown exception");
    }
}
void reader_thread(Shared& S, hm::BamReader& in)
{
    guarded(S, "reader", [&] { reader_body(S, in); });
    S.raw.close();  // workers drain what is queued and stop, also when the reader ended early
}
void gpu_worker(Shared& S, int device, int threads)
{
    guarded(S, "GPU worker", [&] { gpu_worker_body(S, device, threads); });
}
void writer_thread(Shared& S, hm::BamWriter& out)
{
    guarded(S, "writer", [&] { writer_body(S, out); });
}

}  // namespace

static int call_main_impl(int argc, char** argv);
static std::atomic<bool> g_fast_exit{false};
// The CLI executable: do not wait for the engines' teardown once the output is closed; the caller must leave with _exit().
extern "C" void hm_call_fast_exit(int on) { g_fast_exit = on != 0; }

// Nothing throws across the C boundary: an allocation failure or any other exception ends the run with EXIT_FAILURE (the calling
// thread is guarded here, the reader / worker / writer threads by guarded() above, pool tasks by parallel_for itself).
extern "C" int hm_call_main(int argc, char** argv)
{
    try {
        return call_main_impl(argc, argv);
    } catch (const std::exception& e) {
        fprintf(stderr, "[hifimeth-b200] fatal: %s\n", e.what());
    } catch (...) {
        fprintf(stderr, "[hifimeth-b200] fatal: unknown exception\n");
    }
    return EXIT_FAILURE;
}

static int call_main_impl(int argc, char** argv)
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

    const int n_workers = (int)opt.devices.size();
    if (opt.max_bases <= 0) {
        // Batch size from the input size: BAM with kinetics is ~3.7 compressed bytes per base; aim at >= 3 batches per worker so
        // that reading, the GPU and writing overlap, within [2 Mi, 24 Mi] bases (sub-batches of the CNN stage hold ~1 Mi bases).
        struct stat sb;
        long long est = 24ll << 20;
        if (stat(opt.in_path.c_str(), &sb) == 0 && S_ISREG(sb.st_mode)) est = (long long)(sb.st_size / 3.7) / (3ll * n_workers);
        opt.max_bases = std::min<long long>(24ll << 20, std::max<long long>(2ll << 20, est));
    }
    Shared S(opt, (size_t)2 * n_workers);
    S.producers = n_workers;
    std::thread reader(reader_thread, std::ref(S), std::ref(in));
    std::vector<std::thread> workers;
    const int worker_threads = std::max(1, opt.threads / n_workers);
    for (int d : opt.devices) workers.emplace_back(gpu_worker, std::ref(S), d, worker_threads);
    std::thread writer(writer_thread, std::ref(S), std::ref(out));
    reader.join();
    writer.join();  // ends when the last worker has handed over its last batch and the writer has written it
    bool failed = S.failed;
    if (!failed && !out.close(err)) { S.err = opt.out_path + ": " + err; failed = true; }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    // The output is complete and closed here.  The workers are still releasing their engines (pinned host memory, ~25 GB of device
    // maps, the CUDA context: 0.2 - 1.5 s per device, measured); the executable skips that wait and leaves through _exit
    // (hm_call_fast_exit, main.cpp) -- the driver reclaims everything at process exit -- while in-process callers get a clean join.
    const bool fast = g_fast_exit && !failed;
    if (fast) for (auto& w : workers) w.detach();  // they still use S: this function does not return on that path (see the end)
    else for (auto& w : workers) w.join();
    if (failed) { fprintf(stderr, "[hifimeth-b200] %s\n", S.err.c_str()); return EXIT_FAILURE; }
    const uint64_t sites = S.n_sites[0] + S.n_sites[1] + S.n_sites[2];
    fprintf(stderr, "[hifimeth-b200] phases (s, summed per thread role): read+inflate %.2f, engine create %.2f, pack %.2f, submit %.2f, "
                    "collect(wait) %.2f, assemble %.2f, write+deflate %.2f; %llu batches of <= %lld bases on %d worker(s), %d host threads\n",
            S.us_read / 1e6, S.us_create / 1e6, S.us_pack / 1e6, S.us_submit / 1e6, S.us_collect / 1e6, S.us_assemble / 1e6, S.us_write / 1e6,
            (unsigned long long)S.n_batches.load(), opt.max_bases, n_workers, opt.threads);
    if (hm::bgzf_inflate_fallbacks())
        fprintf(stderr, "[hifimeth-b200] note: %llu BGZF block(s) were not accepted by the built-in inflater and were read by zlib instead (output unaffected)\n",
                (unsigned long long)hm::bgzf_inflate_fallbacks());
    fprintf(stderr, "[hifimeth-b200] timeline (s since start): engines ready %.2f, input inflated %.2f, last batch collected %.2f, engines "
                    "destroyed %.2f, output closed %.2f\n",
            S.at_ready / 1e6, S.at_input_done / 1e6, S.at_last_collect / 1e6, S.at_destroyed / 1e6, secs);
    fprintf(stderr, "[hifimeth-b200] %llu reads, %llu bases, CpG %llu, CHG %llu, CHH %llu samples in %.2f s (%.3g sites/s, %.3g reads/s)\n",
            (unsigned long long)S.n_reads.load(), (unsigned long long)S.n_bases.load(), (unsigned long long)S.n_sites[0].load(),
            (unsigned long long)S.n_sites[1].load(), (unsigned long long)S.n_sites[2].load(), secs, sites / secs, S.n_reads.load() / secs);
    // Row N3: the scaled-probability thresholds `hifimeth pileup` would infer from this output (pileup.cpp:355-436), from the
    // histograms the device kept while the ML bytes were resident; --hist FILE also writes the bins.
    static const char* ctx_name[3] = {"CpG", "CHG", "CHH"};
    for (int c = 0; c < 3; ++c) {
        if (!(opt.ctx_mask & (1 << c))) continue;
        uint64_t n = 0;
        const unsigned t = hm_ml_threshold(S.ml_hist[c], &n);
        fprintf(stderr, "[hifimeth-b200] %s scaled probability threshold (pileup rule): %u (%llu samples in range)\n", ctx_name[c], t, (unsigned long long)n);
    }
    if (!opt.hist_path.empty()) {
        FILE* hf = fopen(opt.hist_path.c_str(), "w");
        if (!hf) { fprintf(stderr, "[hifimeth-b200] cannot create %s\n", opt.hist_path.c_str()); return EXIT_FAILURE; }
        fprintf(hf, "scaled_prob\tCpG\tCHG\tCHH\n");
        for (int b = 0; b < 256; ++b)
            fprintf(hf, "%d\t%llu\t%llu\t%llu\n", b, (unsigned long long)S.ml_hist[0][b], (unsigned long long)S.ml_hist[1][b], (unsigned long long)S.ml_hist[2][b]);
        fclose(hf);
    }
    if (fast) {
        fflush(nullptr);
        _exit(EXIT_SUCCESS);
    }
    return EXIT_SUCCESS;
}

// Round trip of the BAM codec alone (tests): read every record of `in_path`, write it unchanged to `out_path`.
extern "C" size_t hm_deflate_block(const uint8_t* in, size_t n, uint8_t* out, size_t cap)
{
    if ((!in && n) || !out) return 0;
    return hm::hm_deflate_rle(in, n, out, cap);
}

extern "C" int hm_inflate_block(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out)
{
    if (!in || (!out && n_out)) return 0;
    return hm::hm_inflate_fast(in, n_in, out, n_out) ? 1 : 0;
}

extern "C" uint32_t hm_crc32_bytes(uint32_t crc, const uint8_t* data, size_t n) { return (!data && n) ? crc : hm::hm_crc32(crc, data, n); }

extern "C" int hm_bam_copy(const char* in_path, const char* out_path, int threads, int level)
{
    try {
    std::string err;
    hm::BamReader in;
    hm::BamHeader hdr;
    if (!in.open(in_path, threads, hdr, err)) return HM_ERR_FORMAT;
    const uint8_t* body;
    size_t len;
    long long n = 0;
    if (level < 0 || !out_path) {  // reader only: inflate, frame, count
        while (in.next(body, len, err)) ++n;
        return err.empty() ? (int)std::min<long long>(n, 0x7fffffff) : HM_ERR_FORMAT;
    }
    hm::BamWriter out;
    const bool pieces = level >= 100;  // 100 + level: hand the records over in finished pieces, as `call` does (BamWriter::write_chunk)
    if (pieces) level -= 100;
    if (!out.open(out_path, threads, level, hdr, err)) return HM_ERR_ARG;
    hm::Bytes piece;
    while (in.next(body, len, err)) {
        if (pieces) {
            const uint32_t l32 = (uint32_t)len;
            const size_t at = piece.size();
            piece.resize(at + 4 + len);
            memcpy(piece.data() + at, &l32, 4);
            memcpy(piece.data() + at + 4, body, len);
            if (piece.size() >= (n % 3 ? (size_t)32 << 20 : (size_t)100000)) {  // large and small pieces alternate
                if (!out.write_chunk(std::move(piece), err)) return HM_ERR_ARG;
                piece = hm::Bytes();
            }
        } else if (!out.write_record(body, len, err)) return HM_ERR_ARG;
        ++n;
    }
    if (pieces && !piece.empty() && !out.write_chunk(std::move(piece), err)) return HM_ERR_ARG;
    if (!err.empty()) return HM_ERR_FORMAT;
    if (!out.close(err)) return HM_ERR_ARG;
    return (int)std::min<long long>(n, 0x7fffffff);
    } catch (...) {
        return HM_ERR_FORMAT;
    }
}
