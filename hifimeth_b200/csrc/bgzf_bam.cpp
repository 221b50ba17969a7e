// bgzf_bam.cpp -- see bgzf_bam.h.  Block-parallel inflate / deflate with zlib; nothing here touches the GPU.
#include <functional>

#include "bgzf_bam.h"
#include "fast_deflate.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <exception>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

namespace hm {

namespace {
constexpr size_t kBlockPayload = 0xff00;      // uncompressed bytes per BGZF block (htslib's BGZF_BLOCK_SIZE)
constexpr size_t kReadSlab = 16u << 20;       // compressed bytes read per slab (~260 blocks: enough to spread over the threads)
constexpr uint32_t kMaxRecord = 1u << 29;    // a block_size above 512 MiB is corruption, not a record
constexpr size_t kReadAhead = 3;              // inflated slabs queued ahead of the consumer
constexpr size_t kWriteBatch = 1024;          // blocks deflated per parallel batch (~64 MB)

uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
void wr16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
void wr32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

// Total size of the BGZF block starting at p (n bytes available), or 0 if the header is incomplete / not BGZF.
size_t bgzf_block_size(const uint8_t* p, size_t n, bool& bad)
{
    bad = false;
    if (n < 18) return 0;
    if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) { bad = true; return 0; }
    const size_t xlen = rd16(p + 10);
    if (n < 12 + xlen) return 0;
    size_t off = 12;
    while (off + 4 <= 12 + xlen) {
        const uint16_t slen = rd16(p + off + 2);
        if (p[off] == 'B' && p[off + 1] == 'C' && slen == 2) return (size_t)rd16(p + off + 4) + 1;
        off += 4 + slen;
    }
    bad = true;
    return 0;
}
}  // namespace

namespace {
// One process-wide pool of worker threads shared by every parallel_for (reader inflate, record packing and assembly, writer
// deflate): spawning threads per call cost ~1 ms per 8 MB slab.  Callers are never pool threads themselves, so a caller may
// block on its own job while the pool serves several jobs at once.
class Pool {
public:
    static Pool& get()
    {
        static Pool p;
        return p;
    }
    void run(size_t n, size_t workers, const std::function<void(size_t)>& fn)
    {
        struct Job {
            std::atomic<size_t> next{0}, left{0};
            std::mutex m;
            std::condition_variable cv;
            std::exception_ptr error;  // first exception of any task: rethrown in the caller, never left to std::terminate a pool thread
        } job;
        job.left = workers;
        auto body = [&job, &fn, n] {
            try {
                for (size_t i = job.next.fetch_add(1); i < n; i = job.next.fetch_add(1)) fn(i);
            } catch (...) {
                job.next = n;  // the other workers stop taking tasks
                std::lock_guard<std::mutex> lk(job.m);
                if (!job.error) job.error = std::current_exception();
            }
            std::lock_guard<std::mutex> lk(job.m);
            if (--job.left == 0) job.cv.notify_all();
        };
        {
            std::lock_guard<std::mutex> lk(m_);
            for (size_t k = 0; k + 1 < workers; ++k) q_.push_back(body);
            cv_.notify_all();
        }
        body();  // the caller works too
        std::unique_lock<std::mutex> lk(job.m);
        job.cv.wait(lk, [&] { return job.left == 0; });
        if (job.error) std::rethrow_exception(job.error);
    }
    size_t size() const { return threads_.size() + 1; }

private:
    Pool()
    {
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        for (unsigned k = 0; k + 1 < hw; ++k)
            threads_.emplace_back([this] {
                for (;;) {
                    std::function<void()> task;
                    {
                        std::unique_lock<std::mutex> lk(m_);
                        cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                        if (q_.empty()) return;
                        task = std::move(q_.front());
                        q_.pop_front();
                    }
                    task();
                }
            });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            cv_.notify_all();
        }
        for (auto& t : threads_) t.join();
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> threads_;
    bool stop_ = false;
};
}  // namespace

void parallel_for(size_t n, int threads, const std::function<void(size_t)>& fn)
{
    if (n == 0) return;
    const size_t t = std::min<size_t>(std::min<size_t>(std::max(threads, 1), n), Pool::get().size());
    if (t == 1) {
        for (size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    Pool::get().run(n, t, fn);
}

// ---- BgzfReader ----------------------------------------------------------------------------------------------------------
namespace {
std::atomic<uint64_t> g_inflate_fallbacks{0};
// HM_ZLIB_ONLY=1: every block through zlib (A/B of the codecs in fast_deflate.h; tools/bam_copy_bench.py)
bool use_zlib_only()
{
    static const bool v = [] { const char* e = getenv("HM_ZLIB_ONLY"); return e && *e && *e != '0'; }();
    return v;
}
constexpr size_t kRawAhead = 2;     // compressed slabs queued between the I/O thread and the drivers
constexpr int kInflateDrivers = 2;  // slabs being inflated at the same time

// One inflate stream per thread for the life of the thread: inflateInit2 / inflateEnd per 64 KiB block was an allocation and a
// reset of the stream state for every block.
struct TlInflate {
    z_stream zs{};
    bool init = false;
    ~TlInflate() { if (init) inflateEnd(&zs); }
};
}  // namespace

uint64_t bgzf_inflate_fallbacks() { return g_inflate_fallbacks.load(std::memory_order_relaxed); }

BgzfReader::~BgzfReader() { close(); }
void BgzfReader::close()
{
    {
        std::lock_guard<std::mutex> lk(m_);
        stop_ = true;
        cv_.notify_all();
    }
    if (io_.joinable()) io_.join();
    for (auto& d : drivers_)
        if (d.joinable()) d.join();
    drivers_.clear();
    if (map_) munmap(const_cast<uint8_t*>(map_), map_size_);
    map_ = nullptr;
    map_size_ = map_pos_ = 0;
    if (f_) fclose(f_);
    f_ = nullptr;
    raw_q_.clear();
    done_.clear();
}

bool BgzfReader::open(const char* path, int threads, std::string& err)
{
    close();
    f_ = fopen(path, "rb");
    if (!f_) { err = std::string("cannot open ") + path; return false; }
    threads_ = std::max(threads, 1);
    slab_bytes_ = kReadSlab;
    if (const char* e = getenv("HM_BGZF_SLAB")) slab_bytes_ = std::max<size_t>(1024, (size_t)atoll(e));  // tests: force records to straddle slabs
    carry_.clear();
    // A regular file is mapped instead of read: fread() is one more copy of every compressed byte on ONE thread -- 14 GB for the
    // 8-GPU bench input, seconds of a run whose GPUs need 2.7 s.  Pipes and HM_NO_MMAP=1 keep the fread path.
    {
        struct stat sb;
        const char* no = getenv("HM_NO_MMAP");
        if (!(no && *no && *no != '0') && fstat(fileno(f_), &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
            void* m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fileno(f_), 0);
            if (m != MAP_FAILED) {
                map_ = static_cast<const uint8_t*>(m);
                map_size_ = (size_t)sb.st_size;
                map_pos_ = 0;
                madvise(m, map_size_, MADV_SEQUENTIAL);
            }
        }
    }
    eof_ = io_done_ = stop_ = false;
    next_seq_ = n_slabs_ = 0;
    err_.clear();
    io_ = std::thread(&BgzfReader::io_loop, this);
    for (int k = 0; k < kInflateDrivers; ++k) drivers_.emplace_back(&BgzfReader::driver_loop, this);
    return true;
}

void BgzfReader::fail(const std::string& err)
{
    std::lock_guard<std::mutex> lk(m_);
    if (err_.empty()) err_ = err.empty() ? std::string("BGZF reader: unknown error") : err;
    stop_ = true;
    cv_.notify_all();
}

// Reads until the buffer holds at least one complete block, then cuts it at the last block boundary (the tail is carried over).
bool BgzfReader::read_raw(RawSlab& rs, std::string& err)
{
    for (;;) {
        rs.blks.clear();
        rs.total = 0;
        size_t off = 0;
        Bytes& buf = carry_;
        while (off < buf.size()) {
            bool bad;
            const size_t bs = bgzf_block_size(buf.data() + off, buf.size() - off, bad);
            if (bad) { err = "not a BGZF block (is the input a BAM file?)"; return false; }
            if (!bs || off + bs > buf.size()) break;
            // header (12) + extra field + at least the empty deflate stream (2) + CRC32 / ISIZE (8): XLEN comes from the file, and a
            // block shorter than that would make the inflate length below wrap around
            if (bs < 12 + (size_t)rd16(buf.data() + off + 10) + 2 + 8) { err = "corrupt BGZF block (BSIZE smaller than its own header and trailer)"; return false; }
            const size_t isize = rd32(buf.data() + off + bs - 4);
            if (isize > 65536) { err = "corrupt BGZF block (ISIZE above 64 KiB)"; return false; }  // BGZF payloads are <= 64 KiB
            rs.blks.push_back({off, bs, isize, rs.total});
            rs.total += isize;
            off += bs;
        }
        if (!rs.blks.empty()) {
            Bytes tail(buf.begin() + off, buf.end());  // < one block
            buf.resize(off);
            rs.raw.swap(buf);
            carry_.swap(tail);
            if (rs.total) return true;
            continue;  // only empty blocks (EOF markers): look for more
        }
        if (eof_) {
            if (!buf.empty()) err = "truncated BGZF block at end of file";
            return false;
        }
        const size_t have = buf.size();
        buf.resize(have + slab_bytes_);
        const size_t got = fread(buf.data() + have, 1, slab_bytes_, f_);
        buf.resize(have + got);
        if (got < slab_bytes_) eof_ = true;
    }
}

// The mapped form of read_raw: a slab is a run of whole blocks of the mapping, nothing is copied.
bool BgzfReader::read_raw_mapped(RawSlab& rs, std::string& err)
{
    for (;;) {
        rs.blks.clear();
        rs.total = 0;
        if (map_pos_ >= map_size_) return false;
        const uint8_t* base = map_ + map_pos_;
        const size_t avail = map_size_ - map_pos_;
        size_t off = 0;
        while (off < avail && off < slab_bytes_) {
            bool bad;
            const size_t bs = bgzf_block_size(base + off, avail - off, bad);
            if (bad) { err = "not a BGZF block (is the input a BAM file?)"; return false; }
            if (!bs || off + bs > avail) { err = "truncated BGZF block at end of file"; return false; }
            if (bs < 12 + (size_t)rd16(base + off + 10) + 2 + 8) { err = "corrupt BGZF block (BSIZE smaller than its own header and trailer)"; return false; }
            const size_t isize = rd32(base + off + bs - 4);
            if (isize > 65536) { err = "corrupt BGZF block (ISIZE above 64 KiB)"; return false; }
            rs.blks.push_back({off, bs, isize, rs.total});
            rs.total += isize;
            off += bs;
        }
        rs.base = base;
        map_pos_ += off;
        if (rs.total) return true;  // else only empty blocks (EOF markers): look for more
    }
}

void BgzfReader::io_loop()
{
    uint64_t seq = 0;
    try {  // an allocation failure ends the stream with an error instead of terminating the process
        for (;;) {
            RawSlab rs;
            std::string err;
            if (!(map_ ? read_raw_mapped(rs, err) : read_raw(rs, err))) {
                if (!err.empty()) { fail(err); return; }
                break;
            }
            rs.seq = seq++;
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return raw_q_.size() < kRawAhead || stop_; });
            if (stop_) return;
            raw_q_.push_back(std::move(rs));
            cv_.notify_all();
        }
    } catch (const std::exception& e) {
        fail(std::string("BGZF reader: ") + e.what());
        return;
    } catch (...) {
        fail("BGZF reader: unknown exception");
        return;
    }
    std::lock_guard<std::mutex> lk(m_);
    n_slabs_ = seq;
    io_done_ = true;
    cv_.notify_all();
}

void BgzfReader::driver_loop()
{
    try {
        for (;;) {
            RawSlab rs;
            {
                std::unique_lock<std::mutex> lk(m_);
                // take the next compressed slab, but do not run more than kReadAhead slabs ahead of the consumer
                cv_.wait(lk, [&] { return stop_ || (!raw_q_.empty() && raw_q_.front().seq < next_seq_ + kReadAhead) || (raw_q_.empty() && io_done_); });
                if (stop_ || raw_q_.empty()) return;
                rs = std::move(raw_q_.front());
                raw_q_.pop_front();
                cv_.notify_all();
            }
            auto slab = std::make_shared<Slab>();
            slab->data.resize(rs.total);
            std::atomic<bool> ok{true};
            uint8_t* out = slab->data.data();
            parallel_for(rs.blks.size(), threads_, [&](size_t i) {
                const Blk& b = rs.blks[i];
                if (!b.isize) return;
                {
                    // own inflater first (fast_deflate.h); whatever it does not accept goes to zlib below, which has the last word
                    const uint8_t* q = (rs.base ? rs.base : rs.raw.data()) + b.off;
                    const size_t xl = rd16(q + 10);
                    if (!use_zlib_only() && hm_inflate_fast(q + 12 + xl, b.size - 12 - xl - 8, out + b.dst, b.isize)) {
                        if (hm_crc32(0, out + b.dst, b.isize) != rd32(q + b.size - 8)) ok = false;
                        return;
                    }
                }
                static thread_local TlInflate tl;
                if (!tl.init) {
                    if (inflateInit2(&tl.zs, -15) != Z_OK) { ok = false; return; }
                    tl.init = true;
                } else if (inflateReset(&tl.zs) != Z_OK) { ok = false; return; }
                const uint8_t* p = (rs.base ? rs.base : rs.raw.data()) + b.off;
                const size_t xlen = rd16(p + 10);
                z_stream& zs = tl.zs;
                zs.next_in = const_cast<Bytef*>(p + 12 + xlen);
                zs.avail_in = (uInt)(b.size - 12 - xlen - 8);
                zs.next_out = out + b.dst;
                zs.avail_out = (uInt)b.isize;
                const int rc = inflate(&zs, Z_FINISH);
                if (rc != Z_STREAM_END || zs.total_out != b.isize ||
                    hm_crc32(0, out + b.dst, b.isize) != rd32(p + b.size - 8))
                    ok = false;
                else if (!use_zlib_only()) g_inflate_fallbacks.fetch_add(1, std::memory_order_relaxed);
            });
            if (!ok) { fail("BGZF block failed to inflate or its CRC does not match"); return; }
            std::lock_guard<std::mutex> lk(m_);
            done_.emplace_back(rs.seq, std::move(slab));
            cv_.notify_all();
        }
    } catch (const std::exception& e) {
        fail(std::string("BGZF reader: ") + e.what());
    } catch (...) {
        fail("BGZF reader: unknown exception");
    }
}

std::shared_ptr<Slab> BgzfReader::next_slab(std::string& err)
{
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
        for (auto it = done_.begin(); it != done_.end(); ++it) {
            if (it->first != next_seq_) continue;
            std::shared_ptr<Slab> s = std::move(it->second);
            done_.erase(it);
            ++next_seq_;
            cv_.notify_all();
            return s;
        }
        if (!err_.empty()) { err = err_; return nullptr; }
        if (stop_ || (io_done_ && next_seq_ == n_slabs_)) return nullptr;
        cv_.wait(lk);
    }
}

// ---- BgzfWriter ----------------------------------------------------------------------------------------------------------
namespace {
struct TlDeflate {
    z_stream zs{};
    bool init = false;
    int level = -2;
    ~TlDeflate() { if (init) deflateEnd(&zs); }
};
}  // namespace

BgzfWriter::~BgzfWriter()
{
    {
        std::lock_guard<std::mutex> lk(m_);
        end_ = true;
        q_.clear();
        cv_.notify_all();
    }
    if (bg_.joinable()) bg_.join();
    {
        std::lock_guard<std::mutex> lk(m_);
        io_end_ = true;
        wq_.clear();
        cv_.notify_all();
    }
    if (io_.joinable()) io_.join();
    if (f_) fclose(f_);
}

bool BgzfWriter::open(const char* path, int threads, int level, std::string& err)
{
    f_ = fopen(path, "wb");
    if (!f_) { err = std::string("cannot create ") + path; return false; }
    threads_ = std::max(threads, 1);
    level_ = level;
    pending_.clear();
    q_.clear();
    bg_err_.clear();
    wq_.clear();
    end_ = io_end_ = false;
    bg_ = std::thread(&BgzfWriter::bg_loop, this);
    io_ = std::thread(&BgzfWriter::io_loop, this);
    return true;
}

// The caller only appends; whole chunks of kWriteBatch blocks go to a background thread that deflates them (blocks spread over the
// shared pool) and writes them out, so that record assembly, the copy into `pending_`, deflate and the file write overlap.  With
// deflate in the caller the one writer thread was busy 2.3 of 2.5 s of a 4-GPU run and set its pace.
bool BgzfWriter::write(const void* data, size_t n, std::string& err)
{
    const uint8_t* p = static_cast<const uint8_t*>(data);
    pending_.insert(pending_.end(), p, p + n);
    if (pending_.size() >= kWriteBatch * kBlockPayload) return hand_over(false, err);
    return true;
}

bool BgzfWriter::write_owned(Bytes&& piece, std::string& err)
{
    if (!hand_over(true, err)) return false;  // what write() collected so far goes first
    if (piece.empty()) return true;
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return q_.size() < 2 || !bg_err_.empty(); });
    if (!bg_err_.empty()) { err = bg_err_; return false; }
    q_.push_back(std::move(piece));
    cv_.notify_all();
    return true;
}

bool BgzfWriter::hand_over(bool all, std::string& err)
{
    const size_t nblk = pending_.size() / kBlockPayload;
    const size_t take = all ? pending_.size() : nblk * kBlockPayload;
    if (take) {
        Bytes chunk;
        if (take == pending_.size()) chunk.swap(pending_);
        else {
            chunk.assign(pending_.begin(), pending_.begin() + take);
            pending_.erase(pending_.begin(), pending_.begin() + take);
        }
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return q_.size() < 2 || !bg_err_.empty(); });
        if (bg_err_.empty()) {
            q_.push_back(std::move(chunk));
            cv_.notify_all();
        }
    }
    std::lock_guard<std::mutex> lk(m_);
    if (!bg_err_.empty()) { err = bg_err_; return false; }
    return true;
}

void BgzfWriter::bg_loop()
{
    for (;;) {
        Bytes chunk;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return !q_.empty() || end_; });
            if (q_.empty()) return;
            chunk = std::move(q_.front());
            q_.pop_front();
            cv_.notify_all();
        }
        std::string err;
        bool ok = false;
        try {
            ok = deflate_chunk(chunk, err);
        } catch (const std::exception& e) {
            err = std::string("BGZF writer: ") + e.what();
        } catch (...) {
            err = "BGZF writer: unknown exception";
        }
        if (!ok) {
            std::lock_guard<std::mutex> lk(m_);
            bg_err_ = err.empty() ? std::string("deflate failed") : err;
            q_.clear();
            cv_.notify_all();
            return;
        }
    }
}

bool BgzfWriter::deflate_chunk(const Bytes& in, std::string& err)
{
    const size_t nblk = (in.size() + kBlockPayload - 1) / kBlockPayload;
    if (!nblk) return true;
    std::vector<Bytes> comp(nblk);
    std::atomic<bool> ok{true};
    const int level = level_;
    parallel_for(nblk, threads_, [&](size_t i) {
        const size_t off = i * kBlockPayload;
        const size_t len = std::min(kBlockPayload, in.size() - off);
        Bytes& c = comp[i];
        c.resize(18 + std::max<size_t>(compressBound((uLong)len), hm_deflate_rle_bound(len)) + 8);
        size_t clen = 0;
        if (level == 1 && !use_zlib_only()) {
            // level 1 = run-length + Huffman, one dynamic block per BGZF block (fast_deflate.h): HiFi records are kinetics codes,
            // packed bases and qualities -- noisy bytes in which zlib's level-1 matcher finds only short, poor matches.  On such
            // records this is several times faster than zlib level 1 and its output is smaller than level 6's
            clen = hm_deflate_rle(in.data() + off, len, c.data() + 18, c.size() - 18 - 8);
            if (!clen) { ok = false; return; }
        } else {
        // one deflate stream per thread and level for the life of the thread: deflateInit2 allocates and clears ~260 KB per call
        static thread_local TlDeflate tl;
        if (!tl.init || tl.level != level) {
            if (tl.init) deflateEnd(&tl.zs);
            tl.zs = z_stream{};
            tl.init = false;
            // (HM_ZLIB_ONLY=1 and level 1: zlib's own run-length strategy, the closest zlib has to hm_deflate_rle)
            if (deflateInit2(&tl.zs, level, Z_DEFLATED, -15, 8, level == 1 ? Z_RLE : Z_DEFAULT_STRATEGY) != Z_OK) { ok = false; return; }
            tl.init = true;
            tl.level = level;
        } else if (deflateReset(&tl.zs) != Z_OK) { ok = false; return; }
        z_stream& zs = tl.zs;
        zs.next_in = const_cast<Bytef*>(in.data() + off);
        zs.avail_in = (uInt)len;
        zs.next_out = c.data() + 18;
        zs.avail_out = (uInt)(c.size() - 18 - 8);
        const int rc = deflate(&zs, Z_FINISH);
        clen = zs.total_out;
        if (rc != Z_STREAM_END) { ok = false; return; }
        }
        if (18 + clen + 8 > 65536) { ok = false; return; }
        static const uint8_t hdr[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
        memcpy(c.data(), hdr, 12);
        c[12] = 'B'; c[13] = 'C';
        wr16(c.data() + 14, 2);
        wr16(c.data() + 16, (uint32_t)(18 + clen + 8 - 1));
        wr32(c.data() + 18 + clen, hm_crc32(0, in.data() + off, len));
        wr32(c.data() + 18 + clen + 4, (uint32_t)len);
        c.resize(18 + clen + 8);
    });
    if (!ok) { err = "deflate failed"; return false; }
    // to the file-write stage (at most two chunks ahead of it)
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return wq_.size() < 2 || !bg_err_.empty(); });
    if (!bg_err_.empty()) { err = bg_err_; return false; }
    wq_.push_back(std::move(comp));
    cv_.notify_all();
    return true;
}

void BgzfWriter::io_loop()
{
    for (;;) {
        std::vector<Bytes> comp;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return !wq_.empty() || io_end_; });
            if (wq_.empty()) return;
            comp = std::move(wq_.front());
            wq_.pop_front();
            cv_.notify_all();
        }
        for (auto& c : comp)
            if (fwrite(c.data(), 1, c.size(), f_) != c.size()) {
                std::lock_guard<std::mutex> lk(m_);
                if (bg_err_.empty()) bg_err_ = "write error";
                q_.clear();
                wq_.clear();
                cv_.notify_all();
                return;
            }
    }
}

bool BgzfWriter::close(std::string& err)
{
    if (!f_) return true;
    bool ok = hand_over(true, err);
    {
        std::lock_guard<std::mutex> lk(m_);
        end_ = true;
        cv_.notify_all();
    }
    if (bg_.joinable()) bg_.join();
    {
        std::lock_guard<std::mutex> lk(m_);
        io_end_ = true;
        cv_.notify_all();
    }
    if (io_.joinable()) io_.join();
    if (ok && !bg_err_.empty()) { err = bg_err_; ok = false; }
    static const uint8_t eof_marker[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (ok && fwrite(eof_marker, 1, 28, f_) != 28) { err = "write error"; ok = false; }
    if (fclose(f_) != 0 && ok) { err = "close error"; ok = false; }
    f_ = nullptr;
    return ok;
}

// ---- BamReader / BamWriter -------------------------------------------------------------------------------------------------
bool BamReader::advance(std::string& err)
{
    std::shared_ptr<Slab> nx = z_.next_slab(err);
    if (!nx) return false;
    cur_ = std::move(nx);
    pos_ = 0;
    return true;
}

bool BamReader::read_bytes(uint8_t* dst, size_t n, std::string& err)
{
    while (n) {
        if (!cur_ || pos_ == cur_->data.size()) {
            if (!advance(err)) return false;
            continue;
        }
        const size_t k = std::min(n, cur_->data.size() - pos_);
        memcpy(dst, cur_->data.data() + pos_, k);
        dst += k;
        pos_ += k;
        n -= k;
    }
    return true;
}

bool BamReader::open(const char* path, int threads, BamHeader& hdr, std::string& err)
{
    if (!z_.open(path, threads, err)) return false;
    cur_.reset();
    pos_ = 0;
    uint8_t h[8];
    if (!read_bytes(h, 8, err)) { if (err.empty()) err = "empty BAM file"; return false; }
    if (memcmp(h, "BAM\1", 4) != 0) { err = "bad BAM magic"; return false; }
    const uint32_t l_text = rd32(h + 4);
    if (l_text > kMaxRecord) { err = "corrupt BAM header"; return false; }
    hdr.text.assign(l_text, '\0');
    uint8_t nr[4];
    if (!read_bytes(reinterpret_cast<uint8_t*>(&hdr.text[0]), l_text, err) || !read_bytes(nr, 4, err)) { if (err.empty()) err = "truncated BAM header"; return false; }
    while (!hdr.text.empty() && hdr.text.back() == '\0') hdr.text.pop_back();
    const uint32_t n_ref = rd32(nr);
    hdr.refs.assign(nr, nr + 4);
    for (uint32_t i = 0; i < n_ref; ++i) {
        uint8_t ln[4];
        if (!read_bytes(ln, 4, err)) { if (err.empty()) err = "truncated BAM references"; return false; }
        const uint32_t l_name = rd32(ln);
        if (l_name > 65536 || hdr.refs.size() > kMaxRecord) { err = "corrupt BAM references"; return false; }
        const size_t at = hdr.refs.size();
        hdr.refs.resize(at + 4 + (size_t)l_name + 4);
        memcpy(hdr.refs.data() + at, ln, 4);
        if (!read_bytes(hdr.refs.data() + at + 4, (size_t)l_name + 4, err)) { if (err.empty()) err = "truncated BAM references"; return false; }
    }
    return true;
}

bool BamReader::next(const uint8_t*& body, size_t& len, std::string& err)
{
    err.clear();
    for (;;) {
        if (!cur_ || pos_ == cur_->data.size()) {  // a clean end of file lands here: advance fails with err empty
            if (!advance(err)) return false;
            continue;
        }
        const size_t avail = cur_->data.size() - pos_;
        if (avail >= 4) {
            const uint32_t bs = rd32(cur_->data.data() + pos_);
            if (bs < 32 || bs > kMaxRecord) { err = "corrupt BAM record"; return false; }
            if (avail >= 4 + (size_t)bs) {  // fast path: the record lies inside the current slab
                body = cur_->data.data() + pos_ + 4;
                len = bs;
                pos_ += 4 + (size_t)bs;
                return true;
            }
        }
        break;
    }
    // the record (or its length word) straddles slabs: gather it into a buffer owned by the slab it ends in
    uint8_t lw[4];
    if (!read_bytes(lw, 4, err)) { if (err.empty()) err = "truncated BAM record"; return false; }
    const uint32_t bs = rd32(lw);
    if (bs < 32 || bs > kMaxRecord) { err = "corrupt BAM record"; return false; }
    std::vector<uint8_t> tmp(bs);
    if (!read_bytes(tmp.data(), bs, err)) { if (err.empty()) err = "truncated BAM record"; return false; }
    cur_->extra.push_back(std::move(tmp));
    body = cur_->extra.back().data();
    len = bs;
    return true;
}

bool BamWriter::open(const char* path, int threads, int level, const BamHeader& hdr, std::string& err)
{
    if (!z_.open(path, threads, level, err)) return false;
    std::vector<uint8_t> h(8);
    memcpy(h.data(), "BAM\1", 4);
    wr32(h.data() + 4, (uint32_t)hdr.text.size());
    h.insert(h.end(), hdr.text.begin(), hdr.text.end());
    h.insert(h.end(), hdr.refs.begin(), hdr.refs.end());
    if (hdr.refs.empty()) { uint8_t z[4] = {0, 0, 0, 0}; h.insert(h.end(), z, z + 4); }
    return z_.write(h.data(), h.size(), err);
}

bool BamWriter::write_record(const uint8_t* body, size_t len, std::string& err)
{
    uint8_t bs[4];
    wr32(bs, (uint32_t)len);
    return z_.write(bs, 4, err) && z_.write(body, len, err);
}

bool BamWriter::close(std::string& err) { return z_.close(err); }

}  // namespace hm
