// bgzf_bam.h -- minimal block-parallel BGZF + BAM record reader / writer over zlib.
//
// Replaces the reference's use of htslib for the `call` path: sam_open / hts_set_threads(8) / sam_hdr_read / sam_read1
// (src/corelib/sam_batch.hpp:12-54) and sam_open("wb") / sam_hdr_write / sam_write1 (src/app/hifimeth/mod_main.cpp:316-321,
// 353-362).  Written from the SAM/BAM specification (SAMv1 section 4: BGZF blocks are gzip members with a BC extra field
// holding BSIZE; a BAM file is magic, l_text, text, n_ref, references, then block_size-prefixed records).  Only what `call`
// needs: sequential reading of records, header pass-through with one @PG line appended, sequential writing.
#pragma once
#include <cstdint>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace hm {

// Runs fn(i) for i in [0, n) on up to `threads` threads (inline when n or threads is small).
void parallel_for(size_t n, int threads, const std::function<void(size_t)>& fn);

// Byte buffers that are resized and then overwritten completely (inflated slabs, compressed input, assembled output): a plain
// std::vector zero-fills on resize, which at 16 - 150 MB per buffer was a serial pass of its own in the reader and worker threads.
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
    template <class U, class... A> void construct(U* p, A&&... a)
    {
        if constexpr (sizeof...(A) == 0) ::new (static_cast<void*>(p)) U;  // default-initialised: no fill
        else ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
};
using Bytes = std::vector<uint8_t, NoInitAlloc<uint8_t>>;

// BGZF blocks the built-in inflater (fast_deflate.h) did not accept and zlib then read correctly, since the start of the process.
// Zero on well-formed input; anything else is a defect of the built-in inflater worth reporting (the data is still right).
uint64_t bgzf_inflate_fallbacks();

// One inflated stretch of the input.  Record bodies handed out by BamReader point into `data`, or -- for a record that
// straddles two slabs -- into one of the `extra` buffers of the slab it ends in; both stay valid while the slab is alive.
struct Slab {
    Bytes data;
    std::deque<std::vector<uint8_t>> extra;
};

// Sequential BGZF inflater with read-ahead, three stages deep: an I/O thread reads compressed slabs and cuts them at block
// boundaries; two driver threads inflate a slab each (its blocks spread over the shared thread pool, one z_stream per pool thread
// re-used with inflateReset); the consumer takes the inflated slabs in file order.  Reading, the serial parts of one slab and the
// inflation of the next overlap -- one thread doing all three in turn delivered 1.6 GB/s of inflated bytes on 32 cores, enough
// for two B200s, not for four.
class BgzfReader {
public:
    ~BgzfReader();
    bool open(const char* path, int threads, std::string& err);
    // Next inflated slab (never empty); nullptr at end of file (err empty) or on error (err set).
    std::shared_ptr<Slab> next_slab(std::string& err);
    void close();

private:
    struct Blk { size_t off, size, isize, dst; };
    struct RawSlab {
        const uint8_t* base = nullptr;  // mapped input: the blocks lie here; otherwise in raw
        Bytes raw;                 // whole BGZF blocks
        std::vector<Blk> blks;
        size_t total = 0;          // inflated bytes
        uint64_t seq = 0;
    };
    void io_loop();
    void driver_loop();
    bool read_raw(RawSlab& rs, std::string& err);   // false at end of file or on error (err set)
    bool read_raw_mapped(RawSlab& rs, std::string& err);
    const uint8_t* map_ = nullptr;  // regular files are mapped: the pool threads inflate straight from the page cache
    size_t map_size_ = 0, map_pos_ = 0;
    void fail(const std::string& err);
    FILE* f_ = nullptr;
    int threads_ = 1;
    size_t slab_bytes_ = 0;     // compressed bytes read per slab
    Bytes carry_;               // tail of the last read: an incomplete block
    bool eof_ = false;
    std::thread io_;
    std::vector<std::thread> drivers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<RawSlab> raw_q_;                              // I/O thread -> drivers
    std::deque<std::pair<uint64_t, std::shared_ptr<Slab>>> done_;  // drivers -> consumer (any order, few entries)
    uint64_t next_seq_ = 0, n_slabs_ = 0;                   // next slab the consumer takes; slabs the file holds (valid once io_done_)
    bool io_done_ = false, stop_ = false;
    std::string err_;
};

class BgzfWriter {
public:
    ~BgzfWriter();
    bool open(const char* path, int threads, int level, std::string& err);
    bool write(const void* data, size_t n, std::string& err);
    // A finished piece of the stream, taken over without a copy: what was written before goes out first (closing its last block
    // short), then the piece becomes deflate work of its own.
    bool write_owned(Bytes&& piece, std::string& err);
    bool close(std::string& err);  // flushes, writes the BGZF end-of-file marker

private:
    bool hand_over(bool all, std::string& err);   // whole blocks (all: everything) of pending_ -> the background thread
    void bg_loop();
    void io_loop();              // writes the deflated chunks to the file, in order, while the next chunk is being deflated
    bool deflate_chunk(const Bytes& in, std::string& err);
    FILE* f_ = nullptr;
    int threads_ = 1, level_ = 6;
    Bytes pending_;
    std::thread bg_;             // deflates and writes the chunks in order
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Bytes> q_;
    bool end_ = false;
    std::string bg_err_;
    std::thread io_;
    std::deque<std::vector<Bytes>> wq_;   // deflated chunks (one Bytes per BGZF block) waiting for the file write
    bool io_end_ = false;
};

struct BamHeader {
    std::string text;              // SAM header text (without the NUL padding some writers add)
    std::vector<uint8_t> refs;     // n_ref + reference records, verbatim
};

class BamReader {
public:
    bool open(const char* path, int threads, BamHeader& hdr, std::string& err);
    // Next alignment record body (SAMv1 4.2 without block_size).  The pointer stays valid as long as slab() -- the slab the
    // record ends in -- is kept alive; nothing is copied except records that straddle two slabs.  false at end of file
    // (err empty) or on error.
    bool next(const uint8_t*& body, size_t& len, std::string& err);
    const std::shared_ptr<Slab>& slab() const { return cur_; }

private:
    bool advance(std::string& err);                                 // cur_ exhausted: take the next slab
    bool read_bytes(uint8_t* dst, size_t n, std::string& err);      // gathers across slabs
    BgzfReader z_;
    std::shared_ptr<Slab> cur_;
    size_t pos_ = 0;
};

class BamWriter {
public:
    bool open(const char* path, int threads, int level, const BamHeader& hdr, std::string& err);
    bool write_record(const uint8_t* body, size_t len, std::string& err);
    // records already in stream form (block_size + body each), handed over without a copy
    bool write_chunk(Bytes&& records, std::string& err) { return z_.write_owned(std::move(records), err); }
    bool close(std::string& err);

private:
    BgzfWriter z_;
};

}  // namespace hm
