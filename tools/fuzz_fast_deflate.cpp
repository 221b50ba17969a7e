// fuzz_fast_deflate.cpp -- address / undefined-behaviour sanitizer run of the BGZF block codec (hifimeth_b200/csrc/fast_deflate.cpp)
// against zlib: random payloads of several shapes -> zlib at random level / strategy / window -> hm_inflate_fast (exact-size
// buffers, so any byte past either end is a sanitizer report), hm_deflate_rle -> zlib inflate, and damaged streams (bit flips,
// truncation, random bytes) on which the only requirements are "no report" and "if accepted, zlib accepts it with the same bytes".
// Build: g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -I../hifimeth_b200/csrc -o fuzz_fast_deflate \
//            fuzz_fast_deflate.cpp ../hifimeth_b200/csrc/fast_deflate.cpp -lz
// Run:   ./fuzz_fast_deflate [seconds] [seed]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <zlib.h>

#include "fast_deflate.h"

static std::mt19937_64 rng;
static uint32_t rnd(uint32_t n) { return n ? (uint32_t)(rng() % n) : 0; }

static std::vector<uint8_t> payload()
{
    const uint32_t n = rnd(8) == 0 ? rnd(300) : rnd(65536);
    std::vector<uint8_t> v(n);
    switch (rnd(6)) {
    case 0: for (auto& b : v) b = (uint8_t)rng(); break;
    case 1: { const uint32_t k = 1 + rnd(16); for (auto& b : v) b = (uint8_t)rnd(k); break; }
    case 2: { size_t i = 0; while (i < n) { const uint8_t c = (uint8_t)rnd(7); size_t r = 1 + rnd(rnd(4) ? 6 : 700); while (r-- && i < n) v[i++] = c; } break; }
    case 3: { std::geometric_distribution<int> g(0.02 + 0.2 * (rnd(100) / 100.0)); for (auto& b : v) b = (uint8_t)std::min(255, g(rng)); break; }
    case 4: { const char* t = "MM:Z:C+m?,0,3,17,2;ML:B:C,250,3,17@PG\tID:x\n"; for (size_t i = 0; i < n; ++i) v[i] = (uint8_t)t[(i + rnd(3)) % 40]; break; }
    default: { for (size_t i = 0; i < n; ++i) v[i] = i >= 64 && rnd(4) ? v[i - 1 - rnd(63)] : (uint8_t)rng(); break; }  // LZ-friendly
    }
    return v;
}

static std::vector<uint8_t> zdeflate(const std::vector<uint8_t>& in, int level, int strategy, int wbits, int memlevel)
{
    z_stream zs{};
    if (deflateInit2(&zs, level, Z_DEFLATED, -wbits, memlevel, strategy) != Z_OK) abort();
    std::vector<uint8_t> out(deflateBound(&zs, (uLong)in.size()) + 64);
    zs.next_in = const_cast<Bytef*>(in.data());
    zs.avail_in = (uInt)in.size();
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    // a few flushes in the middle: extra (empty stored) blocks
    if (in.size() > 100 && rnd(3) == 0) {
        zs.avail_in = (uInt)(in.size() / 2);
        deflate(&zs, rnd(2) ? Z_SYNC_FLUSH : Z_FULL_FLUSH);
        zs.avail_in += (uInt)(in.size() - in.size() / 2);
    }
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) abort();
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return out;
}

// zlib's verdict: true + bytes if `comp` is a complete raw DEFLATE stream of exactly n_out bytes (trailing garbage allowed? no)
static bool zinflate(const std::vector<uint8_t>& comp, size_t n_out, std::vector<uint8_t>& out)
{
    z_stream zs{};
    if (inflateInit2(&zs, -15) != Z_OK) abort();
    out.assign(n_out + 1, 0);
    zs.next_in = const_cast<Bytef*>(comp.data());
    zs.avail_in = (uInt)comp.size();
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && zs.total_out == n_out;
    out.resize(ok ? n_out : 0);
    inflateEnd(&zs);
    return ok;
}

int main(int argc, char** argv)
{
    const double secs = argc > 1 ? atof(argv[1]) : 30;
    rng.seed(argc > 2 ? strtoull(argv[2], nullptr, 10) : 1);
    const auto t0 = std::chrono::steady_clock::now();
    unsigned long long n_ok = 0, n_own = 0, n_bad = 0, n_bad_accepted = 0;
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < secs) {
        const std::vector<uint8_t> data = payload();
        // zlib -> own inflater
        static const int strategies[] = {Z_DEFAULT_STRATEGY, Z_FILTERED, Z_HUFFMAN_ONLY, Z_RLE, Z_FIXED};
        const std::vector<uint8_t> comp = zdeflate(data, (int)rnd(10), strategies[rnd(5)], 9 + (int)rnd(7), 1 + (int)rnd(9));
        {
            std::vector<uint8_t> in_exact(comp), out_exact(data.size());
            if (!hm::hm_inflate_fast(in_exact.data(), in_exact.size(), out_exact.data(), out_exact.size()) || out_exact != data) { printf("FAIL: zlib stream not read\n"); return 1; }
            ++n_ok;
        }
        // own deflater -> zlib and own inflater
        if (data.size() <= 65535) {
            std::vector<uint8_t> own(hm::hm_deflate_rle_bound(data.size()));
            const size_t n = hm::hm_deflate_rle(data.data(), data.size(), own.data(), own.size());
            if (!n) { printf("FAIL: deflate returned 0\n"); return 1; }
            own.resize(n);
            std::vector<uint8_t> back;
            if (!zinflate(own, data.size(), back) || back != data) { printf("FAIL: zlib cannot read own stream\n"); return 1; }
            std::vector<uint8_t> out_exact(data.size());
            if (!hm::hm_inflate_fast(own.data(), own.size(), out_exact.data(), out_exact.size()) || out_exact != data) { printf("FAIL: own stream not read back\n"); return 1; }
            ++n_own;
        }
        // damaged streams
        for (int k = 0; k < 4; ++k) {
            std::vector<uint8_t> bad(comp);
            switch (rnd(4)) {
            case 0: if (!bad.empty()) for (uint32_t f = 1 + rnd(4); f--;) bad[rnd((uint32_t)bad.size())] ^= (uint8_t)(1u << rnd(8)); break;
            case 1: bad.resize(rnd((uint32_t)bad.size() + 1)); break;
            case 2: for (auto& b : bad) if (rnd(50) == 0) b = (uint8_t)rng(); break;
            default: bad.resize(rnd(200)); for (auto& b : bad) b = (uint8_t)rng(); break;
            }
            const size_t want = rnd(3) ? data.size() : rnd(65536);
            std::vector<uint8_t> in_exact(bad), out_exact(want);
            const bool ok = hm::hm_inflate_fast(in_exact.data(), in_exact.size(), out_exact.data(), out_exact.size());
            ++n_bad;
            if (ok) {
                std::vector<uint8_t> ref;
                if (!zinflate(bad, want, ref) || ref != out_exact) { printf("FAIL: accepted a stream zlib rejects or reads differently\n"); return 1; }
                ++n_bad_accepted;
            }
        }
    }
    printf("ok: %llu zlib streams read, %llu own streams round-tripped, %llu damaged streams (%llu of them still valid and read like zlib)\n", n_ok, n_own, n_bad, n_bad_accepted);
    return 0;
}
