// fast_deflate.cpp -- see fast_deflate.h.  Written from RFC 1951 (block formats, code-length alphabet and its transmission order,
// length / distance base tables, canonical code assignment).
#include "fast_deflate.h"

#include <algorithm>
#include <cstring>
#include <memory>

#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace hm {
namespace {

// ---- RFC 1951 section 3.2.5: length and distance symbols ------------------------------------------------------------------
constexpr uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
constexpr uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
constexpr uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
constexpr uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
constexpr uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};  // section 3.2.7

inline uint32_t bit_reverse(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// ---- encoder --------------------------------------------------------------------------------------------------------------

// Code lengths of a Huffman code over freq[0, n) limited to max_bits; symbols with frequency 0 get length 0.  `complete` forces a
// second code when only one symbol is used (the code-length alphabet must not be incomplete; the other two may be).
void huffman_lengths(const uint32_t* freq, int n, int max_bits, bool complete, uint8_t* len)
{
    struct Leaf { uint32_t f; int sym; };
    Leaf leaf[288];
    int used = 0;
    for (int s = 0; s < n; ++s) {
        len[s] = 0;
        if (freq[s]) leaf[used++] = Leaf{freq[s], s};
    }
    if (used == 0) return;
    if (used == 1) {
        len[leaf[0].sym] = 1;
        if (complete) len[leaf[0].sym ? 0 : 1] = 1;
        return;
    }
    std::sort(leaf, leaf + used, [](const Leaf& a, const Leaf& b) { return a.f != b.f ? a.f < b.f : a.sym < b.sym; });
    // two-queue construction: leaves in ascending order, internal nodes are created in ascending order of weight
    uint64_t w[2 * 288];
    int parent[2 * 288];
    for (int i = 0; i < used; ++i) w[i] = leaf[i].f;
    int a = 0, b = used, next = used;  // a: next unmerged leaf, b: next unmerged internal node, next: node being created
    while (next < 2 * used - 1) {
        int pick[2];
        for (int k = 0; k < 2; ++k) {
            if (a < used && (b >= next || w[a] <= w[b])) pick[k] = a++;
            else pick[k] = b++;
        }
        w[next] = w[pick[0]] + w[pick[1]];
        parent[pick[0]] = parent[pick[1]] = next;
        ++next;
    }
    int depth[2 * 288];
    const int root = 2 * used - 2;
    depth[root] = 0;
    for (int i = root - 1; i >= 0; --i) depth[i] = depth[parent[i]] + 1;
    // clamp to max_bits, then repair the Kraft sum by lengthening the shortest possible codes
    int count[32] = {};
    int overflow = 0;
    for (int i = 0; i < used; ++i) {
        int d = depth[i];
        if (d > max_bits) { d = max_bits; ++overflow; }
        ++count[d];
    }
    while (overflow > 0) {
        int bits = max_bits - 1;
        while (count[bits] == 0) --bits;
        --count[bits];
        count[bits + 1] += 2;
        --count[max_bits];
        overflow -= 2;
    }
    // the rarest symbols take the longest codes
    int i = 0;
    for (int bits = max_bits; bits >= 1; --bits)
        for (int k = 0; k < count[bits]; ++k) len[leaf[i++].sym] = (uint8_t)bits;
}

// Canonical codes (section 3.2.2), stored bit-reversed: DEFLATE packs Huffman codes starting from their most significant bit.
void canonical_codes(const uint8_t* len, int n, uint16_t* code)
{
    int count[16] = {};
    for (int s = 0; s < n; ++s) ++count[len[s]];
    count[0] = 0;
    uint32_t next[16] = {};
    uint32_t c = 0;
    for (int bits = 1; bits <= 15; ++bits) {
        c = (c + count[bits - 1]) << 1;
        next[bits] = c;
    }
    for (int s = 0; s < n; ++s) code[s] = len[s] ? (uint16_t)bit_reverse(next[len[s]]++, len[s]) : 0;
}

// Bits are collected in a 64-bit word and flushed without a branch: 8 bytes are stored at every flush, the pointer advances by the
// whole bytes among them.  A flush leaves at most 7 bits, so up to 57 bits may be added before the next one.  The 8-byte stores
// need 8 bytes of room behind the last byte of the stream: hm_deflate_rle_bound().
struct BitWriter {
    uint8_t* p;
    uint64_t acc = 0;
    uint32_t n = 0;
    explicit BitWriter(uint8_t* out) : p(out) {}
    inline void add(uint64_t v, uint32_t bits) { acc |= v << n; n += bits; }
    inline void flush()
    {
        memcpy(p, &acc, 8);
        p += n >> 3;
        acc >>= n & ~7u;
        n &= 7u;
    }
    inline void put(uint32_t v, uint32_t bits) { add(v, bits); flush(); }  // bits <= 32
    inline void align_to_byte()
    {
        flush();
        if (n) { *p++ = (uint8_t)acc; acc = 0; n = 0; }
    }
    inline uint8_t* finish() { align_to_byte(); return p; }
};

struct EncScratch {
    uint16_t tok[65536];   // < 256: literal; >= 256: run of length tok - 256 + 3 at distance 1
    uint8_t len_sym[256];  // run length - 3 -> length symbol - 257
    bool init = false;
};

// A block takes this many input bytes: every block has its own code, and the sections of a record (packed bases, qualities, the
// four kinetics arrays, tag text) have different byte statistics -- with one code per 64 KiB payload the output was 1.7 % larger
// than zlib's, which starts a new block every 16 Ki symbols.
constexpr size_t kEncBlock = 16384;

inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

// One DEFLATE block over in[lo, hi) of the payload in[0, n) (runs may start from the byte in front of lo: it is part of the same
// stream).  Dynamic Huffman, or stored when that is not smaller.
void encode_block(EncScratch& S, const uint8_t* in, size_t lo, size_t hi, bool last, BitWriter& bw)
{
    // ---- pass 1: tokens and symbol frequencies (four histograms: equal neighbours do not wait for each other's increment)
    uint32_t freq[288] = {}, f1[256] = {}, f2[256] = {}, f3[256] = {};
    size_t n_tok = 0, n_run = 0;
    uint64_t extra_bits = 0;
    uint16_t* const tok = S.tok;
    size_t i = lo;
    while (i < hi) {
        // eight bytes at a time while none of them equals its predecessor three times in a row (the usual case)
        if (i > 0 && i + 10 <= hi) {
            const uint64_t x = load64(in + i) ^ load64(in + i - 1);  // byte j == 0: in[i + j] == in[i + j - 1]
            uint64_t z = ~(((x & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | x | 0x7f7f7f7f7f7f7f7full);  // 0x80 where byte == 0
            z &= z >> 8;   // ... and the next byte too
            if (z == 0 || (z & (z >> 8)) == 0) {
                // no run of three equalities starts in bytes 0 .. 5; bytes 6 and 7 wait for the next word
                const uint8_t* q = in + i;
                ++freq[q[0]]; ++f1[q[1]]; ++f2[q[2]]; ++f3[q[3]]; ++freq[q[4]]; ++f1[q[5]];
                tok[n_tok] = q[0]; tok[n_tok + 1] = q[1]; tok[n_tok + 2] = q[2]; tok[n_tok + 3] = q[3]; tok[n_tok + 4] = q[4]; tok[n_tok + 5] = q[5];
                n_tok += 6;
                i += 6;
                continue;
            }
        }
        const uint8_t c = in[i];
        if (i > 0 && i + 3 <= hi && in[i - 1] == c && in[i + 1] == c && in[i + 2] == c) {
            size_t l = 3;
            const size_t lim = std::min<size_t>(258, hi - i);
            while (l < lim && in[i + l] == c) ++l;
            const int sy = S.len_sym[l - 3];
            ++freq[257 + sy];
            extra_bits += kLenExtra[sy];
            tok[n_tok++] = (uint16_t)(256 + l - 3);
            ++n_run;
            i += l;
        } else {
            ++freq[c];
            tok[n_tok++] = c;
            ++i;
        }
    }
    for (int k = 0; k < 256; ++k) freq[k] += f1[k] + f2[k] + f3[k];
    freq[256] = 1;
    // ---- codes
    uint8_t ll_len[288], cl_len[19];
    uint16_t ll_code[288], cl_code[19];
    huffman_lengths(freq, 286, 15, false, ll_len);
    canonical_codes(ll_len, 286, ll_code);
    int hlit = 286;
    while (hlit > 257 && ll_len[hlit - 1] == 0) --hlit;
    // code lengths to transmit: hlit literal/length lengths, then ONE distance code of length 1 (distance 1; with no run in the
    // block it is an unused single code, which section 3.2.7 allows)
    uint8_t seq[288 + 1];
    memcpy(seq, ll_len, hlit);
    seq[hlit] = 1;
    const int n_seq = hlit + 1;
    // run-length form of the sequence in the code-length alphabet: (symbol, extra value)
    uint8_t cl_sym[290], cl_ext[290];
    int n_cl = 0;
    uint32_t cl_freq[19] = {};
    for (int k = 0; k < n_seq;) {
        const uint8_t v = seq[k];
        int run = 1;
        while (k + run < n_seq && seq[k + run] == v) ++run;
        k += run;
        if (v == 0) {
            while (run >= 11) { const int r = std::min(run, 138); cl_sym[n_cl] = 18; cl_ext[n_cl++] = (uint8_t)(r - 11); run -= r; }
            if (run >= 3) { cl_sym[n_cl] = 17; cl_ext[n_cl++] = (uint8_t)(run - 3); run = 0; }
        } else {
            cl_sym[n_cl] = v; cl_ext[n_cl++] = 0; --run;  // the value itself, then repeats of it
            while (run >= 3) { const int r = std::min(run, 6); cl_sym[n_cl] = 16; cl_ext[n_cl++] = (uint8_t)(r - 3); run -= r; }
        }
        while (run-- > 0) { cl_sym[n_cl] = v; cl_ext[n_cl++] = 0; }
    }
    for (int k = 0; k < n_cl; ++k) ++cl_freq[cl_sym[k]];
    huffman_lengths(cl_freq, 19, 7, true, cl_len);
    canonical_codes(cl_len, 19, cl_code);
    int hclen = 19;
    while (hclen > 4 && cl_len[kClOrder[hclen - 1]] == 0) --hclen;
    // ---- size of the dynamic block against a stored one (3 header bits, padding, LEN, NLEN, the bytes)
    uint64_t bits = 3 + 5 + 5 + 4 + 3 * (uint64_t)hclen + extra_bits + n_run /* distance codes */;
    for (int k = 0; k < n_cl; ++k) bits += cl_len[cl_sym[k]] + (cl_sym[k] == 16 ? 2 : cl_sym[k] == 17 ? 3 : cl_sym[k] == 18 ? 7 : 0);
    for (int sy = 0; sy < 286; ++sy) bits += (uint64_t)freq[sy] * ll_len[sy];
    const size_t n = hi - lo;
    if ((bits + 7) / 8 >= 5 + n) {
        bw.put(last ? 1u : 0u, 3);
        bw.align_to_byte();
        uint8_t* q = bw.p;
        q[0] = (uint8_t)n; q[1] = (uint8_t)(n >> 8); q[2] = (uint8_t)~n; q[3] = (uint8_t)(~n >> 8);
        memcpy(q + 4, in + lo, n);
        bw.p = q + 4 + n;
        return;
    }
    // ---- pass 2: emit
    bw.add((last ? 1u : 0u) | (2u << 1), 3);  // BFINAL, BTYPE = 10
    bw.add((uint32_t)(hlit - 257), 5);
    bw.add(0, 5);                             // HDIST: one distance code
    bw.add((uint32_t)(hclen - 4), 4);
    bw.flush();
    for (int k = 0; k < hclen; ++k) bw.put(cl_len[kClOrder[k]], 3);
    for (int k = 0; k < n_cl; ++k) {
        const int sy = cl_sym[k];
        bw.add(cl_code[sy], cl_len[sy]);
        if (sy == 16) bw.add(cl_ext[k], 2);
        else if (sy == 17) bw.add(cl_ext[k], 3);
        else if (sy == 18) bw.add(cl_ext[k], 7);
        bw.flush();
    }
    // one table for both kinds of token: code bits and bit count (runs: length code + extra bits + the one-bit distance code 0)
    uint32_t t_bits[512];
    uint8_t t_n[512];
    for (int k = 0; k < 256; ++k) { t_bits[k] = ll_code[k]; t_n[k] = ll_len[k]; }
    for (int l = 0; l < 256; ++l) {
        const int sy = S.len_sym[l];
        const int cl = ll_len[257 + sy];
        t_bits[256 + l] = ll_code[257 + sy] | ((uint32_t)(l + 3 - kLenBase[sy]) << cl);
        t_n[256 + l] = (uint8_t)(cl + kLenExtra[sy] + 1);
    }
    size_t t = 0;
    for (; t + 2 <= n_tok; t += 2) {  // two tokens (<= 21 bits each) per flush
        const uint32_t k0 = tok[t], k1 = tok[t + 1];
        bw.add(t_bits[k0], t_n[k0]);
        bw.add(t_bits[k1], t_n[k1]);
        bw.flush();
    }
    if (t < n_tok) bw.put(t_bits[tok[t]], t_n[tok[t]]);
    bw.put(ll_code[256], ll_len[256]);
}

}  // namespace

size_t hm_deflate_rle(const uint8_t* in, size_t n, uint8_t* out, size_t cap)
{
    if (n > 65535 || cap < hm_deflate_rle_bound(n)) return 0;
    if (n == 0) {
        out[0] = 3;  // BFINAL = 1, BTYPE = 01, end-of-block code 0000000
        out[1] = 0;
        return 2;
    }
    static thread_local EncScratch S;
    if (!S.init) {
        for (int l = 3; l <= 258; ++l) {
            int sy = 28;
            while (kLenBase[sy] > l) --sy;
            S.len_sym[l - 3] = (uint8_t)sy;
        }
        S.init = true;
    }
    BitWriter bw(out);
    for (size_t lo = 0; lo < n; lo += kEncBlock) {
        const size_t hi = std::min(n, lo + kEncBlock);
        encode_block(S, in, lo, hi, hi == n, bw);
    }
    return (size_t)(bw.finish() - out);
}

// ---- decoder --------------------------------------------------------------------------------------------------------------
namespace {

// table entry: bits 0-4 code bits consumed by this lookup, bits 5-7 kind, bits 8-12 extra-bit count (kind kBase) or width of the
// second-level table (kind kSub), bits 16-31 literal / base value / offset of the second-level table
enum : uint32_t { kLit = 0, kBase = 1, kEob = 2, kSub = 3, kBad = 7 };
inline uint32_t mk(uint32_t nbits, uint32_t kind, uint32_t extra, uint32_t value) { return nbits | (kind << 5) | (extra << 8) | (value << 16); }
inline uint32_t e_nbits(uint32_t e) { return e & 31u; }
inline uint32_t e_kind(uint32_t e) { return (e >> 5) & 7u; }
inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 31u; }
inline uint32_t e_value(uint32_t e) { return e >> 16; }

constexpr int kLitBits = 11, kDistBits = 8;
constexpr int kLitTable = (1 << kLitBits) + 288 * 16, kDistTable = (1 << kDistBits) + 32 * 128;

struct DecScratch {
    uint32_t lit[kLitTable];
    uint32_t dist[kDistTable];
    uint8_t sub_bits[1 << kLitBits];
};

// Builds the lookup table of a canonical code.  is_dist selects what a symbol means.  false: over-subscribed, or incomplete in a
// way section 3.2.7 does not allow (only a single code of length 1, or no distance code at all, may be incomplete).
bool build_table(const uint8_t* len, int n, int first_bits, bool is_dist, uint32_t* tab, int tab_cap, uint8_t* sub_bits)
{
    int count[16] = {};
    for (int s = 0; s < n; ++s) ++count[len[s]];
    const int n_codes = n - count[0];
    int max_len = 15;
    while (max_len > 0 && count[max_len] == 0) --max_len;
    int left = 1;
    for (int b = 1; b <= 15; ++b) {
        left = 2 * left - count[b];
        if (left < 0) return false;
    }
    if (left > 0 && !(n_codes == 1 && max_len == 1) && !(n_codes == 0 && is_dist)) return false;
    const int first = 1 << first_bits;
    for (int i = 0; i < first; ++i) tab[i] = mk(1, kBad, 0, 0);
    uint32_t next[16] = {};
    {
        uint32_t c = 0;
        count[0] = 0;
        for (int b = 1; b <= 15; ++b) {
            c = (c + count[b - 1]) << 1;
            next[b] = c;
        }
    }
    // widths of the second-level tables: the longest code under each first-level prefix
    if (max_len > first_bits) {
        memset(sub_bits, 0, (size_t)first);
        uint32_t nx[16];
        memcpy(nx, next, sizeof(nx));
        for (int s = 0; s < n; ++s) {
            const int l = len[s];
            if (l <= first_bits) { if (l) ++nx[l]; continue; }
            const uint32_t rev = bit_reverse(nx[l]++, l);
            uint8_t& sb = sub_bits[rev & (uint32_t)(first - 1)];
            sb = std::max<uint8_t>(sb, (uint8_t)(l - first_bits));
        }
    }
    int used = first;
    for (int s = 0; s < n; ++s) {
        const int l = len[s];
        if (!l) continue;
        const uint32_t rev = bit_reverse(next[l]++, l);
        uint32_t entry_kind, entry_extra = 0, entry_value;
        if (is_dist) {
            if (s >= 30) { entry_kind = kBad; entry_value = 0; }
            else { entry_kind = kBase; entry_extra = kDistExtra[s]; entry_value = kDistBase[s]; }
        } else if (s < 256) { entry_kind = kLit; entry_value = (uint32_t)s; }
        else if (s == 256) { entry_kind = kEob; entry_value = 0; }
        else if (s < 286) { entry_kind = kBase; entry_extra = kLenExtra[s - 257]; entry_value = kLenBase[s - 257]; }
        else { entry_kind = kBad; entry_value = 0; }
        if (l <= first_bits) {
            const uint32_t e = mk((uint32_t)l, entry_kind, entry_extra, entry_value);
            for (uint32_t i = rev; i < (uint32_t)first; i += 1u << l) tab[i] = e;
        } else {
            const uint32_t prefix = rev & (uint32_t)(first - 1);
            const int sb = sub_bits[prefix];
            if (e_kind(tab[prefix]) != kSub) {
                if (used + (1 << sb) > tab_cap) return false;
                tab[prefix] = mk((uint32_t)first_bits, kSub, (uint32_t)sb, (uint32_t)used);
                for (int i = 0; i < (1 << sb); ++i) tab[used + i] = mk(1, kBad, 0, 0);
                used += 1 << sb;
            }
            const uint32_t base = e_value(tab[prefix]);
            const uint32_t e = mk((uint32_t)(l - first_bits), entry_kind, entry_extra, entry_value);
            for (uint32_t i = rev >> first_bits; i < (1u << sb); i += 1u << (l - first_bits)) tab[base + i] = e;
        }
    }
    return true;
}

struct BitReader {
    const uint8_t* ip;
    const uint8_t* end;
    uint64_t buf = 0;
    uint32_t cnt = 0;       // valid bits in buf
    size_t phantom = 0;     // zero bytes appended past the end of the input
    BitReader(const uint8_t* p, size_t n) : ip(p), end(p + n) {}
    inline void refill()  // afterwards cnt >= 56
    {
        if (end - ip >= 8) {
            uint64_t w;
            memcpy(&w, ip, 8);
            buf |= w << cnt;
            ip += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt < 56) {
                if (ip < end) buf |= (uint64_t)*ip++ << cnt;
                else ++phantom;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(uint32_t n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(uint32_t n) { buf >>= n; cnt -= n; }
    inline uint32_t take(uint32_t n) { const uint32_t v = peek(n); drop(n); return v; }
    // true if no bit past the end of the input has been consumed
    inline bool inside() const { return phantom * 8 <= cnt; }
};

}  // namespace

bool hm_inflate_fast(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out)
{
    static thread_local std::unique_ptr<DecScratch> scratch;  // 43 KB of tables, one per pool thread for the life of the thread
    if (!scratch) scratch.reset(new DecScratch);
    DecScratch* const T = scratch.get();
    BitReader br(in, n_in);
    uint8_t* op = out;
    uint8_t* const out_end = out + n_out;
    for (;;) {
        br.refill();
        const uint32_t bfinal = br.take(1), btype = br.take(2);
        if (btype == 0) {
            br.drop(br.cnt & 7u);
            if (br.cnt < 32) br.refill();
            const uint32_t len = br.take(16), nlen = br.take(16);
            if ((len ^ nlen) != 0xffffu || !br.inside()) return false;
            // hand the unread whole bytes of the bit buffer back (cnt is a multiple of 8 here)
            br.ip -= (br.cnt >> 3) - br.phantom;  // inside(): the appended zero bytes are all still in the buffer
            br.phantom = 0;
            br.buf = 0;
            br.cnt = 0;
            if ((size_t)(br.end - br.ip) < len || (size_t)(out_end - op) < len) return false;
            if (len) memcpy(op, br.ip, len);
            br.ip += len;
            op += len;
        } else if (btype == 1 || btype == 2) {
            uint8_t lens[288 + 32];
            int hlit, hdist;
            if (btype == 1) {
                for (int s = 0; s < 144; ++s) lens[s] = 8;
                for (int s = 144; s < 256; ++s) lens[s] = 9;
                for (int s = 256; s < 280; ++s) lens[s] = 7;
                for (int s = 280; s < 288; ++s) lens[s] = 8;
                for (int s = 0; s < 32; ++s) lens[288 + s] = 5;
                hlit = 288;
                hdist = 32;
            } else {
                hlit = (int)br.take(5) + 257;
                hdist = (int)br.take(5) + 1;
                const int hclen = (int)br.take(4) + 4;
                if (hlit > 286 || hdist > 30) return false;
                uint8_t cl[19] = {};
                br.refill();
                for (int i = 0; i < hclen; ++i) {
                    if (br.cnt < 3) br.refill();
                    cl[kClOrder[i]] = (uint8_t)br.take(3);
                }
                // the code-length code: at most 7 bits, one flat table
                uint32_t clt[128];
                if (!build_table(cl, 19, 7, false, clt, 128, T->sub_bits)) return false;
                // (build_table marks symbols >= 0 as literals here: value = the code-length symbol)
                int i = 0;
                while (i < hlit + hdist) {
                    br.refill();
                    const uint32_t e = clt[br.peek(7)];
                    if (e_kind(e) != kLit) return false;
                    br.drop(e_nbits(e));
                    const uint32_t v = e_value(e);
                    if (v < 16) { lens[i++] = (uint8_t)v; continue; }
                    int rep;
                    uint8_t fill = 0;
                    if (v == 16) {
                        if (i == 0) return false;
                        fill = lens[i - 1];
                        rep = 3 + (int)br.take(2);
                    } else if (v == 17) rep = 3 + (int)br.take(3);
                    else rep = 11 + (int)br.take(7);
                    if (i + rep > hlit + hdist) return false;
                    memset(lens + i, fill, (size_t)rep);
                    i += rep;
                }
                if (!br.inside() || lens[256] == 0) return false;
                // distance lengths behind the literal/length lengths -> fixed position
                if (hlit != 288) memmove(lens + 288, lens + hlit, (size_t)hdist);
            }
            if (!build_table(lens, hlit, kLitBits, false, T->lit, kLitTable, T->sub_bits)) return false;
            if (!build_table(lens + 288, hdist, kDistBits, true, T->dist, kDistTable, T->sub_bits)) return false;
            const uint32_t* const lt = T->lit;
            const uint32_t* const dt = T->dist;
            for (;;) {
                br.refill();  // >= 56 bits: a length (15 + 5) and a distance (15 + 13) fit without another refill
                uint32_t e = lt[br.peek(kLitBits)];
                // up to three short literals per refill (3 x 11 bits) while there is room for them
                if (e_kind(e) == kLit && out_end - op >= 3) {
                    br.drop(e_nbits(e));
                    *op++ = (uint8_t)e_value(e);
                    e = lt[br.peek(kLitBits)];
                    if (e_kind(e) == kLit) {
                        br.drop(e_nbits(e));
                        *op++ = (uint8_t)e_value(e);
                        e = lt[br.peek(kLitBits)];
                        if (e_kind(e) == kLit) {
                            br.drop(e_nbits(e));
                            *op++ = (uint8_t)e_value(e);
                            continue;
                        }
                    }
                    // 22 bits used at most: 34 left, short of the 48 a match may need
                    br.refill();
                }
                if (e_kind(e) == kSub) {
                    br.drop(kLitBits);
                    e = lt[e_value(e) + br.peek(e_extra(e))];
                }
                br.drop(e_nbits(e));
                const uint32_t kind = e_kind(e);
                if (kind == kLit) {
                    if (op >= out_end) return false;
                    *op++ = (uint8_t)e_value(e);
                    continue;
                }
                if (kind == kEob) break;
                if (kind != kBase) return false;
                const uint32_t length = e_value(e) + br.take(e_extra(e));
                uint32_t d = dt[br.peek(kDistBits)];
                if (e_kind(d) == kSub) {
                    br.drop(kDistBits);
                    d = dt[e_value(d) + br.peek(e_extra(d))];
                }
                if (e_kind(d) != kBase) return false;
                br.drop(e_nbits(d));
                const uint32_t distance = e_value(d) + br.take(e_extra(d));
                if (distance > (size_t)(op - out) || length > (size_t)(out_end - op)) return false;
                const uint8_t* src = op - distance;
                if (distance == 1) {
                    memset(op, *src, length);
                    op += length;
                } else if (distance >= 8 && (size_t)(out_end - op) >= length + 8) {
                    uint8_t* const stop = op + length;
                    do {
                        uint64_t w;
                        memcpy(&w, src, 8);
                        memcpy(op, &w, 8);
                        src += 8;
                        op += 8;
                    } while (op < stop);
                    op = stop;
                } else {
                    for (uint32_t k = 0; k < length; ++k) op[k] = src[k];
                    op += length;
                }
            }
            if (!br.inside()) return false;
        } else
            return false;
        if (bfinal) break;
    }
    return br.inside() && op == out_end;
}

// ---- CRC-32 ---------------------------------------------------------------------------------------------------------------
#if defined(__x86_64__)
namespace {
// Folds data[0, n) (n a multiple of 16, n >= 64) into 16 bytes that continue to the same CRC: x^k mod P constants of the
// reflected polynomial for k = 4*128 +- 32 (four lanes in flight) and 128 +- 32 (one lane).
__attribute__((target("pclmul,sse4.1")))
void crc32_fold(uint32_t raw_state, const uint8_t* data, size_t n, uint8_t out[16])
{
    const __m128i k1k2 = _mm_set_epi64x(0x1c6e41596ll, 0x154442bd4ll);
    const __m128i k3k4 = _mm_set_epi64x(0x0ccaa009ell, 0x1751997d0ll);
    const __m128i* p = reinterpret_cast<const __m128i*>(data);
    __m128i x1 = _mm_xor_si128(_mm_loadu_si128(p), _mm_cvtsi32_si128((int)raw_state));
    __m128i x2 = _mm_loadu_si128(p + 1), x3 = _mm_loadu_si128(p + 2), x4 = _mm_loadu_si128(p + 3);
    p += 4;
    n -= 64;
    while (n >= 64) {
        const __m128i l1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), l2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        const __m128i l3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), l4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k1k2, 0x11), l1), _mm_loadu_si128(p));
        x2 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x2, k1k2, 0x11), l2), _mm_loadu_si128(p + 1));
        x3 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x3, k1k2, 0x11), l3), _mm_loadu_si128(p + 2));
        x4 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x4, k1k2, 0x11), l4), _mm_loadu_si128(p + 3));
        p += 4;
        n -= 64;
    }
#define HM_FOLD1(a, next) _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(a, k3k4, 0x11), _mm_clmulepi64_si128(a, k3k4, 0x00)), next)
    x1 = HM_FOLD1(x1, x2);
    x1 = HM_FOLD1(x1, x3);
    x1 = HM_FOLD1(x1, x4);
    while (n >= 16) {
        x1 = HM_FOLD1(x1, _mm_loadu_si128(p));
        ++p;
        n -= 16;
    }
#undef HM_FOLD1
    _mm_storeu_si128(reinterpret_cast<__m128i*>(out), x1);
}
}  // namespace
#endif

uint32_t hm_crc32(uint32_t crc, const uint8_t* data, size_t n)
{
#if defined(__x86_64__)
    static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (have && n >= 128) {
        const size_t bulk = n & ~(size_t)15;
        uint8_t rem[16];
        crc32_fold(~crc, data, bulk, rem);
        // the 16 folded bytes stand for everything read so far: their CRC from the all-zero state, then the tail
        uint32_t c = (uint32_t)crc32(0xffffffffu, rem, 16);
        if (n > bulk) c = (uint32_t)crc32(c, data + bulk, (uInt)(n - bulk));
        return c;
    }
#endif
    return (uint32_t)crc32(crc, data, (uInt)n);
}

}  // namespace hm
