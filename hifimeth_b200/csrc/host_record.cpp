// host_record.cpp -- host-side record helpers of the C ABI: BAM record -> pinned SoA staging
// (hm_pack_record), CodecV1 (hm_codev1_*), and the MM/ML/MN writer (hm_build_mod_record).
//
// Reference behaviour mirrored here:
//   acceptance rules     src/app/hifimeth/mod_main.cpp:189-196, src/corelib/bam_info.cpp:443-453,572-603
//   CodecV1              src/corelib/bam_info.cpp:455-478 (encode), :562-570 (decode table)
//   tag stripping + MM/ML/MN   src/corelib/build_mod_bam.cpp:87-109,125-247
#include <cstring>
#include <vector>

#include "../../include/hm_engine.h"
#include "bgzf_bam.h"

namespace {

struct AuxField {
    size_t tag_off;  // offset of tag[0]
    size_t end_off;  // one past the field
};

// Walks the aux area.  Returns false when a field is truncated or has an unknown type.
bool next_aux(const uint8_t* b, size_t p, size_t len, AuxField& f)
{
    if (p + 3 > len) return false;
    size_t q = p + 3;
    switch (b[p + 2]) {
    case 'A': case 'c': case 'C': q += 1; break;
    case 's': case 'S': q += 2; break;
    case 'i': case 'I': case 'f': q += 4; break;
    case 'd': q += 8; break;
    case 'Z': case 'H':
        while (q < len && b[q]) ++q;
        q += 1;
        break;
    case 'B': {
        if (q + 5 > len) return false;
        size_t es;
        switch (b[q]) {
        case 'c': case 'C': es = 1; break;
        case 's': case 'S': es = 2; break;
        case 'i': case 'I': case 'f': es = 4; break;
        default: return false;
        }
        uint32_t n;
        memcpy(&n, b + q + 1, 4);
        q += 5 + es * (size_t)n;
        break;
    }
    default: return false;
    }
    if (q > len) return false;
    f.tag_off = p;
    f.end_off = q;
    return true;
}

struct RecLayout {
    uint16_t flag;
    int32_t l_seq;
    size_t seq_off, aux_off;
};

bool layout_of(const uint8_t* body, size_t len, RecLayout& r)
{
    if (len < 32) return false;
    uint8_t l_read_name = body[8];
    uint16_t n_cigar;
    memcpy(&n_cigar, body + 12, 2);
    memcpy(&r.flag, body + 14, 2);
    memcpy(&r.l_seq, body + 16, 4);
    if (r.l_seq < 0) return false;
    r.seq_off = 32 + (size_t)l_read_name + 4u * (size_t)n_cigar;
    r.aux_off = r.seq_off + (size_t)((r.l_seq + 1) >> 1) + (size_t)r.l_seq;
    return r.aux_off <= len;
}

size_t put_decimal(uint8_t* dst, uint32_t v)
{
    uint8_t tmp[12];
    size_t n = 0;
    do { tmp[n++] = (uint8_t)('0' + v % 10); v /= 10; } while (v);
    for (size_t i = 0; i < n; ++i) dst[i] = tmp[n - 1 - i];
    return n;
}

// bam_aux_update_int (contract: src/htslib/sam.h:1844-1866; htslib 1.19.1 sam.c, not vendored): a new tag gets the smallest
// unsigned type that holds the value -- htslib compares with `<`, so 255 is stored as 'S' and 65535 as 'I' -- and an EXISTING
// integer tag keeps its width when the value fits in it (the field is reused in place), growing only when it does not.
// old_type = 0: no existing field.
size_t put_mn(uint8_t* dst, int32_t l_seq, char old_type = 0)
{
    int sz = l_seq < 0xff ? 1 : (l_seq < 0xffff ? 2 : 4);
    const int old_sz = (old_type == 'c' || old_type == 'C') ? 1 : (old_type == 's' || old_type == 'S') ? 2 : (old_type == 'i' || old_type == 'I') ? 4 : 0;
    if (old_sz > sz) sz = old_sz;
    dst[0] = 'M'; dst[1] = 'N';
    const uint32_t v = (uint32_t)l_seq;
    dst[2] = sz == 1 ? 'C' : (sz == 2 ? 'S' : 'I');
    memcpy(dst + 3, &v, (size_t)sz);  // little endian
    return 3 + (size_t)sz;
}

}  // namespace

extern "C" {

uint8_t hm_codev1_encode(uint32_t frames)
{
    uint32_t s = frames < 952u ? frames : 952u;
    if (s >= 448) return (uint8_t)((s - 448) / 8 + 192);
    if (s >= 192) return (uint8_t)((s - 192) / 4 + 128);
    if (s >= 64) return (uint8_t)((s - 64) / 2 + 64);
    return (uint8_t)s;
}

uint16_t hm_codev1_decode(uint8_t code)
{
    uint32_t c = code;
    if (c < 64) return (uint16_t)c;
    if (c < 128) return (uint16_t)((c - 64) * 2 + 64);
    if (c < 192) return (uint16_t)((c - 128) * 4 + 192);
    return (uint16_t)((c - 192) * 8 + 448);
}

// Writes record `body` as read i of the batch at base offset b0 / packed-SEQ offset s0 (both already known to fit).
static void pack_at(hm_read_batch* b, uint32_t i, uint32_t b0, uint32_t s0, const uint8_t* body, size_t len, const RecLayout& r,
                    int32_t min_read_len)
{
    const uint32_t l = (uint32_t)r.l_seq;
    // locate the four kinetics tags (first occurrence wins, as bam_aux_get does)
    const uint8_t* tag[4] = {nullptr, nullptr, nullptr, nullptr};
    static const char names[4][2] = {{'f', 'i'}, {'f', 'p'}, {'r', 'i'}, {'r', 'p'}};
    size_t p = r.aux_off;
    AuxField f;
    while (p < len && next_aux(body, p, len, f)) {
        for (int k = 0; k < 4; ++k)
            if (!tag[k] && body[p] == (uint8_t)names[k][0] && body[p + 1] == (uint8_t)names[k][1]) tag[k] = body + p + 2;
        p = f.end_off;
    }
    bool ok = r.l_seq >= min_read_len;
    for (int k = 0; k < 4 && ok; ++k) {
        const uint8_t* t = tag[k];
        if (!t || t[0] != 'B' || (t[1] != 'C' && t[1] != 'S')) { ok = false; break; }
        uint32_t n;
        memcpy(&n, t + 2, 4);
        if (n != l) ok = false;
    }
    uint8_t* planes[4] = {b->fi + b0, b->fp + b0, b->ri + b0, b->rp + b0};
    if (ok) {
        for (int k = 0; k < 4; ++k) {
            const uint8_t* t = tag[k];
            if (t[1] == 'C') memcpy(planes[k], t + 6, l);
            else
                for (uint32_t j = 0; j < l; ++j) {
                    uint16_t v;
                    memcpy(&v, t + 6 + 2 * (size_t)j, 2);
                    planes[k][j] = hm_codev1_encode(v);
                }
        }
    } else {
        for (int k = 0; k < 4; ++k) memset(planes[k], 0, l);
    }
    memcpy(b->seq4 + s0, body + r.seq_off, (l + 1) >> 1);
    b->flag[i] = r.flag;
    b->valid[i] = ok ? 1 : 0;
}

int hm_pack_record(hm_read_batch* b, uint32_t* n_reads, const uint8_t* body, size_t len, int32_t min_read_len)
{
    if (!b || !n_reads || !body) return HM_ERR_ARG;
    RecLayout r;
    if (!layout_of(body, len, r)) return HM_ERR_FORMAT;
    uint32_t i = *n_reads;
    uint32_t l = (uint32_t)r.l_seq;
    if (i >= b->max_reads) return HM_ERR_ARG;
    uint32_t b0 = i ? b->base_off[i] : 0, s0 = i ? b->seq_off[i] : 0;
    if ((uint64_t)b0 + l > b->max_bases) return HM_ERR_ARG;
    if (i == 0) { b->base_off[0] = 0; b->seq_off[0] = 0; }
    pack_at(b, i, b0, s0, body, len, r, min_read_len);
    b->base_off[i + 1] = b0 + l;
    b->seq_off[i + 1] = s0 + ((l + 1) >> 1);
    *n_reads = i + 1;
    return HM_OK;
}

int hm_pack_records(hm_read_batch* b, uint32_t n, const uint8_t* const* bodies, const size_t* lens, int32_t min_read_len, int threads,
                    int32_t* read_index, uint32_t* n_packed)
{
    if (!b || (n && (!bodies || !lens || !read_index)) || !n_packed) return HM_ERR_ARG;
    // pass 1 (serial, touches 32 bytes per record): layouts and offsets
    std::vector<RecLayout> lay(n);
    uint32_t m = 0, b0 = 0, s0 = 0;
    b->base_off[0] = 0;
    b->seq_off[0] = 0;
    for (uint32_t k = 0; k < n; ++k) {
        read_index[k] = -1;
        if (!layout_of(bodies[k], lens[k], lay[k])) continue;  // malformed: the caller passes the bytes through
        const uint32_t l = (uint32_t)lay[k].l_seq;
        if (l > b->max_bases) continue;                         // longer than any batch: passed through uncalled
        if (m >= b->max_reads || (uint64_t)b0 + l > b->max_bases) return HM_ERR_ARG;
        read_index[k] = (int32_t)m;
        b0 += l;
        s0 += (l + 1) >> 1;
        ++m;
        b->base_off[m] = b0;
        b->seq_off[m] = s0;
    }
    // pass 2 (parallel): the copies
    hm::parallel_for(n, threads, [&](size_t k) {
        const int32_t i = read_index[k];
        if (i >= 0) pack_at(b, (uint32_t)i, b->base_off[i], b->seq_off[i], bodies[k], lens[k], lay[k], min_read_len);
    });
    *n_packed = m;
    return HM_OK;
}

// Copies the record up to the aux block and filters the tags the way s_remove_skipped_tags + bam_aux_update_int do
// (src/corelib/build_mod_bam.cpp:87-109,178-247): drops fi/ri/fp/rp (unless -k), the first ML and MM; an existing integer MN is
// rewritten in place when there are calls.  Returns the output length so far.
static size_t strip_tags(const uint8_t* body, size_t len, const RecLayout& r, int keep_kinetics, bool have_calls, uint8_t* out,
                         bool& mn_written)
{
    memcpy(out, body, r.aux_off);
    size_t o = r.aux_off;
    bool dropped[6] = {false, false, false, false, false, false};
    static const char names[6][2] = {{'f', 'i'}, {'r', 'i'}, {'f', 'p'}, {'r', 'p'}, {'M', 'L'}, {'M', 'M'}};
    mn_written = false;
    size_t p = r.aux_off;
    AuxField f;
    while (p < len) {
        if (!next_aux(body, p, len, f)) {  // corrupt tail: keep bytes as they are
            memcpy(out + o, body + p, len - p);
            o += len - p;
            break;
        }
        bool drop = false;
        for (int k = 0; k < 6; ++k) {
            if (dropped[k] || body[p] != (uint8_t)names[k][0] || body[p + 1] != (uint8_t)names[k][1]) continue;
            if (k < 4 && keep_kinetics) continue;
            dropped[k] = true;
            drop = true;
        }
        if (!drop && have_calls && !mn_written && body[p] == 'M' && body[p + 1] == 'N') {
            // an integer MN is updated where it stands, keeping its width when l_seq fits; the reference aborts on an MN of any other
            // type (bam_aux_update_int fails, build_mod_bam.cpp:240-247) -- here that field is dropped and a fresh MN is appended
            if (strchr("cCsSiI", body[p + 2])) {
                o += put_mn(out + o, r.l_seq, (char)body[p + 2]);
                mn_written = true;
            }
            drop = true;
        }
        if (!drop) {
            memcpy(out + o, body + p, f.end_off - p);
            o += f.end_off - p;
        }
        p = f.end_off;
    }
    return o;
}

size_t hm_mod_record_bound(size_t len, uint32_t n_calls) { return len + 64 + 12 * (size_t)n_calls; }

int hm_build_mod_record(const uint8_t* body, size_t len, int keep_kinetics, const int32_t* fwd_qoff,
                        const uint8_t* fwd_ml, uint32_t n_fwd, const int32_t* rev_qoff, const uint8_t* rev_ml,
                        uint32_t n_rev, uint8_t* out, size_t* out_len)
{
    RecLayout r;
    if (!body || !out || !out_len || !layout_of(body, len, r)) return HM_ERR_FORMAT;
    const bool have_calls = (n_fwd + n_rev) > 0;
    bool mn_written = false;
    size_t o = strip_tags(body, len, r, keep_kinetics, have_calls, out, mn_written);
    if (!have_calls) { *out_len = o; return HM_OK; }

    // forward-strand base at original-read offset k: get_bam_fwd_strand_base (bam_info.cpp:224-232)
    const uint8_t* seq = body + r.seq_off;
    const bool is_rev = (r.flag & 16) != 0;
    const int32_t L = r.l_seq;
    auto fwd_nib = [&](int32_t k) -> int {
        int32_t i = is_rev ? L - 1 - k : k;
        int nib = (seq[i >> 1] >> ((~i & 1) << 2)) & 0xf;
        if (!is_rev) return nib;
        switch (nib) { case 1: return 8; case 2: return 4; case 4: return 2; case 8: return 1; default: return nib; }
    };
    out[o++] = 'M'; out[o++] = 'M'; out[o++] = 'Z';
    for (int pass = 0; pass < 2; ++pass) {
        const int32_t* qoff = pass ? rev_qoff : fwd_qoff;
        const uint32_t n = pass ? n_rev : n_fwd;
        const int target = pass ? 4 : 2;  // G : C
        out[o++] = pass ? 'G' : 'C';
        out[o++] = pass ? '-' : '+';
        out[o++] = 'm';
        int32_t last = 0;
        for (uint32_t i = 0; i < n; ++i) {
            if (qoff[i] < last || qoff[i] >= L || fwd_nib(qoff[i]) != target) return HM_ERR_ARG;
            uint32_t delta = 0;
            for (int32_t k = last; k < qoff[i]; ++k) delta += (fwd_nib(k) == target);
            out[o++] = ',';
            o += put_decimal(out + o, delta);
            last = qoff[i] + 1;
        }
        out[o++] = ';';
    }
    out[o++] = 0;
    out[o++] = 'M'; out[o++] = 'L'; out[o++] = 'B'; out[o++] = 'C';
    uint32_t cnt = n_fwd + n_rev;
    memcpy(out + o, &cnt, 4); o += 4;
    if (n_fwd) memcpy(out + o, fwd_ml, n_fwd);
    o += n_fwd;
    if (n_rev) memcpy(out + o, rev_ml, n_rev);
    o += n_rev;
    if (!mn_written) o += put_mn(out + o, r.l_seq);
    *out_len = o;
    return HM_OK;
}

int hm_build_mod_record_mm(const uint8_t* body, size_t len, int keep_kinetics, const uint8_t* mm_fwd, uint32_t mm_fwd_len,
                           const uint8_t* mm_rev, uint32_t mm_rev_len, const uint8_t* ml, uint32_t n_fwd, uint32_t n_rev, uint8_t* out,
                           size_t* out_len)
{
    RecLayout r;
    if (!body || !out || !out_len || !layout_of(body, len, r)) return HM_ERR_FORMAT;
    const bool have_calls = (n_fwd + n_rev) > 0;
    if (have_calls && (!ml || (n_fwd && !mm_fwd) || (n_rev && !mm_rev))) return HM_ERR_ARG;
    bool mn_written = false;
    size_t o = strip_tags(body, len, r, keep_kinetics, have_calls, out, mn_written);
    if (!have_calls) { *out_len = o; return HM_OK; }
    out[o++] = 'M'; out[o++] = 'M'; out[o++] = 'Z';
    out[o++] = 'C'; out[o++] = '+'; out[o++] = 'm';
    if (mm_fwd_len) memcpy(out + o, mm_fwd, mm_fwd_len);
    o += mm_fwd_len;
    out[o++] = ';';
    out[o++] = 'G'; out[o++] = '-'; out[o++] = 'm';
    if (mm_rev_len) memcpy(out + o, mm_rev, mm_rev_len);
    o += mm_rev_len;
    out[o++] = ';';
    out[o++] = 0;
    out[o++] = 'M'; out[o++] = 'L'; out[o++] = 'B'; out[o++] = 'C';
    const uint32_t cnt = n_fwd + n_rev;
    memcpy(out + o, &cnt, 4); o += 4;
    memcpy(out + o, ml, cnt);
    o += cnt;
    if (!mn_written) o += put_mn(out + o, r.l_seq);
    *out_len = o;
    return HM_OK;
}

// ---- row N4: MM/ML parser (round-trip validator of the tags this engine writes) ------------------------------------------

namespace {

// Forward-strand base at forward offset i as a letter (BamQuerySequence::get_bam_fwd_strand_base, src/corelib/bam_info.cpp:
// 224-232): the stored base, or under flag 0x10 the complement of the stored base at l_seq - 1 - i.  0 for an illegal nibble.
char fwd_strand_base(const uint8_t* body, const RecLayout& r, int32_t i)
{
    const bool rev = (r.flag & 16) != 0;
    const int32_t q = rev ? r.l_seq - 1 - i : i;
    const uint32_t nib = (body[r.seq_off + ((size_t)q >> 1)] >> ((~q & 1) << 2)) & 0xfu;
    switch (nib) {
    case 1: return rev ? 'T' : 'A';
    case 2: return rev ? 'G' : 'C';
    case 4: return rev ? 'C' : 'G';
    case 8: return rev ? 'A' : 'T';
    case 15: return 'N';
    default: return 0;
    }
}

// s_chebi_to_iupac_code, src/corelib/bam_mod_parser.cpp:36-76
char chebi_code(long c)
{
    switch (c) {
    case 27551: return 'm'; case 76792: return 'h'; case 76794: return 'f'; case 76793: return 'c'; case 16964: return 'g';
    case 80961: return 'e'; case 17477: return 'b'; case 28871: return 'a'; case 44605: return 'o'; case 18107: return 'n';
    default: return 0;
    }
}

// s_is_valid_unmod_base_and_code, src/corelib/bam_mod_parser.cpp:101-139
bool code_fits_base(char base, char c)
{
    if (c == 'm' || c == 'h' || c == 'f' || c == 'c' || c == 'C') return base == 'C' || base == 'G';
    if (c == 'g' || c == 'e' || c == 'b' || c == 'T') return base == 'T' || base == 'A';
    if (c == 'U') return base == 'U';
    if (c == 'a' || c == 'A') return base == 'A' || base == 'T';
    if (c == 'o' || c == 'G') return base == 'G' || base == 'C';
    if (c == 'n' || c == 'N') return base == 'N';
    return true;
}

const uint8_t* find_tag(const uint8_t* body, size_t len, const RecLayout& r, char t0, char t1)
{
    size_t p = r.aux_off;
    AuxField f;
    while (p < len && next_aux(body, p, len, f)) {
        if (body[p] == (uint8_t)t0 && body[p + 1] == (uint8_t)t1) return body + p + 2;  // at the type byte
        p = f.end_off;
    }
    return nullptr;
}

}  // namespace

int hm_parse_mod_record(const uint8_t* body, size_t len, int32_t* qoff, uint8_t* strand, uint8_t* prob, char* code, uint32_t cap,
                        uint32_t* n_mods)
{
    if (!body || !n_mods) return HM_ERR_ARG;
    *n_mods = 0;
    RecLayout r;
    if (!layout_of(body, len, r)) return HM_ERR_FORMAT;
    // ML: any integer B array, every value in [0, 255] (s_extract_bam_mod_scaled_probs, bam_mod_parser.cpp:8-34)
    const uint8_t* ml = find_tag(body, len, r, 'M', 'L');
    if (!ml) return HM_OK;
    if (ml[0] != 'B') return HM_ERR_FORMAT;
    uint32_t n_probs;
    memcpy(&n_probs, ml + 2, 4);
    const char sub = (char)ml[1];
    const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I') ? 4 : 0;
    if (!es) return HM_ERR_FORMAT;
    auto prob_at = [&](uint32_t i, uint8_t& out) {
        const uint8_t* p = ml + 6 + es * (size_t)i;
        long long v;
        switch (sub) {
        case 'c': v = (int8_t)p[0]; break;
        case 'C': v = p[0]; break;
        case 's': { int16_t t; memcpy(&t, p, 2); v = t; break; }
        case 'S': { uint16_t t; memcpy(&t, p, 2); v = t; break; }
        case 'i': { int32_t t; memcpy(&t, p, 4); v = t; break; }
        default: { uint32_t t; memcpy(&t, p, 4); v = t; break; }
        }
        if (v < 0 || v > 255) return false;
        out = (uint8_t)v;
        return true;
    };
    if (n_probs == 0) return HM_OK;
    const uint8_t* mm = find_tag(body, len, r, 'M', 'M');
    if (!mm) return HM_OK;
    if (mm[0] != 'Z') return HM_ERR_FORMAT;
    const char* s = reinterpret_cast<const char*>(mm + 1);
    const size_t sl = strlen(s);
    if (sl == 0 || s[sl - 1] != ';') return HM_ERR_FORMAT;  // "The MM aux tag must end with ';'"
    uint32_t prob_idx = 0, n_out = 0;
    size_t i = 0;
    while (i < sl) {
        size_t j = i + 1;
        while (j < sl && s[j] != ';') ++j;
        ++j;  // one past the ';' (the string ends with one)
        // ---- one edit series s[i, j): base, strand, codes, skip counts (s_parse_one_mod_list, bam_mod_parser.cpp:141-229) ----
        const char* e = s + i;
        const size_t el = j - i;
        if (el < 4) return HM_ERR_FORMAT;
        const char base = e[0];
        if (base != 'C' && base != 'G' && base != 'T' && base != 'A' && base != 'U' && base != 'N') return HM_ERR_FORMAT;
        if (e[1] != '+' && e[1] != '-') return HM_ERR_FORMAT;
        const uint8_t st = e[1] == '+' ? 0 : 1;  // FWD / REV, src/corelib/hbn_aux.hpp:60-63
        char codes[16];
        int n_code = 0;
        size_t k = 2;
        if (e[2] >= '0' && e[2] <= '9') {
            long c = 0;
            while (k < el && e[k] >= '0' && e[k] <= '9') { c = c * 10 + (e[k] - '0'); if (c > 100000000) return HM_ERR_FORMAT; ++k; }
            const char cc = chebi_code(c);
            if (!cc) return HM_ERR_FORMAT;
            codes[n_code++] = cc;
        } else {
            for (; k < el; ++k) {
                if (e[k] == ',' || e[k] == ';') break;
                if (e[k] != '.' && e[k] != '?') {
                    if (n_code == 16) return HM_ERR_FORMAT;
                    codes[n_code++] = e[k];
                }
            }
        }
        for (int c = 0; c < n_code; ++c)
            if (!code_fits_base(base, codes[c])) return HM_ERR_FORMAT;
        if (k >= el || (e[k] != ',' && e[k] != ';')) return HM_ERR_FORMAT;
        ++k;
        int32_t q = 0;
        while (k < el) {
            if (e[k] < '0' || e[k] > '9') return HM_ERR_FORMAT;
            long long d = 0;
            while (k < el && e[k] != ',' && e[k] != ';') {
                if (e[k] < '0' || e[k] > '9') return HM_ERR_FORMAT;
                d = d * 10 + (e[k] - '0');
                if (d > 0x7fffffff) return HM_ERR_FORMAT;
                ++k;
            }
            ++k;
            // skip d occurrences of the unmodified base, then land on the next one
            long long cnt = 0;
            while (cnt < d) {
                if (q >= r.l_seq) return HM_ERR_FORMAT;
                if (fwd_strand_base(body, r, q) == base) ++cnt;
                ++q;
            }
            for (;;) {
                if (q >= r.l_seq) return HM_ERR_FORMAT;
                if (fwd_strand_base(body, r, q) == base) break;
                ++q;
            }
            for (int c = 0; c < n_code; ++c) {
                if (prob_idx >= n_probs) return HM_ERR_FORMAT;
                uint8_t pv;
                if (!prob_at(prob_idx++, pv)) return HM_ERR_FORMAT;
                if (n_out < cap) {
                    if (qoff) qoff[n_out] = q;
                    if (strand) strand[n_out] = st;
                    if (prob) prob[n_out] = pv;
                    if (code) code[n_out] = codes[c];
                }
                ++n_out;
            }
            ++q;
        }
        i = j;
    }
    *n_mods = n_out;
    return HM_OK;
}

uint8_t hm_ml_threshold(const uint64_t bins[256], uint64_t* n_samples)
{
    if (n_samples) *n_samples = 0;
    if (!bins) return 128;
    int st = 20, en = 256 - 20;
    while (st < 256 && bins[st] < 10) ++st;
    while (en && bins[en - 1] < 10) --en;
    uint64_t sum = 0, min_cnt = ~(uint64_t)0;
    int min_i = -1;
    if (en - st >= 50)
        for (int i = st; i < en; ++i) {
            sum += bins[i];
            if (min_cnt > bins[i]) { min_cnt = bins[i]; min_i = i; }  // first of equal minima, as the reference's strict '>'
        }
    if (n_samples) *n_samples = sum;
    return (sum < 10000 || min_i == -1) ? (uint8_t)128 : (uint8_t)min_i;
}

}  // extern "C"
