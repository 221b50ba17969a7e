/*
 * oracle/ref_driver.cpp  --  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * C-callable entry points over the REFERENCE'S OWN compiled hot-path code.  This file is compiled
 * together with the unmodified reference sources where they lie under /root/reference/src (see
 * oracle/Makefile) into oracle/_ref/libhifimeth_ref.so.  Nothing here restates the algorithm: each
 * function only builds an in-memory bam1_t from a raw BAM record body and calls the reference class:
 *   ref_query_decode   -> BamQuerySequence::init              src/corelib/bam_info.cpp:169-222
 *                         BamKinetics::init/decoded_ipd/pw    src/corelib/bam_info.cpp:520-603
 *   ref_extract_sites  -> EvalKmerFeaturesGenerator::extract_{cpg,chg,chh}_samples
 *                                                             src/app/hifimeth/eval_kmer_features.cpp:67-126
 *   ref_extract_features -> get_next_sample_features          src/app/hifimeth/eval_kmer_features.cpp:9-65,128-136
 *   ref_build_mod_bam  -> build_one_mod_bam                   src/corelib/build_mod_bam.cpp:125-248
 *   ref_parse_mods     -> extract_bam_base_mods               src/corelib/bam_mod_parser.cpp:231-286
 *
 * A "record body" is the BAM on-disk alignment record WITHOUT its leading block_size field
 * (SAMv1 s4.2: refID,pos,l_read_name,mapq,bin,n_cigar_op,flag,l_seq,next_refID,next_pos,tlen,
 * read_name,cigar,seq,qual,aux).
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>

#include <corelib/bam_info.hpp>
#include <corelib/bam_mod_parser.hpp>
#include <corelib/build_mod_bam.hpp>
#include <eval_kmer_features.hpp>

namespace {

struct Rec {
    bam1_t b;
    Rec(const uint8_t* body, size_t len) {
        memset(&b, 0, sizeof(b));
        int32_t refid, pos, l_seq, next_refid, next_pos, tlen;
        uint8_t l_read_name, mapq;
        uint16_t bin, n_cigar, flag;
        memcpy(&refid, body + 0, 4);
        memcpy(&pos, body + 4, 4);
        l_read_name = body[8];
        mapq = body[9];
        memcpy(&bin, body + 10, 2);
        memcpy(&n_cigar, body + 12, 2);
        memcpy(&flag, body + 14, 2);
        memcpy(&l_seq, body + 16, 4);
        memcpy(&next_refid, body + 20, 4);
        memcpy(&next_pos, body + 24, 4);
        memcpy(&tlen, body + 28, 4);
        b.core.tid = refid;
        b.core.pos = pos;
        b.core.l_qname = l_read_name;
        b.core.l_extranul = 0;
        b.core.qual = mapq;
        b.core.bin = bin;
        b.core.n_cigar = n_cigar;
        b.core.flag = flag;
        b.core.l_qseq = l_seq;
        b.core.mtid = next_refid;
        b.core.mpos = next_pos;
        b.core.isize = tlen;
        b.l_data = (int)(len - 32);
        b.m_data = (uint32_t)(len - 32 + 64);
        b.data = (uint8_t*)malloc(b.m_data);
        memcpy(b.data, body + 32, len - 32);
    }
    ~Rec() { free(b.data); }
    size_t body_size() const { return 32 + (size_t)b.l_data; }
    void write_body(uint8_t* out) const {
        int32_t v;
        v = b.core.tid; memcpy(out + 0, &v, 4);
        v = (int32_t)b.core.pos; memcpy(out + 4, &v, 4);
        out[8] = (uint8_t)b.core.l_qname;
        out[9] = b.core.qual;
        memcpy(out + 10, &b.core.bin, 2);
        uint16_t nc = (uint16_t)b.core.n_cigar; memcpy(out + 12, &nc, 2);
        memcpy(out + 14, &b.core.flag, 2);
        v = b.core.l_qseq; memcpy(out + 16, &v, 4);
        v = b.core.mtid; memcpy(out + 20, &v, 4);
        v = (int32_t)b.core.mpos; memcpy(out + 24, &v, 4);
        v = (int32_t)b.core.isize; memcpy(out + 28, &v, 4);
        memcpy(out + 32, b.data, (size_t)b.l_data);
    }
};

} // namespace

extern "C" {

/* Returns 1 on success, 0 if the reference rejects the record (missing / wrong-length kinetics).
 * Outputs (each l_qseq long): fwd_qs, rev_qs = base codes of the two strands; dec[8] = decoded
 * frames in the order fwd_ipd, fwd_pw, rev_ipd, rev_pw (each indexed in its own strand's
 * coordinates, as BamKinetics::decoded_ipd(strand, offset) is). */
int ref_query_decode(const uint8_t* body, size_t len, uint8_t* fwd_qs, uint8_t* rev_qs,
                     int32_t* fwd_ipd, int32_t* fwd_pw, int32_t* rev_ipd, int32_t* rev_pw)
{
    Rec r(body, len);
    BamQuerySequence q;
    BamKinetics k;
    if (!q.init(&r.b)) return 0;
    if (!k.init(&r.b)) return 0;
    for (int i = 0; i < q.size; ++i) {
        fwd_qs[i] = q.fwd_qs[i];
        rev_qs[i] = q.rev_qs[i];
        fwd_ipd[i] = k.decoded_ipd(FWD, i);
        fwd_pw[i] = k.decoded_pw(FWD, i);
        rev_ipd[i] = k.decoded_ipd(REV, i);
        rev_pw[i] = k.decoded_pw(REV, i);
    }
    return 1;
}

/* ctx: 0 = CpG, 1 = CHG, 2 = CHH.  Returns the number of sites (written up to max_sites), or -1
 * if EvalKmerFeaturesGenerator::init fails. */
int ref_extract_sites(const uint8_t* body, size_t len, int ctx, int32_t* offsets, int max_sites)
{
    Rec r(body, len);
    ns_mods::EvalKmerFeaturesGenerator g;
    if (!g.init(&r.b)) return -1;
    if (ctx == 0) g.extract_cpg_samples();
    else if (ctx == 1) g.extract_chg_samples();
    else g.extract_chh_samples();
    int n = g.M_num_samples;
    for (int i = 0; i < n && i < max_sites; ++i) offsets[i] = g.M_sample_offsets[i];
    return n;
}

/* Features for sites [first, first+count) of context ctx: features[count][kmer][fpb] f32, plus
 * offsets[count], strands[count].  Returns number written, or -1 on init failure. */
int ref_extract_features(const uint8_t* body, size_t len, int ctx, int kmer, int fpb,
                         int first, int count, float* features, int32_t* offsets, int32_t* strands)
{
    Rec r(body, len);
    ns_mods::EvalKmerFeaturesGenerator g;
    if (!g.init(&r.b)) return -1;
    if (ctx == 0) g.extract_cpg_samples();
    else if (ctx == 1) g.extract_chg_samples();
    else g.extract_chh_samples();
    std::vector<float> scratch((size_t)kmer * fpb);
    int off, strand, idx = 0, written = 0;
    while (written < count) {
        float* dst = (idx >= first) ? features + (size_t)written * kmer * fpb : scratch.data();
        if (!g.get_next_sample_features(kmer, fpb, dst, off, strand)) break;
        if (idx >= first) {
            offsets[written] = off;
            strands[written] = strand;
            ++written;
        }
        ++idx;
    }
    return written;
}

/* build_one_mod_bam on a record body.  out must have room for len + extra; *out_len receives the
 * new body size.  Returns 0. */
int ref_build_mod_bam(const uint8_t* body, size_t len, int keep_kinetics,
                      const int32_t* fwd_qoff, const uint8_t* fwd_prob, int nf,
                      const int32_t* rev_qoff, const uint8_t* rev_prob, int nr,
                      uint8_t* out, size_t out_cap, size_t* out_len)
{
    Rec r(body, len);
    std::set<int> skipped;
    if (!keep_kinetics) {
        /* the four tags of fill_skipped_tags(), src/app/hifimeth/mod_main.cpp:119-143 */
        skipped.insert(('f' << 8) | 'p');
        skipped.insert(('r' << 8) | 'p');
        skipped.insert(('f' << 8) | 'i');
        skipped.insert(('r' << 8) | 'i');
    }
    std::vector<MolMethyCall> f(nf), v(nr);
    for (int i = 0; i < nf; ++i) { f[i].qid = 0; f[i].qoff = fwd_qoff[i]; f[i].strand = FWD; f[i].scaled_prob = fwd_prob[i]; }
    for (int i = 0; i < nr; ++i) { v[i].qid = 0; v[i].qoff = rev_qoff[i]; v[i].strand = REV; v[i].scaled_prob = rev_prob[i]; }
    build_one_mod_bam(&r.b, skipped, f.data(), nf, v.data(), nr);
    *out_len = r.body_size();
    if (r.body_size() > out_cap) return -1;
    r.write_body(out);
    return 0;
}

/* extract_bam_base_mods: returns count; fills qoff/strand/prob up to max. */
int ref_parse_mods(const uint8_t* body, size_t len, int32_t* qoff, uint8_t* strand, uint8_t* prob, int max)
{
    Rec r(body, len);
    std::vector<BaseModInfo> mods;
    extract_bam_base_mods(&r.b, mods);
    int n = (int)mods.size();
    for (int i = 0; i < n && i < max; ++i) {
        qoff[i] = mods[i].qoff;
        strand[i] = mods[i].observed_strand;
        prob[i] = mods[i].scaled_prob;
    }
    return n;
}

} // extern "C"
