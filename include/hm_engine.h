/*
 * hm_engine.h -- C ABI of the B200-native engine for the read-level `hifimeth call` hot path.
 *
 * This is the drop-in boundary.  It replaces, inside the reference's worker thread
 * (src/app/hifimeth/mod_main.cpp:145-262), everything between SAM_Batch::get_next_sam
 * (src/corelib/sam_batch.hpp:38) and build_one_mod_bam (src/corelib/build_mod_bam.hpp:8-10):
 *
 *   reference interface                                         replaced by
 *   ----------------------------------------------------------  -------------------------------------
 *   ModModels ctor: read_model(.onnx) / reshape / compile_model  hm_engine_create
 *     src/app/hifimeth/mod_main.cpp:32-98
 *   EvalKmerFeaturesGenerator::init(bam1_t*)                     hm_batch_acquire + hm_pack_record
 *     src/app/hifimeth/eval_kmer_features.hpp:17-22                (record -> pinned SoA staging)
 *   extract_{cpg,chg,chh}_samples + ModBatch::call_mods_for_     hm_batch_submit
 *     one_read + ModBatch::call_current_batch + infer()
 *     src/app/hifimeth/eval_kmer_features.cpp:67-136,
 *     src/app/hifimeth/mod_batch.cpp:66-93
 *   per-read regroup of MolMethyCall (sort by qid, split by      hm_batch_collect
 *     strand, sort by qoff)  src/app/hifimeth/mod_main.cpp:217-251   (fwd/rev qoff[] + ml[] per read =
 *                                                                  the argument lists of build_one_mod_bam)
 *   build_one_mod_bam  src/corelib/build_mod_bam.cpp:125-248      hm_build_mod_record (host helper)
 *
 * Conventions: every function returns 0 on success or a negative hm_status; hm_last_error() gives the
 * message.  Nothing throws or aborts across this boundary (the reference aborts: src/corelib/hbn_aux.hpp:
 * 100-101,146-154).  No torch types, plain pointers and sizes only.  There is no CPU fallback: creation
 * fails with HM_ERR_CUDA when no sm_100 device is usable.
 *
 * Threading: one feeder thread per slot may call acquire/submit/collect for that slot; different slots
 * may be driven from different threads.  Each slot owns a CUDA stream; H2D, kernels and D2H of different
 * slots overlap.  Buffers handed out by acquire/collect stay valid until the next acquire of the slot.
 * hm_batch_submit is NOT fully asynchronous: it returns after the decode and scan kernels of the batch have run (tens of
 * microseconds of device time) -- the host cuts the CNN stage's sub-batches from the per-read site counts the scan produces --
 * and with the CNN stage enqueued; everything after that overlaps the caller.  Two slots per device (what the `call` driver and
 * bench.py use) keep the device busy across that hand-over.
 */
#ifndef HM_ENGINE_H
#define HM_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HM_KMER 401            /* window, from the ONNX input shape [batch,401,8] (mod_main.cpp:41-57) */
#define HM_FEATURES_PER_BASE 8
#define HM_CTX_CPG 1
#define HM_CTX_CHG 2
#define HM_CTX_CHH 4

typedef enum hm_status {
    HM_OK = 0,
    HM_ERR_ARG = -1,      /* bad argument / capacity exceeded */
    HM_ERR_MODEL = -2,    /* model file missing or not the expected graph */
    HM_ERR_CUDA = -3,     /* CUDA runtime error (message names the stage) */
    HM_ERR_STATE = -4,    /* call sequence error (collect without submit, ...) */
    HM_ERR_FORMAT = -5    /* malformed BAM record */
} hm_status;

/* CNN arithmetic.  TENSOR is the product path: bf16 split-precision (hi*hi + lo*hi + hi*lo, fp32
 * accumulate) on tcgen05 tensor cores.  FP32_SIMT evaluates the same graph with fp32 FMAs on the CUDA
 * cores; it exists as an on-device cross-check for the validation hooks, not as a fallback. */
typedef enum hm_cnn_mode { HM_CNN_TENSOR = 0, HM_CNN_FP32_SIMT = 1 } hm_cnn_mode;

typedef struct hm_config {
    const char* model_dir;   /* directory holding CpG.onnx, CHG.onnx, CHH.onnx (-m) */
    int32_t ctx_mask;        /* HM_CTX_* bits (-c); 0 means all three */
    int32_t min_read_len;    /* -l, default 1000 (src/app/hifimeth/mod_options.cpp:10) */
    int32_t device;          /* CUDA device ordinal */
    int32_t n_slots;         /* staging slots, 1..4 (2 = double buffered) */
    uint32_t max_reads;      /* capacity of one batch */
    uint32_t max_bases;      /* capacity of one batch, sum of l_qseq (< 2^31) */
    int32_t cnn_mode;        /* hm_cnn_mode */
    int32_t keep_debug;      /* 1 = keep intermediates for the hm_debug_* hooks (costs memory) */
} hm_config;

/* Engine-owned PINNED staging buffers of one slot, struct-of-arrays.  The caller fills them
 * (hm_pack_record does it from a BAM record) and then calls hm_batch_submit. */
typedef struct hm_read_batch {
    uint32_t max_reads, max_bases;
    uint32_t* base_off;  /* [n_reads+1] prefix sum of l_qseq; base_off[0] = 0 */
    uint32_t* seq_off;   /* [n_reads+1] byte offset of each read's packed SEQ in seq4 */
    uint8_t* seq4;       /* packed 4-bit SEQ exactly as BAM stores it; capacity max_bases/2 + max_reads */
    uint16_t* flag;      /* [n_reads] BAM flag (only 0x10 is looked at) */
    uint8_t* valid;      /* [n_reads] 1 = call; 0 = pass through (short read / kinetics missing) */
    uint8_t* fi;         /* [n_bases] CodecV1 codes, forward IPD, forward-read coordinates */
    uint8_t* fp;         /* [n_bases] forward PW */
    uint8_t* ri;         /* [n_bases] reverse IPD, reverse-strand coordinates */
    uint8_t* rp;         /* [n_bases] reverse PW */
} hm_read_batch;

/* HM_SUBMIT_READ_STATS: per-read sum and maximum of the decoded kinetics frames, planes in the order fi, fp, ri, rp.  Diagnostics
 * only: the model is fed frames / 952 exactly as the reference does (src/corelib/bam_info.hpp:108); SURVEY.md s8a row A9. */
typedef struct hm_read_stats {
    uint64_t sum[4];
    uint32_t max[4];
} hm_read_stats;

/* Result of one batch, in engine-owned pinned memory.  For read r the calls are
 * [call_off[r], call_off[r+1]): first n_fwd[r] forward-strand calls (C on the read) with ascending qoff,
 * then the reverse-strand calls (G on the read) with ascending qoff -- exactly the two arrays
 * build_one_mod_bam takes.  qoff is in forward-strand coordinates of BamQuerySequence. */
typedef struct hm_call_batch {
    uint32_t n_reads;
    uint32_t n_calls;
    const uint32_t* call_off; /* [n_reads+1] */
    const uint32_t* n_fwd;    /* [n_reads] */
    const int32_t* qoff;      /* [n_calls] */
    const uint8_t* ml;        /* [n_calls] scaled_prob = min(255,(int)(255*p1)) (mod_batch.cpp:46-64) */
    uint64_t n_sites[3];      /* CpG, CHG, CHH samples in this batch (mod_main.cpp:364-407 statistics) */
    /* HM_SUBMIT_MM_TEXT only (else NULL): the skip counts of the MM tag as decimal text, built on the device
     * (build_one_mod_bam, src/corelib/build_mod_bam.cpp:134-168).  Read r owns mm_text[mm_off[r] .. mm_off[r+1]): first
     * mm_fwd_len[r] bytes ",d,d,..." for its forward-strand calls, then the same for its reverse-strand calls. */
    const uint8_t* mm_text;
    const uint32_t* mm_off;     /* [n_reads+1] */
    const uint32_t* mm_fwd_len; /* [n_reads] */
    /* HM_SUBMIT_ML_HIST only (else NULL): [3][256] histogram of this batch's ML bytes per context (CpG, CHG, CHH) over reads
     * without flag 0x900 -- what `hifimeth pileup` accumulates to infer its thresholds (src/app/hifimeth/pileup.cpp:237-272). */
    const uint32_t* ml_hist;
    /* HM_SUBMIT_READ_STATS only (else NULL): [n_reads] */
    const hm_read_stats* read_stats;
} hm_call_batch;

/* Device-side timing of the last submit of a slot (CUDA events on the slot's stream). */
typedef struct hm_timing {
    float h2d_ms, decode_ms, scan_ms, cnn_ms, d2h_ms, total_ms;
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t kernel_launches;   /* kernels of this library launched by the submit */
    float top_kernel_ms;        /* device time of the dense plan: CUDA events around all launches of the CNN kernel family of this
                                   submit (the tensor-core kernels plus the ~1 % of small kernels between them) */
    uint32_t top_kernel_launches;
    double executed_flops;      /* tensor-core FLOPs those launches ISSUED: 2 x MACs x 3 split-precision passes over the rows they
                                   process (the dense plan shares conv work between windows, so this differs from the per-site
                                   algorithmic count) */
} hm_timing;

#define HM_SUBMIT_SKIP_H2D 1u   /* inputs of this slot are already resident in HBM (re-run) */
#define HM_SUBMIT_SKIP_D2H 2u   /* leave results on the device (kernel-only timing) */
#define HM_SUBMIT_MM_TEXT 4u    /* also build the MM skip-count text on the device (hm_call_batch.mm_*) */
#define HM_SUBMIT_READ_STATS 16u /* also reduce every read's decoded kinetics to sum / max per plane (hm_call_batch.read_stats) */
#define HM_SUBMIT_ML_HIST 8u    /* also histogram the ML bytes per context on the device (hm_call_batch.ml_hist) */

typedef struct hm_engine hm_engine;

int hm_engine_create(const hm_config* cfg, hm_engine** out);
void hm_engine_destroy(hm_engine* e);
const char* hm_last_error(const hm_engine* e); /* e may be NULL: error of the last failed create */
const char* hm_version(void);

/* Weight loader on its own (no device): parses models/<ctx>.onnx -- either ONNX dialect that ships, what the CPU reference reads
 * (src/app/hifimeth/mod_main.cpp:32-67) -- or models/<ctx>.pt, the TorchScript export its app-gpu binary loads
 * (src/app-gpu/hifimeth-gpu/5mc_call_gpu.cpp:48), by extension, and returns the network's parameters flattened in graph order:
 * bn0 weight, bias, mean, var [8 each]; 8 x (conv W [cout][cin][k], b [cout]); fc1 W [256][128], b; fc2 W [2][256], b.
 * out may be NULL to query *n_floats.  hm_engine_create uses <model_dir>/<ctx>.onnx, or <ctx>.pt when HM_MODEL_FORMAT=pt is set
 * or no .onnx is there. */
int hm_model_weights(const char* path, float* out, size_t cap, size_t* n_floats, int32_t* conv1_k);

int hm_batch_acquire(hm_engine* e, int slot, hm_read_batch* out);
int hm_batch_submit(hm_engine* e, int slot, uint32_t n_reads, uint32_t flags);
int hm_batch_collect(hm_engine* e, int slot, hm_call_batch* out);
int hm_batch_timing(hm_engine* e, int slot, hm_timing* out);

/* ---- host helpers on the record path (A2 re-encode, record -> SoA, A8 tag construction) ---------------- */

/* s_encode_signal_value, src/corelib/bam_info.cpp:455-478: raw frames (B:S tags) -> CodecV1 code. */
uint8_t hm_codev1_encode(uint32_t frames);
/* The decode table of src/corelib/bam_info.cpp:562-570. */
uint16_t hm_codev1_decode(uint8_t code);

/* Append one BAM alignment record body (SAMv1 s4.2 without block_size) to a staging batch at index
 * *n_reads.  Applies the reference's acceptance rules: valid = l_seq >= min_read_len and fi, ri, fp, rp all
 * present as B:C or B:S with count == l_seq (src/app/hifimeth/mod_main.cpp:189-196,
 * src/corelib/bam_info.cpp:443-453).  Returns HM_ERR_ARG (and appends nothing) when the batch is full. */
int hm_pack_record(hm_read_batch* b, uint32_t* n_reads, const uint8_t* body, size_t len, int32_t min_read_len);
/* The same for a whole batch, copying on up to `threads` host threads: records bodies[0..n) become reads 0..*n_packed of an
 * EMPTY staging batch in order; read_index[k] = the read index of record k, or -1 for a record that is malformed or longer
 * than max_bases (the caller passes those through).  HM_ERR_ARG when the records do not fit the batch. */
int hm_pack_records(hm_read_batch* b, uint32_t n, const uint8_t* const* bodies, const size_t* lens, int32_t min_read_len,
                    int threads, int32_t* read_index, uint32_t* n_packed);

/* build_one_mod_bam, src/corelib/build_mod_bam.cpp:125-248, on a record body: strips fi/ri/fp/rp unless
 * keep_kinetics, strips old ML/MM, appends MM:Z ML:B:C MN when there are calls.  out needs
 * hm_mod_record_bound(len, n_calls) bytes.  Returns the new body length in *out_len. */
size_t hm_mod_record_bound(size_t len, uint32_t n_calls);
int hm_build_mod_record(const uint8_t* body, size_t len, int keep_kinetics, const int32_t* fwd_qoff,
                        const uint8_t* fwd_ml, uint32_t n_fwd, const int32_t* rev_qoff, const uint8_t* rev_ml,
                        uint32_t n_rev, uint8_t* out, size_t* out_len);

/* The same record, with the MM skip-count text taken from hm_call_batch (HM_SUBMIT_MM_TEXT) instead of being recounted
 * from the sequence: mm_fwd / mm_rev are the ",d,d,..." runs of the read's forward / reverse calls, ml holds the n_fwd
 * forward bytes followed by the n_rev reverse bytes. */
int hm_build_mod_record_mm(const uint8_t* body, size_t len, int keep_kinetics, const uint8_t* mm_fwd, uint32_t mm_fwd_len,
                           const uint8_t* mm_rev, uint32_t mm_rev_len, const uint8_t* ml, uint32_t n_fwd, uint32_t n_rev,
                           uint8_t* out, size_t* out_len);

/* extract_bam_base_mods, src/corelib/bam_mod_parser.cpp:231-286, on a record body (round-trip validator
 * of the tags above): parses ML (any integer B array, values 0..255) and MM:Z (edit series "<base><+|-><codes|ChEBI>,d,d,...;")
 * and returns one entry per (position, code) in tag order -- qoff in forward-strand coordinates, strand 0 for '+' / 1 for '-',
 * the scaled probability and the code letter.  Arrays may be NULL; at most `cap` entries are written, *n_mods is the full
 * count.  Where the reference aborts (malformed series, skip counts running past the read, too few ML values) this returns
 * HM_ERR_FORMAT. */
int hm_parse_mod_record(const uint8_t* body, size_t len, int32_t* qoff, uint8_t* strand, uint8_t* prob, char* code, uint32_t cap,
                        uint32_t* n_mods);

/* s_resolve_scaled_prob_threshold, src/app/hifimeth/pileup.cpp:355-436, for one context: the scaled-probability threshold is
 * the emptiest bin of the histogram between its outermost bins (inside [20, 236)) holding >= 10 samples, provided that range
 * spans >= 50 bins and holds >= 10000 samples; otherwise 128.  *n_samples (may be NULL) = the samples in that range. */
uint8_t hm_ml_threshold(const uint64_t bins[256], uint64_t* n_samples);

/* ---- the `call` driver and its BAM codec ------------------------------------------------------- */

/* `hifimeth call [-m dir] [-l 1000] [-s 32] [-b 10000] [-k] [-c cpg,chg,chh] [-t N] in.bam out.bam` on this engine
 * (src/app/hifimeth/mod_main.cpp:303-412, options src/app/hifimeth/mod_options.cpp:61-181).  argv[0] = program name,
 * argv[1] = "call".  Returns EXIT_SUCCESS / EXIT_FAILURE like the reference's main(). */
int hm_call_main(int argc, char** argv);
/* on != 0: hm_call_main returns as soon as the output is closed, with the GPU workers' engine teardown still running on detached
 * threads -- for executables that leave through _exit() right after (hifimeth_b200/csrc/main.cpp).  Default: join. */
void hm_call_fast_exit(int on);
/* Reads every record of a BAM file and writes it unchanged (block-parallel BGZF inflate / deflate); returns the number of
 * records or a negative hm_status.  level < 0 (or out_path NULL): read and count only; 100 + level: the writer gets
 * the records in finished pieces (the hand-over `call` uses) instead of one by one.  Test hook for the codec. */
int hm_bam_copy(const char* in_path, const char* out_path, int threads, int level);
/* The raw-DEFLATE block codec behind level 1 of the writer and behind the reader (hifimeth_b200/csrc/fast_deflate.h), one BGZF
 * payload at a time.  Test hooks: hm_deflate_block compresses in[0, n) (n <= 65535; cap >= n + 64) and returns the compressed size
 * (0 on a bad argument); hm_inflate_block returns 1 if in[0, n_in) is a complete DEFLATE stream of exactly n_out bytes, else 0. */
size_t hm_deflate_block(const uint8_t* in, size_t n, uint8_t* out, size_t cap);
int hm_inflate_block(const uint8_t* in, size_t n_in, uint8_t* out, size_t n_out);
/* CRC-32 of BGZF trailers as the reader / writer compute it (carry-less-multiply folding on x86-64); same contract as zlib's crc32(). */
uint32_t hm_crc32_bytes(uint32_t crc, const uint8_t* data, size_t n);

/* ---- validation hooks (parity tests; need cfg.keep_debug = 1, call after hm_batch_collect) ------------ */

/* Decoded kinetics as frames, u16, each [n_bases]: fi, fp in forward coordinates, ri, rp in reverse-strand
 * coordinates (= BamKinetics::decoded_ipd/pw(strand, offset), src/corelib/bam_info.cpp:550-560), plus the
 * two strands' base codes (A0 C1 G2 T3 N14, BamQuerySequence::fwd_qs / rev_qs). */
int hm_debug_dump_decode(hm_engine* e, int slot, uint16_t* fi, uint16_t* fp, uint16_t* ri, uint16_t* rp,
                         uint8_t* fwd_qs, uint8_t* rev_qs);
/* Context (0 CpG, 1 CHG, 2 CHH) of every call, [n_calls], in hm_call_batch order. */
int hm_debug_dump_ctx(hm_engine* e, int slot, uint8_t* ctx);
/* Feature tensors [count][401][8] f32 of calls [first, first+count) in hm_call_batch order
 * (s_extract_kmer_features, src/app/hifimeth/eval_kmer_features.cpp:9-65). */
int hm_debug_dump_features(hm_engine* e, int slot, uint32_t first, uint32_t count, float* out);
/* Logits [n_calls][2] f32 in hm_call_batch order. */
int hm_debug_dump_logits(hm_engine* e, int slot, float* out);

/* What the PRODUCT CNN path actually eats: the [count][401][8] window of each call as held in the engine's X map (bf16 hi + lo,
 * re-summed to f32), i.e. the features after track_features_kernel rather than the fp32 re-gather of hm_debug_dump_features.
 * Same values as s_extract_kmer_features (src/app/hifimeth/eval_kmer_features.cpp:9-65) up to the 16-bit mantissa of hi + lo:
 * one-hot columns and out-of-read rows exact, kinetics within 2^-16 relative.  Tensor path only. */
int hm_debug_dump_xmap(hm_engine* e, int slot, uint32_t first, uint32_t count, float* out);
/* Per-site activations of conv layer `layer` (1..8) of context ctx (0 CpG, 1 CHG, 2 CHH) for calls [first, first + count), all of
 * which must be sites of that context: out [count][*n_pos][*channels] f32, position-major (the transpose of the reference's
 * [C][L] layout, training/model_cnn.py:76-85), assembled from the maps the dense plan leaves in HBM.  NaN marks values the
 * product path never stores (interior positions of layer 1: conv1 and conv2 are fused on chip).  Re-runs the context on the
 * resident batch, which must fit one sub-batch.  Tensor path only. */
int hm_debug_dump_acts(hm_engine* e, int slot, int ctx, int layer, uint32_t first, uint32_t count, float* out, size_t out_floats,
                       int32_t* n_pos, int32_t* channels);

/* One op of the tensor-core dense plan on caller-provided fp32 data (unit test of dense_gemm_kernel; no engine needed):
 *   out[r][:] = relu(bias + sum_k src[term_src[k]][r + term_shift[k]][:] . W_k),  r < rows (multiple of 128)
 * src maps are [rows_alloc][cin] f32, W_k = weights + k*cin*cout as [cin][cout].  conv1_taps > 0 selects the conv1 form:
 * cin = 8, one term, weights [taps][8][cout], rows r + shift .. r + shift + taps - 1.  w2 != NULL selects the head
 * form: out = [rows][2] = relu(...) . w2^T + b2 with w2 [2][cout].  Bit k of gather_mask makes term k read row
 * gather_rows[r] + term_shift[k] instead of r + term_shift[k] (the compact ops of the plan; gather_rows [rows]).
 * Arithmetic: bf16 hi/lo split, fp32 accumulate. */
int hm_debug_dense_op(int device, uint32_t rows, uint32_t rows_alloc, int cin, int cout, int n_src, const float* const* src,
                      int n_terms, const int32_t* term_src, const int32_t* term_shift, const float* weights, const float* bias,
                      int conv1_taps, const float* w2, const float* b2, const uint32_t* gather_rows, uint32_t gather_mask, float* out);

/* Device time (ms, CUDA events) of the kernel launched by the last hm_debug_dense_op (tools/dense_microbench.py). */
float hm_debug_last_op_ms(void);

/* ---- kernel microbenchmarks (BASELINE.json config 5) --------------------------------------------------- */
/* Runs one named kernel family `iters` times on the slot's resident inputs and returns the mean device
 * time per launch (CUDA events) and the algorithmic bytes / flops one launch processes.
 * name: "decode", "scan", "gather", "cnn", "mm" (row N1: MM skip counts + text), "stats" (row A9 diagnostics). */
int hm_microbench(hm_engine* e, int slot, const char* name, uint32_t n_sites, int iters, float* ms_per_launch,
                  double* algo_bytes, double* algo_flops);

#ifdef __cplusplus
}
#endif
#endif /* HM_ENGINE_H */
