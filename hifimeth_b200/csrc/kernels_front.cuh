// kernels_front.cuh -- sm_100a kernels for the integer/byte front of the hot path:
//   decode_kernel        A1 + A2: packed 4-bit SEQ -> forward-strand base codes; fi/fp/ri/rp CodecV1 codes
//                        -> frames, re-packed per forward position (HBM bound: 4.5 B/base in, 9 B/base out)
//   scan_count_kernel /  A3: CpG / CHG / CHH site scan on both strands as a two-pass scan-compaction
//   scan_offsets_kernel / scan_write_kernel   (HBM bound: 1 B/base in, 4 B/site + 12 B/site out)
//   gather_features_kernel  A4: [n,401,8] f32 windows (validation / microbench layout, 12 832 B/site)
//
// Reference semantics: src/corelib/bam_info.cpp:169-222,520-570; src/app/hifimeth/eval_kmer_features.cpp:9-126;
// src/corelib/5mc_context.cpp:4-10.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hm {

constexpr int kChunk = 1024;          // positions per scan chunk; a chunk never straddles two reads
constexpr int kKmer = 401;
constexpr int kHalf = 200;
constexpr int kFpb = 8;
constexpr uint32_t kCodeN = 4;        // internal code for N / any non-ACGT base

// class of a site: 0 CpG (fwd), 1 CHG (fwd), 2 CHH fwd, 3 CHH rev
struct __align__(16) ClassCount { uint32_t c[4]; };

// CodecV1 code -> frames without a table: code = g*64 + r  ->  (r << g) + 64 * ((1 << g) - 1)
// (identical to the 256-entry table of src/corelib/bam_info.cpp:562-570).
__device__ __forceinline__ uint32_t codev1_frames(uint32_t code)
{
    uint32_t g = code >> 6, r = code & 63u;
    return (r << g) + 64u * ((1u << g) - 1u);
}

__device__ __forceinline__ uint32_t nib_to_code(uint32_t nib)
{
    // 1 A, 2 C, 4 G, 8 T -> 0..3; anything else (15 = N) -> kCodeN
    return nib == 1 ? 0u : nib == 2 ? 1u : nib == 4 ? 2u : nib == 8 ? 3u : kCodeN;
}

// One block per chunk (kChunk forward positions of one read), 256 threads x 4 positions.
// bcode[B+p]  = forward-strand code at forward position p (flag 0x10: reverse complement of the stored SEQ)
// kinf[B+p]   = frames {fi[p], fp[p], ri[L-1-p], rp[L-1-p]}: the four kinetics of forward position p's base
//               pair, so that a window is one contiguous run in either strand direction.
__global__ void __launch_bounds__(256)
decode_kernel(const uint8_t* __restrict__ seq4, const uint8_t* __restrict__ fi, const uint8_t* __restrict__ fp,
              const uint8_t* __restrict__ ri, const uint8_t* __restrict__ rp, const uint32_t* __restrict__ base_off,
              const uint32_t* __restrict__ seq_off, const uint16_t* __restrict__ flag,
              const uint32_t* __restrict__ chunk_read, const uint32_t* __restrict__ chunk_pos,
              uint8_t* __restrict__ bcode, ushort4* __restrict__ kinf)
{
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t p0 = chunk_pos[blockIdx.x];
    const uint32_t B = base_off[r];
    const uint32_t L = base_off[r + 1] - B;
    const uint32_t S = seq_off[r];
    const bool rev = (flag[r] & 16) != 0;
    #pragma unroll
    for (int it = 0; it < kChunk / 256; ++it) {
        uint32_t p = p0 + it * 256 + threadIdx.x;
        if (p >= L) break;
        uint32_t q = rev ? L - 1 - p : p;  // index into the stored SEQ
        uint32_t nib = (seq4[S + (q >> 1)] >> ((~q & 1u) << 2)) & 0xfu;
        uint32_t code = nib_to_code(nib);
        if (rev && code < 4) code = 3u - code;
        bcode[B + p] = (uint8_t)code;
        ushort4 k;
        k.x = (unsigned short)codev1_frames(fi[B + p]);
        k.y = (unsigned short)codev1_frames(fp[B + p]);
        k.z = (unsigned short)codev1_frames(ri[B + L - 1 - p]);
        k.w = (unsigned short)codev1_frames(rp[B + L - 1 - p]);
        kinf[B + p] = k;
    }
}

// Un-permutes decode_kernel's packed output back into the reference's views (validation hook):
// frames planes in native coordinates and the two strands' BLASTNA codes (N = 14).
__global__ void __launch_bounds__(256)
decode_unpack_kernel(const uint8_t* __restrict__ bcode, const ushort4* __restrict__ kinf,
                     const uint32_t* __restrict__ base_off, const uint32_t* __restrict__ chunk_read,
                     const uint32_t* __restrict__ chunk_pos, uint16_t* __restrict__ ofi, uint16_t* __restrict__ ofp,
                     uint16_t* __restrict__ ori, uint16_t* __restrict__ orp, uint8_t* __restrict__ fwd_qs,
                     uint8_t* __restrict__ rev_qs)
{
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t p0 = chunk_pos[blockIdx.x];
    const uint32_t B = base_off[r];
    const uint32_t L = base_off[r + 1] - B;
    for (int it = 0; it < kChunk / 256; ++it) {
        uint32_t p = p0 + it * 256 + threadIdx.x;
        if (p >= L) break;
        ushort4 k = kinf[B + p];
        ofi[B + p] = k.x;
        ofp[B + p] = k.y;
        ori[B + L - 1 - p] = k.z;
        orp[B + L - 1 - p] = k.w;
        uint32_t c = bcode[B + p];
        fwd_qs[B + p] = (uint8_t)(c < 4 ? c : 14u);
        rev_qs[B + L - 1 - p] = (uint8_t)(c < 4 ? 3u - c : 14u);
    }
}

// Site class at forward position p of a read with codes c[0..L): 0..3, or -1 for none.
// CpG: C,G.  CHG: C,[ACT],G.  CHH fwd: C,[ACT],[ACT].  CHH rev: [AGT],[AGT],G -> the G.
// (eval_kmer_features.cpp:67-126; the three sets are disjoint.)
__device__ __forceinline__ int site_class(const uint8_t* __restrict__ c, uint32_t p, uint32_t L, uint32_t ctx_mask)
{
    uint32_t b = c[p];
    if (b == 1u) {
        if (p + 1 >= L) return -1;
        uint32_t n1 = c[p + 1];
        if (n1 == 2u) return (ctx_mask & 1u) ? 0 : -1;
        if (n1 > 3u || p + 2 >= L) return -1;
        uint32_t n2 = c[p + 2];
        if (n2 == 2u) return (ctx_mask & 2u) ? 1 : -1;
        if (n2 > 3u) return -1;
        return (ctx_mask & 4u) ? 2 : -1;
    }
    if (b == 2u) {
        if (p < 2 || !(ctx_mask & 4u)) return -1;
        uint32_t m1 = c[p - 1], m2 = c[p - 2];
        if (m1 > 3u || m2 > 3u || m1 == 1u || m2 == 1u) return -1;
        return 3;
    }
    return -1;
}

// Pass 1: per-chunk class counts.  One block (kChunk threads) per chunk.
__global__ void __launch_bounds__(kChunk)
scan_count_kernel(const uint8_t* __restrict__ bcode, const uint32_t* __restrict__ base_off,
                  const uint8_t* __restrict__ valid, const uint32_t* __restrict__ chunk_read,
                  const uint32_t* __restrict__ chunk_pos, uint32_t ctx_mask, ClassCount* __restrict__ chunk_cnt)
{
    __shared__ uint32_t s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t B = base_off[r];
    const uint32_t L = base_off[r + 1] - B;
    const uint32_t p = chunk_pos[blockIdx.x] + threadIdx.x;
    int cls = -1;
    if (valid[r] && p < L) cls = site_class(bcode + B, p, L, ctx_mask);
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t m = __ballot_sync(0xffffffffu, cls == k);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt[k], __popc(m));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ClassCount cc;
        cc.c[0] = s_cnt[0]; cc.c[1] = s_cnt[1]; cc.c[2] = s_cnt[2]; cc.c[3] = s_cnt[3];
        chunk_cnt[blockIdx.x] = cc;
    }
}

// Pass 2: exclusive prefix over chunks (single block; n_chunks is ~ bases / 1024).  pref has
// n_chunks + 1 entries; pref[n_chunks] = totals.  Also emits the per-read call offsets and forward
// counts, and the totals record the host reads back: totals[0..3] class totals, totals[4] = n_calls.
__global__ void __launch_bounds__(1024)
scan_offsets_kernel(const ClassCount* __restrict__ chunk_cnt, uint32_t n_chunks, ClassCount* __restrict__ pref,
                    uint32_t* __restrict__ totals)
{
    __shared__ ClassCount s_part[1024];
    const uint32_t per = (n_chunks + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per;
    const uint32_t hi = min(lo + per, n_chunks);
    ClassCount sum = {{0, 0, 0, 0}};
    for (uint32_t i = lo; i < hi; ++i) {
        ClassCount v = chunk_cnt[i];
        #pragma unroll
        for (int k = 0; k < 4; ++k) sum.c[k] += v.c[k];
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        ClassCount add = {{0, 0, 0, 0}};
        if ((int)threadIdx.x >= off) add = s_part[threadIdx.x - off];
        __syncthreads();
        #pragma unroll
        for (int k = 0; k < 4; ++k) s_part[threadIdx.x].c[k] += add.c[k];
        __syncthreads();
    }
    ClassCount run = {{0, 0, 0, 0}};
    if (threadIdx.x > 0) run = s_part[threadIdx.x - 1];
    for (uint32_t i = lo; i < hi; ++i) {
        pref[i] = run;
        ClassCount v = chunk_cnt[i];
        #pragma unroll
        for (int k = 0; k < 4; ++k) run.c[k] += v.c[k];
    }
    if (threadIdx.x == 1023) {
        ClassCount t = s_part[1023];
        pref[n_chunks] = t;
        totals[0] = t.c[0]; totals[1] = t.c[1]; totals[2] = t.c[2]; totals[3] = t.c[3];
        totals[4] = t.c[0] + t.c[1] + t.c[2] + t.c[3];
    }
}

__global__ void read_offsets_kernel(const ClassCount* __restrict__ pref, const uint32_t* __restrict__ read_first_chunk,
                                    uint32_t n_reads, uint32_t* __restrict__ call_off, uint32_t* __restrict__ n_fwd,
                                    uint32_t* __restrict__ read_pref)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    ClassCount a = pref[read_first_chunk[r]];  // read_first_chunk[n_reads] = n_chunks
    reinterpret_cast<ClassCount*>(read_pref)[r] = a;  // per-class sites in reads before r (sub-batching of the CNN stage)
    call_off[r] = a.c[0] + a.c[1] + a.c[2] + a.c[3];
    if (r < n_reads) {
        ClassCount b = pref[read_first_chunk[r + 1]];
        n_fwd[r] = (b.c[0] + b.c[1] + b.c[2]) - (a.c[0] + a.c[1] + a.c[2]);
    }
}

// Pass 3: write.  Output order per read: forward-strand calls (classes 0,1,2) ascending, then reverse
// (class 3) ascending -- the order the worker thread hands to build_one_mod_bam (mod_main.cpp:217-251).
// Per-class work lists (site_read / site_pos / site_out) are laid out class after class.
__global__ void __launch_bounds__(kChunk)
scan_write_kernel(const uint8_t* __restrict__ bcode, const uint32_t* __restrict__ base_off,
                  const uint8_t* __restrict__ valid, const uint32_t* __restrict__ chunk_read,
                  const uint32_t* __restrict__ chunk_pos, const uint32_t* __restrict__ read_first_chunk,
                  const ClassCount* __restrict__ pref, uint32_t n_chunks, uint32_t ctx_mask,
                  int32_t* __restrict__ qoff, uint8_t* __restrict__ call_ctx, uint32_t* __restrict__ site_read,
                  uint32_t* __restrict__ site_pos, uint32_t* __restrict__ site_out)
{
    __shared__ uint32_t s_warp[32][4];
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t B = base_off[r];
    const uint32_t L = base_off[r + 1] - B;
    const uint32_t p = chunk_pos[blockIdx.x] + threadIdx.x;
    int cls = -1;
    if (valid[r] && p < L) cls = site_class(bcode + B, p, L, ctx_mask);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t rank_in_warp[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t m = __ballot_sync(0xffffffffu, cls == k);
        rank_in_warp[k] = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp][k] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = s_warp[lane][k];
            uint32_t incl = v;
            #pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += t;
            }
            s_warp[lane][k] = incl - v;  // exclusive
        }
    }
    __syncthreads();
    if (cls < 0) return;
    uint32_t rk[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k) rk[k] = s_warp[warp][k] + rank_in_warp[k];

    const ClassCount pc = pref[blockIdx.x];
    const ClassCount p0 = pref[read_first_chunk[r]];
    const ClassCount p1 = pref[read_first_chunk[r + 1]];
    const ClassCount tot = pref[n_chunks];
    const uint32_t call_base = p0.c[0] + p0.c[1] + p0.c[2] + p0.c[3];
    uint32_t out;
    if (cls < 3) {
        uint32_t fwd_before = (pc.c[0] + pc.c[1] + pc.c[2]) - (p0.c[0] + p0.c[1] + p0.c[2]);
        out = call_base + fwd_before + rk[0] + rk[1] + rk[2];
    } else {
        uint32_t nfwd = (p1.c[0] + p1.c[1] + p1.c[2]) - (p0.c[0] + p0.c[1] + p0.c[2]);
        out = call_base + nfwd + (pc.c[3] - p0.c[3]) + rk[3];
    }
    qoff[out] = (int32_t)p;
    call_ctx[out] = (uint8_t)(cls == 3 ? 2 : cls);
    uint32_t region = 0;
    #pragma unroll
    for (int k = 0; k < 3; ++k) if (k < cls) region += tot.c[k];
    const uint32_t slot = region + pc.c[cls] + rk[cls];
    site_read[slot] = r;
    site_pos[slot] = p | (cls == 3 ? 0x80000000u : 0u);
    site_out[slot] = out;
}

// A4: one block per site, 128 threads; thread w writes window row w (8 floats, 32 B).
// Row for strand position i:  onehot(seq[i]), lut[own ipd]/952, lut[own pw]/952, lut[opp ipd]/952,
// lut[opp pw]/952 with IEEE fp32 division, zero-filled outside the read (eval_kmer_features.cpp:36-64).
__global__ void __launch_bounds__(128)
gather_features_kernel(const uint8_t* __restrict__ bcode, const ushort4* __restrict__ kinf,
                       const uint32_t* __restrict__ base_off, const uint32_t* __restrict__ site_read,
                       const uint32_t* __restrict__ site_pos, uint32_t first, uint32_t count, float* __restrict__ out)
{
    const uint32_t s = blockIdx.x;
    if (s >= count) return;
    const uint32_t r = site_read[first + s];
    const uint32_t sp = site_pos[first + s];
    const bool rev = (sp & 0x80000000u) != 0;
    const int p = (int)(sp & 0x7fffffffu);
    const uint32_t B = base_off[r];
    const int L = (int)(base_off[r + 1] - B);
    const int o = rev ? L - 1 - p : p;
    float4* dst = reinterpret_cast<float4*>(out + (size_t)s * kKmer * kFpb);
    for (int w = threadIdx.x; w < kKmer; w += 128) {
        int i = o - kHalf + w;  // strand coordinate
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (i >= 0 && i < L) {
            int j = rev ? L - 1 - i : i;  // forward coordinate
            uint32_t c = bcode[B + j];
            if (rev && c < 4) c = 3u - c;
            a.x = c == 0 ? 1.f : 0.f; a.y = c == 1 ? 1.f : 0.f; a.z = c == 2 ? 1.f : 0.f; a.w = c == 3 ? 1.f : 0.f;
            ushort4 k = kinf[B + j];
            float f0 = __fdiv_rn((float)k.x, 952.0f), f1 = __fdiv_rn((float)k.y, 952.0f);
            float f2 = __fdiv_rn((float)k.z, 952.0f), f3 = __fdiv_rn((float)k.w, 952.0f);
            b = rev ? make_float4(f2, f3, f0, f1) : make_float4(f0, f1, f2, f3);
        }
        dst[2 * w] = a;
        dst[2 * w + 1] = b;
    }
}

}  // namespace hm
