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

constexpr int kChunk = 4096;          // positions per chunk (one block); a chunk never straddles two reads
constexpr int kFrontThreads = 256;    // 16 positions per thread
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

// Copies n bytes starting at src (any alignment) into dst + (src & 15) with 16-byte global loads; returns the pointer to the
// first copied byte in dst.  May read up to 15 bytes before src and after src + n (inside the same allocation: every input
// array is allocated with 64 bytes of slack and starts 256-byte aligned).
__device__ __forceinline__ const uint8_t* stage_bytes(uint8_t* dst, const uint8_t* src, uint32_t n)
{
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint4* s4 = reinterpret_cast<const uint4*>(src - mis);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const uint32_t nv = (mis + n + 15u) >> 4;
    for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) d4[i] = __ldg(s4 + i);
    return dst + mis;
}

// One block per chunk (kChunk forward positions of one read), 256 threads.  Inputs are staged in shared memory with
// 16-byte loads (the four kinetics planes are byte arrays at arbitrary alignment, two of them read backwards); outputs are
// written position-major so that a warp's stores are contiguous (8 B x 32 for kinf).
// bcode[B+p]  = forward-strand code at forward position p (flag 0x10: reverse complement of the stored SEQ)
// kinf[B+p]   = frames {fi[p], fp[p], ri[L-1-p], rp[L-1-p]}: the four kinetics of forward position p's base
//               pair, so that a window is one contiguous run in either strand direction.
__global__ void __launch_bounds__(kFrontThreads)
decode_kernel(const uint8_t* __restrict__ seq4, const uint8_t* __restrict__ fi, const uint8_t* __restrict__ fp,
              const uint8_t* __restrict__ ri, const uint8_t* __restrict__ rp, const uint32_t* __restrict__ base_off,
              const uint32_t* __restrict__ seq_off, const uint16_t* __restrict__ flag,
              const uint32_t* __restrict__ chunk_read, const uint32_t* __restrict__ chunk_pos,
              uint8_t* __restrict__ bcode, ushort4* __restrict__ kinf)
{
    __shared__ __align__(16) uint8_t s_buf[4][kChunk + 32];
    __shared__ __align__(16) uint8_t s_seq[kChunk / 2 + 48];
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t p0 = chunk_pos[blockIdx.x];
    const uint32_t B = base_off[r];
    const uint32_t L = base_off[r + 1] - B;
    const uint32_t S = seq_off[r];
    const bool rev = (flag[r] & 16) != 0;
    const uint32_t n = min((uint32_t)kChunk, L - p0);
    // forward planes: positions [p0, p0+n); reverse planes: indices L-1-p, i.e. [L-p0-n, L-p0)
    const uint32_t r0 = L - p0 - n;
    const uint8_t* s_fi = stage_bytes(s_buf[0], fi + B + p0, n);
    const uint8_t* s_fp = stage_bytes(s_buf[1], fp + B + p0, n);
    const uint8_t* s_ri = stage_bytes(s_buf[2], ri + B + r0, n);
    const uint8_t* s_rp = stage_bytes(s_buf[3], rp + B + r0, n);
    // stored-SEQ indices q: p (forward) or L-1-p (flag 0x10) -> nibble range [q0, q0+n)
    const uint32_t q0 = rev ? r0 : p0;
    const uint8_t* s_sq = stage_bytes(s_seq, seq4 + S + (q0 >> 1), ((q0 + n + 1) >> 1) - (q0 >> 1));
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += kFrontThreads) {
        const uint32_t p = p0 + i;
        const uint32_t q = rev ? L - 1 - p : p;  // index into the stored SEQ
        const uint32_t nib = (s_sq[(q >> 1) - (q0 >> 1)] >> ((~q & 1u) << 2)) & 0xfu;
        uint32_t code = nib_to_code(nib);
        if (rev && code < 4) code = 3u - code;
        bcode[B + p] = (uint8_t)code;
        ushort4 k;
        k.x = (unsigned short)codev1_frames(s_fi[i]);
        k.y = (unsigned short)codev1_frames(s_fp[i]);
        k.z = (unsigned short)codev1_frames(s_ri[n - 1 - i]);
        k.w = (unsigned short)codev1_frames(s_rp[n - 1 - i]);
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

// Shared staging of one chunk's forward-strand codes with a 2-position halo on both sides (0xff outside the read), so that
// every thread classifies 16 consecutive positions from registers.  s_code[2 + i] = code at chunk position i.
__device__ __forceinline__ void stage_codes(uint8_t (&s_code)[kChunk + 48], const uint8_t* __restrict__ codes, uint32_t p0, uint32_t n,
                                            uint32_t L)
{
    // positions [p0 - 2, p0 + n + 2) clipped to the read; shared index 2 + (p - p0)
    for (uint32_t i = threadIdx.x; i < n + 4; i += blockDim.x) {
        const int p = (int)p0 - 2 + (int)i;
        s_code[i] = (p >= 0 && p < (int)L) ? codes[p] : (uint8_t)0xff;
    }
}

// Site class from a thread-local window: w[2 + j] is the code at the thread's position j, w[0..1] / w[18..19] the halo.
// CpG: C,G.  CHG: C,[ACT],G.  CHH fwd: C,[ACT],[ACT].  CHH rev: [AGT],[AGT],G -> the G.  0xff (outside the read) and 4 (N) are
// "no base" (eval_kmer_features.cpp:67-126; the three sets are disjoint).
__device__ __forceinline__ int class_at(const uint8_t* w, int j, uint32_t ctx_mask)
{
    const uint32_t b = w[2 + j];
    if (b == 1u) {
        const uint32_t n1 = w[3 + j];
        if (n1 == 2u) return (ctx_mask & 1u) ? 0 : -1;
        if (n1 > 3u) return -1;
        const uint32_t n2 = w[4 + j];
        if (n2 == 2u) return (ctx_mask & 2u) ? 1 : -1;
        if (n2 > 3u) return -1;
        return (ctx_mask & 4u) ? 2 : -1;
    }
    if (b == 2u && (ctx_mask & 4u)) {
        const uint32_t m1 = w[1 + j], m2 = w[j];
        if (m1 > 3u || m2 > 3u || m1 == 1u || m2 == 1u) return -1;
        return 3;
    }
    return -1;
}

constexpr int kPosPerThread = kChunk / kFrontThreads;  // 16

// Loads the thread's 16 positions + halo from the staged chunk into w[20] and returns the packed classes:
// 4 bits per position (0..3, 0xf = none), position j in bits [4j, 4j+4).
__device__ __forceinline__ uint64_t thread_classes(const uint8_t (&s_code)[kChunk + 48], uint32_t n, uint32_t ctx_mask, bool valid)
{
    uint64_t packed = ~0ull;
    const uint32_t t0 = threadIdx.x * kPosPerThread;
    if (!valid || t0 >= n) return packed;
    uint8_t w[kPosPerThread + 4];
    #pragma unroll
    for (int i = 0; i < kPosPerThread + 4; ++i) w[i] = s_code[t0 + i];
    #pragma unroll
    for (int j = 0; j < kPosPerThread; ++j) {
        if (t0 + j >= n) break;
        const int c = class_at(w, j, ctx_mask);
        if (c >= 0) packed = (packed & ~(0xfull << (4 * j))) | ((uint64_t)c << (4 * j));
    }
    return packed;
}

// Pass 1: per-chunk class counts.  One block (256 threads x 16 positions) per chunk.
__global__ void __launch_bounds__(kFrontThreads)
scan_count_kernel(const uint8_t* __restrict__ bcode, const uint32_t* __restrict__ base_off,
                  const uint8_t* __restrict__ valid, const uint32_t* __restrict__ chunk_read,
                  const uint32_t* __restrict__ chunk_pos, uint32_t ctx_mask, ClassCount* __restrict__ chunk_cnt)
{
    __shared__ __align__(16) uint8_t s_code[kChunk + 48];
    __shared__ uint32_t s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t B = base_off[r];
    const uint32_t L = base_off[r + 1] - B;
    const uint32_t p0 = chunk_pos[blockIdx.x];
    const uint32_t n = min((uint32_t)kChunk, L - p0);
    stage_codes(s_code, bcode + B, p0, n, L);
    __syncthreads();
    const uint64_t cls = thread_classes(s_code, n, ctx_mask, valid[r] != 0);
    uint32_t c[4] = {0, 0, 0, 0};
    #pragma unroll
    for (int j = 0; j < kPosPerThread; ++j) {
        const uint32_t v = (uint32_t)(cls >> (4 * j)) & 0xfu;
        #pragma unroll
        for (int k = 0; k < 4; ++k) c[k] += (v == (uint32_t)k);
    }
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t v = c[k];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[k], v);
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
__global__ void __launch_bounds__(kFrontThreads)
scan_write_kernel(const uint8_t* __restrict__ bcode, const uint32_t* __restrict__ base_off,
                  const uint8_t* __restrict__ valid, const uint32_t* __restrict__ chunk_read,
                  const uint32_t* __restrict__ chunk_pos, const uint32_t* __restrict__ read_first_chunk,
                  const ClassCount* __restrict__ pref, uint32_t n_chunks, uint32_t ctx_mask,
                  int32_t* __restrict__ qoff, uint8_t* __restrict__ call_ctx, uint32_t* __restrict__ site_read,
                  uint32_t* __restrict__ site_pos, uint32_t* __restrict__ site_out)
{
    __shared__ __align__(16) uint8_t s_code[kChunk + 48];
    __shared__ uint32_t s_warp[kFrontThreads / 32][4];
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t B = base_off[r];
    const uint32_t L = base_off[r + 1] - B;
    const uint32_t p0 = chunk_pos[blockIdx.x];
    const uint32_t n = min((uint32_t)kChunk, L - p0);
    stage_codes(s_code, bcode + B, p0, n, L);
    __syncthreads();
    const uint64_t cls = thread_classes(s_code, n, ctx_mask, valid[r] != 0);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // exclusive rank of this thread's first site of each class inside the chunk: thread counts -> warp scan -> block scan
    uint32_t rk[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t c = 0;
        #pragma unroll
        for (int j = 0; j < kPosPerThread; ++j) c += (((uint32_t)(cls >> (4 * j)) & 0xfu) == (uint32_t)k);
        uint32_t incl = c;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp][k] = incl;
        rk[k] = incl - c;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        uint32_t run = 0;
        for (int w = 0; w < kFrontThreads / 32; ++w) {
            const uint32_t v = s_warp[w][threadIdx.x];
            s_warp[w][threadIdx.x] = run;
            run += v;
        }
    }
    __syncthreads();
    if (cls == ~0ull) return;
    #pragma unroll
    for (int k = 0; k < 4; ++k) rk[k] += s_warp[warp][k];

    const ClassCount pc = pref[blockIdx.x];
    const ClassCount q0 = pref[read_first_chunk[r]];
    const ClassCount q1 = pref[read_first_chunk[r + 1]];
    const ClassCount tot = pref[n_chunks];
    const uint32_t call_base = q0.c[0] + q0.c[1] + q0.c[2] + q0.c[3];
    const uint32_t fwd_before = (pc.c[0] + pc.c[1] + pc.c[2]) - (q0.c[0] + q0.c[1] + q0.c[2]);
    const uint32_t nfwd = (q1.c[0] + q1.c[1] + q1.c[2]) - (q0.c[0] + q0.c[1] + q0.c[2]);
    const uint32_t region[4] = {0, tot.c[0], tot.c[0] + tot.c[1], tot.c[0] + tot.c[1] + tot.c[2]};
    #pragma unroll
    for (int j = 0; j < kPosPerThread; ++j) {
        const uint32_t c = (uint32_t)(cls >> (4 * j)) & 0xfu;
        if (c > 3u) continue;
        const uint32_t p = p0 + threadIdx.x * kPosPerThread + j;
        uint32_t out;
        if (c < 3u) out = call_base + fwd_before + rk[0] + rk[1] + rk[2];
        else out = call_base + nfwd + (pc.c[3] - q0.c[3]) + rk[3];
        qoff[out] = (int32_t)p;
        call_ctx[out] = (uint8_t)(c == 3u ? 2u : c);
        const uint32_t slot = region[c] + pc.c[c] + rk[c];
        site_read[slot] = r;
        site_pos[slot] = p | (c == 3u ? 0x80000000u : 0u);
        site_out[slot] = out;
        ++rk[c];
    }
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
