// kernels_front.cuh -- sm_100a kernels for the integer/byte front of the hot path:
//   decode_kernel        A1 + A2: packed 4-bit SEQ -> forward-strand base codes; fi/fp/ri/rp CodecV1 codes
//                        -> frames, re-packed per forward position (HBM bound: 4.5 B/base in, 9 B/base out); while the codes
//                        are in shared memory it also classifies every position (first half of A3: 0.5 B/base of class
//                        nibbles + per-chunk class counts)
//   scan_offsets_kernel / scan_write_kernel   second half of A3: prefix over chunks, then the site lists in final order
//                        (HBM bound: 0.5 B/base in, 4 B/site + 1 B/site + 12 B/site out)
//   gather_features_kernel  A4: [n,401,8] f32 windows (validation / microbench layout, 12 832 B/site)
//
// Reference semantics: src/corelib/bam_info.cpp:169-222,520-570; src/app/hifimeth/eval_kmer_features.cpp:9-126;
// src/corelib/5mc_context.cpp:4-10.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "umma.cuh"

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

// 8 bytes at byte offset `off` (any alignment) of a shared array given as words: three aligned loads and two funnel shifts.
__device__ __forceinline__ uint2 lds8(const uint32_t* words, uint32_t off)
{
    const uint32_t* w = words + (off >> 2);
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

template <int K>
__device__ __forceinline__ uint32_t byte_of(const uint2& v)
{
    return K < 4 ? ((v.x >> (8 * (K & 3))) & 0xffu) : ((v.y >> (8 * (K & 3))) & 0xffu);
}

// CodecV1 frames + 64 (the 64 is taken off both halves of a packed pair at once): ((c & 63) | 64) << (c >> 6)
__device__ __forceinline__ uint32_t codev1_biased(uint32_t c) { return ((c & 63u) | 64u) << (c >> 6); }

// BAM nibble -> base code through a 16-entry table held in a 64-bit constant (4 bits per entry): 1 A, 2 C, 4 G, 8 T -> 0..3
// (kNibFwd) or the complement 3..0 (kNibRev); everything else -> kCodeN.
constexpr uint64_t nib_table(bool complement)
{
    uint64_t t = 0;
    for (int n = 0; n < 16; ++n) {
        uint64_t c = n == 1 ? 0 : n == 2 ? 1 : n == 4 ? 2 : n == 8 ? 3 : kCodeN;
        if (complement && c < 4) c = 3 - c;
        t |= c << (4 * n);
    }
    return t;
}
constexpr uint64_t kNibFwd = nib_table(false), kNibRev = nib_table(true);

constexpr int kPosPerThread = kChunk / kFrontThreads;  // 16 positions per thread in the classification phase
constexpr int kCodePad = 16;                           // s_code[kCodePad + a + i] = code at chunk position i (a = global misalignment)

// Site class from a window of codes: w(j) is the code at the thread's position j, j in [-2, 17].
// CpG: C,G.  CHG: C,[ACT],G.  CHH fwd: C,[ACT],[ACT].  CHH rev: [AGT],[AGT],G -> the G.  0xff (outside the read) and 4 (N) are
// "no base" (eval_kmer_features.cpp:67-126; the three sets are disjoint).
__device__ __forceinline__ int class_of(uint32_t m2, uint32_t m1, uint32_t b, uint32_t n1, uint32_t n2, uint32_t ctx_mask)
{
    if (b == 1u) {
        if (n1 == 2u) return (ctx_mask & 1u) ? 0 : -1;
        if (n1 > 3u) return -1;
        if (n2 == 2u) return (ctx_mask & 2u) ? 1 : -1;
        if (n2 > 3u) return -1;
        return (ctx_mask & 4u) ? 2 : -1;
    }
    if (b == 2u && (ctx_mask & 4u)) {
        if (m1 > 3u || m2 > 3u || m1 == 1u || m2 == 1u) return -1;
        return 3;
    }
    return -1;
}

// Number of nibbles of `cls` equal to k (k in 0..3; 0xf = no site).
__device__ __forceinline__ uint32_t count_class(uint64_t cls, uint32_t k)
{
    uint64_t x = cls ^ (0x1111111111111111ull * k);
    x |= x >> 1;
    x |= x >> 2;
    return (uint32_t)__popcll(~x & 0x1111111111111111ull);
}

// A1 + A2 + the classification half of A3, one block per chunk (kChunk forward positions of one read), 256 threads.
//   bcode[B+p]  = forward-strand code at forward position p (flag 0x10: reverse complement of the stored SEQ)
//   kinf[B+p]   = frames {fi[p], fp[p], ri[L-1-p], rp[L-1-p]}: the four kinetics of forward position p's base pair, so that a
//                 window is one contiguous run in either strand direction
//   cls[chunk*256 + t] = site classes of chunk positions 16t .. 16t+15, 4 bits each (0 CpG, 1 CHG, 2 CHH fwd, 3 CHH rev, f none)
//   chunk_cnt[chunk]   = sites per class in the chunk
// Inputs are staged in shared memory with 16-byte loads (the kinetics planes are byte arrays at arbitrary alignment, two of
// them read backwards).  Decode: a thread takes 8 consecutive positions whose GLOBAL index starts at a multiple of 8, so codes
// leave as one 8-byte store and kinetics as four 16-byte stores; the first and last group of a chunk may be partial.
__global__ void __launch_bounds__(kFrontThreads)
decode_kernel(const uint8_t* __restrict__ seq4, const uint8_t* __restrict__ fi, const uint8_t* __restrict__ fp,
              const uint8_t* __restrict__ ri, const uint8_t* __restrict__ rp, const uint32_t* __restrict__ base_off,
              const uint32_t* __restrict__ seq_off, const uint16_t* __restrict__ flag, const uint8_t* __restrict__ valid,
              const uint32_t* __restrict__ chunk_read, const uint32_t* __restrict__ chunk_pos, uint32_t ctx_mask,
              uint8_t* __restrict__ bcode, ushort4* __restrict__ kinf, ClassCount* __restrict__ chunk_cnt, uint64_t* __restrict__ cls_out)
{
    constexpr uint32_t kPlane = kChunk + 48;                 // staged bytes of one kinetics plane (16-byte units + misalignment)
    constexpr uint32_t kSeqOff = 4 * kPlane;                 // packed SEQ after the four planes
    __shared__ __align__(16) uint8_t s_in[4 * kPlane + kChunk / 2 + 64];
    __shared__ __align__(16) uint8_t s_code[kCodePad + 8 + kChunk + 24];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_cnt[4];
    const uint32_t lane = threadIdx.x & 31u;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        umma::mbar_init(&s_bar, 1);
        umma::fence_barrier_init();
    }
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t p0 = chunk_pos[blockIdx.x];
    const uint32_t B = base_off[r];
    const uint32_t L = base_off[r + 1] - B;
    const uint32_t S = seq_off[r];
    const bool rev = (flag[r] & 16) != 0;
    const uint32_t n = min((uint32_t)kChunk, L - p0);
    // forward planes: positions [p0, p0+n); reverse planes: indices L-1-p, i.e. [L-p0-n, L-p0)
    const uint32_t r0 = L - p0 - n;
    // SEQ with a halo of two positions on both sides (clipped to the read): forward positions [pa, pb) are stored indices
    // [pa, pb) or, under flag 0x10, [L-pb, L-pa)
    const uint32_t pa = p0 >= 2u ? p0 - 2u : 0u, pb = min(L, p0 + n + 2u);
    const uint32_t qa = rev ? L - pb : pa;
    // ---- staging: five 1-D bulk copies on the TMA engine (16-byte aligned source ranges covering the bytes wanted; every
    // input array has 64 bytes of slack), issued by five lanes in ONE instruction, completion on an mbarrier -----------------
    const uint8_t* src[5] = {fi + B + p0, fp + B + p0, ri + B + r0, rp + B + r0, seq4 + S + (qa >> 1)};
    const uint32_t len[5] = {n, n, n, n, ((qa + (pb - pa) + 1u) >> 1) - (qa >> 1)};
    uint32_t off[5];  // byte offset in s_in of the first wanted byte of each array
    uint32_t total = 0;
    #pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src[k]) & 15u);
        off[k] = (k < 4 ? (uint32_t)k * kPlane : kSeqOff) + mis;
        total += (mis + len[k] + 15u) & ~15u;
    }
    __syncthreads();  // barrier initialised
    if (threadIdx.x < 32) {
        if (lane == 0) umma::mbar_arrive_expect_tx(&s_bar, total);
        __syncwarp();
        #pragma unroll
        for (int k = 0; k < 5; ++k)
            if (lane == (uint32_t)k) {
                const uint32_t mis = off[k] & 15u;  // plane bases are multiples of 16
                umma::bulk_g2s(s_in + off[k] - mis, src[k] - mis, (mis + len[k] + 15u) & ~15u, &s_bar);
            }
    }
    umma::mbar_wait(&s_bar, 0);
    const uint32_t* s_w = reinterpret_cast<const uint32_t*>(s_in);
    const uint32_t o_fi = off[0], o_fp = off[1], o_ri = off[2], o_rp = off[3], o_sq = off[4];
    const uint64_t lut = rev ? kNibRev : kNibFwd;
    // code of forward position p (pa <= p < pb) from the staged SEQ
    auto code_at = [&](uint32_t p) -> uint32_t {
        const uint32_t q = rev ? L - 1u - p : p;
        const uint32_t nib = (s_in[o_sq + (q >> 1) - (qa >> 1)] >> ((~q & 1u) << 2)) & 0xfu;
        return (uint32_t)(lut >> (4u * nib)) & 0xfu;
    };
    const uint32_t a = (B + p0) & 7u;           // group g covers chunk positions [8g - a, 8g - a + 8)
    const uint32_t o_code = kCodePad + a;       // s_code[o_code + i] = code at chunk position i; group starts are 8-byte aligned
    const uint32_t n_groups = (n + a + 7u) >> 3;
    for (uint32_t g = threadIdx.x; g < n_groups; g += kFrontThreads) {
        const int i0 = (int)(8u * g) - (int)a;
        if (i0 >= 0 && (uint32_t)i0 + 8u <= n) {
            // ---- SEQ: 8 consecutive stored nibbles starting at q_lo, as one 64-bit nibble stream in index order --------
            const uint32_t q_lo = rev ? L - 1u - (p0 + (uint32_t)i0 + 7u) : p0 + (uint32_t)i0;
            const uint2 sv = lds8(s_w, o_sq + (q_lo >> 1) - (qa >> 1));
            uint64_t x = ((uint64_t)sv.y << 32) | sv.x;
            x = ((x & 0x0f0f0f0f0f0f0f0full) << 4) | ((x >> 4) & 0x0f0f0f0f0f0f0f0full);  // even index (high nibble) first
            const uint32_t nibs = (uint32_t)(x >> (4u * (q_lo & 1u)));                      // nibble k = stored index q_lo + k
            uint32_t c[8];
            #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t nib = (nibs >> (4 * (rev ? 7 - j : j))) & 0xfu;
                c[j] = (uint32_t)(lut >> (4u * nib)) & 0xfu;
            }
            const uint2 cv = make_uint2(c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24), c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24));
            *reinterpret_cast<uint2*>(s_code + o_code + i0) = cv;
            *reinterpret_cast<uint2*>(bcode + B + p0 + i0) = cv;
            // ---- kinetics: forward planes ascending, reverse planes descending ---------------------------------------------
            const uint2 vfi = lds8(s_w, o_fi + (uint32_t)i0), vfp = lds8(s_w, o_fp + (uint32_t)i0);
            const uint2 vri = lds8(s_w, o_ri + (n - 8u - (uint32_t)i0)), vrp = lds8(s_w, o_rp + (n - 8u - (uint32_t)i0));
            uint32_t w[16];
            #define HM_KIN(J)                                                                                              \
                w[2 * J] = (codev1_biased(byte_of<J>(vfi)) | (codev1_biased(byte_of<J>(vfp)) << 16)) - 0x00400040u;       \
                w[2 * J + 1] = (codev1_biased(byte_of<7 - J>(vri)) | (codev1_biased(byte_of<7 - J>(vrp)) << 16)) - 0x00400040u;
            HM_KIN(0) HM_KIN(1) HM_KIN(2) HM_KIN(3) HM_KIN(4) HM_KIN(5) HM_KIN(6) HM_KIN(7)
            #undef HM_KIN
            uint4* kd = reinterpret_cast<uint4*>(kinf + B + p0 + i0);
            #pragma unroll
            for (int j = 0; j < 4; ++j) kd[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        } else {
            // partial group at a chunk edge: position by position
            for (int i = max(i0, 0); i < i0 + 8 && (uint32_t)i < n; ++i) {
                const uint32_t code = code_at(p0 + (uint32_t)i);
                s_code[o_code + i] = (uint8_t)code;
                bcode[B + p0 + i] = (uint8_t)code;
                ushort4 k;
                k.x = (unsigned short)codev1_frames(s_in[o_fi + i]);
                k.y = (unsigned short)codev1_frames(s_in[o_fp + i]);
                k.z = (unsigned short)codev1_frames(s_in[o_ri + n - 1 - i]);
                k.w = (unsigned short)codev1_frames(s_in[o_rp + n - 1 - i]);
                kinf[B + p0 + i] = k;
            }
        }
    }
    // halo codes (0xff outside the read)
    if (threadIdx.x < 4) {
        const int i = threadIdx.x < 2 ? (int)threadIdx.x - 2 : (int)n + (int)threadIdx.x - 2;
        const long long p = (long long)p0 + i;
        s_code[(int)o_code + i] = (p >= 0 && p < (long long)L) ? (uint8_t)code_at((uint32_t)p) : (uint8_t)0xff;
    }
    __syncthreads();
    // ---- classification: 16 consecutive positions per thread from a 20-byte window, four positions per step on byte lanes ----
    // Per code a flag byte {bit0 C, bit1 G, bit2 H = A|C|T, bit3 D = A|G|T} (0 for N and outside the read), looked up with one
    // byte-permute per word; neighbours are funnel shifts of the flag words.  CpG: C,G.  CHG: C,H,G.  CHH fwd: C,H,H.
    // CHH rev: D,D,G -> the G (eval_kmer_features.cpp:67-126; the sets are disjoint).
    uint64_t cls = ~0ull;
    const uint32_t t0 = threadIdx.x * kPosPerThread;
    if (valid[r] != 0 && t0 < n) {
        const uint32_t wo = o_code + t0 - 2u;  // byte offset of the window in s_code: window byte i = position t0 - 2 + i
        const uint32_t* cw = reinterpret_cast<const uint32_t*>(s_code) + (wo >> 2);
        const uint32_t sh = (wo & 3u) * 8u;
        uint32_t t[6], f[5];
        #pragma unroll
        for (int k = 0; k < 6; ++k) t[k] = cw[k];
        #pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t v = __funnelshift_r(t[k], t[k + 1], sh) & 0x0f0f0f0fu;
            uint32_t sel = (v | (v >> 4)) & 0x00ff00ffu;           // low nibbles of the four codes -> one selector
            sel = (sel | (sel >> 8)) & 0xffffu;
            f[k] = __byte_perm(0x0c0a050cu, 0u, sel);             // A 0c, C 05, G 0a, T 0c; 4..7 -> 0; >= 8 (0xff) -> 0
        }
        const uint32_t e0 = (ctx_mask & 1u) ? 0x01010101u : 0u, e1 = (ctx_mask & 2u) ? 0x01010101u : 0u, e2 = (ctx_mask & 4u) ? 0x01010101u : 0u;
        uint32_t lo = 0, hi = 0;
        #pragma unroll
        for (int m = 0; m < 4; ++m) {  // positions t0 + 4m .. + 3 = window bytes 4m + 2 .. 4m + 5
            const uint32_t m2 = f[m], m1 = __funnelshift_r(f[m], f[m + 1], 8), s0 = __funnelshift_r(f[m], f[m + 1], 16);
            const uint32_t p1 = __funnelshift_r(f[m], f[m + 1], 24), p2 = f[m + 1];
            const uint32_t cpg = s0 & (p1 >> 1) & e0;
            const uint32_t ch = s0 & (p1 >> 2);
            const uint32_t chg = ch & (p2 >> 1) & e1;
            const uint32_t chhf = ch & (p2 >> 2) & e2;
            const uint32_t chhr = (s0 >> 1) & (m1 >> 3) & (m2 >> 3) & e2;
            const uint32_t code = (chg | chhr) | ((chhf | chhr) << 1);
            const uint32_t none = ~(cpg | chg | chhf | chhr) & 0x01010101u;
            uint32_t nib = code | (none * 15u);                      // one nibble value per byte lane
            nib = (nib | (nib >> 4)) & 0x00ff00ffu;
            nib = (nib | (nib >> 8)) & 0xffffu;
            if (m < 2) lo |= nib << (16 * m);
            else hi |= nib << (16 * (m - 2));
        }
        cls = ((uint64_t)hi << 32) | lo;
        if (t0 + kPosPerThread > n) cls |= ~0ull << (4u * (n - t0));  // positions past the end of the read
    }
    cls_out[(size_t)blockIdx.x * kFrontThreads + threadIdx.x] = cls;
    #pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        uint32_t v = count_class(cls, k);
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&s_cnt[k], v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ClassCount cc;
        cc.c[0] = s_cnt[0]; cc.c[1] = s_cnt[1]; cc.c[2] = s_cnt[2]; cc.c[3] = s_cnt[3];
        chunk_cnt[blockIdx.x] = cc;
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

// Pass 2: exclusive prefix over chunks (single block; n_chunks is ~ bases / 4096), then the per-read tables.  pref has
// n_chunks + 1 entries; pref[n_chunks] = totals.  totals[0..3] = class totals, totals[4] = n_calls (read back by the host).
// Per read: call_off (first call of the read in hm_call_batch order), n_fwd, read_pref (per-class sites in earlier reads).
__global__ void __launch_bounds__(1024)
scan_offsets_kernel(const ClassCount* __restrict__ chunk_cnt, uint32_t n_chunks, const uint32_t* __restrict__ read_first_chunk,
                    uint32_t n_reads, ClassCount* __restrict__ pref, uint32_t* __restrict__ totals, uint32_t* __restrict__ call_off,
                    uint32_t* __restrict__ n_fwd, uint32_t* __restrict__ read_pref)
{
    __shared__ ClassCount s_warp[32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t per = (n_chunks + 1023u) / 1024u;
    const uint32_t lo = min(threadIdx.x * per, n_chunks);
    const uint32_t hi = min(lo + per, n_chunks);
    ClassCount sum = {{0, 0, 0, 0}};
    for (uint32_t i = lo; i < hi; ++i) {
        const ClassCount v = chunk_cnt[i];
        #pragma unroll
        for (int k = 0; k < 4; ++k) sum.c[k] += v.c[k];
    }
    // inclusive scan of the 1024 partials: shuffles inside a warp, one shared round over the 32 warp totals
    ClassCount inc = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc.c[k], o);
            if ((int)lane >= o) inc.c[k] += t;
        }
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        ClassCount w = s_warp[lane], wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi.c[k], o);
                if ((int)lane >= o) wi.c[k] += t;
            }
        }
        #pragma unroll
        for (int k = 0; k < 4; ++k) wi.c[k] -= w.c[k];  // exclusive
        s_warp[lane] = wi;
    }
    __syncthreads();
    ClassCount run = s_warp[warp];
    #pragma unroll
    for (int k = 0; k < 4; ++k) run.c[k] += inc.c[k] - sum.c[k];
    for (uint32_t i = lo; i < hi; ++i) {
        pref[i] = run;
        const ClassCount v = chunk_cnt[i];
        #pragma unroll
        for (int k = 0; k < 4; ++k) run.c[k] += v.c[k];
    }
    if (threadIdx.x == 1023) {
        pref[n_chunks] = run;
        totals[0] = run.c[0]; totals[1] = run.c[1]; totals[2] = run.c[2]; totals[3] = run.c[3];
        totals[4] = run.c[0] + run.c[1] + run.c[2] + run.c[3];
    }
    __syncthreads();  // pref is complete and visible to the block
    for (uint32_t r = threadIdx.x; r <= n_reads; r += 1024u) {
        const ClassCount a = pref[read_first_chunk[r]];  // read_first_chunk[n_reads] = n_chunks
        reinterpret_cast<ClassCount*>(read_pref)[r] = a;
        call_off[r] = a.c[0] + a.c[1] + a.c[2] + a.c[3];
        if (r < n_reads) {
            const ClassCount b = pref[read_first_chunk[r + 1]];
            n_fwd[r] = (b.c[0] + b.c[1] + b.c[2]) - (a.c[0] + a.c[1] + a.c[2]);
        }
    }
}

// Pass 3: write, from the class nibbles decode_kernel left (0.5 B/base).  Output order per read: forward-strand calls (classes
// 0,1,2) ascending, then reverse (class 3) ascending -- the order the worker thread hands to build_one_mod_bam
// (mod_main.cpp:217-251).  Per-class work lists (site_read / site_pos / site_out) are laid out class after class.
// One block per chunk; warp w owns chunk positions [512w, 512w + 512) and walks them 32 at a time with lane = position, so
// that the ranks come from ballots and every store instruction of the warp writes one dense ascending run.
__global__ void __launch_bounds__(kFrontThreads)
scan_write_kernel(const uint64_t* __restrict__ cls_in, const uint32_t* __restrict__ chunk_read, const uint32_t* __restrict__ chunk_pos,
                  const uint32_t* __restrict__ read_first_chunk, const ClassCount* __restrict__ pref, uint32_t n_chunks,
                  int32_t* __restrict__ qoff, uint8_t* __restrict__ call_ctx, uint32_t* __restrict__ site_read,
                  uint32_t* __restrict__ site_pos, uint32_t* __restrict__ site_out)
{
    __shared__ uint32_t s_warp[kFrontThreads / 32][4];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t cls = cls_in[(size_t)blockIdx.x * kFrontThreads + threadIdx.x];  // positions 512*warp + 16*lane .. + 15
    uint32_t c[4];
    #pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        uint32_t v = count_class(cls, k);
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        c[k] = v;
    }
    if (lane == 0) { s_warp[warp][0] = c[0]; s_warp[warp][1] = c[1]; s_warp[warp][2] = c[2]; s_warp[warp][3] = c[3]; }
    __syncthreads();
    if (c[0] + c[1] + c[2] + c[3] == 0) return;  // warp-uniform: nothing to write in this warp's span
    uint32_t run[4] = {0, 0, 0, 0};               // sites of each class in the chunk before this warp's span
    for (uint32_t w = 0; w < warp; ++w) {
        #pragma unroll
        for (int k = 0; k < 4; ++k) run[k] += s_warp[w][k];
    }
    const uint32_t r = chunk_read[blockIdx.x];
    const uint32_t p0 = chunk_pos[blockIdx.x];
    const ClassCount pc = pref[blockIdx.x];
    const ClassCount q0 = pref[read_first_chunk[r]];
    const ClassCount q1 = pref[read_first_chunk[r + 1]];
    const ClassCount tot = pref[n_chunks];
    const uint32_t call_base = q0.c[0] + q0.c[1] + q0.c[2] + q0.c[3];
    const uint32_t nfwd = (q1.c[0] + q1.c[1] + q1.c[2]) - (q0.c[0] + q0.c[1] + q0.c[2]);
    // running output cursors of this warp
    uint32_t fwd_out = call_base + (pc.c[0] + pc.c[1] + pc.c[2]) - (q0.c[0] + q0.c[1] + q0.c[2]) + run[0] + run[1] + run[2];
    uint32_t rev_out = call_base + nfwd + (pc.c[3] - q0.c[3]) + run[3];
    uint32_t slot[4] = {pc.c[0] + run[0], tot.c[0] + pc.c[1] + run[1], tot.c[0] + tot.c[1] + pc.c[2] + run[2],
                        tot.c[0] + tot.c[1] + tot.c[2] + pc.c[3] + run[3]};
    const uint32_t lo32 = (uint32_t)cls, hi32 = (uint32_t)(cls >> 32);
    const uint32_t lt = (1u << lane) - 1u;
    #pragma unroll 4
    for (int j = 0; j < 16; ++j) {
        // position 512*warp + 32*j + lane is nibble (lane & 15) of the word held by lane 2*j + (lane >> 4)
        const int src = 2 * j + (int)(lane >> 4);
        const uint32_t wl = __shfl_sync(0xffffffffu, lo32, src), wh = __shfl_sync(0xffffffffu, hi32, src);
        const uint32_t k = (((lane & 8u) ? wh : wl) >> (4u * (lane & 7u))) & 0xfu;
        const uint32_t m0 = __ballot_sync(0xffffffffu, k == 0u), m1 = __ballot_sync(0xffffffffu, k == 1u);
        const uint32_t m2 = __ballot_sync(0xffffffffu, k == 2u), m3 = __ballot_sync(0xffffffffu, k == 3u);
        const uint32_t mf = m0 | m1 | m2;
        if (k < 4u) {
            const uint32_t p = p0 + 512u * warp + 32u * (uint32_t)j + lane;
            const uint32_t mk = k == 0u ? m0 : k == 1u ? m1 : k == 2u ? m2 : m3;
            const uint32_t out = k < 3u ? fwd_out + __popc(mf & lt) : rev_out + __popc(m3 & lt);
            const uint32_t sl = (k == 0u ? slot[0] : k == 1u ? slot[1] : k == 2u ? slot[2] : slot[3]) + __popc(mk & lt);
            qoff[out] = (int32_t)p;
            call_ctx[out] = (uint8_t)(k == 3u ? 2u : k);
            site_read[sl] = r;
            site_pos[sl] = p | (k == 3u ? 0x80000000u : 0u);
            site_out[sl] = out;
        }
        fwd_out += __popc(mf);
        rev_out += __popc(m3);
        slot[0] += __popc(m0); slot[1] += __popc(m1); slot[2] += __popc(m2); slot[3] += __popc(m3);
    }
}

// A9 (north_star "per-read normalisation"; DIAGNOSTICS ONLY -- the reference normalises by the fixed 952, src/corelib/
// bam_info.hpp:108, and nothing computed here is fed to the model): per read the sum and maximum of the decoded frames of each
// kinetics plane (fi, fp, ri, rp), from which the caller gets per-read means.  One block per chunk (the decode kernel's chunk
// table): coalesced 8-byte loads, warp-shuffle reduction, one set of atomics per block into the read's record (zeroed by the
// caller).  HBM bound: 8 B/base in.
struct __align__(8) ReadKinStats {
    unsigned long long sum[4];
    uint32_t max[4];
};

__global__ void __launch_bounds__(kFrontThreads)
read_stats_kernel(const ushort4* __restrict__ kinf, const uint32_t* __restrict__ base_off, const uint32_t* __restrict__ chunk_read,
                  const uint32_t* __restrict__ chunk_pos, ReadKinStats* __restrict__ out)
{
    __shared__ unsigned long long s_sum[kFrontThreads / 32][4];
    __shared__ uint32_t s_max[kFrontThreads / 32][4];
    const uint32_t r = chunk_read[blockIdx.x], p0 = chunk_pos[blockIdx.x];
    const uint32_t B = base_off[r], L = base_off[r + 1] - B;
    const uint32_t n = min((uint32_t)kChunk, L - p0);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t s[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0};  // <= 16 values of <= 952 per thread: 32-bit partial sums
    for (uint32_t i = threadIdx.x; i < n; i += kFrontThreads) {
        const ushort4 k = kinf[B + p0 + i];
        s[0] += k.x; s[1] += k.y; s[2] += k.z; s[3] += k.w;
        m[0] = max(m[0], (uint32_t)k.x); m[1] = max(m[1], (uint32_t)k.y); m[2] = max(m[2], (uint32_t)k.z); m[3] = max(m[3], (uint32_t)k.w);
    }
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
            m[k] = max(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
        }
        if (lane == 0) { s_sum[warp][k] = s[k]; s_max[warp][k] = m[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long ts = 0;
        uint32_t tm = 0;
        for (int w = 0; w < kFrontThreads / 32; ++w) { ts += s_sum[w][threadIdx.x]; tm = max(tm, s_max[w][threadIdx.x]); }
        atomicAdd(&out[r].sum[threadIdx.x], ts);
        atomicMax(&out[r].max[threadIdx.x], tm);
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
