// site_chain.cuh -- the compact (per-site) part of the dilated dense plan in ONE kernel: F2..F6, G2..G6, T7_0..3, T8_0..1 and the
// FC head of a tile of sites, with every intermediate map kept in shared memory.
//
// Launched op by op (round 1) the compact chain moved ~23 KB of HBM traffic per site -- every op wrote its 256..512 B/site map and
// the next one read it back, next to the scatter copies of the dense layers -- for 415 k MAC per site: each op was HBM-bound
// (53 B per tensor-core cycle and SM against ~30 the memory system delivers) and the chain took 24 ms of a 96 ms step at half of
// its tensor-core bound.  Here a CTA pair (cta_group::2, M = 256 = two tiles of 128 sites) walks the whole chain for its tiles:
//
//   resident   two 64 KiB buffers per CTA hold the running F and G maps (later T7 / T8) of the CTA's 128 sites, in the plane
//              layout the tensor core reads ([hl][g][128 rows][8] bf16): an op's epilogue overwrites its own input in place
//   streamed   through a ring of 8 slots, one slot per (op, 16-channel stage, term): the term's WEIGHT tiles (hi, lo; this CTA's
//              half of N; they come from L2 -- the whole model is 1 MB) and, for terms that read a scatter copy of a dense map or
//              F1 / G1, the 4 activation planes of the stage (8 KiB, HBM -> shared, 1-D bulk copies)
//   MMA warp   (leader CTA) per slot the split-precision triple hi*hi + lo*hi + hi*lo, M = 256; waits for the epilogue of the
//              ops an op depends on (mbarrier) before its first MMA, so F and G ops alternate: MMA(G_l) runs under epilogue(F_l)
//   epilogue   8 warps per CTA: TMEM -> + bias, ReLU -> hi/lo bf16 -> the resident buffer (fence.proxy.async, arrive); the head's
//              epilogue applies fc2 and writes the site's two logits
//
// HBM traffic per site: the scatter copies and F1 / G1 read once (~7.5 KB) and 8 B of logits written.  conv1-form ops (F1, G1:
// gathered from the X map) stay separate launches of dense_gemm_kernel.
#pragma once
#include "dense_gemm2.cuh"

namespace hm {

constexpr int kChainMaxOps = 20;
constexpr int kChainSlots = 8;                 // = kProducerWarps: ring slot s is always filled by producer warp s
constexpr uint32_t kChainSlabBytes = 8192;     // 4 planes x 128 rows x 16 B: one 16-channel stage of a streamed operand
constexpr uint32_t kChainSlotBytes = 12288;    // slab + weight tiles of N <= 128 (4 KiB); a term without slab may use all of it (N = 256: 8 KiB)
constexpr uint32_t kChainResBytes = 65536;     // one resident buffer: 128 rows x 128 channels x {hi, lo}
constexpr uint32_t kChainPlaneBytes = 2048;    // 128 rows x 16 B
constexpr int kChainRegions = 8;               // TMEM in 64-column regions
#ifndef HM_CHAIN_LOOKAHEAD
#define HM_CHAIN_LOOKAHEAD 6
#endif
constexpr int kChainLookahead = HM_CHAIN_LOOKAHEAD;
#ifndef HM_CHAIN_EXPERIMENT
#define HM_CHAIN_EXPERIMENT 0   // timing experiments only (results are wrong): 1 = epilogue does no work, 2 = no MMAs issued, 4 = no slab copies
#endif  // L2 prefetch distance in rounds of the ring (x 8 steps); 0 = none

struct ChainTerm {
    const uint8_t* src;   // streamed term: plane 0 (hi, g = 0), row 0 of the compact source map; resident term: nullptr
    uint32_t res_off;     // resident term: byte offset of the map (its hi plane g = 0) in the resident area
    int32_t dep;          // resident term: index of the chain op that produces the map
};

struct ChainOp {
    ChainTerm term[kMaxTerms];
    const uint8_t* w_img;   // pair lowering of lower_op(): [rank][stage][term][hl] tiles of (n / 2) x 32 bytes
    const float* bias;      // [n]
    uint8_t* spill;         // debug (hm_debug_dump_acts): the output map is ALSO stored here (compact map in HBM); else nullptr
    uint32_t out_off;       // resident byte offset of the output map
    uint32_t tmem_col;      // first accumulator column
    uint32_t regions;       // mask of the 64-column TMEM regions the accumulator covers
    int32_t n, cin, n_terms;
    int32_t head;           // 1: fc1 + ReLU -> fc2 -> logits instead of a resident map
    int32_t wait_op;        // the epilogue waits for the accumulator of THIS op (>= own index): an output that overwrites a buffer
                            // later MMAs still read waits for the last of them
};

struct ChainProgram {
    ChainOp op[kChainMaxOps];
    int32_t n_ops;
    uint32_t n_tiles;                  // 128-site tiles
    unsigned long long plane_stride;   // of every compact map (streamed sources and spill targets)
    // Every CTA pair streams the SAME weight tiles at about the same time: with one copy in memory 74 pairs hit the same L2 lines
    // together and the ring ran at ~460 cycles per slot with no MMAs and no epilogue work at all.  The model blob is therefore
    // replicated: pair p reads copy p % w_copies, w_copy_stride bytes apart.
    unsigned long long w_copy_stride;
    uint32_t w_copies;
    const float* w2;                   // head: [2][256], [2]
    const float* b2;
    float* logits;                     // [rows][2]
};

inline size_t chain_smem_bytes()
{
    return 2 * (size_t)kChainResBytes + (size_t)kChainSlots * kChainSlotBytes +
           (2 * kChainSlots + 2 * kChainMaxOps + kChainRegions + 1) * sizeof(uint64_t) + 16;
}

// hi / lo split of 8 ReLU'd values -> two 16-byte units (same arithmetic as epilogue_store_groups)
__device__ __forceinline__ void split_hilo8(const float* x, uint4& vh, uint4& vl)
{
    uint32_t hi[4], lo[4];
    #pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float x0 = x[2 * j], x1 = x[2 * j + 1];
        const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
        const __nv_bfloat162 e = __floats2bfloat162_rn(x0 - __uint_as_float(hb << 16), x1 - __uint_as_float(hb & 0xffff0000u));
        hi[j] = hb;
        lo[j] = *reinterpret_cast<const uint32_t*>(&e);
    }
    vh = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    vl = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDenseThreads, 1) site_chain_kernel(const __grid_constant__ ChainProgram prog)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = umma::cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    uint8_t* s_res = smem;                                   // [2][64 KiB]
    uint8_t* s_ring = smem + 2 * kChainResBytes;             // [8][12 KiB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)kChainSlots * kChainSlotBytes);
    uint64_t* full = bars;                                   // [8]   leader: own expect_tx arrive + the peer's relay; peer: own arrive
    uint64_t* empty = full + kChainSlots;                    // [8]   multicast commit from the leader
    uint64_t* acc_full = empty + kChainSlots;                // [ops] multicast commit: every MMA of the op (and before it) has completed
    uint64_t* res_ready = acc_full + kChainMaxOps;           // [ops] leader only: one arrive per epilogue warp of both CTAs
    uint64_t* acc_free = res_ready + kChainMaxOps;           // [8]   leader only: the TMEM region has been drained, same arrivals
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_free + kChainRegions);
    const int n_ops = prog.n_ops;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < kChainSlots; ++i) {
                umma::mbar_init(&full[i], rank == 0 ? 2u : 1u);
                umma::mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < kChainMaxOps; ++i) {
                umma::mbar_init(&acc_full[i], 1);
                umma::mbar_init(&res_ready[i], 2 * kEpilogueWarps);
            }
            for (int i = 0; i < kChainRegions; ++i) umma::mbar_init(&acc_free[i], 2 * kEpilogueWarps);
            umma::fence_barrier_init();
        }
        __syncwarp();
        umma::tmem_alloc2(s_tmem, 512);
    }
    umma::tc_fence_before();
    umma::cluster_sync();  // both CTAs' barriers exist before anyone arrives remotely
    umma::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp < (uint32_t)kProducerWarps) {
        // ===================================== producers: one ring slot per (op, stage, term) ================================
        // A slot carries only 3 MMAs (~230 cycles) of work, so the 8 slots in flight cover ~1 900 cycles -- less than an HBM round
        // trip plus the peer's relay (first version: ~410 cycles per step, the chain ran at half of its tensor bound).  Every warp
        // therefore also runs a second cursor kChainLookahead of ITS OWN steps ahead and pulls that step's slab into L2
        // (cp.async.bulk.prefetch.L2), so that the ring's copies are L2 hits; the weights are L2-resident anyway.
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the streamed maps come from earlier launches
        struct Cursor {
            uint32_t t2, s;
            int oi, k;
        };
        auto valid = [&](const Cursor& c) { return 2 * c.t2 < prog.n_tiles; };
        auto advance = [&](Cursor& c) {  // next (tile, op, stage, term) in issue order
            const ChainOp& op = prog.op[c.oi];
            if (++c.k < op.n_terms) return;
            c.k = 0;
            if (++c.s < ((uint32_t)op.cin >> 4)) return;
            c.s = 0;
            if (++c.oi < n_ops) return;
            c.oi = 0;
            c.t2 += n_pairs;
        };
        auto prefetch = [&](const Cursor& c) {
            const ChainOp& op = prog.op[c.oi];
            const uint8_t* src = op.term[c.k].src;
            if (!src || lane >= 4u) return;
            const unsigned long long row0 = (unsigned long long)(2 * c.t2 + rank) * kTileRows;
            const uint8_t* plane = src + (unsigned long long)((lane >> 1) * ((uint32_t)op.cin >> 3) + 2u * c.s + (lane & 1u)) * prog.plane_stride;
            umma::bulk_prefetch_l2(plane + row0 * 16ull, kChainPlaneBytes);
        };
        Cursor cur{pair, 0u, 0, 0}, ahead{pair, 0u, 0, 0};
        for (uint32_t i = 0; i < warp && valid(cur); ++i) advance(cur);  // this warp's first step
        ahead = cur;
        for (int a = 0; a < kChainLookahead && valid(ahead); ++a) {  // warm-up: this warp's first kChainLookahead steps
            prefetch(ahead);
            for (int i = 0; i < kChainSlots && valid(ahead); ++i) advance(ahead);
        }
        for (uint32_t step = warp; valid(cur); step += kChainSlots) {
            const ChainOp& op = prog.op[cur.oi];
            const uint32_t n_stages = (uint32_t)op.cin >> 4, groups = (uint32_t)op.cin >> 3;
            const uint32_t w_step = (uint32_t)op.n * 32u;                          // hi + lo tile of this CTA's half of N
            const uint8_t* w_rank = op.w_img + (unsigned long long)(pair % prog.w_copies) * prog.w_copy_stride + (size_t)rank * n_stages * (uint32_t)op.n_terms * w_step;
            const unsigned long long row0 = (unsigned long long)(2 * cur.t2 + rank) * kTileRows;  // odd n_tiles: the peer's last tile lies in the slack rows
            const uint32_t slot = warp, phase = (step / kChainSlots) & 1u;
            const uint8_t* src = op.term[cur.k].src;
            if (valid(ahead)) prefetch(ahead);
            umma::mbar_wait(&empty[slot], phase ^ 1u);
            uint8_t* dst = s_ring + (size_t)slot * kChainSlotBytes;
            if (lane == 0) umma::mbar_arrive_expect_tx(&full[slot], ((src && !(HM_CHAIN_EXPERIMENT & 4)) ? kChainSlabBytes : 0u) + w_step);
            __syncwarp();
            if (lane < 4u) {
                if (src && !(HM_CHAIN_EXPERIMENT & 4)) {  // planes {hi g0, hi g1, lo g0, lo g1} of the stage
                    const uint8_t* plane = src + (unsigned long long)((lane >> 1) * groups + 2u * cur.s + (lane & 1u)) * prog.plane_stride;
                    umma::bulk_g2s(dst + lane * kChainPlaneBytes, plane + row0 * 16ull, kChainPlaneBytes, &full[slot]);
                }
            } else if (lane < 8u) {  // the weight tiles in four pieces
                const uint32_t per = w_step >> 2, piece = lane - 4u;
                umma::bulk_g2s(dst + (src ? kChainSlabBytes : 0u) + piece * per,
                               w_rank + (size_t)(cur.s * (uint32_t)op.n_terms + (uint32_t)cur.k) * w_step + piece * per, per, &full[slot]);
            }
            __syncwarp();
            if (rank != 0) {
                // the peer tells the leader when its half of the slot has landed
                umma::mbar_wait(&full[slot], phase);
                if (lane == 0) umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&full[slot]), 0));
                __syncwarp();
            }
            for (int i = 0; i < kChainSlots && valid(cur); ++i) advance(cur);
            for (int i = 0; i < kChainSlots && valid(ahead); ++i) advance(ahead);
        }
    } else if (warp == (uint32_t)kProducerWarps) {
        if (rank == 0) {
            // ===================================== MMA issuer (leader) =========================================================
            const uint32_t desc_hi = (uint32_t)(umma::make_desc(0, 0, 128) >> 32);
            const uint32_t ring16 = umma::smem_u32(s_ring) >> 4, res16 = umma::smem_u32(s_res) >> 4;
            uint32_t uses[kChainRegions];
            #pragma unroll
            for (int r = 0; r < kChainRegions; ++r) uses[r] = 0;
            uint32_t step = 0, it = 0;
            for (uint32_t t2 = pair; 2 * t2 < prog.n_tiles; t2 += n_pairs, ++it) {
                for (int oi = 0; oi < n_ops; ++oi) {
                    const ChainOp& op = prog.op[oi];
                    const uint32_t n_stages = (uint32_t)op.cin >> 4, groups = (uint32_t)op.cin >> 3;
                    const uint32_t nh = (uint32_t)op.n >> 1;
                    const uint32_t idesc = umma::make_idesc_bf16_m256((uint32_t)op.n);
                    // inputs written by earlier epilogues (both CTAs), and the accumulator's previous contents drained
                    for (int k = 0; k < op.n_terms; ++k)
                        if (!op.term[k].src) umma::mbar_wait(&res_ready[op.term[k].dep], it & 1u);
                    #pragma unroll
                    for (int r = 0; r < kChainRegions; ++r) {
                        if ((op.regions >> r) & 1u) {
                            if (uses[r]) umma::mbar_wait(&acc_free[r], (uses[r] - 1u) & 1u);
                            ++uses[r];
                        }
                    }
                    umma::tc_fence_after();
                    const uint32_t d_addr = tmem_base + op.tmem_col;
                    uint32_t acc = 0;
                    for (uint32_t s = 0; s < n_stages; ++s) {
                        for (int k = 0; k < op.n_terms; ++k, ++step) {
                            const uint32_t slot = step & (kChainSlots - 1), phase = (step / kChainSlots) & 1u;
                            umma::mbar_wait(&full[slot], phase);
                            umma::tc_fence_after();
                            if (umma::elect_one()) {
                                const uint32_t sb16 = ring16 + slot * (kChainSlotBytes >> 4);
                                const bool stream = op.term[k].src != nullptr;
                                // A: hi view and lo view, K-adjacent core matrices one plane apart
                                const uint32_t a_addr16 = stream ? sb16 : res16 + ((op.term[k].res_off + 2u * s * kChainPlaneBytes) >> 4);
                                const uint32_t a_lo_off16 = stream ? (2u * kChainPlaneBytes) >> 4 : (groups * kChainPlaneBytes) >> 4;
                                const uint32_t a_hi = (uint32_t)umma::make_desc(0, kChainPlaneBytes, 128) + a_addr16;
                                // B: this CTA's nh rows of the hi tile, then of the lo tile
                                const uint32_t b_hi = (uint32_t)umma::make_desc(0, nh * 16u, 128) + sb16 + (stream ? (kChainSlabBytes >> 4) : 0u);
                                const uint32_t b_step = (nh * 32u) >> 4;
                                if (!(HM_CHAIN_EXPERIMENT & 2)) {
                                    umma::mma2_bf16_w(d_addr, a_hi, b_hi, desc_hi, idesc, acc);
                                    umma::mma2_bf16_w(d_addr, a_hi + a_lo_off16, b_hi, desc_hi, idesc, 1);
                                    umma::mma2_bf16_w(d_addr, a_hi, b_hi + b_step, desc_hi, idesc, 1);
                                }
                                umma::mma2_commit_mc(&empty[slot]);
                            }
                            acc = 1;
                            __syncwarp();
                        }
                    }
                    if (umma::elect_one()) umma::mma2_commit_mc(&acc_full[oi]);
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================================== epilogue (own CTA's 128 sites) ====================================================
        const uint32_t lane_grp = (warp & 3u) * 32u;
        const uint32_t half = (warp - (uint32_t)(kProducerWarps + 1)) >> 2;
        const uint32_t m = lane_grp + lane;
        uint32_t it = 0;
        for (uint32_t t2 = pair; 2 * t2 < prog.n_tiles; t2 += n_pairs, ++it) {
            const unsigned long long row = (unsigned long long)(2 * t2 + rank) * kTileRows + m;
            for (int oi = 0; oi < n_ops; ++oi) {
                const ChainOp& op = prog.op[oi];
                const int n = op.n;
                umma::mbar_wait(&acc_full[op.wait_op], it & 1u);
                umma::tc_fence_after();
                const uint32_t t_addr = tmem_base + (lane_grp << 16) + op.tmem_col;
                if (HM_CHAIN_EXPERIMENT & 1) {
                } else if (op.head) {
                    // fc1 + ReLU from the accumulator, fc2 as two dot products per site (the first warp of every lane group does all
                    // the columns, as in dense_gemm_kernel's head form: once per tile, off the critical path)
                    if (half == 0) {
                        float l0 = 0.f, l1 = 0.f;
                        for (int c0 = 0; c0 < n; c0 += 32) {
                            uint32_t v[32];
                            umma::tmem_ld32(t_addr + (uint32_t)c0, v);
                            umma::tmem_ld_wait();
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(op.bias + c0 + j));
                                const float4 w0 = __ldg(reinterpret_cast<const float4*>(prog.w2 + c0 + j));
                                const float4 w1 = __ldg(reinterpret_cast<const float4*>(prog.w2 + n + c0 + j));
                                const float f0 = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f), f1 = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                                const float f2 = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f), f3 = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
                                l0 = fmaf(f0, w0.x, l0); l0 = fmaf(f1, w0.y, l0); l0 = fmaf(f2, w0.z, l0); l0 = fmaf(f3, w0.w, l0);
                                l1 = fmaf(f0, w1.x, l1); l1 = fmaf(f1, w1.y, l1); l1 = fmaf(f2, w1.z, l1); l1 = fmaf(f3, w1.w, l1);
                            }
                        }
                        *reinterpret_cast<float2*>(prog.logits + row * 2ull) = make_float2(l0 + __ldg(prog.b2), l1 + __ldg(prog.b2 + 1));
                    }
                } else {
                    const int mid = ((n >> 1) + 15) & ~15;
                    int c0 = half ? mid : 0;
                    const int c1 = half ? n : mid;
                    const uint32_t out_groups = (uint32_t)n >> 3;
                    uint8_t* out = s_res + op.out_off + m * 16u;
                    while (c0 < c1) {
                        const int nc = (c1 - c0 >= 32) ? 32 : 16;
                        float f[32];
                        if (nc == 32) {
                            uint32_t v[32];
                            umma::tmem_ld32(t_addr + (uint32_t)c0, v);
                            umma::tmem_ld_wait();
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(op.bias + c0 + j));
                                f[j] = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f);
                                f[j + 1] = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                                f[j + 2] = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f);
                                f[j + 3] = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
                            }
                        } else {
                            uint32_t v[16];
                            umma::tmem_ld16(t_addr + (uint32_t)c0, v);
                            umma::tmem_ld_wait();
                            #pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(op.bias + c0 + j));
                                f[j] = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f);
                                f[j + 1] = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                                f[j + 2] = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f);
                                f[j + 3] = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
                            }
                        }
                        const uint32_t g0 = (uint32_t)c0 >> 3;
                        #pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (8 * g < nc) {
                                uint4 vh, vl;
                                split_hilo8(f + 8 * g, vh, vl);
                                *reinterpret_cast<uint4*>(out + (g0 + g) * kChainPlaneBytes) = vh;
                                *reinterpret_cast<uint4*>(out + (out_groups + g0 + g) * kChainPlaneBytes) = vl;
                                if (op.spill) {
                                    uint8_t* q = op.spill + (unsigned long long)(g0 + g) * prog.plane_stride + row * 16ull;
                                    *reinterpret_cast<uint4*>(q) = vh;
                                    *reinterpret_cast<uint4*>(q + (unsigned long long)out_groups * prog.plane_stride) = vl;
                                }
                            }
                        }
                        c0 += nc;
                    }
                    umma::fence_proxy_async();  // the tensor core (async proxy) reads what these threads just wrote
                }
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) {
                        umma::mbar_arrive(&res_ready[oi]);
                        #pragma unroll
                        for (int r = 0; r < kChainRegions; ++r)
                            if ((op.regions >> r) & 1u) umma::mbar_arrive(&acc_free[r]);
                    } else {
                        umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&res_ready[oi]), 0));
                        #pragma unroll
                        for (int r = 0; r < kChainRegions; ++r)
                            if ((op.regions >> r) & 1u) umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&acc_free[r]), 0));
                    }
                }
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync();  // the peer's shared memory and TMEM stay alive until the leader's last MMA and arrive are done
    umma::tc_fence_after();
    if (warp == 0) umma::tmem_dealloc2(tmem_base, 512);
}

}  // namespace hm
