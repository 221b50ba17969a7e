// site_chain.cuh -- the compact (per-site) part of the dilated dense plan in ONE kernel: F1..F6, G1..G6, T7_0..3, T8_0..1 and the
// FC head of a tile of sites, with every intermediate map kept ON CHIP IN TENSOR MEMORY.
//
// Launched op by op (round 1) the compact chain moved ~23 KB of HBM traffic per site -- every op wrote its 256..512 B/site map and
// the next one read it back, next to the scatter copies of the dense layers -- for 415 k MAC per site: each op was HBM-bound and the
// chain took 24 ms of a 96 ms step at half of its tensor-core bound.  Here a CTA pair (cta_group::2, M = 256 = two tiles of 128
// sites) walks the whole chain for its tiles:
//
//   resident   the running F and G maps (later T7 / T8) of the CTA's 128 sites live in TMEM as packed bf16 hi / lo pairs (site = TMEM
//              lane, two channels per 32-bit column) and are the A operand of the next op's MMAs straight from there (the ".ts" form
//              of tcgen05.mma).  512 columns = F map 128 | G map 128 | F accumulator 128 | G accumulator 128; the tail re-uses them
//              (T7: four 64-column accumulators, packed outputs over the dead F6 / G6; head: 256-column accumulator over the dead T7s).
//              The first version of this kernel kept the maps in shared memory (2 x 64 KiB), which left a ring of 8 x 12 KiB and one
//              (op, 16-channel stage, term) per slot: 3 MMAs (~190 cycles) per ring hand-over against ~400 cycles of issue, barrier and
//              commit latency in BOTH the producer warps and the MMA warp (clock64 stamps, HM_CHAIN_STAMPS) -- 30 ms per step against
//              24 ms op by op.  With the maps in TMEM all of shared memory is ring.
//   streamed   through a ring of 9 slots of 24 KiB, one slot per (op, PAIR of 16-channel stages, term): the term's weight tiles (hi, lo;
//              this CTA's half of N; from L2 -- the whole model is 1 MB) and, for terms that read a scatter copy of a dense map,
//              the 8 activation planes of the two stages (16 KiB, HBM -> shared, 1-D bulk copies); 6 MMAs per hand-over.  The
//              conv1-form ops F1 / G1 open the chain: their steps hold 4 taps of the sites' windows, gathered from the X map with
//              one 16-byte cp.async per (site, tap, hi / lo) through the site-row index
//   MMA warp   (leader CTA) per slot and stage the split-precision triple hi*hi + lo*hi + hi*lo, M = 256; waits for the epilogues an
//              op depends on (mbarrier) before its first MMA, so F and G ops alternate: MMA(G_l) runs under epilogue(F_l).  Resident
//              operands also run faster: N = 96 / 64 take 48 / 46 cycles per MMA from TMEM against 64 from shared memory
//              (tools/mma_ts_probe.cu)
//   epilogue   8 warps per CTA: accumulator -> + bias, ReLU -> hi / lo bf16 pairs -> tcgen05.st into the op's packed columns; the
//              head's epilogue applies fc2, the softmax and the `(int)(255 p)` truncation and writes the site's logits and ML byte
//
// Who may overwrite what (the column plan is made by the host, cnn_tensor.cu): MMAs execute in issue order, so an accumulator may
// overwrite columns that EARLIER MMAs read; an MMA waits (mma_wait) for the epilogues that produce its resident inputs and that drain
// the previous contents of its accumulator columns; an epilogue waits for its own op's MMAs, or (wait_op) for the last LATER op whose
// MMAs still read the columns its packed output goes to; the epilogue warps of a CTA meet at a named barrier after the head, because
// the next round's F2 / G2 outputs go where the head's accumulator was.
//
// HBM traffic per site: the scatter copies (6.5 KB) and the site's X rows read once, 9 B of logits + ML byte written.  No other
// launch belongs to the compact part of the plan.
#pragma once
#include "dense_gemm2.cuh"
#include "postprocess.cuh"

namespace hm {

constexpr int kChainMaxOps = 20;
constexpr int kChainSlots = 9;
constexpr uint32_t kChainPlaneBytes = 2048;    // 128 rows x 16 B
constexpr uint32_t kChainWBytes = 8192;        // weight tiles of a step with a slab (two stages x {hi, lo} x N/2 <= 64 rows x 32 B)
constexpr uint32_t kChainSlabBytes = 16384;    // 8 planes: {hi, lo} x 4 channel groups = two 16-channel stages of a streamed operand
constexpr uint32_t kChainSlotBytes = 24576;    // weights first, slab behind them; a step without slab may use all of it (head: 16 KiB)
constexpr int kChainMaxSteps = 192;            // table of the steps of one tile round (124 for the shipped models); last entry = count
constexpr int kChainMaxWait = 4;
// clock64 stamps (HM_CHAIN_STAMPS=1 at run time) are compiled in only with -DHM_CHAIN_STAMPS_BUILD=1 (tools/build_variant.py): even
// predicated off they cost the single MMA-issuing thread instructions on its critical loop
#ifndef HM_CHAIN_STAMPS_BUILD
#define HM_CHAIN_STAMPS_BUILD 0
#endif

struct ChainTerm {
    const uint8_t* src;   // streamed term: plane 0 (hi, g = 0), row 0 of the compact source map; resident term: nullptr
    uint32_t a_hi_col;    // resident term: first TMEM column of the packed map's hi half and of its lo half (cin / 2 columns each)
    uint32_t a_lo_col;
};

struct ChainOp {
    ChainTerm term[kMaxTerms];
    const uint8_t* w_img;   // chain image of rank 0: [stage pair][term][stage][hl] tiles of (n / 2) x 32 bytes; rank 1: w_rank_bytes further
    const float* bias;      // [n]
    uint8_t* spill;         // debug (hm_debug_dump_acts): the output map is ALSO stored here (compact map in HBM); else nullptr
    uint32_t w_rank_bytes;
    uint32_t acc_col;       // first accumulator column
    uint32_t out_hi_col;    // packed output: n / 2 columns of hi pairs, n / 2 columns of lo pairs
    uint32_t out_lo_col;
    int32_t n, cin, n_terms;
    int32_t head;           // 1: fc1 + ReLU -> fc2 -> logits instead of a packed map
    int32_t wait_op;        // the epilogue waits for the MMAs of THIS op (>= own index)
    int32_t gather;         // 1: conv1 form (F1, G1): term 0 is gathered from the X map -- a ring step holds 4 TAPS (= 2 K-steps of two
                            // 8-feature taps) of the 128 sites' windows, rows site_rows[site] + gather_shift + tap; cin = 8 * padded taps
    int32_t gather_shift;
    uint32_t mma_wait;      // the first MMA waits for the epilogues of up to four ops: one byte each, 1 + op index, 0 = unused
    uint32_t pad_[3];
};

struct ChainProgram {
    ChainOp op[kChainMaxOps];
    int32_t n_ops;
    uint32_t n_tiles;                  // 128-site tiles
    unsigned long long plane_stride;   // of every compact map (streamed sources and spill targets)
    const float* w2;                   // head: [2][256], [2]
    const float* b2;
    const uint32_t* site_rows;         // [rows] compact row -> row of the batch-wide X map (gathered conv1-form ops)
    unsigned long long x_lo_off;       // byte distance from the X map's hi plane to its lo plane
    const uint32_t* out_idx;           // [rows] compact row -> index of the site in hm_call_batch order; 0xffffffff = padding row
    float* logits;                     // [sites][2] in hm_call_batch order
    uint8_t* ml;                       // [sites] the quantised ML byte (s_logits_to_methy_probs, src/app/hifimeth/mod_batch.cpp:46-64)
    long long* dbg;                    // HM_CHAIN_STAMPS: clock64 timeline of pair 0 (third tile round), else nullptr
};

inline size_t chain_smem_bytes()
{
    return (size_t)kChainSlots * kChainSlotBytes + (2 * kChainSlots + 2 * kChainMaxOps) * sizeof(uint64_t) + kChainMaxSteps * sizeof(uint32_t) + 32 +
           kChainMaxOps * sizeof(ChainOp) + 3 * 256 * sizeof(float) + kTileRows * sizeof(float2);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDenseThreads, 1) site_chain_kernel(const __grid_constant__ ChainProgram prog)
{
    extern __shared__ __align__(128) uint8_t smem[];
    // the warp index through a shuffle: ptxas then knows it is warp-uniform, role branches become uniform branches and the MMA
    // issuer's loop counters, descriptors and barrier addresses can stay in uniform registers (cutlass::canonical_warp_idx_sync)
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t rank = umma::cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    uint8_t* s_ring = smem;                                  // [9][24 KiB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)kChainSlots * kChainSlotBytes);
    uint64_t* full = bars;                                   // [9]   leader: own expect_tx arrive + the peer's relay; peer: own arrive
    uint64_t* empty = full + kChainSlots;                    // [9]   multicast commit from the leader
    uint64_t* acc_full = empty + kChainSlots;                // [ops] multicast commit: every MMA of the op (and before it) has completed
    uint64_t* res_ready = acc_full + kChainMaxOps;           // [ops] leader only: one arrive per epilogue warp of both CTAs
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(res_ready + kChainMaxOps);  // steps of a tile round: op | stage pair << 8 | term << 16
    uint32_t* s_tmem = s_tab + kChainMaxSteps;
    // The op table is read from shared memory: indexed dynamically in the kernel-parameter bank every new op cost several constant-cache
    // misses in a row (~600 cycles at every op boundary of the MMA warp, clock64 stamps).
    const ChainOp* s_ops = reinterpret_cast<const ChainOp*>(s_tmem + 8);
    float* s_head = reinterpret_cast<float*>(s_tmem + 8) + kChainMaxOps * sizeof(ChainOp) / 4;  // [3][256] fc1 bias, fc2 row 0, fc2 row 1
    float2* s_hand = reinterpret_cast<float2*>(s_head + 3 * 256);                                // [128] partial logits of the second column half
    const int n_ops = prog.n_ops;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&prog.op[0]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_tmem + 8);
        for (uint32_t i = threadIdx.x; i < (uint32_t)(n_ops * sizeof(ChainOp) / 4); i += blockDim.x) dst[i] = src[i];
    }

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < kChainSlots; ++i) {
                umma::mbar_init(&full[i], (rank == 0 ? 2u : 1u) + 32u);  // + one cp.async-tracking arrival per producer lane
                umma::mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < kChainMaxOps; ++i) {
                umma::mbar_init(&acc_full[i], 1);
                umma::mbar_init(&res_ready[i], 2 * kEpilogueWarps);
            }
            umma::fence_barrier_init();
        }
        __syncwarp();
        umma::tmem_alloc2(s_tmem, 512);
    } else if (warp == 1 && lane == 0) {
        // step table of one tile round: op | stage pair << 8 | term << 16 | last step of its op << 20 | (1 + op) << 24 when the step
        // that used this step's ring slot before it (9 steps earlier) was the last of that op
        uint32_t j = 0;
        for (int oi = 0; oi < n_ops; ++oi) {
            const uint32_t ns = (uint32_t)prog.op[oi].cin >> 5, nt = (uint32_t)prog.op[oi].n_terms;
            for (uint32_t S = 0; S < ns; ++S)
                for (uint32_t k = 0; k < nt; ++k)
                    if (j < (uint32_t)kChainMaxSteps - 1u) s_tab[j++] = (uint32_t)oi | (S << 8) | (k << 16) | ((S + 1 == ns && k + 1 == nt) ? 1u << 20 : 0u);
        }
        for (uint32_t i = 0; i < j; ++i) {
            const uint32_t prev = s_tab[(i + j - (uint32_t)kChainSlots) % j];
            if (prev & (1u << 20)) s_tab[i] |= ((prev & 0xffu) + 1u) << 24;
        }
        s_tab[kChainMaxSteps - 1] = j;  // steps per tile round (the host checks that the table holds them and that there are >= 9)
    } else if (warp == 2) {
        // head: fc1 bias and the two fc2 rows, read by every epilogue thread for every site
        for (int oi = 0; oi < n_ops; ++oi)
            if (prog.op[oi].head)
                for (uint32_t i = lane; i < 256u; i += 32u) {
                    s_head[i] = prog.op[oi].bias[i];
                    s_head[256 + i] = prog.w2[i];
                    s_head[512 + i] = prog.w2[256 + i];
                }
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync();  // both CTAs' barriers exist before anyone arrives remotely
    umma::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t spi = s_tab[kChainMaxSteps - 1];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp < (uint32_t)kProducerWarps) {
        // ===================================== producers: one ring slot per (op, stage pair, term) ===========================
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the streamed maps come from earlier launches
        uint32_t j = warp, t2 = pair, it = 0;
        for (uint32_t step = warp;; step += (uint32_t)kProducerWarps, j += (uint32_t)kProducerWarps) {
            while (j >= spi) { j -= spi; t2 += n_pairs; ++it; }
            if (2 * t2 >= prog.n_tiles) break;
            const uint32_t e = s_tab[j];
            const uint32_t oi = e & 0xffu, S = (e >> 8) & 0xffu, k = (e >> 16) & 0xfu, rel = e >> 24;
            const ChainOp& op = s_ops[oi];
            const uint32_t groups = (uint32_t)op.cin >> 3;
            const uint32_t w_bytes = (uint32_t)op.n * 64u;  // two stages x {hi, lo} tiles of this CTA's half of N
            const uint8_t* src = op.term[k].src;
            const uint8_t* w_src = op.w_img + (size_t)rank * op.w_rank_bytes + (size_t)(S * (uint32_t)op.n_terms + k) * w_bytes;
            const unsigned long long row0 = (unsigned long long)(2 * t2 + rank) * kTileRows;  // odd n_tiles: the peer's last tile lies in the slack rows
            const uint32_t slot = step % (uint32_t)kChainSlots, phase = (step / (uint32_t)kChainSlots) & 1u;
            if (rel) {
                // the slot's previous step was the last of op rel - 1: its MMAs were committed to that op's acc_full barrier only (a
                // second tcgen05.commit right behind the slot's own stalled the MMA warp ~600 cycles at every op boundary); the
                // slot's own barrier gets the missing arrival from here, so that its phases keep counting uses
                if (step >= (uint32_t)kChainSlots) {
                    umma::mbar_wait(&acc_full[rel - 1u], (j >= (uint32_t)kChainSlots ? it : it - 1u) & 1u);
                    if (lane == 0) umma::mbar_arrive(&empty[slot]);
                }
            } else {
                umma::mbar_wait(&empty[slot], phase ^ 1u);
            }
            uint8_t* dst = s_ring + (size_t)slot * kChainSlotBytes;
            if (lane == 0) umma::mbar_arrive_expect_tx(&full[slot], ((src && !op.gather) ? kChainSlabBytes : 0u) + w_bytes);
            __syncwarp();
            if (op.gather) {
                // conv1 form: planes {hi, lo} x {tap 4 S .. 4 S + 3}, every row fetched with its own 16-byte copy (the four taps of a
                // site are 64 consecutive bytes of the X map)
                #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t r = lane + 32u * (uint32_t)q;
                    const unsigned long long xr = (unsigned long long)__ldg(prog.site_rows + row0 + r) + (unsigned long long)(op.gather_shift + (int)(4u * S));
                    #pragma unroll
                    for (uint32_t pl = 0; pl < 8u; ++pl)
                        umma::cp_async16(dst + kChainWBytes + pl * kChainPlaneBytes + r * 16u,
                                         src + (pl >> 2) * prog.x_lo_off + (xr + (pl & 3u)) * 16ull);
                }
            } else if (lane < 8u) {
                if (src) {  // planes {hi, lo} x {g = 4 S .. 4 S + 3}
                    const uint8_t* plane = src + (unsigned long long)((lane >> 2) * groups + 4u * S + (lane & 3u)) * prog.plane_stride;
                    umma::bulk_g2s(dst + kChainWBytes + lane * kChainPlaneBytes, plane + row0 * 16ull, kChainPlaneBytes, &full[slot]);
                }
            }
            umma::cp_async_mbar_arrive_noinc(&full[slot]);  // every lane, every step: the barrier counts 32 of these
            if (lane >= 8u && lane < 12u) {  // the weight tiles in four pieces
                const uint32_t per = w_bytes >> 2, piece = lane - 8u;
                umma::bulk_g2s(dst + piece * per, w_src + piece * per, per, &full[slot]);
            }
            __syncwarp();
            if (rank != 0) {
                // the peer tells the leader when its half of the slot has landed
                umma::mbar_wait(&full[slot], phase);
                if (lane == 0) umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&full[slot]), 0));
                __syncwarp();
            }
        }
    } else if (warp == (uint32_t)kProducerWarps) {
        if (rank == 0) {
            // ===================================== MMA issuer (leader) =========================================================
            const uint32_t desc_hi = (uint32_t)(umma::make_desc(0, 0, 128) >> 32);
            const uint32_t ring16 = umma::smem_u32(s_ring) >> 4;
            const uint32_t a_desc = (uint32_t)umma::make_desc(0, kChainPlaneBytes, 128);  // K-adjacent core matrices one plane apart
            // Everything the issue loop needs from an op sits in registers, and the NEXT op's values are fetched while the current
            // op's steps issue: read from the op table at the op boundary they cost the one issuing thread a chain of shared-memory
            // round trips there, on top of a load per step for the term's kind.
            struct OpVals {
                uint32_t n_spairs, n_terms, idesc, b_desc, b_tile16, acc_col, mma_wait, stream_mask;
                uint32_t a_hi[kMaxTerms], a_lo[kMaxTerms];
            };
            auto load_op = [&](int oi) {
                const ChainOp& op = s_ops[oi];
                OpVals v;
                const uint32_t nh = (uint32_t)op.n >> 1;
                v.n_spairs = (uint32_t)op.cin >> 5;
                v.n_terms = (uint32_t)op.n_terms;
                v.idesc = umma::make_idesc_bf16_m256((uint32_t)op.n);
                v.b_desc = (uint32_t)umma::make_desc(0, nh * 16u, 128);
                v.b_tile16 = (nh * 32u) >> 4;  // the {hi} or the {lo} tile of one stage
                v.acc_col = op.acc_col;
                v.mma_wait = op.mma_wait;
                v.stream_mask = 0;
                #pragma unroll
                for (int k = 0; k < kMaxTerms; ++k) {
                    if (op.term[k].src != nullptr) v.stream_mask |= 1u << k;
                    v.a_hi[k] = tmem_base + op.term[k].a_hi_col;
                    v.a_lo[k] = tmem_base + op.term[k].a_lo_col;
                }
                return v;
            };
            uint32_t step = 0, it = 0, mstep = 0;
            bool ready = false;  // the NEXT step's slot is probed before this step's MMAs are issued: hides the barrier round trip
            OpVals cur = load_op(0);
            for (uint32_t t2 = pair; 2 * t2 < prog.n_tiles; t2 += n_pairs, ++it) {
                for (int oi = 0; oi < n_ops; ++oi) {
                    OpVals nxt = cur;
                    const bool sto = HM_CHAIN_STAMPS_BUILD && prog.dbg && pair == 0 && it == 2 && lane == 0;
                    if (sto) prog.dbg[320 + 2 * oi] = clock64();
                    // resident inputs written and accumulator columns drained by earlier epilogues (both CTAs)
                    for (uint32_t w = cur.mma_wait; w; w >>= 8) umma::mbar_wait(&res_ready[(w & 0xffu) - 1u], it & 1u);
                    umma::tc_fence_after();
                    if (sto) prog.dbg[321 + 2 * oi] = clock64();
                    const uint32_t d_addr = tmem_base + cur.acc_col;
                    const uint32_t n_steps = cur.n_spairs * cur.n_terms;
                    uint32_t acc = 0, S = 0, k = 0;
                    for (uint32_t q = 0; q < n_steps; ++q, ++step) {
                        const uint32_t slot = step % (uint32_t)kChainSlots, phase = (step / (uint32_t)kChainSlots) & 1u;
                        const bool st = HM_CHAIN_STAMPS_BUILD && prog.dbg && pair == 0 && it == 2 && lane == 0 && mstep < 124u;
                        if (st) prog.dbg[2 * mstep] = clock64();
                        if (!ready) umma::mbar_wait(&full[slot], phase);
                        {
                            const uint32_t ns = step + 1u;
                            ready = umma::mbar_test_wait(&full[ns % (uint32_t)kChainSlots], (ns / (uint32_t)kChainSlots) & 1u);
                        }
                        if (st) {
                            prog.dbg[2 * mstep + 1] = clock64();
                            ++mstep;
                        }
                        umma::tc_fence_after();
                        const bool stream = (cur.stream_mask >> k) & 1u;
                        const uint32_t ah = k == 0 ? cur.a_hi[0] : k == 1 ? cur.a_hi[1] : cur.a_hi[2];
                        const uint32_t al = k == 0 ? cur.a_lo[0] : k == 1 ? cur.a_lo[1] : cur.a_lo[2];
                        if (umma::elect_one()) {
                            const uint32_t sb16 = ring16 + slot * (kChainSlotBytes >> 4);
                            const uint32_t b0 = cur.b_desc + sb16;
                            // six MMAs in one asm statement: the operands cross into uniform registers once (umma::mma2_stage3_bf16)
                            if (stream) {
                                const uint32_t a0 = a_desc + sb16 + (kChainWBytes >> 4);
                                umma::mma2_step6_bf16(d_addr, a0, a0 + ((4u * kChainPlaneBytes) >> 4), (2u * kChainPlaneBytes) >> 4, b0, cur.b_tile16, desc_hi,
                                                      cur.idesc, acc);
                            } else {
                                umma::mma2_step6_ts_bf16(d_addr, ah + 16u * S, al + 16u * S, b0, cur.b_tile16, desc_hi, cur.idesc, acc);
                            }
                            // the op's last step releases its slot through acc_full (see the producers)
                            umma::mma2_commit_mc(q + 1 == n_steps ? &acc_full[oi] : &empty[slot]);
                        }
                        acc = 1;
                        __syncwarp();
                        if (q == 0) nxt = load_op(oi + 1 == n_ops ? 0 : oi + 1);  // in flight under this op's remaining steps
                        if (++k == cur.n_terms) { k = 0; ++S; }
                    }
                    cur = nxt;
                }
            }
        }
    } else {
        // ===================================== epilogue (own CTA's 128 sites) ====================================================
        const uint32_t lane_grp = (warp & 3u) * 32u;
        const uint32_t half = (warp - (uint32_t)(kProducerWarps + 1)) >> 2;
        const uint32_t m = lane_grp + lane;
        const uint32_t t_lane = tmem_base + (lane_grp << 16);
        uint32_t it = 0;
        for (uint32_t t2 = pair; 2 * t2 < prog.n_tiles; t2 += n_pairs, ++it) {
            const unsigned long long row = (unsigned long long)(2 * t2 + rank) * kTileRows + m;
            for (int oi = 0; oi < n_ops; ++oi) {
                const ChainOp& op = s_ops[oi];
                const int n = op.n;
                const bool st = HM_CHAIN_STAMPS_BUILD && prog.dbg && pair == 0 && it == 2 && rank == 0 && warp == (uint32_t)kProducerWarps + 1u && lane == 0;
                if (st) prog.dbg[256 + 3 * oi] = clock64();
                umma::mbar_wait(&acc_full[op.wait_op], it & 1u);
                if (st) prog.dbg[256 + 3 * oi + 1] = clock64();
                umma::tc_fence_after();
                const uint32_t t_addr = t_lane + op.acc_col;
                if (op.head) {
                    // fc1 + ReLU from the accumulator, fc2 as two dot products per site (bias and fc2 rows staged in shared memory).  The
                    // two warps of a lane group take half of the columns each; the second one hands its partial sums over through
                    // shared memory across a named barrier, which also tells every epilogue warp of this CTA that the head accumulator
                    // has been drained: the next round's F2 / G2 outputs go to those columns.
                    float l0 = 0.f, l1 = 0.f;
                    const int cb = half ? n >> 1 : 0, ce = half ? n : n >> 1;
                    for (int c0 = cb; c0 < ce; c0 += 32) {
                        uint32_t v[32];
                        umma::tmem_ld32(t_addr + (uint32_t)c0, v);
                        umma::tmem_ld_wait();
                        #pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = *reinterpret_cast<const float4*>(s_head + c0 + j);
                            const float4 w0 = *reinterpret_cast<const float4*>(s_head + 256 + c0 + j);
                            const float4 w1 = *reinterpret_cast<const float4*>(s_head + 512 + c0 + j);
                            const float f0 = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f), f1 = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                            const float f2 = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f), f3 = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
                            l0 = fmaf(f0, w0.x, l0); l0 = fmaf(f1, w0.y, l0); l0 = fmaf(f2, w0.z, l0); l0 = fmaf(f3, w0.w, l0);
                            l1 = fmaf(f0, w1.x, l1); l1 = fmaf(f1, w1.y, l1); l1 = fmaf(f2, w1.z, l1); l1 = fmaf(f3, w1.w, l1);
                        }
                    }
                    if (half) s_hand[m] = make_float2(l0, l1);
                    umma::tc_fence_before();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (!half) {
                        // logits -> max-subtracted softmax -> ML byte, written where the caller reads it: no per-site pass of its own
                        const uint32_t site = __ldg(prog.out_idx + row);
                        if (site != 0xffffffffu) {
                            const float2 o = s_hand[m];
                            const float v0 = (l0 + o.x) + __ldg(prog.b2), v1 = (l1 + o.y) + __ldg(prog.b2 + 1);
                            *reinterpret_cast<float2*>(prog.logits + 2ull * site) = make_float2(v0, v1);
                            prog.ml[site] = prob_to_ml(softmax_p1(v0, v1));
                        }
                    }
                } else {
                    const int mid = ((n >> 1) + 15) & ~15;
                    int c0 = half ? mid : 0;
                    const int c1 = half ? n : mid;
                    const uint32_t out_groups = (uint32_t)n >> 3;
                    while (c0 < c1) {
                        const int nc = (c1 - c0 >= 32) ? 32 : 16;
                        uint32_t v[32];
                        if (nc == 32) umma::tmem_ld32(t_addr + (uint32_t)c0, v);
                        else umma::tmem_ld16(t_addr + (uint32_t)c0, reinterpret_cast<uint32_t(&)[16]>(v));
                        umma::tmem_ld_wait();
                        uint32_t hi[16], lo[16];
                        #pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            if (2 * j < nc) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(op.bias + c0 + 2 * j));
                                const float x0 = fmaxf(__uint_as_float(v[2 * j]) + bv.x, 0.f), x1 = fmaxf(__uint_as_float(v[2 * j + 1]) + bv.y, 0.f);
                                const float x2 = fmaxf(__uint_as_float(v[2 * j + 2]) + bv.z, 0.f), x3 = fmaxf(__uint_as_float(v[2 * j + 3]) + bv.w, 0.f);
                                const __nv_bfloat162 h0 = __floats2bfloat162_rn(x0, x1), h1 = __floats2bfloat162_rn(x2, x3);
                                const uint32_t hb0 = *reinterpret_cast<const uint32_t*>(&h0), hb1 = *reinterpret_cast<const uint32_t*>(&h1);
                                const __nv_bfloat162 e0 = __floats2bfloat162_rn(x0 - __uint_as_float(hb0 << 16), x1 - __uint_as_float(hb0 & 0xffff0000u));
                                const __nv_bfloat162 e1 = __floats2bfloat162_rn(x2 - __uint_as_float(hb1 << 16), x3 - __uint_as_float(hb1 & 0xffff0000u));
                                hi[j] = hb0;
                                hi[j + 1] = hb1;
                                lo[j] = *reinterpret_cast<const uint32_t*>(&e0);
                                lo[j + 1] = *reinterpret_cast<const uint32_t*>(&e1);
                            }
                        }
                        const uint32_t pc = (uint32_t)c0 >> 1;
                        if (nc == 32) {
                            umma::tmem_st16(t_lane + op.out_hi_col + pc, hi);
                            umma::tmem_st16(t_lane + op.out_lo_col + pc, lo);
                        } else {
                            umma::tmem_st8(t_lane + op.out_hi_col + pc, hi);
                            umma::tmem_st8(t_lane + op.out_lo_col + pc, lo);
                        }
                        if (op.spill) {
                            const uint32_t g0 = (uint32_t)c0 >> 3;
                            #pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                if (8 * g < nc) {
                                    uint8_t* q = op.spill + (unsigned long long)(g0 + g) * prog.plane_stride + row * 16ull;
                                    *reinterpret_cast<uint4*>(q) = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
                                    *reinterpret_cast<uint4*>(q + (unsigned long long)out_groups * prog.plane_stride) =
                                        make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
                                }
                            }
                        }
                        c0 += nc;
                    }
                    umma::tmem_st_wait();  // the tensor core reads what these threads just wrote
                    umma::tc_fence_before();
                }
                __syncwarp();
                if (st) prog.dbg[256 + 3 * oi + 2] = clock64();
                if (lane == 0) {
                    if (rank == 0) umma::mbar_arrive(&res_ready[oi]);
                    else umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&res_ready[oi]), 0));
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
