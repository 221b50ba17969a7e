// dense_gemm2.cuh -- the CTA-pair (cta_group::2) form of dense_gemm_kernel for the layers whose weights are too large
// to leave room for a deep activation ring in one SM (conv2, conv3: 3 x 128 x 128 x {hi,lo} = 192 KiB).
//
// A cluster of two CTAs computes 256 rows per tcgen05.mma: each CTA stages its own 128 rows of A and HALF of every weight
// tile (its N/2 output channels), so resident weights drop to 96 KiB per SM and the ring deepens from 3 to 8 stages.
//   * rank 0 (leader): warp 8 issues every MMA (M = 256) and commits with a multicast arrive to BOTH CTAs' barriers
//   * rank 1 (peer):   every producer warp relays its own ring slot's `full` barrier to the leader once its copies have
//                      landed (cluster-scope arrive); warp 8 relays the weight barrier
//   * producers and epilogue warps work on their own CTA's tile exactly as in the single-CTA kernel
// Dense (bulk-copy), normal-form, map-output ops only; same DenseOp parameters, w_img = [rank 0 half][rank 1 half],
// w_bytes = bytes of ONE half.
#pragma once
#include "dense_gemm.cuh"

namespace hm {

// clock64 stamps of pair 0 (HM_G2_STAMPS=1 at run time) are compiled in only with -DHM_G2_STAMPS_BUILD=1 (tools/build_variant.py):
// per tile 16 .. 47 of the launch, dbg[16 * (it - 16) + k]: MMA warp k = 0 before / 1 after the wait for the accumulator buffer,
// 2 cycles spent waiting for ring stages, 3 last stage issued; epilogue warp 0 of the leader: 4 before / 5 after the wait for the
// accumulator, 6 tile stored; producer warp 0 of the leader: 8 cycles spent waiting for its ring slot during the tile.
#ifndef HM_G2_STAMPS_BUILD
#define HM_G2_STAMPS_BUILD 0
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDenseThreads, 1) dense_gemm2_kernel(const __grid_constant__ DenseOp op)
{
    extern __shared__ __align__(128) uint8_t smem[];
    // the warp index through a shuffle: ptxas then knows it is warp-uniform, role branches become uniform branches and the MMA
    // issuer's loop counters, descriptors and barrier addresses can stay in uniform registers (cutlass::canonical_warp_idx_sync)
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t rank = umma::cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    uint8_t* s_w = smem;
    uint8_t* s_ring = smem + ((op.w_bytes + 127u) & ~127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)op.ring * op.stage_bytes);
    uint64_t* full = bars;                  // [ring]  leader: own expect_tx arrive + the peer's relay; peer: own arrive
    uint64_t* empty = bars + op.ring;       // [ring]  multicast commit from the leader
    uint64_t* w_full = bars + 2 * op.ring;  // [1]
    uint64_t* t_full = w_full + 1;          // [2]     multicast commit from the leader
    uint64_t* t_empty = t_full + 2;         // [2]     leader only: one arrive per epilogue warp of both CTAs
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_empty + 2);
    // as an offset from the shared-memory array, so that the compiler keeps the accesses in the shared state space (through a
    // uintptr_t round trip they became generic LD / ST: long-scoreboard latency and a queue shared with the global stores)
    float* s_bias = reinterpret_cast<float*>(smem + (((uint32_t)(reinterpret_cast<uint8_t*>(t_empty + 3) - smem) + 15u) & ~15u));

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < op.ring; ++i) {
                umma::mbar_init(&full[i], rank == 0 ? 2u : 1u);
                umma::mbar_init(&empty[i], 1);
            }
            umma::mbar_init(w_full, rank == 0 ? 2u : 1u);
            for (int i = 0; i < 2; ++i) {
                umma::mbar_init(&t_full[i], 1);
                umma::mbar_init(&t_empty[i], 2 * kEpilogueWarps);
            }
            umma::fence_barrier_init();
        }
        __syncwarp();
        umma::tmem_alloc2(s_tmem, op.tmem_cols);
    }
    umma::tc_fence_before();
    umma::cluster_sync();  // both CTAs' barriers exist before anyone arrives remotely
    umma::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t acc_stride = op.tmem_cols >> 1;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp < (uint32_t)kProducerWarps) {
        // ===================================== producers (own CTA's rows, own half of the weights) ==================
        if (warp == 0) {
            if (lane == 0) umma::mbar_arrive_expect_tx(w_full, op.w_bytes);
            __syncwarp();
            const uint32_t per = (((op.w_bytes + 31u) / 32u) + 15u) & ~15u;
            const uint32_t off = lane * per;
            if (off < op.w_bytes) umma::bulk_g2s(s_w + off, op.w_img + (size_t)rank * op.w_bytes + off, min(per, op.w_bytes - off), w_full);
        }
        const uint32_t n_prod = (uint32_t)op.ring;
        asm volatile("griddepcontrol.wait;" ::: "memory");
        uint32_t stage_no = 0, stage_tx = 0, ncopies = 0;
        for (int s = 0; s < op.n_segs; ++s) {
            stage_tx += op.seg[s].nrows * 16u * 4u;
            ncopies += 4u;
        }
        for (uint32_t t2 = pair; 2 * t2 < op.n_tiles; t2 += n_pairs) {
            const long long row0 = (long long)(2 * t2 + rank) * kTileRows;  // odd n_tiles: the peer's last tile lies in the slack rows
            for (int st = 0; st < op.n_stages; ++st, ++stage_no) {
                if (stage_no % n_prod != warp) continue;
                const uint32_t slot = stage_no % (uint32_t)op.ring, phase = (stage_no / (uint32_t)op.ring) & 1u;
                const uint32_t pit = (t2 - pair) / n_pairs;
                const bool pst = HM_G2_STAMPS_BUILD && op.dbg && pair == 0 && rank == 0 && warp == 0 && lane == 0 && pit >= 16u && pit < 48u;
                const long long pt0 = pst ? clock64() : 0;
                umma::mbar_wait(&empty[slot], phase ^ 1u);
                if (pst) op.dbg[16 * (pit - 16u) + 8] += clock64() - pt0;
                uint8_t* stage = s_ring + (size_t)slot * op.stage_bytes;
                if (lane == 0) umma::mbar_arrive_expect_tx(&full[slot], stage_tx);
                __syncwarp();
                if (lane < ncopies) {
                    const uint32_t p = lane & 3u;
                    const DenseSeg& sg = op.seg[lane >> 2];
                    const uint32_t pl_bytes = sg.nrows * 16u;
                    const uint8_t* plane = sg.src + (unsigned long long)((p >> 1) * sg.groups + (uint32_t)(2 * st) + (p & 1u)) * sg.plane_stride;
                    umma::bulk_g2s(stage + sg.smem_off + p * pl_bytes, plane + (row0 + sg.row_off) * 16ll, pl_bytes, &full[slot]);
                }
                __syncwarp();
                if (rank != 0) {
                    // the peer tells the leader when its half of the stage has landed (one relay per slot owner, in parallel)
                    umma::mbar_wait(&full[slot], phase);
                    if (lane == 0) umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&full[slot]), 0));
                    __syncwarp();
                }
            }
        }
    } else if (warp == (uint32_t)kProducerWarps) {
        const int n_terms = op.n_terms, n_stages = op.n_stages, ring = op.ring;
        if (rank == 0) {
            // ===================================== MMA issuer (leader) ============================================
            const uint32_t idesc = umma::make_idesc_bf16_m256((uint32_t)op.n);
            const uint32_t desc_hi = (uint32_t)(umma::make_desc(0, 0, 128) >> 32);
            const uint32_t nh = (uint32_t)op.n >> 1;                       // B rows held by each CTA
            const uint32_t b_step = (nh * 32u) >> 4;
            const uint32_t b_base = (uint32_t)umma::make_desc(umma::smem_u32(s_w), nh * 16u, 128);
            const uint32_t ring16 = umma::smem_u32(s_ring) >> 4, stage16 = op.stage_bytes >> 4;
            uint32_t a_hi[kMaxTerms], a_lo[kMaxTerms];
            #pragma unroll
            for (int k = 0; k < kMaxTerms; ++k) {
                a_hi[k] = (uint32_t)umma::make_desc(op.term[k].a_off, op.term[k].a_lbo, 128);
                a_lo[k] = (uint32_t)umma::make_desc(op.term[k].a_off + op.term[k].a_hl_off, op.term[k].a_lbo, 128);
            }
            umma::mbar_wait(w_full, 0);  // both halves of the weights have landed (own copy + the peer's relay)
            uint32_t slot = 0, phase = 0, it = 0;
            for (uint32_t t2 = pair; 2 * t2 < op.n_tiles; t2 += n_pairs, ++it) {
                const uint32_t buf = it & 1u, use = it >> 1;
                const bool mst = HM_G2_STAMPS_BUILD && op.dbg && pair == 0 && lane == 0 && it >= 16u && it < 48u;
                long long* md = op.dbg + 16 * ((int)it - 16);
                if (mst) md[0] = clock64();
                umma::mbar_wait(&t_empty[buf], (use & 1u) ^ 1u);
                if (mst) md[1] = clock64();
                umma::tc_fence_after();
                const uint32_t d_addr = tmem_base + buf * acc_stride;
                uint32_t acc = 0, b_cur = b_base;
                for (int st = 0; st < n_stages; ++st) {
                    const long long mt0 = mst ? clock64() : 0;
                    umma::mbar_wait(&full[slot], phase);
                    if (mst) md[2] += clock64() - mt0;
                    umma::tc_fence_after();
                    if (n_terms == 3) {
                        // all nine MMAs of the stage in one asm statement: see umma::mma2_stage3_bf16
                        const uint32_t sa = ring16 + slot * stage16;
                        if (umma::elect_one()) {
                            umma::mma2_stage3_bf16(d_addr, a_hi[0] + sa, a_lo[0] + sa, a_hi[1] + sa, a_lo[1] + sa, a_hi[2] + sa, a_lo[2] + sa, b_cur, b_step,
                                                   desc_hi, idesc, acc);
                            umma::mma2_commit_mc(&empty[slot]);
                        }
                    } else if (umma::elect_one()) {
                        const uint32_t sa = ring16 + slot * stage16;
                        uint32_t bq = b_cur, a0 = acc;
                        #pragma unroll
                        for (int k = 0; k < kMaxTerms; ++k) {
                            if (k < n_terms) {
                                umma::mma2_bf16_w(d_addr, a_hi[k] + sa, bq, desc_hi, idesc, a0);
                                a0 = 1;
                                umma::mma2_bf16_w(d_addr, a_lo[k] + sa, bq, desc_hi, idesc, 1);
                                umma::mma2_bf16_w(d_addr, a_hi[k] + sa, bq + b_step, desc_hi, idesc, 1);
                                bq += 2 * b_step;
                            }
                        }
                        umma::mma2_commit_mc(&empty[slot]);
                    }
                    acc = 1;
                    b_cur += 2 * b_step * (uint32_t)n_terms;
                    __syncwarp();
                    if (++slot == (uint32_t)ring) { slot = 0; phase ^= 1u; }
                }
                if (umma::elect_one()) umma::mma2_commit_mc(&t_full[buf]);
                if (mst) md[3] = clock64();
                __syncwarp();
            }
        } else {
            // ===================================== peer: only the weights need relaying here ==========================
            umma::mbar_wait(w_full, 0);
            if (lane == 0) umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(w_full), 0));
        }
    } else {
        // ===================================== epilogue (own CTA's 128 rows) ==========================================
        const uint32_t lane_grp = (warp & 3u) * 32u;
        const uint32_t half = (warp - (uint32_t)(kProducerWarps + 1)) >> 2;
        const uint32_t m = lane_grp + lane;
        const int n = op.n;
        for (int i = (int)threadIdx.x - 32 * (kProducerWarps + 1); i < n; i += 32 * kEpilogueWarps) s_bias[i] = __ldg(op.bias + i);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
        uint32_t it = 0;
        // The site-row lookups (global loads) are issued one tile AHEAD: an epilogue-bound kernel never waits at t_full, so loads
        // issued at the top of the same tile had their whole latency exposed at first use (ncu: long-scoreboard stalls there).
        int msc_next[kMaxScatter];
        scatter_rows(op, (unsigned long long)(2 * pair + rank) * kTileRows + m, msc_next);
        for (uint32_t t2 = pair; 2 * t2 < op.n_tiles; t2 += n_pairs, ++it) {
            const uint32_t buf = it & 1u, use = it >> 1;
            const unsigned long long row = (unsigned long long)(2 * t2 + rank) * kTileRows + m;
            int msc[kMaxScatter];
            #pragma unroll
            for (int k = 0; k < kMaxScatter; ++k) msc[k] = msc_next[k];
            if (2 * (t2 + n_pairs) < op.n_tiles) scatter_rows(op, (unsigned long long)(2 * (t2 + n_pairs) + rank) * kTileRows + m, msc_next);
            const bool est = HM_G2_STAMPS_BUILD && op.dbg && pair == 0 && rank == 0 && warp == (uint32_t)kProducerWarps + 1u && lane == 0 && it >= 16u && it < 48u;
            if (est) op.dbg[16 * ((int)it - 16) + 4] = clock64();
            umma::mbar_wait(&t_full[buf], use & 1u);
            if (est) op.dbg[16 * ((int)it - 16) + 5] = clock64();
            umma::tc_fence_after();
            const uint32_t t_addr = tmem_base + (lane_grp << 16) + buf * acc_stride;
            epilogue_map_row(op, t_addr, row, half, s_bias, msc, false);
            umma::tc_fence_before();
            __syncwarp();
            if (est) op.dbg[16 * ((int)it - 16) + 6] = clock64();
            if (lane == 0) {
                if (rank == 0) umma::mbar_arrive(&t_empty[buf]);
                else umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&t_empty[buf]), 0));
            }
        }
    }
    umma::tc_fence_before();
    umma::cluster_sync();  // the peer's shared memory and TMEM stay alive until the leader's last MMA and arrive are done
    umma::tc_fence_after();
    if (warp == 0) umma::tmem_dealloc2(tmem_base, op.tmem_cols);
}

}  // namespace hm
