// dense_fused12.cuh -- conv1 and conv2 of the dilated dense plan in ONE kernel: the Y1 map never goes to HBM.
//
// Unfused, conv1 writes Y1 (512 B per row) and conv2 reads it back: 1 KB of the ~4.9 KB of HBM traffic a dense row costs,
// and conv1 on its own runs at ~55 % of its HBM bound (epilogue-store bound, tools/dense_microbench.py conv1).  Here a
// CTA pair (cta_group::2, the dense_gemm2_kernel skeleton) computes per CTA a tile of 124 conv2 rows:
//
//   producers   X rows [124 T, +144) (8 channels, hi/lo)  --bulk copy-->  X ring (4 slots)
//   MMA warp    MMA1: conv1 form (11/13 taps along K, LBO = 16 B) -> accumulator A1[T & 1] in TMEM (Y1 rows 124 T .. +127)
//   epilogue-1  8 warps: A1 -> + bias1, ReLU, hi/lo bf16 -> the A RING itself, in the plane layout conv2's bulk copies
//               would have produced (8 stages x {hi g0, hi g1, lo g0, lo g1} x 132 rows x 16 B); stage s may be written as
//               soon as MMA2 of the previous tile has released it and is handed to the MMA warp with an mbarrier arrive
//               (fence.proxy.async in front: generic-proxy writes read by the tensor core) -- the role the producer warps
//               play in dense_gemm2_kernel.  It also writes the compact scatter copies of Y1 (operands of F2 / G2).
//   MMA warp    MMA2: conv2 (3 taps = descriptor row shifts 0 / 2 / 4) over the 8 stages -> accumulator A2[T & 1]
//   epilogue-2  8 warps: A2 -> + bias2, ReLU, hi/lo -> Y2 rows 124 T .. +123 in HBM (+ scatter copies of Y2)
//
// Issue order MMA1(T+1), MMA2(T), MMA1(T+2), ...: epilogue-1 of tile T+1 runs under MMA2(T), so the tensor pipe does not
// wait for it.  Rows 124 .. 127 of a tile see conv2 taps beyond the 128 Y1 rows the CTA holds; they are computed on
// whatever the ring holds there and never stored -- the next tile recomputes them (3 % extra conv1 + conv2 work).
// TMEM: A1 x 2 + A2 x 2 = 4 x 128 columns = all 512.  Shared memory per CTA: W2 half 96 KiB + W1 half 24 KiB + A ring
// 66 KiB + X ring 18 KiB.
#pragma once
#include "dense_gemm2.cuh"

namespace hm {

constexpr int kF12OutRows = 124;  // conv2 rows per CTA tile (128 - the reach of conv2's taps)
constexpr int kF12XRing = 4;
constexpr int kF12Epi1Warps = 8;  // two per TMEM lane group; 32-channel chunks alternate between the two
constexpr int kF12ProducerWarps = kF12XRing;  // one per X ring slot; fewer threads than the other kernels: 96 registers for the epilogues
constexpr int kF12Threads = 32 * (kF12ProducerWarps + 1 + kEpilogueWarps + kF12Epi1Warps);

struct Fused12Op {
    DenseOp c1;        // conv1 form, pair lowering (w_img = [rank 0 half][rank 1 half]); scatter fields = Y1's scatter copies
    DenseOp c2;        // conv2, pair lowering; out / scatter fields = Y2's; seg[] unused (its operand is produced on chip)
    uint32_t n_tiles;  // 124-row tiles
    long long* dbg;    // HM_F12_STAMPS: CTA 0 writes clock64 stamps of its first 48 tiles, 16 per tile (see kStamp* below)
};
// stamp slots per tile: 0 MMA warp reaches issue of MMA1(t+1), 1 MMA1(t+1) issued, 2 t_empty seen, 3..10 full[0..7] seen,
// 11 MMA2(t) issued; 12 epilogue-1 sees a1_full, 13 / 14 its first / second chunk written; 15 epilogue-2 sees t_full

inline size_t fused12_smem_bytes(const Fused12Op& f)
{
    return ((f.c2.w_bytes + 127u) & ~127u) + ((f.c1.w_bytes + 127u) & ~127u) + (size_t)f.c2.n_stages * f.c2.stage_bytes +
           (size_t)kF12XRing * f.c1.stage_bytes + (2 * 8 + 2 * kF12XRing + 1 + 2 + 2 + 2 + 1) * sizeof(uint64_t) + 16 + 2 * 128 * sizeof(float);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kF12Threads, 1) dense_fused12_kernel(const __grid_constant__ Fused12Op f)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const DenseOp& c1 = f.c1;
    const DenseOp& c2 = f.c2;
    // the warp index through a shuffle: ptxas then knows it is warp-uniform, role branches become uniform branches and the MMA
    // issuer's loop counters, descriptors and barrier addresses can stay in uniform registers (cutlass::canonical_warp_idx_sync)
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const uint32_t rank = umma::cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    constexpr int kStages = 8;  // = c2.n_stages: the A ring holds exactly one Y1 tile, slot = stage
    uint8_t* s_w2 = smem;
    uint8_t* s_w1 = s_w2 + ((c2.w_bytes + 127u) & ~127u);
    uint8_t* s_ring = s_w1 + ((c1.w_bytes + 127u) & ~127u);
    uint8_t* s_xring = s_ring + (size_t)kStages * c2.stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_xring + (size_t)kF12XRing * c1.stage_bytes);
    uint64_t* full = bars;                          // [8]  leader: one arrive per epilogue-1 warp of both CTAs
    uint64_t* empty = full + kStages;               // [8]  multicast commit from the leader
    uint64_t* xfull = empty + kStages;              // [4]  leader: own expect_tx arrive + the peer's relay; peer: own arrive
    uint64_t* xempty = xfull + kF12XRing;           // [4]  multicast commit
    uint64_t* w_full = xempty + kF12XRing;          // [1]
    uint64_t* a1_full = w_full + 1;                 // [2]  multicast commit: MMA1 of a tile has completed
    uint64_t* t_full = a1_full + 2;                 // [2]  multicast commit: MMA2 of a tile has completed
    uint64_t* t_empty = t_full + 2;                 // [2]  leader only: one arrive per epilogue-2 warp of both CTAs
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_empty + 2);
    // as an offset from the shared-memory array, so that the compiler keeps the accesses in the shared state space (through a
    // uintptr_t round trip they became generic LD / ST: long-scoreboard latency and a queue shared with the global stores)
    float* s_bias = reinterpret_cast<float*>(smem + (((uint32_t)(reinterpret_cast<uint8_t*>(t_empty + 3) - smem) + 15u) & ~15u));  // [0,128) conv1, [128,256) conv2

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < kStages; ++i) {
                umma::mbar_init(&full[i], 2 * 4);  // the four lane-group warps that own the stage's chunk, in both CTAs
                umma::mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < kF12XRing; ++i) {
                umma::mbar_init(&xfull[i], rank == 0 ? 2u : 1u);
                umma::mbar_init(&xempty[i], 1);
            }
            umma::mbar_init(w_full, rank == 0 ? 2u : 1u);
            for (int i = 0; i < 2; ++i) {
                umma::mbar_init(&a1_full[i], 1);
                umma::mbar_init(&t_full[i], 1);
                umma::mbar_init(&t_empty[i], 2 * kEpilogueWarps);
            }
            umma::fence_barrier_init();
        }
        __syncwarp();
        umma::tmem_alloc2(s_tmem, 512);
    }
    for (int i = (int)threadIdx.x; i < 128; i += kF12Threads) {
        s_bias[i] = __ldg(c1.bias + i);
        s_bias[128 + i] = __ldg(c2.bias + i);
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync();  // both CTAs' barriers exist before anyone arrives remotely
    umma::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;  // columns [0,256): A1 x 2, [256,512): A2 x 2
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    long long* const dbg = (blockIdx.x == 0) ? f.dbg : nullptr;
    auto stamp = [&](uint32_t it, int slot) {
        if (dbg && it < 48u) dbg[16u * it + (uint32_t)slot] = clock64();
    };
    uint32_t n_my = 0;  // tiles of this pair
    for (uint32_t t2 = pair; 2 * t2 < f.n_tiles; t2 += n_pairs) ++n_my;

    if (warp < (uint32_t)kF12ProducerWarps) {
        // ===================================== producers: weights once, then the X ring =====================================
        if (warp == 0) {
            if (lane == 0) umma::mbar_arrive_expect_tx(w_full, c1.w_bytes + c2.w_bytes);
            __syncwarp();
            {
                const uint32_t per = (((c2.w_bytes + 31u) / 32u) + 15u) & ~15u, off = lane * per;
                if (off < c2.w_bytes) umma::bulk_g2s(s_w2 + off, c2.w_img + (size_t)rank * c2.w_bytes + off, min(per, c2.w_bytes - off), w_full);
            }
            {
                const uint32_t per = (((c1.w_bytes + 31u) / 32u) + 15u) & ~15u, off = lane * per;
                if (off < c1.w_bytes) umma::bulk_g2s(s_w1 + off, c1.w_img + (size_t)rank * c1.w_bytes + off, min(per, c1.w_bytes - off), w_full);
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");  // X comes from the previous launch
        if (warp < (uint32_t)kF12XRing) {
            const DenseSeg& sg = c1.seg[0];
            const uint32_t pl_bytes = sg.nrows * 16u;
            uint32_t it = 0;
            for (uint32_t t2 = pair; 2 * t2 < f.n_tiles; t2 += n_pairs, ++it) {
                if (it % (uint32_t)kF12XRing != warp) continue;
                const uint32_t slot = warp, phase = (it / (uint32_t)kF12XRing) & 1u;
                const long long row0 = (long long)(2 * t2 + rank) * kF12OutRows;
                umma::mbar_wait(&xempty[slot], phase ^ 1u);
                uint8_t* stage = s_xring + (size_t)slot * c1.stage_bytes;
                if (lane == 0) umma::mbar_arrive_expect_tx(&xfull[slot], 2u * pl_bytes);
                __syncwarp();
                if (lane < 2u) {  // planes {hi, lo} of the one 8-channel group
                    const uint8_t* plane = sg.src + (unsigned long long)(lane * sg.groups) * sg.plane_stride;
                    umma::bulk_g2s(stage + sg.smem_off + lane * pl_bytes, plane + (row0 + sg.row_off) * 16ll, pl_bytes, &xfull[slot]);
                }
                __syncwarp();
                if (rank != 0) {
                    umma::mbar_wait(&xfull[slot], phase);
                    if (lane == 0) umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&xfull[slot]), 0));
                    __syncwarp();
                }
            }
        }
    } else if (warp == (uint32_t)kF12ProducerWarps) {
        if (rank == 0) {
            // ===================================== MMA issuer (leader) =====================================================
            const uint32_t idesc = umma::make_idesc_bf16_m256(128u);
            const uint32_t desc_hi = (uint32_t)(umma::make_desc(0, 0, 128) >> 32);
            const uint32_t nh = 64u;                                   // B rows held by each CTA (N / 2)
            const uint32_t b_step = (nh * 32u) >> 4;
            const uint32_t b1_base = (uint32_t)umma::make_desc(umma::smem_u32(s_w1), nh * 16u, 128);
            const uint32_t b2_base = (uint32_t)umma::make_desc(umma::smem_u32(s_w2), nh * 16u, 128);
            const uint32_t ring16 = umma::smem_u32(s_ring) >> 4, stage16 = c2.stage_bytes >> 4;
            const uint32_t xring16 = umma::smem_u32(s_xring) >> 4, xstage16 = c1.stage_bytes >> 4, aq16 = c1.a_q_off >> 4;
            const uint32_t x_hi = (uint32_t)umma::make_desc(c1.term[0].a_off, c1.term[0].a_lbo, 128);
            const uint32_t x_lo = (uint32_t)umma::make_desc(c1.term[0].a_off + c1.term[0].a_hl_off, c1.term[0].a_lbo, 128);
            uint32_t a_hi[3], a_lo[3];
            #pragma unroll
            for (int k = 0; k < 3; ++k) {
                a_hi[k] = (uint32_t)umma::make_desc(c2.term[k].a_off, c2.term[k].a_lbo, 128);
                a_lo[k] = (uint32_t)umma::make_desc(c2.term[k].a_off + c2.term[k].a_hl_off, c2.term[k].a_lbo, 128);
            }
            const int ksteps = c1.ksteps;
            umma::mbar_wait(w_full, 0);
            auto issue_mma1 = [&](uint32_t j) {  // conv1 of this pair's tile j -> A1[j & 1]
                const uint32_t slot = j % (uint32_t)kF12XRing, phase = (j / (uint32_t)kF12XRing) & 1u;
                umma::mbar_wait(&xfull[slot], phase);
                umma::tc_fence_after();
                if (umma::elect_one()) {
                    const uint32_t d_addr = tmem_base + (j & 1u) * 128u;
                    uint32_t sa = xring16 + slot * xstage16, bq = b1_base, acc = 0;
                    for (int q = 0; q < ksteps; ++q, sa += aq16) {  // one asm statement per triple (umma::mma2_stage3_bf16 has the reason)
                        umma::mma2_triple_bf16(d_addr, x_hi + sa, x_lo + sa, bq, b_step, desc_hi, idesc, acc);
                        bq += 2 * b_step;
                        acc = 1;
                    }
                    umma::mma2_commit_mc(&xempty[slot]);
                    umma::mma2_commit_mc(&a1_full[j & 1u]);
                }
                __syncwarp();
            };
            if (n_my) issue_mma1(0);
            for (uint32_t it = 0; it < n_my; ++it) {
                if (lane == 0) stamp(it, 0);
                if (it + 1 < n_my) issue_mma1(it + 1);
                if (lane == 0) stamp(it, 1);
                const uint32_t buf = it & 1u, use = it >> 1;
                umma::mbar_wait(&t_empty[buf], (use & 1u) ^ 1u);
                umma::tc_fence_after();
                if (lane == 0) stamp(it, 2);
                const uint32_t d_addr = tmem_base + 256u + buf * 128u;
                uint32_t b_cur = b2_base;
                for (int st = 0; st < kStages; ++st) {
                    umma::mbar_wait(&full[st], it & 1u);
                    umma::tc_fence_after();
                    if (lane == 0) stamp(it, 3 + st);
                    if (umma::elect_one()) {
                        const uint32_t sa = ring16 + (uint32_t)st * stage16;
                        umma::mma2_stage3_bf16(d_addr, a_hi[0] + sa, a_lo[0] + sa, a_hi[1] + sa, a_lo[1] + sa, a_hi[2] + sa, a_lo[2] + sa, b_cur, b_step, desc_hi,
                                               idesc, st ? 1u : 0u);
                        // epilogue-1 refills the ring one 32-channel chunk (two stages) at a time: one release per chunk
                        if (st & 1) umma::mma2_commit_mc(&empty[st]);
                    }
                    b_cur += 2 * b_step * 3u;
                    __syncwarp();
                }
                if (umma::elect_one()) umma::mma2_commit_mc(&t_full[buf]);
                __syncwarp();
                if (lane == 0) stamp(it, 11);
            }
        } else {
            umma::mbar_wait(w_full, 0);
            if (lane == 0) umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(w_full), 0));
        }
    } else if (warp < (uint32_t)(kF12ProducerWarps + 1 + kEpilogueWarps)) {
        // ===================================== epilogue-2: A2 -> Y2 in HBM (own CTA's 124 rows) ===============================
        const uint32_t lane_grp = (warp & 3u) * 32u;
        const uint32_t half = (warp - (uint32_t)(kF12ProducerWarps + 1)) >> 2;
        const uint32_t m = lane_grp + lane;
        uint32_t it = 0;
        // site-row lookups one tile ahead: this warp never waits long at t_full, so loads issued in the same tile had their whole
        // latency exposed at first use
        const bool keep = m < (uint32_t)kF12OutRows;
        int msc_next[kMaxScatter];
        scatter_rows(c2, (unsigned long long)(2 * pair + rank) * kF12OutRows + m, msc_next);
        for (uint32_t t2 = pair; 2 * t2 < f.n_tiles; t2 += n_pairs, ++it) {
            const uint32_t buf = it & 1u, use = it >> 1;
            const unsigned long long row = (unsigned long long)(2 * t2 + rank) * kF12OutRows + m;
            int msc[kMaxScatter];
            #pragma unroll
            for (int k = 0; k < kMaxScatter; ++k) msc[k] = keep ? msc_next[k] : -1;
            if (2 * (t2 + n_pairs) < f.n_tiles) scatter_rows(c2, (unsigned long long)(2 * (t2 + n_pairs) + rank) * kF12OutRows + m, msc_next);
            umma::mbar_wait(&t_full[buf], use & 1u);
            umma::tc_fence_after();
            if (threadIdx.x == 32u * (kF12ProducerWarps + 1)) stamp(it, 15);
            const uint32_t t_addr = tmem_base + (lane_grp << 16) + 256u + buf * 128u;
            for (int c0 = (int)half * 32; c0 < 128; c0 += 64) {
                uint32_t v[32];
                umma::tmem_ld32(t_addr + (uint32_t)c0, v);
                umma::tmem_ld_wait();
                float x[32];
                #pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(s_bias + 128 + c0 + j);
                    x[j] = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f);
                    x[j + 1] = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                    x[j + 2] = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f);
                    x[j + 3] = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
                }
                epilogue_store_chunk(c2, row, c0, x, msc, !keep);
            }
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) umma::mbar_arrive(&t_empty[buf]);
                else umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&t_empty[buf]), 0));
            }
        }
    } else {
        // ===================================== epilogue-1: A1 -> the A ring (+ Y1's scatter copies) ==========================
        const uint32_t lane_grp = (warp & 3u) * 32u;
        const uint32_t half = (warp - (uint32_t)(kF12ProducerWarps + 1 + kEpilogueWarps)) >> 2;  // chunks half, half + 2
        const uint32_t m = lane_grp + lane;
        const uint32_t pl_bytes = c2.seg[0].nrows * 16u;  // one plane of a stage: 132 rows x 16 B
        uint32_t it = 0;
        // site-row lookups one tile ahead (as in epilogue-2); rows 124 .. 127 are rows 0 .. 3 of the next tile, which scatters them
        const bool keep1 = m < (uint32_t)kF12OutRows;
        int msc_next[kMaxScatter];
        scatter_rows(c1, (unsigned long long)(2 * pair + rank) * kF12OutRows + m, msc_next);
        for (uint32_t t2 = pair; 2 * t2 < f.n_tiles; t2 += n_pairs, ++it) {
            const uint32_t buf = it & 1u, use = it >> 1;
            int msc[kMaxScatter];
            #pragma unroll
            for (int k = 0; k < kMaxScatter; ++k) msc[k] = keep1 ? msc_next[k] : -1;
            if (2 * (t2 + n_pairs) < f.n_tiles) scatter_rows(c1, (unsigned long long)(2 * (t2 + n_pairs) + rank) * kF12OutRows + m, msc_next);
            umma::mbar_wait(&a1_full[buf], use & 1u);
            umma::tc_fence_after();
            const bool stamper = threadIdx.x == 32u * (kF12ProducerWarps + 1 + kEpilogueWarps);
            if (stamper) stamp(it, 12);
            const uint32_t t_addr = tmem_base + (lane_grp << 16) + buf * 128u;
            for (int c = (int)half; c < 4; c += 2) {  // 32 channels = stages 2c, 2c + 1
                umma::mbar_wait(&empty[2 * c + 1], (it & 1u) ^ 1u);  // committed after the second stage of the chunk
                uint32_t v[32];
                umma::tmem_ld32(t_addr + 32u * (uint32_t)c, v);
                umma::tmem_ld_wait();
                float x[32];
                #pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(s_bias + 32 * c + j);
                    x[j] = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f);
                    x[j + 1] = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                    x[j + 2] = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f);
                    x[j + 3] = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
                }
                uint4 vh[4], vl[4];
                #pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t hi[4], lo[4];
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float x0 = x[8 * g + 2 * j], x1 = x[8 * g + 2 * j + 1];
                        const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
                        const __nv_bfloat162 e = __floats2bfloat162_rn(x0 - __uint_as_float(hb << 16), x1 - __uint_as_float(hb & 0xffff0000u));
                        hi[j] = hb;
                        lo[j] = *reinterpret_cast<const uint32_t*>(&e);
                    }
                    vh[g] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    vl[g] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                // stage s = 2c + (g >> 1), planes {hi g0, hi g1, lo g0, lo g1}: group g of the chunk is plane (g & 1) / 2 + (g & 1)
                #pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint8_t* st = s_ring + (size_t)(2 * c + (g >> 1)) * c2.stage_bytes + c2.seg[0].smem_off + m * 16u;
                    *reinterpret_cast<uint4*>(st + (uint32_t)(g & 1) * pl_bytes) = vh[g];
                    *reinterpret_cast<uint4*>(st + (uint32_t)(2 + (g & 1)) * pl_bytes) = vl[g];
                }
                umma::fence_proxy_async();  // the tensor core (async proxy) reads what these threads just wrote
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) {
                        umma::mbar_arrive(&full[2 * c]);
                        umma::mbar_arrive(&full[2 * c + 1]);
                    } else {
                        // plain remote arrive: the fence above already ordered this CTA's writes for its own tensor-core reads;
                        // a cluster-scope release here costs a GPU-wide membar per chunk (it waits for the scatter stores)
                        umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&full[2 * c]), 0));
                        umma::mbar_arrive_cluster(umma::mapa(umma::smem_u32(&full[2 * c + 1]), 0));
                    }
                }
                // Y1's scatter copies (compact rows are consecutive, so these fill whole lines).  AFTER the hand-over: the proxy fence
                // above is a memory barrier for this thread, and in front of it these global stores made every chunk wait for their
                // acknowledgement (HM_F12_STAMPS: ~4 000 cycles per chunk; 3 300 with the stores behind the arrive).  Giving them to
                // four dedicated warps instead was slower still: one warp needs ~2 200 cycles per 32-column chunk.
                #pragma unroll
                for (int k = 0; k < kMaxScatter; ++k) {
                    if (msc[k] >= 0) {
                        uint8_t* q_hi = c1.sc_out[k] + (unsigned long long)(4 * c) * c1.sc_plane_stride + (unsigned long long)msc[k] * 16ull;
                        uint8_t* q_lo = q_hi + (unsigned long long)c1.out_groups * c1.sc_plane_stride;
                        #pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            *reinterpret_cast<uint4*>(q_hi + g * c1.sc_plane_stride) = vh[g];
                            *reinterpret_cast<uint4*>(q_lo + g * c1.sc_plane_stride) = vl[g];
                        }
                    }
                }
                if (stamper) stamp(it, c < 2 ? 13 : 14);
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
