// umma.cuh -- thin inline-PTX wrappers for the sm_100a features the CNN kernels use:
// mbarrier, 1-D bulk async copies (TMA engine, UBLKCP), tcgen05 TMEM allocation / MMA / commit / load,
// and the shared-memory matrix descriptor + instruction descriptor encodings.
//
// Layout convention used everywhere in this engine: K-major operands WITHOUT swizzle.  A core matrix is
// 8 rows x 16 bytes stored contiguously (row pitch 16 B); the descriptor gives the byte distance between
// core matrices that are adjacent in K (LBO) and adjacent in M/N (SBO).  Because the start address only has
// to be 16-byte aligned, a row-shifted view of an activation tile (a convolution tap) is just another
// descriptor over the same shared memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hm {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of the (converged) warp; ptxas recognises elect.sync and keeps the guarded tcgen05 operands in uniform registers.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a while when the phase is still open).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// ---- bulk async copy global -> shared (1-D, TMA engine), completion on an mbarrier ------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Pulls a global range into L2 ahead of a later bulk copy (no completion tracking; bytes a multiple of 16).
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// ---- cp.async (16 B, gather path) -----------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// Makes the mbarrier track completion of all prior cp.async of this thread (arrival does not raise the pending count).
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- register redistribution between warp roles (whole warpgroups = 4 consecutive warps, .sync.aligned) -------------------------
template <int kRegs>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// ---- tcgen05: TMEM allocation ---------------------------------------------------------------------------------
// Whole-warp, .sync.aligned.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- clusters and cta_group::2 ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Address of the same shared-memory location in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 over a CTA pair: each CTA supplies its 128 rows of A and its N/2 rows of B; issued by the leader CTA only.
__device__ __forceinline__ void mma2_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the barrier at this offset in BOTH CTAs of the pair when the MMAs issued so far have completed.
__device__ __forceinline__ void mma2_commit_mc(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// Same with the A operand in TENSOR MEMORY (cute's SM100_MMA_F16BF16_2x1SM_TS): row m of this CTA's 128 rows is TMEM lane m, K runs
// along the columns, two bf16 per 32-bit column (element 2c in the low half of column c), so a K = 16 tile is 8 columns.  The layout
// and the rate were checked on a B200 (tools/mma_ts_probe.cu): exact results, and N = 96 / 64 take 48 / 46 cycles per MMA instead of
// the 64 of the shared-memory form (no 4 KB A fetch).
__device__ __forceinline__ void mma2_ts_bf16_w(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, {%6, %6, %6, %6, %6, %6, %6, %6}, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// The nine MMAs of one ring stage of a three-term op (three taps x {hi*hi, lo*hi, hi*lo}) as ONE asm statement.  The thread that
// issues the MMAs is the kernels' real limit: measured in situ they run at 70 - 92 cycles each against 64 in a tight loop, and
// tools/mma_contention_probe.cu reproduces it -- operands recomputed in vector registers in front of every MMA (one integer op + one
// register-to-uniform move each, what nine separate asm statements compile to) cost 95 cycles per MMA, 129 with epilogue warps
// competing for the scheduler.  Here the inputs cross into uniform registers once per stage.
// a_h / a_l: low descriptor words of the three taps' hi and lo views; b0: low word of the stage's first weight tile, tiles
// {hi, lo} per term b_step apart; acc = 0 overwrites the accumulator with the first product.
__device__ __forceinline__ void mma2_stage3_bf16(uint32_t tmem_d, uint32_t ah0, uint32_t al0, uint32_t ah1, uint32_t al1, uint32_t ah2, uint32_t al2,
                                                 uint32_t b0, uint32_t b_step, uint32_t desc_hi, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 da, dl, db, dc;\n\t.reg .b32 bb;\n\t"
        "setp.ne.b32 p, %11, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 da, {%1, %9};\n\t"
        "mov.b64 dl, {%2, %9};\n\t"
        "mov.b64 db, {%7, %9};\n\t"
        "add.u32 bb, %7, %8;\n\t"
        "mov.b64 dc, {bb, %9};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %10, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], dl, db, %10, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, dc, %10, t;\n\t"
        "mov.b64 da, {%3, %9};\n\t"
        "mov.b64 dl, {%4, %9};\n\t"
        "add.u32 bb, bb, %8;\n\t"
        "mov.b64 db, {bb, %9};\n\t"
        "add.u32 bb, bb, %8;\n\t"
        "mov.b64 dc, {bb, %9};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %10, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], dl, db, %10, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, dc, %10, t;\n\t"
        "mov.b64 da, {%5, %9};\n\t"
        "mov.b64 dl, {%6, %9};\n\t"
        "add.u32 bb, bb, %8;\n\t"
        "mov.b64 db, {bb, %9};\n\t"
        "add.u32 bb, bb, %8;\n\t"
        "mov.b64 dc, {bb, %9};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %10, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], dl, db, %10, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, dc, %10, t;\n\t}"
        ::"r"(tmem_d), "r"(ah0), "r"(al0), "r"(ah1), "r"(al1), "r"(ah2), "r"(al2), "r"(b0), "r"(b_step), "r"(desc_hi), "r"(idesc), "r"(acc)
        : "memory");
}
// One split-precision triple (a_h b0, a_l b0, a_h (b0 + b_step)) as one asm statement (conv1-form k-steps of the fused kernel).
__device__ __forceinline__ void mma2_triple_bf16(uint32_t tmem_d, uint32_t a_h, uint32_t a_l, uint32_t b0, uint32_t b_step, uint32_t desc_hi, uint32_t idesc,
                                                 uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 da, dl, db, dc;\n\t.reg .b32 bb;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 dl, {%2, %5};\n\t"
        "mov.b64 db, {%3, %5};\n\t"
        "add.u32 bb, %3, %4;\n\t"
        "mov.b64 dc, {bb, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], dl, db, %6, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, dc, %6, t;\n\t}"
        ::"r"(tmem_d), "r"(a_h), "r"(a_l), "r"(b0), "r"(b_step), "r"(desc_hi), "r"(idesc), "r"(acc)
        : "memory");
}
// The six MMAs of one ring step of the chain kernel (two 16-channel stages of one term) as one asm statement; operand A in shared
// memory (stage 1's views a_step further) ...
__device__ __forceinline__ void mma2_step6_bf16(uint32_t tmem_d, uint32_t a_h, uint32_t a_l, uint32_t a_step, uint32_t b0, uint32_t b_tile, uint32_t desc_hi,
                                                uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 da, dl, db, dc;\n\t.reg .b32 bb, ah, al;\n\t"
        "setp.ne.b32 p, %8, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 da, {%1, %6};\n\t"
        "mov.b64 dl, {%2, %6};\n\t"
        "mov.b64 db, {%4, %6};\n\t"
        "add.u32 bb, %4, %5;\n\t"
        "mov.b64 dc, {bb, %6};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %7, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], dl, db, %7, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, dc, %7, t;\n\t"
        "add.u32 ah, %1, %3;\n\t"
        "add.u32 al, %2, %3;\n\t"
        "mov.b64 da, {ah, %6};\n\t"
        "mov.b64 dl, {al, %6};\n\t"
        "add.u32 bb, bb, %5;\n\t"
        "mov.b64 db, {bb, %6};\n\t"
        "add.u32 bb, bb, %5;\n\t"
        "mov.b64 dc, {bb, %6};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %7, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], dl, db, %7, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, dc, %7, t;\n\t}"
        ::"r"(tmem_d), "r"(a_h), "r"(a_l), "r"(a_step), "r"(b0), "r"(b_tile), "r"(desc_hi), "r"(idesc), "r"(acc)
        : "memory");
}
// ... or in tensor memory (t_h / t_l: TMEM addresses of the hi and lo halves of the packed map; stage 1 is 8 columns further).
__device__ __forceinline__ void mma2_step6_ts_bf16(uint32_t tmem_d, uint32_t t_h, uint32_t t_l, uint32_t b0, uint32_t b_tile, uint32_t desc_hi, uint32_t idesc,
                                                   uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p, t;\n\t.reg .b64 db, dc;\n\t.reg .b32 bb, ah, al, z;\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b32 z, 0;\n\t"
        "mov.b64 db, {%3, %5};\n\t"
        "add.u32 bb, %3, %4;\n\t"
        "mov.b64 dc, {bb, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %6, {z, z, z, z, z, z, z, z}, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%2], db, %6, {z, z, z, z, z, z, z, z}, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], dc, %6, {z, z, z, z, z, z, z, z}, t;\n\t"
        "add.u32 ah, %1, 8;\n\t"
        "add.u32 al, %2, 8;\n\t"
        "add.u32 bb, bb, %4;\n\t"
        "mov.b64 db, {bb, %5};\n\t"
        "add.u32 bb, bb, %4;\n\t"
        "mov.b64 dc, {bb, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [ah], db, %6, {z, z, z, z, z, z, z, z}, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [al], db, %6, {z, z, z, z, z, z, z, z}, t;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [ah], dc, %6, {z, z, z, z, z, z, z, z}, t;\n\t}"
        ::"r"(tmem_d), "r"(t_h), "r"(t_l), "r"(b0), "r"(b_tile), "r"(desc_hi), "r"(idesc), "r"(acc)
        : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_m256(uint32_t n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

// ---- tcgen05: descriptors ---------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor): bits [0,14) address>>4,
// [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout type = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::f16, BF16 x BF16 -> FP32, both operands K-major, M = 128
// (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6), a_format BF16 = 1 @ [7,10), b_format BF16 = 1 @ [10,13),
// N>>3 @ [17,23), M>>4 @ [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16_m128(uint32_t n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- tcgen05: MMA / commit (single thread) -----------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, with both descriptors given as (low word, shared high word) so that descriptor arithmetic stays 32-bit.
// Call under `if (elect_one())` with warp-uniform operands.
__device__ __forceinline__ void mma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- split precision, second form: fp16 x fp16 main product + ONE e4m3 x e4m3 MMA carrying both correction products ---------------
// (a_f + r) (w_f + w_r) ~= a_f w_f + [a_h8 | a_l8] . [w_l8 ; w_h8]: the corrections are accumulated FIRST, scaled by 2^15 so that
// they fit the e4m3 range, and the first fp16 MMA of the tile folds them in with D = A*B + D * 2^-15 (scale-input-d, an immediate
// of kind::f16).  Two tensor-core instructions per (16 channels, tap) instead of the three of the bf16 hi/lo form.
constexpr uint32_t kCorrScaleLog2 = 15;  // = kActLoScaleLog2 + kWgtHiScaleLog2 (dense_gemm.cuh); the ISA allows 0..15
__device__ __forceinline__ void mma_f8_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma2_f8_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f16 with D = A*B + D * 2^-kCorrScaleLog2 (always accumulating)
__device__ __forceinline__ void mma_f16_scaled_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p, %5;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(kCorrScaleLog2)
        : "memory");
}
__device__ __forceinline__ void mma2_f16_scaled_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p, %5;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(kCorrScaleLog2)
        : "memory");
}
// Instruction descriptors of the second form: FP16 x FP16 (kind::f16, a/b format 0) and E4M3 x E4M3 (kind::f8f6f4, a/b format 0)
// have the same bits -- c_format F32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16_m128(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); }
__host__ __device__ constexpr uint32_t make_idesc_f16_m256(uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((256u >> 4) << 24); }

__device__ __forceinline__ void mma_commit_if(uint64_t* bar, uint32_t issue)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)), "r"(issue)
        : "memory");
}
// Arrives on the mbarrier when all tcgen05.mma issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- tcgen05: TMEM -> registers (whole warp; lane i of warp w reads TMEM lane 32*(w%4)+i) -----------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- tcgen05: registers -> TMEM (whole warp; lane i of warp w writes TMEM lane 32*(w%4)+i, N consecutive columns) -------------
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace umma
}  // namespace hm
