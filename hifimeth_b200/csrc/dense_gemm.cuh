// dense_gemm.cuh -- the one tensor-core kernel of the CNN path.
//
//   out[r][0..N) = act( bias + sum_k  In_k[r + shift_k][0..Cin) . W_k[Cin x N] )        r = tile rows (128 per tile)
//
// Every layer of the dilated dense plan (see cnn_tensor.cu) is an instance of this: a few row-shifted views of
// channel-plane activation maps, multiplied by small weight blocks.  bf16 split precision: activations and
// weights are stored as hi + lo bf16 pairs and every product is evaluated as hi*hi + lo*hi + hi*lo on the tcgen05
// tensor cores with fp32 accumulation in TMEM (max probability error 6e-5 vs fp32, tests/test_dense_plan.py).
//
// Activation map layout in HBM ("plane layout"):  [hl][g][row][8] bf16,  hl in {hi, lo}, g = channel / 8.
// A 16-byte unit is 8 consecutive channels of one row; a plane is all rows of one 8-channel group.  Loaded into
// shared memory plane by plane (1-D bulk copies on the TMA engine), a plane IS a column of UMMA core matrices
// (8 rows x 16 B, SWIZZLE_NONE, K-major), so a convolution tap is a descriptor whose start address is advanced by
// shift * 16 bytes -- no im2col, no re-load per tap.
//
// CTA = 17 warps, persistent over tiles (grid = #SMs):
//   warps 0-7 producers: weights once (resident for the whole launch), then the activation ring; ring slot s is always
//             filled by warp s (a warp has one bulk-copy instruction in flight at a time, tools/bulk_copy_probe.cu, so
//             the number of stages in flight is the number of producer warps)
//   warp 8    MMA issuer (one elected lane): 3 tcgen05.mma per term and k-step, commit -> ring slot free
//   warps 9-16 epilogue: tcgen05.ld the fp32 accumulator (double buffered in TMEM), bias + ReLU, split to hi/lo
//             bf16, coalesced 16-byte stores in plane layout (or the fc2 head -> logits)
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "umma.cuh"

namespace hm {

constexpr int kTileRows = 128;
constexpr int kMaxTerms = 3;
constexpr int kMaxSegs = 3;
constexpr int kMaxScatter = 5;
constexpr int kProducerWarps = 8;
constexpr int kEpilogueWarps = 8;  // two per TMEM lane group; each takes half of the accumulator columns
constexpr int kDenseThreads = 32 * (kProducerWarps + 1 + kEpilogueWarps);

struct DenseSeg {
    const uint8_t* src;           // plane 0 (hi, g = 0), row 0 of the input map
    unsigned long long plane_stride;  // bytes between consecutive planes in HBM
    int32_t row_off;              // rows [tile*128 + row_off, +nrows) are staged; gather: rows gather_rows[tile*128 + m] + row_off
    uint32_t nrows;               // rows staged per plane (128 + span of the shifts using this segment; gather: 128)
    uint32_t gather;              // 1: this segment's rows are picked through DenseOp::gather_rows (16-byte cp.async per row)
    uint32_t groups;              // Cin / 8 of the source map (lo planes start at plane index `groups`)
    uint32_t smem_off;            // byte offset of this segment inside a ring stage
};

struct DenseTerm {
    uint32_t a_off;     // byte offset of the term's hi view inside a ring stage (segment + row shift * 16)
    uint32_t a_hl_off;  // distance from the hi planes to the lo planes of that segment
    uint32_t a_lbo;     // byte distance between the two 8-channel halves of a k-step
};

struct DenseOp {
    DenseSeg seg[kMaxSegs];
    DenseTerm term[kMaxTerms];
    int32_t n_segs, n_terms;
    int32_t n_stages;        // ring stages per tile (Cin / 16; 1 for the conv1 form)
    int32_t ksteps;          // k16 steps per stage and term (1; conv1 form: ceil(taps / 2))
    uint32_t a_q_off;        // byte advance of the A view per k-step inside a stage (conv1 form: 32)
    int32_t planes_per_seg;  // planes copied per segment and stage (4 = {hi,lo} x 2 groups; conv1 form: 2;
                             // gathered conv1 form: 2 * gather_taps, plane p = {hl = p / taps, row + p % taps})
    int32_t gather_taps;     // > 0: gathered conv1 form
    const uint32_t* gather_rows;  // [n_tiles * 128] source row of every compact row (gather segments)
    uint32_t stage_bytes;    // bytes of one ring stage
    int32_t ring;            // ring depth
    const uint8_t* w_img;    // packed weights: tiles [stage][kstep][term][hl], each [2][N][8] bf16 (N*32 bytes)
    uint32_t w_bytes;
    const float* bias;       // [N]
    int32_t n;               // output channels (multiple of 16, <= 256)
    uint32_t tmem_cols;      // power of two >= 2 * n
    uint32_t n_tiles;
    int32_t mode;            // 0: ReLU -> hi/lo plane map;  1: ReLU -> fc2 -> logits
    uint8_t* out;            // mode 0: plane 0 row 0 of the output map
    unsigned long long out_plane_stride;
    uint32_t out_groups;     // mode 0: 8-channel groups of the whole output map (lo planes start there)
    uint32_t out_g0;         // mode 0: first group this launch writes (an op may be split over output channels)
    const float* w2;         // mode 1: [2][n]
    const float* b2;         // mode 1: [2]
    float* logits;           // mode 1: [rows][2]
    // Scatter (dense ops whose output is later read at site rows): besides its own map the epilogue copies row r into
    // compact row site_of_row[r - sc_shift[k]] of compact buffer k, so that the compact ops read gathered operands with
    // full-line bulk copies (a 16-byte gather costs a 128-byte DRAM fetch: 8x read amplification, profiles/).
    int32_t n_scatter;
    int32_t sc_shift[kMaxScatter];
    uint8_t* sc_out[kMaxScatter];     // plane 0 (hi, g = 0) of compact buffer k (same channel count as the output map)
    unsigned long long sc_plane_stride;
    const int32_t* site_of_row;       // [rows] compact row of the site whose s-row this is, or -1
    long long* dbg;          // variant & 32: CTA 0 writes clock64 stamps: [0..255] stage issue, [256..511] stage full seen by the MMA warp
    uint32_t variant;        // experiments (tools/dense_microbench.py): 1 = hi*hi pass only, 2 = no MMA, 4 = no epilogue stores, 8 = no epilogue work, 16 = plain arrive instead of tcgen05.commit on the ring (only with 2), 64 = producers free-run (no consumer)
};

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b)
{
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// Epilogue of one chunk of NG 8-column groups (32 or 16 columns) of one row: ReLU'd values -> hi/lo bf16 -> the op's own map
// (plane layout) and the compact scatter copies.  msc[k] = compact row for scatter k, or -1.
template <int NG>
__device__ __forceinline__ void epilogue_store_groups(const DenseOp& op, unsigned long long row, int c0, const float (&f)[8 * NG],
                                                      const int (&msc)[kMaxScatter], bool skip_store)
{
    uint4 vh[NG], vl[NG];
    #pragma unroll
    for (int g = 0; g < NG; ++g) {
        uint32_t hi[4], lo[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            // hi = bf16(x) for two values in one cvt; lo = bf16(x - hi), hi widened back with integer ops
            const float x0 = f[8 * g + 2 * j], x1 = f[8 * g + 2 * j + 1];
            const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
            const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
            const __nv_bfloat162 e = __floats2bfloat162_rn(x0 - __uint_as_float(hb << 16), x1 - __uint_as_float(hb & 0xffff0000u));
            hi[j] = hb;
            lo[j] = *reinterpret_cast<const uint32_t*>(&e);
        }
        vh[g] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        vl[g] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (skip_store) return;
    const uint32_t g0 = op.out_g0 + ((uint32_t)c0 >> 3);
    {
        uint8_t* p_hi = op.out + (unsigned long long)g0 * op.out_plane_stride + row * 16ull;
        uint8_t* p_lo = p_hi + (unsigned long long)op.out_groups * op.out_plane_stride;
        #pragma unroll
        for (int g = 0; g < NG; ++g) {
            *reinterpret_cast<uint4*>(p_hi + g * op.out_plane_stride) = vh[g];
            *reinterpret_cast<uint4*>(p_lo + g * op.out_plane_stride) = vl[g];
        }
    }
    // compact copies: msc[k] = -1 for rows that feed no site (and for k >= n_scatter)
    #pragma unroll
    for (int k = 0; k < kMaxScatter; ++k) {
        if (msc[k] >= 0) {
            uint8_t* q_hi = op.sc_out[k] + (unsigned long long)g0 * op.sc_plane_stride + (unsigned long long)msc[k] * 16ull;
            uint8_t* q_lo = q_hi + (unsigned long long)op.out_groups * op.sc_plane_stride;
            #pragma unroll
            for (int g = 0; g < NG; ++g) {
                *reinterpret_cast<uint4*>(q_hi + g * op.sc_plane_stride) = vh[g];
                *reinterpret_cast<uint4*>(q_lo + g * op.sc_plane_stride) = vl[g];
            }
        }
    }
}

__device__ __forceinline__ void epilogue_store_chunk(const DenseOp& op, unsigned long long row, int c0, const float (&f)[32],
                                                     const int (&msc)[kMaxScatter], bool skip_store)
{
    epilogue_store_groups<4>(op, row, c0, f, msc, skip_store);
}

// Map-form epilogue of one tile row: the N accumulator columns are split evenly between the two warps of a TMEM lane group at
// 16-column granularity (N = 96 -> 48 + 48: with 32-column chunks one warp had two chunks and the other one, and a chunk costs
// one warp ~2 200 cycles -- the N = 96 layers were epilogue-bound at 4 400 cycles per tile against 3 456 cycles of MMAs).
__device__ __forceinline__ void epilogue_map_row(const DenseOp& op, uint32_t t_addr, unsigned long long row, uint32_t half, const float* s_bias,
                                                 const int (&msc)[kMaxScatter], bool skip_store)
{
    const int n = op.n;
    const int mid = ((n >> 1) + 15) & ~15;
    int c0 = half ? mid : 0;
    const int c1 = half ? n : mid;
    while (c0 < c1) {
        if (c1 - c0 >= 32) {
            uint32_t v[32];
            umma::tmem_ld32(t_addr + (uint32_t)c0, v);
            umma::tmem_ld_wait();
            float f[32];
            #pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + j);
                f[j] = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f);
                f[j + 1] = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                f[j + 2] = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f);
                f[j + 3] = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
            }
            epilogue_store_groups<4>(op, row, c0, f, msc, skip_store);
            c0 += 32;
        } else {
            uint32_t v[16];
            umma::tmem_ld16(t_addr + (uint32_t)c0, v);
            umma::tmem_ld_wait();
            float f[16];
            #pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + j);
                f[j] = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f);
                f[j + 1] = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                f[j + 2] = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f);
                f[j + 3] = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
            }
            epilogue_store_groups<2>(op, row, c0, f, msc, skip_store);
            c0 += 16;
        }
    }
}

__device__ __forceinline__ void scatter_rows(const DenseOp& op, unsigned long long row, int (&msc)[kMaxScatter])
{
    #pragma unroll
    for (int k = 0; k < kMaxScatter; ++k) {
        msc[k] = -1;
        if (k < op.n_scatter && row >= (unsigned long long)op.sc_shift[k]) msc[k] = __ldg(op.site_of_row + (row - (unsigned long long)op.sc_shift[k]));
    }
}

// Shared memory: [weights image][ring stages][barriers]
template <bool kDbg>
__global__ void __launch_bounds__(kDenseThreads, 1) dense_gemm_kernel_t(const __grid_constant__ DenseOp op)
{
    const uint32_t variant = kDbg ? op.variant : 0u;  // experiment switches compile away in the product instantiation
    extern __shared__ __align__(128) uint8_t smem[];
    // the warp index through a shuffle: ptxas then knows it is warp-uniform, role branches become uniform branches and the MMA
    // issuer's loop counters, descriptors and barrier addresses can stay in uniform registers (cutlass::canonical_warp_idx_sync)
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    uint8_t* s_w = smem;
    uint8_t* s_ring = smem + ((op.w_bytes + 127u) & ~127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)op.ring * op.stage_bytes);
    uint64_t* full = bars;                  // [ring]
    uint64_t* empty = bars + op.ring;       // [ring]
    uint64_t* w_full = bars + 2 * op.ring;  // [1]
    uint64_t* t_full = w_full + 1;          // [2]
    uint64_t* t_empty = t_full + 2;         // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_empty + 2);
    // as an offset from the shared-memory array, so that the compiler keeps the accesses in the shared state space (through a
    // uintptr_t round trip they became generic LD / ST: long-scoreboard latency and a queue shared with the global stores)
    float* s_bias = reinterpret_cast<float*>(smem + (((uint32_t)(reinterpret_cast<uint8_t*>(t_empty + 3) - smem) + 15u) & ~15u));  // [768]: bias | fc2 row 0 | fc2 row 1

    bool any_gather = false, any_bulk = false;
    for (int s = 0; s < op.n_segs; ++s) {
        any_gather |= op.seg[s].gather != 0;
        any_bulk |= op.seg[s].gather == 0;
    }
    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < op.ring; ++i) {
                umma::mbar_init(&full[i], (any_bulk ? 1u : 0u) + (any_gather ? 32u : 0u));
                umma::mbar_init(&empty[i], 1);
            }
            umma::mbar_init(w_full, 1);
            for (int i = 0; i < 2; ++i) {
                umma::mbar_init(&t_full[i], 1);
                umma::mbar_init(&t_empty[i], 32 * kEpilogueWarps);
            }
            umma::fence_barrier_init();
        }
        __syncwarp();
        umma::tmem_alloc(s_tmem, op.tmem_cols);
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t acc_stride = op.tmem_cols >> 1;

    // Programmatic dependent launch: the next launch of the stream may start its prologue (barrier init, TMEM allocation,
    // weight load) on SMs this grid has left; everything that reads or overwrites activations waits for the previous grid.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp < (uint32_t)kProducerWarps) {
        // ===================================== producers ====================================================
        if (warp == 0) {
            if (lane == 0) umma::mbar_arrive_expect_tx(w_full, op.w_bytes);
            __syncwarp();
            const uint32_t per = (((op.w_bytes + 31u) / 32u) + 15u) & ~15u;  // one copy per lane, one instruction
            const uint32_t off = lane * per;
            if (off < op.w_bytes) umma::bulk_g2s(s_w + off, op.w_img + off, min(per, op.w_bytes - off), w_full);
        }
        // Bulk copies issued by one warp execute one INSTRUCTION at a time (about one memory latency each, however many
        // lanes take part): 1 lane x 2 KB per instruction = 0.75 TB/s chip-wide, 32 lanes = 6.7 TB/s, 8 warps x 1 lane =
        // 6.2 TB/s (tools/bulk_copy_probe.cu).  So all copies of a stage leave in one instruction (lane c issues copy c)
        // and consecutive stages are filled by different warps.
        // A ring slot must always be filled by the same warp (parity waits are only valid one phase ahead): warp s owns
        // slot s, ring <= kProducerWarps.
        const uint32_t n_prod = (uint32_t)op.ring;
        asm volatile("griddepcontrol.wait;" ::: "memory");  // activations (and the site-row index) come from earlier launches
        uint32_t stage_no = 0;
        uint32_t stage_tx = 0, ncopies = 0;
        int bulk_seg[kMaxSegs] = {0, 0, 0};
        for (int s = 0; s < op.n_segs; ++s)
            if (!op.seg[s].gather) {
                stage_tx += op.seg[s].nrows * 16u * op.planes_per_seg;
                bulk_seg[ncopies / (uint32_t)op.planes_per_seg] = s;
                ncopies += (uint32_t)op.planes_per_seg;
            }
        for (uint32_t tile = blockIdx.x; tile < op.n_tiles; tile += gridDim.x) {
            const long long row0 = (long long)tile * kTileRows;
            uint32_t grow[4] = {0, 0, 0, 0};
            if (any_gather) {
                #pragma unroll
                for (int j = 0; j < 4; ++j) grow[j] = __ldg(op.gather_rows + row0 + lane + 32 * j);
            }
            for (int st = 0; st < op.n_stages; ++st, ++stage_no) {
                if (stage_no % n_prod != warp) continue;
                const uint32_t slot = stage_no % (uint32_t)op.ring, phase = (stage_no / (uint32_t)op.ring) & 1u;
                if (variant & 64u) { if (stage_no >= n_prod) umma::mbar_wait(&full[slot], phase ^ 1u); }
                else umma::mbar_wait(&empty[slot], phase ^ 1u);
                uint8_t* stage = s_ring + (size_t)slot * op.stage_bytes;
                if (any_bulk) {
                    if (lane == 0) umma::mbar_arrive_expect_tx(&full[slot], stage_tx);
                    if ((variant & 32u) && blockIdx.x == 0 && lane == 0 && stage_no < 256) op.dbg[stage_no] = clock64();
                    __syncwarp();
                    const uint32_t c = lane;  // this lane's copy of the stage, if < ncopies
                    if (c < ncopies) {
                        const uint32_t p = c % (uint32_t)op.planes_per_seg;
                        const DenseSeg& sg = op.seg[bulk_seg[c / (uint32_t)op.planes_per_seg]];
                        // planes of a stage: normal form {hi g0, hi g1, lo g0, lo g1}; conv1 form {hi, lo}
                        const uint32_t hl = (op.planes_per_seg == 4) ? (p >> 1) : p;
                        const uint32_t g = (op.planes_per_seg == 4) ? (uint32_t)(2 * st) + (p & 1u) : 0u;
                        const uint32_t pl_bytes = sg.nrows * 16u;
                        const uint8_t* plane = sg.src + (unsigned long long)(hl * sg.groups + g) * sg.plane_stride;
                        umma::bulk_g2s(stage + sg.smem_off + p * pl_bytes, plane + (row0 + sg.row_off) * 16ll, pl_bytes, &full[slot]);
                    }
                }
                if (any_gather) {
                    for (int s = 0; s < op.n_segs; ++s) {
                        const DenseSeg& sg = op.seg[s];
                        if (!sg.gather) continue;
                        const uint32_t pl_bytes = sg.nrows * 16u;
                        for (int p = 0; p < op.planes_per_seg; ++p) {
                            // normal form {hi g0, hi g1, lo g0, lo g1}; gathered conv1 form {hi tap 0.., lo tap 0..}
                            uint32_t hl, g, extra = 0;
                            if (op.gather_taps > 0) { hl = (uint32_t)p / (uint32_t)op.gather_taps; g = 0; extra = (uint32_t)p % (uint32_t)op.gather_taps; }
                            else { hl = (uint32_t)p >> 1; g = (uint32_t)(2 * st + (p & 1)); }
                            const uint8_t* plane = sg.src + (unsigned long long)(hl * sg.groups + g) * sg.plane_stride;
                            uint8_t* dst = stage + sg.smem_off + p * pl_bytes;
                            #pragma unroll
                            for (int j = 0; j < 4; ++j)
                                umma::cp_async16(dst + (lane + 32 * j) * 16u, plane + ((long long)grow[j] + sg.row_off + extra) * 16ll);
                        }
                    }
                    umma::cp_async_mbar_arrive_noinc(&full[slot]);
                }
                __syncwarp();
            }
        }
    } else if (warp == (uint32_t)kProducerWarps) {
        // ===================================== MMA issuer ===================================================
        // One lane issues every tcgen05.mma of the CTA, so the per-MMA instruction count matters: all descriptors are
        // built once and advanced by adding 16-byte units to their address field (bits [0,14), never carries out).
        const uint32_t idesc = umma::make_idesc_bf16_m128((uint32_t)op.n);
        // Only the low word of a descriptor changes (address bits [0,14), LBO bits [16,30)); the high word (SBO, version)
        // is the same for every operand, so descriptor arithmetic is 32-bit adds on the low word.
        const uint32_t desc_hi = (uint32_t)(umma::make_desc(0, 0, 128) >> 32);
        const uint32_t b_step = ((uint32_t)op.n * 32u) >> 4;                 // one weight tile, in descriptor units
        const uint32_t b_base = (uint32_t)umma::make_desc(umma::smem_u32(s_w), (uint32_t)op.n * 16u, 128);
        const uint32_t ring16 = umma::smem_u32(s_ring) >> 4, stage16 = op.stage_bytes >> 4, aq16 = op.a_q_off >> 4;
        uint32_t a_hi[kMaxTerms], a_lo[kMaxTerms];
        #pragma unroll
        for (int k = 0; k < kMaxTerms; ++k) {
            a_hi[k] = (uint32_t)umma::make_desc(op.term[k].a_off, op.term[k].a_lbo, 128);
            a_lo[k] = (uint32_t)umma::make_desc(op.term[k].a_off + op.term[k].a_hl_off, op.term[k].a_lbo, 128);
        }
        const int n_terms = op.n_terms, ksteps = op.ksteps, n_stages = op.n_stages, ring = op.ring;
        umma::mbar_wait(w_full, 0);
        uint32_t slot = 0, phase = 0, it = 0;
        for (uint32_t tile = blockIdx.x; tile < op.n_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, use = it >> 1;
            if ((variant & 32u) && blockIdx.x == 0 && lane == 0 && it < 32) op.dbg[768 + 2 * it] = clock64();
            umma::mbar_wait(&t_empty[buf], (use & 1u) ^ 1u);
            umma::tc_fence_after();
            if ((variant & 32u) && blockIdx.x == 0 && lane == 0 && it < 32) op.dbg[768 + 2 * it + 1] = clock64();
            const uint32_t d_addr = tmem_base + buf * acc_stride;
            uint32_t acc = 0;
            uint32_t b_cur = b_base;
            for (int st = 0; st < n_stages; ++st) {
                if (variant & 64u) continue;
                umma::mbar_wait(&full[slot], phase);
                umma::tc_fence_after();
                if ((variant & 32u) && blockIdx.x == 0 && lane == 0 && it * n_stages + st < 256) op.dbg[256 + it * n_stages + st] = clock64();
                if (umma::elect_one()) {
                    uint32_t sa = ring16 + slot * stage16, bq = b_cur;
                    uint32_t a0 = acc;
                    for (int q = 0; q < ksteps; ++q, sa += aq16) {
                        #pragma unroll
                        for (int k = 0; k < kMaxTerms; ++k) {
                            if (k < n_terms) {
                                if (!(variant & 2u)) umma::mma_bf16_w(d_addr, a_hi[k] + sa, bq, desc_hi, idesc, a0);
                                a0 = 1;
                                if (!(variant & 3u)) {
                                    umma::mma_bf16_w(d_addr, a_lo[k] + sa, bq, desc_hi, idesc, 1);
                                    umma::mma_bf16_w(d_addr, a_hi[k] + sa, bq + b_step, desc_hi, idesc, 1);
                                }
                                bq += 2 * b_step;
                            }
                        }
                    }
                    if (variant & 16u) umma::mbar_arrive(&empty[slot]);
                    else umma::mma_commit(&empty[slot]);
                }
                acc = 1;
                b_cur += 2 * b_step * (uint32_t)(ksteps * n_terms);
                __syncwarp();
                if ((variant & 32u) && blockIdx.x == 0 && lane == 0 && it * n_stages + st < 256) op.dbg[512 + it * n_stages + st] = clock64();
                if (++slot == (uint32_t)ring) { slot = 0; phase ^= 1u; }
            }
            if (umma::elect_one()) umma::mma_commit(&t_full[buf]);
            __syncwarp();
        }
    } else {
        // ===================================== epilogue ======================================================
        const uint32_t lane_grp = (warp & 3u) * 32u;  // TMEM lanes this warp may touch
        const uint32_t half = (warp - (uint32_t)(kProducerWarps + 1)) >> 2;  // which 32-column chunks this warp takes
        const uint32_t m = lane_grp + lane;            // row of the tile
        const int n = op.n;
        // bias (and the fc2 weights of the head form) live in shared memory for the whole launch
        for (int i = (int)threadIdx.x - 32 * (kProducerWarps + 1); i < n; i += 32 * kEpilogueWarps) {
            s_bias[i] = __ldg(op.bias + i);
            if (op.mode == 1) { s_bias[256 + i] = __ldg(op.w2 + i); s_bias[512 + i] = __ldg(op.w2 + n + i); }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpilogueWarps) : "memory");
        uint32_t it = 0;
        // site-row lookups one tile ahead (see dense_gemm2.cuh)
        int msc_next[kMaxScatter];
        scatter_rows(op, (unsigned long long)blockIdx.x * kTileRows + m, msc_next);
        for (uint32_t tile = blockIdx.x; tile < op.n_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, use = it >> 1;
            const unsigned long long row = (unsigned long long)tile * kTileRows + m;
            float l0 = 0.f, l1 = 0.f;
            int msc[kMaxScatter];
            #pragma unroll
            for (int k = 0; k < kMaxScatter; ++k) msc[k] = msc_next[k];
            if (tile + gridDim.x < op.n_tiles) scatter_rows(op, (unsigned long long)(tile + gridDim.x) * kTileRows + m, msc_next);
            umma::mbar_wait(&t_full[buf], use & 1u);
            umma::tc_fence_after();
            const uint32_t t_addr = tmem_base + (lane_grp << 16) + buf * acc_stride;
            // map form: the columns are split between the two warps of a lane group; head form: the first warp does all
            if (op.mode == 0 && !(variant & 8u)) epilogue_map_row(op, t_addr, row, half, s_bias, msc, (variant & 4u) != 0);
            const int c_end = (op.mode == 1 && half == 0 && !(variant & 8u)) ? n : 0;
            for (int c0 = 0; c0 < c_end; c0 += 32) {
                uint32_t v[32];
                umma::tmem_ld32(t_addr + (uint32_t)c0, v);
                umma::tmem_ld_wait();
                float f[32];
                #pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + j);
                    f[j] = fmaxf(__uint_as_float(v[j]) + bv.x, 0.f);
                    f[j + 1] = fmaxf(__uint_as_float(v[j + 1]) + bv.y, 0.f);
                    f[j + 2] = fmaxf(__uint_as_float(v[j + 2]) + bv.z, 0.f);
                    f[j + 3] = fmaxf(__uint_as_float(v[j + 3]) + bv.w, 0.f);
                }
                #pragma unroll
                for (int j = 0; j < 32; ++j) {
                    l0 = fmaf(f[j], s_bias[256 + c0 + j], l0);
                    l1 = fmaf(f[j], s_bias[512 + c0 + j], l1);
                }
            }
            umma::tc_fence_before();
            umma::mbar_arrive(&t_empty[buf]);
            if (op.mode == 1 && half == 0) {
                float2 o = make_float2(l0 + __ldg(op.b2), l1 + __ldg(op.b2 + 1));
                *reinterpret_cast<float2*>(op.logits + row * 2ull) = o;
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    if (warp == 0) umma::tmem_dealloc(tmem_base, op.tmem_cols);
}

// Product instantiation (no experiment switches) and the one tools/dense_microbench.py drives.
#define dense_gemm_kernel dense_gemm_kernel_t<false>
#define dense_gemm_kernel_dbg dense_gemm_kernel_t<true>

inline size_t dense_smem_bytes(const DenseOp& op)
{
    return ((op.w_bytes + 127u) & ~127u) + (size_t)op.ring * op.stage_bytes + (2 * op.ring + 5 + 1) * sizeof(uint64_t) + 16 + 768 * sizeof(float);
}

}  // namespace hm
