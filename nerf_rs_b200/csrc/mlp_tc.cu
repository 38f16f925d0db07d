// mlp_tc.cu -- K-mlp: fused multi-layer MLP on tcgen05 tensor cores (sm_100a).
//
// k_chain      persistent, one CTA per SM, one 128-sample tile in flight per CTA. Warp roles:
//                warp 0   weight-ring producer: cp.async.bulk (UBLKCP) of pre-swizzled bf16 weight
//                         chunks [N<=128][64] from the L2-resident packed stream into a 7-stage ring;
//                warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N<=128, K=16) with both
//                         operands in shared memory, accumulators in TMEM (2 blocks x 128 columns);
//                warp 2   TMEM allocator;
//                warps 4-7 epilogue: tcgen05.ld -> +bias -> ReLU -> bf16 -> 128B-swizzled smem panel
//                         that IS the next layer's A operand (never touches HBM in inference);
//                         in training the panels are bulk-stored to HBM for the weight gradients.
//              N is split in two 128-column blocks so block 0's epilogue overlaps block 1's MMAs and
//              the next layer's first K panels (see mlp_tc_plan.cpp for the hazard argument).
//              The positional encoding is computed in the tile prologue straight
//              into the smem A operand of fc1 (and reused by the skip layer).
// k_wgrad      dW^T[in x out] = sum over samples of P^T Q with both operands MN-major straight from
//              the saved panel images; fp32 accumulators stay in TMEM across a segment's whole tile
//              range and are flushed with red.global.add.f32. A CTA works through up to three
//              (unit, tile range) segments cut by tc_wgrad_partition (mlp_tc_plan.cpp) so that the
//              fitted cost is equal across the 148 CTAs; the sigma row of fc8 is a CUDA-core GEMV
//              inside the fc8 feature unit, fed by a d(sigma) loader warp.
// k_pack       gathers the flat f32 [out,in] parameter blob into the bf16 chunk streams.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"
#include "mlp_tc.h"
#include "ptx.cuh"

namespace {

constexpr uint32_t kSlotBytes = NERF_PANEL_BYTES;
constexpr uint32_t kSmemSlots = TC_NUM_SLOTS * kSlotBytes;
constexpr uint32_t kSmemRing = TC_NUM_STAGES * TC_STAGE_BYTES;
constexpr uint32_t kSmemTables = kSmemSlots + kSmemRing;      // bias | ops | jobs copies
constexpr uint32_t kTableBytes = 12288;
constexpr uint32_t kSmemBars = kSmemTables + kTableBytes;
constexpr uint32_t kChainSmem = kSmemBars + TC_NUM_BARS * 8 + 16;
constexpr int kEpiWarps = 8;
constexpr int kChainThreads = 32 * (4 + kEpiWarps);

struct ChainArgs {
    const MmaOp *ops;
    const EpiJob *jobs;
    int32_t n_ops, n_jobs;
    const uint8_t *wpack;
    const float *bias;
    int32_t bias_floats;
    int64_t n_samples;
    int32_t n_tiles, S;
    int32_t xyz_freqs, dir_freqs;
    const float *points;    // fwd in  [n][3]
    const float *dirs;      // fwd in  [rays][3]
    float *sigma;           // fwd out [n]
    float *rgba;            // fwd out [n][4]; bwd in
    const float *d_sigma;   // bwd in [n]
    const float *d_rgba;    // bwd in [n][4]
    uint8_t *save_base;     // per-tile panel area written by this launch (act or grad), or NULL
    int32_t save_slots;
    uint32_t *mask_base;    // [tile][mask_slots][128][8] (v1 kernel; the CTA-pair kernel stores each tile slot word-major, [8][128])
    int32_t mask_slots;
    unsigned long long *trace;  // debug: [3 roles][kTraceEvents][2] clock64 stamps of CTA 0, or NULL
};
constexpr int kTraceEvents = 2048;

__device__ __forceinline__ void trace_event(const ChainArgs &a, int role, int idx, unsigned long long t0, unsigned long long t1) {
    if (a.trace && blockIdx.x == 0 && idx < kTraceEvents) {
        a.trace[((size_t)role * kTraceEvents + idx) * 2] = t0;
        a.trace[((size_t)role * kTraceEvents + idx) * 2 + 1] = t1;
    }
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t panel_chunk_addr(uint32_t slot_addr, uint32_t row, uint32_t chunk) {
    return slot_addr + row * 128u + (((chunk ^ row) & 7u) << 4);
}

// [v, sin(2^k v), cos(2^k v)]_k for a 3-vector -> f[0 .. 3+6*freqs), zero padded; this warp packs and
// stores only the 16-byte chunks [kCh0, kCh1) of the row. sin/cos of the base angle are accurate
// (sincosf); octaves use the double-angle recurrence (abs error <= 2^k * 1e-7, far below bf16 resolution).
template <int kMaxF, int kCh0, int kCh1>
__device__ __forceinline__ void encode_panel(uint32_t slot_addr, uint32_t row, const float v[3], int freqs) {
    float f[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) f[i] = 0.f;
    f[0] = v[0]; f[1] = v[1]; f[2] = v[2];
    float s[3], c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) sincosf(v[d], &s[d], &c[d]);
#pragma unroll
    for (int k = 0; k < kMaxF; ++k) {
        if (k < freqs) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                f[3 + 6 * k + d] = s[d];
                f[3 + 6 * k + 3 + d] = c[d];
                const float s2 = 2.f * s[d] * c[d];
                const float c2 = fmaf(-2.f * s[d], s[d], 1.f);
                s[d] = s2;
                c[d] = c2;
            }
        }
    }
#pragma unroll
    for (int ch = kCh0; ch < kCh1; ++ch) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = ptx::pack_bf16x2(f[8 * ch + 2 * e], f[8 * ch + 2 * e + 1]);
        st_shared_v4(panel_chunk_addr(slot_addr, row, ch), w[0], w[1], w[2], w[3]);
    }
}

// panel whose only non-zero entries are the first four bf16 of each row; half h writes chunks 4h..4h+3
__device__ __forceinline__ void write_sparse_panel(uint32_t slot_addr, uint32_t row, uint32_t h, uint32_t w0, uint32_t w1) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const bool first = (h == 0 && ch == 0);
        st_shared_v4(panel_chunk_addr(slot_addr, row, 4 * h + ch), first ? w0 : 0u, first ? w1 : 0u, 0u, 0u);
    }
}


// One 32-column group of a hidden-layer epilogue: accumulator registers -> (bias, activation | relu mask)
// -> 16 packed bf16x2 words. kKind selects the arithmetic at compile time.
template <bool kSave, uint8_t kKind>
__device__ __forceinline__ void epi_group(const uint32_t (&r)[32], const float *bias_g, uint32_t &mask, uint32_t (&w)[16]) {
    if (kKind == EK_RELU || kKind == EK_LINEAR) {
        const float4 *bp = reinterpret_cast<const float4 *>(bias_g);
        uint32_t signs = 0;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b = bp[j4];
            const float v0 = __uint_as_float(r[4 * j4 + 0]) + b.x;
            const float v1 = __uint_as_float(r[4 * j4 + 1]) + b.y;
            const float v2 = __uint_as_float(r[4 * j4 + 2]) + b.z;
            const float v3 = __uint_as_float(r[4 * j4 + 3]) + b.w;
            if (kSave && kKind == EK_RELU) {
                signs = __funnelshift_l(__float_as_uint(v0), signs, 1);
                signs = __funnelshift_l(__float_as_uint(v1), signs, 1);
                signs = __funnelshift_l(__float_as_uint(v2), signs, 1);
                signs = __funnelshift_l(__float_as_uint(v3), signs, 1);
            }
            if (kKind == EK_RELU) {
                w[2 * j4] = ptx::pack_bf16x2_relu(v0, v1);
                w[2 * j4 + 1] = ptx::pack_bf16x2_relu(v2, v3);
            } else {
                w[2 * j4] = ptx::pack_bf16x2(v0, v1);
                w[2 * j4 + 1] = ptx::pack_bf16x2(v2, v3);
            }
        }
        mask = ~signs;  // bit (31 - col) = pre-activation sign bit clear
    } else {
        const uint32_t m = (kKind == EK_DMASK) ? mask : 0xffffffffu;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const float v0 = (m & (0x80000000u >> (2 * p))) ? __uint_as_float(r[2 * p]) : 0.f;
            const float v1 = (m & (0x80000000u >> (2 * p + 1))) ? __uint_as_float(r[2 * p + 1]) : 0.f;
            w[p] = ptx::pack_bf16x2(v0, v1);
        }
    }
}

__device__ __forceinline__ void store_group(uint32_t slot_addr, uint32_t row, uint32_t cb, const uint32_t (&w)[16]) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        st_shared_v4(panel_chunk_addr(slot_addr, row, cb + c), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

// Hidden-layer epilogue job for one warp: `gcount` (1 or 2) groups starting at group g0 of the block.
template <bool kSave, uint8_t kKind>
__device__ __forceinline__ void epi_hidden(const EpiJob &j, uint32_t taddr, uint32_t sbase, const float *s_bias, uint32_t *mask_ptr,
                                           uint32_t row, int g0, int gcount, uint32_t acc_free_bar, int lane) {
    uint32_t r0[32], r1[32];
    uint32_t m0 = 0, m1 = 0;
    if (kKind == EK_DMASK) {
        m0 = mask_ptr[0];
        if (gcount > 1) m1 = mask_ptr[1];
    }
    ptx::tmem_ld32(taddr + g0 * 32, r0);
    if (gcount > 1) ptx::tmem_ld32(taddr + (g0 + 1) * 32, r1);
    ptx::tmem_ld_wait();
    // this warp's share of the accumulator is in registers: after the GEMM's last job, hand the set back
    if (acc_free_bar) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_free_bar);
    }
    uint32_t w[16];
    epi_group<kSave, kKind>(r0, s_bias + j.bias_off + g0 * 32, m0, w);
    store_group(sbase + (uint32_t)(j.out_slot + (g0 >> 1)) * kSlotBytes, row, (uint32_t)(g0 & 1) * 4u, w);
    if (gcount > 1) {
        epi_group<kSave, kKind>(r1, s_bias + j.bias_off + (g0 + 1) * 32, m1, w);
        store_group(sbase + (uint32_t)(j.out_slot + ((g0 + 1) >> 1)) * kSlotBytes, row, (uint32_t)((g0 + 1) & 1) * 4u, w);
    }
    if (kSave && kKind == EK_RELU && mask_ptr) {
        mask_ptr[0] = m0;
        if (gcount > 1) mask_ptr[1] = m1;
    }
}

template <bool kBwd, bool kSave>
__global__ void __launch_bounds__(kChainThreads, 1) k_chain(const ChainArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = ptx::smem_u32(smem);
    const uint32_t bars = sbase + kSmemBars;
    volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kSmemBars + TC_NUM_BARS * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int id) { return bars + 8u * (uint32_t)id; };

    // tables live in shared memory: with ~220 KB of the SM's 228 KB carved out as smem the L1 is tiny,
    // so per-op / per-job / bias reads from global would each pay an L2 round trip on the critical path
    float *s_bias = reinterpret_cast<float *>(smem + kSmemTables);
    MmaOp *s_ops = reinterpret_cast<MmaOp *>(smem + kSmemTables + ((a.bias_floats * 4 + 15) & ~15));
    EpiJob *s_jobs = reinterpret_cast<EpiJob *>(reinterpret_cast<uint8_t *>(s_ops) + ((a.n_ops * (int)sizeof(MmaOp) + 15) & ~15));
    for (int i = threadIdx.x; i < a.bias_floats; i += blockDim.x) s_bias[i] = a.bias[i];
    for (int i = threadIdx.x; i < a.n_ops * (int)(sizeof(MmaOp) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_ops)[i] = reinterpret_cast<const uint32_t *>(a.ops)[i];
    for (int i = threadIdx.x; i < a.n_jobs * (int)(sizeof(EpiJob) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_jobs)[i] = reinterpret_cast<const uint32_t *>(a.jobs)[i];

    if (threadIdx.x == 0) {
        if (sbase & 1023u) {
            printf("nerf_b200: dynamic smem base not 1024-aligned\n");
            __trap();
        }
        for (int s = 0; s < TC_NUM_STAGES; ++s) {
            ptx::mbar_init(bar(TC_BAR_FULL + s), 1);
            ptx::mbar_init(bar(TC_BAR_EMPTY + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(bar(TC_BAR_ACC_FULL + b), 1);
            ptx::mbar_init(bar(TC_BAR_ACC_FREE + b), kEpiWarps);
        }
        for (int g = 0; g < 4; ++g) ptx::mbar_init(bar(TC_BAR_READY + g), kEpiWarps);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(ptx::smem_u32(const_cast<uint32_t *>(tmem_ptr_smem)));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ================= weight-ring producer =================
        // The whole warp runs the loop (warp-uniform control flow keeps addresses in uniform registers);
        // one elected lane issues the bulk copy.
        uint32_t stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            for (int i = 0; i < a.n_ops; ++i) {
                const uint32_t w_off = s_ops[i].w_off;
                const uint32_t bytes = (uint32_t)s_ops[i].n * 128u;
                const unsigned long long tp0 = a.trace ? clock64() : 0;
                ptx::mbar_wait(bar(TC_BAR_EMPTY + stage), phase ^ 1u);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(bar(TC_BAR_FULL + stage), bytes);
                    ptx::bulk_g2s(sbase + kSmemSlots + stage * TC_STAGE_BYTES, a.wpack + w_off, bytes, bar(TC_BAR_FULL + stage));
                }
                __syncwarp();
                if (a.trace && lane == 0) trace_event(a, 2, (tile - blockIdx.x) / gridDim.x * a.n_ops + i, tp0, clock64());
                if (++stage == TC_NUM_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // Warp-uniform loop; tcgen05.mma / tcgen05.commit are issued by one elected lane. (A plain
        // `if (lane == 0)` makes ptxas wrap every UTCHMMA in an elect/branch convergence loop and
        // rebuild its uniform-register operands, ~100 cycles per instruction.)
        uint32_t stage = 0, phase = 0;
        // waiter-side parity per barrier id; "free"-type barriers start at 1 (first wait passes)
        uint32_t wph = (1u << (TC_BAR_ACC_FREE + 0)) | (1u << (TC_BAR_ACC_FREE + 1));
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            for (int i = 0; i < a.n_ops; ++i) {
                const MmaOp op = s_ops[i];
                const unsigned long long tm0 = a.trace ? clock64() : 0;
                if (op.wait0 != TC_NONE) {
                    ptx::mbar_wait(bar(op.wait0), (wph >> op.wait0) & 1u);
                    wph ^= 1u << op.wait0;
                }
                if (op.wait1 != TC_NONE) {
                    ptx::mbar_wait(bar(op.wait1), (wph >> op.wait1) & 1u);
                    wph ^= 1u << op.wait1;
                }
                const unsigned long long tm1 = a.trace ? clock64() : 0;
                ptx::mbar_wait(bar(TC_BAR_FULL + stage), phase);
                ptx::tc_fence_after();
                const unsigned long long tm2 = a.trace ? clock64() : 0;
                const uint32_t a_addr = sbase + (uint32_t)op.a_slot * kSlotBytes;
                const uint32_t b_addr = sbase + kSmemSlots + stage * TC_STAGE_BYTES;
                const uint32_t idesc = ptx::umma_idesc_bf16(128, op.n, 0, 0);
                const uint32_t d_tmem = tmem_base + (uint32_t)op.acc * 256u;
                const uint64_t ad0 = ptx::umma_desc_sw128(a_addr, 16, 1024);
                const uint64_t bd0 = ptx::umma_desc_sw128(b_addr, 16, 1024);
                const uint32_t acc_first = (op.flags & TC_OP_FIRST) ? 0u : 1u;
                if (ptx::elect_one()) {
                    // +32 B per K16 step == +2 in the descriptor's address field
                    ptx::umma_ss(d_tmem, ad0, bd0, idesc, acc_first);
                    if (op.kcount > 1) ptx::umma_ss(d_tmem, ad0 + 2u, bd0 + 2u, idesc, 1u);
                    if (op.kcount > 2) {
                        ptx::umma_ss(d_tmem, ad0 + 4u, bd0 + 4u, idesc, 1u);
                        ptx::umma_ss(d_tmem, ad0 + 6u, bd0 + 6u, idesc, 1u);
                    }
                    ptx::umma_commit(bar(TC_BAR_EMPTY + stage));
                    if (op.flags & TC_OP_COMMIT_ACC) ptx::umma_commit(bar(TC_BAR_ACC_FULL + op.acc));
                }
                __syncwarp();
                if (a.trace && lane == 0) {
                    const int ev = ((tile - blockIdx.x) / gridDim.x * a.n_ops + i) * 2;
                    trace_event(a, 0, ev, tm0, tm1);
                    trace_event(a, 0, ev + 1, tm2, clock64());
                }
                if (++stage == TC_NUM_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue warps =================
        // warp (q, h): TMEM lane quarter q (rows 32q..32q+31), column half h of each accumulator block
        const uint32_t we = (uint32_t)(warp - 4);
        const uint32_t q = we & 3u, h = we >> 2;
        const uint32_t row = q * 32u + (uint32_t)lane;
        const bool store_lane = (h == 0 && lane == 0);   // issues this row quarter's bulk stores
        uint32_t aph = 0;  // acc_full parities
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            const int64_t gs = (int64_t)tile * NERF_TILE_M + row;
            const bool valid = gs < a.n_samples;
            for (int ji = 0; ji < a.n_jobs; ++ji) {
                const EpiJob j = s_jobs[ji];
                const bool writes_e = (j.acc == TC_NONE) || (j.enc != ENC_NONE);
                if (kSave) {
                    // a panel may still be the source of an earlier bulk store of this row quarter: hidden
                    // panels rotate with period >= 2 store groups, slot E is rewritten back to back
                    if (store_lane) {
                        if (writes_e) ptx::bulk_wait_read<0>();
                        else ptx::bulk_wait_read<1>();
                    }
                    ptx::named_bar_sync(2 + q, 64);
                }
                const unsigned long long te0 = a.trace ? clock64() : 0;
                if (j.flags & TC_JOB_WAIT_ACC) {
                    ptx::mbar_wait(bar(TC_BAR_ACC_FULL + j.acc), (aph >> j.acc) & 1u);
                    aph ^= 1u << j.acc;
                    ptx::tc_fence_after();
                }
                const unsigned long long te1 = a.trace ? clock64() : 0;
                const uint32_t taddr = tmem_base + ((q * 32u) << 16) + (uint32_t)j.acc_col;
                bool wrote_smem = false;

                if (j.kind == EK_RELU || j.kind == EK_LINEAR || j.kind == EK_DMASK || j.kind == EK_DCOPY) {
                    const int gcount = j.ncols >> 6;            // 32-column groups per warp (1 or 2)
                    const int g0 = (int)h * gcount;
                    uint32_t *mask_ptr = nullptr;
                    if (j.mask_slot >= 0)
                        mask_ptr = a.mask_base + ((((size_t)tile * a.mask_slots + j.mask_slot) * NERF_TILE_M + row) * 8 + j.mask_word0 + g0);
                    const uint32_t fb = (j.flags & TC_JOB_RELEASE_ACC) ? bar(TC_BAR_ACC_FREE + j.acc) : 0u;
                    if (!kBwd) {
                        if (j.kind == EK_RELU) epi_hidden<kSave, EK_RELU>(j, taddr, sbase, s_bias, mask_ptr, row, g0, gcount, fb, lane);
                        else epi_hidden<kSave, EK_LINEAR>(j, taddr, sbase, s_bias, mask_ptr, row, g0, gcount, fb, lane);
                    } else {
                        if (j.kind == EK_DMASK) epi_hidden<kSave, EK_DMASK>(j, taddr, sbase, s_bias, mask_ptr, row, g0, gcount, fb, lane);
                        else epi_hidden<kSave, EK_DCOPY>(j, taddr, sbase, s_bias, mask_ptr, row, g0, gcount, fb, lane);
                    }
                    wrote_smem = true;
                } else if (j.kind == EK_SIGMA || j.kind == EK_RGBA) {
                    if (h == 0) {
                        uint32_t r[16];
                        ptx::tmem_ld16(taddr, r);
                        ptx::tmem_ld_wait();
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(bar(TC_BAR_ACC_FREE + j.acc));
                        if (j.kind == EK_SIGMA) {
                            if (valid) a.sigma[gs] = __uint_as_float(r[0]) + s_bias[j.bias_off];
                        } else if (valid) {
                            float4 o;
                            o.x = 1.f / (1.f + expf(-(__uint_as_float(r[0]) + s_bias[j.bias_off + 0])));
                            o.y = 1.f / (1.f + expf(-(__uint_as_float(r[1]) + s_bias[j.bias_off + 1])));
                            o.z = 1.f / (1.f + expf(-(__uint_as_float(r[2]) + s_bias[j.bias_off + 2])));
                            o.w = 1.f / (1.f + expf(-(__uint_as_float(r[3]) + s_bias[j.bias_off + 3])));
                            reinterpret_cast<float4 *>(a.rgba)[gs] = o;
                        }
                    } else if (lane == 0) {
                        ptx::mbar_arrive(bar(TC_BAR_ACC_FREE + j.acc));  // this half reads nothing
                    }
                }

                // ---- slot E producers (each column half writes four of the eight 16-byte chunks per row)
                const uint32_t e_addr = sbase + TC_SLOT_E * kSlotBytes;
                if (j.kind == EK_PROLOGUE_FWD || j.enc == ENC_X) {
                    float v[3] = {0.f, 0.f, 0.f};
                    if (valid) { v[0] = a.points[3 * gs]; v[1] = a.points[3 * gs + 1]; v[2] = a.points[3 * gs + 2]; }
                    if (h == 0) encode_panel<10, 0, 4>(e_addr, row, v, a.xyz_freqs);
                    else encode_panel<10, 4, 8>(e_addr, row, v, a.xyz_freqs);
                    wrote_smem = true;
                } else if (j.enc == ENC_D) {
                    if (h == 0) {
                        float v[3] = {0.f, 0.f, 0.f};
                        if (valid) {
                            const int64_t ray = gs / a.S;
                            v[0] = a.dirs[3 * ray]; v[1] = a.dirs[3 * ray + 1]; v[2] = a.dirs[3 * ray + 2];
                        }
                        encode_panel<4, 0, 4>(e_addr, row, v, a.dir_freqs);
                    } else {
                        write_sparse_panel(e_addr, row, 1, 0u, 0u);
                    }
                    wrote_smem = true;
                } else if (j.enc == ENC_DSIGMA) {
                    const float ds = valid ? a.d_sigma[gs] : 0.f;
                    write_sparse_panel(e_addr, row, h, ptx::pack_bf16x2(ds, 0.f), 0u);
                    wrote_smem = true;
                } else if (j.kind == EK_PROLOGUE_BWD) {
                    float4 y = make_float4(0.f, 0.f, 0.f, 0.f), d = y;
                    if (valid && h == 0) {
                        y = reinterpret_cast<const float4 *>(a.rgba)[gs];
                        d = reinterpret_cast<const float4 *>(a.d_rgba)[gs];
                    }
                    write_sparse_panel(e_addr, row, h, ptx::pack_bf16x2(d.x * y.x * (1.f - y.x), d.y * y.y * (1.f - y.y)),
                                       ptx::pack_bf16x2(d.z * y.z * (1.f - y.z), d.w * y.w * (1.f - y.w)));
                    wrote_smem = true;
                }

                const unsigned long long te2 = a.trace ? clock64() : 0;
                if (wrote_smem) ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to UMMA / bulk store
                __syncwarp();
                if (lane == 0) {
                    if (j.ready_bar != TC_NONE) ptx::mbar_arrive(bar(j.ready_bar));
                    if (j.enc_bar != TC_NONE) ptx::mbar_arrive(bar(j.enc_bar));
                }
                if (a.trace && we == 0 && lane == 0) {
                    const int ev = ((tile - blockIdx.x) / gridDim.x * a.n_jobs + ji) * 2;
                    trace_event(a, 1, ev, te0, te1);
                    trace_event(a, 1, ev + 1, te2, clock64());
                }
                if (kSave && (j.save_slot >= 0 || j.enc_save_slot >= 0)) {
                    ptx::named_bar_sync(2 + q, 64);   // both column halves of this row quarter are written
                    if (store_lane) {
                        // rows 32q..32q+31 of a panel image are one contiguous 4 KB block
                        uint8_t *tile_base = a.save_base + (size_t)tile * a.save_slots * kSlotBytes + q * 4096u;
                        if (j.save_slot >= 0) {
                            for (int p = 0; p < (j.ncols >> 6); ++p)
                                ptx::bulk_s2g(tile_base + (size_t)(j.save_slot + p) * kSlotBytes,
                                              sbase + (uint32_t)(j.out_slot + p) * kSlotBytes + q * 4096u, 4096u);
                        }
                        if (j.enc_save_slot >= 0)
                            ptx::bulk_s2g(tile_base + (size_t)j.enc_save_slot * kSlotBytes, e_addr + q * 4096u, 4096u);
                        ptx::bulk_commit();
                    }
                }
            }
        }
        if (kSave && store_lane) ptx::bulk_wait_all<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------- wgrad
constexpr int kEpiThreads = 128;
constexpr int kWgStages = 3;
constexpr uint32_t kWgStageBytes = 65536;
constexpr uint32_t kWgHalf = 8192;  // 64 sample rows of one panel
constexpr uint32_t kWgBars = kWgStages * kWgStageBytes;
constexpr uint32_t kWgSmem = kWgBars + 8 * 8 + 16 + 256 * 4 + 260 * 4 + kWgStages * 64 * 4;

struct WgradArgs {
    const WgradUnit *units;
    const WgradWork *work;
    const uint8_t *act_base;
    const uint8_t *grad_base;
    int32_t act_slots, grad_slots;
    float *grads;
};

// per-CTA wall-clock marks of the last k_wgrad launch (ns, %globaltimer): start, first stage landed, all MMAs done, end
__device__ unsigned long long g_wgrad_marks[256 * 4];
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256, 1) k_wgrad(const WgradArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = ptx::smem_u32(smem);
    const uint32_t bars = sbase + kWgBars;  // full[3], empty[3], done
    volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kWgBars + 64);
    float *s_bias = reinterpret_cast<float *>(smem + kWgBars + 80);
    float *s_sg = s_bias + 256;     // sigma-row slice (256) + its bias (1)
    float *s_dsg = s_sg + 260;      // [stage][64] d(sigma) of the stage's sample rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    const WgradWork wk = a.work[blockIdx.x];     // (uploaded by the host before the launch, not produced by the predecessor)
    if (threadIdx.x == 0 && blockIdx.x < 256) g_wgrad_marks[4 * blockIdx.x] = global_ns();
    if (wk.n_seg <= 0) return;  // uniform per CTA

    if (threadIdx.x == 0) {
        if (sbase & 1023u) __trap();
        for (int s = 0; s < kWgStages; ++s) {
            ptx::mbar_init(bars + 8 * s, 2);                // producer (expect_tx) + the d(sigma) loader warp
            ptx::mbar_init(bars + 8 * (kWgStages + s), 5);  // MMA commit + 4 epilogue warps
        }
        ptx::mbar_init(bars + 8 * (2 * kWgStages), 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(ptx::smem_u32(const_cast<uint32_t *>(tmem_ptr_smem)));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();   // the dgrad kernel's panels (and the zeroed gradient buffer) are complete from here on

    // A CTA works through up to kWgMaxSeg segments = (unit, tile range) pieces, so the 148 CTAs can split the units' total
    // cost evenly instead of in whole CTAs per unit. Ring stage / phase counters simply run on across segments (every role
    // performs the same number of iterations); the accumulator is flushed and the CTA re-synchronised between segments.
    uint32_t stage = 0, phase = 0;   // per-thread copies; advanced identically by the producer, MMA and epilogue threads
    for (int seg = 0; seg < wk.n_seg; ++seg) {
        const WgradUnit u = a.units[wk.seg[seg].unit];
        const int tile_end = wk.seg[seg].tile_end;
        const int n_p = u.n_p, n_q = u.n_q;
        const int N = 64 * n_q;
        const int mblocks = (n_p + 1) >> 1;
        const uint32_t mb_cols = mblocks > 2 ? (uint32_t)N : 256u;   // TMEM columns between M blocks (3 x 128 for the merged fc9 unit)
        const int n_iters = (tile_end - wk.seg[seg].tile_begin) * 2;
        for (int i = threadIdx.x; i < 256 + 260; i += blockDim.x) s_bias[i] = 0.f;
        __syncthreads();

        if (warp == 0) {
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(n_p + n_q) * kWgHalf;
                for (int it = 0; it < n_iters; ++it) {
                    // last tile first: the tail of what the dgrad kernel just wrote is still in L2
                    const int tile = tile_end - 1 - (it >> 1);
                    const uint32_t half = (uint32_t)(it & 1) * kWgHalf;
                    ptx::mbar_wait(bars + 8 * (kWgStages + stage), phase ^ 1u);
                    ptx::mbar_arrive_expect_tx(bars + 8 * stage, bytes);
                    const uint32_t dst = sbase + stage * kWgStageBytes;
                    const uint8_t *ab = a.act_base + (size_t)tile * a.act_slots * kSlotBytes + half;
                    const uint8_t *gb = a.grad_base + (size_t)tile * a.grad_slots * kSlotBytes + half;
                    for (int i = 0; i < n_p; ++i)
                        ptx::bulk_g2s(dst + i * kWgHalf, ab + (size_t)u.p_slot[i] * kSlotBytes, kWgHalf, bars + 8 * stage);
                    for (int i = 0; i < n_q; ++i)
                        ptx::bulk_g2s(dst + (n_p + i) * kWgHalf, gb + (size_t)u.q_slot[i] * kSlotBytes, kWgHalf, bars + 8 * stage);
                    if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)N, 1, 1);
                for (int it = 0; it < n_iters; ++it) {
                    ptx::mbar_wait(bars + 8 * stage, phase);
                    ptx::tc_fence_after();
                    const uint32_t st_addr = sbase + stage * kWgStageBytes;
                    for (int mb = 0; mb < mblocks; ++mb) {
                        for (uint32_t k = 0; k < 4; ++k) {
                            // 16 sample rows per K step = 2048 B; 64-element M/N blocks are kWgHalf apart
                            const uint64_t ad = ptx::umma_desc_sw128(st_addr + (uint32_t)(2 * mb) * kWgHalf + k * 2048u, kWgHalf, 1024);
                            const uint64_t bd = ptx::umma_desc_sw128(st_addr + (uint32_t)n_p * kWgHalf + k * 2048u, kWgHalf, 1024);
                            ptx::umma_ss(tmem_base + (uint32_t)mb * mb_cols, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    ptx::umma_commit(bars + 8 * (kWgStages + stage));
                    if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
                }
                ptx::umma_commit(bars + 8 * (2 * kWgStages));
            }
        } else if (warp == 3) {
            // second arrival of every full barrier. For an fc8 feature unit it first stages column 0 of the d(sigma) panel for
            // the stage's 64 sample rows (fp32 strip): the scattered global loads run as far ahead as the ring allows and
            // never sit on the consumers' path. Ping-pong registers: half tile it+1 is in flight while it is handed over.
            const bool sg = u.sg_slot >= 0;
            const int r0 = lane, r1 = lane + 32;
            auto fetch = [&](int it, uint16_t &v0, uint16_t &v1) {
                const int tile = tile_end - 1 - (it >> 1);
                const uint8_t *gb = a.grad_base + ((size_t)tile * a.grad_slots + (size_t)u.sg_slot) * kSlotBytes + (size_t)(it & 1) * kWgHalf;
                v0 = __ldg(reinterpret_cast<const uint16_t *>(gb + r0 * 128 + ((r0 & 7) << 4)));
                v1 = __ldg(reinterpret_cast<const uint16_t *>(gb + r1 * 128 + ((r1 & 7) << 4)));
            };
            auto step = [&](int it, uint16_t c0, uint16_t c1, uint16_t &n0, uint16_t &n1) {
                if (sg && it + 1 < n_iters) fetch(it + 1, n0, n1);
                ptx::mbar_wait(bars + 8 * (kWgStages + stage), phase ^ 1u);
                if (sg) {
                    s_dsg[stage * 64 + r0] = __uint_as_float((uint32_t)c0 << 16);
                    s_dsg[stage * 64 + r1] = __uint_as_float((uint32_t)c1 << 16);
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bars + 8 * stage);
                if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
            };
            uint16_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
            if (sg) fetch(0, a0, a1);
            for (int it = 0; it < n_iters; it += 2) {   // two half tiles per tile: n_iters is even
                step(it, a0, a1, b0, b1);
                step(it + 1, b0, b1, a0, a1);
            }
        } else if (warp >= 4) {
            const int tid = threadIdx.x - 128;
            const bool has_bias = u.b_base >= 0;
            const int ccols = N >> 3;            // 16-byte chunk columns of Q
            const int c8 = tid % ccols;
            const int rg = tid / ccols;
            const int n_rg = kEpiThreads / ccols;
            const int rpr = 64 / n_rg;           // rows per row group per half tile
            float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            // sigma row of fc8 on the CUDA cores (units with sg_slot >= 0, n_p <= 4): thread = 8 input features x a group of rows
            const bool has_sg = u.sg_slot >= 0 && n_p <= 4;
            const int pcols = n_p * 8;            // 16-byte chunk columns of P
            const int pc8 = tid % pcols;
            const int prg = tid / pcols;
            const int n_prg = kEpiThreads / pcols;
            const int rpp = (64 + n_prg - 1) / n_prg;   // sample rows per row group per half tile
            const bool sg_active = has_sg && prg < n_prg;
            float ssum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float dsum = 0.f;
            for (int it = 0; it < n_iters; ++it) {
                ptx::mbar_wait(bars + 8 * stage, phase);
                if (seg == 0 && it == 0 && tid == 0 && blockIdx.x < 256) g_wgrad_marks[4 * blockIdx.x + 1] = global_ns();
                if (sg_active) {
                    const uint32_t paddr = sbase + stage * kWgStageBytes + (uint32_t)(pc8 >> 3) * kWgHalf;
                    const float *dsg = s_dsg + stage * 64;
#pragma unroll 4
                    for (int rr = 0; rr < rpp; ++rr) {
                        const uint32_t r = (uint32_t)(prg * rpp + rr);
                        if (r < 64u) {
                            const float d = dsg[r];
                            uint32_t w0, w1, w2, w3;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                         : "r"(panel_chunk_addr(paddr, r, (uint32_t)(pc8 & 7))));
                            ssum[0] += d * __uint_as_float(w0 << 16); ssum[1] += d * __uint_as_float(w0 & 0xffff0000u);
                            ssum[2] += d * __uint_as_float(w1 << 16); ssum[3] += d * __uint_as_float(w1 & 0xffff0000u);
                            ssum[4] += d * __uint_as_float(w2 << 16); ssum[5] += d * __uint_as_float(w2 & 0xffff0000u);
                            ssum[6] += d * __uint_as_float(w3 << 16); ssum[7] += d * __uint_as_float(w3 & 0xffff0000u);
                            if (pc8 == 0) dsum += d;
                        }
                    }
                }
                if (has_bias) {
                    const uint32_t qaddr = sbase + stage * kWgStageBytes + (uint32_t)(n_p + (c8 >> 3)) * kWgHalf;
                    for (int rr = 0; rr < rpr; ++rr) {
                        const uint32_t r = (uint32_t)(rg * rpr + rr);
                        uint32_t w0, w1, w2, w3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                     : "r"(panel_chunk_addr(qaddr, r, (uint32_t)(c8 & 7))));
                        bsum[0] += __uint_as_float(w0 << 16); bsum[1] += __uint_as_float(w0 & 0xffff0000u);
                        bsum[2] += __uint_as_float(w1 << 16); bsum[3] += __uint_as_float(w1 & 0xffff0000u);
                        bsum[4] += __uint_as_float(w2 << 16); bsum[5] += __uint_as_float(w2 & 0xffff0000u);
                        bsum[6] += __uint_as_float(w3 << 16); bsum[7] += __uint_as_float(w3 & 0xffff0000u);
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bars + 8 * (kWgStages + stage));
                if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
            }
            if (has_bias) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(&s_bias[c8 * 8 + e], bsum[e]);
            }
            if (sg_active) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(&s_sg[pc8 * 8 + e], ssum[e]);
                if (pc8 == 0) atomicAdd(&s_sg[256], dsum);
            }
            ptx::named_bar_sync(1, kEpiThreads);
            if (has_bias) {
                for (int n = tid; n < u.n_valid; n += kEpiThreads) atomicAdd(a.grads + u.b_base + n, s_bias[n]);
            }
            if (has_sg) {
                for (int m = tid; m < u.m_valid; m += kEpiThreads) atomicAdd(a.grads + u.sg_w_base + m, s_sg[m]);
                if (tid == 0 && u.sg_b_base >= 0) atomicAdd(a.grads + u.sg_b_base, s_sg[256]);
            }
        }
        // ---- flush the TMEM-resident dW^T block: lane = input index (contiguous in dW rows -> coalesced REDs). All eight
        //      warps take part (warp w reads TMEM lanes 32 (w % 4)..; the service warps take the odd (M block, column group)
        //      pairs): 29 -> 15 us for a 256 x 256 block.
        __syncwarp();
        {
            const uint32_t q = (uint32_t)(warp & 3);
            const int half = warp >> 2;
            ptx::mbar_wait(bars + 8 * (2 * kWgStages), (uint32_t)(seg & 1));
            ptx::tc_fence_after();
            if (seg == wk.n_seg - 1 && threadIdx.x == 128 && blockIdx.x < 256) g_wgrad_marks[4 * blockIdx.x + 2] = global_ns();
            const int n_groups = N >> 5, n_pairs = mblocks * n_groups;
            for (int pi = half; pi < n_pairs; pi += 2) {
                const int mb = pi / n_groups, g = pi % n_groups;
                const int m = mb * 128 + (int)(q * 32) + lane;
                uint32_t r[32];
                ptx::tmem_ld32(tmem_base + ((q * 32u) << 16) + (uint32_t)mb * mb_cols + (uint32_t)g * 32u, r);
                ptx::tmem_ld_wait();
                if (m < u.m_valid) {
#pragma unroll
                    for (int jn = 0; jn < 32; ++jn) {
                        const int n = g * 32 + jn;
                        if (n < u.n_valid) atomicAdd(a.grads + u.w_base + (int64_t)n * u.w_row_stride + m, __uint_as_float(r[jn]));
                    }
                }
            }
        }
        ptx::tc_fence_before();
        __syncthreads();     // every warp has drained its accumulator columns before the next segment's first MMA overwrites them
        ptx::tc_fence_after();
    }
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
    if (threadIdx.x == 0 && blockIdx.x < 256) g_wgrad_marks[4 * blockIdx.x + 3] = global_ns();
}

// ---------------------------------------------------------------------------------- pack
__global__ void k_pack_chunks(const PackChunk *chunks, const float *__restrict__ params, uint8_t *__restrict__ dst) {
    const PackChunk pc = chunks[blockIdx.x];
    for (int i = threadIdx.x; i < pc.n_rows * 8; i += blockDim.x) {
        const int r = i >> 3, c8 = i & 7;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c8 * 8 + 2 * e + h;
                v[h] = (r < pc.valid_rows && c < pc.valid_cols)
                           ? params[pc.src_base + (int64_t)r * pc.row_stride + (int64_t)c * pc.col_stride]
                           : 0.f;
            }
            w[e] = ptx::pack_bf16x2(v[0], v[1]);
        }
        *reinterpret_cast<uint4 *>(dst + pc.dst_off + (uint32_t)r * 128u + (uint32_t)(((c8 ^ r) & 7) << 4)) =
            make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// one launch for everything that follows an Adam step: forward chunk stream, backward (transposed) chunk stream, padded biases
__global__ void k_pack_all(const PackChunk *fwd_chunks, int n_fwd, uint8_t *__restrict__ fwd_dst, const PackChunk *bwd_chunks, int n_bwd,
                           uint8_t *__restrict__ bwd_dst, const PackBias *pb, int n_bias, const float *__restrict__ params,
                           float *__restrict__ bias_dst) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x;
    if (b >= n_fwd + n_bwd) {
        if (blockIdx.y != 0) return;
        const PackBias e = pb[b - n_fwd - n_bwd];
        for (int i = threadIdx.x; i < e.padded; i += blockDim.x) bias_dst[e.dst_off + i] = i < e.count ? params[e.src_base + i] : 0.f;
        return;
    }
    const PackChunk pc = b < n_fwd ? fwd_chunks[b] : bwd_chunks[b - n_fwd];
    uint8_t *dst = b < n_fwd ? fwd_dst : bwd_dst;
    // gridDim.y blocks share a chunk: the kernel is latency-bound (a few dependent-free gathers per thread), not bandwidth-bound
    for (int i = threadIdx.x + blockDim.x * blockIdx.y; i < pc.n_rows * 8; i += blockDim.x * gridDim.y) {
        const int r = i >> 3, c8 = i & 7;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c8 * 8 + 2 * e + h;
                v[h] = (r < pc.valid_rows && c < pc.valid_cols)
                           ? params[pc.src_base + (int64_t)r * pc.row_stride + (int64_t)c * pc.col_stride]
                           : 0.f;
            }
            w[e] = ptx::pack_bf16x2(v[0], v[1]);
        }
        *reinterpret_cast<uint4 *>(dst + pc.dst_off + (uint32_t)r * 128u + (uint32_t)(((c8 ^ r) & 7) << 4)) =
            make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void k_pack_bias(const PackBias *pb, int n, const float *__restrict__ params, float *__restrict__ dst) {
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        const PackBias e = pb[b];
        for (int i = threadIdx.x; i < e.padded; i += blockDim.x) dst[e.dst_off + i] = i < e.count ? params[e.src_base + i] : 0.f;
    }
}

template <typename T>
T *upload(const std::vector<T> &v) {
    T *d = nullptr;
    if (v.empty()) return nullptr;
    if (cudaMalloc(&d, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

struct DevProgram {
    MmaOp *ops = nullptr;
    EpiJob *jobs = nullptr;
    PackChunk *chunks = nullptr;
    uint8_t *wpack = nullptr;
    int n_ops = 0, n_jobs = 0, n_chunks = 0;
};

}  // namespace

struct TcState {
    NetGeom g;
    TcPlan plan;
    int num_sms = 0;
    int64_t max_tiles = 0;
    DevProgram fwd_train, fwd_infer, bwd;
    PackBias *d_pbias = nullptr;
    float *d_bias = nullptr;
    WgradUnit *d_units = nullptr;
    WgradWork *d_work = nullptr;
    int64_t work_tiles = -1;
    uint8_t *d_act = nullptr, *d_grad = nullptr;
    uint32_t *d_mask = nullptr;
    uint64_t bias_version = 1;   // bumped by tc_pack_weights (constant-bank copy of the biases, v2)
    int version = 2;   // 2: CTA-pair two-lane chain (mlp_tc2.cu); 1: one tile per CTA (k_chain above)
    Lane2Program *fwd_train2 = nullptr, *fwd_infer2 = nullptr, *bwd2 = nullptr;
    std::string err;
};

static bool upload_program(const TcProgram &p, DevProgram &d, bool own_weights) {
    d.n_ops = (int)p.ops.size();
    d.n_jobs = (int)p.jobs.size();
    d.n_chunks = (int)p.chunks.size();
    d.ops = upload(p.ops);
    d.jobs = upload(p.jobs);
    if (own_weights) {
        d.chunks = upload(p.chunks);
        if (cudaMalloc(&d.wpack, p.wpack_bytes) != cudaSuccess) return false;
    }
    return d.ops && d.jobs;
}

TcState *tc_create(const NetGeom &g, int64_t max_tiles, int num_sms, int version, std::string &err) {
    TcState *s = new TcState();
    s->g = g;
    s->num_sms = num_sms;
    max_tiles = (max_tiles + 1) & ~(int64_t)1;   // the pair kernel walks 256-sample pair tiles
    s->max_tiles = max_tiles;
    s->version = version;
    if (!tc_build_plan(g, s->plan, err)) { delete s; return nullptr; }
    if (s->plan.np > 4 && version != 2) { err = "hidden > 256 needs the CTA-pair kernel (NERF_MLP_TCGEN05)"; delete s; return nullptr; }
    for (const TcProgram *p : {&s->plan.fwd_train, &s->plan.fwd_infer, &s->plan.bwd}) {
        if (version != 1) break;
        const size_t need = ((s->plan.bias_floats * 4 + 15) & ~15u) + ((p->ops.size() * sizeof(MmaOp) + 15) & ~15u) + p->jobs.size() * sizeof(EpiJob);
        if (need > kTableBytes) { err = "tc_create: program tables exceed the shared-memory table area"; delete s; return nullptr; }
    }
    bool ok = upload_program(s->plan.fwd_train, s->fwd_train, true) && upload_program(s->plan.fwd_infer, s->fwd_infer, false) &&
              upload_program(s->plan.bwd, s->bwd, true);
    s->fwd_infer.wpack = s->fwd_train.wpack;  // identical chunk streams
    s->d_pbias = upload(s->plan.biases);
    s->d_units = upload(s->plan.units);
    ok = ok && s->d_pbias && s->d_units;
    ok = ok && cudaMalloc(&s->d_bias, sizeof(float) * s->plan.bias_floats) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_work, sizeof(WgradWork) * (size_t)num_sms) == cudaSuccess;
    if (max_tiles > 0) {
        ok = ok && cudaMalloc(&s->d_act, (size_t)max_tiles * s->plan.act_slots * kSlotBytes) == cudaSuccess;
        ok = ok && cudaMalloc(&s->d_grad, (size_t)max_tiles * s->plan.grad_slots * kSlotBytes) == cudaSuccess;
        ok = ok && cudaMalloc(&s->d_mask, (size_t)max_tiles * s->plan.mask_slots * NERF_TILE_M * s->plan.mask_words * sizeof(uint32_t)) == cudaSuccess;
    }
    ok = ok && cudaFuncSetAttribute(k_chain<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_chain<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_chain<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmem) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem) == cudaSuccess;
    if (!ok) {
        err = std::string("tc_create: allocation/attribute failure: ") + cudaGetErrorString(cudaGetLastError());
        tc_destroy(s);
        return nullptr;
    }
    if (version == 2) {
        LaneProgram lp;
        const TcProgram *src[3] = {&s->plan.fwd_train, &s->plan.fwd_infer, &s->plan.bwd};
        Lane2Program **dst[3] = {&s->fwd_train2, &s->fwd_infer2, &s->bwd2};
        for (int i = 0; i < 3; ++i) {
            if (!make_lane_program(*src[i], lp, err) || !(*dst[i] = tc2_upload(lp, s->plan.bias_floats, err))) {
                tc_destroy(s);
                return nullptr;
            }
        }
    }
    return s;
}

void tc_destroy(TcState *s) {
    if (!s) return;
    DevProgram *ps[3] = {&s->fwd_train, &s->fwd_infer, &s->bwd};
    for (DevProgram *p : ps) {
        cudaFree(p->ops);
        cudaFree(p->jobs);
        cudaFree(p->chunks);
    }
    cudaFree(s->fwd_train.wpack);
    cudaFree(s->bwd.wpack);
    cudaFree(s->d_pbias);
    cudaFree(s->d_bias);
    cudaFree(s->d_units);
    cudaFree(s->d_work);
    cudaFree(s->d_act);
    cudaFree(s->d_grad);
    cudaFree(s->d_mask);
    tc2_bias_release(s);
    tc2_free(s->fwd_train2);
    tc2_free(s->fwd_infer2);
    tc2_free(s->bwd2);
    delete s;
}

size_t tc_bytes_per_tile(const TcState *s) {
    return (size_t)(s->plan.act_slots + s->plan.grad_slots) * kSlotBytes + (size_t)s->plan.mask_slots * NERF_TILE_M * 4 * s->plan.mask_words;
}

const char *tc_last_error(const TcState *s) { return s->err.c_str(); }

void tc_pack_weights(TcState *s, const float *params, cudaStream_t st) {
    const int nb = (int)s->plan.biases.size();
    launch_pdl(k_pack_all, dim3(s->fwd_train.n_chunks + s->bwd.n_chunks + nb, 4), dim3(256), 0, st, s->fwd_train.chunks, s->fwd_train.n_chunks, s->fwd_train.wpack,
                                                                              s->bwd.chunks, s->bwd.n_chunks, s->bwd.wpack, s->d_pbias, nb,
                                                                              params, s->d_bias);
    ++s->bias_version;
}

int tc_version(const TcState *s) { return s->version; }

int tc_forward(TcState *s, const float *points, const float *dirs, int64_t n, int S, int train, float *sigma, float *rgba,
               cudaStream_t st, const TcRayInputs *fused) {
    const int64_t n_tiles = (n + NERF_TILE_M - 1) / NERF_TILE_M;
    if (n_tiles == 0) return 0;
    if (train && n_tiles > s->max_tiles) { s->err = "tc_forward: batch exceeds the saved-activation capacity"; return -1; }
    const DevProgram &P = train ? s->fwd_train : s->fwd_infer;
    if (s->version == 2) {
        Chain2Launch l;
        memset(&l, 0, sizeof(l));
        l.bwd = false; l.save = train != 0;
        l.wpack = P.wpack; l.bias = s->d_bias; l.bias_floats = (int)s->plan.bias_floats;
        l.bias_slot = tc2_bias_upload(s, s->bias_version, s->d_bias, (int)s->plan.bias_floats, st);
        if (l.bias_slot < 0) { s->err = "tc_forward: bias table exceeds the constant-bank slot"; return -1; }
        l.n_samples = n; l.S = S; l.xyz_freqs = s->g.xyz_freqs; l.dir_freqs = s->g.dir_freqs; l.num_sms = s->num_sms;
        l.points = points; l.dirs = dirs; l.sigma = sigma; l.rgba = rgba;
        if (!points) {
            if (!fused) { s->err = "tc_forward: neither points nor ray inputs"; return -1; }
            l.rays = fused->rays; l.t = fused->t; l.poses = fused->poses;
        } else if (fused && fused->h2d_flag) {
            l.h2d_flag = fused->h2d_flag; l.h2d_chunk_samples = fused->h2d_chunk_samples;
        }
        l.save_base = train ? s->d_act : nullptr; l.save_slots = s->plan.act_slots;
        l.mask_base = s->d_mask; l.mask_slots = s->plan.mask_slots;
        l.wide = s->plan.np > 4; l.e_slot = s->plan.e_slot; l.mask_words = s->plan.mask_words;
        tc2_launch(train ? s->fwd_train2 : s->fwd_infer2, l, st);
        return 0;
    }
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    a.ops = P.ops; a.jobs = P.jobs; a.n_ops = P.n_ops; a.n_jobs = P.n_jobs;
    a.wpack = P.wpack; a.bias = s->d_bias; a.bias_floats = (int)s->plan.bias_floats;
    a.n_samples = n; a.n_tiles = (int)n_tiles; a.S = S;
    a.xyz_freqs = s->g.xyz_freqs; a.dir_freqs = s->g.dir_freqs;
    a.points = points; a.dirs = dirs; a.sigma = sigma; a.rgba = rgba;
    a.save_base = train ? s->d_act : nullptr; a.save_slots = s->plan.act_slots;
    a.mask_base = s->d_mask; a.mask_slots = s->plan.mask_slots;
    const int grid = (int)(n_tiles < s->num_sms ? n_tiles : s->num_sms);
    if (train) k_chain<false, true><<<grid, kChainThreads, kChainSmem, st>>>(a);
    else k_chain<false, false><<<grid, kChainThreads, kChainSmem, st>>>(a);
    return 0;
}

static void build_work(TcState *s, int64_t n_tiles, cudaStream_t st) {
    if (s->work_tiles == n_tiles) return;
    std::vector<WgradWork> work;
    tc_wgrad_partition(s->plan.units, s->num_sms, n_tiles, work);
    cudaMemcpyAsync(s->d_work, work.data(), sizeof(WgradWork) * work.size(), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);  // `work` is a host temporary
    s->work_tiles = n_tiles;
}

int tc_backward(TcState *s, const float *rgba, const float *d_sigma, const float *d_rgba, int64_t n, float *grads,
                cudaStream_t st, void (*between)(void *, const char *), void *user) {
    const int64_t n_tiles = (n + NERF_TILE_M - 1) / NERF_TILE_M;
    if (n_tiles == 0) return 0;
    if (n_tiles > s->max_tiles) { s->err = "tc_backward: batch exceeds the saved-activation capacity"; return -1; }
    if ((int)s->plan.units.size() > kWgMaxSeg * s->num_sms) { s->err = "tc_backward: too few SMs for the weight-gradient units"; return -1; }
    build_work(s, n_tiles, st);
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    a.ops = s->bwd.ops; a.jobs = s->bwd.jobs; a.n_ops = s->bwd.n_ops; a.n_jobs = s->bwd.n_jobs;
    a.wpack = s->bwd.wpack; a.bias = s->d_bias; a.bias_floats = (int)s->plan.bias_floats;
    a.n_samples = n; a.n_tiles = (int)n_tiles; a.S = 1;
    a.rgba = const_cast<float *>(rgba); a.d_sigma = d_sigma; a.d_rgba = d_rgba;
    a.save_base = s->d_grad; a.save_slots = s->plan.grad_slots;
    a.mask_base = s->d_mask; a.mask_slots = s->plan.mask_slots;
    const int grid = (int)(n_tiles < s->num_sms ? n_tiles : s->num_sms);
    if (between) between(user, "mlp_dgrad");
    if (s->version == 2) {
        Chain2Launch l;
        memset(&l, 0, sizeof(l));
        l.bwd = true; l.save = true;
        l.wpack = s->bwd.wpack; l.bias = s->d_bias; l.bias_floats = (int)s->plan.bias_floats;
        l.n_samples = n; l.S = 1; l.xyz_freqs = s->g.xyz_freqs; l.dir_freqs = s->g.dir_freqs; l.num_sms = s->num_sms;
        l.rgba = const_cast<float *>(rgba); l.d_sigma = d_sigma; l.d_rgba = d_rgba;
        l.save_base = s->d_grad; l.save_slots = s->plan.grad_slots;
        l.mask_base = s->d_mask; l.mask_slots = s->plan.mask_slots;
        l.wide = s->plan.np > 4; l.e_slot = s->plan.e_slot; l.mask_words = s->plan.mask_words;
        tc2_launch(s->bwd2, l, st);
    } else
        k_chain<true, true><<<grid, kChainThreads, kChainSmem, st>>>(a);
    if (between) between(user, "mlp_wgrad");
    WgradArgs w;
    w.units = s->d_units; w.work = s->d_work;
    w.act_base = s->d_act; w.grad_base = s->d_grad;
    w.act_slots = s->plan.act_slots; w.grad_slots = s->plan.grad_slots;
    w.grads = grads;
    launch_pdl(k_wgrad, dim3(s->num_sms), dim3(256), kWgSmem, st, w);
    if (between) between(user, nullptr);
    return 0;
}

int tc_debug_read(TcState *s, int area, int64_t tile, int slot, void *out, cudaStream_t st) {
    if (!s || tile < 0 || tile >= s->max_tiles || slot < 0) return -1;
    const void *src = nullptr;
    size_t bytes = kSlotBytes;
    if (area == 0 && slot < s->plan.act_slots) src = s->d_act + ((size_t)tile * s->plan.act_slots + slot) * kSlotBytes;
    else if (area == 1 && slot < s->plan.grad_slots) src = s->d_grad + ((size_t)tile * s->plan.grad_slots + slot) * kSlotBytes;
    else if (area == 2 && slot < s->plan.mask_slots) {
        src = s->d_mask + ((size_t)tile * s->plan.mask_slots + slot) * NERF_TILE_M * s->plan.mask_words;
        bytes = NERF_TILE_M * s->plan.mask_words * sizeof(uint32_t);
    }
    if (!src) return -1;
    if (cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -2;
    if (area == 2 && s->version == 2) {
        // the CTA-pair kernel keeps a tile's masks word-major ([word][row], coalesced warp stores); callers get [row][word]
        const int nw = s->plan.mask_words;
        std::vector<uint32_t> tmp((size_t)NERF_TILE_M * nw);
        memcpy(tmp.data(), out, bytes);
        uint32_t *o = static_cast<uint32_t *>(out);
        for (int r = 0; r < NERF_TILE_M; ++r)
            for (int w = 0; w < nw; ++w) o[(size_t)r * nw + w] = tmp[(size_t)w * NERF_TILE_M + r];
    }
    return 0;
}

// marks of the last k_wgrad launch + the CTA -> (unit, tile range) assignment; out: [num_sms][8] = start, first stage, MMAs
// done, end (ns), unit, tile_begin, tile_end, n_p + n_q
int tc_debug_wgrad_marks(TcState *s, unsigned long long *out, int capacity_ctas, cudaStream_t st) {
    cudaStreamSynchronize(st);
    const int G = s->num_sms < 256 ? s->num_sms : 256;
    if (capacity_ctas < G) return -1;
    std::vector<unsigned long long> marks((size_t)G * 4);
    if (cudaMemcpyFromSymbol(marks.data(), g_wgrad_marks, sizeof(unsigned long long) * 4 * G) != cudaSuccess) return -2;
    std::vector<WgradWork> work((size_t)G);
    if (cudaMemcpy(work.data(), s->d_work, sizeof(WgradWork) * G, cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    for (int i = 0; i < G; ++i) {
        for (int k = 0; k < 4; ++k) out[8 * i + k] = marks[4 * i + k];
        int64_t iters = 0, bytes = 0;
        for (int k = 0; k < work[i].n_seg; ++k) {
            const WgradUnit &u = s->plan.units[work[i].seg[k].unit];
            const int64_t it = 2 * (int64_t)(work[i].seg[k].tile_end - work[i].seg[k].tile_begin);
            iters += it;
            bytes += it * (u.n_p + u.n_q) * 8192;
        }
        out[8 * i + 4] = work[i].n_seg > 0 ? (unsigned long long)work[i].seg[0].unit : 0ull;
        out[8 * i + 5] = (unsigned long long)work[i].n_seg;
        out[8 * i + 6] = (unsigned long long)iters;
        out[8 * i + 7] = (unsigned long long)bytes;
    }
    return G;
}

int tc_debug_trace(TcState *s, const float *points, const float *dirs, int64_t n, int S, int program, const float *rgba,
                   const float *d_sigma, const float *d_rgba, float *sigma_out, float *rgba_out, unsigned long long *host_out,
                   cudaStream_t st) {
    const int64_t n_tiles = (n + NERF_TILE_M - 1) / NERF_TILE_M;
    if (n_tiles == 0 || (program != 1 && n_tiles > s->max_tiles)) return -1;
    unsigned long long *d_trace = nullptr;
    const size_t bytes = sizeof(unsigned long long) * 3 * kTraceEvents * (s->version == 2 ? 4 : 2);
    if (cudaMalloc(&d_trace, bytes) != cudaSuccess) return -2;
    cudaMemsetAsync(d_trace, 0, bytes, st);
    const DevProgram &P = program == 0 ? s->fwd_train : (program == 1 ? s->fwd_infer : s->bwd);
    if (s->version == 2) {
        Chain2Launch l;
        memset(&l, 0, sizeof(l));
        l.bwd = program == 2; l.save = program != 1;
        l.wpack = P.wpack; l.bias = s->d_bias; l.bias_floats = (int)s->plan.bias_floats;
        l.bias_slot = tc2_bias_upload(s, s->bias_version, s->d_bias, (int)s->plan.bias_floats, st);
        if (l.bias_slot < 0) l.bias_slot = 0;
        l.n_samples = n; l.S = S; l.xyz_freqs = s->g.xyz_freqs; l.dir_freqs = s->g.dir_freqs; l.num_sms = s->num_sms;
        l.points = points; l.dirs = dirs; l.sigma = sigma_out; l.rgba = program == 2 ? const_cast<float *>(rgba) : rgba_out;
        l.d_sigma = d_sigma; l.d_rgba = d_rgba;
        l.save_base = program == 0 ? s->d_act : (program == 2 ? s->d_grad : nullptr);
        l.save_slots = program == 2 ? s->plan.grad_slots : s->plan.act_slots;
        l.mask_base = s->d_mask; l.mask_slots = s->plan.mask_slots;
        l.wide = s->plan.np > 4; l.e_slot = s->plan.e_slot; l.mask_words = s->plan.mask_words;
        l.trace = d_trace;
        tc2_launch(program == 0 ? s->fwd_train2 : (program == 1 ? s->fwd_infer2 : s->bwd2), l, st);
        cudaMemcpyAsync(host_out, d_trace, bytes, cudaMemcpyDeviceToHost, st);
        const cudaError_t e2 = cudaStreamSynchronize(st);
        cudaFree(d_trace);
        return e2 == cudaSuccess ? 0 : -2;
    }
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    a.ops = P.ops; a.jobs = P.jobs; a.n_ops = P.n_ops; a.n_jobs = P.n_jobs;
    a.wpack = P.wpack; a.bias = s->d_bias; a.bias_floats = (int)s->plan.bias_floats;
    a.n_samples = n; a.n_tiles = (int)n_tiles; a.S = S;
    a.xyz_freqs = s->g.xyz_freqs; a.dir_freqs = s->g.dir_freqs;
    a.points = points; a.dirs = dirs; a.sigma = sigma_out; a.rgba = program == 2 ? const_cast<float *>(rgba) : rgba_out;
    a.d_sigma = d_sigma; a.d_rgba = d_rgba;
    a.save_base = program == 0 ? s->d_act : (program == 2 ? s->d_grad : nullptr);
    a.save_slots = program == 2 ? s->plan.grad_slots : s->plan.act_slots;
    a.mask_base = s->d_mask; a.mask_slots = s->plan.mask_slots;
    a.trace = d_trace;
    const int grid = (int)(n_tiles < s->num_sms ? n_tiles : s->num_sms);
    if (program == 0) k_chain<false, true><<<grid, kChainThreads, kChainSmem, st>>>(a);
    else if (program == 1) k_chain<false, false><<<grid, kChainThreads, kChainSmem, st>>>(a);
    else k_chain<true, true><<<grid, kChainThreads, kChainSmem, st>>>(a);
    cudaMemcpyAsync(host_out, d_trace, bytes, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d_trace);
    return e == cudaSuccess ? 0 : -2;
}
