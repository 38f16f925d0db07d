// mlp_tc.cu -- K-mlp: host side of the fused tcgen05/TMEM MLP, the weight-gradient kernel and the weight packer.
//
// The chain kernels (forward fc1..fc10 and the backward dgrad chain) live in mlp_tc3.cu (TS mode: activations in tensor
// memory, hidden <= 256) and mlp_tc2.cu (SS mode: activations in shared memory; hidden 449..512, or any width for A/B runs).
// k_wgrad      dW^T[in x out] = sum over samples of P^T Q with both operands MN-major straight from
//              the saved panel images; fp32 accumulators stay in TMEM across a segment's whole tile
//              range and are flushed with red.global.add.f32. A CTA works through up to three
//              (unit, tile range) segments cut by tc_wgrad_partition (mlp_tc_plan.cpp) so that the
//              fitted cost is equal across the 148 CTAs; the sigma row of fc8 is a CUDA-core GEMV
//              inside the fc8 feature unit, fed by a d(sigma) loader warp.
// k_pack       gathers the flat f32 [out,in] parameter blob into the bf16 chunk streams.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "guard.h"
#include "kernels.h"
#include "mlp_tc.h"
#include "ptx.cuh"

namespace {

constexpr uint32_t kSlotBytes = NERF_PANEL_BYTES;
__device__ __forceinline__ uint32_t panel_chunk_addr(uint32_t slot_addr, uint32_t row, uint32_t chunk) {
    return slot_addr + row * 128u + (((chunk ^ row) & 7u) << 4);
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
    float *partials;   // deterministic mode: per-segment partial blocks (WgradWork::part_off), reduced by k_wgrad_reduce; else NULL
};

// per-CTA wall-clock marks of the last k_wgrad launch (ns, %globaltimer): start, first stage landed, all MMAs done, end
__device__ unsigned long long g_wgrad_marks[256 * 4];
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// kChunkMajor: layout of the saved panels -- [half tile][16-byte chunk][row (64)][16 B] (TS-mode chain kernel, mlp_tc3.cu: a
// no-swizzle MN-major operand) or the 128B-swizzled row-major smem image the CTA-pair SS kernel bulk-stores (mlp_tc2.cu).
template <bool kChunkMajor>
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
                            // Saved half panels are [16-byte chunk (8)][sample row (64)][16 B] (written by the chain kernels'
                            // epilogues): a no-swizzle MN-major operand whose 8 x 16 B core matrices are 128 contiguous bytes,
                            // 128 B apart along K (samples) and 1024 B apart along M/N -- uniformly across consecutive panels
                            // (8 chunks x 1024 B = kWgHalf). One K16 step = 16 sample rows = 256 B.
                            // (swizzled row-major images: 16 sample rows per K step = 2048 B; 64-element M/N blocks are kWgHalf apart)
                            const uint64_t ad = kChunkMajor ? ptx::umma_desc_nosw(st_addr + (uint32_t)(2 * mb) * kWgHalf + k * 256u, 128, 1024)
                                                            : ptx::umma_desc_sw128(st_addr + (uint32_t)(2 * mb) * kWgHalf + k * 2048u, kWgHalf, 1024);
                            const uint64_t bd = kChunkMajor ? ptx::umma_desc_nosw(st_addr + (uint32_t)n_p * kWgHalf + k * 256u, 128, 1024)
                                                            : ptx::umma_desc_sw128(st_addr + (uint32_t)n_p * kWgHalf + k * 2048u, kWgHalf, 1024);
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
                // column 0 = first element of chunk 0: rows 16 B apart (chunk-major) / in the row's swizzled chunk slot (row-major)
                v0 = __ldg(reinterpret_cast<const uint16_t *>(gb + (kChunkMajor ? r0 * 16 : r0 * 128 + ((r0 & 7) << 4))));
                v1 = __ldg(reinterpret_cast<const uint16_t *>(gb + (kChunkMajor ? r1 * 16 : r1 * 128 + ((r1 & 7) << 4))));
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
          if constexpr (kChunkMajor) {
            // Bias gradients (column sums of Q) and, for the fc8 feature unit, the sigma row dW[sigma][m] = sum_s dsigma[s] P[s][m],
            // on the CUDA cores from the stage the MMAs are reading. Half panels are [chunk][row][16 B]: the 32 lanes of a warp
            // read 32 consecutive rows of ONE chunk (512 contiguous bytes, conflict-free); epilogue warp w owns chunks w, w+4, ...
            // (up to 8 of a unit's 32) and keeps their per-lane partial sums in registers until the segment ends.
            const int tid = threadIdx.x - 128;
            const int wq = warp - 4;
            const bool has_bias = u.b_base >= 0;
            const int ccols = N >> 3;             // 16-byte chunk columns of Q
            const bool has_sg = u.sg_slot >= 0 && n_p <= 4;
            const int pcols = n_p * 8;            // 16-byte chunk columns of P
            float bsum[8][8], ssum[8][8];
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
#pragma unroll
                for (int e = 0; e < 8; ++e) bsum[sl][e] = ssum[sl][e] = 0.f;
            float dsum = 0.f;
            auto ld_chunk = [&](uint32_t addr, float (&f)[8]) {
                uint32_t w0, w1, w2, w3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(addr));
                f[0] = __uint_as_float(w0 << 16); f[1] = __uint_as_float(w0 & 0xffff0000u);
                f[2] = __uint_as_float(w1 << 16); f[3] = __uint_as_float(w1 & 0xffff0000u);
                f[4] = __uint_as_float(w2 << 16); f[5] = __uint_as_float(w2 & 0xffff0000u);
                f[6] = __uint_as_float(w3 << 16); f[7] = __uint_as_float(w3 & 0xffff0000u);
            };
            for (int it = 0; it < n_iters; ++it) {
                ptx::mbar_wait(bars + 8 * stage, phase);
                if (seg == 0 && it == 0 && tid == 0 && blockIdx.x < 256) g_wgrad_marks[4 * blockIdx.x + 1] = global_ns();
                const uint32_t st_addr = sbase + stage * kWgStageBytes + (uint32_t)lane * 16u;
                if (has_sg) {
                    const float d0 = s_dsg[stage * 64 + lane], d1 = s_dsg[stage * 64 + 32 + lane];
                    if (wq == 0) dsum += d0 + d1;
#pragma unroll
                    for (int sl = 0; sl < 8; ++sl) {
                        const int ch = wq + 4 * sl;
                        if (ch < pcols) {
                            float f[8];
                            ld_chunk(st_addr + (uint32_t)ch * 1024u, f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) ssum[sl][e] += d0 * f[e];
                            ld_chunk(st_addr + (uint32_t)ch * 1024u + 512u, f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) ssum[sl][e] += d1 * f[e];
                        }
                    }
                }
                if (has_bias) {
                    const uint32_t q_addr = st_addr + (uint32_t)n_p * kWgHalf;
#pragma unroll
                    for (int sl = 0; sl < 8; ++sl) {
                        const int ch = wq + 4 * sl;
                        if (ch < ccols) {
                            float f[8];
                            ld_chunk(q_addr + (uint32_t)ch * 1024u, f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) bsum[sl][e] += f[e];
                            ld_chunk(q_addr + (uint32_t)ch * 1024u + 512u, f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) bsum[sl][e] += f[e];
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(bars + 8 * (kWgStages + stage));
                if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
            }
            // lanes hold partial sums over their rows: butterfly-reduce, lane 0 publishes (every chunk has exactly one owner warp,
            // and the shuffle tree has a fixed order: the bias / sigma-row gradients of a segment are deterministic)
            if (has_bias) {
#pragma unroll
                for (int sl = 0; sl < 8; ++sl) {
                    const int ch = wq + 4 * sl;
                    if (ch < ccols) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float v = bsum[sl][e];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            if (lane == 0) s_bias[ch * 8 + e] = v;
                        }
                    }
                }
            }
            if (has_sg) {
#pragma unroll
                for (int sl = 0; sl < 8; ++sl) {
                    const int ch = wq + 4 * sl;
                    if (ch < pcols) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float v = ssum[sl][e];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            if (lane == 0) s_sg[ch * 8 + e] = v;
                        }
                    }
                }
                if (wq == 0) {
                    float v = dsum;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) s_sg[256] = v;
                }
            }
            ptx::named_bar_sync(1, kEpiThreads);
            if (a.partials) {   // deterministic: the segment's own slots, summed in segment order by k_wgrad_reduce
                float *pb = a.partials + wk.part_off[seg] + (int64_t)(128 * mblocks) * N;
                for (int n = tid; n < 256; n += kEpiThreads) pb[n] = has_bias ? s_bias[n] : 0.f;
                for (int m = tid; m < 260; m += kEpiThreads) pb[256 + m] = has_sg ? s_sg[m] : 0.f;
            } else {
                if (has_bias) {
                    for (int n = tid; n < u.n_valid; n += kEpiThreads) atomicAdd(a.grads + u.b_base + n, s_bias[n]);
                }
                if (has_sg) {
                    for (int m = tid; m < u.m_valid; m += kEpiThreads) atomicAdd(a.grads + u.sg_w_base + m, s_sg[m]);
                    if (tid == 0 && u.sg_b_base >= 0) atomicAdd(a.grads + u.sg_b_base, s_sg[256]);
                }
            }
          } else {
            // (row-major 128B-swizzled panel images, written by the wide-mode chain kernel's bulk stores)
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
            // the row groups add their partial sums one after the other (a fixed order: a segment's bias / sigma-row sums are
            // reproducible; once per segment, so the few extra barriers cost nothing)
            if (has_bias) {
                for (int g = 0; g < n_rg; ++g) {
                    if (rg == g) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) s_bias[c8 * 8 + e] += bsum[e];
                    }
                    ptx::named_bar_sync(1, kEpiThreads);
                }
            }
            if (has_sg) {
                for (int g = 0; g < n_prg; ++g) {
                    if (prg == g) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) s_sg[pc8 * 8 + e] += ssum[e];
                        if (pc8 == 0) s_sg[256] += dsum;
                    }
                    ptx::named_bar_sync(1, kEpiThreads);
                }
            }
            ptx::named_bar_sync(1, kEpiThreads);
            if (a.partials) {   // deterministic: the segment's own slots, summed in segment order by k_wgrad_reduce
                float *pb = a.partials + wk.part_off[seg] + (int64_t)(128 * mblocks) * N;
                for (int n = tid; n < 256; n += kEpiThreads) pb[n] = has_bias ? s_bias[n] : 0.f;
                for (int m = tid; m < 260; m += kEpiThreads) pb[256 + m] = has_sg ? s_sg[m] : 0.f;
            } else {
                if (has_bias) {
                    for (int n = tid; n < u.n_valid; n += kEpiThreads) atomicAdd(a.grads + u.b_base + n, s_bias[n]);
                }
                if (has_sg) {
                    for (int m = tid; m < u.m_valid; m += kEpiThreads) atomicAdd(a.grads + u.sg_w_base + m, s_sg[m]);
                    if (tid == 0 && u.sg_b_base >= 0) atomicAdd(a.grads + u.sg_b_base, s_sg[256]);
                }
            }
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
                if (a.partials) {   // deterministic: plain stores into the segment's own [n][m] block (lanes = m: coalesced)
                    float *pw = a.partials + wk.part_off[seg] + m;
#pragma unroll
                    for (int jn = 0; jn < 32; ++jn) pw[(int64_t)(g * 32 + jn) * (128 * mblocks)] = __uint_as_float(r[jn]);
                } else if (m < u.m_valid) {
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

// Deterministic mode: fold the segments' partial blocks into the gradient blob, each unit's segments in ascending tile order
// (a fixed summation order: bit-identical gradients run to run). One thread per element of a unit's block; the segments'
// values of one element are 2-13 coalesced loads. grid = (blocks, units).
__global__ void k_wgrad_reduce(const WgradRedUnit *units, const int64_t *seg_off, const float *__restrict__ partials, float *grads) {
    pdl_trigger();
    pdl_wait();
    const WgradRedUnit u = units[blockIdx.y];
    const int64_t n_w = (int64_t)u.m_pad * u.n_cols, n_all = n_w + 256 + 260;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_all; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t dst = -1;
        if (i < n_w) {
            const int m = (int)(i % u.m_pad), n = (int)(i / u.m_pad);
            if (m < u.m_valid && n < u.n_valid) dst = u.w_base + (int64_t)n * u.w_row_stride + m;
        } else if (i < n_w + 256) {
            const int n = (int)(i - n_w);
            if (u.b_base >= 0 && n < u.n_valid) dst = u.b_base + n;
        } else if (u.has_sg) {
            const int m = (int)(i - n_w - 256);
            if (m < u.m_valid) dst = u.sg_w_base + m;
            else if (m == 256 && u.sg_b_base >= 0) dst = u.sg_b_base;
        }
        if (dst < 0) continue;
        float sum = 0.f;
        for (int k = u.seg_begin; k < u.seg_end; ++k) sum += partials[seg_off[k] + i];
        grads[dst] += sum;   // (+=: a step that runs in micro-batches accumulates launch after launch, in launch order)
    }
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
    if (guard_malloc(&d, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

struct DevStream {          // packed bf16 weight chunk stream of one program + the gather table that builds it
    PackChunk *chunks = nullptr;
    uint8_t *wpack = nullptr;
    int n_chunks = 0;
};

}  // namespace

struct TcState {
    NetGeom g;
    TcPlan plan;
    int num_sms = 0;
    int64_t max_tiles = 0;
    DevStream fwd, bwd;          // the forward stream serves the training and the inference program (identical chunks)
    PackBias *d_pbias = nullptr;
    float *d_bias = nullptr;
    WgradUnit *d_units = nullptr;
    WgradWork *d_work = nullptr;
    int64_t work_tiles = -1;
    uint8_t *d_act = nullptr, *d_grad = nullptr;
    uint32_t *d_mask = nullptr;
    uint64_t bias_version = 1;   // bumped by tc_pack_weights (constant-bank copy of the biases)
    int version = 3;             // 3: TS-mode CTA-pair chain (mlp_tc3.cu, hidden <= 256); 2: SS-mode CTA-pair chain (mlp_tc2.cu)
    bool chunk_major = false;    // layout of the saved panels (see k_wgrad)
    bool deterministic = false;  // weight gradients through per-segment partial blocks + k_wgrad_reduce (fixed order)
    float *d_partials = nullptr;
    size_t partials_floats = 0;
    WgradRedUnit *d_red_units = nullptr;
    int64_t *d_seg_off = nullptr;
    int n_red_units = 0;
    Lane2Program *fwd_train2 = nullptr, *fwd_infer2 = nullptr, *bwd2 = nullptr;
    Ts3Program *fwd_train3 = nullptr, *fwd_infer3 = nullptr, *bwd3 = nullptr;
    std::string err;
};

static bool upload_stream(const std::vector<PackChunk> &chunks, uint32_t bytes, DevStream &d) {
    d.n_chunks = (int)chunks.size();
    d.chunks = upload(chunks);
    return d.chunks && guard_malloc(&d.wpack, bytes) == cudaSuccess;
}

// version: 0 = best for the geometry (TS mode for hidden <= 256, SS mode above), 2 = SS mode, 3 = TS mode
TcState *tc_create(const NetGeom &g, int64_t max_tiles, int num_sms, int version, std::string &err, bool deterministic) {
    TcState *s = new TcState();
    s->deterministic = deterministic;
    s->g = g;
    s->num_sms = num_sms;
    max_tiles = (max_tiles + 1) & ~(int64_t)1;   // the pair kernels walk 256-sample pair tiles
    s->max_tiles = max_tiles;
    if (!tc_build_plan(g, s->plan, err)) { delete s; return nullptr; }
    if (version == 0) version = s->plan.np <= 4 ? 3 : 2;
    if (version == 3 && s->plan.np > 4) { err = "hidden > 256 needs the SS-mode CTA-pair kernel"; delete s; return nullptr; }
    s->version = version;
    s->chunk_major = version == 3;
    bool ok;
    if (version == 3) {
        ok = upload_stream(s->plan.ts_fwd_train.chunks, s->plan.ts_fwd_train.wpack_bytes, s->fwd) &&
             upload_stream(s->plan.ts_bwd.chunks, s->plan.ts_bwd.wpack_bytes, s->bwd);
    } else {
        ok = upload_stream(s->plan.fwd_train.chunks, s->plan.fwd_train.wpack_bytes, s->fwd) &&
             upload_stream(s->plan.bwd.chunks, s->plan.bwd.wpack_bytes, s->bwd);
    }
    s->d_pbias = upload(s->plan.biases);
    s->d_units = upload(s->plan.units);
    ok = ok && s->d_pbias && s->d_units;
    ok = ok && guard_malloc(&s->d_bias, sizeof(float) * s->plan.bias_floats) == cudaSuccess;
    ok = ok && guard_malloc(&s->d_work, sizeof(WgradWork) * (size_t)num_sms) == cudaSuccess;
    if (max_tiles > 0) {
        ok = ok && guard_malloc(&s->d_act, (size_t)max_tiles * s->plan.act_slots * kSlotBytes) == cudaSuccess;
        ok = ok && guard_malloc(&s->d_grad, (size_t)max_tiles * s->plan.grad_slots * kSlotBytes) == cudaSuccess;
        ok = ok && guard_malloc(&s->d_mask, (size_t)max_tiles * s->plan.mask_slots * NERF_TILE_M * s->plan.mask_words * sizeof(uint32_t)) == cudaSuccess;
    }
    ok = ok && cudaFuncSetAttribute(k_wgrad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_wgrad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem) == cudaSuccess;
    if (!ok) {
        err = std::string("tc_create: allocation/attribute failure: ") + cudaGetErrorString(cudaGetLastError());
        tc_destroy(s);
        return nullptr;
    }
    if (version == 3) {
        const TsProgram *src[3] = {&s->plan.ts_fwd_train, &s->plan.ts_fwd_infer, &s->plan.ts_bwd};
        Ts3Program **dst[3] = {&s->fwd_train3, &s->fwd_infer3, &s->bwd3};
        for (int i = 0; i < 3; ++i)
            if (!(*dst[i] = tc3_upload(*src[i], err))) { tc_destroy(s); return nullptr; }
    } else {
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
    for (DevStream *d : {&s->fwd, &s->bwd}) {
        guard_free(d->chunks);
        guard_free(d->wpack);
    }
    guard_free(s->d_pbias);
    guard_free(s->d_bias);
    guard_free(s->d_units);
    guard_free(s->d_work);
    guard_free(s->d_act);
    guard_free(s->d_grad);
    guard_free(s->d_mask);
    guard_free(s->d_partials);
    guard_free(s->d_red_units);
    guard_free(s->d_seg_off);
    tc2_bias_release(s);
    tc3_bias_release(s);
    tc2_free(s->fwd_train2);
    tc2_free(s->fwd_infer2);
    tc2_free(s->bwd2);
    tc3_free(s->fwd_train3);
    tc3_free(s->fwd_infer3);
    tc3_free(s->bwd3);
    delete s;
}

size_t tc_bytes_per_tile(const TcState *s) {
    return (size_t)(s->plan.act_slots + s->plan.grad_slots) * kSlotBytes + (size_t)s->plan.mask_slots * NERF_TILE_M * 4 * s->plan.mask_words;
}

const char *tc_last_error(const TcState *s) { return s->err.c_str(); }

void tc_pack_weights(TcState *s, const float *params, cudaStream_t st) {
    const int nb = (int)s->plan.biases.size();
    launch_pdl(k_pack_all, dim3(s->fwd.n_chunks + s->bwd.n_chunks + nb, 4), dim3(256), 0, st, s->fwd.chunks, s->fwd.n_chunks, s->fwd.wpack,
               s->bwd.chunks, s->bwd.n_chunks, s->bwd.wpack, s->d_pbias, nb, params, s->d_bias);
    ++s->bias_version;
}

int tc_version(const TcState *s) { return s->version; }

// common launch parameters of a chain program (0 fwd-train, 1 fwd-infer, 2 bwd)
static bool chain_launch(TcState *s, int program, Chain2Launch &l, int64_t n, int S, cudaStream_t st) {
    memset(&l, 0, sizeof(l));
    l.bwd = program == 2;
    l.save = program != 1;
    l.wpack = program == 2 ? s->bwd.wpack : s->fwd.wpack;
    l.bias = s->d_bias;
    l.bias_floats = (int)s->plan.bias_floats;
    l.bias_slot = s->version == 3 ? tc3_bias_upload(s, s->bias_version, s->d_bias, (int)s->plan.bias_floats, st)
                                  : tc2_bias_upload(s, s->bias_version, s->d_bias, (int)s->plan.bias_floats, st);
    if (l.bias_slot < 0) { s->err = "tcgen05 MLP: bias table exceeds the constant-bank slot"; return false; }
    l.n_samples = n; l.S = S; l.xyz_freqs = s->g.xyz_freqs; l.dir_freqs = s->g.dir_freqs; l.num_sms = s->num_sms;
    l.save_base = program == 0 ? s->d_act : (program == 2 ? s->d_grad : nullptr);
    l.save_slots = program == 2 ? s->plan.grad_slots : s->plan.act_slots;
    l.mask_base = s->d_mask; l.mask_slots = s->plan.mask_slots;
    l.wide = s->plan.np > 4; l.e_slot = s->plan.e_slot; l.mask_words = s->plan.mask_words;
    return true;
}
static void chain_run(TcState *s, int program, const Chain2Launch &l, cudaStream_t st) {
    if (s->version == 3) tc3_launch(program == 0 ? s->fwd_train3 : (program == 1 ? s->fwd_infer3 : s->bwd3), l, st);
    else tc2_launch(program == 0 ? s->fwd_train2 : (program == 1 ? s->fwd_infer2 : s->bwd2), l, st);
}

int tc_forward(TcState *s, const float *points, const float *dirs, int64_t n, int S, int train, float *sigma, float *rgba,
               cudaStream_t st, const TcRayInputs *fused) {
    const int64_t n_tiles = (n + NERF_TILE_M - 1) / NERF_TILE_M;
    if (n_tiles == 0) return 0;
    if (train && n_tiles > s->max_tiles) { s->err = "tc_forward: batch exceeds the saved-activation capacity"; return -1; }
    Chain2Launch l;
    if (!chain_launch(s, train ? 0 : 1, l, n, S, st)) return -1;
    l.points = points; l.dirs = dirs; l.sigma = sigma; l.rgba = rgba;
    if (!points) {
        if (!fused) { s->err = "tc_forward: neither points nor ray inputs"; return -1; }
        l.rays = fused->rays; l.t = fused->t; l.poses = fused->poses;
    } else if (fused && fused->h2d_flag) {
        l.h2d_flag = fused->h2d_flag; l.h2d_chunk_samples = fused->h2d_chunk_samples;
    }
    chain_run(s, train ? 0 : 1, l, st);
    return 0;
}

static bool build_work(TcState *s, int64_t n_tiles, cudaStream_t st) {
    if (s->work_tiles == n_tiles) return true;
    std::vector<WgradWork> work;
    tc_wgrad_partition(s->plan.units, s->num_sms, n_tiles, work);
    if (s->deterministic) {
        // every segment gets a private partial block; a unit's blocks are listed in ascending tile order for the reduction
        const int U = (int)s->plan.units.size();
        std::vector<std::vector<std::pair<int, int64_t>>> per_unit((size_t)U);   // (tile_begin, offset)
        int64_t off = 0;
        for (WgradWork &w : work) {
            for (int k = 0; k < w.n_seg; ++k) {
                const WgradUnit &u = s->plan.units[w.seg[k].unit];
                const int64_t m_pad = 128 * ((u.n_p + 1) >> 1), n_cols = 64 * u.n_q;
                w.part_off[k] = off;
                per_unit[(size_t)w.seg[k].unit].push_back({w.seg[k].tile_begin, off});
                off += m_pad * n_cols + 256 + 260;
                off = (off + 31) & ~(int64_t)31;   // 128-byte aligned blocks
            }
        }
        std::vector<WgradRedUnit> red((size_t)U);
        std::vector<int64_t> seg_off;
        for (int i = 0; i < U; ++i) {
            const WgradUnit &u = s->plan.units[i];
            std::sort(per_unit[(size_t)i].begin(), per_unit[(size_t)i].end());
            WgradRedUnit r;
            memset(&r, 0, sizeof(r));
            r.w_base = u.w_base; r.b_base = u.b_base; r.sg_w_base = u.sg_w_base; r.sg_b_base = u.sg_b_base;
            r.w_row_stride = u.w_row_stride; r.m_valid = u.m_valid; r.n_valid = u.n_valid;
            r.m_pad = 128 * ((u.n_p + 1) >> 1); r.n_cols = 64 * u.n_q;
            r.has_sg = (u.sg_slot >= 0 && u.n_p <= 4) ? 1 : 0;
            r.seg_begin = (int)seg_off.size();
            for (auto &pr : per_unit[(size_t)i]) seg_off.push_back(pr.second);
            r.seg_end = (int)seg_off.size();
            red[(size_t)i] = r;
        }
        if ((size_t)off > s->partials_floats) {
            cudaStreamSynchronize(st);
            guard_free(s->d_partials);
            s->d_partials = nullptr;
            if (guard_malloc(&s->d_partials, sizeof(float) * (size_t)off) != cudaSuccess) return false;
            s->partials_floats = (size_t)off;
        }
        guard_free(s->d_red_units);
        guard_free(s->d_seg_off);
        s->d_red_units = upload(red);
        s->d_seg_off = upload(seg_off);
        s->n_red_units = U;
        if (!s->d_red_units || !s->d_seg_off) return false;
    }
    cudaMemcpyAsync(s->d_work, work.data(), sizeof(WgradWork) * work.size(), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);  // `work` is a host temporary
    s->work_tiles = n_tiles;
    return true;
}

int tc_backward(TcState *s, const float *rgba, const float *d_sigma, const float *d_rgba, int64_t n, float *grads,
                cudaStream_t st, void (*between)(void *, const char *), void *user) {
    const int64_t n_tiles = (n + NERF_TILE_M - 1) / NERF_TILE_M;
    if (n_tiles == 0) return 0;
    if (n_tiles > s->max_tiles) { s->err = "tc_backward: batch exceeds the saved-activation capacity"; return -1; }
    if ((int)s->plan.units.size() > kWgMaxSeg * s->num_sms) { s->err = "tc_backward: too few SMs for the weight-gradient units"; return -1; }
    if (!build_work(s, n_tiles, st)) { s->err = "tc_backward: could not allocate the partial-gradient blocks"; return -1; }
    if (between) between(user, "mlp_dgrad");
    Chain2Launch l;
    if (!chain_launch(s, 2, l, n, 1, st)) return -1;
    l.rgba = const_cast<float *>(rgba); l.d_sigma = d_sigma; l.d_rgba = d_rgba;
    chain_run(s, 2, l, st);
    if (between) between(user, "mlp_wgrad");
    WgradArgs w;
    w.units = s->d_units; w.work = s->d_work;
    w.act_base = s->d_act; w.grad_base = s->d_grad;
    w.act_slots = s->plan.act_slots; w.grad_slots = s->plan.grad_slots;
    w.grads = grads;
    w.partials = s->deterministic ? s->d_partials : nullptr;
    if (s->chunk_major) launch_pdl(k_wgrad<true>, dim3(s->num_sms), dim3(256), kWgSmem, st, w);
    else launch_pdl(k_wgrad<false>, dim3(s->num_sms), dim3(256), kWgSmem, st, w);
    if (s->deterministic)
        launch_pdl(k_wgrad_reduce, dim3(64, (unsigned)s->n_red_units), dim3(256), 0, st, (const WgradRedUnit *)s->d_red_units,
                   (const int64_t *)s->d_seg_off, (const float *)s->d_partials, grads);
    if (between) between(user, nullptr);
    return 0;
}

// debug: raw image of one saved panel (area 0 activations, 1 gradients) or one layer's ReLU masks (area 2) of a tile.
// Panel images are chunk-major ([half tile][16-byte chunk][row][8 bf16]) from the TS-mode kernel, 128B-swizzled row-major
// from the SS-mode kernel (tc_version tells which).
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
    if (area == 2) {
        // a tile's masks are kept word-major ([word][row], coalesced warp stores); callers get [row][word]
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

// debug: run one chain program (0 fwd-train, 1 fwd-infer, 2 bwd) of the SS-mode kernel with clock64 tracing of CTA 0
int tc_debug_trace(TcState *s, const float *points, const float *dirs, int64_t n, int S, int program, const float *rgba,
                   const float *d_sigma, const float *d_rgba, float *sigma_out, float *rgba_out, unsigned long long *host_out,
                   cudaStream_t st) {
    const int64_t n_tiles = (n + NERF_TILE_M - 1) / NERF_TILE_M;
    if (n_tiles == 0 || (program != 1 && n_tiles > s->max_tiles) || s->version != 2) return -1;
    unsigned long long *d_trace = nullptr;
    const size_t bytes = sizeof(unsigned long long) * 3 * 2048 * 4;
    if (guard_malloc(&d_trace, bytes) != cudaSuccess) return -2;
    cudaMemsetAsync(d_trace, 0, bytes, st);
    Chain2Launch l;
    if (!chain_launch(s, program, l, n, S, st)) { guard_free(d_trace); return -1; }
    l.points = points; l.dirs = dirs; l.sigma = sigma_out; l.rgba = program == 2 ? const_cast<float *>(rgba) : rgba_out;
    l.d_sigma = d_sigma; l.d_rgba = d_rgba;
    l.trace = d_trace;
    chain_run(s, program, l, st);
    cudaMemcpyAsync(host_out, d_trace, bytes, cudaMemcpyDeviceToHost, st);
    const cudaError_t e2 = cudaStreamSynchronize(st);
    guard_free(d_trace);
    return e2 == cudaSuccess ? 0 : -2;
}
