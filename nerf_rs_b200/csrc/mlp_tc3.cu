// mlp_tc3.cu -- K-mlp v3: the fused MLP chain on CTA pairs with the hidden activations in TENSOR MEMORY (tcgen05 TS mode).
//
// Why: the SS-mode pair kernel (mlp_tc2.cu) keeps every layer's activations in shared memory, so per
// 2 048-cycle lane GEMM an SM moves 64 KB of A-operand reads + 64 KB of epilogue stores (+ 64 KB of bulk-store reads when
// training) through a 128 B/clk shared memory that also feeds the B operand and the weight ring: it is shared-memory-bound at
// 50-68 % tensor-pipe activity. Here the epilogue writes the bf16 activations into TMEM (tcgen05.st, 16 packed words per
// thread and 32-column group) and the next layer's MMAs read their A operand from there (tcgen05.mma with A in TMEM):
//   * per lane 128 accumulator columns + 128 activation columns (256 bf16 features); two lanes per CTA fill the 512 columns;
//   * a layer wider than 128 runs as two N = 128 half-GEMMs ("steps") into the same accumulator columns. The first half's
//     converted output waits in 16 registers per thread until the second half's MMAs have finished reading the old
//     activations, then both are written in place;
//   * shared memory holds only the encoded inputs (two slot-E panels per lane, SS-mode operands: K <= 64 is too narrow to
//     matter) and the weight ring (16 x 8 KB half chunks per CTA). A tile's slot-E panels are written during the PREVIOUS
//     tile, as soon as their last reader there has completed, after the writing step's own signal: the positional encoding
//     (32 sin/cos per thread, exposed global loads) is off every critical path;
//   * when training, four SAVER warps read each finished layer's activations back out of tensor memory and store them to
//     HBM in the weight-gradient kernel's operand layout (and derive the ReLU masks from them) while the next layer's MMAs
//     run: the epilogue's critical path is the same as in inference.
// Roles per CTA (640 threads, 768 when training): warp 0 weight producer, warps 1 and 3 the two lanes' MMA issuers (leader;
// in the peer warp 1 relays weight arrivals), warp 2 TMEM allocator, warps 4-19 epilogue (4 TMEM lane quarters x 4 column
// slices of 32), warps 20-23 savers (training).
// The per-tile program (TsOp / TsStep, mlp_tc_plan.cpp) travels as kernel parameters.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "chain_common.cuh"
#include "kernels.h"
#include "mlp_tc.h"
#include "ptx.cuh"
#include "raygeom.cuh"

namespace {
using namespace chain;

constexpr int kStages = 16;
constexpr uint32_t kStageBytes = 8192;                               // half chunk: [<= 64 rows][64 bf16]
constexpr uint32_t kSmemRing = 4 * kSlotBytes;                       // after the two lanes' slots E_A, E_B
constexpr uint32_t kSmemBars = kSmemRing + kStages * kStageBytes;    // 192 KB
// barrier ids
constexpr int kBarFull = 0;                   // +(step mod 16): both halves of all of a step's chunks landed (leader: own bytes + the peer's relay)
constexpr int kBarEmpty = kStages;            // +stage: both lanes' MMAs reading the stage completed (multicast commits)
constexpr int kBarAccFull = 2 * kStages;      // +lane
constexpr int kBarEpiDone = 2 * kStages + 2;  // +lane (leader): both CTAs' epilogue warps finished the lane's step
constexpr int kBarActReady = 2 * kStages + 4; // +lane (training): this CTA's epilogue warps have written a layer's activations
constexpr int kBarActSaved = 2 * kStages + 6; // +lane (training): the saver warps have read them out of tensor memory
constexpr int kNumBars = 2 * kStages + 8;
constexpr uint32_t kChain3Smem = kSmemBars + kNumBars * 8 + 16;
static_assert(kStages == 16, "the FULL barriers are indexed by step counter mod 16");
static_assert(kChain3Smem <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
constexpr int kEpiWarps = 16;
constexpr int kServiceWarps = 4;
constexpr int kSaverWarps = 4;                // training only: one per TMEM lane quarter
constexpr int kThreadsInfer = 32 * (kServiceWarps + kEpiWarps);                  // 640 threads, 96 registers per thread at launch
constexpr int kThreadsTrain = 32 * (kServiceWarps + kEpiWarps + kSaverWarps);    // 768 threads, 80 registers per thread at launch
constexpr int kMaxOps = 80, kMaxSteps = 24;
constexpr uint32_t kLaneCols = 256, kActCol = 128;   // TMEM columns per lane; activations start at column 128 of the lane

// padded biases in the constant bank (see mlp_tc2.cu); hidden <= 256: 7*256 + 16 + 256 + 128 + 16 = 2208 floats
constexpr int kBiasSlots = 3;
constexpr int kBiasSlotFloats = 2304;
__constant__ float c_bias3[kBiasSlots * kBiasSlotFloats];

// -DNERF_TC3_STATS: per-CTA cycle counters of the last launch (debug builds only; tools/tc3_stats.py). Slots:
//  0 MMA thread total, 1 waiting for EPI_DONE, 2 waiting for weights (FULL), 3 steps issued,
//  8 epilogue warp 0 total, 9 waiting for ACC_FULL, 10 tcgen05.ld + wait, 11 waiting for SAVE_FREE, 12 (step, lane) items
#ifdef NERF_TC3_STATS
__device__ unsigned long long g_tc3_stats[160 * 16];
__device__ unsigned long long g_tc3_trace[4096];   // CTA 0's MMA thread: (tag << 48 | clock) events, tools/tc3_stats.py --trace
__device__ __forceinline__ void tc3_trace(int &n, unsigned long long tag) {
    if (blockIdx.x == 0 && n < 4096) g_tc3_trace[n++] = (tag << 48) | (clock64() & 0xffffffffffffull);
}
#define TC3_TRACE(n, tag) tc3_trace(n, tag)
#define TC3_STAT_DECL(name) unsigned long long name = 0
#define TC3_CLK() clock64()
#define TC3_STAT_ADD(name, t0) name += clock64() - (t0)
#define TC3_STAT_PUT(slot, v) do { if (lane_id == 0) g_tc3_stats[blockIdx.x * 16 + (slot)] = (v); } while (0)
#else
#define TC3_TRACE(n, tag)
#define TC3_STAT_DECL(name)
#define TC3_CLK() 0ull
#define TC3_STAT_ADD(name, t0)
#define TC3_STAT_PUT(slot, v)
#endif

struct Chain3Args {
    TsOp ops[kMaxOps];
    TsStep steps[kMaxSteps];   // steps[0] = tile prologue
    int32_t n_steps;           // GEMM steps (without the prologue)
    int32_t bias_slot;
    const uint8_t *wpack;
    int64_t n_samples;
    int32_t n_pairs, S;
    int32_t xyz_freqs, dir_freqs;
    const float *points;
    const RayRec *rays;
    const float *t;
    const ViewPose *poses;
    const float *dirs;
    float *sigma;
    float *rgba;
    const float *d_sigma;
    const float *d_rgba;
    uint8_t *save_base;
    int32_t save_slots;
    uint32_t *mask_base;
    int32_t mask_slots;
    const unsigned int *h2d_flag;
    int64_t h2d_chunk_samples;
};

// ---- slot-E producers: the positional encodings (forward) and the sparse gradient panels (backward) of one row, written by the
// warp's column slice h into the lane's slot E (128B-swizzled SS operand) and, when training, into the chunk-major saved image
// (gsave = the row's position in it, or NULL). Deliberately NOT inlined: it runs once or twice per tile, is large (32 sin/cos
// per call site) and register-hungry; inlined at its call sites it tripled the kernel's SASS (I-cache) and kept the epilogue's
// hot loop spilling around it.
template <bool kH2D>
__device__ __noinline__ void write_enc(const Chain3Args &a, uint32_t row, int h, uint8_t kind, uint8_t enc, uint32_t e_addr, int64_t gs,
                                       bool valid, uint8_t *gsave) {
    const int ech0 = 2 * h, ech1 = ech0 + 2;   // this warp's 16-byte chunks of a slot-E row

            if (kind == EK_PROLOGUE_FWD || enc == ENC_X) {
                float v[3] = {0.f, 0.f, 0.f};
                if (valid) {
                    if (a.points) {
                        if (kH2D && a.h2d_flag) {
                            // the host->device copy of the points runs on another stream, chunk by chunk: wait until the chunk
                            // holding this sample has landed (the copy engine writes the counter after the chunk, in stream order)
                            const unsigned int need = (unsigned int)(gs / a.h2d_chunk_samples) + 1u;
                            unsigned int spins = 0, have;
                            do {
                                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(have) : "l"(a.h2d_flag) : "memory");
                                if (have >= need) break;
                                __nanosleep(256);
                            } while (++spins < (1u << 24));
                            if (have < need) __trap();   // the host copy never arrived
                            v[0] = __ldcg(a.points + 3 * gs); v[1] = __ldcg(a.points + 3 * gs + 1); v[2] = __ldcg(a.points + 3 * gs + 2);
                        } else {
                            v[0] = a.points[3 * gs]; v[1] = a.points[3 * gs + 1]; v[2] = a.points[3 * gs + 2];
                        }
                    } else {
                        // fused sampling: the point never exists in HBM -- same ops as k_sample, bit for bit
                        const RayRec rec = a.rays[gs / a.S];
                        raygeom::sample_point(a.poses[rec.view], rec.to, a.t[gs], v);
                    }
                }
                encode_row(h, e_addr, row, v[0], v[1], v[2], a.xyz_freqs, gsave);
            } else if (enc == ENC_D) {
                float v[3] = {0.f, 0.f, 0.f};
                if (valid && ech0 < 4) {
                    const int64_t ray = gs / a.S;
                    v[0] = a.dirs[3 * ray]; v[1] = a.dirs[3 * ray + 1]; v[2] = a.dirs[3 * ray + 2];
                }
                encode_row(h, e_addr, row, v[0], v[1], v[2], a.dir_freqs, gsave);
            } else if (enc == ENC_DSIGMA) {
                const float ds = valid ? a.d_sigma[gs] : 0.f;
                write_sparse_panel(e_addr, row, ech0, ech1, ptx::pack_bf16x2(ds, 0.f), 0u, gsave);
            } else if (kind == EK_PROLOGUE_BWD) {
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f), d = y;
                if (valid && h == 0) {
                    y = reinterpret_cast<const float4 *>(a.rgba)[gs];
                    d = reinterpret_cast<const float4 *>(a.d_rgba)[gs];
                }
                write_sparse_panel(e_addr, row, ech0, ech1, ptx::pack_bf16x2(d.x * y.x * (1.f - y.x), d.y * y.y * (1.f - y.y)),
                                   ptx::pack_bf16x2(d.z * y.z * (1.f - y.z), d.w * y.w * (1.f - y.w)), gsave);
            }
        }

template <bool kBwd, bool kSave, bool kH2D>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kSave ? kThreadsTrain : kThreadsInfer, 1) k_chain3(const __grid_constant__ Chain3Args a) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = ptx::smem_u32(smem);
    const uint32_t bars = sbase + kSmemBars;
    volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kSmemBars + kNumBars * 8);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane_id = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    auto bar = [&](int id) { return bars + 8u * (uint32_t)id; };
    const int bias_base = a.bias_slot * kBiasSlotFloats;
    const float *g_bias = c_bias3 + bias_base;

    if (threadIdx.x == 0) {
        if (sbase & 1023u) __trap();
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(bar(kBarFull + s), rank == 0 ? 2 : 1);
            ptx::mbar_init(bar(kBarEmpty + s), 2);   // both lanes' issuers
        }
        for (int l = 0; l < 2; ++l) {
            ptx::mbar_init(bar(kBarAccFull + l), 1);
            ptx::mbar_init(bar(kBarEpiDone + l), 2 * kEpiWarps);
        }
        for (int l = 0; l < 2; ++l) {
            ptx::mbar_init(bar(kBarActReady + l), kEpiWarps);
            ptx::mbar_init(bar(kBarActSaved + l), kSaverWarps);
        }
        *reinterpret_cast<volatile uint32_t *>(smem + kSmemBars + kNumBars * 8 + 8) = 0u;   // the issuers' lock
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc2<512>(ptx::smem_u32(const_cast<uint32_t *>(tmem_ptr_smem)));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();

    // Register rebalancing (setmaxnreg). Inference: 128 x 64 + 512 x 104 = the CTA's launch allocation of 640 x 96 registers;
    // training (4 more warps): 128 x 40 (service) + 128 x 56 (savers) + 512 x 96 (epilogue) = 768 x 80.
    if (warp < kServiceWarps) {
        if (kSave) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 0 || (warp == 1 && rank != 0)) {
            // ===== warp 0 (both CTAs): weight producer -- this CTA's half of every chunk, once per lane group
            // ===== warp 1 of the peer: relays "my half landed" to the leader's full barrier, in the same order
            const bool producer = warp == 0;
            // One FULL barrier per STEP (slot = step counter mod 16, as many slots as ring stages: a slot cannot come round again
            // before its step's stages have been released, i.e. before both issuers have seen it): the producer arms it with the
            // step's total bytes, every chunk of the step completes on it, and an issuer waits once per step instead of once
            // per chunk (~140 cycles each, even when the chunk landed long ago).
            LaneSched sch(cluster, n_clusters, a.n_pairs, a.n_steps, 0);
            uint32_t stage = 0, phase = 0, gstep = 0;
            int g, nl, pr0, pr1;
            while (sch.next(g, nl, pr0, pr1)) {
                const TsStep st = a.steps[g + 1];
                const uint32_t full = bar(kBarFull + (int)(gstep & 15u));
                if (producer) {
                    uint32_t total = 0;
                    for (int i = st.op_begin; i < st.op_end; ++i) total += (uint32_t)a.ops[i].n * 64u;   // (n / 2) rows * 128 B each
                    for (int i = st.op_begin; i < st.op_end; ++i) {
                        const uint32_t half = (uint32_t)a.ops[i].n * 64u;
                        const uint8_t *src = a.wpack + a.ops[i].w_off + rank * half;
                        ptx::mbar_wait(bar(kBarEmpty + stage), phase ^ 1u);
                        if (ptx::elect_one()) {
                            if (i == st.op_begin) ptx::mbar_arrive_expect_tx(full, total);
                            ptx::bulk_g2s(sbase + kSmemRing + stage * kStageBytes, src, half, full);
                        }
                        __syncwarp();
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                } else {
                    ptx::mbar_wait(full, (gstep >> 4) & 1u);
                    if (lane_id == 0) ptx::mbar_arrive_cluster(ptx::mapa(full, 0));
                    __syncwarp();
                }
                ++gstep;
            }
        } else if (warp == 1 || (warp == 3 && rank == 0)) {
            // ================= leader: one MMA issuer PER LANE (warp 1: lane 0, warp 3: lane 1) =================
            // A single issuing thread is latency-bound, not pipe-bound: per lane step it pays ~200 cycles for the EPI_DONE wait,
            // ~120 per weight-stage wait, ~45 per tcgen05.mma and a few hundred of loop overhead -- ~1 900 cycles for 1 024 cycles
            // of tensor work (tools/tc3_stats.py --trace). With one issuer per lane those latencies overlap with the other lane's
            // MMAs: each issuer has two step times per step. Both wait for the same weight stages; a stage is released when both
            // lanes' MMAs on it have completed (EMPTY counts two commits; a lone lane commits twice).
            // The two issuers must not issue AT THE SAME TIME: their MMAs would interleave in the tensor pipe's queue, both lanes'
            // steps would finish together and the pipe would idle through both epilogues. A lock in shared memory makes an issue
            // burst exclusive; it is released as soon as the burst is ISSUED (~600 cycles of queued MMAs before it completes), so
            // the other lane's burst queues up right behind it. (A strict hand-over token was tried first: it also makes a lane
            // whose epilogue is late hold up the lane that is ready.)
            const int ln = warp == 1 ? 0 : 1;
            LaneSched sch(cluster, n_clusters, a.n_pairs, a.n_steps, 0);
            uint32_t stage = 0, gstep = 0;
            uint32_t done_phase = 0;
            const uint32_t lock = bars + 8u * (uint32_t)kNumBars + 8u;   // (4 bytes after the TMEM base pointer word pair)
            int g, nl, pr0, pr1;
            TC3_STAT_DECL(s_epi); TC3_STAT_DECL(s_full); TC3_STAT_DECL(s_steps); TC3_STAT_DECL(s_issue);
#ifdef NERF_TC3_STATS
            int tr_n = 0;
#endif
            const unsigned long long s_t0 = TC3_CLK();
            const uint32_t d_tmem = tmem_base + (uint32_t)ln * kLaneCols;
            const uint64_t adA = ptx::umma_desc_sw128(sbase + (uint32_t)(2 * ln) * kSlotBytes, 16, 1024);       // the lane's slot E_A
            const uint64_t adB = ptx::umma_desc_sw128(sbase + (uint32_t)(2 * ln + 1) * kSlotBytes, 16, 1024);   // and E_B
            while (sch.next(g, nl, pr0, pr1)) {
                const int ob = a.steps[g + 1].op_begin, n_ops = a.steps[g + 1].op_end - ob;
                if (ln < nl) {
                    TsOp ops[5];   // (tc3_upload checks: at most 5 ops per step)
#pragma unroll
                    for (int i = 0; i < 5; ++i) if (i < n_ops) ops[i] = a.ops[ob + i];
                    if (ln == 0) TC3_TRACE(tr_n, 1);
                    {   // the weights first: this wait overlaps with the lane's epilogue, which is what the issuer really waits for
                        const unsigned long long t0 = TC3_CLK(); (void)t0;
                        ptx::mbar_wait(bar(kBarFull + (int)(gstep & 15u)), (gstep >> 4) & 1u);
                        TC3_STAT_ADD(s_full, t0);
                    }
                    if (ln == 0) TC3_TRACE(tr_n, 3);
                    { const unsigned long long t0 = TC3_CLK(); (void)t0;
                    ptx::mbar_wait(bar(kBarEpiDone + ln), done_phase);
                    TC3_STAT_ADD(s_epi, t0); }
                    done_phase ^= 1u;
                    ptx::tc_fence_after();
                    if (ln == 0) TC3_TRACE(tr_n, 2);
                    const unsigned long long ti0 = TC3_CLK(); (void)ti0;
                    if (ptx::elect_one()) {
                        if (nl == 2) {   // the pipe is mine for this burst
                            uint32_t old;
                            do {
                                asm volatile("atom.shared.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "r"(lock) : "memory");
                            } while (old != 0u);
                        }
                        uint32_t sg = stage;
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            if (i < n_ops) {
                                const TsOp op = ops[i];
                                const uint64_t bd0 = ptx::umma_desc_sw128(sbase + kSmemRing + sg * kStageBytes, 16, 1024);
                                const uint32_t idesc = ptx::umma_idesc_bf16(256, op.n, 0, 0);
                                const uint32_t acc0 = op.first ? 0u : 1u;
                                if (op.a_src >= TS_A_SMEM_B) {
                                    const uint64_t ad0 = op.a_src == TS_A_SMEM ? adA : adB;
                                    ptx::umma_ss2(d_tmem, ad0, bd0, idesc, acc0);
                                    if (op.kcount > 1) ptx::umma_ss2(d_tmem, ad0 + 2u, bd0 + 2u, idesc, 1u);
                                    if (op.kcount > 2) {
                                        ptx::umma_ss2(d_tmem, ad0 + 4u, bd0 + 4u, idesc, 1u);
                                        ptx::umma_ss2(d_tmem, ad0 + 6u, bd0 + 6u, idesc, 1u);
                                    }
                                } else {
                                    // activations: panel p = 64 features = 32 columns; one K16 step = 8 columns
                                    const uint32_t a_tmem = d_tmem + kActCol + 32u * (uint32_t)op.a_src;
                                    ptx::umma_ts2(d_tmem, a_tmem, bd0, idesc, acc0);
                                    ptx::umma_ts2(d_tmem, a_tmem + 8u, bd0 + 2u, idesc, 1u);
                                    ptx::umma_ts2(d_tmem, a_tmem + 16u, bd0 + 4u, idesc, 1u);
                                    ptx::umma_ts2(d_tmem, a_tmem + 24u, bd0 + 6u, idesc, 1u);
                                }
                                ptx::umma_commit2_mc(bar(kBarEmpty + sg), 3);
                                if (nl == 1) ptx::umma_commit2_mc(bar(kBarEmpty + sg), 3);   // (the other lane has no tile left)
                                if (++sg == kStages) sg = 0;
                                if (ln == 0) TC3_TRACE(tr_n, 4);
                            }
                        }
                        ptx::umma_commit2_mc(bar(kBarAccFull + ln), 3);
                        if (nl == 2) asm volatile("st.shared.u32 [%0], %1;" ::"r"(lock), "r"(0u) : "memory");
                        if (ln == 0) TC3_TRACE(tr_n, 5);
                    }
#ifdef NERF_TC3_STATS
                    tr_n = __shfl_sync(0xffffffffu, tr_n, __ffs(__activemask()) - 1);
                    ++s_steps;
#endif
                    __syncwarp();
                    TC3_STAT_ADD(s_issue, ti0);
                }
                stage += (uint32_t)n_ops;
                if (stage >= (uint32_t)kStages) stage -= kStages;
                ++gstep;
            }
#ifdef NERF_TC3_STATS
            if (ln == 0) { TC3_STAT_PUT(0, clock64() - s_t0); TC3_STAT_PUT(1, s_epi); TC3_STAT_PUT(2, s_full); TC3_STAT_PUT(3, s_steps); TC3_STAT_PUT(4, s_issue); }
#endif
        }
    } else if (kSave && warp >= kServiceWarps + kEpiWarps) {
        // ================= saver warps (training): activations TMEM -> HBM, off the epilogue's critical path =================
        // Warp q reads its lane quarter of a finished layer's bf16 activations back from tensor memory (the MMAs of the next
        // layer read them at the same time), stores them in the weight-gradient kernel's operand layout (chunk-major: the
        // warp's 32 rows are 512 contiguous bytes per 16-byte chunk) and, for the forward ReLU layers, derives the ReLU masks
        // from them (mask bit = bf16 activation != 0, three instructions per packed word). The epilogue only waits (ACT_SAVED) before it overwrites the columns,
        // a layer later.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const uint32_t q = (uint32_t)(warp & 3);
        const uint32_t row = q * 32u + (uint32_t)lane_id;
        LaneSched sch(cluster, n_clusters, a.n_pairs, a.n_steps, 0);
        uint32_t rph = 0;   // bit lane: parity to wait for on ACT_READY
        int p, nl, pr0, pr1;
        while (sch.next(p, nl, pr0, pr1)) {
            const TsStep st = a.steps[p + 1];
            if (!st.final_step || st.save_slot < 0) continue;
            const int n_groups = ((int)st.a_col + ((int)st.ncols >> 1)) >> 4;   // 16 TMEM columns = 32 features = half a panel
            const int slot0 = (int)st.save_slot - ((int)st.a_col >> 5);         // the layer's first saved panel
            for (int ln = 0; ln < nl; ++ln) {
                const int tile = 2 * (ln ? pr1 : pr0) + (int)rank;
                ptx::mbar_wait(bar(kBarActReady + ln), (rph >> ln) & 1u);
                rph ^= 1u << ln;
                ptx::tc_fence_after();
                const uint32_t act = tmem_base + ((q * 32u) << 16) + (uint32_t)ln * kLaneCols + kActCol;
                uint8_t *gbase = a.save_base + ((size_t)tile * a.save_slots + (size_t)slot0) * kSlotBytes + cm_row_off(row);
                uint32_t *mbase = (!kBwd && st.mask_slot >= 0) ? a.mask_base + ((size_t)tile * a.mask_slots + st.mask_slot) * NERF_TILE_M * 8 + row : nullptr;
                for (int g2 = 0; g2 < n_groups; g2 += 2) {
                    uint32_t w0[16], w1[16];
                    ptx::tmem_ld16(act + 16u * (uint32_t)g2, w0);
                    if (g2 + 1 < n_groups) ptx::tmem_ld16(act + 16u * (uint32_t)(g2 + 1), w1);
                    ptx::tmem_ld_wait();
                    uint8_t *gp = gbase + (size_t)(g2 >> 1) * kSlotBytes;
#pragma unroll
                    for (int c = 0; c < 4; ++c) st_global_v4(gp + (uint32_t)c * kCmChunkStride, w0[4 * c], w0[4 * c + 1], w0[4 * c + 2], w0[4 * c + 3]);
                    if (g2 + 1 < n_groups) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) st_global_v4(gp + (uint32_t)(4 + c) * kCmChunkStride, w1[4 * c], w1[4 * c + 1], w1[4 * c + 2], w1[4 * c + 3]);
                    }
                    if (mbase) {
                        uint32_t m0 = 0, m1 = 0;   // (bit layout: chain_common.cuh, ts_mask_word)
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            m0 = ts_mask_word(m0, w0[i], i);
                            m1 = ts_mask_word(m1, w1[i], i);
                        }
                        mbase[(size_t)g2 * NERF_TILE_M] = m0;
                        if (g2 + 1 < n_groups) mbase[(size_t)(g2 + 1) * NERF_TILE_M] = m1;
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane_id == 0) ptx::mbar_arrive(bar(kBarActSaved + ln));
            }
        }
    } else {
        if (kSave) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= epilogue warps =================
        // warp (q, h): TMEM lane quarter q (rows 32q..32q+31), 32-column group h of the step's accumulator
        const uint32_t we = (uint32_t)(warp - kServiceWarps);
        const uint32_t q = we & 3u;
        const int h = (int)(we >> 2);
        const uint32_t row = q * 32u + (uint32_t)lane_id;
        uint32_t aph = 0;     // bit l: parity to wait for on ACC_FULL[l]
        uint32_t sph = 0;     // bit lane: ACT_SAVED phase bookkeeping (training)
        const uint32_t done_bar0 = ptx::mapa(bar(kBarEpiDone), 0);   // leader's EPI_DONE[0] in the cluster window
        TC3_STAT_DECL(s_acc); TC3_STAT_DECL(s_ld); TC3_STAT_DECL(s_free); TC3_STAT_DECL(s_items); TC3_STAT_DECL(s_sig); TC3_STAT_DECL(s_conv); TC3_STAT_DECL(s_st); TC3_STAT_DECL(s_h1); TC3_STAT_DECL(s_h2); TC3_STAT_DECL(s_pl); TC3_STAT_DECL(s_pro);
        const unsigned long long s_e0 = TC3_CLK(); (void)s_e0;
        uint32_t e_dirty = 0;   // bit lane: this warp has written a slot-E panel of the lane since its last signal
        auto signal_done = [&](int ln) {
            const unsigned long long t0 = TC3_CLK(); (void)t0;
            if (e_dirty & (1u << ln)) {   // generic-proxy writes -> visible to the pair's MMAs
                ptx::fence_proxy_async_smem();
                e_dirty &= ~(1u << ln);
            }
            ptx::tc_fence_before();          // (tcgen05.ld / tcgen05.st of this warp have been waited for)
            __syncwarp();
            if (lane_id == 0) ptx::mbar_arrive_cluster(done_bar0 + 8u * (uint32_t)ln);
            TC3_STAT_ADD(s_sig, t0);
        };


        // 32 accumulator columns of this warp's quarter -> 16 packed bf16 pairs, in two passes of 16 columns (16 accumulator
        // registers live). `sig`: release the accumulator to the lane's next MMAs as soon as it is in registers.
        auto convert = [&](uint8_t kind, uint32_t tcol, int bidx, uint32_t m, uint32_t (&w)[16], bool sig, int ln) {
            const unsigned long long tc0 = TC3_CLK(); (void)tc0;
            auto half0 = [&](const uint32_t (&r)[16]) {
                if (!kBwd) { if (kind == EK_RELU) epi_half<EK_RELU, 0>(c_bias3, r, bidx, m, w); else epi_half<EK_LINEAR, 0>(c_bias3, r, bidx, m, w); }
                else { if (kind == EK_DMASK) epi_half<EK_DMASK, 0>(c_bias3, r, bidx, m, w); else epi_half<EK_DCOPY, 0>(c_bias3, r, bidx, m, w); }
            };
            auto half1 = [&](const uint32_t (&r)[16]) {
                if (!kBwd) { if (kind == EK_RELU) epi_half<EK_RELU, 1>(c_bias3, r, bidx, m, w + 8); else epi_half<EK_LINEAR, 1>(c_bias3, r, bidx, m, w + 8); }
                else { if (kind == EK_DMASK) epi_half<EK_DMASK, 1>(c_bias3, r, bidx, m, w + 8); else epi_half<EK_DCOPY, 1>(c_bias3, r, bidx, m, w + 8); }
            };
            if constexpr (!kSave) {
                // inference (104 registers): both 16-column loads in flight at once, the accumulator is released before any arithmetic
                uint32_t r0[16], r1[16];
                ptx::tmem_ld16(tcol, r0);
                ptx::tmem_ld16(tcol + 16u, r1);
                ptx::tmem_ld_wait();
                TC3_STAT_ADD(s_ld, tc0);
                if (sig) signal_done(ln);
                half0(r0);
                half1(r1);
            } else {
                // training (96 registers): two passes of 16 columns
                uint32_t r[16];
                ptx::tmem_ld16(tcol, r);
                ptx::tmem_ld_wait();
                TC3_STAT_ADD(s_ld, tc0);
                half0(r);
                ptx::tmem_ld16(tcol + 16u, r);
                ptx::tmem_ld_wait();
                if (sig) signal_done(ln);
                half1(r);
            }
            TC3_STAT_ADD(s_conv, tc0);
        };
        // training: the saver warps must have read the lane's previous activations out of tensor memory before they are
        // overwritten (normally long done: a layer's MMAs take ~4 000 cycles), and are told when the new ones are in place
        auto wait_act_saved = [&](int ln) {
            if (!kSave) return;
            const unsigned long long t0 = TC3_CLK(); (void)t0;
            ptx::mbar_wait(bar(kBarActSaved + ln), ((sph >> ln) & 1u) ^ 1u);
            TC3_STAT_ADD(s_free, t0);
            sph ^= 1u << ln;
        };
        auto act_ready = [&](int ln) {   // (after tcgen05.wait::st)
            if (!kSave) return;
            ptx::tc_fence_before();
            __syncwarp();
            if (lane_id == 0) ptx::mbar_arrive(bar(kBarActReady + ln));
        };
        auto mask_addr = [&](const TsStep &st, int tile) -> uint32_t * {
            return a.mask_base + ((size_t)tile * a.mask_slots + st.mask_slot) * NERF_TILE_M * 8 + (size_t)(st.mask_word0 + h) * NERF_TILE_M + row;
        };
        auto wait_acc = [&](int ln) {
            const unsigned long long t0 = TC3_CLK(); (void)t0;
            ptx::mbar_wait(bar(kBarAccFull + ln), (aph >> ln) & 1u);
            TC3_STAT_ADD(s_acc, t0);
#ifdef NERF_TC3_STATS
            ++s_items;
#endif
            aph ^= 1u << ln;
            ptx::tc_fence_after();
        };

        // A slot-E panel of pair tile `pr` for lane `ln`: panel A = encoded positions (forward) / fc10 pre-activation gradients
        // (backward), panel B = encoded direction / d(sigma). The lanes' first tiles get theirs before step 0; later tiles'
        // panels are written by the steps the program marks (TsStep.pre_enc), after those steps' own signals.
        auto write_panel = [&](int ln, int pr, bool panel_b) {
            const TsStep pj = a.steps[0];
            const int tile = 2 * pr + (int)rank;
            const int64_t gs = (int64_t)tile * NERF_TILE_M + row;
            const int slot = panel_b ? (int)pj.b_save_slot : (int)pj.enc_save_slot;
            uint8_t *gsave = (kSave && slot >= 0) ? a.save_base + ((size_t)tile * a.save_slots + (size_t)slot) * kSlotBytes + cm_row_off(row) : nullptr;
            write_enc<kH2D>(a, row, h, panel_b ? (uint8_t)EK_RELU : pj.kind, panel_b ? pj.b_enc : pj.enc,
                            sbase + (uint32_t)(2 * ln + (panel_b ? 1 : 0)) * kSlotBytes, gs, gs < a.n_samples, gsave);
            e_dirty |= 1u << ln;
        };
        auto pre_encode = [&](const TsStep &st, int ln, int pr, int stride) {
            if (st.pre_enc != TS_PRE_NONE && pr + stride < a.n_pairs) write_panel(ln, pr + stride, st.pre_enc == TS_PRE_B);
        };

        // ---- one half of a two-step layer (N = 128 columns each, every warp active). kSecond = false: convert, release the
        // accumulator, keep the 16 words in `stash`; kSecond = true: convert, then write both halves over the old activations
        // (all of the layer's MMAs have completed once the second half's accumulator is full) and hand them to the next layer.
        auto half_step = [&](auto second, int ln, int p, int pr, int stride, uint32_t (&stash)[16], uint32_t m) {
            constexpr bool kSecond = decltype(second)::value;
            const TsStep st = a.steps[p + 1];
            wait_acc(ln);
            const uint32_t tlane = tmem_base + ((q * 32u) << 16) + (uint32_t)ln * kLaneCols;
            const int bidx = bias_base + (int)st.bias_off + 32 * h;
            if (!kSecond) {
                convert(st.kind, tlane + 32u * (uint32_t)h, bidx, m, stash, true, ln);
                pre_encode(st, ln, pr, stride);
            } else {
                // the accumulator is full = every MMA of the layer has read the old activations: the first half goes in place
                // right away (its registers are free before the second half's accumulator columns are loaded)
                const uint32_t act = tlane + kActCol + 16u * (uint32_t)h;
                if (st.save_slot >= 0) wait_act_saved(ln);   // (the savers have had the layer's whole MMA time: normally long done)
                ptx::tmem_st16(act, stash);
                uint32_t w[16];
                convert(st.kind, tlane + 32u * (uint32_t)h, bidx, m, w, false, ln);
                { const unsigned long long t0 = TC3_CLK(); (void)t0;
                ptx::tmem_st16(act + (uint32_t)st.a_col, w);
                ptx::tmem_st_wait();
                TC3_STAT_ADD(s_st, t0); }
                if (p != a.n_steps - 1 || pr + stride < a.n_pairs) signal_done(ln);   // (nobody waits after the cluster's very last step)
                if (st.save_slot >= 0) act_ready(ln);
                pre_encode(st, ln, pr, stride);
            }
        };

        // ---- any other step: single-step layers (<= 128 columns), the sigma / rgba heads, the slot-E writes and the tile
        // prologue (of the lane's first tile, or of its next tile inside the current tile's last step)
        auto plain_step = [&](int ln, int p, int pr, int stride) {
            const bool has_next = pr + stride < a.n_pairs;
            const int tile = 2 * pr + (int)rank;
            const int64_t gs = (int64_t)tile * NERF_TILE_M + row;
            const bool valid = gs < a.n_samples;
            const bool last_step = (p == a.n_steps - 1);
            const TsStep st = a.steps[p + 1];
            const bool active = 32 * h < (int)st.ncols;
            const bool small = st.kind == EK_SIGMA || st.kind == EK_RGBA;
            // the lane's MMAs may go on as soon as the accumulator is in registers, unless this step also hands over new
            // activations (final step of a layer)
            const bool need_signal = !last_step || has_next;   // (nobody waits after the cluster's very last step)
            const bool early = need_signal && !st.final_step;
            uint32_t m = 0;
            if (kBwd && st.kind == EK_DMASK && active) m = *mask_addr(st, tile);
            wait_acc(ln);
            const uint32_t tlane = tmem_base + ((q * 32u) << 16) + (uint32_t)ln * kLaneCols;
            if (small) {
                uint32_t r[16];
                if (h == 0) {
                    ptx::tmem_ld16(tlane, r);
                    ptx::tmem_ld_wait();
                }
                if (early) signal_done(ln);
                if (h == 0 && valid) {
                    if (st.kind == EK_SIGMA) {
                        a.sigma[gs] = __uint_as_float(r[0]) + g_bias[st.bias_off];
                    } else {
                        const float4 b = *reinterpret_cast<const float4 *>(g_bias + st.bias_off);
                        float4 o;
                        o.x = 1.f / (1.f + expf(-(__uint_as_float(r[0]) + b.x)));
                        o.y = 1.f / (1.f + expf(-(__uint_as_float(r[1]) + b.y)));
                        o.z = 1.f / (1.f + expf(-(__uint_as_float(r[2]) + b.z)));
                        o.w = 1.f / (1.f + expf(-(__uint_as_float(r[3]) + b.w)));
                        reinterpret_cast<float4 *>(a.rgba)[gs] = o;
                    }
                }
            } else {
                // a single-step layer: its MMAs are done, the new activations go in place
                if (st.save_slot >= 0) wait_act_saved(ln);
                if (active) {
                    uint32_t w[16];
                    convert(st.kind, tlane + 32u * (uint32_t)h, bias_base + (int)st.bias_off + 32 * h, m, w, early, ln);
                    ptx::tmem_st16(tlane + kActCol + (uint32_t)st.a_col + 16u * (uint32_t)h, w);
                    ptx::tmem_st_wait();
                } else if (early) {
                    signal_done(ln);
                }
            }
            if (!early && need_signal) signal_done(ln);
            if (!small && st.save_slot >= 0) act_ready(ln);
            pre_encode(st, ln, pr, stride);
        };

        LaneSched sch(cluster, n_clusters, a.n_pairs, a.n_steps, 0);
        int p, nl, pr0, pr1;
        while (sch.next(p, nl, pr0, pr1)) {
            if (p == 0 && pr0 < sch.stride) {   // the lanes' first tiles: their prologues are steps of their own
                const bool has_b = a.steps[0].b_enc != ENC_NONE;
                write_panel(0, pr0, false);
                if (has_b) write_panel(0, pr0, true);
                signal_done(0);
                if (nl == 2) {
                    write_panel(1, pr1, false);
                    if (has_b) write_panel(1, pr1, true);
                    signal_done(1);
                }
            }
            const TsStep st = a.steps[p + 1];
            if (st.ncols == 128 && !st.final_step && st.kind != EK_SIGMA && st.kind != EK_RGBA) {
                // a two-step layer: both lanes' first halves, then both lanes' second halves; the first halves' words stay in
                // registers that are live only inside this block
                if (kBwd && p + 3 < a.n_steps) {
                    // the next layer's mask lines are prefetched towards the SM a layer ahead (the first item of a layer still
                    // loads its word on the spot: then an L2 hit instead of a DRAM access)
                    const TsStep n0 = a.steps[p + 3], n1 = a.steps[p + 4];
                    if (n0.kind == EK_DMASK && n0.mask_slot >= 0) {
                        auto pf = [&](const TsStep &ns, int pr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(mask_addr(ns, 2 * pr + (int)rank))); };
                        pf(n0, pr0);
                        pf(n1, pr0);
                        if (nl == 2) { pf(n0, pr1); pf(n1, pr1); }
                    }
                }
                uint32_t s0[16], s1[16];
                // backward: every item's ReLU-mask word is loaded one item ahead (an item is too short to hide the load behind
                // its own accumulator wait; loading further ahead keeps more words live than the register budget has)
                const bool masked = kBwd && st.kind == EK_DMASK;
                const int t0i = 2 * pr0 + (int)rank, t1i = 2 * pr1 + (int)rank;
                const TsStep stb = a.steps[p + 2];   // the layer's second step
                uint32_t ma = 0, mb = 0;
                if (masked) {
                    ma = *mask_addr(st, t0i);
                    if (nl == 2) mb = *mask_addr(st, t1i);
                }
                { const unsigned long long t0 = TC3_CLK(); (void)t0;
                half_step(std::false_type{}, 0, p, pr0, sch.stride, s0, ma);
                if (masked) ma = *mask_addr(stb, t0i);
                if (nl == 2) {
                    half_step(std::false_type{}, 1, p, pr1, sch.stride, s1, mb);
                    if (masked) mb = *mask_addr(stb, t1i);
                }
                TC3_STAT_ADD(s_h1, t0); }
                sch.next(p, nl, pr0, pr1);   // the layer's second step: same tiles, same lanes
                { const unsigned long long t0 = TC3_CLK(); (void)t0;
                half_step(std::true_type{}, 0, p, pr0, sch.stride, s0, ma);
                if (nl == 2) half_step(std::true_type{}, 1, p, pr1, sch.stride, s1, mb);
                TC3_STAT_ADD(s_h2, t0); }
            } else {
                const unsigned long long t0 = TC3_CLK(); (void)t0;
                plain_step(0, p, pr0, sch.stride);
                if (nl == 2) plain_step(1, p, pr1, sch.stride);
                TC3_STAT_ADD(s_pl, t0);
            }
        }
#ifdef NERF_TC3_STATS
        if (we == 0 && lane_id == 0) { g_tc3_stats[blockIdx.x * 16 + 5] = s_h1; g_tc3_stats[blockIdx.x * 16 + 6] = s_h2; g_tc3_stats[blockIdx.x * 16 + 7] = s_pl; }
#endif
#ifdef NERF_TC3_STATS
        if (we == 0) { TC3_STAT_PUT(8, clock64() - s_e0); TC3_STAT_PUT(9, s_acc); TC3_STAT_PUT(10, s_ld); TC3_STAT_PUT(11, s_free); TC3_STAT_PUT(12, s_items); TC3_STAT_PUT(13, s_sig); TC3_STAT_PUT(14, s_conv); TC3_STAT_PUT(15, s_st); }
#endif
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();   // the peer's smem/TMEM stay alive until every MMA and epilogue of the pair is done
    if (warp == 2) ptx::tmem_dealloc2<512>(tmem_base);
}

}  // namespace

struct Ts3Program {
    TsProgram prog;
};

Ts3Program *tc3_upload(const TsProgram &p, std::string &err) {
    if (p.ops.size() > (size_t)kMaxOps || p.steps.size() > (size_t)kMaxSteps || p.steps.size() < 2) {
        err = "tc3: program exceeds the kernel-parameter tables";
        return nullptr;
    }
    for (size_t i = 1; i < p.steps.size(); ++i)
        if (p.steps[i].op_end - p.steps[i].op_begin > 5) { err = "tc3: at most 5 ops per step"; return nullptr; }
    static bool attr_done = false;
    if (!attr_done) {
        bool ok = cudaFuncSetAttribute(k_chain3<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain3Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain3<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain3Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain3<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain3Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain3<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain3Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain3<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain3Smem) == cudaSuccess;
        if (!ok) { err = std::string("tc3: cudaFuncSetAttribute failed: ") + cudaGetErrorString(cudaGetLastError()); return nullptr; }
        attr_done = true;
    }
    Ts3Program *d = new Ts3Program();
    d->prog = p;
    return d;
}
void tc3_free(Ts3Program *d) { delete d; }

// ---- constant-bank bias slots (per process; see mlp_tc2.cu)
namespace {
struct BiasSlot3 { const void *owner = nullptr; uint64_t version = 0; };
BiasSlot3 g_slots3[kBiasSlots];
int g_next_slot3 = 0;
}  // namespace
int tc3_bias_upload(const void *owner, uint64_t version, const float *d_bias, int n_floats, cudaStream_t st) {
    if (n_floats > kBiasSlotFloats) return -1;
    int slot = -1;
    for (int i = 0; i < kBiasSlots; ++i) if (g_slots3[i].owner == owner) slot = i;
    if (slot < 0) {
        for (int i = 0; i < kBiasSlots && slot < 0; ++i) if (!g_slots3[i].owner) slot = i;
        if (slot < 0) {   // more live engines than slots: evict round-robin; its kernels may still be reading the bank
            slot = g_next_slot3++ % kBiasSlots;
            cudaDeviceSynchronize();
        }
        g_slots3[slot].owner = owner;
        g_slots3[slot].version = ~version;
    }
    if (g_slots3[slot].version != version) {
        cudaMemcpyToSymbolAsync(c_bias3, d_bias, sizeof(float) * n_floats, sizeof(float) * slot * kBiasSlotFloats, cudaMemcpyDeviceToDevice, st);
        g_slots3[slot].version = version;
    }
    return slot;
}
void tc3_bias_release(const void *owner) {
    for (int i = 0; i < kBiasSlots; ++i) if (g_slots3[i].owner == owner) g_slots3[i] = BiasSlot3();
}

// debug: per-CTA cycle counters of the last k_chain3 launch ([ctas][16]); -1 unless built with -DNERF_TC3_STATS
int tc3_debug_trace(unsigned long long *out, int n) {
#ifdef NERF_TC3_STATS
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, g_tc3_trace, sizeof(unsigned long long) * (size_t)(n < 4096 ? n : 4096)) == cudaSuccess ? 0 : -2;
#else
    (void)out; (void)n;
    return -1;
#endif
}
int tc3_debug_stats(unsigned long long *out, int ctas) {
#ifdef NERF_TC3_STATS
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, g_tc3_stats, sizeof(unsigned long long) * 16 * (size_t)(ctas < 160 ? ctas : 160)) == cudaSuccess ? 0 : -2;
#else
    (void)out; (void)ctas;
    return -1;
#endif
}

void tc3_launch(const Ts3Program *P, const Chain2Launch &l, cudaStream_t st) {
    Chain3Args a;
    memset(&a, 0, sizeof(a));
    memcpy(a.ops, P->prog.ops.data(), P->prog.ops.size() * sizeof(TsOp));
    memcpy(a.steps, P->prog.steps.data(), P->prog.steps.size() * sizeof(TsStep));
    a.n_steps = (int)P->prog.steps.size() - 1;
    a.bias_slot = l.bias_slot;
    a.wpack = l.wpack;
    a.n_samples = l.n_samples;
    const int64_t n_tiles = (l.n_samples + NERF_TILE_M - 1) / NERF_TILE_M;
    a.n_pairs = (int)((n_tiles + 1) / 2);
    a.S = l.S;
    a.xyz_freqs = l.xyz_freqs; a.dir_freqs = l.dir_freqs;
    a.points = l.points; a.rays = l.rays; a.t = l.t; a.poses = l.poses; a.dirs = l.dirs; a.sigma = l.sigma; a.rgba = l.rgba;
    a.d_sigma = l.d_sigma; a.d_rgba = l.d_rgba;
    a.save_base = l.save_base; a.save_slots = l.save_slots; a.mask_base = l.mask_base; a.mask_slots = l.mask_slots;
    a.h2d_flag = l.h2d_flag; a.h2d_chunk_samples = l.h2d_chunk_samples;
    const int max_clusters = l.num_sms / 2;
    const int clusters = a.n_pairs < max_clusters ? a.n_pairs : max_clusters;
    const int grid = 2 * clusters;
    if (l.bwd) launch_pdl(k_chain3<true, true, false>, dim3(grid), dim3(kThreadsTrain), kChain3Smem, st, a);
    else if (l.h2d_flag && l.save) launch_pdl(k_chain3<false, true, true>, dim3(grid), dim3(kThreadsTrain), kChain3Smem, st, a);
    else if (l.h2d_flag) launch_pdl(k_chain3<false, false, true>, dim3(grid), dim3(kThreadsInfer), kChain3Smem, st, a);
    else if (l.save) launch_pdl(k_chain3<false, true, false>, dim3(grid), dim3(kThreadsTrain), kChain3Smem, st, a);
    else launch_pdl(k_chain3<false, false, false>, dim3(grid), dim3(kThreadsInfer), kChain3Smem, st, a);
}
