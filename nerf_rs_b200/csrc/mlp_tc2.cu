// mlp_tc2.cu -- K-mlp v2: the fused MLP chain on CTA PAIRS (tcgen05 cta_group::2), two tiles in flight per CTA.
//
// Why: one 128-sample tile per CTA serialises MMA -> epilogue -> MMA per layer, and every
// tile re-streams the whole 1 MB weight set from L2. Here
//   * a cluster of two CTAs issues M=256 MMAs (tcgen05.mma.cta_group::2): each CTA holds its own 128 rows of
//     activations and HALF of every weight chunk ([N/2 rows][64]), so L2->smem weight traffic per sample halves
//     and the per-CTA weight ring shrinks to 3 x 16 KB;
//   * that frees shared memory for TWO activation sets ("lanes") per CTA. Lane l owns TMEM columns [256 l, 256 l + 256)
//     and five 16 KB panels (4 hidden + 1 encoded-input). The MMA thread and the epilogue warps alternate between the
//     lanes GEMM by GEMM, so lane B's MMAs run while lane A's accumulator is drained, converted and written back as
//     the next layer's A operand (and vice versa): the tensor pipe no longer waits for the epilogue.
// Roles per CTA (384 threads): warp 0 weight producer (cp.async.bulk of this CTA's half chunks), warp 1 MMA issuer in
// the leader CTA / weight-arrival relay in the peer CTA, warp 2 TMEM allocator, warps 4-11 epilogue.
// The per-tile program (ops, GEMM boundaries, epilogue jobs) is the same host-built table the v1 kernel runs
// (mlp_tc_plan.cpp), regrouped per GEMM by make_lane_program().
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "chain_common.cuh"
#include "kernels.h"
#include "mlp_tc.h"
#include "ptx.cuh"
#include "raygeom.cuh"

namespace {
using namespace chain;

constexpr uint32_t kLaneSlots = 5;                                   // 4 hidden panels + E
constexpr uint32_t kLaneBytes = kLaneSlots * kSlotBytes;
constexpr uint32_t kSmemSlots = 2 * kLaneBytes;                      // 160 KB
constexpr int kStages = 4;
constexpr uint32_t kStageBytes = 16384;                              // half chunk: [<=128 rows][64 bf16]
constexpr uint32_t kSmemBars = kSmemSlots + kStages * kStageBytes;   // 224 KB; the program tables travel as kernel parameters
// barrier ids
constexpr int kBarFullL = 0;             // +stage: this CTA's half chunk landed (tx bytes)
constexpr int kBarFullP = kStages;       // +stage: (leader) the peer's half landed
constexpr int kBarEmpty = 2 * kStages;   // +stage: every MMA reading the stage completed (multicast commit)
constexpr int kBarAccFull = 3 * kStages;     // +lane: accumulator complete (multicast commit)
constexpr int kBarEpiDone = 3 * kStages + 2; // +lane: (leader) both CTAs' epilogue warps finished the lane's step
constexpr int kBarSaveReady = 3 * kStages + 4;  // +lane: every epilogue warp has written this step's panels (training)
constexpr int kBarSaveFree = 3 * kStages + 6;   // +lane: the store warp's bulk stores have finished reading them
constexpr int kNumBars = 3 * kStages + 8;
constexpr uint32_t kChain2Smem = kSmemBars + 20 * 8 + 16;
static_assert(kNumBars <= 20, "barrier area");
static_assert(kChain2Smem <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
constexpr int kEpiWarps = 16;                 // 4 lane quarters x kSplit column slices
constexpr int kSplit = kEpiWarps / 4;
constexpr int kThreads = 32 * (4 + kEpiWarps);
constexpr int kMaxOps = 160, kMaxGemms = 16;
constexpr int kMaxGroups = 16 / kSplit;       // 32-column groups per column slice of the widest (512-column) accumulator

// Padded biases live in the constant bank: the epilogue adds them with warp-uniform constant loads, which -- unlike
// ld.shared -- do not queue behind the tensor core's operand reads (tools/ubench_epi.cu: a 128x256 epilogue with
// smem biases takes 1071 cycles on an idle SM but 2486 with MMAs in flight; without the bias loads 697 either way).
// kBiasSlots contexts of one process can hold their biases side by side (tc2_bias_upload arbitrates).
constexpr int kBiasSlots = 3;
constexpr int kBiasSlotFloats = 4608;   // hidden 512: 7*512 + 16 + 512 + 256 + 16 = 4384 padded floats
__constant__ float c_bias[kBiasSlots * kBiasSlotFloats];

struct Chain2Args {
    // the per-tile program lives in the kernel parameter (constant) bank: every field the roles branch on is then
    // warp-uniform by construction (uniform datapath, no shared-memory reads on the critical path)
    LaneOp ops[kMaxOps];
    LaneGemm gemms[kMaxGemms];
    LaneJob jobs[kMaxGemms + 1];   // job 0 = tile prologue
    int32_t n_ops, n_gemms;
    int32_t wide;             // hidden 257..512: ONE lane per CTA (8 hidden panels + slot E = 8), 512-column accumulators
    int32_t e_slot, mask_words;
    const uint8_t *wpack;
    const float *bias;
    int32_t bias_floats, bias_slot;
    int64_t n_samples;
    int32_t n_pairs, S;       // 256-sample pair tiles
    int32_t xyz_freqs, dir_freqs;
    const float *points;      // [n][3] sample positions, or NULL: generate them in the prologue from (rays, t, poses)
    const RayRec *rays;       // [n / S] camera-frame ray directions + view ids (sampler output)
    const float *t;           // [n] sorted depths
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
    unsigned long long *trace;   // debug: [3 roles][kTraceEvents][4] clock64 stamps of CTA 0, or NULL
    const unsigned int *h2d_flag;   // caller-supplied points still arriving from the host (see Chain2Launch), or NULL
    int64_t h2d_chunk_samples;
};
constexpr int kTraceEvents = 2048;
__device__ __forceinline__ void trace4(const Chain2Args &a, int role, int idx, unsigned long long t0, unsigned long long t1,
                                       unsigned long long t2, unsigned long long t3) {
    if (a.trace && blockIdx.x == 0 && idx < kTraceEvents) {
        unsigned long long *p = a.trace + ((size_t)role * kTraceEvents + idx) * 4;
        p[0] = t0; p[1] = t1; p[2] = t2; p[3] = t3;
    }
}

// a GEMM's chunks can be shared by the two lanes when they all fit in the ring at once
__device__ __forceinline__ bool shareable(const LaneGemm &gm) { return (int)(gm.op_end - gm.op_begin) <= kStages; }

__device__ __forceinline__ void store_group(uint32_t lane_base, uint32_t out_slot, uint32_t row, int G, const uint32_t (&w)[16]) {
    const uint32_t slot_addr = lane_base + (out_slot + (uint32_t)(G >> 1)) * kSlotBytes;
    const uint32_t cb = (uint32_t)(G & 1) * 4u;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        st_shared_v4(panel_chunk_addr(slot_addr, row, cb + c), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
}

// One warp's share of a hidden-layer epilogue: 32-column groups [G0, G1) of the accumulator, one at a time
// (with four warps per scheduler the other warps cover the TMEM-load and store latencies).
template <bool kSave, uint8_t kKind, int kG>
__device__ __forceinline__ void epi_hidden(const LaneJob &j, uint32_t taddr, uint32_t lane_base, int bias_base, uint32_t *mask_row,
                                           uint32_t row, int G0, int G1, const uint32_t (&pm)[kG]) {
#pragma unroll
    for (int g = 0; g < kG; ++g) {  // unrolled: G stays in the uniform datapath (2 groups per warp; 4 for a 512-column job)
        const int G = G0 + g;
        if (G < G1) {
            uint32_t r[32];
            uint32_t m = 0;
            if (kKind == EK_DMASK) m = pm[g];   // prefetched before the accumulator wait
            ptx::tmem_ld32(taddr + (uint32_t)G * 32u, r);
            ptx::tmem_ld_wait();
            uint32_t w[16];
            epi_group<kSave, kKind>(c_bias, r, bias_base + (int)j.bias_off + G * 32, m, w);
            store_group(lane_base, j.out_slot, row, G, w);
            if (kSave && kKind == EK_RELU && mask_row) mask_row[G * NERF_TILE_M] = m;   // [word][row]: a warp writes 128 contiguous bytes
        }
    }
}

// kH2D: the caller's points are still arriving from the host (nerf_predict_points) -- a separate instantiation, because even a
// never-taken polling branch in the prologue costs the register-starved epilogue 10-25 % (same-box A/B of the two builds).
// kTrace: clock64 stamps of CTA 0 for tools/trace_chain2.py -- also its own instantiation, for the same reason (the stamps are
// 64-bit values that stay live across a whole epilogue step).
template <bool kBwd, bool kSave, bool kWide, bool kH2D = false, bool kTrace = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_chain2(const __grid_constant__ Chain2Args a) {
    constexpr int kG = kWide ? kMaxGroups : kMaxGroups / 2;   // 32-column groups per column-slice warp (256- vs 512-column jobs)
    pdl_trigger();
    const bool tracing = kTrace && a.trace != nullptr;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = ptx::smem_u32(smem);
    const uint32_t bars = sbase + kSmemBars;
    volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(smem + kSmemBars + 20 * 8);
    // the shuffle tells the compiler the warp index is warp-uniform (role dispatch and all table indices stay uniform)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane_id = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    auto bar = [&](int id) { return bars + 8u * (uint32_t)id; };

    const int bias_base = a.bias_slot * kBiasSlotFloats;
    const float *g_bias = c_bias + bias_base;
    const LaneOp *s_ops = a.ops;
    const LaneGemm *s_gemms = a.gemms;
    const LaneJob *s_jobs = a.jobs;

    if (threadIdx.x == 0) {
        if (sbase & 1023u) {
            printf("nerf_b200: dynamic smem base not 1024-aligned\n");
            __trap();
        }
        for (int s = 0; s < kStages; ++s) {
            // leader: own half (expect_tx arrive) + the peer's relay; peer: own half only
            ptx::mbar_init(bar(kBarFullL + s), rank == 0 ? 2 : 1);
            ptx::mbar_init(bar(kBarFullP + s), 1);
            ptx::mbar_init(bar(kBarEmpty + s), 1);
        }
        for (int l = 0; l < 2; ++l) {
            ptx::mbar_init(bar(kBarAccFull + l), 1);
            ptx::mbar_init(bar(kBarEpiDone + l), 2 * kEpiWarps);
            ptx::mbar_init(bar(kBarSaveReady + l), kEpiWarps);
            ptx::mbar_init(bar(kBarSaveFree + l), 1);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc2<512>(ptx::smem_u32(const_cast<uint32_t *>(tmem_ptr_smem)));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();   // barriers of both CTAs initialised before any remote arrive / multicast commit
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();   // everything above is private to the CTA pair; the predecessor kernel's outputs are read only below

    // Register rebalancing (setmaxnreg): the four service warps (producer, MMA issuer / relay, TMEM allocator, store warp)
    // need few registers; the 16 epilogue warps hold 32 accumulator columns, 16 packed words and addressing.
    // The pool is what the CTA was launched with (96 x 640 registers): 128 x 40 + 512 x 104 fits, 112 would block forever.
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 || (warp == 1 && rank != 0)) {
        // ===== warp 0 (both CTAs): weight producer -- this CTA's half of every chunk, once per lane GROUP when the
        //       GEMM is shareable, once per lane otherwise
        // ===== warp 1 of the peer: relays "my half landed" to the leader's MMA thread, in the same order
        const bool producer = warp == 0;
        LaneSched sch(cluster, n_clusters, a.n_pairs, a.n_gemms, kWide ? 1 : 0);
        uint32_t stage = 0, phase = 0;
        int g, nl, pr0, pr1, ev = 0;
        while (sch.next(g, nl, pr0, pr1)) {
            const LaneGemm gm = s_gemms[g];
            const int reps = (nl == 2 && !shareable(gm)) ? 2 : 1;
            for (int rep = 0; rep < reps; ++rep) {
                for (int i = gm.op_begin; i < gm.op_end; ++i) {
                    if (producer) {
                        const uint32_t half = (uint32_t)s_ops[i].n * 64u;   // (n / 2) rows * 128 B
                        const uint8_t *src = a.wpack + s_ops[i].w_off + rank * half;
                        const unsigned long long tp0 = tracing ? clock64() : 0;
                        ptx::mbar_wait(bar(kBarEmpty + stage), phase ^ 1u);
                        if (tracing && lane_id == 0) trace4(a, 2, ev++, tp0, clock64(), 0, 0);
                        if (ptx::elect_one()) {
                            ptx::mbar_arrive_expect_tx(bar(kBarFullL + stage), half);
                            ptx::bulk_g2s(sbase + kSmemSlots + stage * kStageBytes, src, half, bar(kBarFullL + stage));
                        }
                    } else {
                        ptx::mbar_wait(bar(kBarFullL + stage), phase);
                        if (lane_id == 0) ptx::mbar_arrive_cluster(ptx::mapa(bar(kBarFullL + stage), 0));
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= leader: MMA issuer for the pair =================
        LaneSched sch(cluster, n_clusters, a.n_pairs, a.n_gemms, kWide ? 1 : 0);
        uint32_t stage = 0, phase = 0;   // next stage to be consumed for the first time
        uint32_t done_phase = 0;         // bit l: parity to wait for on EPI_DONE[l]
        int mma_ev = 0;
        int g, nl, pr0, pr1;
        while (sch.next(g, nl, pr0, pr1)) {
            const LaneGemm gm = s_gemms[g];
            const bool shared = nl == 2 && shareable(gm);
            const uint32_t stage0 = stage, phase0 = phase;
            for (int ln = 0; ln < nl; ++ln) {
                // lane 1 of a shared group re-reads the stages lane 0 just used (already landed) and releases them
                const bool reuse = shared && ln == 1;
                const bool release = !shared || ln == 1;
                if (reuse) { stage = stage0; phase = phase0; }
                unsigned long long tm0 = tracing ? clock64() : 0;
                ptx::mbar_wait_cluster(bar(kBarEpiDone + ln), (done_phase >> ln) & 1u);
                done_phase ^= 1u << ln;
                for (int i = gm.op_begin; i < gm.op_end; ++i) {
                    const LaneOp op = s_ops[i];
                    if (i != gm.op_begin && tracing) tm0 = clock64();
                    if (!reuse) {
                        ptx::mbar_wait(bar(kBarFullL + stage), phase);   // both halves: own bytes + the peer's relay
                    }
                    ptx::tc_fence_after();
                    const unsigned long long tm1 = tracing ? clock64() : 0;
                    const uint32_t a_addr = sbase + (uint32_t)ln * kLaneBytes + (uint32_t)op.a_slot * kSlotBytes;
                    const uint32_t b_addr = sbase + kSmemSlots + stage * kStageBytes;
                    const uint32_t idesc = ptx::umma_idesc_bf16(256, op.n, 0, 0);
                    const uint32_t d_tmem = tmem_base + (uint32_t)ln * 256u + ((op.kflags & LANE_OP_HALF1) ? 256u : 0u);
                    const uint64_t ad0 = ptx::umma_desc_sw128(a_addr, 16, 1024);
                    const uint64_t bd0 = ptx::umma_desc_sw128(b_addr, 16, 1024);
                    const uint32_t acc0 = (op.kflags & LANE_OP_FIRST) ? 0u : 1u;
                    const uint32_t kcount = op.kflags & LANE_OP_KCOUNT;
                    if (ptx::elect_one()) {
                        // +32 B per K16 step == +2 in the descriptor's address field
                        ptx::umma_ss2(d_tmem, ad0, bd0, idesc, acc0);
                        if (kcount > 1) ptx::umma_ss2(d_tmem, ad0 + 2u, bd0 + 2u, idesc, 1u);
                        if (kcount > 2) {
                            ptx::umma_ss2(d_tmem, ad0 + 4u, bd0 + 4u, idesc, 1u);
                            ptx::umma_ss2(d_tmem, ad0 + 6u, bd0 + 6u, idesc, 1u);
                        }
                        if (release) ptx::umma_commit2_mc(bar(kBarEmpty + stage), 3);
                        if (i + 1 == gm.op_end) ptx::umma_commit2_mc(bar(kBarAccFull + ln), 3);
                    }
                    __syncwarp();
                    if (tracing && lane_id == 0) trace4(a, 0, mma_ev++, tm0, tm1, tm1, clock64());
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (kSave && warp == 3) {
        // ================= store warp (training): bulk-stores every step's panels to the per-tile save area ==========
        // Steps arrive in the epilogue's order. A lane's panels may be rewritten once its stores have finished READING
        // shared memory (SAVE_FREE); the global writes themselves only have to land before the kernel ends.
        LaneSched sch(cluster, n_clusters, a.n_pairs, a.n_gemms, kWide ? 1 : 0);
        uint32_t rph = 0;   // bit l: parity to wait for on SAVE_READY[l]
        const LaneJob pj = s_jobs[0];
        int p, nl, pr0, pr1;
        auto store_step = [&](int ln, auto issue) {
            ptx::mbar_wait(bar(kBarSaveReady + ln), (rph >> ln) & 1u);
            rph ^= 1u << ln;
            if (lane_id == 0) {
                issue();
                ptx::bulk_commit();
                ptx::bulk_wait_read<0>();
                ptx::mbar_arrive(bar(kBarSaveFree + ln));
            }
            __syncwarp();
        };
        while (sch.next(p, nl, pr0, pr1)) {
            for (int ln = 0; ln < nl; ++ln) {
                const int pr = ln ? pr1 : pr0;
                const bool first_tile = pr < sch.stride;
                const bool has_next = pr + sch.stride < a.n_pairs;
                const uint32_t lane_base = sbase + (uint32_t)ln * kLaneBytes;
                const uint32_t e_addr = lane_base + (uint32_t)a.e_slot * kSlotBytes;
                uint8_t *tile_base = a.save_base + (size_t)(2 * pr + (int)rank) * a.save_slots * kSlotBytes;
                uint8_t *next_base = a.save_base + (size_t)(2 * (pr + sch.stride) + (int)rank) * a.save_slots * kSlotBytes;
                if (p == 0 && first_tile)
                    store_step(ln, [&]() {
                        if (pj.enc_save_slot >= 0) ptx::bulk_s2g(tile_base + (size_t)pj.enc_save_slot * kSlotBytes, e_addr, kSlotBytes);
                    });
                const LaneJob j = s_jobs[p + 1];
                store_step(ln, [&]() {
                    // consecutive panels are contiguous both in shared memory and in the save area: one copy
                    if (j.save_slot >= 0) {
                        for (int p4 = 0; p4 < (j.ncols >> 6); p4 += 4) {   // <= 64 KB per copy
                            const int np4 = (j.ncols >> 6) - p4 < 4 ? (j.ncols >> 6) - p4 : 4;
                            ptx::bulk_s2g(tile_base + (size_t)(j.save_slot + p4) * kSlotBytes,
                                          lane_base + (uint32_t)(j.out_slot + p4) * kSlotBytes, (uint32_t)np4 * kSlotBytes);
                        }
                    }
                    if (j.enc_save_slot >= 0) ptx::bulk_s2g(tile_base + (size_t)j.enc_save_slot * kSlotBytes, e_addr, kSlotBytes);
                    if (p == a.n_gemms - 1 && has_next && pj.enc_save_slot >= 0)
                        ptx::bulk_s2g(next_base + (size_t)pj.enc_save_slot * kSlotBytes, e_addr, kSlotBytes);
                });
            }
        }
        if (lane_id == 0) ptx::bulk_wait_all<0>();
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= epilogue warps =================
        // warp (q, h): TMEM lane quarter q (rows 32q..32q+31), column slice h of kSplit of the accumulator
        const uint32_t we = (uint32_t)(warp - 4);
        const uint32_t q = we & 3u;
        const int h = (int)(we >> 2);
        const uint32_t row = q * 32u + (uint32_t)lane_id;
        const int ech0 = h * (8 / kSplit), ech1 = ech0 + 8 / kSplit;   // this warp's 16-byte chunks of a slot-E row
        uint32_t aph = 0;   // bit l: parity to wait for on ACC_FULL[l]
        uint32_t sph = 0;   // bit l: SAVE_FREE[l] phase bookkeeping (training)
        const uint32_t done_bar0 = ptx::mapa(bar(kBarEpiDone), 0);   // leader's EPI_DONE[0] in the cluster window

        // ---- slot-E producers (each column half writes four of the eight 16-byte chunks of a row)
        auto write_enc = [&](uint8_t kind, uint8_t enc, uint32_t e_addr, int64_t gs, bool valid) {
            if (kind == EK_PROLOGUE_FWD || enc == ENC_X) {
                float v[3] = {0.f, 0.f, 0.f};
                if (valid) {
                    if (a.points) {
                        if (kH2D && a.h2d_flag) {
                            // the host->device copy of the points runs on another stream, chunk by chunk: wait until the chunk
                            // holding this sample has landed (the copy engine writes the counter after the chunk, in stream order)
                            const unsigned int need = (unsigned int)(gs / a.h2d_chunk_samples) + 1u;
                            unsigned int spins = 0, have;
                            do {   // acquire: the point loads below may not be satisfied before the counter is seen
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
                encode_row(h, e_addr, row, v[0], v[1], v[2], a.xyz_freqs);
            } else if (enc == ENC_D) {
                float v[3] = {0.f, 0.f, 0.f};   // 3 + 6*4 = 27 values; features beyond 3 + 6*freqs are written as zeros
                if (valid && ech0 < 4) {
                    const int64_t ray = gs / a.S;
                    v[0] = a.dirs[3 * ray]; v[1] = a.dirs[3 * ray + 1]; v[2] = a.dirs[3 * ray + 2];
                }
                encode_row(h, e_addr, row, v[0], v[1], v[2], a.dir_freqs);
            } else if (enc == ENC_DSIGMA) {
                const float ds = valid ? a.d_sigma[gs] : 0.f;
                write_sparse_panel(e_addr, row, ech0, ech1, ptx::pack_bf16x2(ds, 0.f), 0u);
            } else if (kind == EK_PROLOGUE_BWD) {
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f), d = y;
                if (valid && h == 0) {
                    y = reinterpret_cast<const float4 *>(a.rgba)[gs];
                    d = reinterpret_cast<const float4 *>(a.d_rgba)[gs];
                }
                write_sparse_panel(e_addr, row, ech0, ech1, ptx::pack_bf16x2(d.x * y.x * (1.f - y.x), d.y * y.y * (1.f - y.y)),
                                   ptx::pack_bf16x2(d.z * y.z * (1.f - y.z), d.w * y.w * (1.f - y.w)));
            }
        };
        unsigned long long ts_a = 0, ts_b = 0;
        auto signal_done = [&](int ln) {
            ptx::fence_proxy_async_smem();   // generic-proxy panel writes -> visible to the pair's MMAs / bulk stores
            if (tracing) ts_a = clock64();
            ptx::tc_fence_before();
            __syncwarp();
            if (tracing) ts_b = clock64();
            if (lane_id == 0) ptx::mbar_arrive_cluster(done_bar0 + 8u * (uint32_t)ln);
        };
        LaneSched sch(cluster, n_clusters, a.n_pairs, a.n_gemms, kWide ? 1 : 0);
        int epi_ev = 0;
        int p, nl, pr0, pr1;
        while (sch.next(p, nl, pr0, pr1)) {
          for (int ln = 0; ln < nl; ++ln) {
            const int pr = ln ? pr1 : pr0;
            const bool first_tile = pr < sch.stride;
            const bool has_next = pr + sch.stride < a.n_pairs;
            const uint32_t lane_base = sbase + (uint32_t)ln * kLaneBytes;
            const uint32_t e_addr = lane_base + (uint32_t)a.e_slot * kSlotBytes;
            const int tile = 2 * pr + (int)rank;
            const int64_t gs = (int64_t)tile * NERF_TILE_M + row;
            const bool valid = gs < a.n_samples;
            const int next_tile = 2 * (pr + sch.stride) + (int)rank;
            const int64_t ngs = (int64_t)next_tile * NERF_TILE_M + row;
            const bool last_step = (p == a.n_gemms - 1);
            const unsigned long long te0 = tracing ? clock64() : 0;

            // ---- tile prologue: before job 1 of a lane's first tile (signals), or ahead of time for the NEXT tile
            //      inside the last step of the current one (slot E is free by then; the last job carries the signal)
            const LaneJob pj = s_jobs[0];
            if (p == 0 && first_tile) {
                if (kSave) { ptx::mbar_wait(bar(kBarSaveFree + ln), ((sph >> ln) & 1u) ^ 1u); sph ^= 1u << ln; }
                write_enc(pj.kind, pj.enc, e_addr, gs, valid);
                signal_done(ln);
                if (kSave && lane_id == 0) ptx::mbar_arrive(bar(kBarSaveReady + ln));   // store warp: the prologue panel, as its own step
            }
            // the panels / slot E this step rewrites may still be the source of this lane's previous bulk stores
            if (kSave) { ptx::mbar_wait(bar(kBarSaveFree + ln), ((sph >> ln) & 1u) ^ 1u); sph ^= 1u << ln; }
            const bool pre_next = last_step && has_next;
            if (pre_next) write_enc(pj.kind, pj.enc, e_addr, ngs, ngs < a.n_samples);

            // ---- the GEMM's epilogue job
            const LaneJob j = s_jobs[p + 1];
            const int NG = j.ncols >> 5;                    // 32-column groups of this accumulator
            const int gpw = (NG + kSplit - 1) / kSplit;     // groups per column slice
            const int G0 = h * gpw < NG ? h * gpw : NG;
            const int G1 = G0 + gpw < NG ? G0 + gpw : NG;
            uint32_t *mask_row = nullptr;
            uint32_t pm[kG];
            if (j.mask_slot >= 0 && j.kind != EK_SIGMA && j.kind != EK_RGBA) {
                mask_row = a.mask_base + ((size_t)tile * a.mask_slots + j.mask_slot) * NERF_TILE_M * a.mask_words + row;   // word-major: [G][row]
                if (kBwd && j.kind == EK_DMASK) {   // global loads issued now, consumed after the accumulator wait
#pragma unroll
                    for (int g = 0; g < kG; ++g) pm[g] = (G0 + g < G1) ? mask_row[(G0 + g) * NERF_TILE_M] : 0u;
                }
            }
            const unsigned long long te1 = tracing ? clock64() : 0;
            ptx::mbar_wait(bar(kBarAccFull + ln), (aph >> ln) & 1u);
            aph ^= 1u << ln;
            ptx::tc_fence_after();
            const unsigned long long te2 = tracing ? clock64() : 0;
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + (uint32_t)ln * 256u;
            if (j.kind == EK_RELU || j.kind == EK_LINEAR || j.kind == EK_DMASK || j.kind == EK_DCOPY) {
                if (!kBwd) {
                    if (j.kind == EK_RELU) epi_hidden<kSave, EK_RELU, kG>(j, taddr, lane_base, bias_base, mask_row, row, G0, G1, pm);
                    else epi_hidden<kSave, EK_LINEAR, kG>(j, taddr, lane_base, bias_base, mask_row, row, G0, G1, pm);
                } else {
                    if (j.kind == EK_DMASK) epi_hidden<kSave, EK_DMASK, kG>(j, taddr, lane_base, bias_base, mask_row, row, G0, G1, pm);
                    else epi_hidden<kSave, EK_DCOPY, kG>(j, taddr, lane_base, bias_base, mask_row, row, G0, G1, pm);
                }
            } else if (j.kind == EK_SIGMA || j.kind == EK_RGBA) {
                if (h == 0) {
                    uint32_t r[16];
                    ptx::tmem_ld16(taddr, r);
                    ptx::tmem_ld_wait();
                    if (j.kind == EK_SIGMA) {
                        if (valid) a.sigma[gs] = __uint_as_float(r[0]) + g_bias[j.bias_off];
                    } else if (valid) {
                        const float4 b = *reinterpret_cast<const float4 *>(g_bias + j.bias_off);
                        float4 o;
                        o.x = 1.f / (1.f + expf(-(__uint_as_float(r[0]) + b.x)));
                        o.y = 1.f / (1.f + expf(-(__uint_as_float(r[1]) + b.y)));
                        o.z = 1.f / (1.f + expf(-(__uint_as_float(r[2]) + b.z)));
                        o.w = 1.f / (1.f + expf(-(__uint_as_float(r[3]) + b.w)));
                        reinterpret_cast<float4 *>(a.rgba)[gs] = o;
                    }
                }
            }
            const unsigned long long te3 = tracing ? clock64() : 0;
            if (j.enc != ENC_NONE) write_enc(j.kind, j.enc, e_addr, gs, valid);
            if (!last_step || has_next) signal_done(ln);
            const unsigned long long te4 = tracing ? clock64() : 0;   // the lane's next GEMM may start (accumulator drained, panels written)

            if (kSave) {
                if (last_step && !has_next) ptx::fence_proxy_async_smem();   // (signal_done, which fences, was skipped)
                __syncwarp();
                if (lane_id == 0) ptx::mbar_arrive(bar(kBarSaveReady + ln));     // store warp: this step's panels
            }
            if (tracing && we == 0 && lane_id == 0) {
                trace4(a, 1, 2 * epi_ev, te0, te1, te2, clock64());
                trace4(a, 1, 2 * epi_ev + 1, te3, te4, ts_a, ts_b);
            }
            ++epi_ev;
          }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();   // the peer's smem/TMEM stay alive until every MMA and epilogue of the pair is done
    if (warp == 2) ptx::tmem_dealloc2<512>(tmem_base);
}

}  // namespace

struct Lane2Program {   // host copy of the tables; they are passed to the kernel by value
    LaneProgram prog;
};

bool make_lane_program(const TcProgram &p, LaneProgram &out, std::string &err) {
    out = LaneProgram();
    for (size_t i = 0; i < p.ops.size(); ++i) {
        const MmaOp &o = p.ops[i];
        if (o.flags & TC_OP_GEMM_START) {
            if (!out.gemms.empty()) out.gemms.back().op_end = (uint16_t)i;
            out.gemms.push_back(LaneGemm{(uint16_t)i, (uint16_t)i});
        }
        if (o.n % 16) { err = "lane program: MMA N must be a multiple of 16 for cta_group::2"; return false; }
        const bool half1 = !(o.flags & TC_OP_GEMM_START) && o.acc == 1 && p.wide;
        out.ops.push_back(LaneOp{o.w_off, o.n, o.a_slot,
                                 (uint8_t)(o.kcount | ((o.flags & TC_OP_FIRST) ? LANE_OP_FIRST : 0u) | ((p.wide && o.acc == 1) ? LANE_OP_HALF1 : 0u))});
        (void)half1;
    }
    if (out.gemms.empty()) { err = "lane program: no GEMMs"; return false; }
    out.gemms.back().op_end = (uint16_t)p.ops.size();
    // jobs: job 0 = prologue; then one merged job per GEMM (v1 splits hidden epilogues in two 128-column blocks)
    size_t ji = 0;
    if (p.jobs.empty() || p.jobs[0].acc != TC_NONE) { err = "lane program: first job must be the tile prologue"; return false; }
    auto conv = [](const EpiJob &e) {
        LaneJob j;
        memset(&j, 0, sizeof(j));
        j.kind = e.kind;
        j.enc = e.enc;
        j.ncols = e.ncols;
        j.bias_off = e.bias_off;
        j.save_slot = e.save_slot;
        j.enc_save_slot = e.enc_save_slot;
        j.mask_slot = e.mask_slot;
        j.out_slot = e.out_slot == TC_NONE ? 0 : e.out_slot;
        return j;
    };
    out.jobs.push_back(conv(p.jobs[ji++]));
    while (ji < p.jobs.size()) {
        LaneJob j = conv(p.jobs[ji]);
        bool released = (p.jobs[ji].flags & TC_JOB_RELEASE_ACC) != 0;
        ++ji;
        while (!released && ji < p.jobs.size()) {
            const EpiJob &e = p.jobs[ji];
            if (e.kind != j.kind || e.mask_slot != j.mask_slot) { err = "lane program: cannot merge epilogue blocks"; return false; }
            j.ncols = (uint16_t)(j.ncols + e.ncols);
            if (e.enc != ENC_NONE) { j.enc = e.enc; j.enc_save_slot = e.enc_save_slot; }
            released = (e.flags & TC_JOB_RELEASE_ACC) != 0;
            ++ji;
        }
        out.jobs.push_back(j);
    }
    if (out.jobs.size() != out.gemms.size() + 1) { err = "lane program: job/GEMM count mismatch"; return false; }
    return true;
}

Lane2Program *tc2_upload(const LaneProgram &p, uint32_t, std::string &err) {
    if (p.ops.size() > (size_t)kMaxOps || p.gemms.size() > (size_t)kMaxGemms) { err = "tc2: program exceeds the kernel-parameter tables"; return nullptr; }
    static bool attr_done = false;
    if (!attr_done) {
        bool ok = cudaFuncSetAttribute(k_chain2<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<false, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_chain2<true, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChain2Smem) == cudaSuccess;
        if (!ok) { err = std::string("tc2: cudaFuncSetAttribute failed: ") + cudaGetErrorString(cudaGetLastError()); return nullptr; }
        attr_done = true;
    }
    Lane2Program *d = new Lane2Program();
    d->prog = p;
    return d;
}

// ---- constant-bank bias slots. Slot ownership is per process: `owner` is an opaque id of the context's engine,
// `version` changes whenever its biases do. Returns the slot to pass to tc2_launch.
namespace {
struct BiasSlot { const void *owner = nullptr; uint64_t version = 0; };
BiasSlot g_slots[kBiasSlots];
int g_next_slot = 0;
}  // namespace
int tc2_bias_upload(const void *owner, uint64_t version, const float *d_bias, int n_floats, cudaStream_t st) {
    if (n_floats > kBiasSlotFloats) return -1;
    int slot = -1;
    for (int i = 0; i < kBiasSlots; ++i) if (g_slots[i].owner == owner) slot = i;
    if (slot < 0) {
        for (int i = 0; i < kBiasSlots && slot < 0; ++i) if (!g_slots[i].owner) slot = i;
        if (slot < 0) {   // more live engines than slots: evict round-robin; its kernels may still be reading the bank
            slot = g_next_slot++ % kBiasSlots;
            cudaDeviceSynchronize();
        }
        g_slots[slot].owner = owner;
        g_slots[slot].version = ~version;
    }
    if (g_slots[slot].version != version) {
        cudaMemcpyToSymbolAsync(c_bias, d_bias, sizeof(float) * n_floats, sizeof(float) * slot * kBiasSlotFloats, cudaMemcpyDeviceToDevice, st);
        g_slots[slot].version = version;
    }
    return slot;
}
void tc2_bias_release(const void *owner) {
    for (int i = 0; i < kBiasSlots; ++i) if (g_slots[i].owner == owner) g_slots[i] = BiasSlot();
}

void tc2_free(Lane2Program *d) { delete d; }

void tc2_launch(const Lane2Program *P, const Chain2Launch &l, cudaStream_t st) {
    Chain2Args a;
    memset(&a, 0, sizeof(a));
    memcpy(a.ops, P->prog.ops.data(), P->prog.ops.size() * sizeof(LaneOp));
    memcpy(a.gemms, P->prog.gemms.data(), P->prog.gemms.size() * sizeof(LaneGemm));
    memcpy(a.jobs, P->prog.jobs.data(), P->prog.jobs.size() * sizeof(LaneJob));
    a.n_ops = (int)P->prog.ops.size(); a.n_gemms = (int)P->prog.gemms.size();
    a.wpack = l.wpack; a.bias = l.bias; a.bias_floats = l.bias_floats; a.bias_slot = l.bias_slot;
    a.n_samples = l.n_samples;
    const int64_t n_tiles = (l.n_samples + NERF_TILE_M - 1) / NERF_TILE_M;
    a.n_pairs = (int)((n_tiles + 1) / 2);
    a.S = l.S;
    a.xyz_freqs = l.xyz_freqs; a.dir_freqs = l.dir_freqs;
    a.points = l.points; a.rays = l.rays; a.t = l.t; a.poses = l.poses; a.dirs = l.dirs; a.sigma = l.sigma; a.rgba = l.rgba; a.d_sigma = l.d_sigma; a.d_rgba = l.d_rgba;
    a.save_base = l.save_base; a.save_slots = l.save_slots; a.mask_base = l.mask_base; a.mask_slots = l.mask_slots;
    a.trace = l.trace;
    a.h2d_flag = l.h2d_flag; a.h2d_chunk_samples = l.h2d_chunk_samples;
    a.wide = l.wide; a.e_slot = l.e_slot; a.mask_words = l.mask_words;
    const int max_clusters = l.num_sms / 2;
    const int clusters = a.n_pairs < max_clusters ? a.n_pairs : max_clusters;
    const int grid = 2 * clusters;
    if (l.wide) {
        if (l.bwd) launch_pdl(k_chain2<true, true, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        else if (l.h2d_flag && l.save) launch_pdl(k_chain2<false, true, true, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        else if (l.h2d_flag) launch_pdl(k_chain2<false, false, true, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        else if (l.save) launch_pdl(k_chain2<false, true, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        else launch_pdl(k_chain2<false, false, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
    } else {
        if (l.trace) {   // (debug timeline: narrow mode only)
            if (l.bwd) launch_pdl(k_chain2<true, true, false, false, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
            else if (l.save) launch_pdl(k_chain2<false, true, false, false, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
            else launch_pdl(k_chain2<false, false, false, false, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        } else if (l.bwd) launch_pdl(k_chain2<true, true, false>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        else if (l.h2d_flag && l.save) launch_pdl(k_chain2<false, true, false, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        else if (l.h2d_flag) launch_pdl(k_chain2<false, false, false, true>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        else if (l.save) launch_pdl(k_chain2<false, true, false>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
        else launch_pdl(k_chain2<false, false, false>, dim3(grid), dim3(kThreads), kChain2Smem, st, a);
    }
}
