// mlp_tc.h -- fused tcgen05/TMEM MLP: program tables + host interface.
//
// One persistent "chain" kernel executes a per-tile PROGRAM (built on the host from the
// network geometry) for 128-sample tiles: a list of MMA ops (one 64-wide K panel each, fed
// by a bulk-copy weight ring) and a list of epilogue jobs (TMEM -> registers -> bias/act ->
// bf16 -> swizzled smem panel = next layer's A operand). The same kernel runs the forward
// chain (fc1..fc10, src/model.rs:97-131) and the backward dgrad chain; a second kernel does
// the weight gradients from the saved panels. See DESIGN.md section "K-mlp".
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "common.cuh"

// smem A-operand slots (16 KB panels): hidden panels 0..3 (rewritten in place), E = encoded inputs
#define TC_SLOT_E 4
#define TC_SLOT_E_WIDE 8   // hidden 257..512 (pair kernel only): hidden panels 0..7, encoded inputs in slot 8
#define TC_NUM_SLOTS 5
#define TC_NUM_STAGES 4
#define TC_STAGE_BYTES 32768   // one weight chunk [N <= 256][64 bf16]

// mbarrier ids
#define TC_BAR_FULL 0                       // +stage
#define TC_BAR_EMPTY (TC_NUM_STAGES)        // +stage
#define TC_BAR_ACC_FULL (2 * TC_NUM_STAGES)     // +accumulator set (2)
#define TC_BAR_ACC_FREE (2 * TC_NUM_STAGES + 2) // +accumulator set (2)
#define TC_BAR_READY (2 * TC_NUM_STAGES + 4)    // +group: 0 = slots 0,1; 1 = slots 2,3; 2 = slot E
#define TC_NUM_BARS (2 * TC_NUM_STAGES + 8)
#define TC_NONE 0xFF

// MmaOp.flags
#define TC_OP_FIRST 1u       // first K step of this accumulator block overwrites (accumulate = 0)
#define TC_OP_COMMIT_ACC 2u  // commit to acc_full[acc] after this op
#define TC_OP_GEMM_START 4u  // first op of a GEMM (a wide GEMM has a second TC_OP_FIRST op: its second 256-column half)

struct MmaOp {
    uint32_t w_off;   // byte offset of the weight chunk in the packed stream
    uint16_t n;       // MMA N == chunk rows (<= 256); chunk bytes = n * 128
    uint8_t a_slot;   // smem panel slot holding the A operand
    uint8_t acc;      // accumulator set (TMEM columns acc*256 ..); GEMMs alternate sets
    uint8_t flags;
    uint8_t wait0, wait1;  // barrier ids the MMA thread waits on before issuing (TC_NONE = none)
    uint8_t kcount;   // K16 steps (4 = full 64-wide panel)
};

// epilogue kinds
enum : uint8_t {
    EK_PROLOGUE_FWD = 0,  // positions -> posenc -> slot E
    EK_RELU = 1,          // relu(acc + bias) -> bf16 panels
    EK_LINEAR = 2,        // acc + bias -> bf16 panels (fc8 features)
    EK_SIGMA = 3,         // acc[0] + bias -> sigma[sample]
    EK_RGBA = 4,          // sigmoid(acc[0..3] + bias) -> rgba[sample]
    EK_PROLOGUE_BWD = 5,  // d_rgba * y(1-y) -> slot E (fc10 pre-activation gradient)
    EK_DMASK = 6,         // acc * relu_mask -> bf16 panels
    EK_DCOPY = 7,         // acc -> bf16 panels
};
// EpiJob.enc
enum : uint8_t { ENC_NONE = 0, ENC_X = 1, ENC_D = 2, ENC_DSIGMA = 3 };

// EpiJob.flags
#define TC_JOB_WAIT_ACC 1u     // first job of a GEMM: wait for acc_full[acc]
#define TC_JOB_RELEASE_ACC 2u  // last job of a GEMM: arrive on acc_free[acc] once its TMEM reads are done

struct EpiJob {
    uint8_t kind;
    uint8_t acc;        // accumulator set; TC_NONE for prologue jobs
    uint8_t ncols;      // accumulator columns processed (multiple of 64, or 16 for SIGMA/RGBA)
    uint8_t out_slot;   // first output smem slot (TC_NONE = none)
    uint8_t ready_bar;  // barrier id to arrive on once the output panels are written (TC_NONE = none)
    uint8_t enc;        // extra panel written to slot E by this job
    uint8_t enc_bar;    // barrier id for slot E
    uint8_t flags;
    int16_t save_slot;      // first per-tile global panel slot to save the output to (-1 = none)
    int16_t enc_save_slot;  // per-tile global slot for the slot-E panel (-1 = none)
    int16_t mask_slot;      // relu bit-mask slot: written by EK_RELU (train), read by EK_DMASK (-1 = none)
    uint16_t mask_word0;    // first 32-bit mask word of this block within the row (0 or 4)
    uint16_t bias_off;      // float offset into the padded bias array
    uint16_t acc_col;       // first TMEM column of this job's block (set*256 + 128*block)
};

// weight gradient unit: dW^T[in x out] block = P^T (inputs, M side) x Q (pre-activation grads, N side)
struct WgradUnit {
    uint8_t n_p, n_q;       // panels on the M side (1..5; 5 only with n_q <= 2) and N side (1..4)
    uint8_t pad[2];
    int16_t p_slot[6];      // per-tile slots in the activation area
    int16_t q_slot[4];      // per-tile slots in the gradient area
    int16_t sg_slot;        // >= 0: gradient-area slot of the d(sigma) panel -- the unit also produces the sigma row of fc8,
    int16_t pad1;           //       dW[sigma][m] = sum_s dsigma[s] P[s][m], on the CUDA cores from the P panels it already holds
    int32_t m_valid, n_valid;  // unpadded extents
    int64_t w_base;         // float offset of dW[out 0][in 0] of this block in the flat gradient blob
    int32_t w_row_stride;   // in_dim of the layer (distance between consecutive out rows)
    int64_t b_base;         // float offset of db[out 0], or -1
    int64_t sg_w_base;      // float offset of dW[sigma row][in 0 of this block]
    int64_t sg_b_base;      // float offset of db[sigma], or -1
};
#define kWgMaxSeg 3
struct WgradWork {          // one CTA's assignment: up to kWgMaxSeg (unit, tile range) segments, processed in order
    int32_t n_seg;
    struct { int32_t unit, tile_begin, tile_end; } seg[kWgMaxSeg];
    // deterministic mode: float offset of the segment's private partial block in the scratch buffer (else unused):
    // [n][m] dW^T block of 128*mblocks x N floats, then 256 bias sums, then 256 + 4 sigma-row sums
    int64_t part_off[kWgMaxSeg];
};
struct WgradRedUnit {       // deterministic mode: how k_wgrad_reduce folds a unit's partial blocks into the gradient blob
    int64_t w_base, b_base, sg_w_base, sg_b_base;
    int32_t w_row_stride, m_valid, n_valid, m_pad, n_cols;
    int32_t seg_begin, seg_end;      // range in the offset list, in ascending tile order
    int32_t has_sg;
};

struct PackChunk {
    uint32_t dst_off;
    int32_t n_rows;
    int64_t src_base;
    int32_t row_stride, col_stride;
    int32_t valid_rows, valid_cols;
};
struct PackBias {
    uint32_t dst_off;   // float offset in padded bias array
    int64_t src_base;   // float offset in params
    int32_t count, padded;
};

struct TcProgram {
    bool wide = false;        // hidden 257..512 (see TC_SLOT_E_WIDE)
    std::vector<MmaOp> ops;
    std::vector<EpiJob> jobs;
    std::vector<PackChunk> chunks;
    uint32_t wpack_bytes = 0;
};

// ---- v3 (mlp_tc3.cu): TS-mode CTA-pair chain for hidden <= 256. Hidden activations never live in shared memory: the
// epilogue writes them as bf16 into TENSOR MEMORY (tcgen05.st) and the next layer's MMAs read their A operand from there
// (tcgen05.mma with A in TMEM). Per lane: 128 accumulator columns + 128 activation columns (256 bf16 features), so every
// layer wider than 128 runs as two N = 128 half-GEMMs ("steps") into the same accumulator columns; the first half's
// converted output waits in registers until the second half's MMAs have finished reading the old activations.
#define TS_A_SMEM 0xFF         // TsOp.a_src: A operand = the lane's slot E_A in shared memory (SS-mode MMA): encoded positions / dPre10
#define TS_A_SMEM_B 0xFE       //             the lane's slot E_B: encoded direction / d(sigma)
// TsStep.pre_enc: after its own signal, the step's epilogue writes a slot-E panel of the lane's NEXT tile -- off every critical
// path, as soon as the current tile's last reader of that slot has completed (the first tile's panels are written up front)
#define TS_PRE_NONE 0
#define TS_PRE_A 1
#define TS_PRE_B 2
struct TsOp {
    uint32_t w_off;     // byte offset of the weight chunk [n rows][64 K]; CTA rank r of the pair loads rows [r n/2, (r+1) n/2)
    uint8_t n;          // MMA N (32, 64 or 128)
    uint8_t a_src;      // TS_A_SMEM, or the 64-feature panel p of the lane's TMEM activations (32-bit columns 32 p ..)
    uint8_t kcount;     // K16 steps (1, 2 or 4)
    uint8_t first;      // first K step of the step: overwrite the accumulator
};
struct TsStep {
    uint16_t op_begin, op_end;
    uint8_t kind;           // EK_*
    uint8_t enc;            // (prologue entry) ENC_* of panel A
    uint8_t ncols;          // accumulator columns (32 for SIGMA/RGBA, else 64 or 128)
    uint8_t a_col;          // first 32-bit TMEM column (within the lane's activation region) of this step's bf16 output
    uint8_t final_step;     // last step of its layer: the lane's activation columns may be rewritten after it
    uint8_t writes_a;       // the output is a later step's A operand (0: SIGMA/RGBA, or a gradient nobody consumes)
    uint8_t mask_word0;     // first 32-bit ReLU-mask word of this step's columns
    uint8_t pre_enc;        // TS_PRE_*
    uint16_t bias_off;
    int16_t save_slot, enc_save_slot, mask_slot;
    // prologue entry (steps[0]) only: kind/enc/enc_save_slot describe panel A; b_enc/b_save_slot panel B (ENC_NONE = no panel B)
    uint8_t b_enc, pad;
    int16_t b_save_slot;
};
struct TsProgram {
    std::vector<TsOp> ops;
    std::vector<TsStep> steps;     // steps[0] = tile prologue (no ops)
    std::vector<PackChunk> chunks;
    uint32_t wpack_bytes = 0;
};

struct TcPlan {
    TcProgram fwd_train, fwd_infer, bwd;
    TsProgram ts_fwd_train, ts_fwd_infer, ts_bwd;   // filled when hidden <= 256
    std::vector<PackBias> biases;
    uint32_t bias_floats = 0;
    std::vector<WgradUnit> units;
    int act_slots = 0, grad_slots = 0, mask_slots = 0;
    int np = 0, np2 = 0;  // hidden panels, fc9-output panels
    int e_slot = TC_SLOT_E, mask_words = 8;   // smem slot of the encoded inputs; 32-bit ReLU mask words per row and layer
};

// ---- v2 (mlp_tc2.cu): CTA-pair chain, two tiles ("lanes") in flight per CTA. The lane program is the v1
// program regrouped per GEMM: one op list slice and ONE epilogue job per GEMM, plus the tile prologue (job 0).
#define LANE_OP_KCOUNT 0x07u   // K16 steps (1, 2 or 4)
#define LANE_OP_HALF1 0x40u    // accumulate into TMEM columns +256 (second half of a wide GEMM)
#define LANE_OP_FIRST 0x80u    // first K step of its accumulator block: overwrite instead of accumulate
struct LaneOp {
    uint32_t w_off;   // byte offset of the full chunk; CTA rank r of the pair loads rows [r n/2, (r+1) n/2)
    uint16_t n;       // MMA N (multiple of 16)
    uint8_t a_slot;   // lane-relative smem slot of the A operand
    uint8_t kflags;   // LANE_OP_*
};
struct LaneGemm {
    uint16_t op_begin, op_end;
};
struct LaneJob {
    uint8_t kind, enc;
    uint16_t ncols;         // accumulator columns (64, 128, 256; 16 for SIGMA/RGBA)
    uint16_t bias_off;
    int16_t save_slot, enc_save_slot, mask_slot;
    uint8_t out_slot, pad0;
    uint16_t pad1;
};
struct LaneProgram {
    std::vector<LaneOp> ops;
    std::vector<LaneGemm> gemms;
    std::vector<LaneJob> jobs;   // gemms.size() + 1
};
bool make_lane_program(const TcProgram &p, LaneProgram &out, std::string &err);
struct Lane2Program;
Lane2Program *tc2_upload(const LaneProgram &p, uint32_t bias_floats, std::string &err);
void tc2_free(Lane2Program *d);
struct Chain2Launch {
    bool bwd, save;
    const uint8_t *wpack;
    const float *bias;
    int32_t bias_floats, bias_slot;
    int64_t n_samples;
    int32_t S, xyz_freqs, dir_freqs, num_sms;
    const float *points, *dirs;
    const RayRec *rays;          // fused sampling inputs, used when points == NULL
    const float *t;
    const ViewPose *poses;
    float *sigma, *rgba;
    const float *d_sigma, *d_rgba;
    uint8_t *save_base;
    int32_t save_slots;
    uint32_t *mask_base;
    int32_t mask_slots;
    int32_t wide, e_slot, mask_words;
    unsigned long long *trace;   // debug: [3][2048][4] clock64 stamps of CTA 0, or NULL
    const unsigned int *h2d_flag;   // points still arriving from the host: sample gs may be read once *h2d_flag > gs / h2d_chunk_samples
    int64_t h2d_chunk_samples;
};
void tc2_launch(const Lane2Program *P, const Chain2Launch &l, cudaStream_t st);
int tc2_bias_upload(const void *owner, uint64_t version, const float *d_bias, int n_floats, cudaStream_t st);
void tc2_bias_release(const void *owner);

struct Ts3Program;
Ts3Program *tc3_upload(const TsProgram &p, std::string &err);
void tc3_free(Ts3Program *d);
void tc3_launch(const Ts3Program *P, const Chain2Launch &l, cudaStream_t st);
int tc3_bias_upload(const void *owner, uint64_t version, const float *d_bias, int n_floats, cudaStream_t st);
void tc3_bias_release(const void *owner);
int tc3_debug_stats(unsigned long long *out, int ctas);
int tc3_debug_trace(unsigned long long *out, int n);

// host-only: build all tables for a geometry. Returns false (err set) if unsupported.
bool tc_build_plan(const NetGeom &g, TcPlan &plan, std::string &err);

struct TcState;
TcState *tc_create(const NetGeom &g, int64_t max_tiles, int num_sms, int version, std::string &err, bool deterministic = false);
void tc_destroy(TcState *s);
size_t tc_bytes_per_tile(const TcState *s);
void tc_pack_weights(TcState *s, const float *params, cudaStream_t st);
// points [n][3], dirs [rays][3]; n samples, S samples per ray. train -> save activations.
// Fused-sampling inputs for the CTA-pair kernel: with points == NULL the prologue rebuilds each sample position from
// its ray record, depth and view pose (bit-identical to the standalone sampler).
struct TcRayInputs {
    const RayRec *rays;
    const float *t;
    const ViewPose *poses;
    // caller-supplied points that are still being copied from the host on another stream (nerf_predict_points): the copy
    // stream bumps *h2d_flag after every chunk of h2d_chunk_samples samples; NULL = everything is already resident
    const unsigned int *h2d_flag;
    int64_t h2d_chunk_samples;
};
int tc_forward(TcState *s, const float *points, const float *dirs, int64_t n, int S, int train, float *sigma, float *rgba,
               cudaStream_t st, const TcRayInputs *fused = nullptr);
int tc_version(const TcState *s);
// dgrad chain + weight/bias gradients accumulated (+=) into grads
int tc_backward(TcState *s, const float *rgba, const float *d_sigma, const float *d_rgba, int64_t n, float *grads,
                cudaStream_t st, void (*between)(void *, const char *), void *user);
const char *tc_last_error(const TcState *s);
int tc_debug_read(TcState *s, int area, int64_t tile, int slot, void *out, cudaStream_t st);
// debug: run one chain program (0 fwd-train, 1 fwd-infer, 2 bwd) with clock64 tracing of CTA 0.
// host_out: [3 roles (mma, epilogue warp 0, producer)][2048 events][2] uint64.
void tc_wgrad_partition(const std::vector<WgradUnit> &units, int n_ctas, int64_t n_tiles, std::vector<WgradWork> &work);
int tc_debug_wgrad_marks(TcState *s, unsigned long long *out, int capacity_ctas, cudaStream_t st);
int tc_debug_trace(TcState *s, const float *points, const float *dirs, int64_t n, int S, int program, const float *rgba,
                   const float *d_sigma, const float *d_rgba, float *sigma_out, float *rgba_out, unsigned long long *host_out,
                   cudaStream_t st);
