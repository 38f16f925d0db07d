/*
 * nerf_b200_debug.h -- test-only entry points of libnerf_b200.so (not part of the
 * drop-in boundary; no reference counterpart). They expose host-side tables and saved
 * intermediate panels so tests/ can check the fused MLP kernels layer by layer.
 */
#ifndef NERF_B200_DEBUG_H
#define NERF_B200_DEBUG_H

#include <stdint.h>

#include "nerf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Copy the per-tile program tables the fused MLP kernels execute for `cfg` (host only, no GPU
 * needed). program: 0 = forward (training), 1 = forward (inference), 2 = backward dgrad chain.
 * On entry *n_x holds the capacity of each array (in records), on exit the record count.
 * info[0..9] = sizeof(MmaOp), sizeof(EpiJob), sizeof(PackChunk), sizeof(WgradUnit), packed weight
 * bytes, activation slots, gradient slots, mask slots, padded bias floats, parameter count. */
int nerf_debug_plan(const nerf_config *cfg, int32_t program, void *ops, int32_t *n_ops, void *jobs, int32_t *n_jobs,
                    void *chunks, int32_t *n_chunks, void *units, int32_t *n_units, int32_t *info);

/* The same programs regrouped per GEMM for the CTA-pair kernel (mlp_tc2.cu): ops {u32 w_off; u16 n; u8 a_slot; u8 kcount},
 * gemms {u16 op_begin, op_end}, jobs {u8 kind, enc; u16 ncols, bias_off; i16 save_slot, enc_save_slot, mask_slot;
 * u8 out_slot, pad; u16 pad} (job 0 = tile prologue, job g+1 = epilogue of GEMM g). Capacities in, counts out. */
int nerf_debug_lane_plan(const nerf_config *cfg, int32_t program, void *ops, int32_t *n_ops, void *gemms, int32_t *n_gemms,
                         void *jobs, int32_t *n_jobs);

/* Padded-bias gather table: records of {uint32 dst_off; int64 src_base; int32 count, padded}. */
int nerf_debug_plan_biases(const nerf_config *cfg, void *out, int32_t *n);

/* Host-side pose matrices (rotateYaw 3x4, rotatePitch 3x3, src/ray_sampling.rs:20-69) and the
 * screen offset tan(FOV/2)*HITHER the library feeds the sampler. No GPU needed. */
int nerf_debug_host_pose(float yaw, float pitch, float *yaw3x4, float *pitch3x3, float *off);

/* Read one saved 16 KB panel image back: area 0 = activations, 1 = pre-activation gradients;
 * area 2 = relu bit masks (slot = mask slot, out = 128*8 uint32). */
/* the TS-mode chain programs (mlp_tc.h: TsOp, TsStep, PackChunk; steps[0] = tile prologue); program 0 fwd-train, 1 fwd-infer, 2 bwd.
   UNSUPPORTED for hidden > 256. info: sizeof(TsOp), sizeof(TsStep), sizeof(PackChunk), packed stream bytes */
int nerf_debug_ts_plan(const nerf_config *cfg, int32_t program, void *ops, int32_t *n_ops, void *steps, int32_t *n_steps, void *chunks,
                       int32_t *n_chunks, int32_t *info);
/* per-CTA cycle counters of the last TS-mode chain launch, [ctas][16] (library built with -DNERF_TC3_STATS; else UNSUPPORTED) */
int nerf_debug_tc3_stats(uint64_t *out, int32_t ctas);
/* (tag << 48 | clock) events of CTA 0's MMA thread in the last TS-mode chain launch (same debug build) */
int nerf_debug_tc3_trace(uint64_t *out, int32_t n);
/* NERF_B200_GUARD=1 (read once, before the first allocation): every device buffer gets guard bands and a NaN fill
   (csrc/guard.h). Returns the number of allocations whose bands were overwritten (0 = clean), -1 when the mode is off. */
int nerf_debug_check_guards(int32_t *n_allocations);
int nerf_debug_read_panel(nerf_ctx *ctx, int32_t area, int32_t tile, int32_t slot, void *out);

/* Run one chain program (0 forward-train, 1 forward-inference, 2 backward dgrad) on the resident
 * batch with clock64 tracing of CTA 0. out: [3 roles (MMA issuer, epilogue warp 0, producer)][2048][2]. */
int nerf_debug_trace(nerf_ctx *ctx, int32_t program, uint64_t *out);

/* The weight-gradient work split for `cfg` over n_ctas CTAs and n_tiles 128-sample tiles (host only, no GPU needed):
 * out[cta][10] = n_seg, then (unit, tile_begin, tile_end) x 3. unit_cost_panels[u] = half panels a ring iteration of unit u loads
 * (capacity in *n_units on entry, unit count on exit). */
int nerf_debug_wgrad_partition(const nerf_config *cfg, int32_t n_ctas, int64_t n_tiles, int32_t *out, int32_t *unit_cost_panels, int32_t *n_units);

/* Per-CTA wall-clock marks of the last weight-gradient launch: out[cta][8] = start, first ring stage landed, all MMAs
 * done, end (ns, %globaltimer), first unit, segments, half-tile iterations, bytes loaded. Returns the CTA count (>= 0) or an error. */
int nerf_debug_wgrad_marks(nerf_ctx *ctx, uint64_t *out, int32_t capacity_ctas);

/* Time one HBM-bound stage kernel alone on synthetic device-resident inputs of num_rays x num_samples (sizes above the
 * 126 MB L2 stream HBM every launch): stage 0 = K-sample writing points + t, 1 = K-composite forward, 2 = K-composite
 * backward (+ fused MSE gradient and loss), 3 = K-adam over num_rays*num_samples parameters. CUDA events, mean of iters. */
int nerf_debug_bench_stage(nerf_ctx *ctx, int32_t stage, int32_t num_rays, int32_t num_samples, int32_t iters, float *ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif
