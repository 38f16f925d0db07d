/*
 * nerf_b200.h -- C ABI of the B200-native nerf-rs hot path.
 *
 * The reference (cadddr/nerf-rs) has no FFI or plugin interface: its hot path is a
 * set of Rust functions called from one closure (src/main.rs:56-114). This header
 * is the boundary a Rust host binds (see ffi/rust/ and INTEGRATION.md) so that
 *   dataset::get_multiview_batch -> NeRF::predict -> compositing -> Trainer::step
 * keep their call surface while the work runs in hand-written sm_100a kernels.
 * Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *  - plain C, no torch/libtorch types; host pointers are caller-owned and borrowed
 *    for the duration of the call only;
 *  - every function returns an int status (NERF_OK == 0, negative == error class)
 *    and never aborts or throws across the ABI (the reference panics instead:
 *    assert_eq! at src/model.rs:162-163,315-316, unwrap at :212,216,307,324);
 *  - a context is bound to one GPU and is NOT thread-safe (the reference is single
 *    threaded: src/display.rs:10-24); one process per GPU for multi-GPU;
 *  - all arithmetic on the path is f32 at the boundary; bf16 is used only inside the
 *    fused MLP kernels (fp32 accumulate);
 *  - there is no CPU fallback: nerf_create fails if no sm_100 device is present.
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: nerf_config.deterministic_grads, nerf_get_predictions, NERF_MLP_TCGEN05_SS (round 2); nerf_step returns when the loss is final */
#define NERF_B200_ABI_VERSION 2

/* status codes */
#define NERF_OK 0
#define NERF_ERR_INVALID_ARG (-1)   /* shape/size mismatch -- the reference's assert_eq! panics */
#define NERF_ERR_CUDA (-2)          /* a CUDA runtime call failed; see nerf_last_error */
#define NERF_ERR_UNSUPPORTED (-3)   /* config outside what the kernels implement */
#define NERF_ERR_COMM (-4)          /* NCCL failure */
#define NERF_ERR_STATE (-5)         /* call order violated, e.g. step before predict */
#define NERF_ERR_NO_DEVICE (-6)     /* no sm_100 GPU: there is no CPU fallback */

/* depth sampling modes (src/ray_sampling.rs:107-125) */
#define NERF_DEPTH_REFERENCE 0   /* t = u * 2.0, u sorted ascending per ray (reference behaviour) */
#define NERF_DEPTH_STRATIFIED 1  /* t = ((i + u) / S) * 2.0 (north-star stratified option) */

/* MLP implementations */
#define NERF_MLP_TCGEN05 0  /* fused tcgen05/TMEM kernels (the product path): CTA pairs (cta_group::2), two tiles in flight */
#define NERF_MLP_SIMT 1     /* plain CUDA-core kernels with the same bf16 rounding points;
                               on-device cross-check for sizes the CPU oracle cannot reach */
#define NERF_MLP_SIMT_FP32 2   /* the same without any bf16 rounding */
#define NERF_MLP_TCGEN05_SS 3  /* the SS-mode CTA-pair chain (activations in shared memory) at every width, kept for A/B runs */

typedef struct nerf_ctx nerf_ctx;

/*
 * Replaces the reference's compile-time constants (src/model.rs:7-13,
 * src/ray_sampling.rs:7-12) with a runtime description. nerf_default_config()
 * fills the north-star defaults; nerf_config_as_shipped() the literal reference.
 */
typedef struct nerf_config {
    int32_t struct_size;          /* sizeof(nerf_config), ABI check */
    int32_t image_w, image_h;     /* WIDTH, HEIGHT (ray_sampling.rs:7-8) */
    int32_t num_rays;             /* NUM_RAYS  (model.rs:7)  R */
    int32_t num_samples;          /* NUM_POINTS (model.rs:8) S, <= 256 */
    int32_t hidden;               /* HIDDEN_NODES (model.rs:12) W <= 512; the tensor-core kernels pad it to 64, 128, 256 or 512 */
    int32_t xyz_freqs;            /* positional-encoding octaves for xyz; 0 = raw xyz (INDIM=3, model.rs:11) */
    int32_t dir_freqs;            /* octaves for the view direction; -1 = no direction input (as shipped) */
    int32_t skip_layer;           /* concat [x_enc, h] after this layer's ReLU (5); 0 = none (as shipped) */
    int32_t use_rgb_head;         /* 1: fc9/fc10 output feeds compositing; 0: colours (s,s,s,1) (model.rs:190-206) */
    int32_t sigma_relu;           /* 0: raw sigma (model.rs:168-171); 1: relu */
    int32_t depth_mode;           /* NERF_DEPTH_* */
    int32_t mlp_impl;             /* NERF_MLP_* */
    int32_t max_rays_per_launch;  /* micro-batch bound for saved activations; 0 = auto */
    float learning_rate;          /* cli.rs:64-65, 5e-4 */
    float beta1, beta2, eps;      /* nn::Adam::default(): .9, .999, 1e-8 (model.rs:307) */
    int32_t deterministic_grads;  /* 1: weight gradients are reduced in a fixed order (per-segment partial blocks + one reduction
                                     pass) instead of with fp32 atomics: bit-identical gradients run to run, ~1 % slower */
} nerf_config;

int nerf_default_config(nerf_config *cfg);
int nerf_config_as_shipped(nerf_config *cfg);

/* ---- lifetime: NeRF::new + Trainer::new (src/model.rs:140-150, 306-309) -------- */
int nerf_create(const nerf_config *cfg, int device, nerf_ctx **out);
int nerf_destroy(nerf_ctx *ctx);
const char *nerf_last_error(const nerf_ctx *ctx);
const char *nerf_strerror(int status);
int nerf_abi_version(void);

/* ---- parameters: VarStore save/load surface (src/model.rs:211-217) ------------
 * Flat f32 blob, layer order fc1..fc10 (creation order model.rs:48-55, 89-90), each
 * weight[out,in] row-major then bias[out]. Weights are random-initialised
 * (U(-1/sqrt(in), 1/sqrt(in)), Philox) at create; set_weights injects exact values. */
int64_t nerf_num_params(const nerf_ctx *ctx);
int nerf_set_weights(nerf_ctx *ctx, const float *flat, int64_t n);
int nerf_get_weights(nerf_ctx *ctx, float *flat, int64_t n);
/* d(loss)/d(param) of the last step. With a communicator: the SUM over the ranks of their local gradients (what the
 * all-reduce produced); the update applied was that sum / nranks = the gradient of the concatenated batch's mean loss. */
int nerf_get_grads(nerf_ctx *ctx, float *flat, int64_t n);
int nerf_get_adam_state(nerf_ctx *ctx, float *m, float *v, int64_t n, int64_t *step);
int nerf_set_adam_state(nerf_ctx *ctx, const float *m, const float *v, int64_t n, int64_t step);

/* ---- dataset residency (replaces imgs/view_angles arguments of
 * dataset::get_multiview_batch, src/dataset.rs:63-71) ------------------------------
 * images: [n_views][H*W][4] RGBA f32 in [0,1] (image_loading.rs:13-18), row-major y*W+x.
 * view_angles: [n_angles][2] (yaw, pitch) as produced by get_view_angles
 * (image_loading.rs:67-80). cos/sin are evaluated on the host (libm), matching the
 * reference's f32::cos/sin, and the per-view rotation matrices are uploaded. */
int nerf_set_images(nerf_ctx *ctx, const float *rgba, int32_t n_views);
int nerf_set_view_angles(nerf_ctx *ctx, const float *yaw_pitch, int32_t n_angles);
/* The same residency from RGBA8 bytes (what image_loading.rs:6-24 decodes before its `as f32 / 255.`): images stay
 * 4 bytes per pixel on the device and the sampler's gold gather performs the IEEE division by 255, bit for bit. */
int nerf_set_images_rgba8(nerf_ctx *ctx, const uint8_t *rgba8, int32_t n_views);
/* load_image_as_array's decode step (image_loading.rs:7): 8-bit RGBA PNG -> bytes. out may be NULL to query the size;
 * anything that is not RGBA8 returns NERF_ERR_UNSUPPORTED (the reference produces an empty Vec for it). Host only. */
int nerf_load_png_rgba8(const char *path, uint8_t *out, int64_t capacity_bytes, int32_t *width, int32_t *height);
/* get_view_angles restated (image_loading.rs:67-80): writes 2n(n+1) pairs. */
int nerf_view_angles_grid(int32_t num_views_per_hemisphere, float *yaw_pitch_out, int32_t capacity);

/* ---- dataset::get_multiview_batch (src/dataset.rs:63-139) + the ray sampler
 * sample_and_rotate_ray_points_for_screen_coords (src/ray_sampling.rs:156-178) ------
 * indices_yx [R][2] ([y,x], dataset.rs:29) or NULL  -> Philox picks (stream 0/1)
 * view_index [n_picks] or NULL                      -> Philox picks with replacement (stream 2)
 *   rays are split evenly: R % n_picks must be 0 (dataset.rs:73-82) else NERF_ERR_INVALID_ARG
 * jitter [R][S] uniforms in [0,1) or NULL           -> Philox (stream 3); in REFERENCE mode
 *   caller-supplied jitter must be sorted ascending per ray (the reference sorts, :125);
 *   Philox jitter is sorted on the device.
 * randomize == 0 selects the deterministic branch u = i/S (ray_sampling.rs:112).
 * Optional host read-backs (any may be NULL): points [R][S][3], t [R][S], gold [R][4],
 * dirs [R][3], indices [R][2]. The batch stays resident in the context for predict. */
int nerf_get_batch(nerf_ctx *ctx, const int64_t *indices_yx, const int64_t *view_index, int32_t n_picks,
                   const float *jitter, int32_t randomize, uint64_t seed, float *out_points, float *out_t,
                   float *out_gold, float *out_dirs, int64_t *out_indices);

/* ---- NeRF::predict (src/model.rs:152-209) ------------------------------------------
 * nerf_predict: on the batch held by the context (no host inputs).
 * nerf_predict_points: the literal signature -- host query_points [B*3], distances
 *   [B] (t values), plus dirs [R*3] when the config has a direction input; sizes are
 *   checked like model.rs:162-163 (NERF_ERR_INVALID_ARG instead of a panic).
 * out_rgba [R*4] (may be NULL), out_sigma [R*S] (may be NULL).
 * train != 0 keeps the activations step() needs (the autograd tape the reference's
 * returned tensor carries, main.rs:58 -> :72). */
int nerf_predict(nerf_ctx *ctx, int32_t train, float *out_rgba, float *out_sigma);
int nerf_predict_points(nerf_ctx *ctx, const float *query_points, int64_t n_points_floats, const float *distances,
                        int64_t n_distances, const float *dirs, int32_t train, float *out_rgba, float *out_sigma);
/* The reference's predict returns a device Tensor that Trainer::step consumes (main.rs:58 -> :72) and the host reads only when
 * it draws (draw_predictions every eval_steps, main.rs:86-89). The same here: with BOTH outputs NULL nerf_predict* enqueues the
 * forward and returns without a device synchronisation, and nerf_get_predictions copies the current batch's pixels [R*4] and /
 * or densities [R*S] to the host whenever the caller wants them -- after predict, or after nerf_step / nerf_train_iter on that
 * batch (NERF_ERR_STATE if the context holds no prediction).
 * Host buffers handed to a call that returns without synchronising (nerf_predict_points with NULL outputs) follow the rule of
 * cudaMemcpyAsync: pageable memory (a Rust Vec, a numpy array) has been staged when the call returns; PINNED memory must stay
 * unchanged until the next call that returns a value from the device (nerf_step with a loss pointer, nerf_get_predictions,
 * nerf_sync). */
int nerf_get_predictions(nerf_ctx *ctx, float *out_rgba, float *out_sigma);

/* ---- compositing (src/model.rs:234-249), standalone on caller buffers ---------------
 * densities [R][S], colors [R][S][4], distances [R][S] = deltas between adjacent
 * samples (what the reference passes, model.rs:184-187). out [R][4]. */
int nerf_compositing(nerf_ctx *ctx, const float *densities, const float *colors, const float *distances,
                     int32_t num_rays, int32_t num_samples, float *out);

/* ---- Trainer::step (src/model.rs:311-325) -------------------------------------------
 * gold [R*4] host RGBA, or NULL to use the gold gathered by nerf_get_batch.
 * Runs MSE (model.rs:296-299) + backward + Adam. loss may be NULL (no host sync);
 * otherwise the call blocks for the scalar like f32::try_from(&loss) (model.rs:324) -- for the SCALAR, not for the step: the
 * loss is final after the compositing backward, and the call returns while dgrad, weight gradients and Adam are still running
 * (every later call is ordered behind them on the context's stream; nerf_sync waits for everything). A nerf_predict_points that
 * follows directly copies its inputs to the device under the rest of this step.
 * With a communicator the loss (here and in nerf_last_loss) is the RANK-LOCAL mean over this rank's R*4 elements;
 * the global-batch loss is the mean of the ranks' values. */
int nerf_step(nerf_ctx *ctx, const float *gold, int64_t n_gold, float *loss);

/* Whole iteration without host round trips: get_batch (Philox) -> predict -> step.
 * The loss of iteration k is readable after nerf_sync via nerf_last_loss. */
int nerf_train_iter(nerf_ctx *ctx, uint64_t seed);
int nerf_last_loss(nerf_ctx *ctx, float *loss);
int nerf_sync(nerf_ctx *ctx);

/* ---- novel-view render: the commented draw_valid_predictions (src/display.rs:55-94) -
 * rows [y0,y1) of a full frame at (yaw,pitch); out_rgba [(y1-y0)*W*4] (may be NULL),
 * out_0rgb [(y1-y0)*W] packed 0x00RRGGBB with (c*255) as u8 (display.rs:37-52). */
int nerf_render(nerf_ctx *ctx, float yaw, float pitch, int32_t y0, int32_t y1, int32_t randomize, uint64_t seed,
                float *out_rgba, uint32_t *out_0rgb);

/* The same frame sharded over the ranks of nerf_comm_init_rank: rank r renders the row band [r H/N, (r+1) H/N) (H % N
 * must be 0), then ONE NCCL all-gather assembles the frame on every rank -- the only collective of the render path.
 * out_rgba [H*W*4] / out_0rgb [H*W] receive the FULL frame (either may be NULL). Without a communicator: one band = all. */
int nerf_render_sharded(nerf_ctx *ctx, float yaw, float pitch, int32_t randomize, uint64_t seed, float *out_rgba,
                        uint32_t *out_0rgb);

/* ---- data-parallel training (no reference counterpart: it is single device) --------
 * One process per GPU. Rank 0 calls nerf_comm_unique_id, the host distributes the 128
 * bytes, every rank calls nerf_comm_init_rank. After that nerf_step all-reduces the
 * flat gradient (NCCL, sum) and Adam applies 1/nranks. */
int nerf_comm_unique_id(void *id128);
int nerf_comm_init_rank(nerf_ctx *ctx, const void *id128, int32_t rank, int32_t nranks);
int nerf_comm_destroy(nerf_ctx *ctx);

/* ---- measurement helpers (CUDA events on the context's own stream) ------------------ */
int nerf_timer_start(nerf_ctx *ctx);
int nerf_timer_stop(nerf_ctx *ctx, float *elapsed_ms);
/* per-kernel event timing of subsequent launches; names/ms arrays of length capacity */
int nerf_profile_enable(nerf_ctx *ctx, int32_t on);
int nerf_profile_read(nerf_ctx *ctx, char *names /*[capacity][32]*/, float *total_ms, int32_t *launches,
                      int32_t capacity, int32_t *count);
int64_t nerf_launch_count(const nerf_ctx *ctx);  /* kernels launched by this context so far */
int nerf_flush_l2(nerf_ctx *ctx);                /* overwrite a >L2-sized scratch buffer */

/* ---- per-batch logging projections (src/logging.rs, src/display.rs:96-110) computed on the device from the resident batch.
 * Every pointer is a caller-owned HOST buffer or NULL to skip that output. Density outputs and `prediction` need the
 * batch's forward to have run: nerf_predict*, or a completed nerf_step / nerf_train_iter (NERF_ERR_STATE otherwise). */
typedef struct nerf_metrics {
    double *screen_x;        /* [image_w]  log_screen_coords (logging.rs:13-25): counts of indices[r][0] -- the reference binds `[x, y]` to the stored [y, x] pair */
    double *screen_y;        /* [image_h]  counts of indices[r][1] */
    double *t_hist;          /* [2000]     log_query_distances (logging.rs:27-39): bucket floor(500 t) */
    uint32_t *world_yx;      /* [100*100]  log_query_points_as_maps (logging.rs:41-107): 0x00FFFFFF where a sample lands */
    uint32_t *world_zx;
    uint32_t *world_yz;
    double *density_x;       /* [2000]     log_densities (logging.rs:109-134): density sums per bucket floor(500 (w + 1)) */
    double *density_y;
    double *density_z;
    uint32_t *density_yx;    /* [100*100]  log_density_maps (logging.rs:136-195): last sample's max(density, 0) as 0x00RRGGBB */
    uint32_t *density_zx;
    uint32_t *density_yz;
    uint32_t *prediction;    /* [image_h*image_w] draw_predictions (display.rs:96-110): the batch's pixels as 0x00RRGGBB, 0 elsewhere */
} nerf_metrics;
int nerf_log_metrics(nerf_ctx *ctx, const nerf_metrics *out);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
