// kernels.h -- launcher declarations shared by the kernel translation units and the context.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------- sampling.cu
struct SampleArgs {
    const int32_t *pix_yx;     // [R][2] (y, x)
    const int32_t *view_pick;  // [n_picks] view ids, ray r uses view_pick[r / rays_per_pick]; NULL -> fixed_view
    int32_t rays_per_pick, fixed_view;
    int32_t gen_pix, gen_view, n_views;   // != 0: draw the pixel / view picks here (Philox) and record them
    int32_t *pix_out, *view_out;
    const ViewPose *poses;
    const float *jitter;       // [R][S] or NULL -> Philox
    const float *images;       // [V][H*W][4] or NULL
    const uint8_t *images_u8;  // [V][H*W][4] RGBA8 (alternative residency): gold = byte / 255
    int32_t num_rays, num_samples, img_w, img_h;
    int32_t randomize, depth_mode;
    float off;                 // tan(FOV/2) * HITHER, from the host
    uint64_t seed;
    int64_t ray_index_base;    // global ray offset for the Philox jitter counter (rank / micro-batch)
    RayRec *rays;              // [R]
    float *dirs;               // [R][3]
    float *t;                  // [R][S]
    float *points;             // [R][S][3] or NULL
    float *gold;               // [R][4]
};
void launch_sample(const SampleArgs &a, int num_sms, cudaStream_t st);
void launch_encode(const float *x, float *out, int64_t n, int freqs, int repeat, cudaStream_t st);
void launch_pack_0rgb(const float *rgba, uint32_t *out, int64_t n, cudaStream_t st);
void launch_full_frame_indices(int32_t *pix_yx, int y0, int y1, int img_w, cudaStream_t st);
void launch_flat_pixel_indices(int32_t *pix_yx, int64_t first, int n, int img_w, cudaStream_t st);

// ---------------------------------------------------------------- composite.cu
struct CompositeArgs {
    const float *sigma;        // [R][S]
    const float *colors;       // [R][S][4] or NULL -> (sigma,sigma,sigma,1)
    const float *t_or_delta;   // [R][S]
    int32_t input_is_delta;    // 1: deltas (standalone reference signature); 0: t values, deltas fused
    int32_t sigma_relu;
    int32_t num_rays, num_samples;
    float *out;                // [R][4]
    // backward
    const float *gold;         // [R][4] (fused MSE) -- used when d_out == NULL
    const float *d_out;        // [R][4] explicit upstream gradient or NULL
    float inv_count;           // 1 / (4 R_total)
    float *loss_partials;      // [grid] this launch's per-block sums of squared errors (fused MSE), or NULL
    float *d_sigma;            // [R][S]
    float *d_colors;           // [R][S][4]
    float *loss_out;           // fused mean loss (the last block sums the per-block partials in a fixed order), or NULL
    const float *loss_partials_first;  // first partial of the step (earlier micro-batch launches wrote theirs before this one's)
    int32_t loss_partials_prior;       // partials written by the step's earlier launches
    int32_t loss_partials_total;       // prior + this launch's grid (filled in by the launcher)
    float loss_scale;          // 1 / (4 R)
    unsigned int *done_counter;  // zero-initialised, reset by the kernel
};
void launch_composite_fwd(const CompositeArgs &a, int num_sms, cudaStream_t st);
int launch_composite_bwd(const CompositeArgs &a, int num_sms, cudaStream_t st);   // returns the grid size (= partials written)
void launch_fill_uniform(float *p, int64_t n, uint32_t seed, float lo, float hi, cudaStream_t st);   // synthetic inputs for stage timing

// ---------------------------------------------------------------- metrics.cu
struct MetricsArgs {
    const int32_t *pix_yx;     // [R][2]
    const RayRec *rays;        // [R]
    const ViewPose *poses;
    const float *t;            // [R][S]
    const float *points;       // [R][S][3] or NULL -> rebuilt from the ray records
    const float *sigma;        // [R][S] or NULL (no density outputs)
    const float *pixels;       // [R][4] or NULL (no prediction back buffer)
    int32_t num_rays, num_samples, img_w, img_h;
    unsigned int *screen_hist;            // [img_w + img_h] or NULL
    unsigned int *t_hist;                 // [2000] or NULL
    uint32_t *world_maps;                 // [3][100*100] or NULL
    double *density_hist;                 // [3][2000] (x, y, z) or NULL
    unsigned long long *density_maps;     // [3][100*100] keys or NULL
    unsigned long long *prediction;       // [img_h*img_w] keys or NULL
};
void launch_metrics(const MetricsArgs &a, int num_sms, cudaStream_t st);
void launch_metrics_resolve(const unsigned long long *keys, uint32_t *out, int64_t n, cudaStream_t st);

// ---------------------------------------------------------------- adam.cu
struct AdamArgs {
    float *p, *m, *v, *g;
    int64_t n;
    float lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps, grad_scale;
    int32_t zero_grad;
};
void launch_adam(const AdamArgs &a, int num_sms, cudaStream_t st);

// Fused gradient all-reduce + Adam over NVLink peer memory (data-parallel training, one process per GPU), in ONE kernel:
//   hand-shake 1  every rank's local gradient for this step is complete (flags[r] = step, system-scope release stores)
//   reduce-scatter rank r sums shard r of all ranks' gradient buffers (peer loads, fixed rank order) into the "reduced" half
//                  of its own buffer
//   hand-shake 2  (flags[8 + r] = step) every rank's reduced shard is in place
//   all-gather + Adam  every rank reads the N reduced shards (N - 1 of them over NVLink) and updates its replica
// Per rank 2 (N-1)/N gradient sizes cross NVLink instead of N - 1 (one-shot pull), every element is summed once, in rank
// order, by one rank -- all replicas apply bit-identical updates. Buffers: [0, n_pad) local gradient, [n_pad, 2 n_pad) reduced.
#define NERF_MAX_RANKS 8
struct AdamP2PArgs {
    AdamArgs adam;                        // adam.g = where the summed gradient is written back (nerf_get_grads)
    float *peer_grads[NERF_MAX_RANKS];    // rank r's gradient buffer pair (local | reduced) for this step (index == rank)
    unsigned int *peer_flags[NERF_MAX_RANKS];  // rank r's flag array [2 * NERF_MAX_RANKS]
    unsigned int *my_flags;
    unsigned int *block_counter;          // zero-initialised; the last block of phase 1 publishes hand-shake 2 and resets it
    int64_t n_pad;                        // floats between the local and the reduced half
    int32_t rank, nranks;
    uint32_t step;                        // > 0, monotonic
};
void launch_adam_p2p(const AdamP2PArgs &a, int num_sms, cudaStream_t st);
unsigned int adam_p2p_timeout_step();
void launch_init_uniform(float *p, const NetGeom &g, uint64_t seed, cudaStream_t st);

// ---------------------------------------------------------------- mlp_simt.cu
// Plain CUDA-core MLP (forward + backward) used as an on-device cross-check of the
// tcgen05 path. round_bf16 != 0 rounds weights/activations/pre-activation gradients
// to bf16 at the same points as the tensor-core kernels; 0 is pure fp32.
struct SimtBuffers {
    float *x_enc;    // [B][Cx]
    float *d_enc;    // [R][Cd]
    float *act;      // scratch for activations: layers h1..h7, df(fc8 out), h9 -- [B][max width] each
    float *dact;     // scratch for gradients
    int64_t act_stride;  // floats per saved activation slab
};
void simt_mlp_forward(const NetGeom &g, const float *params, const float *points, const float *dirs, int64_t num_rays,
                      int num_samples, int round_bf16, SimtBuffers &buf, float *sigma, float *rgba, cudaStream_t st);
void simt_mlp_backward(const NetGeom &g, const float *params, float *grads, int64_t num_rays, int num_samples,
                       int round_bf16, SimtBuffers &buf, const float *rgba, const float *d_sigma, const float *d_rgba,
                       cudaStream_t st);
size_t simt_act_floats_per_sample(const NetGeom &g);
