/* nerf_driver.c -- a native (plain C) caller of the C ABI in include/nerf_b200.h, doing what src/main.rs:44-72 does with the
 * reference's own functions: load the scene, then per iteration get_multiview_batch -> NeRF::predict(query_points, distances)
 * -> Trainer::step(pred, gold). No Python, no torch: this is the call sequence a Rust `extern "C"` binding makes
 * (ffi/rust/src/lib.rs); tests/test_cpp_host.py compiles it, runs it on the GPU and compares its dump with the CPU oracle.
 *
 * usage: nerf_driver <input.bin> <output.bin>
 * input : int32 hdr[8] = {image_w, image_h, num_rays, num_samples, hidden, n_views, n_picks, n_steps}
 *         f32 weights[n_params], f32 images[n_views][w*h][4], f32 yaw_pitch[n_views][2],
 *         per step: i64 indices[R][2] (y, x), i64 view_index[n_picks], f32 jitter[R][S]
 * output: i64 n_params; per step: f32 points[R][S][3], t[R][S], gold[R][4], pixels[R][4], sigma[R][S], loss; then f32 weights[n_params] */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "nerf_b200.h"

#define CHECK(call)                                                                                         \
    do {                                                                                                    \
        int rc_ = (call);                                                                                   \
        if (rc_ != NERF_OK) {                                                                               \
            fprintf(stderr, "%s -> %d (%s): %s\n", #call, rc_, nerf_strerror(rc_), ctx ? nerf_last_error(ctx) : ""); \
            return 10 - rc_;                                                                                \
        }                                                                                                   \
    } while (0)

static void *rd(FILE *f, size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read (%zu bytes)\n", bytes); exit(3); }
    return p;
}

int main(int argc, char **argv) {
    nerf_ctx *ctx = NULL;
    if (argc != 3) return 2;
    FILE *in = fopen(argv[1], "rb"), *out = fopen(argv[2], "wb");
    if (!in || !out) return 2;
    int32_t *hdr = rd(in, 8 * sizeof(int32_t));
    const int w = hdr[0], h = hdr[1], R = hdr[2], S = hdr[3], n_views = hdr[5], n_picks = hdr[6], n_steps = hdr[7];

    nerf_config cfg;
    CHECK(nerf_default_config(&cfg));                 /* the north-star network: posenc 10/4, skip into fc6, RGB head */
    cfg.image_w = w; cfg.image_h = h; cfg.num_rays = R; cfg.num_samples = S; cfg.hidden = hdr[4];
    CHECK(nerf_create(&cfg, 0, &ctx));                /* NeRF::new + Trainer::new (model.rs:140-150, 306-309) */
    const int64_t n_params = nerf_num_params(ctx);
    float *weights = rd(in, sizeof(float) * (size_t)n_params);
    CHECK(nerf_set_weights(ctx, weights, n_params));
    float *images = rd(in, sizeof(float) * 4 * (size_t)n_views * w * h);
    CHECK(nerf_set_images(ctx, images, n_views));     /* load_multiple_images_as_arrays (main.rs:44-47) */
    float *angles = rd(in, sizeof(float) * 2 * (size_t)n_views);
    CHECK(nerf_set_view_angles(ctx, angles, n_views));
    fwrite(&n_params, sizeof(n_params), 1, out);

    const size_t B = (size_t)R * S;
    float *points = malloc(sizeof(float) * 3 * B), *t = malloc(sizeof(float) * B), *gold = malloc(sizeof(float) * 4 * R);
    float *dirs = malloc(sizeof(float) * 3 * R), *pixels = malloc(sizeof(float) * 4 * R), *sigma = malloc(sizeof(float) * B);
    for (int it = 0; it < n_steps; ++it) {
        int64_t *idx = rd(in, sizeof(int64_t) * 2 * (size_t)R);
        int64_t *vi = rd(in, sizeof(int64_t) * (size_t)n_picks);
        float *jit = rd(in, sizeof(float) * B);
        /* dataset::get_multiview_batch (main.rs:57-58): host indices + jitter in, query points / distances / gold out */
        CHECK(nerf_get_batch(ctx, idx, vi, n_picks, jit, 1, 0, points, t, gold, dirs, NULL));
        /* NeRF::predict(query_points, distances) (main.rs:61-64, model.rs:152): flat host arrays, pixels + densities back */
        CHECK(nerf_predict_points(ctx, points, (int64_t)(3 * B), t, (int64_t)B, dirs, 1, pixels, sigma));
        /* Trainer::step(&pred, gold, &iter) (main.rs:66-72, model.rs:311): loss back as a host f32 */
        float loss = -1.f;
        CHECK(nerf_step(ctx, gold, (int64_t)(4 * R), &loss));
        fwrite(points, sizeof(float), 3 * B, out); fwrite(t, sizeof(float), B, out); fwrite(gold, sizeof(float), 4 * (size_t)R, out);
        fwrite(pixels, sizeof(float), 4 * (size_t)R, out); fwrite(sigma, sizeof(float), B, out); fwrite(&loss, sizeof(float), 1, out);
        printf("iter %d loss %.6f\n", it, loss);
        free(idx); free(vi); free(jit);
    }
    CHECK(nerf_get_weights(ctx, weights, n_params));
    fwrite(weights, sizeof(float), (size_t)n_params, out);
    fclose(out);
    CHECK(nerf_destroy(ctx));
    return 0;
}
