/*
 * oracle/ray_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Scalar f32 CPU restatement of the reference's ray geometry and depth
 * sampling, written from the formulas, not copied: the reference is Rust on
 * top of the un-vendored `vecmath = "1.0.0"` crate, whose published
 * semantics (left-to-right dot products, scale-by-reciprocal-length
 * normalisation, no FMA) are restated here.
 *
 * Reference anchors (all under /root/reference):
 *   src/ray_sampling.rs:7-16    constants WIDTH/HEIGHT, HITHER, FOV, T_FAR, UP/AT/FROM
 *   src/ray_sampling.rs:20-26   rotateYaw
 *   src/ray_sampling.rs:32-69   rotatePitch (Rodrigues matrix built per call)
 *   src/ray_sampling.rs:79-93   screen_to_world
 *   src/ray_sampling.rs:96-142  sample_points_along_ray_and_rotate
 *   src/ray_sampling.rs:156-178 sample_and_rotate_ray_points_for_screen_coords
 *   src/image_loading.rs:67-80  get_view_angles
 *   src/dataset.rs:63-139       get_multiview_batch (batch layout, [y,x], gold gather)
 *
 * Parity pins (the only golden vectors the reference's own tests hold):
 *   src/ray_sampling.rs:443-449 rotateYaw([1,2,3], pi/2) == [3.0, 2.0, -1.0000001]
 *   src/ray_sampling.rs:70-77   rotatePitch round trip on [0,0,1] is exact
 * plus the three property tests at :368-441. tests/test_oracle_ray.py checks all.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: the reference arithmetic is unfused mul-then-add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef float v3[3];

/* src/ray_sampling.rs:10-16 */
static const float HITHER = 0.05f;
static const float T_FAR = 2.0f;
static const float UP[3] = {0.f, 1.f, 0.f};
static const float AT[3] = {0.f, 0.f, 1.f};
static const float FROM[3] = {0.f, 0.f, -1.f};

static float fov(void) {
    /* std::f32::consts::PI / 3. evaluated in f32 (ray_sampling.rs:11).
     * volatile: tan(FOV/2) lies within 1e-11 of an f32 rounding tie, so a compile-time
     * (MPFR) fold gives 0x3f13cd3a while the runtime libm call the reference makes
     * (f32::tan is not const) gives 0x3f13cd3b. Pinned to the runtime value:
     * off = tanf(FOV/2)*HITHER = 0.028867517. */
    volatile float pi = 3.14159265358979323846f;
    return pi / 3.f;
}

/* ---- vecmath 1.0.0 restated ------------------------------------------- */
static float dot3(const float *a, const float *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static void sub3(const float *a, const float *b, float *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static void add3(const float *a, const float *b, float *o) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static void scale3(const float *a, float s, float *o) { o[0] = a[0] * s; o[1] = a[1] * s; o[2] = a[2] * s; }
static void cross3(const float *a, const float *b, float *o) {
    float r0 = a[1] * b[2] - a[2] * b[1];
    float r1 = a[2] * b[0] - a[0] * b[2];
    float r2 = a[0] * b[1] - a[1] * b[0];
    o[0] = r0; o[1] = r1; o[2] = r2;
}
static void normalized3(const float *a, float *o) {
    /* vec3_normalized = vec3_scale(a, 1/len), len = sqrt(dot(a,a)) */
    float inv = 1.f / sqrtf(dot3(a, a));
    scale3(a, inv, o);
}
static void mat3_mul_row(const float a[3][3], const float b[3][3], float o[3][3]) {
    /* row_mat3_mul: o[i][j] = dot(a[i], column j of b) */
    float t[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float col[3] = {b[0][j], b[1][j], b[2][j]};
            t[i][j] = dot3(a[i], col);
        }
    memcpy(o, t, sizeof(t));
}

/* ---- ray_sampling.rs:20-26 -------------------------------------------- */
void oracle_rotate_yaw(const float *v, float angle, float *out) {
    float c = cosf(angle), s = sinf(angle);
    float rot[3][4] = {{c, 0.f, s, 0.f}, {0.f, 1.f, 0.f, 0.f}, {-s, 0.f, c, 0.f}};
    float r[3];
    for (int i = 0; i < 3; ++i) /* row_mat3x4_transform_pos3 */
        r[i] = ((rot[i][0] * v[0] + rot[i][1] * v[1]) + rot[i][2] * v[2]) + rot[i][3];
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

/* ---- ray_sampling.rs:32-69: the 3x3 the reference rebuilds per point ---- */
void oracle_pitch_matrix(float angle, float rot[3][3]) {
    float d[3], v[3], cr[3], u[3];
    sub3(AT, FROM, d);
    normalized3(d, v);
    cross3(v, UP, cr);
    normalized3(cr, u);
    float ux = u[0], uy = u[1], uz = u[2];
    float cross_mat[3][3] = {{0.f, -uz, uy}, {uz, 0.f, -ux}, {-uy, ux, 0.f}};
    float outer[3][3] = {{ux * ux, ux * uy, ux * uz}, {uy * ux, uy * uy, uy * uz}, {uz * ux, uz * uy, uz * uz}};
    float c = cosf(angle), s = sinf(angle);
    float idc[3][3] = {{c, 0.f, 0.f}, {0.f, c, 0.f}, {0.f, 0.f, c}};
    float ids[3][3] = {{s, 0.f, 0.f}, {0.f, s, 0.f}, {0.f, 0.f, s}};
    float id[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    float cs[3][3], imc[3][3], oc[3][3];
    mat3_mul_row(cross_mat, ids, cs);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) imc[i][j] = id[i][j] - idc[i][j];
    mat3_mul_row(outer, imc, oc);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) rot[i][j] = (idc[i][j] + cs[i][j]) + oc[i][j];
}

void oracle_rotate_pitch(const float *v, float angle, float *out) {
    float rot[3][3];
    oracle_pitch_matrix(angle, rot);
    float r[3];
    for (int i = 0; i < 3; ++i) { /* col_mat3_transform: dot(column i, v) */
        float col[3] = {rot[0][i], rot[1][i], rot[2][i]};
        r[i] = dot3(col, v);
    }
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

/* ---- ray_sampling.rs:79-93 -------------------------------------------- */
void oracle_screen_to_world(float x, float y, float width, float height, float *to) {
    float off = tanf(fov() / 2.f) * HITHER;
    float offset_left = off - 2.f * off * x / width;
    float offset_up = off - 2.f * off * y / height;
    float d[3], view[3], cr[3], left[3], a[3], b[3], c[3], s1[3], s2[3];
    sub3(AT, FROM, d);
    normalized3(d, view);
    cross3(view, UP, cr);
    normalized3(cr, left);
    scale3(view, HITHER, a);
    scale3(left, offset_left, b);
    scale3(UP, offset_up, c);
    add3(a, b, s1);
    add3(s1, c, s2);
    normalized3(s2, to);
}

/* ---- ray_sampling.rs:96-142 -------------------------------------------
 * u: S caller-supplied uniforms (replaces rand::random::<f32>(), :110), or
 * NULL for the deterministic branch t = i/S (:112). The (point, t) pairs are
 * stably sorted ascending by t (:125) before the per-point rotations (:128-132). */
typedef struct { float p[3]; float t; int idx; } pt_t;
static int cmp_pt(const void *a, const void *b) {
    const pt_t *x = (const pt_t *)a, *y = (const pt_t *)b;
    if (x->t < y->t) return -1;
    if (x->t > y->t) return 1;
    return x->idx - y->idx; /* stable, like slice::sort_by */
}

void oracle_sample_points_along_ray_and_rotate(const float *from, const float *to, float yaw, float pitch,
                                               int num_samples, const float *u, float *points /*[S][3]*/,
                                               float *locations /*[S]*/) {
    pt_t *pl = (pt_t *)malloc(sizeof(pt_t) * (size_t)num_samples);
    for (int i = 0; i < num_samples; ++i) {
        float t = u ? u[i] : (float)i / (float)num_samples;
        t *= (T_FAR - HITHER) + HITHER; /* :114, precedence as written */
        float sc[3];
        scale3(to, t, sc);
        add3(from, sc, pl[i].p); /* :115 */
        pl[i].t = t;
        pl[i].idx = i;
    }
    qsort(pl, (size_t)num_samples, sizeof(pt_t), cmp_pt);
    for (int i = 0; i < num_samples; ++i) {
        float y3[3];
        oracle_rotate_yaw(pl[i].p, yaw, y3);
        oracle_rotate_pitch(y3, pitch, points + 3 * i);
        locations[i] = pl[i].t;
    }
    free(pl);
}

/* ---- ray_sampling.rs:156-178 ------------------------------------------
 * indices are [y,x] pairs (dataset.rs:29). u is [n_rays][S] or NULL. */
void oracle_sample_rays_for_screen_coords(const int64_t *indices_yx, int n_rays, int num_points, float yaw, float pitch,
                                          const float *u, int img_w, int img_h, float *points /*[n][S][3]*/,
                                          float *locations /*[n][S]*/) {
    for (int r = 0; r < n_rays; ++r) {
        float to[3];
        oracle_screen_to_world((float)indices_yx[2 * r + 1], (float)indices_yx[2 * r + 0], (float)img_w, (float)img_h, to);
        oracle_sample_points_along_ray_and_rotate(FROM, to, yaw, pitch, num_points, u ? u + (size_t)r * num_points : 0,
                                                  points + (size_t)r * num_points * 3, locations + (size_t)r * num_points);
    }
}

/* world-space view direction of a pixel's ray: the reference rotates points,
 * not the camera (ray_sampling.rs:95 "TODO"), so the direction that matches its
 * points is rotatePitch(rotateYaw(to)). Used by the north-star direction input. */
void oracle_ray_dirs_for_screen_coords(const int64_t *indices_yx, int n_rays, float yaw, float pitch, int img_w,
                                       int img_h, float *dirs /*[n][3]*/) {
    for (int r = 0; r < n_rays; ++r) {
        float to[3], y3[3];
        oracle_screen_to_world((float)indices_yx[2 * r + 1], (float)indices_yx[2 * r + 0], (float)img_w, (float)img_h, to);
        oracle_rotate_yaw(to, yaw, y3);
        oracle_rotate_pitch(y3, pitch, dirs + 3 * r);
    }
}

/* ---- image_loading.rs:67-80: 2n*(n+1) (yaw,pitch) pairs, f32 running sums */
int oracle_get_view_angles(int num_views, float *out /*[2n(n+1)][2]*/) {
    const float pi = 3.14159265358979323846f;
    float rot_ver = 0.f, rot_hor = 0.f;
    int k = 0;
    for (int i = 0; i < 2 * num_views; ++i) {
        for (int j = 0; j < num_views + 1; ++j) {
            out[2 * k] = rot_hor;
            out[2 * k + 1] = rot_ver;
            ++k;
            rot_ver = rot_ver + pi / (float)num_views;
        }
        rot_hor = rot_hor + pi / (float)num_views;
        rot_ver = 0.f;
    }
    return k;
}

/* ---- dataset.rs:63-139: batch assembly with caller-supplied randomness ---
 * indices_yx[R][2], view_index[V] (views picked with replacement, :88-93),
 * rays split evenly: bsz = R / V (:73-82, asserted by the caller).
 * imgs: [n_views][H*W][4] RGBA f32, gathered at y*W+x (:111-114). */
int oracle_get_multiview_batch(const float *imgs, const float *view_angles, int n_views, const int64_t *indices_yx,
                               const int64_t *view_index, int num_rays, int num_points, const float *u, int img_w,
                               int img_h, float *points, float *locations, float *gold) {
    if (num_rays % n_views != 0) return -1;
    int bsz = num_rays / n_views;
    for (int i = 0; i < n_views; ++i) {
        int n = (int)view_index[i];
        float yaw = view_angles[2 * n], pitch = view_angles[2 * n + 1];
        oracle_sample_rays_for_screen_coords(indices_yx + 2 * (size_t)i * bsz, bsz, num_points, yaw, pitch,
                                             u ? u + (size_t)i * bsz * num_points : 0, img_w, img_h,
                                             points + (size_t)i * bsz * num_points * 3,
                                             locations + (size_t)i * bsz * num_points);
        for (int r = 0; r < bsz; ++r) {
            int64_t y = indices_yx[2 * ((size_t)i * bsz + r)], x = indices_yx[2 * ((size_t)i * bsz + r) + 1];
            const float *px = imgs + (((size_t)n * img_h * img_w) + (size_t)(y * img_w + x)) * 4;
            memcpy(gold + ((size_t)i * bsz + r) * 4, px, 4 * sizeof(float));
        }
    }
    return 0;
}

/* ---- Philox4x32-10 (Salmon et al. 2011), the build's counter-based jitter.
 * Not in the reference (it uses rand 0.8's thread RNG, ray_sampling.rs:110);
 * restated here so the CUDA generator can be checked bit-exactly. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

/* u for element `index` of stream `stream`: counter=(index lo, index hi, stream, 0),
 * key=(seed lo, seed hi); u = (x0 >> 8) * 2^-24 in [0,1). */
void oracle_philox_uniform(uint64_t seed, uint32_t stream, uint64_t first_index, int64_t n, float *out) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t idx = first_index + (uint64_t)i;
        uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), stream, 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        out[i] = (float)(c[0] >> 8) * (1.0f / 16777216.0f);
    }
}
