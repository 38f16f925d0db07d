// context.cu -- implementation of the C ABI in include/nerf_b200.h.
//
// One nerf_ctx per GPU owns: the flat f32 parameter / gradient / Adam blobs, the resident
// dataset (images + per-view pose matrices), the current ray batch, the MLP engine state
// (tcgen05 chain kernels or the SIMT cross-check) and one CUDA stream on which everything
// is enqueued. Call surface mirrors src/main.rs:57-72:
//   get_multiview_batch -> NeRF::predict -> (compositing) -> Trainer::step.
#include <nvtx3/nvToolsExt.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/nerf_b200.h"
#include "../../include/nerf_b200_debug.h"
#include "comm.h"
#include "guard.h"
#include "kernels.h"
#include "mlp_tc.h"

namespace {

inline bool is_tc(int impl) { return impl == NERF_MLP_TCGEN05 || impl == NERF_MLP_TCGEN05_SS; }

struct Profiler {
    bool on = false;
    std::vector<std::string> names;
    std::map<std::string, int> index;
    struct Rec { int name; cudaEvent_t e0, e1; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    std::vector<double> total_ms;
    std::vector<int> launches;
    int open = -1;

    cudaEvent_t get_event() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(const char *name, cudaStream_t st) {
        if (!on) return;
        end(st);
        auto it = index.find(name);
        int id;
        if (it == index.end()) {
            id = (int)names.size();
            names.push_back(name);
            index[name] = id;
            total_ms.push_back(0.0);
            launches.push_back(0);
        } else id = it->second;
        Rec r{id, get_event(), get_event()};
        cudaEventRecord(r.e0, st);
        recs.push_back(r);
        open = (int)recs.size() - 1;
    }
    void end(cudaStream_t st) {
        if (!on || open < 0) return;
        cudaEventRecord(recs[open].e1, st);
        open = -1;
    }
    void collect(cudaStream_t st) {
        end(st);
        cudaStreamSynchronize(st);
        for (auto &r : recs) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { total_ms[r.name] += ms; launches[r.name] += 1; }
            pool.push_back(r.e0);
            pool.push_back(r.e1);
        }
        recs.clear();
    }
    void reset() {
        for (auto &t : total_ms) t = 0;
        for (auto &l : launches) l = 0;
    }
};

}  // namespace

struct nerf_ctx {
    nerf_config cfg;
    NetGeom g;
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launch_count = 0;
    Profiler prof;

    // parameters
    float *d_params = nullptr, *d_grads = nullptr, *d_m = nullptr, *d_v = nullptr;
    float *d_gacc[2] = {nullptr, nullptr};   // peer-memory data parallelism: local gradient accumulation buffers, alternating per step
    int64_t step = 0;
    bool weights_dirty = true;

    // dataset
    float *d_images = nullptr;
    uint8_t *d_images_u8 = nullptr;
    int n_img_views = 0;
    ViewPose *d_poses = nullptr;
    int n_poses = 0;
    ViewPose *d_render_pose = nullptr;
    float off = 0.f;

    // batch
    int R = 0, S = 0, chunk = 0;
    int64_t B = 0;
    int32_t *d_pix = nullptr, *d_view_pick = nullptr;
    RayRec *d_rays = nullptr;
    float *d_dirs = nullptr, *d_t = nullptr, *d_points = nullptr, *d_gold = nullptr, *d_jitter = nullptr;
    float *d_sigma = nullptr, *d_rgba = nullptr, *d_out = nullptr, *d_dsigma = nullptr, *d_drgba = nullptr;
    float *d_loss_partials = nullptr, *d_loss = nullptr;
    float *h_loss = nullptr;   // pinned
    int32_t *h_i32 = nullptr;  // pinned staging for index conversion
    size_t h_i32_cap = 0;
    bool batch_valid = false, predicted = false, acts_valid = false;
    bool outputs_valid = false;           // d_sigma / d_out hold the resident batch's densities and pixels (survives nerf_step)
    int gen_pix = 0, gen_view = 0, pick_views = 0;   // Philox picks requested for the next sampler launch
    bool fwd_deferred = false;            // nerf_train_iter on a micro-batched batch: the forward runs inside nerf_step
    bool points_valid = false;            // d_points holds the batch's sample positions (else: fused sampling in the MLP prologue)
    const ViewPose *batch_poses = nullptr;  // pose table the resident batch's ray records index

    // engines
    TcState *tc = nullptr;
    SimtBuffers simt{};
    bool simt_ready = false;

    // comm
    CommState comm;

    // timers / scratch
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaEvent_t ev_stage = nullptr;   // recorded after the asynchronous H2D copies out of the pinned index staging buffer
    bool stage_busy = false;
    uint8_t *d_flush = nullptr;
    size_t flush_bytes = 0;
    float *d_frame_rgba = nullptr;       // full-frame render targets (lazily allocated)
    uint32_t *d_frame_0rgb = nullptr;
    uint8_t *d_metrics = nullptr;        // scratch of nerf_log_metrics (lazily allocated)
    // nerf_predict_points: the host->device copy of the points overlaps the forward kernel (chunked copy on its own stream)
    cudaStream_t copy_stream = nullptr;   // host copies that overlap kernels: chunked points (nerf_predict_points), the early loss read (nerf_step)
    cudaEvent_t ev_main = nullptr, ev_dirs = nullptr, ev_all = nullptr;
    cudaEvent_t ev_cbwd = nullptr;        // recorded after a step's compositing backward: the loss is final, the batch inputs have no reader left
    bool inputs_free = false;             // the last enqueuing API call was a single-launch nerf_step (ev_cbwd covers every reader of the inputs)
    int h2d_word = 0;                     // which of the two arrival counters the next overlapped copy uses
    bool sample_side = false;             // nerf_train_iter: this batch's sampler may run on the copy stream behind ev_cbwd
    unsigned int *d_h2d_flag = nullptr, *h_h2d_seq = nullptr;   // two device counters (alternating per call); pinned 1, 2, 3, ... source values
    const unsigned int *h2d_flag_active = nullptr;               // non-NULL while the resident points are still arriving
    int64_t h2d_chunk_samples = 0;
};

namespace {

int fail(nerf_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}
#define CU(c, expr)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(c, NERF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));        \
    } while (0)

// Every public entry point that enqueues work starts with ENTER: device selection, and "the previous call was not the step
// whose compositing-backward event covers all readers of the batch inputs" (nerf_predict_points reads the flag before ENTER).
#define ENTER(c)                                  \
    do {                                          \
        CU(c, cudaSetDevice((c)->device));        \
        (c)->inputs_free = false;                 \
    } while (0)

// One Scope per kernel (group) launch: counts it, times it with CUDA events when profiling is on (nerf_profile_enable), and
// opens an NVTX range of the same name (header-only NVTX3: a no-op unless a tool such as Nsight Systems is attached).
struct Scope {
    nerf_ctx *c;
    Scope(nerf_ctx *c_, const char *name, int launches = 1) : c(c_) {
        c->launch_count += launches;
        nvtxRangePushA(name);
        c->prof.begin(name, c->stream);
    }
    ~Scope() {
        c->prof.end(c->stream);
        nvtxRangePop();
    }
};

int check_launch(nerf_ctx *c, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, NERF_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return NERF_OK;
}

void build_geom(const nerf_config &cfg, NetGeom &g) {
    memset(&g, 0, sizeof(g));
    g.W = cfg.hidden;
    // the tcgen05 kernels run 1, 2, 4 or 8 hidden panels of 64: every other width is zero-padded up to the next of those
    // (HIDDEN_NODES is a free constant in the reference, model.rs:12; padding costs tensor work, not correctness)
    g.Wp = cfg.hidden <= 64 ? 64 : (cfg.hidden <= 128 ? 128 : (cfg.hidden <= 256 ? 256 : (cfg.hidden <= 512 ? 512 : (cfg.hidden + 63) / 64 * 64)));
    g.W2 = cfg.hidden / 2;
    g.W2p = (g.W2 + 63) / 64 * 64;
    g.xyz_freqs = cfg.xyz_freqs;
    g.dir_freqs = cfg.dir_freqs;
    g.Cx = 3 + 6 * cfg.xyz_freqs;
    g.Cd = cfg.dir_freqs < 0 ? 0 : 3 + 6 * cfg.dir_freqs;
    g.skip_layer = cfg.skip_layer;
    g.use_rgb_head = cfg.use_rgb_head;
    g.sigma_relu = cfg.sigma_relu;
    g.n_layers = 10;
    int64_t off = 0;
    for (int l = 1; l <= 10; ++l) {
        int in, out;
        if (l <= 7) {
            in = (l == 1) ? g.Cx : g.W;
            if (g.skip_layer && l == g.skip_layer + 1) in = g.W + g.Cx;
            out = g.W;
        } else if (l == 8) { in = g.W; out = g.W + 1; }   // sigma || features (model.rs:55)
        else if (l == 9) { in = g.W + g.Cd; out = g.W2; } // model.rs:89
        else { in = g.W2; out = 4; }                      // model.rs:90
        g.L[l - 1].in_dim = in;
        g.L[l - 1].out_dim = out;
        g.L[l - 1].w_off = off;
        off += (int64_t)in * out;
        g.L[l - 1].b_off = off;
        off += out;
    }
    g.n_params = off;
}

// rotateYaw / rotatePitch matrices (ray_sampling.rs:20-26, 32-69), built like the reference
// builds them -- per call, in f32, libm cos/sin -- but once per view instead of once per point.
void mat3_mul_rows(const float a[3][3], const float b[3][3], float o[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o[i][j] = (a[i][0] * b[0][j] + a[i][1] * b[1][j]) + a[i][2] * b[2][j];
}
void make_pose(float yaw, float pitch, ViewPose &vp) {
    memset(&vp, 0, sizeof(vp));
    {
        const float c = cosf(yaw), s = sinf(yaw);
        const float m[3][4] = {{c, 0.f, s, 0.f}, {0.f, 1.f, 0.f, 0.f}, {-s, 0.f, c, 0.f}};
        memcpy(vp.yaw, m, sizeof(m));
    }
    {
        // v = normalize(AT - FROM) = (0,0,1); u = normalize(cross(v, UP))
        const float at_from[3] = {0.f - 0.f, 0.f - 0.f, 1.f - (-1.f)};
        const float inv = 1.f / sqrtf((at_from[0] * at_from[0] + at_from[1] * at_from[1]) + at_from[2] * at_from[2]);
        const float v[3] = {at_from[0] * inv, at_from[1] * inv, at_from[2] * inv};
        const float up[3] = {0.f, 1.f, 0.f};
        const float cr[3] = {v[1] * up[2] - v[2] * up[1], v[2] * up[0] - v[0] * up[2], v[0] * up[1] - v[1] * up[0]};
        const float inv2 = 1.f / sqrtf((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2]);
        const float ux = cr[0] * inv2, uy = cr[1] * inv2, uz = cr[2] * inv2;
        const float cross_m[3][3] = {{0.f, -uz, uy}, {uz, 0.f, -ux}, {-uy, ux, 0.f}};
        const float outer[3][3] = {{ux * ux, ux * uy, ux * uz}, {uy * ux, uy * uy, uy * uz}, {uz * ux, uz * uy, uz * uz}};
        const float c = cosf(pitch), s = sinf(pitch);
        const float idc[3][3] = {{c, 0.f, 0.f}, {0.f, c, 0.f}, {0.f, 0.f, c}};
        const float ids[3][3] = {{s, 0.f, 0.f}, {0.f, s, 0.f}, {0.f, 0.f, s}};
        float imc[3][3], cs[3][3], oc[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) imc[i][j] = (i == j ? 1.f : 0.f) - idc[i][j];
        mat3_mul_rows(cross_m, ids, cs);
        mat3_mul_rows(outer, imc, oc);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) vp.pitch[i][j] = (idc[i][j] + cs[i][j]) + oc[i][j];
    }
}

float screen_offset() {
    // tan(FOV/2) * HITHER evaluated at run time in f32 like f32::tan (ray_sampling.rs:80). The
    // volatile keeps the compiler from folding it: tan(pi/6) sits within 1e-11 of an f32 rounding
    // tie and a compile-time fold rounds the other way ("the tan(FOV/2) tie").
    volatile float pi = 3.14159265358979323846f;
    const float fov = pi / 3.f;
    return tanf(fov / 2.f) * NERF_HITHER;
}

int ensure_i32(nerf_ctx *c, size_t n) {
    if (c->h_i32_cap >= n) return NERF_OK;
    if (c->h_i32) cudaFreeHost(c->h_i32);
    c->h_i32 = nullptr;
    c->h_i32_cap = 0;
    CU(c, cudaMallocHost(&c->h_i32, n * sizeof(int32_t)));
    c->h_i32_cap = n;
    return NERF_OK;
}

void prof_between(void *user, const char *name) {
    nerf_ctx *c = (nerf_ctx *)user;
    if (name) { c->launch_count += 1; c->prof.begin(name, c->stream); }
    else c->prof.end(c->stream);
}

int ensure_packed(nerf_ctx *c) {
    if (c->weights_dirty && c->tc) {
        Scope s(c, "pack_weights", 1);
        tc_pack_weights(c->tc, c->d_params, c->stream);
    }
    c->weights_dirty = false;
    return check_launch(c, "pack_weights");
}

// MLP forward on rays [r0, r0+nr) of the current batch
int mlp_forward(nerf_ctx *c, int r0, int nr, int train) {
    const int64_t n = (int64_t)nr * c->S;
    const int64_t s0 = (int64_t)r0 * c->S;
    if (is_tc(c->cfg.mlp_impl)) {
        int rc = ensure_packed(c);
        if (rc) return rc;
        Scope s(c, train ? "mlp_fwd_train" : "mlp_fwd");
        TcRayInputs fused{c->d_rays + r0, c->d_t + s0, c->batch_poses, nullptr, 0};
        if (c->h2d_flag_active && r0 == 0 && nr == c->R) { fused.h2d_flag = c->h2d_flag_active; fused.h2d_chunk_samples = c->h2d_chunk_samples; }
        if (tc_forward(c->tc, c->points_valid ? c->d_points + 3 * s0 : nullptr, c->d_dirs + 3 * (int64_t)r0, n, c->S, train,
                       c->d_sigma + s0, c->d_rgba + 4 * s0, c->stream, &fused))
            return fail(c, NERF_ERR_INVALID_ARG, tc_last_error(c->tc));
    } else {
        Scope s(c, "mlp_fwd_simt", 12);
        simt_mlp_forward(c->g, c->d_params, c->d_points + 3 * s0, c->d_dirs + 3 * (int64_t)r0, nr, c->S,
                         c->cfg.mlp_impl == NERF_MLP_SIMT ? 1 : 0, c->simt, c->d_sigma + s0, c->d_rgba + 4 * s0, c->stream);
    }
    return check_launch(c, "mlp_forward");
}

int mlp_backward(nerf_ctx *c, int r0, int nr, float *grads) {
    const int64_t n = (int64_t)nr * c->S;
    const int64_t s0 = (int64_t)r0 * c->S;
    if (is_tc(c->cfg.mlp_impl)) {
        if (tc_backward(c->tc, c->d_rgba + 4 * s0, c->d_dsigma + s0, c->d_drgba + 4 * s0, n, grads, c->stream,
                        prof_between, c))
            return fail(c, NERF_ERR_INVALID_ARG, tc_last_error(c->tc));
    } else {
        Scope s(c, "mlp_bwd_simt", 30);
        simt_mlp_backward(c->g, c->d_params, grads, nr, c->S, c->cfg.mlp_impl == NERF_MLP_SIMT ? 1 : 0, c->simt,
                          c->d_rgba + 4 * s0, c->d_dsigma + s0, c->d_drgba + 4 * s0, c->stream);
    }
    return check_launch(c, "mlp_backward");
}

int composite_forward(nerf_ctx *c, int nr, float *out) {
    CompositeArgs a;
    memset(&a, 0, sizeof(a));
    a.sigma = c->d_sigma;
    a.colors = c->cfg.use_rgb_head ? c->d_rgba : nullptr;
    a.t_or_delta = c->d_t;
    a.input_is_delta = 0;
    a.sigma_relu = c->cfg.sigma_relu;
    a.num_rays = nr;
    a.num_samples = c->S;
    a.out = out;
    Scope s(c, "composite_fwd");
    launch_composite_fwd(a, c->num_sms, c->stream);
    return check_launch(c, "composite_fwd");
}

int do_predict(nerf_ctx *c, int train, float *out_rgba, float *out_sigma, bool skip_composite = false, bool forward_done = false) {
    if (!c->batch_valid) return fail(c, NERF_ERR_STATE, "predict: no batch (call nerf_get_batch or nerf_predict_points)");
    c->acts_valid = false;
    c->fwd_deferred = false;
    if (skip_composite && train && c->chunk < c->R) {
        // fused iteration over a batch larger than the saved-activation budget: rays are independent, so each micro-batch
        // runs forward -> compositing backward -> dgrad -> wgrad on its own inside nerf_step; no full forward first
        c->fwd_deferred = true;
        c->predicted = true;
        return NERF_OK;
    }
    // Nobody asked for the pixels yet (device-resident prediction): a training forward leaves compositing to nerf_step's
    // backward kernel, which recomputes the pixels anyway, or to nerf_get_predictions / nerf_log_metrics (ensure_outputs).
    if (!out_rgba && !out_sigma && train && c->chunk >= c->R) skip_composite = true;
    if (forward_done) {   // (nerf_predict_points already ran the single-launch forward under its overlapped host copy)
        if (train) c->acts_valid = true;
    } else {
        for (int r0 = 0; r0 < c->R; r0 += c->chunk) {
            const int nr = (c->R - r0 < c->chunk) ? c->R - r0 : c->chunk;
            const int keep = train && c->chunk >= c->R;
            int rc = mlp_forward(c, r0, nr, keep);
            if (rc) return rc;
            if (keep) c->acts_valid = true;
        }
    }
    // (a fused training iteration gets its pixels from the compositing backward kernel, which recomputes them)
    int rc = skip_composite ? NERF_OK : composite_forward(c, c->R, c->d_out);
    if (rc) return rc;
    c->predicted = true;
    c->outputs_valid = !skip_composite;
    if (out_rgba) CU(c, cudaMemcpyAsync(out_rgba, c->d_out, sizeof(float) * 4 * c->R, cudaMemcpyDeviceToHost, c->stream));
    if (out_sigma) CU(c, cudaMemcpyAsync(out_sigma, c->d_sigma, sizeof(float) * c->B, cudaMemcpyDeviceToHost, c->stream));
    if (out_rgba || out_sigma) CU(c, cudaStreamSynchronize(c->stream));
    return NERF_OK;
}

int ensure_copy_stream(nerf_ctx *c) {
    if (c->copy_stream) return NERF_OK;
    constexpr int kChunks = 8;
    CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&c->ev_dirs, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&c->ev_all, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&c->ev_cbwd, cudaEventDisableTiming));
    CU(c, guard_malloc(&c->d_h2d_flag, 2 * sizeof(unsigned int)));
    CU(c, cudaMemsetAsync(c->d_h2d_flag, 0, 2 * sizeof(unsigned int), c->stream));
    CU(c, cudaMallocHost(&c->h_h2d_seq, sizeof(unsigned int) * kChunks));
    for (int k = 0; k < kChunks; ++k) c->h_h2d_seq[k] = (unsigned int)k + 1u;
    return NERF_OK;
}

// pixels of the current batch on demand (a prediction whose compositing was deferred)
int ensure_outputs(nerf_ctx *c) {
    if (c->outputs_valid) return NERF_OK;
    if (!c->batch_valid || !c->predicted || c->fwd_deferred) return fail(c, NERF_ERR_STATE, "no prediction on this batch (call nerf_predict first)");
    int rc = composite_forward(c, c->R, c->d_out);
    if (rc) return rc;
    c->outputs_valid = true;
    return NERF_OK;
}

int do_step(nerf_ctx *c, const float *gold, int64_t n_gold, float *loss) {
    if (!c->predicted) return fail(c, NERF_ERR_STATE, "step: predict has not run on this batch");
    if (gold) {
        if (n_gold != (int64_t)c->R * 4) return fail(c, NERF_ERR_INVALID_ARG, "step: gold must hold num_rays*4 floats (model.rs:316)");
        CU(c, cudaMemcpyAsync(c->d_gold, gold, sizeof(float) * 4 * c->R, cudaMemcpyHostToDevice, c->stream));
    }
    const int nranks = c->comm.nranks;
    // compositing backward (+ pixels, + fused MSE gradient and loss) of rays [r0, r0+nr); the mean loss over ALL rays is
    // reduced by the launch that covers the last rays
    int loss_partials_done = 0;   // per-block loss partials written so far this step (d_loss_partials holds up to 2R + 16)
    auto composite_backward = [&](int r0, int nr) -> int {
        CompositeArgs a;
        memset(&a, 0, sizeof(a));
        const int64_t s0 = (int64_t)r0 * c->S;
        a.sigma = c->d_sigma + s0;
        a.colors = c->cfg.use_rgb_head ? c->d_rgba + 4 * s0 : nullptr;
        a.t_or_delta = c->d_t + s0;
        a.sigma_relu = c->cfg.sigma_relu;
        a.num_rays = nr;
        a.num_samples = c->S;
        a.gold = c->d_gold + 4 * (int64_t)r0;
        a.inv_count = 1.f / (4.f * (float)c->R);   // mean over R*4 elements (model.rs:298)
        a.loss_partials = c->d_loss_partials + loss_partials_done;
        a.loss_partials_first = c->d_loss_partials;
        a.loss_partials_prior = loss_partials_done;
        a.d_sigma = c->d_dsigma + s0;
        a.d_colors = c->d_drgba + 4 * s0;
        a.out = c->d_out + 4 * (int64_t)r0;            // the backward pass recomputes the pixels on its way
        if (r0 + nr >= c->R) {                         // ... and the mean loss (model.rs:298), reduced by its last block
            a.loss_out = c->d_loss;
            a.loss_scale = 1.f / (4.f * (float)c->R);
            a.done_counter = reinterpret_cast<unsigned int *>(c->d_loss + 2);
        }
        Scope s(c, "composite_bwd");
        loss_partials_done += launch_composite_bwd(a, c->num_sms, c->stream);
        return check_launch(c, "composite_bwd");
    };
    // gradients accumulate (+=) over micro-batches and weight-gradient CTAs: start from zero.
    // With the peer-memory all-reduce the local gradient goes to one of two buffers the other ranks read directly; a
    // buffer is reused two steps later, after every peer has passed the next step's hand-shake.
    const bool p2p = nranks > 1 && c->comm.p2p;
    float *gacc = p2p ? c->d_gacc[(c->comm.p2p_step + 1) & 1] : c->d_grads;
    CU(c, cudaMemsetAsync(gacc, 0, sizeof(float) * c->g.n_params, c->stream));   // (before the kernels, so that they stay back to back for the programmatic launches)
    int rc = NERF_OK;
    bool cbwd_recorded = false;
    bool fwd_in_step = false;   // a forward ran inside this step (micro-batches, or a step after an inference predict): it reads the batch inputs
    if (!c->fwd_deferred) {
        rc = composite_backward(0, c->R);
        if (rc) return rc;
        if (loss || c->copy_stream) {   // (the fused device-resident iteration with no loss read needs neither)
            rc = ensure_copy_stream(c);
            if (rc) return rc;
            CU(c, cudaEventRecord(c->ev_cbwd, c->stream));
            cbwd_recorded = true;
        }
    }
    for (int r0 = 0; r0 < c->R; r0 += c->chunk) {
        const int nr = (c->R - r0 < c->chunk) ? c->R - r0 : c->chunk;
        if (!c->acts_valid) {
            rc = mlp_forward(c, r0, nr, 1);  // (re)compute this micro-batch's activations
            if (rc) return rc;
            fwd_in_step = true;
        }
        if (c->fwd_deferred) {
            rc = composite_backward(r0, nr);
            if (rc) return rc;
        }
        rc = mlp_backward(c, r0, nr, gacc);
        if (rc) return rc;
    }
    c->acts_valid = false;
    c->predicted = false;
    c->fwd_deferred = false;
    c->outputs_valid = true;   // the compositing backward wrote the pixels, every micro-batch's forward the densities
    if (nranks > 1 && !p2p) {
        Scope s(c, "grad_allreduce");
        char eb[256] = {0};
        if (comm_allreduce_sum_f32(c->comm, c->d_grads, c->g.n_params, c->stream, eb, sizeof(eb))) return fail(c, NERF_ERR_COMM, eb);
    }
    c->step += 1;
    {
        AdamArgs a;
        a.p = c->d_params; a.m = c->d_m; a.v = c->d_v; a.g = c->d_grads;
        a.n = c->g.n_params;
        const double b1 = c->cfg.beta1, b2 = c->cfg.beta2;
        const double bc1 = 1.0 - pow(b1, (double)c->step), bc2 = 1.0 - pow(b2, (double)c->step);
        a.lr_over_bc1 = (float)((double)c->cfg.learning_rate / bc1);
        a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
        a.beta1 = c->cfg.beta1; a.beta2 = c->cfg.beta2; a.eps = c->cfg.eps;
        a.grad_scale = 1.f / (float)nranks;
        a.zero_grad = 0;
        if (p2p) {
            // ONE kernel: cross-GPU hand-shake, sum of all ranks' gradients over NVLink peer loads (rank order), Adam
            AdamP2PArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.adam = a;
            const unsigned int pstep = ++c->comm.p2p_step;
            for (int r = 0; r < nranks; ++r) {
                pa.peer_grads[r] = c->comm.peer_grads[pstep & 1][r];
                pa.peer_flags[r] = c->comm.peer_flags[r];
            }
            pa.my_flags = c->comm.my_flags;
            pa.block_counter = c->comm.my_flags + 2 * NERF_MAX_RANKS;
            pa.n_pad = (c->g.n_params + 3) / 4 * 4;
            pa.rank = c->comm.rank;
            pa.nranks = nranks;
            pa.step = pstep;
            Scope s(c, "adam_allreduce_p2p");
            launch_adam_p2p(pa, c->num_sms, c->stream);
        } else {
            Scope s(c, "adam");
            launch_adam(a, c->num_sms, c->stream);
        }
    }
    c->weights_dirty = true;
    rc = ensure_packed(c);
    if (rc) return rc;
    if (loss) {
        // The loss is final once the compositing backward has run -- long before dgrad, wgrad and Adam finish. It is read on the
        // copy stream behind that kernel's event, so the call returns (like f32::try_from(&loss) on an asynchronous backend,
        // model.rs:322-324) while the rest of the step is still running and the host can already enqueue the next batch's copies.
        static const bool sync_step = getenv("NERF_B200_STEP_SYNC") != nullptr;   // A/B: wait for the whole step
        if (cbwd_recorded && !sync_step) {
            CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_cbwd, 0));
            CU(c, cudaMemcpyAsync(c->h_loss, c->d_loss, sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
            CU(c, cudaStreamSynchronize(c->copy_stream));
        } else {
            CU(c, cudaMemcpyAsync(c->h_loss, c->d_loss, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
        }
        float l = *c->h_loss;
        *loss = l;
    }
    rc = check_launch(c, "step");
    // (a step that re-runs the forward -- which reads the batch inputs -- after the compositing backward does not release them)
    c->inputs_free = rc == NERF_OK && cbwd_recorded && !fwd_in_step && c->chunk >= c->R;
    return rc;
}

int run_sampler(nerf_ctx *c, int nr, const int32_t *view_pick, int rays_per_pick, const ViewPose *poses, int fixed_view,
                const float *jitter, int randomize, uint64_t seed, int64_t ray_base, bool gather_gold, bool write_points) {
    // the CTA-pair MLP kernel regenerates sample positions in its prologue: points only go to HBM when somebody reads them
    if (!c->tc) write_points = true;
    c->points_valid = write_points;
    c->batch_poses = poses;
    SampleArgs a;
    memset(&a, 0, sizeof(a));
    a.pix_yx = c->d_pix;
    a.view_pick = view_pick;
    a.rays_per_pick = rays_per_pick;
    a.fixed_view = fixed_view;
    a.gen_pix = view_pick ? c->gen_pix : 0;        // (render passes explicit pixels and a fixed view)
    a.gen_view = view_pick ? c->gen_view : 0;
    a.n_views = c->pick_views;
    a.pix_out = c->d_pix;
    a.view_out = c->d_view_pick;
    a.poses = poses;
    a.jitter = jitter;
    a.images = gather_gold ? c->d_images : nullptr;
    a.images_u8 = gather_gold ? c->d_images_u8 : nullptr;
    a.num_rays = nr;
    a.num_samples = c->S;
    a.img_w = c->cfg.image_w;
    a.img_h = c->cfg.image_h;
    a.randomize = randomize;
    a.depth_mode = c->cfg.depth_mode;
    a.off = c->off;
    a.seed = seed;
    a.ray_index_base = ray_base;
    a.rays = c->d_rays;
    a.dirs = c->d_dirs;
    a.t = c->d_t;
    a.points = write_points ? c->d_points : nullptr;
    a.gold = c->d_gold;
    // A fused iteration that follows another one: everything the sampler writes (picks, ray records, depths, gold) had its
    // last reader in the previous step's compositing backward, and nothing it reads depends on the weights -- it runs on the
    // copy stream behind that kernel's event, under the previous step's dgrad / weight gradients / gradient exchange, instead
    // of between Adam and the forward. (Not while profiling: the per-kernel events are recorded on the main stream.)
    const bool side = c->sample_side && !c->prof.on && c->copy_stream;
    if (side) CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_cbwd, 0));
    {
        Scope s(c, "sample");
        launch_sample(a, c->num_sms, side ? c->copy_stream : c->stream);
    }
    if (side) {
        CU(c, cudaEventRecord(c->ev_all, c->copy_stream));
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_all, 0));
    }
    return check_launch(c, "sample");
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int nerf_abi_version(void) { return NERF_B200_ABI_VERSION; }

int nerf_default_config(nerf_config *cfg) {
    if (!cfg) return NERF_ERR_INVALID_ARG;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (int32_t)sizeof(nerf_config);
    cfg->image_w = 800; cfg->image_h = 800;
    cfg->num_rays = 4096; cfg->num_samples = 64;
    cfg->hidden = 256; cfg->xyz_freqs = 10; cfg->dir_freqs = 4; cfg->skip_layer = 5;
    cfg->use_rgb_head = 1; cfg->sigma_relu = 0;
    cfg->depth_mode = NERF_DEPTH_REFERENCE;
    cfg->mlp_impl = NERF_MLP_TCGEN05;
    cfg->max_rays_per_launch = 0;
    cfg->learning_rate = 5e-4f; cfg->beta1 = 0.9f; cfg->beta2 = 0.999f; cfg->eps = 1e-8f;
    cfg->deterministic_grads = 0;
    return NERF_OK;
}

int nerf_config_as_shipped(nerf_config *cfg) {
    int rc = nerf_default_config(cfg);
    if (rc) return rc;
    cfg->image_w = 128; cfg->image_h = 128;       // ray_sampling.rs:7-8
    cfg->num_rays = 84; cfg->num_samples = 64;    // model.rs:7-8
    cfg->hidden = 100;                             // model.rs:12
    cfg->xyz_freqs = 0; cfg->dir_freqs = -1; cfg->skip_layer = 0;
    cfg->use_rgb_head = 0;                         // model.rs:190-206
    return NERF_OK;
}

const char *nerf_strerror(int status) {
    switch (status) {
        case NERF_OK: return "ok";
        case NERF_ERR_INVALID_ARG: return "invalid argument";
        case NERF_ERR_CUDA: return "CUDA error";
        case NERF_ERR_UNSUPPORTED: return "unsupported configuration";
        case NERF_ERR_COMM: return "communication error";
        case NERF_ERR_STATE: return "invalid call order";
        case NERF_ERR_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
        default: return "unknown status";
    }
}

const char *nerf_last_error(const nerf_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int nerf_destroy(nerf_ctx *c) {
    if (!c) return NERF_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    comm_destroy(c->comm);
    tc_destroy(c->tc);
    void *ptrs[] = {c->d_params, c->d_grads, c->d_m, c->d_v, c->d_images, c->d_poses, c->d_render_pose, c->d_pix, c->d_view_pick,
                    c->d_rays, c->d_dirs, c->d_t, c->d_points, c->d_gold, c->d_jitter, c->d_sigma, c->d_rgba, c->d_out,
                    c->d_dsigma, c->d_drgba, c->d_loss_partials, c->d_loss, c->simt.x_enc, c->simt.d_enc, c->simt.act, c->simt.dact,
                    c->d_flush, c->d_frame_rgba, c->d_frame_0rgb, c->d_gacc[0], c->d_gacc[1], c->d_images_u8, c->d_metrics};
    for (void *p : ptrs) guard_free(p);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
        cudaEventDestroy(c->ev_main); cudaEventDestroy(c->ev_dirs); cudaEventDestroy(c->ev_all); cudaEventDestroy(c->ev_cbwd);
        guard_free(c->d_h2d_flag);
        cudaFreeHost(c->h_h2d_seq);
    }
    if (c->h_loss) cudaFreeHost(c->h_loss);
    if (c->h_i32) cudaFreeHost(c->h_i32);
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    if (c->ev_stage) cudaEventDestroy(c->ev_stage);
    for (auto e : c->prof.pool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return NERF_OK;
}

int nerf_create(const nerf_config *cfg, int device, nerf_ctx **out) {
    if (!cfg || !out) return NERF_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(nerf_config)) return NERF_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return NERF_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NERF_ERR_NO_DEVICE;
    if (prop.major != 10) return NERF_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    if (cfg->num_rays < 1 || cfg->num_samples < 1 || cfg->num_samples > 256 || cfg->hidden < 2 || cfg->hidden > 512 ||
        cfg->xyz_freqs < 0 || cfg->xyz_freqs > 10 || cfg->dir_freqs > 4 || cfg->image_w < 1 || cfg->image_h < 1 ||
        cfg->mlp_impl < 0 || cfg->mlp_impl > 3 || cfg->depth_mode < 0 || cfg->depth_mode > 1)
        return NERF_ERR_UNSUPPORTED;
    nerf_ctx *c = new nerf_ctx();
    c->cfg = *cfg;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    build_geom(c->cfg, c->g);
    c->R = cfg->num_rays;
    c->S = cfg->num_samples;
    c->B = (int64_t)c->R * c->S;
    int64_t def_chunk = (int64_t)(1 << 21) / c->S;
    if (def_chunk < 1) def_chunk = 1;
    int64_t chunk = cfg->max_rays_per_launch > 0 ? cfg->max_rays_per_launch : def_chunk;
    if (chunk > c->R) chunk = c->R;
    c->chunk = (int)chunk;
    c->off = screen_offset();

    auto bail = [&](int code, const std::string &msg) {
        fprintf(stderr, "nerf_create: %s\n", msg.c_str());
        nerf_destroy(c);
        return code;
    };
#define CUB(expr)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (expr);                                                                               \
        if (e_ != cudaSuccess) return bail(NERF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)
    CUB(cudaSetDevice(device));
    CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUB(cudaEventCreate(&c->t0));
    CUB(cudaEventCreate(&c->t1));
    CUB(cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming));
    const int64_t P = c->g.n_params, Pp = (P + 3) / 4 * 4;
    CUB(guard_malloc(&c->d_params, sizeof(float) * Pp));
    CUB(guard_malloc(&c->d_grads, sizeof(float) * Pp));
    CUB(guard_malloc(&c->d_m, sizeof(float) * Pp));
    CUB(guard_malloc(&c->d_v, sizeof(float) * Pp));
    CUB(cudaMemsetAsync(c->d_params, 0, sizeof(float) * Pp, c->stream));
    CUB(cudaMemsetAsync(c->d_grads, 0, sizeof(float) * Pp, c->stream));
    CUB(cudaMemsetAsync(c->d_m, 0, sizeof(float) * Pp, c->stream));
    CUB(cudaMemsetAsync(c->d_v, 0, sizeof(float) * Pp, c->stream));
    const int64_t R = c->R, B = c->B;
    CUB(guard_malloc(&c->d_pix, sizeof(int32_t) * 2 * R));
    CUB(guard_malloc(&c->d_view_pick, sizeof(int32_t) * R));
    CUB(guard_malloc(&c->d_rays, sizeof(RayRec) * R));
    CUB(guard_malloc(&c->d_dirs, sizeof(float) * 3 * R));
    CUB(guard_malloc(&c->d_t, sizeof(float) * B));
    CUB(guard_malloc(&c->d_points, sizeof(float) * 3 * B));
    CUB(guard_malloc(&c->d_gold, sizeof(float) * 4 * R));
    CUB(guard_malloc(&c->d_jitter, sizeof(float) * B));
    CUB(guard_malloc(&c->d_sigma, sizeof(float) * B));
    CUB(guard_malloc(&c->d_rgba, sizeof(float) * 4 * B));
    CUB(guard_malloc(&c->d_out, sizeof(float) * 4 * R));
    CUB(guard_malloc(&c->d_dsigma, sizeof(float) * B));
    CUB(guard_malloc(&c->d_drgba, sizeof(float) * 4 * B));
    CUB(guard_malloc(&c->d_loss_partials, sizeof(float) * (2 * (size_t)R + 16)));   // per-block loss partials: <= ceil(nr/8) per launch
    CUB(guard_malloc(&c->d_loss, sizeof(float) * 4));
    CUB(guard_malloc(&c->d_render_pose, sizeof(ViewPose)));
    CUB(cudaMemsetAsync(c->d_gold, 0, sizeof(float) * 4 * R, c->stream));
    CUB(cudaMemsetAsync(c->d_rgba, 0, sizeof(float) * 4 * B, c->stream));
    CUB(cudaMemsetAsync(c->d_drgba, 0, sizeof(float) * 4 * B, c->stream));
    CUB(cudaMemsetAsync(c->d_loss, 0, sizeof(float) * 4, c->stream));
    CUB(cudaMallocHost(&c->h_loss, sizeof(float) * 4));
    if (is_tc(cfg->mlp_impl)) {
        std::string e;
        const int64_t max_tiles = ((int64_t)c->chunk * c->S + NERF_TILE_M - 1) / NERF_TILE_M;
        c->tc = tc_create(c->g, max_tiles, c->num_sms, cfg->mlp_impl == NERF_MLP_TCGEN05_SS ? 2 : 0, e, cfg->deterministic_grads != 0);
        if (!c->tc) return bail(NERF_ERR_UNSUPPORTED, e);
    } else {
        const int64_t bc = (int64_t)c->chunk * c->S;
        c->simt.act_stride = bc * (int64_t)simt_act_floats_per_sample(c->g);
        CUB(guard_malloc(&c->simt.x_enc, sizeof(float) * bc * c->g.Cx));
        CUB(guard_malloc(&c->simt.d_enc, sizeof(float) * (int64_t)c->chunk * (c->g.Cd > 0 ? c->g.Cd : 1)));
        CUB(guard_malloc(&c->simt.act, sizeof(float) * c->simt.act_stride * 9));
        CUB(guard_malloc(&c->simt.dact, sizeof(float) * c->simt.act_stride * 3));
        c->simt_ready = true;
    }
#undef CUB
    launch_init_uniform(c->d_params, c->g, 0x5eed5eedull, c->stream);
    c->launch_count += 10;
    c->weights_dirty = true;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return bail(NERF_ERR_CUDA, "initialisation kernels failed");
    *out = c;
    return NERF_OK;
}

int64_t nerf_num_params(const nerf_ctx *c) { return c ? c->g.n_params : 0; }
int64_t nerf_launch_count(const nerf_ctx *c) { return c ? c->launch_count : 0; }

int nerf_set_weights(nerf_ctx *c, const float *flat, int64_t n) {
    if (!c || !flat) return NERF_ERR_INVALID_ARG;
    if (n != c->g.n_params) return fail(c, NERF_ERR_INVALID_ARG, "set_weights: wrong parameter count");
    ENTER(c);
    CU(c, cudaMemcpyAsync(c->d_params, flat, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->weights_dirty = true;
    return NERF_OK;
}
static int get_blob(nerf_ctx *c, const float *src, float *dst, int64_t n) {
    if (!c || !dst) return NERF_ERR_INVALID_ARG;
    if (n != c->g.n_params) return fail(c, NERF_ERR_INVALID_ARG, "wrong parameter count");
    ENTER(c);
    CU(c, cudaMemcpyAsync(dst, src, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return NERF_OK;
}
int nerf_get_weights(nerf_ctx *c, float *flat, int64_t n) { return get_blob(c, c ? c->d_params : nullptr, flat, n); }
int nerf_get_grads(nerf_ctx *c, float *flat, int64_t n) { return get_blob(c, c ? c->d_grads : nullptr, flat, n); }
int nerf_get_adam_state(nerf_ctx *c, float *m, float *v, int64_t n, int64_t *step) {
    int rc = get_blob(c, c ? c->d_m : nullptr, m, n);
    if (rc) return rc;
    rc = get_blob(c, c->d_v, v, n);
    if (rc) return rc;
    if (step) *step = c->step;
    return NERF_OK;
}
int nerf_set_adam_state(nerf_ctx *c, const float *m, const float *v, int64_t n, int64_t step) {
    if (!c || !m || !v || step < 0) return NERF_ERR_INVALID_ARG;
    if (n != c->g.n_params) return fail(c, NERF_ERR_INVALID_ARG, "set_adam_state: wrong parameter count");
    ENTER(c);
    CU(c, cudaMemcpyAsync(c->d_m, m, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->d_v, v, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->step = step;
    return NERF_OK;
}

int nerf_set_images(nerf_ctx *c, const float *rgba, int32_t n_views) {
    if (!c || !rgba || n_views < 1) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    const size_t bytes = sizeof(float) * 4 * (size_t)n_views * c->cfg.image_w * c->cfg.image_h;
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->d_images) { guard_free(c->d_images); c->d_images = nullptr; }
    if (c->d_images_u8) { guard_free(c->d_images_u8); c->d_images_u8 = nullptr; }
    CU(c, guard_malloc(&c->d_images, bytes));
    CU(c, cudaMemcpyAsync(c->d_images, rgba, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->n_img_views = n_views;
    return NERF_OK;
}

int nerf_set_images_rgba8(nerf_ctx *c, const uint8_t *rgba8, int32_t n_views) {
    if (!c || !rgba8 || n_views < 1) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    const size_t bytes = 4 * (size_t)n_views * c->cfg.image_w * c->cfg.image_h;
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->d_images) { guard_free(c->d_images); c->d_images = nullptr; }
    if (c->d_images_u8) { guard_free(c->d_images_u8); c->d_images_u8 = nullptr; }
    CU(c, guard_malloc(&c->d_images_u8, bytes));
    CU(c, cudaMemcpyAsync(c->d_images_u8, rgba8, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->n_img_views = n_views;
    return NERF_OK;
}

int nerf_set_view_angles(nerf_ctx *c, const float *yaw_pitch, int32_t n) {
    if (!c || !yaw_pitch || n < 1) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    std::vector<ViewPose> poses((size_t)n);
    for (int i = 0; i < n; ++i) make_pose(yaw_pitch[2 * i], yaw_pitch[2 * i + 1], poses[i]);
    if (c->d_poses) { CU(c, cudaStreamSynchronize(c->stream)); guard_free(c->d_poses); c->d_poses = nullptr; }
    CU(c, guard_malloc(&c->d_poses, sizeof(ViewPose) * (size_t)n));
    CU(c, cudaMemcpyAsync(c->d_poses, poses.data(), sizeof(ViewPose) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->n_poses = n;
    return NERF_OK;
}

int nerf_view_angles_grid(int32_t nv, float *out, int32_t capacity) {
    // image_loading.rs:67-80: yaw advances once per outer step, pitch per inner step, both by
    // repeated f32 addition of pi/n; 2n(n+1) pairs.
    if (nv < 1 || !out || capacity < 2 * nv * (nv + 1)) return NERF_ERR_INVALID_ARG;
    const float pi = 3.14159265358979323846f;
    float rot_ver = 0.f, rot_hor = 0.f;
    int k = 0;
    for (int i = 0; i < 2 * nv; ++i) {
        for (int j = 0; j < nv + 1; ++j) {
            out[2 * k] = rot_hor;
            out[2 * k + 1] = rot_ver;
            ++k;
            rot_ver = rot_ver + pi / (float)nv;
        }
        rot_hor = rot_hor + pi / (float)nv;
        rot_ver = 0.f;
    }
    return NERF_OK;
}

int nerf_get_batch(nerf_ctx *c, const int64_t *indices_yx, const int64_t *view_index, int32_t n_picks, const float *jitter,
                   int32_t randomize, uint64_t seed, float *out_points, float *out_t, float *out_gold, float *out_dirs,
                   int64_t *out_indices) {
    if (!c) return NERF_ERR_INVALID_ARG;
    // Directly after a single-launch nerf_step every buffer this call writes had its last reader in that step's compositing
    // backward: the host inputs, the sampler and the read-back then go through the copy stream behind that kernel's event
    // instead of queueing behind dgrad, weight gradients and Adam on the main stream (nerf_train_iter asks for the same).
    static const bool no_side = getenv("NERF_B200_NO_SAMPLER_OVERLAP") != nullptr;   // A/B
    const bool side = (c->sample_side || (c->inputs_free && !no_side)) && c->copy_stream && !c->prof.on;
    ENTER(c);
    if (!c->d_poses) return fail(c, NERF_ERR_STATE, "get_batch: call nerf_set_view_angles first");
    cudaStream_t st = side ? c->copy_stream : c->stream;
    if (side) CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_cbwd, 0));
    const int R = c->R, S = c->S;
    if (n_picks < 1 || n_picks > R || R % n_picks != 0)
        return fail(c, NERF_ERR_INVALID_ARG, "get_batch: can't divide rays evenly among views (dataset.rs:73-81)");
    const int n_views = ((c->d_images || c->d_images_u8) && c->n_img_views < c->n_poses) ? c->n_img_views : c->n_poses;
    int rc;
    if ((indices_yx || view_index) && c->stage_busy) {
        // the previous call's copies out of the pinned staging buffer are asynchronous: a caller that never reads anything
        // back can be several batches ahead of the stream, so wait for them before the buffer is rewritten
        CU(c, cudaEventSynchronize(c->ev_stage));
        c->stage_busy = false;
    }
    if (indices_yx) {
        rc = ensure_i32(c, (size_t)2 * R + n_picks);
        if (rc) return rc;
        for (int i = 0; i < R; ++i) {
            const int64_t y = indices_yx[2 * i], x = indices_yx[2 * i + 1];
            if (y < 0 || y >= c->cfg.image_h || x < 0 || x >= c->cfg.image_w) return fail(c, NERF_ERR_INVALID_ARG, "get_batch: pixel index out of range");
            c->h_i32[2 * i] = (int32_t)y;
            c->h_i32[2 * i + 1] = (int32_t)x;
        }
        CU(c, cudaMemcpyAsync(c->d_pix, c->h_i32, sizeof(int32_t) * 2 * R, cudaMemcpyHostToDevice, st));
    }
    if (view_index) {
        rc = ensure_i32(c, (size_t)2 * R + n_picks);
        if (rc) return rc;
        for (int i = 0; i < n_picks; ++i) {
            if (view_index[i] < 0 || view_index[i] >= n_views) return fail(c, NERF_ERR_INVALID_ARG, "get_batch: view index out of range");
            c->h_i32[2 * R + i] = (int32_t)view_index[i];
        }
        CU(c, cudaMemcpyAsync(c->d_view_pick, c->h_i32 + 2 * R, sizeof(int32_t) * n_picks, cudaMemcpyHostToDevice, st));
    }
    if (indices_yx || view_index) {
        CU(c, cudaEventRecord(c->ev_stage, st));
        c->stage_busy = true;
    }
    c->gen_pix = indices_yx ? 0 : 1;       // missing picks are drawn inside the sampler (Philox)
    c->gen_view = view_index ? 0 : 1;
    c->pick_views = n_views;
    const float *dj = nullptr;
    if (jitter && randomize) {
        CU(c, cudaMemcpyAsync(c->d_jitter, jitter, sizeof(float) * (size_t)R * S, cudaMemcpyHostToDevice, st));
        dj = c->d_jitter;
    }
    const int64_t ray_base = (int64_t)c->comm.rank * R;
    const bool side_was = c->sample_side;
    c->sample_side = side;
    rc = run_sampler(c, R, c->d_view_pick, R / n_picks, c->d_poses, 0, dj, randomize, seed, ray_base, c->d_images != nullptr || c->d_images_u8 != nullptr,
                     out_points != nullptr);
    c->sample_side = side_was;
    if (rc) return rc;
    c->batch_valid = true;
    c->predicted = false;
    c->outputs_valid = false;
    c->acts_valid = false;
    bool sync = false;
    if (out_points) { CU(c, cudaMemcpyAsync(out_points, c->d_points, sizeof(float) * 3 * c->B, cudaMemcpyDeviceToHost, st)); sync = true; }
    if (out_t) { CU(c, cudaMemcpyAsync(out_t, c->d_t, sizeof(float) * c->B, cudaMemcpyDeviceToHost, st)); sync = true; }
    if (out_gold) { CU(c, cudaMemcpyAsync(out_gold, c->d_gold, sizeof(float) * 4 * R, cudaMemcpyDeviceToHost, st)); sync = true; }
    if (out_dirs) { CU(c, cudaMemcpyAsync(out_dirs, c->d_dirs, sizeof(float) * 3 * R, cudaMemcpyDeviceToHost, st)); sync = true; }
    if (out_indices) {
        rc = ensure_i32(c, (size_t)2 * R + n_picks);
        if (rc) return rc;
        CU(c, cudaMemcpyAsync(c->h_i32, c->d_pix, sizeof(int32_t) * 2 * R, cudaMemcpyDeviceToHost, st));
        sync = true;
    }
    if (sync) CU(c, cudaStreamSynchronize(st));
    if (out_indices) for (int i = 0; i < 2 * R; ++i) out_indices[i] = c->h_i32[i];
    return check_launch(c, "get_batch");
}

int nerf_predict(nerf_ctx *c, int32_t train, float *out_rgba, float *out_sigma) {
    if (!c) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    return do_predict(c, train, out_rgba, out_sigma);
}

int nerf_get_predictions(nerf_ctx *c, float *out_rgba, float *out_sigma) {
    if (!c || (!out_rgba && !out_sigma)) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    int rc = ensure_outputs(c);
    if (rc) return rc;
    if (out_rgba) CU(c, cudaMemcpyAsync(out_rgba, c->d_out, sizeof(float) * 4 * c->R, cudaMemcpyDeviceToHost, c->stream));
    if (out_sigma) CU(c, cudaMemcpyAsync(out_sigma, c->d_sigma, sizeof(float) * c->B, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return NERF_OK;
}

int nerf_predict_points(nerf_ctx *c, const float *query_points, int64_t n_points_floats, const float *distances,
                        int64_t n_distances, const float *dirs, int32_t train, float *out_rgba, float *out_sigma) {
    if (!c || !query_points || !distances) return NERF_ERR_INVALID_ARG;
    const bool pipelined = c->inputs_free;   // (before ENTER clears it)
    ENTER(c);
    // assert_eq!(query_points.size(), [BATCH_SIZE*INDIM]); assert_eq!(distances.size(), [BATCH_SIZE]) (model.rs:162-163)
    if (n_points_floats != c->B * 3) return fail(c, NERF_ERR_INVALID_ARG, "predict: query_points must hold num_rays*num_samples*3 floats (model.rs:162)");
    if (n_distances != c->B) return fail(c, NERF_ERR_INVALID_ARG, "predict: distances must hold num_rays*num_samples floats (model.rs:163)");
    if (c->g.Cd && !dirs) return fail(c, NERF_ERR_INVALID_ARG, "predict: this configuration needs ray directions [num_rays*3]");
    c->points_valid = true;
    c->batch_valid = true;
    c->predicted = false;
    c->outputs_valid = false;
    // The CTA-pair kernel can start on the first tiles while the rest of the points are still crossing PCIe: the copy runs in
    // kChunks pieces on its own stream, each followed by a 4-byte counter update the kernel's prologue polls. (One launch only:
    // micro-batched batches, the v1 kernel and the SIMT cross-check take the plain copy-then-run path.)
    constexpr int kChunks = 8;
    const bool overlap = c->tc && c->chunk >= c->R && c->R >= 4 * kChunks && !getenv("NERF_B200_NO_H2D_OVERLAP");
    if (!overlap) {
        CU(c, cudaMemcpyAsync(c->d_points, query_points, sizeof(float) * 3 * c->B, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_t, distances, sizeof(float) * c->B, cudaMemcpyHostToDevice, c->stream));
        if (dirs) CU(c, cudaMemcpyAsync(c->d_dirs, dirs, sizeof(float) * 3 * c->R, cudaMemcpyHostToDevice, c->stream));
        return do_predict(c, train, out_rgba, out_sigma);
    }
    {
        int rc0 = ensure_copy_stream(c);
        if (rc0) return rc0;
    }
    // What the copies must wait for: the last reader of the batch inputs. Right after a single-launch nerf_step (whose loss read
    // returned early) that is its compositing backward (ev_cbwd; dgrad, wgrad and Adam touch neither points, depths nor
    // directions), so the next batch crosses PCIe under the rest of the previous step; otherwise everything enqueued so far.
    if (pipelined) {
        // The GPU still has the previous step's dgrad, weight gradients and Adam to run (~0.7 ms at the bench shape): the whole
        // batch arrives before the forward kernel's turn comes, so it is three plain copies and the ordinary (non-polling) kernel.
        CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_cbwd, 0));
        if (dirs) CU(c, cudaMemcpyAsync(c->d_dirs, dirs, sizeof(float) * 3 * c->R, cudaMemcpyHostToDevice, c->copy_stream));
        CU(c, cudaMemcpyAsync(c->d_points, query_points, sizeof(float) * 3 * c->B, cudaMemcpyHostToDevice, c->copy_stream));
        CU(c, cudaMemcpyAsync(c->d_t, distances, sizeof(float) * c->B, cudaMemcpyHostToDevice, c->copy_stream));
        CU(c, cudaEventRecord(c->ev_all, c->copy_stream));
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_all, 0));
        return do_predict(c, train, out_rgba, out_sigma);
    }
    unsigned int *const flag = c->d_h2d_flag + c->h2d_word;   // zero: reset in stream order after its previous use (below)
    CU(c, cudaEventRecord(c->ev_main, c->stream));
    CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
    if (dirs) CU(c, cudaMemcpyAsync(c->d_dirs, dirs, sizeof(float) * 3 * c->R, cudaMemcpyHostToDevice, c->copy_stream));
    CU(c, cudaEventRecord(c->ev_dirs, c->copy_stream));
    const int rays_per_chunk = (c->R + kChunks - 1) / kChunks;
    const int64_t chunk_samples = (int64_t)rays_per_chunk * c->S;
    // The kernel is launched BEFORE the point copies are enqueued: its prologue waits for the counter anyway, and the ~20 driver
    // calls below take the host longer (~50 us) than the first chunks take to cross PCIe -- launched last, the kernel started
    // after half of the points had already arrived.
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_dirs, 0));
    c->h2d_flag_active = flag;
    c->h2d_chunk_samples = chunk_samples;
    int rc = ensure_packed(c);
    if (rc == NERF_OK) rc = mlp_forward(c, 0, c->R, train ? 1 : 0);
    c->h2d_flag_active = nullptr;
    // the OTHER counter -- last used by the previous call's kernel, which precedes this point in the stream -- is reset for the
    // next call
    c->h2d_word ^= 1;
    cudaError_t e = cudaMemsetAsync(c->d_h2d_flag + c->h2d_word, 0, sizeof(unsigned int), c->stream);
    // (the copies are enqueued on the error path too: the counter must reach kChunks for a kernel that did get launched)
    for (int k = 0; k < kChunks && e == cudaSuccess; ++k) {
        const int64_t s0 = (int64_t)k * chunk_samples;
        const int64_t ns = (c->B - s0 < chunk_samples) ? c->B - s0 : chunk_samples;
        if (ns > 0) e = cudaMemcpyAsync(c->d_points + 3 * s0, query_points + 3 * s0, sizeof(float) * 3 * ns, cudaMemcpyHostToDevice, c->copy_stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(flag, c->h_h2d_seq + k, sizeof(unsigned int), cudaMemcpyHostToDevice, c->copy_stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_t, distances, sizeof(float) * c->B, cudaMemcpyHostToDevice, c->copy_stream);   // compositing only
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_all, c->copy_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_all, 0);   // (later work must not race the copies)
    if (rc) return rc;
    if (e != cudaSuccess) return fail(c, NERF_ERR_CUDA, cudaGetErrorString(e));
    return do_predict(c, train, out_rgba, out_sigma, false, /*forward_done=*/true);
}

int nerf_compositing(nerf_ctx *c, const float *densities, const float *colors, const float *distances, int32_t num_rays,
                     int32_t num_samples, float *out) {
    if (!c || !densities || !distances || !out) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    if (num_rays < 1 || num_samples < 1 || num_samples > 256) return fail(c, NERF_ERR_INVALID_ARG, "compositing: bad shape");
    const size_t n = (size_t)num_rays * num_samples;
    float *ds = nullptr, *dc = nullptr, *dd = nullptr, *dout = nullptr;
    CU(c, guard_malloc(&ds, sizeof(float) * n));
    CU(c, guard_malloc(&dd, sizeof(float) * n));
    CU(c, guard_malloc(&dout, sizeof(float) * 4 * num_rays));
    if (colors) CU(c, guard_malloc(&dc, sizeof(float) * 4 * n));
    CU(c, cudaMemcpyAsync(ds, densities, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(dd, distances, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    if (colors) CU(c, cudaMemcpyAsync(dc, colors, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, c->stream));
    CompositeArgs a;
    memset(&a, 0, sizeof(a));
    a.sigma = ds; a.colors = dc; a.t_or_delta = dd; a.input_is_delta = 1; a.sigma_relu = 0;
    a.num_rays = num_rays; a.num_samples = num_samples; a.out = dout;
    {
        Scope s(c, "composite_fwd");
        launch_composite_fwd(a, c->num_sms, c->stream);
    }
    CU(c, cudaMemcpyAsync(out, dout, sizeof(float) * 4 * num_rays, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    guard_free(ds); guard_free(dd); guard_free(dout); guard_free(dc);
    return check_launch(c, "compositing");
}

int nerf_step(nerf_ctx *c, const float *gold, int64_t n_gold, float *loss) {
    if (!c) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    return do_step(c, gold, n_gold, loss);
}

int nerf_train_iter(nerf_ctx *c, uint64_t seed) {
    if (!c) return NERF_ERR_INVALID_ARG;
    static const bool no_side = getenv("NERF_B200_NO_SAMPLER_OVERLAP") != nullptr;   // A/B
    const bool after_step = c->inputs_free && !no_side;   // (before ENTER clears it) the previous call was a single-launch step
    ENTER(c);
    if (!c->d_images && !c->d_images_u8) return fail(c, NERF_ERR_STATE, "train_iter: call nerf_set_images first");
    int n_picks = c->n_img_views < c->n_poses ? c->n_img_views : c->n_poses;
    while (n_picks > 1 && c->R % n_picks != 0) --n_picks;   // largest pick count that splits R evenly
    int rc = ensure_copy_stream(c);   // (so that every step records ev_cbwd)
    if (rc) return rc;
    c->sample_side = after_step;
    rc = nerf_get_batch(c, nullptr, nullptr, n_picks, nullptr, 1, seed, nullptr, nullptr, nullptr, nullptr, nullptr);
    c->sample_side = false;
    if (rc) return rc;
    rc = do_predict(c, 1, nullptr, nullptr, /*skip_composite=*/true);
    if (rc) return rc;
    return do_step(c, nullptr, 0, nullptr);
}

int nerf_last_loss(nerf_ctx *c, float *loss) {
    if (!c || !loss) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    CU(c, cudaMemcpyAsync(c->h_loss, c->d_loss, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    *loss = *c->h_loss;
    return NERF_OK;
}

int nerf_sync(nerf_ctx *c) {
    if (!c) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->comm.p2p) {
        // the fused all-reduce + Adam kernel gives up on a peer that never publishes its gradient (instead of trapping or
        // hanging): from then on the replicas have diverged, which the host must hear about
        const unsigned int ts = adam_p2p_timeout_step();
        if (ts) return fail(c, NERF_ERR_COMM, "peer gradient exchange timed out at step " + std::to_string(ts) + ": the replicas have diverged");
    }
    return check_launch(c, "sync");
}

// Rows [y0, y1) of a frame at (yaw, pitch) into the context's device frame buffers (row y at offset y*W).
static int render_rows_device(nerf_ctx *c, float yaw, float pitch, int32_t y0, int32_t y1, int32_t randomize, uint64_t seed, bool pack) {
    const int W = c->cfg.image_w, H = c->cfg.image_h;
    if (!c->d_frame_rgba) {
        CU(c, guard_malloc(&c->d_frame_rgba, sizeof(float) * 4 * (size_t)W * H));
        CU(c, guard_malloc(&c->d_frame_0rgb, sizeof(uint32_t) * (size_t)W * H));
    }
    ViewPose vp;
    make_pose(yaw, pitch, vp);
    CU(c, cudaMemcpyAsync(c->d_render_pose, &vp, sizeof(vp), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));  // vp is a stack temporary
    const int64_t total = (int64_t)(y1 - y0) * W;
    c->batch_valid = false;
    c->predicted = false;
    c->outputs_valid = false;
    c->acts_valid = false;
    const int saveR = c->R;
    int rc = NERF_OK;
    for (int64_t p0 = 0; p0 < total && rc == NERF_OK; p0 += saveR) {
        const int nr = (int)((total - p0 < saveR) ? total - p0 : saveR);
        {   // pixel (y,x) list of this chunk, row-major (display.rs:58-62)
            Scope s(c, "frame_indices");
            launch_flat_pixel_indices(c->d_pix, (int64_t)y0 * W + p0, nr, W, c->stream);
        }
        rc = run_sampler(c, nr, nullptr, 1, c->d_render_pose, 0, nullptr, randomize, seed, (int64_t)y0 * W + p0, false, false);
        if (rc) break;
        c->R = nr;  // temporarily narrow the batch view for the engines
        for (int r0 = 0; r0 < nr && rc == NERF_OK; r0 += c->chunk) {
            const int n2 = (nr - r0 < c->chunk) ? nr - r0 : c->chunk;
            rc = mlp_forward(c, r0, n2, 0);
        }
        float *dst = c->d_frame_rgba + 4 * ((size_t)y0 * W + p0);
        if (rc == NERF_OK) rc = composite_forward(c, nr, dst);
        c->R = saveR;
        if (rc) break;
        if (pack) {
            Scope s(c, "pack_0rgb");
            launch_pack_0rgb(dst, c->d_frame_0rgb + (size_t)y0 * W + p0, nr, c->stream);
        }
    }
    if (rc) return rc;
    return check_launch(c, "render");
}

int nerf_render(nerf_ctx *c, float yaw, float pitch, int32_t y0, int32_t y1, int32_t randomize, uint64_t seed, float *out_rgba,
                uint32_t *out_0rgb) {
    if (!c) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    const int W = c->cfg.image_w, H = c->cfg.image_h;
    if (y0 < 0 || y1 > H || y0 >= y1) return fail(c, NERF_ERR_INVALID_ARG, "render: bad row range");
    int rc = render_rows_device(c, yaw, pitch, y0, y1, randomize, seed, out_0rgb != nullptr);
    if (rc) return rc;
    const size_t n = (size_t)(y1 - y0) * W, off = (size_t)y0 * W;
    if (out_rgba) CU(c, cudaMemcpyAsync(out_rgba, c->d_frame_rgba + 4 * off, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
    if (out_0rgb) CU(c, cudaMemcpyAsync(out_0rgb, c->d_frame_0rgb + off, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return check_launch(c, "render");
}

int nerf_render_sharded(nerf_ctx *c, float yaw, float pitch, int32_t randomize, uint64_t seed, float *out_rgba, uint32_t *out_0rgb) {
    if (!c) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    const int W = c->cfg.image_w, H = c->cfg.image_h, N = c->comm.nranks, r = c->comm.rank;
    if (H % N != 0) return fail(c, NERF_ERR_INVALID_ARG, "render_sharded: image height must divide evenly among the ranks");
    const int rows = H / N;
    int rc = render_rows_device(c, yaw, pitch, r * rows, (r + 1) * rows, randomize, seed, out_0rgb != nullptr);
    if (rc) return rc;
    if (N > 1) {   // the one collective of the render path: gather the row bands (in place: band r already sits at its offset)
        char eb[256] = {0};
        Scope s(c, "render_allgather");
        if (out_rgba && comm_allgather_bytes(c->comm, c->d_frame_rgba + 4 * (size_t)r * rows * W, c->d_frame_rgba, sizeof(float) * 4 * (size_t)rows * W,
                                             c->stream, eb, sizeof(eb)))
            return fail(c, NERF_ERR_COMM, eb);
        if (out_0rgb && comm_allgather_bytes(c->comm, c->d_frame_0rgb + (size_t)r * rows * W, c->d_frame_0rgb, sizeof(uint32_t) * (size_t)rows * W,
                                             c->stream, eb, sizeof(eb)))
            return fail(c, NERF_ERR_COMM, eb);
    }
    if (out_rgba) CU(c, cudaMemcpyAsync(out_rgba, c->d_frame_rgba, sizeof(float) * 4 * (size_t)W * H, cudaMemcpyDeviceToHost, c->stream));
    if (out_0rgb) CU(c, cudaMemcpyAsync(out_0rgb, c->d_frame_0rgb, sizeof(uint32_t) * (size_t)W * H, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return check_launch(c, "render_sharded");
}

int nerf_log_metrics(nerf_ctx *c, const nerf_metrics *m) {
    if (!c || !m) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    if (!c->batch_valid) return fail(c, NERF_ERR_STATE, "log_metrics: no batch (call nerf_get_batch first)");
    const bool want_density = m->density_x || m->density_y || m->density_z || m->density_yx || m->density_zx || m->density_yz;
    if ((want_density || m->prediction) && !c->outputs_valid) {
        if (!c->predicted || c->fwd_deferred)
            return fail(c, NERF_ERR_STATE, "log_metrics: densities / predictions need nerf_predict (or a completed nerf_step / nerf_train_iter) on this batch");
        int rc = ensure_outputs(c);
        if (rc) return rc;
    }
    if (!c->points_valid && !c->batch_poses) return fail(c, NERF_ERR_STATE, "log_metrics: batch has neither points nor ray records");
    const int W = c->cfg.image_w, H = c->cfg.image_h;
    // scratch layout (device): screen[W+H] u32 | t[2000] u32 | world[30000] u32 | pad | density_hist[6000] f64 | density keys[30000] u64 | prediction keys[W*H] u64 | resolved u32 [30000 + W*H]
    const size_t n_u32 = (size_t)W + H + 2000 + 30000;
    const size_t off_f64 = (n_u32 * 4 + 7) / 8 * 8;
    const size_t off_dkeys = off_f64 + 6000 * 8;
    const size_t off_pkeys = off_dkeys + 30000 * 8;
    const size_t off_res = off_pkeys + (size_t)W * H * 8;
    const size_t total = off_res + (30000 + (size_t)W * H) * 4;
    if (!c->d_metrics) CU(c, guard_malloc(&c->d_metrics, total));
    uint8_t *base = c->d_metrics;
    CU(c, cudaMemsetAsync(base, 0, off_res, c->stream));
    MetricsArgs a;
    memset(&a, 0, sizeof(a));
    a.pix_yx = c->d_pix; a.rays = c->d_rays; a.poses = c->batch_poses; a.t = c->d_t;
    a.points = c->points_valid ? c->d_points : nullptr;
    a.sigma = want_density ? c->d_sigma : nullptr;
    a.pixels = m->prediction ? c->d_out : nullptr;
    a.num_rays = c->R; a.num_samples = c->S; a.img_w = W; a.img_h = H;
    unsigned int *u32 = reinterpret_cast<unsigned int *>(base);
    a.screen_hist = u32;
    a.t_hist = u32 + W + H;
    a.world_maps = u32 + W + H + 2000;
    a.density_hist = want_density ? reinterpret_cast<double *>(base + off_f64) : nullptr;
    a.density_maps = want_density ? reinterpret_cast<unsigned long long *>(base + off_dkeys) : nullptr;
    a.prediction = m->prediction ? reinterpret_cast<unsigned long long *>(base + off_pkeys) : nullptr;
    uint32_t *res = reinterpret_cast<uint32_t *>(base + off_res);
    {
        Scope s(c, "metrics", 2);
        launch_metrics(a, c->num_sms, c->stream);
        launch_metrics_resolve(reinterpret_cast<unsigned long long *>(base + off_dkeys), res, 30000 + (m->prediction ? (int64_t)W * H : 0), c->stream);
    }
    int rc = check_launch(c, "metrics");
    if (rc) return rc;
    std::vector<uint32_t> h((size_t)W + H + 2000);
    CU(c, cudaMemcpyAsync(h.data(), u32, h.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    auto d2h = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
        return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
    };
    CU(c, d2h(m->world_yx, a.world_maps, 40000));
    CU(c, d2h(m->world_zx, a.world_maps + 10000, 40000));
    CU(c, d2h(m->world_yz, a.world_maps + 20000, 40000));
    if (want_density) {
        CU(c, d2h(m->density_x, a.density_hist, 16000));
        CU(c, d2h(m->density_y, a.density_hist + 2000, 16000));
        CU(c, d2h(m->density_z, a.density_hist + 4000, 16000));
        CU(c, d2h(m->density_yx, res, 40000));
        CU(c, d2h(m->density_zx, res + 10000, 40000));
        CU(c, d2h(m->density_yz, res + 20000, 40000));
    }
    if (m->prediction) CU(c, d2h(m->prediction, res + 30000, (size_t)W * H * 4));
    CU(c, cudaStreamSynchronize(c->stream));
    // the reference accumulates its counts in f64 (logging.rs:14-15, 32)
    if (m->screen_x) for (int i = 0; i < W; ++i) m->screen_x[i] = (double)h[i];
    if (m->screen_y) for (int i = 0; i < H; ++i) m->screen_y[i] = (double)h[(size_t)W + i];
    if (m->t_hist) for (int i = 0; i < 2000; ++i) m->t_hist[i] = (double)h[(size_t)W + H + i];
    return NERF_OK;
}

int nerf_comm_unique_id(void *id128) {
    if (!id128) return NERF_ERR_INVALID_ARG;
    char eb[256] = {0};
    if (comm_unique_id(id128, eb, sizeof(eb))) {
        fprintf(stderr, "nerf_comm_unique_id: %s\n", eb);
        return NERF_ERR_COMM;
    }
    return NERF_OK;
}

int nerf_comm_init_rank(nerf_ctx *c, const void *id128, int32_t rank, int32_t nranks) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    char eb[256] = {0};
    if (comm_init_rank(c->comm, id128, rank, nranks, eb, sizeof(eb))) return fail(c, NERF_ERR_COMM, eb);
    // peer-memory gradient exchange (fused all-reduce + Adam); NERF_B200_P2P=0 keeps the plain NCCL all-reduce
    const char *env = getenv("NERF_B200_P2P");
    if (nranks > 1 && nranks <= NERF_MAX_RANKS && !(env && env[0] == '0')) {
        const size_t bytes = 2 * sizeof(float) * ((c->g.n_params + 3) / 4 * 4);   // local gradient | reduced gradient (kernels.h)
        for (int b = 0; b < 2; ++b) {
            if (!c->d_gacc[b]) CU(c, cudaMalloc(&c->d_gacc[b], bytes));   // (plain: the IPC export needs the allocation base)
            CU(c, cudaMemsetAsync(c->d_gacc[b], 0, bytes, c->stream));
        }
        if (comm_p2p_setup(c->comm, c->d_gacc[0], c->d_gacc[1], c->stream, eb, sizeof(eb))) return fail(c, NERF_ERR_COMM, eb);
        if (!c->comm.p2p && eb[0]) c->err = eb;   // informational: fell back to NCCL
    }
    return NERF_OK;
}

int nerf_comm_destroy(nerf_ctx *c) {
    if (!c) return NERF_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);   // comm_destroy's good-bye means "my stream has drained"
    comm_destroy(c->comm);
    return NERF_OK;
}

int nerf_timer_start(nerf_ctx *c) {
    if (!c) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    CU(c, cudaEventRecord(c->t0, c->stream));
    return NERF_OK;
}
int nerf_timer_stop(nerf_ctx *c, float *ms) {
    if (!c || !ms) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    CU(c, cudaEventRecord(c->t1, c->stream));
    CU(c, cudaEventSynchronize(c->t1));
    CU(c, cudaEventElapsedTime(ms, c->t0, c->t1));
    return check_launch(c, "timer_stop");
}

int nerf_profile_enable(nerf_ctx *c, int32_t on) {
    if (!c) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    c->prof.collect(c->stream);
    c->prof.reset();
    c->prof.on = on != 0;
    return NERF_OK;
}
int nerf_profile_read(nerf_ctx *c, char *names, float *total_ms, int32_t *launches, int32_t capacity, int32_t *count) {
    if (!c || !count) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    c->prof.collect(c->stream);
    const int n = (int)c->prof.names.size();
    *count = n;
    for (int i = 0; i < n && i < capacity; ++i) {
        if (names) { strncpy(names + 32 * i, c->prof.names[i].c_str(), 31); names[32 * i + 31] = 0; }
        if (total_ms) total_ms[i] = (float)c->prof.total_ms[i];
        if (launches) launches[i] = c->prof.launches[i];
    }
    return NERF_OK;
}

int nerf_flush_l2(nerf_ctx *c) {
    if (!c) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    if (!c->d_flush) {
        c->flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
        CU(c, guard_malloc(&c->d_flush, c->flush_bytes));
    }
    CU(c, cudaMemsetAsync(c->d_flush, (int)(c->launch_count & 0xff), c->flush_bytes, c->stream));
    return NERF_OK;
}

// ------------------------------------------------------------------------------- debug
int nerf_debug_plan(const nerf_config *cfg, int32_t program, void *ops, int32_t *n_ops, void *jobs, int32_t *n_jobs, void *chunks,
                    int32_t *n_chunks, void *units, int32_t *n_units, int32_t *info) {
    if (!cfg) return NERF_ERR_INVALID_ARG;
    NetGeom g;
    build_geom(*cfg, g);
    TcPlan plan;
    std::string err;
    if (!tc_build_plan(g, plan, err)) {
        fprintf(stderr, "nerf_debug_plan: %s\n", err.c_str());
        return NERF_ERR_UNSUPPORTED;
    }
    const TcProgram &P = program == 0 ? plan.fwd_train : (program == 1 ? plan.fwd_infer : plan.bwd);
    if (ops && n_ops && *n_ops >= (int)P.ops.size()) memcpy(ops, P.ops.data(), P.ops.size() * sizeof(MmaOp));
    if (jobs && n_jobs && *n_jobs >= (int)P.jobs.size()) memcpy(jobs, P.jobs.data(), P.jobs.size() * sizeof(EpiJob));
    if (chunks && n_chunks && *n_chunks >= (int)P.chunks.size()) memcpy(chunks, P.chunks.data(), P.chunks.size() * sizeof(PackChunk));
    if (units && n_units && *n_units >= (int)plan.units.size()) memcpy(units, plan.units.data(), plan.units.size() * sizeof(WgradUnit));
    if (n_ops) *n_ops = (int)P.ops.size();
    if (n_jobs) *n_jobs = (int)P.jobs.size();
    if (n_chunks) *n_chunks = (int)P.chunks.size();
    if (n_units) *n_units = (int)plan.units.size();
    if (info) {
        info[0] = (int)sizeof(MmaOp); info[1] = (int)sizeof(EpiJob); info[2] = (int)sizeof(PackChunk); info[3] = (int)sizeof(WgradUnit);
        info[4] = (int)P.wpack_bytes; info[5] = plan.act_slots; info[6] = plan.grad_slots; info[7] = plan.mask_slots;
        info[8] = (int)plan.bias_floats; info[9] = (int)g.n_params;
    }
    return NERF_OK;
}

int nerf_debug_lane_plan(const nerf_config *cfg, int32_t program, void *ops, int32_t *n_ops, void *gemms, int32_t *n_gemms, void *jobs,
                         int32_t *n_jobs) {
    if (!cfg || !n_ops || !n_gemms || !n_jobs) return NERF_ERR_INVALID_ARG;
    NetGeom g;
    build_geom(*cfg, g);
    TcPlan plan;
    std::string err;
    if (!tc_build_plan(g, plan, err)) return NERF_ERR_UNSUPPORTED;
    LaneProgram lp;
    if (!make_lane_program(program == 0 ? plan.fwd_train : (program == 1 ? plan.fwd_infer : plan.bwd), lp, err)) {
        fprintf(stderr, "nerf_debug_lane_plan: %s\n", err.c_str());
        return NERF_ERR_UNSUPPORTED;
    }
    if (ops && *n_ops >= (int)lp.ops.size()) memcpy(ops, lp.ops.data(), lp.ops.size() * sizeof(LaneOp));
    if (gemms && *n_gemms >= (int)lp.gemms.size()) memcpy(gemms, lp.gemms.data(), lp.gemms.size() * sizeof(LaneGemm));
    if (jobs && *n_jobs >= (int)lp.jobs.size()) memcpy(jobs, lp.jobs.data(), lp.jobs.size() * sizeof(LaneJob));
    *n_ops = (int)lp.ops.size();
    *n_gemms = (int)lp.gemms.size();
    *n_jobs = (int)lp.jobs.size();
    return NERF_OK;
}

int nerf_debug_ts_plan(const nerf_config *cfg, int32_t program, void *ops, int32_t *n_ops, void *steps, int32_t *n_steps, void *chunks,
                       int32_t *n_chunks, int32_t *info) {
    if (!cfg || !n_ops || !n_steps || !n_chunks) return NERF_ERR_INVALID_ARG;
    NetGeom g;
    build_geom(*cfg, g);
    TcPlan plan;
    std::string err;
    if (!tc_build_plan(g, plan, err) || plan.np > 4) return NERF_ERR_UNSUPPORTED;
    const TsProgram &P = program == 0 ? plan.ts_fwd_train : (program == 1 ? plan.ts_fwd_infer : plan.ts_bwd);
    if (ops && *n_ops >= (int)P.ops.size()) memcpy(ops, P.ops.data(), P.ops.size() * sizeof(TsOp));
    if (steps && *n_steps >= (int)P.steps.size()) memcpy(steps, P.steps.data(), P.steps.size() * sizeof(TsStep));
    if (chunks && *n_chunks >= (int)P.chunks.size()) memcpy(chunks, P.chunks.data(), P.chunks.size() * sizeof(PackChunk));
    *n_ops = (int)P.ops.size();
    *n_steps = (int)P.steps.size();
    *n_chunks = (int)P.chunks.size();
    if (info) { info[0] = (int)sizeof(TsOp); info[1] = (int)sizeof(TsStep); info[2] = (int)sizeof(PackChunk); info[3] = (int)P.wpack_bytes; }
    return NERF_OK;
}

int nerf_debug_plan_biases(const nerf_config *cfg, void *out, int32_t *n) {
    if (!cfg || !n) return NERF_ERR_INVALID_ARG;
    NetGeom g;
    build_geom(*cfg, g);
    TcPlan plan;
    std::string err;
    if (!tc_build_plan(g, plan, err)) return NERF_ERR_UNSUPPORTED;
    if (out && *n >= (int)plan.biases.size()) memcpy(out, plan.biases.data(), plan.biases.size() * sizeof(PackBias));
    *n = (int)plan.biases.size();
    return NERF_OK;
}

int nerf_debug_wgrad_partition(const nerf_config *cfg, int32_t n_ctas, int64_t n_tiles, int32_t *out, int32_t *unit_cost_panels, int32_t *n_units) {
    if (!cfg || !out || n_ctas < 1 || n_tiles < 1 || !n_units) return NERF_ERR_INVALID_ARG;
    NetGeom g;
    build_geom(*cfg, g);
    TcPlan plan;
    std::string err;
    if (!tc_build_plan(g, plan, err)) return NERF_ERR_UNSUPPORTED;
    std::vector<WgradWork> work;
    tc_wgrad_partition(plan.units, n_ctas, n_tiles, work);
    for (int c = 0; c < n_ctas; ++c) {
        int32_t *o = out + (size_t)c * (1 + 3 * kWgMaxSeg);
        o[0] = work[c].n_seg;
        for (int k = 0; k < kWgMaxSeg; ++k) {
            o[1 + 3 * k] = work[c].seg[k].unit; o[2 + 3 * k] = work[c].seg[k].tile_begin; o[3 + 3 * k] = work[c].seg[k].tile_end;
        }
    }
    if (unit_cost_panels && *n_units >= (int)plan.units.size())
        for (size_t i = 0; i < plan.units.size(); ++i) unit_cost_panels[i] = plan.units[i].n_p + plan.units[i].n_q;
    *n_units = (int)plan.units.size();
    return NERF_OK;
}

int nerf_debug_tc3_stats(uint64_t *out, int32_t ctas) {
    if (!out || ctas < 1) return NERF_ERR_INVALID_ARG;
    const int rc = tc3_debug_stats(reinterpret_cast<unsigned long long *>(out), ctas);
    return rc == 0 ? NERF_OK : (rc == -1 ? NERF_ERR_UNSUPPORTED : NERF_ERR_CUDA);
}

int nerf_debug_tc3_trace(uint64_t *out, int32_t n) {
    if (!out || n < 1) return NERF_ERR_INVALID_ARG;
    const int rc = tc3_debug_trace(reinterpret_cast<unsigned long long *>(out), n);
    return rc == 0 ? NERF_OK : (rc == -1 ? NERF_ERR_UNSUPPORTED : NERF_ERR_CUDA);
}

int nerf_debug_check_guards(int32_t *n_allocations) {
    int n = 0;
    const int bad = guard_check(&n);
    if (n_allocations) *n_allocations = n;
    return bad;   // -1: guard mode off (NERF_B200_GUARD=1 was not set when the library first allocated)
}

int nerf_debug_read_panel(nerf_ctx *c, int32_t area, int32_t tile, int32_t slot, void *out) {
    if (!c || !out) return NERF_ERR_INVALID_ARG;
    if (!c->tc) return fail(c, NERF_ERR_STATE, "debug_read_panel: not a tcgen05 context");
    ENTER(c);
    const int rc = tc_debug_read(c->tc, area, tile, slot, out, c->stream);
    if (rc == -1) return fail(c, NERF_ERR_INVALID_ARG, "debug_read_panel: bad area/tile/slot");
    if (rc) return fail(c, NERF_ERR_CUDA, "debug_read_panel: copy failed");
    return NERF_OK;
}

int nerf_debug_trace(nerf_ctx *c, int32_t program, uint64_t *out) {
    if (!c || !out) return NERF_ERR_INVALID_ARG;
    if (!c->tc) return fail(c, NERF_ERR_STATE, "debug_trace: not a tcgen05 context");
    if (!c->batch_valid) return fail(c, NERF_ERR_STATE, "debug_trace: no batch");
    ENTER(c);
    int rc = ensure_packed(c);
    if (rc) return rc;
    const int nr = c->chunk < c->R ? c->chunk : c->R;
    if (tc_debug_trace(c->tc, c->d_points, c->d_dirs, (int64_t)nr * c->S, c->S, program, c->d_rgba, c->d_dsigma, c->d_drgba,
                       c->d_sigma, c->d_rgba, (unsigned long long *)out, c->stream))
        return fail(c, NERF_ERR_CUDA, "debug_trace failed");
    return check_launch(c, "debug_trace");
}

// Stand-alone timing of one HBM-bound stage kernel on synthetic device-resident inputs of `num_rays` x `num_samples`
// (pick sizes whose working set exceeds the 126 MB L2 so every launch streams HBM). CUDA events on the context's stream.
int nerf_debug_bench_stage(nerf_ctx *c, int32_t stage, int32_t num_rays, int32_t num_samples, int32_t iters, float *ms_per_launch) {
    if (!c || !ms_per_launch || num_rays < 1 || num_samples < 1 || num_samples > 256 || iters < 1) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    const int64_t n = (int64_t)num_rays * num_samples;
    std::vector<void *> bufs;
    auto alloc = [&](size_t bytes) -> void * {
        void *p = nullptr;
        if (guard_malloc(&p, bytes) != cudaSuccess) return nullptr;
        bufs.push_back(p);
        return p;
    };
    auto release = [&]() { for (void *p : bufs) guard_free(p); };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int rc = NERF_OK;
    auto timed = [&](auto &&launch) {
        for (int i = 0; i < 2; ++i) launch(i);
        cudaEventRecord(e0, c->stream);
        for (int i = 0; i < iters; ++i) launch(2 + i);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_per_launch = ms / (float)iters;
    };
    if (stage == 0) {   // K-sample writing points + t (16 B/sample), Philox pixels and jitter, reference (sorted) depth mode
        SampleArgs a;
        memset(&a, 0, sizeof(a));
        ViewPose vp;
        make_pose(0.3f, 0.2f, vp);
        ViewPose *d_vp = (ViewPose *)alloc(sizeof(vp));
        a.pix_out = (int32_t *)alloc(sizeof(int32_t) * 2 * (size_t)num_rays);
        a.rays = (RayRec *)alloc(sizeof(RayRec) * (size_t)num_rays);
        a.dirs = (float *)alloc(sizeof(float) * 3 * (size_t)num_rays);
        a.t = (float *)alloc(sizeof(float) * n);
        a.points = (float *)alloc(sizeof(float) * 3 * n);
        if (!d_vp || !a.pix_out || !a.rays || !a.dirs || !a.t || !a.points) { release(); return fail(c, NERF_ERR_CUDA, "bench_stage: out of memory"); }
        cudaMemcpy(d_vp, &vp, sizeof(vp), cudaMemcpyHostToDevice);
        a.pix_yx = a.pix_out;
        a.gen_pix = 1;
        a.rays_per_pick = 1;
        a.poses = d_vp;
        a.num_rays = num_rays;
        a.num_samples = num_samples;
        a.img_w = c->cfg.image_w;
        a.img_h = c->cfg.image_h;
        a.randomize = 1;
        a.depth_mode = c->cfg.depth_mode;
        a.off = c->off;
        timed([&](int i) { a.seed = 77 + i; launch_sample(a, c->num_sms, c->stream); });
    } else if (stage == 1 || stage == 2) {   // K-composite forward / backward (+ fused MSE gradient and loss)
        CompositeArgs a;
        memset(&a, 0, sizeof(a));
        float *sig = (float *)alloc(sizeof(float) * n), *col = (float *)alloc(sizeof(float) * 4 * n), *dl = (float *)alloc(sizeof(float) * n);
        float *out = (float *)alloc(sizeof(float) * 4 * (size_t)num_rays), *gold = (float *)alloc(sizeof(float) * 4 * (size_t)num_rays);
        float *rl = (float *)alloc(sizeof(float) * (size_t)num_rays), *loss = (float *)alloc(sizeof(float) * 4);
        float *ds = stage == 2 ? (float *)alloc(sizeof(float) * n) : nullptr, *dc = stage == 2 ? (float *)alloc(sizeof(float) * 4 * n) : nullptr;
        if (!sig || !col || !dl || !out || !gold || !rl || !loss || (stage == 2 && (!ds || !dc))) { release(); return fail(c, NERF_ERR_CUDA, "bench_stage: out of memory"); }
        launch_fill_uniform(sig, n, 1, 0.f, 4.f, c->stream);
        launch_fill_uniform(col, 4 * n, 2, 0.f, 1.f, c->stream);
        launch_fill_uniform(dl, n, 3, 0.f, 2.f / (float)num_samples, c->stream);
        launch_fill_uniform(gold, 4 * (int64_t)num_rays, 4, 0.f, 1.f, c->stream);
        cudaMemsetAsync(loss, 0, sizeof(float) * 4, c->stream);
        a.sigma = sig; a.colors = col; a.t_or_delta = dl;
        a.input_is_delta = 1;
        a.sigma_relu = c->cfg.sigma_relu;
        a.num_rays = num_rays; a.num_samples = num_samples;
        a.out = out;
        if (stage == 2) {
            a.gold = gold;
            a.inv_count = 1.f / (4.f * (float)num_rays);
            a.loss_partials = rl; a.d_sigma = ds; a.d_colors = dc;
            a.loss_out = loss; a.loss_partials_first = rl; a.loss_scale = a.inv_count;
            a.done_counter = reinterpret_cast<unsigned int *>(loss + 2);
            timed([&](int) { launch_composite_bwd(a, c->num_sms, c->stream); });
        } else {
            timed([&](int) { launch_composite_fwd(a, c->num_sms, c->stream); });
        }
    } else if (stage == 3) {   // K-adam over num_rays*num_samples parameters (28 B/param)
        AdamArgs a;
        memset(&a, 0, sizeof(a));
        a.p = (float *)alloc(sizeof(float) * n); a.m = (float *)alloc(sizeof(float) * n);
        a.v = (float *)alloc(sizeof(float) * n); a.g = (float *)alloc(sizeof(float) * n);
        if (!a.p || !a.m || !a.v || !a.g) { release(); return fail(c, NERF_ERR_CUDA, "bench_stage: out of memory"); }
        launch_fill_uniform(a.p, n, 1, -1.f, 1.f, c->stream);
        launch_fill_uniform(a.g, n, 2, -1.f, 1.f, c->stream);
        cudaMemsetAsync(a.m, 0, sizeof(float) * n, c->stream);
        cudaMemsetAsync(a.v, 0, sizeof(float) * n, c->stream);
        a.n = n;
        a.lr_over_bc1 = 5e-4f; a.inv_sqrt_bc2 = 1.f; a.beta1 = 0.9f; a.beta2 = 0.999f; a.eps = 1e-8f; a.grad_scale = 1.f;
        timed([&](int) { launch_adam(a, c->num_sms, c->stream); });
    } else {
        rc = fail(c, NERF_ERR_INVALID_ARG, "bench_stage: stage 0 sample, 1 composite fwd, 2 composite bwd, 3 adam");
    }
    cudaStreamSynchronize(c->stream);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    release();
    if (rc) return rc;
    return check_launch(c, "bench_stage");
}

int nerf_debug_wgrad_marks(nerf_ctx *c, uint64_t *out, int32_t capacity_ctas) {
    if (!c || !out || !c->tc) return NERF_ERR_INVALID_ARG;
    ENTER(c);
    const int n = tc_debug_wgrad_marks(c->tc, reinterpret_cast<unsigned long long *>(out), capacity_ctas, c->stream);
    if (n < 0) return fail(c, NERF_ERR_INVALID_ARG, "wgrad_marks: capacity too small or copy failed");
    return n;
}

int nerf_debug_host_pose(float yaw, float pitch, float *yaw3x4, float *pitch3x3, float *off) {
    ViewPose vp;
    make_pose(yaw, pitch, vp);
    if (yaw3x4) memcpy(yaw3x4, vp.yaw, sizeof(vp.yaw));
    if (pitch3x3) memcpy(pitch3x3, vp.pitch, sizeof(vp.pitch));
    if (off) *off = screen_offset();
    return NERF_OK;
}

}  // extern "C"
