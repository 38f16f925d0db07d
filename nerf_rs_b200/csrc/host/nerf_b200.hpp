// nerf_b200.hpp -- C++ host-side mirror of the reference's call surface over the C ABI.
//
// The reference host is Rust (src/main.rs:57-72); the Rust toolchain is absent from the build
// image, so the compiled-language host layer above the C ABI is this header (the Rust binding of
// the same surface ships as source in ffi/rust/). Names and argument meaning follow the reference:
//   nerf::NeRF::predict            <- NeRF::predict            src/model.rs:152-209
//   nerf::compositing              <- compositing              src/model.rs:234-249
//   nerf::Trainer::step            <- Trainer::step            src/model.rs:311-325
//   nerf::get_multiview_batch      <- dataset::get_multiview_batch  src/dataset.rs:63-139
// Where the reference panics (assert_eq!/unwrap) these throw nerf::Error; nothing throws across
// the C ABI itself.
#pragma once
#include <array>
#include <cstdint>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../../include/nerf_b200.h"

namespace nerf {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};

inline void check(nerf_ctx *c, int status) {
    if (status != NERF_OK)
        throw Error(status, std::string(nerf_strerror(status)) + (c ? std::string(": ") + nerf_last_error(c) : std::string()));
}

inline std::vector<std::pair<float, float>> get_view_angles(int num_views) {  // image_loading.rs:67-80
    std::vector<float> buf(4 * (size_t)num_views * (num_views + 1));
    check(nullptr, nerf_view_angles_grid(num_views, buf.data(), (int32_t)buf.size()));
    std::vector<std::pair<float, float>> out(buf.size() / 2);
    for (size_t i = 0; i < out.size(); ++i) out[i] = {buf[2 * i], buf[2 * i + 1]};
    return out;
}

// image_loading::load_image_as_array's decode step (image_loading.rs:7): 8-bit RGBA PNG -> bytes; throws for anything else
inline std::vector<uint8_t> load_image_rgba8(const std::string &path, int *width = nullptr, int *height = nullptr) {
    int32_t w = 0, h = 0;
    check(nullptr, nerf_load_png_rgba8(path.c_str(), nullptr, 0, &w, &h));
    std::vector<uint8_t> out((size_t)w * h * 4);
    check(nullptr, nerf_load_png_rgba8(path.c_str(), out.data(), (int64_t)out.size(), &w, &h));
    if (width) *width = w;
    if (height) *height = h;
    return out;
}
// image_loading.rs:6-24: one [r,g,b,a] per pixel, each `as f32 / 255.`
inline std::vector<std::array<float, 4>> load_image_as_array(const std::string &path) {
    const std::vector<uint8_t> b = load_image_rgba8(path);
    std::vector<std::array<float, 4>> out(b.size() / 4);
    for (size_t i = 0; i < out.size(); ++i)
        out[i] = {(float)b[4 * i] / 255.f, (float)b[4 * i + 1] / 255.f, (float)b[4 * i + 2] / 255.f, (float)b[4 * i + 3] / 255.f};
    return out;
}

class NeRF {
   public:
    explicit NeRF(const nerf_config &cfg, int device = 0) : cfg_(cfg) { check(nullptr, nerf_create(&cfg_, device, &ctx_)); }
    NeRF() : NeRF(default_config()) {}   // NeRF::new() (model.rs:140)
    ~NeRF() { nerf_destroy(ctx_); }
    NeRF(const NeRF &) = delete;
    NeRF &operator=(const NeRF &) = delete;

    static nerf_config default_config() {
        nerf_config c;
        check(nullptr, nerf_default_config(&c));
        return c;
    }
    static nerf_config as_shipped_config() {
        nerf_config c;
        check(nullptr, nerf_config_as_shipped(&c));
        return c;
    }

    // (pixels [R*4], densities [R*S]) -- query_points [B*3], distances [B] (model.rs:152-156)
    std::pair<std::vector<float>, std::vector<float>> predict(const std::vector<float> &query_points,
                                                              const std::vector<float> &distances,
                                                              const std::vector<float> *dirs = nullptr, bool train = true) {
        std::vector<float> out((size_t)cfg_.num_rays * 4), sig((size_t)cfg_.num_rays * cfg_.num_samples);
        ++pred_gen_;
        check(ctx_, nerf_predict_points(ctx_, query_points.data(), (int64_t)query_points.size(), distances.data(),
                                        (int64_t)distances.size(), dirs ? dirs->data() : nullptr, train ? 1 : 0, out.data(), sig.data()));
        return {std::move(out), std::move(sig)};
    }
    // The prediction as the reference has it: a device tensor (main.rs:58). Enqueues the forward and returns without a device
    // synchronisation; Trainer::step(Prediction, gold) consumes it, pixels() / densities() fetch the values on demand.
    struct Prediction {
        NeRF *model;
        uint64_t gen;
        std::vector<float> pixels() const {
            model->check_current(gen);
            std::vector<float> out((size_t)model->cfg_.num_rays * 4);
            check(model->ctx_, nerf_get_predictions(model->ctx_, out.data(), nullptr));
            return out;
        }
        std::vector<float> densities() const {
            model->check_current(gen);
            std::vector<float> sig((size_t)model->cfg_.num_rays * model->cfg_.num_samples);
            check(model->ctx_, nerf_get_predictions(model->ctx_, nullptr, sig.data()));
            return sig;
        }
    };
    Prediction predict_device(const std::vector<float> &query_points, const std::vector<float> &distances,
                              const std::vector<float> *dirs = nullptr, bool train = true) {
        check(ctx_, nerf_predict_points(ctx_, query_points.data(), (int64_t)query_points.size(), distances.data(),
                                        (int64_t)distances.size(), dirs ? dirs->data() : nullptr, train ? 1 : 0, nullptr, nullptr));
        return Prediction{this, ++pred_gen_};
    }
    void check_current(uint64_t gen) const {
        if (gen != pred_gen_) throw Error(NERF_ERR_STATE, "stale prediction handle: the model has run another predict since");
    }
    std::vector<float> predict_resident(bool train = true) {
        std::vector<float> out((size_t)cfg_.num_rays * 4);
        ++pred_gen_;
        check(ctx_, nerf_predict(ctx_, train ? 1 : 0, out.data(), nullptr));
        return out;
    }
    void save_weights(std::vector<float> &flat) {  // NeRF::save surface (model.rs:211-213)
        flat.resize((size_t)nerf_num_params(ctx_));
        check(ctx_, nerf_get_weights(ctx_, flat.data(), (int64_t)flat.size()));
    }
    void load_weights(const std::vector<float> &flat) {  // NeRF::load surface (model.rs:215-217)
        check(ctx_, nerf_set_weights(ctx_, flat.data(), (int64_t)flat.size()));
    }
    void set_images(const std::vector<std::vector<std::array<float, 4>>> &imgs) {
        std::vector<float> flat;
        for (auto &im : imgs)
            for (auto &px : im) flat.insert(flat.end(), px.begin(), px.end());
        check(ctx_, nerf_set_images(ctx_, flat.data(), (int32_t)imgs.size()));
        n_views_ = (int)imgs.size();
    }
    // residency from RGBA8 bytes (4 B/pixel on the device; the gold gather does the `as f32 / 255.` of image_loading.rs:13-18)
    void set_images_rgba8(const std::vector<std::vector<uint8_t>> &imgs) {
        std::vector<uint8_t> flat;
        for (auto &im : imgs) {
            // a decoded file of another size than the context's image_w x image_h must not reach the device copy
            if (im.size() != (size_t)cfg_.image_w * cfg_.image_h * 4) throw std::runtime_error("set_images_rgba8: image size differs from the configuration");
            flat.insert(flat.end(), im.begin(), im.end());
        }
        check(ctx_, nerf_set_images_rgba8(ctx_, flat.data(), (int32_t)imgs.size()));
        n_views_ = (int)imgs.size();
    }
    void set_view_angles(const std::vector<std::pair<float, float>> &va) {
        std::vector<float> flat;
        for (auto &p : va) { flat.push_back(p.first); flat.push_back(p.second); }
        check(ctx_, nerf_set_view_angles(ctx_, flat.data(), (int32_t)va.size()));
    }
    // The per-batch logging projections of src/logging.rs (+ draw_predictions, display.rs:96-110), computed on the device.
    struct BatchLog {
        std::vector<double> screen_x, screen_y, t;                 // log_screen_coords :13-25, log_query_distances :27-39
        std::vector<uint32_t> world_yx, world_zx, world_yz;       // log_query_points_as_maps :41-107 (100 x 100, 0x00RRGGBB)
        std::vector<double> density_x, density_y, density_z;      // log_densities :109-134
        std::vector<uint32_t> density_yx, density_zx, density_yz; // log_density_maps :136-195
        std::vector<uint32_t> prediction;                          // draw_predictions: image_h x image_w back buffer
    };
    BatchLog log_metrics(bool densities = true, bool prediction = true) {
        BatchLog b;
        b.screen_x.resize(cfg_.image_w); b.screen_y.resize(cfg_.image_h); b.t.resize(2000);
        b.world_yx.resize(10000); b.world_zx.resize(10000); b.world_yz.resize(10000);
        nerf_metrics m{};
        m.screen_x = b.screen_x.data(); m.screen_y = b.screen_y.data(); m.t_hist = b.t.data();
        m.world_yx = b.world_yx.data(); m.world_zx = b.world_zx.data(); m.world_yz = b.world_yz.data();
        if (densities) {
            b.density_x.resize(2000); b.density_y.resize(2000); b.density_z.resize(2000);
            b.density_yx.resize(10000); b.density_zx.resize(10000); b.density_yz.resize(10000);
            m.density_x = b.density_x.data(); m.density_y = b.density_y.data(); m.density_z = b.density_z.data();
            m.density_yx = b.density_yx.data(); m.density_zx = b.density_zx.data(); m.density_yz = b.density_yz.data();
        }
        if (prediction) {
            b.prediction.resize((size_t)cfg_.image_w * cfg_.image_h);
            m.prediction = b.prediction.data();
        }
        check(ctx_, nerf_log_metrics(ctx_, &m));
        return b;
    }
    nerf_ctx *ctx() { return ctx_; }
    const nerf_config &config() const { return cfg_; }
    int n_views() const { return n_views_; }

   private:
    nerf_config cfg_;
    nerf_ctx *ctx_ = nullptr;
    int n_views_ = 0;
    uint64_t pred_gen_ = 0;   // bumped by every predict: older Prediction handles become stale
};

// compositing(&densities [R,S], colors [R,S,4], distances [R,S]) -> [R,4]  (model.rs:234)
inline std::vector<float> compositing(NeRF &m, const std::vector<float> &densities, const std::vector<float> *colors,
                                      const std::vector<float> &distances, int num_rays, int num_samples) {
    std::vector<float> out((size_t)num_rays * 4);
    check(m.ctx(), nerf_compositing(m.ctx(), densities.data(), colors ? colors->data() : nullptr, distances.data(), num_rays,
                                    num_samples, out.data()));
    return out;
}

class Trainer {  // Trainer::new(&vs, lr) / step (model.rs:301-325)
   public:
    explicit Trainer(NeRF &m) : m_(m) {}
    float step(const NeRF::Prediction &prediction, const std::vector<float> &gold, size_t /*iter*/ = 0) {
        if (prediction.model != &m_) throw Error(NERF_ERR_INVALID_ARG, "the prediction belongs to another model");
        m_.check_current(prediction.gen);
        float loss = 0.f;
        check(m_.ctx(), nerf_step(m_.ctx(), gold.data(), (int64_t)gold.size(), &loss));
        return loss;
    }
    float step(const std::vector<float> &predictions, const std::vector<float> &gold, size_t /*iter*/ = 0) {
        if (predictions.size() != (size_t)m_.config().num_rays * 4) throw Error(NERF_ERR_INVALID_ARG, "predictions must be [NUM_RAYS, LABELS]");
        float loss = 0.f;
        check(m_.ctx(), nerf_step(m_.ctx(), gold.data(), (int64_t)gold.size(), &loss));
        return loss;
    }

   private:
    NeRF &m_;
};

// get_multiview_batch(&imgs, &view_angles) -> (indices [R][2], query_points [R*S*3], distances [R*S], gold [R*4])
// Images / angles must already be resident (set_images / set_view_angles).
template <class Rng>
inline std::tuple<std::vector<std::array<int64_t, 2>>, std::vector<float>, std::vector<float>, std::vector<float>>
get_multiview_batch(NeRF &m, Rng &rng) {
    const nerf_config &c = m.config();
    const int R = c.num_rays, S = c.num_samples, V = m.n_views();
    if (V < 1 || R % V != 0) throw Error(NERF_ERR_INVALID_ARG, "Can't divide rays evenly among views (dataset.rs:73-81)");
    std::vector<int64_t> idx(2 * (size_t)R), vi((size_t)V);
    std::uniform_int_distribution<int64_t> dy(0, c.image_h - 1), dx(0, c.image_w - 1), dv(0, V - 1);
    for (int i = 0; i < R; ++i) { idx[2 * i] = dy(rng); idx[2 * i + 1] = dx(rng); }
    for (int i = 0; i < V; ++i) vi[i] = dv(rng);
    std::vector<float> pts((size_t)R * S * 3), t((size_t)R * S), gold((size_t)R * 4);
    check(m.ctx(), nerf_get_batch(m.ctx(), idx.data(), vi.data(), V, nullptr, 1, (uint64_t)rng(), pts.data(), t.data(), gold.data(),
                                  nullptr, nullptr));
    std::vector<std::array<int64_t, 2>> indices((size_t)R);
    for (int i = 0; i < R; ++i) indices[i] = {idx[2 * i], idx[2 * i + 1]};
    return {std::move(indices), std::move(pts), std::move(t), std::move(gold)};
}

}  // namespace nerf
