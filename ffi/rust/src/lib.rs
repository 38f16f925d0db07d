//! Thin FFI crate over `include/nerf_b200.h`.
//!
//! `sys` is the raw `extern "C"` surface; the safe wrappers keep the reference's call surface
//! (`src/main.rs:57-72`) so `main.rs` changes only its `use` lines:
//!
//! ```ignore
//! let (indices, query_points, distances, gold) = get_multiview_batch(&mut model, &mut rng);   // dataset.rs:63
//! let (colors, densities) = model.predict(&query_points, &distances, Some(&dirs));           // model.rs:152
//! let loss = trainer.step(&mut model, &colors, &gold, &iter);                                  // model.rs:311
//! ```
//! Panics of the reference (`assert_eq!`, `unwrap`) become `Err(NerfError)`.
//! NOTE: shipped as source; the build image has no Rust toolchain, so this crate is exercised only
//! through the identical C ABI (C++ mirror `csrc/host/nerf_b200.hpp`, ctypes mirror `nerf_rs_b200/api.py`).

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub mod sys {
    use super::*;

    #[repr(C)]
    #[derive(Clone, Copy, Debug)]
    pub struct nerf_config {
        pub struct_size: i32,
        pub image_w: i32,
        pub image_h: i32,
        pub num_rays: i32,
        pub num_samples: i32,
        pub hidden: i32,
        pub xyz_freqs: i32,
        pub dir_freqs: i32,
        pub skip_layer: i32,
        pub use_rgb_head: i32,
        pub sigma_relu: i32,
        pub depth_mode: i32,
        pub mlp_impl: i32,
        pub max_rays_per_launch: i32,
        pub learning_rate: f32,
        pub beta1: f32,
        pub beta2: f32,
        pub eps: f32,
        pub deterministic_grads: i32,
    }

    #[repr(C)]
    pub struct nerf_ctx {
        _private: [u8; 0],
    }

    /// Host output pointers of `nerf_log_metrics` (null = skip); see include/nerf_b200.h.
    #[repr(C)]
    pub struct nerf_metrics {
        pub screen_x: *mut f64, pub screen_y: *mut f64, pub t_hist: *mut f64,
        pub world_yx: *mut u32, pub world_zx: *mut u32, pub world_yz: *mut u32,
        pub density_x: *mut f64, pub density_y: *mut f64, pub density_z: *mut f64,
        pub density_yx: *mut u32, pub density_zx: *mut u32, pub density_yz: *mut u32,
        pub prediction: *mut u32,
    }

    extern "C" {
        pub fn nerf_abi_version() -> c_int;
        pub fn nerf_default_config(cfg: *mut nerf_config) -> c_int;
        pub fn nerf_config_as_shipped(cfg: *mut nerf_config) -> c_int;
        pub fn nerf_create(cfg: *const nerf_config, device: c_int, out: *mut *mut nerf_ctx) -> c_int;
        pub fn nerf_destroy(ctx: *mut nerf_ctx) -> c_int;
        pub fn nerf_last_error(ctx: *const nerf_ctx) -> *const c_char;
        pub fn nerf_strerror(status: c_int) -> *const c_char;
        pub fn nerf_num_params(ctx: *const nerf_ctx) -> i64;
        pub fn nerf_set_weights(ctx: *mut nerf_ctx, flat: *const f32, n: i64) -> c_int;
        pub fn nerf_get_weights(ctx: *mut nerf_ctx, flat: *mut f32, n: i64) -> c_int;
        pub fn nerf_get_grads(ctx: *mut nerf_ctx, flat: *mut f32, n: i64) -> c_int;
        pub fn nerf_get_adam_state(ctx: *mut nerf_ctx, m: *mut f32, v: *mut f32, n: i64, step: *mut i64) -> c_int;
        pub fn nerf_set_adam_state(ctx: *mut nerf_ctx, m: *const f32, v: *const f32, n: i64, step: i64) -> c_int;
        pub fn nerf_set_images(ctx: *mut nerf_ctx, rgba: *const f32, n_views: i32) -> c_int;
        pub fn nerf_log_metrics(ctx: *mut nerf_ctx, out: *const nerf_metrics) -> c_int;
        pub fn nerf_set_images_rgba8(ctx: *mut nerf_ctx, rgba8: *const u8, n_views: i32) -> c_int;
        pub fn nerf_load_png_rgba8(path: *const c_char, out: *mut u8, capacity_bytes: i64, width: *mut i32, height: *mut i32) -> c_int;
        pub fn nerf_set_view_angles(ctx: *mut nerf_ctx, yaw_pitch: *const f32, n_angles: i32) -> c_int;
        pub fn nerf_view_angles_grid(n: i32, out: *mut f32, capacity: i32) -> c_int;
        pub fn nerf_get_batch(ctx: *mut nerf_ctx, indices_yx: *const i64, view_index: *const i64, n_picks: i32,
                              jitter: *const f32, randomize: i32, seed: u64, out_points: *mut f32, out_t: *mut f32,
                              out_gold: *mut f32, out_dirs: *mut f32, out_indices: *mut i64) -> c_int;
        pub fn nerf_predict(ctx: *mut nerf_ctx, train: i32, out_rgba: *mut f32, out_sigma: *mut f32) -> c_int;
        pub fn nerf_predict_points(ctx: *mut nerf_ctx, query_points: *const f32, n_points_floats: i64,
                                   distances: *const f32, n_distances: i64, dirs: *const f32, train: i32,
                                   out_rgba: *mut f32, out_sigma: *mut f32) -> c_int;
        pub fn nerf_get_predictions(ctx: *mut nerf_ctx, out_rgba: *mut f32, out_sigma: *mut f32) -> c_int;
        pub fn nerf_compositing(ctx: *mut nerf_ctx, densities: *const f32, colors: *const f32, distances: *const f32,
                                num_rays: i32, num_samples: i32, out: *mut f32) -> c_int;
        pub fn nerf_step(ctx: *mut nerf_ctx, gold: *const f32, n_gold: i64, loss: *mut f32) -> c_int;
        pub fn nerf_train_iter(ctx: *mut nerf_ctx, seed: u64) -> c_int;
        pub fn nerf_last_loss(ctx: *mut nerf_ctx, loss: *mut f32) -> c_int;
        pub fn nerf_sync(ctx: *mut nerf_ctx) -> c_int;
        pub fn nerf_render(ctx: *mut nerf_ctx, yaw: f32, pitch: f32, y0: i32, y1: i32, randomize: i32, seed: u64,
                           out_rgba: *mut f32, out_0rgb: *mut u32) -> c_int;
        pub fn nerf_render_sharded(ctx: *mut nerf_ctx, yaw: f32, pitch: f32, randomize: i32, seed: u64,
                                   out_rgba: *mut f32, out_0rgb: *mut u32) -> c_int;
        pub fn nerf_comm_unique_id(id128: *mut c_void) -> c_int;
        pub fn nerf_comm_init_rank(ctx: *mut nerf_ctx, id128: *const c_void, rank: i32, nranks: i32) -> c_int;
        pub fn nerf_comm_destroy(ctx: *mut nerf_ctx) -> c_int;
    }
}

#[derive(Debug)]
pub struct NerfError {
    pub status: i32,
    pub message: String,
}

fn check(ctx: *const sys::nerf_ctx, status: c_int) -> Result<(), NerfError> {
    if status == 0 {
        return Ok(());
    }
    let mut message = unsafe { CStr::from_ptr(sys::nerf_strerror(status)) }.to_string_lossy().into_owned();
    if !ctx.is_null() {
        let detail = unsafe { CStr::from_ptr(sys::nerf_last_error(ctx)) }.to_string_lossy();
        if !detail.is_empty() {
            message = format!("{}: {}", message, detail);
        }
    }
    Err(NerfError { status, message })
}

/// `image_loading::get_view_angles` (image_loading.rs:67-80).
pub fn get_view_angles(num_views: usize) -> Vec<(f32, f32)> {
    let mut buf = vec![0f32; 4 * num_views * (num_views + 1)];
    unsafe { sys::nerf_view_angles_grid(num_views as i32, buf.as_mut_ptr(), buf.len() as i32) };
    buf.chunks(2).map(|p| (p[0], p[1])).collect()
}

/// `model::NeRF` (model.rs:133-218): owns the GPU context instead of a `VarStore`.
pub struct NeRF {
    ctx: *mut sys::nerf_ctx,
    pub cfg: sys::nerf_config,
    n_views: usize,
}

/// The decode step of `image_loading::load_image_as_array` (image_loading.rs:7): 8-bit RGBA PNG -> (bytes, width, height).
/// Anything that is not RGBA8 is an error (the reference yields an empty Vec for it).
pub fn load_image_rgba8(path: &str) -> Result<(Vec<u8>, usize, usize), NerfError> {
    let c = std::ffi::CString::new(path).map_err(|_| NerfError { status: -1, message: "path contains NUL".into() })?;
    let (mut w, mut h) = (0i32, 0i32);
    check(std::ptr::null(), unsafe { sys::nerf_load_png_rgba8(c.as_ptr(), std::ptr::null_mut(), 0, &mut w, &mut h) })?;
    let mut out = vec![0u8; (w as usize) * (h as usize) * 4];
    check(std::ptr::null(), unsafe { sys::nerf_load_png_rgba8(c.as_ptr(), out.as_mut_ptr(), out.len() as i64, &mut w, &mut h) })?;
    Ok((out, w as usize, h as usize))
}

/// `image_loading::load_image_as_array` (image_loading.rs:6-24): `[r, g, b, a]` per pixel, each `as f32 / 255.`.
pub fn load_image_as_array(path: &str) -> Result<Vec<[f32; 4]>, NerfError> {
    let (bytes, _, _) = load_image_rgba8(path)?;
    Ok(bytes.chunks_exact(4).map(|p| [p[0] as f32 / 255., p[1] as f32 / 255., p[2] as f32 / 255., p[3] as f32 / 255.]).collect())
}

impl NeRF {
    /// `NeRF::new()` (model.rs:140) with the north-star defaults.
    pub fn new() -> Result<NeRF, NerfError> {
        let mut cfg: sys::nerf_config = unsafe { std::mem::zeroed() };
        check(std::ptr::null(), unsafe { sys::nerf_default_config(&mut cfg) })?;
        NeRF::with_config(cfg, 0)
    }

    pub fn with_config(cfg: sys::nerf_config, device: i32) -> Result<NeRF, NerfError> {
        let mut ctx: *mut sys::nerf_ctx = std::ptr::null_mut();
        check(std::ptr::null(), unsafe { sys::nerf_create(&cfg, device, &mut ctx) })?;
        Ok(NeRF { ctx, cfg, n_views: 0 })
    }

    pub fn set_images(&mut self, imgs: &Vec<Vec<[f32; 4]>>) -> Result<(), NerfError> {
        let flat: Vec<f32> = imgs.iter().flatten().flatten().copied().collect();
        self.n_views = imgs.len();
        check(self.ctx, unsafe { sys::nerf_set_images(self.ctx, flat.as_ptr(), imgs.len() as i32) })
    }

    /// Residency from the decoded RGBA8 bytes of `load_image_rgba8` (4 B/pixel on the device; the sampler's gold gather
    /// performs the `as f32 / 255.` of image_loading.rs:13-18, bit for bit).
    pub fn set_images_rgba8(&mut self, imgs: &Vec<Vec<u8>>) -> Result<(), NerfError> {
        let flat: Vec<u8> = imgs.iter().flatten().copied().collect();
        self.n_views = imgs.len();
        check(self.ctx, unsafe { sys::nerf_set_images_rgba8(self.ctx, flat.as_ptr(), imgs.len() as i32) })
    }

    /// `log_screen_coords` + `log_query_distances` (logging.rs:13-39) for the resident batch, computed on the device:
    /// (screen_x [W], screen_y [H], t [2000]) bucket counts, ready for `log_as_hist` (logging.rs:266-283).
    pub fn log_batch_histograms(&self) -> Result<(Vec<f64>, Vec<f64>, Vec<f64>), NerfError> {
        let (mut sx, mut sy, mut t) = (vec![0f64; self.cfg.image_w as usize], vec![0f64; self.cfg.image_h as usize], vec![0f64; 2000]);
        let mut m: sys::nerf_metrics = unsafe { std::mem::zeroed() };
        m.screen_x = sx.as_mut_ptr();
        m.screen_y = sy.as_mut_ptr();
        m.t_hist = t.as_mut_ptr();
        check(self.ctx, unsafe { sys::nerf_log_metrics(self.ctx, &m) })?;
        Ok((sx, sy, t))
    }

    /// `log_query_points_as_maps` (logging.rs:41-107): the yx / zx / yz occupancy maps, 100x100 0x00RRGGBB each.
    pub fn log_query_point_maps(&self) -> Result<[Vec<u32>; 3], NerfError> {
        let mut maps = [vec![0u32; 10000], vec![0u32; 10000], vec![0u32; 10000]];
        let mut m: sys::nerf_metrics = unsafe { std::mem::zeroed() };
        m.world_yx = maps[0].as_mut_ptr();
        m.world_zx = maps[1].as_mut_ptr();
        m.world_yz = maps[2].as_mut_ptr();
        check(self.ctx, unsafe { sys::nerf_log_metrics(self.ctx, &m) })?;
        Ok(maps)
    }

    /// `draw_predictions` (display.rs:96-110): the batch's predicted pixels scattered into a WIDTH x HEIGHT back buffer.
    pub fn draw_predictions(&self, backbuffer: &mut [u32]) -> Result<(), NerfError> {
        assert_eq!(backbuffer.len(), (self.cfg.image_w * self.cfg.image_h) as usize);
        let mut m: sys::nerf_metrics = unsafe { std::mem::zeroed() };
        m.prediction = backbuffer.as_mut_ptr();
        check(self.ctx, unsafe { sys::nerf_log_metrics(self.ctx, &m) })
    }

    pub fn set_view_angles(&mut self, view_angles: &Vec<(f32, f32)>) -> Result<(), NerfError> {
        let flat: Vec<f32> = view_angles.iter().flat_map(|(y, p)| [*y, *p]).collect();
        check(self.ctx, unsafe { sys::nerf_set_view_angles(self.ctx, flat.as_ptr(), view_angles.len() as i32) })
    }

    /// `NeRF::predict(query_points [B*3], distances [B])` (model.rs:152-156) -> (colors [R*4], densities [R*S]).
    pub fn predict(&self, query_points: &[f32], distances: &[f32], dirs: Option<&[f32]>) -> Result<(Vec<f32>, Vec<f32>), NerfError> {
        let r = self.cfg.num_rays as usize;
        let s = self.cfg.num_samples as usize;
        let mut out = vec![0f32; r * 4];
        let mut sigma = vec![0f32; r * s];
        check(self.ctx, unsafe {
            sys::nerf_predict_points(self.ctx, query_points.as_ptr(), query_points.len() as i64, distances.as_ptr(),
                                     distances.len() as i64, dirs.map_or(std::ptr::null(), |d| d.as_ptr()), 1,
                                     out.as_mut_ptr(), sigma.as_mut_ptr())
        })?;
        Ok((out, sigma))
    }

    /// The prediction as the reference has it -- a device tensor (main.rs:58): enqueues the forward, no device synchronisation.
    /// `Trainer::step_device` consumes it; `predictions()` fetches pixels and densities when the host draws (main.rs:86-89).
    pub fn predict_device(&self, query_points: &[f32], distances: &[f32], dirs: Option<&[f32]>) -> Result<(), NerfError> {
        check(self.ctx, unsafe {
            sys::nerf_predict_points(self.ctx, query_points.as_ptr(), query_points.len() as i64, distances.as_ptr(),
                                     distances.len() as i64, dirs.map_or(std::ptr::null(), |d| d.as_ptr()), 1,
                                     std::ptr::null_mut(), std::ptr::null_mut())
        })
    }
    /// (colors [R*4], densities [R*S]) of the current batch from the device.
    pub fn predictions(&self) -> Result<(Vec<f32>, Vec<f32>), NerfError> {
        let r = self.cfg.num_rays as usize;
        let s = self.cfg.num_samples as usize;
        let mut out = vec![0f32; r * 4];
        let mut sigma = vec![0f32; r * s];
        check(self.ctx, unsafe { sys::nerf_get_predictions(self.ctx, out.as_mut_ptr(), sigma.as_mut_ptr()) })?;
        Ok((out, sigma))
    }

    /// `NeRF::save` / `NeRF::load` (model.rs:211-217) as flat f32 blobs.
    pub fn weights(&self) -> Result<Vec<f32>, NerfError> {
        let n = unsafe { sys::nerf_num_params(self.ctx) };
        let mut flat = vec![0f32; n as usize];
        check(self.ctx, unsafe { sys::nerf_get_weights(self.ctx, flat.as_mut_ptr(), n) })?;
        Ok(flat)
    }
    pub fn set_weights(&mut self, flat: &[f32]) -> Result<(), NerfError> {
        check(self.ctx, unsafe { sys::nerf_set_weights(self.ctx, flat.as_ptr(), flat.len() as i64) })
    }

    pub fn raw(&self) -> *mut sys::nerf_ctx {
        self.ctx
    }
}

impl Drop for NeRF {
    fn drop(&mut self) {
        unsafe { sys::nerf_destroy(self.ctx) };
    }
}

/// `compositing(&densities, colors, distances)` (model.rs:234-249).
pub fn compositing(model: &NeRF, densities: &[f32], colors: Option<&[f32]>, distances: &[f32], num_rays: usize, num_samples: usize) -> Result<Vec<f32>, NerfError> {
    let mut out = vec![0f32; num_rays * 4];
    check(model.ctx, unsafe {
        sys::nerf_compositing(model.ctx, densities.as_ptr(), colors.map_or(std::ptr::null(), |c| c.as_ptr()), distances.as_ptr(),
                              num_rays as i32, num_samples as i32, out.as_mut_ptr())
    })?;
    Ok(out)
}

/// `model::Trainer` (model.rs:301-347). Adam state lives in the context; lr is `cfg.learning_rate`.
pub struct Trainer;

impl Trainer {
    pub fn new(_model: &NeRF, _lr: f64) -> Trainer {
        Trainer
    }

    /// `Trainer::step` on the device-resident prediction of `NeRF::predict_device`.
    pub fn step_device(&mut self, model: &mut NeRF, gold: &[f32], _iter: &usize) -> Result<f32, NerfError> {
        let mut loss = 0f32;
        check(model.raw(), unsafe { sys::nerf_step(model.raw(), gold.as_ptr(), gold.len() as i64, &mut loss) })?;
        Ok(loss)
    }
    /// `Trainer::step(&predictions [R,4], gold [R*4], &iter) -> f32` (model.rs:311).
    pub fn step(&mut self, model: &mut NeRF, predictions: &[f32], gold: &[f32], _iter: &usize) -> Result<f32, NerfError> {
        if predictions.len() != model.cfg.num_rays as usize * 4 {
            return Err(NerfError { status: -1, message: "predictions must be [NUM_RAYS, LABELS] (model.rs:315)".into() });
        }
        let mut loss = 0f32;
        check(model.ctx, unsafe { sys::nerf_step(model.ctx, gold.as_ptr(), gold.len() as i64, &mut loss) })?;
        Ok(loss)
    }
}

/// `dataset::get_multiview_batch(&imgs, &view_angles)` (dataset.rs:63-139). `next_u64` supplies the
/// randomness the reference draws from `Tensor::randint` (dataset.rs:12,19,88).
pub fn get_multiview_batch(model: &mut NeRF, next_u64: &mut dyn FnMut() -> u64)
    -> Result<(Vec<[usize; 2]>, Vec<Vec<[f32; 3]>>, Vec<Vec<f32>>, Vec<[f32; 4]>), NerfError> {
    let r = model.cfg.num_rays as usize;
    let s = model.cfg.num_samples as usize;
    let v = model.n_views;
    if v == 0 || r % v != 0 {
        return Err(NerfError { status: -1, message: format!("Can't divide {:?} rays evenly among {:?} views", r, v) });
    }
    let mut idx = vec![0i64; 2 * r];
    for i in 0..r {
        idx[2 * i] = (next_u64() % model.cfg.image_h as u64) as i64;
        idx[2 * i + 1] = (next_u64() % model.cfg.image_w as u64) as i64;
    }
    let vi: Vec<i64> = (0..v).map(|_| (next_u64() % v as u64) as i64).collect();
    let mut pts = vec![0f32; r * s * 3];
    let mut t = vec![0f32; r * s];
    let mut gold = vec![0f32; r * 4];
    check(model.ctx, unsafe {
        sys::nerf_get_batch(model.ctx, idx.as_ptr(), vi.as_ptr(), v as i32, std::ptr::null(), 1, next_u64(), pts.as_mut_ptr(),
                            t.as_mut_ptr(), gold.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut())
    })?;
    let indices = (0..r).map(|i| [idx[2 * i] as usize, idx[2 * i + 1] as usize]).collect();
    let query_points = pts.chunks(3 * s).map(|ray| ray.chunks(3).map(|p| [p[0], p[1], p[2]]).collect()).collect();
    let distances = t.chunks(s).map(|c| c.to_vec()).collect();
    let gold = gold.chunks(4).map(|g| [g[0], g[1], g[2], g[3]]).collect();
    Ok((indices, query_points, distances, gold))
}
