"""Host-side mirror of the reference's call surface over the C ABI.

Same names, argument meaning and error behaviour as the Rust functions used at
src/main.rs:57-72 (the reference has no plugin interface; these ARE its hot-path API):

  get_view_angles(n)                              image_loading.rs:67-80
  NeRF(config)          -> NeRF::new              model.rs:140-150
  NeRF.predict(points, distances[, dirs])         model.rs:152-209
  compositing(model, densities, colors, deltas)   model.rs:234-249
  Trainer(model, lr)    -> Trainer::new           model.rs:306-309
  Trainer.step(predictions, gold, iter)           model.rs:311-325
  get_multiview_batch(model, imgs, view_angles)   dataset.rs:63-139
  NeRF.save / NeRF.load                           model.rs:211-217
  load_image_as_array / get_image_paths           image_loading.rs:6-54

Where the reference panics (assert_eq!/unwrap), these raise NerfError. Arrays are numpy
float32/int64 on the HOST, like the Vecs the reference passes; all device work happens
inside libnerf_b200.so. No torch on this path.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import NerfConfig, NerfError

T_FAR = 2.0  # ray_sampling.rs:12


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def default_config(**over):
    """North-star defaults (800x800, 4096x64, W=256, L=10/4, skip 5, RGB head)."""
    cfg = NerfConfig()
    _check(None, _lib.load().nerf_default_config(ctypes.byref(cfg)))
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def as_shipped_config(**over):
    """The reference exactly as shipped: 128x128, 84x64, W=100, raw xyz, colours (s,s,s,1)."""
    cfg = NerfConfig()
    _check(None, _lib.load().nerf_config_as_shipped(ctypes.byref(cfg)))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def _check(handle, status):
    if status != _lib.NERF_OK:
        lib = _lib.load()
        msg = lib.nerf_strerror(status).decode()
        if handle is not None:
            detail = lib.nerf_last_error(handle).decode()
            if detail:
                msg = f"{msg}: {detail}"
        raise NerfError(status, msg)


def get_view_angles(num_views):
    """image_loading.rs:67-80 -> float32 [2n(n+1), 2] of (yaw, pitch)."""
    out = np.empty((2 * num_views * (num_views + 1), 2), dtype=np.float32)
    _check(None, _lib.load().nerf_view_angles_grid(num_views, _ptr(out), out.size))
    return out


def load_image_rgba8(path):
    """The decode step of load_image_as_array (image_loading.rs:7): 8-bit RGBA PNG -> uint8 [H, W, 4].
    Like the reference, anything that is not RGBA8 is refused (it would yield an empty Vec there)."""
    lib = _lib.load()
    w, h = ctypes.c_int32(), ctypes.c_int32()
    _check(None, lib.nerf_load_png_rgba8(str(path).encode(), None, 0, ctypes.byref(w), ctypes.byref(h)))
    out = np.empty((h.value, w.value, 4), dtype=np.uint8)
    _check(None, lib.nerf_load_png_rgba8(str(path).encode(), _ptr(out), out.nbytes, ctypes.byref(w), ctypes.byref(h)))
    return out


def load_image_as_array(path):
    """image_loading.rs:6-24 -> float32 [H*W, 4], each channel `as f32 / 255.`."""
    return (load_image_rgba8(path).reshape(-1, 4).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def get_image_paths(directory, start, end, step):
    """image_loading.rs:37-54: `{dir}/image-{i}.png` for i in (start..end).step_by(step), same asserts."""
    assert start < end and (end - start) % step == 0 and (end - start) // step > 0
    return [f"{directory}/image-{i}.png" for i in range(start, end, step)]


class NeRF:
    """NeRF::new (model.rs:140-150) + the context that replaces the VarStore/autograd tape."""

    def __init__(self, config=None, device=0):
        self.lib = _lib.load()
        self.cfg = config if config is not None else default_config()
        self.cfg.struct_size = ctypes.sizeof(NerfConfig)
        h = ctypes.c_void_p()
        status = self.lib.nerf_create(ctypes.byref(self.cfg), device, ctypes.byref(h))
        if status != _lib.NERF_OK:
            raise NerfError(status, self.lib.nerf_strerror(status).decode())
        self.h = h
        self.num_rays, self.num_points = self.cfg.num_rays, self.cfg.num_samples
        self.batch_size = self.num_rays * self.num_points
        self.has_dirs = self.cfg.dir_freqs >= 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.nerf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters (VarStore surface)
    @property
    def num_params(self):
        return int(self.lib.nerf_num_params(self.h))

    def set_weights(self, flat):
        flat = _f32(flat)
        _check(self.h, self.lib.nerf_set_weights(self.h, _ptr(flat), flat.size))

    def get_weights(self):
        out = np.empty(self.num_params, dtype=np.float32)
        _check(self.h, self.lib.nerf_get_weights(self.h, _ptr(out), out.size))
        return out

    def get_grads(self):
        out = np.empty(self.num_params, dtype=np.float32)
        _check(self.h, self.lib.nerf_get_grads(self.h, _ptr(out), out.size))
        return out

    def get_adam_state(self):
        m = np.empty(self.num_params, dtype=np.float32)
        v = np.empty(self.num_params, dtype=np.float32)
        step = ctypes.c_int64()
        _check(self.h, self.lib.nerf_get_adam_state(self.h, _ptr(m), _ptr(v), m.size, ctypes.byref(step)))
        return m, v, step.value

    def set_adam_state(self, m, v, step):
        m, v = _f32(m), _f32(v)
        _check(self.h, self.lib.nerf_set_adam_state(self.h, _ptr(m), _ptr(v), m.size, step))

    def save(self, path, with_adam=True):
        """NeRF::save (model.rs:211-213). `*.ot` writes a tch VarStore checkpoint the reference can load (checkpoint.py);
        anything else a flat .npz. Unlike the reference the Adam state travels too (the .ot file loses it)."""
        m, v, step = self.get_adam_state()
        if str(path).endswith(".ot"):
            from . import checkpoint
            checkpoint.save_varstore(path, self.get_weights(), checkpoint.layer_dims(self.cfg), adam=(m, v, step) if with_adam else None)
            return
        np.savez(path, weights=self.get_weights(), adam_m=m, adam_v=v, step=np.int64(step))

    def load(self, path):
        """NeRF::load (model.rs:215-217): a tch `.ot` VarStore file (as written by the reference or by save()) or the .npz."""
        if str(path).endswith(".ot"):
            from . import checkpoint
            try:
                flat, adam = checkpoint.load_varstore(path, checkpoint.layer_dims(self.cfg))
            except ValueError as e:
                raise NerfError(_lib.ERR_INVALID_ARG, str(e))
            self.set_weights(flat)
            if adam is not None:
                self.set_adam_state(*adam)
            return
        z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        self.set_weights(z["weights"])
        self.set_adam_state(z["adam_m"], z["adam_v"], int(z["step"]))

    # ---- dataset residency
    def set_images(self, imgs):
        imgs = _f32(imgs)
        assert imgs.ndim == 3 and imgs.shape[1] == self.cfg.image_w * self.cfg.image_h and imgs.shape[2] == 4
        _check(self.h, self.lib.nerf_set_images(self.h, _ptr(imgs), imgs.shape[0]))
        self.n_views = imgs.shape[0]

    def set_images_rgba8(self, imgs_u8):
        """Residency from RGBA8 bytes [V, H*W, 4] (or [V, H, W, 4]): 4 B/pixel on the device; the gold gather divides by 255."""
        a = np.ascontiguousarray(imgs_u8, dtype=np.uint8).reshape(len(imgs_u8), -1, 4)
        assert a.shape[1] == self.cfg.image_w * self.cfg.image_h
        _check(self.h, self.lib.nerf_set_images_rgba8(self.h, _ptr(a), a.shape[0]))
        self.n_views = a.shape[0]

    def set_view_angles(self, view_angles):
        va = _f32(view_angles)
        _check(self.h, self.lib.nerf_set_view_angles(self.h, _ptr(va), va.shape[0]))
        self.n_angles = va.shape[0]

    # ---- hot path
    def get_batch(self, indices=None, view_index=None, n_picks=None, jitter=None, randomize=True, seed=0,
                  want=("points", "t", "gold", "dirs", "indices")):
        """dataset::get_multiview_batch body (dataset.rs:72-138) with caller-supplied or Philox randomness.
        Returns a dict of the requested host arrays; the batch also stays resident for predict()."""
        r, s = self.num_rays, self.num_points
        self._pred_gen = getattr(self, "_pred_gen", 0) + 1   # (a new batch invalidates Prediction handles)
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int64).reshape(r, 2)
        vi = None if view_index is None else np.ascontiguousarray(view_index, dtype=np.int64)
        if n_picks is None:
            n_picks = len(vi) if vi is not None else 1
        jit = None if jitter is None else _f32(jitter).reshape(r, s)
        out = {}
        if "points" in want:
            out["points"] = np.empty((r, s, 3), dtype=np.float32)
        if "t" in want:
            out["t"] = np.empty((r, s), dtype=np.float32)
        if "gold" in want:
            out["gold"] = np.empty((r, 4), dtype=np.float32)
        if "dirs" in want:
            out["dirs"] = np.empty((r, 3), dtype=np.float32)
        if "indices" in want:
            out["indices"] = np.empty((r, 2), dtype=np.int64)
        _check(self.h, self.lib.nerf_get_batch(self.h, _ptr(idx), _ptr(vi), n_picks, _ptr(jit), 1 if randomize else 0, seed,
                                              _ptr(out.get("points")), _ptr(out.get("t")), _ptr(out.get("gold")),
                                              _ptr(out.get("dirs")), _ptr(out.get("indices"))))
        return out

    def predict(self, query_points=None, distances=None, dirs=None, train=True, want_sigma=True, lazy=False):
        """NeRF::predict (model.rs:152-209): (pixels [R,4], densities [R,S]).

        With arrays: the literal signature -- flat query_points [B*3], distances [B] (t values),
        dirs [R*3] if the config has a direction input. Without: runs on the resident batch.
        lazy=True: the prediction stays on the device, like the Tensor the reference's predict returns (main.rs:58): the
        call enqueues the forward and returns (Prediction, None) without a device synchronisation; Trainer.step takes the
        handle, `.numpy()` / `.densities()` fetch the values when the host wants them (draw_predictions, main.rs:86-89)."""
        r, s = self.num_rays, self.num_points
        self._pred_gen = getattr(self, "_pred_gen", 0) + 1
        out = None if lazy else np.empty((r, 4), dtype=np.float32)
        sig = np.empty((r, s), dtype=np.float32) if (want_sigma and not lazy) else None
        if query_points is None:
            _check(self.h, self.lib.nerf_predict(self.h, 1 if train else 0, _ptr(out), _ptr(sig)))
        else:
            qp, di, dr = _f32(query_points), _f32(distances), _f32(dirs)
            if qp.ndim != 1 or di.ndim != 1:
                raise NerfError(_lib.ERR_INVALID_ARG, "predict expects 1-D tensors (model.rs:162-163)")
            _check(self.h, self.lib.nerf_predict_points(self.h, _ptr(qp), qp.size, _ptr(di), di.size, _ptr(dr), 1 if train else 0,
                                                        _ptr(out), _ptr(sig)))
        if lazy:
            return Prediction(self), None
        return out, sig

    def get_predictions(self, want_pixels=True, want_sigma=False):
        """Pixels [R,4] and / or densities [R,S] of the current batch from the device (nerf_get_predictions)."""
        out = np.empty((self.num_rays, 4), dtype=np.float32) if want_pixels else None
        sig = np.empty((self.num_rays, self.num_points), dtype=np.float32) if want_sigma else None
        _check(self.h, self.lib.nerf_get_predictions(self.h, _ptr(out), _ptr(sig)))
        return out, sig

    def render(self, yaw, pitch, y0=0, y1=None, randomize=False, seed=0, packed=False):
        """Full-frame rows [y0,y1) at (yaw,pitch): the commented draw_valid_predictions (display.rs:55-94)."""
        w, h = self.cfg.image_w, self.cfg.image_h
        y1 = h if y1 is None else y1
        rgba = np.empty((y1 - y0, w, 4), dtype=np.float32)
        pk = np.empty((y1 - y0, w), dtype=np.uint32) if packed else None
        _check(self.h, self.lib.nerf_render(self.h, yaw, pitch, y0, y1, 1 if randomize else 0, seed, _ptr(rgba), _ptr(pk)))
        return (rgba, pk) if packed else rgba

    def render_sharded(self, yaw, pitch, randomize=False, seed=0, packed=False):
        """Full frame with the rows sharded over the data-parallel ranks (comm_init_rank); one all-gather at the end."""
        w, h = self.cfg.image_w, self.cfg.image_h
        rgba = np.empty((h, w, 4), dtype=np.float32)
        pk = np.empty((h, w), dtype=np.uint32) if packed else None
        _check(self.h, self.lib.nerf_render_sharded(self.h, yaw, pitch, 1 if randomize else 0, seed, _ptr(rgba), _ptr(pk)))
        return (rgba, pk) if packed else rgba

    def log_metrics(self, densities=True, prediction=True):
        """The per-batch TensorBoard projections of src/logging.rs + draw_predictions (display.rs:96-110), computed on the
        device from the resident batch. Returns a dict: screen_x/screen_y/t (f64 counts, log_screen_coords :13-25,
        log_query_distances :27-39), world_yx/zx/yz (u32 [100,100], log_query_points_as_maps :41-107), and after predict()
        density_x/y/z (f64 [2000], log_densities :109-134), density_yx/zx/yz (log_density_maps :136-195), prediction [H,W]."""
        w, h = self.cfg.image_w, self.cfg.image_h
        out = {"screen_x": np.empty(w, np.float64), "screen_y": np.empty(h, np.float64), "t": np.empty(2000, np.float64)}
        for k in ("world_yx", "world_zx", "world_yz"):
            out[k] = np.empty((100, 100), np.uint32)
        if densities:
            for k in ("density_x", "density_y", "density_z"):
                out[k] = np.empty(2000, np.float64)
            for k in ("density_yx", "density_zx", "density_yz"):
                out[k] = np.empty((100, 100), np.uint32)
        if prediction:
            out["prediction"] = np.empty((h, w), np.uint32)
        m = _lib.NerfMetrics()
        for name, _ in _lib.NerfMetrics._fields_:
            key = "t" if name == "t_hist" else name
            setattr(m, name, out[key].ctypes.data if key in out else None)
        _check(self.h, self.lib.nerf_log_metrics(self.h, ctypes.byref(m)))
        return out

    def train_iter(self, seed):
        self._pred_gen = getattr(self, "_pred_gen", 0) + 1
        _check(self.h, self.lib.nerf_train_iter(self.h, seed))

    def last_loss(self):
        v = ctypes.c_float()
        _check(self.h, self.lib.nerf_last_loss(self.h, ctypes.byref(v)))
        return v.value

    def sync(self):
        _check(self.h, self.lib.nerf_sync(self.h))

    # ---- data parallel
    @staticmethod
    def comm_unique_id():
        buf = (ctypes.c_uint8 * 128)()
        _check(None, _lib.load().nerf_comm_unique_id(buf))
        return bytes(buf)

    def comm_init_rank(self, uid, rank, nranks):
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(uid)
        _check(self.h, self.lib.nerf_comm_init_rank(self.h, buf, rank, nranks))

    def comm_destroy(self):
        """Leave the communicator (collective: every rank calls it; peers wait for each other's last exchange to finish)."""
        _check(self.h, self.lib.nerf_comm_destroy(self.h))

    # ---- measurement
    def timer_start(self):
        _check(self.h, self.lib.nerf_timer_start(self.h))

    def timer_stop(self):
        v = ctypes.c_float()
        _check(self.h, self.lib.nerf_timer_stop(self.h, ctypes.byref(v)))
        return v.value

    def profile(self, on):
        _check(self.h, self.lib.nerf_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        cap = 64
        names = ctypes.create_string_buffer(32 * cap)
        ms = (ctypes.c_float * cap)()
        n = (ctypes.c_int32 * cap)()
        cnt = ctypes.c_int32()
        _check(self.h, self.lib.nerf_profile_read(self.h, names, ms, n, cap, ctypes.byref(cnt)))
        out = {}
        for i in range(min(cnt.value, cap)):
            nm = names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode()
            out[nm] = (float(ms[i]), int(n[i]))
        return out

    @property
    def launch_count(self):
        return int(self.lib.nerf_launch_count(self.h))

    def flush_l2(self):
        _check(self.h, self.lib.nerf_flush_l2(self.h))

    STAGES = {"sample": 0, "composite_fwd": 1, "composite_bwd": 2, "adam": 3}

    def bench_stage(self, stage, num_rays, num_samples, iters=10):
        """ms per launch of one HBM-bound stage kernel alone on synthetic device-resident inputs (nerf_debug_bench_stage)."""
        v = ctypes.c_float()
        _check(self.h, self.lib.nerf_debug_bench_stage(self.h, self.STAGES[stage], num_rays, num_samples, iters, ctypes.byref(v)))
        return v.value

    def debug_read_panel(self, area, tile, slot):
        if area == 2:
            out = np.empty((128, 8), dtype=np.uint32)
        else:
            out = np.empty(8192, dtype=np.uint16)
        _check(self.h, self.lib.nerf_debug_read_panel(self.h, area, tile, slot, _ptr(out)))
        return out


def compositing(model, densities, colors, distances):
    """compositing (model.rs:234-249): densities [R,S], colors [R,S,4] (None = (s,s,s,1)), distances [R,S] = deltas."""
    d = _f32(densities)
    r, s = d.shape
    c = None if colors is None else _f32(colors).reshape(r, s, 4)
    dl = _f32(distances).reshape(r, s)
    out = np.empty((r, 4), dtype=np.float32)
    _check(model.h, model.lib.nerf_compositing(model.h, _ptr(d), _ptr(c), _ptr(dl), r, s, _ptr(out)))
    return out


class Prediction:
    """Device-resident result of NeRF.predict(..., lazy=True): the counterpart of the Tensor the reference's predict returns and
    Trainer::step consumes (main.rs:58 -> :72). Valid until the model's next predict."""

    def __init__(self, model):
        self.model = model
        self.gen = model._pred_gen
        self.shape = (model.num_rays, 4)

    def _current(self):
        if self.model.h is None or self.gen != getattr(self.model, "_pred_gen", 0):
            raise NerfError(_lib.ERR_STATE, "stale prediction handle: the model has run another predict since")

    def numpy(self):
        """pixels [R,4] on the host"""
        self._current()
        return self.model.get_predictions(True, False)[0]

    def densities(self):
        """densities [R,S] on the host"""
        self._current()
        return self.model.get_predictions(False, True)[1]

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)


class Trainer:
    """Trainer::new / Trainer::step (model.rs:301-325). Adam state lives in the model's context;
    lr is fixed at context creation (cli.rs:64-65), so `lr` here must match the config."""

    def __init__(self, model, lr=None):
        self.model = model
        if lr is not None and abs(lr - model.cfg.learning_rate) > 1e-12 * max(1.0, abs(lr)):
            if abs(np.float32(lr) - np.float32(model.cfg.learning_rate)) > 0:
                raise NerfError(_lib.ERR_INVALID_ARG, "Trainer lr differs from the context's learning_rate")

    def step(self, predictions, gold, it=0, want_loss=True):
        """predictions: the array predict() returned (kept for signature parity: the tape is in the
        context); gold: flat [R*4] (model.rs:316). Returns the loss as a host float (model.rs:324)."""
        m = self.model
        if isinstance(predictions, Prediction):
            if predictions.model is not m:
                raise NerfError(_lib.ERR_INVALID_ARG, "the prediction belongs to another model")
            predictions._current()
        elif predictions is not None and tuple(np.shape(predictions)) != (m.num_rays, 4):
            raise NerfError(_lib.ERR_INVALID_ARG, "predictions must be [NUM_RAYS, LABELS] (model.rs:315)")
        g = _f32(gold)
        if g is not None and g.ndim != 1:
            raise NerfError(_lib.ERR_INVALID_ARG, "gold must be a flat [NUM_RAYS*LABELS] tensor (model.rs:316)")
        loss = ctypes.c_float()
        _check(m.h, m.lib.nerf_step(m.h, _ptr(g), 0 if g is None else g.size, ctypes.byref(loss) if want_loss else None))
        return loss.value if want_loss else None


def get_multiview_batch(model, imgs=None, view_angles=None, rng=None):
    """dataset::get_multiview_batch(&imgs, &view_angles) (dataset.rs:63-139).

    imgs/view_angles are made resident on first use (or pass None if already set). Pixel and view
    picks come from `rng` (numpy Generator; the reference uses tch randint) and the depth jitter
    from the device Philox stream seeded by rng. Returns (indices, query_points, distances, gold)
    with the reference's shapes [R][2], [R][S][3], [R][S], [R][4]."""
    if imgs is not None:
        model.set_images(imgs)
    if view_angles is not None:
        model.set_view_angles(view_angles)
    rng = rng if rng is not None else np.random.default_rng()
    v = model.n_views
    r = model.num_rays
    if r % v != 0:
        raise NerfError(_lib.ERR_INVALID_ARG, f"Can't divide {r} rays evenly among {v} views, got extra {r % v}")
    idx = np.stack([rng.integers(0, model.cfg.image_h, r), rng.integers(0, model.cfg.image_w, r)], axis=1).astype(np.int64)
    vi = rng.integers(0, v, v).astype(np.int64)
    b = model.get_batch(idx, vi, v, None, True, int(rng.integers(0, 2 ** 62)), want=("points", "t", "gold"))
    return idx, b["points"], b["t"], b["gold"]
