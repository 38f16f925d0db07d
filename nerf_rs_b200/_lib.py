"""ctypes loader for libnerf_b200.so (the C ABI in include/nerf_b200.h).

There is no CPU fallback: if the library is missing or no sm_100 GPU is present the
product raises. ``load()`` only dlopen()s; nothing here touches a GPU until nerf_create.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnerf_b200.so")
_lib = None

NERF_OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_COMM, ERR_STATE, ERR_NO_DEVICE = -1, -2, -3, -4, -5, -6
DEPTH_REFERENCE, DEPTH_STRATIFIED = 0, 1
MLP_TCGEN05, MLP_SIMT, MLP_SIMT_FP32, MLP_TCGEN05_SS = 0, 1, 2, 3


class NerfConfig(ctypes.Structure):
    """nerf_config (include/nerf_b200.h) -- runtime form of the reference's consts
    (src/model.rs:7-13, src/ray_sampling.rs:7-12)."""
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("image_w", ctypes.c_int32), ("image_h", ctypes.c_int32),
        ("num_rays", ctypes.c_int32), ("num_samples", ctypes.c_int32),
        ("hidden", ctypes.c_int32), ("xyz_freqs", ctypes.c_int32), ("dir_freqs", ctypes.c_int32),
        ("skip_layer", ctypes.c_int32), ("use_rgb_head", ctypes.c_int32), ("sigma_relu", ctypes.c_int32),
        ("depth_mode", ctypes.c_int32), ("mlp_impl", ctypes.c_int32), ("max_rays_per_launch", ctypes.c_int32),
        ("learning_rate", ctypes.c_float), ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float),
        ("deterministic_grads", ctypes.c_int32),
    ]


class NerfError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"nerf_b200 error {status}: {message}")
        self.status = status


vp = ctypes.c_void_p
i32, i64, u64, f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float
P = ctypes.POINTER

# name -> (restype, argtypes); exactly the declarations of include/nerf_b200.h + nerf_b200_debug.h
class NerfMetrics(ctypes.Structure):
    """nerf_metrics (include/nerf_b200.h): host output pointers of nerf_log_metrics, NULL = skip."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("screen_x", "screen_y", "t_hist", "world_yx", "world_zx", "world_yz", "density_x",
                                                 "density_y", "density_z", "density_yx", "density_zx", "density_yz", "prediction")]


SIGNATURES = {
    "nerf_abi_version": (ctypes.c_int, []),
    "nerf_default_config": (ctypes.c_int, [P(NerfConfig)]),
    "nerf_config_as_shipped": (ctypes.c_int, [P(NerfConfig)]),
    "nerf_create": (ctypes.c_int, [P(NerfConfig), ctypes.c_int, P(vp)]),
    "nerf_destroy": (ctypes.c_int, [vp]),
    "nerf_last_error": (ctypes.c_char_p, [vp]),
    "nerf_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "nerf_num_params": (i64, [vp]),
    "nerf_set_weights": (ctypes.c_int, [vp, vp, i64]),
    "nerf_get_weights": (ctypes.c_int, [vp, vp, i64]),
    "nerf_get_grads": (ctypes.c_int, [vp, vp, i64]),
    "nerf_get_adam_state": (ctypes.c_int, [vp, vp, vp, i64, P(i64)]),
    "nerf_set_adam_state": (ctypes.c_int, [vp, vp, vp, i64, i64]),
    "nerf_set_images": (ctypes.c_int, [vp, vp, i32]),
    "nerf_set_images_rgba8": (ctypes.c_int, [vp, vp, i32]),
    "nerf_load_png_rgba8": (ctypes.c_int, [ctypes.c_char_p, vp, ctypes.c_int64, vp, vp]),
    "nerf_set_view_angles": (ctypes.c_int, [vp, vp, i32]),
    "nerf_view_angles_grid": (ctypes.c_int, [i32, vp, i32]),
    "nerf_get_batch": (ctypes.c_int, [vp, vp, vp, i32, vp, i32, u64, vp, vp, vp, vp, vp]),
    "nerf_predict": (ctypes.c_int, [vp, i32, vp, vp]),
    "nerf_predict_points": (ctypes.c_int, [vp, vp, i64, vp, i64, vp, i32, vp, vp]),
    "nerf_get_predictions": (ctypes.c_int, [vp, vp, vp]),
    "nerf_compositing": (ctypes.c_int, [vp, vp, vp, vp, i32, i32, vp]),
    "nerf_step": (ctypes.c_int, [vp, vp, i64, P(f32)]),
    "nerf_train_iter": (ctypes.c_int, [vp, u64]),
    "nerf_last_loss": (ctypes.c_int, [vp, P(f32)]),
    "nerf_sync": (ctypes.c_int, [vp]),
    "nerf_render": (ctypes.c_int, [vp, f32, f32, i32, i32, i32, u64, vp, vp]),
    "nerf_render_sharded": (ctypes.c_int, [vp, f32, f32, i32, u64, vp, vp]),
    "nerf_comm_unique_id": (ctypes.c_int, [vp]),
    "nerf_comm_init_rank": (ctypes.c_int, [vp, vp, i32, i32]),
    "nerf_comm_destroy": (ctypes.c_int, [vp]),
    "nerf_timer_start": (ctypes.c_int, [vp]),
    "nerf_timer_stop": (ctypes.c_int, [vp, P(f32)]),
    "nerf_profile_enable": (ctypes.c_int, [vp, i32]),
    "nerf_profile_read": (ctypes.c_int, [vp, vp, vp, vp, i32, P(i32)]),
    "nerf_launch_count": (i64, [vp]),
    "nerf_flush_l2": (ctypes.c_int, [vp]),
    "nerf_log_metrics": (ctypes.c_int, [vp, P(NerfMetrics)]),
    "nerf_debug_plan": (ctypes.c_int, [P(NerfConfig), i32, vp, P(i32), vp, P(i32), vp, P(i32), vp, P(i32), P(i32)]),
    "nerf_debug_plan_biases": (ctypes.c_int, [P(NerfConfig), vp, P(i32)]),
    "nerf_debug_trace": (ctypes.c_int, [vp, i32, vp]),
    "nerf_debug_lane_plan": (ctypes.c_int, [vp, i32, vp, vp, vp, vp, vp, vp]),
    "nerf_debug_host_pose": (ctypes.c_int, [f32, f32, vp, vp, vp]),
    "nerf_debug_ts_plan": (ctypes.c_int, [P(NerfConfig), i32, vp, P(i32), vp, P(i32), vp, P(i32), P(i32)]),
    "nerf_debug_tc3_stats": (ctypes.c_int, [vp, i32]),
    "nerf_debug_tc3_trace": (ctypes.c_int, [vp, i32]),
    "nerf_debug_check_guards": (ctypes.c_int, [P(i32)]),
    "nerf_debug_read_panel": (ctypes.c_int, [vp, i32, i32, i32, vp]),
    "nerf_debug_wgrad_partition": (ctypes.c_int, [P(NerfConfig), i32, i64, vp, vp, P(i32)]),
    "nerf_debug_wgrad_marks": (ctypes.c_int, [vp, vp, i32]),
    "nerf_debug_bench_stage": (ctypes.c_int, [vp, i32, i32, i32, i32, P(f32)]),
}


def load():
    """dlopen the in-tree library; raise (never fall back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -m nerf_rs_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
