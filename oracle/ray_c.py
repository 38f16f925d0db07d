"""ctypes binding of ``ray_oracle.c`` -- TEST INFRASTRUCTURE ONLY.

``build()`` compiles the C restatement with gcc (``-O2 -ffp-contract=off``);
``lib()`` loads it. No part of the product imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libray_oracle.so")
_SRC = os.path.join(_HERE, "ray_oracle.c")
_lib = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_i64_p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_get_view_angles.restype = ctypes.c_int
        _lib.oracle_get_multiview_batch.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _ip(a):
    return a.ctypes.data_as(c_i64_p)


def rotate_yaw(v, angle):
    v = np.ascontiguousarray(v, dtype=np.float32)
    out = np.empty(3, dtype=np.float32)
    lib().oracle_rotate_yaw(_fp(v), ctypes.c_float(angle), _fp(out))
    return out


def rotate_pitch(v, angle):
    v = np.ascontiguousarray(v, dtype=np.float32)
    out = np.empty(3, dtype=np.float32)
    lib().oracle_rotate_pitch(_fp(v), ctypes.c_float(angle), _fp(out))
    return out


def pitch_matrix(angle):
    out = np.empty((3, 3), dtype=np.float32)
    lib().oracle_pitch_matrix(ctypes.c_float(angle), _fp(out))
    return out


def screen_to_world(x, y, width, height):
    out = np.empty(3, dtype=np.float32)
    lib().oracle_screen_to_world(ctypes.c_float(x), ctypes.c_float(y), ctypes.c_float(width), ctypes.c_float(height), _fp(out))
    return out


def sample_points_along_ray_and_rotate(frm, to, yaw, pitch, num_samples, u=None):
    frm = np.ascontiguousarray(frm, dtype=np.float32)
    to = np.ascontiguousarray(to, dtype=np.float32)
    pts = np.empty((num_samples, 3), dtype=np.float32)
    loc = np.empty(num_samples, dtype=np.float32)
    up = None
    if u is not None:
        u = np.ascontiguousarray(u, dtype=np.float32)
        up = _fp(u)
    lib().oracle_sample_points_along_ray_and_rotate(_fp(frm), _fp(to), ctypes.c_float(yaw), ctypes.c_float(pitch),
                                                    ctypes.c_int(num_samples), up, _fp(pts), _fp(loc))
    return pts, loc


def sample_rays(indices_yx, num_points, yaw, pitch, u, width, height):
    idx = np.ascontiguousarray(indices_yx, dtype=np.int64)
    n = idx.shape[0]
    pts = np.empty((n, num_points, 3), dtype=np.float32)
    loc = np.empty((n, num_points), dtype=np.float32)
    up = None
    if u is not None:
        u = np.ascontiguousarray(u, dtype=np.float32)
        up = _fp(u)
    lib().oracle_sample_rays_for_screen_coords(_ip(idx), ctypes.c_int(n), ctypes.c_int(num_points), ctypes.c_float(yaw),
                                               ctypes.c_float(pitch), up, ctypes.c_int(width), ctypes.c_int(height),
                                               _fp(pts), _fp(loc))
    return pts, loc


def ray_dirs(indices_yx, yaw, pitch, width, height):
    idx = np.ascontiguousarray(indices_yx, dtype=np.int64)
    n = idx.shape[0]
    out = np.empty((n, 3), dtype=np.float32)
    lib().oracle_ray_dirs_for_screen_coords(_ip(idx), ctypes.c_int(n), ctypes.c_float(yaw), ctypes.c_float(pitch),
                                            ctypes.c_int(width), ctypes.c_int(height), _fp(out))
    return out


def get_view_angles(num_views):
    out = np.empty((2 * num_views * (num_views + 1), 2), dtype=np.float32)
    k = lib().oracle_get_view_angles(ctypes.c_int(num_views), _fp(out))
    assert k == out.shape[0]
    return out


def get_multiview_batch(imgs, view_angles, indices_yx, view_index, num_points, u, width, height):
    imgs = np.ascontiguousarray(imgs, dtype=np.float32)
    va = np.ascontiguousarray(view_angles, dtype=np.float32)
    idx = np.ascontiguousarray(indices_yx, dtype=np.int64)
    vi = np.ascontiguousarray(view_index, dtype=np.int64)
    r = idx.shape[0]
    pts = np.empty((r, num_points, 3), dtype=np.float32)
    loc = np.empty((r, num_points), dtype=np.float32)
    gold = np.empty((r, 4), dtype=np.float32)
    up = None
    if u is not None:
        u = np.ascontiguousarray(u, dtype=np.float32)
        up = _fp(u)
    rc = lib().oracle_get_multiview_batch(_fp(imgs), _fp(va), ctypes.c_int(imgs.shape[0]), _ip(idx), _ip(vi),
                                          ctypes.c_int(r), ctypes.c_int(num_points), up, ctypes.c_int(width),
                                          ctypes.c_int(height), _fp(pts), _fp(loc), _fp(gold))
    if rc != 0:
        raise ValueError("Can't divide rays evenly among views (dataset.rs:73-81)")
    return idx, pts, loc, gold


def philox_uniform(seed, stream, first_index, n):
    out = np.empty(n, dtype=np.float32)
    lib().oracle_philox_uniform(ctypes.c_uint64(seed), ctypes.c_uint32(stream), ctypes.c_uint64(first_index),
                                ctypes.c_int64(n), _fp(out))
    return out
