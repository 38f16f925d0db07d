"""numpy-f32 twin of ``ray_oracle.c`` -- TEST INFRASTRUCTURE ONLY.

Vectorised over rays/samples, same operation order as the scalar C restatement
(numpy never fuses mul+add), so the two must agree bit for bit; the tests check
that. Also holds the build-decision pieces the reference has no code for:
the sinusoidal positional encoding (paper convention, no pi
factor, include-input) and stratified depth sampling.

Reference anchors: src/ray_sampling.rs:20-26, :32-69, :79-93, :96-142, :156-178;
src/image_loading.rs:67-80; src/dataset.rs:63-139.
"""
import ctypes
import ctypes.util

import numpy as np

f32 = np.float32

# Scalar trig comes from the platform libm (cosf/sinf/tanf), like Rust's f32::cos on the
# reference side; numpy's own SIMD float32 sin/cos differ from glibc in the last ulp.
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "tanf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]


def cosf(x):
    return f32(_libm.cosf(float(f32(x))))


def sinf(x):
    return f32(_libm.sinf(float(f32(x))))


def tanf(x):
    return f32(_libm.tanf(float(f32(x))))

HITHER = f32(0.05)
T_FAR = f32(2.0)
FOV = f32(np.pi) / f32(3.0)
UP = np.array([0.0, 1.0, 0.0], dtype=f32)
AT = np.array([0.0, 0.0, 1.0], dtype=f32)
FROM = np.array([0.0, 0.0, -1.0], dtype=f32)
# (T_FAR - HITHER) + HITHER evaluated in f32 (ray_sampling.rs:114); == 2.0 exactly
T_SCALE = f32(f32(T_FAR - HITHER) + HITHER)


def _dot(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def _normalized(a):
    inv = f32(1.0) / np.sqrt(_dot(a, a))
    return a * inv[..., None]


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1).astype(f32)


def _row_mat3_mul(a, b):
    o = np.zeros((3, 3), dtype=f32)
    for i in range(3):
        for j in range(3):
            o[i, j] = (a[i, 0] * b[0, j] + a[i, 1] * b[1, j]) + a[i, 2] * b[2, j]
    return o


def yaw_matrix(angle):
    """3x4 row matrix of rotateYaw (ray_sampling.rs:21-23)."""
    c, s = cosf(angle), sinf(angle)
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0]], dtype=f32)


def pitch_matrix(angle):
    """The Rodrigues matrix rotatePitch rebuilds per point (ray_sampling.rs:34-65)."""
    v = _normalized((AT - FROM).astype(f32))
    u = _normalized(_cross(v, UP))
    ux, uy, uz = u
    cross_mat = np.array([[0, -uz, uy], [uz, 0, -ux], [-uy, ux, 0]], dtype=f32)
    outer = np.array([[ux * ux, ux * uy, ux * uz], [uy * ux, uy * uy, uy * uz], [uz * ux, uz * uy, uz * uz]], dtype=f32)
    c, s = cosf(angle), sinf(angle)
    idc = (np.eye(3, dtype=f32) * c).astype(f32)
    ids = (np.eye(3, dtype=f32) * s).astype(f32)
    imc = (np.eye(3, dtype=f32) - idc).astype(f32)
    return ((idc + _row_mat3_mul(cross_mat, ids)) + _row_mat3_mul(outer, imc)).astype(f32)


def rotate_yaw(v, angle):
    m = yaw_matrix(angle)
    v = np.asarray(v, dtype=f32)
    out = [((m[i, 0] * v[..., 0] + m[i, 1] * v[..., 1]) + m[i, 2] * v[..., 2]) + m[i, 3] for i in range(3)]
    return np.stack(out, axis=-1).astype(f32)


def rotate_pitch(v, angle):
    m = pitch_matrix(angle)
    v = np.asarray(v, dtype=f32)
    out = [(m[0, i] * v[..., 0] + m[1, i] * v[..., 1]) + m[2, i] * v[..., 2] for i in range(3)]
    return np.stack(out, axis=-1).astype(f32)


def screen_to_world(x, y, width, height):
    """ray_sampling.rs:79-93, vectorised over pixels. x = column, y = row (no +0.5)."""
    x = np.asarray(x, dtype=f32)
    y = np.asarray(y, dtype=f32)
    off = tanf(FOV / f32(2.0)) * HITHER
    offset_left = off - f32(2.0) * off * x / f32(width)
    offset_up = off - f32(2.0) * off * y / f32(height)
    view = _normalized((AT - FROM).astype(f32))
    left = _normalized(_cross(view, UP))
    a = view * HITHER
    b = left[None, :] * offset_left[..., None]
    c = UP[None, :] * offset_up[..., None]
    return _normalized(((a[None, :] + b) + c).astype(f32))


def depth_t(u=None, num_samples=None, n_rays=None, mode="reference"):
    """Depths t[R,S].

    reference (ray_sampling.rs:107-125): t = u * 2.0, sorted ascending per ray;
    u=None is the randomize=false branch u = i/S.
    stratified (build decision, north star): t = ((i + u)/S) * 2.0 -- already sorted.
    """
    if u is None:
        i = np.arange(num_samples, dtype=f32)
        base = (i / f32(num_samples)).astype(f32)
        t = np.broadcast_to(base * T_SCALE, (n_rays, num_samples)).astype(f32)
        return t
    u = np.asarray(u, dtype=f32)
    if mode == "reference":
        t = (u * T_SCALE).astype(f32)
        return np.sort(t, axis=-1, kind="stable")
    if mode == "stratified":
        s = u.shape[-1]
        i = np.arange(s, dtype=f32)
        return (((i[None, :] + u) / f32(s)) * T_SCALE).astype(f32)
    raise ValueError(mode)


def sample_rays(indices_yx, num_points, yaw, pitch, u, width, height, mode="reference"):
    """sample_and_rotate_ray_points_for_screen_coords (ray_sampling.rs:156-178).

    Returns (points[n,S,3], t[n,S]). Sorting t before or after forming the points
    is equivalent: p is a function of t alone for a fixed ray."""
    idx = np.asarray(indices_yx)
    to = screen_to_world(idx[:, 1].astype(f32), idx[:, 0].astype(f32), width, height)
    t = depth_t(u, num_points, idx.shape[0], mode)
    p = (FROM[None, None, :] + (to[:, None, :] * t[:, :, None]).astype(f32)).astype(f32)
    p = rotate_pitch(rotate_yaw(p, yaw), pitch)
    return p, t


def ray_dirs(indices_yx, yaw, pitch, width, height):
    idx = np.asarray(indices_yx)
    to = screen_to_world(idx[:, 1].astype(f32), idx[:, 0].astype(f32), width, height)
    return rotate_pitch(rotate_yaw(to, yaw), pitch)


def get_view_angles(num_views):
    """image_loading.rs:67-80 -- repeated f32 addition, 2n(n+1) (yaw,pitch) pairs."""
    out = []
    step = f32(np.pi) / f32(num_views)
    rot_ver, rot_hor = f32(0), f32(0)
    for _ in range(2 * num_views):
        for _ in range(num_views + 1):
            out.append((rot_hor, rot_ver))
            rot_ver = f32(rot_ver + step)
        rot_hor = f32(rot_hor + step)
        rot_ver = f32(0)
    return np.array(out, dtype=f32)


def get_multiview_batch(imgs, view_angles, indices_yx, view_index, num_points, u, width, height, mode="reference"):
    """dataset.rs:63-139 with the randomness supplied by the caller.

    imgs[V,H*W,4]; indices_yx[R,2]; view_index[V] (with replacement); u[R,S] or None.
    Returns (indices, points[R,S,3], t[R,S], gold[R,4], dirs[R,3])."""
    n_views = imgs.shape[0]
    r = indices_yx.shape[0]
    assert r % n_views == 0, "Can't divide rays evenly among views (dataset.rs:73-81)"
    bsz = r // n_views
    pts, ts, gold, dirs = [], [], [], []
    for i, n in enumerate(view_index):
        yaw, pitch = view_angles[n]
        idx = indices_yx[i * bsz:(i + 1) * bsz]
        ub = None if u is None else u[i * bsz:(i + 1) * bsz]
        p, t = sample_rays(idx, num_points, yaw, pitch, ub, width, height, mode)
        pts.append(p)
        ts.append(t)
        gold.append(imgs[n][idx[:, 0] * width + idx[:, 1]])
        dirs.append(ray_dirs(idx, yaw, pitch, width, height))
    return indices_yx, np.concatenate(pts), np.concatenate(ts), np.concatenate(gold).astype(f32), np.concatenate(dirs)


def posenc(x, num_freqs):
    """[x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] per 3-vector, no pi.

    x[..., 3] -> [..., 3 + 6*num_freqs]. num_freqs == 0 returns x (as-shipped raw xyz)."""
    x = np.asarray(x, dtype=f32)
    outs = [x]
    for k in range(num_freqs):
        a = (x * f32(2.0 ** k)).astype(f32)
        outs.append(np.sin(a, dtype=f32))
        outs.append(np.cos(a, dtype=f32))
    return np.concatenate(outs, axis=-1).astype(f32)


# ---- Philox4x32-10, twin of oracle_philox_uniform --------------------------
def philox_uniform(seed, stream, first_index, n):
    idx = (np.uint64(first_index) + np.arange(n, dtype=np.uint64))
    c0 = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    c1 = (idx >> np.uint64(32)).astype(np.uint64)
    c2 = np.full(n, stream, dtype=np.uint64)
    c3 = np.zeros(n, dtype=np.uint64)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m32
        n1 = p1 & m32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m32
        n3 = p0 & m32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return ((c0 >> np.uint64(8)).astype(f32) * f32(1.0 / 16777216.0)).astype(f32)
