"""CPU restatement of the reference's per-batch logging projections -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/src/logging.rs line by line (plain loops over the batch, like the reference) and
src/display.rs:37-52,96-110 for the 0x00RRGGBB packing and the prediction back buffer. Parity unpinned: the reference has no
test for any of these functions; the restatement is checked by closed forms in tests/test_metrics.py.
"""
import numpy as np


def _as_usize(v):
    """Rust `f32 as usize`: saturating, NaN -> 0."""
    v = np.float32(v)
    if not (v == v) or v <= 0:
        return 0
    return int(min(float(v), 9.0e18))


def _as_u8(v):
    v = np.float32(v)
    if not (v == v):
        return 0
    return int(min(max(float(v), 0.0), 255.0))


def prediction_array_as_u32(rgba):
    """display.rs:46-52 + from_u8_rgb :37-40."""
    r, g, b = (_as_u8(np.float32(c) * np.float32(255.0)) for c in rgba[:3])
    return (r << 16) | (g << 8) | b


def log_screen_coords(indices, width, height):
    """logging.rs:13-25. `for [x, y] in indices` binds x to the FIRST stored element (which dataset.rs stores as y)."""
    sy = np.zeros(height, np.float64)
    sx = np.zeros(width, np.float64)
    for x, y in np.asarray(indices):
        sy[y] += 1.0
        sx[x] += 1.0
    return sx, sy


def log_query_distances(distances):
    """logging.rs:27-39."""
    b = np.zeros(2000, np.float64)
    for t in np.asarray(distances, np.float32).reshape(-1):
        b[_as_usize(np.floor(np.float32(500.0) * t))] += 1.0
    return b


def _cells(p):
    wx, wy, wz = (np.float32(c) for c in p)
    y = _as_usize(np.floor(np.float32(50.0) * (wy + np.float32(1.0))))
    x = _as_usize(np.floor(np.float32(50.0) * (wx + np.float32(1.0))))
    z = _as_usize(np.floor(np.float32(25.0) * (wz + np.float32(1.0))))
    return y, x, z


def log_query_points_as_maps(query_points):
    """logging.rs:41-107 -> (yx, zx, yz) u32 [100,100]."""
    yx, zx, yz = (np.zeros(10000, np.uint32) for _ in range(3))
    white = prediction_array_as_u32([1.0, 1.0, 1.0, 1.0])
    for p in np.asarray(query_points, np.float32).reshape(-1, 3):
        y, x, z = _cells(p)
        yx[min(y * 100 + x, 9999)] = white
        zx[min(z * 100 + x, 9999)] = white
        yz[min(y * 100 + z, 9999)] = white
    return yx.reshape(100, 100), zx.reshape(100, 100), yz.reshape(100, 100)


def log_densities(query_points, densities):
    """logging.rs:109-134 -> (x, y, z) f64 [2000]. Cells past 1999 panic in the reference; they are dropped here."""
    bx, by, bz = (np.zeros(2000, np.float64) for _ in range(3))
    for p, d in zip(np.asarray(query_points, np.float32).reshape(-1, 3), np.asarray(densities, np.float32).reshape(-1)):
        y = _as_usize(np.floor(np.float32(500.0) * (p[1] + np.float32(1.0))))
        x = _as_usize(np.floor(np.float32(500.0) * (p[0] + np.float32(1.0))))
        z = _as_usize(np.floor(np.float32(500.0) * (p[2] + np.float32(1.0))))
        if y < 2000:
            by[y] += float(d)
        if x < 2000:
            bx[x] += float(d)
        if z < 2000:
            bz[z] += float(d)
    return bx, by, bz


def log_density_maps(query_points, densities):
    """logging.rs:136-195 -> (yx, zx, yz): sequential overwrite, so the last sample landing in a cell wins. Out-of-range
    cells panic in the reference; they are dropped here."""
    yx, zx, yz = (np.zeros(10000, np.uint32) for _ in range(3))
    for p, d in zip(np.asarray(query_points, np.float32).reshape(-1, 3), np.asarray(densities, np.float32).reshape(-1)):
        y, x, z = _cells(p)
        dc = max(np.float32(d), np.float32(0.0))
        c = prediction_array_as_u32([dc, dc, dc, 1.0])
        if y * 100 + x < 10000:
            yx[y * 100 + x] = c
        if z * 100 + x < 10000:
            zx[z * 100 + x] = c
        if y * 100 + z < 10000:
            yz[y * 100 + z] = c
    return yx.reshape(100, 100), zx.reshape(100, 100), yz.reshape(100, 100)


def draw_predictions(indices, predictions, width, height):
    """display.rs:96-110: backbuffer[y * WIDTH + x] = 0RGB(pred[0..3]) in batch order."""
    bb = np.zeros(width * height, np.uint32)
    for (y, x), pr in zip(np.asarray(indices), np.asarray(predictions, np.float32)):
        bb[y * width + x] = prediction_array_as_u32([pr[0], pr[1], pr[2], 1.0])
    return bb.reshape(height, width)
