"""Dataset ingest: the reference's image loader (src/image_loading.rs:6-54) restated for the drop-in.
CPU: PNG decode (all five scan-line filters) and the `as f32 / 255.` conversion. GPU: RGBA8 residency + the gold gather
with the division fused in is bit-identical to the reference's host-side conversion followed by the f32 gather."""
import os
import struct
import zlib

import numpy as np
import pytest

import nerf_rs_b200 as nb
from oracle import ray_c


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png(path, img, color_type=6, filters=None):
    """Minimal PNG encoder (test helper): 8-bit, non-interlaced; `filters[y]` picks the scan-line filter type 0..4."""
    h, w, ch = img.shape
    raw = bytearray()
    for y in range(h):
        f = (filters[y] if filters is not None else y % 5)
        raw.append(f)
        row = img[y].reshape(-1).astype(int)
        up = img[y - 1].reshape(-1).astype(int) if y else np.zeros(w * ch, int)
        for x in range(w * ch):
            a = row[x - ch] if x >= ch else 0
            b = up[x]
            c = up[x - ch] if x >= ch else 0
            pred = [0, a, b, (a + b) >> 1, _paeth(a, b, c)][f]
            raw.append((row[x] - pred) & 0xFF)
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    comp = zlib.compress(bytes(raw), 6)
    half = len(comp) // 2     # two IDAT chunks, like real encoders emit
    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, color_type, 0, 0, 0)) +
                 chunk(b"IDAT", comp[:half]) + chunk(b"IDAT", comp[half:]) + chunk(b"IEND", b""))


def test_png_rgba8_decode_all_filters(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (23, 17, 4), dtype=np.uint8)
    p = str(tmp_path / "image-0.png")
    write_png(p, img)
    assert np.array_equal(nb.load_image_rgba8(p), img)
    arr = nb.load_image_as_array(p)                       # image_loading.rs:6-24
    assert arr.shape == (23 * 17, 4) and arr.dtype == np.float32
    assert arr.tobytes() == (img.reshape(-1, 4).astype(np.float32) / np.float32(255.0)).tobytes()
    try:                                                  # cross-check with an independent decoder when one is installed
        from PIL import Image
        assert np.array_equal(np.asarray(Image.open(p).convert("RGBA")), img)
    except ImportError:
        pass


def test_non_rgba8_is_refused_like_the_reference(tmp_path):
    rng = np.random.default_rng(1)
    p = str(tmp_path / "rgb.png")
    write_png(p, rng.integers(0, 256, (5, 6, 3), dtype=np.uint8), color_type=2)
    with pytest.raises(nb.NerfError):                    # the reference yields an empty Vec for non-RGBA8 bitmaps
        nb.load_image_rgba8(p)
    with pytest.raises(nb.NerfError):
        nb.load_image_rgba8(str(tmp_path / "missing.png"))


def test_get_image_paths():
    assert nb.get_image_paths("d", 0, 6, 2) == ["d/image-0.png", "d/image-2.png", "d/image-4.png"]   # image_loading.rs:37-54
    with pytest.raises(AssertionError):
        nb.get_image_paths("d", 3, 3, 1)


@pytest.mark.gpu
def test_rgba8_residency_gold_gather_is_bit_exact(tmp_path):
    rng = np.random.default_rng(2)
    w, h, v, r, s = 40, 32, 4, 64, 16
    imgs8 = rng.integers(0, 256, (v, h, w, 4), dtype=np.uint8)
    paths = nb.get_image_paths(str(tmp_path), 0, v, 1)
    for path, im in zip(paths, imgs8):
        write_png(path, im)
    loaded8 = np.stack([nb.load_image_rgba8(p) for p in paths])
    imgs_f = np.stack([nb.load_image_as_array(p) for p in paths])          # the reference's host-side f32 images
    m = nb.NeRF(nb.default_config(image_w=w, image_h=h, num_rays=r, num_samples=s, hidden=64, mlp_impl=1))
    angles = nb.get_view_angles(6)[:v]
    m.set_view_angles(angles)
    idx = np.stack([rng.integers(0, h, r), rng.integers(0, w, r)], 1).astype(np.int64)
    vi = rng.integers(0, v, v).astype(np.int64)
    u = rng.random((r, s)).astype(np.float32)
    m.set_images_rgba8(loaded8)
    gold8 = m.get_batch(idx, vi, v, u, True, 0)["gold"]
    m.set_images(imgs_f)
    goldf = m.get_batch(idx, vi, v, u, True, 0)["gold"]
    _, _, _, want = ray_c.get_multiview_batch(imgs_f, angles, idx, vi, s, u, w, h)
    assert gold8.tobytes() == goldf.tobytes() == want.tobytes()
