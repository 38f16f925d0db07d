// host_io.cpp -- dataset ingest: the reference's image loader (src/image_loading.rs:6-24) restated for the drop-in.
// load_image_as_array decodes a PNG to RGBA8 and divides by 255 on the host; here the RGBA8 bytes are decoded on the
// host (zlib inflate + PNG scan-line unfiltering) and stay RGBA8 on the device -- the /255 (IEEE f32 division, exactly
// `rgba.r as f32 / 255.`) is fused into the sampler's gold gather, so a view costs 4 bytes per pixel of HBM instead of 16.
// Like the reference (which yields an empty Vec for anything but ImageData::RGBA8) only 8-bit RGBA PNGs are accepted.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/nerf_b200.h"

namespace {

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

static int load_png_rgba8(const char *path, uint8_t *out, int64_t capacity_bytes, int32_t *width, int32_t *height);

// Nothing may be thrown across the C ABI (a crafted or truncated file must not terminate the Rust / ctypes caller).
extern "C" int nerf_load_png_rgba8(const char *path, uint8_t *out, int64_t capacity_bytes, int32_t *width, int32_t *height) {
    try {
        return load_png_rgba8(path, out, capacity_bytes, width, height);
    } catch (...) {   // std::bad_alloc, std::length_error
        return NERF_ERR_INVALID_ARG;
    }
}

static int load_png_rgba8(const char *path, uint8_t *out, int64_t capacity_bytes, int32_t *width, int32_t *height) {
    if (!path || !width || !height) return NERF_ERR_INVALID_ARG;
    FILE *f = fopen(path, "rb");
    if (!f) return NERF_ERR_INVALID_ARG;
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || memcmp(file.data(), sig, 8) != 0) return NERF_ERR_INVALID_ARG;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t *type = &file[pos + 4], *data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) return NERF_ERR_INVALID_ARG;
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    // IHDR is untrusted: bound the dimensions before any size arithmetic or allocation (65535^2 * 4 fits size_t comfortably)
    if (w == 0 || h == 0 || w > 65535u || h > 65535u) return NERF_ERR_INVALID_ARG;
    *width = (int32_t)w;
    *height = (int32_t)h;
    if (depth != 8 || ctype != 6 || interlace != 0) return NERF_ERR_UNSUPPORTED;   // not ImageData::RGBA8
    if (!out) return NERF_OK;                                                        // size query
    const size_t stride = (size_t)w * 4, need = stride * h;
    if ((size_t)capacity_bytes < need) return NERF_ERR_INVALID_ARG;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return NERF_ERR_INVALID_ARG;
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t *src = &raw[(stride + 1) * y];
        const int filter = src[0];
        ++src;
        uint8_t *dst = out + stride * y;
        const uint8_t *up = y ? dst - stride : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= 4 ? dst[x - 4] : 0, b = up ? up[x] : 0, c = (up && x >= 4) ? up[x - 4] : 0;
            int v = src[x];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: return NERF_ERR_INVALID_ARG;
            }
            dst[x] = (uint8_t)v;
        }
    }
    return NERF_OK;
}
