// metrics.cu -- K-metrics: the reference's per-batch TensorBoard projections (src/logging.rs) computed on the device from the
// resident batch, so a logging step copies a few KB of histograms / 100x100 maps instead of the batch itself.
//   log_screen_coords        logging.rs:13-25    pixel-index histograms
//   log_query_distances      logging.rs:27-39    histogram of floor(500 t), 2000 buckets
//   log_query_points_as_maps logging.rs:41-107   yx / zx / yz occupancy maps (white where a sample lands, index clamped to 9999)
//   log_densities            logging.rs:109-134  per-axis density sums, 2000 buckets of floor(500 (w + 1)), f64
//   log_density_maps         logging.rs:136-195  yx / zx / yz maps of the LAST sample's clamped density (sequential overwrite)
//   draw_predictions         display.rs:96-110   batch pixels scattered into a WxH 0x00RRGGBB back buffer (last ray wins)
// "Last writer wins" is made order-exact with a 64-bit atomicMax on (sequence number << 32 | colour). Sample positions come
// from the points buffer when it exists, else they are rebuilt from the ray records exactly like the MLP prologue does.
#include "common.cuh"
#include "kernels.h"
#include "raygeom.cuh"

namespace {

// Rust `f32 as usize`: saturating, NaN -> 0
__device__ __forceinline__ long long as_usize(float v) {
    if (!(v == v) || v <= 0.f) return 0;
    if (v >= 9.0e18f) return 9000000000000000000ll;
    return (long long)v;
}
// (c * 255.) as u8 per channel, packed 0x00RRGGBB (display.rs:37-52)
__device__ __forceinline__ uint32_t pack_0rgb(float r, float g, float b) {
    auto u8 = [](float c) -> uint32_t {
        float v = __fmul_rn(c, 255.f);
        v = (v != v) ? 0.f : fminf(fmaxf(v, 0.f), 255.f);
        return (uint32_t)v;
    };
    return (u8(r) << 16) | (u8(g) << 8) | u8(b);
}

__global__ void __launch_bounds__(256) k_metrics(MetricsArgs a) {
    __shared__ unsigned int s_t[2000];
    for (int i = threadIdx.x; i < 2000; i += blockDim.x) s_t[i] = 0u;
    __syncthreads();
    const int64_t n = (int64_t)a.num_rays * a.num_samples;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ray = i / a.num_samples;
        const float t = a.t[i];
        {   // log_query_distances: bucket_counts_t[floor(500 t)] += 1 (:33-35); the reference panics past 1999
            long long b = as_usize(floorf(__fmul_rn(500.f, t)));
            atomicAdd(&s_t[b > 1999 ? 1999 : b], 1u);
        }
        float w[3];
        if (a.points) {
            w[0] = a.points[3 * i]; w[1] = a.points[3 * i + 1]; w[2] = a.points[3 * i + 2];
        } else {
            const RayRec rec = a.rays[ray];
            raygeom::sample_point(a.poses[rec.view], rec.to, t, w);
        }
        // map cells (:60-62, :157-159): y = floor(50 (wy + 1)), x = floor(50 (wx + 1)), z = floor(25 (wz + 1))
        const long long my = as_usize(floorf(__fmul_rn(50.f, __fadd_rn(w[1], 1.f))));
        const long long mx = as_usize(floorf(__fmul_rn(50.f, __fadd_rn(w[0], 1.f))));
        const long long mz = as_usize(floorf(__fmul_rn(25.f, __fadd_rn(w[2], 1.f))));
        const long long cyx = my * 100 + mx, czx = mz * 100 + mx, cyz = my * 100 + mz;
        if (a.world_maps) {   // .min(10000 - 1) (:64-69)
            a.world_maps[cyx < 9999 ? cyx : 9999] = 0x00FFFFFFu;
            a.world_maps[10000 + (czx < 9999 ? czx : 9999)] = 0x00FFFFFFu;
            a.world_maps[20000 + (cyz < 9999 ? cyz : 9999)] = 0x00FFFFFFu;
        }
        if (a.sigma) {
            const float d = a.sigma[i];
            if (a.density_hist) {   // (:118-127) buckets floor(500 (w + 1)); out-of-range cells (a panic in the reference) are dropped
                const long long by = as_usize(floorf(__fmul_rn(500.f, __fadd_rn(w[1], 1.f))));
                const long long bx = as_usize(floorf(__fmul_rn(500.f, __fadd_rn(w[0], 1.f))));
                const long long bz = as_usize(floorf(__fmul_rn(500.f, __fadd_rn(w[2], 1.f))));
                if (bx < 2000) atomicAdd(a.density_hist + bx, (double)d);
                if (by < 2000) atomicAdd(a.density_hist + 2000 + by, (double)d);
                if (bz < 2000) atomicAdd(a.density_hist + 4000 + bz, (double)d);
            }
            if (a.density_maps) {   // (:161-170) sequential overwrite == the write with the largest sample index
                const float dc = fmaxf(d, 0.f);
                const unsigned long long key = ((unsigned long long)(i + 1) << 32) | pack_0rgb(dc, dc, dc);
                if (cyx < 10000) atomicMax(a.density_maps + cyx, key);
                if (czx < 10000) atomicMax(a.density_maps + 10000 + czx, key);
                if (cyz < 10000) atomicMax(a.density_maps + 20000 + cyz, key);
            }
        }
        if (i % a.num_samples == 0) {   // per-ray work, done by the ray's first sample
            const int p0 = a.pix_yx[2 * ray], p1 = a.pix_yx[2 * ray + 1];
            // log_screen_coords binds `[x, y]` to the stored [y, x] pair (:17-20): screen_x counts element 0, screen_y element 1
            if (a.screen_hist) {
                if (p0 < a.img_w) atomicAdd(a.screen_hist + p0, 1u);
                if (p1 < a.img_h) atomicAdd(a.screen_hist + a.img_w + p1, 1u);
            }
            if (a.prediction && a.pixels) {   // draw_predictions: backbuffer[y * WIDTH + x] = 0RGB(pred) (display.rs:102-108)
                const float4 p = reinterpret_cast<const float4 *>(a.pixels)[ray];
                const unsigned long long key = ((unsigned long long)(ray + 1) << 32) | pack_0rgb(p.x, p.y, p.z);
                atomicMax(a.prediction + (size_t)p0 * a.img_w + p1, key);
            }
        }
    }
    __syncthreads();
    if (a.t_hist)
        for (int i = threadIdx.x; i < 2000; i += blockDim.x)
            if (s_t[i]) atomicAdd(a.t_hist + i, s_t[i]);
}

// keys (sequence << 32 | colour) -> colours
__global__ void k_metrics_resolve(const unsigned long long *__restrict__ keys, uint32_t *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)keys[i];
}

}  // namespace

void launch_metrics(const MetricsArgs &a, int num_sms, cudaStream_t st) {
    const int64_t n = (int64_t)a.num_rays * a.num_samples;
    int64_t blocks = (n + 255) / 256;
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    if (blocks < 1) blocks = 1;
    k_metrics<<<(unsigned)blocks, 256, 0, st>>>(a);
}
void launch_metrics_resolve(const unsigned long long *keys, uint32_t *out, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    k_metrics_resolve<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, out, n);
}
