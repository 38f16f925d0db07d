// sampling.cu -- K-sample: pixel -> ray -> depth samples -> rotated world points.
//
// Replaces the host loops of src/ray_sampling.rs:79-178 and the gold gather of
// src/dataset.rs:111-114. One warp per ray. Bit-exactness rules:
// every f32 op is an explicit round-to-nearest intrinsic (no FMA contraction),
// sqrt/div are IEEE, and cos/sin/tan never run on the device -- the per-view
// matrices and tan(FOV/2)*HITHER come from the host.
#include "common.cuh"
#include "kernels.h"
#include "raygeom.cuh"

namespace {

using raygeom::add;
using raygeom::mul;
using raygeom::rotate_yaw_pitch;
using raygeom::sub;

// screen_to_world (ray_sampling.rs:79-93), op for op. view=(0,0,1), left=(-1,0,0), UP=(0,1,0)
// are the normalised constants the reference recomputes per call.
__device__ __forceinline__ void screen_to_world(float x, float y, float width, float height, float off, float to[3]) {
    const float two_off = mul(2.f, off);
    const float ol = sub(off, __fdiv_rn(mul(two_off, x), width));
    const float ou = sub(off, __fdiv_rn(mul(two_off, y), height));
    // (view*HITHER + left*ol) + UP*ou, component-wise with the zero products kept
    const float ax = mul(0.f, NERF_HITHER), ay = mul(0.f, NERF_HITHER), az = mul(1.f, NERF_HITHER);
    const float bx = mul(-1.f, ol), by = mul(0.f, ol), bz = mul(0.f, ol);
    const float cx = mul(0.f, ou), cy = mul(1.f, ou), cz = mul(0.f, ou);
    const float sx = add(add(ax, bx), cx), sy = add(add(ay, by), cy), sz = add(add(az, bz), cz);
    const float d = add(add(mul(sx, sx), mul(sy, sy)), mul(sz, sz));
    const float inv = __fdiv_rn(1.f, __fsqrt_rn(d));
    to[0] = mul(sx, inv);
    to[1] = mul(sy, inv);
    to[2] = mul(sz, inv);
}

constexpr int kWarpsPerBlock = 8;

// Ascending bitonic sort of 32*SPL floats held one per (register k, lane): element i = 32 k + lane. Partners closer than
// 32 apart are exchanged with a shuffle, the others are register pairs of the same lane whose direction is a compile-time
// constant after unrolling -- no shared memory and no barriers (the reference sorts (p,t) pairs by t, ray_sampling.rs:125;
// t = 2u, so sorting the uniforms is the same order).
template <int SPL>
__device__ __forceinline__ void warp_bitonic_sort(float (&v)[SPL], int lane) {
    constexpr int N = 32 * SPL;
#pragma unroll
    for (int kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    const int l = k ^ (j >> 5);
                    if (l > k) {
                        const bool up = (((32 * k) & kk) == 0);
                        const float lo = fminf(v[k], v[l]), hi = fmaxf(v[k], v[l]);
                        v[k] = up ? lo : hi;
                        v[l] = up ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    const bool up = (((32 * k + lane) & kk) == 0);
                    const bool lower = ((lane & j) == 0);
                    const float o = __shfl_xor_sync(0xffffffffu, v[k], j);
                    v[k] = (lower == up) ? fminf(v[k], o) : fmaxf(v[k], o);
                }
            }
        }
    }
}

// SPL = registers per lane for the ray's depths: the next power of two of ceil(S / 32) (S <= 256).
template <int SPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_sample(SampleArgs a) {
    __shared__ float s_p[kWarpsPerBlock][96];
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = a.num_samples;
    for (int r = blockIdx.x * kWarpsPerBlock + warp; r < a.num_rays; r += gridDim.x * kWarpsPerBlock) {
        // pixel / view picks: caller-supplied, or Philox in place (replaces Tensor::randint, dataset.rs:12,19,88)
        // Every Philox stream is indexed by the GLOBAL ray (ray_index_base + r; base = rank * R in data-parallel training), so
        // the ranks of a job draw disjoint pixels, views and jitter from one seed: N ranks x R rays = one batch of N*R rays.
        const uint64_t gr = (uint64_t)(a.ray_index_base + r);
        int y, x;
        if (a.gen_pix) {
            y = min((int)(philox_uniform(a.seed, NERF_STREAM_PIX_Y, gr) * (float)a.img_h), a.img_h - 1);
            x = min((int)(philox_uniform(a.seed, NERF_STREAM_PIX_X, gr) * (float)a.img_w), a.img_w - 1);
            if (lane == 0) { a.pix_out[2 * r] = y; a.pix_out[2 * r + 1] = x; }
        } else {
            y = a.pix_yx[2 * r];
            x = a.pix_yx[2 * r + 1];
        }
        int view = a.fixed_view;
        if (a.gen_view) {
            const int pick = r / a.rays_per_pick;
            view = min((int)(philox_uniform(a.seed, NERF_STREAM_VIEW, gr / (uint64_t)a.rays_per_pick) * (float)a.n_views), a.n_views - 1);
            if (lane == 0 && r % a.rays_per_pick == 0) a.view_out[pick] = view;
        } else if (a.view_pick) {
            view = a.view_pick[r / a.rays_per_pick];
        }
        const ViewPose vp = a.poses[view];
        float to[3];
        screen_to_world((float)x, (float)y, (float)a.img_w, (float)a.img_h, a.off, to);

        // ---- depths (ray_sampling.rs:107-125): sample i = 32 k + lane lives in tv[k]
        float tv[SPL];
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int i = 32 * k + lane;
            float t = __int_as_float(0x7f800000);   // +inf padding sorts to the end
            if (i < S) {
                if (!a.randomize) {
                    t = mul(__fdiv_rn((float)i, (float)S), 2.0f);  // :112,:114
                } else {
                    const float u = a.jitter ? a.jitter[(size_t)r * S + i]
                                             : philox_uniform(a.seed, NERF_STREAM_JITTER, gr * S + i);
                    if (a.depth_mode == 0) t = mul(u, 2.0f);                                       // :110,:114
                    else t = mul(__fdiv_rn(add((float)i, u), (float)S), 2.0f);                      // stratified
                }
            }
            tv[k] = t;
        }
        if (a.randomize && a.depth_mode == 0) warp_bitonic_sort<SPL>(tv, lane);   // sort ascending by t (:125)

        // ---- per-ray outputs
        if (lane == 0) {
            RayRec rec;
            rec.to[0] = to[0]; rec.to[1] = to[1]; rec.to[2] = to[2];
            rec.view = view;
            a.rays[r] = rec;
            float d[3];
            rotate_yaw_pitch(vp, to, d);
            a.dirs[3 * r] = d[0]; a.dirs[3 * r + 1] = d[1]; a.dirs[3 * r + 2] = d[2];
        }
        if (a.images && lane < 4)  // gold = imgs[n][y*W+x] (dataset.rs:111-114)
            a.gold[4 * (size_t)r + lane] = a.images[(((size_t)view * a.img_h + y) * a.img_w + x) * 4 + lane];
        else if (a.images_u8 && lane < 4)  // ... with the loader's `as f32 / 255.` (image_loading.rs:13-18) fused in
            a.gold[4 * (size_t)r + lane] = __fdiv_rn((float)a.images_u8[(((size_t)view * a.img_h + y) * a.img_w + x) * 4 + lane], 255.f);

        // ---- points: p = FROM + to*t (:115), then yaw, pitch per point (:128-132)
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int base = 32 * k;
            if (base < S) {   // warp-uniform
                const int i = base + lane;
                if (i < S) a.t[(size_t)r * S + i] = tv[k];
                if (a.points) {
                    if (i < S) {
                        float q[3];
                        raygeom::sample_point(vp, to, tv[k], q);
                        s_p[warp][3 * lane] = q[0]; s_p[warp][3 * lane + 1] = q[1]; s_p[warp][3 * lane + 2] = q[2];
                    }
                    __syncwarp();
                    const int nvalid = min(32, S - base) * 3;
                    float *dst = a.points + ((size_t)r * S + base) * 3;
                    for (int q = lane; q < nvalid; q += 32) dst[q] = s_p[warp][q];
                    __syncwarp();
                }
            }
        }
    }
}

// Accurate sinusoidal encoding to f32 (parity / SIMT path). out[n][3+6L]:
// [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] per 3-vector, no pi.
__global__ void k_encode(const float *__restrict__ x, float *__restrict__ out, int64_t n, int freqs, int repeat) {
    const int C = 3 + 6 * freqs;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *src = x + 3 * (i / repeat);
    float v[3] = {src[0], src[1], src[2]};
    float *o = out + i * C;
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
    for (int k = 0; k < freqs; ++k) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s, co;
            sincosf(__fmul_rn(v[c], f), &s, &co);
            o[3 + 6 * k + c] = s;
            o[3 + 6 * k + 3 + c] = co;
        }
    }
}

// (c*255) as u8 saturating-truncating cast, packed 0x00RRGGBB (display.rs:37-52)
__global__ void k_pack_0rgb(const float *__restrict__ rgba, uint32_t *__restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = __fmul_rn(rgba[4 * i + k], 255.f);
        v = (v != v) ? 0.f : fminf(fmaxf(v, 0.f), 255.f);  // Rust `as u8`: NaN -> 0, saturate
        c[k] = (uint32_t)v;
    }
    out[i] = (c[0] << 16) | (c[1] << 8) | c[2];
}

// pixel (y, x) of the flat row-major pixel range [first, first + n)
__global__ void k_flat_pixel_indices(int32_t *pix_yx, int64_t first, int n, int img_w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = first + i;
    pix_yx[2 * i] = (int)(p / img_w);
    pix_yx[2 * i + 1] = (int)(p % img_w);
}

__global__ void k_full_frame_indices(int32_t *pix_yx, int y0, int y1, int img_w) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = (int64_t)(y1 - y0) * img_w;
    if (i >= n) return;
    pix_yx[2 * i] = y0 + (int)(i / img_w);
    pix_yx[2 * i + 1] = (int)(i % img_w);
}

}  // namespace

template <int SPL>
static void launch_sample_spl(const SampleArgs &a, int num_sms, cudaStream_t st) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sample<SPL>, kWarpsPerBlock * 32, 0);
    int blocks = (a.num_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int cap = num_sms * (occ < 1 ? 1 : occ);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    launch_pdl(k_sample<SPL>, dim3(blocks), dim3(kWarpsPerBlock * 32), 0, st, a);
}
void launch_sample(const SampleArgs &a, int num_sms, cudaStream_t st) {
    const int spl = (a.num_samples + 31) / 32;
    if (spl <= 1) launch_sample_spl<1>(a, num_sms, st);
    else if (spl <= 2) launch_sample_spl<2>(a, num_sms, st);
    else if (spl <= 4) launch_sample_spl<4>(a, num_sms, st);
    else launch_sample_spl<8>(a, num_sms, st);
}

void launch_encode(const float *x, float *out, int64_t n, int freqs, int repeat, cudaStream_t st) {
    if (n <= 0) return;
    k_encode<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, out, n, freqs, repeat);
}

void launch_pack_0rgb(const float *rgba, uint32_t *out, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    k_pack_0rgb<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rgba, out, n);
}

void launch_flat_pixel_indices(int32_t *pix_yx, int64_t first, int n, int img_w, cudaStream_t st) {
    if (n <= 0) return;
    k_flat_pixel_indices<<<(n + 255) / 256, 256, 0, st>>>(pix_yx, first, n, img_w);
}

void launch_full_frame_indices(int32_t *pix_yx, int y0, int y1, int img_w, cudaStream_t st) {
    int64_t n = (int64_t)(y1 - y0) * img_w;
    if (n <= 0) return;
    k_full_frame_indices<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pix_yx, y0, y1, img_w);
}
