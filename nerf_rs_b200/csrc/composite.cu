// composite.cu -- K-composite: front-to-back alpha compositing, one warp per ray.
//
// Forward restates src/model.rs:184-187 (deltas), :221-232 (transmittance) and
// :234-249 (weights, weighted colour sum) as a segmented warp scan instead of the
// reference's S separate slice/mul/sum/neg/exp chains:
//   x_i = sigma_i * delta_i,  T_i = exp(-sum_{j<i} x_j),  w_i = T_i (1 - exp(-x_i)),
//   out_c = sum_i w_i col_{i,c}.
// Backward is the analytic gradient fused with the MSE gradient
// g = 2 (out - gold) / (4R) of src/model.rs:296-299.
// HBM-bound: 24 B/sample read forward; +20 B/sample written backward.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxPerLane = 8;  // S <= 256

// inclusive warp scan; total = the warp's sum
__device__ __forceinline__ float warp_incl_scan_add(float v, int lane, float &total) {
    float inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// One warp per ray; samples are interleaved over the lanes (sample i = 32 k + lane, k < SPL), so every global load and
// store of a warp is one contiguous run (128 B of sigma / t, 512 B of rgba per instruction) and ALL of a ray's loads are
// issued before the first dependent instruction. SPL = ceil(S / 32) is a template parameter: registers scale with the
// ray length instead of the S <= 256 worst case, which is what buys the occupancy a streaming kernel needs.
template <bool kBackward, int SPL>
__global__ void __launch_bounds__(kWarps * 32)
k_composite(CompositeArgs a) {
    pdl_trigger();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = a.num_samples;
    float loss_sum = 0.f;   // this warp's squared errors, rays in loop order (identical in every lane)
    for (int r = blockIdx.x * kWarps + warp; r < a.num_rays; r += gridDim.x * kWarps) {
        const float *sig = a.sigma + (size_t)r * S;
        const float *tt = a.t_or_delta + (size_t)r * S;
        const float4 *colp = reinterpret_cast<const float4 *>(a.colors) + (size_t)r * S;
        float sg[SPL], dl[SPL];
        float4 col[SPL];
        // ---- loads (independent of each other)
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int i = 32 * k + lane;
            sg[k] = 0.f; dl[k] = 0.f;
            col[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < S) {
                sg[k] = sig[i];
                if (a.input_is_delta) dl[k] = tt[i];
                else dl[k] = ((i + 1 < S) ? tt[i + 1] : NERF_T_FAR) - tt[i];  // model.rs:184-187
                if (a.colors) col[k] = colp[i];
            }
        }
        float g[4] = {0.f, 0.f, 0.f, 0.f}, gold4[4] = {0.f, 0.f, 0.f, 0.f};
        if (kBackward) {
            if (a.d_out) {
#pragma unroll
                for (int c = 0; c < 4; ++c) g[c] = a.d_out[4 * (size_t)r + c];
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) gold4[c] = a.gold[4 * (size_t)r + c];
            }
        }
        // ---- x_i = sigma_i delta_i, exclusive prefix over the ray, T_i = exp(-prefix), w_i = T_i (1 - exp(-x_i))
        float x[SPL], w[SPL], Tn[SPL];
        float carry = 0.f;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int i = 32 * k + lane;
            float s = sg[k];
            if (a.sigma_relu) s = fmaxf(s, 0.f);
            if (!a.colors) col[k] = (i < S) ? make_float4(s, s, s, 1.f) : col[k];  // as shipped: (sigma,sigma,sigma,1) (model.rs:192-204)
            x[k] = s * dl[k];
            float tot;
            const float incl = warp_incl_scan_add(x[k], lane, tot);
            const float p = carry + (incl - x[k]);   // sum_{j<i} x_j
            carry += tot;
            // accurate expf -- the reference uses ATen exp; keeps parity within 1e-6
            const float T = expf(-p);
            const float av = expf(-x[k]);
            w[k] = (i < S) ? T * (1.f - av) : 0.f;   // model.rs:243
            Tn[k] = T * av;                          // T_{i+1}
            acc[0] += w[k] * col[k].x; acc[1] += w[k] * col[k].y; acc[2] += w[k] * col[k].z; acc[3] += w[k] * col[k].w;
        }
        float out[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) out[c] = warp_sum(acc[c]);  // model.rs:245-246

        if (!kBackward) {
            if (lane < 4) a.out[4 * (size_t)r + lane] = out[lane];
            continue;
        }

        // ---- backward: g = dL/dout, either supplied or the fused MSE gradient
        if (!a.d_out) {
            float l = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float diff = out[c] - gold4[c];
                g[c] = 2.f * diff * a.inv_count;   // d mean((x-y)^2) / dx, count = 4R (model.rs:296-299)
                l += diff * diff;
            }
            loss_sum += l;
            if (lane < 4 && a.out) a.out[4 * (size_t)r + lane] = out[lane];
        }
        // s_i = g . col_i ;  dL/dsigma_i = delta_i (T_{i+1} s_i - sum_{k>i} w_k s_k)
        float ws[SPL], incl_ws[SPL];
        float carry_ws = 0.f;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            ws[k] = g[0] * col[k].x + g[1] * col[k].y + g[2] * col[k].z + g[3] * col[k].w;
            float tot;
            incl_ws[k] = carry_ws + warp_incl_scan_add(w[k] * ws[k], lane, tot);
            carry_ws += tot;
        }
        float4 *dcol = reinterpret_cast<float4 *>(a.d_colors) + (size_t)r * S;
        float *dsg = a.d_sigma + (size_t)r * S;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int i = 32 * k + lane;
            if (i < S) {
                const float suffix = carry_ws - incl_ws[k];                 // sum_{k>i} w_k s_k
                float dsig = dl[k] * (Tn[k] * ws[k] - suffix);              // dL/dsigma_i
                if (a.colors) dcol[i] = make_float4(w[k] * g[0], w[k] * g[1], w[k] * g[2], w[k] * g[3]);
                else dsig += w[k] * (g[0] + g[1] + g[2]);                   // direct colour = sigma terms
                if (a.sigma_relu && sg[k] <= 0.f) dsig = 0.f;
                dsg[i] = dsig;
            }
        }
    }
    // ---- fused loss (model.rs:296-299), deterministic: warps -> block partial in a fixed order; the last block of the
    //      launch that covers the step's last rays sums the partials of all the step's launches in a fixed order
    if (kBackward && a.loss_partials) {
        __shared__ float sh[kWarps * 32];
        __shared__ bool is_last;
        if (lane == 0) sh[warp] = loss_sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < kWarps; ++k) s += sh[k];
            a.loss_partials[blockIdx.x] = s;
        }
        if (!a.loss_out) return;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = atomicAdd(a.done_counter, 1u) == gridDim.x - 1;
        __syncthreads();
        if (is_last) {
            __threadfence();
            float s = 0.f;
            for (int i = threadIdx.x; i < a.loss_partials_total; i += blockDim.x) s += __ldcg(a.loss_partials_first + i);
            __syncthreads();
            sh[threadIdx.x] = s;
            __syncthreads();
            for (int d = (kWarps * 32) >> 1; d > 0; d >>= 1) {
                if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                *a.loss_out = sh[0] * a.loss_scale;
                *a.done_counter = 0u;
            }
        }
    }
}

template <bool kBackward, int SPL>
int launch_composite_spl(CompositeArgs a, int num_sms, cudaStream_t st) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_composite<kBackward, SPL>, kWarps * 32, 0);
    int blocks = (a.num_rays + kWarps - 1) / kWarps;
    const int cap = num_sms * (occ < 1 ? 1 : occ);   // one resident wave; warps stride over the rays
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    a.loss_partials_total = a.loss_partials_prior + blocks;
    launch_pdl(k_composite<kBackward, SPL>, dim3(blocks), dim3(kWarps * 32), 0, st, a);
    return blocks;
}
template <bool kBackward>
int launch_composite_any(const CompositeArgs &a, int num_sms, cudaStream_t st) {
    switch ((a.num_samples + 31) / 32) {
        case 1: return launch_composite_spl<kBackward, 1>(a, num_sms, st);
        case 2: return launch_composite_spl<kBackward, 2>(a, num_sms, st);
        case 3: return launch_composite_spl<kBackward, 3>(a, num_sms, st);
        case 4: return launch_composite_spl<kBackward, 4>(a, num_sms, st);
        case 5: return launch_composite_spl<kBackward, 5>(a, num_sms, st);
        case 6: return launch_composite_spl<kBackward, 6>(a, num_sms, st);
        case 7: return launch_composite_spl<kBackward, 7>(a, num_sms, st);
        default: return launch_composite_spl<kBackward, kMaxPerLane>(a, num_sms, st);
    }
}

// synthetic inputs for nerf_debug_bench_stage: hashed uniforms in [lo, hi)
__global__ void k_fill_uniform(float *__restrict__ p, int64_t n, uint32_t seed, float lo, float hi) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ (uint32_t)(i >> 32) ^ (seed * 0x9E3779B9u);
        h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
        p[i] = lo + (hi - lo) * ((float)(h >> 8) * (1.f / 16777216.f));
    }
}

}  // namespace

void launch_fill_uniform(float *p, int64_t n, uint32_t seed, float lo, float hi, cudaStream_t st) {
    if (n <= 0) return;
    k_fill_uniform<<<148 * 8, 256, 0, st>>>(p, n, seed, lo, hi);
}
void launch_composite_fwd(const CompositeArgs &a, int num_sms, cudaStream_t st) { launch_composite_any<false>(a, num_sms, st); }
int launch_composite_bwd(const CompositeArgs &a, int num_sms, cudaStream_t st) { return launch_composite_any<true>(a, num_sms, st); }
