// adam.cu -- K-adam: one fused elementwise pass over the flat parameter blob.
//
// Restates nn::Adam::default() as used by Trainer::new/step (src/model.rs:306-309, :322):
// betas (.9,.999), eps 1e-8, no weight decay, no amsgrad, in libtorch's op order
//   m = m*b1 + g*(1-b1);  v = v*b2 + (1-b2)*g*g;
//   denom = sqrt(v)/sqrt(1-b2^t) + eps;  p -= (lr/(1-b1^t)) * m/denom.
// HBM-bound: p,m,v read+write, g read (+ zeroed for the next step) = 28(+4) B/param.
// grad_scale folds the 1/nranks of the data-parallel all-reduce.
#include <cstdio>

#include "common.cuh"
#include "kernels.h"

namespace {

__global__ void __launch_bounds__(256) k_adam(AdamArgs a) {
    pdl_trigger();
    pdl_wait();
    const int64_t n4 = a.n >> 2;
    const float omb1 = 1.f - a.beta1, omb2 = 1.f - a.beta2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = reinterpret_cast<float4 *>(a.p)[i];
        float4 m = reinterpret_cast<float4 *>(a.m)[i];
        float4 v = reinterpret_cast<float4 *>(a.v)[i];
        float4 g = reinterpret_cast<float4 *>(a.g)[i];
        float *pp = &p.x, *mm = &m.x, *vv = &v.x, *gg = &g.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = gg[k] * a.grad_scale;
            mm[k] = mm[k] * a.beta1 + gr * omb1;
            vv[k] = vv[k] * a.beta2 + omb2 * gr * gr;
            const float denom = sqrtf(vv[k]) * a.inv_sqrt_bc2 + a.eps;
            pp[k] = pp[k] - a.lr_over_bc1 * (mm[k] / denom);
        }
        reinterpret_cast<float4 *>(a.p)[i] = p;
        reinterpret_cast<float4 *>(a.m)[i] = m;
        reinterpret_cast<float4 *>(a.v)[i] = v;
        if (a.zero_grad) reinterpret_cast<float4 *>(a.g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail
    if (blockIdx.x == 0) {
        for (int64_t i = (n4 << 2) + threadIdx.x; i < a.n; i += blockDim.x) {
            const float gr = a.g[i] * a.grad_scale;
            const float m = a.m[i] * a.beta1 + gr * omb1;
            const float v = a.v[i] * a.beta2 + omb2 * gr * gr;
            const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
            a.p[i] = a.p[i] - a.lr_over_bc1 * (m / denom);
            a.m[i] = m;
            a.v[i] = v;
            if (a.zero_grad) a.g[i] = 0.f;
        }
    }
}

// ---- fused all-reduce + Adam over peer memory ------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float *p) {   // peer memory: do not keep in L1
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// a dead or stalled peer must not tear the context down (__trap) nor hang it forever: after ~2^31 polls (minutes) the kernel
// gives up on the exchange, records the error for the host (nerf_last_error via the context's error word) and returns
__device__ unsigned int g_p2p_timeout = 0;
__device__ __forceinline__ bool wait_flags(const unsigned int *flags, int nranks, uint32_t step) {
    for (int r = 0; r < nranks; ++r) {
        unsigned long long spins = 0;
        while ((int)(ld_acquire_sys(flags + r) - step) < 0) {
            if (++spins > (1ull << 31)) { atomicExch(&g_p2p_timeout, step); return false; }
            __nanosleep(64);
        }
    }
    return true;
}
__device__ __forceinline__ void st_f4(float *p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256) k_adam_p2p(const __grid_constant__ AdamP2PArgs a) {
    pdl_trigger();
    pdl_wait();
    __shared__ int s_ok;
    // hand-shake 1: publish "my gradient for `step` is complete" (the weight-gradient kernel precedes this one in the stream)
    // in every rank's flag array, slot = my rank; then wait until every rank has published (local polls)
    if (blockIdx.x == 0 && (int)threadIdx.x < a.nranks) {
        __threadfence_system();
        st_release_sys(a.peer_flags[threadIdx.x] + a.rank, a.step);
    }
    if (threadIdx.x == 0) s_ok = wait_flags(a.my_flags, a.nranks, a.step) ? 1 : 0;
    __syncthreads();
    if (!s_ok) return;
    const AdamArgs &ad = a.adam;
    const int64_t n4 = ad.n >> 2;
    const int64_t shard4 = (n4 + a.nranks - 1) / a.nranks;
    // reduce-scatter: my shard of every rank's local gradient, all peer loads in flight together, fixed rank-order sum
    {
        const int64_t lo = shard4 * a.rank, hi = lo + shard4 < n4 ? lo + shard4 : n4;
        float *red = a.peer_grads[a.rank] + a.n_pad;
        for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
            float4 t[NERF_MAX_RANKS];
#pragma unroll
            for (int r = 0; r < NERF_MAX_RANKS; ++r)
                t[r] = r < a.nranks ? ld_peer_f4(a.peer_grads[r] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 g = t[0];
#pragma unroll
            for (int r = 1; r < NERF_MAX_RANKS; ++r) { g.x += t[r].x; g.y += t[r].y; g.z += t[r].z; g.w += t[r].w; }
            st_f4(red + 4 * i, g);
        }
    }
    // hand-shake 2: the LAST block of this rank to finish its part of the shard publishes it to every rank
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(a.block_counter, 1u);
        if (done == gridDim.x - 1) {
            *a.block_counter = 0u;
            __threadfence_system();
            for (int r = 0; r < a.nranks; ++r) st_release_sys(a.peer_flags[r] + NERF_MAX_RANKS + a.rank, a.step);
        }
        s_ok = wait_flags(a.my_flags + NERF_MAX_RANKS, a.nranks, a.step) ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok) return;
    // all-gather + Adam on my replica
    const float omb1 = 1.f - ad.beta1, omb2 = 1.f - ad.beta2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int owner = (int)(i / shard4);
        float4 g = ld_peer_f4(a.peer_grads[owner] + a.n_pad + 4 * i);
        float4 p = reinterpret_cast<float4 *>(ad.p)[i];
        float4 m = reinterpret_cast<float4 *>(ad.m)[i];
        float4 v = reinterpret_cast<float4 *>(ad.v)[i];
        float *pp = &p.x, *mm = &m.x, *vv = &v.x, *gg = &g.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = gg[k] * ad.grad_scale;
            mm[k] = mm[k] * ad.beta1 + gr * omb1;
            vv[k] = vv[k] * ad.beta2 + omb2 * gr * gr;
            const float denom = sqrtf(vv[k]) * ad.inv_sqrt_bc2 + ad.eps;
            pp[k] = pp[k] - ad.lr_over_bc1 * (mm[k] / denom);
        }
        reinterpret_cast<float4 *>(ad.p)[i] = p;
        reinterpret_cast<float4 *>(ad.m)[i] = m;
        reinterpret_cast<float4 *>(ad.v)[i] = v;
        reinterpret_cast<float4 *>(ad.g)[i] = g;      // the summed gradient, like ncclAllReduce leaves it
    }
    if (blockIdx.x == 0) {   // tail (n % 4): every rank sums it directly, in rank order
        for (int64_t i = (n4 << 2) + threadIdx.x; i < ad.n; i += blockDim.x) {
            float g = 0.f;
            for (int r = 0; r < a.nranks; ++r) g += *reinterpret_cast<const volatile float *>(a.peer_grads[r] + i);
            const float gr = g * ad.grad_scale;
            const float m = ad.m[i] * ad.beta1 + gr * omb1;
            const float v = ad.v[i] * ad.beta2 + omb2 * gr * gr;
            const float denom = sqrtf(v) * ad.inv_sqrt_bc2 + ad.eps;
            ad.p[i] = ad.p[i] - ad.lr_over_bc1 * (m / denom);
            ad.m[i] = m;
            ad.v[i] = v;
            ad.g[i] = g;
        }
    }
}

// The one-shot variant (round 1): one hand-shake, then every rank pulls ALL ranks' full gradients ((N - 1) gradient sizes over
// NVLink per rank). One hand-shake less than the two-phase kernel above, N/2 times its traffic; see launch_adam_p2p for which
// one runs.
__global__ void __launch_bounds__(256) k_adam_p2p_oneshot(const __grid_constant__ AdamP2PArgs a) {
    pdl_trigger();
    pdl_wait();
    // 1. publish "my gradient for `step` is complete" (the weight-gradient kernel precedes this one in the stream) in
    //    every rank's flag array, slot = my rank
    if (blockIdx.x == 0 && (int)threadIdx.x < a.nranks) {
        __threadfence_system();
        st_release_sys(a.peer_flags[threadIdx.x] + a.rank, a.step);
    }
    // 2. wait until every rank has published this step (local polls; a dead peer becomes a trap, not a hang)
    if (threadIdx.x == 0) {
        for (int r = 0; r < a.nranks; ++r) {
            unsigned int spins = 0;
            while ((int)(ld_acquire_sys(a.my_flags + r) - a.step) < 0) {
                if (++spins > (1u << 30)) { atomicExch(&g_p2p_timeout, a.step); break; }
                __nanosleep(64);
            }
        }
    }
    __syncthreads();
    // 3. sum the peers' gradients in rank order and apply Adam to my replica
    const AdamArgs &ad = a.adam;
    const int64_t n4 = ad.n >> 2;
    const float omb1 = 1.f - ad.beta1, omb2 = 1.f - ad.beta2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        // all peer loads in flight together (NVLink round trips overlap), then a fixed rank-order sum
        float4 t[NERF_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < NERF_MAX_RANKS; ++r)
            t[r] = r < a.nranks ? ld_peer_f4(a.peer_grads[r] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 g = t[0];
#pragma unroll
        for (int r = 1; r < NERF_MAX_RANKS; ++r) { g.x += t[r].x; g.y += t[r].y; g.z += t[r].z; g.w += t[r].w; }
        float4 p = reinterpret_cast<float4 *>(ad.p)[i];
        float4 m = reinterpret_cast<float4 *>(ad.m)[i];
        float4 v = reinterpret_cast<float4 *>(ad.v)[i];
        float *pp = &p.x, *mm = &m.x, *vv = &v.x, *gg = &g.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = gg[k] * ad.grad_scale;
            mm[k] = mm[k] * ad.beta1 + gr * omb1;
            vv[k] = vv[k] * ad.beta2 + omb2 * gr * gr;
            const float denom = sqrtf(vv[k]) * ad.inv_sqrt_bc2 + ad.eps;
            pp[k] = pp[k] - ad.lr_over_bc1 * (mm[k] / denom);
        }
        reinterpret_cast<float4 *>(ad.p)[i] = p;
        reinterpret_cast<float4 *>(ad.m)[i] = m;
        reinterpret_cast<float4 *>(ad.v)[i] = v;
        reinterpret_cast<float4 *>(ad.g)[i] = g;      // the summed gradient, like ncclAllReduce leaves it
    }
    if (blockIdx.x == 0) {   // tail (n % 4)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < ad.n; i += blockDim.x) {
            float g = 0.f;
            for (int r = 0; r < a.nranks; ++r) g += *reinterpret_cast<const volatile float *>(a.peer_grads[r] + i);
            const float gr = g * ad.grad_scale;
            const float m = ad.m[i] * ad.beta1 + gr * omb1;
            const float v = ad.v[i] * ad.beta2 + omb2 * gr * gr;
            const float denom = sqrtf(v) * ad.inv_sqrt_bc2 + ad.eps;
            ad.p[i] = ad.p[i] - ad.lr_over_bc1 * (m / denom);
            ad.m[i] = m;
            ad.v[i] = v;
            ad.g[i] = g;
        }
    }
}

// U(-1/sqrt(in), 1/sqrt(in)) for weights and biases -- the bound nn.Linear / tch nn::linear
// defaults use (kaiming_uniform(a=sqrt(5)) reduces to it). Exact values are irrelevant once
// set_weights injects a blob; this only makes a fresh context trainable.
__global__ void k_init_uniform(float *p, LayerGeom lg, uint64_t seed, int layer) {
    const int64_t nw = (int64_t)lg.in_dim * lg.out_dim;
    const int64_t n = nw + lg.out_dim;
    const float bound = rsqrtf((float)lg.in_dim);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float u = philox_uniform(seed, NERF_STREAM_INIT + (uint32_t)layer, (uint64_t)i);
        const float val = (2.f * u - 1.f) * bound;
        if (i < nw) p[lg.w_off + i] = val;
        else p[lg.b_off + (i - nw)] = val;
    }
}

}  // namespace

void launch_adam(const AdamArgs &a, int num_sms, cudaStream_t st) {
    int64_t n4 = a.n >> 2;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    if (blocks < 1) blocks = 1;
    launch_pdl(k_adam, dim3(blocks), dim3(256), 0, st, a);
}

void launch_adam_p2p(const AdamP2PArgs &a, int num_sms, cudaStream_t st) {
    int64_t n4 = a.adam.n >> 2;
    int blocks = (int)((n4 + 255) / 256);
    // every block waits for its rank's last block (hand-shake 2): the grid must be co-resident (256 threads, no shared
    // memory: 8 blocks per SM fit; 4 per SM leaves room beside the tail of the weight-gradient kernel)
    if (blocks > num_sms * 4) blocks = num_sms * 4;
    if (blocks < 1) blocks = 1;
    // Which kernel: measured on 8 B200s at 530 181 parameters (profiles/r02_scaling_notes.md) the step costs the same with
    // either (the exchange waits for the slowest GPU, not for bytes), so the one-shot kernel -- one hand-shake less -- is used
    // while (N - 1) gradient sizes per rank are a few microseconds of NVLink time, the two-phase kernel beyond that.
    // NERF_B200_P2P=2 / =3 force the one-shot / the two-phase kernel (A/B runs).
    const char *env = getenv("NERF_B200_P2P");
    const char sel = env ? env[0] : 0;
    const bool big = (double)a.adam.n * 4.0 * (a.nranks - 1) > 32e6;
    const bool oneshot = sel == '2' ? true : (sel == '3' ? false : !big);
    if (oneshot) launch_pdl(k_adam_p2p_oneshot, dim3(blocks), dim3(256), 0, st, a);
    else launch_pdl(k_adam_p2p, dim3(blocks), dim3(256), 0, st, a);
}
// step of the last peer-flag time-out seen by this process (0 = none): the exchange was abandoned, replicas have diverged
unsigned int adam_p2p_timeout_step() {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, g_p2p_timeout, sizeof(v));
    return v;
}

void launch_init_uniform(float *p, const NetGeom &g, uint64_t seed, cudaStream_t st) {
    for (int l = 0; l < g.n_layers; ++l) k_init_uniform<<<64, 256, 0, st>>>(p, g.L[l], seed, l);
}
