// mlp_simt.cu -- CUDA-core MLP forward/backward (NERF_MLP_SIMT).
//
// Same network as the tcgen05 path (DensityNet/RadianceNet, src/model.rs:44-131, with the
// north-star skip/direction options), written as plain tiled f32 GEMMs. It exists to
// cross-check the fused tensor-core kernels ON THE DEVICE at sizes the CPU oracle cannot
// reach, and to validate layouts in pure fp32 (round_bf16 == 0) against the torch oracle.
// It is a CUDA path, not a fallback: the product default is NERF_MLP_TCGEN05.
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace {

constexpr int TS = 32;  // tile edge
constexpr int TK = 16;

__device__ __forceinline__ float rnd(float x, int on) { return on ? ptx::bf16_round(x) : x; }

// C[m,n] (+)= sum_k A(m,k) * W[n*ldw + k]; A row = m / arep. Epilogue: bias, act, rounding.
// act: 0 none, 1 relu, 2 sigmoid
__global__ void __launch_bounds__(TS * 8)
k_gemm_nt(const float *__restrict__ A, int lda, int arep, const float *__restrict__ W, int ldw, const float *__restrict__ bias,
          float *__restrict__ C, int ldc, int64_t M, int N, int K, int accumulate, int act, int round_w, int round_out) {
    __shared__ float sA[TK][TS + 1];
    __shared__ float sW[TK][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty 0..7, each thread 4 rows
    const int64_t m0 = (int64_t)blockIdx.x * TS;
    const int n0 = blockIdx.y * TS;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int e = threadIdx.x; e < TS * TK; e += TS * 8) {
            const int kk = e % TK, rr = e / TK;
            const int64_t m = m0 + rr;
            const int k = k0 + kk;
            sA[kk][rr] = (m < M && k < K) ? A[(m / arep) * lda + k] : 0.f;
            const int n = n0 + rr;
            sW[kk][rr] = (n < N && k < K) ? rnd(W[(int64_t)n * ldw + k], round_w) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            const float w = sW[kk][tx];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += sA[kk][ty * 4 + i] * w;
        }
        __syncthreads();
    }
    const int n = n0 + tx;
    if (n >= N) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float v = acc[i];
        if (accumulate) v += C[m * ldc + n];
        if (bias) v += bias[n];
        if (act == 1) v = fmaxf(v, 0.f);
        else if (act == 2) v = 1.f / (1.f + expf(-v));
        C[m * ldc + n] = rnd(v, round_out);
    }
}

// dA[m,k] = sum_n dC[m,n] * W[n*ldw + k]; optional relu mask from H[m,k] > 0; rounding.
__global__ void __launch_bounds__(TS * 8)
k_gemm_nn(const float *__restrict__ dC, int ldc, const float *__restrict__ W, int ldw, const float *__restrict__ H, int ldh,
          float *__restrict__ dA, int lda, int64_t M, int N, int K, int round_w, int round_out) {
    __shared__ float sC[TK][TS + 1];
    __shared__ float sW[TK][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t m0 = (int64_t)blockIdx.x * TS;
    const int k0 = blockIdx.y * TS;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n0 = 0; n0 < N; n0 += TK) {
        for (int e = threadIdx.x; e < TS * TK; e += TS * 8) {
            const int nn = e % TK, rr = e / TK;
            const int64_t m = m0 + rr;
            const int n = n0 + nn;
            sC[nn][rr] = (m < M && n < N) ? dC[m * ldc + n] : 0.f;
        }
        for (int e = threadIdx.x; e < TS * TK; e += TS * 8) {
            const int kk = e % TS, nn = e / TS;
            const int n = n0 + nn, k = k0 + kk;
            sW[nn][kk] = (n < N && k < K) ? rnd(W[(int64_t)n * ldw + k], round_w) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int nn = 0; nn < TK; ++nn) {
            const float w = sW[nn][tx];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += sC[nn][ty * 4 + i] * w;
        }
        __syncthreads();
    }
    const int k = k0 + tx;
    if (k >= K) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float v = acc[i];
        if (H && !(H[m * ldh + k] > 0.f)) v = 0.f;
        dA[m * lda + k] = rnd(v, round_out);
    }
}

// dW[n*ldw + k] += sum_m dC[m,n] * A(m,k). One block per (n-tile, k-tile, m-chunk); atomics over chunks.
__global__ void __launch_bounds__(TS * 8)
k_gemm_tn(const float *__restrict__ dC, int ldc, const float *__restrict__ A, int lda, int arep, float *__restrict__ dW, int ldw,
          int64_t M, int N, int K, int64_t m_chunk) {
    __shared__ float sC[TK][TS + 1];
    __shared__ float sA[TK][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n0 = blockIdx.x * TS, k0 = blockIdx.y * TS;
    const int64_t mb = (int64_t)blockIdx.z * m_chunk;
    const int64_t me = (mb + m_chunk < M) ? mb + m_chunk : M;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t m0 = mb; m0 < me; m0 += TK) {
        for (int e = threadIdx.x; e < TS * TK; e += TS * 8) {
            const int cc = e % TS, mm = e / TS;
            const int64_t m = m0 + mm;
            const int n = n0 + cc, k = k0 + cc;
            sC[mm][cc] = (m < me && n < N) ? dC[m * ldc + n] : 0.f;
            sA[mm][cc] = (m < me && k < K) ? A[(m / arep) * lda + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < TK; ++mm) {
            const float a = sA[mm][tx];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += sC[mm][ty * 4 + i] * a;
        }
        __syncthreads();
    }
    const int k = k0 + tx;
    if (k >= K) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n < N) atomicAdd(&dW[(int64_t)n * ldw + k], acc[i]);
    }
}

// db[n] += sum_m dC[m,n]
__global__ void k_colsum(const float *__restrict__ dC, int ldc, float *__restrict__ db, int64_t M, int N, int64_t m_chunk) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int64_t mb = (int64_t)blockIdx.y * m_chunk;
    const int64_t me = (mb + m_chunk < M) ? mb + m_chunk : M;
    float s = 0.f;
    for (int64_t m = mb; m < me; ++m) s += dC[m * ldc + n];
    atomicAdd(&db[n], s);
}

// dz10 = d_rgba * rgba * (1 - rgba)
__global__ void k_sigmoid_bwd(const float *__restrict__ rgba, const float *__restrict__ d_rgba, float *__restrict__ dz, int64_t n,
                              int round_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float y = rgba[i];
    dz[i] = rnd(d_rgba[i] * y * (1.f - y), round_out);
}

// ddf[m, 0] = dsigma[m]; ddf[m, 1..W] = dfeat[m, :]
__global__ void k_join_sigma_feat(const float *__restrict__ dsig, const float *__restrict__ dfeat, int W, float *__restrict__ ddf,
                                  int64_t M, int round_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * (W + 1)) return;
    const int64_t m = i / (W + 1);
    const int c = (int)(i % (W + 1));
    ddf[i] = (c == 0) ? rnd(dsig[m], round_out) : (dfeat ? dfeat[m * W + (c - 1)] : 0.f);
}

__global__ void k_split_sigma(const float *__restrict__ df, int W, float *__restrict__ sigma, int64_t M) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) sigma[m] = df[m * (W + 1)];
}

__global__ void k_round_inplace(float *x, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = ptx::bf16_round(x[i]);
}

inline dim3 grid2(int64_t M, int N) { return dim3((unsigned)((M + TS - 1) / TS), (unsigned)((N + TS - 1) / TS)); }

void gemm_nt(const float *A, int lda, int arep, const float *W, int ldw, const float *bias, float *C, int ldc, int64_t M,
             int N, int K, int accumulate, int act, int rw, int ro, cudaStream_t st) {
    k_gemm_nt<<<grid2(M, N), TS * 8, 0, st>>>(A, lda, arep, W, ldw, bias, C, ldc, M, N, K, accumulate, act, rw, ro);
}
void gemm_nn(const float *dC, int ldc, const float *W, int ldw, const float *H, int ldh, float *dA, int lda, int64_t M, int N,
             int K, int rw, int ro, cudaStream_t st) {
    k_gemm_nn<<<grid2(M, K), TS * 8, 0, st>>>(dC, ldc, W, ldw, H, ldh, dA, lda, M, N, K, rw, ro);
}
void gemm_tn(const float *dC, int ldc, const float *A, int lda, int arep, float *dW, int ldw, float *db, int64_t M, int N,
             int K, cudaStream_t st) {
    const int64_t m_chunk = 4096;
    const unsigned chunks = (unsigned)((M + m_chunk - 1) / m_chunk);
    dim3 g((N + TS - 1) / TS, (K + TS - 1) / TS, chunks);
    k_gemm_tn<<<g, TS * 8, 0, st>>>(dC, ldc, A, lda, arep, dW, ldw, M, N, K, m_chunk);
    if (db) {
        dim3 gb((N + 127) / 128, chunks);
        k_colsum<<<gb, 128, 0, st>>>(dC, ldc, db, M, N, m_chunk);
    }
}

// slab indices inside buf.act (each slab is act_stride floats = B * maxw)
enum { SL_H1 = 0, SL_DF = 7, SL_H9 = 8, SL_COUNT = 9 };

}  // namespace

size_t simt_act_floats_per_sample(const NetGeom &g) { return (size_t)(g.W + 1); }

void simt_mlp_forward(const NetGeom &g, const float *params, const float *points, const float *dirs, int64_t num_rays,
                      int num_samples, int rb, SimtBuffers &buf, float *sigma, float *rgba, cudaStream_t st) {
    const int64_t B = num_rays * num_samples;
    const int W = g.W, Cx = g.Cx, Cd = g.Cd;
    launch_encode(points, buf.x_enc, B, g.xyz_freqs, 1, st);
    if (Cd) launch_encode(dirs, buf.d_enc, num_rays, g.dir_freqs, 1, st);
    if (rb) {
        k_round_inplace<<<(unsigned)((B * Cx + 255) / 256), 256, 0, st>>>(buf.x_enc, B * Cx);
        if (Cd) k_round_inplace<<<(unsigned)((num_rays * Cd + 255) / 256), 256, 0, st>>>(buf.d_enc, num_rays * Cd);
    }
    const float *h = buf.x_enc;
    int hk = Cx;
    for (int l = 1; l <= 7; ++l) {
        const LayerGeom &L = g.L[l - 1];
        float *out = buf.act + (int64_t)(SL_H1 + l - 1) * buf.act_stride;
        if (g.skip_layer && l == g.skip_layer + 1) {
            // input = [x_enc | h]: two K segments of the same weight rows
            gemm_nt(buf.x_enc, Cx, 1, params + L.w_off, L.in_dim, nullptr, out, W, B, W, Cx, 0, 0, rb, 0, st);
            gemm_nt(h, hk, 1, params + L.w_off + Cx, L.in_dim, params + L.b_off, out, W, B, W, W, 1, 1, rb, rb, st);
        } else {
            gemm_nt(h, hk, 1, params + L.w_off, L.in_dim, params + L.b_off, out, W, B, W, hk, 0, 1, rb, rb, st);
        }
        h = out;
        hk = W;
    }
    // fc8: [sigma | feat], no activation (model.rs:113, 168-176). sigma stays f32; feat is an MMA operand.
    const LayerGeom &L8 = g.L[7];
    float *df = buf.act + (int64_t)SL_DF * buf.act_stride;
    gemm_nt(h, W, 1, params + L8.w_off, W, params + L8.b_off, df, W + 1, B, W + 1, W, 0, 0, rb, 0, st);
    k_split_sigma<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(df, W, sigma, B);
    if (!g.use_rgb_head) return;
    if (rb) k_round_inplace<<<(unsigned)((B * (W + 1) + 255) / 256), 256, 0, st>>>(df, B * (W + 1));
    // fc9 on [feat | d_enc] (model.rs:121-124 + direction), fc10 + sigmoid (:125-126)
    const LayerGeom &L9 = g.L[8], &L10 = g.L[9];
    float *h9 = buf.act + (int64_t)SL_H9 * buf.act_stride;
    if (Cd) {
        gemm_nt(df + 1, W + 1, 1, params + L9.w_off, L9.in_dim, nullptr, h9, g.W2, B, g.W2, W, 0, 0, rb, 0, st);
        gemm_nt(buf.d_enc, Cd, num_samples, params + L9.w_off + W, L9.in_dim, params + L9.b_off, h9, g.W2, B, g.W2, Cd, 1, 1,
                rb, rb, st);
    } else {
        gemm_nt(df + 1, W + 1, 1, params + L9.w_off, L9.in_dim, params + L9.b_off, h9, g.W2, B, g.W2, W, 0, 1, rb, rb, st);
    }
    gemm_nt(h9, g.W2, 1, params + L10.w_off, g.W2, params + L10.b_off, rgba, 4, B, 4, g.W2, 0, 2, rb, 0, st);
}

void simt_mlp_backward(const NetGeom &g, const float *params, float *grads, int64_t num_rays, int num_samples, int rb,
                       SimtBuffers &buf, const float *rgba, const float *d_sigma, const float *d_rgba, cudaStream_t st) {
    const int64_t B = num_rays * num_samples;
    const int W = g.W, Cx = g.Cx, Cd = g.Cd;
    float *df = buf.act + (int64_t)SL_DF * buf.act_stride;
    float *h9 = buf.act + (int64_t)SL_H9 * buf.act_stride;
    float *ga = buf.dact;                       // current gradient  [B][<=W+1]
    float *gb = buf.dact + buf.act_stride;      // next gradient
    float *ddf = buf.dact + 2 * buf.act_stride; // [B][W+1]
    const float *dfeat = nullptr;
    if (g.use_rgb_head) {
        const LayerGeom &L9 = g.L[8], &L10 = g.L[9];
        k_sigmoid_bwd<<<(unsigned)((B * 4 + 255) / 256), 256, 0, st>>>(rgba, d_rgba, ga, B * 4, rb);       // dz10 [B][4]
        gemm_tn(ga, 4, h9, g.W2, 1, grads + L10.w_off, g.W2, grads + L10.b_off, B, 4, g.W2, st);
        gemm_nn(ga, 4, params + L10.w_off, g.W2, h9, g.W2, gb, g.W2, B, 4, g.W2, rb, rb, st);             // dz9 [B][W2]
        gemm_tn(gb, g.W2, df + 1, W + 1, 1, grads + L9.w_off, L9.in_dim, grads + L9.b_off, B, g.W2, W, st);
        if (Cd) gemm_tn(gb, g.W2, buf.d_enc, Cd, num_samples, grads + L9.w_off + W, L9.in_dim, nullptr, B, g.W2, Cd, st);
        gemm_nn(gb, g.W2, params + L9.w_off, L9.in_dim, nullptr, 0, ga, W, B, g.W2, W, rb, rb, st);        // dfeat [B][W]
        dfeat = ga;
    }
    k_join_sigma_feat<<<(unsigned)((B * (W + 1) + 255) / 256), 256, 0, st>>>(d_sigma, dfeat, W, ddf, B, rb);
    const LayerGeom &L8 = g.L[7];
    const float *h7 = buf.act + (int64_t)(SL_H1 + 6) * buf.act_stride;
    gemm_tn(ddf, W + 1, h7, W, 1, grads + L8.w_off, W, grads + L8.b_off, B, W + 1, W, st);
    gemm_nn(ddf, W + 1, params + L8.w_off, W, h7, W, ga, W, B, W + 1, W, rb, rb, st);                      // dz7
    float *cur = ga, *nxt = gb;
    for (int l = 7; l >= 1; --l) {
        const LayerGeom &L = g.L[l - 1];
        const float *hin = (l == 1) ? buf.x_enc : buf.act + (int64_t)(SL_H1 + l - 2) * buf.act_stride;
        const int hk = (l == 1) ? Cx : W;
        const bool skip = g.skip_layer && l == g.skip_layer + 1;
        if (skip) {
            gemm_tn(cur, W, buf.x_enc, Cx, 1, grads + L.w_off, L.in_dim, nullptr, B, W, Cx, st);
            gemm_tn(cur, W, hin, W, 1, grads + L.w_off + Cx, L.in_dim, grads + L.b_off, B, W, W, st);
        } else {
            gemm_tn(cur, W, hin, hk, 1, grads + L.w_off, L.in_dim, grads + L.b_off, B, W, hk, st);
        }
        if (l > 1) {
            gemm_nn(cur, W, params + L.w_off + (skip ? Cx : 0), L.in_dim, hin, W, nxt, W, B, W, W, rb, rb, st);
            float *t = cur; cur = nxt; nxt = t;
        }
    }
}
