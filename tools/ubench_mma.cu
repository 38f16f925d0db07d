// ubench_mma.cu -- tensor-pipe rate of the CTA-pair MMAs the chain kernels issue (cta_group::2, M = 256, K = 16, bf16):
// SS (A and B from shared memory) vs TS (A from tensor memory), N = 64/128/256, with `batch` MMAs between commits + waits
// (batch = 0: one commit at the very end = a free-running issue thread). Operands are garbage; only time matters.
// Build (on the GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nerf_rs_b200/csrc tools/ubench_mma.cu -o gpurun_out/ubench_mma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

template <bool kTS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k_mma(int n, int total, int batch, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_ptr;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = ptx::smem_u32(smem);
    const uint32_t rank = ptx::cluster_ctarank();
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
    if (warp == 2) ptx::tmem_alloc2<512>(ptx::smem_u32(&tmem_ptr));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(256, (uint32_t)n, 0, 0);
        const uint64_t ad0 = ptx::umma_desc_sw128(sbase, 16, 1024);             // 16 KB A panel
        const uint64_t bd0 = ptx::umma_desc_sw128(sbase + 16384u, 16, 1024);    // <= 16 KB half chunk
        uint32_t ph = 0;
        unsigned long long t_issue = 0;
        const unsigned long long t0 = clock64();
        int done = 0;
        while (done < total) {
            const int nb = batch > 0 ? batch : total;
            const unsigned long long ti = clock64();
            if (ptx::elect_one()) {
                for (int i = 0; i < nb; ++i) {
                    const uint32_t k = (uint32_t)(i & 3);
                    if (kTS) ptx::umma_ts2(tmem_base, tmem_base + 256u + 8u * k, bd0 + 2u * k, idesc, 1u);
                    else ptx::umma_ss2(tmem_base, ad0 + 2u * k, bd0 + 2u * k, idesc, 1u);
                }
                ptx::umma_commit2_mc(ptx::smem_u32(&bar), 1);
            }
            __syncwarp();
            t_issue += clock64() - ti;
            ptx::mbar_wait(ptx::smem_u32(&bar), ph);
            ph ^= 1u;
            done += nb;
        }
        const unsigned long long t1 = clock64();
        if (lane == 0) { out[2 * (blockIdx.x >> 1)] = t1 - t0; out[2 * (blockIdx.x >> 1) + 1] = t_issue; }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2<512>(tmem_base);
}

// Free-running "steps" of 16 TS MMAs (N = 128) with `commits` tcgen05.commit per step spread evenly (to barriers nobody waits
// on; flavour 0: cta_group::2 multicast to both CTAs, 1: multicast to the leader only), one wait at the very end:
// does a commit stall the issuing thread / the tensor pipe?
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k_commit(int steps, int commits, int flavour, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_ptr;
    __shared__ __align__(8) uint64_t bar[9];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = ptx::smem_u32(smem);
    const uint32_t rank = ptx::cluster_ctarank();
    if (threadIdx.x == 0) { for (int i = 0; i < 9; ++i) ptx::mbar_init(ptx::smem_u32(&bar[i]), 1); ptx::fence_mbar_init(); }
    if (warp == 2) ptx::tmem_alloc2<512>(ptx::smem_u32(&tmem_ptr));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (flavour & 2) {   // random bf16 operands in [-1, 1] instead of whatever the memories hold: does the data change the rate?
        uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
        auto rnd = [&]() { x ^= x << 13; x ^= x >> 17; x ^= x << 5; const uint32_t lo = 0x3f00u | ((x >> 3) & 0x80ffu), hi = 0x3f00u | ((x >> 19) & 0x80ffu); return lo | (hi << 16); };
        for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = rnd();
        uint32_t r[16];
        for (int c = 0; c < 8; ++c) {
            for (int i = 0; i < 16; ++i) r[i] = rnd();
            ptx::tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + 256u + 16u * c, r);
        }
        ptx::tmem_st_wait();
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncthreads();
        ptx::tc_fence_after();
    }
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(256, 128, 0, 0);
        const uint64_t bd0 = ptx::umma_desc_sw128(sbase + 16384u, 16, 1024);
        const unsigned long long t0 = clock64();
        const int every = commits > 0 ? 16 / commits : 1 << 30;
        if (ptx::elect_one()) {
            for (int s = 0; s < steps; ++s) {
                for (int i = 0; i < 16; ++i) {
                    ptx::umma_ts2(tmem_base, tmem_base + 256u + 8u * (uint32_t)(i & 3), bd0 + 2u * (uint32_t)(i & 3), idesc, 1u);
                    if ((i + 1) % every == 0) ptx::umma_commit2_mc(ptx::smem_u32(&bar[(i / every) & 7]), (flavour & 1) == 0 ? 3 : 1);
                }
            }
            ptx::umma_commit2_mc(ptx::smem_u32(&bar[8]), 1);
        }
        __syncwarp();
        const unsigned long long t1 = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bar[8]), 0);
        const unsigned long long t2 = clock64();
        if (lane == 0) { out[2 * (blockIdx.x >> 1)] = t2 - t0; out[2 * (blockIdx.x >> 1) + 1] = t1 - t0; }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc2<512>(tmem_base);
}

int main() {
    unsigned long long *d_out, h_out[2 * 74];
    cudaMalloc(&d_out, sizeof(h_out));
    const int smem = 49152;
    cudaFuncSetAttribute(k_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int total = 4096;
    printf("mode  N  batch  clk/MMA (median cluster)  issue clk/MMA   [ideal: N/2 clk at 8192 MAC/clk/pair]\n");
    for (int ts = 0; ts < 2; ++ts)
        for (int n : {64, 128, 256})
            for (int batch : {0, 16}) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (ts) k_mma<true><<<148, 128, smem>>>(n, total, batch, d_out);
                    else k_mma<false><<<148, 128, smem>>>(n, total, batch, d_out);
                }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost);
                unsigned long long v[74], w[74];
                for (int i = 0; i < 74; ++i) { v[i] = h_out[2 * i]; w[i] = h_out[2 * i + 1]; }
                for (int i = 0; i < 74; ++i) for (int j = i + 1; j < 74; ++j) if (v[j] < v[i]) { auto t = v[i]; v[i] = v[j]; v[j] = t; t = w[i]; w[i] = w[j]; w[j] = t; }
                printf("%s  %3d  %4d   %8.1f   %8.1f\n", ts ? "TS" : "SS", n, batch, (double)v[37] / total, (double)w[37] / total);
            }
    cudaFuncSetAttribute(k_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("free-running steps of 16 TS MMAs (N=128): commits/step  flavour  clk/step total  clk/step issue   [ideal 1024]\n");
    for (int flavour : {0, 2})
        for (int commits : {0, 4}) {
            for (int rep = 0; rep < 2; ++rep) k_commit<<<148, 128, smem>>>(flavour ? 4096 : 256, commits, flavour, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost);
            double a = 0, b = 0;
            for (int i = 0; i < 74; ++i) { a += (double)h_out[2 * i]; b += (double)h_out[2 * i + 1]; }
            printf("  %2d  %d   %8.1f   %8.1f\n", commits, flavour, a / 74 / (flavour ? 4096 : 256), b / 74 / (flavour ? 4096 : 256));
        }
    return 0;
}
