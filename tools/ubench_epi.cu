// ubench_epi.cu -- microbenchmark of the MLP epilogue's building blocks on one SM:
// tcgen05.ld throughput, bias fetch path (smem vs global/L1), bf16 pack, swizzled st.shared.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nerf_rs_b200/csrc tools/ubench_epi.cu -o gpurun_out/ubench_epi
// (built on the GPU box by tools/gpu_ubench.sh). Prints cycles per 128x256 fp32 accumulator tile.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t panel_chunk_addr(uint32_t slot_addr, uint32_t row, uint32_t chunk) {
    return slot_addr + row * 128u + (((chunk ^ row) & 7u) << 4);
}

// mode bits: 1 = tcgen05.ld, 2 = bias from smem, 4 = bias from global (__ldg), 8 = relu+pack, 16 = st.shared, 32 = sign mask
// kMma: warp 1 keeps the tensor pipe busy (M=128 x N=256 x K=16 SS MMAs on smem garbage, accumulating into the
// OTHER 256 TMEM columns) while the epilogue warps run -- measures smem/TMEM contention between the two.
template <int kWarps, int kMode, bool kMma>
__global__ void __launch_bounds__(kWarps * 32 + 128, 1) k_epi(const float *gbias, int iters, unsigned long long *out, uint32_t *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile uint32_t stop_flag;
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ __align__(16) float s_bias[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = ptx::smem_u32(smem);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bias[i] = gbias[i];
    if (threadIdx.x == 0) { stop_flag = 0; ptx::mbar_init(ptx::smem_u32(&mma_bar), 1); ptx::fence_mbar_init(); }
    if (warp == 2) ptx::tmem_alloc<512>(ptx::smem_u32(&tmem_ptr));
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    unsigned long long t0 = 0, t1 = 0;
    uint32_t acc_sink = 0;
    if (kMma && warp == 1) {
        // A = panel 4 (16 KB), B = panels 5,6 (32 KB chunk) of the dynamic smem
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 256, 0, 0);
        const uint64_t ad0 = ptx::umma_desc_sw128(sbase + 4 * 16384u, 16, 1024);
        const uint64_t bd0 = ptx::umma_desc_sw128(sbase + 5 * 16384u, 16, 1024);
        unsigned long long n_mma = 0;
        uint32_t ph = 0;
        const unsigned long long tm0 = clock64();
        while (!stop_flag) {
            if (ptx::elect_one()) {
                for (int rep = 0; rep < 4; ++rep) {
                    ptx::umma_ss(tmem_base + 256u, ad0, bd0, idesc, 1u);
                    ptx::umma_ss(tmem_base + 256u, ad0 + 2u, bd0 + 2u, idesc, 1u);
                    ptx::umma_ss(tmem_base + 256u, ad0 + 4u, bd0 + 4u, idesc, 1u);
                    ptx::umma_ss(tmem_base + 256u, ad0 + 6u, bd0 + 6u, idesc, 1u);
                }
                ptx::umma_commit(ptx::smem_u32(&mma_bar));
            }
            __syncwarp();
            ptx::mbar_wait(ptx::smem_u32(&mma_bar), ph);   // keeps at most 16 MMAs queued
            ph ^= 1u;
            n_mma += 16;
        }
        const unsigned long long tm1 = clock64();
        if (lane == 0 && blockIdx.x == 0) { out[1] = n_mma; out[2] = tm1 - tm0; }
    }
    if (warp >= 4) {
        const uint32_t we = warp - 4;
        const uint32_t q = we & 3u, h = we >> 2;           // lane quarter, column split
        constexpr int kSplit = kWarps / 4;                  // column split ways
        constexpr int kGroups = 8 / kSplit;                 // 32-column groups per warp per tile
        const uint32_t row = q * 32u + lane;
        const uint32_t taddr = tmem_base + ((q * 32u) << 16);
        ptx::named_bar_sync(1, kWarps * 32);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t tcol = kMma ? 0u : (it & 1) * 256u;
#pragma unroll 1
            for (int gi = 0; gi < kGroups; gi += 2) {
                const int G = h * kGroups + gi;
                uint32_t r0[32], r1[32];
                if (kMode & 1) {
                    ptx::tmem_ld32(taddr + tcol + G * 32, r0);
                    if (kGroups > 1) ptx::tmem_ld32(taddr + tcol + (G + 1) * 32, r1);
                    ptx::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) { r0[i] = it * 7 + i + lane; r1[i] = it * 5 + i; }
                }
#pragma unroll
                for (int half = 0; half < (kGroups > 1 ? 2 : 1); ++half) {
                    const uint32_t(&r)[32] = half ? r1 : r0;
                    const int GG = G + half;
                    uint32_t w[16];
                    uint32_t signs = 0;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (kMode & 2) b = reinterpret_cast<const float4 *>(s_bias + GG * 32)[j4];
                        if (kMode & 4) b = __ldg(reinterpret_cast<const float4 *>(gbias + GG * 32) + j4);
                        const float v0 = __uint_as_float(r[4 * j4 + 0]) + b.x, v1 = __uint_as_float(r[4 * j4 + 1]) + b.y;
                        const float v2 = __uint_as_float(r[4 * j4 + 2]) + b.z, v3 = __uint_as_float(r[4 * j4 + 3]) + b.w;
                        if (kMode & 32) {
                            signs = __funnelshift_l(__float_as_uint(v0), signs, 1);
                            signs = __funnelshift_l(__float_as_uint(v1), signs, 1);
                            signs = __funnelshift_l(__float_as_uint(v2), signs, 1);
                            signs = __funnelshift_l(__float_as_uint(v3), signs, 1);
                        }
                        if (kMode & 8) {
                            w[2 * j4] = ptx::pack_bf16x2_relu(v0, v1);
                            w[2 * j4 + 1] = ptx::pack_bf16x2_relu(v2, v3);
                        } else {
                            w[2 * j4] = __float_as_uint(v0) ^ __float_as_uint(v1);
                            w[2 * j4 + 1] = __float_as_uint(v2) ^ __float_as_uint(v3);
                        }
                    }
                    acc_sink ^= signs;
                    if (kMode & 16) {
                        const uint32_t slot = sbase + (GG >> 1) * 16384u;
                        const uint32_t cb = (GG & 1) * 4u;
#pragma unroll
                        for (int c = 0; c < 4; ++c) st_shared_v4(panel_chunk_addr(slot, row, cb + c), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) acc_sink ^= w[c];
                    }
                }
            }
        }
        t1 = clock64();
        ptx::named_bar_sync(1, kWarps * 32);
        if (we == 0 && lane == 0) stop_flag = 1;
        if (we == 0 && lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
        if (acc_sink == 0x12345678u) sink[threadIdx.x] = acc_sink;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

template <int kWarps, int kMode, bool kMma = false>
void run(const char *name, const float *gbias, unsigned long long *d_out, uint32_t *sink, int grid) {
    const int iters = 200;
    const size_t smem = 7 * 16384 + 1024;
    cudaFuncSetAttribute(k_epi<kWarps, kMode, kMma>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(d_out, 0, 64);
    k_epi<kWarps, kMode, kMma><<<grid, kWarps * 32 + 128, smem>>>(gbias, iters, d_out, sink);
    k_epi<kWarps, kMode, kMma><<<grid, kWarps * 32 + 128, smem>>>(gbias, iters, d_out, sink);
    unsigned long long h[3] = {0, 0, 0};
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost);
    printf("%-46s warps %2d grid %3d mma %d : %7.0f cycles per 128x256 tile", name, kWarps, grid, (int)kMma, (double)h[0] / iters);
    if (kMma) printf("  | MMA: %.1f cycles per K16 MMA (floor 128)", h[1] ? (double)h[2] / (double)h[1] : 0.0);
    printf("  (%s)\n", cudaGetErrorString(e));
}

int main() {
    float *gbias;
    unsigned long long *d_out;
    uint32_t *sink;
    cudaMalloc(&gbias, 4096 * 4);
    cudaMemset(gbias, 0, 4096 * 4);
    cudaMalloc(&d_out, 64);
    cudaMalloc(&sink, 4096 * 4);
    for (int grid : {1, 148}) {
        run<8, 1>("tcgen05.ld only", gbias, d_out, sink, grid);
        run<16, 1>("tcgen05.ld only", gbias, d_out, sink, grid);
        run<8, 8>("pack only (no ld, no bias, no st)", gbias, d_out, sink, grid);
        run<8, 8 + 16>("pack + st.shared", gbias, d_out, sink, grid);
        run<8, 1 + 8 + 16>("ld + pack + st", gbias, d_out, sink, grid);
        run<8, 1 + 2 + 8 + 16>("ld + bias(smem) + pack + st", gbias, d_out, sink, grid);
        run<8, 1 + 4 + 8 + 16>("ld + bias(global) + pack + st", gbias, d_out, sink, grid);
        run<8, 1 + 2 + 8 + 16 + 32>("ld + bias(smem) + pack + st + signmask", gbias, d_out, sink, grid);
        run<16, 1 + 8 + 16>("ld + pack + st", gbias, d_out, sink, grid);
        run<16, 1 + 2 + 8 + 16>("ld + bias(smem) + pack + st", gbias, d_out, sink, grid);
        run<16, 1 + 4 + 8 + 16>("ld + bias(global) + pack + st", gbias, d_out, sink, grid);
        run<16, 1 + 2 + 8 + 16 + 32>("ld + bias(smem) + pack + st + signmask", gbias, d_out, sink, grid);
        run<8, 0, true>("nothing (MMA alone)", gbias, d_out, sink, grid);
        run<8, 1, true>("tcgen05.ld only", gbias, d_out, sink, grid);
        run<8, 8 + 16, true>("pack + st.shared", gbias, d_out, sink, grid);
        run<8, 1 + 8 + 16, true>("ld + pack + st", gbias, d_out, sink, grid);
        run<8, 1 + 2 + 8 + 16, true>("ld + bias(smem) + pack + st", gbias, d_out, sink, grid);
        run<16, 1 + 8 + 16, true>("ld + pack + st", gbias, d_out, sink, grid);
        run<16, 1 + 2 + 8 + 16, true>("ld + bias(smem) + pack + st", gbias, d_out, sink, grid);
    }
    return 0;
}
