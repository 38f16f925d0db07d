// ubench_s2g.cu -- how fast can one SM push shared memory to HBM?
//   mode 0: cp.async.bulk S2G, one thread, 64 KB per copy, `depth` copies in flight
//   mode 1: st.global.v4 from 16 warps, thread = panel row, four 16-byte chunks of a 64-byte half-row per "group"
//           (the register-direct alternative: what the epilogue warps could store without the copy engine)
// Prints bytes/cycle/SM and chip GB/s at grid = 148 (every SM busy, like the chain kernel).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

__global__ void __launch_bounds__(640, 1) k_s2g(uint8_t *dst, int iters, int depth, int mode, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = ptx::smem_u32(smem);
    for (int i = threadIdx.x; i < 65536 * 2 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(i, i, i, i);
    ptx::fence_proxy_async_smem();
    __syncthreads();
    const size_t per_cta = (size_t)iters * 65536;
    uint8_t *base = dst + (size_t)blockIdx.x * per_cta;
    const unsigned long long t0 = clock64();
    if (mode == 0) {
        if (threadIdx.x == 0) {
            for (int it = 0; it < iters; ++it) {
                ptx::bulk_s2g(base + (size_t)it * 65536, sbase + (it & 1) * 65536, 65536);
                ptx::bulk_commit();
                if (depth == 1) ptx::bulk_wait_read<0>();
                else ptx::bulk_wait_read<1>();
            }
            ptx::bulk_wait_all<0>();
        }
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp >= 4) {
            const int we = warp - 4, q = we & 3, h = we >> 2;
            const int row = q * 32 + lane;
            for (int it = 0; it < iters; ++it) {
                uint8_t *tile = base + (size_t)it * 65536 + (size_t)h * 16384;   // this warp's panel
#pragma unroll
                for (int g = 0; g < 2; ++g) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int chunk = g * 4 + c;
                        uint4 v = make_uint4(it, row, chunk, lane);
                        *reinterpret_cast<uint4 *>(tile + row * 128 + (((chunk ^ row) & 7) << 4)) = v;
                    }
                }
            }
        }
    }
    __syncthreads();
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    const int grid = 148, iters = 256;
    uint8_t *dst;
    unsigned long long *d_out;
    cudaMalloc(&dst, (size_t)grid * iters * 65536);
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(k_s2g, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024);
    for (int mode = 0; mode < 2; ++mode) {
        for (int depth = 1; depth <= (mode == 0 ? 2 : 1); ++depth) {
            for (int g : {1, 148}) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                k_s2g<<<g, 640, 131072 + 1024>>>(dst, iters, depth, mode, d_out);
                cudaEventRecord(e0);
                k_s2g<<<g, 640, 131072 + 1024>>>(dst, iters, depth, mode, d_out);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                unsigned long long cyc = 0;
                cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
                printf("mode %d (%s) depth %d grid %3d: %6.1f B/cycle/SM, %7.1f GB/s chip  (%s)\n", mode, mode ? "st.global.v4 x16 warps" : "cp.async.bulk S2G",
                       depth, g, (double)iters * 65536 / (double)cyc, (double)g * iters * 65536 / (ms * 1e-3) / 1e9, cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
