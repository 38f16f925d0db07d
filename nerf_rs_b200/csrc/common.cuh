// common.cuh -- shared geometry / tables for the nerf_b200 kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define NERF_TILE_M 128          // samples per MLP tile (one UMMA M)
#define NERF_PANEL_BYTES 16384   // [128 rows][64 bf16] 128B-swizzled panel image
#define NERF_MAX_LAYERS 10

// reference constants (src/ray_sampling.rs:10-16)
#define NERF_HITHER 0.05f
#define NERF_T_FAR 2.0f

struct LayerGeom {
    int32_t in_dim, out_dim;
    int64_t w_off, b_off;  // float offsets into the flat parameter blob
};

// Network geometry derived from nerf_config (fc1..fc10, reference order model.rs:48-55, 89-90)
struct NetGeom {
    int32_t W, Wp;        // hidden width and its 64-padded size
    int32_t W2, W2p;      // W/2 and padded
    int32_t Cx, Cd;       // encoded xyz / dir widths (3+6L), Cd = 0 -> no direction input
    int32_t xyz_freqs, dir_freqs;
    int32_t skip_layer;   // 0 none
    int32_t use_rgb_head, sigma_relu;
    int32_t n_layers;     // 10
    LayerGeom L[NERF_MAX_LAYERS];
    int64_t n_params;
};

// Per-view pose table entry: matrices the reference rebuilds per point
// (rotateYaw ray_sampling.rs:20-26, rotatePitch :32-69), computed once on the host.
struct ViewPose {
    float yaw[3][4];
    float pitch[3][3];
    float pad[3];
};

// Per-ray record produced by the sampler and consumed by the fused MLP prologue.
struct RayRec {
    float to[3];      // unit direction in camera frame (screen_to_world)
    int32_t view;     // index into the pose table
};

// ---- Philox4x32-10 (device + host) ----------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__host__ __device__ inline float philox_uniform(uint64_t seed, uint32_t stream, uint64_t index) {
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), stream, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (float)(c[0] >> 8) * (1.0f / 16777216.0f);
}

// ---- programmatic dependent launch (PDL) ---------------------------------------
// The kernels of a training step are launched with the programmatic-stream-serialization attribute: a kernel announces
// at its very start that its successor may begin (pdl_trigger), the successor's CTAs become resident as SMs free up and run
// their private prologue (barrier init, TMEM allocation, table loads), and only then wait for the predecessor grid to have
// completed and flushed (pdl_wait) before touching anything it wrote. NERF_B200_NO_PDL=1 launches plainly (A/B).
#ifdef __CUDACC__
#include <cstdlib>
#include <utility>
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    static const bool off = getenv("NERF_B200_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

#define NERF_STREAM_PIX_Y 0u
#define NERF_STREAM_PIX_X 1u
#define NERF_STREAM_VIEW 2u
#define NERF_STREAM_JITTER 3u
#define NERF_STREAM_INIT 4u
