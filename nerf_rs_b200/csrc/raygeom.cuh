// raygeom.cuh -- the reference's per-point ray arithmetic (src/ray_sampling.rs), op for op, shared by the standalone
// sampler (sampling.cu) and the fused MLP prologue (mlp_tc2.cu). Every f32 operation is an explicit round-to-nearest
// intrinsic (no FMA contraction) so both users produce bit-identical points.
#pragma once
#include "common.cuh"

namespace raygeom {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

// rotateYaw = row_mat3x4_transform_pos3 (ray_sampling.rs:20-26), then rotatePitch = col_mat3_transform (:68)
__device__ __forceinline__ void rotate_yaw_pitch(const ViewPose &vp, const float v[3], float out[3]) {
    float y3[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        y3[i] = add(add(add(mul(vp.yaw[i][0], v[0]), mul(vp.yaw[i][1], v[1])), mul(vp.yaw[i][2], v[2])), vp.yaw[i][3]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        out[i] = add(add(mul(vp.pitch[0][i], y3[0]), mul(vp.pitch[1][i], y3[1])), mul(vp.pitch[2][i], y3[2]));
}

// p = FROM + to*t (:115, FROM = (0,0,-1), unfused mul then add), rotated about the world origin (:128-132)
__device__ __forceinline__ void sample_point(const ViewPose &vp, const float to[3], float t, float out[3]) {
    const float p[3] = {add(0.f, mul(to[0], t)), add(0.f, mul(to[1], t)), add(-1.f, mul(to[2], t))};
    rotate_yaw_pitch(vp, p, out);
}

}  // namespace raygeom
