// chain_common.cuh -- device helpers shared by the two CTA-pair chain kernels (mlp_tc2.cu: SS mode, activations in shared
// memory; mlp_tc3.cu: TS mode, activations in tensor memory): the tile schedule every role derives its order from, the
// slot-E producers (positional encoding and the sparse gradient panels) and the per-32-column epilogue arithmetic.
#pragma once
#include "mlp_tc.h"
#include "ptx.cuh"

namespace chain {

constexpr uint32_t kSlotBytes = NERF_PANEL_BYTES;

// Both lanes walk the same per-tile step list in lockstep: group k = (lane 0: step g of its tile, lane 1: step g of
// its tile). While both lanes have tiles a group holds two steps that use the SAME weight chunks, so a chunk is
// loaded once and consumed by lane 0's MMAs and then by lane 1's (it is released by the second consumer). Every role
// (producer, relay, MMA issuer, epilogue, store warp) derives its order from this one deterministic schedule.
struct LaneSched {
    int pair0, pair1;     // current pair-tile of lane 0 / lane 1 (lane l starts at cluster + l*C, stride 2C)
    int pos, stride, n_pairs, n_pos, wide;   // wide: a single lane per CTA walks every pair tile of the cluster (stride C)
    __device__ LaneSched(int cluster, int n_clusters, int n_pairs_, int n_pos_, int wide_)
        : pair0(cluster), pair1(cluster + n_clusters), pos(0), stride(wide_ ? n_clusters : 2 * n_clusters), n_pairs(n_pairs_),
          n_pos(n_pos_), wide(wide_) {}
    // next group: step index g, number of live lanes (1 or 2: lane 1 never outlives lane 0), their pair tiles
    __device__ bool next(int &g, int &n_lanes, int &pr0, int &pr1) {
        if (pair0 >= n_pairs) return false;
        g = pos;
        pr0 = pair0;
        pr1 = pair1;
        n_lanes = (!wide && pair1 < n_pairs) ? 2 : 1;
        if (++pos == n_pos) {
            pos = 0;
            pair0 += stride;
            if (!wide) pair1 += stride;
        }
        return true;
    }
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Saved panels are written once and read once, a whole kernel later, by the weight-gradient kernel: streaming stores
// (st.global.cs, evict-first) keep 1.2 GB per kernel from displacing what IS reused in L2 (-0.8 % of the step at burst clocks,
// -0.45 % sustained, same-box A/B; streaming only the forward kernel's panels gains half of that).
__device__ __forceinline__ void st_global_v4(uint8_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 128B-swizzled K-major panel [128 rows][64 bf16]: the smem image of an SS-mode A operand
__device__ __forceinline__ uint32_t panel_chunk_addr(uint32_t slot_addr, uint32_t row, uint32_t chunk) {
    return slot_addr + row * 128u + (((chunk ^ row) & 7u) << 4);
}
// "Chunk-major" panel image [half tile (64 rows)][16-byte chunk (8)][row (64)][16 B]: what the weight-gradient kernel
// consumes as a no-swizzle MN-major operand. The 32 rows of a warp are 512 contiguous bytes per chunk.
constexpr uint32_t kCmChunkStride = 1024;   // 64 rows x 16 B
__device__ __forceinline__ uint32_t cm_row_off(uint32_t row) { return (row >> 6) * (kSlotBytes / 2) + (row & 63u) * 16u; }

// Positional encoding [v, sin(2^k v), cos(2^k v)]_k of a 3-vector, zero padded to 64 features, written as the bf16
// 16-byte chunks [kCh0, kCh1) of a panel row (each column-slice warp writes its own chunks). Every feature is evaluated
// independently with the SFU (sin.approx / cos.approx of the exactly scaled argument 2^k v): at |2^k v| <= ~1.5e3 the
// absolute error is ~1e-4, an order of magnitude below the bf16 rounding the value gets next, and there is no serial
// dependency. (An accurate sincosf + double-angle recurrence, inlined at three call sites, made the kernel 155 KB of
// SASS and cost 1-5 k cycles per tile on the critical path; a compact run-time-indexed loop was no better.)
// gsave != NULL: the row's position in a chunk-major image of the panel in global memory (saved for the weight gradients).
template <int kCh0, int kCh1>
__device__ __forceinline__ void encode_chunks(uint32_t slot_addr, uint32_t row, float v0, float v1, float v2, int freqs, uint8_t *gsave) {
    const int n_feat = 3 + 6 * freqs;
#pragma unroll
    for (int ch = kCh0; ch < kCh1; ++ch) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int fi = 8 * ch + e;                                 // compile-time after unrolling
            const int g = fi - 3;
            const int k = g / 6, r = g - 6 * k;                        // octave, slot within the octave (sin xyz | cos xyz)
            const int d = fi < 3 ? fi : (r < 3 ? r : r - 3);
            const float x = d == 0 ? v0 : (d == 1 ? v1 : v2);
            const float arg = x * (float)(1 << (fi < 3 ? 0 : k));      // x * 2^k, exact
            const float val = fi < 3 ? x : (r < 3 ? __sinf(arg) : __cosf(arg));
            f[e] = fi < n_feat ? val : 0.f;
        }
        const uint32_t w0 = ptx::pack_bf16x2(f[0], f[1]), w1 = ptx::pack_bf16x2(f[2], f[3]), w2 = ptx::pack_bf16x2(f[4], f[5]),
                       w3 = ptx::pack_bf16x2(f[6], f[7]);
        st_shared_v4(panel_chunk_addr(slot_addr, row, (uint32_t)ch), w0, w1, w2, w3);
        if (gsave) st_global_v4(gsave + (uint32_t)ch * kCmChunkStride, w0, w1, w2, w3);
    }
}
// column slice h of 4 writes chunks [2h, 2h + 2) of the row
__device__ __forceinline__ void encode_row(int h, uint32_t slot_addr, uint32_t row, float v0, float v1, float v2, int freqs,
                                           uint8_t *gsave = nullptr) {
    switch (h) {
        case 0: encode_chunks<0, 2>(slot_addr, row, v0, v1, v2, freqs, gsave); break;
        case 1: encode_chunks<2, 4>(slot_addr, row, v0, v1, v2, freqs, gsave); break;
        case 2: encode_chunks<4, 6>(slot_addr, row, v0, v1, v2, freqs, gsave); break;
        default: encode_chunks<6, 8>(slot_addr, row, v0, v1, v2, freqs, gsave); break;
    }
}
// panel whose only non-zero entries are the first four bf16 of each row; this warp writes chunks [ch0, ch1)
__device__ __forceinline__ void write_sparse_panel(uint32_t slot_addr, uint32_t row, int ch0, int ch1, uint32_t w0, uint32_t w1,
                                                   uint8_t *gsave = nullptr) {
    for (int ch = ch0; ch < ch1; ++ch) {
        st_shared_v4(panel_chunk_addr(slot_addr, row, (uint32_t)ch), ch == 0 ? w0 : 0u, ch == 0 ? w1 : 0u, 0u, 0u);
        if (gsave) st_global_v4(gsave + (uint32_t)ch * kCmChunkStride, ch == 0 ? w0 : 0u, ch == 0 ? w1 : 0u, 0u, 0u);
    }
}

// 32 accumulator columns -> 16 packed bf16x2 words (bias + ReLU/linear, or ReLU-mask select for the backward chain).
// The bias comes from the constant bank `cbias` at a warp-uniform index (bidx, a multiple of 4 floats).
template <bool kSave, uint8_t kKind>
__device__ __forceinline__ void epi_group(const float *cbias, const uint32_t (&r)[32], int bidx, uint32_t &mask, uint32_t (&w)[16]) {
    if (kKind == EK_RELU || kKind == EK_LINEAR) {
        uint32_t signs = 0;
#pragma unroll
        for (int j2 = 0; j2 < 16; ++j2) {
            // packed fp32x2 add (sm_100): one instruction and one 64-bit constant operand per column pair
            // (bias offsets are multiples of 16 floats: one 128-bit uniform constant load serves two column pairs)
            const float4 b4 = reinterpret_cast<const float4 *>(cbias)[(bidx >> 2) + (j2 >> 1)];
            const float2 b = (j2 & 1) ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y);
            unsigned long long acc2, bias2, sum2;
            asm("mov.b64 %0, {%1, %2};" : "=l"(acc2) : "r"(r[2 * j2]), "r"(r[2 * j2 + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(bias2) : "f"(b.x), "f"(b.y));
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sum2) : "l"(acc2), "l"(bias2));
            float v0, v1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(sum2));
            if (kSave && kKind == EK_RELU) {
                signs = __funnelshift_l(__float_as_uint(v0), signs, 1);
                signs = __funnelshift_l(__float_as_uint(v1), signs, 1);
            }
            w[j2] = (kKind == EK_RELU) ? ptx::pack_bf16x2_relu(v0, v1) : ptx::pack_bf16x2(v0, v1);
        }
        mask = ~signs;  // bit (31 - col) set = pre-activation sign bit clear
    } else {
        const uint32_t m = (kKind == EK_DMASK) ? mask : 0xffffffffu;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const float v0 = (m & (0x80000000u >> (2 * p))) ? __uint_as_float(r[2 * p]) : 0.f;
            const float v1 = (m & (0x80000000u >> (2 * p + 1))) ? __uint_as_float(r[2 * p + 1]) : 0.f;
            w[p] = ptx::pack_bf16x2(v0, v1);
        }
    }
}

// TS-mode kernel (mlp_tc3.cu): 16 accumulator columns (columns [16 kHalf, 16 kHalf + 16) of a 32-column group) -> 8 packed
// words; a group is converted in two passes to keep 16 instead of 32 accumulator registers live.
// ReLU masks of the TS kernels are built by the saver warps from the bf16 activations with three instructions per packed
// word (ts_mask_word below), which fixes the bit layout: feature 2i of the group -> bit 15 - i, feature 2i + 1 -> bit 31 - i.
__device__ __forceinline__ uint32_t ts_mask_bit(int feature) { return (feature & 1) ? (0x80000000u >> (feature >> 1)) : (0x8000u >> (feature >> 1)); }
// accumulate the mask bits of packed word i (two non-negative bf16: nonzero <=> bit 15 of x + 0x7fff) into m
__device__ __forceinline__ uint32_t ts_mask_word(uint32_t m, uint32_t w, int i) { return m | (((w + 0x7fff7fffu) & 0x80008000u) >> i); }
template <uint8_t kKind, int kHalf>
__device__ __forceinline__ void epi_half(const float *cbias, const uint32_t (&r)[16], int bidx, uint32_t mask, uint32_t *w) {
    if (kKind == EK_RELU || kKind == EK_LINEAR) {
#pragma unroll
        for (int j2 = 0; j2 < 8; ++j2) {
#ifdef NERF_TC3_NOBIAS   // timing experiment (wrong results): what the bias add costs
            float v0 = __uint_as_float(r[2 * j2]), v1 = __uint_as_float(r[2 * j2 + 1]);
#else
            const float4 b4 = reinterpret_cast<const float4 *>(cbias)[(bidx >> 2) + 4 * kHalf + (j2 >> 1)];
            const float2 b = (j2 & 1) ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y);
            unsigned long long acc2, bias2, sum2;
            asm("mov.b64 %0, {%1, %2};" : "=l"(acc2) : "r"(r[2 * j2]), "r"(r[2 * j2 + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(bias2) : "f"(b.x), "f"(b.y));
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sum2) : "l"(acc2), "l"(bias2));
            float v0, v1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(sum2));
#endif
            w[j2] = (kKind == EK_RELU) ? ptx::pack_bf16x2_relu(v0, v1) : ptx::pack_bf16x2(v0, v1);
        }
    } else if (kKind == EK_DMASK) {
        // Mask select without predicates (32 predicate-setting LOP3 + FSEL per group serialise on the 7 predicate registers):
        // shifted left by the pair index P, the mask has feature 2P's bit at bit 15 and feature 2P+1's at bit 31 -- the sign
        // bits of bytes 1 and 3 -- and PRMT's sign-replicate mode turns them into a 0xffff / 0x0000 mask per bf16 half.
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            uint32_t sel;
            asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(sel) : "r"(mask << (8 * kHalf + p)));
            w[p] = ptx::pack_bf16x2(__uint_as_float(r[2 * p]), __uint_as_float(r[2 * p + 1])) & sel;
        }
    } else {
#pragma unroll
        for (int p = 0; p < 8; ++p) w[p] = ptx::pack_bf16x2(__uint_as_float(r[2 * p]), __uint_as_float(r[2 * p + 1]));
    }
}

}  // namespace chain
