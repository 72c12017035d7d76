// vskin.cuh — tiling constants of the fused blend-shape + skinning forward kernel with LANE = VERTEX (vskin.cu),
// shared with the pose stage that writes its bone operand (mano_pose_lh.cu).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "blend_tc.cuh"

namespace mb {

constexpr int VS_M = 128;                          // vertices per tile = TMEM lanes = MMA M
constexpr int VS_NT = (NV + VS_M - 1) / VS_M;      // 7 vertex tiles (the last one holds 10 vertices)
static_assert(VS_NH == 64, "hands per hand tile = MMA N of the blend products (common.cuh)");
constexpr int VS_HC = 4;                           // hands per transform chunk
constexpr int VS_NCH = VS_NH / VS_HC;              // 16 chunks per hand tile
constexpr int VS_TN = VS_HC * BONE_F;              // 48 = MMA N of the transform products: (hand, 3x4 element)
constexpr int VS_A_STAGE_BYTES = 2 * VS_M * TC_K_CHUNK * 2;      // 16 KB: hi + lo of one (tile, plane, K chunk)
constexpr int VS_STAGES_PER_TILE = 3 * TC_K_CHUNKS;              // 15: (plane, K chunk)
constexpr int VS_W_SPLITS = 3;
constexpr int VS_W_TILE_BYTES = VS_W_SPLITS * VS_M * NJ * 2;     // 12 KB: three fp16 splits of W[128 vertices][16 bones]
constexpr int VS_BONE_SPLITS = 3;
constexpr int VS_BONE_CHUNK_BYTES = VS_TN * NJ * 2;              // 1536 B: one split of one chunk, MN-major [6 n-groups][2 k-groups][8 k][8 n]
static_assert(VS_BONE_TILE_BYTES == VS_NCH * VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES, "73 728 B per 64 hands (common.cuh)");
constexpr int VS_W_SCALE_LOG2 = 8;                 // skinning weights (<= 1) are pre-scaled by 2^8 before the fp16 split
constexpr int VS_BONE_SCALE_LOG2 = 4;              // bone transforms (|R| <= 1, |t| < ~1 m) by 2^4: fp16 overflow only beyond 4 km
constexpr int VS_MIN_HANDS = 8192;                 // the fused forward is used from here on (one-thread-per-hand pose kernels)

// extra constant-blob section behind the blend_tc images
struct VsBlobLayout {
    size_t basis;       // fp16 hi/lo A-operand images of the blend basis: [7 tiles][3 planes][5 K chunks][2][8 KB], K-major
    size_t w;           // fp16 x3 A-operand images of the skinning weights: [7 tiles][3 splits][4 KB], K-major (K = bone)
    size_t tmpl;        // float4 [896]: v_template x, y, z of the vertex; .w = bits of int: 3 * block-order position (v_posed_t row) or -1
    size_t total;
};
__host__ __device__ inline VsBlobLayout vs_blob_layout() {
    VsBlobLayout L;
    size_t o = 0;
    L.basis = o; o = align256(o + (size_t)VS_NT * VS_STAGES_PER_TILE * VS_A_STAGE_BYTES);
    L.w = o;     o = align256(o + (size_t)VS_NT * VS_W_TILE_BYTES);
    L.tmpl = o;  o = align256(o + sizeof(float) * 4 * VS_NT * VS_M);
    L.total = o;
    return L;
}

inline size_t vs_bone16_bytes(long long B) { return (size_t)((B + VS_NH - 1) / VS_NH) * VS_BONE_TILE_BYTES; }

// The pose stage writes bone k of hand h (3x4 transform A[12], global rotation folded in) into the transform
// products' B operand: three fp16 splits a = a1 + a2 + a3 of 2^4 * A, MN-major canonical (no swizzle) core matrices
// [n-group][k-group][8 k][8 n], n = (h % 4) * 12 + element, k = bone.  A hand's 12 elements of one bone are 24
// contiguous bytes that straddle two 16-byte groups: one 16-byte and one 8-byte store per split.
__device__ __forceinline__ void vs_emit_bone16(unsigned char* __restrict__ bone16, long long hand, int k, const float (&A)[BONE_F]) {
    const int hl = (int)(hand & 3);
    unsigned char* base = bone16 + (size_t)(hand >> 6) * VS_BONE_TILE_BYTES + (size_t)((hand & 63) >> 2) * (VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES) +
                          (k >> 3) * 128 + (k & 7) * 16;
    float r[BONE_F];
#pragma unroll
    for (int j = 0; j < BONE_F; ++j) r[j] = A[j] * (float)(1 << VS_BONE_SCALE_LOG2);
#pragma unroll
    for (int s = 0; s < VS_BONE_SPLITS; ++s) {
        uint32_t p[BONE_F / 2];                                   // packed pairs (element 2i in the low half)
#pragma unroll
        for (int i = 0; i < BONE_F / 2; ++i) {
            const __half lo = __float2half_rn(r[2 * i]), hi = __float2half_rn(r[2 * i + 1]);
            r[2 * i] -= __half2float(lo);
            r[2 * i + 1] -= __half2float(hi);
            p[i] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
        }
        unsigned char* d = base + s * VS_BONE_CHUNK_BYTES;
        // n = hl * 12 + j -> byte (n >> 3) * 256 + (n & 7) * 2
        const int g = (hl * BONE_F) >> 3;
        if ((hl & 1) == 0) {            // n0 = 0 or 24: elements 0-7 fill a group, 8-11 the first half of the next
            *reinterpret_cast<uint4*>(d + g * 256) = make_uint4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<uint2*>(d + (g + 1) * 256) = make_uint2(p[4], p[5]);
        } else {                        // n0 = 12 or 36: elements 0-3 are the second half of a group, 4-11 fill the next
            *reinterpret_cast<uint2*>(d + g * 256 + 8) = make_uint2(p[0], p[1]);
            *reinterpret_cast<uint4*>(d + (g + 1) * 256) = make_uint4(p[2], p[3], p[4], p[5]);
        }
    }
}

size_t vskin_blob_bytes();
// basis [FEAT_K][2334], skin_w / skin_b [778][8] ELL, sk_perm from the packed blob (block order) -> the section above
void vskin_pack(const float* basis, const float* skin_w, const int32_t* skin_b, const int32_t* sk_perm, int basis_scale_log2,
                void* host_section);
// verts[B][778][3], fingertip joints; v_posed_t (nullable): the rest-pose scratch of the skinning backward, hand-minor block order
int launch_vskin_forward(const void* blob, const unsigned char* featp, const unsigned char* bone16, int B, int mode,
                         float* verts, float* joints, float* v_posed_t, float* dbg, int variant, cudaStream_t s);

}  // namespace mb
