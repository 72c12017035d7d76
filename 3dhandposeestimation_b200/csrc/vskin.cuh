// vskin.cuh — tiling constants of the fused blend-shape + skinning forward kernel with LANE = VERTEX (vskin.cu),
// and the layout of its section of the constant blob.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "blend_tc.cuh"

namespace mb {

constexpr int VS_M = 128;                          // vertices per tile = TMEM lanes = MMA M
constexpr int VS_NT = (NV + VS_M - 1) / VS_M;      // 7 vertex tiles (the last one holds 10 vertices)
constexpr int VS_NH = 64;                          // hands per hand tile = MMA N of the blend products
constexpr int VS_HC = 8;                           // hands per transform chunk
constexpr int VS_NCH = VS_NH / VS_HC;              // 8 chunks per hand tile
constexpr int VS_TN = VS_HC * BONE_F;              // 96 = MMA N of the transform products: (hand, 3x4 element)
constexpr int VS_A_STAGE_BYTES = 2 * VS_M * TC_K_CHUNK * 2;      // 16 KB: hi + lo of one (tile, plane, K chunk)
constexpr int VS_STAGES_PER_TILE = 3 * TC_K_CHUNKS;              // 15: (plane, K chunk)
constexpr int VS_W_SPLITS = 2;
constexpr int VS_W_TILE_BYTES = VS_W_SPLITS * VS_M * NJ * 2;     // 8 KB: two fp16 splits of W[128 vertices][16 bones]
constexpr int VS_BONE_SPLITS = 3;
constexpr int VS_BONE_CHUNK_BYTES = VS_TN * NJ * 2;              // 3072 B: one split of one chunk, MN-major [12 n-groups][2 k-groups][8 k][8 n]
constexpr int VS_BONE_TILE_BYTES = VS_NCH * VS_BONE_SPLITS * VS_BONE_CHUNK_BYTES;   // 73 728 B per 64 hands
constexpr int VS_W_SCALE_LOG2 = 8;                 // skinning weights (<= 1) are pre-scaled by 2^8 before the fp16 split
constexpr int VS_BONE_SCALE_LOG2 = 4;              // bone transforms (|R| <= 1, |t| < ~1 m) by 2^4: fp16 overflow only beyond 4 km
constexpr int VS_MIN_HANDS = 8192;                 // the fused forward is used from here on (one-thread-per-hand pose kernels)

// extra constant-blob section behind the blend_tc images
struct VsBlobLayout {
    size_t basis;       // fp16 hi/lo A-operand images of the blend basis: [7 tiles][3 planes][5 K chunks][2][8 KB], K-major
    size_t w;           // fp16 x2 rows of the dense skinning weights: [896 vertices][2 splits][16 bones] (the kernel stores them into TMEM)
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

size_t vskin_blob_bytes();
// basis [FEAT_K][2334], skin_w / skin_b [778][8] ELL, sk_perm from the packed blob (block order) -> the section above
void vskin_pack(const float* basis, const float* skin_w, const int32_t* skin_b, const int32_t* sk_perm, int basis_scale_log2,
                void* host_section);
// verts[B][778][3], fingertip joints; v_posed_t (nullable): the rest-pose scratch of the skinning backward, hand-minor block order
// bone_t: the pose stage's fp32 transforms [groups][16][32][12]; bones_op: scratch of vskin_bones_op_bytes(B) for their fp16 x3
// operand images (written by a conversion kernel launched first on the same stream)
inline size_t vskin_bones_op_bytes(long long B) { return (size_t)((B + VS_NH - 1) / VS_NH) * VS_BONE_TILE_BYTES; }
int launch_vskin_forward(const void* blob, const unsigned char* featp, const float* bone_t, unsigned char* bones_op, int B, int mode,
                         float* verts, float* joints, float* v_posed_t, float* dbg, int variant, cudaStream_t s);

}  // namespace mb
