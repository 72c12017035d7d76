// blend_tc.cuh — tiling constants of the tcgen05 blend-shape contraction, shared between the
// GEMM kernel (blend_tc.cu) and the pose stage that writes its A operand (mano_pose.cu).
#pragma once
#include "common.cuh"

namespace mb {

constexpr int TC_M = 128;                 // hands per tile (TMEM lanes)
constexpr int TC_N = 160;                 // vertex coordinates per tile
constexpr int TC_N_TILES = 15;            // 15 * 160 = 2400 >= 2334
constexpr int TC_K = 160;                 // padded feature length
constexpr int TC_K_REAL = 145;            // beta(10) + pose feature(135); v_template is added in the epilogue
constexpr int TC_K_CHUNK = 32;            // K elements per pipeline stage
constexpr int TC_K_CHUNKS = TC_K / TC_K_CHUNK;
// hand tiles from which the hand-tile-resident forward kernel is used — the measured crossover (blend_fwd in us, resident
// hand tile vs resident basis): 4 096 hands 43.5 / 29.2, 6 144 43.6 / 37.4, 8 192 45.1 / 45.6, 12 288 45.5 / 59.9,
// 18 944 48.0 / 86.5, 65 536 164 / 262
constexpr int TC_MRES_MIN_TILES = 64;
constexpr int TC_FEAT_SCALE_LOG2 = 4;     // features are pre-scaled by 2^4 before the fp16 split

// UMMA canonical K-major no-swizzle blocks: [row-group][k-group (4 per chunk)][8 rows][8 halves]
constexpr uint32_t TC_LBO = 128;                                        // between K core matrices
constexpr uint32_t TC_SBO = (TC_K_CHUNK / 8) * 128;                     // between 8-row groups (512 B)
constexpr int TC_A_BLOCK_BYTES = TC_M * TC_K_CHUNK * 2;                 // 8 KB: one split of one chunk
constexpr int TC_A_STAGE_BYTES = 2 * TC_A_BLOCK_BYTES;                  // hi + lo
constexpr int TC_A_TILE_BYTES = TC_K_CHUNKS * TC_A_STAGE_BYTES;         // 80 KB per 128 hands
constexpr int TC_B_BLOCK_BYTES = TC_N * TC_K_CHUNK * 2;                 // 10 KB
constexpr int TC_B_TILE_BYTES = TC_K_CHUNKS * 2 * TC_B_BLOCK_BYTES;     // 100 KB per n-tile

struct TcBlobHeader {
    int32_t basis_scale_log2;
    int32_t feat_scale_log2;
    int32_t pad[2];
};

// byte offset of the 16-byte group holding features [8*kg8, 8*kg8+8) of hand h, split sp (0 hi, 1 lo)
__host__ __device__ inline size_t tc_feat_group_offset(long long h, int kg8, int sp) {
    const long long tile = h >> 7;
    const int r = (int)(h & 127);
    const int c = kg8 >> 2, kg = kg8 & 3;
    return (size_t)tile * TC_A_TILE_BYTES + (size_t)c * TC_A_STAGE_BYTES + (size_t)sp * TC_A_BLOCK_BYTES +
           ((size_t)((r >> 3) * (TC_K_CHUNK / 8) + kg) * 8 + (r & 7)) * 16;
}

inline size_t tc_featp_bytes(long long B) { return (size_t)((B + TC_M - 1) / TC_M) * TC_A_TILE_BYTES; }

// ---- backward contraction  dfeat[h][n] = sum_k dv_posed[h][k] * basis[n][k]   (K = 2368 = 74 chunks of 32, block order)
// A operand: dv_posed as bf16 hi + mid tiles written by the skinning backward kernel, per 128-hand
// tile [K chunk 74][split 2][8 KB canonical block]; B operand: basis as bf16 hi + mid, [74][2][10 KB].
constexpr int TCB_K_CHUNKS = (SK_NCOORD + TC_K_CHUNK - 1) / TC_K_CHUNK; // 74 (2352 block-order coordinates + 16 zero columns)
constexpr int TCB_A_CHUNK_BYTES = 2 * TC_A_BLOCK_BYTES;                 // 16 KB: hi + mid of one chunk of one hand tile
constexpr size_t TCB_A_TILE_BYTES = (size_t)TCB_K_CHUNKS * TCB_A_CHUNK_BYTES;   // 1.17 MB per 128 hands
constexpr int TCB_B_CHUNK_BYTES = 2 * TC_B_BLOCK_BYTES;                 // 20 KB
constexpr size_t TCB_B_BYTES = (size_t)TCB_K_CHUNKS * TCB_B_CHUNK_BYTES;

inline size_t tc_dvp_bytes(long long B) { return (size_t)((B + TC_M - 1) / TC_M) * TCB_A_TILE_BYTES; }

size_t blend_tc_blob_bytes();
void blend_tc_pack(const float* basis, const int32_t* coord_map, void* host_blob_tc);

// dfeat: [B][148] rows, or hand-minor [groups][160][32] when hand_minor != 0
// ksplit K ranges, each into its own dfeat copy part_stride floats apart (the pose backward adds them)
int launch_blend_tc_backward(const void* blob, const unsigned char* dvp, float* dfeat, int B, int hand_minor, int ksplit,
                             size_t part_stride, cudaStream_t s);
int launch_blend_tc_forward(const void* blob, const unsigned char* featp, float* v_posed_t, int B, int mode, cudaStream_t s);

}  // namespace mb
