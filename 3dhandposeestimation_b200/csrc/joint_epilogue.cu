// joint_epilogue.cu — the heads' joint epilogue in one pass: match_mano_to_RHD -> pinhole projection,
// forward and backward.
//   reference: network/Resnet50MANO3DHandPose.py:35-60 (same body network/MANO3DHandPose.py:30-55) and
//   utils/coordinate_trans.py:48-65 as called at Resnet50MANO3DHandPose.py:73
//
//   p_i   = joints[slot(i)]                      slot = per-finger reversal when joint_order_switched is False
//   r_i   = p_i - p_0 ; s = ||r_12|| ; n_i = r_i / s           -> rel_normalized
//   x_i   = n_i * index_root_bone_length + kp_coord_xyz_root    -> joint_xyz21
//   uv_i  = project(K, x_i)                                     -> uv21 (optional)
//
// HBM-bound and tiny (252 B in, <= 672 B out per hand): one thread per hand, every global row moved as
// contiguous 128-byte warp accesses through a per-warp shared-memory tile (pitch 63 = -1 mod 32 banks).
#include "common.cuh"
#include "ptx.cuh"
#include "fk_math.cuh"
#include "../../include/mano_b200.h"

namespace mb {
namespace {

constexpr int JE_WARPS = 4;
constexpr int JN = NOUTJ * 3;                      // 63 floats per hand

// joints[base .. base + n) rows <-> tile (row pitch `w`, contiguous, so the copy is flat)
__device__ __forceinline__ int je_slot(int i, int swap) { return i == 0 ? 0 : fk_out_slot(i, swap); }

__global__ void __launch_bounds__(JE_WARPS * 32)
joint_epilogue_forward_kernel(const float* __restrict__ joints, const float* __restrict__ scale, const float* __restrict__ root,
                              const float* __restrict__ K, int B, int swap, float* __restrict__ rel, float* __restrict__ xyz,
                              float* __restrict__ uv) {
    __shared__ __align__(16) float tiles[JE_WARPS][32 * JN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tiles[warp];
    float* mine = tile + lane * JN;
    const long long ngroups = ((long long)B + 31) >> 5;
    for (long long g = (long long)blockIdx.x * JE_WARPS + warp; g < ngroups; g += (long long)gridDim.x * JE_WARPS) {
        const long long base = g * 32;
        const int n = (B - base) < 32 ? (int)(B - base) : 32;
        const bool on = lane < n;
        tile_load(tile, joints, base, n, JN, lane);
        float p[JN];
#pragma unroll
        for (int i = 0; i < NOUTJ; ++i) {
            const int s = je_slot(i, swap);
#pragma unroll
            for (int c = 0; c < 3; ++c) p[3 * i + c] = on ? mine[3 * s + c] : 0.f;
        }
        __syncwarp();
        const float p0x = p[0], p0y = p[1], p0z = p[2];
#pragma unroll
        for (int i = 0; i < NOUTJ; ++i) { p[3 * i] -= p0x; p[3 * i + 1] -= p0y; p[3 * i + 2] -= p0z; }
        const float s = sqrtf(p[36] * p[36] + p[37] * p[37] + p[38] * p[38]);
#pragma unroll
        for (int i = 0; i < JN; ++i) p[i] = p[i] / s;          // the reference divides; so does this
        if (rel != nullptr) {
            if (on) {
#pragma unroll
                for (int i = 0; i < JN; ++i) mine[i] = p[i];
            }
            tile_store(tile, rel, base, n, JN, lane);
        }
        const float L = on ? scale[base + lane] : 0.f;
        float rt[3] = {0.f, 0.f, 0.f};
        if (on) { rt[0] = root[(base + lane) * 3]; rt[1] = root[(base + lane) * 3 + 1]; rt[2] = root[(base + lane) * 3 + 2]; }
#pragma unroll
        for (int i = 0; i < JN; ++i) p[i] = fmaf(p[i], L, rt[i % 3]);
        if (on) {
#pragma unroll
            for (int i = 0; i < JN; ++i) mine[i] = p[i];
        }
        tile_store(tile, xyz, base, n, JN, lane);
        if (uv != nullptr) {
            if (on) {
                float k[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) k[i] = K[(base + lane) * 9 + i];
#pragma unroll
                for (int i = 0; i < NOUTJ; ++i) {
                    float u, v;
                    project_point(k, p[3 * i], p[3 * i + 1], p[3 * i + 2], u, v);
                    tile[lane * 42 + 2 * i] = u; tile[lane * 42 + 2 * i + 1] = v;
                }
            }
            tile_store(tile, uv, base, n, 42, lane);
        }
    }
}

__global__ void __launch_bounds__(JE_WARPS * 32)
joint_epilogue_backward_kernel(const float* __restrict__ joints, const float* __restrict__ scale, const float* __restrict__ root,
                               const float* __restrict__ K, const float* __restrict__ g_rel, const float* __restrict__ g_xyz,
                               const float* __restrict__ g_uv, int B, int swap, float* __restrict__ g_joints,
                               float* __restrict__ g_scale, float* __restrict__ g_root) {
    __shared__ __align__(16) float tiles[JE_WARPS][32 * JN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tiles[warp];
    float* mine = tile + lane * JN;
    const long long ngroups = ((long long)B + 31) >> 5;
    for (long long g = (long long)blockIdx.x * JE_WARPS + warp; g < ngroups; g += (long long)gridDim.x * JE_WARPS) {
        const long long base = g * 32;
        const int n = (B - base) < 32 ? (int)(B - base) : 32;
        const bool on = lane < n;
        tile_load(tile, joints, base, n, JN, lane);
        float r[JN], gx[JN];                                   // r_i = p_i - p_0 ; gx: d/dx_i, later d/dr_i
#pragma unroll
        for (int i = 0; i < NOUTJ; ++i) {
            const int s = je_slot(i, swap);
#pragma unroll
            for (int c = 0; c < 3; ++c) r[3 * i + c] = on ? mine[3 * s + c] : (i == 12 ? 1.f : 0.f);
        }
        __syncwarp();
        const float p0x = r[0], p0y = r[1], p0z = r[2];
#pragma unroll
        for (int i = 0; i < NOUTJ; ++i) { r[3 * i] -= p0x; r[3 * i + 1] -= p0y; r[3 * i + 2] -= p0z; }
        const float s = sqrtf(r[36] * r[36] + r[37] * r[37] + r[38] * r[38]);
        const float L = on ? scale[base + lane] : 0.f;
        float rt[3] = {0.f, 0.f, 0.f};
        if (on) { rt[0] = root[(base + lane) * 3]; rt[1] = root[(base + lane) * 3 + 1]; rt[2] = root[(base + lane) * 3 + 2]; }
#pragma unroll
        for (int i = 0; i < JN; ++i) gx[i] = 0.f;
        if (g_uv != nullptr) {
            tile_load(tile, g_uv, base, n, 42, lane);
            if (on) {
                float k[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) k[i] = K[(base + lane) * 9 + i];
#pragma unroll
                for (int i = 0; i < NOUTJ; ++i) {
                    const float x = fmaf(r[3 * i] / s, L, rt[0]), y = fmaf(r[3 * i + 1] / s, L, rt[1]), z = fmaf(r[3 * i + 2] / s, L, rt[2]);
                    const V3 d = project_point_bwd(k, x, y, z, tile[lane * 42 + 2 * i], tile[lane * 42 + 2 * i + 1]);
                    gx[3 * i] = d.x; gx[3 * i + 1] = d.y; gx[3 * i + 2] = d.z;
                }
            }
            __syncwarp();
        }
        if (g_xyz != nullptr) {
            tile_load(tile, g_xyz, base, n, JN, lane);
            if (on) {
#pragma unroll
                for (int i = 0; i < JN; ++i) gx[i] += mine[i];
            }
            __syncwarp();
        }
        // x = n L + root
        float gL = 0.f, gr0 = 0.f, gr1 = 0.f, gr2 = 0.f;
#pragma unroll
        for (int i = 0; i < NOUTJ; ++i) {
            gL += (r[3 * i] * gx[3 * i] + r[3 * i + 1] * gx[3 * i + 1] + r[3 * i + 2] * gx[3 * i + 2]);
            gr0 += gx[3 * i]; gr1 += gx[3 * i + 1]; gr2 += gx[3 * i + 2];
        }
        if (on) {
            if (g_scale != nullptr) g_scale[base + lane] = gL / s;
            if (g_root != nullptr) { g_root[(base + lane) * 3] = gr0; g_root[(base + lane) * 3 + 1] = gr1; g_root[(base + lane) * 3 + 2] = gr2; }
        }
#pragma unroll
        for (int i = 0; i < JN; ++i) gx[i] *= L;               // now d/dn_i
        if (g_rel != nullptr) {
            tile_load(tile, g_rel, base, n, JN, lane);
            if (on) {
#pragma unroll
                for (int i = 0; i < JN; ++i) gx[i] += mine[i];
            }
            __syncwarp();
        }
        // n = r / s, s = ||r_12||
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < JN; ++i) dot = fmaf(gx[i], r[i], dot);
        const float gs = -dot / (s * s);
#pragma unroll
        for (int i = 0; i < JN; ++i) gx[i] = gx[i] / s;
#pragma unroll
        for (int c = 0; c < 3; ++c) gx[36 + c] = fmaf(gs, r[36 + c] / s, gx[36 + c]);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;                     // p_0 enters every r_i with -1
#pragma unroll
        for (int i = 1; i < NOUTJ; ++i) { s0 += gx[3 * i]; s1 += gx[3 * i + 1]; s2 += gx[3 * i + 2]; }
        gx[0] = -s0; gx[1] = -s1; gx[2] = -s2;
        if (on) {
#pragma unroll
            for (int i = 0; i < NOUTJ; ++i) {
                const int sl = je_slot(i, swap);
#pragma unroll
                for (int c = 0; c < 3; ++c) mine[3 * sl + c] = gx[3 * i + c];
            }
        }
        tile_store(tile, g_joints, base, n, JN, lane);
    }
}

inline int je_grid(int B) {
    const long long nblk = (((long long)B + 31) / 32 + JE_WARPS - 1) / JE_WARPS;
    return (int)(nblk < NUM_SMS * 16 ? nblk : NUM_SMS * 16);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_joint_epilogue_forward(const float* joints, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                                         const float* K, int B, int swap_order, float* rel_normalized, float* xyz, float* uv,
                                         mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!joints || !index_root_bone_length || !kp_coord_xyz_root || !xyz || (uv && !K)) return MB_E_NULL;
    joint_epilogue_forward_kernel<<<je_grid(B), JE_WARPS * 32, 0, (cudaStream_t)stream>>>(
        joints, index_root_bone_length, kp_coord_xyz_root, K, B, swap_order != 0, rel_normalized, xyz, uv);
    return cuda_rc();
}

extern "C" int mb_joint_epilogue_backward(const float* joints, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                                          const float* K, const float* g_rel, const float* g_xyz, const float* g_uv, int B,
                                          int swap_order, float* g_joints, float* g_scale, float* g_root, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!joints || !index_root_bone_length || !kp_coord_xyz_root || !g_joints || (g_uv && !K)) return MB_E_NULL;
    joint_epilogue_backward_kernel<<<je_grid(B), JE_WARPS * 32, 0, (cudaStream_t)stream>>>(
        joints, index_root_bone_length, kp_coord_xyz_root, K, g_rel, g_xyz, g_uv, B, swap_order != 0, g_joints, g_scale, g_root);
    return cuda_rc();
}
