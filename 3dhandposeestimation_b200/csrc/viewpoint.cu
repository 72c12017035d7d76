// viewpoint.cu — the viewpoint epilogue of the canonical-pose heads, forward and backward, one pass each:
//   utils/general.py:191-226            _get_rot_mat: (ux, uy, uz) -> R, theta = sqrt(|u|^2 + 1e-8), axis u / theta
//   network/Hand3DPoseNet.py:41-43      coord_xyz_rel_normed = can_xyz_kps21 @ R          (row vectors times R)
//   network/Hand3DPoseNet.py:46-50      inference: * index_root_bone_length + root, batch_project_xyz_to_uv
// (the same lines in network/Hand3DPosePriorNetwork.py:38-40 and the ground-truth side trainval_hand3DPose.py:389).
// One thread per hand; the 21 x 3 rows go through a pitch-63 shared tile as flat 128-byte warp accesses.
#include "common.cuh"
#include "ptx.cuh"
#include "fk_math.cuh"
#include "../../include/mano_b200.h"

namespace mb {
namespace {

constexpr int VP_WARPS = 4;
constexpr int JN = NOUTJ * 3;


struct AxisAngle { float nx, ny, nz, th, st, ct; };
__device__ __forceinline__ AxisAngle axis_angle(float ux, float uy, float uz) {
    AxisAngle a;
    a.th = sqrtf(ux * ux + uy * uy + uz * uz + 1e-8f);
    sincosf(a.th, &a.st, &a.ct);
    const float inv = 1.0f / a.th;                             // the reference multiplies by 1 / u_norm
    a.nx = ux * inv; a.ny = uy * inv; a.nz = uz * inv;
    return a;
}
__device__ __forceinline__ void rot_from(const AxisAngle& a, float (&R)[9]) {
    const float oc = 1.0f - a.ct;
    R[0] = a.ct + a.nx * a.nx * oc;        R[1] = a.nx * a.ny * oc - a.nz * a.st; R[2] = a.nx * a.nz * oc + a.ny * a.st;
    R[3] = a.ny * a.nx * oc + a.nz * a.st; R[4] = a.ct + a.ny * a.ny * oc;        R[5] = a.ny * a.nz * oc - a.nx * a.st;
    R[6] = a.nz * a.nx * oc - a.ny * a.st; R[7] = a.nz * a.ny * oc + a.nx * a.st; R[8] = a.ct + a.nz * a.nz * oc;
}

__global__ void __launch_bounds__(VP_WARPS * 32)
viewpoint_forward_kernel(const float* __restrict__ can, const float* __restrict__ ux, const float* __restrict__ uy,
                         const float* __restrict__ uz, const float* __restrict__ scale, const float* __restrict__ root,
                         const float* __restrict__ K, int B, float* __restrict__ rot, float* __restrict__ rel,
                         float* __restrict__ xyz, float* __restrict__ uv) {
    __shared__ __align__(16) float tiles[VP_WARPS][32 * JN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tiles[warp];
    float* mine = tile + lane * JN;
    const long long ngroups = ((long long)B + 31) >> 5;
    for (long long g = (long long)blockIdx.x * VP_WARPS + warp; g < ngroups; g += (long long)gridDim.x * VP_WARPS) {
        const long long base = g * 32;
        const int n = (B - base) < 32 ? (int)(B - base) : 32;
        const bool on = lane < n;
        const long long h = base + lane;
        float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
        if (on) rot_from(axis_angle(ux[h], uy[h], uz[h]), R);
        if (can != nullptr) {
            tile_load(tile, can, base, n, JN, lane);
            float p[JN];
            if (on) {
#pragma unroll
                for (int j = 0; j < NOUTJ; ++j) {
                    const float x = mine[3 * j], y = mine[3 * j + 1], z = mine[3 * j + 2];
                    p[3 * j]     = x * R[0] + y * R[3] + z * R[6];
                    p[3 * j + 1] = x * R[1] + y * R[4] + z * R[7];
                    p[3 * j + 2] = x * R[2] + y * R[5] + z * R[8];
                }
            }
            __syncwarp();
            if (rel != nullptr) {
                if (on) {
#pragma unroll
                    for (int i = 0; i < JN; ++i) mine[i] = p[i];
                }
                tile_store(tile, rel, base, n, JN, lane);
            }
            if (xyz != nullptr) {
                if (on) {
                    const float L = scale[h];
                    const float r0 = root[h * 3], r1 = root[h * 3 + 1], r2 = root[h * 3 + 2];
#pragma unroll
                    for (int j = 0; j < NOUTJ; ++j) {
                        p[3 * j] = fmaf(p[3 * j], L, r0); p[3 * j + 1] = fmaf(p[3 * j + 1], L, r1); p[3 * j + 2] = fmaf(p[3 * j + 2], L, r2);
                    }
#pragma unroll
                    for (int i = 0; i < JN; ++i) mine[i] = p[i];
                }
                tile_store(tile, xyz, base, n, JN, lane);
                if (uv != nullptr) {
                    if (on) {
                        float k[9];
#pragma unroll
                        for (int i = 0; i < 9; ++i) k[i] = K[h * 9 + i];
#pragma unroll
                        for (int j = 0; j < NOUTJ; ++j) {
                            float u, v;
                            project_point(k, p[3 * j], p[3 * j + 1], p[3 * j + 2], u, v);
                            tile[lane * 42 + 2 * j] = u; tile[lane * 42 + 2 * j + 1] = v;
                        }
                    }
                    tile_store(tile, uv, base, n, 42, lane);
                }
            }
        }
        if (rot != nullptr) {
            if (on) {
#pragma unroll
                for (int i = 0; i < 9; ++i) tile[lane * 9 + i] = R[i];
            }
            tile_store(tile, rot, base, n, 9, lane);
        }
    }
}

// gradients of (rot_mat, coord_xyz_rel_normed) w.r.t. (can_xyz_kps21, ux, uy, uz)
__global__ void __launch_bounds__(VP_WARPS * 32)
viewpoint_backward_kernel(const float* __restrict__ can, const float* __restrict__ ux, const float* __restrict__ uy,
                          const float* __restrict__ uz, const float* __restrict__ g_rot, const float* __restrict__ g_rel, int B,
                          float* __restrict__ g_can, float* __restrict__ g_ux, float* __restrict__ g_uy, float* __restrict__ g_uz) {
    __shared__ __align__(16) float tiles[VP_WARPS][32 * JN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tiles[warp];
    float* mine = tile + lane * JN;
    const long long ngroups = ((long long)B + 31) >> 5;
    for (long long g = (long long)blockIdx.x * VP_WARPS + warp; g < ngroups; g += (long long)gridDim.x * VP_WARPS) {
        const long long base = g * 32;
        const int n = (B - base) < 32 ? (int)(B - base) : 32;
        const bool on = lane < n;
        const long long h = base + lane;
        AxisAngle a = axis_angle(on ? ux[h] : 0.f, on ? uy[h] : 0.f, on ? uz[h] : 0.f);
        float R[9];
        rot_from(a, R);
        float G[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // dL/dR
        if (g_rot != nullptr) {
            tile_load(tile, g_rot, base, n, 9, lane);
            if (on) {
#pragma unroll
                for (int i = 0; i < 9; ++i) G[i] = tile[lane * 9 + i];
            }
            __syncwarp();
        }
        if (g_rel != nullptr && can != nullptr) {
            float c[JN];
            tile_load(tile, can, base, n, JN, lane);
#pragma unroll
            for (int i = 0; i < JN; ++i) c[i] = on ? mine[i] : 0.f;
            __syncwarp();
            tile_load(tile, g_rel, base, n, JN, lane);
            if (on) {
#pragma unroll
                for (int j = 0; j < NOUTJ; ++j) {
                    const float gx = mine[3 * j], gy = mine[3 * j + 1], gz = mine[3 * j + 2];
                    // rel_j = sum_i can_i R[i][j]
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        G[3 * i] = fmaf(c[3 * j + i], gx, G[3 * i]);
                        G[3 * i + 1] = fmaf(c[3 * j + i], gy, G[3 * i + 1]);
                        G[3 * i + 2] = fmaf(c[3 * j + i], gz, G[3 * i + 2]);
                    }
                    if (g_can != nullptr) {
                        mine[3 * j]     = gx * R[0] + gy * R[1] + gz * R[2];
                        mine[3 * j + 1] = gx * R[3] + gy * R[4] + gz * R[5];
                        mine[3 * j + 2] = gx * R[6] + gy * R[7] + gz * R[8];
                    }
                }
            }
            if (g_can != nullptr) tile_store(tile, g_can, base, n, JN, lane);
            else __syncwarp();
        } else if (g_can != nullptr) {
            for (int i = lane; i < n * JN; i += 32) g_can[base * JN + i] = 0.f;
        }
        if (on && g_ux != nullptr) {
            // R = ct I + (1 - ct) n n^T + st [n]x with n = u / theta, theta = sqrt(|u|^2 + 1e-8) (n is not exactly unit)
            const float oc = 1.f - a.ct;
            const float nx = a.nx, ny = a.ny, nz = a.nz;
            const float tr = G[0] + G[4] + G[8];
            const float nGn = nx * (G[0] * nx + G[1] * ny + G[2] * nz) + ny * (G[3] * nx + G[4] * ny + G[5] * nz) +
                              nz * (G[6] * nx + G[7] * ny + G[8] * nz);
            const float ax0 = G[7] - G[5], ax1 = G[2] - G[6], ax2 = G[3] - G[1];      // <G, d[n]x / dn>
            const float d_st = nx * ax0 + ny * ax1 + nz * ax2;
            const float d_ct = tr - nGn;
            const float d_th = d_st * a.ct - d_ct * a.st;
            const float dn0 = oc * ((G[0] + G[0]) * nx + (G[1] + G[3]) * ny + (G[2] + G[6]) * nz) + a.st * ax0;
            const float dn1 = oc * ((G[3] + G[1]) * nx + (G[4] + G[4]) * ny + (G[5] + G[7]) * nz) + a.st * ax1;
            const float dn2 = oc * ((G[6] + G[2]) * nx + (G[7] + G[5]) * ny + (G[8] + G[8]) * nz) + a.st * ax2;
            // n = u / theta: du = dn / theta + (d_th - (dn . n) / theta) * dtheta/du, dtheta/du = u / theta = n
            const float inv = 1.f / a.th;
            const float k = d_th - (dn0 * nx + dn1 * ny + dn2 * nz) * inv;
            g_ux[h] = fmaf(k, nx, dn0 * inv);
            g_uy[h] = fmaf(k, ny, dn1 * inv);
            g_uz[h] = fmaf(k, nz, dn2 * inv);
        }
    }
}

inline int vp_grid(int B) {
    const long long nblk = (((long long)B + 31) / 32 + VP_WARPS - 1) / VP_WARPS;
    return (int)(nblk < NUM_SMS * 16 ? nblk : NUM_SMS * 16);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_viewpoint_forward(const float* can_xyz, const float* ux, const float* uy, const float* uz,
                                    const float* index_root_bone_length, const float* kp_coord_xyz_root, const float* K, int B,
                                    float* rot_mat, float* rel_normed, float* xyz, float* uv, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!ux || !uy || !uz) return MB_E_NULL;
    if ((rel_normed || xyz || uv) && !can_xyz) return MB_E_NULL;
    if (xyz && (!index_root_bone_length || !kp_coord_xyz_root)) return MB_E_NULL;
    if (uv && (!xyz || !K)) return MB_E_NULL;
    viewpoint_forward_kernel<<<vp_grid(B), VP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        can_xyz, ux, uy, uz, index_root_bone_length, kp_coord_xyz_root, K, B, rot_mat, rel_normed, xyz, uv);
    return cuda_rc();
}

extern "C" int mb_viewpoint_backward(const float* can_xyz, const float* ux, const float* uy, const float* uz, const float* g_rot,
                                     const float* g_rel, int B, float* g_can, float* g_ux, float* g_uy, float* g_uz,
                                     mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!ux || !uy || !uz) return MB_E_NULL;
    if (g_rel && !can_xyz) return MB_E_NULL;
    if ((g_ux || g_uy || g_uz) && !(g_ux && g_uy && g_uz)) return MB_E_NULL;
    viewpoint_backward_kernel<<<vp_grid(B), VP_WARPS * 32, 0, (cudaStream_t)stream>>>(can_xyz, ux, uy, uz, g_rot, g_rel, B, g_can,
                                                                                      g_ux, g_uy, g_uz);
    return cuda_rc();
}
