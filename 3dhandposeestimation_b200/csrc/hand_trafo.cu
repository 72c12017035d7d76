// hand_trafo.cu — the keypoint re-parameterisations the reference runs per sample in its dataloaders
// (dataloaderRHD.py:242-250), batched: one thread per hand, forward only.
//   utils/relative_trafo.py:167-216  bone_rel_trafo      xyz[B,21,3] -> (length, angle_x, angle_y)[B,21,3]
//   utils/relative_trafo.py:219-270  bone_rel_trafo_inv  the inverse
//   utils/canonical_trafo.py:93-159  canonical_trafo     xyz -> canonical frame + total rotation [B,3,3]
//   utils/canonical_trafo.py:163-184 flip_right_hand     z -> -z where cond_right
// The reference's 4x4 transforms are affine [R | t]: a bone vector is a difference of two points in the
// parent frame (t cancels) and inverse(T) applied to the origin is parent + length * (third row of R), so only
// the 3x3 rotation is carried; torch.inverse (relative_trafo.py:98) is not needed.
// HBM-bound: 252 B in, 252 (+36) B out per hand; rows move as flat 128-byte warp accesses through a
// pitch-63 shared tile, as in joint_epilogue.cu.
#include "common.cuh"
#include "ptx.cuh"
#include "hand_math.cuh"
#include "../../include/mano_b200.h"

namespace mb {
namespace {

constexpr int HT_WARPS = 4;
constexpr int JN = NOUTJ * 3;


struct R3 { float m[9]; };
__device__ __forceinline__ R3 r3_identity() { return {{1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}}; }
// R <- Rx(-ax) Ry(-ay) R   (relative_trafo.py:118-122 / :89-93 without the translation)
__device__ __forceinline__ R3 step_frame(const R3& R, float ax, float ay) {
    float sx, cx, sy, cy;
    sincosf(ax, &sx, &cx);
    sincosf(ay, &sy, &cy);
    // Ry(-ay) = [[cy, 0, -sy], [0, 1, 0], [sy, 0, cy]] ; Rx(-ax) = [[1, 0, 0], [0, cx, sx], [0, -sx, cx]]
    R3 o;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float a = cy * R.m[j] - sy * R.m[6 + j];        // row 0 of Ry(-ay) R
        const float b = R.m[3 + j];                            // row 1
        const float c = sy * R.m[j] + cy * R.m[6 + j];         // row 2
        o.m[j] = a;
        o.m[3 + j] = cx * b + sx * c;
        o.m[6 + j] = -sx * b + cx * c;
    }
    return o;
}

__global__ void __launch_bounds__(HT_WARPS * 32)
bone_rel_trafo_kernel(const float* __restrict__ xyz, int B, float* __restrict__ rel) {
    __shared__ __align__(16) float tiles[HT_WARPS][32 * JN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tiles[warp];
    float* mine = tile + lane * JN;
    const long long ngroups = ((long long)B + 31) >> 5;
    for (long long g = (long long)blockIdx.x * HT_WARPS + warp; g < ngroups; g += (long long)gridDim.x * HT_WARPS) {
        const long long base = g * 32;
        const int n = (B - base) < 32 ? (int)(B - base) : 32;
        tile_load(tile, xyz, base, n, JN, lane);
        if (lane < n) {
            // joint 0 and the five chains 4f+4 -> 4f+1, every chain starting at the origin with R = I
            auto bone = [&](int b, float dx, float dy, float dz, R3& R) {
                const float len = sqrtf(dx * dx + dy * dy + dz * dz);
                const float ay = atan2f(dx, dz + 1e-8f);
                float s, c;
                sincosf(ay, &s, &c);
                const float tz = s * dx + c * dz;                 // Ry(-ay) delta: x -> 0
                const float ax = atan2f(-dy, tz + 1e-8f);
                R = step_frame(R, ax, ay);
                return make_float3(len, ax, ay);
            };
            float out[JN];
            {
                R3 R = r3_identity();
                const float3 o = bone(0, mine[0], mine[1], mine[2], R);
                out[0] = o.x; out[1] = o.y; out[2] = o.z;
            }
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                R3 R = r3_identity();
                float px = 0.f, py = 0.f, pz = 0.f;                // parent position ('root' = the origin)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int b = 4 * f + 4 - k;
                    const float cx = mine[3 * b], cy = mine[3 * b + 1], cz = mine[3 * b + 2];
                    const float ex = cx - px, ey = cy - py, ez = cz - pz;
                    const float dx = R.m[0] * ex + R.m[1] * ey + R.m[2] * ez;
                    const float dy = R.m[3] * ex + R.m[4] * ey + R.m[5] * ez;
                    const float dz = R.m[6] * ex + R.m[7] * ey + R.m[8] * ez;
                    const float3 o = bone(b, dx, dy, dz, R);
                    out[3 * b] = o.x; out[3 * b + 1] = o.y; out[3 * b + 2] = o.z;
                    px = cx; py = cy; pz = cz;
                }
            }
#pragma unroll
            for (int i = 0; i < JN; ++i) mine[i] = out[i];
        }
        tile_store(tile, rel, base, n, JN, lane);
    }
}

__global__ void __launch_bounds__(HT_WARPS * 32)
bone_rel_trafo_inv_kernel(const float* __restrict__ rel, int B, float* __restrict__ xyz) {
    __shared__ __align__(16) float tiles[HT_WARPS][32 * JN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tiles[warp];
    float* mine = tile + lane * JN;
    const long long ngroups = ((long long)B + 31) >> 5;
    for (long long g = (long long)blockIdx.x * HT_WARPS + warp; g < ngroups; g += (long long)gridDim.x * HT_WARPS) {
        const long long base = g * 32;
        const int n = (B - base) < 32 ? (int)(B - base) : 32;
        tile_load(tile, rel, base, n, JN, lane);
        if (lane < n) {
            float out[JN];
            {
                const R3 R = step_frame(r3_identity(), mine[1], mine[2]);
                out[0] = mine[0] * R.m[6]; out[1] = mine[0] * R.m[7]; out[2] = mine[0] * R.m[8];
            }
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                R3 R = r3_identity();
                float px = 0.f, py = 0.f, pz = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int b = 4 * f + 4 - k;
                    R = step_frame(R, mine[3 * b + 1], mine[3 * b + 2]);
                    const float len = mine[3 * b];
                    px = fmaf(len, R.m[6], px); py = fmaf(len, R.m[7], py); pz = fmaf(len, R.m[8], pz);
                    out[3 * b] = px; out[3 * b + 1] = py; out[3 * b + 2] = pz;
                }
            }
#pragma unroll
            for (int i = 0; i < JN; ++i) mine[i] = out[i];
        }
        tile_store(tile, xyz, base, n, JN, lane);
    }
}

// canonical_trafo.py:23-41: atan(y / (x + 1e-8)) moved to (-pi, pi] by quadrant
__device__ __forceinline__ float atan2_reference(float y, float x) {
    const float PI = 3.141592653589793f;
    const float xe = x + 1e-8f;
    float t = atanf(y / xe);
    t += xe < 0.f ? PI : 0.f;
    t += t < 0.f ? 2.f * PI : 0.f;
    t += t > PI ? -2.f * PI : 0.f;
    return t;
}

__global__ void __launch_bounds__(HT_WARPS * 32)
canonical_trafo_kernel(const float* __restrict__ xyz, const unsigned char* __restrict__ cond_right, int B,
                       float* __restrict__ can, float* __restrict__ rot) {
    __shared__ __align__(16) float tiles[HT_WARPS][32 * JN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = tiles[warp];
    float* mine = tile + lane * JN;
    const long long ngroups = ((long long)B + 31) >> 5;
    for (long long g = (long long)blockIdx.x * HT_WARPS + warp; g < ngroups; g += (long long)gridDim.x * HT_WARPS) {
        const long long base = g * 32;
        const int n = (B - base) < 32 ? (int)(B - base) : 32;
        tile_load(tile, xyz, base, n, JN, lane);
        float tot[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (lane < n) {
            const float PI = 3.141592653589793f;
            const float ox = mine[0], oy = mine[1], oz = mine[2];
            auto at = [&](int j, float& x, float& y, float& z) { x = mine[3 * j] - ox; y = mine[3 * j + 1] - oy; z = mine[3 * j + 2] - oz; };
            float x, y, z;
            // 1. about z: joint 12 into the yz-plane
            at(12, x, y, z);
            float sa, ca;
            sincosf(atan2_reference(x, y), &sa, &ca);
            const float y1 = sa * x + ca * y;                       // joint 12 after Rz: (ca x - sa y, y1, z)
            // 2. about x: joint 12 onto the y axis
            float sb, cb;
            sincosf(-atan2_reference(z, y1) + PI, &sb, &cb);
            // 3. about y: joint 20 into the xy-plane with x > 0
            at(20, x, y, z);
            const float x20 = ca * x - sa * y, y20 = sa * x + ca * y;
            const float z20 = sb * y20 + cb * z;
            float sg, cg;
            sincosf(atan2_reference(z20, x20), &sg, &cg);
            // total = Rz Rx Ry (canonical_trafo.py:134,141,148); a point goes through Ry Rx Rz
            const float Rz[9] = {ca, -sa, 0.f, sa, ca, 0.f, 0.f, 0.f, 1.f};
            const float Rx[9] = {1.f, 0.f, 0.f, 0.f, cb, -sb, 0.f, sb, cb};
            const float Ry[9] = {cg, 0.f, sg, 0.f, 1.f, 0.f, -sg, 0.f, cg};
            float t1[9];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) t1[3 * i + j] = Rz[3 * i] * Rx[j] + Rz[3 * i + 1] * Rx[3 + j] + Rz[3 * i + 2] * Rx[6 + j];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) tot[3 * i + j] = t1[3 * i] * Ry[j] + t1[3 * i + 1] * Ry[3 + j] + t1[3 * i + 2] * Ry[6 + j];
            const float zs = (cond_right != nullptr && cond_right[base + lane]) ? -1.f : 1.f;   // flip_right_hand
#pragma unroll
            for (int j = 0; j < NOUTJ; ++j) {
                at(j, x, y, z);
                const float xa = ca * x - sa * y, ya = sa * x + ca * y;          // Rz
                const float yb = cb * ya - sb * z, zb = sb * ya + cb * z;        // Rx
                mine[3 * j] = cg * xa + sg * zb;                                  // Ry
                mine[3 * j + 1] = yb;
                mine[3 * j + 2] = zs * (-sg * xa + cg * zb);
            }
        }
        tile_store(tile, can, base, n, JN, lane);
        if (rot != nullptr) {
            if (lane < n) {
#pragma unroll
                for (int i = 0; i < 9; ++i) tile[lane * 9 + i] = tot[i];
            }
            tile_store(tile, rot, base, n, 9, lane);
        }
    }
}

__global__ void flip_right_hand_kernel(const float* __restrict__ xyz, const unsigned char* __restrict__ cond, long long total,
                                       int cond_per_joint, int N, int axis, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long joint = i / 3;
        const bool flip = (i - joint * 3 == axis) && cond[cond_per_joint ? joint : joint / N];
        out[i] = flip ? -xyz[i] : xyz[i];
    }
}

inline int ht_grid(int B) {
    const long long nblk = (((long long)B + 31) / 32 + HT_WARPS - 1) / HT_WARPS;
    return (int)(nblk < NUM_SMS * 16 ? nblk : NUM_SMS * 16);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_bone_rel_trafo(const float* coords_xyz, int B, float* coords_rel, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!coords_xyz || !coords_rel) return MB_E_NULL;
    bone_rel_trafo_kernel<<<ht_grid(B), HT_WARPS * 32, 0, (cudaStream_t)stream>>>(coords_xyz, B, coords_rel);
    return cuda_rc();
}

extern "C" int mb_bone_rel_trafo_inv(const float* coords_rel, int B, float* coords_xyz, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!coords_xyz || !coords_rel) return MB_E_NULL;
    bone_rel_trafo_inv_kernel<<<ht_grid(B), HT_WARPS * 32, 0, (cudaStream_t)stream>>>(coords_rel, B, coords_xyz);
    return cuda_rc();
}

extern "C" int mb_canonical_trafo(const float* coords_xyz, const unsigned char* cond_right, int B, float* coords_can,
                                  float* total_rot_mat, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!coords_xyz || !coords_can) return MB_E_NULL;
    canonical_trafo_kernel<<<ht_grid(B), HT_WARPS * 32, 0, (cudaStream_t)stream>>>(coords_xyz, cond_right, B, coords_can, total_rot_mat);
    return cuda_rc();
}

extern "C" int mb_flip_right_hand(const float* coords_xyz, const unsigned char* cond_right, int B, int N, int cond_per_joint,
                                  float* out, mb_stream_t stream) {
    return mb_mirror_hand(coords_xyz, cond_right, B, N, cond_per_joint, 2, out, stream);
}

extern "C" int mb_mirror_hand(const float* coords_xyz, const unsigned char* cond_right, int B, int N, int cond_per_joint, int axis,
                              float* out, mb_stream_t stream) {
    if (B < 0 || N < 0 || axis < 0 || axis > 2) return MB_E_RANGE;
    if (B == 0 || N == 0) return 0;
    if (!coords_xyz || !cond_right || !out) return MB_E_NULL;
    const long long total = (long long)B * N * 3;
    const long long blocks = (total + 255) / 256;
    flip_right_hand_kernel<<<(unsigned)(blocks < NUM_SMS * 16 ? blocks : NUM_SMS * 16), 256, 0, (cudaStream_t)stream>>>(
        coords_xyz, cond_right, total, cond_per_joint != 0, N, axis, out);
    return cuda_rc();
}
