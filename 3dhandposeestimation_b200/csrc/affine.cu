// affine.cu — optional per-hand scale and translation of the MANO layer's outputs (the `transl=` / `scale=`
// keyword extension of ManoLayer.forward; BASELINE north_star "global rotation and translation in").
//
// The reference's layer has no such argument (MANOLayer.py:238); its callers apply the equivalent post-ops themselves
// on the layer's outputs: resnet50MANO.py:77-81 (uv = trans + scale * joints[:, :, :2], and the commented
// x3d = scale * x3d; x3d[:, :, :2] += trans) and Resnet50MANO3DHandPose.py:35-60 (* index_root_bone_length +
// kp_coord_xyz_root).  Here:   p' = scale[h] * p + transl[h]   for every vertex and joint of hand h, in place,
// one warp per hand, rows moved as flat coalesced accesses.  Backward: the MANO backward is linear in the upstream
// gradient, so instead of scaling g_verts (a 9.3 KB/hand pass) the parameter gradients are scaled by scale[h];
//   g_transl[h] = sum_p g_p ,  g_scale[h] = sum_p <g_p, (p' - transl[h]) / scale[h]>.
#include "common.cuh"

namespace mb {
namespace {

constexpr int AFF_WARPS = 8;

__global__ void __launch_bounds__(AFF_WARPS * 32)
affine_forward_kernel(float* __restrict__ verts, float* __restrict__ joints, const float* __restrict__ scale,
                      const float* __restrict__ transl, int B) {
    const int lane = threadIdx.x & 31;
    for (long long h = (long long)blockIdx.x * AFF_WARPS + (threadIdx.x >> 5); h < B; h += (long long)gridDim.x * AFF_WARPS) {
        const float s = scale ? scale[h] : 1.f;
        float t[3] = {0.f, 0.f, 0.f};
        if (transl) { t[0] = transl[h * 3]; t[1] = transl[h * 3 + 1]; t[2] = transl[h * 3 + 2]; }
        if (verts) {
            float* v = verts + h * NVC;
            for (int i = lane; i < NVC; i += 32) v[i] = fmaf(s, v[i], t[i % 3]);
        }
        float* j = joints + h * (NOUTJ * 3);
        for (int i = lane; i < NOUTJ * 3; i += 32) j[i] = fmaf(s, j[i], t[i % 3]);
    }
}

__global__ void __launch_bounds__(AFF_WARPS * 32)
affine_backward_kernel(const float* __restrict__ g_verts, const float* __restrict__ g_joints,
                       const float* __restrict__ verts_out, const float* __restrict__ joints_out,
                       const float* __restrict__ scale, const float* __restrict__ transl, int B, int nc,
                       float* __restrict__ g_scale, float* __restrict__ g_transl,
                       float* __restrict__ g_rot, float* __restrict__ g_coeffs, float* __restrict__ g_betas) {
    const int lane = threadIdx.x & 31;
    for (long long h = (long long)blockIdx.x * AFF_WARPS + (threadIdx.x >> 5); h < B; h += (long long)gridDim.x * AFF_WARPS) {
        const float s = scale ? scale[h] : 1.f;
        float t[3] = {0.f, 0.f, 0.f};
        if (transl) { t[0] = transl[h * 3]; t[1] = transl[h * 3 + 1]; t[2] = transl[h * 3 + 2]; }
        // lane-strided sums: lane i handles floats i, i + 32, ... ; 32 = 2 (mod 3), so a lane's coordinate index cycles
        float gs = 0.f, gt[3] = {0.f, 0.f, 0.f};
        if (g_verts) {
            const float* g = g_verts + h * NVC;
            const float* p = verts_out + h * NVC;
            for (int i = lane; i < NVC; i += 32) {
                const float gi = g[i];
                const int c = i % 3;
                gt[c] += gi;
                gs = fmaf(gi, p[i] - t[c], gs);
            }
        }
        if (g_joints) {
            const float* g = g_joints + h * (NOUTJ * 3);
            const float* p = joints_out + h * (NOUTJ * 3);
            for (int i = lane; i < NOUTJ * 3; i += 32) {
                const float gi = g[i];
                const int c = i % 3;
                gt[c] += gi;
                gs = fmaf(gi, p[i] - t[c], gs);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gs += __shfl_xor_sync(0xffffffffu, gs, o);
            gt[0] += __shfl_xor_sync(0xffffffffu, gt[0], o);
            gt[1] += __shfl_xor_sync(0xffffffffu, gt[1], o);
            gt[2] += __shfl_xor_sync(0xffffffffu, gt[2], o);
        }
        if (lane == 0) {
            if (g_scale) g_scale[h] = s != 0.f ? gs / s : 0.f;
            if (g_transl) { g_transl[h * 3] = gt[0]; g_transl[h * 3 + 1] = gt[1]; g_transl[h * 3 + 2] = gt[2]; }
        }
        if (scale) {                                    // d(params) of the unscaled layer x scale[h]
            if (lane < 3) g_rot[h * 3 + lane] *= s;
            for (int i = lane; i < nc; i += 32) g_coeffs[h * nc + i] *= s;
            if (lane < NB) g_betas[h * NB + lane] *= s;
        }
    }
}

inline unsigned aff_grid(int B) {
    long long b = ((long long)B + AFF_WARPS - 1) / AFF_WARPS;
    const long long cap = (long long)NUM_SMS * 8;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_affine_forward(float* verts, float* joints, const float* scale, const float* transl, int B, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0 || (!scale && !transl)) return 0;
    if (!joints) return MB_E_NULL;
    affine_forward_kernel<<<aff_grid(B), AFF_WARPS * 32, 0, (cudaStream_t)stream>>>(verts, joints, scale, transl, B);
    return cuda_rc();
}

extern "C" int mb_affine_backward(const float* g_verts, const float* g_joints, const float* verts_out, const float* joints_out,
                                  const float* scale, const float* transl, int B, int nc, float* g_scale, float* g_transl,
                                  float* g_rot, float* g_coeffs, float* g_betas, mb_stream_t stream) {
    if (B < 0 || nc < 1 || nc > NAA) return MB_E_RANGE;
    if (B == 0 || (!scale && !transl)) return 0;
    if ((g_verts && !verts_out) || (g_joints && !joints_out)) return MB_E_NULL;
    if (scale && (!g_rot || !g_coeffs || !g_betas)) return MB_E_NULL;
    affine_backward_kernel<<<aff_grid(B), AFF_WARPS * 32, 0, (cudaStream_t)stream>>>(g_verts, g_joints, verts_out, joints_out, scale, transl,
                                                                                    B, nc, g_scale, g_transl, g_rot, g_coeffs, g_betas);
    return cuda_rc();
}
