// head_loss.cu — the MANO heads' tail as ONE call in each direction:
//   MANO parameters -> 21 joints (joints-only kernels, no 778-vertex contraction) [-> scale * p + transl]
//   [-> match_mano_to_RHD] -> pinhole projection -> L2Loss on xyz, L2Loss on uv, MANO regulariser.
// Reference chain: network/sub_modules/resnet50MANO.py:76-87 (mano_layer, scale / translation post-ops),
// network/Resnet50MANO3DHandPose.py:35-60 (match_mano_to_RHD), :71-73 (batch_project_xyz_to_uv),
// criterions/loss.py:10-25 (L2Loss), :83-87 (xyz / uv losses), :113-117 (regulariser), trainval.py:328-358.
//
// The two entry points enqueue the kernels that already implement each piece (and are pinned to the reference
// piece by piece) back to back on the caller's stream: no host round trip, no torch op and no allocation between
// them — at the heads' batch sizes (config.py:79: 200) the path is launch-latency-bound, and one ctypes call that
// can be captured in a CUDA graph replaces ~10 autograd nodes.  All scratch lives in a caller-provided workspace
// that must stay untouched between the forward and its backward.
#include "common.cuh"

using namespace mb;

namespace {

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}
__global__ void zero_kernel(float* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = 0.f;
}
int launch_add(float* dst, const float* src, long long n, cudaStream_t s) {
    const long long b = (n + 255) / 256;
    add_inplace_kernel<<<(unsigned)(b < NUM_SMS * 8 ? b : NUM_SMS * 8), 256, 0, s>>>(dst, src, n);
    return cuda_rc();
}
int launch_zero(float* dst, long long n, cudaStream_t s) {
    const long long b = (n + 255) / 256;
    zero_kernel<<<(unsigned)(b < NUM_SMS * 8 ? b : NUM_SMS * 8), 256, 0, s>>>(dst, n);
    return cuda_rc();
}

struct HeadWs {
    size_t joints;      // float [B][21][3]  MANO joints (after scale / transl) when match_mano_to_RHD follows
    size_t g_xyz;       // float [B][21][3]
    size_t g_uv;        // float [B][21][2]
    size_t g_joints;    // float [B][21][3]
    size_t g_theta;     // float [B][45]
    size_t g_beta;      // float [B][10]
    size_t accum;       // double [6]: {sum, count} of the xyz loss, of the uv loss, {sum theta^2, sum beta^2}
    size_t total;
};
HeadWs head_ws(long long B) {
    HeadWs W;
    size_t o = 0;
    W.joints = o;   o = align256(o + sizeof(float) * B * 63);
    W.g_xyz = o;    o = align256(o + sizeof(float) * B * 63);
    W.g_uv = o;     o = align256(o + sizeof(float) * B * 42);
    W.g_joints = o; o = align256(o + sizeof(float) * B * 63);
    W.g_theta = o;  o = align256(o + sizeof(float) * B * NAA);
    W.g_beta = o;   o = align256(o + sizeof(float) * B * 10);
    W.accum = o;    o = align256(o + sizeof(double) * 6);
    W.total = o;
    return W;
}

}  // namespace

extern "C" size_t mb_mano_head_loss_workspace_bytes(int B) { return B < 0 ? 0 : head_ws(B).total; }

extern "C" int mb_mano_head_loss_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                         const float* transl, const float* scale,
                                         const float* index_root_bone_length, const float* kp_coord_xyz_root, const float* K,
                                         const float* gt_xyz, const float* gt_uv, const float* keypoint_vis,
                                         int B, int mode, int flags, int swap_order, float alpha_beta,
                                         float* joint_xyz21, float* uv21, float* losses,
                                         void* workspace, size_t workspace_bytes, mb_stream_t stream) {
    if (B < 0 || (flags & ~(MB_HEAD_XYZ | MB_HEAD_UV | MB_HEAD_REG | MB_HEAD_MATCH))) return MB_E_RANGE;
    if (!losses) return MB_E_NULL;
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (B == 0) return launch_zero(losses, 3, s);
    if (!blob || !rot || !coeffs || !betas || !K || !joint_xyz21 || !uv21 || !workspace) return MB_E_NULL;
    if (((flags & MB_HEAD_XYZ) && !gt_xyz) || ((flags & MB_HEAD_UV) && !gt_uv) || ((flags & (MB_HEAD_XYZ | MB_HEAD_UV)) && !keypoint_vis))
        return MB_E_NULL;
    if ((flags & MB_HEAD_MATCH) && (!index_root_bone_length || !kp_coord_xyz_root)) return MB_E_NULL;
    if (workspace_bytes < mb_mano_head_loss_workspace_bytes(B)) return MB_E_WORKSPACE;
    const HeadWs W = head_ws(B);
    char* ws = reinterpret_cast<char*>(workspace);
    double* accum = reinterpret_cast<double*>(ws + W.accum);
    const bool match = (flags & MB_HEAD_MATCH) != 0;
    float* joints = match ? reinterpret_cast<float*>(ws + W.joints) : joint_xyz21;
    // 1. joints-only MANO forward (+ the callers' scale / translation)
    if ((rc = mb_mano_forward(blob, nc, rot, coeffs, betas, B, mode, nullptr, joints, nullptr, 0, stream))) return rc;
    if (transl || scale)
        if ((rc = mb_affine_forward(nullptr, joints, scale, transl, B, stream))) return rc;
    // 2. match_mano_to_RHD + projection in one kernel, or the projection alone
    if (match) {
        if ((rc = mb_joint_epilogue_forward(joints, index_root_bone_length, kp_coord_xyz_root, K, B, swap_order, nullptr,
                                            joint_xyz21, uv21, stream))) return rc;
    } else {
        if ((rc = mb_project_uv_forward(joints, K, B, 21, uv21, stream))) return rc;
    }
    // 3. the three loss terms, each a device scalar (terms not asked for are 0)
    if ((rc = launch_zero(losses, 3, s))) return rc;
    if (flags & MB_HEAD_XYZ)
        if ((rc = mb_masked_joint_reduce(joint_xyz21, gt_xyz, keypoint_vis, MB_VIS_F32, (long long)B * 21, 3, MB_REDUCE_L2, accum,
                                         losses, stream))) return rc;
    if (flags & MB_HEAD_UV)
        if ((rc = mb_masked_joint_reduce(uv21, gt_uv, keypoint_vis, MB_VIS_F32, (long long)B * 21, 2, MB_REDUCE_L2, accum + 2,
                                         losses + 1, stream))) return rc;
    if (flags & MB_HEAD_REG)
        if ((rc = mb_regulariser_forward(coeffs, (long long)B * nc, betas, (long long)B * 10, alpha_beta, accum + 4, losses + 2,
                                         stream))) return rc;
    return 0;
}

extern "C" int mb_mano_head_loss_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                          const float* transl, const float* scale,
                                          const float* index_root_bone_length, const float* kp_coord_xyz_root, const float* K,
                                          const float* gt_xyz, const float* gt_uv, const float* keypoint_vis,
                                          int B, int mode, int flags, int swap_order, float alpha_beta,
                                          const float* joint_xyz21, const float* uv21, const float* g_losses,
                                          float* g_rot, float* g_coeffs, float* g_betas, float* g_transl, float* g_scale,
                                          void* workspace, size_t workspace_bytes, mb_stream_t stream) {
    if (B < 0 || (flags & ~(MB_HEAD_XYZ | MB_HEAD_UV | MB_HEAD_REG | MB_HEAD_MATCH))) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!blob || !rot || !coeffs || !betas || !K || !joint_xyz21 || !uv21 || !g_losses || !g_rot || !g_coeffs || !g_betas || !workspace)
        return MB_E_NULL;
    if (workspace_bytes < mb_mano_head_loss_workspace_bytes(B)) return MB_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const HeadWs W = head_ws(B);
    char* ws = reinterpret_cast<char*>(workspace);
    double* accum = reinterpret_cast<double*>(ws + W.accum);
    const bool match = (flags & MB_HEAD_MATCH) != 0;
    const float* joints = match ? reinterpret_cast<const float*>(ws + W.joints) : joint_xyz21;
    float* g_xyz = reinterpret_cast<float*>(ws + W.g_xyz);
    float* g_uv = reinterpret_cast<float*>(ws + W.g_uv);
    float* g_joints = reinterpret_cast<float*>(ws + W.g_joints);
    int rc;
    // 1. d(loss terms) / d(xyz), d(uv), scaled by the upstream gradients of the terms
    const bool has_xyz = (flags & MB_HEAD_XYZ) != 0, has_uv = (flags & MB_HEAD_UV) != 0;
    if (has_xyz)
        if ((rc = mb_masked_l2_backward(joint_xyz21, gt_xyz, keypoint_vis, MB_VIS_F32, (long long)B * 21, 3, accum, g_losses,
                                        g_xyz, stream))) return rc;
    if (has_uv)
        if ((rc = mb_masked_l2_backward(uv21, gt_uv, keypoint_vis, MB_VIS_F32, (long long)B * 21, 2, accum + 2, g_losses + 1,
                                        g_uv, stream))) return rc;
    // 2. back through the projection (and match_mano_to_RHD) to the MANO joints
    if (match) {
        if ((rc = mb_joint_epilogue_backward(joints, index_root_bone_length, kp_coord_xyz_root, K, nullptr, has_xyz ? g_xyz : nullptr,
                                             has_uv ? g_uv : nullptr, B, swap_order, g_joints, nullptr, nullptr, stream))) return rc;
    } else {
        if (has_uv) {
            if ((rc = mb_project_uv_backward(joints, K, g_uv, B, 21, g_joints, stream))) return rc;
            if (has_xyz && (rc = launch_add(g_joints, g_xyz, (long long)B * 63, s))) return rc;
        } else if (has_xyz) {
            g_joints = g_xyz;
        } else {
            if ((rc = launch_zero(g_joints, (long long)B * 63, s))) return rc;
        }
    }
    // 3. joints-only MANO backward (+ the scale / translation post-op's own gradients)
    if ((rc = mb_mano_backward(blob, nc, rot, coeffs, betas, nullptr, g_joints, B, mode, 0, g_rot, g_coeffs, g_betas, nullptr, 0,
                               stream))) return rc;
    if (transl || scale)
        if ((rc = mb_affine_backward(nullptr, g_joints, nullptr, joints, scale, transl, B, nc, g_scale, g_transl, g_rot, g_coeffs,
                                     g_betas, stream))) return rc;
    // 4. the regulariser's gradient joins the pose / shape gradients
    if (flags & MB_HEAD_REG) {
        float* g_theta = reinterpret_cast<float*>(ws + W.g_theta);
        float* g_beta = reinterpret_cast<float*>(ws + W.g_beta);
        if ((rc = mb_regulariser_backward(coeffs, (long long)B * nc, betas, (long long)B * 10, alpha_beta, accum + 4, g_losses + 2,
                                          g_theta, g_beta, stream))) return rc;
        if ((rc = launch_add(g_coeffs, g_theta, (long long)B * nc, s))) return rc;
        if ((rc = launch_add(g_betas, g_beta, (long long)B * 10, s))) return rc;
    }
    return 0;
}
