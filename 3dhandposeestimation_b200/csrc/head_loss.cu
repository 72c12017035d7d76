// head_loss.cu — the MANO heads' tail as ONE call in each direction:
//   MANO parameters -> 21 joints (joints-only kernels, no 778-vertex contraction) [-> scale * p + transl]
//   [-> match_mano_to_RHD] -> pinhole projection -> L2Loss on xyz, L2Loss on uv, MANO regulariser.
// Reference chain: network/sub_modules/resnet50MANO.py:76-87 (mano_layer, scale / translation post-ops),
// network/Resnet50MANO3DHandPose.py:35-60 (match_mano_to_RHD), :71-73 (batch_project_xyz_to_uv),
// criterions/loss.py:10-25 (L2Loss), :83-87 (xyz / uv losses), :113-117 (regulariser), trainval.py:328-358.
//
// The two entry points enqueue, back to back on the caller's stream, the kernels that already implement the geometric
// pieces (joints-only MANO forward / backward, scale + translation, joint epilogue — each pinned to the reference on
// its own) and three small kernels of this file for the loss terms: all three reductions + their finalisation in one
// launch, both L2 gradients in one, the regulariser's gradient accumulated into the pose / shape gradients in one.
// No host round trip, no torch op and no allocation in between — at the heads' batch sizes (config.py:79: 200) the
// path is launch-latency-bound, and one ctypes call that can be captured in a CUDA graph replaces ~10 autograd
// nodes.  Forward: 3-4 launches + one 64-byte memset; backward: 4-5 launches.  All scratch lives in a caller-provided
// workspace that must stay untouched between the forward and its backward.
#include "common.cuh"

using namespace mb;

namespace {

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}
__global__ void zero_kernel(float* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = 0.f;
}
constexpr int HL_THREADS = 256;

// All three loss terms in ONE launch: fp64 block sums -> atomics into accum[0..5] = {S_xyz, N_xyz, S_uv, N_uv, sum theta^2,
// sum beta^2}; the block that finishes last (ticket in accum[6]) writes losses[0..2].  The same arithmetic as
// masked_reduce_kernel / sumsq2_kernel + their finalisers (reduce.cu), which remain the stand-alone drop-ins.
__global__ void __launch_bounds__(HL_THREADS)
head_reduce_kernel(const float* __restrict__ xyz, const float* __restrict__ gt_xyz, const float* __restrict__ uv,
                   const float* __restrict__ gt_uv, const float* __restrict__ vis, long long nj,
                   const float* __restrict__ theta, long long n_theta, const float* __restrict__ beta, long long n_beta,
                   int flags, float alpha_beta, double* __restrict__ accum, float* __restrict__ losses) {
    double v[6] = {0, 0, 0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (flags & (MB_HEAD_XYZ | MB_HEAD_UV))
        for (long long i = i0; i < nj; i += stride)
            if (vis[i] != 0.f) {
                if (flags & MB_HEAD_XYZ) {
                    const float a = xyz[i * 3] - gt_xyz[i * 3], b = xyz[i * 3 + 1] - gt_xyz[i * 3 + 1], c = xyz[i * 3 + 2] - gt_xyz[i * 3 + 2];
                    v[0] += (double)(a * a + b * b + c * c);       // same order as the reference's sum(dim=2)
                    v[1] += 1.0;
                }
                if (flags & MB_HEAD_UV) {
                    const float a = uv[i * 2] - gt_uv[i * 2], b = uv[i * 2 + 1] - gt_uv[i * 2 + 1];
                    v[2] += (double)(a * a + b * b);
                    v[3] += 1.0;
                }
            }
    if (flags & MB_HEAD_REG) {
        for (long long i = i0; i < n_theta; i += stride) { const float x = theta[i]; v[4] += (double)x * x; }
        for (long long i = i0; i < n_beta; i += stride) { const float x = beta[i]; v[5] += (double)x * x; }
    }
    __shared__ double sh[HL_THREADS / 32][6];
    __shared__ bool last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 6; ++k) {
            double x = 0.0;
            for (int w = 0; w < HL_THREADS / 32; ++w) x += sh[w][k];
            if (x != 0.0) atomicAdd(&accum[k], x);
        }
        __threadfence();
        last = atomicAdd(reinterpret_cast<unsigned long long*>(&accum[6]), 1ULL) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const volatile double* a = accum;
        losses[0] = (flags & MB_HEAD_XYZ) && a[1] > 0.0 ? (float)(a[0] / a[1]) : 0.f;
        losses[1] = (flags & MB_HEAD_UV) && a[3] > 0.0 ? (float)(a[2] / a[3]) : 0.f;
        losses[2] = (flags & MB_HEAD_REG) ? (sqrtf((float)a[4]) + alpha_beta * sqrtf((float)a[5])) / 100.f : 0.f;
    }
}

// d(loss_xyz) / d(xyz) and d(loss_uv) / d(uv) in one launch (masked_l2_backward_kernel twice)
__global__ void __launch_bounds__(HL_THREADS)
head_l2_backward_kernel(const float* __restrict__ xyz, const float* __restrict__ gt_xyz, const float* __restrict__ uv,
                        const float* __restrict__ gt_uv, const float* __restrict__ vis, long long nj, int flags,
                        const double* __restrict__ accum, const float* __restrict__ g_losses, float* __restrict__ g_xyz,
                        float* __restrict__ g_uv) {
    const float sx = (flags & MB_HEAD_XYZ) && accum[1] > 0.0 ? (float)(2.0 * (double)g_losses[0] / accum[1]) : 0.f;
    const float su = (flags & MB_HEAD_UV) && accum[3] > 0.0 ? (float)(2.0 * (double)g_losses[1] / accum[3]) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nj; i += (long long)gridDim.x * blockDim.x) {
        const bool m = vis[i] != 0.f;
        if (flags & MB_HEAD_XYZ) {
            const float k = m ? sx : 0.f;
            for (int c = 0; c < 3; ++c) g_xyz[i * 3 + c] = k * (xyz[i * 3 + c] - gt_xyz[i * 3 + c]);
        }
        if (flags & MB_HEAD_UV) {
            const float k = m ? su : 0.f;
            for (int c = 0; c < 2; ++c) g_uv[i * 2 + c] = k * (uv[i * 2 + c] - gt_uv[i * 2 + c]);
        }
    }
}

// the regulariser's gradient ADDED to the pose / shape gradients: g_theta += g theta / (100 ||theta||),
// g_beta += g alpha beta / (100 ||beta||), 0 at a zero norm (regulariser_backward_kernel + two accumulations)
__global__ void __launch_bounds__(HL_THREADS)
head_reg_backward_add_kernel(const float* __restrict__ theta, long long n_theta, const float* __restrict__ beta, long long n_beta,
                             float alpha_beta, const double* __restrict__ accum, const float* __restrict__ g_losses,
                             float* __restrict__ g_theta, float* __restrict__ g_beta) {
    const float nt = sqrtf((float)accum[4]), nb = sqrtf((float)accum[5]), g = g_losses[2];
    const float kt = nt > 0.f ? g / (100.f * nt) : 0.f, kb = nb > 0.f ? g * alpha_beta / (100.f * nb) : 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = i0; i < n_theta; i += stride) g_theta[i] += kt * theta[i];
    for (long long i = i0; i < n_beta; i += stride) g_beta[i] += kb * beta[i];
}
unsigned hl_grid(long long n) {
    const long long b = (n + HL_THREADS - 1) / HL_THREADS;
    return (unsigned)(b < 1 ? 1 : (b < NUM_SMS * 4 ? b : NUM_SMS * 4));
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
    size_t accum;       // double [8]: {sum, count} of the xyz loss, of the uv loss, {sum theta^2, sum beta^2}, block ticket
    size_t total;
};
HeadWs head_ws(long long B) {
    HeadWs W;
    size_t o = 0;
    W.joints = o;   o = align256(o + sizeof(float) * B * 63);
    W.g_xyz = o;    o = align256(o + sizeof(float) * B * 63);
    W.g_uv = o;     o = align256(o + sizeof(float) * B * 42);
    W.g_joints = o; o = align256(o + sizeof(float) * B * 63);
    W.accum = o;    o = align256(o + sizeof(double) * 8);
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
    // 3. the three loss terms in one launch, each a device scalar (terms not asked for are 0)
    if (cudaError_t e = cudaMemsetAsync(accum, 0, 8 * sizeof(double), s)) return (int)e;
    const long long nj = (long long)B * 21, nt = (long long)B * nc, nb = (long long)B * 10;
    head_reduce_kernel<<<hl_grid(nj > nt ? nj : nt), HL_THREADS, 0, s>>>(joint_xyz21, gt_xyz, uv21, gt_uv, keypoint_vis, nj, coeffs, nt, betas,
                                                                         nb, flags, alpha_beta, accum, losses);
    if ((rc = cuda_rc())) return rc;
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
    if (has_xyz || has_uv) {
        head_l2_backward_kernel<<<hl_grid((long long)B * 21), HL_THREADS, 0, s>>>(joint_xyz21, gt_xyz, uv21, gt_uv, keypoint_vis, (long long)B * 21,
                                                                                  flags, accum, g_losses, g_xyz, g_uv);
        if ((rc = cuda_rc())) return rc;
    }
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
        const long long nt = (long long)B * nc, nb = (long long)B * 10;
        head_reg_backward_add_kernel<<<hl_grid(nt), HL_THREADS, 0, s>>>(coeffs, nt, betas, nb, alpha_beta, accum, g_losses, g_coeffs, g_betas);
        if ((rc = cuda_rc())) return rc;
    }
    return 0;
}
