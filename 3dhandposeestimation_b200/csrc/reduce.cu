// reduce.cu — visible-joint reductions (MPJPE / masked L2) and the fused Adam update.
//
// Reference: MPJPE.forward (criterions/metrics.py:10-27) and L2Loss.forward
// (criterions/loss.py:10-25): d = ||pred - gt|| (or its square) per joint, global mean over
// the joints whose visibility flag is non-zero, 0 when there is none.  The reference
// materialises masked_select (dynamic shape) and syncs the host on numel(); here the sum and
// the count are reduced on the device (fp64 accumulators, one atomic pair per block) and a
// one-thread finalise selects 0 for the empty case — no host round trip.
#include "common.cuh"

namespace mb {
namespace {

constexpr int RED_THREADS = 256;

__device__ __forceinline__ bool visible(const void* vis, int kind, long long i) {
    return kind == MB_VIS_U8 ? (reinterpret_cast<const uint8_t*>(vis)[i] != 0)
                             : (reinterpret_cast<const float*>(vis)[i] != 0.f);
}

__global__ void __launch_bounds__(RED_THREADS)
masked_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const void* __restrict__ vis,
                     int vis_kind, long long n, int kind, double* __restrict__ accum) {
    double sum = 0.0;
    double cnt = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (visible(vis, vis_kind, i)) {
            const float dx = pred[i * 3] - gt[i * 3], dy = pred[i * 3 + 1] - gt[i * 3 + 1], dz = pred[i * 3 + 2] - gt[i * 3 + 2];
            const float d2 = dx * dx + dy * dy + dz * dz;
            sum += (kind == MB_REDUCE_MPJPE_MM) ? (double)sqrtf(d2) : (double)d2;
            cnt += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ double s_sum[RED_THREADS / 32], s_cnt[RED_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < RED_THREADS / 32; ++w) { a += s_sum[w]; c += s_cnt[w]; }
        if (c > 0.0) { atomicAdd(&accum[0], a); atomicAdd(&accum[1], c); }
    }
}

__global__ void masked_finalize_kernel(const double* __restrict__ accum, int kind, float* __restrict__ out) {
    const double c = accum[1];
    double v = c > 0.0 ? accum[0] / c : 0.0;
    if (kind == MB_REDUCE_MPJPE_MM) v *= 1000.0;
    out[0] = (float)v;
}

__global__ void __launch_bounds__(RED_THREADS)
masked_l2_backward_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const void* __restrict__ vis,
                          int vis_kind, long long n, const double* __restrict__ accum, const float* __restrict__ g_out,
                          float* __restrict__ g_pred) {
    const double c = accum[1];
    const float scale = c > 0.0 ? (float)(2.0 * (double)g_out[0] / c) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = visible(vis, vis_kind, i) ? scale : 0.f;
        g_pred[i * 3] = m * (pred[i * 3] - gt[i * 3]);
        g_pred[i * 3 + 1] = m * (pred[i * 3 + 1] - gt[i * 3 + 1]);
        g_pred[i * 3 + 2] = m * (pred[i * 3 + 2] - gt[i * 3 + 2]);
    }
}

// LossCalculation.compute_hand_mask_loss (criterions/loss.py:92-111): uv -> int64 (truncation), clamp to
// [0, W-1] on both axes (the reference clamps rows with the last dimension too), sample hand_mask[b][v][u] at the
// predicted and the ground-truth keypoints; accum = {sum of predicted samples, sum of ground-truth samples}.
__device__ __forceinline__ double mask_at(const void* mask, int kind, long long b, long long HW, int W, float u, float v) {
    long long x = __float2ll_rz(u), y = __float2ll_rz(v);
    x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
    y = y < 0 ? 0 : (y > W - 1 ? W - 1 : y);
    const long long i = b * HW + y * W + x;
    return kind == MB_VIS_U8 ? (double)reinterpret_cast<const uint8_t*>(mask)[i] : (double)reinterpret_cast<const float*>(mask)[i];
}

__global__ void __launch_bounds__(RED_THREADS)
hand_mask_kernel(const float2* __restrict__ pred_uv, const float2* __restrict__ gt_uv, const void* __restrict__ mask, int mask_kind,
                 long long n, int N, int H, int W, double* __restrict__ accum) {
    double sp = 0.0, sg = 0.0;
    const long long HW = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / N;
        const float2 p = pred_uv[i], g = gt_uv[i];
        sp += mask_at(mask, mask_kind, b, HW, W, p.x, p.y);
        sg += mask_at(mask, mask_kind, b, HW, W, g.x, g.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
    }
    __shared__ double s_p[RED_THREADS / 32], s_g[RED_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_p[warp] = sp; s_g[warp] = sg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < RED_THREADS / 32; ++w) { a += s_p[w]; c += s_g[w]; }
        if (a != 0.0) atomicAdd(&accum[0], a);
        if (c != 0.0) atomicAdd(&accum[1], c);
    }
}

__global__ void hand_mask_finalize_kernel(const double* __restrict__ accum, float* __restrict__ out) {
    out[0] = 1.f - (float)accum[0] / ((float)accum[1] + 1e-8f);      // fp32 like the reference (epsilon :108)
}

// torch.optim.Adam (no weight decay, no amsgrad):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(RED_THREADS)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            float b1, float b2, float eps, float step_size, float inv_sqrt_bc2) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
        const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

inline unsigned red_grid(long long n) {
    long long blocks = (n + RED_THREADS - 1) / RED_THREADS;
    long long cap = (long long)NUM_SMS * 8;
    return (unsigned)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_masked_joint_reduce(const float* pred, const float* gt, const void* vis, int vis_kind,
                                      long long n_joints, int kind, double* accum, float* out, mb_stream_t stream) {
    if (n_joints < 0 || (kind != MB_REDUCE_MPJPE_MM && kind != MB_REDUCE_L2) ||
        (vis_kind != MB_VIS_F32 && vis_kind != MB_VIS_U8))
        return MB_E_RANGE;
    if (!accum || !out) return MB_E_NULL;
    if (n_joints > 0 && (!pred || !gt || !vis)) return MB_E_NULL;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(accum, 0, 2 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    if (n_joints > 0) {
        masked_reduce_kernel<<<red_grid(n_joints), RED_THREADS, 0, s>>>(pred, gt, vis, vis_kind, n_joints, kind, accum);
        int rc = cuda_rc();
        if (rc) return rc;
    }
    masked_finalize_kernel<<<1, 1, 0, s>>>(accum, kind, out);
    return cuda_rc();
}

extern "C" int mb_masked_l2_backward(const float* pred, const float* gt, const void* vis, int vis_kind,
                                     long long n_joints, const double* accum, const float* g_out,
                                     float* g_pred, mb_stream_t stream) {
    if (n_joints < 0 || (vis_kind != MB_VIS_F32 && vis_kind != MB_VIS_U8)) return MB_E_RANGE;
    if (n_joints == 0) return 0;
    if (!pred || !gt || !vis || !accum || !g_out || !g_pred) return MB_E_NULL;
    masked_l2_backward_kernel<<<red_grid(n_joints), RED_THREADS, 0, (cudaStream_t)stream>>>(pred, gt, vis, vis_kind, n_joints,
                                                                                              accum, g_out, g_pred);
    return cuda_rc();
}

extern "C" int mb_hand_mask_loss(const float* pred_uv, const float* gt_uv, const void* hand_mask, int mask_kind, int B, int N,
                                 int H, int W, double* accum, float* out, mb_stream_t stream) {
    if (B < 0 || N < 0 || H < 1 || W < 1 || W > H || (mask_kind != MB_VIS_F32 && mask_kind != MB_VIS_U8)) return MB_E_RANGE;
    if (!accum || !out) return MB_E_NULL;
    const long long n = (long long)B * N;
    if (n > 0 && (!pred_uv || !gt_uv || !hand_mask)) return MB_E_NULL;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(accum, 0, 2 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    if (n > 0) {
        hand_mask_kernel<<<red_grid(n), RED_THREADS, 0, s>>>(reinterpret_cast<const float2*>(pred_uv),
                                                             reinterpret_cast<const float2*>(gt_uv), hand_mask, mask_kind, n, N, H,
                                                             W, accum);
        int rc = cuda_rc();
        if (rc) return rc;
    }
    hand_mask_finalize_kernel<<<1, 1, 0, s>>>(accum, out);
    return cuda_rc();
}

extern "C" int mb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                            float lr, float beta1, float beta2, float eps, int step, mb_stream_t stream) {
    if (n < 0 || step < 1) return MB_E_RANGE;
    if (n == 0) return 0;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return MB_E_NULL;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<red_grid(n), RED_THREADS, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                                        (float)(lr / bc1), (float)(1.0 / sqrt(bc2)));
    return cuda_rc();
}
