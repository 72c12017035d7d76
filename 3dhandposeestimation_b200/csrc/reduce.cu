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
                     int vis_kind, long long n, int dim, int kind, double* __restrict__ accum) {
    // dim = components per joint: 3 for xyz, 2 for the uv loss (loss.py:86-87 sends [B,21,2] through the same L2Loss)
    double sum = 0.0;
    double cnt = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (visible(vis, vis_kind, i)) {
            float d2 = 0.f;
            for (int c = 0; c < dim; ++c) {                     // same order as the reference's sum(dim=2)
                const float d = pred[i * dim + c] - gt[i * dim + c];
                d2 = c == 0 ? d * d : d2 + d * d;
            }
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
                          int vis_kind, long long n, int dim, const double* __restrict__ accum, const float* __restrict__ g_out,
                          float* __restrict__ g_pred) {
    const double c = accum[1];
    const float scale = c > 0.0 ? (float)(2.0 * (double)g_out[0] / c) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float m = visible(vis, vis_kind, i) ? scale : 0.f;
        for (int k = 0; k < dim; ++k) g_pred[i * dim + k] = m * (pred[i * dim + k] - gt[i * dim + k]);
    }
}

// LossCalculation.compute_regularization_loss (criterions/loss.py:113-117): (||theta||_F + alpha ||beta||_F) / 100
// over the whole batch.  accum = {sum theta^2, sum beta^2} in fp64.
__global__ void __launch_bounds__(RED_THREADS)
sumsq2_kernel(const float* __restrict__ a, long long na, const float* __restrict__ b, long long nb, double* __restrict__ accum) {
    double sa = 0.0, sb = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += stride) { const float x = a[i]; sa += (double)x * x; }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += stride) { const float x = b[i]; sb += (double)x * x; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    __shared__ double s_a[RED_THREADS / 32], s_b[RED_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_a[warp] = sa; s_b[warp] = sb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
        for (int w = 0; w < RED_THREADS / 32; ++w) { x += s_a[w]; y += s_b[w]; }
        atomicAdd(&accum[0], x);
        atomicAdd(&accum[1], y);
    }
}
__global__ void regulariser_finalize_kernel(const double* __restrict__ accum, float alpha_beta, float* __restrict__ out) {
    out[0] = (sqrtf((float)accum[0]) + alpha_beta * sqrtf((float)accum[1])) / 100.f;
}
// d/dtheta = g theta / (100 ||theta||), d/dbeta = g alpha beta / (100 ||beta||); 0 at a zero norm (torch.norm's subgradient)
__global__ void __launch_bounds__(RED_THREADS)
regulariser_backward_kernel(const float* __restrict__ a, long long na, const float* __restrict__ b, long long nb,
                            const double* __restrict__ accum, float alpha_beta, const float* __restrict__ g_out,
                            float* __restrict__ ga, float* __restrict__ gb) {
    const float na2 = sqrtf((float)accum[0]), nb2 = sqrtf((float)accum[1]);
    const float g = g_out[0] / 100.f;
    const float ka = na2 > 0.f ? g / na2 : 0.f, kb = nb2 > 0.f ? g * alpha_beta / nb2 : 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += stride) ga[i] = ka * a[i];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += stride) gb[i] = kb * b[i];
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

// End of a fused fitting iteration (fitting.py), after the caller all-reduced the kernel's partials
// {sum vis |d|^2, sum theta_new^2, sum beta_new^2}: the loss of the iteration just done — L2 term over the global
// visible count + the regulariser of the parameters the iteration STARTED from (globals) — then the norms of the
// updated parameters become the next iteration's globals.  One thread: replaces ~10 tiny elementwise launches.
__global__ void fit_finalize_kernel(double* __restrict__ globals, const double* __restrict__ partials, int regularize,
                                    double* __restrict__ out4, double* __restrict__ loss) {
    const double n = globals[0], t = regularize ? globals[1] : 0.0, b = regularize ? globals[2] : 0.0;
    out4[0] = partials[0]; out4[1] = n; out4[2] = t; out4[3] = b;
    loss[0] = (n > 0.0 ? partials[0] / n : 0.0) + (sqrt(t) + 10.0 * sqrt(b)) / 100.0;
    globals[1] = partials[1];
    globals[2] = partials[2];
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
                                      long long n_joints, int dim, int kind, double* accum, float* out, mb_stream_t stream) {
    if (n_joints < 0 || dim < 1 || dim > 4 || (kind != MB_REDUCE_MPJPE_MM && kind != MB_REDUCE_L2) ||
        (vis_kind != MB_VIS_F32 && vis_kind != MB_VIS_U8))
        return MB_E_RANGE;
    if (!accum || !out) return MB_E_NULL;
    if (n_joints > 0 && (!pred || !gt || !vis)) return MB_E_NULL;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(accum, 0, 2 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    if (n_joints > 0) {
        masked_reduce_kernel<<<red_grid(n_joints), RED_THREADS, 0, s>>>(pred, gt, vis, vis_kind, n_joints, dim, kind, accum);
        int rc = cuda_rc();
        if (rc) return rc;
    }
    masked_finalize_kernel<<<1, 1, 0, s>>>(accum, kind, out);
    return cuda_rc();
}

extern "C" int mb_masked_l2_backward(const float* pred, const float* gt, const void* vis, int vis_kind,
                                     long long n_joints, int dim, const double* accum, const float* g_out,
                                     float* g_pred, mb_stream_t stream) {
    if (n_joints < 0 || dim < 1 || dim > 4 || (vis_kind != MB_VIS_F32 && vis_kind != MB_VIS_U8)) return MB_E_RANGE;
    if (n_joints == 0) return 0;
    if (!pred || !gt || !vis || !accum || !g_out || !g_pred) return MB_E_NULL;
    masked_l2_backward_kernel<<<red_grid(n_joints), RED_THREADS, 0, (cudaStream_t)stream>>>(pred, gt, vis, vis_kind, n_joints,
                                                                                              dim, accum, g_out, g_pred);
    return cuda_rc();
}

extern "C" int mb_regulariser_forward(const float* theta, long long n_theta, const float* beta, long long n_beta,
                                      float alpha_beta, double* accum, float* out, mb_stream_t stream) {
    if (n_theta < 0 || n_beta < 0) return MB_E_RANGE;
    if (!accum || !out) return MB_E_NULL;
    if ((n_theta > 0 && !theta) || (n_beta > 0 && !beta)) return MB_E_NULL;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(accum, 0, 2 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    if (n_theta + n_beta > 0) {
        sumsq2_kernel<<<red_grid(n_theta > n_beta ? n_theta : n_beta), RED_THREADS, 0, s>>>(theta, n_theta, beta, n_beta, accum);
        int rc = cuda_rc();
        if (rc) return rc;
    }
    regulariser_finalize_kernel<<<1, 1, 0, s>>>(accum, alpha_beta, out);
    return cuda_rc();
}

extern "C" int mb_regulariser_backward(const float* theta, long long n_theta, const float* beta, long long n_beta,
                                       float alpha_beta, const double* accum, const float* g_out, float* g_theta,
                                       float* g_beta, mb_stream_t stream) {
    if (n_theta < 0 || n_beta < 0) return MB_E_RANGE;
    if (n_theta + n_beta == 0) return 0;
    if (!accum || !g_out || (n_theta > 0 && (!theta || !g_theta)) || (n_beta > 0 && (!beta || !g_beta))) return MB_E_NULL;
    regulariser_backward_kernel<<<red_grid(n_theta > n_beta ? n_theta : n_beta), RED_THREADS, 0, (cudaStream_t)stream>>>(
        theta, n_theta, beta, n_beta, accum, alpha_beta, g_out, g_theta, g_beta);
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

extern "C" int mb_fit_finalize(double* globals, const double* partials, int regularize, double* out4, double* loss,
                               mb_stream_t stream) {
    if (!globals || !partials || !out4 || !loss) return MB_E_NULL;
    fit_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(globals, partials, regularize, out4, loss);
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
