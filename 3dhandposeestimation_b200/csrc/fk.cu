// fk.cu — RHD 21-joint forward-kinematics layer + pinhole projection, forward and backward.
// One thread per sample (59 floats in, 105 out — no cross-sample data), rows staged through
// shared memory so every global access is coalesced; odd smem row pitches keep the per-thread
// row walks bank-conflict free.  Math in fk_math.cuh (reference citations there).
#include "fk_math.cuh"

namespace mb {
namespace {

constexpr int FK_THREADS = 64;
// smem row pitches (floats), all odd
constexpr int P_RA = 3, P_OA = 23, P_BL = 21, P_K = 9, P_S = 1, P_ROOT = 3, P_XYZ = 63, P_UV = 43;

// coalesced copy of `rows` rows of width w from global (dense) to smem (pitch p), as asynchronous 4-byte copies
// (cp.async): a thread issues all of its ~60-165 requests back to back and waits once (fk_stage_wait), instead of
// paying one global-load round trip per element — with 64-thread blocks that latency chain was 40 % of the kernel
__device__ __forceinline__ void stage_in(float* dst, int p, const float* __restrict__ src, int w, long long row0, int rows) {
    const float* g = src + row0 * w;
    const int n = rows * w;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int r = i / w, c = i - r * w;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(dst + r * p + c)), "l"(g + i)
                     : "memory");
    }
}
__device__ __forceinline__ void fk_stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void stage_out(float* __restrict__ dst, int w, long long row0, int rows, const float* src, int p) {
    float* g = dst + row0 * w;
    const int n = rows * w;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int r = i / w, c = i - r * w;
        g[i] = src[r * p + c];
    }
}

__global__ void __launch_bounds__(FK_THREADS)
fk_forward_kernel(const float* __restrict__ ra, const float* __restrict__ oa, const float* __restrict__ bl,
                  const float* __restrict__ K, const float* __restrict__ sc, const float* __restrict__ root,
                  int B, int swap, float* __restrict__ xyz, float* __restrict__ uv) {
    __shared__ float s_ra[FK_THREADS * P_RA], s_oa[FK_THREADS * P_OA], s_bl[FK_THREADS * P_BL], s_K[FK_THREADS * P_K],
        s_s[FK_THREADS * P_S], s_root[FK_THREADS * P_ROOT], s_xyz[FK_THREADS * P_XYZ], s_uv[FK_THREADS * P_UV];
    const int nblk = (B + FK_THREADS - 1) / FK_THREADS;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long row0 = (long long)blk * FK_THREADS;
        const int rows = (B - row0) < FK_THREADS ? (int)(B - row0) : FK_THREADS;
        __syncthreads();
        stage_in(s_ra, P_RA, ra, 3, row0, rows);
        stage_in(s_oa, P_OA, oa, 23, row0, rows);
        stage_in(s_bl, P_BL, bl, 20, row0, rows);
        stage_in(s_K, P_K, K, 9, row0, rows);
        stage_in(s_s, P_S, sc, 1, row0, rows);
        stage_in(s_root, P_ROOT, root, 3, row0, rows);
        fk_stage_wait();
        __syncthreads();
        const int t = threadIdx.x;
        if (t < rows)
            fk_forward_sample(s_ra + t * P_RA, s_oa + t * P_OA, s_bl + t * P_BL, s_K + t * P_K, s_s[t], s_root + t * P_ROOT,
                              swap, s_xyz + t * P_XYZ, s_uv + t * P_UV);
        __syncthreads();
        stage_out(xyz, 63, row0, rows, s_xyz, P_XYZ);
        stage_out(uv, 42, row0, rows, s_uv, P_UV);
    }
}

__global__ void __launch_bounds__(FK_THREADS)
fk_backward_kernel(const float* __restrict__ ra, const float* __restrict__ oa, const float* __restrict__ bl,
                   const float* __restrict__ K, const float* __restrict__ sc, const float* __restrict__ root,
                   const float* __restrict__ g_xyz, const float* __restrict__ g_uv, int B, int swap,
                   float* __restrict__ g_ra, float* __restrict__ g_oa, float* __restrict__ g_bl) {
    __shared__ float s_ra[FK_THREADS * P_RA], s_oa[FK_THREADS * P_OA], s_bl[FK_THREADS * P_BL], s_K[FK_THREADS * P_K],
        s_s[FK_THREADS * P_S], s_root[FK_THREADS * P_ROOT], s_xyz[FK_THREADS * P_XYZ], s_uv[FK_THREADS * P_UV];
    const int nblk = (B + FK_THREADS - 1) / FK_THREADS;
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long row0 = (long long)blk * FK_THREADS;
        const int rows = (B - row0) < FK_THREADS ? (int)(B - row0) : FK_THREADS;
        __syncthreads();
        stage_in(s_ra, P_RA, ra, 3, row0, rows);
        stage_in(s_oa, P_OA, oa, 23, row0, rows);
        stage_in(s_bl, P_BL, bl, 20, row0, rows);
        stage_in(s_K, P_K, K, 9, row0, rows);
        stage_in(s_s, P_S, sc, 1, row0, rows);
        stage_in(s_root, P_ROOT, root, 3, row0, rows);
        if (g_xyz) stage_in(s_xyz, P_XYZ, g_xyz, 63, row0, rows);
        if (g_uv) stage_in(s_uv, P_UV, g_uv, 42, row0, rows);
        fk_stage_wait();
        __syncthreads();
        const int t = threadIdx.x;
        float r_gra[3], r_goa[FK_OA], r_gbl[FK_NODES];
        if (t < rows)
            fk_backward_sample(s_ra + t * P_RA, s_oa + t * P_OA, s_bl + t * P_BL, s_K + t * P_K, s_s[t], s_root + t * P_ROOT,
                               swap, g_xyz ? s_xyz + t * P_XYZ : nullptr, g_uv ? s_uv + t * P_UV : nullptr,
                               r_gra, r_goa, r_gbl);
        __syncthreads();                      // everyone is done reading the staged inputs
        if (t < rows) {                       // reuse the input tiles as output tiles (same pitches)
#pragma unroll
            for (int i = 0; i < 3; ++i) s_ra[t * P_RA + i] = r_gra[i];
#pragma unroll
            for (int i = 0; i < FK_OA; ++i) s_oa[t * P_OA + i] = r_goa[i];
#pragma unroll
            for (int i = 0; i < FK_NODES; ++i) s_bl[t * P_BL + i] = r_gbl[i];
        }
        __syncthreads();
        stage_out(g_ra, 3, row0, rows, s_ra, P_RA);
        stage_out(g_oa, 23, row0, rows, s_oa, P_OA);
        stage_out(g_bl, 20, row0, rows, s_bl, P_BL);
    }
}

__global__ void project_forward_kernel(const float* __restrict__ xyz, const float* __restrict__ K, long long total, int N,
                                       float* __restrict__ uv) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float* k = K + (i / N) * 9;
        float u, v;
        project_point(k, xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], u, v);
        uv[i * 2] = u; uv[i * 2 + 1] = v;
    }
}
__global__ void project_backward_kernel(const float* __restrict__ xyz, const float* __restrict__ K,
                                        const float* __restrict__ g_uv, long long total, int N, float* __restrict__ g_xyz) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float* k = K + (i / N) * 9;
        V3 g = project_point_bwd(k, xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], g_uv[i * 2], g_uv[i * 2 + 1]);
        g_xyz[i * 3] = g.x; g_xyz[i * 3 + 1] = g.y; g_xyz[i * 3 + 2] = g.z;
    }
}

inline int fk_grid(int B) {
    int nblk = (B + FK_THREADS - 1) / FK_THREADS;
    int cap = NUM_SMS * 16;
    return nblk < cap ? nblk : cap;
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" int mb_fk_forward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                             const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                             int B, int swap_order, float* xyz, float* uv, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!root_angles || !other_angles || !bone_lengths || !K || !index_root_bone_length || !kp_coord_xyz_root || !xyz || !uv)
        return MB_E_NULL;
    StageTimer t(ST_FK_FWD, (cudaStream_t)stream);
    fk_forward_kernel<<<fk_grid(B), FK_THREADS, 0, (cudaStream_t)stream>>>(root_angles, other_angles, bone_lengths, K,
                                                                            index_root_bone_length, kp_coord_xyz_root, B,
                                                                            swap_order != 0, xyz, uv);
    return cuda_rc();
}

extern "C" int mb_fk_backward(const float* root_angles, const float* other_angles, const float* bone_lengths,
                              const float* K, const float* index_root_bone_length, const float* kp_coord_xyz_root,
                              const float* g_xyz, const float* g_uv, int B, int swap_order,
                              float* g_root_angles, float* g_other_angles, float* g_bone_lengths, mb_stream_t stream) {
    if (B < 0) return MB_E_RANGE;
    if (B == 0) return 0;
    if (!root_angles || !other_angles || !bone_lengths || !K || !index_root_bone_length || !kp_coord_xyz_root ||
        !g_root_angles || !g_other_angles || !g_bone_lengths)
        return MB_E_NULL;
    StageTimer t(ST_FK_BWD, (cudaStream_t)stream);
    fk_backward_kernel<<<fk_grid(B), FK_THREADS, 0, (cudaStream_t)stream>>>(root_angles, other_angles, bone_lengths, K,
                                                                             index_root_bone_length, kp_coord_xyz_root,
                                                                             g_xyz, g_uv, B, swap_order != 0,
                                                                             g_root_angles, g_other_angles, g_bone_lengths);
    return cuda_rc();
}

extern "C" int mb_project_uv_forward(const float* xyz, const float* K, int B, int N, float* uv, mb_stream_t stream) {
    if (B < 0 || N < 0) return MB_E_RANGE;
    if (B == 0 || N == 0) return 0;
    if (!xyz || !K || !uv) return MB_E_NULL;
    const long long total = (long long)B * N;
    const long long blocks = (total + 255) / 256;
    project_forward_kernel<<<(unsigned)(blocks < NUM_SMS * 16 ? blocks : NUM_SMS * 16), 256, 0, (cudaStream_t)stream>>>(
        xyz, K, total, N, uv);
    return cuda_rc();
}

extern "C" int mb_project_uv_backward(const float* xyz, const float* K, const float* g_uv, int B, int N,
                                      float* g_xyz, mb_stream_t stream) {
    if (B < 0 || N < 0) return MB_E_RANGE;
    if (B == 0 || N == 0) return 0;
    if (!xyz || !K || !g_uv || !g_xyz) return MB_E_NULL;
    const long long total = (long long)B * N;
    const long long blocks = (total + 255) / 256;
    project_backward_kernel<<<(unsigned)(blocks < NUM_SMS * 16 ? blocks : NUM_SMS * 16), 256, 0, (cudaStream_t)stream>>>(
        xyz, K, g_uv, total, N, g_xyz);
    return cuda_rc();
}
