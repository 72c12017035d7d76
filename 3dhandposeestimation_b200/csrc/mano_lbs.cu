// mano_lbs.cu — linear blend skinning over 778 vertices x 16 bones, forward and backward.
//
// Reference: MANOLayer.py:177-185 (T = sum_k w_vk A_k ; v' = T [v_posed;1]), :190-202 (the five
// fingertip vertices become joints 4,8,12,16,20) and :188,:204-205 (global rotation — already
// folded into the bone transforms by the pose stage, so it costs nothing here).
//
// The stage is HBM-bound: per hand it reads v_posed (9.3 KB) + 16 bone transforms (768 B) and
// writes verts (9.3 KB) + 5 tip joints.  Rows are staged through shared memory so that every
// global access is a coalesced 16-byte vector: hands are processed in PAIRS because one hand's
// 2334 floats are only 8-byte aligned in the [B][778][3] output, two hands are 16-byte aligned.
// Skinning weights live in shared memory slot-major ([slot][vertex]) so a warp's reads are
// conflict-free; each thread owns fixed vertices (tid, tid+256, ...).
#include "common.cuh"

namespace mb {
namespace {

constexpr int LBS_THREADS = 256;
constexpr int VPAD = 784;                       // 778 rounded up (slot-major weight rows)
constexpr int ROW4 = VP_PITCH / 4;              // 584 float4 per padded row
constexpr int MAX_CSC = 3200;                   // >= total skin nnz supported by the backward kernel
__constant__ int c_tip_vert[5] = {333, 444, 672, 555, 745};
__constant__ int c_tip_slot[5] = {4, 8, 12, 16, 20};

struct SkinShared {
    float w[MAX_INFL][VPAD];
    uint8_t b[MAX_INFL][VPAD];
    uint8_t cnt[VPAD];
};

__device__ __forceinline__ void stage_skin(SkinShared& S, const void* blob) {
    const BlobLayout L = blob_layout();
    const float* sw = blob_ptr<float>(blob, L.skin_w);
    const uint8_t* sb = blob_ptr<uint8_t>(blob, L.skin_b);
    const uint8_t* sc = blob_ptr<uint8_t>(blob, L.skin_cnt);
    for (int i = threadIdx.x; i < NV * MAX_INFL; i += blockDim.x) {
        int v = i / MAX_INFL, s = i % MAX_INFL;
        S.w[s][v] = sw[i];
        S.b[s][v] = sb[i];
    }
    for (int i = threadIdx.x; i < NV; i += blockDim.x) S.cnt[i] = sc[i];
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// same without .nc: for rows that this kernel overwrites later (dv_posed aliases v_posed)
__device__ __forceinline__ float4 ld_stream_rw(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(LBS_THREADS)
lbs_forward_kernel(const void* __restrict__ blob, const float* __restrict__ v_posed, int pitch,
                   const float* __restrict__ bone, int B, float* __restrict__ verts, float* __restrict__ joints) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SkinShared& S = *reinterpret_cast<SkinShared*>(smem_raw);
    float* s_vp = reinterpret_cast<float*>(smem_raw + ((sizeof(SkinShared) + 15) & ~size_t(15)));  // [2][VP_PITCH]
    float* s_bone = s_vp + 2 * VP_PITCH;                                                            // [2][192]
    stage_skin(S, blob);
    const int tid = threadIdx.x;
    const int npairs = (B + 1) / 2;
    const int pitch4 = pitch / 4;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const long long h0 = 2LL * pair;
        const int nh = (B - h0) >= 2 ? 2 : 1;
        __syncthreads();                          // previous iteration's readers are done (also covers stage_skin)
        for (int i = tid; i < nh * ROW4; i += LBS_THREADS) {
            int hh = i / ROW4, q = i - hh * ROW4;
            float4 v = (q < pitch4) ? ld_stream(reinterpret_cast<const float4*>(v_posed + (h0 + hh) * pitch) + q)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4*>(s_vp)[hh * ROW4 + q] = v;
        }
        for (int i = tid; i < nh * (NJ * BONE_F / 4); i += LBS_THREADS)
            reinterpret_cast<float4*>(s_bone)[i] = reinterpret_cast<const float4*>(bone + h0 * (NJ * BONE_F))[i];
        __syncthreads();
        for (int hh = 0; hh < nh; ++hh) {
            float* vp = s_vp + hh * VP_PITCH;
            const float* A0 = s_bone + hh * (NJ * BONE_F);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int v = tid + r * LBS_THREADS;
                if (v < NV) {
                    const float x = vp[v * 3], y = vp[v * 3 + 1], z = vp[v * 3 + 2];
                    float ox = 0.f, oy = 0.f, oz = 0.f;
                    const int cnt = S.cnt[v];
                    for (int s = 0; s < cnt; ++s) {
                        const float w = S.w[s][v];
                        const float4* A = reinterpret_cast<const float4*>(A0 + S.b[s][v] * BONE_F);
                        const float4 r0 = A[0], r1 = A[1], r2 = A[2];
                        ox = fmaf(w, fmaf(r0.x, x, fmaf(r0.y, y, fmaf(r0.z, z, r0.w))), ox);
                        oy = fmaf(w, fmaf(r1.x, x, fmaf(r1.y, y, fmaf(r1.z, z, r1.w))), oy);
                        oz = fmaf(w, fmaf(r2.x, x, fmaf(r2.y, y, fmaf(r2.z, z, r2.w))), oz);
                    }
                    vp[v * 3] = ox; vp[v * 3 + 1] = oy; vp[v * 3 + 2] = oz;   // in place: this thread owns vertex v
                }
            }
        }
        __syncthreads();
        // coalesced float4 stores of the pair's nh*2334 contiguous output floats
        float* out = verts + h0 * NVC;
        const int n = nh * NVC, nq = n >> 2;
        for (int q = tid; q < nq; q += LBS_THREADS) {
            float4 v;
            int e = q * 4;
            int a0 = e < NVC ? e : e + (VP_PITCH - NVC);            ++e;
            int a1 = e < NVC ? e : e + (VP_PITCH - NVC);            ++e;
            int a2 = e < NVC ? e : e + (VP_PITCH - NVC);            ++e;
            int a3 = e < NVC ? e : e + (VP_PITCH - NVC);
            v.x = s_vp[a0]; v.y = s_vp[a1]; v.z = s_vp[a2]; v.w = s_vp[a3];
            st_stream(reinterpret_cast<float4*>(out) + q, v);
        }
        if (tid < (n & 3)) out[nq * 4 + tid] = s_vp[nq * 4 + tid];      // nh == 1: two tail floats (< NVC)
        if (joints != nullptr && tid < nh * 15) {
            int hh = tid / 15, tc = tid % 15;
            joints[(h0 + hh) * (NOUTJ * 3) + c_tip_slot[tc / 3] * 3 + tc % 3] =
                s_vp[hh * VP_PITCH + c_tip_vert[tc / 3] * 3 + tc % 3];
        }
    }
}

// ----------------------------------------------------------------- backward
// dv_posed_v = sum_k w_vk R'_k^T g_v ;  dA'_k = sum_v w_vk g_v (x) [v_posed_v ; 1]   (A.2 steps 1-2)
struct CscShared {
    int ptr[NJ + 1];
    int v[MAX_CSC];
    float w[MAX_CSC];
};

__global__ void __launch_bounds__(LBS_THREADS)
lbs_backward_kernel(const void* __restrict__ blob, const float* v_posed, int pitch,
                    const float* __restrict__ bone, const float* __restrict__ g_verts,
                    const float* __restrict__ g_joints, int B,
                    float* dv_posed, float* __restrict__ dbone) {   // dv_posed may alias v_posed (row-wise read-then-write)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SkinShared& S = *reinterpret_cast<SkinShared*>(smem_raw);
    size_t off = (sizeof(SkinShared) + 15) & ~size_t(15);
    CscShared& C = *reinterpret_cast<CscShared*>(smem_raw + off);
    off = (off + sizeof(CscShared) + 15) & ~size_t(15);
    float* s_vp = reinterpret_cast<float*>(smem_raw + off);     // [VP_PITCH]
    float* s_g = s_vp + VP_PITCH;                               // [VP_PITCH]
    float* s_dv = s_g + VP_PITCH;                               // [VP_PITCH]
    float* s_bone = s_dv + VP_PITCH;                            // [192]
    stage_skin(S, blob);
    {
        const BlobLayout L = blob_layout();
        const int* cp = blob_ptr<int>(blob, L.csc_ptr);
        const int* cv = blob_ptr<int>(blob, L.csc_v);
        const float* cw = blob_ptr<float>(blob, L.csc_w);
        for (int i = threadIdx.x; i <= NJ; i += blockDim.x) C.ptr[i] = cp[i];
        const int nnz = cp[NJ];
        for (int i = threadIdx.x; i < nnz && i < MAX_CSC; i += blockDim.x) { C.v[i] = cv[i]; C.w[i] = cw[i]; }
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pitch4 = pitch / 4;
    for (long long hand = blockIdx.x; hand < B; hand += gridDim.x) {
        __syncthreads();
        for (int q = tid; q < ROW4; q += LBS_THREADS)
            reinterpret_cast<float4*>(s_vp)[q] = (q < pitch4)
                ? ld_stream_rw(reinterpret_cast<const float4*>(v_posed + hand * pitch) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        {   // g_verts rows are 8-byte aligned: float2 loads
            const float2* g2 = reinterpret_cast<const float2*>(g_verts + hand * NVC);
            for (int q = tid; q < NVC / 2; q += LBS_THREADS) reinterpret_cast<float2*>(s_g)[q] = g2[q];
        }
        if (tid < NJ * BONE_F / 4)
            reinterpret_cast<float4*>(s_bone)[tid] = reinterpret_cast<const float4*>(bone + hand * (NJ * BONE_F))[tid];
        __syncthreads();
        if (tid < 15)   // fingertip joints are vertices: their upstream gradient joins g_verts (A.2 step 1)
            s_g[c_tip_vert[tid / 3] * 3 + tid % 3] += g_joints[hand * (NOUTJ * 3) + c_tip_slot[tid / 3] * 3 + tid % 3];
        __syncthreads();
        // (a) per-vertex gather: dv = sum_s w R'^T g
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int v = tid + r * LBS_THREADS;
            if (v < NV) {
                const float gx = s_g[v * 3], gy = s_g[v * 3 + 1], gz = s_g[v * 3 + 2];
                float dx = 0.f, dy = 0.f, dz = 0.f;
                const int cnt = S.cnt[v];
                for (int s = 0; s < cnt; ++s) {
                    const float w = S.w[s][v];
                    const float4* A = reinterpret_cast<const float4*>(s_bone + S.b[s][v] * BONE_F);
                    const float4 r0 = A[0], r1 = A[1], r2 = A[2];
                    const float wx = w * gx, wy = w * gy, wz = w * gz;
                    dx = fmaf(r0.x, wx, fmaf(r1.x, wy, fmaf(r2.x, wz, dx)));
                    dy = fmaf(r0.y, wx, fmaf(r1.y, wy, fmaf(r2.y, wz, dy)));
                    dz = fmaf(r0.z, wx, fmaf(r1.z, wy, fmaf(r2.z, wz, dz)));
                }
                s_dv[v * 3] = dx; s_dv[v * 3 + 1] = dy; s_dv[v * 3 + 2] = dz;
            }
        }
        if (tid < VP_PITCH - NVC) s_dv[NVC + tid] = 0.f;
        // (b) per-bone reduction over the bone's vertex list (CSC), one warp per bone
        for (int k = warp; k < NJ; k += LBS_THREADS / 32) {
            float acc[BONE_F];
#pragma unroll
            for (int e = 0; e < BONE_F; ++e) acc[e] = 0.f;
            for (int i = C.ptr[k] + lane; i < C.ptr[k + 1]; i += 32) {
                const int v = C.v[i];
                const float w = C.w[i];
                const float wx = w * s_g[v * 3], wy = w * s_g[v * 3 + 1], wz = w * s_g[v * 3 + 2];
                const float x = s_vp[v * 3], y = s_vp[v * 3 + 1], z = s_vp[v * 3 + 2];
                acc[0] = fmaf(wx, x, acc[0]); acc[1] = fmaf(wx, y, acc[1]); acc[2] = fmaf(wx, z, acc[2]);   acc[3] += wx;
                acc[4] = fmaf(wy, x, acc[4]); acc[5] = fmaf(wy, y, acc[5]); acc[6] = fmaf(wy, z, acc[6]);   acc[7] += wy;
                acc[8] = fmaf(wz, x, acc[8]); acc[9] = fmaf(wz, y, acc[9]); acc[10] = fmaf(wz, z, acc[10]); acc[11] += wz;
            }
#pragma unroll
            for (int e = 0; e < BONE_F; ++e) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
            }
            if (lane == 0) {
                float4* d = reinterpret_cast<float4*>(dbone + hand * (NJ * BONE_F) + k * BONE_F);
                d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                d[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
            }
        }
        __syncthreads();
        for (int q = tid; q < ROW4; q += LBS_THREADS)
            st_stream(reinterpret_cast<float4*>(dv_posed + hand * pitch) + q, reinterpret_cast<const float4*>(s_dv)[q]);
    }
}

size_t lbs_fwd_smem() { return ((sizeof(SkinShared) + 15) & ~size_t(15)) + sizeof(float) * (2 * VP_PITCH + 2 * NJ * BONE_F); }
size_t lbs_bwd_smem() {
    size_t off = (sizeof(SkinShared) + 15) & ~size_t(15);
    off = (off + sizeof(CscShared) + 15) & ~size_t(15);
    return off + sizeof(float) * (3 * VP_PITCH + NJ * BONE_F);
}

}  // namespace

int launch_lbs_forward(const void* blob, const float* v_posed, int pitch, const float* bone, int B,
                       float* verts, float* joints, cudaStream_t s) {
    if (B <= 0) return 0;
    static bool attr_done = false;
    const size_t smem = lbs_fwd_smem();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lbs_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    const int npairs = (B + 1) / 2;
    const int cap = NUM_SMS * 4;
    const int grid = npairs < cap ? npairs : cap;
    lbs_forward_kernel<<<grid, LBS_THREADS, smem, s>>>(blob, v_posed, pitch, bone, B, verts, joints);
    return cuda_rc();
}

int launch_lbs_backward(const void* blob, const float* v_posed, int pitch, const float* bone,
                        const float* g_verts, const float* g_joints, int B,
                        float* dv_posed, float* dbone, cudaStream_t s) {
    if (B <= 0) return 0;
    static bool attr_done = false;
    const size_t smem = lbs_bwd_smem();
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lbs_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    const int cap = NUM_SMS * 2;
    const int grid = B < cap ? B : cap;
    lbs_backward_kernel<<<grid, LBS_THREADS, smem, s>>>(blob, v_posed, pitch, bone, g_verts, g_joints, B, dv_posed, dbone);
    return cuda_rc();
}

}  // namespace mb
