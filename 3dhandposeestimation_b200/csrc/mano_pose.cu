// mano_pose.cu — per-hand "pose stage" of the MANO layer, one warp per hand.
//
//   forward :  PCA coefficients -> axis-angle (MANOLayer.py:126-128), 17 Rodrigues
//              (:82-112; lanes 0..15 = chain joints, lane 16 = global rotation),
//              pose feature (:114-120), folded joint regression (:139-141),
//              kinematic chain by tree level with parent state fetched by __shfl_sync
//              (:159-165), rest-pose removal with the global rotation folded in
//              (:169-175, :188, :204-205)  ->  feat[B][148], bone_t[B/32][16][32][12], joints (chain slots)
//   backward:  SURVEY Appendix A.2 steps 3-7 (reverse chain, Rodrigues backward, PCA^T).
//
// JOINTS_ONLY variants additionally evaluate the five fingertip vertices
// (333,444,672,555,745) from a 15-column slice of the blend basis, so the 21 joints and
// their gradients are produced without touching the 778-vertex contraction at all —
// the workload of every MANO head of the reference (resnet50MANO.py:76,87).
#include <cuda_fp16.h>
#include "hand_math.cuh"
#include "blend_tc.cuh"

namespace mb {

namespace {

constexpr int WARPS = 4;
constexpr int NTIP = 5;
__constant__ int c_chain_slot[NJ] = {0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 13, 14, 15, 17, 18, 19};
__constant__ int c_tip_vert[NTIP] = {333, 444, 672, 555, 745};
__constant__ int c_tip_slot[NTIP] = {4, 8, 12, 16, 20};

constexpr unsigned FULL = 0xffffffffu;

struct PoseShared {
    float pca[NAA * NAA];
    float mean[NAA + 3];
    float j0[NJ * 3];
    float jb[NJ * 3 * NB];
    int parents[NJ];
    int depth[NJ];
    int n_children[NJ];
    int children[NJ][NJ];
    int max_children_at_depth[NJ];
    int max_depth;
    // joints-only extras
    float tip_basis[FEAT_K][16];         // basis[k][tip*3+c] for the 15 tip coordinates
    float tip_w[NTIP][MAX_INFL];
    int tip_b[NTIP][MAX_INFL];
    int tip_cnt[NTIP];
    // per-warp scratch
    alignas(16) float theta[WARPS][48];
    alignas(16) float feat[WARPS][FEAT_K];
    alignas(16) float bone[WARPS][NJ * BONE_F];
    alignas(16) float dbone[WARPS][NJ * BONE_F];
    alignas(16) float dfeat[WARPS][FEAT_K];
    alignas(16) float dtheta[WARPS][48];
    float dJ[WARPS][48];
    float tipv[WARPS][16];
};

__device__ __forceinline__ M3 shfl_m3(const M3& a, int src) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.m[i] = __shfl_sync(FULL, a.m[i], src);
    return r;
}
__device__ __forceinline__ V3 shfl_v3(const V3& a, int src) {
    return v3(__shfl_sync(FULL, a.x, src), __shfl_sync(FULL, a.y, src), __shfl_sync(FULL, a.z, src));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

template <bool JOINTS_ONLY>
__device__ void stage_constants(PoseShared& S, const void* blob, int nc) {
    const BlobLayout L = blob_layout();
    const BlobHeader* H = blob_ptr<BlobHeader>(blob, L.header);
    const float* pca = blob_ptr<float>(blob, L.pca);
    const float* mean = blob_ptr<float>(blob, L.pose_mean);
    const float* j0 = blob_ptr<float>(blob, L.j0);
    const float* jb = blob_ptr<float>(blob, L.jb);
    const int t = threadIdx.x, nt = blockDim.x;
    for (int i = t; i < nc * NAA; i += nt) S.pca[i] = pca[i];
    for (int i = t; i < NAA; i += nt) S.mean[i] = mean[i];
    for (int i = t; i < NJ * 3; i += nt) S.j0[i] = j0[i];
    for (int i = t; i < NJ * 3 * NB; i += nt) S.jb[i] = jb[i];
    for (int i = t; i < NJ; i += nt) {
        S.parents[i] = H->parents[i];
        S.depth[i] = H->depth[i];
        S.n_children[i] = H->n_children[i];
        S.max_children_at_depth[i] = H->max_children_at_depth[i];
    }
    for (int i = t; i < NJ * NJ; i += nt) S.children[i / NJ][i % NJ] = H->children[i / NJ][i % NJ];
    if (t == 0) S.max_depth = H->max_depth;
    if (JOINTS_ONLY) {
        const float* basis = blob_ptr<float>(blob, L.basis);
        const float* sw = blob_ptr<float>(blob, L.skin_w);
        const uint8_t* sb = blob_ptr<uint8_t>(blob, L.skin_b);
        const uint8_t* sc = blob_ptr<uint8_t>(blob, L.skin_cnt);
        for (int i = t; i < FEAT_K * 16; i += nt) {
            int k = i / 16, tc = i % 16;
            S.tip_basis[k][tc] = (tc < 15) ? basis[(size_t)k * VP_PITCH + c_tip_vert[tc / 3] * 3 + tc % 3] : 0.f;
        }
        for (int i = t; i < NTIP * MAX_INFL; i += nt) {
            int tp = i / MAX_INFL, s = i % MAX_INFL;
            S.tip_w[tp][s] = sw[c_tip_vert[tp] * MAX_INFL + s];
            S.tip_b[tp][s] = sb[c_tip_vert[tp] * MAX_INFL + s];
        }
        for (int i = t; i < NTIP; i += nt) S.tip_cnt[i] = sc[c_tip_vert[i]];
    }
    __syncthreads();
}

// Per-lane state of the forward chain.  lane 0..15 = joint, lane 16 = global rotation.
struct Lane {
    V3 r;       // axis-angle
    M3 R;       // local rotation
    V3 J;       // rest joint
    M3 Rg;      // global rotation of the joint
    V3 tg;      // global joint position (before the global rotation Rq)
    M3 Rgp;     // parent's global rotation
    V3 Jp;      // parent's rest joint
    M3 Rq;      // global rotation R(rot), on every lane
    float beta; // lane < 10
};

// Forward chain for one hand; all 32 lanes call it.  Fills s_feat (the blend feature row).
__device__ __forceinline__ void chain_forward(const PoseShared& S, float* s_theta, float* s_feat, int nc,
                                              const float* __restrict__ rot, const float* __restrict__ coeffs,
                                              const float* __restrict__ betas, long long hand, int lane, Lane& st) {
    // ---- PCA coefficients -> 45-D axis-angle (MANOLayer.py:126)
    float c0 = (lane < nc) ? coeffs[hand * nc + lane] : 0.f;
    float c1 = (lane + 32 < nc) ? coeffs[hand * nc + lane + 32] : 0.f;
    float th0 = 0.f, th1 = 0.f;
    const bool hi = lane + 32 < NAA;
    for (int i = 0; i < nc; ++i) {
        float ci = __shfl_sync(FULL, (i < 32) ? c0 : c1, i & 31);
        th0 = fmaf(ci, S.pca[i * NAA + lane], th0);
        if (hi) th1 = fmaf(ci, S.pca[i * NAA + lane + 32], th1);
    }
    s_theta[lane] = th0 + S.mean[lane];
    if (hi) s_theta[lane + 32] = th1 + S.mean[lane + 32];
    __syncwarp();

    // ---- per-lane axis-angle: lane 0 = constant root [pi,0,0] (:76,:128), 1..15 = theta, 16 = rot
    const int j = lane < NJ ? lane : 0;
    if (lane == 0) st.r = v3(3.14159274101257324f, 0.f, 0.f);
    else if (lane < NJ) st.r = v3(s_theta[3 * (lane - 1)], s_theta[3 * (lane - 1) + 1], s_theta[3 * (lane - 1) + 2]);
    else if (lane == NJ) st.r = v3(rot[hand * 3], rot[hand * 3 + 1], rot[hand * 3 + 2]);
    else st.r = v3(0.f, 0.f, 0.f);
    st.R = rodrigues(st.r);
    st.Rq = shfl_m3(st.R, NJ);

    // ---- folded joint regression  J = J0 + Jb beta  (:139-141)
    st.beta = (lane < NB) ? betas[hand * NB + lane] : 0.f;
    float jx = S.j0[j * 3], jy = S.j0[j * 3 + 1], jz = S.j0[j * 3 + 2];
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        float b = __shfl_sync(FULL, st.beta, s);
        jx = fmaf(S.jb[(j * 3 + 0) * NB + s], b, jx);
        jy = fmaf(S.jb[(j * 3 + 1) * NB + s], b, jy);
        jz = fmaf(S.jb[(j * 3 + 2) * NB + s], b, jz);
    }
    st.J = v3(jx, jy, jz);

    // ---- blend feature row  f = [beta | vec(R_j - I), j = 1..15 | 1 | 0 0]  (:116-119)
    if (lane < NB) s_feat[lane] = st.beta;
    if (lane >= 1 && lane < NJ) {
        float* f = s_feat + NB + 9 * (lane - 1);
#pragma unroll
        for (int e = 0; e < 9; ++e) f[e] = st.R.m[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
    }
    if (lane == 0) { s_feat[FEAT_ONE] = 1.f; s_feat[FEAT_ONE + 1] = 0.f; s_feat[FEAT_ONE + 2] = 0.f; }

    // ---- kinematic chain by tree level (:159-165)
    const int p = S.parents[j] < 0 ? 0 : S.parents[j];
    const int dep = lane < NJ ? S.depth[j] : -1;
    st.Rg = st.R;
    st.tg = st.J;
    st.Jp = shfl_v3(st.J, p);
    st.Rgp = m3_identity();
    for (int level = 1; level <= S.max_depth; ++level) {
        M3 Rp = shfl_m3(st.Rg, p);
        V3 tp = shfl_v3(st.tg, p);
        if (dep == level) {
            st.Rgp = Rp;
            st.Rg = m3_mul(Rp, st.R);
            st.tg = v3_add(tp, m3_vec(Rp, v3_sub(st.J, st.Jp)));
        }
    }
    __syncwarp();
}

// bone transform with the global rotation folded in: A' = [Rq Rg | Rq (tg - Rg J)]
__device__ __forceinline__ void bone_transform(const Lane& st, float* out12) {
    M3 Rp = m3_mul(st.Rq, st.Rg);
    V3 tA = v3_sub(st.tg, m3_vec(st.Rg, st.J));
    V3 tp = m3_vec(st.Rq, tA);
    out12[0] = Rp.m[0]; out12[1] = Rp.m[1]; out12[2] = Rp.m[2];  out12[3] = tp.x;
    out12[4] = Rp.m[3]; out12[5] = Rp.m[4]; out12[6] = Rp.m[5];  out12[7] = tp.y;
    out12[8] = Rp.m[6]; out12[9] = Rp.m[7]; out12[10] = Rp.m[8]; out12[11] = tp.z;
}

// Rest-pose tip vertices from the 15-column basis slice; lanes 0..14 -> s_tipv[tc].
__device__ __forceinline__ void tips_rest_pose(const PoseShared& S, const float* s_feat, float* s_tipv, int lane) {
    if (lane < 15) {
        float acc = 0.f;
        for (int k = 0; k <= FEAT_ONE; ++k) acc = fmaf(s_feat[k], S.tip_basis[k][lane], acc);
        s_tipv[lane] = acc;
    }
    __syncwarp();
}

template <bool JOINTS_ONLY>
__global__ void __launch_bounds__(WARPS * 32)
pose_forward_kernel(const void* __restrict__ blob, int nc, const float* __restrict__ rot,
                    const float* __restrict__ coeffs, const float* __restrict__ betas, int B,
                    float* __restrict__ feat, unsigned char* __restrict__ featp, float* __restrict__ bone_t,
                    float* __restrict__ joints) {
    __shared__ alignas(16) PoseShared S;
    stage_constants<JOINTS_ONLY>(S, blob, nc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * WARPS;
    for (long long hand = (long long)blockIdx.x * WARPS + warp; hand < B; hand += nwarps) {
        Lane st;
        chain_forward(S, S.theta[warp], S.feat[warp], nc, rot, coeffs, betas, hand, lane, st);
        if (lane < NJ) {
            bone_transform(st, &S.bone[warp][lane * BONE_F]);
            V3 jt = m3_vec(st.Rq, st.tg);
            float* o = joints + hand * (NOUTJ * 3) + c_chain_slot[lane] * 3;
            o[0] = jt.x; o[1] = jt.y; o[2] = jt.z;
        }
        __syncwarp();
        if (!JOINTS_ONLY) {
            if (feat != nullptr) {
                float4* fo = reinterpret_cast<float4*>(feat + hand * FEAT_K);
                const float4* fs = reinterpret_cast<const float4*>(S.feat[warp]);
                for (int i = lane; i < FEAT_K / 4; i += 32) fo[i] = fs[i];
            }
            if (featp != nullptr) {
                // A operand of the tcgen05 contraction: x * 2^4 split into fp16 hi + lo, written as
                // 16-byte K-groups into the UMMA canonical tile layout (blend_tc.cuh)
                const float fs = (float)(1 << TC_FEAT_SCALE_LOG2);
                for (int p = lane; p < 2 * (TC_K / 8); p += 32) {
                    const int sp = p / (TC_K / 8), kg8 = p - sp * (TC_K / 8);
                    __align__(16) __half h[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int k = kg8 * 8 + e;
                        const float x = k < TC_K_REAL ? S.feat[warp][k] * fs : 0.f;
                        const __half hi = __float2half_rn(x);
                        h[e] = sp == 0 ? hi : __float2half_rn(x - __half2float(hi));
                    }
                    *reinterpret_cast<uint4*>(featp + tc_feat_group_offset(hand, kg8, sp)) = *reinterpret_cast<const uint4*>(h);
                }
            }
            // bone transforms leave grouped by 32 hands: bone_t[hand / 32][bone][hand % 32][12] — one bone
            // of a hand group is 1.5 KB contiguous, one hand's transform 48 B (what the lane = hand
            // skinning kernels fetch with three 16-byte async copies per lane)
            float* bo = bone_t + (hand >> 5) * (NJ * BONE_F * 32) + (hand & 31) * BONE_F;
            for (int i = lane; i < NJ * BONE_F; i += 32) bo[(i / BONE_F) * (BONE_F * 32) + (i % BONE_F)] = S.bone[warp][i];
        } else {
            tips_rest_pose(S, S.feat[warp], S.tipv[warp], lane);
            if (lane < NTIP) {
                const float x = S.tipv[warp][lane * 3], y = S.tipv[warp][lane * 3 + 1], z = S.tipv[warp][lane * 3 + 2];
                float ox = 0.f, oy = 0.f, oz = 0.f;
                for (int s = 0; s < S.tip_cnt[lane]; ++s) {
                    const float w = S.tip_w[lane][s];
                    const float* A = &S.bone[warp][S.tip_b[lane][s] * BONE_F];
                    ox = fmaf(w, fmaf(A[0], x, fmaf(A[1], y, fmaf(A[2], z, A[3]))), ox);
                    oy = fmaf(w, fmaf(A[4], x, fmaf(A[5], y, fmaf(A[6], z, A[7]))), oy);
                    oz = fmaf(w, fmaf(A[8], x, fmaf(A[9], y, fmaf(A[10], z, A[11]))), oz);
                }
                float* o = joints + hand * (NOUTJ * 3) + c_tip_slot[lane] * 3;
                o[0] = ox; o[1] = oy; o[2] = oz;
            }
        }
        __syncwarp();
    }
}

template <bool JOINTS_ONLY>
__global__ void __launch_bounds__(WARPS * 32)
pose_backward_kernel(const void* __restrict__ blob, int nc, const float* __restrict__ rot,
                     const float* __restrict__ coeffs, const float* __restrict__ betas,
                     const float* __restrict__ dfeat_g, int dfeat_parts, size_t dfeat_stride,
                     const float* __restrict__ dbone_g, const float* __restrict__ g_joints, int B,
                     float* __restrict__ g_rot, float* __restrict__ g_coeffs, float* __restrict__ g_betas) {
    __shared__ alignas(16) PoseShared S;
    stage_constants<JOINTS_ONLY>(S, blob, nc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * WARPS;
    for (long long hand = (long long)blockIdx.x * WARPS + warp; hand < B; hand += nwarps) {
        Lane st;
        chain_forward(S, S.theta[warp], S.feat[warp], nc, rot, coeffs, betas, hand, lane, st);
        const float* gj_hand = g_joints + hand * (NOUTJ * 3);

        // ---- upstream gradients of the bone transforms and the blend features
        float* s_dbone = S.dbone[warp];
        float* s_dfeat = S.dfeat[warp];
        if (!JOINTS_ONLY) {
            const float4* src = reinterpret_cast<const float4*>(dbone_g + hand * (NJ * BONE_F));
            float4* dst = reinterpret_cast<float4*>(s_dbone);
            for (int i = lane; i < NJ * BONE_F / 4; i += 32) dst[i] = src[i];
            const float4* fsrc = reinterpret_cast<const float4*>(dfeat_g + hand * FEAT_K);
            float4* fdst = reinterpret_cast<float4*>(s_dfeat);
            for (int i = lane; i < FEAT_K / 4; i += 32) {
                float4 a = fsrc[i];
                for (int part = 1; part < dfeat_parts; ++part) {     // K ranges of a split backward contraction
                    const float4 b = reinterpret_cast<const float4*>(dfeat_g + part * dfeat_stride + hand * FEAT_K)[i];
                    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                }
                fdst[i] = a;
            }
            __syncwarp();
        } else {
            // skinning backward restricted to the five tip vertices (A.2 steps 1-2)
            if (lane < NJ) bone_transform(st, &S.bone[warp][lane * BONE_F]);
            for (int i = lane; i < NJ * BONE_F; i += 32) s_dbone[i] = 0.f;
            __syncwarp();
            tips_rest_pose(S, S.feat[warp], S.tipv[warp], lane);
            float dvx = 0.f, dvy = 0.f, dvz = 0.f;
            if (lane < NTIP) {
                const float x = S.tipv[warp][lane * 3], y = S.tipv[warp][lane * 3 + 1], z = S.tipv[warp][lane * 3 + 2];
                const float* g = gj_hand + c_tip_slot[lane] * 3;
                const float gx = g[0], gy = g[1], gz = g[2];
                for (int s = 0; s < S.tip_cnt[lane]; ++s) {
                    const float w = S.tip_w[lane][s];
                    const int b = S.tip_b[lane][s];
                    const float* A = &S.bone[warp][b * BONE_F];
                    float* D = &s_dbone[b * BONE_F];
                    const float wx = w * gx, wy = w * gy, wz = w * gz;
                    atomicAdd(&D[0], wx * x); atomicAdd(&D[1], wx * y); atomicAdd(&D[2], wx * z);  atomicAdd(&D[3], wx);
                    atomicAdd(&D[4], wy * x); atomicAdd(&D[5], wy * y); atomicAdd(&D[6], wy * z);  atomicAdd(&D[7], wy);
                    atomicAdd(&D[8], wz * x); atomicAdd(&D[9], wz * y); atomicAdd(&D[10], wz * z); atomicAdd(&D[11], wz);
                    dvx = fmaf(A[0], wx, fmaf(A[4], wy, fmaf(A[8], wz, dvx)));
                    dvy = fmaf(A[1], wx, fmaf(A[5], wy, fmaf(A[9], wz, dvy)));
                    dvz = fmaf(A[2], wx, fmaf(A[6], wy, fmaf(A[10], wz, dvz)));
                }
            }
            __syncwarp();
            if (lane < NTIP) { S.tipv[warp][lane * 3] = dvx; S.tipv[warp][lane * 3 + 1] = dvy; S.tipv[warp][lane * 3 + 2] = dvz; }
            __syncwarp();
            for (int k = lane; k < FEAT_K; k += 32) {
                float acc = 0.f;
#pragma unroll
                for (int tc = 0; tc < 15; ++tc) acc = fmaf(S.tip_basis[k][tc], S.tipv[warp][tc], acc);
                s_dfeat[k] = acc;
            }
            __syncwarp();
        }

        // ---- A.2 step 3: split A'_k = Rq [Rg_k | tg_k - Rg_k J_k], joint_k = Rq tg_k
        float dAp[BONE_F];
        V3 gj = v3(0.f, 0.f, 0.f);
        if (lane < NJ) {
#pragma unroll
            for (int e = 0; e < BONE_F; ++e) dAp[e] = s_dbone[lane * BONE_F + e];
            const float* g = gj_hand + c_chain_slot[lane] * 3;
            gj = v3(g[0], g[1], g[2]);
        } else {
#pragma unroll
            for (int e = 0; e < BONE_F; ++e) dAp[e] = 0.f;
        }
        const V3 tA = v3_sub(st.tg, m3_vec(st.Rg, st.J));
        M3 dRq;
        {
            const float tav[3] = {tA.x, tA.y, tA.z};
            const float tgv[3] = {st.tg.x, st.tg.y, st.tg.z};
            const float gjv[3] = {gj.x, gj.y, gj.z};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    float v = dAp[i * 4 + 0] * st.Rg.m[jj * 3 + 0] + dAp[i * 4 + 1] * st.Rg.m[jj * 3 + 1]
                            + dAp[i * 4 + 2] * st.Rg.m[jj * 3 + 2] + dAp[i * 4 + 3] * tav[jj] + gjv[i] * tgv[jj];
                    dRq.m[i * 3 + jj] = warp_sum(lane < NJ ? v : 0.f);
                }
        }
        M3 dAR;   // Rq^T dA'[:, :3]
        V3 dAt;   // Rq^T dA'[:, 3]
        {
            M3 dApR;
            dApR.m[0] = dAp[0]; dApR.m[1] = dAp[1]; dApR.m[2] = dAp[2];
            dApR.m[3] = dAp[4]; dApR.m[4] = dAp[5]; dApR.m[5] = dAp[6];
            dApR.m[6] = dAp[8]; dApR.m[7] = dAp[9]; dApR.m[8] = dAp[10];
            dAR = m3_tmul(st.Rq, dApR);
            dAt = m3_tvec(st.Rq, v3(dAp[3], dAp[7], dAp[11]));
        }
        M3 dRg = dAR;
        m3_add_outer(dRg, v3(-dAt.x, -dAt.y, -dAt.z), st.J);
        V3 dtg = v3_add(dAt, m3_tvec(st.Rq, gj));
        V3 dJ = m3_tvec(st.Rg, v3(-dAt.x, -dAt.y, -dAt.z));

        // ---- A.2 step 4: reverse chain, deepest level first
        const int j = lane < NJ ? lane : 0;
        const int dep = lane < NJ ? S.depth[j] : -1;
        const int nch = lane < NJ ? S.n_children[j] : 0;
        M3 dRl = m3_zero();
        for (int level = S.max_depth; level >= 1; --level) {
            const bool child = (dep == level);
            M3 cR = m3_zero();
            V3 ct = v3(0.f, 0.f, 0.f), cJ = v3(0.f, 0.f, 0.f);
            if (child) {
                dRl = m3_tmul(st.Rgp, dRg);
                cR = m3_mult(dRg, st.R);
                m3_add_outer(cR, dtg, v3_sub(st.J, st.Jp));
                ct = dtg;
                V3 dd = m3_tvec(st.Rgp, dtg);
                dJ = v3_add(dJ, dd);
                cJ = v3(-dd.x, -dd.y, -dd.z);
            }
            const bool par = (dep == level - 1);
            const int rounds = S.max_children_at_depth[level - 1];
            for (int c = 0; c < rounds; ++c) {
                const bool take = par && c < nch;
                const int src = take ? S.children[j][c] : 0;
                M3 gR = shfl_m3(cR, src);
                V3 gt = shfl_v3(ct, src);
                V3 gJ = shfl_v3(cJ, src);
                if (take) { m3_acc(dRg, gR); dtg = v3_add(dtg, gt); dJ = v3_add(dJ, gJ); }
            }
        }
        if (lane == 0) dJ = v3_add(dJ, dtg);      // tg_0 = J_0 ; R_0 is a constant

        // ---- A.2 steps 5-6: pose-feature gradient joins dR_j ; Rodrigues backward
        if (lane >= 1 && lane < NJ) {
#pragma unroll
            for (int e = 0; e < 9; ++e) dRl.m[e] += s_dfeat[NB + 9 * (lane - 1) + e];
            V3 dth = rodrigues_bwd(st.r, dRl);
            S.dtheta[warp][3 * (lane - 1)] = dth.x;
            S.dtheta[warp][3 * (lane - 1) + 1] = dth.y;
            S.dtheta[warp][3 * (lane - 1) + 2] = dth.z;
        }
        if (lane == NJ) {
            V3 dr = rodrigues_bwd(st.r, dRq);
            g_rot[hand * 3] = dr.x; g_rot[hand * 3 + 1] = dr.y; g_rot[hand * 3 + 2] = dr.z;
        }
        if (lane < NJ) { S.dJ[warp][lane * 3] = dJ.x; S.dJ[warp][lane * 3 + 1] = dJ.y; S.dJ[warp][lane * 3 + 2] = dJ.z; }
        __syncwarp();

        // ---- A.2 step 7: d_coeffs = C[:nc] dtheta ; d_beta = S^T dv_posed + Jb^T dJ
        for (int i = lane; i < nc; i += 32) {
            float acc = 0.f;
            for (int e = 0; e < NAA; ++e) acc = fmaf(S.pca[i * NAA + e], S.dtheta[warp][e], acc);
            g_coeffs[hand * nc + i] = acc;
        }
        if (lane < NB) {
            float acc = s_dfeat[lane];
            for (int e = 0; e < NJ * 3; ++e) acc = fmaf(S.jb[e * NB + lane], S.dJ[warp][e], acc);
            g_betas[hand * NB + lane] = acc;
        }
        __syncwarp();
    }
}

inline int pose_grid(int B) {
    long long blocks = ((long long)B + WARPS - 1) / WARPS;
    long long cap = (long long)NUM_SMS * 16;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace

int launch_pose_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                        int B, float* feat, unsigned char* featp, float* bone, float* joints, cudaStream_t s) {
    pose_forward_kernel<false><<<pose_grid(B), WARPS * 32, 0, s>>>(blob, nc, rot, coeffs, betas, B, feat, featp, bone, joints);
    return cuda_rc();
}

int launch_joints_only_forward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                               int B, float* joints, cudaStream_t s) {
    pose_forward_kernel<true><<<pose_grid(B), WARPS * 32, 0, s>>>(blob, nc, rot, coeffs, betas, B, nullptr, nullptr, nullptr, joints);
    return cuda_rc();
}

int launch_pose_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                         const float* dfeat, int dfeat_parts, size_t dfeat_stride, const float* dbone, const float* g_joints,
                         int B, float* g_rot, float* g_coeffs, float* g_betas, cudaStream_t s) {
    pose_backward_kernel<false><<<pose_grid(B), WARPS * 32, 0, s>>>(blob, nc, rot, coeffs, betas, dfeat, dfeat_parts, dfeat_stride, dbone, g_joints, B,
                                                                     g_rot, g_coeffs, g_betas);
    return cuda_rc();
}

int launch_joints_only_backward(const void* blob, int nc, const float* rot, const float* coeffs, const float* betas,
                                const float* g_joints, int B, float* g_rot, float* g_coeffs, float* g_betas,
                                cudaStream_t s) {
    pose_backward_kernel<true><<<pose_grid(B), WARPS * 32, 0, s>>>(blob, nc, rot, coeffs, betas, nullptr, 0, 0, nullptr, g_joints, B,
                                                                    g_rot, g_coeffs, g_betas);
    return cuda_rc();
}

}  // namespace mb
